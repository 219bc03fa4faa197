// gemm_simt.cu -- the exact-fp32 arm of dmc_gemm: same contract, FFMA pipes, no tensor cores.
// It exists to cross-check the tcgen05 kernel on the device and as the strict-fp32 fallback for
// shapes the tensor-core kernel's TMA descriptors cannot express (unaligned strides).  It is a
// plain 64x64x16 shared-memory tiled kernel; it is not a performance path.
#include "dmc_common.cuh"

namespace dmc {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct SimtArgs {
  int M, N, K;
  const float* A; long long sam, sak;   // A(m,k) = A[m*sam + k*sak]
  const float* B; long long sbn, sbk;   // B(n,k) = B[n*sbn + k*sbk]
  void* D; long long ldd; int out_dtype;
  const float* col_scale; const float* bias; const float* alpha_dev; float alpha;
  int act; void* aux; long long ldaux; int aux_dtype;
};

__device__ __forceinline__ float ld_any(const void* p, long long i, int dt) {
  return dt == DMC_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_any(void* p, long long i, int dt, float v) {
  if (dt == DMC_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[i] = v;
}

__global__ void __launch_bounds__(256)
gemm_simt_kernel(const SimtArgs a) {
  pdl_prologue();
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;     // thread computes rows ty*4.., cols tx*4..
  float acc[4][4] = {};
  const bool a_kc = (a.sak == 1), b_kc = (a.sbk == 1);
  for (int k0 = 0; k0 < a.K; k0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = threadIdx.x + i * 256;
      int m, k;
      if (a_kc) { k = e % TK; m = e / TK; } else { m = e % TM; k = e / TM; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < a.M && gk < a.K) ? a.A[gm * a.sam + gk * a.sak] : 0.f;
      int n, kb;
      if (b_kc) { kb = e % TK; n = e / TK; } else { n = e % TN; kb = e / TN; }
      const int gn = n0 + n, gkb = k0 + kb;
      Bs[kb][n] = (gn < a.N && gkb < a.K) ? a.B[gn * a.sbn + gkb * a.sbk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = As[k][ty * 4 + i]; bv[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float alpha = a.alpha;
  if (a.alpha_dev) alpha *= __ldg(a.alpha_dev);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long row = m0 + ty * 4 + i;
    if (row >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= a.N) continue;
      float v = acc[i][j];
      if (a.col_scale) v *= a.col_scale[col];
      v *= alpha;
      if (a.bias) v += a.bias[col];
      if (a.act == DMC_ACT_GELU) {
        if (a.aux) st_any(a.aux, row * a.ldaux + col, a.aux_dtype, v);
        v = gelu_f(v);
      } else if (a.act == DMC_ACT_GELU_BWD) {
        v *= gelu_grad_f(ld_any(a.aux, row * a.ldaux + col, a.aux_dtype));
      } else if (a.act == DMC_ACT_GELU_DG) {
        st_any(a.aux, row * a.ldaux + col, a.aux_dtype, gelu_grad_f(v));
        v = gelu_f(v);
      } else if (a.act == DMC_ACT_MUL_AUX) {
        v *= ld_any(a.aux, row * a.ldaux + col, a.aux_dtype);
      }
      st_any(a.D, row * a.ldd + col, a.out_dtype, v);
    }
  }
}

}  // namespace
}  // namespace dmc

using namespace dmc;

extern "C" int dmc_gemm_simt(const dmc_gemm_args* g, void* stream) {
  DMC_REQUIRE(g != nullptr, "dmc_gemm_simt: null args");
  DMC_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, "dmc_gemm_simt: empty problem");
  DMC_REQUIRE(g->M < (1ll << 31) && g->N < (1ll << 31) && g->K < (1ll << 31), "dmc_gemm_simt: dimension too large");
  DMC_REQUIRE(g->in_dtype == DMC_F32, "dmc_gemm_simt: operands must be fp32");
  DMC_REQUIRE(g->A && g->B && g->D, "dmc_gemm_simt: null operand");
  DMC_REQUIRE(g->act != DMC_ACT_GELU_BWD || g->aux != nullptr, "dmc_gemm_simt: DMC_ACT_GELU_BWD needs aux");
  SimtArgs a{};
  a.M = (int)g->M; a.N = (int)g->N; a.K = (int)g->K;
  a.A = static_cast<const float*>(g->A); a.B = static_cast<const float*>(g->B);
  if (g->a_mn_major) { a.sam = 1; a.sak = g->lda; } else { a.sam = g->lda; a.sak = 1; }
  if (g->b_mn_major) { a.sbn = 1; a.sbk = g->ldb; } else { a.sbn = g->ldb; a.sbk = 1; }
  a.D = g->D; a.ldd = g->ldd; a.out_dtype = g->out_dtype;
  a.col_scale = g->col_scale; a.bias = g->bias; a.alpha_dev = g->alpha_dev; a.alpha = g->alpha;
  a.act = g->act; a.aux = g->aux; a.ldaux = g->ldaux; a.aux_dtype = g->aux_dtype;
  dim3 grid((unsigned)ceil_div(g->N, TN), (unsigned)ceil_div(g->M, TM));
  DMC_REQUIRE(grid.y <= 65535, "dmc_gemm_simt: M too large for this kernel (%lld)", (long long)g->M);
  launch_kernel(gemm_simt_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), a);
  DMC_LAUNCH_CHECK("gemm_simt_kernel launch");
  return 0;
}
