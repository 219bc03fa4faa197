// rowops.cu -- the small row-wise kernels of DINOHead around the last-layer GEMM:
// F.normalize (utils/vision_transformer.py:292) forward/backward, the weight_norm re-parameterisation
// (utils/vision_transformer.py:279) forward/backward, dtype casts / TF32 splits and a column sum
// (bias gradients).  All are HBM-bound streaming kernels: one warp per row, coalesced accesses,
// warp-shuffle reductions, no shared memory.
#include "dmc_common.cuh"

namespace dmc {
namespace {

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ float ld_in(const void* p, long long i, int dt) {
  return dt == DMC_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}

constexpr int kWarpsPerBlock = 8;
constexpr int kRowsPerWarp = 3;       // weightnorm_fwd, width-256 path: 3 rows x 2 x 128 bits in flight per lane within 40 registers

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
normalize_fwd_kernel(const void* __restrict__ z, int z_dtype, long long n_rows, int dim, long long ld, float eps,
                     float* __restrict__ zhat_f32, __nv_bfloat16* __restrict__ zhat_bf16, float* __restrict__ zhat_lo,
                     float* __restrict__ inv_den) {
  pdl_prologue();
  const long long row = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  float ss = 0.f;
  for (int c = lane; c < dim; c += 32) { float v = ld_in(z, row * ld + c, z_dtype); ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  const float den = fmaxf(sqrtf(ss), eps);
  const float inv = 1.0f / den;
  if (lane == 0) inv_den[row] = inv;
  for (int c = lane; c < dim; c += 32) {
    const float v = ld_in(z, row * ld + c, z_dtype) / den;
    const long long o = row * dim + c;
    if (zhat_lo) {                 // 3xTF32 operand pair: hi is tf32-exact, lo the residual
      const float hi = tf32_round(v);
      if (zhat_f32) zhat_f32[o] = hi;
      zhat_lo[o] = tf32_round(v - hi);
    } else if (zhat_f32) {
      zhat_f32[o] = v;
    }
    if (zhat_bf16) zhat_bf16[o] = __float2bfloat16_rn(v);
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
normalize_bwd_kernel(const float* __restrict__ dzhat, const float* __restrict__ zhat, const float* __restrict__ inv_den,
                     long long n_rows, int dim, float eps, void* __restrict__ dz, int dz_dtype) {
  pdl_prologue();
  const long long row = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const float inv = inv_den[row];
  const bool clamped = (inv * eps >= 1.0f);          // ||z|| < eps: zhat = z/eps, no projection term
  float proj = 0.f;
  if (!clamped)
    for (int c = lane; c < dim; c += 32) proj = fmaf(dzhat[row * dim + c], zhat[row * dim + c], proj);
  proj = warp_sum(proj);
  for (int c = lane; c < dim; c += 32) {
    const long long o = row * dim + c;
    const float v = (dzhat[o] - proj * zhat[o]) * inv;
    if (dz_dtype == DMC_BF16) reinterpret_cast<__nv_bfloat16*>(dz)[o] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(dz)[o] = v;
  }
}

// Vector path (dim % 4 == 0, 16-byte aligned rows): each lane owns float4 #lane, #lane+32, ... of its row; the
// usual bottleneck width (256) keeps the whole row in registers between the two passes.
// <= 40 registers (6 CTAs / SM bound): one 256-thread CTA then fits next to a resident GEMM CTA (168 regs x 320 threads)
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 6)
weightnorm_fwd_kernel(const float* __restrict__ v, const float* __restrict__ g, long long K, int dim,
                      float* __restrict__ w_f32, float* __restrict__ w_lo, __nv_bfloat16* __restrict__ w_bf16,
                      float* __restrict__ scale, float* __restrict__ inv_vnorm, bool vec_ok, float* __restrict__ gmax) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  (void)gmax;
  if (vec_ok && dim == 256) {                          // the DINO bottleneck width: whole rows in registers
    // kRowsPerWarp consecutive rows per warp, all their loads issued before the first reduction: four times the bytes in
    // flight per warp and a quarter of the blocks of the one-row-per-warp form (ncu: 52 % issue-active, 46 % DRAM before)
    const long long nblk = (K + kWarpsPerBlock * kRowsPerWarp - 1) / (kWarpsPerBlock * kRowsPerWarp);
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {     // grid-stride over row groups (see streaming_grid)
    const long long row0 = (blk * kWarpsPerBlock + (threadIdx.x >> 5)) * kRowsPerWarp;
    if (row0 >= K) continue;
    float4 x0[kRowsPerWarp], x1[kRowsPerWarp];
    float ss[kRowsPerWarp], gr[kRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const long long row = min(row0 + r, K - 1);      // rows past the end repeat the last one and are not stored
      const float4* v4 = reinterpret_cast<const float4*>(v + row * 256);
      x0[r] = __ldg(v4 + lane); x1[r] = __ldg(v4 + lane + 32);
      gr[r] = __ldg(g + row);
    }
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      float a = x0[r].x * x0[r].x;
      a = fmaf(x0[r].y, x0[r].y, a); a = fmaf(x0[r].z, x0[r].z, a); a = fmaf(x0[r].w, x0[r].w, a);
      a = fmaf(x1[r].x, x1[r].x, a); a = fmaf(x1[r].y, x1[r].y, a); a = fmaf(x1[r].z, x1[r].z, a); a = fmaf(x1[r].w, x1[r].w, a);
      ss[r] = a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r) ss[r] += __shfl_xor_sync(0xffffffffu, ss[r], o);
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const long long row = row0 + r;
      if (row >= K) break;                             // warp-uniform
      const float nrm = sqrtf(ss[r]);
      const float sc = gr[r] / nrm;                    // torch _weight_norm: v * (g / ||v||)
      if (lane == 0) { scale[row] = sc; inv_vnorm[row] = 1.0f / nrm; }
      const float w[8] = {x0[r].x * sc, x0[r].y * sc, x0[r].z * sc, x0[r].w * sc, x1[r].x * sc, x1[r].y * sc, x1[r].z * sc, x1[r].w * sc};
      const long long o0 = row * 256 + 4 * lane, o1 = o0 + 128;
      if (w_lo) {
        float hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { hi[e] = tf32_round(w[e]); lo[e] = tf32_round(w[e] - hi[e]); }
        if (w_f32) {
          *reinterpret_cast<float4*>(w_f32 + o0) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<float4*>(w_f32 + o1) = make_float4(hi[4], hi[5], hi[6], hi[7]);
        }
        *reinterpret_cast<float4*>(w_lo + o0) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<float4*>(w_lo + o1) = make_float4(lo[4], lo[5], lo[6], lo[7]);
      } else if (w_f32) {
        *reinterpret_cast<float4*>(w_f32 + o0) = make_float4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<float4*>(w_f32 + o1) = make_float4(w[4], w[5], w[6], w[7]);
      }
      if (w_bf16) {
        *reinterpret_cast<uint2*>(w_bf16 + o0) = make_uint2(pack_bf16(w[0], w[1]), pack_bf16(w[2], w[3]));
        *reinterpret_cast<uint2*>(w_bf16 + o1) = make_uint2(pack_bf16(w[4], w[5]), pack_bf16(w[6], w[7]));
      }
    }
    }
    return;
  }
  for (long long row = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5); row < K;
       row += static_cast<long long>(gridDim.x) * kWarpsPerBlock) {
  const float* vr = v + row * dim;
  float ss = 0.f;
  for (int c = lane; c < dim; c += 32) { float x = vr[c]; ss = fmaf(x, x, ss); }
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  const float sc = g[row] / nrm;                      // torch _weight_norm: v * (g / ||v||)
  if (lane == 0) { scale[row] = sc; inv_vnorm[row] = 1.0f / nrm; }
  for (int c = lane; c < dim; c += 32) {
    const float w = vr[c] * sc;
    const long long o = row * dim + c;
    if (w_lo) {
      const float hi = tf32_round(w);
      if (w_f32) w_f32[o] = hi;
      w_lo[o] = tf32_round(w - hi);
    } else if (w_f32) {
      w_f32[o] = w;
    }
    if (w_bf16) w_bf16[o] = __float2bfloat16_rn(w);
  }
  }
}

// TD = storage type of dW: fp32 (the wgrad GEMM's default output) or bf16 (the data-parallel exchange averages dW over
// the ranks in bf16 BEFORE this pass -- the pass is linear in dW, so mean(dv) = weightnorm_bwd(mean(dW))).
template <typename TD>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
weightnorm_bwd_kernel(const TD* __restrict__ dw, const float* __restrict__ v, const float* __restrict__ scale,
                      const float* __restrict__ inv_vnorm, long long K, int dim, float* __restrict__ dv, float* __restrict__ dg,
                      bool vec_ok) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  for (long long row = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5); row < K;
       row += static_cast<long long>(gridDim.x) * kWarpsPerBlock) {          // grid-stride (see streaming_grid)
  const float iv = inv_vnorm[row], sc = scale[row];
  if (vec_ok && dim == 256) {
    using Q4 = Quad<TD>;
    const float4* v4 = reinterpret_cast<const float4*>(v + row * 256);
    float4* dv4 = reinterpret_cast<float4*>(dv + row * 256);
    float e0[4], e1[4];
    {
      const typename Q4::Raw r0 = Q4::load(dw + row * 256 + lane * 4), r1 = Q4::load(dw + row * 256 + (lane + 32) * 4);
      Q4::unpack(r0, e0); Q4::unpack(r1, e1);
    }
    const float4 a0 = make_float4(e0[0], e0[1], e0[2], e0[3]), a1 = make_float4(e1[0], e1[1], e1[2], e1[3]);
    float4 b0 = __ldg(v4 + lane), b1 = __ldg(v4 + lane + 32);
    b0.x *= iv; b0.y *= iv; b0.z *= iv; b0.w *= iv; b1.x *= iv; b1.y *= iv; b1.z *= iv; b1.w *= iv;   // v_hat
    float dot = a0.x * b0.x;
    dot = fmaf(a0.y, b0.y, dot); dot = fmaf(a0.z, b0.z, dot); dot = fmaf(a0.w, b0.w, dot);
    dot = fmaf(a1.x, b1.x, dot); dot = fmaf(a1.y, b1.y, dot); dot = fmaf(a1.z, b1.z, dot); dot = fmaf(a1.w, b1.w, dot);
    dot = warp_sum(dot);
    if (dg && lane == 0) dg[row] = dot;
    dv4[lane] = make_float4(sc * (a0.x - dot * b0.x), sc * (a0.y - dot * b0.y), sc * (a0.z - dot * b0.z), sc * (a0.w - dot * b0.w));
    dv4[lane + 32] = make_float4(sc * (a1.x - dot * b1.x), sc * (a1.y - dot * b1.y), sc * (a1.z - dot * b1.z), sc * (a1.w - dot * b1.w));
    continue;
  }
  float dot = 0.f;
  for (int c = lane; c < dim; c += 32) dot = fmaf(static_cast<float>(dw[row * dim + c]), v[row * dim + c] * iv, dot);
  dot = warp_sum(dot);                                 // dW . v_hat
  if (dg && lane == 0) dg[row] = dot;
  for (int c = lane; c < dim; c += 32) {
    const long long o = row * dim + c;
    dv[o] = sc * (static_cast<float>(dw[o]) - dot * (v[o] * iv));
  }
  }
}

// gmax = max_k |g_k| (single block; K floats are a few hundred KB at most).
__global__ void __launch_bounds__(1024)
absmax_kernel(const float* __restrict__ g, long long K, float* __restrict__ out) {
  pdl_prologue();
  __shared__ float red[32];
  float m = 0.f;
  for (long long i = threadIdx.x; i < K; i += 1024) m = fmaxf(m, fabsf(g[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = warp_max(red[threadIdx.x]);
    if (threadIdx.x == 0) *out = m;
  }
}

__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long long n) {
  pdl_prologue();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    const float h = tf32_round(v);
    hi[i] = h;
    lo[i] = tf32_round(v - h);
  }
}

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  pdl_prologue();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long n4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0) ? n / 4 : 0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
  for (long long i = n4 * 4 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = __float2bfloat16_rn(x[i]);
}

// Up to 8 independent fp32 -> bf16 casts in one launch (blockIdx.y = tensor): the operand copies of one head forward.
struct CastBatch {
  const float* src[8];
  __nv_bfloat16* dst[8];
  long long n[8];
};
__global__ void __launch_bounds__(256)
cast_bf16_batch_kernel(const CastBatch b) {
  pdl_prologue();
  const float* __restrict__ x = b.src[blockIdx.y];
  __nv_bfloat16* __restrict__ y = b.dst[blockIdx.y];
  const long long n = b.n[blockIdx.y];
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long n4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0) ? n / 4 : 0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
  for (long long i = n4 * 4 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = __float2bfloat16_rn(x[i]);
}

// The way back: up to 8 bf16 -> fp32 widenings in one launch (gradients after a bf16 all-reduce).
struct WidenBatch {
  const __nv_bfloat16* src[8];
  float* dst[8];
  long long n[8];
};
__global__ void __launch_bounds__(256)
widen_bf16_batch_kernel(const WidenBatch b) {
  pdl_prologue();
  const __nv_bfloat16* __restrict__ x = b.src[blockIdx.y];
  float* __restrict__ y = b.dst[blockIdx.y];
  const long long n = b.n[blockIdx.y];
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long n4 = ((reinterpret_cast<uintptr_t>(x) & 7) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) ? n / 4 : 0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint2 r = reinterpret_cast<const uint2*>(x)[i];
    reinterpret_cast<float4*>(y)[i] = make_float4(bf16_lo(r.x), bf16_hi(r.x), bf16_lo(r.y), bf16_hi(r.y));
  }
  for (long long i = n4 * 4 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = __bfloat162float(x[i]);
}

// Column sum in two deterministic steps: [row_splits][N] partials, then a fixed-order reduction.
// A thread owns 4 consecutive columns (one packed load per row), a block = 32 column groups x 8 row lanes.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const T* __restrict__ X, long long M, int N, long long ld, int rows_per_split, bool vec_ok,
                      float* __restrict__ partial) {
  pdl_prologue();
  using Q4 = Quad<T>;
  __shared__ float red[8][32][5];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const long long col = (static_cast<long long>(blockIdx.x) * 32 + cx) * 4;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_split;
  const long long r1 = min(r0 + rows_per_split, M);
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  if (col < N) {
    const bool fast = vec_ok && (col + 4 <= N);
    for (long long r = r0 + ry; r < r1; r += 8) {
      float x[4];
      const T* p = X + r * ld + col;
      Q4::unpack(fast ? Q4::load(p) : Q4::load_guard(p, col, N), x);
#pragma unroll
      for (int e = 0; e < 4; ++e) s[e] += x[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) red[ry][cx][e] = s[e];
  __syncthreads();
  if (ry == 0 && col < N) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][cx][e];
      if (col + e < N) partial[static_cast<long long>(blockIdx.y) * N + col + e] = t;
    }
  }
}
__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ partial, int splits, int N, float* __restrict__ out) {
  pdl_prologue();
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= N) return;
  float t = 0.f;
  for (int s = 0; s < splits; ++s) t += partial[static_cast<long long>(s) * N + col];
  out[col] = t;
}

int colsum_splits(int64_t M, int64_t N) {
  const int64_t col_blocks = ceil_div(N, 128);
  int64_t s = ceil_div(4 * kNumSMs, col_blocks);
  if (s > ceil_div(M, 16)) s = ceil_div(M, 16);
  if (s < 1) s = 1;
  if (s > 1024) s = 1024;
  return static_cast<int>(s);
}

int grid_1d(long long n, int per_block) {
  long long b = ceil_div(n, per_block);
  const long long cap = static_cast<long long>(kNumSMs) * 16;
  return static_cast<int>(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace
}  // namespace dmc

using namespace dmc;

extern "C" int dmc_normalize_rows_fwd(const void* z, int32_t z_dtype, int64_t n_rows, int64_t dim, int64_t ld, float eps,
                                      float* zhat_f32, void* zhat_bf16, float* zhat_lo, float* inv_den, void* stream) {
  DMC_REQUIRE(z && inv_den, "dmc_normalize_rows_fwd: null pointer");
  DMC_REQUIRE(n_rows > 0 && dim > 0 && dim < (1 << 30) && ld >= dim, "dmc_normalize_rows_fwd: bad shape n_rows=%lld dim=%lld ld=%lld", (long long)n_rows, (long long)dim, (long long)ld);
  DMC_REQUIRE(z_dtype == DMC_F32 || z_dtype == DMC_BF16, "dmc_normalize_rows_fwd: bad dtype");
  DMC_REQUIRE(zhat_f32 || zhat_bf16, "dmc_normalize_rows_fwd: no output requested");
  launch_kernel(normalize_fwd_kernel, dim3((unsigned)ceil_div(n_rows, kWarpsPerBlock)), dim3(kWarpsPerBlock * 32), 0, (cudaStream_t)stream, z, z_dtype, n_rows, (int)dim, ld, eps, zhat_f32, static_cast<__nv_bfloat16*>(zhat_bf16), zhat_lo, inv_den);
  DMC_LAUNCH_CHECK("normalize_fwd_kernel launch");
  return 0;
}

extern "C" int dmc_normalize_rows_bwd(const float* dzhat, const float* zhat, const float* inv_den, int64_t n_rows,
                                      int64_t dim, float eps, void* dz, int32_t dz_dtype, void* stream) {
  DMC_REQUIRE(dzhat && zhat && inv_den && dz, "dmc_normalize_rows_bwd: null pointer");
  DMC_REQUIRE(n_rows > 0 && dim > 0 && dim < (1 << 30), "dmc_normalize_rows_bwd: bad shape");
  DMC_REQUIRE(dz_dtype == DMC_F32 || dz_dtype == DMC_BF16, "dmc_normalize_rows_bwd: bad dtype");
  launch_kernel(normalize_bwd_kernel, dim3((unsigned)ceil_div(n_rows, kWarpsPerBlock)), dim3(kWarpsPerBlock * 32), 0, (cudaStream_t)stream, dzhat, zhat, inv_den, n_rows, (int)dim, eps, dz, dz_dtype);
  DMC_LAUNCH_CHECK("normalize_bwd_kernel launch");
  return 0;
}

extern "C" int dmc_weightnorm_fwd(const float* v, const float* g, int64_t K, int64_t dim, float* w_f32, float* w_lo,
                                  void* w_bf16, float* scale, float* inv_vnorm, float* gmax, void* stream) {
  DMC_REQUIRE(v && g && scale && inv_vnorm, "dmc_weightnorm_fwd: null pointer");
  DMC_REQUIRE(K > 0 && dim > 0 && dim < (1 << 30), "dmc_weightnorm_fwd: bad shape");
  auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec_ok = (dim % 4 == 0) && al16(v) && al16(w_f32) && al16(w_lo) && al16(w_bf16);
  if (gmax != nullptr) {
    launch_kernel(absmax_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, g, K, gmax);
    DMC_LAUNCH_CHECK("absmax_kernel launch");
  }
  const int64_t rows_per_block = (vec_ok && dim == 256) ? kWarpsPerBlock * kRowsPerWarp : kWarpsPerBlock;
  launch_kernel(weightnorm_fwd_kernel, dim3(streaming_grid(ceil_div(K, rows_per_block))), dim3(kWarpsPerBlock * 32), 0, (cudaStream_t)stream, v, g, K, (int)dim, w_f32, w_lo, static_cast<__nv_bfloat16*>(w_bf16), scale, inv_vnorm, vec_ok, gmax);
  DMC_LAUNCH_CHECK("weightnorm_fwd_kernel launch");
  return 0;
}

extern "C" int dmc_weightnorm_bwd(const float* dw, const float* v, const float* scale, const float* inv_vnorm,
                                  int64_t K, int64_t dim, float* dv, float* dg, void* stream) {
  DMC_REQUIRE(dw && v && scale && inv_vnorm && dv, "dmc_weightnorm_bwd: null pointer");
  DMC_REQUIRE(K > 0 && dim > 0 && dim < (1 << 30), "dmc_weightnorm_bwd: bad shape");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec_ok = (dim % 4 == 0) && al16(dw) && al16(v) && al16(dv);
  launch_kernel(weightnorm_bwd_kernel<float>, dim3(streaming_grid(ceil_div(K, kWarpsPerBlock))), dim3(kWarpsPerBlock * 32), 0, (cudaStream_t)stream, dw, v, scale, inv_vnorm, K, (int)dim, dv, dg, vec_ok);
  DMC_LAUNCH_CHECK("weightnorm_bwd_kernel launch");
  return 0;
}

extern "C" int dmc_weightnorm_bwd_bf16(const void* dw_bf16, const float* v, const float* scale, const float* inv_vnorm,
                                       int64_t K, int64_t dim, float* dv, float* dg, void* stream) {
  DMC_REQUIRE(dw_bf16 && v && scale && inv_vnorm && dv, "dmc_weightnorm_bwd_bf16: null pointer");
  DMC_REQUIRE(K > 0 && dim > 0 && dim < (1 << 30), "dmc_weightnorm_bwd_bf16: bad shape");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec_ok = (dim % 4 == 0) && al16(dw_bf16) && al16(v) && al16(dv);
  launch_kernel(weightnorm_bwd_kernel<__nv_bfloat16>, dim3(streaming_grid(ceil_div(K, kWarpsPerBlock))), dim3(kWarpsPerBlock * 32), 0, (cudaStream_t)stream,
                static_cast<const __nv_bfloat16*>(dw_bf16), v, scale, inv_vnorm, K, (int)dim, dv, dg, vec_ok);
  DMC_LAUNCH_CHECK("weightnorm_bwd_kernel<bf16> launch");
  return 0;
}

extern "C" int dmc_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream) {
  DMC_REQUIRE(x && hi && lo && n > 0, "dmc_split_tf32: bad arguments");
  launch_kernel(split_tf32_kernel, dim3(grid_1d(n, 256)), dim3(256), 0, (cudaStream_t)stream, x, hi, lo, n);
  DMC_LAUNCH_CHECK("split_tf32_kernel launch");
  return 0;
}

extern "C" int dmc_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream) {
  DMC_REQUIRE(x && y && n > 0, "dmc_cast_f32_to_bf16: bad arguments");
  launch_kernel(cast_bf16_kernel, dim3(grid_1d(n, 1024)), dim3(256), 0, (cudaStream_t)stream, x, static_cast<__nv_bfloat16*>(y), n);
  DMC_LAUNCH_CHECK("cast_bf16_kernel launch");
  return 0;
}

extern "C" int dmc_cast_f32_to_bf16_batch(const float* const* srcs_host, void* const* dsts_host, const int64_t* ns_host,
                                          int32_t count, void* stream) {
  DMC_REQUIRE(srcs_host && dsts_host && ns_host && count >= 1 && count <= 8, "dmc_cast_f32_to_bf16_batch: 1..8 tensors per call");
  CastBatch b{};
  long long nmax = 0;
  for (int i = 0; i < count; ++i) {
    DMC_REQUIRE(srcs_host[i] && dsts_host[i] && ns_host[i] > 0, "dmc_cast_f32_to_bf16_batch: bad tensor %d", i);
    b.src[i] = srcs_host[i]; b.dst[i] = static_cast<__nv_bfloat16*>(dsts_host[i]); b.n[i] = ns_host[i];
    nmax = ns_host[i] > nmax ? ns_host[i] : nmax;
  }
  dim3 grid((unsigned)grid_1d(nmax, 1024), (unsigned)count);
  launch_kernel(cast_bf16_batch_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, b);
  DMC_LAUNCH_CHECK("cast_bf16_batch_kernel launch");
  return 0;
}

extern "C" int dmc_cast_bf16_to_f32_batch(const void* const* srcs_host, float* const* dsts_host, const int64_t* ns_host,
                                          int32_t count, void* stream) {
  DMC_REQUIRE(srcs_host && dsts_host && ns_host && count >= 1 && count <= 8, "dmc_cast_bf16_to_f32_batch: 1..8 tensors per call");
  WidenBatch b{};
  long long nmax = 0;
  for (int i = 0; i < count; ++i) {
    DMC_REQUIRE(srcs_host[i] && dsts_host[i] && ns_host[i] > 0, "dmc_cast_bf16_to_f32_batch: bad tensor %d", i);
    b.src[i] = static_cast<const __nv_bfloat16*>(srcs_host[i]); b.dst[i] = dsts_host[i]; b.n[i] = ns_host[i];
    nmax = ns_host[i] > nmax ? ns_host[i] : nmax;
  }
  dim3 grid((unsigned)grid_1d(nmax, 1024), (unsigned)count);
  launch_kernel(widen_bf16_batch_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, b);
  DMC_LAUNCH_CHECK("widen_bf16_batch_kernel launch");
  return 0;
}

extern "C" size_t dmc_colsum_workspace_bytes(int64_t M, int64_t N) {
  if (M <= 0 || N <= 0) return 0;
  return static_cast<size_t>(colsum_splits(M, N)) * N * sizeof(float);
}

extern "C" int dmc_colsum(const void* X, int32_t dtype, int64_t M, int64_t N, int64_t ld, float* out, void* workspace,
                          size_t workspace_bytes, void* stream) {
  DMC_REQUIRE(X && out && workspace, "dmc_colsum: null pointer");
  DMC_REQUIRE(M > 0 && N > 0 && N < (1ll << 31) && ld >= N, "dmc_colsum: bad shape");
  DMC_REQUIRE(dtype == DMC_F32 || dtype == DMC_BF16, "dmc_colsum: bad dtype");
  const int splits = colsum_splits(M, N);
  DMC_REQUIRE(workspace_bytes >= static_cast<size_t>(splits) * N * sizeof(float), "dmc_colsum: workspace too small");
  const int rows_per_split = static_cast<int>(ceil_div(M, splits));
  dim3 grid((unsigned)ceil_div(N, 128), (unsigned)splits);
  const int esz = dtype == DMC_BF16 ? 2 : 4;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(X) % (4 * esz)) == 0) && ((ld * esz) % (4 * esz) == 0);
  if (dtype == DMC_BF16)
    launch_kernel(colsum_partial_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, static_cast<const __nv_bfloat16*>(X), M, (int)N, ld,
                                                                                rows_per_split, vec_ok, static_cast<float*>(workspace));
  else
    launch_kernel(colsum_partial_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, static_cast<const float*>(X), M, (int)N, ld, rows_per_split,
                                                                        vec_ok, static_cast<float*>(workspace));
  DMC_LAUNCH_CHECK("colsum_partial_kernel launch");
  launch_kernel(colsum_final_kernel, dim3((unsigned)ceil_div(N, 256)), dim3(256), 0, (cudaStream_t)stream, static_cast<const float*>(workspace), splits, (int)N, out);
  DMC_LAUNCH_CHECK("colsum_final_kernel launch");
  return 0;
}

extern "C" int dmc_absmax(const float* x, int64_t n, float* out, void* stream) {
  DMC_REQUIRE(x && out && n > 0, "dmc_absmax: bad arguments");
  launch_kernel(absmax_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, x, n, out);
  DMC_LAUNCH_CHECK("absmax_kernel launch");
  return 0;
}
