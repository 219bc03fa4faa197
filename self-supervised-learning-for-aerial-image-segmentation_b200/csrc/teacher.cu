// teacher.cu -- fused teacher kernels of DINOLoss.
//
//   main_dino_mc.py:446   softmax((teacher_output - center) / temp)     -> per-row (max * log2(e), 1/sum exp)
//   main_dino_mc.py:468   torch.sum(teacher_output, dim=0)              -> per-GPU batch column sum
//   main_dino_mc.py:470-473  center EMA                                  -> dmc_center_update
//
// teacher_pass_kernel reads every teacher logit exactly once and produces BOTH reductions: each warp
// walks rows of a 512-column chunk four rows at a time (16 packed loads in flight per lane), a lane keeps
// its 16 columns' running sums in registers (column direction) and reduces max / sum-exp across the warp
// with shuffles (row direction).  Partials go to a
// small workspace and are merged in a fixed order (deterministic, no atomics).
// HBM-bound: algorithmic bytes = Nt*K*sizeof(logit) read; everything else is O(Nt + K).
#include <math.h>
#include <stdlib.h>

#include "dmc_common.cuh"

namespace dmc {
namespace {

constexpr int kChunk = 512;      // columns per CTA: a lane owns 4 vectors of 4 consecutive columns
constexpr int kWarps = 8;
constexpr int kRowsInFlight = 4; // max rows a warp loads before it starts reducing (<= 16 packed loads in flight per lane)
constexpr int kNV = kChunk / (32 * 4);

// FULL: the whole 512-column chunk lies inside K and is aligned (packed loads, no per-element guards).
template <typename T> struct RowsInFlight { static constexpr int value = 16 / sizeof(T) / 2; };   // 4 (bf16) / 2 (fp32)

// FIXED: every y2 = (t - center) * ct is known to lie below `shift` (caller-supplied bounds on |t| and |center|): the sum
// 2^(y2 - shift) needs no running maximum -- ONE pass over the registers, no warp max reduction, half the instructions
// (the pass is instruction-bound).  `cb` then already holds -center * ct - shift and the reported maximum is `shift`.
template <typename T, bool FULL, bool FIXED>
__device__ __forceinline__ void teacher_rows(const T* __restrict__ t, long long K, long long ld, long long col0, int lane,
                                             long long r0, long long r_end, const float (&cb)[kNV][4], float ct,
                                             float (&cs)[kNV][4], float2* __restrict__ ws_stats, int nchunks, int chunk, float shift) {
  using Q4 = Quad<T>;
  constexpr int RIF = RowsInFlight<T>::value;
  typename Q4::Raw raw[RIF][kNV];
#pragma unroll
  for (int j = 0; j < RIF; ++j) {
    const long long r = r0 + j;
    if (r < r_end) {
      const T* rowp = t + r * ld;
#pragma unroll
      for (int i = 0; i < kNV; ++i) {
        const long long c = col0 + (i * 32 + lane) * 4;
        raw[j][i] = FULL ? Q4::load(rowp + c) : Q4::load_guard(rowp + c, c, K);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < RIF; ++j) {
    const long long r = r0 + j;
    if (FIXED && r < r_end) {                            // warp-uniform
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int i = 0; i < kNV; ++i) {
        float x[4];
        Q4::unpack(raw[j][i], x);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          cs[i][e] += x[e];
          float v = fmaf(x[e], ct, cb[i][e]);            // (t - center) / temp * log2(e) - shift
          if (!FULL && (col0 + (i * 32 + lane) * 4 + e >= K)) v = -INFINITY;
          if (e & 1) l1 += ex2(v); else l0 += ex2(v);
        }
      }
      const float l = warp_sum(l0 + l1);
      if (lane == 0) ws_stats[r * nchunks + chunk] = make_float2(shift, l);
    } else if (r < r_end) {                              // warp-uniform
      float m = -INFINITY;
#pragma unroll
      for (int i = 0; i < kNV; ++i) {
        float x[4];
        Q4::unpack(raw[j][i], x);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          cs[i][e] += x[e];
          float v = fmaf(x[e], ct, cb[i][e]);            // (t - center) / temp * log2(e)
          if (!FULL && (col0 + (i * 32 + lane) * 4 + e >= K)) v = -INFINITY;
          m = fmaxf(m, v);
        }
      }
      m = warp_max(m);
      float l = 0.f;
#pragma unroll
      for (int i = 0; i < kNV; ++i) {                    // second pass over the packed registers (one FFMA to redo)
        float x[4];
        Q4::unpack(raw[j][i], x);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = fmaf(x[e], ct, cb[i][e]) - m;
          if (!FULL && (col0 + (i * 32 + lane) * 4 + e >= K)) v = -INFINITY;
          l += ex2(v);
        }
      }
      l = warp_sum(l);
      if (lane == 0) ws_stats[r * nchunks + chunk] = make_float2(m, l);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kWarps * 32, 2)
teacher_pass_kernel(const T* __restrict__ t, long long Nt, long long K, long long ld, const float* __restrict__ center,
                    float inv_temp, float2* __restrict__ ws_stats, float* __restrict__ ws_colsum, int rows_per_block,
                    int nchunks, bool vec_ok, const float* __restrict__ bounds) {
  pdl_prologue();
  __shared__ __align__(16) float sm[kWarps][kChunk];  // 16 KiB: per-warp column sums
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x;
  const long long col0 = static_cast<long long>(chunk) * kChunk;
  const bool full = vec_ok && (col0 + kChunk <= K);
  const float ct = inv_temp * kLog2e;
  float shift = 0.f;
  bool fixed = false;
  if (bounds != nullptr) {                             // {max |t|, max |center|}: fixed shift while 2^(-2 shift) stays a normal fp32
    shift = (__ldg(bounds) + __ldg(bounds + 1)) * fabsf(ct) * 1.01f + 0.05f;
    fixed = shift < 55.f;
  }

  float cb[kNV][4], cs[kNV][4];                        // cb = -center * ct (folded into one FFMA per logit)
#pragma unroll
  for (int i = 0; i < kNV; ++i) {
    const long long c = col0 + (i * 32 + lane) * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      cb[i][e] = ((c + e < K) ? -__ldg(center + c + e) * ct : 0.f) - (fixed ? shift : 0.f);
      cs[i][e] = 0.f;
    }
  }
  const long long r_begin = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r_end = min(r_begin + rows_per_block, Nt);
  constexpr int RIF = RowsInFlight<T>::value;
  for (long long r0 = r_begin + warp * RIF; r0 < r_end; r0 += kWarps * RIF) {
    if (full && fixed) teacher_rows<T, true, true>(t, K, ld, col0, lane, r0, r_end, cb, ct, cs, ws_stats, nchunks, chunk, shift);
    else if (full) teacher_rows<T, true, false>(t, K, ld, col0, lane, r0, r_end, cb, ct, cs, ws_stats, nchunks, chunk, shift);
    else if (fixed) teacher_rows<T, false, true>(t, K, ld, col0, lane, r0, r_end, cb, ct, cs, ws_stats, nchunks, chunk, shift);
    else teacher_rows<T, false, false>(t, K, ld, col0, lane, r0, r_end, cb, ct, cs, ws_stats, nchunks, chunk, shift);
  }
#pragma unroll
  for (int i = 0; i < kNV; ++i)
    *reinterpret_cast<float4*>(&sm[warp][(i * 32 + lane) * 4]) = make_float4(cs[i][0], cs[i][1], cs[i][2], cs[i][3]);
  __syncthreads();
  for (int c = threadIdx.x; c < kChunk; c += kWarps * 32) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) acc += sm[w][c];
    if (col0 + c < K) ws_colsum[static_cast<long long>(blockIdx.y) * K + col0 + c] = acc;
  }
}

// blocks [0, row_blocks): one warp per teacher row merges the per-chunk (max, sum) partials;
// blocks [row_blocks, ...): one thread per column adds the per-row-block column sums in fixed order.
__global__ void __launch_bounds__(256)
teacher_finalize_kernel(const float2* __restrict__ ws_stats, const float* __restrict__ ws_colsum, long long Nt, long long K,
                        int nchunks, int nrb, int row_blocks, float2* __restrict__ row_stats, float* __restrict__ colsum) {
  pdl_prologue();
  if (static_cast<int>(blockIdx.x) < row_blocks) {
    const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (r >= Nt) return;
    const int lane = threadIdx.x & 31;
    float m = -INFINITY, l = 0.f;                    // base-2 domain: l = sum 2^(y2 - m)
    for (int c = lane; c < nchunks; c += 32) {
      const float2 p = ws_stats[r * nchunks + c];
      online_merge2(m, l, p.x, p.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
      const float l2 = __shfl_xor_sync(0xffffffffu, l, o);
      online_merge2(m, l, m2, l2);
    }
    if (lane == 0) row_stats[r] = make_float2(m, 1.0f / l);
  } else {
    const long long k = static_cast<long long>(blockIdx.x - row_blocks) * 256 + threadIdx.x;
    if (k >= K) return;
    float s = 0.f;
    for (int b = 0; b < nrb; ++b) s += ws_colsum[static_cast<long long>(b) * K + k];
    colsum[k] = s;
  }
}

__global__ void __launch_bounds__(256)
center_update_kernel(const float* center_in, float* center_out, const float* __restrict__ colsum, long long K, float count, float mom, float omm) {
  pdl_prologue();
  const long long k = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (k >= K) return;
  const float bc = __fdiv_rn(colsum[k], count);                        // batch_center / (len * world)
  center_out[k] = __fadd_rn(__fmul_rn(center_in[k], mom), __fmul_rn(bc, omm));  // center * m + bc * (1 - m), no FMA
}

// out[k] = sum_j W[k, j] * x[j]  (W row-major [K, dim] bf16 or fp32, x fp32 [dim]); one warp per row, grid-stride.
// With x = the column sum of the teacher's normalised bottleneck rows this is the per-GPU batch column sum of the teacher
// logits (main_dino_mc.py:468: sum_rows (zhat_r . W_k) = (sum_rows zhat_r) . W_k) without a pass over the [Nt, K] logits.
template <typename T>
__global__ void __launch_bounds__(256)
rowdot_kernel(const T* __restrict__ W, const float* __restrict__ x, long long K, int dim, float* __restrict__ out, bool vec_ok) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  for (long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5); row < K; row += static_cast<long long>(gridDim.x) * 8) {
    const T* wr = W + row * dim;
    float acc = 0.f;
    if (vec_ok) {
      constexpr int N = Vec<T>::N;
      for (int c = lane * N; c < dim; c += 32 * N) {
        float w[N];
        Vec<T>::load(wr + c, w);
#pragma unroll
        for (int e = 0; e < N; ++e) acc = fmaf(w[e], __ldg(x + c + e), acc);
      }
    } else {
      for (int c = lane; c < dim; c += 32) acc = fmaf(Vec<T>::load1(wr + c), __ldg(x + c), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[row] = acc;
  }
}

struct TeacherPlan { int nchunks, nrb, rows_per_block; size_t stats_bytes, colsum_bytes; };

TeacherPlan teacher_plan(int64_t Nt, int64_t K) {
  TeacherPlan p{};
  p.nchunks = static_cast<int>(ceil_div(K, kChunk));
  const int64_t rows_unit = kWarps * kRowsInFlight;
  // grid size target: 8 CTAs per SM's worth of (chunk, row block) pairs.  DMC_TEACHER_TARGET_CTAS (read once) overrides it
  // for timing experiments: fewer, longer-lived CTAs amortise the per-CTA prologue (center loads, smem merge).
  static const int64_t target_ctas = [] {
    const char* e = getenv("DMC_TEACHER_TARGET_CTAS");
    const long v = e ? strtol(e, nullptr, 10) : 0;
    return static_cast<int64_t>(v > 0 ? v : 8 * kNumSMs);
  }();
  int64_t nrb = ceil_div(target_ctas, p.nchunks);
  const int64_t max_rb = ceil_div(Nt, rows_unit);
  if (nrb > max_rb) nrb = max_rb;
  if (nrb < 1) nrb = 1;
  int64_t rpb = ceil_div(Nt, nrb);
  rpb = ceil_div(rpb, rows_unit) * rows_unit;
  p.rows_per_block = static_cast<int>(rpb);
  p.nrb = static_cast<int>(ceil_div(Nt, rpb));
  p.stats_bytes = (static_cast<size_t>(Nt) * p.nchunks * sizeof(float2) + 255) & ~static_cast<size_t>(255);
  p.colsum_bytes = static_cast<size_t>(p.nrb) * K * sizeof(float);
  return p;
}

}  // namespace
}  // namespace dmc

using namespace dmc;

extern "C" size_t dmc_teacher_workspace_bytes(int64_t Nt, int64_t K) {
  if (Nt <= 0 || K <= 0) return 0;
  TeacherPlan p = teacher_plan(Nt, K);
  return p.stats_bytes + p.colsum_bytes;
}

static int teacher_stats_impl(const void* t, int32_t dtype, int64_t Nt, int64_t K, int64_t ld, const float* center,
                              float inv_temp, const float* bounds, float* row_stats, float* colsum, void* workspace,
                              size_t workspace_bytes, void* stream) {
  DMC_REQUIRE(t && center && row_stats && colsum && workspace, "dmc_teacher_stats_colsum: null pointer");
  DMC_REQUIRE(Nt > 0 && K > 0 && ld >= K, "dmc_teacher_stats_colsum: bad shape Nt=%lld K=%lld ld=%lld", (long long)Nt, (long long)K, (long long)ld);
  DMC_REQUIRE(dtype == DMC_F32 || dtype == DMC_BF16, "dmc_teacher_stats_colsum: bad dtype %d", dtype);
  TeacherPlan p = teacher_plan(Nt, K);
  DMC_REQUIRE(workspace_bytes >= p.stats_bytes + p.colsum_bytes, "dmc_teacher_stats_colsum: workspace too small (%zu < %zu)",
              workspace_bytes, p.stats_bytes + p.colsum_bytes);
  DMC_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "dmc_teacher_stats_colsum: workspace must be 16-byte aligned");
  DMC_REQUIRE(p.nrb <= 65535, "dmc_teacher_stats_colsum: too many row blocks");
  float2* ws_stats = static_cast<float2*>(workspace);
  float* ws_colsum = reinterpret_cast<float*>(static_cast<char*>(workspace) + p.stats_bytes);
  const int esz = dtype == DMC_BF16 ? 2 : 4;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(t) % (4 * esz)) == 0) && ((ld * esz) % (4 * esz) == 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)p.nchunks, (unsigned)p.nrb);
  if (dtype == DMC_BF16)
    launch_kernel(teacher_pass_kernel<__nv_bfloat16>, dim3(grid), dim3(kWarps * 32), 0, st, static_cast<const __nv_bfloat16*>(t), Nt, K, ld, center, inv_temp,
                                                                     ws_stats, ws_colsum, p.rows_per_block, p.nchunks, vec_ok, bounds);
  else
    launch_kernel(teacher_pass_kernel<float>, dim3(grid), dim3(kWarps * 32), 0, st, static_cast<const float*>(t), Nt, K, ld, center, inv_temp, ws_stats,
                                                             ws_colsum, p.rows_per_block, p.nchunks, vec_ok, bounds);
  DMC_LAUNCH_CHECK("teacher_pass_kernel launch");
  const int row_blocks = static_cast<int>(ceil_div(Nt, 8));
  const int col_blocks = static_cast<int>(ceil_div(K, 256));
  launch_kernel(teacher_finalize_kernel, dim3(row_blocks + col_blocks), dim3(256), 0, st, ws_stats, ws_colsum, Nt, K, p.nchunks, p.nrb, row_blocks,
                                                                   reinterpret_cast<float2*>(row_stats), colsum);
  DMC_LAUNCH_CHECK("teacher_finalize_kernel launch");
  return 0;
}

extern "C" int dmc_teacher_stats_colsum(const void* t, int32_t dtype, int64_t Nt, int64_t K, int64_t ld, const float* center,
                                        float inv_temp, float* row_stats, float* colsum, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  return teacher_stats_impl(t, dtype, Nt, K, ld, center, inv_temp, nullptr, row_stats, colsum, workspace, workspace_bytes, stream);
}

extern "C" int dmc_teacher_stats_colsum_bounded(const void* t, int32_t dtype, int64_t Nt, int64_t K, int64_t ld, const float* center,
                                                float inv_temp, const float* bounds_dev, float* row_stats, float* colsum,
                                                void* workspace, size_t workspace_bytes, void* stream) {
  DMC_REQUIRE(bounds_dev != nullptr, "dmc_teacher_stats_colsum_bounded: null bounds");
  return teacher_stats_impl(t, dtype, Nt, K, ld, center, inv_temp, bounds_dev, row_stats, colsum, workspace, workspace_bytes, stream);
}

extern "C" int dmc_center_update(const float* center_in, float* center_out, const float* colsum, int64_t K, float count,
                                 float momentum, float one_minus_momentum, void* stream) {
  DMC_REQUIRE(center_in && center_out && colsum && K > 0 && count > 0.f, "dmc_center_update: bad arguments");
  launch_kernel(center_update_kernel, dim3((unsigned)ceil_div(K, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), center_in, center_out, colsum, K, count,
                                                                                                 momentum, one_minus_momentum);
  DMC_LAUNCH_CHECK("center_update_kernel launch");
  return 0;
}

extern "C" int dmc_teacher_finalize(const float* row_partials, const float* colsum_partials, int64_t Nt, int64_t K, int64_t parts,
                                    int64_t row_groups, float* row_stats, float* colsum, void* stream) {
  DMC_REQUIRE(row_partials && row_stats, "dmc_teacher_finalize: null pointer");
  DMC_REQUIRE((colsum_partials == nullptr) || colsum, "dmc_teacher_finalize: colsum_partials given without colsum");
  DMC_REQUIRE(Nt > 0 && K > 0 && parts > 0 && parts < (1 << 30) && row_groups < (1 << 30) && (colsum_partials == nullptr || row_groups > 0),
              "dmc_teacher_finalize: bad shape");
  const int row_blocks = static_cast<int>(ceil_div(Nt, 8));
  const int col_blocks = colsum_partials ? static_cast<int>(ceil_div(K, 256)) : 0;      // no column partials: row statistics only
  launch_kernel(teacher_finalize_kernel, dim3(row_blocks + col_blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), reinterpret_cast<const float2*>(row_partials), colsum_partials, Nt, K, (int)parts, (int)row_groups, row_blocks,
      reinterpret_cast<float2*>(row_stats), colsum);
  DMC_LAUNCH_CHECK("teacher_finalize_kernel launch");
  return 0;
}

extern "C" int dmc_rowdot(const void* W, int32_t dtype, int64_t K, int64_t dim, const float* x, float* out, void* stream) {
  DMC_REQUIRE(W && x && out, "dmc_rowdot: null pointer");
  DMC_REQUIRE(K > 0 && dim > 0 && dim < (1 << 30), "dmc_rowdot: bad shape");
  DMC_REQUIRE(dtype == DMC_F32 || dtype == DMC_BF16, "dmc_rowdot: bad dtype %d", dtype);
  const int n = (dtype == DMC_BF16) ? 8 : 4;
  const bool vec_ok = (dim % n == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  const unsigned grid = streaming_grid(ceil_div(K, 8));
  if (dtype == DMC_BF16)
    launch_kernel(rowdot_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(W), x, K, (int)dim, out, vec_ok);
  else
    launch_kernel(rowdot_kernel<float>, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const float*>(W), x, K, (int)dim, out, vec_ok);
  DMC_LAUNCH_CHECK("rowdot_kernel launch");
  return 0;
}
