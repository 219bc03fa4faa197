// ce.cu -- the crop-pair cross-entropy of DINOLoss (main_dino_mc.py:441-459) and its backward.
//
// The reference evaluates G*C - min(G,C) (teacher view, student view) pairs, each with its own
// log_softmax over [B,K] (14 logit-sized passes forward, the same again in autograd, 14 saved [B,K]
// log-prob tensors).  Here the closed form is used:
//
//   L = 1/(n B) sum_b [ sum_v n_v lse_v[b]  -  sum_k ( Q[b,k] S[b,k] - sum_{i<min(G,C)} q_i[b,k] x_i[b,k] ) ]
//   dL/ds_v[b,k] = (n_v p_v[b,k] - Q[b,k] + [v<G] q_v[b,k]) / (n B tau_s)
//
// with x_v = s_v/tau_s, lse_v = logsumexp_k x_v, p_v = softmax(x_v), q_i = softmax((t_i - c)/tau_t),
// Q = sum_i q_i, S = sum_v x_v, n_v = G - [v<G].  A CTA owns (sample b, column chunk): it loads the C
// student and G teacher vectors of that chunk once, so every logit is read exactly once per pass and
// the gradient is written exactly once.  HBM-bound streaming kernels: packed 4-element loads, per-thread online
// softmax statistics, warp-shuffle + shared-memory block reductions, deterministic partials (no atomics).
#include <math.h>

#include "dmc_common.cuh"

namespace dmc {
namespace {

constexpr int kThreads = 256;
constexpr int kIters = 8;          // 4-element vectors per thread per row per CTA
constexpr int kChunkCols = kThreads * 4 * kIters;   // 8192 columns per CTA
#ifndef DMC_FUSED_MINBLOCKS
#define DMC_FUSED_MINBLOCKS 2
#endif
constexpr int kMinBlocks = 3;      // CTAs per SM the register budget is tuned for (<= 85 registers/thread)

struct CeArgs {
  const void* s; long long lds;
  const void* t; long long ldt;
  const float* center; const float2* t_stats;
  long long B, K;
  int C, G;
  float inv_ts, inv_tt;
  int nchunks; bool vec_ok;
  // forward
  float2* ws_s; float* ws_x;
  // backward
  const float* s_lse; const float* gout; float coef;
  void* ds; long long ldds;
};

__device__ __forceinline__ void load_center4(const float* center, long long col, long long K, bool fast, float (&c)[4]) {
  if (fast) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(center + col));
    c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) c[e] = (col + e < K) ? __ldg(center + col + e) : 0.f;
  }
}

// Per-thread running softmax statistics of the student rows, in the base-2 domain:
//   m[v] = max of the RAW logits seen so far, l[v] = sum 2^{(x - m[v]) * c2},  c2 = log2(e)/tau_s.
template <int MAXC>
struct RowStats {
  float m[MAXC], l[MAXC];
};

// The packed logits of one 4-column vector: C student rows + G teacher rows of sample b, plus the center.
template <typename T, int MAXC, int MAXG>
struct CeVec {
  typename Quad<T>::Raw rs[MAXC], rt[MAXG];
  float cen[4];
};

// FAST: the vector is fully inside K and aligned (packed loads, no per-element guards); !FAST: ragged row end.
template <typename T, int MAXC, int MAXG, bool FAST>
__device__ __forceinline__ void ce_load(const CeArgs& a, const T* s, const T* t, long long b, long long col, int C, int G,
                                        CeVec<T, MAXC, MAXG>& v) {
  using Q4 = Quad<T>;
#pragma unroll
  for (int i = 0; i < MAXG; ++i)
    if (i < G) {
      const T* p = t + (i * a.B + b) * a.ldt + col;
      v.rt[i] = FAST ? Q4::load(p) : Q4::load_guard(p, col, a.K);
    }
#pragma unroll
  for (int r = 0; r < MAXC; ++r)
    if (r < C) {
      const T* p = s + (r * a.B + b) * a.lds + col;
      v.rs[r] = FAST ? Q4::load(p) : Q4::load_guard(p, col, a.K);
    }
  load_center4(a.center, col, a.K, FAST && ((reinterpret_cast<uintptr_t>(a.center) & 15) == 0), v.cen);
}

template <typename T, int MAXC, int MAXG, bool FAST>
__device__ __forceinline__ void ce_fwd_compute(const CeArgs& a, const CeVec<T, MAXC, MAXG>& in, long long col, int C, int G,
                                               const float (&tmc)[MAXG], const float (&tinv)[MAXG], float c2, float ct,
                                               RowStats<MAXC>& st, float& cross) {
  using Q4 = Quad<T>;
  float Q[4] = {0.f, 0.f, 0.f, 0.f}, S[4] = {0.f, 0.f, 0.f, 0.f};     // S = sum_v RAW student logits
  float qx = 0.f;                                                       // sum_{v<G} q_v . s_v (raw)
#pragma unroll
  for (int v = 0; v < MAXC; ++v) {
    if (v < C) {
      float x[4];
      Q4::unpack(in.rs[v], x);
      float vm;
      if (FAST) {
        vm = fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3]));
      } else {
        vm = -INFINITY;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (col + e < a.K) vm = fmaxf(vm, x[e]);
      }
      if (vm > st.m[v]) { st.l[v] *= ex2((st.m[v] - vm) * c2); st.m[v] = vm; }     // rare after the first vectors
      const float mb = -st.m[v] * c2;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        S[e] += x[e];                                                   // padded columns hold 0 and q = 0 there
        const float p = ex2(fmaf(x[e], c2, mb));
        st.l[v] += (FAST || col + e < a.K) ? p : 0.f;
      }
      if (v < MAXG && v < G) {                                          // same-view pair is skipped: subtract q_v . x_v
        float tq[4];
        Q4::unpack(in.rt[v < MAXG ? v : 0], tq);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float q = ex2(fmaf(tq[e] - in.cen[e], ct, tmc[v < MAXG ? v : 0])) * tinv[v < MAXG ? v : 0];
          if (!FAST && col + e >= a.K) q = 0.f;
          Q[e] += q;
          qx = fmaf(q, x[e], qx);
        }
      }
    }
  }
  // teacher views without a student row of the same index (only when G > C)
#pragma unroll
  for (int i = 0; i < MAXG; ++i)
    if (i < G && i >= C) {
      float tq[4];
      Q4::unpack(in.rt[i], tq);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float q = ex2(fmaf(tq[e] - in.cen[e], ct, tmc[i])) * tinv[i];
        if (!FAST && col + e >= a.K) q = 0.f;
        Q[e] += q;
      }
    }
  float qs = -qx;
#pragma unroll
  for (int e = 0; e < 4; ++e) qs = fmaf(Q[e], S[e], qs);
  cross = fmaf(qs, a.inv_ts, cross);                                    // raw logits -> x = s / tau_s
}

// CT / GT: compile-time crop counts (0 = runtime, bounded by 16 / 4).
template <typename T, int CT, int GT>
__global__ void __launch_bounds__(kThreads, 2)
ce_fwd_kernel(const CeArgs a) {
  pdl_prologue();
  constexpr int MAXC = CT ? CT : 16, MAXG = GT ? GT : 4;
  const int C = CT ? CT : a.C, G = GT ? GT : a.G;
  const long long b = blockIdx.y;
  const int chunk = blockIdx.x;
  const long long col_begin = static_cast<long long>(chunk) * kChunkCols;
  const long long col_end = min(a.K, col_begin + kChunkCols);
  const T* s = static_cast<const T*>(a.s);
  const T* t = static_cast<const T*>(a.t);
  const float c2 = a.inv_ts * kLog2e, ct = a.inv_tt * kLog2e;

  float tmc[MAXG], tinv[MAXG];                        // -(row max in the base-2 domain) and 1/sum of the teacher rows of sample b
#pragma unroll
  for (int i = 0; i < MAXG; ++i) {
    tmc[i] = 0.f; tinv[i] = 0.f;
    if (i < G) { const float2 st = a.t_stats[i * a.B + b]; tmc[i] = -st.x; tinv[i] = st.y; }
  }
  RowStats<MAXC> st;
#pragma unroll
  for (int v = 0; v < MAXC; ++v) { st.m[v] = -INFINITY; st.l[v] = 0.f; }
  float cross = 0.f;

  constexpr long long kStride = kThreads * 4;
  long long col = col_begin + threadIdx.x * 4;
  if (a.vec_ok) {
    // Software pipeline over the full vectors: the C+G loads of vector i+1 are in flight while vector i is reduced.
    const long long full_end = min(col_end, (a.K / 4) * 4);
    CeVec<T, MAXC, MAXG> va, vb;
    if (col < full_end) ce_load<T, MAXC, MAXG, true>(a, s, t, b, col, C, G, va);
    while (col < full_end) {
      const long long c1 = col + kStride, c2n = col + 2 * kStride;
      if (c1 < full_end) ce_load<T, MAXC, MAXG, true>(a, s, t, b, c1, C, G, vb);
      ce_fwd_compute<T, MAXC, MAXG, true>(a, va, col, C, G, tmc, tinv, c2, ct, st, cross);
      if (c1 >= full_end) { col = c1; break; }
      if (c2n < full_end) ce_load<T, MAXC, MAXG, true>(a, s, t, b, c2n, C, G, va);
      ce_fwd_compute<T, MAXC, MAXG, true>(a, vb, c1, C, G, tmc, tinv, c2, ct, st, cross);
      col = c2n;
    }
  }
  for (; col < col_end; col += kStride) {              // ragged row end / unaligned tensors
    CeVec<T, MAXC, MAXG> v;
    ce_load<T, MAXC, MAXG, false>(a, s, t, b, col, C, G, v);
    ce_fwd_compute<T, MAXC, MAXG, false>(a, v, col, C, G, tmc, tinv, c2, ct, st, cross);
  }

  // ---- block reduction: (m,l) per student row by online merge (natural-log domain), cross by sum ----
  __shared__ float red[kThreads / 32][2 * MAXC + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int v = 0; v < MAXC; ++v)
    if (v < C) {
      float mv = st.m[v] * a.inv_ts, lv = st.l[v];                      // max of x = s/tau_s; l is already sum e^{x - max}
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, mv, o);
        const float l2 = __shfl_xor_sync(0xffffffffu, lv, o);
        online_merge(mv, lv, m2, l2);
      }
      if (lane == 0) { red[warp][2 * v] = mv; red[warp][2 * v + 1] = lv; }
    }
  cross = warp_sum(cross);
  if (lane == 0) red[warp][2 * MAXC] = cross;
  __syncthreads();
  if (threadIdx.x < C) {
    const int v = threadIdx.x;
    float mm = -INFINITY, ll = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) online_merge(mm, ll, red[w][2 * v], red[w][2 * v + 1]);
    a.ws_s[(v * a.B + b) * a.nchunks + chunk] = make_float2(mm, ll);
  }
  if (threadIdx.x == 32) {
    float c = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) c += red[w][2 * MAXC];
    a.ws_x[b * a.nchunks + chunk] = c;
  }
}

// Single CTA: merges the per-chunk partials into s_lse[C*B] and the scalar loss (fixed order).
__global__ void __launch_bounds__(1024)
ce_finalize_kernel(const float2* __restrict__ ws_s, const float* __restrict__ ws_x, long long B, int C, int G, int nchunks,
                   float* __restrict__ s_lse, float* __restrict__ loss) {
  pdl_prologue();
  __shared__ double red[32];
  double acc = 0.0;
  const long long rows = static_cast<long long>(C) * B;
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
    float m = -INFINITY, l = 0.f;
    for (int c = 0; c < nchunks; ++c) { const float2 p = ws_s[r * nchunks + c]; online_merge(m, l, p.x, p.y); }
    const float lse = m + logf(l);
    s_lse[r] = lse;
    const int v = static_cast<int>(r / B);
    acc += static_cast<double>((v < G) ? (G - 1) : G) * static_cast<double>(lse);
  }
  for (long long i = threadIdx.x; i < B * nchunks; i += blockDim.x) acc -= static_cast<double>(ws_x[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    const int n_terms = G * C - (G < C ? G : C);
    loss[0] = static_cast<float>(tot / (static_cast<double>(n_terms) * static_cast<double>(B)));
  }
}

template <typename T, int MAXC, int MAXG, bool FAST>
__device__ __forceinline__ void ce_bwd_vector(const CeArgs& a, const T* s, const T* t, T* ds, long long b, long long col, int C,
                                              int G, const float (&tmc)[MAXG], const float (&tinv)[MAXG],
                                              const float (&lse2)[MAXC], float c2, float ct, float scale) {
  using Q4 = Quad<T>;
  typename Q4::Raw rt[MAXG], rs[MAXC];
#pragma unroll
  for (int i = 0; i < MAXG; ++i)
    if (i < G) {
      const T* p = t + (i * a.B + b) * a.ldt + col;
      rt[i] = FAST ? Q4::load(p) : Q4::load_guard(p, col, a.K);
    }
#pragma unroll
  for (int v = 0; v < MAXC; ++v)
    if (v < C) {
      const T* p = s + (v * a.B + b) * a.lds + col;
      rs[v] = FAST ? Q4::load(p) : Q4::load_guard(p, col, a.K);
    }
  float cen[4];
  load_center4(a.center, col, a.K, FAST && ((reinterpret_cast<uintptr_t>(a.center) & 15) == 0), cen);
  float q[MAXG][4], Q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < MAXG; ++i)
    if (i < G) {
      float tq[4];
      Q4::unpack(rt[i], tq);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        q[i][e] = ex2(fmaf(tq[e] - cen[e], ct, tmc[i])) * tinv[i];
        Q[e] += q[i][e];
      }
    }
#pragma unroll
  for (int v = 0; v < MAXC; ++v)
    if (v < C) {
      const float nvs = scale * static_cast<float>((v < G) ? (G - 1) : G);
      float x[4], d[4];
      Q4::unpack(rs[v], x);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p = ex2(fmaf(x[e], c2, lse2[v]));                   // softmax(s/tau_s)
        float qs = Q[e];
        if (v < MAXG && v < G) qs -= q[v < MAXG ? v : 0][e];
        d[e] = fmaf(nvs, p, -scale * qs);
      }
      T* dst = ds + (v * a.B + b) * a.ldds + col;
      if (FAST) {
        Q4::store(dst, d);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (col + e < a.K) Q4::store1(dst + e, d[e]);
      }
    }
}

template <typename T, int CT, int GT>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
ce_bwd_kernel(const CeArgs a) {
  pdl_prologue();
  constexpr int MAXC = CT ? CT : 16, MAXG = GT ? GT : 4;
  const int C = CT ? CT : a.C, G = GT ? GT : a.G;
  const long long b = blockIdx.y;
  const long long col_begin = static_cast<long long>(blockIdx.x) * kChunkCols;
  const long long col_end = min(a.K, col_begin + kChunkCols);
  const T* s = static_cast<const T*>(a.s);
  const T* t = static_cast<const T*>(a.t);
  T* ds = static_cast<T*>(a.ds);
  const float scale = a.coef * __ldg(a.gout);
  const float c2 = a.inv_ts * kLog2e, ct = a.inv_tt * kLog2e;

  float tmc[MAXG], tinv[MAXG], lse2[MAXC];
#pragma unroll
  for (int i = 0; i < MAXG; ++i) {
    tmc[i] = 0.f; tinv[i] = 0.f;
    if (i < G) { const float2 st = a.t_stats[i * a.B + b]; tmc[i] = -st.x; tinv[i] = st.y; }
  }
#pragma unroll
  for (int v = 0; v < MAXC; ++v) {
    lse2[v] = 0.f;
    if (v < C) lse2[v] = -a.s_lse[v * a.B + b] * kLog2e;
  }
  for (long long col = col_begin + threadIdx.x * 4; col < col_end; col += kThreads * 4) {
    if (a.vec_ok && (col + 4 <= a.K)) ce_bwd_vector<T, MAXC, MAXG, true>(a, s, t, ds, b, col, C, G, tmc, tinv, lse2, c2, ct, scale);
    else ce_bwd_vector<T, MAXC, MAXG, false>(a, s, t, ds, b, col, C, G, tmc, tinv, lse2, c2, ct, scale);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Fused loss + gradient pass.  When the per-row log-sum-exp of the student logits is already known (the GEMM
// epilogue produced it, dmc_gemm stat_row_partials -> dmc_lse_finalize), ONE pass over the logits yields both the
// loss's cross term and the gradient (for an upstream gradient of 1; dmc_scale_inplace_if rescales otherwise),
// i.e. the forward statistics pass over [N_s + N_t, K] disappears.
// ---------------------------------------------------------------------------------------------------------
template <typename T, int MAXC, int MAXG, bool FAST>
__device__ __forceinline__ void ce_fused_compute(const CeArgs& a, const CeVec<T, MAXC, MAXG>& in, T* ds, long long b, long long col,
                                                 int C, int G, const float (&tmc)[MAXG], const float (&tinv)[MAXG],
                                                 const float (&lse2)[MAXC], float c2, float ct, float scale, float& cross) {
  using Q4 = Quad<T>;
  float q[MAXG][4], Q[4] = {0.f, 0.f, 0.f, 0.f}, S[4] = {0.f, 0.f, 0.f, 0.f};
  float qx = 0.f;
#pragma unroll
  for (int i = 0; i < MAXG; ++i)
    if (i < G) {
      float tq[4];
      Q4::unpack(in.rt[i], tq);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float qv = ex2(fmaf(tq[e] - in.cen[e], ct, tmc[i])) * tinv[i];
        if (!FAST && col + e >= a.K) qv = 0.f;
        q[i][e] = qv;
        Q[e] += qv;
      }
    }
#pragma unroll
  for (int v = 0; v < MAXC; ++v)
    if (v < C) {
      const float nvs = scale * static_cast<float>((v < G) ? (G - 1) : G);
      float x[4], d[4];
      Q4::unpack(in.rs[v], x);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        S[e] += x[e];
        const float p = ex2(fmaf(x[e], c2, lse2[v]));                   // softmax(s/tau_s)
        float qs = Q[e];
        if (v < MAXG && v < G) { qs -= q[v < MAXG ? v : 0][e]; qx = fmaf(q[v < MAXG ? v : 0][e], x[e], qx); }
        d[e] = fmaf(nvs, p, -scale * qs);
      }
      T* dst = ds + (v * a.B + b) * a.ldds + col;
      if (FAST) {
        Q4::store(dst, d);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (col + e < a.K) Q4::store1(dst + e, d[e]);
      }
    }
  float qs = -qx;
#pragma unroll
  for (int e = 0; e < 4; ++e) qs = fmaf(Q[e], S[e], qs);
  cross = fmaf(qs, a.inv_ts, cross);
}

template <typename T, int MAXC, int MAXG, bool FAST>
__device__ __forceinline__ void ce_fused_vector(const CeArgs& a, const T* s, const T* t, T* ds, long long b, long long col, int C,
                                                int G, const float (&tmc)[MAXG], const float (&tinv)[MAXG],
                                                const float (&lse2)[MAXC], float c2, float ct, float scale, float& cross) {
  CeVec<T, MAXC, MAXG> in;
  ce_load<T, MAXC, MAXG, FAST>(a, s, t, b, col, C, G, in);
  ce_fused_compute<T, MAXC, MAXG, FAST>(a, in, ds, b, col, C, G, tmc, tinv, lse2, c2, ct, scale, cross);
}

// ALLFAST: every vector of every row is full and aligned (K % 4 == 0, aligned pointers and strides): the ragged path is
// not even compiled in, which keeps the hot kernel inside its register budget.
template <typename T, int CT, int GT, bool ALLFAST>
__global__ void __launch_bounds__(kThreads, DMC_FUSED_MINBLOCKS)
ce_fused_kernel(const CeArgs a) {
  pdl_prologue();
  constexpr int MAXC = CT ? CT : 16, MAXG = GT ? GT : 4;
  const int C = CT ? CT : a.C, G = GT ? GT : a.G;
  const long long b = blockIdx.y;
  const int chunk = blockIdx.x;
  const long long col_begin = static_cast<long long>(chunk) * kChunkCols;
  const long long col_end = min(a.K, col_begin + kChunkCols);
  const T* s = static_cast<const T*>(a.s);
  const T* t = static_cast<const T*>(a.t);
  T* ds = static_cast<T*>(a.ds);
  const float scale = a.coef;
  const float c2 = a.inv_ts * kLog2e, ct = a.inv_tt * kLog2e;
  float tmc[MAXG], tinv[MAXG], lse2[MAXC];
#pragma unroll
  for (int i = 0; i < MAXG; ++i) {
    tmc[i] = 0.f; tinv[i] = 0.f;
    if (i < G) { const float2 st = a.t_stats[i * a.B + b]; tmc[i] = -st.x; tinv[i] = st.y; }
  }
#pragma unroll
  for (int v = 0; v < MAXC; ++v) {
    lse2[v] = 0.f;
    if (v < C) lse2[v] = -a.s_lse[v * a.B + b] * kLog2e;
  }
  float cross = 0.f;
  if constexpr (ALLFAST && sizeof(T) == 2) {
    // software pipeline: the packed loads of the NEXT vector (C + G rows, 8 bytes each) are in flight while this one is
    // computed -- twice the bytes in flight per thread at the same occupancy (the pass is bound by memory-level
    // parallelism: 16 warps per SM x 96 bytes per thread otherwise)
    CeVec<T, MAXC, MAXG> cur, nxt;
    long long col = col_begin + threadIdx.x * 4;
    if (col < col_end) ce_load<T, MAXC, MAXG, true>(a, s, t, b, col, C, G, cur);
    for (; col < col_end; col += kThreads * 4) {
      const long long cn = col + kThreads * 4;
      if (cn < col_end) ce_load<T, MAXC, MAXG, true>(a, s, t, b, cn, C, G, nxt);
      ce_fused_compute<T, MAXC, MAXG, true>(a, cur, ds, b, col, C, G, tmc, tinv, lse2, c2, ct, scale, cross);
      cur = nxt;
    }
  } else {
    for (long long col = col_begin + threadIdx.x * 4; col < col_end; col += kThreads * 4) {
      if constexpr (ALLFAST) {
        ce_fused_vector<T, MAXC, MAXG, true>(a, s, t, ds, b, col, C, G, tmc, tinv, lse2, c2, ct, scale, cross);
      } else {
        if (a.vec_ok && (col + 4 <= a.K)) ce_fused_vector<T, MAXC, MAXG, true>(a, s, t, ds, b, col, C, G, tmc, tinv, lse2, c2, ct, scale, cross);
        else ce_fused_vector<T, MAXC, MAXG, false>(a, s, t, ds, b, col, C, G, tmc, tinv, lse2, c2, ct, scale, cross);
      }
    }
  }
  __shared__ float red[kThreads / 32];
  cross = warp_sum(cross);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cross;
  __syncthreads();
  if (threadIdx.x == 0) {
    float c = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) c += red[w];
    a.ws_x[b * a.nchunks + chunk] = c;
  }
}

// loss = ( sum_rows n_v lse - sum cross partials ) / (n B), from final per-row lse values (fixed order).
__global__ void __launch_bounds__(1024)
ce_finalize_lse_kernel(const float* __restrict__ s_lse, const float* __restrict__ ws_x, long long B, int C, int G, int nchunks,
                       float* __restrict__ loss) {
  pdl_prologue();
  __shared__ double red[32];
  double acc = 0.0;
  const long long rows = static_cast<long long>(C) * B;
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
    const int v = static_cast<int>(r / B);
    acc += static_cast<double>((v < G) ? (G - 1) : G) * static_cast<double>(s_lse[r]);
  }
  for (long long i = threadIdx.x; i < B * nchunks; i += blockDim.x) acc -= static_cast<double>(ws_x[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    const int n_terms = G * C - (G < C ? G : C);
    loss[0] = static_cast<float>(tot / (static_cast<double>(n_terms) * static_cast<double>(B)));
  }
}

// lse[m] = (max2 + log2(sum)) * ln2 from the GEMM epilogue's per-part (max2, sum) pairs.  One warp per row.
__global__ void __launch_bounds__(256)
lse_finalize_kernel(const float2* __restrict__ partials, long long M, int parts, float* __restrict__ lse) {
  pdl_prologue();
  const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= M) return;
  const int lane = threadIdx.x & 31;
  float m = -INFINITY, l = 0.f;
  for (int c = lane; c < parts; c += 32) { const float2 p = partials[r * parts + c]; online_merge2(m, l, p.x, p.y); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float l2 = __shfl_xor_sync(0xffffffffu, l, o);
    online_merge2(m, l, m2, l2);
  }
  if (lane == 0) lse[r] = (m + log2f(l)) * 0.6931471805599453f;
}

// x *= (*scale / expected) unless *scale == expected (then every CTA exits after one 4-byte read).
template <typename T>
__global__ void __launch_bounds__(256)
scale_if_kernel(T* __restrict__ x, long long n, const float* __restrict__ scale, float expected) {
  pdl_prologue();
  const float sc = __ldg(scale);
  if (sc == expected) return;
  const float f = sc / expected;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    Vec<T>::store1(x + i, Vec<T>::load1(x + i) * f);
}

template <typename T>
int launch_fused(const CeArgs& a, dim3 grid, cudaStream_t st) {
  const bool allfast = a.vec_ok && (a.K % 4 == 0);
  if (a.C == 8 && a.G == 2 && allfast) launch_kernel(ce_fused_kernel<T, 8, 2, true>, dim3(grid), dim3(kThreads), 0, st, a);
  else if (a.C == 8 && a.G == 2) launch_kernel(ce_fused_kernel<T, 8, 2, false>, dim3(grid), dim3(kThreads), 0, st, a);
  else if (a.C == 9 && a.G == 3 && allfast) launch_kernel(ce_fused_kernel<T, 9, 3, true>, dim3(grid), dim3(kThreads), 0, st, a);
  else if (a.C == 9 && a.G == 3) launch_kernel(ce_fused_kernel<T, 9, 3, false>, dim3(grid), dim3(kThreads), 0, st, a);
  else launch_kernel(ce_fused_kernel<T, 0, 0, false>, dim3(grid), dim3(kThreads), 0, st, a);
  DMC_LAUNCH_CHECK("ce_fused_kernel launch");
  return 0;
}

template <typename T>
int launch_fwd(const CeArgs& a, dim3 grid, cudaStream_t st) {
  if (a.C == 8 && a.G == 2) launch_kernel(ce_fwd_kernel<T, 8, 2>, dim3(grid), dim3(kThreads), 0, st, a);
  else if (a.C == 9 && a.G == 3) launch_kernel(ce_fwd_kernel<T, 9, 3>, dim3(grid), dim3(kThreads), 0, st, a);
  else launch_kernel(ce_fwd_kernel<T, 0, 0>, dim3(grid), dim3(kThreads), 0, st, a);
  DMC_LAUNCH_CHECK("ce_fwd_kernel launch");
  return 0;
}
template <typename T>
int launch_bwd(const CeArgs& a, dim3 grid, cudaStream_t st) {
  if (a.C == 8 && a.G == 2) launch_kernel(ce_bwd_kernel<T, 8, 2>, dim3(grid), dim3(kThreads), 0, st, a);
  else if (a.C == 9 && a.G == 3) launch_kernel(ce_bwd_kernel<T, 9, 3>, dim3(grid), dim3(kThreads), 0, st, a);
  else launch_kernel(ce_bwd_kernel<T, 0, 0>, dim3(grid), dim3(kThreads), 0, st, a);
  DMC_LAUNCH_CHECK("ce_bwd_kernel launch");
  return 0;
}

int chunk_cols(int /*dtype*/) { return kChunkCols; }

int check_common(const char* who, const void* s, int s_dtype, int64_t lds, const void* t, int t_dtype, int64_t ldt,
                 const float* center, const float* t_row_stats, int64_t B, int C, int G, int64_t K) {
  DMC_REQUIRE(s && t && center && t_row_stats, "%s: null pointer", who);
  DMC_REQUIRE(s_dtype == t_dtype && (s_dtype == DMC_F32 || s_dtype == DMC_BF16), "%s: student/teacher logits must share one dtype (F32 or BF16)", who);
  DMC_REQUIRE(B > 0 && K > 0 && lds >= K && ldt >= K, "%s: bad shape B=%lld K=%lld", who, (long long)B, (long long)K);
  DMC_REQUIRE(C >= 1 && C <= 16 && G >= 1 && G <= 4, "%s: ncrops must be in [1,16] and teacher crops in [1,4] (got %d, %d)", who, C, G);
  DMC_REQUIRE(G * C - (G < C ? G : C) > 0, "%s: no (teacher, student) pair left: ncrops=%d teacher_crops=%d", who, C, G);
  DMC_REQUIRE(B <= 65535, "%s: batch per GPU too large for the launch grid (%lld)", who, (long long)B);
  return 0;
}

}  // namespace
}  // namespace dmc

using namespace dmc;

extern "C" size_t dmc_ce_workspace_bytes(int64_t B, int32_t C, int32_t G, int64_t K) {
  if (B <= 0 || C <= 0 || K <= 0) return 0;
  (void)G;
  const int64_t nchunks = ceil_div(K, chunk_cols(DMC_F32));   // fp32 has the most chunks
  const size_t a = (static_cast<size_t>(C) * B * nchunks * sizeof(float2) + 255) & ~static_cast<size_t>(255);
  return a + static_cast<size_t>(B) * nchunks * sizeof(float);
}

extern "C" int dmc_ce_fwd(const void* s, int32_t s_dtype, int64_t lds, const void* t, int32_t t_dtype, int64_t ldt,
                          const float* center, const float* t_row_stats, int64_t B, int32_t C, int32_t G, int64_t K,
                          float inv_student_temp, float inv_teacher_temp, float* s_lse, float* loss, void* workspace,
                          size_t workspace_bytes, void* stream) {
  int rc = check_common("dmc_ce_fwd", s, s_dtype, lds, t, t_dtype, ldt, center, t_row_stats, B, C, G, K);
  if (rc) return rc;
  DMC_REQUIRE(s_lse && loss && workspace, "dmc_ce_fwd: null output");
  DMC_REQUIRE(workspace_bytes >= dmc_ce_workspace_bytes(B, C, G, K), "dmc_ce_fwd: workspace too small");
  DMC_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "dmc_ce_fwd: workspace must be 16-byte aligned");
  const int esz = s_dtype == DMC_BF16 ? 2 : 4;
  CeArgs a{};
  a.s = s; a.lds = lds; a.t = t; a.ldt = ldt; a.center = center; a.t_stats = reinterpret_cast<const float2*>(t_row_stats);
  a.B = B; a.K = K; a.C = C; a.G = G; a.inv_ts = inv_student_temp; a.inv_tt = inv_teacher_temp;
  a.nchunks = static_cast<int>(ceil_div(K, chunk_cols(s_dtype)));
  const int va = 4 * esz;     // bytes per packed 4-element access
  a.vec_ok = ((reinterpret_cast<uintptr_t>(s) % va) == 0) && ((reinterpret_cast<uintptr_t>(t) % va) == 0) &&
             ((lds * esz) % va == 0) && ((ldt * esz) % va == 0);
  const size_t s_bytes = (static_cast<size_t>(C) * B * a.nchunks * sizeof(float2) + 255) & ~static_cast<size_t>(255);
  a.ws_s = static_cast<float2*>(workspace);
  a.ws_x = reinterpret_cast<float*>(static_cast<char*>(workspace) + s_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)a.nchunks, (unsigned)B);
  rc = (s_dtype == DMC_BF16) ? launch_fwd<__nv_bfloat16>(a, grid, st) : launch_fwd<float>(a, grid, st);
  if (rc) return rc;
  launch_kernel(ce_finalize_kernel, dim3(1), dim3(1024), 0, st, a.ws_s, a.ws_x, B, C, G, a.nchunks, s_lse, loss);
  DMC_LAUNCH_CHECK("ce_finalize_kernel launch");
  return 0;
}

extern "C" int dmc_ce_bwd(const void* s, int32_t s_dtype, int64_t lds, const void* t, int32_t t_dtype, int64_t ldt,
                          const float* center, const float* t_row_stats, const float* s_lse, const float* grad_out, int64_t B,
                          int32_t C, int32_t G, int64_t K, float inv_student_temp, float inv_teacher_temp, void* ds,
                          int32_t ds_dtype, int64_t ldds, void* stream) {
  int rc = check_common("dmc_ce_bwd", s, s_dtype, lds, t, t_dtype, ldt, center, t_row_stats, B, C, G, K);
  if (rc) return rc;
  DMC_REQUIRE(s_lse && grad_out && ds, "dmc_ce_bwd: null pointer");
  DMC_REQUIRE(ds_dtype == s_dtype && ldds >= K, "dmc_ce_bwd: gradient must have the logits' dtype and ld >= K");
  const int esz = s_dtype == DMC_BF16 ? 2 : 4;
  CeArgs a{};
  a.s = s; a.lds = lds; a.t = t; a.ldt = ldt; a.center = center; a.t_stats = reinterpret_cast<const float2*>(t_row_stats);
  a.B = B; a.K = K; a.C = C; a.G = G; a.inv_ts = inv_student_temp; a.inv_tt = inv_teacher_temp;
  a.nchunks = static_cast<int>(ceil_div(K, chunk_cols(s_dtype)));
  const int va = 4 * esz;
  a.vec_ok = ((reinterpret_cast<uintptr_t>(s) % va) == 0) && ((reinterpret_cast<uintptr_t>(t) % va) == 0) &&
             ((reinterpret_cast<uintptr_t>(ds) % va) == 0) && ((lds * esz) % va == 0) && ((ldt * esz) % va == 0) &&
             ((ldds * esz) % va == 0);
  a.s_lse = s_lse; a.gout = grad_out;
  const int n_terms = G * C - (G < C ? G : C);
  a.coef = static_cast<float>(static_cast<double>(inv_student_temp) / (static_cast<double>(n_terms) * static_cast<double>(B)));
  a.ds = ds; a.ldds = ldds;
  dim3 grid((unsigned)a.nchunks, (unsigned)B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return (s_dtype == DMC_BF16) ? launch_bwd<__nv_bfloat16>(a, grid, st) : launch_bwd<float>(a, grid, st);
}

extern "C" int dmc_lse_finalize(const float* row_partials, int64_t M, int64_t parts, float* lse, void* stream) {
  DMC_REQUIRE(row_partials && lse && M > 0 && parts > 0 && parts < (1 << 30), "dmc_lse_finalize: bad arguments");
  launch_kernel(lse_finalize_kernel, dim3((unsigned)ceil_div(M, 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), reinterpret_cast<const float2*>(row_partials), M, (int)parts, lse);
  DMC_LAUNCH_CHECK("lse_finalize_kernel launch");
  return 0;
}

extern "C" int dmc_ce_fused(const void* s, int32_t s_dtype, int64_t lds, const void* t, int32_t t_dtype, int64_t ldt,
                            const float* center, const float* t_row_stats, const float* s_lse, int64_t B, int32_t C, int32_t G,
                            int64_t K, float inv_student_temp, float inv_teacher_temp, void* ds, int32_t ds_dtype, int64_t ldds,
                            float* loss, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common("dmc_ce_fused", s, s_dtype, lds, t, t_dtype, ldt, center, t_row_stats, B, C, G, K);
  if (rc) return rc;
  DMC_REQUIRE(s_lse && ds && loss && workspace, "dmc_ce_fused: null pointer");
  DMC_REQUIRE(ds_dtype == s_dtype && ldds >= K, "dmc_ce_fused: gradient must have the logits' dtype and ld >= K");
  DMC_REQUIRE(workspace_bytes >= dmc_ce_workspace_bytes(B, C, G, K), "dmc_ce_fused: workspace too small");
  const int esz = s_dtype == DMC_BF16 ? 2 : 4;
  CeArgs a{};
  a.s = s; a.lds = lds; a.t = t; a.ldt = ldt; a.center = center; a.t_stats = reinterpret_cast<const float2*>(t_row_stats);
  a.B = B; a.K = K; a.C = C; a.G = G; a.inv_ts = inv_student_temp; a.inv_tt = inv_teacher_temp;
  a.nchunks = static_cast<int>(ceil_div(K, chunk_cols(s_dtype)));
  const int va = 4 * esz;
  a.vec_ok = ((reinterpret_cast<uintptr_t>(s) % va) == 0) && ((reinterpret_cast<uintptr_t>(t) % va) == 0) &&
             ((reinterpret_cast<uintptr_t>(ds) % va) == 0) && ((lds * esz) % va == 0) && ((ldt * esz) % va == 0) &&
             ((ldds * esz) % va == 0);
  a.s_lse = s_lse;
  const int n_terms = G * C - (G < C ? G : C);
  a.coef = static_cast<float>(static_cast<double>(inv_student_temp) / (static_cast<double>(n_terms) * static_cast<double>(B)));
  a.ds = ds; a.ldds = ldds;
  a.ws_x = static_cast<float*>(workspace);
  dim3 grid((unsigned)a.nchunks, (unsigned)B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rc = (s_dtype == DMC_BF16) ? launch_fused<__nv_bfloat16>(a, grid, st) : launch_fused<float>(a, grid, st);
  if (rc) return rc;
  launch_kernel(ce_finalize_lse_kernel, dim3(1), dim3(1024), 0, st, s_lse, a.ws_x, B, C, G, a.nchunks, loss);
  DMC_LAUNCH_CHECK("ce_finalize_lse_kernel launch");
  return 0;
}

extern "C" int dmc_scale_inplace_if(void* x, int32_t dtype, int64_t n, const float* scale, float expected, void* stream) {
  DMC_REQUIRE(x && scale && n > 0 && expected != 0.f, "dmc_scale_inplace_if: bad arguments");
  DMC_REQUIRE(dtype == DMC_F32 || dtype == DMC_BF16, "dmc_scale_inplace_if: bad dtype");
  const int blocks = kNumSMs * 8;
  if (dtype == DMC_BF16)
    launch_kernel(scale_if_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<__nv_bfloat16*>(x), n, scale, expected);
  else
    launch_kernel(scale_if_kernel<float>, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<float*>(x), n, scale, expected);
  DMC_LAUNCH_CHECK("scale_if_kernel launch");
  return 0;
}
