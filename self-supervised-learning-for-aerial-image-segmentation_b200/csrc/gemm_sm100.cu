// gemm_sm100.cu -- the tensor-core path of libdinomc: a persistent, warp-specialised
// tcgen05 / TMEM / TMA GEMM for sm_100a.
//
//   D[M,N] = epilogue( sum_k A(m,k) B(n,k) )            (contract in include/dinomc.h: dmc_gemm)
//
// It serves the weight-normed last layer of DINOHead (utils/vision_transformer.py:293: M = crops*batch,
// N = out_dim = 65536, K = bottleneck = 256), its dgrad (contraction over out_dim, split-K) and wgrad
// (MN-major operands straight from the row-major gradient), and the MLP Linears (:291).
//
// Structure (one CTA per SM, 320 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor tiles into a ring of 128B-swizzled smem stages
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma (M=128, N=block_n, K=32 bytes/row)
//               into one of two TMEM accumulators (2 x 256 fp32 columns = all 512 TMEM columns)
//   warps 2..9  epilogue: tcgen05.ld the finished accumulator (two warps per TMEM lane quadrant, one per half of
//               the tile's columns), apply scale/bias/activation, store; overlaps the MMAs of the next tile
//               thanks to the double-buffered accumulator.
// Three mbarrier pipelines: smem full/empty (TMA<->MMA), TMEM full/empty (MMA<->epilogue).
//
// Output path: the epilogue warps write the converted tile into 128B-swizzled smem staging buffers and one
// lane issues a TMA store (cp.async.bulk.tensor, coalesced 128-byte rows, clipped at the matrix edge).
// "B-resident" schedule (short contractions, e.g. the last layer's K = 256): every CTA owns a contiguous range
// of tiles ordered n-tile-major, keeps the whole B (weight) tile of its current n-tile in smem and streams only
// the A tiles, which cuts the L2->smem operand traffic from 3x to ~1x the output bytes.
//
// Operands may be K-major or MN-major (tcgen05 reads both through the smem matrix descriptor), in
// bf16 (kind::f16) or fp32-as-TF32 (kind::tf32) with an optional 3-pass hi/lo split ("3xTF32").
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>

#include <cstdio>
#include <cstdlib>
#include "dmc_common.cuh"
#include "dmc_ptx.cuh"

namespace dmc {
namespace {

constexpr int kBlockM = 128;
constexpr int kRowBytes = 128;          // one swizzle row: 64 bf16 or 32 tf32 along the contiguous dim
constexpr int kMaxStages = 8;
constexpr int kMmaPerKBlock = 4;        // 128 B / 32 B: four tcgen05.mma per k-block
constexpr int kTmemCols = 512;
constexpr int kAccCols = 256;
constexpr int kEpiWarps = 8;           // two epilogue warps per TMEM lane quadrant (one per half of the tile columns);
                                       // 16 warps (register cap 96, 3 smem stages) measured slower
constexpr int kColGroups = kEpiWarps / 4;   // column groups per tile = statistics parts per tile
constexpr int kThreads = 64 + kEpiWarps * 32;
constexpr uint32_t kABytes = kBlockM * kRowBytes;  // 16 KiB per stage for A

struct GemmDev {
  int M, N;
  int block_n;
  int m_tiles, n_tiles;
  int kb_total;       // k-blocks per pass
  int vk_total;       // passes * kb_total "virtual" k-blocks
  int splits, vk_per_split;
  int stages;
  uint32_t b_bytes;   // block_n * 128
  int resident;       // 1: B-resident schedule (contiguous tile ranges, B slabs loaded once per n-tile)
  int dual;           // 1: "dual-M": a work item is 256 rows = two accumulators that share every B k-block
  int banks;          // > 1 (3xTF32 only): the contraction is spread over `banks` TMEM accumulators of block_n columns that the
                      // epilogue adds in fp32 registers.  tcgen05 accumulation TRUNCATES (measured: results shrink towards zero
                      // by ~4e-8 per MMA of a chain), so fp32-grade results need short chains: hi*hi k-blocks rotate over
                      // banks 0..banks-2, the small hi*lo / lo*hi terms share the last bank
  // fused output statistics (EPI == 2): softmax row partials and 32-row column sums of the stored values
  float stat_sc2;               // stat_scale * log2(e)
  const float* stat_center;     // [N] or nullptr
  float2* stat_row_partials;    // [M][kColGroups * n_tiles]
  float* stat_colsum_partials;  // [ceil(M/32)][N] or nullptr
  const float* stat_bound;      // device scalar b with |D| <= b, or nullptr
  const float* stat_bound2;     // optional second device scalar added to it (e.g. max |center|)
  int dbg;                      // DMC_GEMM_FLAGS >> 3 (timing experiments only): 1 = no TMA store issue, 2 = no staging writes, 4 = no TMEM loads
  int tma_store;      // 1: epilogue stores D through smem staging + TMA
  long long* trace;   // DMC_GEMM_TRACE=1 (debug): clock64() stamps of CTA 0's pipeline events, [8][512]
  // epilogue
  void* D; long long ldd; int out_dtype;
  float* partial;     // split-K partial sums [splits][M][N] (raw accumulators) or nullptr
  const float* col_scale; const float* bias; const float* alpha_dev; float alpha;
  int act; void* aux; long long ldaux; int aux_dtype;
};

struct Epilogue {
  const float* col_scale; const float* bias; float alpha; int act;
  void* aux; long long ldaux; int aux_dtype;
  void* D; long long ldd; int out_dtype;
  int N;
};

__device__ __forceinline__ float load_elem(const void* p, long long idx, int dtype) {
  return dtype == DMC_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx])
                           : reinterpret_cast<const float*>(p)[idx];
}
__device__ __forceinline__ void store_elem(void* p, long long idx, int dtype, float v) {
  if (dtype == DMC_BF16) reinterpret_cast<__nv_bfloat16*>(p)[idx] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[idx] = v;
}

// Epilogue arithmetic on the 32 consecutive columns [col0, col0+32) of one row (raw fp32 accumulators in,
// finished values out).  Every option sits behind ONE warp-uniform branch so the common plain case costs
// nothing; columns >= N are computed on clamped vector entries and never stored.  `n` = valid columns.
__device__ __forceinline__ void epilogue_math(const Epilogue& e, float (&acc)[32], long long row, int col0, int n,
                                              bool aux_vec_ok) {
  if (e.col_scale != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] *= __ldg(e.col_scale + min(col0 + j, e.N - 1));
  }
  if (e.alpha != 1.0f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] *= e.alpha;
  }
  if (e.bias != nullptr) {
    if (n == 32 && (reinterpret_cast<uintptr_t>(e.bias) & 15) == 0) {     // col0 is a multiple of 32: 8 aligned 128-bit loads
      const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = __ldg(b4 + j);
        acc[4 * j] += b.x; acc[4 * j + 1] += b.y; acc[4 * j + 2] += b.z; acc[4 * j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] += __ldg(e.bias + min(col0 + j, e.N - 1));
    }
  }
  if (e.act == DMC_ACT_GELU) {
    if (e.aux != nullptr) {                                  // save the pre-activation for backward
      if (aux_vec_ok && n == 32) {
        if (e.aux_dtype == DMC_BF16) {
          uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.aux) + row * e.ldaux + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            p[j] = make_uint4(pack_bf16(acc[8 * j], acc[8 * j + 1]), pack_bf16(acc[8 * j + 2], acc[8 * j + 3]),
                              pack_bf16(acc[8 * j + 4], acc[8 * j + 5]), pack_bf16(acc[8 * j + 6], acc[8 * j + 7]));
        } else {
          float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.aux) + row * e.ldaux + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) p[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < n) store_elem(e.aux, row * e.ldaux + col0 + j, e.aux_dtype, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = gelu_f(acc[j]);
  } else if (e.act == DMC_ACT_GELU_DG) {
    // D = gelu(z) and aux = gelu'(z) from ONE evaluation (they share exp(-z^2/2)): the backward epilogue is then a plain
    // multiply (DMC_ACT_MUL_AUX) instead of a second, exposed GELU' evaluation per element
    float dg[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float phi, dens;
      gelu_parts(acc[j], phi, dens);
      dg[j] = fmaf(acc[j] * dens, 0.39894228040143268f, phi);
      acc[j] *= phi;
    }
    if (aux_vec_ok && n == 32) {
      if (e.aux_dtype == DMC_BF16) {
        uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.aux) + row * e.ldaux + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          p[j] = make_uint4(pack_bf16(dg[8 * j], dg[8 * j + 1]), pack_bf16(dg[8 * j + 2], dg[8 * j + 3]),
                            pack_bf16(dg[8 * j + 4], dg[8 * j + 5]), pack_bf16(dg[8 * j + 6], dg[8 * j + 7]));
      } else {
        float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.aux) + row * e.ldaux + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) p[j] = make_float4(dg[4 * j], dg[4 * j + 1], dg[4 * j + 2], dg[4 * j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < n) store_elem(e.aux, row * e.ldaux + col0 + j, e.aux_dtype, dg[j]);
    }
  } else if (e.act == DMC_ACT_MUL_AUX) {
    if (aux_vec_ok && n == 32 && e.aux_dtype == DMC_BF16) {
      const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(e.aux) + row * e.ldaux + col0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 w = p[j];
        acc[8 * j + 0] *= bf16_lo(w.x); acc[8 * j + 1] *= bf16_hi(w.x); acc[8 * j + 2] *= bf16_lo(w.y); acc[8 * j + 3] *= bf16_hi(w.y);
        acc[8 * j + 4] *= bf16_lo(w.z); acc[8 * j + 5] *= bf16_hi(w.z); acc[8 * j + 6] *= bf16_lo(w.w); acc[8 * j + 7] *= bf16_hi(w.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < n) acc[j] *= load_elem(e.aux, row * e.ldaux + col0 + j, e.aux_dtype);
    }
  } else if (e.act == DMC_ACT_GELU_BWD) {
    if (aux_vec_ok && n == 32 && e.aux_dtype == DMC_BF16) {
      const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(e.aux) + row * e.ldaux + col0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 w = p[j];
        acc[8 * j + 0] *= gelu_grad_f(bf16_lo(w.x)); acc[8 * j + 1] *= gelu_grad_f(bf16_hi(w.x));
        acc[8 * j + 2] *= gelu_grad_f(bf16_lo(w.y)); acc[8 * j + 3] *= gelu_grad_f(bf16_hi(w.y));
        acc[8 * j + 4] *= gelu_grad_f(bf16_lo(w.z)); acc[8 * j + 5] *= gelu_grad_f(bf16_hi(w.z));
        acc[8 * j + 6] *= gelu_grad_f(bf16_lo(w.w)); acc[8 * j + 7] *= gelu_grad_f(bf16_hi(w.w));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < n) acc[j] *= gelu_grad_f(load_elem(e.aux, row * e.ldaux + col0 + j, e.aux_dtype));
    }
  }
}

// Direct (non-TMA) store of one row's 32-column chunk; used for unaligned outputs only.
__device__ __forceinline__ void epilogue_store_row(const Epilogue& e, float (&acc)[32], long long row, int col0, int n,
                                                   bool vec_ok) {
  if (vec_ok && n == 32) {
    if (e.out_dtype == DMC_BF16) {
      uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.D) + row * e.ldd + col0);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        p[j] = make_uint4(pack_bf16(acc[8 * j], acc[8 * j + 1]), pack_bf16(acc[8 * j + 2], acc[8 * j + 3]),
                          pack_bf16(acc[8 * j + 4], acc[8 * j + 5]), pack_bf16(acc[8 * j + 6], acc[8 * j + 7]));
    } else {
      float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.D) + row * e.ldd + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) p[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < n) store_elem(e.D, row * e.ldd + col0 + j, e.out_dtype, acc[j]);
  }
}

// Pipeline trace (clock64 stamps of CTA 0): compiled in only with -DDMC_GEMM_TRACE_BUILD (tools/gemm_trace.py); the
// production kernels carry none of it.
__device__ __forceinline__ void trace_at(const GemmDev& p, int kind, int idx) {
#ifdef DMC_GEMM_TRACE_BUILD
  if (p.trace != nullptr && blockIdx.x == 0 && idx < 512 && (threadIdx.x & 31) == 0) p.trace[kind * 512 + idx] = clock64();
#endif
}

constexpr int kStagingBytesPerWarp = 4096;       // one 32-row x 128-byte box per epilogue warp
constexpr int kStagingBytes = kEpiWarps * kStagingBytesPerWarp;

// ---- lean epilogue of one finished accumulator -----------------------------------------------------------------
// The common case: output through smem staging + TMA store, tile entirely inside N (rows >= M are clipped by the TMA
// store), no split-K partials.  The generic chunk code below handles everything else; it spends ~9 of 10 instructions on
// its runtime options, which is what bounded the epilogue (ncu: 745 warp-instructions per tile and warp for 84 useful).
// A warp drains columns [c_begin, c_end) (a multiple of 64) of its 32 TMEM lanes, 64 columns per step; the TMEM loads
// of step i+1 are issued before the statistics of step i so their latency hides behind the SFU work.
template <int EPI, bool OUT_BF16, typename Release>
__device__ __forceinline__ void epilogue_fast_acc(const GemmDev& p, const CUtensorMap* tmD, const Epilogue& e,
                                                  uint32_t t_addr, int row0, int lane, int n0, int c_begin, int c_end,
                                                  uint8_t* buf, uint32_t& n_boxes, bool release_after, Release release,
                                                  float& st_m, float& st_l) {
  const float stat_shift = st_m;                       // EPI 2: the fixed shift (bound); EPI 3 updates st_m as it goes
  const long long row = static_cast<long long>(row0) + lane;
  const bool row_ok = row < p.M;
  uint8_t* rowp = buf + lane * 128;
  const uint32_t sw = static_cast<uint32_t>(lane & 7);
  uint32_t ra[32], rb[32];
  ptx::tmem_ld_32x32(t_addr + c_begin, ra);
  ptx::tmem_ld_32x32(t_addr + c_begin + 32, rb);
  for (int c = c_begin; c < c_end; c += 64) {
    const bool last = (c + 64 >= c_end);
    ptx::tmem_ld_wait();
    if (last && release_after) release();                     // accumulator fully read: hand it back to the MMA warp
    float (&va)[32] = reinterpret_cast<float (&)[32]>(ra);
    float (&vb)[32] = reinterpret_cast<float (&)[32]>(rb);
    if constexpr (EPI == 0) {
      if (row_ok) {
        epilogue_math(e, va, row, n0 + c, 32, true);
        epilogue_math(e, vb, row, n0 + c + 32, 32, true);
      }
    } else {
      if (e.alpha != 1.0f) {
#pragma unroll
        for (int j = 0; j < 32; ++j) { va[j] *= e.alpha; vb[j] *= e.alpha; }
      }
    }
    if constexpr (OUT_BF16) {
      // one box: 32 rows x 64 bf16 columns; 16-byte chunk j of row r at r*128 + ((j ^ (r & 7)) << 4)
      uint32_t pk[32];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        pk[j] = pack_bf16(va[2 * j], va[2 * j + 1]);
        pk[16 + j] = pack_bf16(vb[2 * j], vb[2 * j + 1]);
      }
      if (n_boxes != 0) {                                     // the previous TMA store must have read the buffer
        if (lane == 0) ptx::tma_store_wait_read<0>();
        __syncwarp();
      }
#pragma unroll
      for (uint32_t j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      ptx::fence_proxy_async();                               // generic-proxy smem writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_2d(tmD, buf, n0 + c, row0);
        ptx::tma_store_commit();
      }
      ++n_boxes;
      if (!last) {
        ptx::tmem_ld_32x32(t_addr + c + 64, ra);
        ptx::tmem_ld_32x32(t_addr + c + 96, rb);
      }
      if constexpr (EPI == 3) {
        // Teacher statistics of the stored values: y2 = (t - center) * scale * log2e; online softmax pair (max, sum 2^(y2-max))
        // of this row over the step's 64 columns, and the column sums of the warp's 32 rows read back from the staged box
        // (lane j owns the 32-bit word j of every 128-byte row: conflict-free, 2 bf16 columns per lane).
        if (p.stat_colsum_partials != nullptr && row0 < p.M) {          // optional: the caller may get the column sums elsewhere
          const int nrows = min(32, p.M - row0);
          float s0 = 0.f, s1 = 0.f;
          for (int r = 0; r < nrows; ++r) {
            const uint32_t wv = *reinterpret_cast<const uint32_t*>(buf + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
            s0 += bf16_lo(wv); s1 += bf16_hi(wv);
          }
          float* dst = p.stat_colsum_partials + static_cast<long long>(row0 >> 5) * p.N + n0 + c;
          *reinterpret_cast<float2*>(dst + 2 * lane) = make_float2(s0, s1);
        }
        // pk[j] (j < 16) holds columns c+2j, c+2j+1; pk[16+j] holds c+32+2j, c+33+2j
        const float4* cen = reinterpret_cast<const float4*>(p.stat_center + n0 + c);
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int q = 0; q < 16; ++q) {                          // pass 1: max of (t - center) over the 64 columns
          const float4 cc = __ldg(cen + q);
          const uint32_t w0 = pk[2 * q], w1 = pk[2 * q + 1];
          m0 = fmaxf(m0, fmaxf(bf16_lo(w0) - cc.x, bf16_hi(w0) - cc.y));
          m1 = fmaxf(m1, fmaxf(bf16_lo(w1) - cc.z, bf16_hi(w1) - cc.w));
        }
        const float mnew = fmaxf(st_m, fmaxf(m0, m1) * p.stat_sc2);
        st_l *= ex2(st_m - mnew);                               // 0 * 2^(-inf) = 0 on the first step
        st_m = mnew;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) {                          // pass 2: sum 2^(y2 - max)
          const float4 cc = __ldg(cen + q);
          const uint32_t w0 = pk[2 * q], w1 = pk[2 * q + 1];
          a0 += ex2(fmaf(bf16_lo(w0) - cc.x, p.stat_sc2, -mnew));
          a1 += ex2(fmaf(bf16_hi(w0) - cc.y, p.stat_sc2, -mnew));
          a2 += ex2(fmaf(bf16_lo(w1) - cc.z, p.stat_sc2, -mnew));
          a3 += ex2(fmaf(bf16_hi(w1) - cc.w, p.stat_sc2, -mnew));
        }
        st_l += (a0 + a1) + (a2 + a3);
      }
      if constexpr (EPI == 2) {                               // statistics of exactly the stored (rounded) values
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (p.stat_center != nullptr) {
          // teacher form, fixed shift: sum 2^((t - center) * sc2 - shift).  pk[2q], pk[2q+1] hold columns c+4q .. c+4q+3 (q < 8)
          // and c+32+4(q-8) .. (q >= 8); the center values are the same for every lane (row) of the warp: broadcast loads
          const float4* cen = reinterpret_cast<const float4*>(p.stat_center + n0 + c);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float4 cc = __ldg(cen + q);
            const uint32_t w0 = pk[2 * q], w1 = pk[2 * q + 1];
            a0 += ex2(fmaf(bf16_lo(w0) - cc.x, p.stat_sc2, -stat_shift));
            a1 += ex2(fmaf(bf16_hi(w0) - cc.y, p.stat_sc2, -stat_shift));
            a2 += ex2(fmaf(bf16_lo(w1) - cc.z, p.stat_sc2, -stat_shift));
            a3 += ex2(fmaf(bf16_hi(w1) - cc.w, p.stat_sc2, -stat_shift));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            a0 += ex2(fmaf(bf16_lo(pk[j]), p.stat_sc2, -stat_shift));
            a1 += ex2(fmaf(bf16_hi(pk[j]), p.stat_sc2, -stat_shift));
            a2 += ex2(fmaf(bf16_lo(pk[j + 1]), p.stat_sc2, -stat_shift));
            a3 += ex2(fmaf(bf16_hi(pk[j + 1]), p.stat_sc2, -stat_shift));
          }
        }
        st_l += (a0 + a1) + (a2 + a3);
      }
    } else {
      // two boxes of 32 rows x 32 fp32 columns through the same staging buffer
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float (&v)[32] = half ? vb : va;
        if (n_boxes != 0) {
          if (lane == 0) ptx::tma_store_wait_read<0>();
          __syncwarp();
        }
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(rowp + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_store_2d(tmD, buf, n0 + c + 32 * half, row0);
          ptx::tma_store_commit();
        }
        ++n_boxes;
      }
      if constexpr (EPI == 2) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          a0 += ex2(fmaf(va[j], p.stat_sc2, -stat_shift));
          a1 += ex2(fmaf(va[j + 1], p.stat_sc2, -stat_shift));
          a2 += ex2(fmaf(vb[j], p.stat_sc2, -stat_shift));
          a3 += ex2(fmaf(vb[j + 1], p.stat_sc2, -stat_shift));
        }
        st_l += (a0 + a1) + (a2 + a3);
      }
      if (!last) {
        ptx::tmem_ld_32x32(t_addr + c + 64, ra);
        ptx::tmem_ld_32x32(t_addr + c + 96, rb);
      }
    }
  }
}

// EPI 0: full epilogue (column scale / bias / activation / aux).
// EPI 1: "plain" -- scale by alpha, convert, store; compiled separately so the hot last-layer kernels carry none of
//        the optional epilogue code.
// EPI 2: plain + fused statistics of the stored (rounded) logits: per row and 128-column part the online-softmax
//        pair (max, sum 2^(y - max)) of y = (D - center) * scale * log2e, and per 32-row group the column sums.
//        This is what lets DINOLoss skip its separate statistics passes over the logits.
// EPI 3: EPI 2 for the TEACHER logits (center subtracted, running maximum, 32-row column sums), lean path for bf16 output.
// CG2: CTA pair (cluster of 2, tcgen05 cta_group::2).  One work item is 256 rows x block_n columns: each CTA of the
//      pair owns 128 rows of A and of the accumulator and loads only HALF of the B tile; the leader CTA's single MMA
//      thread drives both tensor cores (M = 256) reading both B halves.  A third less shared-memory fill per unit of
//      MMA work than the single-CTA tile, which is what bounds the long-K GEMMs.
template <int ESZ, bool A_MN, bool B_MN, int EPI, bool CG2>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
               const __grid_constant__ CUtensorMap tmD, const GemmDev p) {
  constexpr int BLOCK_K = kRowBytes / ESZ;       // elements per k-block (64 bf16 / 32 tf32)
  constexpr int UMMA_K = 32 / ESZ;               // elements per tcgen05.mma (16 / 8)
  constexpr int BOX_MN = kRowBytes / ESZ;        // MN-major box width in elements (64 / 32)
  constexpr uint32_t kBoxBytes = BLOCK_K * kRowBytes;   // one MN-major box: BLOCK_K k-rows x 128 B
  constexpr bool kTf32 = (ESZ == 4);
  constexpr uint32_t kMnSbo = kTf32 ? 512 : 1024;        // MN-major stride between swizzle-pattern repeats along K
  constexpr uint32_t kMnLayout = kTf32 ? 1 : 2;          // SWIZZLE_128B_BASE32B for tf32, SWIZZLE_128B for bf16

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);   // 1024-B aligned (SWIZZLE_128B atoms)

  // smem carve-up: [resident B slabs][stage ring][epilogue staging][barriers]
  const uint32_t cta_rank = CG2 ? ptx::cluster_ctarank() : 0u;      // rank in the CTA pair (0 = MMA leader)
  const uint32_t b_cta_bytes = CG2 ? (p.b_bytes >> 1) : p.b_bytes;  // CTA pair: this CTA holds half of the B tile
  const uint32_t a_stage_bytes = p.dual ? 2u * kABytes : kABytes;   // dual-M: two 128-row A tiles per stage
  const uint32_t stage_bytes = p.resident ? a_stage_bytes : (a_stage_bytes + b_cta_bytes);
  const int tile_m = (p.dual || CG2) ? 2 * kBlockM : kBlockM;
  const uint32_t res_bytes = p.resident ? static_cast<uint32_t>(p.vk_total) * p.b_bytes : 0u;
  uint8_t* b_res = smem;
  uint8_t* tiles = smem + res_bytes;
  uint8_t* staging = tiles + static_cast<size_t>(p.stages) * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + (p.tma_store ? kStagingBytes : 0));
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* bfull_bar = empty_bar + kMaxStages;          // resident B slabs (<= kMaxStages of them)
  uint64_t* bempty_bar = bfull_bar + kMaxStages;
  uint64_t* tmem_full = bempty_bar + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // warp-uniform for the compiler too
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) trace_at(p, 7, 0);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA0); ptx::prefetch_tensormap(&tmB0);
    ptx::prefetch_tensormap(&tmA1); ptx::prefetch_tensormap(&tmB1);
    if (p.tma_store) ptx::prefetch_tensormap(&tmD);
    for (int i = 0; i < kMaxStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1);
      ptx::mbar_init(&bfull_bar[i], 1); ptx::mbar_init(&bempty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], CG2 ? 2 * kEpiWarps : kEpiWarps);   // pair: both CTAs' epilogues release the leader
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {               // this warp owns the TMEM allocation (alloc + dealloc)
    if constexpr (CG2) { ptx::tmem_alloc_pair(tmem_ptr, kTmemCols); ptx::tmem_relinquish_pair(); }
    else { ptx::tmem_alloc(tmem_ptr, kTmemCols); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CG2) ptx::cluster_sync();       // barrier inits and TMEM of both CTAs are in place before any remote signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  // Everything above (barrier init, TMEM allocation, tensor-map prefetch) touched no global memory: it may overlap the
  // tail of the previous kernel in the stream.  From here on operands are read and outputs written.
  pdl_prologue();
  if (threadIdx.x == 0) trace_at(p, 7, 1);

  // Work enumeration.  Work item w -> (mt = w % m_tiles, nt = (w / m_tiles) % n_tiles, split = w / (m_tiles n_tiles)).
  // round-robin: w = blockIdx.x, +gridDim.x, ...   resident: a contiguous range (same n-tile for consecutive items).
  const int num_work = p.m_tiles * p.n_tiles * p.splits;
  int w_begin, w_end, w_step;
  if constexpr (CG2) {
    w_begin = blockIdx.x >> 1; w_end = num_work; w_step = gridDim.x >> 1;    // both CTAs of a pair walk the same items
  } else if (p.resident) {
    w_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * num_work / gridDim.x);
    w_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * num_work / gridDim.x);
    w_step = 1;
  } else {
    w_begin = blockIdx.x; w_end = num_work; w_step = gridDim.x;
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The WHOLE warp walks the loop (warp-uniform control flow and addresses, so the TMA operands sit in uniform
    // registers) and one elected lane issues.  A `lane == 0` branch around the loop instead makes the compiler wrap
    // every UTMALDG in an ELECT / R2UR.BROADCAST retry loop: measured ~800 cycles per k-block in this thread alone.
    int stage = 0; uint32_t phase = 0;
    int prev_nt = -1; uint32_t b_gen = 0;
    int tr = 0;
    for (int w = w_begin; w < w_end; w += w_step) {
      const int mt = w % p.m_tiles;
      const int rest = w / p.m_tiles;
      const int nt = rest % p.n_tiles;
      const int sp = rest / p.n_tiles;
      const int m0 = mt * tile_m + (CG2 ? static_cast<int>(cta_rank) * kBlockM : 0);
      const int n0 = nt * p.block_n + (CG2 ? static_cast<int>(cta_rank) * (p.block_n >> 1) : 0);
      const int nb_rows = CG2 ? (p.block_n >> 1) : p.block_n;       // B rows (N extent) this CTA loads
      const int vk0 = sp * p.vk_per_split;
      const int vk1 = min(vk0 + p.vk_per_split, p.vk_total);
      const bool load_b = p.resident && (nt != prev_nt);
      int pass = vk0 / p.kb_total;                                   // passes: hi*hi, hi*lo, lo*hi
      int kb = vk0 - pass * p.kb_total;
      for (int vk = vk0; vk < vk1; ++vk) {
        const int k0 = kb * BLOCK_K;
        const CUtensorMap* ta = (pass == 2) ? &tmA1 : &tmA0;
        const CUtensorMap* tb = (pass == 1) ? &tmB1 : &tmB0;
        uint8_t* sA = tiles + static_cast<uint32_t>(stage) * stage_bytes;
        uint8_t* sB = p.resident ? (b_res + static_cast<uint32_t>(vk) * p.b_bytes) : (sA + a_stage_bytes);
        uint64_t* bar_b = p.resident ? &bfull_bar[vk] : &full_bar[stage];
        if (load_b) ptx::mbar_wait(&bempty_bar[vk], (b_gen & 1u) ^ 1u);   // new n-tile: the MMAs that read slab vk retired
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
        trace_at(p, 0, tr);
        if (ptx::elect_one()) {
          if (load_b) ptx::mbar_arrive_expect_tx(&bfull_bar[vk], p.b_bytes);
          if constexpr (CG2) {
            // Pair: both CTAs fill their own smem but report the bytes to the LEADER's full barrier, which therefore
            // expects both CTAs' stage bytes.
            const uint32_t lead_full = ptx::mapa(ptx::smem_u32(&full_bar[stage]), 0);
            if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2u * stage_bytes);
            if constexpr (!B_MN) {
              ptx::tma_load_2d_pair(sB, tb, lead_full, k0, n0);                  // box {BLOCK_K, block_n / 2}
            } else {
              for (int j = 0; j < nb_rows / BOX_MN; ++j)
                ptx::tma_load_2d_pair(sB + j * kBoxBytes, tb, lead_full, n0 + j * BOX_MN, k0);
            }
            if constexpr (!A_MN) {
              ptx::tma_load_2d_pair(sA, ta, lead_full, k0, m0);                  // box {BLOCK_K, 128}
            } else {
#pragma unroll
              for (int j = 0; j < kBlockM / BOX_MN; ++j)
                ptx::tma_load_2d_pair(sA + j * kBoxBytes, ta, lead_full, m0 + j * BOX_MN, k0);
            }
          } else {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
            if (!p.resident || load_b) {
              if constexpr (!B_MN) {
                ptx::tma_load_2d(sB, tb, bar_b, k0, n0);                         // box {BLOCK_K, block_n}
              } else {
                for (int j = 0; j < p.block_n / BOX_MN; ++j)
                  ptx::tma_load_2d(sB + j * kBoxBytes, tb, bar_b, n0 + j * BOX_MN, k0);
              }
            }
            for (int h = 0; h <= p.dual; ++h) {                                   // one or two 128-row A tiles
              uint8_t* sAh = sA + h * kABytes;
              const int mh = m0 + h * kBlockM;
              if constexpr (!A_MN) {
                ptx::tma_load_2d(sAh, ta, &full_bar[stage], k0, mh);            // box {BLOCK_K, 128}
              } else {
#pragma unroll
                for (int j = 0; j < kBlockM / BOX_MN; ++j)                        // boxes {BOX_MN, BLOCK_K}
                  ptx::tma_load_2d(sAh + j * kBoxBytes, ta, &full_bar[stage], mh + j * BOX_MN, k0);
              }
            }
          }
        }
        __syncwarp();
        trace_at(p, 1, tr++);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        if (++kb == p.kb_total) { kb = 0; ++pass; }
      }
      if (load_b) { ++b_gen; prev_nt = nt; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair: leader CTA only) =====================
    // Same structure: the whole warp walks the loop and waits on the barriers, one elected lane issues the MMAs.
    if (cta_rank == 0) {
      const uint32_t idesc = ptx::make_instr_desc(kTf32 ? 2u : 1u, A_MN, B_MN, CG2 ? 2 * kBlockM : kBlockM,
                                                  static_cast<uint32_t>(p.block_n));
      // Shared-memory descriptors: everything but the 14-bit start address is constant; one MMA step advances the address
      // field by a constant.  K-major: rows of 128 B, 8-row groups 1024 B apart (SBO), 32 B per MMA inside the swizzle
      // row.  MN-major: boxes of BLOCK_K k-rows x 128 B; LBO = box stride along MN, SBO = 1024 B (8 k-rows), UMMA_K
      // k-rows = UMMA_K * 128 B per MMA; 32-bit operands use the 32-byte-atom swizzle there (SBO = 512 B).
      const uint64_t da_base = A_MN ? ptx::make_smem_desc_sw128(0, kBoxBytes, kMnSbo, kMnLayout) : ptx::make_smem_desc_sw128(0, 16, 1024);
      const uint64_t db_base = B_MN ? ptx::make_smem_desc_sw128(0, kBoxBytes, kMnSbo, kMnLayout) : ptx::make_smem_desc_sw128(0, 16, 1024);
      constexpr uint32_t kStepA = (A_MN ? UMMA_K * kRowBytes : 32) >> 4;
      constexpr uint32_t kStepB = (B_MN ? UMMA_K * kRowBytes : 32) >> 4;
      const uint32_t tiles_addr = ptx::smem_u32(tiles);
      const uint32_t bres_addr = ptx::smem_u32(b_res);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      int prev_nt = -1; uint32_t b_gen = 0;
      int tr = 0;
      for (int w = w_begin; w < w_end; w += w_step, ++it) {
        const int rest = w / p.m_tiles;
        const int nt = rest % p.n_tiles;
        const int sp = rest / p.n_tiles;
        const int vk0 = sp * p.vk_per_split;
        const int vk1 = min(vk0 + p.vk_per_split, p.vk_total);
        const bool new_b = p.resident && (nt != prev_nt);
        // last tile of this CTA that uses the resident B of this n-tile -> release the slabs afterwards
        const bool last_of_nt = p.resident && ((w + 1 >= w_end) || (((w + 1) / p.m_tiles) % p.n_tiles != nt));
        // single-M: two accumulators ping-pong between MMA and epilogue; dual-M: both belong to this work item
        const bool one_set = p.dual || p.banks > 1;            // the work item owns all of TMEM: no ping-pong
        const int acc = one_set ? 0 : (it & 1);
        const uint32_t acc_phase = one_set ? (it & 1u) : ((it >> 1) & 1u);
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);      // epilogue has drained this accumulator
        trace_at(p, 4, it);
        ptx::tc_fence_after();
        const uint32_t d_tmem0 = tmem_base + static_cast<uint32_t>(acc * kAccCols);
        uint32_t banks_used = 0;
        int bpass = vk0 / p.kb_total;
        int bkb = vk0 - bpass * p.kb_total;
        for (int vk = vk0; vk < vk1; ++vk) {
          uint32_t d_tmem = d_tmem0;
          uint32_t bank_started = (vk > vk0) ? 1u : 0u;
          if (p.banks > 1) {
            const int bank = (bpass == 0) ? (bkb % (p.banks - 1)) : (p.banks - 1);
            d_tmem = tmem_base + static_cast<uint32_t>(bank * p.block_n);
            bank_started = (banks_used >> bank) & 1u;
            banks_used |= 1u << bank;
            if (++bkb == p.kb_total) { bkb = 0; ++bpass; }
          }
          if (new_b) ptx::mbar_wait(&bfull_bar[vk], b_gen & 1u);
          ptx::mbar_wait(&full_bar[stage], phase);              // TMA bytes have landed
          trace_at(p, 2, tr);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t a_addr = tiles_addr + static_cast<uint32_t>(stage) * stage_bytes;
            const uint32_t b_addr = p.resident ? (bres_addr + static_cast<uint32_t>(vk) * p.b_bytes) : (a_addr + a_stage_bytes);
            const uint64_t da0 = da_base | static_cast<uint64_t>((a_addr >> 4) & 0x3FFFu);
            const uint64_t db0 = db_base | static_cast<uint64_t>((b_addr >> 4) & 0x3FFFu);
            const uint32_t acc0 = bank_started;
#pragma unroll
            for (int k = 0; k < kMmaPerKBlock; ++k) {
              const uint64_t da = da0 + k * kStepA;
              const uint64_t db = db0 + k * kStepB;
              const uint32_t accum = (k > 0) ? 1u : acc0;
              if constexpr (CG2) ptx::umma_pair<kTf32>(d_tmem, da, db, idesc, accum);
              else ptx::umma<kTf32>(d_tmem, da, db, idesc, accum);
              if (p.dual)                                         // rows 128..255 of the item: same B, second accumulator
                ptx::umma<kTf32>(d_tmem + kAccCols, da + (kABytes >> 4), db, idesc, accum);
            }
            if constexpr (CG2) {
              ptx::umma_commit_pair(&empty_bar[stage]);           // frees the stage in BOTH CTAs when these MMAs retire
            } else {
              ptx::umma_commit(&empty_bar[stage]);                // frees the smem stage when these MMAs retire
              if (last_of_nt) ptx::umma_commit(&bempty_bar[vk]);  // ... and the resident B slab
            }
            if (vk + 1 == vk1) {                                  // accumulator complete -> epilogue (pair: both CTAs')
              if constexpr (CG2) ptx::umma_commit_pair(&tmem_full[acc]);
              else ptx::umma_commit(&tmem_full[acc]);
            }
          }
          __syncwarp();
          trace_at(p, 3, tr++);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        if (new_b) { ++b_gen; prev_nt = nt; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;                                     // TMEM lane quadrant this warp may access
    const int ew = warp - 2;
    const int quarter = ew >> 2;                                // which column group of the tile this warp drains
    const int cols_per_group = max(64, p.block_n / kColGroups);
    const int c_begin = min(quarter * cols_per_group, p.block_n);
    const int c_end = min(c_begin + cols_per_group, p.block_n); // empty for the upper groups of narrow tiles
    Epilogue e{p.col_scale, p.bias, p.alpha, p.act, p.aux, p.ldaux, p.aux_dtype, p.D, p.ldd, p.out_dtype, p.N};
    if (p.alpha_dev) e.alpha *= __ldg(p.alpha_dev);
    const int out_esz = (p.out_dtype == DMC_BF16) ? 2 : 4;
    bool vec_ok = ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0) && ((p.ldd * out_esz) % 16 == 0);
    bool aux_vec_ok = true;
    if (p.aux) {
      const int aux_esz = (p.aux_dtype == DMC_BF16) ? 2 : 4;
      aux_vec_ok = ((reinterpret_cast<uintptr_t>(p.aux) & 15) == 0) && ((p.ldaux * aux_esz) % 16 == 0);
    }
    uint8_t* my_staging = staging + ew * kStagingBytesPerWarp;
    const int chunks_per_box = (p.out_dtype == DMC_BF16) ? 2 : 1;   // 32-column chunks per 128-byte-wide box
    uint32_t n_boxes = 0;                                            // boxes this warp has stored so far
    // EPI 2, bounded logits: with |D| <= bound (device scalar: max gain of the weight-normed rows) and no center, every
    // y2 lies in [-shift, shift]; summing 2^(y2 - shift) directly is safe while 2*shift stays inside fp32's exponent range.
    bool stat_fixed = false;                    // fixed shift, no center: lean and generic paths
    bool stat_fixed_center = false;             // fixed shift WITH a center (teacher, bound = max |t| + max |center|): lean bf16 path only
    float stat_shift = 0.f;
    if constexpr (EPI >= 2) {
      if (p.stat_bound != nullptr) {
        const float b = __ldg(p.stat_bound) + (p.stat_bound2 != nullptr ? __ldg(p.stat_bound2) : 0.f);
        stat_shift = fabsf(b * p.stat_sc2) * 1.01f + 0.05f;
        if (p.stat_center == nullptr) stat_fixed = (stat_shift < 55.f);
        else stat_fixed_center = (stat_shift < 55.f) && p.out_dtype == DMC_BF16 && (reinterpret_cast<uintptr_t>(p.stat_center) & 15) == 0;
        if (!stat_fixed && !stat_fixed_center) stat_shift = 0.f;
      }
    }
    // lean path (epilogue_fast_acc) for full tiles stored through TMA; everything else takes the generic chunk code
    bool fast_ok = p.tma_store && p.partial == nullptr && p.dbg == 0 && c_begin < c_end && ((c_end - c_begin) & 63) == 0 &&
                   p.banks <= 1;
    if constexpr (EPI == 0) fast_ok = fast_ok && p.col_scale == nullptr && aux_vec_ok;
    if constexpr (EPI == 2) fast_ok = fast_ok && (stat_fixed || stat_fixed_center) && p.stat_colsum_partials == nullptr;
    if constexpr (EPI == 3)                            // lean teacher statistics: bf16 output, center + column sums requested
      fast_ok = fast_ok && p.stat_center != nullptr && p.out_dtype == DMC_BF16 &&
                (reinterpret_cast<uintptr_t>(p.stat_center) & 15) == 0;
    const bool out_bf16 = (p.out_dtype == DMC_BF16);
    int it = 0;
    for (int w = w_begin; w < w_end; w += w_step, ++it) {
      const int mt = w % p.m_tiles;
      const int rest = w / p.m_tiles;
      const int nt = rest % p.n_tiles;
      const int sp = rest / p.n_tiles;
      const int n0 = nt * p.block_n;
      const bool one_set = p.dual || p.banks > 1;
      const int acc = one_set ? 0 : (it & 1);
      const uint32_t acc_phase = one_set ? (it & 1u) : ((it >> 1) & 1u);
      uint32_t banks_used = 0;                                  // which accumulator banks this work item's MMAs wrote
      if (p.banks > 1) {
        const int vk0 = sp * p.vk_per_split;
        const int vk1 = min(vk0 + p.vk_per_split, p.vk_total);
        const int n_hh = max(0, min(vk1, p.kb_total) - vk0);    // hi*hi k-blocks of this split: consecutive from k-block vk0
        for (int i = 0; i < min(n_hh, p.banks - 1); ++i) banks_used |= 1u << ((vk0 + i) % (p.banks - 1));
        if (vk1 > p.kb_total) banks_used |= 1u << (p.banks - 1);
      }
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      if (warp == 2) trace_at(p, 5, it);
      ptx::tc_fence_after();
      if (fast_ok && n0 + p.block_n <= p.N) {
        auto release = [&]() {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG2) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tmem_empty[acc]), 0));   // leader's barrier
            else ptx::mbar_arrive(&tmem_empty[acc]);
          }
        };
        for (int h = 0; h <= p.dual; ++h) {
          const int row0 = mt * tile_m + (CG2 ? static_cast<int>(cta_rank) : h) * kBlockM + q * 32;
          const uint32_t t_addr = tmem_base + static_cast<uint32_t>((acc + h) * kAccCols) + (static_cast<uint32_t>(q * 32) << 16);
          float st_l = 0.f;
          float st_m = (EPI == 3) ? -INFINITY : stat_shift;       // EPI 3 tracks the running maximum, EPI 2 uses the bound
          if (out_bf16)
            epilogue_fast_acc<EPI, true>(p, &tmD, e, t_addr, row0, lane, n0, c_begin, c_end, my_staging, n_boxes, h == p.dual,
                                         release, st_m, st_l);
          else
            epilogue_fast_acc<EPI, false>(p, &tmD, e, t_addr, row0, lane, n0, c_begin, c_end, my_staging, n_boxes, h == p.dual,
                                          release, st_m, st_l);
          if constexpr (EPI >= 2) {
            const long long row = static_cast<long long>(row0) + lane;
            if (row < p.M) p.stat_row_partials[row * (kColGroups * p.n_tiles) + kColGroups * nt + quarter] = make_float2(st_m, st_l);
          }
        }
        if (warp == 2) trace_at(p, 6, it);
        continue;
      }
      const int ncols = min(p.block_n, p.N - n0);
      for (int h = 0; h <= p.dual; ++h) {                       // dual-M: drain both accumulators of the item
      const int row0 = mt * tile_m + (CG2 ? static_cast<int>(cta_rank) : h) * kBlockM + q * 32;
      const long long row = static_cast<long long>(row0) + lane;
      const uint32_t t_addr = tmem_base + static_cast<uint32_t>((acc + h) * kAccCols) + (static_cast<uint32_t>(q * 32) << 16);
      const bool last_h = (h == p.dual);
      float st_m = stat_fixed ? stat_shift : -INFINITY, st_l = 0.f;   // EPI 2: this row's softmax partial over [c_begin, c_end)
      // One 32-column chunk: accumulators -> epilogue -> split-K partials | direct store | smem staging + TMA store.
      auto process = [&](uint32_t (&r)[32], int c) {
        const int n = min(32, ncols - c);
        if (n <= 0) return;                                     // whole chunk beyond N (warp-uniform)
        float (&v)[32] = reinterpret_cast<float (&)[32]>(r);    // accumulators in place (EPI 2 leaves the stored values here)
        if (p.partial) {                                        // split-K: raw partial sums, epilogue runs in the reducer
          if (row < p.M) {
            float* dst = p.partial + (static_cast<long long>(sp) * p.M + row) * p.N + n0 + c;
            if (n == 32 && (p.N % 4 == 0)) {
#pragma unroll
              for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < n) dst[j] = v[j];
            }
          }
          return;
        }
        if constexpr (EPI >= 1) {
          if (e.alpha != 1.0f) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= e.alpha;
          }
        } else {
          if (row < p.M) epilogue_math(e, v, row, n0 + c, n, aux_vec_ok);
        }
        if (!p.tma_store) {
          if (row < p.M) epilogue_store_row(e, v, row, n0 + c, n, vec_ok);
          return;
        }
        // ---- smem-staged TMA store.  Box = 32 rows x 128 bytes (64 bf16 / 32 fp32 columns), 128B-swizzled:
        //      16-byte chunk j of row r lives at r*128 + ((j ^ (r & 7)) << 4)  -> conflict-free warp stores.
        const int sub = (c >> 5) % chunks_per_box;              // which half of the box this chunk fills
        uint8_t* buf = my_staging;
        if (p.dbg & 2) return;
        if (sub == 0 && n_boxes >= 1) {                         // the previous box's TMA store must have read the buffer
          if (lane == 0) ptx::tma_store_wait_read<0>();
          __syncwarp();
        }
        uint8_t* rowp = buf + lane * 128;
        if (p.out_dtype == DMC_BF16) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 pk = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                        pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
            *reinterpret_cast<uint4*>(rowp + ((((sub << 2) + j) ^ (lane & 7)) << 4)) = pk;
            if constexpr (EPI >= 2) {                           // statistics see exactly the values that are stored
              v[8 * j + 0] = bf16_lo(pk.x); v[8 * j + 1] = bf16_hi(pk.x); v[8 * j + 2] = bf16_lo(pk.y); v[8 * j + 3] = bf16_hi(pk.y);
              v[8 * j + 4] = bf16_lo(pk.z); v[8 * j + 5] = bf16_hi(pk.z); v[8 * j + 6] = bf16_lo(pk.w); v[8 * j + 7] = bf16_hi(pk.w);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(rowp + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        const bool box_done = (sub == chunks_per_box - 1) || (c + 32 >= ncols);
        if (box_done) {
          ptx::fence_proxy_async();                             // generic-proxy smem writes -> visible to the TMA engine
          __syncwarp();
          if (lane == 0 && !(p.dbg & 1)) {
            ptx::tma_store_2d(&tmD, buf, n0 + c - sub * 32, row0);
            ptx::tma_store_commit();
          }
          ++n_boxes;
          if constexpr (EPI >= 2) {
            if (p.stat_colsum_partials != nullptr && row0 < p.M) {
              // column sums over this warp's 32 rows, read back from the staged (swizzled) box: lane j owns the
              // 32-bit word j of every 128-byte row -> conflict-free; 2 bf16 columns or 1 fp32 column per lane.
              const int nrows = min(32, p.M - row0);
              float s0 = 0.f, s1 = 0.f;
              for (int r = 0; r < nrows; ++r) {
                const uint32_t wv = *reinterpret_cast<const uint32_t*>(buf + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
                if (p.out_dtype == DMC_BF16) { s0 += bf16_lo(wv); s1 += bf16_hi(wv); }
                else s0 += __uint_as_float(wv);
              }
              float* dst = p.stat_colsum_partials + static_cast<long long>(row0 >> 5) * p.N;
              const int cbox = n0 + c - sub * 32;
              if (p.out_dtype == DMC_BF16) {
                if (cbox + 2 * lane < p.N) dst[cbox + 2 * lane] = s0;
                if (cbox + 2 * lane + 1 < p.N) dst[cbox + 2 * lane + 1] = s1;
              } else {
                if (cbox + lane < p.N) dst[cbox + lane] = s0;
              }
            }
          }
        }
      };
      // EPI 2: softmax partial of this thread's row over the 32 columns of chunk c (base-2 domain, four independent
      // chains).  Runs AFTER the chunk pair has been staged and its TMA store issued, so the SFU work overlaps the
      // store's shared-memory read instead of sitting in front of it.
      auto chunk_stats = [&](uint32_t (&r)[32], int c) {
        const int n = min(32, ncols - c);
        if (n <= 0) return;
        float (&v)[32] = reinterpret_cast<float (&)[32]>(r);
        if (stat_fixed && n == 32) {
          // |y2| <= stat_shift is known (rows and weights are unit / g-bounded): no running max, 3 instr/logit
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            a0 += ex2(fmaf(v[j + 0], p.stat_sc2, -stat_shift));
            a1 += ex2(fmaf(v[j + 1], p.stat_sc2, -stat_shift));
            a2 += ex2(fmaf(v[j + 2], p.stat_sc2, -stat_shift));
            a3 += ex2(fmaf(v[j + 3], p.stat_sc2, -stat_shift));
          }
          st_l += (a0 + a1) + (a2 + a3);
        } else if (stat_fixed) {
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            a0 += (j + 0 < n) ? ex2(fmaf(v[j + 0], p.stat_sc2, -stat_shift)) : 0.f;
            a1 += (j + 1 < n) ? ex2(fmaf(v[j + 1], p.stat_sc2, -stat_shift)) : 0.f;
            a2 += (j + 2 < n) ? ex2(fmaf(v[j + 2], p.stat_sc2, -stat_shift)) : 0.f;
            a3 += (j + 3 < n) ? ex2(fmaf(v[j + 3], p.stat_sc2, -stat_shift)) : 0.f;
          }
          st_l += (a0 + a1) + (a2 + a3);
        } else {
          float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
          if (p.stat_center != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float cb = -__ldg(p.stat_center + min(n0 + c + j, p.N - 1)) * p.stat_sc2;
              v[j] = (j < n) ? fmaf(v[j], p.stat_sc2, cb) : -INFINITY;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (j < n) ? v[j] * p.stat_sc2 : -INFINITY;
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            m0 = fmaxf(m0, v[j]); m1 = fmaxf(m1, v[j + 1]); m2 = fmaxf(m2, v[j + 2]); m3 = fmaxf(m3, v[j + 3]);
          }
          const float cm = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          if (cm > st_m) { st_l *= ex2(st_m - cm); st_m = cm; }
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            a0 += ex2(v[j] - st_m); a1 += ex2(v[j + 1] - st_m); a2 += ex2(v[j + 2] - st_m); a3 += ex2(v[j + 3] - st_m);
          }
          st_l += (a0 + a1) + (a2 + a3);
        }
      };
      if (p.banks > 1) {
        // 3xTF32 with accumulator banks: sum the banks in fp32 registers (exact IEEE adds), 32 columns at a time
        for (int c = c_begin; c < c_end; c += 32) {
          uint32_t ra[32], rt[32];
          float (&va)[32] = reinterpret_cast<float (&)[32]>(ra);
          float (&vt)[32] = reinterpret_cast<float (&)[32]>(rt);
#pragma unroll
          for (int j = 0; j < 32; ++j) va[j] = 0.f;
          for (int b = 0; b < p.banks; ++b) {
            if (!((banks_used >> b) & 1u)) continue;            // warp-uniform
            ptx::tmem_ld_32x32(tmem_base + static_cast<uint32_t>(b * p.block_n + c) + (static_cast<uint32_t>(q * 32) << 16), rt);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) va[j] += vt[j];
          }
          if (c + 32 >= c_end) {                                // this warp's last read of the accumulators: hand back
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
          }
          process(ra, c);
          if constexpr (EPI >= 2) chunk_stats(ra, c);
        }
      } else
      for (int c = c_begin; c < c_end; c += 64) {               // this warp's column range, 64 columns at a time
        uint32_t ra[32], rb[32];
        if (!(p.dbg & 4)) {
          ptx::tmem_ld_32x32(t_addr + c, ra);                   // two TMEM loads in flight per wait
          ptx::tmem_ld_32x32(t_addr + c + 32, rb);
          ptx::tmem_ld_wait();
        }
        if (last_h && c + 64 >= c_end) {                        // this warp's last read of the accumulator(s): hand back
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG2) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tmem_empty[acc]), 0));   // leader's barrier
            else ptx::mbar_arrive(&tmem_empty[acc]);
          }
        }
        process(ra, c);
        process(rb, c + 32);
        if constexpr (EPI >= 2) {
          chunk_stats(ra, c);
          chunk_stats(rb, c + 32);
        }
      }
      if constexpr (EPI >= 2) {
        if (row < p.M && c_begin < c_end)
          p.stat_row_partials[row * (kColGroups * p.n_tiles) + kColGroups * nt + quarter] = make_float2(st_m, st_l);
      }
      if (last_h && c_begin >= c_end) {                         // nothing to drain (block_n == 64, upper half): still release
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG2) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tmem_empty[acc]), 0));
          else ptx::mbar_arrive(&tmem_empty[acc]);
        }
      }
      }  // h
      if (warp == 2) trace_at(p, 6, it);
    }
    if (p.tma_store && lane == 0) ptx::tma_store_wait_all<0>();  // all bulk stores complete before the CTA exits
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_at(p, 7, 2);
  if constexpr (CG2) ptx::cluster_sync();       // the peer may still signal our barriers / the leader still writes our TMEM
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (CG2) ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
    else ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Sums the split-K partials and applies the epilogue.  One thread per 4 consecutive columns.
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int splits, long long M, int N, Epilogue e, const float* alpha_dev) {
  pdl_prologue();
  const long long groups_per_row = (N + 3) / 4;
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (g >= M * groups_per_row) return;
  const long long row = g / groups_per_row;
  const int col0 = static_cast<int>(g % groups_per_row) * 4;
  const int n = min(4, N - col0);
  if (alpha_dev) e.alpha *= __ldg(alpha_dev);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s = 0; s < splits; ++s) {
    const float* src = partial + (static_cast<long long>(s) * M + row) * N + col0;
    if (n == 4 && (N % 4 == 0)) {
      float4 t = *reinterpret_cast<const float4*>(src);
      acc[0] += t.x; acc[1] += t.y; acc[2] += t.z; acc[3] += t.w;
    } else {
      for (int j = 0; j < n; ++j) acc[j] += src[j];
    }
  }
  for (int j = 0; j < n; ++j) {
    float v = acc[j];
    const int col = col0 + j;
    if (e.col_scale) v *= __ldg(e.col_scale + col);
    v *= e.alpha;
    if (e.bias) v += __ldg(e.bias + col);
    if (e.act == DMC_ACT_GELU) {
      if (e.aux) store_elem(e.aux, row * e.ldaux + col, e.aux_dtype, v);
      v = gelu_f(v);
    } else if (e.act == DMC_ACT_GELU_BWD) {
      v *= gelu_grad_f(load_elem(e.aux, row * e.ldaux + col, e.aux_dtype));
    } else if (e.act == DMC_ACT_GELU_DG) {
      store_elem(e.aux, row * e.ldaux + col, e.aux_dtype, gelu_grad_f(v));
      v = gelu_f(v);
    } else if (e.act == DMC_ACT_MUL_AUX) {
      v *= load_elem(e.aux, row * e.ldaux + col, e.aux_dtype);
    }
    store_elem(e.D, row * e.ldd + col, e.out_dtype, v);
  }
}

// Split-K reduction fused with the backward of F.normalize (DMC_ACT_NORMALIZE_BWD): one warp per output row (N <= 1024).
// Sums the partials in fixed order (deterministic), then dz = (dzhat - (dzhat . zhat) zhat) / ||z|| and the store in the
// output dtype -- replaces splitk_reduce + normalize_bwd (+ the bf16 cast of dz in the bf16-GEMM mode): three launches on the
// critical path between the last layer's dgrad and the MLP backward.
__global__ void __launch_bounds__(256)
splitk_normalize_bwd_kernel(const float* __restrict__ partial, int splits, long long M, int N, float alpha, const float* alpha_dev,
                            const float* __restrict__ zhat, long long ldz, const float* __restrict__ row_scale, float eps,
                            void* __restrict__ D, long long ldd, int out_dtype) {
  pdl_prologue();
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  if (alpha_dev) alpha *= __ldg(alpha_dev);
  const float inv = row_scale[row];
  const bool clamped = (inv * eps >= 1.0f);
  constexpr int kMaxPerLane = 8;                       // N <= 1024: up to 8 float4 per lane
  float4 acc[kMaxPerLane];
  const int nv = N >> 2;                               // N % 4 == 0 (checked on the host)
  float proj = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int v = lane + 32 * i;
    if (v < nv) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < splits; ++s) {
        const float4 t = *reinterpret_cast<const float4*>(partial + (static_cast<long long>(s) * M + row) * N + 4 * v);
        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
      }
      a.x *= alpha; a.y *= alpha; a.z *= alpha; a.w *= alpha;
      acc[i] = a;
      if (!clamped) {
        const float4 z = *reinterpret_cast<const float4*>(zhat + row * ldz + 4 * v);
        proj = fmaf(a.x, z.x, proj); proj = fmaf(a.y, z.y, proj); proj = fmaf(a.z, z.z, proj); proj = fmaf(a.w, z.w, proj);
      }
    }
  }
  proj = warp_sum(proj);
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int v = lane + 32 * i;
    if (v < nv) {
      float4 a = acc[i];
      if (!clamped) {
        const float4 z = *reinterpret_cast<const float4*>(zhat + row * ldz + 4 * v);
        a.x -= proj * z.x; a.y -= proj * z.y; a.z -= proj * z.z; a.w -= proj * z.w;
      }
      a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
      if (out_dtype == DMC_BF16)
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(D) + row * ldd + 4 * v) = make_uint2(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w));
      else
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(D) + row * ldd + 4 * v) = a;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;      // benign race: every thread computes the same pointer
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Tensor map over a row-major matrix with `rows` rows of `cols` contiguous elements (row stride ld),
// box = {box_cols (inner, 128 bytes), box_rows}, 128B swizzle, out-of-bounds elements read as zero.
int make_tmap(CUtensorMap* tm, const void* base, int esz, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
              bool atom32 = false) {
  EncodeTiledFn fn = get_encode_fn();
  DMC_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  DMC_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "dmc_gemm: operand base pointer must be 16-byte aligned");
  DMC_REQUIRE((ld * esz) % 16 == 0, "dmc_gemm: operand row stride must be a multiple of 16 bytes (ld=%lld)", (long long)ld);
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * esz};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DMC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

struct Plan {
  int block_n, m_tiles, n_tiles, kb_total, passes, vk_total, splits, vk_per_split, stages, resident, tma_store, dual, cg2, banks;
  size_t smem_bytes, workspace_bytes;
};

// DMC_GEMM_FLAGS (debug / A-B measurements): bit 0 = no TMA-store epilogue, bit 1 = no B-resident schedule,
// bit 2 = no dual-M work items, bits 3-5 = epilogue ablations (timing only), bit 6 = use CTA pairs (cta_group::2) where M > 128,
// bit 8 = no CTA pairs for long contractions.
int debug_flags() {
  static int flags = -1;
  if (flags < 0) {
    const char* e = getenv("DMC_GEMM_FLAGS");
    flags = e ? atoi(e) : 0;
  }
  return flags;
}

Plan make_plan(int64_t M, int64_t N, int64_t K, int in_dtype, bool three_pass, int forced_split, bool store_ok = false,
               bool want_stats = false) {
  Plan pl{};
  const int esz = (in_dtype == DMC_BF16) ? 2 : 4;
  const int block_k = kRowBytes / esz;
  // Tile width: as wide as possible (fewest re-reads of A), but for short contractions prefer enough tiles to fill
  // the 148 SMs over a split-K pass; long contractions keep the wide tile and split K instead (A is read once).
  // CTA pairs (cta_group::2): 256-row items, half a B tile per CTA.  Correct in every layout (tests run both), but on
  // these shapes measured slower than the single-CTA schedule (8192^3: 1.11 vs 1.43 PFLOP/s), so it is opt-in:
  // DMC_GEMM_FLAGS bit 6.
  // Exception: very long contractions with few output tiles (the last layer's dgrad: 2048 x 256 x 65536).  There the pair
  // halves the split-K partial traffic (9 instead of 18 splits for the same 144 CTAs) and the smem reads per MMA:
  // measured 84 -> 76 us.  DMC_GEMM_FLAGS bit 8 disables it.
  // 3xTF32 (fp32 parity mode): accumulator banks keep every tcgen05 accumulation chain short (see GemmDev::banks).  At most
  // kMaxChainKb hi*hi k-blocks (x4 MMAs) per bank: the widest tile whose banks reach that without splitting the contraction,
  // else 64-wide tiles (8 banks) plus split-K.  DMC_GEMM_FLAGS bit 9 turns the banks off (A/B measurements of the error).
  constexpr int kMaxChainKb = 22;
  const bool use_banks = three_pass && !(debug_flags() & 512);
  const bool long_k = (K >= 32768 && N <= 256 && !(debug_flags() & 256)) && !use_banks;
  pl.cg2 = (M > kBlockM && !use_banks && ((debug_flags() & 64) || long_k)) ? 1 : 0;
  const int units = pl.cg2 ? kNumSMs / 2 : kNumSMs;                 // schedulable units: CTA pairs or CTAs
  const int64_t mt = ceil_div(M, pl.cg2 ? 2 * kBlockM : kBlockM);
  int bn = N > 128 ? 256 : ((N > 64 || pl.cg2) ? 128 : 64);
  if (K < 8192 && forced_split == 0 && !want_stats) {
    const int bn_min = pl.cg2 ? 128 : 64;       // a CTA pair splits the B tile in two: each half needs >= one 64-wide box
    while (bn > bn_min && mt * ceil_div(N, bn) < (units * 2) / 3) bn >>= 1;
  }
  if (want_stats) { bn = 256; forced_split = 1; }         // statistics parts are defined on 256-wide unsplit tiles
  pl.banks = 1;
  if (use_banks) {
    const int kb = static_cast<int>(ceil_div(K, block_k));
    if (!want_stats) {
      bn = 64;
      for (int cand = 256; cand >= 64; cand >>= 1) {
        if (cand > 64 && cand / 2 >= N) continue;             // no wider than the output needs
        if (ceil_div(kb, kTmemCols / cand - 1) <= kMaxChainKb) { bn = cand; break; }
      }
    }
    pl.banks = kTmemCols / bn;
    if (forced_split == 0) {
      const int64_t per_split = static_cast<int64_t>(kMaxChainKb) * (pl.banks - 1);   // virtual k-blocks per split
      forced_split = static_cast<int>(ceil_div(3 * static_cast<int64_t>(kb), per_split));
      if (ceil_div(kb, pl.banks - 1) <= kMaxChainKb) forced_split = 1;
    }
  }
  if ((debug_flags() & 128) && !want_stats && bn == 256 && K <= 512 && mt * ceil_div(N, 128) >= 4 * units) bn = 128;
  pl.block_n = bn;
  pl.n_tiles = static_cast<int>(ceil_div(N, pl.block_n));
  pl.kb_total = static_cast<int>(ceil_div(K, block_k));
  pl.passes = three_pass ? 3 : 1;
  pl.vk_total = pl.kb_total * pl.passes;
  // dual-M (single-CTA kernels only): 256-row work items whose two accumulators share every B k-block from smem.
  pl.dual = 0;
  if (!pl.cg2 && !(debug_flags() & 4) && M > kBlockM && pl.vk_total >= 16 && !want_stats && pl.banks == 1) {
    const int64_t items = ceil_div(M, 2 * kBlockM) * pl.n_tiles;
    if (K >= 8192 || items >= kNumSMs) pl.dual = 1;
  }
  pl.m_tiles = static_cast<int>(ceil_div(M, (pl.dual || pl.cg2) ? 2 * kBlockM : kBlockM));
  const int tiles = pl.m_tiles * pl.n_tiles;
  int splits = 1;
  if (forced_split >= 1) {
    splits = forced_split;
  } else if (tiles * 3 <= units) {                         // under a third of a wave: split the contraction
    splits = units / tiles;
    const int max_by_k = pl.vk_total / 4 > 0 ? pl.vk_total / 4 : 1;   // keep >= 4 k-blocks per split
    if (splits > max_by_k) splits = max_by_k;
  }
  if (splits > pl.vk_total) splits = pl.vk_total;
  if (splits < 1) splits = 1;
  pl.vk_per_split = static_cast<int>(ceil_div(pl.vk_total, splits));
  pl.splits = static_cast<int>(ceil_div(pl.vk_total, pl.vk_per_split));   // every split owns >= 1 k-block
  const size_t b_bytes = static_cast<size_t>(pl.block_n) * kRowBytes;
  pl.tma_store = (store_ok && pl.splits == 1 && !(debug_flags() & 1)) ? 1 : 0;
  const size_t kBarrierBytes = 512;
  // DMC_GEMM_SMEM_RESERVE_KB (timing experiments): shared memory left unused so that small CTAs of other kernels
  // (1 KiB reserved each) can be co-resident with the one-CTA-per-SM GEMM
  static const size_t smem_reserve = [] { const char* e = getenv("DMC_GEMM_SMEM_RESERVE_KB"); return e ? static_cast<size_t>(atoi(e)) * 1024 : size_t(0); }();
  size_t budget = 227 * 1024 - 1024 /*alignment slack*/ - kBarrierBytes - (pl.tma_store ? kStagingBytes : 0) - smem_reserve;
  // B-resident schedule: the whole contraction's worth of B for one n-tile stays in smem (<= 8 slabs, <= 128 KiB),
  // leaving >= 4 A-only stages.  Pays off when several m-tiles share an n-tile.
  const size_t res_bytes = static_cast<size_t>(pl.vk_total) * b_bytes;
  pl.resident = (pl.splits == 1 && pl.vk_total <= kMaxStages && pl.m_tiles >= 2 && !pl.dual && !pl.cg2 && !(debug_flags() & 2) &&
                 res_bytes + 4 * kABytes <= budget) ? 1 : 0;
  const size_t stage_bytes = pl.resident ? kABytes : ((pl.dual ? 2 : 1) * kABytes + (pl.cg2 ? b_bytes / 2 : b_bytes));
  if (pl.resident) budget -= res_bytes;
  int stages = static_cast<int>(budget / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  pl.stages = stages;
  pl.smem_bytes = (pl.resident ? res_bytes : 0) + static_cast<size_t>(stages) * stage_bytes +
                  (pl.tma_store ? kStagingBytes : 0) + kBarrierBytes + 1024;
  pl.workspace_bytes = pl.splits > 1 ? static_cast<size_t>(pl.splits) * M * N * sizeof(float) : 0;
  return pl;
}

template <int ESZ, bool A_MN, bool B_MN, int EPI, bool CG2>
int launch_tc2(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0, const CUtensorMap& b1,
               const CUtensorMap& td, const GemmDev& dev, size_t smem_bytes, int grid, cudaStream_t st) {
  auto kern = gemm_tc_kernel<ESZ, A_MN, B_MN, EPI, CG2>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes));
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(gemm_tc_kernel)");
  if constexpr (CG2) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    e = cudaLaunchKernelEx(&cfg, kern, a0, a1, b0, b1, td, dev);
    if (e != cudaSuccess) return cuda_status(e, "cudaLaunchKernelEx(gemm_tc_kernel, cluster 2)");
  } else {
    launch_kernel(kern, dim3(grid), dim3(kThreads), smem_bytes, st, a0, a1, b0, b1, td, dev);
  }
  DMC_LAUNCH_CHECK("gemm_tc_kernel launch");
  return 0;
}

template <int ESZ, bool A_MN, bool B_MN, int EPI>
int launch_tc(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0, const CUtensorMap& b1,
              const CUtensorMap& td, const GemmDev& dev, size_t smem_bytes, int grid, cudaStream_t st, bool cg2) {
  return cg2 ? launch_tc2<ESZ, A_MN, B_MN, EPI, true>(a0, a1, b0, b1, td, dev, smem_bytes, grid, st)
             : launch_tc2<ESZ, A_MN, B_MN, EPI, false>(a0, a1, b0, b1, td, dev, smem_bytes, grid, st);
}

}  // namespace
}  // namespace dmc

using namespace dmc;

extern "C" int64_t dmc_gemm_stats_parts(int64_t N) { return N > 0 ? kColGroups * ceil_div(N, 256) : 0; }

extern "C" size_t dmc_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int32_t in_dtype) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  Plan a = make_plan(M, N, K, in_dtype, false, 0);
  Plan b = make_plan(M, N, K, in_dtype, in_dtype == DMC_F32, 0);
  return a.workspace_bytes > b.workspace_bytes ? a.workspace_bytes : b.workspace_bytes;
}

extern "C" int dmc_gemm(const dmc_gemm_args* a, void* stream) {
  DMC_REQUIRE(a != nullptr, "dmc_gemm: null args");
  DMC_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "dmc_gemm: empty problem M=%lld N=%lld K=%lld", (long long)a->M, (long long)a->N, (long long)a->K);
  DMC_REQUIRE(a->M < (1ll << 31) && a->N < (1ll << 31) && a->K < (1ll << 31), "dmc_gemm: dimension too large");
  DMC_REQUIRE(a->A && a->B && a->D, "dmc_gemm: null operand");
  DMC_REQUIRE(a->in_dtype == DMC_BF16 || a->in_dtype == DMC_F32, "dmc_gemm: bad in_dtype %d", a->in_dtype);
  DMC_REQUIRE(a->out_dtype == DMC_BF16 || a->out_dtype == DMC_F32, "dmc_gemm: bad out_dtype %d", a->out_dtype);
  DMC_REQUIRE(a->act >= DMC_ACT_NONE && a->act <= DMC_ACT_MUL_AUX, "dmc_gemm: bad act %d", a->act);
  DMC_REQUIRE((a->act != DMC_ACT_GELU_DG && a->act != DMC_ACT_MUL_AUX) || a->aux != nullptr, "dmc_gemm: DMC_ACT_GELU_DG / DMC_ACT_MUL_AUX need aux");
  const bool norm_bwd = (a->act == DMC_ACT_NORMALIZE_BWD);
  if (norm_bwd) {
    DMC_REQUIRE(a->aux != nullptr && a->aux_dtype == DMC_F32 && a->row_scale != nullptr, "dmc_gemm: DMC_ACT_NORMALIZE_BWD needs aux (fp32 rows) and row_scale");
    DMC_REQUIRE(a->N % 4 == 0 && a->N <= 1024 && a->col_scale == nullptr && a->bias == nullptr, "dmc_gemm: DMC_ACT_NORMALIZE_BWD needs N %% 4 == 0, N <= 1024 and no column scale / bias");
    DMC_REQUIRE((reinterpret_cast<uintptr_t>(a->aux) & 15) == 0 && (a->ldaux % 4) == 0 && (reinterpret_cast<uintptr_t>(a->D) & 15) == 0 &&
                (a->ldd * ((a->out_dtype == DMC_BF16) ? 2 : 4)) % 16 == 0, "dmc_gemm: DMC_ACT_NORMALIZE_BWD needs 16-byte aligned rows");
  }
  DMC_REQUIRE(a->act != DMC_ACT_GELU_BWD || a->aux != nullptr, "dmc_gemm: DMC_ACT_GELU_BWD needs aux");
  DMC_REQUIRE((a->A_lo == nullptr) == (a->B_lo == nullptr), "dmc_gemm: A_lo and B_lo must be given together");
  const bool three = (a->A_lo != nullptr);
  DMC_REQUIRE(!three || a->in_dtype == DMC_F32, "dmc_gemm: hi/lo split operands require in_dtype F32");
  const int esz = (a->in_dtype == DMC_BF16) ? 2 : 4;
  const int block_k = kRowBytes / esz, box_mn = kRowBytes / esz;
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  const int out_esz = (a->out_dtype == DMC_BF16) ? 2 : 4;
  const bool store_ok = ((reinterpret_cast<uintptr_t>(a->D) & 15) == 0) && ((a->ldd * out_esz) % 16 == 0);
  Plan pl = make_plan(a->M, a->N, a->K, a->in_dtype, three, a->split_k, store_ok, a->stat_row_partials != nullptr);
  if (norm_bwd && pl.splits < 2 && a->K >= 2 * block_k)        // the row operation lives in the split-K reducer
    pl = make_plan(a->M, a->N, a->K, a->in_dtype, three, 2, store_ok, false);
  DMC_REQUIRE(!norm_bwd || pl.splits >= 2, "dmc_gemm: DMC_ACT_NORMALIZE_BWD needs a contraction of at least two k-blocks");
  if (pl.splits > 1) {
    DMC_REQUIRE(a->workspace != nullptr && a->workspace_bytes >= pl.workspace_bytes,
                "dmc_gemm: split-K needs a workspace of %zu bytes (got %zu)", pl.workspace_bytes, a->workspace_bytes);
    DMC_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0, "dmc_gemm: workspace must be 16-byte aligned");
  }

  CUtensorMap tA0, tA1, tB0, tB1, tD;
  int rc;
  auto mapA = [&](CUtensorMap* tm, const void* base) {
    return a->a_mn_major ? make_tmap(tm, base, esz, a->K, a->M, a->lda, box_mn, block_k, esz == 4)   // stored [K,M]
                         : make_tmap(tm, base, esz, a->M, a->K, a->lda, block_k, kBlockM);    // stored [M,K]
  };
  auto mapB = [&](CUtensorMap* tm, const void* base) {
    return a->b_mn_major ? make_tmap(tm, base, esz, a->K, a->N, a->ldb, box_mn, block_k, esz == 4)   // stored [K,N]
                         : make_tmap(tm, base, esz, a->N, a->K, a->ldb, block_k, pl.cg2 ? pl.block_n / 2 : pl.block_n); // [N,K]
  };
  if ((rc = mapA(&tA0, a->A))) return rc;
  if ((rc = mapB(&tB0, a->B))) return rc;
  if ((rc = mapA(&tA1, three ? a->A_lo : a->A))) return rc;
  if ((rc = mapB(&tB1, three ? a->B_lo : a->B))) return rc;
  if (pl.tma_store) {            // output boxes: 32 rows x 128 bytes, same 128B swizzle as the staging writes
    if ((rc = make_tmap(&tD, a->D, out_esz, a->M, a->N, a->ldd, kRowBytes / out_esz, 32))) return rc;
  } else {
    tD = tA0;
  }

  GemmDev d{};
  d.M = static_cast<int>(a->M); d.N = static_cast<int>(a->N);
  d.block_n = pl.block_n; d.m_tiles = pl.m_tiles; d.n_tiles = pl.n_tiles;
  d.kb_total = pl.kb_total; d.vk_total = pl.vk_total; d.splits = pl.splits; d.vk_per_split = pl.vk_per_split;
  d.stages = pl.stages; d.b_bytes = static_cast<uint32_t>(pl.block_n) * kRowBytes;
  d.resident = pl.resident; d.tma_store = pl.tma_store; d.dual = pl.dual; d.banks = pl.banks;
  d.stat_sc2 = a->stat_scale * 1.4426950408889634f; d.stat_center = a->stat_center;
  d.stat_row_partials = reinterpret_cast<float2*>(a->stat_row_partials); d.stat_colsum_partials = a->stat_colsum_partials;
  d.stat_bound = a->stat_bound; d.stat_bound2 = a->stat_bound2;
  d.dbg = (debug_flags() >> 3) & 7;
  static const bool trace_on = (getenv("DMC_GEMM_TRACE") != nullptr);       // debug only: never set in production
  static long long* trace_dev = nullptr;
  if (trace_on) {
    if (trace_dev == nullptr) cudaMalloc(&trace_dev, 8 * 512 * sizeof(long long));
    cudaMemsetAsync(trace_dev, 0, 8 * 512 * sizeof(long long), st);
    d.trace = trace_dev;
  }
  if (a->stat_row_partials != nullptr) {
    DMC_REQUIRE(a->col_scale == nullptr && a->bias == nullptr && a->act == DMC_ACT_NONE && !a->a_mn_major && !a->b_mn_major,
                "dmc_gemm: fused statistics need a plain epilogue and K-major operands");
    DMC_REQUIRE(pl.tma_store && pl.block_n == 256 && !pl.dual && pl.splits == 1,
                "dmc_gemm: fused statistics need a 16-byte aligned output (pointer and row stride)");
  }
  d.D = a->D; d.ldd = a->ldd; d.out_dtype = a->out_dtype;
  d.partial = pl.splits > 1 ? static_cast<float*>(a->workspace) : nullptr;
  d.col_scale = a->col_scale; d.bias = a->bias; d.alpha_dev = a->alpha_dev; d.alpha = a->alpha;
  d.act = a->act; d.aux = a->aux; d.ldaux = a->ldaux; d.aux_dtype = a->aux_dtype;

  const int num_work = pl.m_tiles * pl.n_tiles * pl.splits;
  const int cap = (a->max_ctas > 0 && a->max_ctas < kNumSMs) ? a->max_ctas : kNumSMs;
  int grid = num_work < cap ? num_work : cap;
  if (pl.cg2) {                                   // CTA pairs: one cluster of 2 per work item slot
    const int pairs_cap = cap / 2;
    grid = 2 * (num_work < pairs_cap ? num_work : pairs_cap);
  }
  const bool amn = a->a_mn_major != 0, bmn = a->b_mn_major != 0;
const bool plain = (a->col_scale == nullptr && a->bias == nullptr && (a->act == DMC_ACT_NONE || norm_bwd));
  const bool stats = (a->stat_row_partials != nullptr);
#define DMC_LAUNCH(ESZ_, AMN_, BMN_)                                                                          \
  (plain ? launch_tc<ESZ_, AMN_, BMN_, 1>(tA0, tA1, tB0, tB1, tD, d, pl.smem_bytes, grid, st, pl.cg2)                 \
         : launch_tc<ESZ_, AMN_, BMN_, 0>(tA0, tA1, tB0, tB1, tD, d, pl.smem_bytes, grid, st, pl.cg2))
#define DMC_DISPATCH(ESZ_)                                                                                    \
  (amn ? (bmn ? DMC_LAUNCH(ESZ_, true, true) : DMC_LAUNCH(ESZ_, true, false))                                 \
       : (bmn ? DMC_LAUNCH(ESZ_, false, true) : DMC_LAUNCH(ESZ_, false, false)))
  if (stats) {          // last-layer forward only: K-major operands
    if (a->stat_center != nullptr && esz == 2 && a->stat_bound == nullptr)   // teacher: center + running maximum (+ column sums), lean path in EPI 3
      rc = launch_tc<2, false, false, 3>(tA0, tA1, tB0, tB1, tD, d, pl.smem_bytes, grid, st, pl.cg2);
    else
      rc = (esz == 2) ? launch_tc<2, false, false, 2>(tA0, tA1, tB0, tB1, tD, d, pl.smem_bytes, grid, st, pl.cg2)
                      : launch_tc<4, false, false, 2>(tA0, tA1, tB0, tB1, tD, d, pl.smem_bytes, grid, st, pl.cg2);
  } else {
    rc = (esz == 2) ? DMC_DISPATCH(2) : DMC_DISPATCH(4);
  }
#undef DMC_LAUNCH
#undef DMC_DISPATCH
  if (rc) return rc;
  if (trace_on) {
    static long long h[8 * 512];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost);
    const long long t0 = h[7 * 512];
    fprintf(stderr, "TRACE M=%lld N=%lld K=%lld bn=%d stages=%d resident=%d dual=%d splits=%d grid=%d | setup=%lld end=%lld\n",
            (long long)a->M, (long long)a->N, (long long)a->K, pl.block_n, pl.stages, pl.resident, pl.dual, pl.splits, grid,
            h[7 * 512 + 1] - t0, h[7 * 512 + 2] - t0);
    for (int i = 0; i < 512 && h[2 * 512 + i]; ++i)
      fprintf(stderr, "  kb%-3d P.empty=%-7lld P.issued=%-7lld M.full=%-7lld M.commit=%-7lld\n", i, h[i] - t0, h[512 + i] - t0,
              h[2 * 512 + i] - t0, h[3 * 512 + i] - t0);
    for (int i = 0; i < 512 && h[5 * 512 + i]; ++i)
      fprintf(stderr, "  tile%-3d M.tmem_empty=%-7lld E.tmem_full=%-7lld E.done=%-7lld\n", i, h[4 * 512 + i] - t0, h[5 * 512 + i] - t0,
              h[6 * 512 + i] - t0);
  }

  if (pl.splits > 1 && norm_bwd) {
    launch_kernel(splitk_normalize_bwd_kernel, dim3(static_cast<unsigned>(ceil_div(a->M, 8))), dim3(256), 0, st,
                  static_cast<const float*>(a->workspace), pl.splits, a->M, static_cast<int>(a->N), a->alpha, a->alpha_dev,
                  static_cast<const float*>(a->aux), a->ldaux, a->row_scale, a->row_eps, a->D, a->ldd, a->out_dtype);
    DMC_LAUNCH_CHECK("splitk_normalize_bwd_kernel launch");
  } else if (pl.splits > 1) {
    Epilogue e{a->col_scale, a->bias, a->alpha, a->act, a->aux, a->ldaux, a->aux_dtype, a->D, a->ldd, a->out_dtype,
               static_cast<int>(a->N)};
    const long long groups = a->M * ((a->N + 3) / 4);
    const int blocks = static_cast<int>(ceil_div(groups, 256));
    launch_kernel(splitk_reduce_kernel, dim3(blocks), dim3(256), 0, st, static_cast<const float*>(a->workspace), pl.splits, a->M,
                                                 static_cast<int>(a->N), e, a->alpha_dev);
    DMC_LAUNCH_CHECK("splitk_reduce_kernel launch");
  }
  return 0;
}
