// xrank.cu -- the step's cross-rank exchanges as ONE kernel each over NVLink / NVSwitch peer memory.
//
// The data-parallel path has exactly two exchanges: the mean of the head gradients over the ranks (what DDP
// does at main_dino_mc.py:260) and the sum of the K-float teacher column sum (dist.all_reduce at main_dino_mc.py:469).
// Both are all-reduces of buffers that every rank holds at the same symmetric address (torch's symmetric-memory
// allocator maps every rank's buffer into every process and, on NVSwitch, also behind ONE multicast address).
//
// dmc_xrank_allreduce is a two-shot all-reduce written for that memory:
//   barrier (all ranks' producers have finished)                                -- hierarchical: rank-local counter, then one
//                                                                                  signal per rank pair on the signal pads
//   shot 1: rank r reduces slice r of the buffer over all ranks
//             multicast:  ONE multimem.ld_reduce per 16 bytes -- the NVSwitch adds the 8 ranks' values in flight
//             peer-to-peer fallback: W-1 remote 16-byte loads + local adds
//   shot 2: the reduced (and scaled: 1/W for a mean) slice goes to every rank
//             multicast:  ONE multimem.st per 16 bytes -- the switch replicates it
//             fallback:   W-1 remote stores
//   barrier (all slices have landed everywhere)
//   optional epilogue: widen the (now complete) bf16 buffer into fp32 destination tensors, so gradients exchanged in
//             bf16 land in their fp32 .grad tensors without another launch.
// Per GPU and direction this moves ~1x the buffer over NVLink (multicast) instead of the 2(W-1)/W x of a ring, and runs on
// one 256-thread, 32-register CTA per SM: it fits next to a resident tcgen05 GEMM CTA on every SM, so the backward GEMMs
// keep all 148 SMs while the exchange runs (NCCL's kernels needed 16 reserved SMs).
//
// bf16 buffers are summed with fp32 accumulation (multimem .acc::f32 / fp32 adds in the fallback) and rounded once.
#include "dmc_common.cuh"

namespace dmc {
namespace {

constexpr int kMaxWorld = 16;
constexpr int kXThreads = 256;

struct XrankArgs {
  void* mc;                       // multicast address of the buffer (nullptr: peer-to-peer path)
  void* peers[kMaxWorld];         // every rank's buffer, mapped here (peers[rank] = local)
  uint32_t* pads[kMaxWorld];      // every rank's signal pad
  long long n_vec;                // 16-byte vectors in the buffer
  int rank, world;
  float scale;                    // applied to the reduced values (1/world for a mean)
  int is_bf16;
  // optional epilogue: bf16 buffer -> fp32 tensors
  int n_out;
  float* out[8];
  long long out_off[8];           // element offset of tensor i inside the buffer (multiple of 8)
  long long out_n[8];
};

// Signal-pad protocol (same as torch's symmetric-memory barrier, so both can share a pad): slot value 0 = empty.
// put: CAS 0 -> 1 at the PEER's slot (spins while the previous signal has not been consumed); wait: CAS 1 -> 0 at OWN slot.
__device__ __forceinline__ void put_signal(uint32_t* addr) {
  uint32_t old;
  for (;;) {
    asm volatile("atom.global.release.sys.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(addr) : "memory");
    if (old == 0u) break;
    __nanosleep(128);                               // the previous signal has not been consumed yet: back off, do not hammer NVLink
  }
}
__device__ __forceinline__ void wait_signal(uint32_t* addr) {
  // Poll with a plain acquire LOAD and back off between polls: 148 CTAs x `world` threads spinning on system-scope
  // atomics measurably slowed the GEMM sharing the SMs (dgrad 66 -> 240 us at 2 GPUs).  Only this thread consumes the
  // slot, so once the signal is seen a plain store resets it.
  long long spins = 0;
  for (;;) {
    uint32_t v;
    asm volatile("ld.global.acquire.sys.b32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    if (v == 1u) break;
    __nanosleep(256);
    if (++spins > (1ll << 26)) __trap();            // ~20 s: a peer never arrived -- fail loudly instead of hanging the box
  }
  asm volatile("st.global.relaxed.sys.b32 [%0], %1;" :: "l"(addr), "r"(0u) : "memory");
}

// Barrier over ALL CTAs of ALL ranks, hierarchical: the CTAs of a rank meet at a rank-local counter (cheap local atomics);
// the LAST one to arrive exchanges ONE signal with every peer (threads 0..world-1 in parallel: put at the peer's pad, wait at
// the own pad) and then releases its rank's CTAs through a generation word.  Remote atomics per barrier and rank: `world`
// instead of `ctas x world` -- with one signal per CTA pair the 1.6 MB exchange at the end of the step took 47-67 us at 8 GPUs
// (1184 remote CAS into each GPU's signal pad per barrier), i.e. it was all barrier.  Needs every CTA of the kernel resident
// (<= 148 CTAs of 256 threads / 32 registers: true next to a GEMM as well).  Self-resetting, reusable launch after launch.
// Pad layout (32-bit words): [channel 0..1][sender rank 0..world-1] signals, then [channel 0..1]{counter, generation}.
__device__ __forceinline__ void rank_barrier(const XrankArgs& a, int channel) {
  __shared__ uint32_t s_last;
  uint32_t* own = a.pads[a.rank];
  uint32_t* ctr = own + 2 * a.world + 2 * channel;
  uint32_t* gen = ctr + 1;
  __syncthreads();
  uint32_t g0 = 0;
  if (threadIdx.x == 0) {
    uint32_t old;
    asm volatile("ld.global.acquire.gpu.b32 %0, [%1];" : "=r"(g0) : "l"(gen) : "memory");
    asm volatile("atom.global.acq_rel.gpu.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(ctr) : "memory");
    s_last = (old == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {                                   // uniform per CTA
    if (threadIdx.x < static_cast<unsigned>(a.world)) {
      const int peer = threadIdx.x;
      put_signal(a.pads[peer] + channel * a.world + a.rank);     // this rank has arrived
      wait_signal(own + channel * a.world + peer);               // ... and so has `peer`
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("st.global.relaxed.gpu.b32 [%0], %1;" :: "l"(ctr), "r"(0u) : "memory");
      asm volatile("st.global.release.gpu.b32 [%0], %1;" :: "l"(gen), "r"(g0 + 1u) : "memory");
    }
  } else if (threadIdx.x == 0) {
    uint32_t g1;
    long long spins = 0;
    do {
      __nanosleep(64);
      asm volatile("ld.global.acquire.gpu.b32 %0, [%1];" : "=r"(g1) : "l"(gen) : "memory");
      if (++spins > (1ll << 27)) __trap();        // a rank never arrived: fail loudly instead of hanging the box
    } while (g1 == g0);
  }
  __syncthreads();
}

__device__ __forceinline__ uint4 mc_ld_reduce_bf16(const void* p) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ uint4 mc_ld_reduce_f32(const void* p) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void mc_st(void* p, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_sys(const void* p) {
  uint4 r;
  asm volatile("ld.global.relaxed.sys.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_sys(void* p, const uint4& v) {
  asm volatile("st.global.relaxed.sys.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint4 scale_vec(const uint4& v, float s, bool bf16) {
  if (s == 1.0f) return v;
  uint4 r;
  if (bf16) {
    r.x = pack_bf16(bf16_lo(v.x) * s, bf16_hi(v.x) * s); r.y = pack_bf16(bf16_lo(v.y) * s, bf16_hi(v.y) * s);
    r.z = pack_bf16(bf16_lo(v.z) * s, bf16_hi(v.z) * s); r.w = pack_bf16(bf16_lo(v.w) * s, bf16_hi(v.w) * s);
  } else {
    r.x = __float_as_uint(__uint_as_float(v.x) * s); r.y = __float_as_uint(__uint_as_float(v.y) * s);
    r.z = __float_as_uint(__uint_as_float(v.z) * s); r.w = __float_as_uint(__uint_as_float(v.w) * s);
  }
  return r;
}

__global__ void __launch_bounds__(kXThreads, 8)   // <= 32 registers: 8 warps x 1024 regs fit beside a GEMM CTA (10 warps x 5632: the register file is allocated in 512-register units per warp)
xrank_allreduce_kernel(const XrankArgs a) {
  pdl_prologue();
  const bool bf16 = a.is_bf16 != 0;
  rank_barrier(a, 0);                        // every rank's producer kernels precede this kernel in its stream: data is final
  // slice of this rank, in 16-byte vectors
  const long long per = (a.n_vec + a.world - 1) / a.world;
  const long long v0 = min(per * a.rank, a.n_vec), v1 = min(v0 + per, a.n_vec);
  const long long stride = static_cast<long long>(gridDim.x) * kXThreads;
  const long long first = v0 + static_cast<long long>(blockIdx.x) * kXThreads + threadIdx.x;
  if (a.mc != nullptr) {
    uint8_t* mc = static_cast<uint8_t*>(a.mc);
    constexpr int U = 4;                     // 4 x 16 bytes in flight per thread
    for (long long i = first; i < v1; i += U * stride) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long j = i + u * stride;
        if (j < v1) v[u] = bf16 ? mc_ld_reduce_bf16(mc + 16 * j) : mc_ld_reduce_f32(mc + 16 * j);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long j = i + u * stride;
        if (j < v1) mc_st(mc + 16 * j, scale_vec(v[u], a.scale, bf16));
      }
    }
  } else {
    for (long long i = first; i < v1; i += stride) {
      float acc[8];
      const int ne = bf16 ? 8 : 4;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
      for (int p = 0; p < a.world; ++p) {     // fixed rank order: every rank would compute the same bits for a slice
        const uint4 x = ld_sys(static_cast<const uint8_t*>(a.peers[p]) + 16 * i);
        if (bf16) {
          acc[0] += bf16_lo(x.x); acc[1] += bf16_hi(x.x); acc[2] += bf16_lo(x.y); acc[3] += bf16_hi(x.y);
          acc[4] += bf16_lo(x.z); acc[5] += bf16_hi(x.z); acc[6] += bf16_lo(x.w); acc[7] += bf16_hi(x.w);
        } else {
          acc[0] += __uint_as_float(x.x); acc[1] += __uint_as_float(x.y); acc[2] += __uint_as_float(x.z); acc[3] += __uint_as_float(x.w);
        }
      }
      for (int e = 0; e < ne; ++e) acc[e] *= a.scale;
      uint4 r;
      if (bf16) r = make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
      else r = make_uint4(__float_as_uint(acc[0]), __float_as_uint(acc[1]), __float_as_uint(acc[2]), __float_as_uint(acc[3]));
      for (int p = 0; p < a.world; ++p) st_sys(static_cast<uint8_t*>(a.peers[p]) + 16 * i, r);
    }
  }
  __threadfence_system();                    // this CTA's remote stores are visible before it signals
  rank_barrier(a, 1);                        // every CTA of every rank has finished: all slices have landed in every buffer
  // epilogue: widen the (now complete) bf16 buffer into fp32 tensors (gradients exchanged in bf16 -> fp32 .grad)
  if (a.n_out > 0) {
    const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(a.peers[a.rank]);
    for (int t = 0; t < a.n_out; ++t) {
      const __nv_bfloat16* s = src + a.out_off[t];
      float* d = a.out[t];
      const long long n8 = a.out_n[t] >> 3;
      const bool al = (reinterpret_cast<uintptr_t>(d) & 15) == 0;
      for (long long i = static_cast<long long>(blockIdx.x) * kXThreads + threadIdx.x; i < n8; i += stride) {
        const uint4 x = *reinterpret_cast<const uint4*>(s + 8 * i);      // plain (coherent) load: written by peers before the barrier
        if (al) {
          *reinterpret_cast<float4*>(d + 8 * i) = make_float4(bf16_lo(x.x), bf16_hi(x.x), bf16_lo(x.y), bf16_hi(x.y));
          *reinterpret_cast<float4*>(d + 8 * i + 4) = make_float4(bf16_lo(x.z), bf16_hi(x.z), bf16_lo(x.w), bf16_hi(x.w));
        } else {
          d[8 * i] = bf16_lo(x.x); d[8 * i + 1] = bf16_hi(x.x); d[8 * i + 2] = bf16_lo(x.y); d[8 * i + 3] = bf16_hi(x.y);
          d[8 * i + 4] = bf16_lo(x.z); d[8 * i + 5] = bf16_hi(x.z); d[8 * i + 6] = bf16_lo(x.w); d[8 * i + 7] = bf16_hi(x.w);
        }
      }
      for (long long i = (n8 << 3) + static_cast<long long>(blockIdx.x) * kXThreads + threadIdx.x; i < a.out_n[t]; i += stride)
        d[i] = __bfloat162float(s[i]);
    }
  }
}

}  // namespace
}  // namespace dmc

using namespace dmc;

extern "C" size_t dmc_xrank_signal_bytes(int32_t world, int32_t ctas) {
  if (world <= 0 || ctas <= 0) return 0;
  (void)ctas;
  return (static_cast<size_t>(2) * world + 4) * sizeof(uint32_t);      // two channels of per-rank signals + {counter, generation} each
}

extern "C" int dmc_xrank_allreduce(void* multicast_ptr, void* const* peer_ptrs_host, void* const* signal_pads_host, int64_t numel,
                                   int32_t dtype, int32_t rank, int32_t world, float scale, int32_t ctas, int32_t n_out,
                                   float* const* out_ptrs_host, const int64_t* out_offsets_host, const int64_t* out_numels_host,
                                   void* stream) {
  DMC_REQUIRE(peer_ptrs_host && signal_pads_host, "dmc_xrank_allreduce: null pointer table");
  DMC_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "dmc_xrank_allreduce: bad rank %d / world %d", rank, world);
  DMC_REQUIRE(dtype == DMC_F32 || dtype == DMC_BF16, "dmc_xrank_allreduce: bad dtype %d", dtype);
  const int esz = dtype == DMC_BF16 ? 2 : 4;
  DMC_REQUIRE(numel > 0 && (numel * esz) % 16 == 0, "dmc_xrank_allreduce: the buffer must be a multiple of 16 bytes (numel=%lld)", (long long)numel);
  DMC_REQUIRE(ctas > 0 && ctas <= 1024, "dmc_xrank_allreduce: bad CTA count %d", ctas);
  DMC_REQUIRE(n_out >= 0 && n_out <= 8 && (n_out == 0 || (dtype == DMC_BF16 && out_ptrs_host && out_offsets_host && out_numels_host)),
              "dmc_xrank_allreduce: the widening epilogue takes <= 8 tensors and a bf16 buffer");
  XrankArgs a{};
  a.mc = multicast_ptr;
  for (int p = 0; p < world; ++p) {
    DMC_REQUIRE(peer_ptrs_host[p] && signal_pads_host[p], "dmc_xrank_allreduce: null peer pointer %d", p);
    DMC_REQUIRE((reinterpret_cast<uintptr_t>(peer_ptrs_host[p]) & 15) == 0, "dmc_xrank_allreduce: peer buffer %d is not 16-byte aligned", p);
    a.peers[p] = peer_ptrs_host[p];
    a.pads[p] = static_cast<uint32_t*>(signal_pads_host[p]);
  }
  a.n_vec = numel * esz / 16;
  a.rank = rank; a.world = world; a.scale = scale; a.is_bf16 = dtype == DMC_BF16;
  a.n_out = n_out;
  for (int t = 0; t < n_out; ++t) {
    DMC_REQUIRE(out_ptrs_host[t] && out_offsets_host[t] >= 0 && out_offsets_host[t] % 8 == 0 && out_numels_host[t] >= 0 &&
                out_offsets_host[t] + out_numels_host[t] <= numel, "dmc_xrank_allreduce: bad output slice %d", t);
    a.out[t] = out_ptrs_host[t]; a.out_off[t] = out_offsets_host[t]; a.out_n[t] = out_numels_host[t];
  }
  prefer_max_smem_carveout(reinterpret_cast<const void*>(xrank_allreduce_kernel));    // stay co-resident with the tcgen05 GEMMs
  launch_kernel(xrank_allreduce_kernel, dim3(static_cast<unsigned>(ctas)), dim3(kXThreads), 0, static_cast<cudaStream_t>(stream), a);
  DMC_LAUNCH_CHECK("xrank_allreduce_kernel launch");
  return 0;
}
