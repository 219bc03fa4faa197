// dmc_common.cuh -- shared host/device helpers for libdinomc (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "dinomc.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libdinomc is written for sm_100a (B200) only"
#endif

namespace dmc {

// ---- host side: error reporting -------------------------------------------------------------
void set_error(const char* fmt, ...);                 // api.cu
int cuda_status(cudaError_t e, const char* what);     // api.cu : 0 if ok else (int)e, message recorded

#define DMC_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::dmc::set_error(__VA_ARGS__);      \
      return -1;                          \
    }                                     \
  } while (0)

#define DMC_LAUNCH_CHECK(what)                                         \
  do {                                                                 \
    cudaError_t e__ = cudaGetLastError();                              \
    if (e__ != cudaSuccess) return ::dmc::cuda_status(e__, what);      \
  } while (0)

// ---- host side: kernel launch with programmatic dependent launch (PDL) -----------------------
// Every kernel of the library starts with pdl_prologue(): it lets the NEXT kernel in the stream be scheduled early
// (griddepcontrol.launch_dependents) and then waits for the PREVIOUS kernel to complete and flush
// (griddepcontrol.wait) before it touches global memory.  Launched through launch_kernel() with the programmatic
// stream-serialization attribute, consecutive kernels overlap launch latency, block scheduling and prologues
// (barrier init, TMEM allocation, tensor-map prefetch) with the tail of their predecessor; the data dependencies are
// unchanged because every kernel waits before its first global access.  DMC_PDL=0 turns the attribute off (the
// prologue instructions are then no-ops).  Works under CUDA-graph stream capture (programmatic edges).
bool pdl_enabled();                                   // api.cu
// Grid cap for the row-streaming kernels (weight-norm forward / backward), calling-thread local, 0 = none.  With a cap of
// one CTA per SM (148) such a kernel leaves registers and shared memory for a one-CTA-per-SM GEMM on every SM, so the two
// run side by side; uncapped, its thousands of CTAs keep every SM full and a GEMM queued behind it cannot start.
int streaming_ctas();                                 // api.cu
inline unsigned streaming_grid(int64_t blocks) {
  const int cap = streaming_ctas();
  return static_cast<unsigned>((cap > 0 && blocks > cap) ? cap : blocks);
}
// The split of an SM's 256 KB between L1 and shared memory is per-SM state that can only change while the SM is idle.  The
// tcgen05 GEMM needs the maximum shared-memory split; a streaming kernel launched with the default preference ("more L1")
// flips the SMs it lands on, and a GEMM CTA queued behind it cannot become co-resident -- measured: the last layer's dgrad
// did not start while the 148 CTAs of the exchange kernel were resident (66 -> 229 us at 2 GPUs), weight-norm passes
// blocked the next MLP GEMM.  Kernels that are MEANT to run next to a GEMM (the cross-rank exchange kernel) therefore ask for
// the maximum shared-memory carveout, so the SMs they sit on stay in the GEMM's configuration.  Applied to EVERY kernel it
// costs the streaming kernels their L1 (measured: step 0.774 -> 0.812 ms at 1 GPU), so it is opt-in per kernel.
void prefer_max_smem_carveout(const void* func);      // api.cu
#ifdef __CUDACC__
template <typename... Params, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<Params>(args)...);
}
#endif

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device side ----------------------------------------------------------------------------
#ifdef __CUDACC__

// See launch_kernel(): first statements of every kernel.  No global memory access may precede pdl_prologue().
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_trigger(); pdl_wait(); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 2^x on the SFU as a single MUFU.EX2 (flush-to-zero: no denormal fix-up sequence around it).  The softmax
// kernels work in the base-2 domain: exp(a*x - m) = ex2(fma(x, a*log2e, -m*log2e)) -> one FFMA + one MUFU.
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;

// GELU of nn.GELU() (erf form) and its derivative, through Phi(x) = 1 - erfc(x / sqrt 2) / 2 with erfc from
// Abramowitz-Stegun 7.1.26 (absolute error <= 1.5e-7, i.e. what an fp32 erff evaluation gives: measured 4.2e-7 on
// GELU and 3.0e-7 on GELU' against fp64 over [-12, 12]).  Branch-free: one MUFU.RCP, one MUFU.EX2 and ~12 FMA-pipe
// instructions per value -- erff costs ~2.5x that in the GEMM epilogues, where it sits on the critical path.
// exp(-x^2/2) is shared between Phi and the density term of the derivative.
__device__ __forceinline__ void gelu_parts(float x, float& phi, float& dens) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float q = 1.061405429f;
  q = fmaf(q, t, -1.453152027f);
  q = fmaf(q, t, 1.421413741f);
  q = fmaf(q, t, -0.284496736f);
  q = fmaf(q, t, 0.254829592f);
  dens = ex2((x * -0.72134752044448170f) * x);          // exp(-x^2 / 2)
  const float h = 0.5f * (q * t) * dens;                  // erfc(|x| / sqrt 2) / 2
  phi = (x < 0.f) ? h : 1.0f - h;
}
__device__ __forceinline__ float gelu_f(float x) {
  float phi, dens;
  gelu_parts(x, phi, dens);
  return x * phi;
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float phi, dens;
  gelu_parts(x, phi, dens);
  return fmaf(x * dens, 0.39894228040143268f, phi);
}

// Streaming 128-bit global accesses (data touched once: do not allocate in L1).
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {  // a -> low half, b -> high half (RNE)
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// A "vector" of logits is what one 128-bit load brings: 4 fp32 or 8 bf16.
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void load(const float* p, float (&v)[4]) {
    uint4 r = ld_stream_u4(p);
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  }
  __device__ static __forceinline__ void store(float* p, const float (&v)[4]) {
    st_stream_u4(p, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])));
  }
  __device__ static __forceinline__ float load1(const float* p) { return *p; }
  __device__ static __forceinline__ void store1(float* p, float v) { *p = v; }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 r = ld_stream_u4(p);
    v[0] = bf16_lo(r.x); v[1] = bf16_hi(r.x); v[2] = bf16_lo(r.y); v[3] = bf16_hi(r.y);
    v[4] = bf16_lo(r.z); v[5] = bf16_hi(r.z); v[6] = bf16_lo(r.w); v[7] = bf16_hi(r.w);
  }
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    st_stream_u4(p, make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7])));
  }
  __device__ static __forceinline__ float load1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  __device__ static __forceinline__ void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// Four consecutive logits per thread: one 128-bit (fp32) or 64-bit (bf16) streaming load, kept packed ("raw")
// in registers until the row is processed -- keeps the register footprint of the multi-row loss kernels small.
__device__ __forceinline__ uint2 ld_stream_u2(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_u2(void* p, const uint2& v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
template <typename T> struct Quad;
template <> struct Quad<float> {
  using Raw = uint4;
  __device__ static __forceinline__ Raw load(const float* p) { return ld_stream_u4(p); }
  __device__ static __forceinline__ Raw load_guard(const float* p, long long col, long long K) {
    Raw r;
    r.x = (col + 0 < K) ? __float_as_uint(p[0]) : 0u; r.y = (col + 1 < K) ? __float_as_uint(p[1]) : 0u;
    r.z = (col + 2 < K) ? __float_as_uint(p[2]) : 0u; r.w = (col + 3 < K) ? __float_as_uint(p[3]) : 0u;
    return r;
  }
  __device__ static __forceinline__ void unpack(const Raw& r, float (&v)[4]) {
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  }
  __device__ static __forceinline__ void store(float* p, const float (&v)[4]) {
    st_stream_u4(p, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])));
  }
  __device__ static __forceinline__ void store1(float* p, float v) { *p = v; }
};
template <> struct Quad<__nv_bfloat16> {
  using Raw = uint2;
  __device__ static __forceinline__ Raw load(const __nv_bfloat16* p) { return ld_stream_u2(p); }
  __device__ static __forceinline__ Raw load_guard(const __nv_bfloat16* p, long long col, long long K) {
    const unsigned short* q = reinterpret_cast<const unsigned short*>(p);
    const uint32_t a = (col + 0 < K) ? q[0] : 0u, b = (col + 1 < K) ? q[1] : 0u;
    const uint32_t c = (col + 2 < K) ? q[2] : 0u, d = (col + 3 < K) ? q[3] : 0u;
    return make_uint2(a | (b << 16), c | (d << 16));
  }
  __device__ static __forceinline__ void unpack(const Raw& r, float (&v)[4]) {
    v[0] = bf16_lo(r.x); v[1] = bf16_hi(r.x); v[2] = bf16_lo(r.y); v[3] = bf16_hi(r.y);
  }
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
    st_stream_u2(p, make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3])));
  }
  __device__ static __forceinline__ void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// online softmax statistic merge: (m, l) <- (m, l) (+) (m2, l2), l = sum exp(x - m)
__device__ __forceinline__ void online_merge(float& m, float& l, float m2, float l2) {
  float mn = fmaxf(m, m2);
  float a = (m == -INFINITY) ? 0.f : ex2((m - mn) * kLog2e);
  float b = (m2 == -INFINITY) ? 0.f : ex2((m2 - mn) * kLog2e);
  l = l * a + l2 * b;
  m = mn;
}

// same merge for statistics kept in the base-2 domain: l = sum 2^(y - m)
__device__ __forceinline__ void online_merge2(float& m, float& l, float m2, float l2) {
  float mn = fmaxf(m, m2);
  float a = (m == -INFINITY) ? 0.f : ex2(m - mn);
  float b = (m2 == -INFINITY) ? 0.f : ex2(m2 - mn);
  l = l * a + l2 * b;
  m = mn;
}

#endif  // __CUDACC__
}  // namespace dmc
