// ema.cu -- EMA teacher update (main_dino_mc.py:403-406) as ONE multi-tensor launch.
//
// The reference issues mul_, scalar*tensor and add_ per parameter (3 launches x ~160 tensors per step).
// Here a host-built "plan" (one entry per <= 16384-element chunk of a (teacher, student) pair) drives a
// single kernel: one CTA per chunk, 128-bit coalesced loads/stores when both pointers are 16-byte
// aligned.  Arithmetic reproduces the reference's three fp32 roundings exactly:
//     p_k <- fl( fl(p_k * m) + fl((1-m) * p_q) ),   m and (1-m) cast to fp32 from float64.
// HBM-bound: 12 bytes per parameter (read student, read teacher, write teacher).
#include "dmc_common.cuh"

namespace dmc {
namespace {

constexpr long long kChunkElems = 16384;

struct EmaChunk {
  float* teacher;
  const float* student;
  long long n;
};

__global__ void __launch_bounds__(256)
ema_kernel(const EmaChunk* __restrict__ plan, float m, float omm) {
  pdl_prologue();
  const EmaChunk c = plan[blockIdx.x];
  float* __restrict__ pk = c.teacher;
  const float* __restrict__ pq = c.student;
  const bool vec = (((reinterpret_cast<uintptr_t>(pk) | reinterpret_cast<uintptr_t>(pq)) & 15) == 0);
  long long done = 0;
  if (vec) {
    const long long n4 = c.n >> 2;
    for (long long i = threadIdx.x; i < n4; i += 256) {
      const uint4 a = ld_stream_u4(pk + 4 * i);           // teacher (read once, rewritten)
      const uint4 b = ld_stream_u4(pq + 4 * i);           // student
      uint4 r;
      r.x = __float_as_uint(__fadd_rn(__fmul_rn(__uint_as_float(a.x), m), __fmul_rn(omm, __uint_as_float(b.x))));
      r.y = __float_as_uint(__fadd_rn(__fmul_rn(__uint_as_float(a.y), m), __fmul_rn(omm, __uint_as_float(b.y))));
      r.z = __float_as_uint(__fadd_rn(__fmul_rn(__uint_as_float(a.z), m), __fmul_rn(omm, __uint_as_float(b.z))));
      r.w = __float_as_uint(__fadd_rn(__fmul_rn(__uint_as_float(a.w), m), __fmul_rn(omm, __uint_as_float(b.w))));
      st_stream_u4(pk + 4 * i, r);
    }
    done = n4 << 2;
  }
  for (long long i = done + threadIdx.x; i < c.n; i += 256)
    pk[i] = __fadd_rn(__fmul_rn(pk[i], m), __fmul_rn(omm, pq[i]));
}

}  // namespace
}  // namespace dmc

using namespace dmc;

extern "C" size_t dmc_ema_plan_bytes(const int64_t* numels_host, int64_t n_tensors) {
  if (!numels_host || n_tensors <= 0) return 0;
  size_t chunks = 0;
  for (int64_t i = 0; i < n_tensors; ++i)
    if (numels_host[i] > 0) chunks += static_cast<size_t>(ceil_div(numels_host[i], kChunkElems));
  return chunks * sizeof(EmaChunk);
}

extern "C" int dmc_ema_build_plan(const void* const* teacher_ptrs_host, const void* const* student_ptrs_host,
                                  const int64_t* numels_host, int64_t n_tensors, void* plan_host, size_t plan_bytes,
                                  int64_t* n_chunks_out) {
  DMC_REQUIRE(teacher_ptrs_host && student_ptrs_host && numels_host && plan_host && n_chunks_out, "dmc_ema_build_plan: null pointer");
  DMC_REQUIRE(n_tensors > 0, "dmc_ema_build_plan: no tensors");
  DMC_REQUIRE(plan_bytes >= dmc_ema_plan_bytes(numels_host, n_tensors), "dmc_ema_build_plan: plan buffer too small");
  EmaChunk* out = static_cast<EmaChunk*>(plan_host);
  int64_t n = 0;
  for (int64_t i = 0; i < n_tensors; ++i) {
    DMC_REQUIRE(numels_host[i] >= 0, "dmc_ema_build_plan: negative numel at %lld", (long long)i);
    DMC_REQUIRE(numels_host[i] == 0 || (teacher_ptrs_host[i] && student_ptrs_host[i]), "dmc_ema_build_plan: null tensor at %lld", (long long)i);
    DMC_REQUIRE((reinterpret_cast<uintptr_t>(teacher_ptrs_host[i]) & 3) == 0 && (reinterpret_cast<uintptr_t>(student_ptrs_host[i]) & 3) == 0,
                "dmc_ema_build_plan: tensor %lld is not 4-byte aligned", (long long)i);
    for (int64_t off = 0; off < numels_host[i]; off += kChunkElems) {
      out[n].teacher = const_cast<float*>(static_cast<const float*>(teacher_ptrs_host[i])) + off;
      out[n].student = static_cast<const float*>(student_ptrs_host[i]) + off;
      out[n].n = (numels_host[i] - off < kChunkElems) ? (numels_host[i] - off) : kChunkElems;
      ++n;
    }
  }
  *n_chunks_out = n;
  return 0;
}

extern "C" int dmc_ema_multi_tensor(const void* plan_dev, int64_t n_chunks, float m, float one_minus_m, void* stream) {
  DMC_REQUIRE(plan_dev && n_chunks > 0 && n_chunks < (1ll << 31), "dmc_ema_multi_tensor: bad plan");
  launch_kernel(ema_kernel, dim3(static_cast<unsigned>(n_chunks)), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const EmaChunk*>(plan_dev), m, one_minus_m);
  DMC_LAUNCH_CHECK("ema_kernel launch");
  return 0;
}
