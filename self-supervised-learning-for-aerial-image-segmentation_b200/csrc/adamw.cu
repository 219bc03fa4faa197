// adamw.cu -- torch.optim.AdamW (main_dino_mc.py:281-282, stepped at :391 / :399) as ONE multi-tensor launch per
// parameter group instead of ~6 elementwise kernels per parameter (or the foreach equivalents).
//
// Per element, in the order of torch's _single_tensor_adam with decoupled weight decay (fp32 throughout, scalars
// formed in float64 on the host like the Python floats they are in torch):
//     p   <- p * (1 - lr * wd)
//     m   <- m + (1 - beta1) * (g - m)                      exp_avg.lerp_(grad, 1 - beta1)
//     v   <- v * beta2 ;  v <- v + (1 - beta2) * g * g      exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
//     den <- sqrt(v) / sqrt(1 - beta2^t) + eps
//     p   <- p + (-(lr / (1 - beta1^t))) * (m / den)        param.addcdiv_(exp_avg, denom, value=-step_size)
// A host-built plan (one entry per <= 16384-element chunk of a (param, grad, exp_avg, exp_avg_sq) quadruple) drives
// the kernel; 128-bit accesses when all four pointers are 16-byte aligned.
// HBM-bound: 16 bytes read + 12 written per parameter.
#include <math.h>

#include "dmc_common.cuh"

namespace dmc {
namespace {

constexpr long long kAdamChunk = 16384;

struct AdamChunk {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
};

struct AdamScalars {
  float decay;        // 1 - lr * wd
  float w1;           // 1 - beta1
  float beta2, w2;    // beta2, 1 - beta2
  float bc2_sqrt;     // sqrt(1 - beta2^t)
  float eps;
  float neg_step;     // -(lr / (1 - beta1^t))
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamScalars& s) {
  p = __fmul_rn(p, s.decay);
  m = fmaf(s.w1, __fsub_rn(g, m), m);
  v = __fmul_rn(v, s.beta2);
  v = fmaf(__fmul_rn(s.w2, g), g, v);
  const float den = __fadd_rn(__fdiv_rn(sqrtf(v), s.bc2_sqrt), s.eps);
  p = fmaf(s.neg_step, __fdiv_rn(m, den), p);
}

__global__ void __launch_bounds__(256)
adamw_kernel(const AdamChunk* __restrict__ plan, const AdamScalars s) {
  pdl_prologue();
  const AdamChunk c = plan[blockIdx.x];
  const bool vec = (((reinterpret_cast<uintptr_t>(c.p) | reinterpret_cast<uintptr_t>(c.g) | reinterpret_cast<uintptr_t>(c.m) |
                      reinterpret_cast<uintptr_t>(c.v)) & 15) == 0);
  long long done = 0;
  if (vec) {
    const long long n4 = c.n >> 2;
    for (long long i = threadIdx.x; i < n4; i += 256) {
      float4 p = *reinterpret_cast<float4*>(c.p + 4 * i);
      const float4 g = *reinterpret_cast<const float4*>(c.g + 4 * i);
      float4 m = *reinterpret_cast<float4*>(c.m + 4 * i);
      float4 v = *reinterpret_cast<float4*>(c.v + 4 * i);
      adam_one(p.x, g.x, m.x, v.x, s); adam_one(p.y, g.y, m.y, v.y, s);
      adam_one(p.z, g.z, m.z, v.z, s); adam_one(p.w, g.w, m.w, v.w, s);
      *reinterpret_cast<float4*>(c.p + 4 * i) = p;
      *reinterpret_cast<float4*>(c.m + 4 * i) = m;
      *reinterpret_cast<float4*>(c.v + 4 * i) = v;
    }
    done = n4 << 2;
  }
  for (long long i = done + threadIdx.x; i < c.n; i += 256) {
    float p = c.p[i], m = c.m[i], v = c.v[i];
    adam_one(p, c.g[i], m, v, s);
    c.p[i] = p; c.m[i] = m; c.v[i] = v;
  }
}

}  // namespace
}  // namespace dmc

using namespace dmc;

extern "C" size_t dmc_adamw_plan_bytes(const int64_t* numels_host, int64_t n_tensors) {
  if (!numels_host || n_tensors <= 0) return 0;
  size_t chunks = 0;
  for (int64_t i = 0; i < n_tensors; ++i)
    if (numels_host[i] > 0) chunks += static_cast<size_t>(ceil_div(numels_host[i], kAdamChunk));
  return chunks * sizeof(AdamChunk);
}

extern "C" int dmc_adamw_build_plan(const void* const* param_ptrs_host, const void* const* grad_ptrs_host,
                                    const void* const* exp_avg_ptrs_host, const void* const* exp_avg_sq_ptrs_host,
                                    const int64_t* numels_host, int64_t n_tensors, void* plan_host, size_t plan_bytes,
                                    int64_t* n_chunks_out) {
  DMC_REQUIRE(param_ptrs_host && grad_ptrs_host && exp_avg_ptrs_host && exp_avg_sq_ptrs_host && numels_host && plan_host && n_chunks_out,
              "dmc_adamw_build_plan: null pointer");
  DMC_REQUIRE(n_tensors > 0, "dmc_adamw_build_plan: no tensors");
  DMC_REQUIRE(plan_bytes >= dmc_adamw_plan_bytes(numels_host, n_tensors), "dmc_adamw_build_plan: plan buffer too small");
  AdamChunk* out = static_cast<AdamChunk*>(plan_host);
  int64_t n = 0;
  for (int64_t i = 0; i < n_tensors; ++i) {
    DMC_REQUIRE(numels_host[i] >= 0, "dmc_adamw_build_plan: negative numel at %lld", (long long)i);
    const void* ptrs[4] = {param_ptrs_host[i], grad_ptrs_host[i], exp_avg_ptrs_host[i], exp_avg_sq_ptrs_host[i]};
    for (const void* q : ptrs) {
      DMC_REQUIRE(numels_host[i] == 0 || q != nullptr, "dmc_adamw_build_plan: null tensor at %lld", (long long)i);
      DMC_REQUIRE((reinterpret_cast<uintptr_t>(q) & 3) == 0, "dmc_adamw_build_plan: tensor %lld is not 4-byte aligned", (long long)i);
    }
    for (int64_t off = 0; off < numels_host[i]; off += kAdamChunk) {
      out[n].p = const_cast<float*>(static_cast<const float*>(ptrs[0])) + off;
      out[n].g = static_cast<const float*>(ptrs[1]) + off;
      out[n].m = const_cast<float*>(static_cast<const float*>(ptrs[2])) + off;
      out[n].v = const_cast<float*>(static_cast<const float*>(ptrs[3])) + off;
      out[n].n = (numels_host[i] - off < kAdamChunk) ? (numels_host[i] - off) : kAdamChunk;
      ++n;
    }
  }
  *n_chunks_out = n;
  return 0;
}

extern "C" int dmc_adamw_multi_tensor(const void* plan_dev, int64_t n_chunks, double lr, double beta1, double beta2, double eps,
                                      double weight_decay, int64_t step, void* stream) {
  DMC_REQUIRE(plan_dev && n_chunks > 0 && n_chunks < (1ll << 31), "dmc_adamw_multi_tensor: bad plan");
  DMC_REQUIRE(step >= 1, "dmc_adamw_multi_tensor: step must be >= 1 (got %lld)", (long long)step);
  DMC_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0, "dmc_adamw_multi_tensor: bad hyper-parameters");
  AdamScalars s{};
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
  s.decay = static_cast<float>(1.0 - lr * weight_decay);
  s.w1 = static_cast<float>(1.0 - beta1);
  s.beta2 = static_cast<float>(beta2);
  s.w2 = static_cast<float>(1.0 - beta2);
  s.bc2_sqrt = static_cast<float>(sqrt(bc2));
  s.eps = static_cast<float>(eps);
  s.neg_step = static_cast<float>(-(lr / bc1));
  launch_kernel(adamw_kernel, dim3(static_cast<unsigned>(n_chunks)), dim3(256), 0, static_cast<cudaStream_t>(stream),
                static_cast<const AdamChunk*>(plan_dev), s);
  DMC_LAUNCH_CHECK("adamw_kernel launch");
  return 0;
}
