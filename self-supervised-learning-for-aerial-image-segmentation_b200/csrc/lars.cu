// lars.cu -- the reference's LARS optimizer step (utils/utils.py:570-608, selected by `--optimizer lars`,
// main_dino_mc.py:285-286) for one parameter group as TWO multi-tensor launches instead of, per parameter, two norm
// kernels, three `where`s, and five elementwise kernels (~12 launches x ~160 parameters per step).
//
//   for every parameter p with gradient g (fp32):
//     adapt = (p.ndim != 1)                                   biases / norm scales are neither decayed nor adapted
//     d  = adapt ? g + wd * p : g
//     q  = adapt && ||p|| > 0 && ||d|| > 0 ? eta * ||p|| / ||d|| : 1
//     mu = mu * momentum + d * q
//     p  = p - lr * mu
//
// A host-built plan (one entry per <= 16384-element chunk plus the chunk range of its tensor, like clip.cu) drives
//   pass 1  one CTA per chunk of an adapted tensor: partial sums of p^2 and (g + wd p)^2 -> workspace [n_chunks][2]
//           (fixed-order block reduction);
//   pass 2  one CTA per chunk: re-adds its tensor's partials in index order (every CTA of a tensor forms the identical
//           fp32 trust ratio), then updates mu and p of its chunk.
// Deterministic (no atomics), no host synchronisation.  HBM-bound: 8 B read per element in pass 1 (adapted tensors),
// 12 B read (p, g mostly from L2) + 8 B written in pass 2.
#include "dmc_common.cuh"

namespace dmc {
namespace {

constexpr long long kLarsChunk = 16384;

struct LarsChunk {
  float* p;
  const float* g;
  float* mu;
  long long n;
  int first, count;     // chunk range [first, first + count) of this tensor in the plan
  int adapt;            // p.ndim != 1
  int pad;
};

__device__ __forceinline__ float2 block_sum2_256(float a, float b, float (*red)[2]) {
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = a; red[threadIdx.x >> 5][1] = b; }
  __syncthreads();
  float ta = 0.f, tb = 0.f;
  if (threadIdx.x < 32) {
    ta = (threadIdx.x < 8) ? red[threadIdx.x][0] : 0.f;
    tb = (threadIdx.x < 8) ? red[threadIdx.x][1] : 0.f;
    ta = warp_sum(ta);
    tb = warp_sum(tb);
  }
  return make_float2(ta, tb);       // valid in warp 0
}

__global__ void __launch_bounds__(256)
lars_norms_kernel(const LarsChunk* __restrict__ plan, float wd, float* __restrict__ partial) {
  pdl_prologue();
  __shared__ float red[8][2];
  const LarsChunk c = plan[blockIdx.x];
  if (!c.adapt) return;
  const float* __restrict__ p = c.p;
  const float* __restrict__ g = c.g;
  float sp = 0.f, sd = 0.f;
  long long done = 0;
  if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g)) & 15) == 0) {
    const long long n4 = c.n >> 2;
    for (long long i = threadIdx.x; i < n4; i += 256) {
      const float4 a = *reinterpret_cast<const float4*>(p + 4 * i);
      const float4 b = *reinterpret_cast<const float4*>(g + 4 * i);
      const float d0 = fmaf(wd, a.x, b.x), d1 = fmaf(wd, a.y, b.y), d2 = fmaf(wd, a.z, b.z), d3 = fmaf(wd, a.w, b.w);
      sp = fmaf(a.x, a.x, sp); sp = fmaf(a.y, a.y, sp); sp = fmaf(a.z, a.z, sp); sp = fmaf(a.w, a.w, sp);
      sd = fmaf(d0, d0, sd); sd = fmaf(d1, d1, sd); sd = fmaf(d2, d2, sd); sd = fmaf(d3, d3, sd);
    }
    done = n4 << 2;
  }
  for (long long i = done + threadIdx.x; i < c.n; i += 256) {
    const float a = p[i], d = fmaf(wd, a, g[i]);
    sp = fmaf(a, a, sp);
    sd = fmaf(d, d, sd);
  }
  const float2 t = block_sum2_256(sp, sd, red);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = t.x; partial[2 * blockIdx.x + 1] = t.y; }
}

__device__ __forceinline__ void lars_one(float& p, float g, float& mu, float wd, float q, float momentum, float neg_lr, bool adapt) {
  float d = g;
  if (adapt) d = __fmul_rn(fmaf(wd, p, g), q);               // dp.add(p, alpha=wd).mul(q)
  mu = __fadd_rn(__fmul_rn(mu, momentum), d);                // mu.mul_(momentum).add_(dp)
  p = fmaf(neg_lr, mu, p);                                   // p.add_(mu, alpha=-lr)
}

__global__ void __launch_bounds__(256)
lars_update_kernel(const LarsChunk* __restrict__ plan, const float* __restrict__ partial, float wd, float eta, float momentum,
                   float neg_lr) {
  pdl_prologue();
  __shared__ float red[8][2];
  __shared__ float q_s;
  const LarsChunk c = plan[blockIdx.x];
  const bool adapt = c.adapt != 0;
  float q = 1.f;
  if (adapt) {
    float sp = 0.f, sd = 0.f;
    for (int i = threadIdx.x; i < c.count; i += 256) { sp += partial[2 * (c.first + i)]; sd += partial[2 * (c.first + i) + 1]; }
    const float2 t = block_sum2_256(sp, sd, red);
    if (threadIdx.x == 0) {
      const float pn = sqrtf(t.x), un = sqrtf(t.y);
      q_s = (pn > 0.f && un > 0.f) ? __fdiv_rn(__fmul_rn(eta, pn), un) : 1.f;     // eta * param_norm / update_norm
    }
    __syncthreads();
    q = q_s;
  }
  long long done = 0;
  if (((reinterpret_cast<uintptr_t>(c.p) | reinterpret_cast<uintptr_t>(c.g) | reinterpret_cast<uintptr_t>(c.mu)) & 15) == 0) {
    const long long n4 = c.n >> 2;
    for (long long i = threadIdx.x; i < n4; i += 256) {
      float4 p = *reinterpret_cast<float4*>(c.p + 4 * i);
      const float4 g = *reinterpret_cast<const float4*>(c.g + 4 * i);
      float4 m = *reinterpret_cast<float4*>(c.mu + 4 * i);
      lars_one(p.x, g.x, m.x, wd, q, momentum, neg_lr, adapt); lars_one(p.y, g.y, m.y, wd, q, momentum, neg_lr, adapt);
      lars_one(p.z, g.z, m.z, wd, q, momentum, neg_lr, adapt); lars_one(p.w, g.w, m.w, wd, q, momentum, neg_lr, adapt);
      *reinterpret_cast<float4*>(c.p + 4 * i) = p;
      *reinterpret_cast<float4*>(c.mu + 4 * i) = m;
    }
    done = n4 << 2;
  }
  for (long long i = done + threadIdx.x; i < c.n; i += 256) {
    float p = c.p[i], m = c.mu[i];
    lars_one(p, c.g[i], m, wd, q, momentum, neg_lr, adapt);
    c.p[i] = p; c.mu[i] = m;
  }
}

}  // namespace
}  // namespace dmc

using namespace dmc;

extern "C" size_t dmc_lars_plan_bytes(const int64_t* numels_host, int64_t n_tensors) {
  if (!numels_host || n_tensors <= 0) return 0;
  size_t chunks = 0;
  for (int64_t i = 0; i < n_tensors; ++i)
    if (numels_host[i] > 0) chunks += static_cast<size_t>(ceil_div(numels_host[i], kLarsChunk));
  return chunks * sizeof(LarsChunk);
}

extern "C" int dmc_lars_build_plan(const void* const* param_ptrs_host, const void* const* grad_ptrs_host,
                                   const void* const* mu_ptrs_host, const int64_t* numels_host, const int32_t* adapt_host,
                                   int64_t n_tensors, void* plan_host, size_t plan_bytes, int64_t* n_chunks_out) {
  DMC_REQUIRE(param_ptrs_host && grad_ptrs_host && mu_ptrs_host && numels_host && adapt_host && plan_host && n_chunks_out,
              "dmc_lars_build_plan: null pointer");
  DMC_REQUIRE(n_tensors > 0 && n_tensors < (1ll << 31), "dmc_lars_build_plan: bad tensor count");
  DMC_REQUIRE(plan_bytes >= dmc_lars_plan_bytes(numels_host, n_tensors), "dmc_lars_build_plan: plan buffer too small");
  LarsChunk* out = static_cast<LarsChunk*>(plan_host);
  int64_t n = 0;
  for (int64_t i = 0; i < n_tensors; ++i) {
    DMC_REQUIRE(numels_host[i] >= 0, "dmc_lars_build_plan: negative numel at %lld", (long long)i);
    const void* ptrs[3] = {param_ptrs_host[i], grad_ptrs_host[i], mu_ptrs_host[i]};
    for (const void* q : ptrs) {
      DMC_REQUIRE(numels_host[i] == 0 || q != nullptr, "dmc_lars_build_plan: null tensor at %lld", (long long)i);
      DMC_REQUIRE((reinterpret_cast<uintptr_t>(q) & 3) == 0, "dmc_lars_build_plan: tensor %lld is not 4-byte aligned", (long long)i);
    }
    const int64_t count = ceil_div(numels_host[i], kLarsChunk);
    DMC_REQUIRE(n + count < (1ll << 31), "dmc_lars_build_plan: too many chunks");
    const int64_t first = n;
    for (int64_t off = 0; off < numels_host[i]; off += kLarsChunk) {
      out[n].p = const_cast<float*>(static_cast<const float*>(ptrs[0])) + off;
      out[n].g = static_cast<const float*>(ptrs[1]) + off;
      out[n].mu = const_cast<float*>(static_cast<const float*>(ptrs[2])) + off;
      out[n].n = (numels_host[i] - off < kLarsChunk) ? (numels_host[i] - off) : kLarsChunk;
      out[n].first = static_cast<int>(first);
      out[n].count = static_cast<int>(count);
      out[n].adapt = adapt_host[i] ? 1 : 0;
      out[n].pad = 0;
      ++n;
    }
  }
  *n_chunks_out = n;
  return 0;
}

extern "C" int dmc_lars_multi_tensor(const void* plan_dev, int64_t n_chunks, double lr, double weight_decay, double momentum,
                                     double eta, float* workspace, size_t workspace_bytes, void* stream) {
  DMC_REQUIRE(plan_dev && workspace, "dmc_lars_multi_tensor: null pointer");
  DMC_REQUIRE(n_chunks > 0 && n_chunks < (1ll << 31), "dmc_lars_multi_tensor: bad plan");
  DMC_REQUIRE(workspace_bytes >= static_cast<size_t>(n_chunks) * 2 * sizeof(float), "dmc_lars_multi_tensor: workspace too small (%zu < %zu)",
              workspace_bytes, static_cast<size_t>(n_chunks) * 2 * sizeof(float));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const LarsChunk* plan = static_cast<const LarsChunk*>(plan_dev);
  const float wd = static_cast<float>(weight_decay);
  launch_kernel(lars_norms_kernel, dim3(static_cast<unsigned>(n_chunks)), dim3(256), 0, st, plan, wd, workspace);
  DMC_LAUNCH_CHECK("lars_norms_kernel launch");
  launch_kernel(lars_update_kernel, dim3(static_cast<unsigned>(n_chunks)), dim3(256), 0, st, plan, static_cast<const float*>(workspace), wd,
                static_cast<float>(eta), static_cast<float>(momentum), static_cast<float>(-lr));
  DMC_LAUNCH_CHECK("lars_update_kernel launch");
  return 0;
}
