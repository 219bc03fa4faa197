// clip.cu -- per-parameter gradient clipping (utils/utils.py:145-154, called at main_dino_mc.py:387-388 / :394-396)
// as TWO multi-tensor launches over all gradients instead of, per parameter, a norm kernel, an .item() host
// sync and a mul_ (~500 launches and ~160 synchronisations per step in the reference).
//
//   for every parameter with a gradient:  n = ||grad||_2 ;  c = clip / (n + 1e-6) ;  if c < 1: grad *= c
//
// A host-built plan (same chunking as the EMA plan: one entry per <= 16384-element chunk, plus the chunk range
// of its tensor) drives both kernels:
//   pass 1  one CTA per chunk: partial sum of squares -> workspace (fixed-order block reduction)
//   pass 2  one CTA per chunk: re-adds the partials of its tensor in index order (every CTA of a tensor gets
//           the identical fp32 norm), forms the coefficient with the reference's fp32 operations and scales
//           its chunk if c < 1; the tensor's first chunk also records the norm.
// Deterministic (no atomics), never synchronises; norms stay on the device until the caller reads them.
// HBM-bound: 8 bytes read per gradient element (the second read mostly hits L2), 4 written where clipped.
#include "dmc_common.cuh"

namespace dmc {
namespace {

constexpr long long kClipChunk = 16384;

struct ClipChunk {
  float* grad;
  long long n;
  int tensor;           // index into norms[]
  int first, count;     // chunk range [first, first + count) of this tensor in the plan
  int pad;
};

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < 8) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
  }
  return t;             // valid in warp 0
}

__global__ void __launch_bounds__(256)
clip_sumsq_kernel(const ClipChunk* __restrict__ plan, float* __restrict__ partial) {
  pdl_prologue();
  __shared__ float red[8];
  const ClipChunk c = plan[blockIdx.x];
  const float* __restrict__ g = c.grad;
  float acc = 0.f;
  long long done = 0;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const long long n4 = c.n >> 2;
    for (long long i = threadIdx.x; i < n4; i += 256) {
      const float4 v = *reinterpret_cast<const float4*>(g + 4 * i);
      acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    }
    done = n4 << 2;
  }
  for (long long i = done + threadIdx.x; i < c.n; i += 256) acc = fmaf(g[i], g[i], acc);
  const float t = block_sum_256(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__(256)
clip_scale_kernel(const ClipChunk* __restrict__ plan, const float* __restrict__ partial, float clip, float* __restrict__ norms) {
  pdl_prologue();
  __shared__ float red[8];
  __shared__ float coef_s;
  const ClipChunk c = plan[blockIdx.x];
  // the tensor's sum of squares: its partials in index order, strided over the block, then the fixed block reduction
  float acc = 0.f;
  for (int i = threadIdx.x; i < c.count; i += 256) acc += partial[c.first + i];
  const float ss = block_sum_256(acc, red);
  if (threadIdx.x == 0) {
    const float nrm = sqrtf(ss);
    if (static_cast<int>(blockIdx.x) == c.first) norms[c.tensor] = nrm;
    coef_s = __fdiv_rn(clip, __fadd_rn(nrm, 1e-6f));       // clip / (param_norm + 1e-6), fp32 like the reference
  }
  __syncthreads();
  const float coef = coef_s;
  if (!(coef < 1.0f)) return;                               // also leaves NaN norms untouched, like `if clip_coef < 1`
  float* __restrict__ g = c.grad;
  long long done = 0;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const long long n4 = c.n >> 2;
    for (long long i = threadIdx.x; i < n4; i += 256) {
      float4 v = *reinterpret_cast<float4*>(g + 4 * i);
      v.x *= coef; v.y *= coef; v.z *= coef; v.w *= coef;
      *reinterpret_cast<float4*>(g + 4 * i) = v;
    }
    done = n4 << 2;
  }
  for (long long i = done + threadIdx.x; i < c.n; i += 256) g[i] *= coef;
}

}  // namespace
}  // namespace dmc

using namespace dmc;

extern "C" size_t dmc_clip_plan_bytes(const int64_t* numels_host, int64_t n_tensors) {
  if (!numels_host || n_tensors <= 0) return 0;
  size_t chunks = 0;
  for (int64_t i = 0; i < n_tensors; ++i)
    if (numels_host[i] > 0) chunks += static_cast<size_t>(ceil_div(numels_host[i], kClipChunk));
  return chunks * sizeof(ClipChunk);
}

extern "C" int dmc_clip_build_plan(const void* const* grad_ptrs_host, const int64_t* numels_host, int64_t n_tensors,
                                   void* plan_host, size_t plan_bytes, int64_t* n_chunks_out) {
  DMC_REQUIRE(grad_ptrs_host && numels_host && plan_host && n_chunks_out, "dmc_clip_build_plan: null pointer");
  DMC_REQUIRE(n_tensors > 0 && n_tensors < (1ll << 31), "dmc_clip_build_plan: bad tensor count");
  DMC_REQUIRE(plan_bytes >= dmc_clip_plan_bytes(numels_host, n_tensors), "dmc_clip_build_plan: plan buffer too small");
  ClipChunk* out = static_cast<ClipChunk*>(plan_host);
  int64_t n = 0;
  for (int64_t i = 0; i < n_tensors; ++i) {
    DMC_REQUIRE(numels_host[i] >= 0, "dmc_clip_build_plan: negative numel at %lld", (long long)i);
    DMC_REQUIRE(numels_host[i] == 0 || grad_ptrs_host[i], "dmc_clip_build_plan: null tensor at %lld", (long long)i);
    DMC_REQUIRE((reinterpret_cast<uintptr_t>(grad_ptrs_host[i]) & 3) == 0, "dmc_clip_build_plan: tensor %lld is not 4-byte aligned", (long long)i);
    const int64_t count = ceil_div(numels_host[i], kClipChunk);
    DMC_REQUIRE(n + count < (1ll << 31), "dmc_clip_build_plan: too many chunks");
    const int64_t first = n;
    for (int64_t off = 0; off < numels_host[i]; off += kClipChunk) {
      out[n].grad = const_cast<float*>(static_cast<const float*>(grad_ptrs_host[i])) + off;
      out[n].n = (numels_host[i] - off < kClipChunk) ? (numels_host[i] - off) : kClipChunk;
      out[n].tensor = static_cast<int>(i);
      out[n].first = static_cast<int>(first);
      out[n].count = static_cast<int>(count);
      out[n].pad = 0;
      ++n;
    }
  }
  *n_chunks_out = n;
  return 0;
}

extern "C" int dmc_clip_grads(const void* plan_dev, int64_t n_chunks, float clip, float* norms, float* workspace,
                              size_t workspace_bytes, void* stream) {
  DMC_REQUIRE(plan_dev && norms && workspace, "dmc_clip_grads: null pointer");
  DMC_REQUIRE(n_chunks > 0 && n_chunks < (1ll << 31), "dmc_clip_grads: bad plan");
  DMC_REQUIRE(workspace_bytes >= static_cast<size_t>(n_chunks) * sizeof(float), "dmc_clip_grads: workspace too small (%zu < %zu)",
              workspace_bytes, static_cast<size_t>(n_chunks) * sizeof(float));
  DMC_REQUIRE(clip > 0.f, "dmc_clip_grads: clip must be positive");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const ClipChunk* plan = static_cast<const ClipChunk*>(plan_dev);
  launch_kernel(clip_sumsq_kernel, dim3(static_cast<unsigned>(n_chunks)), dim3(256), 0, st, plan, workspace);
  DMC_LAUNCH_CHECK("clip_sumsq_kernel launch");
  launch_kernel(clip_scale_kernel, dim3(static_cast<unsigned>(n_chunks)), dim3(256), 0, st, plan, static_cast<const float*>(workspace), clip, norms);
  DMC_LAUNCH_CHECK("clip_scale_kernel launch");
  return 0;
}
