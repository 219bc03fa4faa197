// ema.cu -- EMA teacher update (main_dino_mc.py:403-406) as ONE multi-tensor launch.
//
// The reference issues mul_, scalar*tensor and add_ per parameter (3 launches x ~160 tensors per step).
// Here a host-built "plan" (one entry per <= 16384-element chunk of a (teacher, student) pair) drives a
// single kernel: one CTA per chunk, 128-bit coalesced loads/stores when both pointers are 16-byte
// aligned.  Arithmetic reproduces the reference's three fp32 roundings exactly:
//     p_k <- fl( fl(p_k * m) + fl((1-m) * p_q) ),   m and (1-m) cast to fp32 from float64.
// HBM-bound: 12 bytes per parameter (read student, read teacher, write teacher).
//
// Plan v2 (dmc_ema_build_plan2 / dmc_ema_multi_tensor2) additionally
//   * reads (m, 1-m) from DEVICE memory, so a captured CUDA graph follows the per-iteration momentum schedule
//     (main_dino_mc.py:404) instead of freezing the value it was captured with;
//   * emits the teacher head's GEMM operands as a by-product of the update it already streams: a bf16 copy of the new
//     MLP weights ("shadow"), and for the weight-normed last layer (utils/vision_transformer.py:279) the new
//     W = g v/||v|| in bf16 together with g/||v|| and 1/||v|| -- the next step's teacher forward then needs neither
//     its cast launches nor its weight-norm pass (64 MiB read per step at out_dim 65536).  The arithmetic of those rows
//     mirrors weightnorm_fwd_kernel (rowops.cu) operation for operation, so both routes give identical operands.
#include "dmc_common.cuh"

namespace dmc {
namespace {

constexpr long long kChunkElems = 16384;

struct EmaChunk {
  float* teacher;
  const float* student;
  long long n;
};

__device__ __forceinline__ float ema1(float pk, float pq, float m, float omm) {
  return __fadd_rn(__fmul_rn(pk, m), __fmul_rn(omm, pq));
}

__device__ __forceinline__ uint4 ema4(const uint4& a, const uint4& b, float m, float omm) {
  uint4 r;
  r.x = __float_as_uint(ema1(__uint_as_float(a.x), __uint_as_float(b.x), m, omm));
  r.y = __float_as_uint(ema1(__uint_as_float(a.y), __uint_as_float(b.y), m, omm));
  r.z = __float_as_uint(ema1(__uint_as_float(a.z), __uint_as_float(b.z), m, omm));
  r.w = __float_as_uint(ema1(__uint_as_float(a.w), __uint_as_float(b.w), m, omm));
  return r;
}

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
      st_stream_u4(pk + 4 * i, ema4(a, b, m, omm));
    }
    done = n4 << 2;
  }
  for (long long i = done + threadIdx.x; i < c.n; i += 256) pk[i] = ema1(pk[i], pq[i], m, omm);
}

// ---- plan v2 --------------------------------------------------------------------------------------------------
enum : int { kPlain = 0, kShadow = 1, kWeightNorm = 2 };

struct EmaChunk2 {
  float* teacher;
  const float* student;
  long long n;                 // elements (kWeightNorm: rows * dim)
  __nv_bfloat16* shadow;       // kShadow: bf16 copy of the new values; kWeightNorm: bf16 g v/||v|| of the rows
  float* g_teacher;            // kWeightNorm: gains of these rows (updated here, NOT by a plain chunk)
  const float* g_student;
  float* scale;                // kWeightNorm outputs per row: g/||v||, 1/||v||
  float* inv_norm;
  int kind, dim;
};

__global__ void __launch_bounds__(256)
ema2_kernel(const EmaChunk2* __restrict__ plan, const float* __restrict__ scalars) {
  pdl_prologue();
  const float m = __ldg(scalars), omm = __ldg(scalars + 1);
  const EmaChunk2 c = plan[blockIdx.x];
  float* __restrict__ pk = c.teacher;
  const float* __restrict__ pq = c.student;
  if (c.kind == kWeightNorm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rows = static_cast<int>(c.n / c.dim);
    if (c.dim == 256) {
      // kWnRows rows per warp, all their loads (teacher + student, 2 x 128 bits per lane and row) issued before the first
      // use: 16 independent 128-bit loads in flight per lane.  Lane owns float4 #lane and #lane+32 of a row (same mapping,
      // same FMA chain and same butterfly as weightnorm_fwd_kernel -> identical ||v||, g/||v|| and bf16 operand).
      constexpr int kWnRows = 4;
      for (int rb = warp * kWnRows; rb < rows; rb += 8 * kWnRows) {
        uint4 a0[kWnRows], a1[kWnRows], b0[kWnRows], b1[kWnRows];
        float g[kWnRows];
#pragma unroll
        for (int j = 0; j < kWnRows; ++j) {
          const int r = min(rb + j, rows - 1);                   // rows past the end repeat the last one and are not stored
          const float* ap = pk + static_cast<long long>(r) * 256;
          const float* bp = pq + static_cast<long long>(r) * 256;
          a0[j] = ld_stream_u4(ap + 4 * lane); a1[j] = ld_stream_u4(ap + 4 * (lane + 32));
          b0[j] = ld_stream_u4(bp + 4 * lane); b1[j] = ld_stream_u4(bp + 4 * (lane + 32));
          g[j] = ema1(c.g_teacher[r], __ldg(c.g_student + r), m, omm);
        }
        float ss[kWnRows];
#pragma unroll
        for (int j = 0; j < kWnRows; ++j) {
          a0[j] = ema4(a0[j], b0[j], m, omm);
          a1[j] = ema4(a1[j], b1[j], m, omm);
          const float v[8] = {__uint_as_float(a0[j].x), __uint_as_float(a0[j].y), __uint_as_float(a0[j].z), __uint_as_float(a0[j].w),
                              __uint_as_float(a1[j].x), __uint_as_float(a1[j].y), __uint_as_float(a1[j].z), __uint_as_float(a1[j].w)};
          float acc = v[0] * v[0];
#pragma unroll
          for (int e = 1; e < 8; ++e) acc = fmaf(v[e], v[e], acc);
          ss[j] = acc;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int j = 0; j < kWnRows; ++j) ss[j] += __shfl_xor_sync(0xffffffffu, ss[j], o);
#pragma unroll
        for (int j = 0; j < kWnRows; ++j) {
          const int r = rb + j;
          if (r >= rows) break;                                  // warp-uniform
          float* dst = pk + static_cast<long long>(r) * 256;
          st_stream_u4(dst + 4 * lane, a0[j]);
          st_stream_u4(dst + 4 * (lane + 32), a1[j]);
          const float nrm = sqrtf(ss[j]);
          const float sc = g[j] / nrm;
          if (lane == 0) { c.g_teacher[r] = g[j]; c.scale[r] = sc; c.inv_norm[r] = 1.0f / nrm; }
          __nv_bfloat16* w = c.shadow + static_cast<long long>(r) * 256;
          *reinterpret_cast<uint2*>(w + 4 * lane) =
              make_uint2(pack_bf16(__uint_as_float(a0[j].x) * sc, __uint_as_float(a0[j].y) * sc),
                         pack_bf16(__uint_as_float(a0[j].z) * sc, __uint_as_float(a0[j].w) * sc));
          *reinterpret_cast<uint2*>(w + 4 * (lane + 32)) =
              make_uint2(pack_bf16(__uint_as_float(a1[j].x) * sc, __uint_as_float(a1[j].y) * sc),
                         pack_bf16(__uint_as_float(a1[j].z) * sc, __uint_as_float(a1[j].w) * sc));
        }
      }
    } else {
      for (int r = warp; r < rows; r += 8) {                     // generic width: the scalar path of weightnorm_fwd_kernel
        float* vr = pk + static_cast<long long>(r) * c.dim;
        const float* sr = pq + static_cast<long long>(r) * c.dim;
        float ss = 0.f;
        for (int col = lane; col < c.dim; col += 32) {
          const float x = ema1(vr[col], sr[col], m, omm);
          vr[col] = x;
          ss = fmaf(x, x, ss);
        }
        ss = warp_sum(ss);
        const float nrm = sqrtf(ss);
        const float g = ema1(c.g_teacher[r], __ldg(c.g_student + r), m, omm);
        const float sc = g / nrm;
        __syncwarp();
        if (lane == 0) { c.g_teacher[r] = g; c.scale[r] = sc; c.inv_norm[r] = 1.0f / nrm; }
        for (int col = lane; col < c.dim; col += 32)             // each lane re-reads only what it wrote itself
          c.shadow[static_cast<long long>(r) * c.dim + col] = __float2bfloat16_rn(vr[col] * sc);
      }
    }
    return;
  }
  const bool vec = (((reinterpret_cast<uintptr_t>(pk) | reinterpret_cast<uintptr_t>(pq)) & 15) == 0) &&
                   (c.kind != kShadow || (reinterpret_cast<uintptr_t>(c.shadow) & 7) == 0);
  long long done = 0;
  if (vec) {
    const long long n4 = c.n >> 2;
    for (long long i = threadIdx.x; i < n4; i += 256) {
      const uint4 a = ld_stream_u4(pk + 4 * i);
      const uint4 b = ld_stream_u4(pq + 4 * i);
      const uint4 r = ema4(a, b, m, omm);
      st_stream_u4(pk + 4 * i, r);
      if (c.kind == kShadow)
        *reinterpret_cast<uint2*>(c.shadow + 4 * i) = make_uint2(pack_bf16(__uint_as_float(r.x), __uint_as_float(r.y)),
                                                                 pack_bf16(__uint_as_float(r.z), __uint_as_float(r.w)));
    }
    done = n4 << 2;
  }
  for (long long i = done + threadIdx.x; i < c.n; i += 256) {
    const float r = ema1(pk[i], pq[i], m, omm);
    pk[i] = r;
    if (c.kind == kShadow) c.shadow[i] = __float2bfloat16_rn(r);
  }
}

int64_t wn_rows_per_chunk(int64_t dim) { return kChunkElems / dim > 0 ? kChunkElems / dim : 1; }

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

extern "C" size_t dmc_ema_plan2_bytes(const int64_t* numels_host, int64_t n_tensors, int64_t wn_v_index, int64_t wn_dim) {
  if (!numels_host || n_tensors <= 0) return 0;
  size_t chunks = 0;
  for (int64_t i = 0; i < n_tensors; ++i) {
    if (numels_host[i] <= 0) continue;
    if (i == wn_v_index && wn_dim > 0) chunks += static_cast<size_t>(ceil_div(numels_host[i] / wn_dim, wn_rows_per_chunk(wn_dim)));
    else chunks += static_cast<size_t>(ceil_div(numels_host[i], kChunkElems));
  }
  return chunks * sizeof(EmaChunk2);
}

extern "C" int dmc_ema_build_plan2(const void* const* teacher_ptrs_host, const void* const* student_ptrs_host,
                                   const int64_t* numels_host, void* const* shadow_bf16_ptrs_host, int64_t n_tensors,
                                   int64_t wn_v_index, int64_t wn_g_index, int64_t wn_dim, void* wn_w_bf16, float* wn_scale,
                                   float* wn_inv_norm, void* plan_host, size_t plan_bytes, int64_t* n_chunks_out) {
  DMC_REQUIRE(teacher_ptrs_host && student_ptrs_host && numels_host && plan_host && n_chunks_out, "dmc_ema_build_plan2: null pointer");
  DMC_REQUIRE(n_tensors > 0, "dmc_ema_build_plan2: no tensors");
  const bool wn = wn_v_index >= 0;
  if (wn) {
    DMC_REQUIRE(wn_v_index < n_tensors && wn_g_index >= 0 && wn_g_index < n_tensors && wn_g_index != wn_v_index,
                "dmc_ema_build_plan2: bad weight-norm tensor indices");
    DMC_REQUIRE(wn_dim > 0 && wn_dim <= kChunkElems && numels_host[wn_v_index] % wn_dim == 0 &&
                numels_host[wn_g_index] == numels_host[wn_v_index] / wn_dim,
                "dmc_ema_build_plan2: weight_v must be [rows, dim] and weight_g [rows]");
    DMC_REQUIRE(wn_w_bf16 && wn_scale && wn_inv_norm, "dmc_ema_build_plan2: weight-norm outputs missing");
    DMC_REQUIRE((reinterpret_cast<uintptr_t>(wn_w_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(teacher_ptrs_host[wn_v_index]) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(student_ptrs_host[wn_v_index]) & 15) == 0,
                "dmc_ema_build_plan2: weight_v tensors and the operand buffer must be 16-byte aligned");
  } else {
    wn_dim = 0;
  }
  DMC_REQUIRE(plan_bytes >= dmc_ema_plan2_bytes(numels_host, n_tensors, wn ? wn_v_index : -1, wn_dim), "dmc_ema_build_plan2: plan buffer too small");
  EmaChunk2* out = static_cast<EmaChunk2*>(plan_host);
  int64_t n = 0;
  for (int64_t i = 0; i < n_tensors; ++i) {
    DMC_REQUIRE(numels_host[i] >= 0, "dmc_ema_build_plan2: negative numel at %lld", (long long)i);
    DMC_REQUIRE(numels_host[i] == 0 || (teacher_ptrs_host[i] && student_ptrs_host[i]), "dmc_ema_build_plan2: null tensor at %lld", (long long)i);
    DMC_REQUIRE((reinterpret_cast<uintptr_t>(teacher_ptrs_host[i]) & 3) == 0 && (reinterpret_cast<uintptr_t>(student_ptrs_host[i]) & 3) == 0,
                "dmc_ema_build_plan2: tensor %lld is not 4-byte aligned", (long long)i);
    float* tp = const_cast<float*>(static_cast<const float*>(teacher_ptrs_host[i]));
    const float* sp = static_cast<const float*>(student_ptrs_host[i]);
    if (wn && i == wn_g_index) continue;                         // the gains travel with their rows (kWeightNorm chunks)
    if (wn && i == wn_v_index) {
      const int64_t rows = numels_host[i] / wn_dim, rpc = wn_rows_per_chunk(wn_dim);
      float* gt = const_cast<float*>(static_cast<const float*>(teacher_ptrs_host[wn_g_index]));
      const float* gs = static_cast<const float*>(student_ptrs_host[wn_g_index]);
      for (int64_t r = 0; r < rows; r += rpc) {
        const int64_t nr = (rows - r < rpc) ? (rows - r) : rpc;
        EmaChunk2& c = out[n++];
        c.teacher = tp + r * wn_dim; c.student = sp + r * wn_dim; c.n = nr * wn_dim;
        c.shadow = static_cast<__nv_bfloat16*>(wn_w_bf16) + r * wn_dim;
        c.g_teacher = gt + r; c.g_student = gs + r; c.scale = wn_scale + r; c.inv_norm = wn_inv_norm + r;
        c.kind = kWeightNorm; c.dim = static_cast<int>(wn_dim);
      }
      continue;
    }
    __nv_bfloat16* sh = shadow_bf16_ptrs_host ? static_cast<__nv_bfloat16*>(shadow_bf16_ptrs_host[i]) : nullptr;
    for (int64_t off = 0; off < numels_host[i]; off += kChunkElems) {
      EmaChunk2& c = out[n++];
      c.teacher = tp + off; c.student = sp + off;
      c.n = (numels_host[i] - off < kChunkElems) ? (numels_host[i] - off) : kChunkElems;
      c.shadow = sh ? sh + off : nullptr;
      c.g_teacher = nullptr; c.g_student = nullptr; c.scale = nullptr; c.inv_norm = nullptr;
      c.kind = sh ? kShadow : kPlain; c.dim = 0;
    }
  }
  *n_chunks_out = n;
  return 0;
}

extern "C" int dmc_ema_multi_tensor2(const void* plan_dev, int64_t n_chunks, const float* scalars_dev, void* stream) {
  DMC_REQUIRE(plan_dev && scalars_dev && n_chunks > 0 && n_chunks < (1ll << 31), "dmc_ema_multi_tensor2: bad plan");
  launch_kernel(ema2_kernel, dim3(static_cast<unsigned>(n_chunks)), dim3(256), 0, static_cast<cudaStream_t>(stream),
                static_cast<const EmaChunk2*>(plan_dev), scalars_dev);
  DMC_LAUNCH_CHECK("ema2_kernel launch");
  return 0;
}
