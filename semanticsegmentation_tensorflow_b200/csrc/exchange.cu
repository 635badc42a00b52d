// Gradient exchange of data-parallel training as our own kernels over NVLink, on a
// symmetric buffer (the same virtual layout on every rank).  A sum-all-reduce is split by ownership: rank
// r reduces the r-th 1/world share of the range and broadcasts the result, so every element is summed
// once, in one fixed order, and every rank ends up with bit-identical values.
//   * multimem (NVLS): one multimem.ld_reduce through the NVSwitch multicast address returns the sum
//     over all ranks, reduced inside the switch; one multimem.st writes it back to all ranks.
//   * peer: plain loads from / stores to the peers' buffers, summed in rank order.
// 128-thread CTAs with a handful of registers and no shared memory: unlike an NCCL CTA they fit on an
// SM beside a 200 KB / 320-thread tensor-core CTA, so the persistent conv kernels keep all their SMs
// while a bucket is in flight.  The caller brackets a launch with cross-rank barriers on the same stream.
#include "common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kUnroll = 8;

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float4* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(kThreads, 8) multimem_allreduce_kernel(float4* __restrict__ mc, int64_t lo4, int64_t hi4) {
  const int64_t stride = (int64_t)gridDim.x * kThreads * kUnroll;
  for (int64_t base = lo4 + (int64_t)blockIdx.x * kThreads * kUnroll + threadIdx.x; base < hi4; base += stride) {
    float4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (base + u * kThreads < hi4) v[u] = multimem_ld_reduce_add(mc + base + u * kThreads);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (base + u * kThreads < hi4) multimem_st(mc + base + u * kThreads, v[u]);
  }
}

// Fused exchange + optimizer (ZeRO-1 style): rank r owns a 1/world share of the range.  For its share it
// pulls the gradient SUM over all ranks out of the switch (multimem.ld_reduce), applies TF's ApplyAdam
// (same arithmetic as adam_kernel in elementwise.cu) to its local parameter / slot values, and multicasts
// the new parameters into every rank's parameter arena (multimem.st).  Gradients are never written back,
// and each rank touches 1/world of the Adam state: 28 B/param of optimizer traffic become 28/world + 4.
__global__ void __launch_bounds__(kThreads, 8) multimem_allreduce_adam_kernel(
    const float4* __restrict__ g_mc, float4* __restrict__ p_mc, const float4* __restrict__ p_local,
    float4* __restrict__ m, float4* __restrict__ v, int64_t lo4, int64_t hi4, float lr_t, float b1, float b2, float eps) {
  const float c1 = 1.f - b1, c2 = 1.f - b2;
  const int64_t stride = (int64_t)gridDim.x * kThreads * 2;
  for (int64_t base = lo4 + (int64_t)blockIdx.x * kThreads * 2 + threadIdx.x; base < hi4; base += stride) {
    float4 gg[2], pp[2], mm[2], vv[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t i = base + u * kThreads;
      if (i < hi4) {
        gg[u] = multimem_ld_reduce_add(g_mc + i);
        pp[u] = p_local[i]; mm[u] = m[i]; vv[u] = v[i];
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t i = base + u * kThreads;
      if (i < hi4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float gj = (&gg[u].x)[j];
          const float mj = (&mm[u].x)[j] + (gj - (&mm[u].x)[j]) * c1;
          const float vj = (&vv[u].x)[j] + (gj * gj - (&vv[u].x)[j]) * c2;
          (&mm[u].x)[j] = mj;
          (&vv[u].x)[j] = vj;
          (&pp[u].x)[j] -= lr_t * mj / (sqrtf(vj) + eps);
        }
        m[i] = mm[u];
        v[i] = vv[u];
        multimem_st(p_mc + i, pp[u]);
      }
    }
  }
}

constexpr int kMaxWorld = 16;
struct PeerPtrs {
  float4* p[kMaxWorld];
};

__global__ void __launch_bounds__(kThreads, 8) peer_allreduce_kernel(const PeerPtrs ptrs, int world, int64_t lo4, int64_t hi4) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = lo4 + (int64_t)blockIdx.x * kThreads + threadIdx.x; i < hi4; i += stride) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {          // fixed rank order: the same sum on every rank
      const float4 v = __ldcv(ptrs.p[r] + i);  // volatile: never a stale cached copy of a peer's line
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    for (int r = 0; r < world; ++r) ptrs.p[r][i] = a;
  }
}

}  // namespace

extern "C" int segk_allreduce_f32(segk_ctx* ctx, void* multicast_ptr, const uint64_t* peer_ptrs, int64_t offset,
                                  int64_t n, int rank, int world, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, (multicast_ptr || peer_ptrs) && n > 0 && offset >= 0, "allreduce: bad args");
  SEGK_REQUIRE(ctx, world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "allreduce: rank %d of %d", rank, world);
  SEGK_REQUIRE(ctx, offset % 4 == 0 && n % 4 == 0, "allreduce: offset and count must be multiples of 4 floats (got %lld, %lld)",
               (long long)offset, (long long)n);
  const int64_t n4 = n / 4, off4 = offset / 4;
  const int64_t share = ceil_div64(n4, world);
  const int64_t lo4 = off4 + share * rank;
  int64_t hi4 = lo4 + share;
  if (hi4 > off4 + n4) hi4 = off4 + n4;
  if (hi4 <= lo4) return SEGK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (multicast_ptr) {
    int64_t blocks = ceil_div64(hi4 - lo4, (int64_t)kThreads * kUnroll);
    if (blocks > 2 * ctx->sm_count) blocks = 2 * ctx->sm_count;
    multimem_allreduce_kernel<<<(int)blocks, kThreads, 0, st>>>((float4*)multicast_ptr, lo4, hi4);
  } else {
    PeerPtrs pp;
    for (int r = 0; r < world; ++r) pp.p[r] = reinterpret_cast<float4*>((uintptr_t)peer_ptrs[r]);
    int64_t blocks = ceil_div64(hi4 - lo4, kThreads);
    if (blocks > 4 * ctx->sm_count) blocks = 4 * ctx->sm_count;
    peer_allreduce_kernel<<<(int)blocks, kThreads, 0, st>>>(pp, world, lo4, hi4);
  }
  SEGK_LAUNCHED(ctx, "allreduce");
  return SEGK_OK;
}

extern "C" int segk_allreduce_adam_f32(segk_ctx* ctx, const void* grad_multicast_ptr, void* param_multicast_ptr,
                                       const float* param_local, float* m, float* v, int64_t offset, int64_t n, int rank,
                                       int world, float lr_t, float beta1, float beta2, float eps, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, grad_multicast_ptr && param_multicast_ptr && param_local && m && v && n > 0 && offset >= 0,
               "allreduce_adam: bad args (needs the multicast addresses of both arenas)");
  SEGK_REQUIRE(ctx, world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "allreduce_adam: rank %d of %d", rank, world);
  SEGK_REQUIRE(ctx, offset % 4 == 0 && n % 4 == 0, "allreduce_adam: offset and count must be multiples of 4 floats");
  SEGK_REQUIRE(ctx, (((uintptr_t)param_local | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "allreduce_adam: 16-byte alignment");
  const int64_t n4 = n / 4, off4 = offset / 4;
  const int64_t share = ceil_div64(n4, world);
  const int64_t lo4 = off4 + share * rank;
  int64_t hi4 = lo4 + share;
  if (hi4 > off4 + n4) hi4 = off4 + n4;
  if (hi4 <= lo4) return SEGK_OK;
  int64_t blocks = ceil_div64(hi4 - lo4, (int64_t)kThreads * 2);
  if (blocks > 4 * ctx->sm_count) blocks = 4 * ctx->sm_count;
  multimem_allreduce_adam_kernel<<<(int)blocks, kThreads, 0, (cudaStream_t)stream>>>(
      (const float4*)grad_multicast_ptr, (float4*)param_multicast_ptr, (const float4*)param_local, (float4*)m, (float4*)v, lo4,
      hi4, lr_t, beta1, beta2, eps);
  SEGK_LAUNCHED(ctx, "allreduce_adam");
  return SEGK_OK;
}
