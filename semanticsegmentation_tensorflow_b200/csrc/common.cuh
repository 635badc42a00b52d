// Shared helpers for libsegk.so (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/segk.h"

struct segk_ctx {
  int device = 0;
  int sm_count = 148;
  std::atomic<int64_t> launches{0};
  char err[512] = {0};
  // tuning overrides, read once from the environment (tools/sweep_tiles.py); 0 = heuristics
  int force_bn = 0, force_ksplit = 0, force_wsplit = 0;
  int slab_mode = 1;        // SEGK_SLAB: 0 off, 1 auto, 2 wherever legal
  int tma_store = 1;        // SEGK_TMA_STORE: bf16 conv outputs leave through smem + TMA store
  int slab3 = 1;            // SEGK_SLAB3: kx-fused N = 192 slab kernel with resident weights for Ck = 64
  int tail_wide = 1;        // SEGK_TAIL_WIDE: 16-byte-access forms of the few-pixel / wide-channel tail layers (conv8, conv_t1)
  int teamk = 1;            // SEGK_TEAMK: lockstep tap-split schedule instead of plain split-K for few-tile / long-K layers
                            //   (conv6 dgrad; IgemmParams::ts in tcconv.cu).  0 = plain split-K
  int pair = 1;             // SEGK_PAIR: 256-column igemm launches as CTA pairs (cta_group::2, M = 256 MMAs; igemm_pair_kernel);
                            //   2 = also the 128-column haloed-slab tiles (slab_pair_kernel<128>; measured no faster, off by default)
  int hybrid = 0;           // SEGK_HYBRID: whole waves + K-split remainder tiles when the last wave of an igemm launch is partly
                            //   filled.  Off by default: measured on conv5_x (B=32, 180 tiles on 148 SMs) 58 us vs 51 us for two
                            //   plain waves -- the two epilogues after the last MMA and the finish kernel (launch + 10 us) cost
                            //   more than the 54 idle k-steps they save (profiles/r2_hybrid_probe.md)
  // one-shot channel pitches of the next conv-family call (segk_set_pitch: zero-copy Concat views); 0 = dense.  The
  // entry points that understand them take (and clear) them before their first launch; any other launch with a pitch
  // still pending fails (SEGK_LAUNCHED)
  int pitch_in = 0, pitch_out = 0;
  void* ws = nullptr;       // grow-only scratch for split-K partial sums (tcconv.cu)
  size_t ws_bytes = 0;
  void* ws2 = nullptr;      // grow-only scratch for per-block BiasAddGrad partials (elementwise.cu)
  size_t ws2_bytes = 0;
  int wslab = 1;            // SEGK_WSLAB: slab-formulated wgrad for 3x3 layers with Cin 64/128 on large maps (2 = wherever legal)
  void* ws4 = nullptr;      // grow-only scratch for the slab wgrad's per-split partial sums (wslab.cu; own buffer:
  size_t ws4_bytes = 0;     //   that kernel may run on a different stream than the users of ws)
  void* ws6 = nullptr;      // grow-only scratch for the dgrad epilogues' column-sum partial rows (tcconv.cu; main stream)
  size_t ws6_bytes = 0;
  void* ws5 = nullptr;      // grow-only scratch for the pool-backward's BiasAddGrad partial rows (elementwise.cu; main stream)
  size_t ws5_bytes = 0;
  void* ws7 = nullptr;      // grow-only scratch for the depthwise conv's filter-gradient partial rows (opfam.cu)
  size_t ws7_bytes = 0;
  void* ws3 = nullptr;      // grow-only scratch for the full-resolution 1x1 head's wgrad partials (smallconv.cu)
  size_t ws3_bytes = 0;
  // driver entry point resolved at segk_create (no link-time libcuda dependency)
  CUresult (*encode_tiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                           CUtensorMapFloatOOBfill) = nullptr;
};

int segk_tc_init(segk_ctx* ctx);   // tcconv.cu

inline int segk_fail(segk_ctx* ctx, int code, const char* fmt, ...) {
  if (ctx) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

// channel pitches set for this call by segk_set_pitch (consumed: the next call is dense again)
struct SegkPitch {
  int in, out;
};
inline SegkPitch segk_take_pitch(segk_ctx* ctx) {
  SegkPitch p{ctx->pitch_in, ctx->pitch_out};
  ctx->pitch_in = ctx->pitch_out = 0;
  return p;
}

#define SEGK_REQUIRE(ctx, cond, ...)                                   \
  do {                                                                 \
    if (!(cond)) return segk_fail((ctx), SEGK_EINVAL, __VA_ARGS__);    \
  } while (0)

// call after a kernel launch: counts it and converts launch errors into a status
#define SEGK_LAUNCHED(ctx, what)                                                        \
  do {                                                                                  \
    (ctx)->launches.fetch_add(1, std::memory_order_relaxed);                            \
    if ((ctx)->pitch_in | (ctx)->pitch_out) {                                           \
      (ctx)->pitch_in = (ctx)->pitch_out = 0;                                           \
      return segk_fail((ctx), SEGK_EINVAL, "%s: a channel pitch (segk_set_pitch) is pending, but this call takes dense tensors only", (what)); \
    }                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess)                                                             \
      return segk_fail((ctx), SEGK_ECUDA, "%s: %s", (what), cudaGetErrorString(e__));   \
  } while (0)

// Grow-only context scratch: (re)allocates *buf on the CONTEXT's device to at least `need` bytes (the
// caller's current device is saved and restored; cudaFree synchronises, so nothing still reads the old one).
inline int segk_grow(segk_ctx* ctx, void** buf, size_t* cur, size_t need, const char* what) {
  if (*cur >= need) return SEGK_OK;
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != ctx->device) cudaSetDevice(ctx->device);
  if (*buf) cudaFree(*buf);
  *buf = nullptr;
  *cur = 0;
  cudaError_t e = cudaMalloc(buf, need);
  if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
  if (e != cudaSuccess) return segk_fail(ctx, SEGK_ENOMEM, "%s: scratch of %zu bytes: %s", what, need, cudaGetErrorString(e));
  *cur = need;
  return SEGK_OK;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float bf2f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ bf16 f2bf(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(t);
}

// [bf16 value > 0] of the two halves of a packed pair (bit 0: low half, bit 1: high half): a bf16 pattern is positive
// iff it is > 0 as a signed 16-bit integer
__device__ __forceinline__ uint32_t bf16x2_pos_bits(uint32_t pk) {
  return ((int)(short)(pk & 0xffffu) > 0 ? 1u : 0u) | ((int)pk > 0xffff ? 2u : 0u);
}

// ---- dropout pattern shared by segk_dropout and the kernels that apply it on the fly ---------------------------
// Philox4x32-10 keyed by the seed, counter = element index / 4 (tf.nn.dropout, FCN.py:165-167 / utils.py:318).
__device__ __forceinline__ uint4 segk_philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// keep bits (bit j = element 8 i + j is kept) of the i-th 8-element group of a dropout tensor: from the injected
// u8 mask when there is one, else from the Philox stream -- ONE Philox4x32-10 block per group (counter = i), 16 random
// bits per element: element j is kept iff its 16-bit uniform < keep_prob * 65536 (keep probabilities are resolved to
// 1 / 65536; the generator, not memory, bounds the kernels that apply dropout on the fly)
__device__ __forceinline__ uint32_t segk_dropout_keep8(const uint2* __restrict__ mask, int64_t i, float keep, uint64_t seed) {
  uint32_t kp = 0;
  if (mask) {
    const uint2 m = __ldg(mask + i);
#pragma unroll
    for (int j = 0; j < 8; ++j) kp |= ((((&m.x)[j >> 2] >> (8 * (j & 3))) & 0xffu) != 0 ? 1u : 0u) << j;
  } else {
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint4 r = segk_philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0u, 0u), key);
    const uint32_t thr = (uint32_t)(keep * 65536.0f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      kp |= (((&r.x)[j] & 0xffffu) < thr ? 1u : 0u) << (2 * j);
      kp |= (((&r.x)[j] >> 16) < thr ? 1u : 0u) << (2 * j + 1);
    }
  }
  return kp;
}
