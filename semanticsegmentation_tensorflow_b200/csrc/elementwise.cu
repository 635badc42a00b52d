// HBM-bound kernels of the segmentation hot path: max-pool fwd/bwd (bit-exact argmax
// routing), dropout, fused softmax-xent + gradient + argmax + confusion counts, Adam /
// momentum, casts, bias-grad and weight-layout packing.  All are pure streaming kernels:
// 16-byte vector accesses, grid = multiple of the SM count, grid-stride loops.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

inline int stream_grid(segk_ctx* ctx, int64_t work_items, int per_sm = 8) {
  int64_t blocks = ceil_div64(work_items, kThreads);
  int64_t cap = (int64_t)ctx->sm_count * per_sm;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

// ------------------------------------------------------------------------------------
// max-pool 2x2 / stride 2 / VALID  (FCN.py:161-163).  One thread = 8 channels (16 B) of one
// pooled pixel.  idx = first maximal element under a strict '>' scan in (dy,dx) row-major
// order: exactly the element TF's MaxPoolGrad (and torch) route the gradient to.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) maxpool_fwd_kernel(const uint4* __restrict__ x,
                                                               uint4* __restrict__ y,
                                                               uint2* __restrict__ idx, int N, int H,
                                                               int W, int C8, int P8) {
  // P8: 8-channel groups between consecutive pixels of x (= C8 when dense; a channel slice of a wider tensor otherwise)
  const int OH = H >> 1, OW = W >> 1;
  const int64_t total = (int64_t)N * OH * OW * C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c8 = (int)(i % C8);
    int64_t p = i / C8;
    int ox = (int)(p % OW);
    p /= OW;
    int oy = (int)(p % OH);
    int n = (int)(p / OH);
    const int64_t row0 = ((int64_t)(n * H + 2 * oy) * W + 2 * ox) * P8 + c8;
    uint4 v[4];
    v[0] = __ldg(x + row0);
    v[1] = __ldg(x + row0 + P8);
    v[2] = __ldg(x + row0 + (int64_t)W * P8);
    v[3] = __ldg(x + row0 + (int64_t)W * P8 + P8);
    uint32_t outw[4];
    uint32_t idxw[2] = {0u, 0u};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t w0 = (&v[0].x)[j];
      float2 best = unpack_bf16x2(w0);
      uint32_t bits_lo = w0 & 0xffffu, bits_hi = w0 >> 16;
      uint32_t k_lo = 0, k_hi = 0;
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        const uint32_t wk = (&v[k].x)[j];
        float2 f = unpack_bf16x2(wk);
        if (f.x > best.x) { best.x = f.x; bits_lo = wk & 0xffffu; k_lo = k; }
        if (f.y > best.y) { best.y = f.y; bits_hi = wk >> 16; k_hi = k; }
      }
      outw[j] = bits_lo | (bits_hi << 16);
      const int e = 2 * j;  // element index within the 8-channel group
      idxw[e >> 2] |= (k_lo << (8 * (e & 3))) | (k_hi << (8 * ((e + 1) & 3)));
    }
    y[i] = make_uint4(outw[0], outw[1], outw[2], outw[3]);
    idx[i] = make_uint2(idxw[0], idxw[1]);
  }
}

// MaxPoolGrad from the stored index when the pre-pool tensor is a ReLU output and no second gradient path
// exists: ReluGrad needs [act > 0] only at the routed element, and there act == the pooled value, so the
// mask comes from the POOLED activation (1/4 of the bytes): dx = idx == k && pooled > 0 ? dy : 0 -- the
// same bits as masking with the full-resolution activation (a window whose max is 0 routes nothing).
//
// BIAS: BiasAddGrad of the conv in front of the pool falls out of the same pass.  Every dy lands on exactly one
// element of the window, so the column sums of dx are the column sums of the masked dy: each thread always owns
// the same 8 channels (the thread count is a multiple of C/8), accumulates them in registers, the block reduces
// over its threads in a fixed order and writes one partial row; reduce_rows_kernel adds the rows.  Saves the
// separate pass over the full-resolution dz (conv1_2: 377 MB).
template <bool BIAS>
__global__ void __launch_bounds__(kThreads) maxpool_bwd_pooled_kernel(const uint4* __restrict__ dy,
                                                                      const uint2* __restrict__ idx,
                                                                      const uint4* __restrict__ pooled,
                                                                      uint4* __restrict__ dx, float* __restrict__ db_part,
                                                                      int N, int H, int W, int C8) {
  const int OH = H >> 1, OW = W >> 1;
  const int64_t total = (int64_t)N * OH * OW * C8;
  float bsum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bsum[j] = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c8 = (int)(i % C8);
    int64_t p = i / C8;
    int ox = (int)(p % OW);
    p /= OW;
    int oy = (int)(p % OH);
    int n = (int)(p / OH);
    const int64_t row0 = ((int64_t)(n * H + 2 * oy) * W + 2 * ox) * C8 + c8;
    const int64_t offs[4] = {row0, row0 + C8, row0 + (int64_t)W * C8, row0 + (int64_t)W * C8 + C8};
    uint4 g = __ldg(dy + i);
    const uint2 id = __ldg(idx + i);
    const uint4 a = __ldg(pooled + i);
    uint32_t klo[4], khi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = 2 * j;
      const float2 af = unpack_bf16x2((&a.x)[j]);
      // a masked element gets an index no window position matches
      klo[j] = af.x > 0.f ? (((&id.x)[e >> 2] >> (8 * (e & 3))) & 0xffu) : 4u;
      khi[j] = af.y > 0.f ? (((&id.x)[e >> 2] >> (8 * ((e + 1) & 3))) & 0xffu) : 4u;
      if (BIAS) {
        const float2 gf = unpack_bf16x2((&g.x)[j]);
        if (af.x > 0.f) bsum[e] += gf.x;
        if (af.y > 0.f) bsum[e + 1] += gf.y;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t gw = (&g.x)[j];
        o[j] = (klo[j] == (uint32_t)k ? (gw & 0xffffu) : 0u) | (khi[j] == (uint32_t)k ? (gw & 0xffff0000u) : 0u);
      }
      dx[offs[k]] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  if (BIAS) {
    __shared__ float sh[kThreads][9];
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[threadIdx.x][j] = bsum[j];
    __syncthreads();
    if ((int)threadIdx.x < C8) {            // thread t owns channel group t % C8: sum t, t + C8, t + 2 C8, ...
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = 0.f;
        for (int k = threadIdx.x; k < kThreads; k += C8) t += sh[k][j];
        db_part[(int64_t)blockIdx.x * (C8 * 8) + threadIdx.x * 8 + j] = t;
      }
    }
  }
}

// MaxPoolGrad from the stored index fused with ReluGrad of the pooled activation.
__global__ void __launch_bounds__(kThreads) maxpool_bwd_kernel(const uint4* __restrict__ dy,
                                                               const uint2* __restrict__ idx,
                                                               const uint4* __restrict__ act,
                                                               const uint4* __restrict__ res,
                                                               uint4* __restrict__ dx, int N, int H,
                                                               int W, int C8, int P8) {
  // P8: 8-channel groups between consecutive pixels of dx / act / res (= C8 when dense)
  const int OH = H >> 1, OW = W >> 1;
  const int64_t total = (int64_t)N * OH * OW * C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c8 = (int)(i % C8);
    int64_t p = i / C8;
    int ox = (int)(p % OW);
    p /= OW;
    int oy = (int)(p % OH);
    int n = (int)(p / OH);
    const int64_t row0 = ((int64_t)(n * H + 2 * oy) * W + 2 * ox) * P8 + c8;
    const int64_t offs[4] = {row0, row0 + P8, row0 + (int64_t)W * P8, row0 + (int64_t)W * P8 + P8};
    const uint4 g = __ldg(dy + i);
    const uint2 id = __ldg(idx + i);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint4 a = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);  // +1.0 pairs
      if (act) a = __ldg(act + offs[k]);
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = 2 * j;
        const uint32_t k_lo = ((&id.x)[e >> 2] >> (8 * (e & 3))) & 0xffu;
        const uint32_t k_hi = ((&id.x)[e >> 2] >> (8 * ((e + 1) & 3))) & 0xffu;
        const float2 af = unpack_bf16x2((&a.x)[j]);
        const uint32_t gw = (&g.x)[j];
        uint32_t lo = (k_lo == (uint32_t)k && af.x > 0.f) ? (gw & 0xffffu) : 0u;
        uint32_t hi = (k_hi == (uint32_t)k && af.y > 0.f) ? (gw & 0xffff0000u) : 0u;
        o[j] = lo | hi;
      }
      if (res) {   // second gradient path into the same tensor (added before the mask)
        const uint4 rv = __ldg(res + offs[k]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 af = unpack_bf16x2((&a.x)[j]);
          const float2 rf = unpack_bf16x2((&rv.x)[j]);
          const float2 of = unpack_bf16x2(o[j]);
          o[j] = pack_bf16x2(af.x > 0.f ? of.x + rf.x : 0.f, af.y > 0.f ? of.y + rf.y : 0.f);
        }
      }
      dx[offs[k]] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ------------------------------------------------------------------------------------
// dropout (FCN.py:165-167).  Philox4x32-10 keyed by (seed), counter = element index / 8, 16 random bits per element.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) { return segk_philox4x32_10(ctr, key); }

__global__ void __launch_bounds__(kThreads) dropout_kernel(const bf16* __restrict__ x,
                                                           bf16* __restrict__ y,
                                                           const uint8_t* __restrict__ mask, int64_t n,
                                                           float keep, float inv_keep, uint64_t seed) {
  // scalar form (any n / alignment): thread = one 8-element group, the same pattern as segk_dropout_keep8
  const int64_t n8 = (n + 7) >> 3;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t kp = 0;
    if (mask) {
      for (int j = 0; j < 8 && i * 8 + j < n; ++j) kp |= (mask[i * 8 + j] != 0 ? 1u : 0u) << j;
    } else {
      kp = segk_dropout_keep8(nullptr, i, keep, seed);
    }
    for (int j = 0; j < 8; ++j) {
      const int64_t e = i * 8 + j;
      if (e >= n) break;
      y[e] = (kp >> j) & 1u ? f2bf(bf2f(x[e]) * inv_keep) : f2bf(0.f);
    }
  }
}

// The same for n % 8 == 0 and 16-byte aligned buffers: 8 elements per thread (one 16-byte load / store), the same
// keep pattern as the scalar form.
__global__ void __launch_bounds__(kThreads) dropout_vec8_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                                const uint2* __restrict__ mask, int64_t n8,
                                                                float keep, float inv_keep, uint64_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(x + i);
    const uint32_t kp = segk_dropout_keep8(mask, i, keep, seed);      // bit j: keep element j
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2((&v.x)[j]);
      o[j] = pack_bf16x2((kp >> (2 * j)) & 1u ? f.x * inv_keep : 0.f, (kp >> (2 * j + 1)) & 1u ? f.y * inv_keep : 0.f);
    }
    y[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------
// fused softmax-xent + dlogits + argmax + confusion counts (FCN.py:334,111), C == 2 fast
// path (float2 / pixel) and a generic small-C path.  Loss: per-block partial sums written to
// the workspace, then summed in a fixed order by a single block => deterministic.
// ------------------------------------------------------------------------------------
constexpr int kXentBlocks = 148 * 8;

__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = 0.f;
  if (w == 0) {
    t = (l < (int)(blockDim.x >> 5)) ? sh[l] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;  // valid in warp 0
}

template <int C>
__global__ void __launch_bounds__(kThreads) xent_kernel(const float* __restrict__ logits,
                                                        const uint8_t* __restrict__ labels,
                                                        float* __restrict__ dlogits,
                                                        uint8_t* __restrict__ pred,
                                                        float* __restrict__ partial,
                                                        unsigned long long* __restrict__ cm,
                                                        int64_t npix, int Crt, float scale) {
  __shared__ float sh[32];
  __shared__ unsigned int cm_sh[4];
  if (threadIdx.x < 4) cm_sh[threadIdx.x] = 0;
  __syncthreads();
  float loss = 0.f;
  unsigned int cnt[4] = {0, 0, 0, 0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npix;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int lab = labels[i];
    const bool ignore = lab >= (C == 2 ? 2 : Crt);       // label outside [0, C): no loss, zero gradient
    int am = 0;
    if (C == 2) {
      const float2 l = __ldg(reinterpret_cast<const float2*>(logits) + i);
      const float mx = fmaxf(l.x, l.y);
      const float z0 = l.x - mx, z1 = l.y - mx;
      const float e0 = expf(z0), e1 = expf(z1);
      const float s = e0 + e1;
      if (!ignore) loss += logf(s) - (lab ? z1 : z0);
      am = l.y > l.x ? 1 : 0;
      if (dlogits) {
        const float inv = 1.f / s;
        float2 d;
        d.x = ignore ? 0.f : (e0 * inv - (lab == 0 ? 1.f : 0.f)) * scale;
        d.y = ignore ? 0.f : (e1 * inv - (lab == 1 ? 1.f : 0.f)) * scale;
        reinterpret_cast<float2*>(dlogits)[i] = d;
      }
    } else {
      const float* l = logits + i * Crt;
      float mx = l[0];
      for (int c = 1; c < Crt; ++c) {
        if (l[c] > mx) { mx = l[c]; am = c; }
      }
      float s = 0.f;
      for (int c = 0; c < Crt; ++c) s += expf(l[c] - mx);
      if (!ignore) loss += logf(s) - (l[lab] - mx);
      if (dlogits) {
        const float inv = 1.f / s;
        for (int c = 0; c < Crt; ++c)
          dlogits[i * Crt + c] = ignore ? 0.f : (expf(l[c] - mx) * inv - (c == lab ? 1.f : 0.f)) * scale;
      }
    }
    if (pred) pred[i] = (uint8_t)am;
    if (C == 2 && !ignore) cnt[lab * 2 + am]++;
  }
  if (cm && C == 2) {
    // warp-aggregated: one shared-memory atomic per warp per bin, one global per block
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      unsigned int w = __reduce_add_sync(0xffffffffu, cnt[b]);
      if ((threadIdx.x & 31) == 0 && w) atomicAdd(&cm_sh[b], w);
    }
  }
  const float tot = block_sum(loss, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
  if (cm && C == 2 && threadIdx.x < 4 && cm_sh[threadIdx.x])
    atomicAdd(&cm[threadIdx.x], (unsigned long long)cm_sh[threadIdx.x]);
}

__global__ void __launch_bounds__(kThreads) sum_partials_kernel(const float* __restrict__ partial,
                                                                int n, float* __restrict__ out, float mean_scale) {
  __shared__ float sh[32];
  float v = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += partial[i];
  const float t = block_sum(v, sh);
  if (threadIdx.x == 0) {
    out[0] = t;                 // sum over pixels
    out[1] = t * mean_scale;    // reduce_mean (FCN.py:334)
  }
}

__global__ void __launch_bounds__(kThreads) softmax_infer_kernel(const float* __restrict__ logits,
                                                                 float* __restrict__ prob,
                                                                 uint8_t* __restrict__ mask,
                                                                 uint8_t* __restrict__ argmax,
                                                                 int64_t npix, int C) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npix;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float* l = logits + i * C;
    float mx = l[0];
    int am = 0;
    for (int c = 1; c < C; ++c)
      if (l[c] > mx) { mx = l[c]; am = c; }       // strict '>': first index on ties (tf.argmax)
    if (argmax) argmax[i] = (uint8_t)am;
    if (!prob && !mask) continue;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(l[c] - mx);
    const float inv = 1.f / s;
    float p1 = 0.f;
    for (int c = 0; c < C; ++c) {
      const float p = expf(l[c] - mx) * inv;
      if (prob) prob[i * C + c] = p;
      if (c == 1) p1 = p;
    }
    if (mask) mask[i] = p1 > 0.5f ? 1 : 0;
  }
}

__global__ void __launch_bounds__(kThreads) confusion_kernel(const uint8_t* __restrict__ gt,
                                                             const uint8_t* __restrict__ pred,
                                                             unsigned long long* __restrict__ cm,
                                                             int64_t npix) {
  __shared__ unsigned int cm_sh[4];
  if (threadIdx.x < 4) cm_sh[threadIdx.x] = 0;
  __syncthreads();
  unsigned int cnt[4] = {0, 0, 0, 0};
  // 16 pixels per thread per iteration (uint4 of u8) where aligned
  const int64_t n16 = npix >> 4;
  const uint4* g4 = reinterpret_cast<const uint4*>(gt);
  const uint4* p4 = reinterpret_cast<const uint4*>(pred);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n16;
       i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 g = __ldg(g4 + i), p = __ldg(p4 + i);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t gw = (&g.x)[j] & 0x01010101u, pw = (&p.x)[j] & 0x01010101u;
      const uint32_t both = gw & pw, gonly = gw & ~pw, ponly = pw & ~gw;
      const unsigned c11 = __popc(both), c10 = __popc(gonly), c01 = __popc(ponly);
      cnt[3] += c11; cnt[2] += c10; cnt[1] += c01; cnt[0] += 4 - c11 - c10 - c01;
    }
  }
  if (blockIdx.x == 0) {
    for (int64_t i = (n16 << 4) + threadIdx.x; i < npix; i += blockDim.x)
      cnt[(gt[i] & 1) * 2 + (pred[i] & 1)]++;
  }
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    unsigned int w = __reduce_add_sync(0xffffffffu, cnt[b]);
    if ((threadIdx.x & 31) == 0 && w) atomicAdd(&cm_sh[b], w);
  }
  __syncthreads();
  if (threadIdx.x < 4 && cm_sh[threadIdx.x])
    atomicAdd(&cm[threadIdx.x], (unsigned long long)cm_sh[threadIdx.x]);
}

// annotation placeholder [N,H,W,C] one-hot (bool / u8 / f32, FCN.py:313; channel 0 = background, FCN.py:195-201)
// -> u8 class ids: first maximal channel (what tf.argmax(annotation, 3) gives; exact for one-hot rows)
template <typename T>
__global__ void __launch_bounds__(kThreads) onehot_to_ids_kernel(const T* __restrict__ onehot, uint8_t* __restrict__ ids,
                                                                 int64_t npix, int C) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
    const T* r = onehot + i * C;
    float mx = (float)r[0];
    int am = 0;
    for (int c = 1; c < C; ++c) {
      const float v = (float)r[c];
      if (v > mx) { mx = v; am = c; }
    }
    ids[i] = (uint8_t)am;
  }
}

// paste_mask (FCN.py:203-211): where mask != 0 blend the RGBA colour over the image with PIL's
// integer arithmetic  t = dst*(255-a) + src*a ; out = ((t+128) + ((t+128) >> 8)) >> 8 ; else copy.
__global__ void __launch_bounds__(kThreads) overlay_kernel(const uint8_t* __restrict__ img,
                                                           const uint8_t* __restrict__ mask,
                                                           uint8_t* __restrict__ out, int64_t npix, int C,
                                                           uint32_t rgba) {
  const uint32_t a = rgba >> 24;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix;
       p += (int64_t)gridDim.x * blockDim.x) {
    const bool on = mask[p] != 0;
    for (int c = 0; c < C; ++c) {
      const uint32_t d = img[p * C + c];
      uint32_t v = d;
      if (on) {
        const uint32_t s = (rgba >> (8 * c)) & 0xffu;
        const uint32_t t = d * (255u - a) + s * a + 128u;
        v = (t + (t >> 8)) >> 8;
      }
      out[p * C + c] = (uint8_t)v;
    }
  }
}

// ------------------------------------------------------------------------------------
// optimizers (FCN.py:338-340).  28 B/param: p,m,v read+write (24) + g read (4).
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) adam_kernel(float* __restrict__ p, float* __restrict__ m,
                                                        float* __restrict__ v,
                                                        const float* __restrict__ g, int64_t n,
                                                        float lr_t, float b1, float b2, float eps,
                                                        float gs) {
  const int64_t n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  const float c1 = 1.f - b1, c2 = 1.f - b2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = __ldg(g4 + i);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // TF's ApplyAdam form: m += (g - m)(1 - b1); v += (g^2 - v)(1 - b2)
      const float gj = (&gg.x)[j] * gs;
      float mj = (&mm.x)[j] + (gj - (&mm.x)[j]) * c1;
      float vj = (&vv.x)[j] + (gj * gj - (&vv.x)[j]) * c2;
      (&mm.x)[j] = mj;
      (&vv.x)[j] = vj;
      (&pp.x)[j] -= lr_t * mj / (sqrtf(vj) + eps);
    }
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  if (blockIdx.x == 0) {
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      const float gj = g[i] * gs;
      const float mj = m[i] + (gj - m[i]) * c1;
      const float vj = v[i] + (gj * gj - v[i]) * c2;
      m[i] = mj; v[i] = vj;
      p[i] -= lr_t * mj / (sqrtf(vj) + eps);
    }
  }
}

// ApplyAdam on a LIST of arena ranges in one launch (the variables of a gradient bucket that are not conv weights of the
// tensor-core layers -- those take adam_pack_blocked_kernel): blockIdx.y = range, grid-stride over its elements.
constexpr int kMaxRanges = 64;
struct AdamRanges {
  int64_t off[kMaxRanges];
  int64_t len[kMaxRanges];
};
__global__ void __launch_bounds__(kThreads) adam_ranges_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                               const float* __restrict__ g, const AdamRanges r, float lr_t, float b1,
                                                               float b2, float eps, float gs) {
  const int64_t o = r.off[blockIdx.y], n = r.len[blockIdx.y];
  const float c1 = 1.f - b1, c2 = 1.f - b2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gj = g[o + i] * gs;
    const float mj = m[o + i] + (gj - m[o + i]) * c1;
    const float vj = v[o + i] + (gj * gj - v[o + i]) * c2;
    m[o + i] = mj;
    v[o + i] = vj;
    p[o + i] -= lr_t * mj / (sqrtf(vj) + eps);
  }
}

__global__ void __launch_bounds__(kThreads) momentum_kernel(float* __restrict__ p,
                                                            float* __restrict__ a,
                                                            const float* __restrict__ g, int64_t n,
                                                            float lr, float mu, float gs) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float aj = mu * a[i] + g[i] * gs;
    a[i] = aj;
    p[i] -= lr * aj;
  }
}

// ------------------------------------------------------------------------------------
// casts, bias-grad, weight packing
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) cast_bf16_kernel(const T* __restrict__ x,
                                                             bf16* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    y[i] = f2bf((float)x[i]);
}

// db[c] += sum_r dy[r][c].  Block = 256 threads = (256/CT channel-threads) ... generic:
// thread t handles channel (blockIdx.y*blockDim.x + t), block x handles a row chunk.
template <typename T>
__global__ void __launch_bounds__(kThreads) bias_grad_kernel(const T* __restrict__ dy,
                                                             float* __restrict__ db, int64_t rows,
                                                             int C, int64_t rows_per_block) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int64_t r0 = blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc = 0.f;
  for (int64_t r = r0; r < r1; ++r) acc += (float)dy[r * C + c];
  atomicAdd(db + c, acc);
}

// Vectorised BiasAddGrad for bf16 [rows][C], C % 8 == 0: thread = 8 channels of one row per
// iteration (16-byte loads), block-level reduce over the row lanes, one atomicAdd per channel/block.
__global__ void __launch_bounds__(kThreads) bias_grad_bf16x8_kernel(const uint4* __restrict__ dy,
                                                                     float* __restrict__ db, int64_t rows,
                                                                     int C8, int P8) {
  __shared__ float sh[kThreads][9];
  const int cpb = C8 < kThreads ? C8 : kThreads;       // column groups handled by this block
  const int R = kThreads / cpb;                         // row lanes
  const int cg = blockIdx.y * cpb + (threadIdx.x % cpb);
  const int rl = threadIdx.x / cpb;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (cg < C8 && rl < R) {
    const int64_t step = (int64_t)gridDim.x * R;
    int64_t r = (int64_t)blockIdx.x * R + rl;
    // four independent 16-byte loads in flight per thread (one load per iteration ran at 2 TB/s)
    for (; r + 3 * step < rows; r += 4 * step) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = __ldg(dy + (r + k * step) * P8 + cg);
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2((&u[k].x)[j]);
          acc[2 * j] += f.x;
          acc[2 * j + 1] += f.y;
        }
    }
    for (; r < rows; r += step) {
      const uint4 u = __ldg(dy + r * P8 + cg);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&u.x)[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[threadIdx.x][j] = acc[j];
  __syncthreads();
  if (rl == 0 && cg < C8) {
    for (int k = 1; k < R; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += sh[threadIdx.x + k * cpb][j];
    // per-block partial row: db is [gridDim.x][C] here (summed in fixed order by reduce_rows_kernel).
    // Atomics onto C addresses from hundreds of blocks serialise in L2 (~0.5 us per link of the chain).
    float* out = db + (int64_t)blockIdx.x * (C8 * 8) + cg * 8;
    *reinterpret_cast<float4*>(out) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(out + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// out[c] = sum_r part[r][c] in a fixed order: block = 32 channels x 8 row lanes (independent
// accumulators keep 4 loads in flight per thread), then a shared-memory tree over the row lanes.
__global__ void __launch_bounds__(kThreads) reduce_rows_kernel(const float* __restrict__ part,
                                                                float* __restrict__ out, int rows, int C) {
  __shared__ float sh[8][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  if (c < C) {
    int r = rl;
    for (; r + 56 < rows; r += 64) {        // eight loads in flight per thread (this small kernel is L2-latency bound)
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += part[(int64_t)(r + 8 * j) * C + c];
    }
    for (; r < rows; r += 8) a[0] += part[(int64_t)r * C + c];
  }
  sh[rl][cl] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  __syncthreads();
  if (rl == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][cl];
    out[c] = t;
  }
}

// BiasAddGrad for fp32 [rows][C] with C in {1,2,4} (the logits gradient): float4 loads.
__global__ void __launch_bounds__(kThreads) bias_grad_f32_small_kernel(const float4* __restrict__ dy,
                                                                        float* __restrict__ db, int64_t n4, int C) {
  __shared__ float sh[32];
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(dy + i);
    a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
  }
  // element j of every float4 belongs to channel j % C
  if (C == 1) { a[0] += a[1] + a[2] + a[3]; }
  else if (C == 2) { a[0] += a[2]; a[1] += a[3]; }
  for (int c = 0; c < C; ++c) {
    const float t = block_sum(a[c], sh);
    if (threadIdx.x == 0) atomicAdd(db + c, t);
  }
}

// Kernel-layout weights are BLOCKED by 64-wide k-chunks: an operand [T taps][rows][K] is stored as
// [T][ceil(K/64)][rows][64] (zero-padded in K), so the tile of one (tap, k-chunk) is one contiguous run of
// rows x 128 bytes (tcconv.cu: encode_weight_map_blocked).
//   w fp32 [T][A][B]  ->  cp bf16 [T'][ceil(B/64)][A][64]   (rows = A, k = B; T' = T-1-t when rev_cp: rot180 for dgrad)
//                     ->  tr bf16 [T ][ceil(A/64)][B][64]   (rows = B, k = A: the per-tap transpose)
// 64 x 64 tiles through shared memory, 16-byte loads and 16-byte bf16x8 stores on both outputs: every tile of
// either output is a contiguous 8 KB block.  A, B arbitrary (edges guarded, padding written as zeros).
__global__ void __launch_bounds__(kThreads) pack_blocked_kernel(const float* __restrict__ w, bf16* __restrict__ cp,
                                                                bf16* __restrict__ tr, int A, int B, int rev_cp) {
  __shared__ float tile[64][65];
  const int t = blockIdx.z;
  const int tc = rev_cp ? (int)gridDim.z - 1 - t : t;
  const int a0 = blockIdx.y * 64, b0 = blockIdx.x * 64;
  const int KCa = (int)gridDim.y, KCb = (int)gridDim.x;
  const float* wt = w + (int64_t)t * A * B;
  {
    const int c4 = threadIdx.x & 15, r0 = threadIdx.x >> 4;      // 16 float4 per row, 16 rows per pass
    const bool vec = (B & 3) == 0 && ((reinterpret_cast<uintptr_t>(w) & 15) == 0);
#pragma unroll
    for (int r = r0; r < 64; r += 16) {
      const int a = a0 + r, b = b0 + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a < A) {
        if (vec && b + 3 < B) {
          v = *reinterpret_cast<const float4*>(wt + (int64_t)a * B + b);
        } else {
          if (b < B) v.x = wt[(int64_t)a * B + b];
          if (b + 1 < B) v.y = wt[(int64_t)a * B + b + 1];
          if (b + 2 < B) v.z = wt[(int64_t)a * B + b + 2];
          if (b + 3 < B) v.w = wt[(int64_t)a * B + b + 3];
        }
      }
      tile[r][c4 * 4] = v.x; tile[r][c4 * 4 + 1] = v.y; tile[r][c4 * 4 + 2] = v.z; tile[r][c4 * 4 + 3] = v.w;
    }
  }
  __syncthreads();
  const int pc = threadIdx.x & 7, q0 = threadIdx.x >> 3;         // 8 bf16x8 pieces per row, 32 rows per pass
  if (cp) {
    bf16* dst = cp + ((int64_t)tc * KCb + blockIdx.x) * A * 64;
#pragma unroll
    for (int r = q0; r < 64; r += 32) {
      if (a0 + r >= A) continue;
      const float* s = &tile[r][pc * 8];
      *reinterpret_cast<uint4*>(dst + (int64_t)(a0 + r) * 64 + pc * 8) =
          make_uint4(pack_bf16x2(s[0], s[1]), pack_bf16x2(s[2], s[3]), pack_bf16x2(s[4], s[5]), pack_bf16x2(s[6], s[7]));
    }
  }
  if (tr) {
    bf16* dst = tr + ((int64_t)t * KCa + blockIdx.y) * B * 64;
#pragma unroll
    for (int r = q0; r < 64; r += 32) {       // output row b0 + r holds A-values a0 + pc*8 .. +7
      if (b0 + r >= B) continue;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = tile[pc * 8 + j][r];
      *reinterpret_cast<uint4*>(dst + (int64_t)(b0 + r) * 64 + pc * 8) =
          make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
  }
}

// ApplyAdam (adam_kernel's arithmetic) and the bf16 repack of a conv layer's weights in ONE pass: a 64 x 64 tile of
// p, m, v, g is read once, updated, written back, and its bf16 image goes to both kernel layouts (cp: k = B, tr: k = A,
// as pack_blocked_kernel).  Saves the packer's second read of the fresh parameters (4 B/param) and one launch per layer.
__global__ void __launch_bounds__(kThreads) adam_pack_blocked_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                                     const float* __restrict__ g, bf16* __restrict__ cp,
                                                                     bf16* __restrict__ tr, int A, int B, int rev_cp, float lr_t,
                                                                     float b1, float b2, float eps, float gs,
                                                                     const float* __restrict__ col_scale, float col_mult) {
  __shared__ float tile[64][65];
  const int t = blockIdx.z;
  const int tc = rev_cp ? (int)gridDim.z - 1 - t : t;
  const int a0 = blockIdx.y * 64, b0 = blockIdx.x * 64;
  const int KCa = (int)gridDim.y, KCb = (int)gridDim.x;
  const int64_t base = (int64_t)t * A * B;
  const float c1 = 1.f - b1, c2 = 1.f - b2;
  {
    const int c4 = threadIdx.x & 15, r0 = threadIdx.x >> 4;      // A % 64 == 0 and B % 64 == 0: 16-byte accesses, no edges
    // folded BN scale of this thread's four columns (the packed copies only; p stays the unscaled master weight)
    // ((p * gamma) * mult, the rounding order of segk_scale_columns: the fused and the separate path agree bit for bit)
    float4 cs = make_float4(1.f, 1.f, 1.f, 1.f);
    float cm = 1.f;
    if (col_scale) {
      cs = __ldg(reinterpret_cast<const float4*>(col_scale + b0 + c4 * 4));
      cm = col_mult;
    }
#pragma unroll
    for (int r = r0; r < 64; r += 16) {
      const int64_t o = base + (int64_t)(a0 + r) * B + b0 + c4 * 4;
      float4 pp = *reinterpret_cast<const float4*>(p + o), mm = *reinterpret_cast<const float4*>(m + o);
      float4 vv = *reinterpret_cast<const float4*>(v + o);
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g + o));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gj = (&gg.x)[j] * gs;
        const float mj = (&mm.x)[j] + (gj - (&mm.x)[j]) * c1;
        const float vj = (&vv.x)[j] + (gj * gj - (&vv.x)[j]) * c2;
        (&mm.x)[j] = mj;
        (&vv.x)[j] = vj;
        (&pp.x)[j] -= lr_t * mj / (sqrtf(vj) + eps);
      }
      *reinterpret_cast<float4*>(p + o) = pp;
      *reinterpret_cast<float4*>(m + o) = mm;
      *reinterpret_cast<float4*>(v + o) = vv;
      tile[r][c4 * 4] = pp.x * cs.x * cm; tile[r][c4 * 4 + 1] = pp.y * cs.y * cm; tile[r][c4 * 4 + 2] = pp.z * cs.z * cm;
      tile[r][c4 * 4 + 3] = pp.w * cs.w * cm;
    }
  }
  __syncthreads();
  const int pc = threadIdx.x & 7, q0 = threadIdx.x >> 3;
  if (cp) {
    bf16* dst = cp + ((int64_t)tc * KCb + blockIdx.x) * A * 64;
#pragma unroll
    for (int r = q0; r < 64; r += 32) {
      const float* s = &tile[r][pc * 8];
      *reinterpret_cast<uint4*>(dst + (int64_t)(a0 + r) * 64 + pc * 8) =
          make_uint4(pack_bf16x2(s[0], s[1]), pack_bf16x2(s[2], s[3]), pack_bf16x2(s[4], s[5]), pack_bf16x2(s[6], s[7]));
    }
  }
  if (tr) {
    bf16* dst = tr + ((int64_t)t * KCa + blockIdx.y) * B * 64;
#pragma unroll
    for (int r = q0; r < 64; r += 32) {
      float vv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) vv[j] = tile[pc * 8 + j][r];
      *reinterpret_cast<uint4*>(dst + (int64_t)(b0 + r) * 64 + pc * 8) =
          make_uint4(pack_bf16x2(vv[0], vv[1]), pack_bf16x2(vv[2], vv[3]), pack_bf16x2(vv[4], vv[5]), pack_bf16x2(vv[6], vv[7]));
    }
  }
}

// deconv forward phase packing: wk[ay*s+ax][uy*2+ux][co][ci] = w[ay+s*(1-uy)][ax+s*(1-ux)][co][ci], blocked over ci:
// wk[phase*4 + u][ci/64][co][ci%64]
__global__ void __launch_bounds__(kThreads) pack_deconv_phase_kernel(const float* __restrict__ w,
                                                                     bf16* __restrict__ wk, int k, int s,
                                                                     int Cin, int Cout) {
  const int KC = (Cin + 63) / 64;
  const int64_t per = (int64_t)KC * Cout * 64;
  const int64_t total = (int64_t)s * s * 4 * per;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i % per;
    const int pu = (int)(i / per);
    const int kk = (int)(e % 64), co = (int)((e / 64) % Cout), kc = (int)(e / (64 * (int64_t)Cout));
    const int ci = kc * 64 + kk;
    const int u = pu & 3, ph = pu >> 2;
    const int uy = u >> 1, ux = u & 1, ay = ph / s, ax = ph % s;
    const int ky = ay + s * (1 - uy), kx = ax + s * (1 - ux);
    wk[i] = ci < Cin ? f2bf(w[((int64_t)(ky * k + kx) * Cout + co) * Cin + ci]) : f2bf(0.f);
  }
}

}  // namespace

extern "C" {

int segk_maxpool2x2_fwd(segk_ctx* ctx, const void* x, void* y, uint8_t* idx, int N, int H, int W,
                        int C, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  const SegkPitch pitch = segk_take_pitch(ctx);
  SEGK_REQUIRE(ctx, pitch.out == 0 && (pitch.in == 0 || pitch.in >= C), "maxpool_fwd: only x may be pitched (>= %d channels)", C);
  const int P8 = (pitch.in > 0 ? pitch.in : C) / 8;
  SEGK_REQUIRE(ctx, x && y && idx, "maxpool_fwd: null pointer");
  SEGK_REQUIRE(ctx, N > 0 && H >= 2 && W >= 2 && (H % 2 == 0) && (W % 2 == 0) && C % 8 == 0,
               "maxpool_fwd: need even H,W and C%%8==0 (got %dx%dx%dx%d)", N, H, W, C);
  const int64_t items = (int64_t)N * (H / 2) * (W / 2) * (C / 8);
  maxpool_fwd_kernel<<<stream_grid(ctx, items), kThreads, 0, (cudaStream_t)stream>>>(
      (const uint4*)x, (uint4*)y, (uint2*)idx, N, H, W, C / 8, P8);
  SEGK_LAUNCHED(ctx, "maxpool_fwd");
  return SEGK_OK;
}

int segk_maxpool2x2_bwd(segk_ctx* ctx, const void* dy, const uint8_t* idx, const void* act, int act_is_pooled,
                        const void* residual, void* dx, float* dbias, int N, int H, int W, int C, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  const SegkPitch pitch = segk_take_pitch(ctx);
  SEGK_REQUIRE(ctx, pitch.in == 0 && (pitch.out == 0 || pitch.out >= C), "maxpool_bwd: only dx (with act / residual) may be pitched");
  const int P8 = (pitch.out > 0 ? pitch.out : C) / 8;
  SEGK_REQUIRE(ctx, P8 == C / 8 || !(act && act_is_pooled), "maxpool_bwd: the pooled-mask mode writes a dense dx");
  SEGK_REQUIRE(ctx, dy && idx && dx, "maxpool_bwd: null pointer");
  SEGK_REQUIRE(ctx, N > 0 && H >= 2 && W >= 2 && (H % 2 == 0) && (W % 2 == 0) && C % 8 == 0,
               "maxpool_bwd: need even H,W and C%%8==0 (got %dx%dx%dx%d)", N, H, W, C);
  SEGK_REQUIRE(ctx, !dbias || (act && act_is_pooled), "maxpool_bwd: dbias needs the pooled-activation mask mode");
  const int64_t items = (int64_t)N * (H / 2) * (W / 2) * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (act && act_is_pooled) {
    SEGK_REQUIRE(ctx, !residual, "maxpool_bwd: a second gradient path needs the full-resolution activation as mask");
    int grid = stream_grid(ctx, items);
    if (dbias) {
      // one partial row per block: fewer, longer-lived blocks keep the row reduction short (1184 rows cost ~35 us)
      if (grid > ctx->sm_count * 3) grid = ctx->sm_count * 3;
      const int C8 = C / 8;
      SEGK_REQUIRE(ctx, C8 <= kThreads && kThreads % C8 == 0, "maxpool_bwd: dbias needs C/8 to divide %d (got C = %d)", kThreads, C);
      const int rc = segk_grow(ctx, &ctx->ws5, &ctx->ws5_bytes, sizeof(float) * (size_t)ctx->sm_count * 8 * 2048, "pool-bwd bias partials");
      if (rc) return rc;
      maxpool_bwd_pooled_kernel<true><<<grid, kThreads, 0, st>>>((const uint4*)dy, (const uint2*)idx, (const uint4*)act,
                                                                 (uint4*)dx, (float*)ctx->ws5, N, H, W, C8);
      SEGK_LAUNCHED(ctx, "maxpool_bwd_pooled_bias");
      reduce_rows_kernel<<<ceil_div(C, 32), kThreads, 0, st>>>((const float*)ctx->ws5, dbias, grid, C);
      SEGK_LAUNCHED(ctx, "maxpool_bwd_bias_reduce");
      return SEGK_OK;
    }
    maxpool_bwd_pooled_kernel<false><<<grid, kThreads, 0, st>>>((const uint4*)dy, (const uint2*)idx, (const uint4*)act,
                                                                (uint4*)dx, nullptr, N, H, W, C / 8);
    SEGK_LAUNCHED(ctx, "maxpool_bwd_pooled");
    return SEGK_OK;
  }
  maxpool_bwd_kernel<<<stream_grid(ctx, items), kThreads, 0, st>>>(
      (const uint4*)dy, (const uint2*)idx, (const uint4*)act, (const uint4*)residual, (uint4*)dx, N, H, W, C / 8, P8);
  SEGK_LAUNCHED(ctx, "maxpool_bwd");
  return SEGK_OK;
}

int segk_dropout(segk_ctx* ctx, const void* x, void* y, const uint8_t* mask, int64_t n,
                 float keep_prob, uint64_t seed, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && y && n > 0, "dropout: null pointer / empty");
  SEGK_REQUIRE(ctx, keep_prob > 0.f && keep_prob <= 1.f, "dropout: keep_prob %f out of (0,1]",
               keep_prob);
  if (n % 8 == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0 && (((uintptr_t)mask) & 7) == 0) {
    dropout_vec8_kernel<<<stream_grid(ctx, n / 8), kThreads, 0, (cudaStream_t)stream>>>(
        (const uint4*)x, (uint4*)y, (const uint2*)mask, n / 8, keep_prob, 1.0f / keep_prob, seed);
    SEGK_LAUNCHED(ctx, "dropout_vec8");
    return SEGK_OK;
  }
  dropout_kernel<<<stream_grid(ctx, (n + 7) / 8), kThreads, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, (bf16*)y, mask, n, keep_prob, 1.0f / keep_prob, seed);
  SEGK_LAUNCHED(ctx, "dropout");
  return SEGK_OK;
}

size_t segk_xent_workspace_bytes(int64_t npix) {
  (void)npix;
  return (size_t)kXentBlocks * sizeof(float) + 64;
}

int segk_softmax_xent_fwd_bwd(segk_ctx* ctx, const float* logits, const uint8_t* labels,
                              float* dlogits, uint8_t* pred, float* loss_sum, int64_t* cm,
                              void* workspace, int64_t npix, int C, float grad_scale, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, logits && labels && loss_sum && workspace, "xent: null pointer");
  SEGK_REQUIRE(ctx, npix > 0 && C >= 2 && C <= 64, "xent: bad npix/C (%lld, %d)", (long long)npix, C);
  SEGK_REQUIRE(ctx, cm == nullptr || C == 2, "xent: confusion counts need C == 2");
  int blocks = (int)ceil_div64(npix, kThreads);
  if (blocks > kXentBlocks) blocks = kXentBlocks;
  float* partial = (float*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 2)
    xent_kernel<2><<<blocks, kThreads, 0, st>>>(logits, labels, dlogits, pred, partial,
                                               (unsigned long long*)cm, npix, C, grad_scale);
  else
    xent_kernel<0><<<blocks, kThreads, 0, st>>>(logits, labels, dlogits, pred, partial,
                                               (unsigned long long*)cm, npix, C, grad_scale);
  SEGK_LAUNCHED(ctx, "xent");
  sum_partials_kernel<<<1, kThreads, 0, st>>>(partial, blocks, loss_sum, 1.0f / (float)npix);
  SEGK_LAUNCHED(ctx, "xent_sum");
  return SEGK_OK;
}

int segk_softmax_infer(segk_ctx* ctx, const float* logits, float* prob, uint8_t* mask, uint8_t* argmax,
                       int64_t npix, int C, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, logits && (prob || mask || argmax) && npix > 0 && C >= 2 && C <= 255, "softmax_infer: bad args");
  softmax_infer_kernel<<<stream_grid(ctx, npix), kThreads, 0, (cudaStream_t)stream>>>(
      logits, prob, mask, argmax, npix, C);
  SEGK_LAUNCHED(ctx, "softmax_infer");
  return SEGK_OK;
}

int segk_onehot_to_ids(segk_ctx* ctx, const void* onehot, int dtype, uint8_t* ids, int64_t npix, int C, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, onehot && ids && npix > 0 && C >= 1 && C <= 255, "onehot_to_ids: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == 1)
    onehot_to_ids_kernel<float><<<stream_grid(ctx, npix), kThreads, 0, st>>>((const float*)onehot, ids, npix, C);
  else if (dtype == 2)
    onehot_to_ids_kernel<uint8_t><<<stream_grid(ctx, npix), kThreads, 0, st>>>((const uint8_t*)onehot, ids, npix, C);
  else
    return segk_fail(ctx, SEGK_EINVAL, "onehot_to_ids: dtype must be 1 (f32) or 2 (u8 / bool)");
  SEGK_LAUNCHED(ctx, "onehot_to_ids");
  return SEGK_OK;
}

int segk_overlay_mask(segk_ctx* ctx, const uint8_t* image, const uint8_t* mask, uint8_t* out, int64_t npix,
                      int C, int r, int g, int b, int a, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, image && mask && out && npix > 0 && (C == 3 || C == 4), "overlay: bad args (C must be 3 or 4)");
  const uint32_t rgba = (uint32_t)(r & 255) | ((uint32_t)(g & 255) << 8) | ((uint32_t)(b & 255) << 16) |
                        ((uint32_t)(a & 255) << 24);
  overlay_kernel<<<stream_grid(ctx, npix), kThreads, 0, (cudaStream_t)stream>>>(image, mask, out, npix, C, rgba);
  SEGK_LAUNCHED(ctx, "overlay");
  return SEGK_OK;
}

int segk_confusion_matrix(segk_ctx* ctx, const uint8_t* gt, const uint8_t* pred, int64_t* cm,
                          int64_t npix, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, gt && pred && cm && npix > 0, "confusion: bad args");
  SEGK_REQUIRE(ctx, (((uintptr_t)gt | (uintptr_t)pred) & 15) == 0, "confusion: 16-byte alignment");
  // 2 blocks / SM: each block ends with 4 global atomics on the same 4 counters (serialised in L2)
  confusion_kernel<<<stream_grid(ctx, (npix + 15) / 16, 2), kThreads, 0, (cudaStream_t)stream>>>(
      gt, pred, (unsigned long long*)cm, npix);
  SEGK_LAUNCHED(ctx, "confusion");
  return SEGK_OK;
}

int segk_adam_step(segk_ctx* ctx, float* p, float* m, float* v, const float* g, int64_t n, float lr_t,
                   float beta1, float beta2, float eps, float grad_scale, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, p && m && v && g && n > 0, "adam: bad args");
  SEGK_REQUIRE(ctx, (((uintptr_t)p | (uintptr_t)m | (uintptr_t)v | (uintptr_t)g) & 15) == 0,
               "adam: arenas must be 16-byte aligned");
  adam_kernel<<<stream_grid(ctx, n / 4 + 1), kThreads, 0, (cudaStream_t)stream>>>(
      p, m, v, g, n, lr_t, beta1, beta2, eps, grad_scale);
  SEGK_LAUNCHED(ctx, "adam");
  return SEGK_OK;
}

int segk_momentum_step(segk_ctx* ctx, float* p, float* a, const float* g, int64_t n, float lr,
                       float mu, float grad_scale, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, p && a && g && n > 0, "momentum: bad args");
  momentum_kernel<<<stream_grid(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(p, a, g, n, lr, mu,
                                                                             grad_scale);
  SEGK_LAUNCHED(ctx, "momentum");
  return SEGK_OK;
}

int segk_cast_to_bf16(segk_ctx* ctx, const void* x, int x_dtype, void* y, int64_t n, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && y && n > 0, "cast: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == 1)
    cast_bf16_kernel<float><<<stream_grid(ctx, n), kThreads, 0, st>>>((const float*)x, (bf16*)y, n);
  else if (x_dtype == 2)
    cast_bf16_kernel<uint8_t><<<stream_grid(ctx, n), kThreads, 0, st>>>((const uint8_t*)x, (bf16*)y, n);
  else
    return segk_fail(ctx, SEGK_EINVAL, "cast: x_dtype must be 1 (f32) or 2 (u8)");
  SEGK_LAUNCHED(ctx, "cast");
  return SEGK_OK;
}

int segk_bias_grad(segk_ctx* ctx, const void* dy, int dy_is_f32, float* db, int64_t rows, int C,
                   void* stream) {
  if (!ctx) return SEGK_EINVAL;
  const SegkPitch pitch = segk_take_pitch(ctx);
  SEGK_REQUIRE(ctx, pitch.out == 0 && (pitch.in == 0 || pitch.in >= C), "bias_grad: only dy may be pitched (>= %d channels)", C);
  const int ld = pitch.in > 0 ? pitch.in : C;
  SEGK_REQUIRE(ctx, dy && db && rows > 0 && C > 0, "bias_grad: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (!dy_is_f32 && C % 8 == 0 && (((uintptr_t)dy) & 15) == 0) {
    const int C8 = C / 8;
    const int cpb = C8 < kThreads ? C8 : kThreads;
    if (kThreads % cpb == 0 && C8 % cpb == 0) {
      const int R = kThreads / cpb;
      const int gy = C8 / cpb;
      int64_t gx = ceil_div64(rows, (int64_t)R * 4);
      const int64_t cap = ceil_div64((int64_t)ctx->sm_count * 4, gy);
      if (gx > cap) gx = cap;
      if (gx < 1) gx = 1;
      const size_t need = sizeof(float) * (size_t)gx * C;
      {
        const int rc = segk_grow(ctx, &ctx->ws2, &ctx->ws2_bytes, need < (size_t)(8 << 20) ? (size_t)(8 << 20) : need, "bias_grad");
        if (rc) return rc;
      }
      bias_grad_bf16x8_kernel<<<dim3((unsigned)gx, gy), kThreads, 0, st>>>((const uint4*)dy, (float*)ctx->ws2, rows, C8, ld / 8);
      SEGK_LAUNCHED(ctx, "bias_grad_bf16x8");
      reduce_rows_kernel<<<ceil_div(C, 32), kThreads, 0, st>>>((const float*)ctx->ws2, db, (int)gx, C);
      SEGK_LAUNCHED(ctx, "bias_grad_reduce");
      return SEGK_OK;
    }
  }
  SEGK_REQUIRE(ctx, ld == C, "bias_grad: a pitched dy needs bf16 with C %% 8 == 0 and C/8 dividing (or a multiple of) %d", kThreads);
  cudaError_t e = cudaMemsetAsync(db, 0, sizeof(float) * C, st);
  if (e != cudaSuccess) return segk_fail(ctx, SEGK_ECUDA, "bias_grad memset: %s", cudaGetErrorString(e));
  if (dy_is_f32 && (C == 1 || C == 2 || C == 4) && (rows * C) % 4 == 0 && (((uintptr_t)dy) & 15) == 0) {
    const int64_t n4 = rows * C / 4;
    bias_grad_f32_small_kernel<<<stream_grid(ctx, n4, 4), kThreads, 0, st>>>((const float4*)dy, db, n4, C);
    SEGK_LAUNCHED(ctx, "bias_grad_f32_small");
    return SEGK_OK;
  }
  const int tx = C < kThreads ? ((C + 31) / 32) * 32 : kThreads;
  const int gy = ceil_div(C, tx);
  int64_t want = (int64_t)ctx->sm_count * 8 / gy;
  if (want < 1) want = 1;
  int64_t rpb = ceil_div64(rows, want);
  if (rpb < 16) rpb = 16;
  dim3 grid((unsigned)ceil_div64(rows, rpb), gy);
  if (dy_is_f32)
    bias_grad_kernel<float><<<grid, tx, 0, st>>>((const float*)dy, db, rows, C, rpb);
  else
    bias_grad_kernel<bf16><<<grid, tx, 0, st>>>((const bf16*)dy, db, rows, C, rpb);
  SEGK_LAUNCHED(ctx, "bias_grad");
  return SEGK_OK;
}

int segk_pack_conv_weights(segk_ctx* ctx, const float* w, void* wk, void* wd, int kh, int kw, int Cin,
                           int Cout, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, w && (wk || wd) && kh > 0 && kw > 0 && Cin > 0 && Cout > 0, "pack_conv: bad args");
  SEGK_REQUIRE(ctx, (((uintptr_t)wk | (uintptr_t)wd) & 15) == 0, "pack_conv: outputs must be 16-byte aligned");
  const int T = kh * kw;
  // w[t][Cin][Cout]: wd = cast copy with taps reversed (k = Cout), wk = per-tap transpose (k = Cin); both blocked
  dim3 grid(ceil_div(Cout, 64), ceil_div(Cin, 64), T);
  SEGK_REQUIRE(ctx, grid.y <= 65535 && grid.z <= 65535, "pack_conv: dims too large");
  pack_blocked_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(w, (bf16*)wd, (bf16*)wk, Cin, Cout, 1);
  SEGK_LAUNCHED(ctx, "pack_conv_weights");
  return SEGK_OK;
}

int segk_adam_step_ranges(segk_ctx* ctx, float* p, float* m, float* v, const float* g, const int64_t* offsets,
                          const int64_t* lengths, int nranges, float lr_t, float beta1, float beta2, float eps, float grad_scale,
                          void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, p && m && v && g && offsets && lengths && nranges > 0, "adam_ranges: bad args");
  for (int r0 = 0; r0 < nranges; r0 += kMaxRanges) {
    AdamRanges r;
    memset(&r, 0, sizeof(r));
    const int nr = nranges - r0 < kMaxRanges ? nranges - r0 : kMaxRanges;
    int64_t longest = 0;
    for (int i = 0; i < nr; ++i) {
      SEGK_REQUIRE(ctx, offsets[r0 + i] >= 0 && lengths[r0 + i] >= 0, "adam_ranges: negative range");
      r.off[i] = offsets[r0 + i];
      r.len[i] = lengths[r0 + i];
      if (r.len[i] > longest) longest = r.len[i];
    }
    int64_t gx = ceil_div64(longest, kThreads);
    if (gx > 2 * ctx->sm_count) gx = 2 * ctx->sm_count;
    if (gx < 1) gx = 1;
    adam_ranges_kernel<<<dim3((unsigned)gx, nr), kThreads, 0, (cudaStream_t)stream>>>(p, m, v, g, r, lr_t, beta1, beta2, eps, grad_scale);
    SEGK_LAUNCHED(ctx, "adam_ranges");
  }
  return SEGK_OK;
}

int segk_adam_pack_conv_weights(segk_ctx* ctx, float* p, float* m, float* v, const float* g, void* wk, void* wd,
                                const float* col_scale, float col_mult, int kh, int kw, int Cin, int Cout, float lr_t,
                                float beta1, float beta2, float eps, float grad_scale, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, p && m && v && g && (wk || wd) && kh > 0 && kw > 0, "adam_pack_conv: bad args");
  SEGK_REQUIRE(ctx, Cin % 64 == 0 && Cout % 64 == 0 && Cin > 0 && Cout > 0, "adam_pack_conv: needs Cin, Cout multiples of 64 (got %d, %d)", Cin, Cout);
  SEGK_REQUIRE(ctx, (((uintptr_t)p | (uintptr_t)m | (uintptr_t)v | (uintptr_t)g | (uintptr_t)wk | (uintptr_t)wd | (uintptr_t)col_scale) & 15) == 0,
               "adam_pack_conv: 16-byte alignment");
  dim3 grid(Cout / 64, Cin / 64, kh * kw);
  SEGK_REQUIRE(ctx, grid.y <= 65535 && grid.z <= 65535, "adam_pack_conv: dims too large");
  adam_pack_blocked_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(p, m, v, g, (bf16*)wd, (bf16*)wk, Cin, Cout, 1, lr_t, beta1,
                                                                       beta2, eps, grad_scale, col_scale, col_mult);
  SEGK_LAUNCHED(ctx, "adam_pack_conv_weights");
  return SEGK_OK;
}

int segk_pack_matrix(segk_ctx* ctx, const float* w, void* cp, void* tr, int T, int A, int B, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, w && (cp || tr) && T > 0 && A > 0 && B > 0, "pack_matrix: bad args");
  SEGK_REQUIRE(ctx, (((uintptr_t)cp | (uintptr_t)tr) & 15) == 0, "pack_matrix: outputs must be 16-byte aligned");
  dim3 grid(ceil_div(B, 64), ceil_div(A, 64), T);
  SEGK_REQUIRE(ctx, grid.y <= 65535 && grid.z <= 65535, "pack_matrix: dims too large");
  pack_blocked_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(w, (bf16*)cp, (bf16*)tr, A, B, 0);
  SEGK_LAUNCHED(ctx, "pack_matrix");
  return SEGK_OK;
}

int segk_pack_deconv_weights(segk_ctx* ctx, const float* w, void* wk, void* wd, int k, int s, int Cin,
                             int Cout, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, w && (wk || wd) && k == 2 * s && s > 0 && Cin > 0 && Cout > 0, "pack_deconv: bad args");
  SEGK_REQUIRE(ctx, (((uintptr_t)wk | (uintptr_t)wd) & 15) == 0, "pack_deconv: outputs must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (wd) {
    // w[t][Cout][Cin] -> wd[t][Cout/64][Cin][64]  (rows = Cin, k = Cout)
    dim3 grid(ceil_div(Cin, 64), ceil_div(Cout, 64), k * k);
    SEGK_REQUIRE(ctx, grid.y <= 65535 && grid.z <= 65535, "pack_deconv: dims too large");
    pack_blocked_kernel<<<grid, kThreads, 0, st>>>(w, nullptr, (bf16*)wd, Cout, Cin, 0);
    SEGK_LAUNCHED(ctx, "pack_deconv_wd");
  }
  if (wk) {
    const int64_t n = (int64_t)s * s * 4 * ceil_div(Cin, 64) * 64 * Cout;
    pack_deconv_phase_kernel<<<stream_grid(ctx, n), kThreads, 0, st>>>(w, (bf16*)wk, k, s, Cin, Cout);
    SEGK_LAUNCHED(ctx, "pack_deconv_wk");
  }
  return SEGK_OK;
}

}  // extern "C"
