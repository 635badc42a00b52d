// HBM-bound kernels of the dense-block builders (Network/model/FCDenseNet.py: bottleneck_layer :23-35,
// Transition_Layer :37-46) that the VGG-style models do not need:
//   * pre-activation  Batch_Normalization (inference-mode affine, utils.py:300-301) + ReLU (utils.py:303) on a
//     channel PREFIX of a concat buffer (the zero-copy view of Concat(layers_concat), FCDenseNet.py:55) and its
//     gradient (dx accumulated into the concat buffer's gradient, d gamma / d beta by a two-stage reduction);
//   * Avg_Pooling 2x2 / stride 2 / VALID (utils.py:309) forward and backward;
//   * logical <-> physical remapping of weights and per-channel parameters.  Tensors are stored with their
//     channel segments padded to multiples of 8 and the total to a multiple of 64 (pads are always zero) so that
//     every conv is a tcgen05 GEMM on 64-wide channel chunks; `map[physical] = logical index or -1`.
// All kernels: 16-byte accesses over rows of `ld` elements (strided channel views), grid-stride loops.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

inline int sgrid(segk_ctx* ctx, int64_t items, int per_sm = 8) {
  int64_t b = ceil_div64(items, kThreads), cap = (int64_t)ctx->sm_count * per_sm;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

// Dropout (utils.py:318) of the kernel's INPUT tensor applied on the fly: in FCDenseNet every conv is followed by a
// Dropout whose only reader is the next BN/ReLU (bottleneck conv1) or the concat slot copy (conv2), so the dropped
// tensor is never materialised.  The keep pattern is segk_dropout's (Philox counter = element index in the
// [rows][ld] tensor, or the injected u8 mask), and the value is rounded to bf16 where segk_dropout would have
// stored it: both ways give the same bits.
struct Drop {
  int on;
  const uint2* mask;
  float keep, inv_keep;
  uint64_t seed;
};
__device__ __forceinline__ uint4 drop_apply(uint4 u, uint32_t kp, float inv_keep) {
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack_bf16x2((&u.x)[j]);
    o[j] = pack_bf16x2((kp >> (2 * j)) & 1u ? f.x * inv_keep : 0.f, (kp >> (2 * j + 1)) & 1u ? f.y * inv_keep : 0.f);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}
inline Drop make_drop(const uint8_t* mask, float keep_prob, uint64_t seed) {
  Drop d;
  d.on = (keep_prob > 0.f && keep_prob < 1.f) ? 1 : 0;
  d.mask = (const uint2*)mask;
  d.keep = keep_prob;
  d.inv_keep = d.on ? 1.0f / keep_prob : 1.f;
  d.seed = seed;
  return d;
}

// y[r][c] = act(x[r][c] * scale[c] + shift[c]) for c < C (C % 8 == 0); thread = 8 channels of one row
__global__ void __launch_bounds__(kThreads) bn_act_fwd_kernel(const bf16* __restrict__ x, int ldx, bf16* __restrict__ y,
                                                              int ldy, const float* __restrict__ scale,
                                                              const float* __restrict__ shift, int64_t rows, int C8,
                                                              int relu, Drop drop) {
  const int64_t total = rows * C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % C8);
    const int64_t r = i / C8;
    uint4 u = __ldg(reinterpret_cast<const uint4*>(x + r * ldx) + g);
    if (drop.on) u = drop_apply(u, segk_dropout_keep8(drop.mask, r * (ldx >> 3) + g, drop.keep, drop.seed), drop.inv_keep);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale) + 2 * g), s1 = __ldg(reinterpret_cast<const float4*>(scale) + 2 * g + 1);
    const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift) + 2 * g), h1 = __ldg(reinterpret_cast<const float4*>(shift) + 2 * g + 1);
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
    float v[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2((&u.x)[j]);
      v[2 * j] = f.x * sc[2 * j] + sh[2 * j];
      v[2 * j + 1] = f.y * sc[2 * j + 1] + sh[2 * j + 1];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    reinterpret_cast<uint4*>(y + r * ldy)[g] =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

// g = dy * [y > 0] (relu) ; dx (+)= g * scale ; per-block partial sums of g * x (-> d scale) and g (-> d shift).
// Every thread keeps the same 8 channels over its grid-stride iterations (the thread count is a multiple of C8).
__global__ void __launch_bounds__(kThreads) bn_act_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y, int ldy,
                                                              const bf16* __restrict__ x, bf16* __restrict__ dx, int ldx,
                                                              const float* __restrict__ scale, float* __restrict__ part,
                                                              int64_t rows, int C8, int relu, int accumulate, Drop drop) {
  __shared__ float sh[kThreads][17];
  const int cpb = C8 < kThreads ? C8 : kThreads;          // channel groups per block row
  const int R = kThreads / cpb;
  const int g = blockIdx.y * cpb + (threadIdx.x % cpb);
  const int rl = threadIdx.x / cpb;
  float ax[8], as[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ax[j] = as[j] = 0.f;
  if (g < C8 && rl < R) {
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale) + 2 * g), s1 = __ldg(reinterpret_cast<const float4*>(scale) + 2 * g + 1);
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    for (int64_t r = (int64_t)blockIdx.x * R + rl; r < rows; r += (int64_t)gridDim.x * R) {
      const uint4 ud = __ldg(reinterpret_cast<const uint4*>(dy + r * ldy) + g);
      const uint4 uy = __ldg(reinterpret_cast<const uint4*>(y + r * ldy) + g);
      uint4 ux = __ldg(reinterpret_cast<const uint4*>(x + r * ldx) + g);
      uint32_t kp = 0xffu;
      if (drop.on) {
        kp = segk_dropout_keep8(drop.mask, r * (ldx >> 3) + g, drop.keep, drop.seed);
        ux = drop_apply(ux, kp, drop.inv_keep);            // the BN input was the dropped tensor
      }
      uint4* dxp = reinterpret_cast<uint4*>(dx + r * ldx) + g;
      uint4 uo = make_uint4(0, 0, 0, 0);
      if (accumulate) uo = *dxp;
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fd = unpack_bf16x2((&ud.x)[j]), fy = unpack_bf16x2((&uy.x)[j]), fx = unpack_bf16x2((&ux.x)[j]);
        const float2 fo = unpack_bf16x2((&uo.x)[j]);
        const float g0 = (!relu || fy.x > 0.f) ? fd.x : 0.f, g1 = (!relu || fy.y > 0.f) ? fd.y : 0.f;
        ax[2 * j] += g0 * fx.x; ax[2 * j + 1] += g1 * fx.y;
        as[2 * j] += g0; as[2 * j + 1] += g1;
        float d0 = fo.x + g0 * sc[2 * j], d1 = fo.y + g1 * sc[2 * j + 1];
        if (drop.on) {      // DropoutGrad: the gradient reaching the conv in front of the dropout (never accumulated)
          d0 = (kp >> (2 * j)) & 1u ? bf2f(f2bf(d0)) * drop.inv_keep : 0.f;
          d1 = (kp >> (2 * j + 1)) & 1u ? bf2f(f2bf(d1)) * drop.inv_keep : 0.f;
        }
        o[j] = pack_bf16x2(d0, d1);
      }
      *dxp = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { sh[threadIdx.x][j] = ax[j]; sh[threadIdx.x][8 + j] = as[j]; }
  __syncthreads();
  if (rl == 0 && g < C8) {
    for (int k = 1; k < R; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) { ax[j] += sh[threadIdx.x + k * cpb][j]; as[j] += sh[threadIdx.x + k * cpb][8 + j]; }
    // partial rows [gridDim.x][2][C]: d scale then d shift
    float* o = part + (int64_t)blockIdx.x * (2 * C8 * 8) + g * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j] = ax[j]; o[C8 * 8 + j] = as[j]; }
  }
}

// out[c] = sum_r part[r][c] (fixed order), 32 columns x 8 row lanes per block
__global__ void __launch_bounds__(kThreads) dense_reduce_rows_kernel(const float* __restrict__ part, float* __restrict__ out0,
                                                                     float* __restrict__ out1, int rows, int C) {
  __shared__ float sh[8][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;          // over 2C columns
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  if (c < 2 * C) {
    int r = rl;
    for (; r + 56 < rows; r += 64) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += part[(int64_t)(r + 8 * j) * 2 * C + c];
    }
    for (; r < rows; r += 8) a[0] += part[(int64_t)r * 2 * C + c];
  }
  sh[rl][cl] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  __syncthreads();
  if (rl == 0 && c < 2 * C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][cl];
    if (c < C) out0[c] = t;
    else out1[c - C] = t;
  }
}

// Avg_Pooling 2x2 / s2 / VALID: y = mean of the window (fp32 sum, one rounding); thread = 8 channels of a pooled pixel
__global__ void __launch_bounds__(kThreads) avgpool_fwd_kernel(const bf16* __restrict__ x, int ldx, bf16* __restrict__ y,
                                                               int ldy, int N, int H, int W, int C8) {
  const int OH = H >> 1, OW = W >> 1;
  const int64_t total = (int64_t)N * OH * OW * C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % C8);
    int64_t p = i / C8;
    const int ox = (int)(p % OW);
    p /= OW;
    const int oy = (int)(p % OH);
    const int n = (int)(p / OH);
    const int64_t r0 = ((int64_t)n * H + 2 * oy) * W + 2 * ox;
    const int64_t rr[4] = {r0, r0 + 1, r0 + W, r0 + W + 1};
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + rr[k] * ldx) + g);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&u.x)[j]);
        a[2 * j] += f.x; a[2 * j + 1] += f.y;
      }
    }
    const int64_t ro = ((int64_t)n * OH + oy) * OW + ox;
    reinterpret_cast<uint4*>(y + ro * ldy)[g] = make_uint4(pack_bf16x2(0.25f * a[0], 0.25f * a[1]), pack_bf16x2(0.25f * a[2], 0.25f * a[3]),
                                                           pack_bf16x2(0.25f * a[4], 0.25f * a[5]), pack_bf16x2(0.25f * a[6], 0.25f * a[7]));
  }
}

// AvgPoolGrad: every window element receives dy / 4 (overwrites dx)
__global__ void __launch_bounds__(kThreads) avgpool_bwd_kernel(const bf16* __restrict__ dy, int lddy, bf16* __restrict__ dx,
                                                               int lddx, int N, int H, int W, int C8) {
  const int OH = H >> 1, OW = W >> 1;
  const int64_t total = (int64_t)N * OH * OW * C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % C8);
    int64_t p = i / C8;
    const int ox = (int)(p % OW);
    p /= OW;
    const int oy = (int)(p % OH);
    const int n = (int)(p / OH);
    const int64_t ro = ((int64_t)n * OH + oy) * OW + ox;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(dy + ro * lddy) + g);
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2((&u.x)[j]);
      o[j] = pack_bf16x2(0.25f * f.x, 0.25f * f.y);
    }
    const uint4 v = make_uint4(o[0], o[1], o[2], o[3]);
    const int64_t r0 = ((int64_t)n * H + 2 * oy) * W + 2 * ox;
    reinterpret_cast<uint4*>(dx + r0 * lddx)[g] = v;
    reinterpret_cast<uint4*>(dx + (r0 + 1) * lddx)[g] = v;
    reinterpret_cast<uint4*>(dx + (r0 + W) * lddx)[g] = v;
    reinterpret_cast<uint4*>(dx + (r0 + W + 1) * lddx)[g] = v;
  }
}

// logical w[T][A][B] <-> physical wp[T][Ap][Bp]; amap[Ap], bmap[Bp] = logical index or -1 (NULL = identity on the
// first A / B entries).  to_phys: wp = mapped ? w : 0;  else: w[mapped] = wp * mul
__global__ void __launch_bounds__(kThreads) remap_weights_kernel(float* __restrict__ w, float* __restrict__ wp, int T, int A,
                                                                 int B, int Ap, int Bp, const int* __restrict__ amap,
                                                                 const int* __restrict__ bmap, int to_phys) {
  const int64_t total = (int64_t)T * Ap * Bp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int bp = (int)(i % Bp), ap = (int)((i / Bp) % Ap), t = (int)(i / ((int64_t)Ap * Bp));
    const int la = amap ? amap[ap] : (ap < A ? ap : -1);
    const int lb = bmap ? bmap[bp] : (bp < B ? bp : -1);
    if (to_phys) {
      wp[i] = (la >= 0 && lb >= 0) ? w[((int64_t)t * A + la) * B + lb] : 0.f;
    } else if (la >= 0 && lb >= 0) {
      w[((int64_t)t * A + la) * B + lb] = wp[i];
    }
  }
}

// dst[i] = map[i] >= 0 ? src[map[i]] * mul + add : 0   (logical -> physical per-channel parameters)
__global__ void __launch_bounds__(kThreads) gather_f32_kernel(const float* __restrict__ src, const int* __restrict__ map,
                                                              float* __restrict__ dst, int n, float mul, float add) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = map[i] >= 0 ? src[map[i]] * mul + add : 0.f;
}
// dst[map[i]] = src[i] * mul for map[i] >= 0   (physical -> logical gradients)
__global__ void __launch_bounds__(kThreads) scatter_f32_kernel(const float* __restrict__ src, const int* __restrict__ map,
                                                               float* __restrict__ dst, int n, float mul) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && map[i] >= 0) dst[map[i]] = src[i] * mul;
}

// ---------------------------------------------------------------------------------------
// Remaining op families of Network/utils/utils.py
// ---------------------------------------------------------------------------------------
// Resize_Bilinear (utils.py:329-330: tf.image.resize_bilinear(x, size, align_corners=True)): src = dst * (in-1)/(out-1);
// top = tl + (tr - tl) * fx; bottom = bl + (br - bl) * fx; out = top + (bottom - top) * fy   (TF's evaluation order, fp32)
__device__ __forceinline__ void bilinear_src(int o, float scale, int in, int& i0, int& i1, float& f) {
  const float s = (float)o * scale;
  i0 = (int)floorf(s);
  i1 = i0 + 1 < in ? i0 + 1 : in - 1;
  f = s - (float)i0;
}

__global__ void __launch_bounds__(kThreads) resize_bilinear_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int N, int H,
                                                                       int W, int OH, int OW, int C8, float sy, float sx) {
  const int64_t total = (int64_t)N * OH * OW * C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % C8);
    int64_t p = i / C8;
    const int ox = (int)(p % OW);
    p /= OW;
    const int oy = (int)(p % OH);
    const int n = (int)(p / OH);
    int y0, y1, x0, x1;
    float fy, fx;
    bilinear_src(oy, sy, H, y0, y1, fy);
    bilinear_src(ox, sx, W, x0, x1, fx);
    const uint4* base = reinterpret_cast<const uint4*>(x) + (int64_t)n * H * W * C8 + g;
    const uint4 tl = __ldg(base + ((int64_t)y0 * W + x0) * C8), tr = __ldg(base + ((int64_t)y0 * W + x1) * C8);
    const uint4 bl = __ldg(base + ((int64_t)y1 * W + x0) * C8), br = __ldg(base + ((int64_t)y1 * W + x1) * C8);
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = unpack_bf16x2((&tl.x)[j]), b = unpack_bf16x2((&tr.x)[j]), c = unpack_bf16x2((&bl.x)[j]), d = unpack_bf16x2((&br.x)[j]);
      const float t0 = a.x + (b.x - a.x) * fx, t1 = a.y + (b.y - a.y) * fx;
      const float u0 = c.x + (d.x - c.x) * fx, u1 = c.y + (d.y - c.y) * fx;
      o[j] = pack_bf16x2(t0 + (u0 - t0) * fy, t1 + (u1 - t1) * fy);
    }
    reinterpret_cast<uint4*>(y)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ResizeBilinearGrad in GATHER form (deterministic, no atomics): input pixel (iy, ix) collects dy from every output
// whose bilinear footprint touches it.  Output o touches input i as its lower tap (weight 1 - f) when floor(o*s) == i and
// as its upper tap (weight f) when min(floor(o*s) + 1, in - 1) == i.
__device__ __forceinline__ float bilinear_weight(int o, float scale, int in, int i) {
  int i0, i1;
  float f;
  bilinear_src(o, scale, in, i0, i1, f);
  float w = 0.f;
  if (i0 == i) w += 1.f - f;
  if (i1 == i) w += f;
  return w;
}

__global__ void __launch_bounds__(kThreads) resize_bilinear_bwd_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int N,
                                                                       int H, int W, int OH, int OW, int C8, float sy, float sx) {
  const int64_t total = (int64_t)N * H * W * C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % C8);
    int64_t p = i / C8;
    const int ix = (int)(p % W);
    p /= W;
    const int iy = (int)(p % H);
    const int n = (int)(p / H);
    // candidate outputs: o * s in (i - 1, i + 1)  (everything when s == 0, i.e. a single output row / column)
    int oy0 = 0, oy1 = OH - 1, ox0 = 0, ox1 = OW - 1;
    if (sy > 0.f) { oy0 = max(0, (int)floorf((iy - 1) / sy)); oy1 = min(OH - 1, (int)ceilf((iy + 1) / sy)); }
    if (sx > 0.f) { ox0 = max(0, (int)floorf((ix - 1) / sx)); ox1 = min(OW - 1, (int)ceilf((ix + 1) / sx)); }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int oy = oy0; oy <= oy1; ++oy) {
      const float wy = bilinear_weight(oy, sy, H, iy);
      if (wy == 0.f) continue;
      for (int ox = ox0; ox <= ox1; ++ox) {
        const float w = wy * bilinear_weight(ox, sx, W, ix);
        if (w == 0.f) continue;
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(dy) + (((int64_t)n * OH + oy) * OW + ox) * C8 + g);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2((&u.x)[j]);
          acc[2 * j] += w * f.x;
          acc[2 * j + 1] += w * f.y;
        }
      }
    }
    reinterpret_cast<uint4*>(dx)[i] = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                                                 pack_bf16x2(acc[6], acc[7]));
  }
}

// Global_Avg_Pool (utils.py:312-313: tflearn global_avg_pool = reduce_mean over H, W): block = 32 channel groups x 8 pixel lanes
__global__ void __launch_bounds__(kThreads) global_avgpool_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int HW, int C8) {
  __shared__ float sh[8][32][9];
  const int gl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int g = blockIdx.x * 32 + gl, n = blockIdx.y;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  if (g < C8) {
    const uint4* base = reinterpret_cast<const uint4*>(x) + (int64_t)n * HW * C8 + g;
    for (int p = pl; p < HW; p += 8) {
      const uint4 u = __ldg(base + (int64_t)p * C8);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&u.x)[j]);
        a[2 * j] += f.x; a[2 * j + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[pl][gl][j] = a[j];
  __syncthreads();
  if (pl == 0 && g < C8) {
    const float inv = 1.f / (float)HW;
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      t[j] = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t[j] += sh[k][gl][j];
      t[j] *= inv;
    }
    reinterpret_cast<uint4*>(y)[(int64_t)n * C8 + g] =
        make_uint4(pack_bf16x2(t[0], t[1]), pack_bf16x2(t[2], t[3]), pack_bf16x2(t[4], t[5]), pack_bf16x2(t[6], t[7]));
  }
}

__global__ void __launch_bounds__(kThreads) global_avgpool_bwd_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int N, int HW,
                                                                      int C8) {
  const int64_t total = (int64_t)N * HW * C8;
  const float inv = 1.f / (float)HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % C8);
    const int n = (int)(i / ((int64_t)HW * C8));
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(dy) + (int64_t)n * C8 + g);
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2((&u.x)[j]);
      o[j] = pack_bf16x2(f.x * inv, f.y * inv);
    }
    reinterpret_cast<uint4*>(dx)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Global_Max_Pool (utils.py:315-316: tflearn global_max_pool = reduce_max over H, W).  Forward also counts the maximal
// elements per (n, c): TF's gradient of reduce_max spreads dy equally over ties (indicators / num_selected).
__global__ void __launch_bounds__(kThreads) global_maxpool_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y,
                                                                      int* __restrict__ count, int HW, int C8) {
  __shared__ float shm[8][32][9];
  __shared__ int shc[8][32][9];
  const int gl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int g = blockIdx.x * 32 + gl, n = blockIdx.y;
  float m[8];
  int cnt[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { m[j] = -INFINITY; cnt[j] = 0; }
  if (g < C8) {
    const uint4* base = reinterpret_cast<const uint4*>(x) + (int64_t)n * HW * C8 + g;
    for (int p = pl; p < HW; p += 8) {
      const uint4 u = __ldg(base + (int64_t)p * C8);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&u.x)[j]);
        const float v[2] = {f.x, f.y};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int e = 2 * j + h;
          if (v[h] > m[e]) { m[e] = v[h]; cnt[e] = 1; }
          else if (v[h] == m[e]) ++cnt[e];
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { shm[pl][gl][j] = m[j]; shc[pl][gl][j] = cnt[j]; }
  __syncthreads();
  if (pl == 0 && g < C8) {
    float t[8];
    int c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      t[j] = shm[0][gl][j]; c[j] = shc[0][gl][j];
#pragma unroll
      for (int k = 1; k < 8; ++k) {
        const float v = shm[k][gl][j];
        if (v > t[j]) { t[j] = v; c[j] = shc[k][gl][j]; }
        else if (v == t[j]) c[j] += shc[k][gl][j];
      }
    }
    reinterpret_cast<uint4*>(y)[(int64_t)n * C8 + g] =
        make_uint4(pack_bf16x2(t[0], t[1]), pack_bf16x2(t[2], t[3]), pack_bf16x2(t[4], t[5]), pack_bf16x2(t[6], t[7]));
    int4* cp = reinterpret_cast<int4*>(count + ((int64_t)n * C8 + g) * 8);
    cp[0] = make_int4(c[0], c[1], c[2], c[3]);
    cp[1] = make_int4(c[4], c[5], c[6], c[7]);
  }
}

// dx = (x == y) ? dy / count : 0
__global__ void __launch_bounds__(kThreads) global_maxpool_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                                      const bf16* __restrict__ y, const int* __restrict__ count,
                                                                      bf16* __restrict__ dx, int N, int HW, int C8) {
  const int64_t total = (int64_t)N * HW * C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % C8);
    const int n = (int)(i / ((int64_t)HW * C8));
    const int64_t nc = (int64_t)n * C8 + g;
    const uint4 ud = __ldg(reinterpret_cast<const uint4*>(dy) + nc), uy = __ldg(reinterpret_cast<const uint4*>(y) + nc);
    const uint4 ux = __ldg(reinterpret_cast<const uint4*>(x) + i);
    const int4 c0 = __ldg(reinterpret_cast<const int4*>(count + nc * 8)), c1 = __ldg(reinterpret_cast<const int4*>(count + nc * 8) + 1);
    const int c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fd = unpack_bf16x2((&ud.x)[j]), fy = unpack_bf16x2((&uy.x)[j]), fx = unpack_bf16x2((&ux.x)[j]);
      o[j] = pack_bf16x2(fx.x == fy.x ? fd.x / (float)c[2 * j] : 0.f, fx.y == fy.y ? fd.y / (float)c[2 * j + 1] : 0.f);
    }
    reinterpret_cast<uint4*>(dx)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Zero_Padding (utils.py:325-327: tf.pad with `pad` zeros on every side of H and W); backward = the centre crop
__global__ void __launch_bounds__(kThreads) zero_pad_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W,
                                                            int C8, int pad, int crop) {
  // crop == 0: y [N,H+2p,W+2p,C] <- x [N,H,W,C];  crop == 1: y [N,H,W,C] <- x [N,H+2p,W+2p,C]
  const int OH = crop ? H : H + 2 * pad, OW = crop ? W : W + 2 * pad;
  const int IH = crop ? H + 2 * pad : H, IW = crop ? W + 2 * pad : W;
  const int64_t total = (int64_t)N * OH * OW * C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % C8);
    int64_t p = i / C8;
    const int ox = (int)(p % OW);
    p /= OW;
    const int oy = (int)(p % OH);
    const int n = (int)(p / OH);
    const int iy = crop ? oy + pad : oy - pad, ix = crop ? ox + pad : ox - pad;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < IH && ix >= 0 && ix < IW) v = __ldg(x + (((int64_t)n * IH + iy) * IW + ix) * C8 + g);
    y[i] = v;
  }
}

// dst[r][j] = j < c ? bf16(src[r][j]) : 0 for j < C (C % 8 == 0): a narrow (class-count) gradient padded to a 64-channel
// GEMM operand; thread = 8 output channels
template <typename ST>
__global__ void __launch_bounds__(kThreads) pad_channels_kernel(const ST* __restrict__ src, bf16* __restrict__ dst, int64_t rows, int c,
                                                                int C8) {
  const int64_t total = rows * C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % C8);
    const int64_t r = i / C8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = g * 8 + j;
      v[j] = col < c ? (float)src[r * c + col] : 0.f;
    }
    reinterpret_cast<uint4*>(dst)[i] =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

}  // namespace

extern "C" {

int segk_pad_channels(segk_ctx* ctx, const void* src, int src_is_f32, void* dst, int64_t rows, int c, int C, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, src && dst && rows > 0 && c > 0 && c <= C && C % 8 == 0, "pad_channels: bad args (0 < c <= C, C %% 8 == 0)");
  const int64_t items = rows * (C / 8);
  if (src_is_f32)
    pad_channels_kernel<float><<<sgrid(ctx, items), kThreads, 0, (cudaStream_t)stream>>>((const float*)src, (bf16*)dst, rows, c, C / 8);
  else
    pad_channels_kernel<bf16><<<sgrid(ctx, items), kThreads, 0, (cudaStream_t)stream>>>((const bf16*)src, (bf16*)dst, rows, c, C / 8);
  SEGK_LAUNCHED(ctx, "pad_channels");
  return SEGK_OK;
}

int segk_global_maxpool_fwd(segk_ctx* ctx, const void* x, void* y, int* count, int N, int H, int W, int C, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && y && count && N > 0 && H > 0 && W > 0 && C % 8 == 0 && C > 0 && N <= 65535,
               "global_maxpool_fwd: bad args (C %% 8 == 0)");
  global_maxpool_fwd_kernel<<<dim3(ceil_div(C / 8, 32), N), kThreads, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, count, H * W,
                                                                                                  C / 8);
  SEGK_LAUNCHED(ctx, "global_maxpool_fwd");
  return SEGK_OK;
}

int segk_global_maxpool_bwd(segk_ctx* ctx, const void* dy, const void* x, const void* y, const int* count, void* dx, int N, int H,
                            int W, int C, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, dy && x && y && count && dx && N > 0 && H > 0 && W > 0 && C % 8 == 0 && C > 0, "global_maxpool_bwd: bad args (C %% 8 == 0)");
  global_maxpool_bwd_kernel<<<sgrid(ctx, (int64_t)N * H * W * (C / 8)), kThreads, 0, (cudaStream_t)stream>>>(
      (const bf16*)dy, (const bf16*)x, (const bf16*)y, count, (bf16*)dx, N, H * W, C / 8);
  SEGK_LAUNCHED(ctx, "global_maxpool_bwd");
  return SEGK_OK;
}

int segk_zero_pad(segk_ctx* ctx, const void* x, void* y, int N, int H, int W, int C, int pad, int crop, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && y && N > 0 && H > 0 && W > 0 && C % 8 == 0 && C > 0 && pad >= 0, "zero_pad: bad args (C %% 8 == 0)");
  const int64_t items = (int64_t)N * (crop ? H : H + 2 * pad) * (crop ? W : W + 2 * pad) * (C / 8);
  zero_pad_kernel<<<sgrid(ctx, items), kThreads, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y, N, H, W, C / 8, pad, crop ? 1 : 0);
  SEGK_LAUNCHED(ctx, "zero_pad");
  return SEGK_OK;
}

int segk_bn_act_fwd(segk_ctx* ctx, const void* x, int ldx, void* y, int ldy, const float* scale, const float* shift,
                    int64_t rows, int C, int relu, const uint8_t* drop_mask, float drop_keep, uint64_t drop_seed, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, !(drop_keep > 0.f && drop_keep < 1.f) || ((uintptr_t)drop_mask & 7) == 0, "bn_act_fwd: dropout mask alignment");
  SEGK_REQUIRE(ctx, x && y && scale && shift && rows > 0 && C > 0, "bn_act_fwd: bad args");
  SEGK_REQUIRE(ctx, C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && ldx >= C && ldy >= C &&
                        (((uintptr_t)x | (uintptr_t)y | (uintptr_t)scale | (uintptr_t)shift) & 15) == 0,
               "bn_act_fwd: C, ldx, ldy multiples of 8, 16-byte aligned pointers");
  bn_act_fwd_kernel<<<sgrid(ctx, rows * (C / 8)), kThreads, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, (bf16*)y, ldy, scale,
                                                                                       shift, rows, C / 8, relu,
                                                                                       make_drop(drop_mask, drop_keep, drop_seed));
  SEGK_LAUNCHED(ctx, "bn_act_fwd");
  return SEGK_OK;
}

size_t segk_bn_act_bwd_workspace_bytes(segk_ctx* ctx, int C) {
  return sizeof(float) * (size_t)(ctx ? ctx->sm_count : 148) * 4 * 2 * (size_t)C;
}

int segk_bn_act_bwd(segk_ctx* ctx, const void* dy, const void* y, int ldy, const void* x, void* dx, int ldx,
                    const float* scale, float* dscale, float* dshift, void* workspace, size_t workspace_bytes, int64_t rows,
                    int C, int relu, int accumulate, const uint8_t* drop_mask, float drop_keep, uint64_t drop_seed, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, !(drop_keep > 0.f && drop_keep < 1.f) || (!accumulate && ((uintptr_t)drop_mask & 7) == 0),
               "bn_act_bwd: the fused DropoutGrad overwrites dx (no accumulate); 8-byte aligned mask");
  SEGK_REQUIRE(ctx, dy && y && x && dx && scale && dscale && dshift && workspace && rows > 0 && C > 0, "bn_act_bwd: bad args");
  SEGK_REQUIRE(ctx, C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 &&
                        (((uintptr_t)dy | (uintptr_t)y | (uintptr_t)x | (uintptr_t)dx | (uintptr_t)scale) & 15) == 0,
               "bn_act_bwd: C, ldx, ldy multiples of 8, 16-byte aligned pointers");
  const int C8 = C / 8;
  const int cpb = C8 < kThreads ? C8 : kThreads;
  const int R = kThreads / cpb;             // row lanes per block (threads beyond R * cpb idle)
  const int gy = ceil_div(C8, cpb);
  int64_t gx = ceil_div64(rows, (int64_t)R * 4);
  const int64_t cap = ceil_div64((int64_t)ctx->sm_count * 4, gy);
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  SEGK_REQUIRE(ctx, workspace_bytes >= sizeof(float) * (size_t)gx * 2 * C, "bn_act_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  bn_act_bwd_kernel<<<dim3((unsigned)gx, gy), kThreads, 0, st>>>((const bf16*)dy, (const bf16*)y, ldy, (const bf16*)x, (bf16*)dx,
                                                                 ldx, scale, (float*)workspace, rows, C8, relu, accumulate,
                                                                 make_drop(drop_mask, drop_keep, drop_seed));
  SEGK_LAUNCHED(ctx, "bn_act_bwd");
  dense_reduce_rows_kernel<<<ceil_div(2 * C, 32), kThreads, 0, st>>>((const float*)workspace, dscale, dshift, (int)gx, C);
  SEGK_LAUNCHED(ctx, "bn_act_bwd_reduce");
  return SEGK_OK;
}

int segk_avgpool2x2_fwd(segk_ctx* ctx, const void* x, int ldx, void* y, int ldy, int N, int H, int W, int C, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && y && N > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "avgpool_fwd: bad args (even H, W)");
  SEGK_REQUIRE(ctx, C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0,
               "avgpool_fwd: C, ldx, ldy multiples of 8, 16-byte aligned pointers");
  const int64_t items = (int64_t)N * (H / 2) * (W / 2) * (C / 8);
  avgpool_fwd_kernel<<<sgrid(ctx, items), kThreads, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, (bf16*)y, ldy, N, H, W, C / 8);
  SEGK_LAUNCHED(ctx, "avgpool_fwd");
  return SEGK_OK;
}

int segk_avgpool2x2_bwd(segk_ctx* ctx, const void* dy, int lddy, void* dx, int lddx, int N, int H, int W, int C, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, dy && dx && N > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "avgpool_bwd: bad args (even H, W)");
  SEGK_REQUIRE(ctx, C % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0 && (((uintptr_t)dy | (uintptr_t)dx) & 15) == 0,
               "avgpool_bwd: C, ld multiples of 8, 16-byte aligned pointers");
  const int64_t items = (int64_t)N * (H / 2) * (W / 2) * (C / 8);
  avgpool_bwd_kernel<<<sgrid(ctx, items), kThreads, 0, (cudaStream_t)stream>>>((const bf16*)dy, lddy, (bf16*)dx, lddx, N, H, W, C / 8);
  SEGK_LAUNCHED(ctx, "avgpool_bwd");
  return SEGK_OK;
}

int segk_remap_weights(segk_ctx* ctx, float* w, float* wp, int T, int A, int B, int Ap, int Bp, const int* amap, const int* bmap,
                       int to_phys, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, w && wp && T > 0 && A > 0 && B > 0 && Ap >= 1 && Bp >= 1, "remap_weights: bad args");
  remap_weights_kernel<<<sgrid(ctx, (int64_t)T * Ap * Bp), kThreads, 0, (cudaStream_t)stream>>>(w, wp, T, A, B, Ap, Bp, amap, bmap,
                                                                                               to_phys);
  SEGK_LAUNCHED(ctx, "remap_weights");
  return SEGK_OK;
}

int segk_gather_f32(segk_ctx* ctx, const float* src, const int* map, float* dst, int n, float mul, float add, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, src && map && dst && n > 0, "gather_f32: bad args");
  gather_f32_kernel<<<ceil_div(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(src, map, dst, n, mul, add);
  SEGK_LAUNCHED(ctx, "gather_f32");
  return SEGK_OK;
}

int segk_scatter_f32(segk_ctx* ctx, const float* src, const int* map, float* dst, int n, float mul, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, src && map && dst && n > 0, "scatter_f32: bad args");
  scatter_f32_kernel<<<ceil_div(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(src, map, dst, n, mul);
  SEGK_LAUNCHED(ctx, "scatter_f32");
  return SEGK_OK;
}

int segk_resize_bilinear_fwd(segk_ctx* ctx, const void* x, void* y, int N, int H, int W, int OH, int OW, int C, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && y && N > 0 && H > 0 && W > 0 && OH > 0 && OW > 0 && C % 8 == 0 && C > 0, "resize_bilinear_fwd: bad args (C %% 8 == 0)");
  const float sy = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.f, sx = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
  resize_bilinear_fwd_kernel<<<sgrid(ctx, (int64_t)N * OH * OW * (C / 8)), kThreads, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, (bf16*)y, N, H, W, OH, OW, C / 8, sy, sx);
  SEGK_LAUNCHED(ctx, "resize_bilinear_fwd");
  return SEGK_OK;
}

int segk_resize_bilinear_bwd(segk_ctx* ctx, const void* dy, void* dx, int N, int H, int W, int OH, int OW, int C, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, dy && dx && N > 0 && H > 0 && W > 0 && OH > 0 && OW > 0 && C % 8 == 0 && C > 0, "resize_bilinear_bwd: bad args (C %% 8 == 0)");
  const float sy = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.f, sx = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
  resize_bilinear_bwd_kernel<<<sgrid(ctx, (int64_t)N * H * W * (C / 8)), kThreads, 0, (cudaStream_t)stream>>>(
      (const bf16*)dy, (bf16*)dx, N, H, W, OH, OW, C / 8, sy, sx);
  SEGK_LAUNCHED(ctx, "resize_bilinear_bwd");
  return SEGK_OK;
}

int segk_global_avgpool_fwd(segk_ctx* ctx, const void* x, void* y, int N, int H, int W, int C, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && y && N > 0 && H > 0 && W > 0 && C % 8 == 0 && C > 0 && N <= 65535, "global_avgpool_fwd: bad args (C %% 8 == 0)");
  global_avgpool_fwd_kernel<<<dim3(ceil_div(C / 8, 32), N), kThreads, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, H * W, C / 8);
  SEGK_LAUNCHED(ctx, "global_avgpool_fwd");
  return SEGK_OK;
}

int segk_global_avgpool_bwd(segk_ctx* ctx, const void* dy, void* dx, int N, int H, int W, int C, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, dy && dx && N > 0 && H > 0 && W > 0 && C % 8 == 0 && C > 0, "global_avgpool_bwd: bad args (C %% 8 == 0)");
  global_avgpool_bwd_kernel<<<sgrid(ctx, (int64_t)N * H * W * (C / 8)), kThreads, 0, (cudaStream_t)stream>>>((const bf16*)dy, (bf16*)dx, N,
                                                                                                      H * W, C / 8);
  SEGK_LAUNCHED(ctx, "global_avgpool_bwd");
  return SEGK_OK;
}

}  // extern "C"
