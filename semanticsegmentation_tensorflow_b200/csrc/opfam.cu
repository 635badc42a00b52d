// HBM-bound kernels of the remaining op families of Network/model, NHWC bf16, C % 8 == 0, 16-byte
// accesses, thread = 8 channels of one output element, grid-stride loops, no atomics (every reduction is a fixed-order
// two-stage sum):
//   * Avg_Pooling(x, kh, kw, stride_h, stride_w, VALID)  (utils.py:309; the pyramid windows of PSPNet.py:147-165,546-567)
//   * Max_Pooling(x, kh, kw, stride, VALID | SAME)        (utils.py:306; the 3x3 / stride-2 stem pools of PSPNet.py:34,190)
//     with the first-max index of every window (u8) for MaxPoolGrad -- overlapping windows accumulate
//   * tf.nn.depthwise_conv2d(x, filter[kh,kw,C,1], stride, SAME, rate)  (DeepLabv3Plus.py:49,117, EfficientNet.py:173,453,
//     Generative_Segmentation_net.py:53) forward, input gradient, filter gradient
//   * sigmoid / swish and the squeeze-excite multiply x * s[n, c]      (EfficientNet.py MBConv / SE blocks)
// TF's SAME geometry: out = ceil(in / s), pad_total = max((out - 1) s + k_eff - in, 0), pad_before = pad_total / 2.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

inline int sgrid(segk_ctx* ctx, int64_t items, int per_sm = 8) {
  int64_t b = ceil_div64(items, kThreads), cap = (int64_t)ctx->sm_count * per_sm;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

struct Geo {
  int N, H, W, C8, OH, OW, kh, kw, sh, sw, ph, pw, rh, rw;      // ph / pw: padding before; rh / rw: dilation
};

inline int same_out(int in, int s) { return (in + s - 1) / s; }
inline int same_pad_before(int in, int k_eff, int s) {
  const int out = same_out(in, s);
  int total = (out - 1) * s + k_eff - in;
  if (total < 0) total = 0;
  return total / 2;
}

__device__ __forceinline__ void add8(float (&a)[8], uint4 u) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack_bf16x2((&u.x)[j]);
    a[2 * j] += f.x;
    a[2 * j + 1] += f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&a)[8]) {
  return make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
}
__device__ __forceinline__ void unpack8(uint4 u, float (&a)[8]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack_bf16x2((&u.x)[j]);
    a[2 * j] = f.x;
    a[2 * j + 1] = f.y;
  }
}

// ---- average pool, VALID ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) avgpool_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, Geo g) {
  const int64_t total = (int64_t)g.N * g.OH * g.OW * g.C8;
  const float inv = 1.f / (float)(g.kh * g.kw);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % g.C8);
    int64_t p = i / g.C8;
    const int ox = (int)(p % g.OW);
    p /= g.OW;
    const int oy = (int)(p % g.OH), n = (int)(p / g.OH);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int ky = 0; ky < g.kh; ++ky)
      for (int kx = 0; kx < g.kw; ++kx)
        add8(a, __ldg(x + (((int64_t)n * g.H + oy * g.sh + ky) * g.W + ox * g.sw + kx) * g.C8 + c8));
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= inv;
    y[i] = pack8(a);
  }
}

// dx[p] = (1 / (kh kw)) * sum of dy over the windows that contain p (gather: deterministic)
__global__ void __launch_bounds__(kThreads) avgpool_bwd_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, Geo g) {
  const int64_t total = (int64_t)g.N * g.H * g.W * g.C8;
  const float inv = 1.f / (float)(g.kh * g.kw);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % g.C8);
    int64_t p = i / g.C8;
    const int ix = (int)(p % g.W);
    p /= g.W;
    const int iy = (int)(p % g.H), n = (int)(p / g.H);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // windows oy with oy*sh <= iy < oy*sh + kh
    int oy0 = iy - g.kh + 1;
    oy0 = oy0 <= 0 ? 0 : (oy0 + g.sh - 1) / g.sh;
    int ox0 = ix - g.kw + 1;
    ox0 = ox0 <= 0 ? 0 : (ox0 + g.sw - 1) / g.sw;
    for (int oy = oy0; oy < g.OH && oy * g.sh <= iy; ++oy)
      for (int ox = ox0; ox < g.OW && ox * g.sw <= ix; ++ox)
        add8(a, __ldg(dy + (((int64_t)n * g.OH + oy) * g.OW + ox) * g.C8 + c8));
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= inv;
    dx[i] = pack8(a);
  }
}

// ---- max pool, VALID / SAME, first max in row-major window order ---------------------------------------------
__global__ void __launch_bounds__(kThreads) maxpool_gen_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                                   uint2* __restrict__ idx, Geo g) {
  const int64_t total = (int64_t)g.N * g.OH * g.OW * g.C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % g.C8);
    int64_t p = i / g.C8;
    const int ox = (int)(p % g.OW);
    p /= g.OW;
    const int oy = (int)(p % g.OH), n = (int)(p / g.OH);
    float best[8];
    uint32_t bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; }
    bool any = false;
    for (int ky = 0; ky < g.kh; ++ky) {
      const int iy = oy * g.sh - g.ph + ky;
      if (iy < 0 || iy >= g.H) continue;
      for (int kx = 0; kx < g.kw; ++kx) {
        const int ix = ox * g.sw - g.pw + kx;
        if (ix < 0 || ix >= g.W) continue;
        float v[8];
        unpack8(__ldg(x + (((int64_t)n * g.H + iy) * g.W + ix) * g.C8 + c8), v);
        const uint32_t k = (uint32_t)(ky * g.kw + kx);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (!any || v[j] > best[j]) { best[j] = v[j]; bi[j] = k; }
        any = true;
      }
    }
    y[i] = pack8(best);
    idx[i] = make_uint2(bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24), bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24));
  }
}

// dx[p] = sum over the windows containing p of [idx(window) == position of p in it] * dy(window)   (gather)
__global__ void __launch_bounds__(kThreads) maxpool_gen_bwd_kernel(const uint4* __restrict__ dy, const uint2* __restrict__ idx,
                                                                   uint4* __restrict__ dx, Geo g) {
  const int64_t total = (int64_t)g.N * g.H * g.W * g.C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % g.C8);
    int64_t p = i / g.C8;
    const int ix = (int)(p % g.W);
    p /= g.W;
    const int iy = (int)(p % g.H), n = (int)(p / g.H);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int oy0 = iy + g.ph - g.kh + 1;
    oy0 = oy0 <= 0 ? 0 : (oy0 + g.sh - 1) / g.sh;
    int ox0 = ix + g.pw - g.kw + 1;
    ox0 = ox0 <= 0 ? 0 : (ox0 + g.sw - 1) / g.sw;
    for (int oy = oy0; oy < g.OH && oy * g.sh - g.ph <= iy; ++oy)
      for (int ox = ox0; ox < g.OW && ox * g.sw - g.pw <= ix; ++ox) {
        const uint32_t k = (uint32_t)((iy - (oy * g.sh - g.ph)) * g.kw + (ix - (ox * g.sw - g.pw)));
        const int64_t o = (((int64_t)n * g.OH + oy) * g.OW + ox) * g.C8 + c8;
        const uint2 id = __ldg(idx + o);
        float v[8];
        unpack8(__ldg(dy + o), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t kj = ((j < 4 ? id.x : id.y) >> (8 * (j & 3))) & 0xffu;
          if (kj == k) a[j] += v[j];
        }
      }
    dx[i] = pack8(a);
  }
}

// ---- depthwise conv, SAME, stride s, dilation r, channel multiplier 1; weights fp32 [kh][kw][C] -------------------
__global__ void __launch_bounds__(kThreads) depthwise_fwd_kernel(const uint4* __restrict__ x, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, uint4* __restrict__ y, Geo g,
                                                                 int relu) {
  const int64_t total = (int64_t)g.N * g.OH * g.OW * g.C8;
  const int C = g.C8 * 8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % g.C8);
    int64_t p = i / g.C8;
    const int ox = (int)(p % g.OW);
    p /= g.OW;
    const int oy = (int)(p % g.OH), n = (int)(p / g.OH);
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = bias ? __ldg(bias + c8 * 8 + j) : 0.f;
    for (int ky = 0; ky < g.kh; ++ky) {
      const int iy = oy * g.sh - g.ph + ky * g.rh;
      if (iy < 0 || iy >= g.H) continue;
      for (int kx = 0; kx < g.kw; ++kx) {
        const int ix = ox * g.sw - g.pw + kx * g.rw;
        if (ix < 0 || ix >= g.W) continue;
        float v[8];
        unpack8(__ldg(x + (((int64_t)n * g.H + iy) * g.W + ix) * g.C8 + c8), v);
        const float4* w4 = reinterpret_cast<const float4*>(w + (int64_t)(ky * g.kw + kx) * C + c8 * 8);
        const float4 w0 = __ldg(w4), w1 = __ldg(w4 + 1);
        a[0] += v[0] * w0.x; a[1] += v[1] * w0.y; a[2] += v[2] * w0.z; a[3] += v[3] * w0.w;
        a[4] += v[4] * w1.x; a[5] += v[5] * w1.y; a[6] += v[6] * w1.z; a[7] += v[7] * w1.w;
      }
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = fmaxf(a[j], 0.f);
    }
    y[i] = pack8(a);
  }
}

// dx[iy,ix] = sum over taps of dy[(iy + ph - ky r) / s, ...] * w[ky,kx]   where the division is exact
__global__ void __launch_bounds__(kThreads) depthwise_dgrad_kernel(const uint4* __restrict__ dy, const float* __restrict__ w,
                                                                   uint4* __restrict__ dx, Geo g) {
  const int64_t total = (int64_t)g.N * g.H * g.W * g.C8;
  const int C = g.C8 * 8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % g.C8);
    int64_t p = i / g.C8;
    const int ix = (int)(p % g.W);
    p /= g.W;
    const int iy = (int)(p % g.H), n = (int)(p / g.H);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int ky = 0; ky < g.kh; ++ky) {
      const int ty = iy + g.ph - ky * g.rh;
      if (ty < 0 || ty % g.sh) continue;
      const int oy = ty / g.sh;
      if (oy >= g.OH) continue;
      for (int kx = 0; kx < g.kw; ++kx) {
        const int tx = ix + g.pw - kx * g.rw;
        if (tx < 0 || tx % g.sw) continue;
        const int ox = tx / g.sw;
        if (ox >= g.OW) continue;
        float v[8];
        unpack8(__ldg(dy + (((int64_t)n * g.OH + oy) * g.OW + ox) * g.C8 + c8), v);
        const float4* w4 = reinterpret_cast<const float4*>(w + (int64_t)(ky * g.kw + kx) * C + c8 * 8);
        const float4 w0 = __ldg(w4), w1 = __ldg(w4 + 1);
        a[0] += v[0] * w0.x; a[1] += v[1] * w0.y; a[2] += v[2] * w0.z; a[3] += v[3] * w0.w;
        a[4] += v[4] * w1.x; a[5] += v[5] * w1.y; a[6] += v[6] * w1.z; a[7] += v[7] * w1.w;
      }
    }
    dx[i] = pack8(a);
  }
}

// dw[tap][c] = sum over output pixels of x[shifted] * dy.  Stage 1: block b takes the output pixels b, b + gridDim.x, ...;
// thread = (8-channel group, pixel lane); per-thread sums over its pixels for every tap, block-reduced over the pixel
// lanes in a fixed order into one partial row part[b][tap][C].  Stage 2 adds the partial rows in order.
constexpr int kDwMaxTaps = 49;
__global__ void __launch_bounds__(kThreads) depthwise_wgrad_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy,
                                                                   float* __restrict__ part, Geo g, int tap0, int ntap) {
  extern __shared__ float sh[];      // [lanes][ntap * 8] per channel group column
  const int groups = g.C8 < kThreads ? g.C8 : kThreads;     // channel groups handled per block pass
  const int lanes = kThreads / groups;                      // pixel lanes (threads beyond lanes * groups idle)
  const int gi = threadIdx.x % groups, lane = threadIdx.x / groups;
  const int C = g.C8 * 8;
  const int64_t npix = (int64_t)g.N * g.OH * g.OW;
  for (int base = 0; base < g.C8; base += groups) {         // (every thread takes every pass: the barriers below stay matched)
    const int c8 = base + gi;
    const bool live = c8 < g.C8 && lane < lanes;
    float acc[9][8];                 // up to 9 taps per pass (host loops over tap ranges)
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
    if (live)
      for (int64_t q = (int64_t)blockIdx.x * lanes + lane; q < npix; q += (int64_t)gridDim.x * lanes) {
        const int ox = (int)(q % g.OW);
        const int oy = (int)((q / g.OW) % g.OH), n = (int)(q / ((int64_t)g.OW * g.OH));
        float d[8];
        unpack8(__ldg(dy + q * g.C8 + c8), d);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          if (t >= ntap) break;
          const int tap = tap0 + t, ky = tap / g.kw, kx = tap % g.kw;
          const int iy = oy * g.sh - g.ph + ky * g.rh, ix = ox * g.sw - g.pw + kx * g.rw;
          if (iy < 0 || iy >= g.H || ix < 0 || ix >= g.W) continue;
          float v[8];
          unpack8(__ldg(x + (((int64_t)n * g.H + iy) * g.W + ix) * g.C8 + c8), v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[t][j] += v[j] * d[j];
        }
      }
    // reduce over the pixel lanes (fixed order), one tap at a time through shared memory
    for (int t = 0; t < ntap; ++t) {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j) sh[(lane * groups + gi) * 8 + j] = acc[t][j];
      __syncthreads();
      if (lane == 0 && c8 < g.C8) {
        float s[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = sh[gi * 8 + j];
        for (int l = 1; l < lanes; ++l)
#pragma unroll
          for (int j = 0; j < 8; ++j) s[j] += sh[(l * groups + gi) * 8 + j];
        float* o = part + ((int64_t)blockIdx.x * g.kh * g.kw + tap0 + t) * C + c8 * 8;
        *reinterpret_cast<float4*>(o) = make_float4(s[0], s[1], s[2], s[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(s[4], s[5], s[6], s[7]);
      }
    }
  }
}
__global__ void __launch_bounds__(kThreads) sum_rows_kernel(const float* __restrict__ part, float* __restrict__ out, int rows,
                                                            int64_t n, int accumulate) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = accumulate ? out[i] : 0.f;
    for (int r = 0; r < rows; ++r) s += part[(int64_t)r * n + i];
    out[i] = s;
  }
}

// ---- sigmoid / swish, squeeze-excite multiply ------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + __expf(-v)); }
__global__ void __launch_bounds__(kThreads) act_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int64_t n8, int kind) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
    unpack8(__ldg(x + i), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = sigmoidf_(v[j]);
      v[j] = kind == 0 ? s : v[j] * s;
    }
    y[i] = pack8(v);
  }
}
// sigmoid: dx = dy s (1 - s);  swish: dx = dy (s + x s (1 - s))
__global__ void __launch_bounds__(kThreads) act_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy,
                                                           uint4* __restrict__ dx, int64_t n8, int kind) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8], d[8];
    unpack8(__ldg(x + i), v);
    unpack8(__ldg(dy + i), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = sigmoidf_(v[j]);
      const float ds = s * (1.f - s);
      d[j] *= kind == 0 ? ds : (s + v[j] * ds);
    }
    dx[i] = pack8(d);
  }
}
// y[n, p, c] = x[n, p, c] * s[n, c]
__global__ void __launch_bounds__(kThreads) chscale_fwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ s,
                                                               uint4* __restrict__ y, int64_t total, int64_t hw, int C8) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int64_t n = i / C8 / hw;
    float v[8], m[8];
    unpack8(__ldg(x + i), v);
    unpack8(__ldg(s + n * C8 + c8), m);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= m[j];
    y[i] = pack8(v);
  }
}
// dx = dy * s;  ds[n, c] = sum_p dy x: block (n, channel-group chunk) with pixel lanes, fixed-order reduction
__global__ void __launch_bounds__(kThreads) chscale_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ s,
                                                               const uint4* __restrict__ dy, uint4* __restrict__ dx,
                                                               float* __restrict__ ds, int64_t hw, int C8) {
  __shared__ float sh[kThreads][9];
  const int groups = C8 < kThreads ? C8 : kThreads;
  const int lanes = kThreads / groups;
  const int gi = threadIdx.x % groups, lane = threadIdx.x / groups;
  const int n = blockIdx.y;
  const int c8 = blockIdx.x * groups + gi;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c8 < C8 && lane < lanes) {
    float m[8];
    unpack8(__ldg(s + (int64_t)n * C8 + c8), m);
    for (int64_t p = lane; p < hw; p += lanes) {
      const int64_t o = ((int64_t)n * hw + p) * C8 + c8;
      float v[8], d[8];
      unpack8(__ldg(x + o), v);
      unpack8(__ldg(dy + o), d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] += v[j] * d[j];
        d[j] *= m[j];
      }
      dx[o] = pack8(d);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[threadIdx.x][j] = acc[j];
  __syncthreads();
  if (lane == 0 && c8 < C8) {
    for (int l = 1; l < lanes; ++l)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += sh[l * groups + gi][j];
    float* o = ds + ((int64_t)n * C8 + c8) * 8;
    *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

int pool_geo(segk_ctx* ctx, const char* what, Geo& g, int N, int H, int W, int C, int kh, int kw, int sh, int sw, int same) {
  SEGK_REQUIRE(ctx, N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "%s: need C %% 8 == 0 (got %dx%dx%dx%d)", what, N, H, W, C);
  SEGK_REQUIRE(ctx, kh >= 1 && kw >= 1 && sh >= 1 && sw >= 1, "%s: bad window %dx%d / stride %dx%d", what, kh, kw, sh, sw);
  g.N = N; g.H = H; g.W = W; g.C8 = C / 8; g.kh = kh; g.kw = kw; g.sh = sh; g.sw = sw; g.rh = g.rw = 1;
  if (same) {
    g.OH = same_out(H, sh); g.OW = same_out(W, sw);
    g.ph = same_pad_before(H, kh, sh); g.pw = same_pad_before(W, kw, sw);
  } else {
    SEGK_REQUIRE(ctx, H >= kh && W >= kw, "%s: VALID window %dx%d larger than the %dx%d map", what, kh, kw, H, W);
    g.OH = (H - kh) / sh + 1; g.OW = (W - kw) / sw + 1;
    g.ph = g.pw = 0;
  }
  return SEGK_OK;
}

int dw_geo(segk_ctx* ctx, const char* what, Geo& g, int N, int H, int W, int C, int kh, int kw, int s, int rate) {
  SEGK_REQUIRE(ctx, N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "%s: need C %% 8 == 0 (got %dx%dx%dx%d)", what, N, H, W, C);
  SEGK_REQUIRE(ctx, kh >= 1 && kw >= 1 && kh * kw <= kDwMaxTaps && s >= 1 && rate >= 1 && (s == 1 || rate == 1),
               "%s: window %dx%d (<= %d taps), stride %d, rate %d (tf.nn.depthwise_conv2d: rate > 1 needs stride 1)", what, kh, kw,
               kDwMaxTaps, s, rate);
  g.N = N; g.H = H; g.W = W; g.C8 = C / 8; g.kh = kh; g.kw = kw; g.sh = g.sw = s; g.rh = g.rw = rate;
  g.OH = same_out(H, s); g.OW = same_out(W, s);
  g.ph = same_pad_before(H, (kh - 1) * rate + 1, s); g.pw = same_pad_before(W, (kw - 1) * rate + 1, s);
  return SEGK_OK;
}

}  // namespace

extern "C" {

int segk_avgpool_fwd(segk_ctx* ctx, const void* x, void* y, int N, int H, int W, int C, int kh, int kw, int sh, int sw, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && y && (((uintptr_t)x | (uintptr_t)y) & 15) == 0, "avgpool_fwd: null or unaligned (16 B) pointer");
  Geo g;
  const int rc = pool_geo(ctx, "avgpool_fwd", g, N, H, W, C, kh, kw, sh, sw, 0);
  if (rc) return rc;
  avgpool_fwd_kernel<<<sgrid(ctx, (int64_t)N * g.OH * g.OW * g.C8), kThreads, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y, g);
  SEGK_LAUNCHED(ctx, "avgpool_fwd");
  return SEGK_OK;
}

int segk_avgpool_bwd(segk_ctx* ctx, const void* dy, void* dx, int N, int H, int W, int C, int kh, int kw, int sh, int sw, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, dy && dx && (((uintptr_t)dy | (uintptr_t)dx) & 15) == 0, "avgpool_bwd: null or unaligned (16 B) pointer");
  Geo g;
  const int rc = pool_geo(ctx, "avgpool_bwd", g, N, H, W, C, kh, kw, sh, sw, 0);
  if (rc) return rc;
  avgpool_bwd_kernel<<<sgrid(ctx, (int64_t)N * H * W * g.C8), kThreads, 0, (cudaStream_t)stream>>>((const uint4*)dy, (uint4*)dx, g);
  SEGK_LAUNCHED(ctx, "avgpool_bwd");
  return SEGK_OK;
}

int segk_maxpool_fwd(segk_ctx* ctx, const void* x, void* y, uint8_t* idx, int N, int H, int W, int C, int kh, int kw, int stride,
                     int same, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && y && idx && (((uintptr_t)x | (uintptr_t)y) & 15) == 0 && (((uintptr_t)idx) & 7) == 0, "maxpool_fwd: null or unaligned pointer");
  Geo g;
  SEGK_REQUIRE(ctx, kh * kw <= 256, "maxpool_fwd: the window position is stored in a byte (got %dx%d)", kh, kw);
  const int rc = pool_geo(ctx, "maxpool_fwd", g, N, H, W, C, kh, kw, stride, stride, same);
  if (rc) return rc;
  maxpool_gen_fwd_kernel<<<sgrid(ctx, (int64_t)N * g.OH * g.OW * g.C8), kThreads, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y,
                                                                                                        (uint2*)idx, g);
  SEGK_LAUNCHED(ctx, "maxpool_fwd (general)");
  return SEGK_OK;
}

int segk_maxpool_bwd(segk_ctx* ctx, const void* dy, const uint8_t* idx, void* dx, int N, int H, int W, int C, int kh, int kw, int stride,
                     int same, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, dy && idx && dx && (((uintptr_t)dy | (uintptr_t)dx) & 15) == 0 && (((uintptr_t)idx) & 7) == 0, "maxpool_bwd: null or unaligned pointer");
  Geo g;
  SEGK_REQUIRE(ctx, kh * kw <= 256, "maxpool_bwd: the window position is stored in a byte (got %dx%d)", kh, kw);
  const int rc = pool_geo(ctx, "maxpool_bwd", g, N, H, W, C, kh, kw, stride, stride, same);
  if (rc) return rc;
  maxpool_gen_bwd_kernel<<<sgrid(ctx, (int64_t)N * H * W * g.C8), kThreads, 0, (cudaStream_t)stream>>>((const uint4*)dy, (const uint2*)idx,
                                                                                                  (uint4*)dx, g);
  SEGK_LAUNCHED(ctx, "maxpool_bwd (general)");
  return SEGK_OK;
}

int segk_depthwise_conv2d_fwd(segk_ctx* ctx, const void* x, const float* w, const float* bias, void* y, int N, int H, int W, int C,
                              int kh, int kw, int stride, int rate, unsigned flags, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && w && y && !(flags & SEGK_EPI_OUT_F32) && (((uintptr_t)w | (uintptr_t)bias | (uintptr_t)x | (uintptr_t)y) & 15) == 0,
               "depthwise_conv2d_fwd: null / unaligned (16 B) pointer, or fp32 output requested (bf16 only)");
  Geo g;
  const int rc = dw_geo(ctx, "depthwise_conv2d_fwd", g, N, H, W, C, kh, kw, stride, rate);
  if (rc) return rc;
  depthwise_fwd_kernel<<<sgrid(ctx, (int64_t)N * g.OH * g.OW * g.C8), kThreads, 0, (cudaStream_t)stream>>>(
      (const uint4*)x, w, bias, (uint4*)y, g, (flags & SEGK_EPI_RELU) ? 1 : 0);
  SEGK_LAUNCHED(ctx, "depthwise_conv2d_fwd");
  return SEGK_OK;
}

int segk_depthwise_conv2d_dgrad(segk_ctx* ctx, const void* dy, const float* w, void* dx, int N, int H, int W, int C, int kh, int kw,
                                int stride, int rate, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, dy && w && dx && (((uintptr_t)w | (uintptr_t)dy | (uintptr_t)dx) & 15) == 0, "depthwise_conv2d_dgrad: null or unaligned (16 B) pointer");
  Geo g;
  const int rc = dw_geo(ctx, "depthwise_conv2d_dgrad", g, N, H, W, C, kh, kw, stride, rate);
  if (rc) return rc;
  depthwise_dgrad_kernel<<<sgrid(ctx, (int64_t)N * H * W * g.C8), kThreads, 0, (cudaStream_t)stream>>>((const uint4*)dy, w, (uint4*)dx, g);
  SEGK_LAUNCHED(ctx, "depthwise_conv2d_dgrad");
  return SEGK_OK;
}

int segk_depthwise_conv2d_wgrad(segk_ctx* ctx, const void* x, const void* dy, float* dw, int N, int H, int W, int C, int kh, int kw,
                                int stride, int rate, int accumulate, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && dy && dw && (((uintptr_t)dw | (uintptr_t)x | (uintptr_t)dy) & 15) == 0, "depthwise_conv2d_wgrad: null or unaligned (16 B) pointer");
  Geo g;
  int rc = dw_geo(ctx, "depthwise_conv2d_wgrad", g, N, H, W, C, kh, kw, stride, rate);
  if (rc) return rc;
  const int groups = g.C8 < kThreads ? g.C8 : kThreads;
  const int lanes = kThreads / groups;
  const int64_t npix = (int64_t)N * g.OH * g.OW;
  int blocks = (int)(ceil_div64(npix, (int64_t)lanes * 16) < (int64_t)ctx->sm_count * 2 ? ceil_div64(npix, (int64_t)lanes * 16)
                                                                                       : (int64_t)ctx->sm_count * 2);
  if (blocks < 1) blocks = 1;
  const int64_t n = (int64_t)kh * kw * C;
  rc = segk_grow(ctx, &ctx->ws7, &ctx->ws7_bytes, sizeof(float) * (size_t)blocks * n, "depthwise wgrad partials");
  if (rc) return rc;
  for (int tap0 = 0; tap0 < kh * kw; tap0 += 9) {
    const int ntap = kh * kw - tap0 < 9 ? kh * kw - tap0 : 9;
    depthwise_wgrad_kernel<<<blocks, kThreads, sizeof(float) * kThreads * 8, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)dy,
                                                                                                  (float*)ctx->ws7, g, tap0, ntap);
    SEGK_LAUNCHED(ctx, "depthwise_conv2d_wgrad");
  }
  sum_rows_kernel<<<sgrid(ctx, n), kThreads, 0, (cudaStream_t)stream>>>((const float*)ctx->ws7, dw, blocks, n, accumulate);
  SEGK_LAUNCHED(ctx, "depthwise_conv2d_wgrad reduce");
  return SEGK_OK;
}

int segk_activation_fwd(segk_ctx* ctx, const void* x, void* y, int64_t n, int kind, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && y && n > 0 && n % 8 == 0 && (kind == 0 || kind == 1) && (((uintptr_t)x | (uintptr_t)y) & 15) == 0,
               "activation_fwd: 16-byte aligned tensors, n %% 8 == 0, kind 0 (sigmoid) or 1 (swish)");
  act_fwd_kernel<<<sgrid(ctx, n / 8), kThreads, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y, n / 8, kind);
  SEGK_LAUNCHED(ctx, "activation_fwd");
  return SEGK_OK;
}

int segk_activation_bwd(segk_ctx* ctx, const void* x, const void* dy, void* dx, int64_t n, int kind, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && dy && dx && n > 0 && n % 8 == 0 && (kind == 0 || kind == 1) && (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx) & 15) == 0,
               "activation_bwd: 16-byte aligned tensors, n %% 8 == 0, kind 0 (sigmoid) or 1 (swish)");
  act_bwd_kernel<<<sgrid(ctx, n / 8), kThreads, 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)dy, (uint4*)dx, n / 8, kind);
  SEGK_LAUNCHED(ctx, "activation_bwd");
  return SEGK_OK;
}

int segk_channel_scale_fwd(segk_ctx* ctx, const void* x, const void* s, void* y, int N, int64_t HW, int C, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && s && y && N > 0 && HW > 0 && C > 0 && C % 8 == 0 && (((uintptr_t)x | (uintptr_t)s | (uintptr_t)y) & 15) == 0,
               "channel_scale_fwd: need C %% 8 == 0 and 16-byte aligned tensors");
  const int64_t total = (int64_t)N * HW * (C / 8);
  chscale_fwd_kernel<<<sgrid(ctx, total), kThreads, 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)s, (uint4*)y, total, HW, C / 8);
  SEGK_LAUNCHED(ctx, "channel_scale_fwd");
  return SEGK_OK;
}

int segk_channel_scale_bwd(segk_ctx* ctx, const void* x, const void* s, const void* dy, void* dx, float* ds, int N, int64_t HW, int C,
                           void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && s && dy && dx && ds && N > 0 && HW > 0 && C > 0 && C % 8 == 0 &&
                        (((uintptr_t)ds | (uintptr_t)x | (uintptr_t)s | (uintptr_t)dy | (uintptr_t)dx) & 15) == 0,
               "channel_scale_bwd: need C %% 8 == 0 and 16-byte aligned tensors");
  const int C8 = C / 8, groups = C8 < kThreads ? C8 : kThreads;
  chscale_bwd_kernel<<<dim3(ceil_div(C8, groups), N), kThreads, 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)s, (const uint4*)dy,
                                                                                            (uint4*)dx, ds, HW, C8);
  SEGK_LAUNCHED(ctx, "channel_scale_bwd");
  return SEGK_OK;
}

}  // extern "C"
