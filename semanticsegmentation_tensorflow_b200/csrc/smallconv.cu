// CUDA-core kernels for the layers whose channel counts cannot feed a tensor-core tile:
//   conv1_1 (Cin = 3/4, K = 27/36)            -> "tiny-K" direct conv fwd / wgrad
//   conv8   (1x1, Cout = num_classes = 2)      -> "skinny" 1x1 fwd / dgrad / wgrad
//   conv_t1 (Cin = 2) and, for now, conv_t3    -> gather-form transposed conv fwd/dgrad/wgrad
// They are bandwidth / latency bound and tiny next to the 3x3 stack.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------
// tiny-K conv forward: thread = (pixel, block of 32 output channels). Weights (fp32 HWIO
// [K][Cout]) staged in shared memory and read as broadcast float4.
// ---------------------------------------------------------------------------------------
constexpr int kTinyKMax = 64;

__device__ __forceinline__ float ld_in(const bf16* p) { return bf2f(*p); }
__device__ __forceinline__ float ld_in(const uint8_t* p) { return (float)*p; }

template <typename XT>
__global__ void __launch_bounds__(kThreads) conv_tinyk_fwd_kernel(
    const XT* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
    bf16* __restrict__ y, int N, int H, int W, int Cin, int Cout, int kh, int kw, int relu) {
  extern __shared__ float wsm[];  // [K][32] for this block's channel block
  const int K = kh * kw * Cin;
  const int cb = blockIdx.y * 32;
  for (int i = threadIdx.x; i < K * 32; i += blockDim.x) {
    const int k = i >> 5, c = i & 31;
    wsm[i] = (cb + c < Cout) ? w[(int64_t)k * Cout + cb + c] : 0.f;
  }
  __syncthreads();
  const int64_t npix = (int64_t)N * H * W;
  const int ph = kh / 2, pw = kw / 2;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix;
       p += (int64_t)gridDim.x * blockDim.x) {
    const int xw = (int)(p % W);
    const int yh = (int)((p / W) % H);
    const int n = (int)(p / ((int64_t)W * H));
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = bias ? ((cb + c < Cout) ? bias[cb + c] : 0.f) : 0.f;
    int k = 0;
    for (int ky = 0; ky < kh; ++ky) {
      const int yy = yh + ky - ph;
      for (int kx = 0; kx < kw; ++kx) {
        const int xx = xw + kx - pw;
        const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
        const XT* xp = x + (((int64_t)n * H + yy) * W + xx) * Cin;
        for (int ci = 0; ci < Cin; ++ci, ++k) {
          const float xv = in ? ld_in(xp + ci) : 0.f;
          const float4* wr = reinterpret_cast<const float4*>(wsm + k * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 wv = wr[q];
            acc[4 * q + 0] += xv * wv.x;
            acc[4 * q + 1] += xv * wv.y;
            acc[4 * q + 2] += xv * wv.z;
            acc[4 * q + 3] += xv * wv.w;
          }
        }
      }
    }
    bf16* yp = y + p * Cout + cb;
    if (cb + 32 <= Cout) {
      uint4* y4 = reinterpret_cast<uint4*>(yp);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = acc[8 * q + 2 * j], b = acc[8 * q + 2 * j + 1];
          if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
          o[j] = pack_bf16x2(a, b);
        }
        y4[q] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    } else {
      for (int c = 0; c < 32 && cb + c < Cout; ++c) yp[c] = f2bf(relu ? fmaxf(acc[c], 0.f) : acc[c]);
    }
  }
}

// tiny-K wgrad: dW[k][co] = sum_p x[p + tap(k)][ci(k)] * dy[p][co].  Block = 256 threads =
// 64 channels x 4 k-groups; each block reduces a contiguous pixel chunk, then atomicAdd.
template <typename XT>
__global__ void __launch_bounds__(kThreads) conv_tinyk_wgrad_kernel(
    const XT* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw, int N, int H,
    int W, int Cin, int Cout, int kh, int kw, int64_t pix_per_block) {
  constexpr int PT = 32;  // pixels staged per tile
  __shared__ float xs[PT][kTinyKMax];
  const int K = kh * kw * Cin;
  const int co = blockIdx.y * 64 + (threadIdx.x & 63);
  const int kg = threadIdx.x >> 6;              // 0..3
  const int kper = (K + 3) / 4;                 // <= 16
  const int k0 = kg * kper;
  const int64_t npix = (int64_t)N * H * W;
  const int64_t p0 = blockIdx.x * pix_per_block;
  const int64_t p1 = p0 + pix_per_block < npix ? p0 + pix_per_block : npix;
  const int ph = kh / 2, pw = kw / 2;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int64_t pt = p0; pt < p1; pt += PT) {
    __syncthreads();
    for (int i = threadIdx.x; i < PT * K; i += blockDim.x) {
      const int pp = i / K, k = i % K;
      const int64_t p = pt + pp;
      float v = 0.f;
      if (p < p1) {
        const int ci = k % Cin, t = k / Cin;
        const int ky = t / kw, kx = t % kw;
        const int xw = (int)(p % W), yh = (int)((p / W) % H);
        const int n = (int)(p / ((int64_t)W * H));
        const int yy = yh + ky - ph, xx = xw + kx - pw;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W)
          v = ld_in(x + (((int64_t)n * H + yy) * W + xx) * Cin + ci);
      }
      xs[pp][k] = v;
    }
    __syncthreads();
    if (co < Cout) {
      const int lim = (int)((p1 - pt) < PT ? (p1 - pt) : PT);
      for (int pp = 0; pp < lim; ++pp) {
        const float g = bf2f(dy[(pt + pp) * Cout + co]);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (i < kper && k0 + i < K) acc[i] += xs[pp][k0 + i] * g;
      }
    }
  }
  if (co < Cout) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < kper && k0 + i < K) atomicAdd(dw + (int64_t)(k0 + i) * Cout + co, acc[i]);
  }
}

// ---------------------------------------------------------------------------------------
// skinny 1x1 conv (Cout <= 8): one warp per pixel.
// ---------------------------------------------------------------------------------------
template <int CO>
__global__ void __launch_bounds__(kThreads) conv_skinny_fwd_kernel(
    const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
    void* __restrict__ y, int64_t npix, int Cin, int relu, int out_f32) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp; p < npix; p += nwarps) {
    float acc[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = 0.f;
    const bf16* xp = x + p * Cin;
    for (int ci = lane * 8; ci < Cin; ci += 256) {
      const uint4 xv = __ldg(reinterpret_cast<const uint4*>(xp + ci));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&xv.x)[j]);
        const float* w0 = w + (int64_t)(ci + 2 * j) * CO;
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[c] += f.x * __ldg(w0 + c) + f.y * __ldg(w0 + CO + c);
      }
    }
#pragma unroll
    for (int c = 0; c < CO; ++c)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < CO; ++c) {
        float v = acc[c] + (bias ? bias[c] : 0.f);
        v = relu ? fmaxf(v, 0.f) : v;
        if (out_f32) reinterpret_cast<float*>(y)[p * CO + c] = v;
        else reinterpret_cast<bf16*>(y)[p * CO + c] = f2bf(v);
      }
    }
  }
}

// dx[p][ci] = sum_co dy[p][co] * w[ci][co], masked by relu_mask[p][ci] > 0
template <int CO>
__global__ void __launch_bounds__(kThreads) conv_skinny_dgrad_kernel(
    const bf16* __restrict__ dy, const float* __restrict__ w, const bf16* __restrict__ mask,
    bf16* __restrict__ dx, int64_t npix, int Cin, float scale) {
  const int64_t total = npix * Cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / Cin;
    const int ci = (int)(i % Cin);
    float a = 0.f;
#pragma unroll
    for (int c = 0; c < CO; ++c) a += bf2f(dy[p * CO + c]) * __ldg(w + (int64_t)ci * CO + c);
    if (mask && !(bf2f(mask[i]) > 0.f)) a = 0.f;
    dx[i] = f2bf(a * scale);
  }
}

// dW[ci][co] += sum_p x[p][ci] dy[p][co]
template <int CO>
__global__ void __launch_bounds__(kThreads) conv_skinny_wgrad_kernel(
    const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw, int64_t npix,
    int Cin, int64_t pix_per_block) {
  const int ci = blockIdx.y * blockDim.x + threadIdx.x;
  if (ci >= Cin) return;
  const int64_t p0 = blockIdx.x * pix_per_block;
  const int64_t p1 = p0 + pix_per_block < npix ? p0 + pix_per_block : npix;
  float acc[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) acc[c] = 0.f;
  for (int64_t p = p0; p < p1; ++p) {
    const float xv = bf2f(x[p * Cin + ci]);
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] += xv * bf2f(dy[p * CO + c]);
  }
#pragma unroll
  for (int c = 0; c < CO; ++c) atomicAdd(dw + (int64_t)ci * CO + c, acc[c]);
}

// ---- skinny 1x1 conv on LARGE pixel counts with small Cin (the 64 -> num_classes head of the
// encoder-decoder builders at full resolution): HBM-bound streaming forms.
// fwd: one thread per pixel; the Cin x CO weights live in shared memory (fp32).
template <int CO>
__global__ void __launch_bounds__(kThreads) conv_skinny_fwd_pixel_kernel(
    const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
    void* __restrict__ y, int64_t npix, int Cin, int relu, int out_f32) {
  extern __shared__ float wsm[];   // [Cin][CO]
  for (int i = threadIdx.x; i < Cin * CO; i += blockDim.x) wsm[i] = w[i];
  __syncthreads();
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    float acc[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = bias ? bias[c] : 0.f;
    const uint4* xp = reinterpret_cast<const uint4*>(x + p * Cin);
    for (int g = 0; g < Cin / 8; ++g) {
      const uint4 v = __ldg(xp + g);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&v.x)[j]);
        const float* w0 = wsm + (g * 8 + 2 * j) * CO;
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[c] += f.x * w0[c] + f.y * w0[CO + c];
      }
    }
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      const float v = relu ? fmaxf(acc[c], 0.f) : acc[c];
      if (out_f32) reinterpret_cast<float*>(y)[p * CO + c] = v;
      else reinterpret_cast<bf16*>(y)[p * CO + c] = f2bf(v);
    }
  }
}

// dgrad: thread = 8 input channels of one pixel (16-byte store)
template <int CO>
__global__ void __launch_bounds__(kThreads) conv_skinny_dgrad_vec_kernel(
    const bf16* __restrict__ dy, const float* __restrict__ w, const bf16* __restrict__ mask,
    bf16* __restrict__ dx, int64_t npix, int Cin, float scale) {
  extern __shared__ float wsm[];   // [Cin][CO]
  for (int i = threadIdx.x; i < Cin * CO; i += blockDim.x) wsm[i] = w[i];
  __syncthreads();
  const int C8 = Cin >> 3;
  const int64_t total = npix * C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % C8);
    const int64_t p = i / C8;
    float d[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) d[c] = bf2f(dy[p * CO + c]);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = 0.f;
#pragma unroll
      for (int c = 0; c < CO; ++c) a += d[c] * wsm[(g * 8 + j) * CO + c];
      v[j] = a * scale;
    }
    if (mask) {
      const uint4 m = __ldg(reinterpret_cast<const uint4*>(mask + p * Cin) + g);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&m.x)[j]);
        if (!(f.x > 0.f)) v[2 * j] = 0.f;
        if (!(f.y > 0.f)) v[2 * j + 1] = 0.f;
      }
    }
    reinterpret_cast<uint4*>(dx + p * Cin)[g] =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

// wgrad stage 1: per-block partial dW[ci][co] over the block's pixels (thread = 8 channels of a pixel
// per iteration), stage 2 sums the partials in a fixed order (no atomic chains).
template <int CO>
__global__ void __launch_bounds__(kThreads) conv_skinny_wgrad_partial_kernel(
    const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ partial, int64_t npix, int Cin) {
  __shared__ float sh[kThreads][8 * CO + 1];
  const int C8 = Cin >> 3;                 // host guarantees C8 divides kThreads
  const int R = kThreads / C8;
  const int g = threadIdx.x % C8, rl = threadIdx.x / C8;
  float acc[8][CO];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[j][c] = 0.f;
  for (int64_t p = (int64_t)blockIdx.x * R + rl; p < npix; p += (int64_t)gridDim.x * R) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + p * Cin) + g);
    float d[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) d[c] = bf2f(dy[p * CO + c]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2((&v.x)[j]);
#pragma unroll
      for (int c = 0; c < CO; ++c) {
        acc[2 * j][c] += f.x * d[c];
        acc[2 * j + 1][c] += f.y * d[c];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int c = 0; c < CO; ++c) sh[threadIdx.x][j * CO + c] = acc[j][c];
  __syncthreads();
  if (rl == 0) {
    for (int k = 1; k < R; ++k)
#pragma unroll
      for (int j = 0; j < 8 * CO; ++j) sh[threadIdx.x][j] += sh[threadIdx.x + k * C8][j];
    float* out = partial + ((int64_t)blockIdx.x * Cin + g * 8) * CO;    // [block][ci][co]
#pragma unroll
    for (int j = 0; j < 8 * CO; ++j) out[j] = sh[threadIdx.x][j];
  }
}

// ---- skinny k x k conv (odd k, SAME) with Cout in {2,4}: the 3x3 64 -> num_classes head of the reference's
// SegNet (`SegNet.py:80`, Conv2D_Layer default 3x3) at full resolution.  Same three forms as the 1x1 head
// above with a tap loop; x is re-read kh*kw times through L1/L2, DRAM traffic stays one pass.
// Shared-memory weight layout of the two kernels below: [tap][8-channel group][8*CO + 4 pad] floats.  A
// quarter-warp reads the 8 groups of one pixel at once; without the pad their 16-byte loads sit 64 B apart
// and collide four ways on the banks.
template <int CO>
__device__ __forceinline__ void load_skinny_weights(float* wsm, const float* __restrict__ w, int taps, int Cin) {
  constexpr int kRow = 8 * CO + 4;
  const int C8 = Cin >> 3;
  for (int i = threadIdx.x; i < taps * Cin * CO; i += blockDim.x) {
    const int c = i % (8 * CO), gg = (i / (8 * CO)) % C8, t = i / (8 * CO * C8);
    wsm[(t * C8 + gg) * kRow + c] = w[i];
  }
}

// forward: 8 lanes per pixel, one 16-byte piece of its Cin*2-byte row at a time (a warp reads
// contiguous memory; one thread per pixel made every load touch 32 cache lines), partial sums over the
// lane's 8 channels, three shuffle steps to combine.
template <int CO>
__global__ void __launch_bounds__(kThreads) conv_skinny_kxk_fwd_kernel(
    const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, void* __restrict__ y,
    int N, int H, int W, int Cin, int kh, int kw, int relu, int out_f32) {
  extern __shared__ float wsm[];
  constexpr int kRow = 8 * CO + 4;
  load_skinny_weights<CO>(wsm, w, kh * kw, Cin);
  __syncthreads();
  const int C8 = Cin >> 3;                       // host: 8 % C8 == 0 or C8 % 8 == 0 handled by the loop below
  const int npix = N * H * W;                    // < 2^31 / Cin (host): 32-bit index arithmetic
  const int ph = kh / 2, pw = kw / 2;
  const int lane8 = threadIdx.x & 7;
  const int ppb = blockDim.x >> 3;               // pixels per block pass
  for (int p0 = blockIdx.x * ppb; p0 < npix; p0 += gridDim.x * ppb) {
    const int p = p0 + (threadIdx.x >> 3);
    const bool live = p < npix;
    const int xw = live ? p % W : 0, yh = live ? (p / W) % H : 0;
    float acc[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = 0.f;
    if (live) {
      for (int t = 0; t < kh * kw; ++t) {
        const int dyo = t / kw - ph, dxo = t % kw - pw;
        if ((unsigned)(yh + dyo) >= (unsigned)H || (unsigned)(xw + dxo) >= (unsigned)W) continue;
        const uint4* xp = reinterpret_cast<const uint4*>(x + (int64_t)(p + dyo * W + dxo) * Cin);
        for (int g = lane8; g < C8; g += 8) {
          const uint4 v = __ldg(xp + g);
          const float4* wt = reinterpret_cast<const float4*>(wsm + (t * C8 + g) * kRow);
          float xv[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2((&v.x)[j]);
            xv[2 * j] = f.x; xv[2 * j + 1] = f.y;
          }
#pragma unroll
          for (int q = 0; q < 2 * CO; ++q) {       // 8*CO weights = 2*CO float4
            const float4 ww = wt[q];
            const float wl[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[(4 * q + e) % CO] += xv[(4 * q + e) / CO] * wl[e];
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 2);
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 4);
    }
    if (live && lane8 == 0) {
#pragma unroll
      for (int c = 0; c < CO; ++c) {
        float v = acc[c] + (bias ? bias[c] : 0.f);
        if (relu) v = fmaxf(v, 0.f);
        if (out_f32) reinterpret_cast<float*>(y)[(int64_t)p * CO + c] = v;
        else reinterpret_cast<bf16*>(y)[(int64_t)p * CO + c] = f2bf(v);
      }
    }
  }
}

// dx[p][ci] = sum_taps sum_co dy[p - tap][co] * w[tap][ci][co]; thread = 8 input channels of one pixel
template <int CO>
__global__ void __launch_bounds__(kThreads) conv_skinny_kxk_dgrad_kernel(
    const bf16* __restrict__ dy, const float* __restrict__ w, const bf16* __restrict__ mask, bf16* __restrict__ dx,
    int N, int H, int W, int Cin, int kh, int kw, float scale) {
  extern __shared__ float wsm[];
  constexpr int kRow = 8 * CO + 4;
  load_skinny_weights<CO>(wsm, w, kh * kw, Cin);
  __syncthreads();
  const int C8 = Cin >> 3;
  const int total = N * H * W * C8;        // < 2^31 (host)
  const int ph = kh / 2, pw = kw / 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int g = i % C8, p = i / C8;
    const int xw = p % W, yh = (p / W) % H;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    for (int t = 0; t < kh * kw; ++t) {
      const int dyo = -(t / kw - ph), dxo = -(t % kw - pw);      // the output pixel this tap came from
      if ((unsigned)(yh + dyo) >= (unsigned)H || (unsigned)(xw + dxo) >= (unsigned)W) continue;
      const bf16* dp = dy + (int64_t)(p + dyo * W + dxo) * CO;
      float d[CO];
#pragma unroll
      for (int c = 0; c < CO; ++c) d[c] = bf2f(dp[c]);
      const float4* wt = reinterpret_cast<const float4*>(wsm + (t * C8 + g) * kRow);
#pragma unroll
      for (int q = 0; q < 2 * CO; ++q) {
        const float4 ww = wt[q];
        const float wl[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) v[(4 * q + e) / CO] += d[(4 * q + e) % CO] * wl[e];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= scale;
    if (mask) {
      const uint4 m = __ldg(reinterpret_cast<const uint4*>(mask + (int64_t)p * Cin) + g);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&m.x)[j]);
        if (!(f.x > 0.f)) v[2 * j] = 0.f;
        if (!(f.y > 0.f)) v[2 * j + 1] = 0.f;
      }
    }
    reinterpret_cast<uint4*>(dx + (int64_t)p * Cin)[g] =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

// wgrad stage 1 for one tap (blockIdx.y): per-block partial dW[tap][ci][co] over the block's pixels
template <int CO>
__global__ void __launch_bounds__(kThreads) conv_skinny_kxk_wgrad_partial_kernel(
    const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ partial, int N, int H, int W, int Cin,
    int kh, int kw) {
  __shared__ float sh[kThreads][8 * CO + 1];
  const int C8 = Cin >> 3;                 // host guarantees C8 divides kThreads
  const int R = kThreads / C8;
  const int g = threadIdx.x % C8, rl = threadIdx.x / C8;
  const int t = blockIdx.y;
  const int dyo = t / kw - kh / 2, dxo = t % kw - kw / 2;
  const int npix = N * H * W;              // < 2^31 / Cin (host)
  float acc[8][CO];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[j][c] = 0.f;
  // (x, y) of the pixel advance incrementally with the grid stride: no division in the loop
  const int S = gridDim.x * R, sx = S % W, sy = (S / W) % H;
  int p = blockIdx.x * R + rl;
  int xw = p % W, yh = (p / W) % H;
  for (; p < npix; p += S, xw += sx, yh += sy) {
    if (xw >= W) { xw -= W; ++yh; }
    while (yh >= H) yh -= H;
    if ((unsigned)(yh + dyo) >= (unsigned)H || (unsigned)(xw + dxo) >= (unsigned)W) continue;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (int64_t)(p + dyo * W + dxo) * Cin) + g);
    float d[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) d[c] = bf2f(dy[(int64_t)p * CO + c]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2((&v.x)[j]);
#pragma unroll
      for (int c = 0; c < CO; ++c) {
        acc[2 * j][c] += f.x * d[c];
        acc[2 * j + 1][c] += f.y * d[c];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int c = 0; c < CO; ++c) sh[threadIdx.x][j * CO + c] = acc[j][c];
  __syncthreads();
  if (rl == 0) {
    for (int k = 1; k < R; ++k)
#pragma unroll
      for (int j = 0; j < 8 * CO; ++j) sh[threadIdx.x][j] += sh[threadIdx.x + k * C8][j];
    // [block][tap][ci][co]
    float* out = partial + (((int64_t)blockIdx.x * gridDim.y + t) * Cin + g * 8) * CO;
#pragma unroll
    for (int j = 0; j < 8 * CO; ++j) out[j] = sh[threadIdx.x][j];
  }
}

__global__ void __launch_bounds__(kThreads) sum_partials_rows_kernel(const float* __restrict__ partial,
                                                                     float* __restrict__ out, int rows, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;      // four loads in flight; the order of the sum is fixed
  int r = 0;
  for (; r + 3 < rows; r += 4) {
    a0 += partial[(int64_t)r * n + i];
    a1 += partial[(int64_t)(r + 1) * n + i];
    a2 += partial[(int64_t)(r + 2) * n + i];
    a3 += partial[(int64_t)(r + 3) * n + i];
  }
  for (; r < rows; ++r) a0 += partial[(int64_t)r * n + i];
  out[i] = (a0 + a1) + (a2 + a3);
}

// ---------------------------------------------------------------------------------------
// transposed conv, k = 2s, SAME (FCN.py:138-159) on CUDA cores, gather form:
//   y[n,oy,ox,co] = b[co] + sum_{ty,tx in {0,1}} sum_ci x[n, qy-ty, qx-tx, ci] * W[ay+s*ty, ax+s*tx, co, ci]
//   with qy = (oy+p)/s, ay = (oy+p)%s, p = s/2  (SURVEY Appendix B.2).
// ---------------------------------------------------------------------------------------
template <typename OT>
__global__ void __launch_bounds__(kThreads) deconv_fwd_kernel(
    const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
    const bf16* __restrict__ res, OT* __restrict__ y, int N, int H, int W, int Cin, int Cout, int k,
    int s, int relu) {
  const int OH = H * s, OW = W * s, p = s / 2;
  const int64_t total = (int64_t)N * OH * OW * Cout;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    int64_t r = i / Cout;
    const int ox = (int)(r % OW);
    r /= OW;
    const int oy = (int)(r % OH);
    const int n = (int)(r / OH);
    const int qy = (oy + p) / s, ay = (oy + p) % s;
    const int qx = (ox + p) / s, ax = (ox + p) % s;
    float acc = bias ? bias[co] : 0.f;
#pragma unroll
    for (int ty = 0; ty < 2; ++ty) {
      const int iy = qy - ty;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int tx = 0; tx < 2; ++tx) {
        const int ix = qx - tx;
        if (ix < 0 || ix >= W) continue;
        const bf16* xp = x + (((int64_t)n * H + iy) * W + ix) * Cin;
        const float* wp = w + (((int64_t)(ay + s * ty) * k + (ax + s * tx)) * Cout + co) * Cin;
        if ((Cin & 7) == 0) {
          for (int ci = 0; ci < Cin; ci += 8) {
            const uint4 xv = __ldg(reinterpret_cast<const uint4*>(xp + ci));
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp + ci));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + ci + 4));
            float2 f;
            f = unpack_bf16x2(xv.x); acc += f.x * w0.x + f.y * w0.y;
            f = unpack_bf16x2(xv.y); acc += f.x * w0.z + f.y * w0.w;
            f = unpack_bf16x2(xv.z); acc += f.x * w1.x + f.y * w1.y;
            f = unpack_bf16x2(xv.w); acc += f.x * w1.z + f.y * w1.w;
          }
        } else {
          for (int ci = 0; ci < Cin; ++ci) acc += bf2f(xp[ci]) * __ldg(wp + ci);
        }
      }
    }
    if (res) acc += bf2f(res[i]);
    if (relu) acc = fmaxf(acc, 0.f);
    y[i] = (OT)acc;
  }
}

// dx[n,i,j,ci] = sum_{ky,kx,co} dy[n, i*s-p+ky, j*s-p+kx, co] * W[ky,kx,co,ci]; thread = (pixel, ci)
template <typename GT>
__global__ void __launch_bounds__(kThreads) deconv_dgrad_kernel(
    const GT* __restrict__ dy, const float* __restrict__ w, const bf16* __restrict__ mask,
    bf16* __restrict__ dx, int N, int H, int W, int Cin, int Cout, int k, int s) {
  const int OH = H * s, OW = W * s, p = s / 2;
  const int64_t total = (int64_t)N * H * W * Cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    int64_t r = i / Cin;
    const int ix = (int)(r % W);
    r /= W;
    const int iy = (int)(r % H);
    const int n = (int)(r / H);
    float acc = 0.f;
    for (int ky = 0; ky < k; ++ky) {
      const int oy = iy * s - p + ky;
      if (oy < 0 || oy >= OH) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ox = ix * s - p + kx;
        if (ox < 0 || ox >= OW) continue;
        const GT* gp = dy + (((int64_t)n * OH + oy) * OW + ox) * Cout;
        const float* wp = w + ((int64_t)(ky * k + kx) * Cout) * Cin + ci;
        for (int co = 0; co < Cout; ++co) acc += (float)gp[co] * __ldg(wp + (int64_t)co * Cin);
      }
    }
    if (mask && !(bf2f(mask[i]) > 0.f)) acc = 0.f;
    dx[i] = f2bf(acc);
  }
}

// Same gradient for tiny Cin (conv_t1: Cin = 2, Cout = 512): one WARP per (pixel, ci); lanes stride
// over co so the dy reads are coalesced, then a shuffle reduction.
template <typename GT>
__global__ void __launch_bounds__(kThreads) deconv_dgrad_warp_kernel(
    const GT* __restrict__ dy, const float* __restrict__ w, const bf16* __restrict__ mask,
    bf16* __restrict__ dx, int N, int H, int W, int Cin, int Cout, int k, int s) {
  const int OH = H * s, OW = W * s, p = s / 2;
  const int lane = threadIdx.x & 31;
  const int64_t total = (int64_t)N * H * W * Cin;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; i < total; i += nwarps) {
    const int ci = (int)(i % Cin);
    int64_t r = i / Cin;
    const int ix = (int)(r % W);
    r /= W;
    const int iy = (int)(r % H);
    const int n = (int)(r / H);
    float acc = 0.f;
    for (int ky = 0; ky < k; ++ky) {
      const int oy = iy * s - p + ky;
      if (oy < 0 || oy >= OH) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ox = ix * s - p + kx;
        if (ox < 0 || ox >= OW) continue;
        const GT* gp = dy + (((int64_t)n * OH + oy) * OW + ox) * Cout;
        const float* wp = w + ((int64_t)(ky * k + kx) * Cout) * Cin + ci;
        for (int co = lane; co < Cout; co += 32) acc += (float)gp[co] * __ldg(wp + (int64_t)co * Cin);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      if (mask && !(bf2f(mask[i]) > 0.f)) acc = 0.f;
      dx[i] = f2bf(acc);
    }
  }
}

// dW[ky,kx,co,ci] += sum_{n,i,j} x[n,i,j,ci] * dy[n, i*s-p+ky, j*s-p+kx, co]
// thread = one weight element (ci fastest); blockIdx.y = slice of the (n,i) rows.
template <typename GT>
__global__ void __launch_bounds__(kThreads) deconv_wgrad_kernel(
    const bf16* __restrict__ x, const GT* __restrict__ dy, float* __restrict__ dw, int N, int H,
    int W, int Cin, int Cout, int k, int s, int rows_per_slice) {
  const int OH = H * s, OW = W * s, p = s / 2;
  const int64_t nw = (int64_t)k * k * Cout * Cin;
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nw) return;
  const int ci = (int)(e % Cin);
  int64_t r = e / Cin;
  const int co = (int)(r % Cout);
  r /= Cout;
  const int kx = (int)(r % k), ky = (int)(r / k);
  const int row0 = blockIdx.y * rows_per_slice;
  const int row1 = min(row0 + rows_per_slice, N * H);
  float acc = 0.f;
  for (int row = row0; row < row1; ++row) {
    const int n = row / H, iy = row % H;
    const int oy = iy * s - p + ky;
    if (oy < 0 || oy >= OH) continue;
    const bf16* xp = x + ((int64_t)row * W) * Cin + ci;
    const GT* gp = dy + (((int64_t)n * OH + oy) * OW) * Cout + co;
    for (int ix = 0; ix < W; ++ix) {
      const int ox = ix * s - p + kx;
      if (ox < 0 || ox >= OW) continue;
      acc += bf2f(xp[(int64_t)ix * Cin]) * (float)gp[(int64_t)ox * Cout];
    }
  }
  atomicAdd(dw + e, acc);
}

// ---------------------------------------------------------------------------------------
// The FCN tail at training batch sizes: few pixels (N*5*18), wide channels.  conv8 (1x1, 4096 -> 2) and
// conv_t1 (4x4 s2, 2 -> 512) and their gradients are HBM-bound streams of 12-24 MB each; the per-element /
// per-warp forms above ran them at 0.2-0.4 TB/s.  These forms use 16-byte accesses and keep many loads in flight.
// ---------------------------------------------------------------------------------------
constexpr int kWidePx = 8;

// conv8 forward: a block handles 8 pixels at a time; thread t owns channel groups t and t + 256 (8 channels
// each, weights in registers), so every thread has 16 independent 16-byte loads in flight per tile.
template <int CO>
__global__ void __launch_bounds__(kThreads) conv_skinny_fwd_wide_kernel(
    const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, void* __restrict__ y,
    int64_t npix, int Cin, int relu, int out_f32) {
  __shared__ float red[kThreads / 32][kWidePx * CO];
  const int G = Cin >> 3;
  const int g0 = threadIdx.x, g1 = threadIdx.x + kThreads;
  float wr[2][8][CO];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int g = u ? g1 : g0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int c = 0; c < CO; ++c) wr[u][j][c] = g < G ? __ldg(w + (int64_t)(g * 8 + j) * CO + c) : 0.f;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t p0 = (int64_t)blockIdx.x * kWidePx; p0 < npix; p0 += (int64_t)gridDim.x * kWidePx) {
    uint4 xv[2][kWidePx];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int g = u ? g1 : g0;
#pragma unroll
      for (int q = 0; q < kWidePx; ++q)
        xv[u][q] = (g < G && p0 + q < npix) ? __ldg(reinterpret_cast<const uint4*>(x + (p0 + q) * Cin) + g)
                                            : make_uint4(0, 0, 0, 0);
    }
    float acc[kWidePx][CO];
#pragma unroll
    for (int q = 0; q < kWidePx; ++q) {
#pragma unroll
      for (int c = 0; c < CO; ++c) acc[q][c] = 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2((&xv[u][q].x)[j]);
#pragma unroll
          for (int c = 0; c < CO; ++c) acc[q][c] += f.x * wr[u][2 * j][c] + f.y * wr[u][2 * j + 1][c];
        }
    }
#pragma unroll
    for (int q = 0; q < kWidePx; ++q)
#pragma unroll
      for (int c = 0; c < CO; ++c) {
        float v = acc[q][c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][q * CO + c] = v;
      }
    __syncthreads();
    if (threadIdx.x < kWidePx * CO) {
      const int q = threadIdx.x / CO, c = threadIdx.x % CO;
      float v = bias ? bias[c] : 0.f;
#pragma unroll
      for (int k = 0; k < kThreads / 32; ++k) v += red[k][threadIdx.x];     // fixed order: deterministic
      v = relu ? fmaxf(v, 0.f) : v;
      if (p0 + q < npix) {
        if (out_f32) reinterpret_cast<float*>(y)[(p0 + q) * CO + c] = v;
        else reinterpret_cast<bf16*>(y)[(p0 + q) * CO + c] = f2bf(v);
      }
    }
    __syncthreads();
  }
}

// conv8 weight gradient: block = 32 channel groups (8 channels each) x 8 pixel lanes over a slice of the pixels;
// the pixel lanes are reduced through shared memory, per-slice partial sums [slices][Cin][CO] go to the workspace
// and are added in a fixed order by sum_partials_rows_kernel.
template <int CO>
__global__ void __launch_bounds__(kThreads) conv_skinny_wgrad_wide_kernel(
    const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ partial, int64_t npix, int Cin,
    int pix_per_slice) {
  __shared__ float sh[8][32][8 * CO + 1];
  const int G = Cin >> 3;
  const int gl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int g = blockIdx.x * 32 + gl;
  const int64_t p0 = (int64_t)blockIdx.y * pix_per_slice;
  const int64_t p1 = p0 + pix_per_slice < npix ? p0 + pix_per_slice : npix;
  float acc[8][CO];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[j][c] = 0.f;
  if (g < G) {
#pragma unroll 4
    for (int64_t p = p0 + pl; p < p1; p += 8) {
      const uint4 xv = __ldg(reinterpret_cast<const uint4*>(x + p * Cin) + g);
      float d[CO];
#pragma unroll
      for (int c = 0; c < CO; ++c) d[c] = bf2f(dy[p * CO + c]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&xv.x)[j]);
#pragma unroll
        for (int c = 0; c < CO; ++c) {
          acc[2 * j][c] += f.x * d[c];
          acc[2 * j + 1][c] += f.y * d[c];
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int c = 0; c < CO; ++c) sh[pl][gl][j * CO + c] = acc[j][c];
  __syncthreads();
  // 256 threads -> 32 groups x (8 * CO) values: thread t sums the 8 pixel lanes of value t % (8*CO) ... in order
  for (int e = threadIdx.x; e < 32 * 8 * CO; e += kThreads) {
    const int gg = e / (8 * CO), v = e % (8 * CO);
    if (blockIdx.x * 32 + gg >= G) continue;
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][gg][v];
    partial[((int64_t)blockIdx.y * Cin + (int64_t)(blockIdx.x * 32 + gg) * 8) * CO + v] = t;
  }
}

// conv_t1 forward (k = 2s, tiny Cin): thread = 8 output channels of one output pixel, 16-byte residual load
// and store; the <= 4 contributing input pixels and their CIN*8 weights each are read per thread.
template <int CIN>
__global__ void __launch_bounds__(kThreads) deconv_fwd_vec8_kernel(
    const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, const bf16* __restrict__ res,
    bf16* __restrict__ y, int N, int H, int W, int Cout, int k, int s, int relu) {
  const int OH = H * s, OW = W * s, p = s / 2, C8 = Cout >> 3;
  const int64_t total = (int64_t)N * OH * OW * C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    int64_t r = i / C8;
    const int ox = (int)(r % OW);
    r /= OW;
    const int oy = (int)(r % OH);
    const int n = (int)(r / OH);
    const int qy = (oy + p) / s, ay = (oy + p) % s;
    const int qx = (ox + p) / s, ax = (ox + p) % s;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias ? __ldg(bias + c8 * 8 + j) : 0.f;
#pragma unroll
    for (int ty = 0; ty < 2; ++ty) {
      const int iy = qy - ty;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int tx = 0; tx < 2; ++tx) {
        const int ix = qx - tx;
        if (ix < 0 || ix >= W) continue;
        const bf16* xp = x + (((int64_t)n * H + iy) * W + ix) * CIN;
        float xv[CIN];
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) xv[ci] = bf2f(xp[ci]);
        // w[(ky*k+kx)][co][ci]: 8 consecutive co x CIN values are contiguous -> 16-byte loads (scalar loads at a
        // 32*CIN-byte lane stride cost one L1 wavefront per sector: they, not HBM, bounded the first version)
        const float4* wp = reinterpret_cast<const float4*>(w + (((int64_t)(ay + s * ty) * k + (ax + s * tx)) * Cout + c8 * 8) * CIN);
        float wv[8 * CIN];
#pragma unroll
        for (int q = 0; q < 2 * CIN; ++q) {
          const float4 t4 = __ldg(wp + q);
          wv[4 * q] = t4.x; wv[4 * q + 1] = t4.y; wv[4 * q + 2] = t4.z; wv[4 * q + 3] = t4.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) acc[j] += xv[ci] * wv[j * CIN + ci];
      }
    }
    if (res) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(res) + i);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&u.x)[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);
    }
    reinterpret_cast<uint4*>(y)[i] = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]),
                                                 pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
  }
}

// conv_t1 input gradient: one warp per input pixel; lanes stride over groups of 8 output channels of each of
// the k*k taps (16-byte dy loads), CIN accumulators per lane, shuffle reduction.
template <int CIN>
__global__ void __launch_bounds__(kThreads) deconv_dgrad_pix_kernel(
    const bf16* __restrict__ dy, const float* __restrict__ w, const bf16* __restrict__ mask, bf16* __restrict__ dx, int N,
    int H, int W, int Cout, int k, int s) {
  const int OH = H * s, OW = W * s, p = s / 2, C8 = Cout >> 3;
  const int lane = threadIdx.x & 31;
  const int64_t npix = (int64_t)N * H * W;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; i < npix; i += nwarps) {
    const int ix = (int)(i % W), iy = (int)((i / W) % H), n = (int)(i / ((int64_t)W * H));
    float acc[CIN];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) acc[ci] = 0.f;
    for (int ky = 0; ky < k; ++ky) {
      const int oy = iy * s - p + ky;
      if (oy < 0 || oy >= OH) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ox = ix * s - p + kx;
        if (ox < 0 || ox >= OW) continue;
        const uint4* gp = reinterpret_cast<const uint4*>(dy + (((int64_t)n * OH + oy) * OW + ox) * Cout);
        const float* wp = w + ((int64_t)(ky * k + kx) * Cout) * CIN;
        for (int c8 = lane; c8 < C8; c8 += 32) {
          const uint4 g = __ldg(gp + c8);
          const float4* w4 = reinterpret_cast<const float4*>(wp + (int64_t)c8 * 8 * CIN);
          float wv[8 * CIN];
#pragma unroll
          for (int q = 0; q < 2 * CIN; ++q) {
            const float4 t4 = __ldg(w4 + q);
            wv[4 * q] = t4.x; wv[4 * q + 1] = t4.y; wv[4 * q + 2] = t4.z; wv[4 * q + 3] = t4.w;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2((&g.x)[j]);
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci)
              acc[ci] += f.x * wv[(2 * j) * CIN + ci] + f.y * wv[(2 * j + 1) * CIN + ci];
          }
        }
      }
    }
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[ci] += __shfl_xor_sync(0xffffffffu, acc[ci], o);
    if (lane < CIN) {
      float v = acc[0];
#pragma unroll
      for (int ci = 1; ci < CIN; ++ci) v = lane == ci ? acc[ci] : v;
      if (mask && !(bf2f(mask[i * CIN + lane]) > 0.f)) v = 0.f;
      dx[i * CIN + lane] = f2bf(v);
    }
  }
}

// conv_t1 weight gradient: thread = (tap, 8 output channels) over a slice of the input rows; per-slice partial
// sums [slices][k*k][Cout][CIN] in the workspace, added in a fixed order afterwards.
template <int CIN>
__global__ void __launch_bounds__(kThreads) deconv_wgrad_vec8_kernel(
    const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ partial, int N, int H, int W, int Cout,
    int k, int s, int rows_per_slice) {
  const int OH = H * s, OW = W * s, p = s / 2, C8 = Cout >> 3;
  const int e = blockIdx.x * kThreads + threadIdx.x;          // (tap, c8)
  if (e >= k * k * C8) return;
  const int c8 = e % C8, tap = e / C8;
  const int ky = tap / k, kx = tap % k;
  const int row0 = blockIdx.y * rows_per_slice;
  const int row1 = min(row0 + rows_per_slice, N * H);
  float acc[8][CIN];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) acc[j][ci] = 0.f;
  for (int row = row0; row < row1; ++row) {
    const int n = row / H, iy = row % H;
    const int oy = iy * s - p + ky;
    if (oy < 0 || oy >= OH) continue;
    const bf16* xp = x + (int64_t)row * W * CIN;
    const uint4* gp = reinterpret_cast<const uint4*>(dy + (((int64_t)n * OH + oy) * OW) * Cout) + c8;
#pragma unroll 6
    for (int ix = 0; ix < W; ++ix) {
      const int ox = ix * s - p + kx;
      if (ox < 0 || ox >= OW) continue;
      const uint4 g = __ldg(gp + (int64_t)ox * C8);
      float xv[CIN];
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) xv[ci] = bf2f(xp[ix * CIN + ci]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&g.x)[j]);
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          acc[2 * j][ci] += f.x * xv[ci];
          acc[2 * j + 1][ci] += f.y * xv[ci];
        }
      }
    }
  }
  // dW[ky,kx,co,ci]: the 8 x CIN values of this thread are contiguous
  float* out = partial + (int64_t)blockIdx.y * ((int64_t)k * k * Cout * CIN) + ((int64_t)tap * Cout + c8 * 8) * CIN;
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) out[j * CIN + ci] = acc[j][ci];
}

inline int sgrid(segk_ctx* ctx, int64_t items, int per_sm = 8) {
  int64_t b = ceil_div64(items, kThreads), cap = (int64_t)ctx->sm_count * per_sm;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

extern "C" {

int segk_conv2d_small_fwd(segk_ctx* ctx, const void* x, int x_dtype, const float* w, const float* bias,
                          void* y, int N, int H, int W, int Cin, int Cout, int kh, int kw, unsigned flags,
                          void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && w && y && N > 0 && H > 0 && W > 0, "conv_small_fwd: bad args");
  SEGK_REQUIRE(ctx, x_dtype == 0 || x_dtype == 2, "conv_small_fwd: x_dtype must be 0 (bf16) or 2 (u8)");
  const int relu = (flags & SEGK_EPI_RELU) ? 1 : 0;
  const int out_f32 = (flags & SEGK_EPI_OUT_F32) ? 1 : 0;
  const int K = kh * kw * Cin;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t npix = (int64_t)N * H * W;
  if (K <= kTinyKMax && (kh & 1) && (kw & 1) && !out_f32) {
    dim3 grid(sgrid(ctx, npix, 4), ceil_div(Cout, 32));
    if (x_dtype == 2)
      conv_tinyk_fwd_kernel<uint8_t><<<grid, kThreads, K * 32 * sizeof(float), st>>>(
          (const uint8_t*)x, w, bias, (bf16*)y, N, H, W, Cin, Cout, kh, kw, relu);
    else
      conv_tinyk_fwd_kernel<bf16><<<grid, kThreads, K * 32 * sizeof(float), st>>>(
          (const bf16*)x, w, bias, (bf16*)y, N, H, W, Cin, Cout, kh, kw, relu);
    SEGK_LAUNCHED(ctx, "conv_tinyk_fwd");
    return SEGK_OK;
  }
  if (x_dtype == 0 && kh == 1 && kw == 1 && Cin % 8 == 0 && Cin <= 256 && npix >= 65536 &&
      (Cout == 2 || Cout == 4)) {
    // full-resolution head: thread per pixel, weights in shared memory
    const int grid = sgrid(ctx, npix, 8);
    const size_t sm = sizeof(float) * (size_t)Cin * Cout;
    if (Cout == 2)
      conv_skinny_fwd_pixel_kernel<2><<<grid, kThreads, sm, st>>>((const bf16*)x, w, bias, y, npix, Cin, relu, out_f32);
    else
      conv_skinny_fwd_pixel_kernel<4><<<grid, kThreads, sm, st>>>((const bf16*)x, w, bias, y, npix, Cin, relu, out_f32);
    SEGK_LAUNCHED(ctx, "conv_skinny_fwd_pixel");
    return SEGK_OK;
  }
  if (x_dtype == 0 && kh == 1 && kw == 1 && Cin % 8 == 0 && Cin >= 512 && Cin / 8 <= 2 * kThreads &&
      (Cout == 2 || Cout == 4) && ctx->tail_wide) {
    // few pixels, wide channels (conv8 at training batch sizes): block per 8 pixels, weights in registers
    int64_t grid = ceil_div64(npix, kWidePx);
    if (grid > (int64_t)ctx->sm_count * 4) grid = (int64_t)ctx->sm_count * 4;
    if (Cout == 2)
      conv_skinny_fwd_wide_kernel<2><<<(unsigned)grid, kThreads, 0, st>>>((const bf16*)x, w, bias, y, npix, Cin, relu, out_f32);
    else
      conv_skinny_fwd_wide_kernel<4><<<(unsigned)grid, kThreads, 0, st>>>((const bf16*)x, w, bias, y, npix, Cin, relu, out_f32);
    SEGK_LAUNCHED(ctx, "conv_skinny_fwd_wide");
    return SEGK_OK;
  }
  if (x_dtype == 0 && kh == 1 && kw == 1 && Cin % 8 == 0 && (Cout == 2 || Cout == 4 || Cout == 8)) {
    const int grid = sgrid(ctx, npix * 32, 8);
    if (Cout == 2)
      conv_skinny_fwd_kernel<2><<<grid, kThreads, 0, st>>>((const bf16*)x, w, bias, y, npix, Cin, relu, out_f32);
    else if (Cout == 4)
      conv_skinny_fwd_kernel<4><<<grid, kThreads, 0, st>>>((const bf16*)x, w, bias, y, npix, Cin, relu, out_f32);
    else
      conv_skinny_fwd_kernel<8><<<grid, kThreads, 0, st>>>((const bf16*)x, w, bias, y, npix, Cin, relu, out_f32);
    SEGK_LAUNCHED(ctx, "conv_skinny_fwd");
    return SEGK_OK;
  }
  if (x_dtype == 0 && (kh & 1) && (kw & 1) && kh * kw <= 25 && Cin % 8 == 0 && (Cout == 2 || Cout == 4) &&
      (size_t)kh * kw * (Cin / 8) * (8 * Cout + 4) * sizeof(float) <= 48 * 1024 && npix * Cin < ((int64_t)1 << 31)) {
    const int grid = sgrid(ctx, npix * 8, 8);      // 8 lanes per pixel
    const size_t sm = sizeof(float) * (size_t)kh * kw * (Cin / 8) * (8 * Cout + 4);
    if (Cout == 2)
      conv_skinny_kxk_fwd_kernel<2><<<grid, kThreads, sm, st>>>((const bf16*)x, w, bias, y, N, H, W, Cin, kh, kw, relu, out_f32);
    else
      conv_skinny_kxk_fwd_kernel<4><<<grid, kThreads, sm, st>>>((const bf16*)x, w, bias, y, N, H, W, Cin, kh, kw, relu, out_f32);
    SEGK_LAUNCHED(ctx, "conv_skinny_kxk_fwd");
    return SEGK_OK;
  }
  return segk_fail(ctx, SEGK_EINVAL,
                   "conv_small_fwd: unsupported shape k=%dx%d Cin=%d Cout=%d dtype=%d (no fallback)", kh, kw,
                   Cin, Cout, x_dtype);
}

int segk_conv2d_small_dgrad(segk_ctx* ctx, const void* dy, const float* w, const void* relu_mask, void* dx,
                            float scale, int N, int H, int W, int Cin, int Cout, int kh, int kw,
                            void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, dy && w && dx && N > 0, "conv_small_dgrad: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t npix = (int64_t)N * H * W;
  if (kh * kw > 1) {
    SEGK_REQUIRE(ctx, (kh & 1) && (kw & 1) && kh * kw <= 25 && Cin % 8 == 0 && (Cout == 2 || Cout == 4) &&
                          (size_t)kh * kw * (Cin / 8) * (8 * Cout + 4) * sizeof(float) <= 48 * 1024 && npix * Cin < ((int64_t)1 << 31),
                 "conv_small_dgrad: k x k needs odd k <= 5, Cin %% 8 == 0, Cout in {2,4}, < 2^31 elements (got %dx%d %d -> %d)", kh, kw, Cin, Cout);
    const int g = sgrid(ctx, npix * (Cin / 8), 8);
    const size_t sm = sizeof(float) * (size_t)kh * kw * (Cin / 8) * (8 * Cout + 4);
    if (Cout == 2)
      conv_skinny_kxk_dgrad_kernel<2><<<g, kThreads, sm, st>>>((const bf16*)dy, w, (const bf16*)relu_mask, (bf16*)dx, N, H, W, Cin, kh, kw, scale);
    else
      conv_skinny_kxk_dgrad_kernel<4><<<g, kThreads, sm, st>>>((const bf16*)dy, w, (const bf16*)relu_mask, (bf16*)dx, N, H, W, Cin, kh, kw, scale);
    SEGK_LAUNCHED(ctx, "conv_skinny_kxk_dgrad");
    return SEGK_OK;
  }
  SEGK_REQUIRE(ctx, kh == 1 && kw == 1 && (Cout == 2 || Cout == 4 || Cout == 8),
               "conv_small_dgrad: only 1x1 with Cout in {2,4,8} (got %dx%d Cout=%d)", kh, kw, Cout);
  if (Cin % 8 == 0 && (Cout == 2 || Cout == 4) &&
      ((Cin <= 256 && npix >= 65536) || (ctx->tail_wide && Cin >= 512 && (size_t)Cin * Cout * sizeof(float) <= 48 * 1024))) {
    // (wide channels: every block stages Cin*Cout weights in shared memory, so fewer, longer-lived blocks)
    const int g = sgrid(ctx, npix * (Cin / 8), Cin >= 512 ? 3 : 8);
    const size_t sm = sizeof(float) * (size_t)Cin * Cout;
    if (Cout == 2)
      conv_skinny_dgrad_vec_kernel<2><<<g, kThreads, sm, st>>>((const bf16*)dy, w, (const bf16*)relu_mask, (bf16*)dx, npix, Cin, scale);
    else
      conv_skinny_dgrad_vec_kernel<4><<<g, kThreads, sm, st>>>((const bf16*)dy, w, (const bf16*)relu_mask, (bf16*)dx, npix, Cin, scale);
    SEGK_LAUNCHED(ctx, "conv_skinny_dgrad_vec");
    return SEGK_OK;
  }
  const int grid = sgrid(ctx, npix * Cin);
  if (Cout == 2)
    conv_skinny_dgrad_kernel<2><<<grid, kThreads, 0, st>>>((const bf16*)dy, w, (const bf16*)relu_mask, (bf16*)dx, npix, Cin, scale);
  else if (Cout == 4)
    conv_skinny_dgrad_kernel<4><<<grid, kThreads, 0, st>>>((const bf16*)dy, w, (const bf16*)relu_mask, (bf16*)dx, npix, Cin, scale);
  else
    conv_skinny_dgrad_kernel<8><<<grid, kThreads, 0, st>>>((const bf16*)dy, w, (const bf16*)relu_mask, (bf16*)dx, npix, Cin, scale);
  SEGK_LAUNCHED(ctx, "conv_skinny_dgrad");
  return SEGK_OK;
}

int segk_conv2d_small_wgrad(segk_ctx* ctx, const void* x, int x_dtype, const void* dy, float* dw, int N,
                            int H, int W, int Cin, int Cout, int kh, int kw, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && dy && dw && N > 0, "conv_small_wgrad: bad args");
  SEGK_REQUIRE(ctx, x_dtype == 0 || x_dtype == 2, "conv_small_wgrad: x_dtype must be 0 (bf16) or 2 (u8)");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t npix = (int64_t)N * H * W;
  const int K = kh * kw * Cin;
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)K * Cout, st);
  if (e != cudaSuccess) return segk_fail(ctx, SEGK_ECUDA, "wgrad memset: %s", cudaGetErrorString(e));
  if (x_dtype == 0 && kh == 1 && kw == 1 && (Cout == 2 || Cout == 4) && Cin % 8 == 0 &&
             kThreads % (Cin / 8) == 0 && npix >= 65536) {
    // full-resolution head: two-stage reduction over pixels
    const int R = kThreads / (Cin / 8);
    int64_t gx = ceil_div64(npix, (int64_t)R * 8);
    if (gx > (int64_t)ctx->sm_count * 4) gx = (int64_t)ctx->sm_count * 4;
    const size_t need = sizeof(float) * (size_t)gx * Cin * Cout;
    // own scratch (ws3): BiasAddGrad may run concurrently on another stream with ws2
    {
      const int rc = segk_grow(ctx, &ctx->ws3, &ctx->ws3_bytes, need < (size_t)(4 << 20) ? (size_t)(4 << 20) : need, "skinny wgrad");
      if (rc) return rc;
    }
    if (Cout == 2)
      conv_skinny_wgrad_partial_kernel<2><<<(unsigned)gx, kThreads, 0, st>>>((const bf16*)x, (const bf16*)dy, (float*)ctx->ws3, npix, Cin);
    else
      conv_skinny_wgrad_partial_kernel<4><<<(unsigned)gx, kThreads, 0, st>>>((const bf16*)x, (const bf16*)dy, (float*)ctx->ws3, npix, Cin);
    SEGK_LAUNCHED(ctx, "conv_skinny_wgrad_partial");
    const int n = Cin * Cout;
    sum_partials_rows_kernel<<<ceil_div(n, kThreads), kThreads, 0, st>>>((const float*)ctx->ws3, dw, (int)gx, n);
    SEGK_LAUNCHED(ctx, "conv_skinny_wgrad_sum");
  } else if (x_dtype == 0 && kh * kw > 1 && (kh & 1) && (kw & 1) && kh * kw <= 25 && (Cout == 2 || Cout == 4) &&
             Cin % 8 == 0 && kThreads % (Cin / 8) == 0 && K > kTinyKMax && npix * Cin < ((int64_t)1 << 31)) {
    // k x k head: two-stage reduction over pixels, one grid row per tap
    const int R = kThreads / (Cin / 8);
    const int T = kh * kw;
    int64_t gx = ceil_div64(npix, (int64_t)R * 8);
    const int64_t cap = ceil_div64((int64_t)ctx->sm_count * 4, T);
    if (gx > cap) gx = cap;
    const size_t need = sizeof(float) * (size_t)gx * K * Cout;
    {
      const int rc = segk_grow(ctx, &ctx->ws3, &ctx->ws3_bytes, need < (size_t)(4 << 20) ? (size_t)(4 << 20) : need, "skinny wgrad");
      if (rc) return rc;
    }
    dim3 grid((unsigned)gx, T);
    if (Cout == 2)
      conv_skinny_kxk_wgrad_partial_kernel<2><<<grid, kThreads, 0, st>>>((const bf16*)x, (const bf16*)dy, (float*)ctx->ws3, N, H, W, Cin, kh, kw);
    else
      conv_skinny_kxk_wgrad_partial_kernel<4><<<grid, kThreads, 0, st>>>((const bf16*)x, (const bf16*)dy, (float*)ctx->ws3, N, H, W, Cin, kh, kw);
    SEGK_LAUNCHED(ctx, "conv_skinny_kxk_wgrad_partial");
    const int n = K * Cout;
    sum_partials_rows_kernel<<<ceil_div(n, kThreads), kThreads, 0, st>>>((const float*)ctx->ws3, dw, (int)gx, n);
    SEGK_LAUNCHED(ctx, "conv_skinny_kxk_wgrad_sum");
  } else if (K <= kTinyKMax && (kh & 1) && (kw & 1)) {
    const int gy = ceil_div(Cout, 64);
    int64_t blocks = (int64_t)ctx->sm_count * 4 / gy;
    if (blocks < 1) blocks = 1;
    int64_t ppb = ceil_div64(npix, blocks);
    ppb = ceil_div64(ppb, 32) * 32;
    dim3 grid((unsigned)ceil_div64(npix, ppb), gy);
    if (x_dtype == 2)
      conv_tinyk_wgrad_kernel<uint8_t><<<grid, kThreads, 0, st>>>((const uint8_t*)x, (const bf16*)dy, dw, N, H,
                                                                  W, Cin, Cout, kh, kw, ppb);
    else
      conv_tinyk_wgrad_kernel<bf16><<<grid, kThreads, 0, st>>>((const bf16*)x, (const bf16*)dy, dw, N, H, W,
                                                               Cin, Cout, kh, kw, ppb);
    SEGK_LAUNCHED(ctx, "conv_tinyk_wgrad");
  } else if (x_dtype == 0 && kh == 1 && kw == 1 && (Cout == 2 || Cout == 4) && Cin % 8 == 0 && Cin >= 512 && ctx->tail_wide) {
    // few pixels, wide channels (conv8): 16-byte loads, per-slice partial sums + ordered reduction
    const int G = Cin / 8;
    int slices = ceil_div(2 * ctx->sm_count, ceil_div(G, 32));       // ~two blocks per SM
    if (slices > 32) slices = 32;
    if ((int64_t)slices * 8 > npix) slices = (int)ceil_div64(npix, 8);
    const int pps = (int)ceil_div64(npix, slices);
    slices = (int)ceil_div64(npix, pps);
    const int n = Cin * Cout;
    {
      const size_t need = sizeof(float) * (size_t)slices * n;
      const int rc = segk_grow(ctx, &ctx->ws3, &ctx->ws3_bytes, need < (size_t)(4 << 20) ? (size_t)(4 << 20) : need, "skinny wgrad");
      if (rc) return rc;
    }
    dim3 grid(ceil_div(G, 32), slices);
    if (Cout == 2)
      conv_skinny_wgrad_wide_kernel<2><<<grid, kThreads, 0, st>>>((const bf16*)x, (const bf16*)dy, (float*)ctx->ws3, npix, Cin, pps);
    else
      conv_skinny_wgrad_wide_kernel<4><<<grid, kThreads, 0, st>>>((const bf16*)x, (const bf16*)dy, (float*)ctx->ws3, npix, Cin, pps);
    SEGK_LAUNCHED(ctx, "conv_skinny_wgrad_wide");
    sum_partials_rows_kernel<<<ceil_div(n, kThreads), kThreads, 0, st>>>((const float*)ctx->ws3, dw, slices, n);
    SEGK_LAUNCHED(ctx, "conv_skinny_wgrad_sum");
  } else if (x_dtype == 0 && kh == 1 && kw == 1 && (Cout == 2 || Cout == 4 || Cout == 8)) {
    const int tx = 128;
    const int gy = ceil_div(Cin, tx);
    int64_t blocks = (int64_t)ctx->sm_count * 8 / gy;
    if (blocks < 1) blocks = 1;
    int64_t ppb = ceil_div64(npix, blocks);
    if (ppb < 8) ppb = 8;
    dim3 grid((unsigned)ceil_div64(npix, ppb), gy);
    if (Cout == 2)
      conv_skinny_wgrad_kernel<2><<<grid, tx, 0, st>>>((const bf16*)x, (const bf16*)dy, dw, npix, Cin, ppb);
    else if (Cout == 4)
      conv_skinny_wgrad_kernel<4><<<grid, tx, 0, st>>>((const bf16*)x, (const bf16*)dy, dw, npix, Cin, ppb);
    else
      conv_skinny_wgrad_kernel<8><<<grid, tx, 0, st>>>((const bf16*)x, (const bf16*)dy, dw, npix, Cin, ppb);
    SEGK_LAUNCHED(ctx, "conv_skinny_wgrad");
  } else {
    return segk_fail(ctx, SEGK_EINVAL,
                     "conv_small_wgrad: unsupported shape k=%dx%d Cin=%d Cout=%d (no fallback)", kh, kw,
                     Cin, Cout);
  }
  return SEGK_OK;
}

int segk_deconv2d_small_fwd(segk_ctx* ctx, const void* x, const float* w, const float* bias,
                            const void* residual, void* y, int N, int H, int W, int Cin, int Cout,
                            int k, int s, unsigned flags, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && w && y && N > 0, "deconv_small_fwd: bad args");
  SEGK_REQUIRE(ctx, k == 2 * s && (s % 2 == 0), "deconv: need k == 2*stride, even stride (k=%d s=%d)", k, s);
  const int relu = (flags & SEGK_EPI_RELU) ? 1 : 0;
  const int64_t total = (int64_t)N * H * s * W * s * Cout;
  cudaStream_t st = (cudaStream_t)stream;
  if (!(flags & SEGK_EPI_OUT_F32) && Cout % 8 == 0 && (Cin == 2 || Cin == 4) && ctx->tail_wide &&
      (((uintptr_t)y | (uintptr_t)residual) & 15) == 0) {
    const int g = sgrid(ctx, total / 8, 16);
    if (Cin == 2)
      deconv_fwd_vec8_kernel<2><<<g, kThreads, 0, st>>>((const bf16*)x, w, bias, (const bf16*)residual, (bf16*)y, N, H, W, Cout, k, s, relu);
    else
      deconv_fwd_vec8_kernel<4><<<g, kThreads, 0, st>>>((const bf16*)x, w, bias, (const bf16*)residual, (bf16*)y, N, H, W, Cout, k, s, relu);
    SEGK_LAUNCHED(ctx, "deconv_small_fwd_vec8");
    return SEGK_OK;
  }
  if (flags & SEGK_EPI_OUT_F32)
    deconv_fwd_kernel<float><<<sgrid(ctx, total, 16), kThreads, 0, st>>>(
        (const bf16*)x, w, bias, (const bf16*)residual, (float*)y, N, H, W, Cin, Cout, k, s, relu);
  else
    deconv_fwd_kernel<bf16><<<sgrid(ctx, total, 16), kThreads, 0, st>>>(
        (const bf16*)x, w, bias, (const bf16*)residual, (bf16*)y, N, H, W, Cin, Cout, k, s, relu);
  SEGK_LAUNCHED(ctx, "deconv_small_fwd");
  return SEGK_OK;
}

int segk_deconv2d_small_dgrad(segk_ctx* ctx, const void* dy, int dy_is_f32, const float* w,
                              const void* relu_mask, void* dx, int N, int H, int W, int Cin, int Cout,
                              int k, int s, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, dy && w && dx && N > 0, "deconv_small_dgrad: bad args");
  SEGK_REQUIRE(ctx, k == 2 * s && (s % 2 == 0), "deconv: need k == 2*stride, even stride (k=%d s=%d)", k, s);
  const int64_t total = (int64_t)N * H * W * Cin;
  cudaStream_t st = (cudaStream_t)stream;
  if (!dy_is_f32 && Cout % 8 == 0 && (Cin == 2 || Cin == 4) && ctx->tail_wide && (((uintptr_t)dy) & 15) == 0) {
    const int g = sgrid(ctx, (int64_t)N * H * W * 32, 16);
    if (Cin == 2)
      deconv_dgrad_pix_kernel<2><<<g, kThreads, 0, st>>>((const bf16*)dy, w, (const bf16*)relu_mask, (bf16*)dx, N, H, W, Cout, k, s);
    else
      deconv_dgrad_pix_kernel<4><<<g, kThreads, 0, st>>>((const bf16*)dy, w, (const bf16*)relu_mask, (bf16*)dx, N, H, W, Cout, k, s);
    SEGK_LAUNCHED(ctx, "deconv_small_dgrad_pix");
    return SEGK_OK;
  }
  if (Cin <= 8 && Cout >= 64) {
    if (dy_is_f32)
      deconv_dgrad_warp_kernel<float><<<sgrid(ctx, total * 32, 16), kThreads, 0, st>>>(
          (const float*)dy, w, (const bf16*)relu_mask, (bf16*)dx, N, H, W, Cin, Cout, k, s);
    else
      deconv_dgrad_warp_kernel<bf16><<<sgrid(ctx, total * 32, 16), kThreads, 0, st>>>(
          (const bf16*)dy, w, (const bf16*)relu_mask, (bf16*)dx, N, H, W, Cin, Cout, k, s);
    SEGK_LAUNCHED(ctx, "deconv_small_dgrad_warp");
    return SEGK_OK;
  }
  if (dy_is_f32)
    deconv_dgrad_kernel<float><<<sgrid(ctx, total, 16), kThreads, 0, st>>>(
        (const float*)dy, w, (const bf16*)relu_mask, (bf16*)dx, N, H, W, Cin, Cout, k, s);
  else
    deconv_dgrad_kernel<bf16><<<sgrid(ctx, total, 16), kThreads, 0, st>>>(
        (const bf16*)dy, w, (const bf16*)relu_mask, (bf16*)dx, N, H, W, Cin, Cout, k, s);
  SEGK_LAUNCHED(ctx, "deconv_small_dgrad");
  return SEGK_OK;
}

int segk_deconv2d_small_wgrad(segk_ctx* ctx, const void* x, const void* dy, int dy_is_f32, float* dw, int N,
                              int H, int W, int Cin, int Cout, int k, int s, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && dy && dw && N > 0, "deconv_small_wgrad: bad args");
  SEGK_REQUIRE(ctx, k == 2 * s && (s % 2 == 0), "deconv: need k == 2*stride, even stride (k=%d s=%d)", k, s);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nw = (int64_t)k * k * Cout * Cin;
  if (!dy_is_f32 && Cout % 8 == 0 && (Cin == 2 || Cin == 4) && ctx->tail_wide && (((uintptr_t)dy) & 15) == 0) {
    // tiny Cin (conv_t1): 16-byte dy loads, per-slice partial sums + ordered reduction (no atomics)
    const int rows = N * H;
    int slices = rows < 40 ? rows : 40;
    const int rps = ceil_div(rows, slices);
    slices = ceil_div(rows, rps);
    {
      const size_t need = sizeof(float) * (size_t)slices * nw;
      const int rc = segk_grow(ctx, &ctx->ws3, &ctx->ws3_bytes, need < (size_t)(4 << 20) ? (size_t)(4 << 20) : need, "deconv wgrad");
      if (rc) return rc;
    }
    dim3 grid(ceil_div(k * k * (Cout / 8), kThreads), slices);
    if (Cin == 2)
      deconv_wgrad_vec8_kernel<2><<<grid, kThreads, 0, st>>>((const bf16*)x, (const bf16*)dy, (float*)ctx->ws3, N, H, W, Cout, k, s, rps);
    else
      deconv_wgrad_vec8_kernel<4><<<grid, kThreads, 0, st>>>((const bf16*)x, (const bf16*)dy, (float*)ctx->ws3, N, H, W, Cout, k, s, rps);
    SEGK_LAUNCHED(ctx, "deconv_small_wgrad_vec8");
    sum_partials_rows_kernel<<<ceil_div((int)nw, kThreads), kThreads, 0, st>>>((const float*)ctx->ws3, dw, slices, (int)nw);
    SEGK_LAUNCHED(ctx, "deconv_small_wgrad_sum");
    return SEGK_OK;
  }
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)nw, st);
  if (e != cudaSuccess) return segk_fail(ctx, SEGK_ECUDA, "wgrad memset: %s", cudaGetErrorString(e));
  const int gx = (int)ceil_div64(nw, kThreads);
  int slices = ceil_div(ctx->sm_count * 8, gx);
  if (slices < 1) slices = 1;
  if (slices > N * H) slices = N * H;
  const int rps = ceil_div(N * H, slices);
  dim3 grid(gx, ceil_div(N * H, rps));
  if (dy_is_f32)
    deconv_wgrad_kernel<float><<<grid, kThreads, 0, st>>>((const bf16*)x, (const float*)dy, dw, N, H, W, Cin, Cout, k, s, rps);
  else
    deconv_wgrad_kernel<bf16><<<grid, kThreads, 0, st>>>((const bf16*)x, (const bf16*)dy, dw, N, H, W, Cin, Cout, k, s, rps);
  SEGK_LAUNCHED(ctx, "deconv_small_wgrad");
  return SEGK_OK;
}

}  // extern "C"
