// Layout-changing helper kernels that turn the ragged-channel layers into tensor-core GEMMs:
//   conv1_1 (Cin = 3/4, K = 27/36)  : im2col of the u8 image into a [N,H,W,64] bf16 patch tensor
//                                     (K zero-padded to one 64-wide k-step) -> 1x1 igemm / wgrad
//   conv_t3 (16x16 s8, Cout = 2)    : forward  = 1x1 GEMM into patch space [N,H,W,k*k*Cout] fp32
//                                                + col2im gather (4 overlapping patches / pixel);
//                                     backward = gather dlogits into the same patch space (bf16)
//                                                -> dgrad and wgrad are plain 1x1 GEMMs.
// All HBM-bound streaming kernels: 16-byte stores, grid-stride, grid = multiple of the SM count.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

inline int sgrid(segk_ctx* ctx, int64_t items, int per_sm = 8) {
  int64_t b = ceil_div64(items, kThreads), cap = (int64_t)ctx->sm_count * per_sm;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

__device__ __forceinline__ float ld_f(const bf16* p) { return bf2f(*p); }
__device__ __forceinline__ float ld_f(const uint8_t* p) { return (float)*p; }
__device__ __forceinline__ float ld_f(const float* p) { return *p; }

// P[n,y,x,kk] = kk < K ? x[n, y+ky-ph, x+kx-pw, ci] : 0,  kk = (ky*kw+kx)*Cin+ci ; thread = 8 kk of one
// pixel (8 lanes write one 128-byte patch row).  The (ky,kx,ci) decode of a thread's 8 kk is hoisted
// out of the grid-stride loop (the stride is a multiple of 8, so a thread keeps its kk group).
template <typename XT>
__global__ void __launch_bounds__(kThreads) im2col_k64_kernel(const XT* __restrict__ x, uint4* __restrict__ P,
                                                              int N, int H, int W, int Cin, int kh, int kw) {
  const int K = kh * kw * Cin;
  const int ph = kh / 2, pw = kw / 2;
  const int g = threadIdx.x & 7;
  int ody[8], odx[8], oci[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int kk = g * 8 + j;
    if (kk < K) {
      const int t = kk / Cin;
      oci[j] = kk % Cin;
      ody[j] = t / kw - ph;
      odx[j] = t % kw - pw;
    } else {
      oci[j] = -1; ody[j] = 0; odx[j] = 0;
    }
  }
  const int64_t npix = (int64_t)N * H * W;
  for (int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3; p < npix;
       p += ((int64_t)gridDim.x * blockDim.x) >> 3) {
    const int xw = (int)(p % W);
    const int yh = (int)((p / W) % H);
    const XT* base = x + p * Cin;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int yy = yh + ody[j], xx = xw + odx[j];
      v[j] = (oci[j] >= 0 && yy >= 0 && yy < H && xx >= 0 && xx < W)
                 ? ld_f(base + ((int64_t)ody[j] * W + odx[j]) * Cin + oci[j]) : 0.f;
    }
    P[p * 8 + g] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                              pack_bf16x2(v[6], v[7]));
  }
}

// wk[co][kk] (bf16, 64 wide) = kk < K ? w[kk][co] : 0
__global__ void __launch_bounds__(kThreads) pack_im2col_weights_kernel(const float* __restrict__ w,
                                                                       bf16* __restrict__ wk, int K, int Cout) {
  const int total = Cout * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kk = i & 63, co = i >> 6;
    wk[i] = f2bf(kk < K ? w[(int64_t)kk * Cout + co] : 0.f);
  }
}

// P[n,i,j,(ky,kx,co)] = dy[n, s*i-p+ky, s*j-p+kx, co] (0 outside) ; thread = 8 consecutive elements
template <typename GT>
__global__ void __launch_bounds__(kThreads) patch_gather_kernel(const GT* __restrict__ dy, uint4* __restrict__ P,
                                                                int N, int H, int W, int Co, int k, int s) {
  const int OH = H * s, OW = W * s, p = s / 2;
  const int E = k * k * Co;       // multiple of 8
  const int E8 = E >> 3;
  const int64_t total = (int64_t)N * H * W * E8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % E8);
    int64_t r = i / E8;
    const int j = (int)(r % W);
    r /= W;
    const int ii = (int)(r % H);
    const int n = (int)(r / H);
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int e = g * 8 + q;
      const int co = e % Co, t = e / Co;
      const int oy = ii * s - p + t / k, ox = j * s - p + t % k;
      v[q] = (oy >= 0 && oy < OH && ox >= 0 && ox < OW) ? ld_f(dy + (((int64_t)n * OH + oy) * OW + ox) * Co + co) : 0.f;
    }
    P[i] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                      pack_bf16x2(v[6], v[7]));
  }
}

// y[n,oy,ox,co] = b[co] + sum_{ty,tx in {0,1}} Yp[n, qy-ty, qx-tx, ((ay+s*ty)*k + ax+s*tx)*Co + co]
//   qy = (oy+p)/s, ay = (oy+p)%s  (k = 2s, SAME: SURVEY Appendix B.2)
template <typename OT>
__global__ void __launch_bounds__(kThreads) col2im_kernel(const float* __restrict__ yp, const float* __restrict__ bias,
                                                          const bf16* __restrict__ res, OT* __restrict__ y, int N,
                                                          int H, int W, int Co, int k, int s) {
  const int OH = H * s, OW = W * s, p = s / 2;
  const int E = k * k * Co;
  const int64_t total = (int64_t)N * OH * OW * Co;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % Co);
    int64_t r = i / Co;
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
        acc += __ldg(yp + (((int64_t)n * H + iy) * W + ix) * E + ((ay + s * ty) * k + ax + s * tx) * Co + co);
      }
    }
    if (res) acc += bf2f(res[i]);
    y[i] = (OT)acc;
  }
}


// out[r][c] = in[r][c] * scale[c] * mult   (BN-affine gamma/sqrt(1+eps) folded into conv weights and
// back out of their gradients, utils.py:300-301)
__global__ void __launch_bounds__(kThreads) scale_columns_kernel(const float* __restrict__ in,
                                                                 const float* __restrict__ scale, float mult,
                                                                 float* __restrict__ out, int64_t n, int C) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = in[i] * __ldg(scale + (int)(i % C)) * mult;
}

// Gradients of the folded BN scale from the WEIGHT gradient (no pass over activations):  y = s_c (W * x)_c + beta_c with
// s_c = gamma_c * mult is computed as a conv with W'[k][c] = W[k][c] s_c, and the wgrad kernels produce dW'.  Then
//   dL/ds_c = sum_pix dz_c (W * x)_c = sum_k W[k][c] dW'[k][c]      ->  dgamma[c] = mult * sum_k W[k][c] dW'[k][c]
//   dL/dW[k][c] = dW'[k][c] s_c                                      (in place)
// Two stages, fixed summation order (deterministic): grid (C/32, S) blocks each take a slice of the rows (8 row lanes x 4
// rows in flight) and write one partial row; the finish kernel adds the S partial rows.
__global__ void __launch_bounds__(kThreads) bn_unfold_grads_kernel(float* __restrict__ gw, const float* __restrict__ w,
                                                                   const float* __restrict__ gamma, float mult,
                                                                   float* __restrict__ partial, int64_t rows, int C,
                                                                   int rows_per_block) {
  __shared__ float sh[8][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    const float s = __ldg(gamma + c) * mult;
    for (int64_t r = r0 + rl; r < r1; r += 32) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t rr = r + 8 * j;
        if (rr < r1) {
          const float g = gw[rr * C + c];
          acc[j] += g * __ldg(w + rr * C + c);
          gw[rr * C + c] = g * s;
        }
      }
    }
  }
  sh[rl][cl] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  __syncthreads();
  if (rl == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][cl];
    partial[(int64_t)blockIdx.y * C + c] = t;
  }
}

__global__ void __launch_bounds__(kThreads) bn_unfold_finish_kernel(const float* __restrict__ partial, float* __restrict__ dgamma,
                                                                    float mult, int S, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float t = 0.f;
  for (int s = 0; s < S; ++s) t += partial[(int64_t)s * C + c];
  dgamma[c] = t * mult;
}

// partial[b][0][c] = sum over the block's rows of dz[r][c] * (y[r][c] - beta[c])
// partial[b][1][c] = sum over the block's rows of dz[r][c]            (BiasAddGrad = d beta, when kBeta)
// (fixed-order second stage: bn_gamma_finish_kernel divides the first by gamma)
template <bool kBeta>
__global__ void __launch_bounds__(kThreads) bn_gamma_partial_kernel(const uint4* __restrict__ dz,
                                                                    const uint4* __restrict__ y,
                                                                    const float* __restrict__ beta,
                                                                    float* __restrict__ partial, int64_t rows, int C8) {
  __shared__ float sh[kThreads][kBeta ? 17 : 9];
  const int cpb = C8 < kThreads ? C8 : kThreads;
  const int R = kThreads / cpb;
  const int cg = blockIdx.y * cpb + (threadIdx.x % cpb);
  const int rl = threadIdx.x / cpb;
  float acc[8], sum[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[j] = 0.f; sum[j] = 0.f; b[j] = (cg < C8) ? __ldg(beta + cg * 8 + j) : 0.f; }
  if (cg < C8 && rl < R) {
    for (int64_t r = (int64_t)blockIdx.x * R + rl; r < rows; r += (int64_t)gridDim.x * R) {
      const uint4 g = __ldg(dz + r * C8 + cg), v = __ldg(y + r * C8 + cg);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 gf = unpack_bf16x2((&g.x)[j]), vf = unpack_bf16x2((&v.x)[j]);
        acc[2 * j] += gf.x * (vf.x - b[2 * j]);
        acc[2 * j + 1] += gf.y * (vf.y - b[2 * j + 1]);
        if (kBeta) { sum[2 * j] += gf.x; sum[2 * j + 1] += gf.y; }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sh[threadIdx.x][j] = acc[j];
    if (kBeta) sh[threadIdx.x][8 + j] = sum[j];
  }
  __syncthreads();
  if (rl == 0 && cg < C8) {
    for (int k = 1; k < R; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] += sh[threadIdx.x + k * cpb][j];
        if (kBeta) sum[j] += sh[threadIdx.x + k * cpb][8 + j];
      }
    const int C = C8 * 8;
    float* out = partial + (int64_t)blockIdx.x * (kBeta ? 2 : 1) * C + cg * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      out[j] = acc[j];
      if (kBeta) out[C + j] = sum[j];
    }
  }
}

// block = 32 channels x 8 lanes over the partial rows (a single thread per channel walking all the
// partial rows was latency-bound: 118 us for a 24 MB layer)
__global__ void __launch_bounds__(kThreads) bn_gamma_finish_kernel(const float* __restrict__ partial,
                                                                   const float* __restrict__ gamma,
                                                                   float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                   int nparts, int C) {
  __shared__ float sa[8][33], sb[8][33];
  const int cx = threadIdx.x & 31, l8 = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int stride = dbeta ? 2 * C : C;
  float a = 0.f, s = 0.f;
  if (c < C) {
#pragma unroll 4
    for (int r = l8; r < nparts; r += 8) {
      a += partial[(int64_t)r * stride + c];
      if (dbeta) s += partial[(int64_t)r * stride + C + c];
    }
  }
  sa[l8][cx] = a;
  sb[l8][cx] = s;
  __syncthreads();
  if (l8 == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < 8; ++i) { a += sa[i][cx]; s += sb[i][cx]; }
    dgamma[c] = a / gamma[c];
    if (dbeta) dbeta[c] = s;
  }
}

// BN-affine gradients of an fp32 head with few channels (SegNet's conv26 + Batch_Normalization on the
// logits, SegNet.py:80-81): partial[b][0][c] = sum dz*(y - beta), partial[b][1][c] = sum dz; thread = one row.
template <int C>
__global__ void __launch_bounds__(kThreads) bn_grads_f32_partial_kernel(const float* __restrict__ dz, const float* __restrict__ y,
                                                                        const float* __restrict__ beta,
                                                                        float* __restrict__ partial, int64_t rows) {
  __shared__ float sh[kThreads / 32][2 * C];
  float a[C], s[C], b[C];
#pragma unroll
  for (int c = 0; c < C; ++c) { a[c] = 0.f; s[c] = 0.f; b[c] = __ldg(beta + c); }
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float g = dz[r * C + c];
      a[c] += g * (y[r * C + c] - b[c]);
      s[c] += g;
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c)
    for (int o = 16; o > 0; o >>= 1) {
      a[c] += __shfl_xor_sync(0xffffffffu, a[c], o);
      s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
    }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < C; ++c) { sh[warp][c] = a[c]; sh[warp][C + c] = s[c]; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * C) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += sh[w][threadIdx.x];
    partial[(int64_t)blockIdx.x * 2 * C + threadIdx.x] = t;      // [block][{dgamma*gamma, dbeta}][C]
  }
}

// dst[r][coff_dst + c] (= or +=) src[r][coff_src + c], optionally zeroed where mask[r][c] <= 0; 8 channels/thread
__global__ void __launch_bounds__(kThreads) channel_copy_kernel(const bf16* __restrict__ src, int ld_src, int coff_src,
                                                                bf16* __restrict__ dst, int ld_dst, int coff_dst,
                                                                const bf16* __restrict__ mask, int accumulate,
                                                                int64_t rows, int C8, int drop_side, const uint2* drop_mask,
                                                                float drop_keep, float drop_inv_keep, uint64_t drop_seed) {
  const int64_t total = rows * C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    const int64_t r = i / C8;
    uint4 v = *reinterpret_cast<const uint4*>(src + r * ld_src + coff_src + c);
    if (drop_side) {
      // dropout of the copied tensor on the fly (drop_side 1: the pattern is indexed by the SOURCE tensor's elements --
      // forward, conv output -> concat slot; 2: by the DESTINATION's -- backward, slot gradient -> conv output gradient)
      const int64_t e8 = drop_side == 1 ? (r * ld_src + coff_src + c) >> 3 : (r * ld_dst + coff_dst + c) >> 3;
      const uint32_t kp = segk_dropout_keep8(drop_mask, e8, drop_keep, drop_seed);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_bf16x2((&v.x)[j]);
        (&v.x)[j] = pack_bf16x2((kp >> (2 * j)) & 1u ? t.x * drop_inv_keep : 0.f, (kp >> (2 * j + 1)) & 1u ? t.y * drop_inv_keep : 0.f);
      }
    }
    uint4* d = reinterpret_cast<uint4*>(dst + r * ld_dst + coff_dst + c);
    float f[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = unpack_bf16x2((&v.x)[j]);
      f[2 * j] = t.x; f[2 * j + 1] = t.y;
    }
    if (accumulate) {
      const uint4 o = *d;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_bf16x2((&o.x)[j]);
        f[2 * j] += t.x; f[2 * j + 1] += t.y;
      }
    }
    if (mask) {
      const uint4 m = *reinterpret_cast<const uint4*>(mask + r * (int64_t)(C8 * 8) + c);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_bf16x2((&m.x)[j]);
        if (!(t.x > 0.f)) f[2 * j] = 0.f;
        if (!(t.y > 0.f)) f[2 * j + 1] = 0.f;
      }
    }
    *d = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
}

// ---------------------------------------------------------------------------------------
// Phase-packed transposed conv for tiny Cout (conv_t3: 16x16 s8, 256 -> 2, FCN.py:98-107).
// k = 2s, SAME: output block (r, c) = the s x s output pixels [r*s - p, (r+1)*s - p) x [c*s - p, (c+1)*s - p)
// depends only on the 2 x 2 input neighbourhood (r-1+dy, c-1+dx), dy,dx in {0,1}:
//   y[block (r,c)][(a,b,co)] = sum_{dy,dx,ci} x[r-1+dy, c-1+dx, ci] * W[a + s(1-dy), b + s(1-dx), co, ci]
// i.e. ONE stride-1 4-tap implicit GEMM over the (H+1) x (W+1) block grid with R = s*s*Cout columns, whose
// epilogue writes each row's s x s x Cout block of logits directly -- no patch-space tensor, no col2im.
// Backward: dyb[n, r, c, (a,b,co)] = dy[n, r*s-p+a, c*s-p+b, co] (0 outside) is a plain re-blocking of dy (no
// overlap), and dx / dW are a 4-tap igemm / wgrad over it with the same packed weights.
//   bf[t][ci/64][R rows][64]   rows (a,b,co), k = ci        (forward B operand)
//   bt[t][R/64][Cin rows][64]  rows ci,       k = (a,b,co)  (dgrad B operand)      t = dy*2 + dx
__global__ void __launch_bounds__(kThreads) pack_deconv_packed_kernel(const float* __restrict__ w, bf16* __restrict__ bf,
                                                                     bf16* __restrict__ bt, int k, int s, int Cin, int Cout) {
  const int R = s * s * Cout;
  const int64_t total = (int64_t)4 * Cin * R;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int row = (int)((i / Cin) % R);
    const int t = (int)(i / ((int64_t)Cin * R));
    const int dy = t >> 1, dx = t & 1;
    const int co = row % Cout, b = (row / Cout) % s, a = row / (Cout * s);
    const int ky = a + s * (1 - dy), kx = b + s * (1 - dx);
    const bf16 v = f2bf(w[(((int64_t)ky * k + kx) * Cout + co) * Cin + ci]);
    if (bf) bf[(((int64_t)t * (Cin / 64) + ci / 64) * R + row) * 64 + (ci & 63)] = v;
    if (bt) bt[(((int64_t)t * (R / 64) + row / 64) * Cin + ci) * 64 + (row & 63)] = v;
  }
}

// dyb[n, r, c, (a,b,co)] bf16 <- dy[n, r*s-p+a, c*s-p+b, co]; thread = 8 consecutive (b,co) of one a
template <typename GT>
__global__ void __launch_bounds__(kThreads) deconv_pack_dy_kernel(const GT* __restrict__ dy, uint4* __restrict__ dyb, int N,
                                                                  int H, int W, int Co, int s) {
  const int OH = H * s, OW = W * s, p = s / 2, R = s * s * Co, R8 = R >> 3, BH = H + 1, BW = W + 1;
  const int64_t total = (int64_t)N * BH * BW * R8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % R8);
    int64_t q = i / R8;
    const int c = (int)(q % BW);
    q /= BW;
    const int r = (int)(q % BH);
    const int n = (int)(q / BH);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = g * 8 + j;
      const int co = e % Co, b = (e / Co) % s, a = e / (Co * s);
      const int oy = r * s - p + a, ox = c * s - p + b;
      v[j] = (oy >= 0 && oy < OH && ox >= 0 && ox < OW) ? ld_f(dy + (((int64_t)n * OH + oy) * OW + ox) * Co + co) : 0.f;
    }
    dyb[i] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

// dW[ky,kx,co,ci] fp32 <- dwt[t][ci][(a,b,co)] with ky = a + s(1-dy), kx = b + s(1-dx), t = dy*2 + dx
__global__ void __launch_bounds__(kThreads) deconv_unpack_dw_kernel(const float* __restrict__ dwt, float* __restrict__ dw,
                                                                    int k, int s, int Cin, int Cout, int accumulate) {
  const int R = s * s * Cout;
  const int64_t total = (int64_t)k * k * Cout * Cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    int64_t q = i / Cin;
    const int co = (int)(q % Cout);
    q /= Cout;
    const int kx = (int)(q % k), ky = (int)(q / k);
    const int dy = 1 - ky / s, a = ky % s, dx = 1 - kx / s, b = kx % s;
    const float v = dwt[((int64_t)(dy * 2 + dx) * Cin + ci) * R + (a * s + b) * Cout + co];
    dw[i] = accumulate ? dw[i] + v : v;
  }
}

}  // namespace

extern "C" {

int segk_im2col_k64(segk_ctx* ctx, const void* x, int x_dtype, void* P, int N, int H, int W, int Cin, int kh,
                    int kw, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && P && N > 0 && H > 0 && W > 0, "im2col_k64: bad args");
  SEGK_REQUIRE(ctx, kh * kw * Cin <= 64 && (kh & 1) && (kw & 1), "im2col_k64: need kh*kw*Cin <= 64, odd kernel");
  const int64_t items = (int64_t)N * H * W * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == 2)
    im2col_k64_kernel<uint8_t><<<sgrid(ctx, items, 16), kThreads, 0, st>>>((const uint8_t*)x, (uint4*)P, N, H, W, Cin, kh, kw);
  else if (x_dtype == 0)
    im2col_k64_kernel<bf16><<<sgrid(ctx, items, 16), kThreads, 0, st>>>((const bf16*)x, (uint4*)P, N, H, W, Cin, kh, kw);
  else
    return segk_fail(ctx, SEGK_EINVAL, "im2col_k64: x_dtype must be 0 (bf16) or 2 (u8)");
  SEGK_LAUNCHED(ctx, "im2col_k64");
  return SEGK_OK;
}

int segk_pack_im2col_weights(segk_ctx* ctx, const float* w, void* wk, int K, int Cout, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, w && wk && K > 0 && K <= 64 && Cout > 0, "pack_im2col_weights: bad args");
  pack_im2col_weights_kernel<<<ceil_div(Cout * 64, kThreads), kThreads, 0, (cudaStream_t)stream>>>(w, (bf16*)wk, K, Cout);
  SEGK_LAUNCHED(ctx, "pack_im2col_weights");
  return SEGK_OK;
}

int segk_deconv_patch_gather(segk_ctx* ctx, const void* dy, int dy_is_f32, void* P, int N, int H, int W, int Cout,
                             int k, int s, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, dy && P && N > 0 && H > 0 && W > 0, "patch_gather: bad args");
  SEGK_REQUIRE(ctx, k == 2 * s && s % 2 == 0 && (k * k * Cout) % 8 == 0, "patch_gather: need k == 2*stride, k*k*Cout %% 8 == 0");
  const int64_t items = (int64_t)N * H * W * (k * k * Cout / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dy_is_f32)
    patch_gather_kernel<float><<<sgrid(ctx, items, 16), kThreads, 0, st>>>((const float*)dy, (uint4*)P, N, H, W, Cout, k, s);
  else
    patch_gather_kernel<bf16><<<sgrid(ctx, items, 16), kThreads, 0, st>>>((const bf16*)dy, (uint4*)P, N, H, W, Cout, k, s);
  SEGK_LAUNCHED(ctx, "patch_gather");
  return SEGK_OK;
}

int segk_deconv_col2im(segk_ctx* ctx, const float* yp, const float* bias, const void* residual, void* y, int out_f32,
                       int N, int H, int W, int Cout, int k, int s, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, yp && y && N > 0 && H > 0 && W > 0, "col2im: bad args");
  SEGK_REQUIRE(ctx, k == 2 * s && s % 2 == 0, "col2im: need k == 2*stride, even stride");
  const int64_t items = (int64_t)N * H * s * W * s * Cout;
  cudaStream_t st = (cudaStream_t)stream;
  if (out_f32)
    col2im_kernel<float><<<sgrid(ctx, items, 16), kThreads, 0, st>>>(yp, bias, (const bf16*)residual, (float*)y, N, H, W, Cout, k, s);
  else
    col2im_kernel<bf16><<<sgrid(ctx, items, 16), kThreads, 0, st>>>(yp, bias, (const bf16*)residual, (bf16*)y, N, H, W, Cout, k, s);
  SEGK_LAUNCHED(ctx, "col2im");
  return SEGK_OK;
}

int segk_scale_columns(segk_ctx* ctx, const float* in, const float* scale, float mult, float* out, int64_t rows,
                       int C, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, in && scale && out && rows > 0 && C > 0, "scale_columns: bad args");
  const int64_t n = rows * C;
  scale_columns_kernel<<<sgrid(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(in, scale, mult, out, n, C);
  SEGK_LAUNCHED(ctx, "scale_columns");
  return SEGK_OK;
}

int segk_bn_unfold_grads(segk_ctx* ctx, float* gw, const float* w, const float* gamma, float mult, float* dgamma,
                         void* workspace, size_t workspace_bytes, int64_t rows, int C, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, gw && w && gamma && dgamma && workspace && rows > 0 && C > 0, "bn_unfold_grads: bad args");
  const int cb = (C + 31) / 32;
  // enough blocks to fill the GPU, at least 64 rows each
  int S = (int)((2 * (int64_t)ctx->sm_count + cb - 1) / cb);
  const int64_t max_s = (rows + 63) / 64;
  if (S > max_s) S = (int)max_s;
  if (S < 1) S = 1;
  const int rpb = (int)((rows + S - 1) / S);
  S = (int)((rows + rpb - 1) / rpb);
  SEGK_REQUIRE(ctx, workspace_bytes >= sizeof(float) * (size_t)S * C, "bn_unfold_grads: workspace too small (%zu < %zu)",
               workspace_bytes, sizeof(float) * (size_t)S * C);
  bn_unfold_grads_kernel<<<dim3(cb, S), kThreads, 0, (cudaStream_t)stream>>>(gw, w, gamma, mult, (float*)workspace, rows, C, rpb);
  SEGK_LAUNCHED(ctx, "bn_unfold_grads");
  bn_unfold_finish_kernel<<<(C + kThreads - 1) / kThreads, kThreads, 0, (cudaStream_t)stream>>>((const float*)workspace, dgamma, mult,
                                                                                               S, C);
  SEGK_LAUNCHED(ctx, "bn_unfold_finish");
  return SEGK_OK;
}

int segk_bn_gamma_grad(segk_ctx* ctx, const void* dz, const void* y, const float* beta, const float* gamma,
                       float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes, int64_t rows, int C,
                       void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, dz && y && beta && gamma && dgamma && workspace && rows > 0, "bn_gamma_grad: bad args");
  SEGK_REQUIRE(ctx, C % 8 == 0 && (C / 8 <= kThreads ? kThreads % (C / 8) == 0 : (C / 8) % kThreads == 0),
               "bn_gamma_grad: C must be a multiple of 8 with C/8 dividing 256 (got %d)", C);
  const int C8 = C / 8;
  const int cpb = C8 < kThreads ? C8 : kThreads;
  const int R = kThreads / cpb, gy = C8 / cpb;
  int64_t gx = ceil_div64(rows, (int64_t)R * 4);
  int64_t cap = ceil_div64((int64_t)ctx->sm_count * 4, gy);
  const int64_t few = ceil_div64(rows, (int64_t)R * 16);      // >= 16 row steps per block: small layers get fewer partials
  if (cap > few) cap = few;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  const size_t need = sizeof(float) * (size_t)gx * C * (dbeta ? 2 : 1);
  SEGK_REQUIRE(ctx, workspace_bytes >= need, "bn_gamma_grad: workspace too small (%zu < %zu)", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  if (dbeta)
    bn_gamma_partial_kernel<true><<<dim3((unsigned)gx, gy), kThreads, 0, st>>>((const uint4*)dz, (const uint4*)y, beta,
                                                                            (float*)workspace, rows, C8);
  else
    bn_gamma_partial_kernel<false><<<dim3((unsigned)gx, gy), kThreads, 0, st>>>((const uint4*)dz, (const uint4*)y, beta,
                                                                             (float*)workspace, rows, C8);
  SEGK_LAUNCHED(ctx, "bn_gamma_partial");
  bn_gamma_finish_kernel<<<ceil_div(C, 32), kThreads, 0, st>>>((const float*)workspace, gamma, dgamma, dbeta, (int)gx, C);
  SEGK_LAUNCHED(ctx, "bn_gamma_finish");
  return SEGK_OK;
}

int segk_bn_grads_f32(segk_ctx* ctx, const float* dz, const float* y, const float* beta, const float* gamma, float* dgamma,
                      float* dbeta, void* workspace, size_t workspace_bytes, int64_t rows, int C, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, dz && y && beta && gamma && dgamma && dbeta && workspace && rows > 0, "bn_grads_f32: bad args");
  SEGK_REQUIRE(ctx, C == 2 || C == 4 || C == 8, "bn_grads_f32: C must be 2, 4 or 8 (got %d)", C);
  int64_t gx = ceil_div64(rows, (int64_t)kThreads * 8);
  if (gx > (int64_t)ctx->sm_count * 4) gx = (int64_t)ctx->sm_count * 4;
  if (gx < 1) gx = 1;
  const size_t need = sizeof(float) * (size_t)gx * 2 * C;
  SEGK_REQUIRE(ctx, workspace_bytes >= need, "bn_grads_f32: workspace too small (%zu < %zu)", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 2) bn_grads_f32_partial_kernel<2><<<(unsigned)gx, kThreads, 0, st>>>(dz, y, beta, (float*)workspace, rows);
  else if (C == 4) bn_grads_f32_partial_kernel<4><<<(unsigned)gx, kThreads, 0, st>>>(dz, y, beta, (float*)workspace, rows);
  else bn_grads_f32_partial_kernel<8><<<(unsigned)gx, kThreads, 0, st>>>(dz, y, beta, (float*)workspace, rows);
  SEGK_LAUNCHED(ctx, "bn_grads_f32_partial");
  bn_gamma_finish_kernel<<<ceil_div(C, 32), kThreads, 0, st>>>((const float*)workspace, gamma, dgamma, dbeta, (int)gx, C);
  SEGK_LAUNCHED(ctx, "bn_gamma_finish");
  return SEGK_OK;
}

int segk_channel_copy(segk_ctx* ctx, const void* src, int ld_src, int coff_src, void* dst, int ld_dst, int coff_dst,
                      const void* mask, int accumulate, int64_t rows, int C, int drop_side, const uint8_t* drop_mask,
                      float drop_keep, uint64_t drop_seed, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  if (!(drop_keep > 0.f && drop_keep < 1.f)) drop_side = 0;
  SEGK_REQUIRE(ctx, drop_side >= 0 && drop_side <= 2 && ((uintptr_t)drop_mask & 7) == 0, "channel_copy: bad dropout arguments");
  SEGK_REQUIRE(ctx, src && dst && rows > 0 && C > 0, "channel_copy: bad args");
  SEGK_REQUIRE(ctx, C % 8 == 0 && ld_src % 8 == 0 && ld_dst % 8 == 0 && coff_src % 8 == 0 && coff_dst % 8 == 0,
               "channel_copy: channel counts / offsets must be multiples of 8");
  const int64_t items = rows * (C / 8);
  channel_copy_kernel<<<sgrid(ctx, items, 16), kThreads, 0, (cudaStream_t)stream>>>(
      (const bf16*)src, ld_src, coff_src, (bf16*)dst, ld_dst, coff_dst, (const bf16*)mask, accumulate, rows, C / 8, drop_side,
      (const uint2*)drop_mask, drop_keep, drop_side ? 1.0f / drop_keep : 1.f, drop_seed);
  SEGK_LAUNCHED(ctx, "channel_copy");
  return SEGK_OK;
}


int segk_pack_deconv_packed(segk_ctx* ctx, const float* w, void* bf, void* bt, int k, int s, int Cin, int Cout, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, w && (bf || bt) && k == 2 * s && s > 0 && s % 2 == 0, "pack_deconv_packed: bad args");
  SEGK_REQUIRE(ctx, Cin % 64 == 0 && (s * s * Cout) % 64 == 0, "pack_deconv_packed: needs Cin %% 64 == 0 and s*s*Cout %% 64 == 0");
  pack_deconv_packed_kernel<<<sgrid(ctx, (int64_t)4 * Cin * s * s * Cout), kThreads, 0, (cudaStream_t)stream>>>(
      w, (bf16*)bf, (bf16*)bt, k, s, Cin, Cout);
  SEGK_LAUNCHED(ctx, "pack_deconv_packed");
  return SEGK_OK;
}

int segk_deconv_pack_dy(segk_ctx* ctx, const void* dy, int dy_is_f32, void* dyb, int N, int H, int W, int Cout, int s,
                        void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, dy && dyb && N > 0 && H > 0 && W > 0 && s > 0 && s % 2 == 0, "deconv_pack_dy: bad args");
  SEGK_REQUIRE(ctx, (s * s * Cout) % 8 == 0 && (((uintptr_t)dyb) & 15) == 0, "deconv_pack_dy: s*s*Cout %% 8, 16-byte alignment");
  const int64_t total = (int64_t)N * (H + 1) * (W + 1) * (s * s * Cout / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dy_is_f32)
    deconv_pack_dy_kernel<float><<<sgrid(ctx, total), kThreads, 0, st>>>((const float*)dy, (uint4*)dyb, N, H, W, Cout, s);
  else
    deconv_pack_dy_kernel<bf16><<<sgrid(ctx, total), kThreads, 0, st>>>((const bf16*)dy, (uint4*)dyb, N, H, W, Cout, s);
  SEGK_LAUNCHED(ctx, "deconv_pack_dy");
  return SEGK_OK;
}

int segk_deconv_unpack_dw(segk_ctx* ctx, const float* dwt, float* dw, int k, int s, int Cin, int Cout, int accumulate,
                          void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, dwt && dw && k == 2 * s && s > 0, "deconv_unpack_dw: bad args");
  deconv_unpack_dw_kernel<<<sgrid(ctx, (int64_t)k * k * Cout * Cin), kThreads, 0, (cudaStream_t)stream>>>(
      dwt, dw, k, s, Cin, Cout, accumulate);
  SEGK_LAUNCHED(ctx, "deconv_unpack_dw");
  return SEGK_OK;
}

}  // extern "C"
