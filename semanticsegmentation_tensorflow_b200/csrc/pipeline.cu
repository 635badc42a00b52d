// Input pipeline on the GPU: what get_batches_fn does per decoded image
// (FCN.py:176-201, 257-304) after PNG decode — crop / horizontal flip / scipy.misc.imresize (= PIL
// BILINEAR resize with its 22-bit fixed-point coefficients) / brightness-contrast / label colour match.
// Bit-exact with PIL: the coefficient tables are computed on the host exactly as PIL's
// precompute_coeffs + normalize_coeffs_8bpc do (semanticsegmentation_tensorflow_b200/pipeline.py), the kernels only
// evaluate  out = clip8((2^21 + sum_x in[x] * k[x]) >> 22)  per pass, horizontal then vertical.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kPrec = 22;

inline int pgrid(segk_ctx* ctx, int64_t items) {
  int64_t b = ceil_div64(items, kThreads), cap = (int64_t)ctx->sm_count * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

__device__ __forceinline__ int clip8(long long v) {
  v >>= kPrec;
  return v < 0 ? 0 : (v > 255 ? 255 : (int)v);
}

// C == 4 is PIL's RGBA (the reference's 4-channel "merge" PNGs, FCN.py:225,312): Image.resize converts to
// premultiplied alpha (RGBA -> RGBa: c' = MULDIV255(c, a)) before filtering and back afterwards
// (RGBa -> RGBA: c = min(255, 255 c' / a) unless a is 0 or 255) -- restated here bit for bit.
__device__ __forceinline__ int muldiv255(int a, int b) {
  const int t = a * b + 128;
  return ((t >> 8) + t) >> 8;
}

// horizontal pass over the crop [y0, y0+ch) x [x0, x0+cw) of src [H][W][C]: dst [ch][ow][C]
__global__ void __launch_bounds__(kThreads) resize_h_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                            const int* __restrict__ kk, const int* __restrict__ bounds,
                                                            int ksize, int W, int C, int x0, int y0, int cw, int ch,
                                                            int ow, int flip) {
  const int64_t total = (int64_t)ch * ow;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % ow), y = (int)(i / ow);
    const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
    const int* k = kk + (int64_t)xx * ksize;
    const uint8_t* row = src + ((int64_t)(y0 + y) * W) * C;
    for (int c = 0; c < C; ++c) {
      long long ss = 1ll << (kPrec - 1);
      for (int x = 0; x < n; ++x) {
        const int xi = xmin + x;                                   // position inside the (possibly flipped) crop
        const int xs = x0 + (flip ? (cw - 1 - xi) : xi);
        int v = row[(int64_t)xs * C + c];
        if (C == 4 && c < 3) v = muldiv255(v, row[(int64_t)xs * C + 3]);
        ss += (long long)v * k[x];
      }
      dst[((int64_t)y * ow + xx) * C + c] = (uint8_t)clip8(ss);
    }
  }
}

// vertical pass: tmp [ch][ow][C] -> out [oh][ow][C], with the per-image epilogue:
//   mode 0: plain;  mode 1: bc_img (FCN.py:186-192): trunc(clamp(v*s + m, 0, 255)) in double;
//   mode 2: process_gt_image (FCN.py:194-201): class id = 0 if pixel == (255,0,0) else 1 (u8 [oh][ow])
__global__ void __launch_bounds__(kThreads) resize_v_kernel(const uint8_t* __restrict__ tmp, uint8_t* __restrict__ out,
                                                            const int* __restrict__ kk, const int* __restrict__ bounds,
                                                            int ksize, int C, int ow, int oh, int mode, double s,
                                                            double m) {
  const int64_t total = (int64_t)oh * ow;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % ow), yy = (int)(i / ow);
    const int ymin = bounds[2 * yy], n = bounds[2 * yy + 1];
    const int* k = kk + (int64_t)yy * ksize;
    int v[4];
    for (int c = 0; c < C; ++c) {
      long long ss = 1ll << (kPrec - 1);
      for (int y = 0; y < n; ++y) ss += (long long)tmp[((int64_t)(ymin + y) * ow + xx) * C + c] * k[y];
      v[c] = clip8(ss);
    }
    if (C == 4 && v[3] != 0 && v[3] != 255) {
      for (int c = 0; c < 3; ++c) {
        const int u = (255 * v[c]) / v[3];
        v[c] = u > 255 ? 255 : u;
      }
    }
    if (mode == 2) {
      out[i] = (C >= 3 && v[0] == 255 && v[1] == 0 && v[2] == 0) ? 0 : 1;
    } else {
      for (int c = 0; c < C; ++c) {
        int o = v[c];
        if (mode == 1) {
          double d = (double)v[c] * s + m;
          d = d > 255.0 ? 255.0 : (d < 0.0 ? 0.0 : d);
          o = (int)d;                                              // astype(np.uint8): truncation
        }
        out[i * C + c] = (uint8_t)o;
      }
    }
  }
}

}  // namespace

extern "C" {

int segk_resize_h_u8(segk_ctx* ctx, const uint8_t* src, uint8_t* dst, const int* coeffs, const int* bounds, int ksize,
                     int W, int C, int x0, int y0, int crop_w, int crop_h, int out_w, int flip, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, src && dst && coeffs && bounds && ksize > 0 && crop_w > 0 && crop_h > 0 && out_w > 0,
               "resize_h: bad args");
  SEGK_REQUIRE(ctx, C == 1 || C == 3 || C == 4, "resize_h: C must be 1, 3 or 4 (4 = RGBA with PIL's premultiplied-alpha resize)");
  resize_h_kernel<<<pgrid(ctx, (int64_t)crop_h * out_w), kThreads, 0, (cudaStream_t)stream>>>(
      src, dst, coeffs, bounds, ksize, W, C, x0, y0, crop_w, crop_h, out_w, flip);
  SEGK_LAUNCHED(ctx, "resize_h");
  return SEGK_OK;
}

int segk_resize_v_u8(segk_ctx* ctx, const uint8_t* tmp, uint8_t* out, const int* coeffs, const int* bounds, int ksize,
                     int C, int out_w, int out_h, int mode, double contrast, double brightness, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, tmp && out && coeffs && bounds && ksize > 0 && out_w > 0 && out_h > 0, "resize_v: bad args");
  SEGK_REQUIRE(ctx, (C == 1 || C == 3 || C == 4) && mode >= 0 && mode <= 2, "resize_v: C must be 1, 3 or 4, mode 0..2");
  resize_v_kernel<<<pgrid(ctx, (int64_t)out_h * out_w), kThreads, 0, (cudaStream_t)stream>>>(
      tmp, out, coeffs, bounds, ksize, C, out_w, out_h, mode, contrast, brightness);
  SEGK_LAUNCHED(ctx, "resize_v");
  return SEGK_OK;
}

}  // extern "C"
