// First conv of the network (conv1_1, FCN.py:52: 3x3 SAME on the raw image, Cin = 3/4) on the tensor
// cores WITHOUT a patch tensor in HBM.  The im2col route (patch.cu) writes a [N,H,W,64] bf16 patch
// tensor (128 B/pixel) that the GEMM reads back: 2 x 377 MB per step at B=32 160x576 for 8.8 MB of
// image.  Here producer warps build each 128-pixel x K patch tile directly in shared memory, in the
// SWIZZLE_128B layout a TMA load would have produced, and tcgen05.mma consumes it from there:
//
//   first_fwd_kernel    y = relu(patch(x) . W + b)          reads the image, writes y once (TMA store)
//   first_wgrad_kernel  dW = patch(x)^T . dy  (+ db)        reads the image and dy once; one extra patch
//                                                           column of ones makes BiasAddGrad a row of dW
//
// Patch column order kk = (ky*3 + kx)*Cin + ci = the row order of the HWIO filter (FCN.py:125), so
// `segk_pack_im2col_weights` provides the forward operand and dW lands in the reference layout.
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace {

using namespace tc;

constexpr int kTileRows = 128;
constexpr int kABytes = kTileRows * 128;     // one patch tile: 128 pixels x 64 bf16, SWIZZLE_128B
constexpr int kFwdStages = 4;
constexpr int kFwdThreads = 288;             // warps 0-3 patch producers, 4 MMA, 5-8 epilogue
constexpr int kWgThreads = 160;              // warps 0-3 patch producers (+ final epilogue), 4 TMA + MMA
constexpr int kStageOutBytes = kTileRows * 128;
constexpr uint64_t kDescKMajor = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
constexpr uint64_t kDescMNBase = (64ull << 32) | (1ull << 46) | (2ull << 61);   // + LBO

struct FirstMaps {
  CUtensorMap w;    // forward: packed weights [1][Cout][64]
  CUtensorMap dy;   // wgrad: dy, box (64 ch, bw, bh, bn) = 128 pixels
  CUtensorMap y;    // forward: output, same box
};

struct FirstParams {
  const void* x;
  int N, H, W;
  int bw, bh, bn, rows;
  int tiles_w, tiles_h, tiles_n;
  const float* bias;
  int relu;
  float* ws;        // wgrad: per-CTA partial [grid][64][Cout]
  uint32_t* bits_out;   // forward: 1-bit ReLU mask of the output, [pixel][Cout/32] (see IgemmParams in tcconv.cu), or NULL
};

struct Pipe {
  int stage = 0;
  uint32_t phase = 0;
  template <int S>
  __device__ __forceinline__ void advance() {
    if (++stage == S) { stage = 0; phase ^= 1; }
  }
};

// raw element as loaded (u8 value, or the bf16 bit pattern) and its bf16 bit pattern
__device__ __forceinline__ uint32_t ld_raw(const uint8_t* p) { return (uint32_t)__ldg(p); }
__device__ __forceinline__ uint32_t ld_raw(const bf16* p) {
  return (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p));
}
// two u8 values -> packed bf16x2, exact, without the quarter-rate I2F: 2^23 + v has v in its low
// mantissa bits; subtracting 2^23 leaves float(v), whose upper 16 bits are its bf16 pattern
template <typename XT>
__device__ __forceinline__ uint32_t pack2(uint32_t lo, uint32_t hi) {
  if (sizeof(XT) == 1) {
    const float flo = __uint_as_float(0x4B000000u | lo) - 8388608.f;
    const float fhi = __uint_as_float(0x4B000000u | hi) - 8388608.f;
    return __byte_perm(__float_as_uint(flo), __float_as_uint(fhi), 0x7632);
  }
  return lo | (hi << 16);
}

template <int CIN>
struct Patch {
  static constexpr int K = 9 * CIN;
  static constexpr int kPieces = (K + 1 + 7) / 8;       // 16-byte pieces written per row (incl. the ones column)
  static constexpr int kSteps = (kPieces + 1) / 2;      // K = 16 MMA steps
};

// One thread owns the patch row of one pixel.  load_patch issues the (predicated) global loads of
// its 3x3xCin neighbourhood into registers; store_patch converts and writes row `row` of the tile:
// 16-byte piece pc at row*128 + ((pc ^ (row & 7)) << 4)  (the SWIZZLE_128B pattern).  Pieces >= kPieces
// stay zero from the one-time clear of the ring.  Split in two so the loads of the NEXT tile are in
// flight while the current one waits for its stage and is written (the producers are latency-bound
// otherwise: measured 2.4 TB/s of output with load-then-use).
template <typename XT, int CIN>
__device__ __forceinline__ bool load_patch(const FirstParams& p, int row, int tile, uint32_t (&raw)[9 * CIN]) {
  int x0 = (tile % p.tiles_w) * p.bw;
  tile /= p.tiles_w;
  int y0 = (tile % p.tiles_h) * p.bh;
  int n0 = (tile / p.tiles_h) * p.bn;
  const int iw = row % p.bw, ih = (row / p.bw) % p.bh, in = row / (p.bw * p.bh);
  const int px = x0 + iw, py = y0 + ih, n = n0 + in;
  const bool inside = row < p.rows && px < p.W && py < p.H && n < p.N;
#pragma unroll
  for (int i = 0; i < 9 * CIN; ++i) raw[i] = 0;
  if (inside) {
    const XT* xin = reinterpret_cast<const XT*>(p.x);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = py + ky - 1;
      if (yy >= 0 && yy < p.H) {
        const XT* r = xin + (((int64_t)n * p.H + yy) * p.W + px) * CIN;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = px + kx - 1;
          if (xx >= 0 && xx < p.W) {
#pragma unroll
            for (int c = 0; c < CIN; ++c) raw[(ky * 3 + kx) * CIN + c] = ld_raw(r + (kx - 1) * CIN + c);
          }
        }
      }
    }
  }
  return inside;
}

template <typename XT, int CIN>
__device__ __forceinline__ void store_patch(const uint32_t (&raw)[9 * CIN], bool inside, int row, uint8_t* tile,
                                            bool ones) {
  using PT = Patch<CIN>;
  uint32_t w[PT::kPieces * 4];
#pragma unroll
  for (int i = 0; i < PT::kPieces * 4; ++i) {
    const uint32_t lo = 2 * i < PT::K ? raw[2 * i < PT::K ? 2 * i : 0] : 0u;
    const uint32_t hi = 2 * i + 1 < PT::K ? raw[2 * i + 1 < PT::K ? 2 * i + 1 : 0] : 0u;
    w[i] = pack2<XT>(lo, hi);
  }
  if (ones && inside) {                      // bf16 1.0 in patch column K
    if (PT::K & 1) w[PT::K / 2] |= 0x3F800000u; else w[PT::K / 2] |= 0x3F80u;
  }
  uint8_t* dst = tile + row * 128;
#pragma unroll
  for (int pc = 0; pc < PT::kPieces; ++pc)
    *reinterpret_cast<uint4*>(dst + ((pc ^ (row & 7)) << 4)) = make_uint4(w[4 * pc], w[4 * pc + 1], w[4 * pc + 2], w[4 * pc + 3]);
}

__device__ __forceinline__ void decode_tile(const FirstParams& p, int tile, int& x0, int& y0, int& n0) {
  x0 = (tile % p.tiles_w) * p.bw;
  tile /= p.tiles_w;
  y0 = (tile % p.tiles_h) * p.bh;
  n0 = (tile / p.tiles_h) * p.bn;
}

template <int BLOCK_N>
struct FwdCfg {
  static constexpr int kBBytes = BLOCK_N * 128;
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kSmemBytes = kFwdStages * kABytes + kBBytes + 2 * kStageOutBytes + 1024 + 256;
};

template <typename XT, int CIN, int BLOCK_N>
__global__ void __launch_bounds__(kFwdThreads, BLOCK_N == 64 ? 2 : 1)
first_fwd_kernel(const __grid_constant__ FirstMaps maps, const FirstParams p) {
  using C = FwdCfg<BLOCK_N>;
  using PT = Patch<CIN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem + kFwdStages * kABytes;
  uint8_t* smem_out = smem_b + C::kBBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_out + 2 * kStageOutBytes);
  uint64_t* empty_bar = full_bar + kFwdStages;
  uint64_t* tfull_bar = empty_bar + kFwdStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* bfull_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfull_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.w);
    tma_prefetch_desc(&maps.y);
    for (int i = 0; i < kFwdStages; ++i) {
      mbar_init(&full_bar[i], 128);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    mbar_init(bfull_bar, 1);
    fence_barrier_init();
  }
  // patch pieces beyond the ones a row writes (K padding) must read as zero
  for (int i = threadIdx.x; i < kFwdStages * kABytes / 16; i += kFwdThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 4) tmem_alloc<C::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ---- patch producers: thread = one pixel row of the tile ----
    const int row = threadIdx.x;
    Pipe ps;
    uint32_t cur[9 * CIN], nxt[9 * CIN];
    bool cur_in = false, nxt_in = false;
    if ((int)blockIdx.x < total_tiles) cur_in = load_patch<XT, CIN>(p, row, blockIdx.x, cur);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int next = tile + gridDim.x;
      if (next < total_tiles) nxt_in = load_patch<XT, CIN>(p, row, next, nxt);   // in flight during the wait below
      mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
      store_patch<XT, CIN>(cur, cur_in, row, smem + ps.stage * kABytes, false);
      fence_proxy_async();                 // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(&full_bar[ps.stage]);
      ps.advance<kFwdStages>();
#pragma unroll
      for (int i = 0; i < 9 * CIN; ++i) cur[i] = nxt[i];
      cur_in = nxt_in;
    }
  } else if (warp == 4) {
    // ---- weights once (TMA), then one short MMA chain per tile ----
    if (elect_one()) {
      mbar_arrive_expect_tx(bfull_bar, (uint32_t)C::kBBytes);
      tma_load_3d(&maps.w, bfull_bar, smem_b, 0, 0, 0);
    }
    __syncwarp();
    mbar_wait(bfull_bar, 0);
    Pipe ps;
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr uint32_t idesc = make_idesc(kTileRows, BLOCK_N, 0, 0);
    const uint32_t a_lo0 = smem_u32(smem) >> 4, b_lo = smem_u32(smem_b) >> 4;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      mbar_wait(&full_bar[ps.stage], ps.phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = a_lo0 + (uint32_t)ps.stage * (uint32_t)(kABytes >> 4);
        const uint32_t d_addr = tmem_base + (uint32_t)(acc * BLOCK_N);
#pragma unroll
        for (int k = 0; k < PT::kSteps; ++k)
          umma_f16(d_addr, kDescKMajor | (uint64_t)(a_lo + 2 * k), kDescKMajor | (uint64_t)(b_lo + 2 * k), idesc,
                   k != 0 ? 1u : 0u);
        umma_commit(&empty_bar[ps.stage]);
        umma_commit(&tfull_bar[acc]);
      }
      __syncwarp();
      ps.advance<kFwdStages>();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ---- epilogue: warps 5..8, TMEM lane quarter = warp % 4 ----
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool ep_leader = (warp == 5 && lane == 0);
    uint32_t sg = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int x0, y0, n0;
      decode_tile(p, tile, x0, y0, n0);
      uint32_t* bits_row = nullptr;      // this thread's row of the 1-bit ReLU mask, when it is inside the tensor
      if (p.bits_out) {
        const int px = x0 + row % p.bw, py = y0 + (row / p.bw) % p.bh, pn = n0 + row / (p.bw * p.bh);
        if (row < p.rows && px < p.W && py < p.H && pn < p.N)
          bits_row = p.bits_out + (((int64_t)pn * p.H + py) * p.W + px) * (BLOCK_N / 32);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint8_t* sbuf = smem_out + (sg & 1) * kStageOutBytes;
        if ((c0 & 32) == 0) {
          if (ep_leader) tma_store_wait_read<1>();   // the store that last used this buffer has read it
          named_bar_sync(1, 128);
        }
        float4 bv[8];
        if (p.bias) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + c0);
#pragma unroll
          for (int i = 0; i < 8; ++i) bv[i] = __ldg(b4 + i);
        }
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[4 * i] += bv[i].x; v[4 * i + 1] += bv[i].y; v[4 * i + 2] += bv[i].z; v[4 * i + 3] += bv[i].w;
          }
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        const int pbase = (c0 & 32) ? 4 : 0;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(sbuf + row * 128 + (((pbase + i) ^ (row & 7)) << 4)) =
              make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        if (bits_row) {
          uint32_t word = 0;
#pragma unroll
          for (int i = 0; i < 16; ++i) word |= bf16x2_pos_bits(pk[i]) << (2 * i);
          bits_row[c0 >> 5] = word;
        }
        if (c0 & 32) {
          fence_proxy_async();
          named_bar_sync(1, 128);
          if (ep_leader) {
            tma_store_4d(&maps.y, sbuf, c0 - 32, x0, y0, n0);   // clipped at the tensor edge
            tma_store_commit();
          }
          ++sg;
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (ep_leader) tma_store_wait_read<0>();
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4) tmem_dealloc<C::kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------
// wgrad: D[kk][co] += sum over the CTA's pixels  patch[pixel][kk] * dy[pixel][co].  Both operands are
// MN-major (rows = pixels = the GEMM K).  M = 128: rows 0..63 are the patch columns, rows 64..127 read
// an all-zero block (LBO points at it) and are discarded.  Per-CTA partials go to the workspace and
// are summed by first_wgrad_reduce_kernel (no atomics: 2 x SM-count CTAs on 1.8 k addresses serialise).
// ------------------------------------------------------------------------------------------
template <int BLOCK_N>
struct WgCfg {
  static constexpr int kBBytes = (BLOCK_N / 64) * kABytes;      // 64-channel blocks of [128 pixels][64 ch]
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BLOCK_N == 64 ? 3 : (BLOCK_N == 128 ? 4 : 2);
  static constexpr int kZeroBytes = 2048;                        // 16 pixel rows of zeros (one MMA's K extent)
  static constexpr int kSmemBytes = kStages * kStageBytes + kZeroBytes + 1024 + 256;
};

template <typename XT, int CIN, int BLOCK_N>
__global__ void __launch_bounds__(kWgThreads, BLOCK_N == 64 ? 2 : 1)
first_wgrad_kernel(const __grid_constant__ FirstMaps maps, const FirstParams p) {
  using C = WgCfg<BLOCK_N>;
  using PT = Patch<CIN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_zero = smem + C::kStages * C::kStageBytes;
  uint64_t* afull_bar = reinterpret_cast<uint64_t*>(smem_zero + C::kZeroBytes);
  uint64_t* bfull_bar = afull_bar + C::kStages;
  uint64_t* empty_bar = bfull_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.dy);
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&afull_bar[i], 128);
      mbar_init(&bfull_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < (C::kStages * C::kStageBytes + C::kZeroBytes) / 16; i += kWgThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 4) tmem_alloc<BLOCK_N>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // (the host launches at most one CTA per tile, so every CTA owns at least one tile)

  if (warp < 4) {
    const int row = threadIdx.x;
    Pipe ps;
    uint32_t cur[9 * CIN], nxt[9 * CIN];
    bool cur_in = false, nxt_in = false;
    cur_in = load_patch<XT, CIN>(p, row, blockIdx.x, cur);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int next = tile + gridDim.x;
      if (next < total_tiles) nxt_in = load_patch<XT, CIN>(p, row, next, nxt);
      mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
      store_patch<XT, CIN>(cur, cur_in, row, smem + ps.stage * C::kStageBytes, true);
      fence_proxy_async();
      mbar_arrive(&afull_bar[ps.stage]);
      ps.advance<C::kStages>();
#pragma unroll
      for (int i = 0; i < 9 * CIN; ++i) cur[i] = nxt[i];
      cur_in = nxt_in;
    }
    // ---- epilogue: the CTA's partial sums, patch columns 0..K (K = the ones column = bias gradient) ----
    const int q = warp;                 // TMEM lane quarter
    float* dst = p.ws + ((int64_t)blockIdx.x * 64 + row) * BLOCK_N;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    if (q < 2) {                        // patch columns live in TMEM lanes 0..63 (warp-uniform branch)
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        if (row <= PT::K) {
          float4* d4 = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            d4[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                __uint_as_float(r[4 * i + 3]));
        }
      }
    }
  } else {
    // ---- warp 4: TMA of the dy tiles + MMA issue ----
    Pipe pl, pm;
    constexpr uint32_t idesc = make_idesc(kTileRows, BLOCK_N, 1, 1);
    const uint32_t smem_lo = smem_u32(smem) >> 4;
    const uint32_t zero_lo = smem_u32(smem_zero) >> 4;
    const uint64_t bdesc_hi = kDescMNBase | ((uint64_t)(kABytes >> 4) << 16);      // 64-channel blocks 16 KB apart
    auto load_dy = [&](int t) {
      mbar_wait(&empty_bar[pl.stage], pl.phase ^ 1);
      if (elect_one()) {
        int x0, y0, n0;
        decode_tile(p, t, x0, y0, n0);
        uint8_t* sb = smem + pl.stage * C::kStageBytes + kABytes;
        mbar_arrive_expect_tx(&bfull_bar[pl.stage], (uint32_t)C::kBBytes);
#pragma unroll
        for (int j = 0; j < BLOCK_N / 64; ++j)
          tma_load_4d(&maps.dy, &bfull_bar[pl.stage], sb + j * kABytes, j * 64, x0, y0, n0);
      }
      __syncwarp();
      pl.advance<C::kStages>();
    };
    // prologue: fill the ring; afterwards the stage of tile j-1 is refilled right after the MMAs of
    // tile j have been issued (so the wait for tile j-1's completion overlaps tile j's MMAs)
    int load_tile = blockIdx.x;
    for (int i = 0; i < C::kStages && load_tile < total_tiles; ++i, load_tile += gridDim.x) load_dy(load_tile);
    bool first = true;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&afull_bar[pm.stage], pm.phase);
      mbar_wait(&bfull_bar[pm.stage], pm.phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = smem_lo + (uint32_t)pm.stage * (uint32_t)(C::kStageBytes >> 4);
        const uint32_t b_lo = a_lo + (uint32_t)(kABytes >> 4);
#pragma unroll
        for (int k = 0; k < kTileRows / 16; ++k) {   // 16 pixels = 2048 B = 128 descriptor units per MMA
          // rows 64..127 of the A operand: the zero block (LBO = its distance from this step's start)
          const uint64_t adesc = kDescMNBase | ((uint64_t)((zero_lo - (a_lo + 128 * k)) & 0x3FFFu) << 16) |
                                 (uint64_t)(a_lo + 128 * k);
          umma_f16(tmem_base, adesc, bdesc_hi | (uint64_t)(b_lo + 128 * k), idesc, (!first || k != 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[pm.stage]);
        if (tile + (int)gridDim.x >= total_tiles) umma_commit(tfull_bar);
      }
      __syncwarp();
      pm.advance<C::kStages>();
      if (!first && load_tile < total_tiles) {
        load_dy(load_tile);
        load_tile += gridDim.x;
      }
      first = false;
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4) tmem_dealloc<BLOCK_N>(tmem_base);
}

// dw[kk][co] = sum over CTAs ws[cta][kk][co] (kk < K);  db[co] = the same for the ones row kk = K.
// block = 32 columns x 8 CTA lanes
__global__ void __launch_bounds__(256) first_wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw,
                                                                 float* __restrict__ db, int nparts, int K, int Cout) {
  __shared__ float sh[8][33];
  const int cx = threadIdx.x & 31, lane8 = threadIdx.x >> 5;
  const int cols_per_row = Cout / 32;
  const int kk = blockIdx.x / cols_per_row;
  const int c = (blockIdx.x % cols_per_row) * 32 + cx;
  float s = 0.f;
  for (int part = lane8; part < nparts; part += 8) s += ws[((int64_t)part * 64 + kk) * Cout + c];
  sh[lane8][cx] = s;
  __syncthreads();
  if (lane8 == 0) {
#pragma unroll
    for (int i = 1; i < 8; ++i) s += sh[i][cx];
    if (kk < K)
      dw[(int64_t)kk * Cout + c] = s;
    else if (db)
      db[c] = s;
  }
}

template <typename XT, int CIN>
int launch_fwd(segk_ctx* ctx, int block_n, const FirstMaps& maps, const FirstParams& p, int tiles, cudaStream_t st) {
  const int per_sm = block_n == 64 ? 2 : 1;
  const int grid = tiles < ctx->sm_count * per_sm ? tiles : ctx->sm_count * per_sm;
  switch (block_n) {
    case 64: first_fwd_kernel<XT, CIN, 64><<<grid, kFwdThreads, FwdCfg<64>::kSmemBytes, st>>>(maps, p); break;
    case 128: first_fwd_kernel<XT, CIN, 128><<<grid, kFwdThreads, FwdCfg<128>::kSmemBytes, st>>>(maps, p); break;
    default: first_fwd_kernel<XT, CIN, 256><<<grid, kFwdThreads, FwdCfg<256>::kSmemBytes, st>>>(maps, p); break;
  }
  SEGK_LAUNCHED(ctx, "first conv fwd");
  return SEGK_OK;
}

template <typename XT, int CIN>
int launch_wgrad(segk_ctx* ctx, int block_n, const FirstMaps& maps, const FirstParams& p, int grid, cudaStream_t st) {
  switch (block_n) {
    case 64: first_wgrad_kernel<XT, CIN, 64><<<grid, kWgThreads, WgCfg<64>::kSmemBytes, st>>>(maps, p); break;
    case 128: first_wgrad_kernel<XT, CIN, 128><<<grid, kWgThreads, WgCfg<128>::kSmemBytes, st>>>(maps, p); break;
    default: first_wgrad_kernel<XT, CIN, 256><<<grid, kWgThreads, WgCfg<256>::kSmemBytes, st>>>(maps, p); break;
  }
  SEGK_LAUNCHED(ctx, "first conv wgrad");
  return SEGK_OK;
}

int check_args(segk_ctx* ctx, const char* what, int x_dtype, int N, int H, int W, int Cin, int Cout, int kh, int kw) {
  SEGK_REQUIRE(ctx, N > 0 && H > 0 && W > 0, "%s: empty tensor", what);
  SEGK_REQUIRE(ctx, kh == 3 && kw == 3 && (Cin == 1 || Cin == 3 || Cin == 4),
               "%s: fused first-layer path is 3x3 with Cin in {1,3,4} (got %dx%d, Cin %d); use the im2col route", what, kh,
               kw, Cin);
  SEGK_REQUIRE(ctx, Cout == 64 || Cout == 128 || Cout == 256, "%s: Cout must be 64, 128 or 256 (got %d)", what, Cout);
  SEGK_REQUIRE(ctx, x_dtype == 0 || x_dtype == 2, "%s: x_dtype must be 0 (bf16) or 2 (u8)", what);
  return SEGK_OK;
}

#define FIRST_DISPATCH(FN, ...)                                                        \
  do {                                                                                 \
    if (x_dtype == 2) {                                                                \
      if (Cin == 1) return FN<uint8_t, 1>(__VA_ARGS__);                                \
      if (Cin == 3) return FN<uint8_t, 3>(__VA_ARGS__);                                \
      return FN<uint8_t, 4>(__VA_ARGS__);                                              \
    }                                                                                  \
    if (Cin == 1) return FN<bf16, 1>(__VA_ARGS__);                                     \
    if (Cin == 3) return FN<bf16, 3>(__VA_ARGS__);                                     \
    return FN<bf16, 4>(__VA_ARGS__);                                                   \
  } while (0)

}  // namespace

extern "C" {

int segk_conv2d_first_fwd(segk_ctx* ctx, const void* x, int x_dtype, const void* wk, const float* bias, void* y,
                          uint32_t* relu_bits, int N, int H, int W, int Cin, int Cout, int kh, int kw, unsigned flags,
                          void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && wk && y, "conv2d_first_fwd: null pointer");
  int rc = check_args(ctx, "conv2d_first_fwd", x_dtype, N, H, W, Cin, Cout, kh, kw);
  if (rc) return rc;
  SEGK_REQUIRE(ctx, !(flags & SEGK_EPI_OUT_F32), "conv2d_first_fwd: output is bf16");
  SEGK_REQUIRE(ctx, (((uintptr_t)wk | (uintptr_t)y) & 15) == 0, "conv2d_first_fwd: 16-byte alignment");
  const Box b = tch::pick_box(N, H, W, kTileRows, false);
  SEGK_REQUIRE(ctx, b.rows > 0, "conv2d_first_fwd: no pixel box");
  FirstMaps maps;
  memset(&maps, 0, sizeof(maps));
  rc = tch::weight_map(ctx, &maps.w, wk, 64, Cout, 1, Cout);
  if (rc) return rc;
  rc = tch::act_map(ctx, &maps.y, y, N, H, W, Cout, b.bw, b.bh, b.bn);
  if (rc) return rc;
  FirstParams p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.N = N; p.H = H; p.W = W;
  p.bw = b.bw; p.bh = b.bh; p.bn = b.bn; p.rows = b.rows;
  p.tiles_w = ceil_div(W, b.bw); p.tiles_h = ceil_div(H, b.bh); p.tiles_n = ceil_div(N, b.bn);
  p.bias = bias; p.relu = (flags & SEGK_EPI_RELU) ? 1 : 0;
  p.bits_out = relu_bits;
  const int tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  FIRST_DISPATCH(launch_fwd, ctx, Cout, maps, p, tiles, (cudaStream_t)stream);
}

int segk_conv2d_first_wgrad(segk_ctx* ctx, const void* x, int x_dtype, const void* dy, float* dw, float* dbias, int N,
                            int H, int W, int Cin, int Cout, int kh, int kw, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && dy && dw, "conv2d_first_wgrad: null pointer");
  int rc = check_args(ctx, "conv2d_first_wgrad", x_dtype, N, H, W, Cin, Cout, kh, kw);
  if (rc) return rc;
  SEGK_REQUIRE(ctx, (((uintptr_t)dy | (uintptr_t)dw) & 15) == 0, "conv2d_first_wgrad: 16-byte alignment");
  const Box b = tch::pick_box(N, H, W, kTileRows, true);   // exact: TMA zero-fills what lies outside
  SEGK_REQUIRE(ctx, b.rows == kTileRows, "conv2d_first_wgrad: cannot tile %dx%dx%d into 128-pixel boxes", N, H, W);
  FirstMaps maps;
  memset(&maps, 0, sizeof(maps));
  rc = tch::act_map(ctx, &maps.dy, dy, N, H, W, Cout, b.bw, b.bh, b.bn);
  if (rc) return rc;
  FirstParams p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.N = N; p.H = H; p.W = W;
  p.bw = b.bw; p.bh = b.bh; p.bn = b.bn; p.rows = b.rows;
  p.tiles_w = ceil_div(W, b.bw); p.tiles_h = ceil_div(H, b.bh); p.tiles_n = ceil_div(N, b.bn);
  const int tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int per_sm = Cout == 64 ? 2 : 1;
  const int grid = tiles < ctx->sm_count * per_sm ? tiles : ctx->sm_count * per_sm;
  rc = tch::workspace(ctx, sizeof(float) * (size_t)grid * 64 * Cout);
  if (rc) return rc;
  p.ws = (float*)ctx->ws;
  cudaStream_t st = (cudaStream_t)stream;
  auto run = [&]() -> int { FIRST_DISPATCH(launch_wgrad, ctx, Cout, maps, p, grid, st); };
  rc = run();
  if (rc) return rc;
  const int K = 9 * Cin;
  first_wgrad_reduce_kernel<<<(K + 1) * (Cout / 32), 256, 0, st>>>(p.ws, dw, dbias, grid, K, Cout);
  SEGK_LAUNCHED(ctx, "first conv wgrad reduce");
  return SEGK_OK;
}

}  // extern "C"

int segk_first_init(segk_ctx* ctx) {
  cudaError_t e = cudaSuccess;
#define FIRST_ATTR(XT, CIN, BN)                                                                                    \
  if (e == cudaSuccess)                                                                                             \
    e = cudaFuncSetAttribute(first_fwd_kernel<XT, CIN, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                             FwdCfg<BN>::kSmemBytes);                                                               \
  if (e == cudaSuccess)                                                                                             \
    e = cudaFuncSetAttribute(first_wgrad_kernel<XT, CIN, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                             WgCfg<BN>::kSmemBytes)
#define FIRST_ATTR_ALL(XT, CIN) FIRST_ATTR(XT, CIN, 64); FIRST_ATTR(XT, CIN, 128); FIRST_ATTR(XT, CIN, 256)
  FIRST_ATTR_ALL(uint8_t, 1); FIRST_ATTR_ALL(uint8_t, 3); FIRST_ATTR_ALL(uint8_t, 4);
  FIRST_ATTR_ALL(bf16, 1); FIRST_ATTR_ALL(bf16, 3); FIRST_ATTR_ALL(bf16, 4);
#undef FIRST_ATTR_ALL
#undef FIRST_ATTR
  if (e != cudaSuccess) return segk_fail(ctx, SEGK_ECUDA, "first-layer kernel setup: %s", cudaGetErrorString(e));
  return SEGK_OK;
}
