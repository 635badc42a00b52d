// Tensor-core convolution family for sm_100a: implicit GEMM on tcgen05.mma with TMEM
// accumulators and TMA-staged operand tiles.
//
//   igemm_kernel  (K-major operands)  : conv_layer forward, its Conv2DBackpropInput (dgrad), the
//                                       phase-decomposed transposed conv and the strided conv that is
//                                       the transposed conv's input gradient.
//   wgrad_kernel  (MN-major operands) : Conv2DBackpropFilter, reduction over pixels, split-K.
//
// GEMM view:  rows = 128 output pixels (a TMA box bw x bh x bn of the NHWC
// tensor), cols = BLOCK_N output channels, K = taps x Cin walked in 64-channel steps.  SAME
// padding is the TMA unit's out-of-bounds zero fill: a tap (dy,dx) simply shifts the box
// start coordinate, which may be negative.  Nothing is im2col'ed in memory.
//
// Warp roles (320 threads, 1 CTA / SM, persistent over tiles):
//   warp 0    TMA producer (one elected lane)        smem ring of kStages {A 16 KB, B BLOCK_N*128 B}
//   warp 1    tcgen05.mma issuer (one lane) + TMEM allocator; 2 accumulators of BLOCK_N columns
//   warps 2-9 epilogue: tcgen05.ld -> bias / residual / ReLU / ReLU-mask / scale -> staging -> TMA store.
//             Two warps per TMEM lane quarter, one per 32-column half of each 64-column group: with
//             one warp per quarter the epilogue of a 64-channel tile (~2300 cycles of dependent issue,
//             ncu r1d) took longer than its MMAs (1152 tensor cycles) and set the pace of conv1_2.
#include "tc_common.cuh"
#include "tc_host.cuh"

#include <algorithm>
#include <cstdlib>
#include <utility>
#include <vector>

namespace {

using namespace tc;

constexpr int kThreads = 320;            // igemm / slab kernels: warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int kWgradThreads = 320;       // wgrad: warps 2-9 epilogue, two per TMEM lane quarter (one 32-column half of every 64-column group each):
                                         // a 128 x 256 fp32 tile is 128 KB of row-strided stores per item, which four warps could not hide under conv6's 10.6 us items
constexpr int kEpiThreads = 256;
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;             // bf16 elements = one 128-byte swizzle row
constexpr int kABytes = kBlockM * 128;  // 16 KB
constexpr int kMaxTaps = 64;
constexpr int kMaxRbp = 512;            // row-block pairs a wgrad launch can reorder by cost
constexpr int kMaxCols = 64;            // pixel tiles per channel tile of the lockstep tap-split schedule
constexpr int kSmemBudget = 192 * 1024;
constexpr int kStageOutBytes = kBlockM * 128;     // epilogue staging: 128 rows x 64 bf16 channels, SWIZZLE_128B
constexpr int kOutBytes = 2 * kStageOutBytes;     // double-buffered

template <int BLOCK_N>
struct Cfg {
  static constexpr int kBBytes = BLOCK_N * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = kSmemBudget / kStageBytes;  // 256:4  128:6  64:8
  static constexpr int kTmemCols = 2 * BLOCK_N;              // double-buffered accumulator
  static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// wgrad ring: a stage holds kPB 64-pixel boxes, so one barrier round-trip feeds 4*kPB MMAs
template <int BLOCK_N>
struct WCfg {
  static constexpr int kPB = BLOCK_N == 256 ? 1 : 2;
  static constexpr int kBoxBytes = kABytes + BLOCK_N * 128;      // one pixel box: A 2x8 KB + B BLOCK_N/64 x 8 KB
  static constexpr int kStageBytes = kPB * kBoxBytes;
  static constexpr int kStages = kSmemBudget / kStageBytes;      // 256:4  128:3  64:4
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

// UMMA shared-memory descriptor minus the start address (tc_common.cuh make_smem_desc): SWIZZLE_128B,
// version 1, SBO = 1024 B; K-major: LBO field 1 (ignored); MN-major: LBO = 8192 B between 64-wide blocks.
constexpr uint64_t kDescKMajor = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
constexpr uint64_t kDescMNMajor = (512ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);

struct alignas(64) TensorMaps {
  CUtensorMap a[4];
  CUtensorMap b;
  CUtensorMap c;   // output tensor (TMA-store epilogue), box (64 ch, bw, bh, bn)
  CUtensorMap b2;  // CTA-pair mode: the weights with a box of BLOCK_N / 2 rows (each CTA stages half of the B tile)
  CUtensorMap cph[4];  // stride-2 transposed conv: the output pixels of phase (ay, ax) as a decimated view (TMA-store epilogue)
};

struct TapTable {
  int8_t dy[kMaxTaps];
  int8_t dx[kMaxTaps];
  int8_t map[kMaxTaps];
};

struct IgemmParams {
  // row space: boxes of the q-grid [N, H, W] (for a plain conv the output pixel grid)
  int N, H, W;
  int bw, bh, bn, rows;
  int tiles_w, tiles_h, tiles_n;
  int n_tiles;   // GEMM-N tiles = Cout / BLOCK_N
  int phases;    // s*s output phases of a transposed conv (1 otherwise)
  int s;         // phases per dim
  int ntaps, kchunks;
  int in_H, in_W;  // extent of the tensor view the taps index (tile-level tap skipping)
  // output mapping: out pixel = (qy*os + ay - opad, qx*os + ax - opad), phase = ay*s + ax
  int out_H, out_W, ldo, os, opad;
  void* out;
  int out_f32;
  const float* bias;
  const bf16* residual;
  const bf16* mask;
  float scale;
  int relu;
  // split-K over the (tap, k-chunk) walk: unit = tile * ksplits + split; split s stores its partial
  // sums (plain 16-byte stores, no atomics: measured ~20 us per million scattered fp32 REDs) into slice s
  // of the fp32 workspace `ws` [ksplits][pixels][ldo]; epilogue_finish_kernel adds the slices up
  int ksplits;
  float* ws;
  int64_t ws_slice;   // elements per slice
  int tma_store;   // 1: bf16 output leaves through shared memory + TMA store (coalesced, async, clipped)
  int m_fastest;   // tile order: 0 = channel tile fastest (activations shared), 1 = pixel tile fastest (weights shared)
  // BiasAddGrad of the PRODUCER layer from the dgrad epilogue: column sums of the output (its pre-activation gradient),
  // per-(CTA, TMEM lane quarter) partial rows [n_tiles][(grid / n_tiles) * 4][BLOCK_N]; needs channel-tile-fastest
  // order, no split-K and grid % n_tiles == 0 (every tile of a CTA then has the same channel tile)
  float* colsum;
  // phase-packed transposed conv (segk_deconv2d_packed_fwd): GEMM column = (a*pack_s + b)*pack_co + co of the
  // pack_s x pack_s x pack_co output block of row (n, qy, qx); fp32 stores at out pixel (qy*pack_s - opad + a, ...)
  int pack_s, pack_co;
  // Lockstep tap-split schedule for layers with few output tiles and a long, weight-heavy K walk (conv6 dgrad:
  // 50 tiles x ~2060 k-steps over 205 MB of weights, which do not fit in L2).  Plain split-K leaves SMs idle
  // (50 x 2 units on 148 SMs) and lets CTAs whose tiles skip different taps walk the weights at different
  // positions, so every (y-tile, channel tile) streams its own copy from DRAM (ncu r2: 720 MB for 205 MB).
  // Here the work of a channel tile is the list of (pixel tile, active tap) pairs -- ts_total of them, exclusive
  // prefix sums per pixel tile in ts_prefix -- cut into ts_G equal ranges, one per CTA (ts_G * n_tiles CTAs, all
  // resident).  A CTA owns at most two pixel tiles (one TMEM accumulator each) and walks  k-chunk (outer) ->
  // its (tile, tap) pairs (inner): every CTA is at the same k-chunk at the same time, so a weight slice
  // (tap, k-chunk) comes from DRAM once and from L2 for everyone else, and every SM gets the same number of
  // k-steps (+-1 tap).  A CTA's piece of a tile goes to partial slice (CTA - first CTA of that tile);
  // epilogue_finish_ts_kernel adds a tile's slices in order (deterministic, no atomics).
  int ts, ts_G, ts_tiles, ts_total;
  int ts_prefix[kMaxCols + 1];
  int hyb, hyb_full; // hyb = 1: hybrid schedule -- tiles [0, hyb_full) whole, the rest in ksplits pieces
  // max_pool 2x2/2 of the output fused into the epilogue (conv -> ReLU -> pool, FCN.py:56,60,65,71,76): pooled values
  // [N,H/2,W/2,ldo] bf16 and first-max indices (u8) computed from the staged bf16 tile; pool_only = 1 skips the store of
  // the full-resolution tensor (nothing in FCN-8s training reads it again)
  bf16* pool_out;
  uint8_t* pool_idx;
  int pool_only;
  // 1-bit ReLU masks: bits_out[pixel][C/32] (bit i of word w = the stored bf16 output of channel 32 w + i is > 0) written
  // by a forward epilogue; mask_bits = the producer layer's words, read by its consumer's dgrad epilogue instead of the
  // bf16 activation (1/16 of the bytes, one coalesced 4-byte load per thread and column group)
  uint32_t* bits_out;
  const uint32_t* mask_bits;
  // narrow fp32 output (class-count heads: a 3x3 conv to 2 classes runs as a 64-column tile whose padded columns are
  // zero weights): only the first out_cols columns are stored, as fp32 [pixels][out_cols]
  int out_cols;
  // CTA-pair mode (igemm_pair_kernel, cta_group::2): CTAs 2q and 2q+1 take two consecutive pixel tiles of the same channel
  // tile and run them as ONE M = 256 MMA -- each stages its own A box and half of the weight tile (plain schedule only)
  int pair;
};

struct PipeState {
  int stage = 0;
  uint32_t phase = 0;
  template <int STAGES>
  __device__ __forceinline__ void advance() {
    if (++stage == STAGES) {
      stage = 0;
      phase ^= 1;
    }
  }
};

struct TileCoord {
  int nt, x0, y0, n0, phase, split;
};

// Transposed conv (phases > 1): out pixel = q * s + a - opad.  A phase with a < opad produces nothing at q = 0 and its last
// pixel at q = H (W), one with a >= opad covers q = 0 .. H - 1: every phase has exactly H x W rows, starting at q = 1 or 0.
// The tiles are laid over that H x W range and shifted here (no (H+1) x (W+1) grid with an empty border row per phase).
__device__ __forceinline__ void phase_origin(const IgemmParams& p, TileCoord& t) {
  if (p.phases > 1) {
    t.x0 += (t.phase % p.s) < p.opad ? 1 : 0;
    t.y0 += (t.phase / p.s) < p.opad ? 1 : 0;
  }
}

__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int tile) {
  TileCoord t;
  t.split = 0;
  if (p.m_fastest) {
    // weight-heavy layers (conv6: 205 MB of weights, 3 MB of activations): consecutive CTAs take
    // different pixel tiles of the SAME channel tile, so a weight tile is fetched from DRAM once and
    // shared through L2 by all the CTAs that need it at that moment
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    int r = tile % m_tiles;
    const int q = tile / m_tiles;
    t.nt = q % p.n_tiles;
    t.phase = q / p.n_tiles;
    t.x0 = (r % p.tiles_w) * p.bw;
    r /= p.tiles_w;
    t.y0 = (r % p.tiles_h) * p.bh;
    t.n0 = (r / p.tiles_h) * p.bn;
    phase_origin(p, t);
    return t;
  }
  t.nt = tile % p.n_tiles;
  int r = tile / p.n_tiles;
  t.x0 = (r % p.tiles_w) * p.bw;
  r /= p.tiles_w;
  t.y0 = (r % p.tiles_h) * p.bh;
  r /= p.tiles_h;
  t.n0 = (r % p.tiles_n) * p.bn;
  t.phase = r / p.tiles_n;
  phase_origin(p, t);
  return t;
}

// pixel tile r (x fastest, then y, then n -- the order of decode_tile) of channel tile nt; r may be one past the last tile
// (the odd tile's partner in CTA-pair mode): its box lies outside the batch, loads zero-fill and nothing is stored
__device__ __forceinline__ TileCoord decode_mtile(const IgemmParams& p, int r, int nt) {
  TileCoord t;
  t.split = 0; t.phase = 0; t.nt = nt;
  t.x0 = (r % p.tiles_w) * p.bw;
  r /= p.tiles_w;
  t.y0 = (r % p.tiles_h) * p.bh;
  t.n0 = (r / p.tiles_h) * p.bn;
  return t;
}

// A tap whose shifted box lies entirely outside the input contributes only zeros.
__device__ __forceinline__ bool tap_active(const IgemmParams& p, const TileCoord& t, int dy, int dx) {
  const int ya = t.y0 + dy, xa = t.x0 + dx;
  return !(ya + p.bh <= 0 || ya >= p.in_H || xa + p.bw <= 0 || xa >= p.in_W);
}

__device__ __forceinline__ uint64_t tap_mask(const IgemmParams& p, const TapTable& taps, const TileCoord& t) {
  uint64_t m = 0;
  for (int i = 0; i < p.ntaps; ++i)
    if (tap_active(p, t, taps.dy[i], taps.dx[i])) m |= (1ull << i);
  if (m == 0) m = (p.ntaps >= 64) ? ~0ull : ((1ull << p.ntaps) - 1);  // all-zero tile: still produce zeros
  return m;
}

// owner of entry e when `total` entries are cut into T ranges [total*t/T, total*(t+1)/T)
__host__ __device__ __forceinline__ int ts_owner(int e, int T, int total) {
  return (int)((((int64_t)e + 1) * T - 1) / total);
}

// The sequence of (tile, k-step range) work units of one CTA; all three warp roles walk it identically.
struct Work {
  TileCoord t;       // t.split = partial-sum slice
  uint64_t tm;       // taps of the tile this unit walks
  int lo, hi;        // k-step range within the unit's (tap, k-chunk) walk
  int partial;       // 1: the unit stores fp32 partial sums into its workspace slice (a finish kernel completes the tile)
};

struct WorkIter {
  int tile, tlo, thi, cta;
  __device__ __forceinline__ void init(const IgemmParams& p) {
    tile = p.pair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    tlo = thi = cta = 0;
    if (p.ts) {
      cta = blockIdx.x % p.ts_G;
      tlo = (int)(((int64_t)p.ts_total * cta) / p.ts_G);
      thi = (int)(((int64_t)p.ts_total * (cta + 1)) / p.ts_G);
      tile = 0;
    }
  }
  __device__ __forceinline__ bool next(const IgemmParams& p, const TapTable& taps, Work& w) {
    if (p.pair) {
      // unit = (channel tile, pair of consecutive pixel tiles); both CTAs walk the union of their tiles' active taps so
      // that the pair issues one k-step sequence (a tap that is all padding for one of them loads zeros there)
      const int m_tiles = p.tiles_n * p.tiles_h * p.tiles_w;
      const int pairs = (m_tiles + 1) >> 1;
      if (tile >= pairs * p.n_tiles) return false;
      const int unit = tile;
      tile += (int)(gridDim.x >> 1);
      int mp, nt;
      if (p.m_fastest) { mp = unit % pairs; nt = unit / pairs; } else { nt = unit % p.n_tiles; mp = unit / p.n_tiles; }
      const int rank = (int)(blockIdx.x & 1);
      w.t = decode_mtile(p, 2 * mp + rank, nt);
      const TileCoord peer = decode_mtile(p, 2 * mp + (rank ^ 1), nt);
      w.tm = tap_mask(p, taps, w.t) | tap_mask(p, taps, peer);
      w.partial = 0;
      w.lo = 0;
      w.hi = __popcll(w.tm) * p.kchunks;
      return true;
    }
    if (!p.ts) {
      const int tiles = p.phases * p.tiles_n * p.tiles_h * p.tiles_w * p.n_tiles;
      // units: every tile in ksplits pieces -- or, hybrid (p.hyb): the first hyb_full tiles whole (whole waves
      // of the persistent grid) and only the remainder in ksplits pieces, so the last, partly filled wave is spread
      // over all SMs (conv5_x at B=32: 180 tiles on 148 SMs ran as two waves)
      const int whole = p.hyb ? p.hyb_full : 0;
      const int total_units = whole + (tiles - whole) * p.ksplits;
      if (tile >= total_units) return false;
      const int unit = tile;
      tile += gridDim.x;
      int tl, split, nsp;
      const int nsplit_units = (tiles - whole) * p.ksplits;
      // the split units come FIRST: their fp32 partial stores then run under the MMAs of the CTA's whole tiles
      if (unit >= nsplit_units) {
        tl = unit - nsplit_units; split = 0; nsp = 1;
      } else {
        const int r = unit, rem = tiles - whole;
        nsp = p.ksplits;
        if (p.m_fastest || p.hyb) {
          // split slowest: the CTAs running at the same time walk the same K range, i.e. share weight slices
          split = r / rem; tl = whole + r % rem;
        } else {
          split = r % nsp; tl = whole + r / nsp;
        }
      }
      w.t = decode_tile(p, tl);
      w.t.split = split;
      w.partial = nsp > 1;
      w.tm = tap_mask(p, taps, w.t);
      const int nall = __popcll(w.tm) * p.kchunks;
      // balanced partition: no split is empty because ksplits <= kchunks <= nall
      w.lo = (int)(((int64_t)nall * split) / nsp);
      w.hi = (int)(((int64_t)nall * (split + 1)) / nsp);
      return true;
    }
    while (tile < p.ts_tiles) {
      const int tau = tile++;
      const int cs = p.ts_prefix[tau], ce = p.ts_prefix[tau + 1];
      if (ce <= tlo) continue;
      if (cs >= thi) return false;
      const int a = (tlo > cs ? tlo : cs) - cs, b = (thi < ce ? thi : ce) - cs;
      const int col = tau / p.tiles_n;
      w.t.x0 = (col % p.tiles_w) * p.bw;
      w.t.y0 = (col / p.tiles_w) * p.bh;
      w.t.n0 = (tau % p.tiles_n) * p.bn;
      w.t.nt = blockIdx.x / p.ts_G;
      w.t.phase = 0;
      w.t.split = cta - ts_owner(cs, p.ts_G, p.ts_total);
      const uint64_t full = tap_mask(p, taps, w.t);
      uint64_t sel = 0;
      int ord = 0;
      for (int i = 0; i < p.ntaps; ++i)
        if ((full >> i) & 1) {
          if (ord >= a && ord < b) sel |= (1ull << i);
          ++ord;
        }
      w.tm = sel;
      w.lo = 0;
      w.hi = (b - a) * p.kchunks;
      w.partial = 1;
      return true;
    }
    return false;
  }
};

// Column sums over the 32 rows held by the 32 lanes of a warp, for the 32 columns each lane has in registers:
// a butterfly transpose-reduce (16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 32 x 5).  Lane L returns the sum of
// column L over the warp's rows.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    const bool hi = (lane & m) != 0;
#pragma unroll
    for (int i = 0; i < m; ++i) {
      const float keep = hi ? v[i + m] : v[i];
      const float send = hi ? v[i] : v[i + m];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
    }
  }
  return v[0];
}

// Shared tail of every conv epilogue: v = the fp32 accumulators of 32 consecutive output channels of one
// output pixel (element offset `off` into the output / residual / mask tensors), bv = their bias.
// bias -> residual (skip-add / AddN) -> ReLU -> ReLU mask of the producer -> scale, then the store: fp32,
// bf16 into the swizzled staging row of a TMA store (16-byte pieces pbase..pbase+3 of row `srow`), or
// bf16 straight to global memory.  P = IgemmParams or SlabParams.
// The residual / ReLU-mask values of those 32 channels.  They are loaded ahead of the accumulator (before the wait on
// the MMA barrier for a tile's first column group, during the previous group's arithmetic for the others): issued after
// the TMEM load they cost a full global-memory latency per column group, which made the dgrad epilogues of the
// 64-channel layers longer than their MMAs (ncu r1d: conv1_2 dgrad epilogue-bound).
struct EpiPre {
  uint4 a[4];     // the residual when there is one, else the mask (both at once -- no layer of the three nets -- loads the mask late)
  uint32_t mb;    // the 32 mask bits of these channels (p.mask_bits)
};
template <class P>
__device__ __forceinline__ void epilogue_prefetch(const P& p, int64_t off, EpiPre& e) {
  if (p.mask_bits) e.mb = __ldg(p.mask_bits + (off >> 5));
  const bf16* src = p.residual ? p.residual : p.mask;
  if (src) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src + off);
#pragma unroll
    for (int i = 0; i < 4; ++i) e.a[i] = __ldg(s4 + i);
  }
}

template <class P>
__device__ __forceinline__ void epilogue_apply_store(const P& p, float (&v)[32], const float4 (&bv)[8], const EpiPre& pre,
                                                     int64_t off, uint8_t* sbuf, int srow, int pbase) {
  if (p.bias) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[4 * i] += bv[i].x; v[4 * i + 1] += bv[i].y; v[4 * i + 2] += bv[i].z; v[4 * i + 3] += bv[i].w;
    }
  }
  if (p.residual) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 u = pre.a[i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&u.x)[j]);
        v[8 * i + 2 * j] += f.x;
        v[8 * i + 2 * j + 1] += f.y;
      }
    }
  }
  if (p.relu) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (p.mask) {
    const uint4* m4 = reinterpret_cast<const uint4*>(p.mask + off);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 u = p.residual ? __ldg(m4 + i) : pre.a[i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&u.x)[j]);
        if (!(f.x > 0.f)) v[8 * i + 2 * j] = 0.f;
        if (!(f.y > 0.f)) v[8 * i + 2 * j + 1] = 0.f;
      }
    }
  }
  if (p.mask_bits) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (!((pre.mb >> i) & 1u)) v[i] = 0.f;
  }
  if (p.scale != 1.f) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= p.scale;
  }
  if (p.bits_out) {
    // from the rounded bf16 values, so that the bits equal [stored tensor > 0] exactly
    uint32_t word = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) word |= bf16x2_pos_bits(pack_bf16x2(v[2 * i], v[2 * i + 1])) << (2 * i);
    p.bits_out[off >> 5] = word;
  }
  if (p.out_f32 && p.out_cols) {
    const int col0 = (int)(off % p.ldo);
    if (col0 < p.out_cols) {
      float* o = reinterpret_cast<float*>(p.out) + (off / p.ldo) * p.out_cols + col0;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.out_cols) o[i] = v[i];
    }
  } else if (p.out_f32) {
    float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off);
#pragma unroll
    for (int i = 0; i < 8; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else if (p.tma_store) {
    // staging row = box-linear pixel index, 128 B per row, 16-byte pieces XOR-swizzled by row & 7
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<uint4*>(sbuf + srow * 128 + (((pbase + i) ^ (srow & 7)) << 4)) =
          make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                     pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
  } else {
    uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + off);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      o4[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                         pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
  }
}

// 2x2 / stride-2 max-pool of one staged tile: `sbuf` holds the tile's bf16 outputs as written by epilogue_apply_store
// (row = box-linear pixel index at pitch bw, 128 B = 64 channels per row, 16-byte pieces XOR-swizzled by row & 7).
// Item = (pooled pixel of the box, 16-byte piece): the first maximal element under a strict '>' scan in (dy,dx)
// row-major order, exactly as maxpool_fwd_kernel (elementwise.cu) computes it from the stored tensor -- the values
// pooled here are the same rounded bf16 values, so the fused and the two-kernel paths agree bit for bit.
template <class P>
__device__ __forceinline__ void epilogue_pool_store(const P& p, const uint8_t* sbuf, int item, int bw, int bh, int x0, int y0,
                                                    int n0, int ch0) {
  const int pc = item & 7;
  int pp = item >> 3;
  const int hw = bw >> 1, hh = bh >> 1;
  const int ppx = pp % hw;
  pp /= hw;
  const int ppy = pp % hh;
  const int pn = pp / hh;
  const int ix = x0 + 2 * ppx, iy = y0 + 2 * ppy, n = n0 + pn;
  if (ix >= p.W || iy >= p.H || n >= p.N) return;
  const int r00 = (pn * bh + 2 * ppy) * bw + 2 * ppx;
  const int rows[4] = {r00, r00 + 1, r00 + bw, r00 + bw + 1};
  uint4 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const uint4*>(sbuf + rows[k] * 128 + ((pc ^ (rows[k] & 7)) << 4));
  uint32_t outw[4];
  uint32_t idxw[2] = {0u, 0u};
  if (p.relu) {
    // ReLU outputs are non-negative (never -0: max(-0, +0) = +0), and non-negative bf16 patterns order like 16-bit
    // integers: per-halfword SIMD max / compare instead of unpacking to fp32.  First-max index = the first window
    // position whose value equals the maximum -- the element a strict '>' scan stops at.
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t w0 = (&v[0].x)[j], w1 = (&v[1].x)[j], w2 = (&v[2].x)[j], w3 = (&v[3].x)[j];
      const uint32_t m = __vmaxu2(__vmaxu2(w0, w1), __vmaxu2(w2, w3));
      const uint32_t e0 = __vcmpeq2(w0, m), e1 = __vcmpeq2(w1, m), e2 = __vcmpeq2(w2, m);      // 0xffff per equal half
      // per half: e0 ? 0 : e1 ? 1 : e2 ? 2 : 3
      const uint32_t k = ~e0 & ((e1 & 0x00010001u) | (~e1 & ((e2 & 0x00020002u) | (~e2 & 0x00030003u))));
      outw[j] = m;
      const int e = 2 * j;
      idxw[e >> 2] |= ((k & 0xffu) << (8 * (e & 3))) | (((k >> 16) & 0xffu) << (8 * ((e + 1) & 3)));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t w0 = (&v[0].x)[j];
      float2 best = unpack_bf16x2(w0);
      uint32_t bits_lo = w0 & 0xffffu, bits_hi = w0 >> 16;
      uint32_t k_lo = 0, k_hi = 0;
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        const uint32_t wk = (&v[k].x)[j];
        const float2 f = unpack_bf16x2(wk);
        if (f.x > best.x) { best.x = f.x; bits_lo = wk & 0xffffu; k_lo = k; }
        if (f.y > best.y) { best.y = f.y; bits_hi = wk >> 16; k_hi = k; }
      }
      outw[j] = bits_lo | (bits_hi << 16);
      const int e = 2 * j;
      idxw[e >> 2] |= (k_lo << (8 * (e & 3))) | (k_hi << (8 * ((e + 1) & 3)));
    }
  }
  const int64_t o = (((int64_t)n * (p.H >> 1) + (iy >> 1)) * (p.W >> 1) + (ix >> 1)) * p.ldo + ch0 + pc * 8;
  *reinterpret_cast<uint4*>(p.pool_out + o) = make_uint4(outw[0], outw[1], outw[2], outw[3]);
  *reinterpret_cast<uint2*>(p.pool_idx + o) = make_uint2(idxw[0], idxw[1]);
}

// PAIR (BLOCK_N = 256 only; launched as clusters of two CTAs, igemm_pair_kernel): the two CTAs of a cluster run two
// consecutive pixel tiles of one channel tile as a single cta_group::2 MMA with M = 256.  Each CTA stages its own A box and
// HALF of the weight tile (128 rows), so a stage is 32 KB instead of 48 (six stages instead of four), and per SM both the
// operand reads from shared memory (96 -> 64 B/clk) and the weight bytes arriving by TMA halve.  The leader (rank 0)
// owns the full barriers (both CTAs' TMA bytes complete on them), issues the MMAs and multicasts their commits to both
// CTAs' empty / accumulator-full barriers; both CTAs' epilogue threads arrive on the leader's accumulator-empty barriers.
// Epilogues are unchanged: each CTA's TMEM holds the 128 accumulator rows of its own pixel tile.
template <int BLOCK_N, bool PAIR>
__device__ __forceinline__ void igemm_body(const TensorMaps& maps, const IgemmParams& p, const TapTable& taps) {
  using C = Cfg<BLOCK_N>;
  static_assert(!PAIR || BLOCK_N == 256, "CTA-pair mode is built for the 256-column tile");
  constexpr int kBHalf = BLOCK_N * 64;                                      // bytes of half a weight tile
  constexpr int STB = PAIR ? kABytes + kBHalf : C::kStageBytes;             // stage bytes
  constexpr int NST = PAIR ? kSmemBudget / STB : C::kStages;                // stages
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_out = smem + NST * STB;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_out + kOutBytes);
  uint64_t* empty_bar = full_bar + NST;
  uint64_t* tfull_bar = empty_bar + NST;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t crank = 0;
  if constexpr (PAIR) crank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(PAIR ? &maps.b2 : &maps.b);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], PAIR ? 2 * kEpiThreads : kEpiThreads);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair<C::kTmemCols>(tmem_slot);
    else tmem_alloc<C::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();       // both CTAs' barriers exist before either one's TMA / commits reach them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Producer and MMA warps run warp-converged and elect one lane only around the issue itself: the
  // single-lane (divergent) form made ptxas marshal every descriptor through ELECT / R2UR loops,
  // ~125 SASS instructions per k-step, which left the BLOCK_N = 64 layers MMA-issue-bound (ncu r1).
  if (warp == 0) {
    PipeState ps;
    const uint32_t tx_bytes = (uint32_t)p.rows * 128u + (uint32_t)C::kBBytes;
    WorkIter it;
    it.init(p);
    Work w;
    if (p.ts) {
      // lockstep tap-split: k-chunk outer, this CTA's (tile, tap) pairs inner
      Work u[2];
      int nu = 0;
      while (nu < 2 && it.next(p, taps, u[nu])) ++nu;
      for (int kc = 0; kc < p.kchunks; ++kc)
        for (int k = 0; k < nu; ++k) {
          const TileCoord& t = u[k].t;
          for (uint64_t m = u[k].tm; m; m &= m - 1) {
            const int i = __ffsll((long long)m) - 1;
            mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
            if (elect_one()) {
              uint8_t* sa = smem + ps.stage * STB;
              mbar_arrive_expect_tx(&full_bar[ps.stage], tx_bytes);
              tma_load_4d(&maps.a[taps.map[i]], &full_bar[ps.stage], sa, kc * kBlockK, t.x0 + taps.dx[i], t.y0 + taps.dy[i], t.n0);
              tma_load_4d(&maps.b, &full_bar[ps.stage], sa + kABytes, 0, t.nt * BLOCK_N, kc, i);
            }
            __syncwarp();
            ps.advance<NST>();
          }
        }
    } else
    while (it.next(p, taps, w)) {
      const TileCoord& t = w.t;
      int step = 0;
      for (int i = 0; i < p.ntaps; ++i) {
        if (!((w.tm >> i) & 1)) continue;
        const int dy = taps.dy[i], dx = taps.dx[i], mi = taps.map[i];
        for (int kc = 0; kc < p.kchunks; ++kc, ++step) {
          if (step < w.lo || step >= w.hi) continue;
          mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
          if (elect_one()) {
            uint8_t* sa = smem + ps.stage * STB;
            if constexpr (PAIR) {
              // the leader's barrier counts the bytes of both CTAs' loads
              const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[ps.stage]), 0);
              if (crank == 0) mbar_arrive_expect_tx(&full_bar[ps.stage], 2u * ((uint32_t)p.rows * 128u + (uint32_t)kBHalf));
              tma_load_4d_pair(&maps.a[mi], lead_full, sa, kc * kBlockK, t.x0 + dx, t.y0 + dy, t.n0);
              tma_load_4d_pair(&maps.b2, lead_full, sa + kABytes, 0, t.nt * BLOCK_N + (int)crank * (BLOCK_N / 2), kc, i);
            } else {
              mbar_arrive_expect_tx(&full_bar[ps.stage], tx_bytes);
              tma_load_4d(&maps.a[mi], &full_bar[ps.stage], sa, kc * kBlockK, t.x0 + dx, t.y0 + dy, t.n0);
              tma_load_4d(&maps.b, &full_bar[ps.stage], sa + kABytes, 0, t.nt * BLOCK_N, kc, t.phase * p.ntaps + i);
            }
          }
          __syncwarp();
          ps.advance<NST>();
        }
      }
    }
  } else if (warp == 1 && crank == 0) {
    PipeState ps;
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr uint32_t idesc = make_idesc(PAIR ? 2 * kBlockM : kBlockM, BLOCK_N, 0, 0);
    const uint32_t smem_lo = smem_u32(smem) >> 4;          // descriptor start-address units (16 B)
    WorkIter it;
    it.init(p);
    Work w;
    if (p.ts) {
      // both accumulators stay live for the whole walk (first and only use: no wait on tempty)
      Work u[2];
      int nu = 0;
      while (nu < 2 && it.next(p, taps, u[nu])) ++nu;
      for (int kc = 0; kc < p.kchunks; ++kc)
        for (int k = 0; k < nu; ++k) {
          const uint32_t d_addr = tmem_base + (uint32_t)(k * BLOCK_N);
          const int n = __popcll(u[k].tm);
          for (int sidx = 0; sidx < n; ++sidx) {
            mbar_wait(&full_bar[ps.stage], ps.phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_lo = smem_lo + (uint32_t)ps.stage * (uint32_t)(STB >> 4);
              const uint32_t b_lo = a_lo + (uint32_t)(kABytes >> 4);
#pragma unroll
              for (int kk = 0; kk < kBlockK / 16; ++kk)
                umma_f16(d_addr, kDescKMajor | (uint64_t)(a_lo + 2 * kk), kDescKMajor | (uint64_t)(b_lo + 2 * kk), idesc,
                         (kc | sidx | kk) != 0 ? 1u : 0u);
              umma_commit(&empty_bar[ps.stage]);
            }
            __syncwarp();
            ps.advance<NST>();
          }
        }
      if (elect_one()) {
        umma_commit(&tfull_bar[0]);
        if (nu > 1) umma_commit(&tfull_bar[1]);
      }
      __syncwarp();
    } else
    while (it.next(p, taps, w)) {
      const int nsteps = w.hi - w.lo;   // > 0
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_addr = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int ks = 0; ks < nsteps; ++ks) {
        mbar_wait(&full_bar[ps.stage], ps.phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = smem_lo + (uint32_t)ps.stage * (uint32_t)(STB >> 4);
          const uint32_t b_lo = a_lo + (uint32_t)(kABytes >> 4);
          if constexpr (PAIR) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_f16_pair(d_addr, kDescKMajor | (uint64_t)(a_lo + 2 * k), kDescKMajor | (uint64_t)(b_lo + 2 * k), idesc,
                            (ks | k) != 0 ? 1u : 0u);
            umma_commit_pair(&empty_bar[ps.stage]);
            if (ks == nsteps - 1) umma_commit_pair(&tfull_bar[acc]);
          } else {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_f16(d_addr, kDescKMajor | (uint64_t)(a_lo + 2 * k), kDescKMajor | (uint64_t)(b_lo + 2 * k), idesc,
                       (ks | k) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[ps.stage]);
            if (ks == nsteps - 1) umma_commit(&tfull_bar[acc]);
          }
        }
        __syncwarp();
        ps.advance<NST>();
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 2) {
    // ---- epilogue: warps 2..9, TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 ----
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int iw = row % p.bw;
    const int ih = (row / p.bw) % p.bh;
    const int in = row / (p.bw * p.bh);
    const bool ep_leader = (threadIdx.x == 64);     // first epilogue thread: issues / tracks the TMA stores
    uint32_t sg = 0;                                // running 64-column group counter -> staging buffer
    float ca0 = 0.f, ca1 = 0.f, ca2 = 0.f, ca3 = 0.f;      // column sums of this thread's columns (p.colsum)
    WorkIter it;
    it.init(p);
    Work w;
    while (it.next(p, taps, w)) {
      const TileCoord& t = w.t;
      const bool partial = w.partial != 0;
      const int ay = t.phase / p.s, ax = t.phase % p.s;
      const int qx = t.x0 + iw, qy = t.y0 + ih, n = t.n0 + in;
      const int ox = qx * p.os + ax - p.opad, oy = qy * p.os + ay - p.opad;
      const bool valid = row < p.rows && qx < p.W && qy < p.H && n < p.N && ox >= 0 && ox < p.out_W &&
                         oy >= 0 && oy < p.out_H;
      const int64_t opix = ((int64_t)n * p.out_H + oy) * p.out_W + ox;
      const int64_t obase = opix * p.ldo + (int64_t)t.nt * BLOCK_N;
      const bool full_epi = valid && !partial && !p.pack_s;
      const bool do_tma = p.tma_store && !partial;      // (uniform over the epilogue warps: barriers stay matched)
      EpiPre pre;
      if (full_epi) epilogue_prefetch(p, obase + half * 32, pre);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
      for (int g0 = 0; g0 < BLOCK_N; g0 += 64) {
        const int c0 = g0 + half * 32;
        uint8_t* sbuf = smem_out + (sg & 1) * kStageOutBytes;
        if (do_tma) {
          // the store that used this staging buffer two groups ago must have finished reading it
          if (ep_leader) tma_store_wait_read<1>();
          named_bar_sync(1, kEpiThreads);
        }
        // bias for these 32 columns: issued before the TMEM load so its L2/L1 latency hides behind it
        float4 bv[8];
        if (p.bias && !p.pack_s) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + t.nt * BLOCK_N + c0);
#pragma unroll
          for (int i = 0; i < 8; ++i) bv[i] = __ldg(b4 + i);
        }
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
        if (p.pack_s) {
          // 32 columns = 32 / (pack_s * pack_co) output rows `a`, each pack_s pixels x pack_co classes contiguous
          const bool rv = row < p.rows && qx < p.W && qy < p.H && n < p.N;
          const int run = p.pack_s * p.pack_co;
          if (rv) {
            float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const int col = t.nt * BLOCK_N + c0 + i;
              const int a = col / run, b = (col % run) / p.pack_co, co = col % p.pack_co;
              const int oy2 = qy * p.pack_s - p.opad + a, ox2 = qx * p.pack_s - p.opad + b;
              if (oy2 < 0 || oy2 >= p.out_H || ox2 < 0 || ox2 >= p.out_W) continue;
              const int64_t off = (((int64_t)n * p.out_H + oy2) * p.out_W + ox2) * p.pack_co + co;
              float v0 = __uint_as_float(r[i]), v1 = __uint_as_float(r[i + 1]);
              if (p.bias) { v0 += __ldg(p.bias + co); v1 += __ldg(p.bias + co + 1); }
              *reinterpret_cast<float2*>(o + off) = make_float2(v0, v1);      // pack_co is even: (co, co+1) of one pixel
            }
          }
        } else if (valid && partial) {
          float4* w4 = reinterpret_cast<float4*>(p.ws + (int64_t)t.split * p.ws_slice + obase + c0);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            w4[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                __uint_as_float(r[4 * i + 3]));
        } else {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = valid ? __uint_as_float(r[i]) : 0.f;
          const EpiPre cur = pre;
          if (full_epi && g0 + 64 < BLOCK_N) epilogue_prefetch(p, obase + c0 + 64, pre);
          if (valid) epilogue_apply_store(p, v, bv, cur, obase + c0, sbuf, row, half * 4);
          if (p.colsum) {        // rows outside the tensor contribute zeros
            const float cs = warp_colsum32(v, lane);
            if (g0 == 0) ca0 += cs; else if (g0 == 64) ca1 += cs; else if (g0 == 128) ca2 += cs; else ca3 += cs;
          }
        }
        if (do_tma) {
          fence_proxy_async();                        // generic-proxy smem writes -> visible to the TMA engine
          named_bar_sync(1, kEpiThreads);
          if (ep_leader && !p.pool_only && t.n0 < p.N) {      // (n0 >= N: the odd tile's partner in CTA-pair mode)
            if (p.phases > 1) {
              // transposed conv: out pixel = q * s + a - opad; phase (ay, ax)'s pixels are a decimated view of y whose
              // index is q (a >= opad) or q - 1 (a < opad: q = 0 falls outside the view and is clipped by the store)
              tma_store_4d(&maps.cph[t.phase], sbuf, t.nt * BLOCK_N + g0, t.x0 - (ax < p.opad ? 1 : 0), t.y0 - (ay < p.opad ? 1 : 0), t.n0);
            } else {
              tma_store_4d(&maps.c, sbuf, t.nt * BLOCK_N + g0, t.x0, t.y0, t.n0);
            }
            tma_store_commit();
          }
          ++sg;
          // (the staging buffer is rewritten two groups later, behind that group's barrier: every thread has finished
          // these reads by then)
          if (p.pool_out) {
            const int items = (p.rows >> 2) * 8;
            for (int item = (int)threadIdx.x - 64; item < items; item += kEpiThreads)
              epilogue_pool_store(p, sbuf, item, p.bw, p.bh, t.x0, t.y0, t.n0, t.nt * BLOCK_N + g0);
          }
        }
      }
      tc_fence_before();
      if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));   // the leader's MMA warp waits for both epilogues
      else mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.tma_store && ep_leader) tma_store_wait_read<0>();
    if (p.colsum) {
      // partial row of this CTA: all its tiles have channel tile nt; slot = its index among the CTAs of that channel tile
      const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
      const int nt = unit0 % p.n_tiles;
      const int slot = PAIR ? (unit0 / p.n_tiles) * 2 + (int)(blockIdx.x & 1) : unit0 / p.n_tiles;
      const int rows_per_nt = (gridDim.x / p.n_tiles) * 4;
      float* dst = p.colsum + ((int64_t)nt * rows_per_nt + slot * 4 + q) * BLOCK_N + half * 32 + lane;
      dst[0] = ca0;
      if (BLOCK_N > 64) dst[64] = ca1;
      if (BLOCK_N > 128) { dst[128] = ca2; dst[192] = ca3; }
    }
  }
  __syncwarp();
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();       // the peer's shared memory and barriers stay alive until the leader's last MMA / commit
  else __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    if constexpr (PAIR) tmem_dealloc_pair<C::kTmemCols>(tmem_base);
    else tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
igemm_kernel(const __grid_constant__ TensorMaps maps, const IgemmParams p, const TapTable taps) {
  igemm_body<BLOCK_N, false>(maps, p, taps);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
igemm_pair_kernel(const __grid_constant__ TensorMaps maps, const IgemmParams p, const TapTable taps) {
  igemm_body<256, true>(maps, p, taps);
}

// ------------------------------------------------------------------------------------------
// slab_kernel: 3x3 stride-1 conv (fwd / dgrad) for large maps with few output channels, where the
// tap-wise igemm is bound by re-reading the activations 9x from L2 (measured: conv1_2 at 0.36 PF/s).
// One TMA loads a haloed slab (kSlabH+2 image rows x 32 columns x 64 channels) per 64-channel step;
// all 9 taps are UMMA descriptors into that SAME slab: output "row" o = i*32 + j (i < 4 image rows,
// j < 32 slab columns, the last two are discarded halo columns) reads slab row o + ty*32 + tx, so a
// tap is just a start-row offset, including offsets that are not multiples of the 8-row swizzle atom:
// measured on B200, the 128B swizzle is a function of the absolute shared-memory address on both the
// TMA write and the UMMA read side, so the descriptor's base-offset field stays 0 (setting it to
// (start >> 7) & 7 gave wrong results in the round-1 diagnostic run).  Activation traffic drops from 9x to
// (6/4)*(32/30) = 1.6x.  Separate rings for slabs and weight tiles.
// ------------------------------------------------------------------------------------------
constexpr int kSlabP = 32;        // slab pitch in pixels
constexpr int kSlabH = 4;         // output image rows per tile
constexpr int kSlabWV = 30;       // valid output columns per tile
constexpr int kSlabRows = (kSlabH + 2) * kSlabP;          // 192 rows written by TMA (+2 spill rows read)
constexpr int kSlabBytes = 25600;                          // 194 rows * 128 B rounded up to 1024

template <int BLOCK_N>
struct SlabCfg {
  // weight tiles travel in groups of kGroup taps per ring stage: one barrier round-trip then feeds
  // 4*kGroup MMAs (with BLOCK_N = 64 a single tap is only 128 tensor cycles, less than the wait costs)
  static constexpr int kGroup = BLOCK_N == 256 ? 1 : 3;
  static constexpr int kTapBytes = BLOCK_N * 128;
  static constexpr int kBBytes = kGroup * kTapBytes;
  static constexpr int kAStages = BLOCK_N == 256 ? 2 : 3;
  static constexpr int kBStages = BLOCK_N == 256 ? 3 : (BLOCK_N == 128 ? 2 : 4);
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kSmemBytes = kAStages * kSlabBytes + kBStages * kBBytes + kOutBytes + 1024 + 512;
};

struct SlabParams {
  int N, H, W;
  int tiles_w, tiles_h;   // ceil(W/30), H/4
  int n_tiles;            // Cn / BLOCK_N
  int kchunks;            // Ck / 64
  int ldo;
  int tma_store;
  float* colsum;          // see IgemmParams
  bf16* pool_out;         // fused 2x2 max-pool of the output (see IgemmParams)
  uint8_t* pool_idx;
  int pool_only;
  uint32_t* bits_out;     // 1-bit ReLU masks out / in (see IgemmParams)
  const uint32_t* mask_bits;
  int out_cols;           // narrow fp32 output (see IgemmParams)
  void* out;
  int out_f32;
  const float* bias;
  const bf16* residual;
  const bf16* mask;
  float scale;
  int relu;
};

// One tile of the slab schedule: channel tile nt, output origin (x0, y0) of image n.  PAIR: CTAs 2q / 2q+1 of a cluster take two
// consecutive pixel tiles of the same channel tile (n == p.N for the partner of an odd last tile: its slab loads zero-fill,
// nothing is stored).
struct SlabTile {
  int nt, x0, y0, n;
};
template <bool PAIR>
__device__ __forceinline__ bool slab_next(const SlabParams& p, int& cursor, SlabTile& t) {
  const int m_tiles = p.N * p.tiles_h * p.tiles_w;
  int r;
  if (PAIR) {
    const int pairs = (m_tiles + 1) >> 1;
    if (cursor >= pairs * p.n_tiles) return false;
    t.nt = cursor % p.n_tiles;
    r = 2 * (cursor / p.n_tiles) + (int)(blockIdx.x & 1);
    cursor += (int)(gridDim.x >> 1);
  } else {
    if (cursor >= m_tiles * p.n_tiles) return false;
    t.nt = cursor % p.n_tiles;
    r = cursor / p.n_tiles;
    cursor += (int)gridDim.x;
  }
  t.x0 = (r % p.tiles_w) * kSlabWV;
  r /= p.tiles_w;
  t.y0 = (r % p.tiles_h) * kSlabH;
  t.n = r / p.tiles_h;
  return true;
}

// PAIR: the CTA-pair form of igemm_body (cta_group::2, M = 256): each CTA stages its own slab and HALF of every weight tap
// (BLOCK_N / 2 rows, TensorMaps::b2); the leader owns the full barriers of both rings, issues the MMAs and multicasts the
// commits.  With BLOCK_N = 128 the operand reads per SM drop from 128 B/clk -- the shared-memory limit -- to 96.
template <int BLOCK_N, bool PAIR>
__device__ __forceinline__ void slab_body(const TensorMaps& maps, const SlabParams& p) {
  using C = SlabCfg<BLOCK_N>;
  constexpr int kTapHalf = C::kTapBytes / 2;
  constexpr int kTapStride = PAIR ? kTapHalf : C::kTapBytes;       // bytes between the taps of a weight stage
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::kAStages * kSlabBytes;
  uint8_t* smem_out = smem_b + C::kBStages * C::kBBytes;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(smem_out + kOutBytes);
  uint64_t* emptyA = fullA + C::kAStages;
  uint64_t* fullB = emptyA + C::kAStages;
  uint64_t* emptyB = fullB + C::kBStages;
  uint64_t* tfull_bar = emptyB + C::kBStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t crank = 0;
  if constexpr (PAIR) crank = cluster_ctarank();
  const int cursor0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(PAIR ? &maps.b2 : &maps.b);
    for (int i = 0; i < C::kAStages; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < C::kBStages; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], PAIR ? 2 * kEpiThreads : kEpiThreads); }
    fence_barrier_init();
  }
  // the two spill rows past each slab are read (for discarded outputs only) but never written by
  // TMA: zero the slab ring once so they can never hold NaN patterns that matter to no one
  for (int i = threadIdx.x; i < C::kAStages * kSlabBytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(smem_a)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair<C::kTmemCols>(tmem_slot);
    else tmem_alloc<C::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    PipeState pa, pb;
    int cursor = cursor0;
    SlabTile t;
    while (slab_next<PAIR>(p, cursor, t)) {
      const int nt = t.nt, x0 = t.x0, y0 = t.y0, n = t.n;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&emptyA[pa.stage], pa.phase ^ 1);
        if (elect_one()) {
          if constexpr (PAIR) {
            if (crank == 0) mbar_arrive_expect_tx(&fullA[pa.stage], 2u * (uint32_t)kSlabRows * 128u);
            tma_load_4d_pair(&maps.a[0], mapa_u32(smem_u32(&fullA[pa.stage]), 0), smem_a + pa.stage * kSlabBytes, kc * kBlockK, x0 - 1,
                             y0 - 1, n);
          } else {
            mbar_arrive_expect_tx(&fullA[pa.stage], (uint32_t)kSlabRows * 128u);
            tma_load_4d(&maps.a[0], &fullA[pa.stage], smem_a + pa.stage * kSlabBytes, kc * kBlockK, x0 - 1, y0 - 1, n);
          }
        }
        __syncwarp();
        pa.advance<C::kAStages>();
        for (int tg = 0; tg < 9; tg += C::kGroup) {
          mbar_wait(&emptyB[pb.stage], pb.phase ^ 1);
          if (elect_one()) {
            if constexpr (PAIR) {
              const uint32_t lead = mapa_u32(smem_u32(&fullB[pb.stage]), 0);
              if (crank == 0) mbar_arrive_expect_tx(&fullB[pb.stage], (uint32_t)C::kBBytes);        // both CTAs' halves
#pragma unroll
              for (int j = 0; j < C::kGroup; ++j)
                tma_load_4d_pair(&maps.b2, lead, smem_b + pb.stage * C::kBBytes + j * kTapHalf, 0,
                                 nt * BLOCK_N + (int)crank * (BLOCK_N / 2), kc, tg + j);
            } else {
              mbar_arrive_expect_tx(&fullB[pb.stage], (uint32_t)C::kBBytes);
#pragma unroll
              for (int j = 0; j < C::kGroup; ++j)
                tma_load_4d(&maps.b, &fullB[pb.stage], smem_b + pb.stage * C::kBBytes + j * C::kTapBytes, 0, nt * BLOCK_N, kc,
                            tg + j);
            }
          }
          __syncwarp();
          pb.advance<C::kBStages>();
        }
      }
    }
  } else if (warp == 1 && crank == 0) {
    PipeState pa, pb;
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr uint32_t idesc = make_idesc(PAIR ? 2 * kBlockM : kBlockM, BLOCK_N, 0, 0);
    const uint32_t a_lo0 = smem_u32(smem_a) >> 4, b_lo0 = smem_u32(smem_b) >> 4;
    int cursor = cursor0;
    SlabTile t;
    while (slab_next<PAIR>(p, cursor, t)) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_addr = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&fullA[pa.stage], pa.phase);
        const uint32_t slab_lo = a_lo0 + (uint32_t)pa.stage * (uint32_t)(kSlabBytes >> 4);
#pragma unroll 1
        for (int tg = 0; tg < 9; tg += C::kGroup) {
          mbar_wait(&fullB[pb.stage], pb.phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b_lo = b_lo0 + (uint32_t)pb.stage * (uint32_t)(C::kBBytes >> 4);
#pragma unroll
            for (int j = 0; j < C::kGroup; ++j) {
              const int tap = tg + j;
              // tap (ty,tx) = slab row offset ty*32 + tx; one row = 128 B = 8 descriptor units
              const uint32_t a_lo = slab_lo + (uint32_t)((tap / 3) * kSlabP + (tap % 3)) * 8u;
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                const uint64_t da = kDescKMajor | (uint64_t)(a_lo + 2 * k);
                const uint64_t db = kDescKMajor | (uint64_t)(b_lo + j * (kTapStride >> 4) + 2 * k);
                if constexpr (PAIR) umma_f16_pair(d_addr, da, db, idesc, (kc | tap | k) != 0 ? 1u : 0u);
                else umma_f16(d_addr, da, db, idesc, (kc | tap | k) != 0 ? 1u : 0u);
              }
            }
            if constexpr (PAIR) {
              umma_commit_pair(&emptyB[pb.stage]);
              if (tg + C::kGroup >= 9) {
                umma_commit_pair(&emptyA[pa.stage]);
                if (kc == p.kchunks - 1) umma_commit_pair(&tfull_bar[acc]);
              }
            } else {
              umma_commit(&emptyB[pb.stage]);
              if (tg + C::kGroup >= 9) {
                umma_commit(&emptyA[pa.stage]);
                if (kc == p.kchunks - 1) umma_commit(&tfull_bar[acc]);
              }
            }
          }
          __syncwarp();
          pb.advance<C::kBStages>();
        }
        pa.advance<C::kAStages>();
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 2) {
    const int q = warp & 3;       // TMEM lane quarter == image row of the tile
    const int half = (warp - 2) >> 2;   // which 32 columns of each 64-column group
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool ep_leader = (threadIdx.x == 64);
    const int srow = q * kSlabWV + lane;            // staging row: box (64 ch, 30 w, 4 h) is packed at pitch 30
    uint32_t sg = 0;
    float ca0 = 0.f, ca1 = 0.f, ca2 = 0.f, ca3 = 0.f;      // column sums of this thread's columns (p.colsum)
    int cursor = cursor0;
    SlabTile t;
    while (slab_next<PAIR>(p, cursor, t)) {
      const int nt = t.nt, x0 = t.x0, y0 = t.y0, n = t.n;
      const int ox = x0 + lane, oy = y0 + q;
      const bool valid = lane < kSlabWV && ox < p.W && oy < p.H && n < p.N;
      const int64_t obase = (((int64_t)n * p.H + oy) * p.W + ox) * p.ldo + (int64_t)nt * BLOCK_N;
      EpiPre pre;
      if (valid) epilogue_prefetch(p, obase + half * 32, pre);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
      for (int g0 = 0; g0 < BLOCK_N; g0 += 64) {
        const int c0 = g0 + half * 32;
        uint8_t* sbuf = smem_out + (sg & 1) * kStageOutBytes;
        if (p.tma_store) {
          if (ep_leader) tma_store_wait_read<1>();
          named_bar_sync(1, kEpiThreads);
        }
        float4 bv[8];
        if (p.bias) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + nt * BLOCK_N + c0);
#pragma unroll
          for (int i = 0; i < 8; ++i) bv[i] = __ldg(b4 + i);
        }
        uint32_t rr[32];
        tmem_ld32(taddr + c0, rr);
        tmem_ld_wait();
        {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = valid ? __uint_as_float(rr[i]) : 0.f;
          const EpiPre cur = pre;
          if (valid && g0 + 64 < BLOCK_N) epilogue_prefetch(p, obase + c0 + 64, pre);
          if (valid) epilogue_apply_store(p, v, bv, cur, obase + c0, sbuf, srow, half * 4);
          if (p.colsum) {
            const float cs = warp_colsum32(v, lane);
            if (g0 == 0) ca0 += cs; else if (g0 == 64) ca1 += cs; else if (g0 == 128) ca2 += cs; else ca3 += cs;
          }
        }
        if (p.tma_store) {
          fence_proxy_async();
          named_bar_sync(1, kEpiThreads);
          if (ep_leader && !p.pool_only && n < p.N) {
            tma_store_4d(&maps.c, sbuf, nt * BLOCK_N + g0, x0, y0, n);
            tma_store_commit();
          }
          ++sg;
          if (p.pool_out) {
            for (int item = (int)threadIdx.x - 64; item < (kSlabWV / 2) * (kSlabH / 2) * 8; item += kEpiThreads)
              epilogue_pool_store(p, sbuf, item, kSlabWV, kSlabH, x0, y0, n, nt * BLOCK_N + g0);
          }
        }
      }
      tc_fence_before();
      if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
      else mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.tma_store && ep_leader) tma_store_wait_read<0>();
    if (p.colsum) {
      const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
      const int nt = unit0 % p.n_tiles;
      const int slot = PAIR ? (unit0 / p.n_tiles) * 2 + (int)(blockIdx.x & 1) : unit0 / p.n_tiles;
      const int rows_per_nt = (gridDim.x / p.n_tiles) * 4;
      float* dst = p.colsum + ((int64_t)nt * rows_per_nt + slot * 4 + q) * BLOCK_N + half * 32 + lane;
      dst[0] = ca0;
      if (BLOCK_N > 64) dst[64] = ca1;
      if (BLOCK_N > 128) { dst[128] = ca2; dst[192] = ca3; }
    }
  }
  __syncwarp();
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    if constexpr (PAIR) tmem_dealloc_pair<C::kTmemCols>(tmem_base);
    else tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
slab_kernel(const __grid_constant__ TensorMaps maps, const SlabParams p) {
  slab_body<BLOCK_N, false>(maps, p);
}

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
slab_pair_kernel(const __grid_constant__ TensorMaps maps, const SlabParams p) {
  slab_body<BLOCK_N, true>(maps, p);
}

// ------------------------------------------------------------------------------------------
// slab3_kernel: the slab kernel for Ck = 64 and 64-channel output tiles (conv1_2 fwd / dgrad, conv2_1).
// ncu (profiles/r1d): with N = 64 the slab kernel is bound by shared-memory bandwidth, not by the
// tensor pipe: a 128x64x16 MMA reads 6 KB for 32 tensor cycles, and every tile re-loads all nine
// 8 KB weight tiles.  Two changes:
//   * the three kx taps of one ky share the A operand: B = [W(ky,0); W(ky,1); W(ky,2)] is one N = 192
//     operand and column block kx of the accumulator holds  D_kx[r] = sum_ky slab[ky*32 + r] . W(ky,kx).
//     The output is  out[o] = D_0[o] + D_1[o+1] + D_2[o+2]:  rows are TMEM lanes = threads of the
//     epilogue warp, so the kx shift is two warp shuffles per value (o = i*32 + j; the lanes that would
//     wrap, j >= 30, are the discarded halo columns anyway).  10 KB per 96 tensor cycles.
//   * the 72 KB of weights of the CTA's channel tile stay resident in shared memory (every tile of a
//     CTA has the same channel tile: the grid is a multiple of the number of channel tiles).
// ------------------------------------------------------------------------------------------
constexpr int kSlab3Bytes = kSlabRows * 128;                 // 24 KB: A rows ky*32 + r never leave the slab
template <int KCH>                                            // 64-channel chunks of the input (1 or 2)
struct Slab3Cfg {
  static constexpr int kWBytes = KCH * 9 * 64 * 128;         // 72 KB of resident weights per chunk
  static constexpr int kStages = KCH == 1 ? 4 : 2;           // slab ring (KCH = 2: 144 KB of weights leave room for two)
  static constexpr int kSmem = kStages * kSlab3Bytes + kWBytes + kOutBytes + 1024 + 512;
};

template <int KCH>
__global__ void __launch_bounds__(kThreads, 1)
slab3_kernel(const __grid_constant__ TensorMaps maps, const SlabParams p) {
  constexpr int kN = 192;
  constexpr int kSlab3Stages = Slab3Cfg<KCH>::kStages;
  constexpr int kSlab3WBytes = Slab3Cfg<KCH>::kWBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_w = smem + kSlab3Stages * kSlab3Bytes;
  uint8_t* smem_out = smem_w + kSlab3WBytes;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(smem_out + kOutBytes);
  uint64_t* emptyA = fullA + kSlab3Stages;
  uint64_t* tfull_bar = emptyA + kSlab3Stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* wfull = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.N * p.tiles_h * p.tiles_w * p.n_tiles;
  const int nt = blockIdx.x % p.n_tiles;          // the same for every tile of this CTA (grid % n_tiles == 0)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b);
    for (int i = 0; i < kSlab3Stages; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 128); }   // one epilogue group per accumulator
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);      // 2 accumulators x 192 columns
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(wfull, (uint32_t)kSlab3WBytes);
      for (int kc = 0; kc < KCH; ++kc)
        for (int ky = 0; ky < 3; ++ky)     // box (64 k, 64 channels, 3 taps) -> rows kx*64 + c
          tma_load_4d(&maps.b, wfull, smem_w + (kc * 3 + ky) * (3 * 64 * 128), 0, nt * 64, kc, ky * 3);
    }
    __syncwarp();
    PipeState pa;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int r = tile / p.n_tiles;
      const int x0 = (r % p.tiles_w) * kSlabWV;
      r /= p.tiles_w;
      const int y0 = (r % p.tiles_h) * kSlabH;
      const int n = r / p.tiles_h;
      for (int kc = 0; kc < KCH; ++kc) {
        mbar_wait(&emptyA[pa.stage], pa.phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&fullA[pa.stage], (uint32_t)kSlab3Bytes);
          tma_load_4d(&maps.a[0], &fullA[pa.stage], smem_a + pa.stage * kSlab3Bytes, kc * 64, x0 - 1, y0 - 1, n);
        }
        __syncwarp();
        pa.advance<kSlab3Stages>();
      }
    }
  } else if (warp == 1) {
    PipeState pa;
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr uint32_t idesc = make_idesc(kBlockM, kN, 0, 0);
    const uint32_t a_lo0 = smem_u32(smem_a) >> 4, w_lo0 = smem_u32(smem_w) >> 4;
    mbar_wait(wfull, 0);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      for (int kc = 0; kc < KCH; ++kc) {
        mbar_wait(&fullA[pa.stage], pa.phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d_addr = tmem_base + (uint32_t)(acc * 256);
          const uint32_t slab_lo = a_lo0 + (uint32_t)pa.stage * (uint32_t)(kSlab3Bytes >> 4);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const uint32_t a_lo = slab_lo + (uint32_t)(ky * kSlabP) * 8u;            // 32 rows = 4 swizzle atoms down
            const uint32_t b_lo = w_lo0 + (uint32_t)(kc * 3 + ky) * (uint32_t)((3 * 64 * 128) >> 4);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_f16(d_addr, kDescKMajor | (uint64_t)(a_lo + 2 * k), kDescKMajor | (uint64_t)(b_lo + 2 * k), idesc,
                       (kc | ky | k) != 0 ? 1u : 0u);
          }
          umma_commit(&emptyA[pa.stage]);
          if (kc == KCH - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        pa.advance<kSlab3Stages>();
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // Two epilogue groups of four warps (one per TMEM lane quarter); group g owns accumulator g, i.e.
    // every other tile of this CTA.  The epilogue of a tile is a long dependent chain (three TMEM loads at
    // 64 B/clk, two shuffles per value, bias / mask / pack, staging, TMA store: ~2900 cycles measured with
    // all eight warps on one tile) against 1152 tensor cycles of MMAs; with the groups on alternate tiles
    // one group's TMEM reads run under the other's arithmetic and stores.
    const int q = warp & 3;       // TMEM lane quarter == image row of the tile
    const int grp = (warp - 2) >> 2;
    uint32_t acc_phase = 0;
    const bool ep_leader = (threadIdx.x == 64 + grp * 128);
    const int srow = q * kSlabWV + lane;
    uint8_t* sbuf = smem_out + grp * kStageOutBytes;
    for (int tile = blockIdx.x + grp * gridDim.x; tile < total_tiles; tile += 2 * gridDim.x) {
      int r = tile / p.n_tiles;
      const int x0 = (r % p.tiles_w) * kSlabWV;
      r /= p.tiles_w;
      const int y0 = (r % p.tiles_h) * kSlabH;
      const int n = r / p.tiles_h;
      const int ox = x0 + lane, oy = y0 + q;
      const bool valid = lane < kSlabWV && ox < p.W && oy < p.H;
      const int64_t obase = (((int64_t)n * p.H + oy) * p.W + ox) * p.ldo + (int64_t)nt * 64;
      EpiPre pre0, pre1;
      if (valid) {
        epilogue_prefetch(p, obase, pre0);
        epilogue_prefetch(p, obase + 32, pre1);
      }
      mbar_wait(&tfull_bar[grp], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(grp * 256);
      if (p.tma_store) {
        if (ep_leader) tma_store_wait_read<0>();     // this group's previous store has read the staging buffer
        named_bar_sync(1 + grp, 128);
      }
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        float4 bv[8];
        if (p.bias) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + nt * 64 + c0);
#pragma unroll
          for (int i = 0; i < 8; ++i) bv[i] = __ldg(b4 + i);
        }
        float v[32];
        {
          uint32_t r0[32], r1[32], r2[32];
          tmem_ld32(taddr + c0, r0);
          tmem_ld32(taddr + 64 + c0, r1);
          tmem_ld32(taddr + 128 + c0, r2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            v[i] = __uint_as_float(r0[i]) + __shfl_down_sync(0xffffffffu, __uint_as_float(r1[i]), 1) +
                   __shfl_down_sync(0xffffffffu, __uint_as_float(r2[i]), 2);
        }
        if (valid) epilogue_apply_store(p, v, bv, c0 ? pre1 : pre0, obase + c0, sbuf, srow, c0 ? 4 : 0);
      }
      // the accumulator is in registers / staging now: hand it back to the MMA warp before the store
      tc_fence_before();
      mbar_arrive(&tempty_bar[grp]);
      acc_phase ^= 1;
      if (p.tma_store) {
        fence_proxy_async();
        named_bar_sync(1 + grp, 128);
        if (ep_leader && !p.pool_only) {
          tma_store_4d(&maps.c, sbuf, nt * 64, x0, y0, n);
          tma_store_commit();
        }
        // (the group's next tile starts behind a barrier: these reads are done before the buffer is rewritten)
        if (p.pool_out) {
          for (int item = (int)threadIdx.x - 64 - grp * 128; item < (kSlabWV / 2) * (kSlabH / 2) * 8; item += 128)
            epilogue_pool_store(p, sbuf, item, kSlabWV, kSlabH, x0, y0, n, nt * 64);
        }
      }
    }
    if (p.tma_store && ep_leader) tma_store_wait_read<0>();
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------
// wgrad: dW[rb*64 + i][co] = sum over pixels  X[pixel + shift(rb)][ci(rb)*64 + i] * dY[pixel][co]
// A row block rb = (tap, 64-channel chunk of Cin).  One work item = (k-split, pair of row blocks,
// N tile); it walks its range of 64-pixel boxes.  Both operands are MN-major in shared memory
// (a TMA box [64 pixels][64 channels] is exactly one MN-major SWIZZLE_128B block).
// ------------------------------------------------------------------------------------------
struct WgradParams {
  int N, H, W;
  int bw, bh, bn;  // 64-pixel box
  int tiles_w, tiles_h, tiles_n;
  int splits, n_rb, n_rbp, kchunks_in, n_tiles;
  int Cin_total, Cout_total;  // leading dims of dW rows / cols
  int dw_tap_stride;          // elements between taps in dW (= Cin*Cout for HWIO)
  int dw_row_stride;          // elements between consecutive ci rows (= Cout for HWIO)
  int dw_col_stride;          // elements between consecutive co (1 for HWIO)
  int direct;                 // always 1: plain 16-byte stores into dW (single split) or into per-split partial buffers
  int skip_oob;               // 1: pixel boxes whose shifted x boxes lie entirely in the SAME padding (for both row blocks
                              //    of the item) are skipped: no TMA, no MMA (conv6: 7x7 taps on a 5x18 map, 38 % of the pairs)
  int64_t part_stride;        // elements between the partial buffers of consecutive splits (0: dw is the result itself)
  float* dw;
  // with skip_oob the items cost between a few and all of the pixel boxes (conv6: corner taps see 8 of 45): the row-block
  // pairs are visited in order of decreasing cost (host-sorted), so the static round-robin over CTAs stays balanced
  int wide_store;             // dW (or the partial buffers) and its strides are 32-byte aligned: 256-bit stores
  int use_perm;
  uint16_t rbp_perm[kMaxRbp];
  uint16_t rbp_nact[kMaxRbp];   // contributing pixel boxes of the pair at sorted position r (splits == 1): counting them in
                                // the kernel took ~3 us per item in the TMA and MMA warps, a third of a conv6 item
};

// wgrad: does pixel box (x0, y0) contribute to the item's row blocks?  A tap whose shifted box is entirely outside
// the image reads only TMA zero fill.
__device__ __forceinline__ bool wgrad_box_active(const WgradParams& p, int x0, int y0, int dx0, int dy0, int dx1, int dy1) {
  const int xa = x0 + dx0, ya = y0 + dy0, xb = x0 + dx1, yb = y0 + dy1;
  const bool a = !(ya + p.bh <= 0 || ya >= p.H || xa + p.bw <= 0 || xa >= p.W);
  const bool b = !(yb + p.bh <= 0 || yb >= p.H || xb + p.bw <= 0 || xb >= p.W);
  return a || b;
}

// number of contributing boxes among pixel tiles [pt0, pt1) (w fastest, then h, then n); 0 active -> all (the item still
// has to produce its zeros)
__device__ __forceinline__ int wgrad_count_active(const WgradParams& p, int pt0, int pt1, int dx0, int dy0, int dx1, int dy1) {
  if (!p.skip_oob) return pt1 - pt0;
  int x0 = (pt0 % p.tiles_w) * p.bw;
  int y0 = ((pt0 / p.tiles_w) % p.tiles_h) * p.bh;
  int n = 0;
  for (int pt = pt0; pt < pt1; ++pt) {
    n += wgrad_box_active(p, x0, y0, dx0, dy0, dx1, dy1) ? 1 : 0;
    x0 += p.bw;
    if (x0 >= p.tiles_w * p.bw) {
      x0 = 0;
      y0 += p.bh;
      if (y0 >= p.tiles_h * p.bh) y0 = 0;
    }
  }
  return n;
}

__device__ __forceinline__ int wgrad_rbp(const WgradParams& p, int item) {
  const int r = (item / p.n_tiles) % p.n_rbp;
  return p.use_perm ? (int)p.rbp_perm[r] : r;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kWgradThreads, 1)
wgrad_kernel(const __grid_constant__ TensorMaps maps, const __grid_constant__ WgradParams p, const TapTable taps) {
  using C = WCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_ptiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_items = p.splits * p.n_rbp * p.n_tiles;
  const int per_split = (n_ptiles + p.splits - 1) / p.splits;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kWgradThreads - 64);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    PipeState ps;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int nt = item % p.n_tiles;
      const int rbp = wgrad_rbp(p, item);
      const int split = item / (p.n_tiles * p.n_rbp);
      const int rb0 = 2 * rbp, rb1 = (2 * rbp + 1 < p.n_rb) ? 2 * rbp + 1 : 2 * rbp;
      const int tap0 = rb0 / p.kchunks_in, c0 = (rb0 % p.kchunks_in) * 64;
      const int tap1 = rb1 / p.kchunks_in, c1 = (rb1 % p.kchunks_in) * 64;
      const int dy0 = taps.dy[tap0], dx0 = taps.dx[tap0], dy1 = taps.dy[tap1], dx1 = taps.dx[tap1];
      const int m0 = taps.map[tap0], m1 = taps.map[tap1];
      const int pt0 = split * per_split;
      const int pt1 = min(pt0 + per_split, n_ptiles);
      int x0 = (pt0 % p.tiles_w) * p.bw;
      int y0 = ((pt0 / p.tiles_w) % p.tiles_h) * p.bh;
      int n0 = (pt0 / (p.tiles_w * p.tiles_h)) * p.bn;
      int nact = p.use_perm ? (int)p.rbp_nact[(item / p.n_tiles) % p.n_rbp] : wgrad_count_active(p, pt0, pt1, dx0, dy0, dx1, dy1);
      const bool all = (nact == 0) || !p.skip_oob;
      if (nact == 0) nact = pt1 - pt0;
      int done = 0;
      const bool leader = elect_one();
      for (int pt = pt0; pt < pt1; ++pt) {
        if (all || wgrad_box_active(p, x0, y0, dx0, dy0, dx1, dy1)) {
          const int bi = done % C::kPB;
          if (bi == 0) {
            mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
            if (leader) mbar_arrive_expect_tx(&full_bar[ps.stage], (uint32_t)(min(C::kPB, nact - done) * C::kBoxBytes));
          }
          if (leader) {
            uint8_t* sa = smem + ps.stage * C::kStageBytes + bi * C::kBoxBytes;
            tma_load_4d(&maps.a[m0], &full_bar[ps.stage], sa, c0, x0 + dx0, y0 + dy0, n0);
            tma_load_4d(&maps.a[m1], &full_bar[ps.stage], sa + 8192, c1, x0 + dx1, y0 + dy1, n0);
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_4d(&maps.b, &full_bar[ps.stage], sa + kABytes + j * 8192, nt * BLOCK_N + j * 64, x0, y0, n0);
          }
          ++done;
          if (bi == C::kPB - 1 || done == nact) {
            __syncwarp();
            ps.advance<C::kStages>();
          }
        }
        // next pixel box (w fastest, then h, then n) without divisions
        x0 += p.bw;
        if (x0 >= p.tiles_w * p.bw) {
          x0 = 0;
          y0 += p.bh;
          if (y0 >= p.tiles_h * p.bh) { y0 = 0; n0 += p.bn; }
        }
      }
    }
  } else if (warp == 1) {
    PipeState ps;
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr uint32_t idesc = make_idesc(kBlockM, BLOCK_N, 1, 1);
    const uint32_t smem_lo = smem_u32(smem) >> 4;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int split = item / (p.n_tiles * p.n_rbp);
      const int pt0 = split * per_split;
      const int pt1 = min(pt0 + per_split, n_ptiles);
      if (pt1 - pt0 <= 0) continue;
      int nboxes = pt1 - pt0;
      if (p.skip_oob) {
        const int rbp = wgrad_rbp(p, item);
        const int rb0 = 2 * rbp, rb1 = (2 * rbp + 1 < p.n_rb) ? 2 * rbp + 1 : 2 * rbp;
        const int tap0 = rb0 / p.kchunks_in, tap1 = rb1 / p.kchunks_in;
        const int na = p.use_perm ? (int)p.rbp_nact[(item / p.n_tiles) % p.n_rbp]
                                  : wgrad_count_active(p, pt0, pt1, taps.dx[tap0], taps.dy[tap0], taps.dx[tap1], taps.dy[tap1]);
        if (na > 0) nboxes = na;
      }
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_addr = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int b0 = 0; b0 < nboxes; b0 += C::kPB) {
        const int nb = min(C::kPB, nboxes - b0);
        mbar_wait(&full_bar[ps.stage], ps.phase);
        tc_fence_after();
        if (elect_one()) {
          for (int bi = 0; bi < nb; ++bi) {
            const uint32_t a_lo = smem_lo + (uint32_t)(ps.stage * C::kStageBytes + bi * C::kBoxBytes) / 16u;
            const uint32_t b_lo = a_lo + (uint32_t)(kABytes >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)   // 16 pixels (2048 B = 128 units) per MMA
              umma_f16(d_addr, kDescMNMajor | (uint64_t)(a_lo + 128 * k), kDescMNMajor | (uint64_t)(b_lo + 128 * k),
                       idesc, (b0 | bi | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[ps.stage]);
          if (b0 + C::kPB >= nboxes) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        ps.advance<C::kStages>();
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int nt = item % p.n_tiles;
      const int rbp = wgrad_rbp(p, item);
      const int split = item / (p.n_tiles * p.n_rbp);
      const int pt0 = split * per_split;
      const int pt1 = min(pt0 + per_split, n_ptiles);
      if (pt1 - pt0 <= 0) continue;
      const int rb = 2 * rbp + (row >> 6);
      const bool valid = rb < p.n_rb;
      const int tap = rb / p.kchunks_in;
      const int ci = (rb % p.kchunks_in) * 64 + (row & 63);
      float* dst = p.dw + (int64_t)split * p.part_stride + (int64_t)tap * p.dw_tap_stride + (int64_t)ci * p.dw_row_stride +
                   (int64_t)(nt * BLOCK_N) * p.dw_col_stride;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
      for (int c0 = half * 32; c0 < BLOCK_N; c0 += 64) {
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
        if (valid) {
          if (p.wide_store) {
            // 32-byte stores: every lane writes whole sectors of its own dW row (rows are Cout * 4 bytes apart, so a
            // warp store touches 32 lines either way; this halves the requests)
#pragma unroll
            for (int i = 0; i < 4; ++i)
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + c0 + 8 * i), "r"(r[8 * i]),
                           "r"(r[8 * i + 1]), "r"(r[8 * i + 2]), "r"(r[8 * i + 3]), "r"(r[8 * i + 4]), "r"(r[8 * i + 5]),
                           "r"(r[8 * i + 6]), "r"(r[8 * i + 7])
                           : "memory");
          } else {
            float4* d4 = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              d4[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                  __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<C::kTmemCols>(tmem_base);
}

// wgrad_pair_kernel: the (dx, dy) of the four row blocks of a CTA pair (row-block pairs 2 rp and 2 rp + 1) and which pixel boxes
// the pair walks: those at least one of the four shifted x boxes of which touches the image (all of them without skip_oob,
// or when none does: the item still has to produce its zeros)
struct PairTaps {
  int dx[4], dy[4];
};
__device__ __forceinline__ PairTaps pair_taps(const WgradParams& p, const TapTable& taps, int rp) {
  PairTaps t;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int rb = 4 * rp + i;
    if (rb >= p.n_rb) rb = p.n_rb - 1;
    const int tap = rb / p.kchunks_in;
    t.dx[i] = taps.dx[tap];
    t.dy[i] = taps.dy[tap];
  }
  return t;
}
__device__ __forceinline__ bool pair_box_active(const WgradParams& p, const PairTaps& t, int x0, int y0) {
  bool any = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int xa = x0 + t.dx[i], ya = y0 + t.dy[i];
    any |= !(ya + p.bh <= 0 || ya >= p.H || xa + p.bw <= 0 || xa >= p.W);
  }
  return any;
}
// active boxes among pixel tiles [pt0, pt1), counted by the whole warp
__device__ __forceinline__ int pair_count_active(const WgradParams& p, const PairTaps& t, int pt0, int pt1, int lane) {
  if (!p.skip_oob) return pt1 - pt0;
  int n = 0;
  for (int pt = pt0 + lane; pt < pt1; pt += 32) {
    const int x0 = (pt % p.tiles_w) * p.bw, y0 = ((pt / p.tiles_w) % p.tiles_h) * p.bh;
    n += pair_box_active(p, t, x0, y0) ? 1 : 0;
  }
  return __reduce_add_sync(0xffffffffu, n);
}

// wgrad_pair_kernel: the 256-column weight-gradient GEMM as CTA pairs (cta_group::2, M = 256).  CTAs 2q / 2q+1 of a cluster take
// two CONSECUTIVE row-block pairs (different taps / input-channel chunks: 2 x 128 dW rows) of the same channel tile and pixel
// split; both need the same dy boxes as their B operand, so each CTA stages its own two 8 KB x boxes and HALF of the dy box
// (128 of the 256 columns): a stage is 32 KB instead of 48 (six stages), and per SM the operand reads from shared memory
// fall from 96 to 64 B/clk.  Barrier protocol as in igemm_body<.., PAIR>.  Launched only when every pixel box contributes
// to every item (!skip_oob: the large-map layers) and n_rbp is even; each CTA's epilogue stores its own 128 rows.
constexpr int kWPairStageBytes = kABytes + 2 * 8192;                 // A (2 x 8 KB) + half of B (2 x 8 KB)
constexpr int kWPairStages = kSmemBudget / kWPairStageBytes;         // 6
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWgradThreads, 1)
wgrad_pair_kernel(const __grid_constant__ TensorMaps maps, const __grid_constant__ WgradParams p, const TapTable taps) {
  constexpr int BLOCK_N = 256;
  using C = WCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kWPairStages * kWPairStageBytes);
  uint64_t* empty_bar = full_bar + kWPairStages;
  uint64_t* tfull_bar = empty_bar + kWPairStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int n_ptiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int half_rbp = p.n_rbp >> 1;
  const int total_units = p.splits * half_rbp * p.n_tiles;
  const int per_split = (n_ptiles + p.splits - 1) / p.splits;
  const int ustride = (int)(gridDim.x >> 1);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    for (int i = 0; i < kWPairStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * (kWgradThreads - 64));
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair<C::kTmemCols>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    PipeState ps;
    for (int u = (int)(blockIdx.x >> 1); u < total_units; u += ustride) {
      const int nt = u % p.n_tiles;
      const int rbp = 2 * ((u / p.n_tiles) % half_rbp) + (int)crank;
      const int split = u / (p.n_tiles * half_rbp);
      const int rb0 = 2 * rbp, rb1 = (2 * rbp + 1 < p.n_rb) ? 2 * rbp + 1 : 2 * rbp;
      const int tap0 = rb0 / p.kchunks_in, c0 = (rb0 % p.kchunks_in) * 64;
      const int tap1 = rb1 / p.kchunks_in, c1 = (rb1 % p.kchunks_in) * 64;
      const int dy0 = taps.dy[tap0], dx0 = taps.dx[tap0], dy1 = taps.dy[tap1], dx1 = taps.dx[tap1];
      const int m0 = taps.map[tap0], m1 = taps.map[tap1];
      const int pt0 = split * per_split;
      const int pt1 = min(pt0 + per_split, n_ptiles);
      int x0 = (pt0 % p.tiles_w) * p.bw;
      int y0 = ((pt0 / p.tiles_w) % p.tiles_h) * p.bh;
      int n0 = (pt0 / (p.tiles_w * p.tiles_h)) * p.bn;
      const PairTaps pt4 = pair_taps(p, taps, (u / p.n_tiles) % half_rbp);
      const bool all = !p.skip_oob || pair_count_active(p, pt4, pt0, pt1, lane) == 0;
      for (int pt = pt0; pt < pt1; ++pt) {
        if (all || pair_box_active(p, pt4, x0, y0)) {
        mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + ps.stage * kWPairStageBytes;
          const uint32_t lead = mapa_u32(smem_u32(&full_bar[ps.stage]), 0);
          if (crank == 0) mbar_arrive_expect_tx(&full_bar[ps.stage], 2u * (uint32_t)kWPairStageBytes);
          tma_load_4d_pair(&maps.a[m0], lead, sa, c0, x0 + dx0, y0 + dy0, n0);
          tma_load_4d_pair(&maps.a[m1], lead, sa + 8192, c1, x0 + dx1, y0 + dy1, n0);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_4d_pair(&maps.b, lead, sa + kABytes + j * 8192, nt * BLOCK_N + (int)crank * 128 + j * 64, x0, y0, n0);
        }
        __syncwarp();
        ps.advance<kWPairStages>();
        }
        x0 += p.bw;
        if (x0 >= p.tiles_w * p.bw) {
          x0 = 0;
          y0 += p.bh;
          if (y0 >= p.tiles_h * p.bh) { y0 = 0; n0 += p.bn; }
        }
      }
    }
  } else if (warp == 1 && crank == 0) {
    PipeState ps;
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr uint32_t idesc = make_idesc(2 * kBlockM, BLOCK_N, 1, 1);
    const uint32_t smem_lo = smem_u32(smem) >> 4;
    for (int u = (int)(blockIdx.x >> 1); u < total_units; u += ustride) {
      const int split = u / (p.n_tiles * half_rbp);
      const int pt0 = split * per_split;
      const int pt1 = min(pt0 + per_split, n_ptiles);
      if (pt1 - pt0 <= 0) continue;
      int nboxes = pt1 - pt0;
      if (p.skip_oob) {
        const int na = pair_count_active(p, pair_taps(p, taps, (u / p.n_tiles) % half_rbp), pt0, pt1, lane);
        if (na > 0) nboxes = na;
      }
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_addr = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int b0 = 0; b0 < nboxes; ++b0) {
        mbar_wait(&full_bar[ps.stage], ps.phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = smem_lo + (uint32_t)(ps.stage * kWPairStageBytes) / 16u;
          const uint32_t b_lo = a_lo + (uint32_t)(kABytes >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)   // 16 pixels (2048 B = 128 units) per MMA
            umma_f16_pair(d_addr, kDescMNMajor | (uint64_t)(a_lo + 128 * k), kDescMNMajor | (uint64_t)(b_lo + 128 * k), idesc,
                          (b0 | k) != 0 ? 1u : 0u);
          umma_commit_pair(&empty_bar[ps.stage]);
          if (b0 + 1 >= nboxes) umma_commit_pair(&tfull_bar[acc]);
        }
        __syncwarp();
        ps.advance<kWPairStages>();
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 2) {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t lead_tempty0 = mapa_u32(smem_u32(&tempty_bar[0]), 0), lead_tempty1 = mapa_u32(smem_u32(&tempty_bar[1]), 0);
    for (int u = (int)(blockIdx.x >> 1); u < total_units; u += ustride) {
      const int nt = u % p.n_tiles;
      const int rbp = 2 * ((u / p.n_tiles) % half_rbp) + (int)crank;
      const int split = u / (p.n_tiles * half_rbp);
      const int pt0 = split * per_split;
      const int pt1 = min(pt0 + per_split, n_ptiles);
      if (pt1 - pt0 <= 0) continue;
      const int rb = 2 * rbp + (row >> 6);
      const bool valid = rb < p.n_rb;
      const int tap = rb / p.kchunks_in;
      const int ci = (rb % p.kchunks_in) * 64 + (row & 63);
      float* dst = p.dw + (int64_t)split * p.part_stride + (int64_t)tap * p.dw_tap_stride + (int64_t)ci * p.dw_row_stride +
                   (int64_t)(nt * BLOCK_N) * p.dw_col_stride;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
      for (int c0 = half * 32; c0 < BLOCK_N; c0 += 64) {
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
        if (valid) {
          if (p.wide_store) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + c0 + 8 * i), "r"(r[8 * i]),
                           "r"(r[8 * i + 1]), "r"(r[8 * i + 2]), "r"(r[8 * i + 3]), "r"(r[8 * i + 4]), "r"(r[8 * i + 5]),
                           "r"(r[8 * i + 6]), "r"(r[8 * i + 7])
                           : "memory");
          } else {
            float4* d4 = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              d4[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                  __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(acc ? lead_tempty1 : lead_tempty0);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 1) tmem_dealloc_pair<C::kTmemCols>(tmem_base);
}

// Finishes a split-K igemm: out = epilogue(sum of the ksplits slices of ws) over [rows][C], 8 channels per thread.
__global__ void __launch_bounds__(256) epilogue_finish_kernel(const float* __restrict__ ws, int ksplits, int64_t slice,
                                                              const float* __restrict__ bias,
                                                              const bf16* __restrict__ residual,
                                                              const bf16* __restrict__ mask, void* __restrict__ out,
                                                              int out_f32, int relu, float scale, int64_t rows, int C,
                                                              const uint32_t* __restrict__ mask_bits, int ldo) {
  const int C8 = C >> 3;
  const int64_t total = rows * C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    const int64_t base = (i / C8) * C + c;
    const int64_t obase = (i / C8) * ldo + c;      // (`out` may be a channel slice of a wider tensor: pixels ldo channels apart)
    float v[8];
    const float4 a = *reinterpret_cast<const float4*>(ws + base), b = *reinterpret_cast<const float4*>(ws + base + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    for (int s = 1; s < ksplits; ++s) {
      const float4 a2 = *reinterpret_cast<const float4*>(ws + s * slice + base);
      const float4 b2 = *reinterpret_cast<const float4*>(ws + s * slice + base + 4);
      v[0] += a2.x; v[1] += a2.y; v[2] += a2.z; v[3] += a2.w; v[4] += b2.x; v[5] += b2.y; v[6] += b2.z; v[7] += b2.w;
    }
    if (bias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += __ldg(bias + c + j);
    }
    if (residual) {
      const uint4 u = *reinterpret_cast<const uint4*>(residual + base);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&u.x)[j]);
        v[2 * j] += f.x; v[2 * j + 1] += f.y;
      }
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (mask) {
      const uint4 u = *reinterpret_cast<const uint4*>(mask + base);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&u.x)[j]);
        if (!(f.x > 0.f)) v[2 * j] = 0.f;
        if (!(f.y > 0.f)) v[2 * j + 1] = 0.f;
      }
    }
    if (mask_bits) {
      const uint32_t mb = __ldg(mask_bits + (base >> 5)) >> (base & 31);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (!((mb >> j) & 1u)) v[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= scale;
    if (out_f32) {
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + obase);
      o[0] = make_float4(v[0], v[1], v[2], v[3]);
      o[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(out) + obase) =
          make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
  }
}

// out[nt * BN + c] = sum over the partial rows of channel tile nt (fixed order): grid (BN / 32, n_tiles), 32 x 8 threads
__global__ void __launch_bounds__(256) colsum_reduce_kernel(const float* __restrict__ part, float* __restrict__ out,
                                                            int rows_per_nt, int BN) {
  __shared__ float sh[8][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const float* base = part + (int64_t)blockIdx.y * rows_per_nt * BN + c;
  // eight independent loads in flight per thread: with two, this 8-16 block kernel was pure L2 latency (~15 us)
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  int r = rl;
  for (; r + 56 < rows_per_nt; r += 64) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += base[(int64_t)(r + 8 * j) * BN];
  }
  for (; r < rows_per_nt; r += 8) a[0] += base[(int64_t)r * BN];
  sh[rl][cl] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  __syncthreads();
  if (rl == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][cl];
    out[blockIdx.y * BN + c] = t;
  }
}

// column sums of a finished bf16 [rows][C] tensor (the path taken when the fused epilogue sums are not available:
// split-K, pixel-tile-fastest order): per-block partial rows + colsum_reduce_kernel
__global__ void __launch_bounds__(256) colsum_rows_kernel(const uint4* __restrict__ x, float* __restrict__ part, int64_t rows,
                                                          int C8) {
  __shared__ float sh[256][9];
  const int cpb = C8 < 256 ? C8 : 256;
  const int R = 256 / cpb;
  const int cg = blockIdx.y * cpb + (threadIdx.x % cpb);
  const int rl = threadIdx.x / cpb;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (cg < C8 && rl < R) {
    for (int64_t r = (int64_t)blockIdx.x * R + rl; r < rows; r += (int64_t)gridDim.x * R) {
      const uint4 u = __ldg(x + r * C8 + cg);
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
    float* o = part + (int64_t)blockIdx.x * (C8 * 8) + cg * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = acc[j];
  }
}

// The same for the lockstep tap-split schedule: the number of partial slices differs per (pixel tile, channel tile).
struct TsFinish {
  int G, total, tiles_w, tiles_n, bw, bh, bn, H, W;
  int prefix[kMaxCols + 1];
};

__global__ void __launch_bounds__(256) epilogue_finish_ts_kernel(const float* __restrict__ ws, const TsFinish tf, int64_t slice,
                                                                 const float* __restrict__ bias,
                                                                 const bf16* __restrict__ residual,
                                                                 const bf16* __restrict__ mask, void* __restrict__ out,
                                                                 int out_f32, int relu, float scale, int64_t rows, int C,
                                                                 const uint32_t* __restrict__ mask_bits, int ldo) {
  const int C8 = C >> 3;
  const int64_t total = rows * C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    const int64_t row = i / C8;
    const int64_t base = row * C + c;
    const int64_t obase = row * ldo + c;           // (`out` may be a channel slice of a wider tensor)
    const int x = (int)(row % tf.W), y = (int)((row / tf.W) % tf.H), n = (int)(row / ((int64_t)tf.W * tf.H));
    const int tau = ((y / tf.bh) * tf.tiles_w + x / tf.bw) * tf.tiles_n + n / tf.bn;
    const int cs = tf.prefix[tau], ce = tf.prefix[tau + 1];
    const int nparts = ts_owner(ce - 1, tf.G, tf.total) - ts_owner(cs, tf.G, tf.total) + 1;
    float v[8];
    const float4 a = *reinterpret_cast<const float4*>(ws + base), b = *reinterpret_cast<const float4*>(ws + base + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    for (int s = 1; s < nparts; ++s) {
      const float4 a2 = *reinterpret_cast<const float4*>(ws + s * slice + base);
      const float4 b2 = *reinterpret_cast<const float4*>(ws + s * slice + base + 4);
      v[0] += a2.x; v[1] += a2.y; v[2] += a2.z; v[3] += a2.w; v[4] += b2.x; v[5] += b2.y; v[6] += b2.z; v[7] += b2.w;
    }
    if (bias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += __ldg(bias + c + j);
    }
    if (residual) {
      const uint4 u = *reinterpret_cast<const uint4*>(residual + base);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&u.x)[j]);
        v[2 * j] += f.x; v[2 * j + 1] += f.y;
      }
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (mask) {
      const uint4 u = *reinterpret_cast<const uint4*>(mask + base);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&u.x)[j]);
        if (!(f.x > 0.f)) v[2 * j] = 0.f;
        if (!(f.y > 0.f)) v[2 * j + 1] = 0.f;
      }
    }
    if (mask_bits) {
      const uint32_t mb = __ldg(mask_bits + (base >> 5)) >> (base & 31);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (!((mb >> j) & 1u)) v[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= scale;
    if (out_f32) {
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + obase);
      o[0] = make_float4(v[0], v[1], v[2], v[3]);
      o[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(out) + obase) =
          make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
  }
}

// Finishes the split tiles [first, first + ntiles) of a hybrid launch: the same arithmetic as epilogue_finish_kernel, over
// the pixel boxes of those tiles only (8 channels per thread).
__global__ void __launch_bounds__(256) epilogue_finish_tiles_kernel(const float* __restrict__ ws, const IgemmParams p, int first,
                                                                    int ntiles, int block_n) {
  const int C8 = block_n >> 3;
  const int64_t total = (int64_t)ntiles * p.rows * C8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    const int row = (int)((i / C8) % p.rows);
    const TileCoord t = decode_tile(p, first + (int)(i / ((int64_t)C8 * p.rows)));
    const int qx = t.x0 + row % p.bw, qy = t.y0 + (row / p.bw) % p.bh, n = t.n0 + row / (p.bw * p.bh);
    if (qx >= p.W || qy >= p.H || n >= p.N) continue;
    const int ch = t.nt * block_n + c;
    const int64_t base = (((int64_t)n * p.out_H + qy) * p.out_W + qx) * p.ldo + ch;
    float v[8];
    const float4 a = *reinterpret_cast<const float4*>(ws + base), b = *reinterpret_cast<const float4*>(ws + base + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    for (int s = 1; s < p.ksplits; ++s) {
      const float4 a2 = *reinterpret_cast<const float4*>(ws + s * p.ws_slice + base);
      const float4 b2 = *reinterpret_cast<const float4*>(ws + s * p.ws_slice + base + 4);
      v[0] += a2.x; v[1] += a2.y; v[2] += a2.z; v[3] += a2.w; v[4] += b2.x; v[5] += b2.y; v[6] += b2.z; v[7] += b2.w;
    }
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += __ldg(p.bias + ch + j);
    }
    if (p.residual) {
      const uint4 u = *reinterpret_cast<const uint4*>(p.residual + base);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&u.x)[j]);
        v[2 * j] += f.x; v[2 * j + 1] += f.y;
      }
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (p.mask) {
      const uint4 u = *reinterpret_cast<const uint4*>(p.mask + base);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2((&u.x)[j]);
        if (!(f.x > 0.f)) v[2 * j] = 0.f;
        if (!(f.y > 0.f)) v[2 * j + 1] = 0.f;
      }
    }
    if (p.mask_bits) {
      const uint32_t mb = __ldg(p.mask_bits + (base >> 5)) >> (base & 31);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (!((mb >> j) & 1u)) v[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= p.scale;
    if (p.out_f32) {
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + base);
      o[0] = make_float4(v[0], v[1], v[2], v[3]);
      o[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + base) =
          make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
  }
}

// grow-only device scratch owned by the context (first use allocates; never on the steady-state path)
int ensure_workspace(segk_ctx* ctx, size_t bytes) { return segk_grow(ctx, &ctx->ws, &ctx->ws_bytes, bytes, "split-K workspace"); }

// Scratch for the fused column sums (ctx->ws6, main stream) and the two ways of finishing them.
int colsum_scratch(segk_ctx* ctx, int grid, int block_n, float** out) {
  const int rc = segk_grow(ctx, &ctx->ws6, &ctx->ws6_bytes, sizeof(float) * (size_t)(ctx->sm_count * 8) * 4096, "column-sum partials");
  if (rc) return rc;
  (void)grid; (void)block_n;
  *out = (float*)ctx->ws6;
  return SEGK_OK;
}

int colsum_finish(segk_ctx* ctx, float* colsum_out, int grid, int n_tiles, int block_n, cudaStream_t st) {
  colsum_reduce_kernel<<<dim3(block_n / 32, n_tiles), 256, 0, st>>>((const float*)ctx->ws6, colsum_out, (grid / n_tiles) * 4, block_n);
  SEGK_LAUNCHED(ctx, "column-sum reduce");
  return SEGK_OK;
}

int colsum_fallback(segk_ctx* ctx, const void* y, int64_t rows, int C, float* colsum_out, cudaStream_t st) {
  SEGK_REQUIRE(ctx, C % 32 == 0 && (C / 8 <= 256 ? 256 % (C / 8) == 0 : (C / 8) % 256 == 0),
               "dgrad column sums: unsupported channel count %d", C);
  float* part = nullptr;
  int rc = colsum_scratch(ctx, 0, 0, &part);
  if (rc) return rc;
  const int C8 = C / 8, cpb = C8 < 256 ? C8 : 256, R = 256 / cpb, gy = C8 / cpb;
  int64_t gx = ceil_div64(rows, (int64_t)R * 4);
  const int64_t cap = ceil_div64((int64_t)ctx->sm_count * 4, gy);
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  colsum_rows_kernel<<<dim3((unsigned)gx, gy), 256, 0, st>>>((const uint4*)y, part, rows, C8);
  SEGK_LAUNCHED(ctx, "column sums");
  colsum_reduce_kernel<<<dim3(C / 32, 1), 256, 0, st>>>(part, colsum_out, (int)gx, C);
  SEGK_LAUNCHED(ctx, "column-sum reduce");
  return SEGK_OK;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// Pick a pixel box bw x bh x bn with at most `max_rows` rows (exactly, when `exact`) that
// covers [N,H,W] with the fewest tiles; ties -> wider bw (longer contiguous runs for TMA).
// Number of (tile, tap) pairs that survive tile-level tap skipping for a kh x kw SAME conv when
// [H, W] is tiled by bh x bw boxes (1-D factorises: rows and columns are independent).
int64_t active_taps_1d(int n, int b, int k) {
  const int p = k / 2;
  int64_t tot = 0;
  for (int t0 = 0; t0 < n; t0 += b)
    for (int d = -p; d <= k - 1 - p; ++d)
      if (!(t0 + d + b <= 0 || t0 + d >= n)) ++tot;
  return tot;
}

// Pick a pixel box bw x bh x bn with at most `max_rows` rows (exactly, when `exact`) that covers
// [N,H,W] with the least work = sum over tiles of active taps (kh = kw = 1: the tile count);
// ties -> wider bw (longer contiguous runs for TMA).
Box choose_box(int N, int H, int W, int max_rows, bool exact, int lim_h = 0, int lim_w = 0, int kh = 1, int kw = 1,
               bool even = false) {
  Box best{0, 0, 0, 0, 0};
  int64_t best_cost = -1;
  if (lim_h <= 0) lim_h = H;
  if (lim_w <= 0) lim_w = W;
  for (int bw = 1; bw <= lim_w && bw <= max_rows && bw <= 256; ++bw) {
    const int64_t ax = active_taps_1d(W, bw, kw);
    if (even && (bw & 1)) continue;                 // fused 2x2 pool: whole windows per box
    for (int bh = 1; bh <= lim_h && bw * bh <= max_rows && bh <= 256; ++bh) {
      if (even && (bh & 1)) continue;
      int bn = max_rows / (bw * bh);
      // exact boxes may run past the batch (TMA zero-fills the out-of-bounds images)
      if (!exact && bn > N) bn = N;
      if (bn > 256) continue;
      const int rows = bw * bh * bn;
      if (exact && rows != max_rows) continue;
      const int tiles = ceil_div(W, bw) * ceil_div(H, bh) * ceil_div(N, bn);
      const int64_t work = ax * active_taps_1d(H, bh, kh) * ceil_div(N, bn);
      const int64_t cost = work * 1024 - bw;
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best = Box{bw, bh, bn, rows, tiles};
      }
    }
  }
  return best;
}

int encode_act_map(segk_ctx* ctx, CUtensorMap* m, const void* base, int N, int H, int W, int C, int64_t pix_stride_w,
                   int64_t pix_stride_h, int64_t pix_stride_n, int bw, int bh, int bn) {
  // dims innermost first: (C, W, H, N); strides in bytes for dims 1..3
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)pix_stride_w * 2, (cuuint64_t)pix_stride_h * 2, (cuuint64_t)pix_stride_n * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = ctx->encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return segk_fail(ctx, SEGK_ECUDA, "cuTensorMapEncodeTiled(act %dx%dx%dx%d box %dx%dx%d) failed: %d", N, H, W, C,
                     bw, bh, bn, (int)r);
  return SEGK_OK;
}

// NHWC tensor whose pixels sit `ld` channels apart (ld = C: dense; ld > C: a channel slice of a wider tensor -- the
// zero-copy Concat views of segk_set_pitch)
int encode_nhwc_map(segk_ctx* ctx, CUtensorMap* m, const void* base, int N, int H, int W, int C, int ld, int bw, int bh, int bn) {
  return encode_act_map(ctx, m, base, N, H, W, C, ld, (int64_t)W * ld, (int64_t)H * W * ld, bw, bh, bn);
}

int encode_weight_map(segk_ctx* ctx, CUtensorMap* m, const void* base, int K, int Nrows, int T, int block_n) {
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)Nrows, (cuuint64_t)T};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * Nrows * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)block_n, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = ctx->encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return segk_fail(ctx, SEGK_ECUDA, "cuTensorMapEncodeTiled(weights K=%d N=%d T=%d) failed: %d", K, Nrows, T, (int)r);
  return SEGK_OK;
}

// Kernel-layout weights are BLOCKED: [T taps][KC = ceil(K/64) chunks][Nrows][64 k] bf16, i.e. the operand tile of
// one (tap, k-chunk) -- Nrows x 128 bytes -- is one contiguous run in memory (an N = 256 tile: 32 KB).  With the
// plain [T][Nrows][K] layout a tile is 256 pieces of 128 B at a stride of 2K bytes; for conv6's dgrad (K = 4096: 8 KB
// stride, 205 MB of weights) ncu showed the tiles evicted from L2 before the CTAs sharing them arrived (L2 hit 49 %,
// 2.2 GB of DRAM reads for 0.68 GB requested).  4-D map (64, Nrows, KC, T), box (64, box_rows, 1, box_taps).
int encode_weight_map_blocked(segk_ctx* ctx, CUtensorMap* m, const void* base, int K, int Nrows, int T, int box_rows,
                              int box_taps = 1) {
  const int KC = ceil_div(K, 64);
  cuuint64_t dims[4] = {64, (cuuint64_t)Nrows, (cuuint64_t)KC, (cuuint64_t)T};
  cuuint64_t strides[3] = {128, (cuuint64_t)Nrows * 128, (cuuint64_t)KC * Nrows * 128};
  cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, (cuuint32_t)box_taps};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = ctx->encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return segk_fail(ctx, SEGK_ECUDA, "cuTensorMapEncodeTiled(blocked weights K=%d N=%d T=%d) failed: %d", K, Nrows, T, (int)r);
  return SEGK_OK;
}

// tuning overrides for sweeps (tools/sweep_tiles.py), read once per context in segk_tc_init;
// 0 = use the heuristics
int env_int(const char* name, int dflt = 0) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

int pick_block_n(const segk_ctx* ctx, int Cout) {
  const int f = ctx->force_bn;
  if ((f == 64 || f == 128 || f == 256) && Cout % f == 0) return f;
  if (Cout % 256 == 0) return 256;
  if (Cout % 128 == 0) return 128;
  return 64;
}

template <int BLOCK_N>
int launch_igemm_t(segk_ctx* ctx, const TensorMaps& maps, const IgemmParams& p, const TapTable& taps, int grid,
                   cudaStream_t st) {
  using C = Cfg<BLOCK_N>;
  igemm_kernel<BLOCK_N><<<grid, kThreads, C::kSmemBytes, st>>>(maps, p, taps);
  SEGK_LAUNCHED(ctx, "igemm");
  return SEGK_OK;
}

// `pair_ok`: maps.b2 is encoded and the launch may run as CTA pairs (igemm_pair_kernel); *grid_out = the grid used
int launch_igemm(segk_ctx* ctx, int block_n, const TensorMaps& maps, const IgemmParams& p, const TapTable& taps,
                 cudaStream_t st, int fixed_grid = 0, bool pair_ok = false, int* grid_out = nullptr) {
  const int tiles_all = p.phases * p.tiles_n * p.tiles_h * p.tiles_w * p.n_tiles;
  if (pair_ok && ctx->pair && block_n == 256 && fixed_grid == 0 && !p.ts && !p.hyb && p.ksplits == 1 && p.phases == 1 && !p.pack_s) {
    const int m_tiles = p.tiles_n * p.tiles_h * p.tiles_w;
    const int units = ((m_tiles + 1) / 2) * p.n_tiles;
    int gp = units < ctx->sm_count / 2 ? units : ctx->sm_count / 2;
    if (p.colsum) gp = (gp / p.n_tiles) * p.n_tiles;
    // A pair walks the UNION of its two tiles' active taps (one k-step sequence for both), and an odd tile count leaves one
    // CTA of the last pair idle: pair only when that costs under 4 % more k-steps than single-CTA tiles.  Large maps lose
    // nothing (tap sets differ at the image border only); conv6 (7x7 taps on a 5x18 map, every tile a different tap set)
    // would walk 13 % more (ncu r2: 433 k vs 382 k active cycles) and stays single.
    bool worth = gp > 0 && m_tiles >= 16;
    if (worth && p.ntaps > 1) {
      const int cols = p.tiles_w * p.tiles_h;
      std::vector<uint64_t> cm((size_t)cols);
      for (int c = 0; c < cols; ++c) {
        const int x0 = (c % p.tiles_w) * p.bw, y0 = (c / p.tiles_w) * p.bh;
        uint64_t m = 0;
        for (int i = 0; i < p.ntaps; ++i) {
          const int ya = y0 + taps.dy[i], xa = x0 + taps.dx[i];
          if (!(ya + p.bh <= 0 || ya >= p.in_H || xa + p.bw <= 0 || xa >= p.in_W)) m |= 1ull << i;
        }
        if (m == 0) m = p.ntaps >= 64 ? ~0ull : ((1ull << p.ntaps) - 1);      // (mirrors tap_mask())
        cm[c] = m;
      }
      int64_t single = 0, paired = 0;
      for (int r = 0; r < m_tiles; r += 2) {
        const uint64_t a = cm[r % cols], b = cm[(r + 1) % cols];
        single += __builtin_popcountll(a) + (r + 1 < m_tiles ? __builtin_popcountll(b) : 0);
        paired += 2 * __builtin_popcountll(a | b);
      }
      worth = paired * 100 <= single * 104;
    } else if (worth) {
      worth = (m_tiles & 1) == 0 || m_tiles >= 50;
    }
    if (worth) {
      IgemmParams q = p;
      q.pair = 1;
      igemm_pair_kernel<<<2 * gp, kThreads, Cfg<256>::kSmemBytes, st>>>(maps, q, taps);
      SEGK_LAUNCHED(ctx, "igemm (CTA pairs)");
      if (grid_out) *grid_out = 2 * gp;
      return SEGK_OK;
    }
  }
  const int total = p.hyb ? p.hyb_full + (tiles_all - p.hyb_full) * p.ksplits : tiles_all * p.ksplits;
  int grid = fixed_grid > 0 ? fixed_grid : (total < ctx->sm_count ? total : ctx->sm_count);
  if (p.colsum) grid = (grid / p.n_tiles) * p.n_tiles;      // every tile of a CTA has the same channel tile
  if (grid_out) *grid_out = grid;
  switch (block_n) {
    case 256: return launch_igemm_t<256>(ctx, maps, p, taps, grid, st);
    case 128: return launch_igemm_t<128>(ctx, maps, p, taps, grid, st);
    default: return launch_igemm_t<64>(ctx, maps, p, taps, grid, st);
  }
}

template <int BLOCK_N>
int launch_wgrad_t(segk_ctx* ctx, const TensorMaps& maps, const WgradParams& p, const TapTable& taps, int grid,
                   cudaStream_t st) {
  using C = WCfg<BLOCK_N>;
  WgradParams q = p;
  q.wide_store = (((uintptr_t)p.dw & 31) == 0 && p.dw_col_stride == 1 && p.dw_row_stride % 8 == 0 && p.dw_tap_stride % 8 == 0 &&
                  p.part_stride % 8 == 0) ? 1 : 0;
  if (BLOCK_N == 256 && ctx->pair && !q.use_perm && (q.n_rbp & 1) == 0 && q.n_rb == 2 * q.n_rbp) {
    // CTA pairs: two consecutive row-block pairs share the dy boxes (wgrad_pair_kernel)
    const int units = q.splits * (q.n_rbp / 2) * q.n_tiles;
    const int gp = units < ctx->sm_count / 2 ? units : ctx->sm_count / 2;
    wgrad_pair_kernel<<<2 * gp, kWgradThreads, WCfg<256>::kSmemBytes, st>>>(maps, q, taps);
    SEGK_LAUNCHED(ctx, "wgrad (CTA pairs)");
    return SEGK_OK;
  }
  wgrad_kernel<BLOCK_N><<<grid, kWgradThreads, C::kSmemBytes, st>>>(maps, q, taps);
  SEGK_LAUNCHED(ctx, "wgrad");
  return SEGK_OK;
}

template <int BLOCK_N>
int launch_slab_t(segk_ctx* ctx, const TensorMaps& maps, const SlabParams& p, int grid, cudaStream_t st) {
  using C = SlabCfg<BLOCK_N>;
  slab_kernel<BLOCK_N><<<grid, kThreads, C::kSmemBytes, st>>>(maps, p);
  SEGK_LAUNCHED(ctx, "slab");
  return SEGK_OK;
}

// CTA pairs for the 128-column slab tiles (slab_pair_kernel<128>; maps.b2 = weights with a 64-row box): grid of `pairs` clusters
int launch_slab_pair128(segk_ctx* ctx, const TensorMaps& maps, const SlabParams& p, int pairs, cudaStream_t st) {
  slab_pair_kernel<128><<<2 * pairs, kThreads, SlabCfg<128>::kSmemBytes, st>>>(maps, p);
  SEGK_LAUNCHED(ctx, "slab (CTA pairs)");
  return SEGK_OK;
}

// 3x3 layers on large maps whose GEMM-N is small are activation-traffic bound in the tap-wise igemm
bool slab_applicable(const segk_ctx* ctx, int N, int H, int W, int Ck, int Cn, int kh, int kw) {
  const int mode = ctx->slab_mode;   // 0 off, 1 auto, 2 whenever legal
  if (mode == 0) return false;
  const bool legal = kh == 3 && kw == 3 && H % kSlabH == 0 && W >= kSlabWV && Ck % 64 == 0 && Cn % 64 == 0;
  if (!legal) return false;
  if (mode == 2) return true;
  return Cn <= 128 && (int64_t)H * W >= 4096;
}

// fused 2x2 max-pool of a forward conv's output (segk_conv2d_fwd_pool)
struct PoolArgs {
  void* pooled;
  uint8_t* idx;
  int pool_only;
  bool fused;      // out: the launch produced pooled / idx itself
  uint32_t* bits_out = nullptr;          // 1-bit ReLU mask of the output (forward)
  const uint32_t* mask_bits = nullptr;   // 1-bit ReLU mask of the producer layer (dgrad)
  bool bits_done = false;                // out: the launch wrote bits_out itself
  int out_cols = 0;                      // narrow fp32 output: store only the first out_cols columns
};

// bits[r][c / 32] from a finished bf16 tensor [rows][C] (forward paths whose epilogue runs in a finish kernel)
__global__ void __launch_bounds__(256) relu_bits_kernel(const uint4* __restrict__ y, uint32_t* __restrict__ bits, int64_t nwords) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nwords; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t word = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 u = __ldg(y + i * 4 + q);
#pragma unroll
      for (int j = 0; j < 4; ++j) word |= bf16x2_pos_bits((&u.x)[j]) << (8 * q + 2 * j);
    }
    bits[i] = word;
  }
}

int conv_slab(segk_ctx* ctx, const char* what, const void* x, const void* wt, const float* bias, const void* residual,
              const void* mask, float scale, int relu, int out_f32, void* y, int N, int H, int W, int Ck, int Cn,
              void* stream, float* colsum_out, PoolArgs* pool = nullptr, int ldx = 0, int ldy = 0) {
  if (ldx <= 0) ldx = Ck;
  if (ldy <= 0) ldy = Cn;
  // Ck = 64: kx-fused N = 192 MMAs on resident weights (slab3_kernel), 64-channel output tiles.  Measured
  // (tools/time_n64.py, B=32 160x576): 64 -> 64 forward 228 vs 297 us, its dgrad 345 vs 372 us; with two
  // channel tiles (64 -> 128) it is a wash (116 vs 112 us), so the tap-wise slab keeps those.  slab3 = 2 forces it.
  const bool fused3 = (Ck == 64 || Ck == 128) && (Cn / 64) <= ctx->sm_count &&
                      (ctx->slab3 == 2 || (ctx->slab3 == 1 && Cn == 64));
  const int block_n = fused3 ? 64 : pick_block_n(ctx, Cn);
  TensorMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc = encode_nhwc_map(ctx, &maps.a[0], x, N, H, W, Ck, ldx, kSlabP, kSlabH + 2, 1);
  if (rc) return rc;
  maps.a[1] = maps.a[2] = maps.a[3] = maps.a[0];
  if (fused3) {
    // box (64 k, 64 channels, 1 chunk, 3 taps): the three kx taps of one ky land as one [192][64] operand
    rc = encode_weight_map_blocked(ctx, &maps.b, wt, Ck, Cn, 9, 64, 3);
    if (rc) return rc;
  } else {
    rc = encode_weight_map_blocked(ctx, &maps.b, wt, Ck, Cn, 9, block_n);
    if (rc) return rc;
  }
  // 128-column tiles as CTA pairs (slab_pair_kernel<128>): a 128 x 128 x 16 MMA reads 8 KB of operands per 64 tensor cycles =
  // the 128 B/clk a SM's shared memory delivers, a pair's M = 256 MMA halves the weight part of it (96 B/clk).  Measured
  // (profiles/r2_pair_ab.md) it does not pay: conv2_2 forward 199.7 -> 195.6 us, its dgrad 197.6 -> 203.8, conv2_1 forward
  // (one k-chunk per tile: nine MMA groups between two cross-CTA barrier round trips) 130.0 -> 162.7.  Kept behind
  // segk_set_tuning("pair", 2) with its parity test; off by default.
  const int m_tiles_all = N * (H / kSlabH) * ceil_div(W, kSlabWV);
  const bool pair = !fused3 && block_n == 128 && ctx->pair >= 2 && m_tiles_all >= 16;
  if (pair) {
    rc = encode_weight_map_blocked(ctx, &maps.b2, wt, Ck, Cn, 9, block_n / 2);
    if (rc) return rc;
  }
  SlabParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W;
  p.tiles_w = ceil_div(W, kSlabWV); p.tiles_h = H / kSlabH;
  p.n_tiles = Cn / block_n; p.kchunks = Ck / 64; p.ldo = Cn;
  p.out = y; p.out_f32 = out_f32; p.bias = bias; p.residual = (const bf16*)residual; p.mask = (const bf16*)mask;
  p.scale = scale; p.relu = relu;
  p.tma_store = (!out_f32 && (ctx->tma_store || ldy != Cn)) ? 1 : 0;
  SEGK_REQUIRE(ctx, ldy == Cn || (p.tma_store && !colsum_out && !residual && !mask && !(pool && (pool->bits_out || pool->mask_bits))),
               "%s: a pitched output needs the plain bf16 epilogue (no residual / mask / column sums)", what);
  if (p.tma_store) {
    rc = encode_nhwc_map(ctx, &maps.c, y, N, H, W, Cn, ldy, kSlabWV, kSlabH, 1);
    if (rc) return rc;
  }
  if (pool && pool->pooled && p.tma_store && !colsum_out && (W & 1) == 0) {      // 4 x 30 tiles at even origins: whole pool windows
    p.pool_out = (bf16*)pool->pooled; p.pool_idx = pool->idx; p.pool_only = pool->pool_only;
    pool->fused = true;
  }
  if (pool) {
    p.mask_bits = pool->mask_bits;
    p.out_cols = out_f32 ? pool->out_cols : 0;
    if (pool->bits_out && !out_f32) { p.bits_out = pool->bits_out; pool->bits_done = true; }
  }
  const int total = N * p.tiles_h * p.tiles_w * p.n_tiles;
  int grid = total < ctx->sm_count ? total : ctx->sm_count;
  if (colsum_out) {
    if (fused3) {
      // slab3 has no column-sum epilogue (in FCN-8s the producers of its inputs are the first layer and the pools)
      const int g3 = total < ctx->sm_count ? total : (ctx->sm_count / p.n_tiles) * p.n_tiles;
      if (Ck == 64) slab3_kernel<1><<<g3, kThreads, Slab3Cfg<1>::kSmem, (cudaStream_t)stream>>>(maps, p);
      else slab3_kernel<2><<<g3, kThreads, Slab3Cfg<2>::kSmem, (cudaStream_t)stream>>>(maps, p);
      SEGK_LAUNCHED(ctx, "slab3");
      return colsum_fallback(ctx, y, (int64_t)N * H * W, Cn, colsum_out, (cudaStream_t)stream);
    }
    rc = colsum_scratch(ctx, grid, block_n, &p.colsum);
    if (rc) return rc;
    if (pair) {
      const int units = ((m_tiles_all + 1) / 2) * p.n_tiles;
      int gp = units < ctx->sm_count / 2 ? units : ctx->sm_count / 2;
      gp = (gp / p.n_tiles) * p.n_tiles;
      if (gp > 0) {
        rc = launch_slab_pair128(ctx, maps, p, gp, (cudaStream_t)stream);
        if (rc) return rc;
        return colsum_finish(ctx, colsum_out, 2 * gp, p.n_tiles, block_n, (cudaStream_t)stream);
      }
    }
    grid = (grid / p.n_tiles) * p.n_tiles;
    switch (block_n) {
      case 256: rc = launch_slab_t<256>(ctx, maps, p, grid, (cudaStream_t)stream); break;
      case 128: rc = launch_slab_t<128>(ctx, maps, p, grid, (cudaStream_t)stream); break;
      default: rc = launch_slab_t<64>(ctx, maps, p, grid, (cudaStream_t)stream); break;
    }
    if (rc) return rc;
    return colsum_finish(ctx, colsum_out, grid, p.n_tiles, block_n, (cudaStream_t)stream);
  }
  if (fused3) {
    // every CTA keeps one channel tile: grid = a multiple of n_tiles (total is one by construction)
    const int g3 = total < ctx->sm_count ? total : (ctx->sm_count / p.n_tiles) * p.n_tiles;
    if (Ck == 64) slab3_kernel<1><<<g3, kThreads, Slab3Cfg<1>::kSmem, (cudaStream_t)stream>>>(maps, p);
    else slab3_kernel<2><<<g3, kThreads, Slab3Cfg<2>::kSmem, (cudaStream_t)stream>>>(maps, p);
    SEGK_LAUNCHED(ctx, "slab3");
    return SEGK_OK;
  }
  if (pair) {
    const int units = ((m_tiles_all + 1) / 2) * p.n_tiles;
    return launch_slab_pair128(ctx, maps, p, units < ctx->sm_count / 2 ? units : ctx->sm_count / 2, (cudaStream_t)stream);
  }
  switch (block_n) {
    case 256: return launch_slab_t<256>(ctx, maps, p, grid, (cudaStream_t)stream);
    case 128: return launch_slab_t<128>(ctx, maps, p, grid, (cudaStream_t)stream);
    default: return launch_slab_t<64>(ctx, maps, p, grid, (cudaStream_t)stream);
  }
}

// taps of a SAME conv; rate > 1: tf.nn.atrous_conv2d (utils.py:210-231) -- the taps simply sit `rate` pixels apart
void conv_taps(TapTable& t, int kh, int kw, int rate = 1) {
  memset(&t, 0, sizeof(t));
  for (int i = 0; i < kh * kw; ++i) {
    t.dy[i] = (int8_t)((i / kw - kh / 2) * rate);
    t.dx[i] = (int8_t)((i % kw - kw / 2) * rate);
    t.map[i] = 0;
  }
}

// shared body of conv fwd and dgrad: y[N,H,W,Cn] = epilogue( sum_taps x[.. + tap][Ck] * wt[tap][Cn][Ck] )
int conv_igemm(segk_ctx* ctx, const char* what, const void* x, const void* wt, const float* bias, const void* residual,
               const void* mask, float scale, int relu, int out_f32, void* y, int N, int H, int W, int Ck, int Cn,
               int kh, int kw, void* stream, float* colsum_out = nullptr, int rate = 1, PoolArgs* pool = nullptr, int ldx = 0,
               int ldy = 0) {
  SEGK_REQUIRE(ctx, x && wt && y, "%s: null pointer", what);
  // ldx / ldy: channels between consecutive pixels of x / y (0 = dense).  A pitched y leaves through the TMA store only:
  // its epilogue offsets (residual, mask, bits, finish kernels of the split schedules) all assume the dense row stride
  if (ldx <= 0) ldx = Ck;
  if (ldy <= 0) ldy = Cn;
  SEGK_REQUIRE(ctx, ldx >= Ck && ldy >= Cn && ldx % 8 == 0 && ldy % 8 == 0, "%s: bad channel pitch (%d for %d, %d for %d)", what, ldx,
               Ck, ldy, Cn);
  const bool pitched_out = ldy != Cn;
  SEGK_REQUIRE(ctx, !pitched_out || (!out_f32 && !colsum_out && !residual && !mask && !(pool && (pool->bits_out || pool->mask_bits))),
               "%s: a pitched output needs the plain bf16 epilogue (no residual / mask / bits / column sums)", what);
  SEGK_REQUIRE(ctx, N > 0 && H > 0 && W > 0, "%s: empty tensor", what);
  SEGK_REQUIRE(ctx, Ck % 64 == 0 && Cn % 64 == 0 && Ck > 0 && Cn > 0,
               "%s: tensor-core path needs channel counts that are multiples of 64 (got %d -> %d); no fallback", what,
               Ck, Cn);
  SEGK_REQUIRE(ctx, (kh & 1) && (kw & 1) && kh * kw <= kMaxTaps, "%s: odd kernel sizes up to %d taps (got %dx%d)", what,
               kMaxTaps, kh, kw);
  SEGK_REQUIRE(ctx, (((uintptr_t)x | (uintptr_t)wt | (uintptr_t)y | (uintptr_t)residual | (uintptr_t)mask) & 15) == 0,
               "%s: pointers must be 16-byte aligned", what);
  SEGK_REQUIRE(ctx, !colsum_out || !out_f32, "%s: column sums need a bf16 output", what);
  SEGK_REQUIRE(ctx, rate >= 1 && (kh / 2) * rate <= 127 && (kw / 2) * rate <= 127, "%s: dilation rate %d out of range", what, rate);
  if (rate == 1 && slab_applicable(ctx, N, H, W, Ck, Cn, kh, kw))
    return conv_slab(ctx, what, x, wt, bias, residual, mask, scale, relu, out_f32, y, N, H, W, Ck, Cn, stream, colsum_out, pool, ldx, ldy);
  // the fused pool needs boxes of whole 2x2 windows, the TMA-store epilogue and unsplit tiles
  bool want_pool = pool && pool->pooled && !out_f32 && ctx->tma_store && !colsum_out && (H & 1) == 0 && (W & 1) == 0;
  Box b = choose_box(N, H, W, kBlockM, false, 0, 0, kh, kw, want_pool);
  if (want_pool && b.rows <= 0) {
    want_pool = false;
    b = choose_box(N, H, W, kBlockM, false, 0, 0, kh, kw);
  }
  SEGK_REQUIRE(ctx, b.rows > 0, "%s: no pixel box for %dx%dx%d", what, N, H, W);
  const int block_n = pick_block_n(ctx, Cn);
  TensorMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc = encode_nhwc_map(ctx, &maps.a[0], x, N, H, W, Ck, ldx, b.bw, b.bh, b.bn);
  if (rc) return rc;
  maps.a[1] = maps.a[2] = maps.a[3] = maps.a[0];
  rc = encode_weight_map_blocked(ctx, &maps.b, wt, Ck, Cn, kh * kw, block_n);
  if (rc) return rc;
  const bool pair_ok = block_n == 256 && ctx->pair != 0;
  if (pair_ok) {      // CTA-pair mode: each CTA of a pair stages half of the weight tile
    rc = encode_weight_map_blocked(ctx, &maps.b2, wt, Ck, Cn, kh * kw, block_n / 2);
    if (rc) return rc;
  }
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W;
  p.bw = b.bw; p.bh = b.bh; p.bn = b.bn; p.rows = b.rows;
  p.tiles_w = ceil_div(W, b.bw); p.tiles_h = ceil_div(H, b.bh); p.tiles_n = ceil_div(N, b.bn);
  p.n_tiles = Cn / block_n;
  p.phases = 1; p.s = 1;
  p.ntaps = kh * kw; p.kchunks = Ck / 64;
  p.in_H = H; p.in_W = W;
  p.out_H = H; p.out_W = W; p.ldo = Cn; p.os = 1; p.opad = 0;
  p.out = y; p.out_f32 = out_f32;
  p.bias = bias; p.residual = (const bf16*)residual; p.mask = (const bf16*)mask;
  p.scale = scale; p.relu = relu;
  p.ksplits = 1; p.ws = nullptr;
  if (pool) p.mask_bits = pool->mask_bits;         // (bits_out is set below, on the paths whose epilogue runs in the igemm kernel)
  const bool narrow = pool && out_f32 && pool->out_cols > 0;
  if (narrow) p.out_cols = pool->out_cols;
  // order the tiles so that what is larger (weights vs activations) is what concurrent CTAs share
  p.m_fastest = ((int64_t)kh * kw * Cn > (int64_t)N * H * W) ? 1 : 0;
  TapTable taps;
  conv_taps(taps, kh, kw, rate);
  // few output tiles but a long K walk (conv6 dgrad: 48 tiles x 3136 k-steps): split K across SMs
  const int tiles = p.tiles_n * p.tiles_h * p.tiles_w * p.n_tiles;
  const int force_ks = ctx->force_ksplit;
  if (want_pool) {
    // (few-tile layers are not split when the pool is fused: their tiles are small anyway)
    p.tma_store = 1;
    rc = encode_nhwc_map(ctx, &maps.c, y, N, H, W, Cn, ldy, b.bw, b.bh, b.bn);
    if (rc) return rc;
    p.pool_out = (bf16*)pool->pooled; p.pool_idx = pool->idx; p.pool_only = pool->pool_only;
    pool->fused = true;
    if (pool->bits_out) { p.bits_out = pool->bits_out; pool->bits_done = true; }
    return launch_igemm(ctx, block_n, maps, p, taps, (cudaStream_t)stream, 0, pair_ok);
  }
  if (!narrow && ((tiles * 2 <= ctx->sm_count && p.ntaps * p.kchunks >= 32 && p.kchunks >= 2) || force_ks > 0)) {
    // one wave of work units.  Measured (tools/time_conv6.py, conv6 dgrad: 50 tiles x 2240 k-steps): 2 splits
    // 483 us, 1: 766, 3: 596, 5: 562, 8: 692, 14: 918 -- more splits than one wave lets the m-tiles that share
    // a weight slice drift apart in K, and the 205 MB of weights stream from DRAM several times over
    // lockstep tap-split (IgemmParams::ts): every CTA gets the same number of k-steps and all CTAs sweep the k-chunks
    // together, so a weight slice is fetched from DRAM once.  Needs >= 2 (tile, tap) pairs per CTA and at most two
    // pixel tiles per CTA (two TMEM accumulators).
    const int tiles_per_nt = p.tiles_h * p.tiles_w * p.tiles_n;
    if (force_ks <= 0 && ctx->teamk && tiles_per_nt <= kMaxCols && p.n_tiles <= ctx->sm_count) {
      int total_taps = 0, min_act = p.ntaps;
      for (int tau = 0; tau < tiles_per_nt; ++tau) {
        const int col = tau / p.tiles_n;
        const int x0 = (col % p.tiles_w) * b.bw, y0 = (col / p.tiles_w) * b.bh;
        int act = 0;
        for (int i = 0; i < p.ntaps; ++i) {
          const int ya = y0 + taps.dy[i], xa = x0 + taps.dx[i];
          act += !(ya + b.bh <= 0 || ya >= H || xa + b.bw <= 0 || xa >= W);
        }
        if (act == 0) act = p.ntaps;                      // mirrors tap_mask(): an all-padding tile still produces zeros
        if (act < min_act) min_act = act;
        p.ts_prefix[tau] = total_taps;
        total_taps += act;
      }
      p.ts_prefix[tiles_per_nt] = total_taps;
      int G = ctx->sm_count / p.n_tiles;
      if (G > total_taps / 2) G = total_taps / 2;
      const int share = G > 0 ? ceil_div(total_taps, G) : 0;
      if (G >= 1 && share <= min_act && G > tiles_per_nt) {
        int max_parts = 1;
        for (int tau = 0; tau < tiles_per_nt; ++tau) {
          const int np = ts_owner(p.ts_prefix[tau + 1] - 1, G, total_taps) - ts_owner(p.ts_prefix[tau], G, total_taps) + 1;
          if (np > max_parts) max_parts = np;
        }
        const size_t slice = (size_t)N * H * W * Cn;
        rc = ensure_workspace(ctx, sizeof(float) * slice * max_parts);
        if (rc) return rc;
        p.ts = 1; p.ts_G = G; p.ts_tiles = tiles_per_nt; p.ts_total = total_taps;
        p.ws = (float*)ctx->ws;
        p.ws_slice = (int64_t)slice;
        rc = launch_igemm(ctx, block_n, maps, p, taps, (cudaStream_t)stream, G * p.n_tiles);
        if (rc) return rc;
        TsFinish tf;
        memset(&tf, 0, sizeof(tf));
        tf.G = G; tf.total = total_taps; tf.tiles_w = p.tiles_w; tf.tiles_n = p.tiles_n;
        tf.bw = b.bw; tf.bh = b.bh; tf.bn = b.bn; tf.H = H; tf.W = W;
        memcpy(tf.prefix, p.ts_prefix, sizeof(tf.prefix));
        const int64_t rows = (int64_t)N * H * W;
        int64_t blocks = ceil_div64(rows * (Cn / 8), 256);
        if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
        epilogue_finish_ts_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
            (const float*)ctx->ws, tf, (int64_t)slice, bias, (const bf16*)residual, (const bf16*)mask, y, out_f32, relu, scale,
            rows, Cn, p.mask_bits, ldy);
        SEGK_LAUNCHED(ctx, "igemm tap-split finish");
        if (colsum_out) return colsum_fallback(ctx, y, rows, Cn, colsum_out, (cudaStream_t)stream);
        return SEGK_OK;
      }
    }
    int ks = ctx->sm_count / tiles;
    if (force_ks > 0) ks = force_ks;
    if (ks > p.kchunks) ks = p.kchunks;   // ksplits <= kchunks <= k-steps of any tile: no split is empty
    if (ks > 16) ks = 16;
    if (ks > 1) {
      const size_t slice = (size_t)N * H * W * Cn;
      rc = ensure_workspace(ctx, sizeof(float) * slice * ks);
      if (rc) return rc;
      p.ksplits = ks;
      p.ws = (float*)ctx->ws;
      p.ws_slice = (int64_t)slice;
      rc = launch_igemm(ctx, block_n, maps, p, taps, (cudaStream_t)stream);
      if (rc) return rc;
      const int64_t rows = (int64_t)N * H * W;
      int64_t blocks = ceil_div64(rows * (Cn / 8), 256);
      if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
      epilogue_finish_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
          (const float*)ctx->ws, ks, (int64_t)slice, bias, (const bf16*)residual, (const bf16*)mask, y, out_f32, relu, scale,
          rows, Cn, p.mask_bits, ldy);
      SEGK_LAUNCHED(ctx, "igemm split-K finish");
      if (colsum_out) return colsum_fallback(ctx, y, rows, Cn, colsum_out, (cudaStream_t)stream);
      return SEGK_OK;
    }
  }
  p.tma_store = (!out_f32 && (ctx->tma_store || pitched_out)) ? 1 : 0;
  if (p.tma_store) {
    rc = encode_nhwc_map(ctx, &maps.c, y, N, H, W, Cn, ldy, b.bw, b.bh, b.bn);
    if (rc) return rc;
  }
  if (pool && pool->bits_out && !out_f32 && !(ctx->hybrid)) { p.bits_out = pool->bits_out; pool->bits_done = true; }
  // Hybrid schedule for a partly filled last wave (conv5_x: 180 tiles on 148 SMs, conv6/conv7 forward: 368): whole waves
  // run as they are, the remaining tiles are split in K so that their units fill one more (short) wave.
  {
    const int sm = ctx->sm_count;
    const int waves = tiles / sm, rem = tiles % sm;
    int ksr = rem > 0 ? sm / rem : 0;
    if (ksr > p.kchunks) ksr = p.kchunks;            // ksplits <= kchunks <= k-steps of any tile: no piece is empty
    if (ksr > 8) ksr = 8;
    if (ctx->hybrid && !narrow && !pitched_out && waves >= 1 && waves <= 4 && ksr >= 2 && p.ntaps * p.kchunks >= 8 * ksr) {
      const size_t slice = (size_t)N * H * W * Cn;
      rc = ensure_workspace(ctx, sizeof(float) * slice * ksr);
      if (rc) return rc;
      p.hyb = 1; p.hyb_full = tiles - rem; p.ksplits = ksr;
      p.ws = (float*)ctx->ws;
      p.ws_slice = (int64_t)slice;
      rc = launch_igemm(ctx, block_n, maps, p, taps, (cudaStream_t)stream);
      if (rc) return rc;
      int64_t blocks = ceil_div64((int64_t)rem * p.rows * (block_n / 8), 256);
      if (blocks > (int64_t)sm * 8) blocks = (int64_t)sm * 8;
      epilogue_finish_tiles_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const float*)ctx->ws, p, p.hyb_full, rem, block_n);
      SEGK_LAUNCHED(ctx, "igemm hybrid finish");
      if (colsum_out) return colsum_fallback(ctx, y, (int64_t)N * H * W, Cn, colsum_out, (cudaStream_t)stream);
      return SEGK_OK;
    }
  }
  if (colsum_out) {
    const int total = tiles < ctx->sm_count ? tiles : ctx->sm_count;
    if (p.m_fastest || total < p.n_tiles) {        // the tiles of a CTA change their channel tile: sum the finished tensor
      rc = launch_igemm(ctx, block_n, maps, p, taps, (cudaStream_t)stream);
      if (rc) return rc;
      return colsum_fallback(ctx, y, (int64_t)N * H * W, Cn, colsum_out, (cudaStream_t)stream);
    }
    rc = colsum_scratch(ctx, total, block_n, &p.colsum);
    if (rc) return rc;
    int grid_used = 0;
    rc = launch_igemm(ctx, block_n, maps, p, taps, (cudaStream_t)stream, 0, pair_ok, &grid_used);
    if (rc) return rc;
    return colsum_finish(ctx, colsum_out, grid_used, p.n_tiles, block_n, (cudaStream_t)stream);
  }
  return launch_igemm(ctx, block_n, maps, p, taps, (cudaStream_t)stream, 0, pair_ok);
}


// decimated views of a [N, s*H, s*W, C] tensor: view (py,px) holds pixels (s*i+py, s*j+px)
int encode_decimated_maps(segk_ctx* ctx, TensorMaps& maps, const void* base, int N, int H, int W, int C, int s,
                          const Box& b, int ld = 0) {
  const int OH = H * s, OW = W * s;
  if (ld <= 0) ld = C;
  for (int py = 0; py < s; ++py)
    for (int px = 0; px < s; ++px) {
      const bf16* v = (const bf16*)base + ((int64_t)py * OW + px) * ld;
      int rc = encode_act_map(ctx, &maps.a[py * s + px], v, N, H, W, C, (int64_t)s * ld, (int64_t)s * OW * ld,
                              (int64_t)OH * OW * ld, b.bw, b.bh, b.bn);
      if (rc) return rc;
    }
  return SEGK_OK;
}

// taps of the stride-s conv that is the transposed conv's gradient: dY row s*i - p + ky
void strided_taps(TapTable& t, int k, int s) {
  memset(&t, 0, sizeof(t));
  const int p = s / 2;
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx) {
      const int ry = ky - p, rx = kx - p;
      const int py = ((ry % s) + s) % s, px = ((rx % s) + s) % s;
      const int i = ky * k + kx;
      t.dy[i] = (int8_t)((ry - py) / s);
      t.dx[i] = (int8_t)((rx - px) / s);
      t.map[i] = (int8_t)(py * s + px);
    }
}

}  // namespace

namespace tch {
Box pick_box(int N, int H, int W, int max_rows, bool exact) { return choose_box(N, H, W, max_rows, exact); }
int act_map(segk_ctx* ctx, CUtensorMap* m, const void* base, int N, int H, int W, int C, int bw, int bh, int bn, int ld) {
  if (ld <= 0) ld = C;
  return encode_act_map(ctx, m, base, N, H, W, C, ld, (int64_t)W * ld, (int64_t)H * W * ld, bw, bh, bn);
}
int weight_map(segk_ctx* ctx, CUtensorMap* m, const void* base, int K, int Nrows, int T, int block_n) {
  return encode_weight_map(ctx, m, base, K, Nrows, T, block_n);
}
int workspace(segk_ctx* ctx, size_t bytes) { return ensure_workspace(ctx, bytes); }
}  // namespace tch

extern "C" {

int segk_deconv2d_fwd(segk_ctx* ctx, const void* x, const void* wk, const float* bias, const void* residual, void* y,
                      int N, int H, int W, int Cin, int Cout, int k, int s, unsigned flags, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  const SegkPitch pitch = segk_take_pitch(ctx);
  SEGK_REQUIRE(ctx, x && wk && y && N > 0 && H > 0 && W > 0, "deconv2d_fwd: bad args");
  SEGK_REQUIRE(ctx, pitch.in == 0 || pitch.in == Cin, "deconv2d_fwd: the input is dense");
  const int ldy = pitch.out > 0 ? pitch.out : Cout;
  SEGK_REQUIRE(ctx, ldy == Cout || (ldy > Cout && !residual && !(flags & SEGK_EPI_OUT_F32) && (((uintptr_t)y) & 15) == 0),
               "deconv2d_fwd: a pitched output (%d channels apart) needs a bf16 output without residual", ldy);
  SEGK_REQUIRE(ctx, k == 2 * s && s >= 2 && s % 2 == 0 && s * s * 4 <= 32767, "deconv2d_fwd: need k == 2*stride, even stride");
  SEGK_REQUIRE(ctx, Cin % 64 == 0 && Cout % 64 == 0 && Cin > 0 && Cout > 0,
               "deconv2d_fwd: tensor-core path needs channels %% 64 == 0 (got %d -> %d); use segk_deconv2d_small_fwd", Cin,
               Cout);
  const Box b = choose_box(N, H, W, kBlockM, false);
  SEGK_REQUIRE(ctx, b.rows > 0, "deconv2d_fwd: no pixel box");
  const int block_n = pick_block_n(ctx, Cout);
  TensorMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc = encode_act_map(ctx, &maps.a[0], x, N, H, W, Cin, Cin, (int64_t)W * Cin, (int64_t)H * W * Cin, b.bw, b.bh, b.bn);
  if (rc) return rc;
  maps.a[1] = maps.a[2] = maps.a[3] = maps.a[0];
  rc = encode_weight_map_blocked(ctx, &maps.b, wk, Cin, Cout, s * s * 4, block_n);
  if (rc) return rc;
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H + 1; p.W = W + 1;                 // bounds of q; the tiles cover H x W of it per phase (phase_origin)
  p.bw = b.bw; p.bh = b.bh; p.bn = b.bn; p.rows = b.rows;
  p.tiles_w = ceil_div(W, b.bw); p.tiles_h = ceil_div(H, b.bh); p.tiles_n = ceil_div(N, b.bn);
  p.n_tiles = Cout / block_n;
  p.phases = s * s; p.s = s;
  p.ntaps = 4; p.kchunks = Cin / 64;
  p.in_H = H; p.in_W = W;
  p.out_H = H * s; p.out_W = W * s; p.ldo = ldy; p.os = s; p.opad = s / 2;      // (ldo is the row stride of `out` only here)
  p.out = y; p.out_f32 = (flags & SEGK_EPI_OUT_F32) ? 1 : 0;
  p.bias = bias; p.residual = (const bf16*)residual; p.mask = nullptr;
  p.scale = 1.f; p.relu = (flags & SEGK_EPI_RELU) ? 1 : 0;
  p.ksplits = 1; p.ws = nullptr;
  if (s == 2 && !p.out_f32 && ctx->tma_store) {
    // bf16 output through shared memory + TMA store, one decimated view of y per phase: phase (ay, ax) owns the pixels
    // (2 j + (ay + 1) % 2, 2 i + (ax + 1) % 2) -- coalesced 128-byte rows instead of 64-byte pieces at every other pixel
    const int OH = H * s, OW = W * s;
    for (int ay = 0; ay < s; ++ay)
      for (int ax = 0; ax < s; ++ax) {
        const int oy0 = (ay - s / 2 + s) % s, ox0 = (ax - s / 2 + s) % s;
        const bf16* v = (const bf16*)y + ((int64_t)oy0 * OW + ox0) * ldy;
        rc = encode_act_map(ctx, &maps.cph[ay * s + ax], v, N, H, W, Cout, (int64_t)s * ldy, (int64_t)s * OW * ldy, (int64_t)OH * OW * ldy,
                            b.bw, b.bh, b.bn);
        if (rc) return rc;
      }
    p.tma_store = 1;
  }
  TapTable taps;
  memset(&taps, 0, sizeof(taps));
  for (int u = 0; u < 4; ++u) {
    taps.dy[u] = (int8_t)(u / 2 - 1);
    taps.dx[u] = (int8_t)(u % 2 - 1);
  }
  return launch_igemm(ctx, block_n, maps, p, taps, (cudaStream_t)stream);
}

static int strided_conv_impl(segk_ctx* ctx, const void* dy, const void* wd, const void* relu_mask, const float* bias, int relu,
                            void* dx, float* dx_colsum, int N, int H, int W, int Cin, int Cout, int k, int s, void* stream, int dy_ld);

int segk_deconv2d_dgrad(segk_ctx* ctx, const void* dy, const void* wd, const void* relu_mask, void* dx, float* dx_colsum,
                        int N, int H, int W, int Cin, int Cout, int k, int s, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  const SegkPitch pitch = segk_take_pitch(ctx);
  SEGK_REQUIRE(ctx, pitch.out == 0 || pitch.out == Cin, "deconv2d_dgrad: the output is dense");
  SEGK_REQUIRE(ctx, pitch.in == 0 || pitch.in >= Cout, "deconv2d_dgrad: dy pitch %d < %d channels", pitch.in, Cout);
  return strided_conv_impl(ctx, dy, wd, relu_mask, nullptr, 0, dx, dx_colsum, N, H, W, Cin, Cout, k, s, stream, pitch.in);
}

// Conv2D with kernel 2s x 2s, stride s, SAME (LidCamNet.py:28-33's 4x4 stride-2 encoder convs) = the strided conv that is
// the transposed conv's input gradient, with the conv_layer epilogue (bias, ReLU).  x [N,sH,sW,Cin] -> y [N,H,W,Cout];
// wd = the dgrad layout of segk_pack_deconv_weights applied to the HWIO weights [k,k,Cin,Cout] (the layout of a transposed
// conv Cout -> Cin).  Gradients: segk_deconv2d_fwd (input gradient) and segk_deconv2d_wgrad with the roles of x / dy swapped.
int segk_conv2d_strided_fwd(segk_ctx* ctx, const void* x, const void* wd, const float* bias, void* y, int N, int H, int W,
                            int Cin, int Cout, int k, int s, unsigned flags, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, !(flags & SEGK_EPI_OUT_F32), "conv2d_strided_fwd: bf16 output only");
  return strided_conv_impl(ctx, x, wd, nullptr, bias, (flags & SEGK_EPI_RELU) ? 1 : 0, y, nullptr, N, H, W, Cout, Cin, k, s, stream, 0);
}

static int strided_conv_impl(segk_ctx* ctx, const void* dy, const void* wd, const void* relu_mask, const float* bias, int relu,
                            void* dx, float* dx_colsum, int N, int H, int W, int Cin, int Cout, int k, int s, void* stream, int dy_ld) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, dy && wd && dx && N > 0 && H > 0 && W > 0, "deconv2d_dgrad: bad args");
  SEGK_REQUIRE(ctx, k == 4 && s == 2, "deconv2d_dgrad: tensor-core path supports k=4, stride 2 (got k=%d s=%d)", k, s);
  SEGK_REQUIRE(ctx, Cin % 64 == 0 && Cout % 64 == 0 && Cin > 0 && Cout > 0,
               "deconv2d_dgrad: tensor-core path needs channels %% 64 == 0 (got %d -> %d)", Cin, Cout);
  const Box b = choose_box(N, H, W, kBlockM, false);
  SEGK_REQUIRE(ctx, b.rows > 0, "deconv2d_dgrad: no pixel box");
  const int block_n = pick_block_n(ctx, Cin);
  TensorMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc = encode_decimated_maps(ctx, maps, dy, N, H, W, Cout, s, b, dy_ld);
  if (rc) return rc;
  rc = encode_weight_map_blocked(ctx, &maps.b, wd, Cout, Cin, k * k, block_n);
  if (rc) return rc;
  const bool pair_ok = block_n == 256 && ctx->pair != 0;      // CTA pairs (igemm_pair_kernel): each CTA stages half of the weight tile
  if (pair_ok) {
    rc = encode_weight_map_blocked(ctx, &maps.b2, wd, Cout, Cin, k * k, block_n / 2);
    if (rc) return rc;
  }
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W;
  p.bw = b.bw; p.bh = b.bh; p.bn = b.bn; p.rows = b.rows;
  p.tiles_w = ceil_div(W, b.bw); p.tiles_h = ceil_div(H, b.bh); p.tiles_n = ceil_div(N, b.bn);
  p.n_tiles = Cin / block_n;
  p.phases = 1; p.s = 1;
  p.ntaps = k * k; p.kchunks = Cout / 64;
  p.in_H = H; p.in_W = W;
  p.out_H = H; p.out_W = W; p.ldo = Cin; p.os = 1; p.opad = 0;
  p.out = dx; p.out_f32 = 0;
  p.mask = (const bf16*)relu_mask;
  p.bias = bias; p.relu = relu;
  p.scale = 1.f;
  p.ksplits = 1; p.ws = nullptr;
  TapTable taps;
  strided_taps(taps, k, s);
  if (dx_colsum) {
    const int tiles = p.tiles_n * p.tiles_h * p.tiles_w * p.n_tiles;
    const int total = tiles < ctx->sm_count ? tiles : ctx->sm_count;
    if (total < p.n_tiles) {
      rc = launch_igemm(ctx, block_n, maps, p, taps, (cudaStream_t)stream);
      if (rc) return rc;
      return colsum_fallback(ctx, dx, (int64_t)N * H * W, Cin, dx_colsum, (cudaStream_t)stream);
    }
    rc = colsum_scratch(ctx, total, block_n, &p.colsum);
    if (rc) return rc;
    int grid_used = 0;
    rc = launch_igemm(ctx, block_n, maps, p, taps, (cudaStream_t)stream, 0, pair_ok, &grid_used);
    if (rc) return rc;
    return colsum_finish(ctx, dx_colsum, grid_used, p.n_tiles, block_n, (cudaStream_t)stream);
  }
  return launch_igemm(ctx, block_n, maps, p, taps, (cudaStream_t)stream, 0, pair_ok);
}

int segk_deconv2d_wgrad(segk_ctx* ctx, const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout,
                        int k, int s, int accumulate, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  const SegkPitch pitch = segk_take_pitch(ctx);
  SEGK_REQUIRE(ctx, pitch.out == 0 && (pitch.in == 0 || pitch.in >= Cout), "deconv2d_wgrad: only dy may be pitched (>= %d channels)", Cout);
  SEGK_REQUIRE(ctx, x && dy && dw && N > 0 && H > 0 && W > 0, "deconv2d_wgrad: bad args");
  SEGK_REQUIRE(ctx, k == 4 && s == 2, "deconv2d_wgrad: tensor-core path supports k=4, stride 2 (got k=%d s=%d)", k, s);
  SEGK_REQUIRE(ctx, Cin % 64 == 0 && Cout % 64 == 0 && Cin > 0 && Cout > 0,
               "deconv2d_wgrad: tensor-core path needs channels %% 64 == 0 (got %d -> %d)", Cin, Cout);
  cudaStream_t st = (cudaStream_t)stream;
  const Box b = choose_box(N, H, W, 64, true);
  SEGK_REQUIRE(ctx, b.rows == 64, "deconv2d_wgrad: cannot tile %dx%dx%d into 64-pixel boxes", N, H, W);
  const int block_n = pick_block_n(ctx, Cin);
  TensorMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc = encode_decimated_maps(ctx, maps, dy, N, H, W, Cout, s, b, pitch.in);  // A side: dY (rows = (tap, co))
  if (rc) return rc;
  rc = encode_act_map(ctx, &maps.b, x, N, H, W, Cin, Cin, (int64_t)W * Cin, (int64_t)H * W * Cin, b.bw, b.bh, b.bn);
  if (rc) return rc;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W;
  p.bw = b.bw; p.bh = b.bh; p.bn = b.bn;
  p.tiles_w = ceil_div(W, b.bw); p.tiles_h = ceil_div(H, b.bh); p.tiles_n = ceil_div(N, b.bn);
  const int n_ptiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.kchunks_in = Cout / 64;
  p.n_rb = k * k * p.kchunks_in;
  p.n_rbp = (p.n_rb + 1) / 2;
  p.n_tiles = Cin / block_n;
  const int base_items = p.n_rbp * p.n_tiles;
  int splits = base_items >= 16 ? ctx->sm_count / base_items : ceil_div(2 * ctx->sm_count, base_items);
  const int max_splits = ceil_div(n_ptiles, 8);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const int per_split = ceil_div(n_ptiles, splits);
  p.splits = ceil_div(n_ptiles, per_split);
  p.Cin_total = Cout; p.Cout_total = Cin;
  p.dw_tap_stride = Cout * Cin; p.dw_row_stride = Cin; p.dw_col_stride = 1;
  p.dw = dw;
  p.direct = (!accumulate && p.splits == 1 && p.n_rb % 2 == 0) ? 1 : 0;
  const size_t n_dw = (size_t)k * k * Cin * Cout;
  const bool partials = !p.direct && n_dw % 4 == 0 && sizeof(float) * n_dw * p.splits <= ((size_t)1 << 30);
  if (partials) {
    rc = segk_ws4(ctx, sizeof(float) * n_dw * p.splits);
    if (rc) return rc;
    p.dw = (float*)ctx->ws4;
    p.part_stride = (int64_t)n_dw;
    p.direct = 1;
  }
  SEGK_REQUIRE(ctx, p.direct, "deconv2d_wgrad: %d splits of %zu weights exceed the 1 GiB partial-sum workspace", p.splits, n_dw);
  TapTable taps;
  strided_taps(taps, k, s);
  const int total = p.splits * p.n_rbp * p.n_tiles;
  const int grid = total < ctx->sm_count ? total : ctx->sm_count;
  switch (block_n) {
    case 256: rc = launch_wgrad_t<256>(ctx, maps, p, taps, grid, st); break;
    case 128: rc = launch_wgrad_t<128>(ctx, maps, p, taps, grid, st); break;
    default: rc = launch_wgrad_t<64>(ctx, maps, p, taps, grid, st); break;
  }
  if (rc) return rc;
  if (partials) return segk_reduce_partials(ctx, (const float*)ctx->ws4, dw, n_dw, p.splits, accumulate, st);
  return SEGK_OK;
}

// ---- phase-packed transposed conv for tiny Cout (conv_t3; patch.cu has the layout kernels) ----------------------
static int packed_check(segk_ctx* ctx, const char* what, int N, int H, int W, int Cin, int Cout, int k, int s) {
  SEGK_REQUIRE(ctx, N > 0 && H > 0 && W > 0, "%s: empty tensor", what);
  SEGK_REQUIRE(ctx, k == 2 * s && s >= 2 && s % 2 == 0, "%s: need k == 2*stride, even stride", what);
  const int R = s * s * Cout;
  SEGK_REQUIRE(ctx, Cin % 64 == 0 && (R == 64 || R == 128 || R == 256) && Cout % 2 == 0 && 32 % (s * Cout) == 0,
               "%s: packed form needs Cin %% 64 == 0, s*s*Cout in {64,128,256}, even Cout, s*Cout | 32 (got %d -> %d, s=%d); no fallback",
               what, Cin, Cout, s);
  return SEGK_OK;
}

int segk_deconv2d_packed_fwd(segk_ctx* ctx, const void* x, const void* bf, const float* bias, float* y, int N, int H, int W,
                             int Cin, int Cout, int k, int s, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && bf && y, "deconv2d_packed_fwd: null pointer");
  int rc = packed_check(ctx, "deconv2d_packed_fwd", N, H, W, Cin, Cout, k, s);
  if (rc) return rc;
  const int R = s * s * Cout;
  const Box b = choose_box(N, H + 1, W + 1, kBlockM, false, H, W);
  SEGK_REQUIRE(ctx, b.rows > 0, "deconv2d_packed_fwd: no pixel box");
  const int block_n = R;                  // one channel tile: 64 / 128 / 256 columns
  TensorMaps maps;
  memset(&maps, 0, sizeof(maps));
  rc = encode_act_map(ctx, &maps.a[0], x, N, H, W, Cin, Cin, (int64_t)W * Cin, (int64_t)H * W * Cin, b.bw, b.bh, b.bn);
  if (rc) return rc;
  maps.a[1] = maps.a[2] = maps.a[3] = maps.a[0];
  rc = encode_weight_map_blocked(ctx, &maps.b, bf, Cin, R, 4, block_n);
  if (rc) return rc;
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H + 1; p.W = W + 1;
  p.bw = b.bw; p.bh = b.bh; p.bn = b.bn; p.rows = b.rows;
  p.tiles_w = ceil_div(W + 1, b.bw); p.tiles_h = ceil_div(H + 1, b.bh); p.tiles_n = ceil_div(N, b.bn);
  p.n_tiles = 1;
  p.phases = 1; p.s = 1;
  p.ntaps = 4; p.kchunks = Cin / 64;
  p.in_H = H; p.in_W = W;
  p.out_H = H * s; p.out_W = W * s; p.ldo = Cout; p.os = 1; p.opad = s / 2;
  p.out = y; p.out_f32 = 1;
  p.bias = bias; p.scale = 1.f;
  p.ksplits = 1;
  p.pack_s = s; p.pack_co = Cout;
  TapTable taps;
  memset(&taps, 0, sizeof(taps));
  for (int u = 0; u < 4; ++u) {            // tap t = dy*2 + dx reads x[r - 1 + dy, c - 1 + dx]
    taps.dy[u] = (int8_t)(u / 2 - 1);
    taps.dx[u] = (int8_t)(u % 2 - 1);
  }
  return launch_igemm(ctx, block_n, maps, p, taps, (cudaStream_t)stream);
}

int segk_deconv2d_packed_dgrad(segk_ctx* ctx, const void* dyb, const void* bt, const void* relu_mask, void* dx,
                               float* dx_colsum, int N, int H, int W, int Cin, int Cout, int k, int s, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, dyb && bt && dx, "deconv2d_packed_dgrad: null pointer");
  int rc = packed_check(ctx, "deconv2d_packed_dgrad", N, H, W, Cin, Cout, k, s);
  if (rc) return rc;
  const int R = s * s * Cout;
  const Box b = choose_box(N, H, W, kBlockM, false);
  SEGK_REQUIRE(ctx, b.rows > 0, "deconv2d_packed_dgrad: no pixel box");
  const int block_n = pick_block_n(ctx, Cin);
  TensorMaps maps;
  memset(&maps, 0, sizeof(maps));
  // A = dyb [N, H+1, W+1, R]; tap t = dy*2 + dx reads block (i + 1 - dy, j + 1 - dx): always inside the block grid
  rc = encode_act_map(ctx, &maps.a[0], dyb, N, H + 1, W + 1, R, R, (int64_t)(W + 1) * R, (int64_t)(H + 1) * (W + 1) * R, b.bw,
                      b.bh, b.bn);
  if (rc) return rc;
  maps.a[1] = maps.a[2] = maps.a[3] = maps.a[0];
  rc = encode_weight_map_blocked(ctx, &maps.b, bt, R, Cin, 4, block_n);
  if (rc) return rc;
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W;
  p.bw = b.bw; p.bh = b.bh; p.bn = b.bn; p.rows = b.rows;
  p.tiles_w = ceil_div(W, b.bw); p.tiles_h = ceil_div(H, b.bh); p.tiles_n = ceil_div(N, b.bn);
  p.n_tiles = Cin / block_n;
  p.phases = 1; p.s = 1;
  p.ntaps = 4; p.kchunks = R / 64;
  p.in_H = H + 1; p.in_W = W + 1;
  p.out_H = H; p.out_W = W; p.ldo = Cin; p.os = 1; p.opad = 0;
  p.out = dx; p.out_f32 = 0;
  p.mask = (const bf16*)relu_mask;
  p.scale = 1.f;
  p.ksplits = 1;
  p.tma_store = ctx->tma_store ? 1 : 0;
  if (p.tma_store) {
    rc = encode_act_map(ctx, &maps.c, dx, N, H, W, Cin, Cin, (int64_t)W * Cin, (int64_t)H * W * Cin, b.bw, b.bh, b.bn);
    if (rc) return rc;
  }
  TapTable taps;
  memset(&taps, 0, sizeof(taps));
  for (int u = 0; u < 4; ++u) {
    taps.dy[u] = (int8_t)(1 - u / 2);
    taps.dx[u] = (int8_t)(1 - u % 2);
  }
  if (dx_colsum) {
    const int tiles = p.tiles_n * p.tiles_h * p.tiles_w * p.n_tiles;
    const int total = tiles < ctx->sm_count ? tiles : ctx->sm_count;
    if (total < p.n_tiles) {
      rc = launch_igemm(ctx, block_n, maps, p, taps, (cudaStream_t)stream);
      if (rc) return rc;
      return colsum_fallback(ctx, dx, (int64_t)N * H * W, Cin, dx_colsum, (cudaStream_t)stream);
    }
    rc = colsum_scratch(ctx, total, block_n, &p.colsum);
    if (rc) return rc;
    rc = launch_igemm(ctx, block_n, maps, p, taps, (cudaStream_t)stream);
    if (rc) return rc;
    return colsum_finish(ctx, dx_colsum, (total / p.n_tiles) * p.n_tiles, p.n_tiles, block_n, (cudaStream_t)stream);
  }
  return launch_igemm(ctx, block_n, maps, p, taps, (cudaStream_t)stream);
}

int segk_deconv2d_packed_wgrad(segk_ctx* ctx, const void* x, const void* dyb, float* dwt, int N, int H, int W, int Cin,
                               int Cout, int k, int s, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, x && dyb && dwt, "deconv2d_packed_wgrad: null pointer");
  int rc = packed_check(ctx, "deconv2d_packed_wgrad", N, H, W, Cin, Cout, k, s);
  if (rc) return rc;
  const int R = s * s * Cout;
  cudaStream_t st = (cudaStream_t)stream;
  // pixels = blocks (n, r, c) of the (H+1) x (W+1) block grid; rows (t, ci) from x shifted by (dy-1, dx-1), cols = R
  const Box b = choose_box(N, H + 1, W + 1, 64, true);
  SEGK_REQUIRE(ctx, b.rows == 64, "deconv2d_packed_wgrad: cannot tile %dx%dx%d into 64-block boxes", N, H + 1, W + 1);
  const int block_n = R;
  TensorMaps maps;
  memset(&maps, 0, sizeof(maps));
  rc = encode_act_map(ctx, &maps.a[0], x, N, H, W, Cin, Cin, (int64_t)W * Cin, (int64_t)H * W * Cin, b.bw, b.bh, b.bn);
  if (rc) return rc;
  maps.a[1] = maps.a[2] = maps.a[3] = maps.a[0];
  rc = encode_act_map(ctx, &maps.b, dyb, N, H + 1, W + 1, R, R, (int64_t)(W + 1) * R, (int64_t)(H + 1) * (W + 1) * R, b.bw, b.bh,
                      b.bn);
  if (rc) return rc;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H + 1; p.W = W + 1;
  p.bw = b.bw; p.bh = b.bh; p.bn = b.bn;
  p.tiles_w = ceil_div(W + 1, b.bw); p.tiles_h = ceil_div(H + 1, b.bh); p.tiles_n = ceil_div(N, b.bn);
  const int n_ptiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.kchunks_in = Cin / 64;
  p.n_rb = 4 * p.kchunks_in;
  p.n_rbp = (p.n_rb + 1) / 2;
  p.n_tiles = 1;
  const int base_items = p.n_rbp;
  int splits = base_items >= 16 ? ctx->sm_count / base_items : ceil_div(2 * ctx->sm_count, base_items);
  const int max_splits = ceil_div(n_ptiles, 8);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const int per_split = ceil_div(n_ptiles, splits);
  p.splits = ceil_div(n_ptiles, per_split);
  p.Cin_total = Cin; p.Cout_total = R;
  p.dw_tap_stride = Cin * R; p.dw_row_stride = R; p.dw_col_stride = 1;
  p.dw = dwt;
  p.direct = 1;
  const size_t n_dw = (size_t)4 * Cin * R;
  const bool partials = p.splits > 1;
  if (partials) {
    rc = segk_ws4(ctx, sizeof(float) * n_dw * p.splits);
    if (rc) return rc;
    p.dw = (float*)ctx->ws4;
    p.part_stride = (int64_t)n_dw;
  }
  TapTable taps;
  memset(&taps, 0, sizeof(taps));
  for (int u = 0; u < 4; ++u) {
    taps.dy[u] = (int8_t)(u / 2 - 1);
    taps.dx[u] = (int8_t)(u % 2 - 1);
  }
  const int total = p.splits * p.n_rbp * p.n_tiles;
  const int grid = total < ctx->sm_count ? total : ctx->sm_count;
  switch (block_n) {
    case 256: rc = launch_wgrad_t<256>(ctx, maps, p, taps, grid, st); break;
    case 128: rc = launch_wgrad_t<128>(ctx, maps, p, taps, grid, st); break;
    default: rc = launch_wgrad_t<64>(ctx, maps, p, taps, grid, st); break;
  }
  if (rc) return rc;
  if (partials) return segk_reduce_partials(ctx, (const float*)ctx->ws4, dwt, n_dw, p.splits, 0, st);
  return SEGK_OK;
}

// bits of a finished bf16 tensor (paths whose epilogue did not write them)
static int relu_bits_of(segk_ctx* ctx, const void* y, uint32_t* bits, int64_t rows, int C, void* stream) {
  const int64_t nwords = rows * (C / 32);
  int64_t blocks = ceil_div64(nwords, 256);
  if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
  relu_bits_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)y, bits, nwords);
  SEGK_LAUNCHED(ctx, "relu bits");
  return SEGK_OK;
}

int segk_conv2d_fwd_narrow(segk_ctx* ctx, const void* x, const void* wk, const float* bias, float* y, int out_cols, int N, int H,
                           int W, int Cin, int Cout, int kh, int kw, unsigned flags, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, y && out_cols > 0 && out_cols <= Cout, "conv2d_fwd_narrow: 0 < out_cols <= Cout (got %d of %d)", out_cols, Cout);
  PoolArgs ex{nullptr, nullptr, 0, false};
  ex.out_cols = out_cols;
  return conv_igemm(ctx, "conv2d_fwd_narrow", x, wk, bias, nullptr, nullptr, 1.f, (flags & SEGK_EPI_RELU) ? 1 : 0, 1, y, N, H, W, Cin,
                    Cout, kh, kw, stream, nullptr, 1, &ex);
}

int segk_relu_bits(segk_ctx* ctx, const void* y, uint32_t* bits, int64_t rows, int C, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, y && bits && rows > 0 && C > 0 && C % 32 == 0 && (((uintptr_t)y) & 15) == 0, "relu_bits: bad args (C %% 32 == 0)");
  return relu_bits_of(ctx, y, bits, rows, C, stream);
}

int segk_conv2d_fwd(segk_ctx* ctx, const void* x, const void* wk, const float* bias, const void* residual, void* y,
                    uint32_t* relu_bits, int N, int H, int W, int Cin, int Cout, int kh, int kw, unsigned flags, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  const SegkPitch pitch = segk_take_pitch(ctx);
  if (!relu_bits)
    return conv_igemm(ctx, "conv2d_fwd", x, wk, bias, residual, nullptr, 1.f, (flags & SEGK_EPI_RELU) ? 1 : 0,
                      (flags & SEGK_EPI_OUT_F32) ? 1 : 0, y, N, H, W, Cin, Cout, kh, kw, stream, nullptr, 1, nullptr, pitch.in, pitch.out);
  SEGK_REQUIRE(ctx, !(flags & SEGK_EPI_OUT_F32) && Cout % 32 == 0, "conv2d_fwd: relu_bits need a bf16 output with Cout %% 32 == 0");
  PoolArgs ex{nullptr, nullptr, 0, false};
  ex.bits_out = relu_bits;
  const int rc = conv_igemm(ctx, "conv2d_fwd", x, wk, bias, residual, nullptr, 1.f, (flags & SEGK_EPI_RELU) ? 1 : 0, 0, y, N, H, W,
                            Cin, Cout, kh, kw, stream, nullptr, 1, &ex, pitch.in, pitch.out);
  if (rc || ex.bits_done) return rc;
  return relu_bits_of(ctx, y, relu_bits, (int64_t)N * H * W, Cout, stream);
}

int segk_conv2d_fwd_pool(segk_ctx* ctx, const void* x, const void* wk, const float* bias, void* y, void* pooled, uint8_t* idx,
                         int pool_only, int N, int H, int W, int Cin, int Cout, int kh, int kw, unsigned flags, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  const SegkPitch pitch = segk_take_pitch(ctx);
  SEGK_REQUIRE(ctx, y && pooled && idx, "conv2d_fwd_pool: null pointer");
  SEGK_REQUIRE(ctx, !(flags & SEGK_EPI_OUT_F32), "conv2d_fwd_pool: bf16 output only");
  SEGK_REQUIRE(ctx, (H & 1) == 0 && (W & 1) == 0, "conv2d_fwd_pool: need even H, W (got %dx%d)", H, W);
  PoolArgs pool{pooled, idx, pool_only ? 1 : 0, false};
  const int rc = conv_igemm(ctx, "conv2d_fwd_pool", x, wk, bias, nullptr, nullptr, 1.f, (flags & SEGK_EPI_RELU) ? 1 : 0, 0, y, N, H,
                            W, Cin, Cout, kh, kw, stream, nullptr, 1, &pool, pitch.in, pitch.out);
  if (rc || pool.fused) return rc;
  ctx->pitch_in = pitch.out;                                                   // (the pool reads the tensor just written)
  return segk_maxpool2x2_fwd(ctx, y, pooled, idx, N, H, W, Cout, stream);     // (tile geometry without whole windows)
}

int segk_conv2d_dgrad(segk_ctx* ctx, const void* dy, const void* wd, const void* relu_mask, const uint32_t* relu_mask_bits,
                      const void* residual, void* dx, float* dx_colsum, float scale, int N, int H, int W, int Cin, int Cout,
                      int kh, int kw, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  const SegkPitch pitch = segk_take_pitch(ctx);
  SEGK_REQUIRE(ctx, pitch.out == 0 || pitch.out == Cin, "conv2d_dgrad: the output is dense");
  SEGK_REQUIRE(ctx, !(relu_mask && relu_mask_bits), "conv2d_dgrad: pass the ReLU mask as a tensor OR as bits");
  SEGK_REQUIRE(ctx, !relu_mask_bits || Cin % 32 == 0, "conv2d_dgrad: mask bits need Cin %% 32 == 0");
  // GEMM-K = Cout (channels of dy), GEMM-N = Cin (channels of dx); wd holds the taps reversed.
  PoolArgs ex{nullptr, nullptr, 0, false};
  ex.mask_bits = relu_mask_bits;
  return conv_igemm(ctx, "conv2d_dgrad", dy, wd, nullptr, residual, relu_mask, scale, 0, 0, dx, N, H, W, Cout, Cin, kh,
                    kw, stream, dx_colsum, 1, relu_mask_bits ? &ex : nullptr, pitch.in, 0);
}

static int conv2d_wgrad_impl(segk_ctx* ctx, const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout,
                             int kh, int kw, int accumulate, void* stream, int rate) {
  if (!ctx) return SEGK_EINVAL;
  const SegkPitch pitch = segk_take_pitch(ctx);
  SEGK_REQUIRE(ctx, pitch.out == 0 && (pitch.in == 0 || pitch.in >= Cout), "conv2d_wgrad: only dy may be pitched (>= %d channels)", Cout);
  const int dy_ld = pitch.in > 0 ? pitch.in : Cout;
  SEGK_REQUIRE(ctx, x && dy && dw, "conv2d_wgrad: null pointer");
  SEGK_REQUIRE(ctx, rate >= 1 && (kh / 2) * rate <= 127 && (kw / 2) * rate <= 127, "conv2d_wgrad: dilation rate %d out of range", rate);
  SEGK_REQUIRE(ctx, N > 0 && H > 0 && W > 0, "conv2d_wgrad: empty tensor");
  SEGK_REQUIRE(ctx, Cin % 64 == 0 && Cout % 64 == 0 && Cin > 0 && Cout > 0,
               "conv2d_wgrad: tensor-core path needs channel counts that are multiples of 64 (got %d -> %d); no fallback",
               Cin, Cout);
  SEGK_REQUIRE(ctx, (kh & 1) && (kw & 1) && kh * kw <= kMaxTaps, "conv2d_wgrad: odd kernel sizes up to %d taps", kMaxTaps);
  SEGK_REQUIRE(ctx, (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dw) & 15) == 0, "conv2d_wgrad: 16-byte alignment");
  if (rate == 1) {
    const int handled = segk_wslab_try(ctx, x, dy, dw, N, H, W, Cin, Cout, kh, kw, accumulate, stream, dy_ld);
    if (handled != 0) return handled < 0 ? handled : SEGK_OK;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const Box b = choose_box(N, H, W, 64, true);
  SEGK_REQUIRE(ctx, b.rows == 64, "conv2d_wgrad: cannot tile %dx%dx%d into 64-pixel boxes (need N*H*W >= 64)", N, H, W);
  const int block_n = pick_block_n(ctx, Cout);
  TensorMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc = encode_act_map(ctx, &maps.a[0], x, N, H, W, Cin, Cin, (int64_t)W * Cin, (int64_t)H * W * Cin, b.bw, b.bh, b.bn);
  if (rc) return rc;
  maps.a[1] = maps.a[2] = maps.a[3] = maps.a[0];
  rc = encode_nhwc_map(ctx, &maps.b, dy, N, H, W, Cout, dy_ld, b.bw, b.bh, b.bn);
  if (rc) return rc;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W;
  p.bw = b.bw; p.bh = b.bh; p.bn = b.bn;
  p.tiles_w = ceil_div(W, b.bw); p.tiles_h = ceil_div(H, b.bh); p.tiles_n = ceil_div(N, b.bn);
  const int n_ptiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.kchunks_in = Cin / 64;
  p.n_rb = kh * kw * p.kchunks_in;
  p.n_rbp = (p.n_rb + 1) / 2;
  p.n_tiles = Cout / block_n;
  const int base_items = p.n_rbp * p.n_tiles;
  // measured (profiles/, sweep): one wave of items is best once there are >= 16 base items (the
  // partial buffers and ordered reduction of extra splits cost more than the tail); tiny item
  // counts want two waves
  int splits = base_items >= 16 ? ctx->sm_count / base_items : ceil_div(2 * ctx->sm_count, base_items);
  if (ctx->force_wsplit > 0) splits = ctx->force_wsplit;
  const int max_splits = ceil_div(n_ptiles, 8);  // at least 8 k-steps per item
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  // make every split non-empty
  const int per_split = ceil_div(n_ptiles, splits);
  splits = ceil_div(n_ptiles, per_split);
  p.splits = splits;
  p.Cin_total = Cin; p.Cout_total = Cout;
  p.dw_tap_stride = Cin * Cout; p.dw_row_stride = Cout; p.dw_col_stride = 1;
  p.dw = dw;
  p.direct = (!accumulate && p.splits == 1) ? 1 : 0;
  // several splits (or accumulate): every split stores its partial sums into its own buffer and an
  // ordered reduction adds them up -- deterministic, and far cheaper than scattered fp32 atomics
  const size_t n_dw = (size_t)kh * kw * Cin * Cout;
  const bool partials = !p.direct && n_dw % 4 == 0 && sizeof(float) * n_dw * p.splits <= ((size_t)1 << 30);
  if (partials) {
    rc = segk_ws4(ctx, sizeof(float) * n_dw * p.splits);
    if (rc) return rc;
    p.dw = (float*)ctx->ws4;
    p.part_stride = (int64_t)n_dw;
    p.direct = 1;
  }
  SEGK_REQUIRE(ctx, p.direct, "conv2d_wgrad: %d splits of %zu weights exceed the 1 GiB partial-sum workspace", p.splits, n_dw);
  TapTable taps;
  conv_taps(taps, kh, kw, rate);
  // (pixel box, tap) pairs that only see SAME padding carry no information: skip them when there are any
  p.skip_oob = (rate > 1 || active_taps_1d(W, b.bw, kw) * active_taps_1d(H, b.bh, kh) < (int64_t)p.tiles_w * p.tiles_h * kh * kw) ? 1 : 0;
  const int total = p.splits * p.n_rbp * p.n_tiles;
  const int grid = total < ctx->sm_count ? total : ctx->sm_count;
  if (p.skip_oob && p.n_rbp <= kMaxRbp && p.splits == 1 && n_ptiles < 65536) {
    // cost of a row-block pair = pixel boxes (x-tile, y-tile) that either of its taps can see
    std::vector<std::pair<int, int>> cost(p.n_rbp);
    for (int r = 0; r < p.n_rbp; ++r) {
      const int rb0 = 2 * r, rb1 = (2 * r + 1 < p.n_rb) ? 2 * r + 1 : 2 * r;
      const int t0 = rb0 / p.kchunks_in, t1 = rb1 / p.kchunks_in;
      int act = 0;
      for (int ty = 0; ty < p.tiles_h; ++ty)
        for (int tx = 0; tx < p.tiles_w; ++tx) {
          const int x0 = tx * b.bw, y0 = ty * b.bh;
          const int xa = x0 + taps.dx[t0], ya = y0 + taps.dy[t0], xb = x0 + taps.dx[t1], yb = y0 + taps.dy[t1];
          const bool a = !(ya + b.bh <= 0 || ya >= H || xa + b.bw <= 0 || xa >= W);
          const bool bb = !(yb + b.bh <= 0 || yb >= H || xb + b.bw <= 0 || xb >= W);
          act += (a || bb) ? 1 : 0;
        }
      cost[r] = std::make_pair(-act, r);
    }
    std::sort(cost.begin(), cost.end());
    for (int r = 0; r < p.n_rbp; ++r) {
      p.rbp_perm[r] = (uint16_t)cost[r].second;
      p.rbp_nact[r] = (uint16_t)(-cost[r].first * p.tiles_n);
    }
    p.use_perm = 1;
  }
  switch (block_n) {
    case 256: rc = launch_wgrad_t<256>(ctx, maps, p, taps, grid, st); break;
    case 128: rc = launch_wgrad_t<128>(ctx, maps, p, taps, grid, st); break;
    default: rc = launch_wgrad_t<64>(ctx, maps, p, taps, grid, st); break;
  }
  if (rc) return rc;
  if (partials) return segk_reduce_partials(ctx, (const float*)ctx->ws4, dw, n_dw, p.splits, accumulate, st);
  return SEGK_OK;
}

int segk_conv2d_wgrad(segk_ctx* ctx, const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout,
                      int kh, int kw, int accumulate, void* stream) {
  return conv2d_wgrad_impl(ctx, x, dy, dw, N, H, W, Cin, Cout, kh, kw, accumulate, stream, 1);
}

// Atrous_Conv2D_Layer (utils.py:210-231: tf.nn.atrous_conv2d(x, W, rate, SAME)): the same implicit GEMM with the taps
// `rate` pixels apart; forward, Conv2DBackpropInput and Conv2DBackpropFilter.
int segk_atrous_conv2d_fwd(segk_ctx* ctx, const void* x, const void* wk, const float* bias, const void* residual, void* y,
                           int N, int H, int W, int Cin, int Cout, int kh, int kw, int rate, unsigned flags, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  return conv_igemm(ctx, "atrous_conv2d_fwd", x, wk, bias, residual, nullptr, 1.f, (flags & SEGK_EPI_RELU) ? 1 : 0,
                    (flags & SEGK_EPI_OUT_F32) ? 1 : 0, y, N, H, W, Cin, Cout, kh, kw, stream, nullptr, rate);
}

int segk_atrous_conv2d_dgrad(segk_ctx* ctx, const void* dy, const void* wd, const void* relu_mask, const void* residual,
                             void* dx, float* dx_colsum, float scale, int N, int H, int W, int Cin, int Cout, int kh, int kw,
                             int rate, void* stream) {
  if (!ctx) return SEGK_EINVAL;
  return conv_igemm(ctx, "atrous_conv2d_dgrad", dy, wd, nullptr, residual, relu_mask, scale, 0, 0, dx, N, H, W, Cout, Cin, kh,
                    kw, stream, dx_colsum, rate);
}

int segk_atrous_conv2d_wgrad(segk_ctx* ctx, const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout,
                             int kh, int kw, int rate, int accumulate, void* stream) {
  return conv2d_wgrad_impl(ctx, x, dy, dw, N, H, W, Cin, Cout, kh, kw, accumulate, stream, rate);
}

}  // extern "C"

// one-time per-context setup of the tensor-core kernels: opt in to > 48 KB dynamic shared memory for
// every instantiation and read the tuning overrides (called from segk_create)
int segk_tc_init(segk_ctx* ctx) {
  ctx->force_bn = env_int("SEGK_FORCE_BN");
  ctx->force_ksplit = env_int("SEGK_FORCE_KSPLIT");
  ctx->force_wsplit = env_int("SEGK_FORCE_WSPLIT");
  ctx->slab_mode = env_int("SEGK_SLAB", 1);
  ctx->tma_store = env_int("SEGK_TMA_STORE", 1);
  ctx->slab3 = env_int("SEGK_SLAB3", 1);
  ctx->teamk = env_int("SEGK_TEAMK", 1);
  ctx->hybrid = env_int("SEGK_HYBRID", 0);
  ctx->pair = env_int("SEGK_PAIR", 1);
  cudaError_t e = cudaSuccess;
#define SEGK_SMEM_ATTR(kern, bytes) \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)
  SEGK_SMEM_ATTR(igemm_kernel<64>, Cfg<64>::kSmemBytes);
  SEGK_SMEM_ATTR(igemm_kernel<128>, Cfg<128>::kSmemBytes);
  SEGK_SMEM_ATTR(igemm_kernel<256>, Cfg<256>::kSmemBytes);
  SEGK_SMEM_ATTR(igemm_pair_kernel, Cfg<256>::kSmemBytes);
  SEGK_SMEM_ATTR(wgrad_kernel<64>, WCfg<64>::kSmemBytes);
  SEGK_SMEM_ATTR(wgrad_kernel<128>, WCfg<128>::kSmemBytes);
  SEGK_SMEM_ATTR(wgrad_kernel<256>, WCfg<256>::kSmemBytes);
  SEGK_SMEM_ATTR(wgrad_pair_kernel, WCfg<256>::kSmemBytes);
  SEGK_SMEM_ATTR(slab_kernel<64>, SlabCfg<64>::kSmemBytes);
  SEGK_SMEM_ATTR(slab_kernel<128>, SlabCfg<128>::kSmemBytes);
  SEGK_SMEM_ATTR(slab_kernel<256>, SlabCfg<256>::kSmemBytes);
  SEGK_SMEM_ATTR(slab_pair_kernel<128>, SlabCfg<128>::kSmemBytes);
  SEGK_SMEM_ATTR(slab3_kernel<1>, Slab3Cfg<1>::kSmem);
  SEGK_SMEM_ATTR(slab3_kernel<2>, Slab3Cfg<2>::kSmem);
#undef SEGK_SMEM_ATTR
  if (e != cudaSuccess) return segk_fail(ctx, SEGK_ECUDA, "tensor-core kernel setup: %s", cudaGetErrorString(e));
  int rc = segk_first_init(ctx);
  if (rc) return rc;
  return segk_wslab_init(ctx);
}
