// Conv2DBackpropFilter for 3x3 layers on large maps with few input channels (conv1_2, conv2_x,
// conv3_1; FCN.py:340), "slab" formulation.
//
// ncu (profiles/r1d): the tap-wise wgrad_kernel is bound by shared-memory fill, not by the tensor
// pipe, on these layers: every (tap pair, pixel box) item re-loads x and dy from L2 (conv1_2: 120 KB of
// TMA writes per 64 pixels for 16 KB of distinct data; tensor pipe 38 %).  Here a CTA owns a range of
// 4-row x 32-column pixel tiles and, per tile, loads ONE haloed x slab (6 rows x 34 columns x 64
// channels) and ONE dy tile; every tap is a start-row offset into that slab:
//
//   dW[ky,kx][ci][co] += sum_{i<4, j<32}  xslab[(i+ky)*34 + j+kx][ci] * dy[i*32 + j][co]
//
// Both operands are MN-major (rows = pixels = GEMM-K); one tcgen05.mma covers 16 pixels = half an
// image row, so the x and dy start rows of an MMA are independent and the two pitches (34 / 32) never
// have to agree.  M = 128 = two 64-row blocks: two taps of the same slab (LBO = their row distance) or,
// for Cin = 128, the two channel chunks of one tap (LBO = distance of the two slabs).  All accumulators
// of an item (<= 512 TMEM columns) stay resident over its whole pixel range; per-split partial sums go
// to a workspace and are added up by wslab_reduce_kernel (deterministic, no atomics).
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace {

using namespace tc;

constexpr int kThreads = 192;              // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int kXP = 34;                    // haloed x slab pitch: 32 + 2 columns
constexpr int kDyP = 32, kDyH = 4;
constexpr int kDyBytes = kDyP * kDyH * 128;   // 16 KB per 64-channel block
constexpr int kMaxPairs = 8;
constexpr int kMaxGroups = 3;
constexpr int kSmemBudget = 208 * 1024;
constexpr uint64_t kDescMNBase = (64ull << 32) | (1ull << 46) | (2ull << 61);   // SBO 1024, v1, SWIZZLE_128B; + LBO

struct WsPair {
  int a_row;     // slab row of block 0 at (i = 0, h = 0): ky*34 + kx
  int a_chunk;   // which 64-channel slab block 0 reads
  int lbo16;     // distance block 0 -> block 1 in 16-byte units
  int dst0;      // dW row (tap*Cin + channel offset) of block 0
  int dst1;      // ... of block 1, -1 = unused (odd tap count)
};

struct WsParams {
  int N, H, W;
  int tiles_w, tiles_h;
  int splits, per_split, groups, n_tiles, npairs;
  int gy[kMaxGroups];         // first x row of the group's slab relative to y0 - 1 (filter row of a ky group)
  WsPair pairs[kMaxGroups][kMaxPairs];
  int Cout, rows_total;       // dW is [rows_total = 9*Cin][Cout]
  float* out;                 // dW itself (splits == 1) or the partial-sum workspace [splits][rows_total][Cout]
};

struct WsMaps {
  CUtensorMap x;    // box (64 ch, 34, 6 or 4, 1)
  CUtensorMap dy;   // box (64 ch, 32, 4, 1)
};

struct Pipe {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int stages) {
    if (++stage == stages) { stage = 0; phase ^= 1; }
  }
};

template <int BLOCK_N, int CHUNKS>
struct WsCfg {
  // CHUNKS == 1: one item holds all nine taps -> 6 slab rows; CHUNKS == 2: one filter row per item -> 4
  static constexpr int kXH = CHUNKS == 1 ? 6 : 4;
  static constexpr int kXLoadBytes = kXP * kXH * 128;                   // written by TMA
  static constexpr int kXBytes = (kXLoadBytes + 1023) / 1024 * 1024;    // slot; the pad rows feed discarded M rows only
  static constexpr int kDyBlocks = BLOCK_N / 64;
  static constexpr int kStageBytes = CHUNKS * kXBytes + kDyBlocks * kDyBytes;
  static constexpr int kStages = kSmemBudget / kStageBytes >= 4 ? 4 : kSmemBudget / kStageBytes;
  static constexpr uint32_t kTxBytes = CHUNKS * kXLoadBytes + kDyBlocks * kDyBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int BLOCK_N, int CHUNKS>
__global__ void __launch_bounds__(kThreads, 1)
wslab_kernel(const __grid_constant__ WsMaps maps, const __grid_constant__ WsParams p) {
  using C = WsCfg<BLOCK_N, CHUNKS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_ptiles = p.N * p.tiles_h * p.tiles_w;
  const int total_items = p.splits * p.groups * p.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.dy);
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 128);
    fence_barrier_init();
  }
  // the pad rows behind each x slab are read (into discarded accumulator rows) but never written
  for (int i = threadIdx.x; i < C::kStages * C::kStageBytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    Pipe ps;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int nt = item % p.n_tiles;
      const int gy = p.gy[(item / p.n_tiles) % p.groups];
      const int split = item / (p.n_tiles * p.groups);
      const int pt0 = split * p.per_split, pt1 = min(pt0 + p.per_split, n_ptiles);
      int x0 = (pt0 % p.tiles_w) * kDyP;
      int y0 = ((pt0 / p.tiles_w) % p.tiles_h) * kDyH;
      int n = pt0 / (p.tiles_w * p.tiles_h);
      for (int pt = pt0; pt < pt1; ++pt) {
        mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
        if (elect_one()) {
          uint8_t* s = smem + ps.stage * C::kStageBytes;
          mbar_arrive_expect_tx(&full_bar[ps.stage], C::kTxBytes);
#pragma unroll
          for (int c = 0; c < CHUNKS; ++c)
            tma_load_4d(&maps.x, &full_bar[ps.stage], s + c * C::kXBytes, c * 64, x0 - 1, y0 - 1 + gy, n);
#pragma unroll
          for (int j = 0; j < C::kDyBlocks; ++j)
            tma_load_4d(&maps.dy, &full_bar[ps.stage], s + CHUNKS * C::kXBytes + j * kDyBytes, nt * BLOCK_N + j * 64, x0, y0, n);
        }
        __syncwarp();
        ps.advance(C::kStages);
        x0 += kDyP;
        if (x0 >= p.tiles_w * kDyP) {
          x0 = 0;
          y0 += kDyH;
          if (y0 >= p.tiles_h * kDyH) { y0 = 0; ++n; }
        }
      }
    }
  } else if (warp == 1) {
    Pipe ps;
    uint32_t item_phase = 0;
    constexpr uint32_t idesc = make_idesc(128, BLOCK_N, 1, 1);
    const uint32_t smem_lo = smem_u32(smem) >> 4;
    constexpr uint64_t bdesc_hi = kDescMNBase | ((uint64_t)(kDyBytes >> 4) << 16);   // 64-channel dy blocks 16 KB apart
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int grp = (item / p.n_tiles) % p.groups;
      const int split = item / (p.n_tiles * p.groups);
      const int pt0 = split * p.per_split, pt1 = min(pt0 + p.per_split, n_ptiles);
      mbar_wait(tempty_bar, item_phase ^ 1);
      tc_fence_after();
      for (int pt = pt0; pt < pt1; ++pt) {
        mbar_wait(&full_bar[ps.stage], ps.phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t x_lo = smem_lo + (uint32_t)ps.stage * (uint32_t)(C::kStageBytes >> 4);
          const uint32_t dy_lo = x_lo + (uint32_t)((CHUNKS * C::kXBytes) >> 4);
          for (int pr = 0; pr < p.npairs; ++pr) {
            const WsPair& w = p.pairs[grp][pr];
            const uint64_t adesc_hi = kDescMNBase | ((uint64_t)(uint32_t)w.lbo16 << 16);
            const uint32_t a_base = x_lo + (uint32_t)w.a_chunk * (uint32_t)(C::kXBytes >> 4) + (uint32_t)w.a_row * 8u;
            const uint32_t d_addr = tmem_base + (uint32_t)(pr * BLOCK_N);
#pragma unroll
            for (int i = 0; i < kDyH; ++i)
#pragma unroll
              for (int h = 0; h < 2; ++h)     // 16 pixels (half an image row) per MMA; one row = 8 descriptor units
                umma_f16(d_addr, adesc_hi | (uint64_t)(a_base + (uint32_t)(i * kXP + 16 * h) * 8u),
                         bdesc_hi | (uint64_t)(dy_lo + (uint32_t)(i * kDyP + 16 * h) * 8u), idesc,
                         (pt != pt0 || i != 0 || h != 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[ps.stage]);
          if (pt == pt1 - 1) umma_commit(tfull_bar);
        }
        __syncwarp();
        ps.advance(C::kStages);
      }
      item_phase ^= 1;
    }
  } else {
    const int q = warp & 3;
    uint32_t item_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int nt = item % p.n_tiles;
      const int grp = (item / p.n_tiles) % p.groups;
      const int split = item / (p.n_tiles * p.groups);
      mbar_wait(tfull_bar, item_phase);
      tc_fence_after();
      float* base = p.out + (int64_t)split * p.rows_total * p.Cout + (int64_t)nt * BLOCK_N;
      for (int pr = 0; pr < p.npairs; ++pr) {
        const WsPair& w = p.pairs[grp][pr];
        const int d = (q >> 1) ? w.dst1 : w.dst0;              // warp-uniform
        if (d < 0) continue;
        float* dst = base + (int64_t)(d + (q & 1) * 32 + lane) * p.Cout;
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(pr * BLOCK_N + c0), r);
          tmem_ld_wait();
          float4* d4 = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            d4[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                __uint_as_float(r[4 * i + 3]));
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar);
      item_phase ^= 1;
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// dw[i] (+)= sum over splits part[s][i]; 4 floats per thread
__global__ void __launch_bounds__(256) wslab_reduce_kernel(const float4* __restrict__ part, float4* __restrict__ dw,
                                                          int n4, int splits, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = accumulate ? dw[i] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int s = 0; s < splits; ++s) {
    const float4 v = part[(int64_t)s * n4 + i];
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  dw[i] = a;
}

template <int BLOCK_N, int CHUNKS>
int launch(segk_ctx* ctx, const WsMaps& maps, const WsParams& p, int grid, cudaStream_t st) {
  wslab_kernel<BLOCK_N, CHUNKS><<<grid, kThreads, WsCfg<BLOCK_N, CHUNKS>::kSmemBytes, st>>>(maps, p);
  SEGK_LAUNCHED(ctx, "wslab");
  return SEGK_OK;
}

}  // namespace

int segk_ws4(segk_ctx* ctx, size_t bytes) { return segk_grow(ctx, &ctx->ws4, &ctx->ws4_bytes, bytes, "wgrad partial sums"); }

int segk_reduce_partials(segk_ctx* ctx, const float* part, float* dw, size_t n, int splits, int accumulate, cudaStream_t st) {
  const int n4 = (int)(n / 4);
  wslab_reduce_kernel<<<ceil_div(n4, 256), 256, 0, st>>>((const float4*)part, (float4*)dw, n4, splits, accumulate);
  SEGK_LAUNCHED(ctx, "wgrad partial-sum reduce");
  return SEGK_OK;
}

// -> 1 handled, 0 not applicable (caller uses the tap-wise kernel), < 0 error
int segk_wslab_try(segk_ctx* ctx, const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout, int kh,
                   int kw, int accumulate, void* stream, int dy_ld) {
  if (!ctx->wslab) return 0;
  if (!(kh == 3 && kw == 3 && (Cin == 64 || Cin == 128) && Cout % 64 == 0)) return 0;
  if (ctx->wslab == 1 && !((int64_t)H * W >= 4096 && W >= 64)) return 0;
  const int chunks = Cin / 64;
  const int block_n = (chunks == 2 && Cout % 128 == 0) ? 128 : 64;
  WsParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W;
  p.tiles_w = ceil_div(W, kDyP); p.tiles_h = ceil_div(H, kDyH);
  p.n_tiles = Cout / block_n;
  p.Cout = Cout; p.rows_total = 9 * Cin;
  if (chunks == 1) {
    // all nine taps in one item: pairs of taps out of the same slab, the ninth alone
    p.groups = 1; p.npairs = 5;
    for (int pr = 0; pr < 5; ++pr) {
      const int t0 = 2 * pr, t1 = 2 * pr + 1;
      const int r0 = (t0 / 3) * kXP + t0 % 3;
      WsPair& w = p.pairs[0][pr];
      w.a_row = r0; w.a_chunk = 0; w.dst0 = t0 * Cin;
      if (t1 < 9) {
        const int r1 = (t1 / 3) * kXP + t1 % 3;
        w.lbo16 = (r1 - r0) * 8; w.dst1 = t1 * Cin;
      } else {
        w.lbo16 = 8; w.dst1 = -1;
      }
    }
  } else {
    // one item per filter row: pair = (tap, both channel chunks)
    p.groups = 3; p.npairs = 3;
    for (int ky = 0; ky < 3; ++ky) {
      p.gy[ky] = ky;
      for (int kx = 0; kx < 3; ++kx) {
        WsPair& w = p.pairs[ky][kx];
        const int t = ky * 3 + kx;
        w.a_row = kx; w.a_chunk = 0; w.lbo16 = WsCfg<64, 2>::kXBytes >> 4;
        w.dst0 = t * Cin; w.dst1 = t * Cin + 64;
      }
    }
  }
  const int n_ptiles = N * p.tiles_h * p.tiles_w;
  const int base_items = p.groups * p.n_tiles;
  int splits = ctx->force_wsplit > 0 ? ctx->force_wsplit : ctx->sm_count / base_items;
  if (splits > n_ptiles) splits = n_ptiles;
  if (splits < 1) splits = 1;
  p.per_split = ceil_div(n_ptiles, splits);
  p.splits = ceil_div(n_ptiles, p.per_split);
  WsMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc = tch::act_map(ctx, &maps.x, x, N, H, W, Cin, kXP, chunks == 1 ? 6 : 4, 1);
  if (rc) return rc;
  rc = tch::act_map(ctx, &maps.dy, dy, N, H, W, Cout, kDyP, kDyH, 1, dy_ld);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)p.rows_total * Cout;
  const bool direct = p.splits == 1 && !accumulate;
  if (direct) {
    p.out = dw;
  } else {
    rc = segk_ws4(ctx, sizeof(float) * n * p.splits);
    if (rc) return rc;
    p.out = (float*)ctx->ws4;
  }
  const int total = p.splits * base_items;
  const int grid = total < ctx->sm_count ? total : ctx->sm_count;
  if (chunks == 1) rc = launch<64, 1>(ctx, maps, p, grid, st);
  else if (block_n == 128) rc = launch<128, 2>(ctx, maps, p, grid, st);
  else rc = launch<64, 2>(ctx, maps, p, grid, st);
  if (rc) return rc;
  if (!direct) {
    rc = segk_reduce_partials(ctx, p.out, dw, n, p.splits, accumulate, st);
    if (rc) return rc;
  }
  return 1;
}

int segk_wslab_init(segk_ctx* ctx) {
  ctx->wslab = 1;
  const char* v = getenv("SEGK_WSLAB");
  if (v) ctx->wslab = atoi(v);
  cudaError_t e = cudaFuncSetAttribute(wslab_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WsCfg<64, 1>::kSmemBytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(wslab_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WsCfg<128, 2>::kSmemBytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(wslab_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WsCfg<64, 2>::kSmemBytes);
  if (e != cudaSuccess) return segk_fail(ctx, SEGK_ECUDA, "wslab kernel setup: %s", cudaGetErrorString(e));
  return SEGK_OK;
}
