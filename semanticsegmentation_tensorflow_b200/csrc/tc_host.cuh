// Host-side helpers shared by the tensor-core translation units (defined in tcconv.cu).
#pragma once

#include "common.cuh"

struct Box {
  int bw, bh, bn, rows, tiles;
};

namespace tch {
// pixel box bw x bh x bn of at most (exactly, when `exact`) max_rows rows covering [N,H,W]
Box pick_box(int N, int H, int W, int max_rows, bool exact);
// 4-D NHWC bf16 tensor map (C, W, H, N), box (64, bw, bh, bn), SWIZZLE_128B, zero fill outside
// (ld = channels between consecutive pixels; 0 = dense, ld = C)
int act_map(segk_ctx* ctx, CUtensorMap* m, const void* base, int N, int H, int W, int C, int bw, int bh, int bn, int ld = 0);
// 3-D weight map [T][Nrows][K] bf16, box (64, block_n, 1), SWIZZLE_128B
int weight_map(segk_ctx* ctx, CUtensorMap* m, const void* base, int K, int Nrows, int T, int block_n);
// grow-only context scratch (ctx->ws)
int workspace(segk_ctx* ctx, size_t bytes);
}  // namespace tch

// grow-only scratch for per-split weight-gradient partial sums (ctx->ws4) and their ordered reduction:
// dw[i] (+)= sum_s part[s*n + i]   (n % 4 == 0)
int segk_ws4(segk_ctx* ctx, size_t bytes);
int segk_reduce_partials(segk_ctx* ctx, const float* part, float* dw, size_t n, int splits, int accumulate, cudaStream_t st);

int segk_first_init(segk_ctx* ctx);
int segk_wslab_init(segk_ctx* ctx);   // wslab.cu
// slab-formulated Conv2DBackpropFilter: 1 = handled, 0 = not applicable, < 0 = error
int segk_wslab_try(segk_ctx* ctx, const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout, int kh,
                   int kw, int accumulate, void* stream, int dy_ld = 0);   // firstconv.cu: shared-memory opt-in of its kernels
