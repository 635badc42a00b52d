// Context management for libsegk.so.
#include "common.cuh"

extern "C" {

int segk_abi_version(void) { return SEGK_ABI_VERSION; }

int segk_create(int device, segk_ctx** out) {
  if (!out) return SEGK_EINVAL;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || device < 0 || device >= ndev) return SEGK_ECUDA;
  segk_ctx* ctx = new (std::nothrow) segk_ctx();
  if (!ctx) return SEGK_ENOMEM;
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete ctx;
    return SEGK_ECUDA;
  }
  if (prop.major != 10) {  // sm_100a only: no fallback code path exists
    delete ctx;
    return SEGK_EINVAL;
  }
  ctx->sm_count = prop.multiProcessorCount;
  int prev_device = -1;
  cudaGetDevice(&prev_device);        // the kernel attributes below are per device; restored before returning
  cudaSetDevice(device);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    if (prev_device >= 0 && prev_device != device) cudaSetDevice(prev_device);
    delete ctx;
    return SEGK_ECUDA;
  }
  ctx->encode_tiled = reinterpret_cast<decltype(ctx->encode_tiled)>(fn);
  const int rc_init = segk_tc_init(ctx);
  if (prev_device >= 0 && prev_device != device) cudaSetDevice(prev_device);
  if (rc_init != SEGK_OK) {
    delete ctx;
    return SEGK_ECUDA;
  }
  *out = ctx;
  return SEGK_OK;
}

int segk_destroy(segk_ctx* ctx) {
  if (ctx && ctx->ws) cudaFree(ctx->ws);
  if (ctx && ctx->ws2) cudaFree(ctx->ws2);
  if (ctx && ctx->ws3) cudaFree(ctx->ws3);
  if (ctx && ctx->ws4) cudaFree(ctx->ws4);
  if (ctx && ctx->ws5) cudaFree(ctx->ws5);
  if (ctx && ctx->ws6) cudaFree(ctx->ws6);
  if (ctx && ctx->ws7) cudaFree(ctx->ws7);
  delete ctx;
  return SEGK_OK;
}

const char* segk_last_error(segk_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }

int segk_set_tuning(segk_ctx* ctx, const char* key, int value) {
  if (!ctx || !key) return SEGK_EINVAL;
  if (!strcmp(key, "slab")) ctx->slab_mode = value;
  else if (!strcmp(key, "tma_store")) ctx->tma_store = value;
  else if (!strcmp(key, "slab3")) ctx->slab3 = value;
  else if (!strcmp(key, "wslab")) ctx->wslab = value;
  else if (!strcmp(key, "teamk")) ctx->teamk = value;
  else if (!strcmp(key, "hybrid")) ctx->hybrid = value;
  else if (!strcmp(key, "pair")) ctx->pair = value;
  else if (!strcmp(key, "tail_wide")) ctx->tail_wide = value;
  else if (!strcmp(key, "force_bn")) ctx->force_bn = value;
  else if (!strcmp(key, "force_ksplit")) ctx->force_ksplit = value;
  else if (!strcmp(key, "force_wsplit")) ctx->force_wsplit = value;
  else return segk_fail(ctx, SEGK_EINVAL, "set_tuning: unknown key '%s'", key);
  return SEGK_OK;
}

int segk_set_pitch(segk_ctx* ctx, int in_pitch, int out_pitch) {
  if (!ctx) return SEGK_EINVAL;
  SEGK_REQUIRE(ctx, in_pitch >= 0 && out_pitch >= 0 && in_pitch % 8 == 0 && out_pitch % 8 == 0,
               "set_pitch: pitches are channel counts, multiples of 8 (got %d, %d)", in_pitch, out_pitch);
  ctx->pitch_in = in_pitch;
  ctx->pitch_out = out_pitch;
  return SEGK_OK;
}

int64_t segk_launch_count(segk_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

int segk_sm_count(segk_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

}  // extern "C"
