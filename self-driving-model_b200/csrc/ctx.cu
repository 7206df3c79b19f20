// Context, error reporting and small utility entry points of the C-ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void amoe_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int amoe_conv_tc_init(amoe_ctx* ctx);    // conv_tc.cu
int amoe_conv_flat_init(amoe_ctx* ctx);  // conv_flat.cu
int amoe_stem_init(amoe_ctx* ctx);       // stem_tc.cu
int amoe_wgrad_tc_init(amoe_ctx* ctx);   // wgrad_tc.cu

extern "C" {

int amoe_abi_version(void) { return 1; }

const char* amoe_last_error(void) { return g_err; }

int amoe_create(int device, amoe_ctx** out) {
  AMOE_REQUIRE(out != nullptr, "amoe_create: out is NULL");
  *out = nullptr;
  AMOE_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  AMOE_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  AMOE_REQUIRE(prop.major == 10, "amoe_create: device %d is sm_%d%d; this library is sm_100a only",
               device, prop.major, prop.minor);
  amoe_ctx* ctx = new amoe_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->launches.store(0);
  ctx->walk_reverse = 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    delete ctx;
    amoe_set_error("amoe_create: cuTensorMapEncodeTiled not available (%s)", cudaGetErrorString(e));
    return -1;
  }
  ctx->encode_tiled = reinterpret_cast<decltype(ctx->encode_tiled)>(fn);
  if (amoe_conv_tc_init(ctx) != 0 || amoe_conv_flat_init(ctx) != 0 || amoe_stem_init(ctx) != 0 ||
      amoe_wgrad_tc_init(ctx) != 0) {
    delete ctx;
    return -1;
  }
  *out = ctx;
  return 0;
}

int amoe_set_walk_reverse(amoe_ctx* ctx, int reverse) {
  AMOE_REQUIRE(ctx != nullptr, "amoe_set_walk_reverse: NULL ctx");
  ctx->walk_reverse = reverse ? 1 : 0;
  return 0;
}

int amoe_destroy(amoe_ctx* ctx) {
  delete ctx;
  return 0;
}

int amoe_sm_count(amoe_ctx* ctx) { return ctx ? ctx->sm_count : -1; }

int64_t amoe_launch_count(amoe_ctx* ctx) { return ctx ? ctx->launches.load() : -1; }

}  // extern "C"
