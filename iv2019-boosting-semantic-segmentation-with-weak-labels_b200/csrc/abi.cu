// Version / error plumbing and small layout helpers of the C ABI.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace wlseg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Opt-in (WLSEG_PDL=1): measured on B200 the training step got SLOWER with it, 12.12 -> 12.53 ms (the
// early-resident, waiting CTAs of the batch-norm kernels take thread / register slots from the running grid),
// and the evaluation step did not move (9.00 vs 9.02 ms: its 200 KB persistent convolution CTAs cannot
// co-reside anyway).  Read per launch so that one process can A/B both modes.
bool pdl_enabled() { return getenv("WLSEG_PDL") != nullptr; }

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = __float2bfloat16_rn(src[i]);
}

__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = __bfloat162float(src[i]);
}

// dst[c][R-1-r][S-1-s][k] = src[k][r][s][c]
template <typename T>
__global__ void transpose_flip_kernel(const T* __restrict__ src, T* __restrict__ dst, int K, int R, int S, int C) {
  int64_t n = (int64_t)K * R * S * C;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    // i indexes dst: [c][r'][s'][k]
    int k = (int)(i % K);
    int64_t t = i / K;
    int s2 = (int)(t % S); t /= S;
    int r2 = (int)(t % R);
    int c = (int)(t / R);
    int r = R - 1 - r2, s = S - 1 - s2;
    dst[i] = src[(((int64_t)k * R + r) * S + s) * C + c];
  }
}

}  // namespace wlseg

using namespace wlseg;

extern "C" int wlseg_version(void) { return WLSEG_VERSION; }
extern "C" const char* wlseg_last_error(void) { return g_err; }

extern "C" int wlseg_cast_f32_to_bf16(const float* src, void* dst, int64_t n, wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(src && dst && n >= 0, "cast: null pointer");
  if (n == 0) return 0;
  cast_f32_bf16_kernel<<<bw_grid(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_cast_bf16_to_f32(const void* src, float* dst, int64_t n, wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(src && dst && n >= 0, "cast: null pointer");
  if (n == 0) return 0;
  cast_bf16_f32_kernel<<<bw_grid(n, 256, 8), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, dst, n);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_weights_transpose_flip(const void* src, void* dst, int32_t K, int32_t R, int32_t S,
                                            int32_t C, int32_t dtype, wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(src && dst && K > 0 && R > 0 && S > 0 && C > 0, "weights_transpose_flip: bad args");
  int64_t n = (int64_t)K * R * S * C;
  int grid = bw_grid(n, 256, 8);
  if (dtype == WLSEG_BF16)
    transpose_flip_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, K, R, S, C);
  else if (dtype == WLSEG_F32)
    transpose_flip_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src, (float*)dst, K, R, S, C);
  else
    WLSEG_CHECK_ARG(false, "weights_transpose_flip: bad dtype %d", dtype);
  WLSEG_LAUNCH_CHECK();
  return 0;
}
