// Shared helpers for libwlseg (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/wlseg.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libwlseg is written for sm_100a (B200) only"
#endif

namespace wlseg {

// thread-local error text behind wlseg_last_error()
void set_error(const char* fmt, ...);

#define WLSEG_CHECK_ARG(cond, ...)             \
  do {                                         \
    if (!(cond)) {                             \
      ::wlseg::set_error(__VA_ARGS__);         \
      return -1;                               \
    }                                          \
  } while (0)

#define WLSEG_CUDA(expr)                                                            \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::wlseg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                         __FILE__, __LINE__);                                       \
      return (int)_e;                                                               \
    }                                                                               \
  } while (0)

#define WLSEG_LAUNCH_CHECK()                                                        \
  do {                                                                              \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      ::wlseg::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),\
                         __FILE__, __LINE__);                                       \
      return (int)_e;                                                               \
    }                                                                               \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// SMs the persistent tensor-core kernels (one 200+ KB CTA per SM, statically strided tiles) spread over.  Data-parallel
// training sets WLSEG_CONV_SMS=144 (wlseg/trainer.py): an NCCL all-reduce kernel overlapping backward needs SMs of its
// own - its CTAs cannot co-reside with a convolution CTA (shared memory), and a convolution CTA that has to wait for an
// SM delays the whole statically scheduled grid by the length of the collective (measured: +4-7 % per step).
inline int conv_sms() {
  const char* e = getenv("WLSEG_CONV_SMS");
  if (e == nullptr) return kNumSMs;
  int v = atoi(e);
  v -= v % 2;
  return v < 2 ? 2 : (v > kNumSMs ? kNumSMs : v);
}

// ---- storage type helpers (fp32 / bf16), arithmetic always in fp32 ----
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// 8 consecutive channels as one 16-byte (bf16) or two 16-byte (fp32) accesses
template <typename T> struct Vec8;
template <> struct Vec8<__nv_bfloat16> {
  uint4 raw;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ __forceinline__ void pack(const float (&f)[8]) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  }
};
template <> struct Vec8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = a;
    *reinterpret_cast<float4*>(p + 4) = b;
  }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  __device__ __forceinline__ void pack(const float (&f)[8]) {
    a = make_float4(f[0], f[1], f[2], f[3]);
    b = make_float4(f[4], f[5], f[6], f[7]);
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------
// One training step is ~520 dependent launches of 10-50 us kernels; without PDL every kernel boundary costs the
// launch latency plus the ramp-up of the next grid.  Kernels on the hot chain (convolutions, batch-norm passes)
// call pdl_launch_dependents() first thing - the NEXT kernel's CTAs may then become resident and run their
// prologue (barrier init, TMEM allocation, constant loads) as SM resources free up - and pdl_wait() before
// their first access to global memory, which returns once the PREVIOUS kernel has completed and flushed.
// Both are no-ops when the launch does not carry the attribute - the default: see pdl_enabled() in abi.cu for
// the measurement that made this opt-in (WLSEG_PDL=1).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();   // abi.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Division by a launch-time constant without the ~25-instruction dependent chain (I2F, MUFU.RCP, F2I, fix-ups) the
// compiler emits for a runtime divisor: q = (umulhi(n, mul) + n) >> shift, exact for 0 <= n < 2^31, 1 <= d < 2^31.
// The persistent tile loops decode (image, row, column, channel-tile) coordinates in single elected threads - the TMA
// producer, the residual cursors of the epilogue warps - where that chain sat on the critical path: ncu's source view
// attributed 38 % of an epilogue step of the conv3 + residual layers to it (profiles/r2_epilogue_divisions.md).
struct FastDiv {
  uint32_t mul, shift, d;
  __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
#ifdef __CUDA_ARCH__
    return (__umulhi(n, mul) + n) >> shift;
#else
    return (uint32_t)((((uint64_t)n * mul) >> 32) + n) >> shift;
#endif
  }
  __host__ __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
    q = div(n);
    r = n - q * d;
  }
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  if (d <= 1) { f.mul = 0; f.shift = 0; f.d = 1; return f; }
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  f.shift = l;
  f.mul = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << l) - d)) / d + 1);
  return f;
}

// grid size for grid-stride bandwidth kernels: a multiple of the SM count
inline int bw_grid(int64_t work_items, int threads, int ctas_per_sm) {
  int64_t need = ceil_div(work_items, threads);
  int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
  if (need >= cap) return (int)cap;
  return (int)(need < 1 ? 1 : need);
}

}  // namespace wlseg
