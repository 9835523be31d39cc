// Fused SGD / momentum / Nesterov step with L2 weight decay over one flat parameter arena.
//
// Replaces tf.train.MomentumOptimizer / GradientDescentOptimizer (ApplyMomentum, ~200 tiny
// launches in the reference) from code/estimator/define_optimizer.py:17-22 plus the gradient and
// value of slim.l2_regularizer (code/models/resnet50_extended_model_hierarchical.py:336,
// code/estimator/define_losses_hierarchical.py:205).
//
// HBM-bound: per parameter reads w, g, acc (12 B) and writes w, acc (+2 B bf16 operand copy).
// The host keeps every trainable tensor in ONE fp32 arena with the conv kernels first, so the
// weight-decay mask is a single split index and one launch updates all 26 M parameters.
#include "common.cuh"

namespace wlseg {

__global__ void __launch_bounds__(256)
sgdm_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ acc,
            __nv_bfloat16* __restrict__ wb, int64_t n, int64_t n_decay, const float* __restrict__ lr_dev,
            float momentum, int nesterov, float wd, float grad_scale, double* __restrict__ reg_loss) {
  const float lr = __ldg(lr_dev);
  float sq = 0.f;
  const int64_t n4 = n >> 2;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n4; v += (int64_t)gridDim.x * blockDim.x) {
    float4 w4 = reinterpret_cast<float4*>(w)[v];
    const float4 g4 = reinterpret_cast<const float4*>(g)[v];
    float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (acc) a4 = reinterpret_cast<float4*>(acc)[v];
    float ww[4] = {w4.x, w4.y, w4.z, w4.w};
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
    float aa[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool decay = (v * 4 + j) < n_decay;
      float gr = gg[j] * grad_scale;
      if (decay) { gr += wd * ww[j]; sq += ww[j] * ww[j]; }
      if (acc) {
        aa[j] = momentum * aa[j] + gr;
        ww[j] -= nesterov ? lr * (gr + momentum * aa[j]) : lr * aa[j];
      } else {
        ww[j] -= lr * gr;
      }
    }
    reinterpret_cast<float4*>(w)[v] = make_float4(ww[0], ww[1], ww[2], ww[3]);
    if (acc) reinterpret_cast<float4*>(acc)[v] = make_float4(aa[0], aa[1], aa[2], aa[3]);
    if (wb) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(ww[0], ww[1]);
      __nv_bfloat162 hi = __floats2bfloat162_rn(ww[2], ww[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(wb)[v] = pk;
    }
  }
  // tail (< 4 elements)
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t i = (n4 << 2) + threadIdx.x;
    float gr = g[i] * grad_scale;
    float wv = w[i];
    if (i < n_decay) { gr += wd * wv; sq += wv * wv; }
    if (acc) {
      float a = momentum * acc[i] + gr;
      acc[i] = a;
      wv -= nesterov ? lr * (gr + momentum * a) : lr * a;
    } else {
      wv -= lr * gr;
    }
    w[i] = wv;
    if (wb) wb[i] = __float2bfloat16_rn(wv);
  }
  if (reg_loss != nullptr) {
    double s = warp_sum((double)sq);
    __shared__ double part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < 8; ++i) t += part[i];
      if (t != 0.0) atomicAdd(reg_loss, 0.5 * (double)wd * t);
    }
  }
}

// shadow update of tf.train.ExponentialMovingAverage(decay, num_updates, zero_debias=True)
// [TF-1.12 assign_moving_average/_zero_debias]: biased <- biased - (1-d)(biased - w);
// unbiased = biased / (1 - d^local_step), where d = min(decay, (1+t)/(10+t)).
__global__ void __launch_bounds__(256)
ema_kernel(float* biased, float* shadow, const float* __restrict__ w, int64_t n,
           float d, float inv_correction) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float b = biased[i];
    b -= (1.0f - d) * (b - w[i]);
    biased[i] = b;
    shadow[i] = b * inv_correction;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) add_inplace_kernel(T* __restrict__ dst, const T* __restrict__ src, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float a[8], b[8];
    Vec8<T> va, vb;
    va.load(dst + i * 8);
    vb.load(src + i * 8);
    va.unpack(a);
    vb.unpack(b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    va.pack(a);
    va.store(dst + i * 8);
  }
}

}  // namespace wlseg

using namespace wlseg;

extern "C" int wlseg_ema_update(float* biased, float* shadow, const float* w, int64_t n, float decay,
                                float inv_correction, wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(n >= 0 && (n == 0 || (biased && shadow && w)), "ema_update: bad args");
  if (n == 0) return 0;
  ema_kernel<<<bw_grid(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(biased, shadow, w, n, decay, inv_correction);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_add_inplace(void* dst, const void* src, int64_t n, int32_t dtype, wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(n >= 0 && n % 8 == 0, "add_inplace: n (%lld) must be a multiple of 8", (long long)n);
  if (n == 0) return 0;
  WLSEG_CHECK_ARG(dst && src, "add_inplace: null pointer");
  int grid = bw_grid(n / 8, 256, 8);
  if (dtype == WLSEG_BF16)
    add_inplace_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)src, n / 8);
  else if (dtype == WLSEG_F32)
    add_inplace_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((float*)dst, (const float*)src, n / 8);
  else
    WLSEG_CHECK_ARG(false, "add_inplace: bad dtype %d", dtype);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_sgdm_step(float* w, const float* g, float* acc, void* w_bf16, int64_t n, int64_t n_decay,
                               const float* lr_dev, float momentum, int32_t nesterov, float wd, float grad_scale,
                               double* reg_loss, wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(n >= 0 && n_decay >= 0 && n_decay <= n, "sgdm: bad sizes n=%lld n_decay=%lld", (long long)n,
                  (long long)n_decay);
  if (n == 0) return 0;
  WLSEG_CHECK_ARG(w && g && lr_dev, "sgdm: null pointer");
  WLSEG_CHECK_ARG(((((uintptr_t)w) | ((uintptr_t)g) | ((uintptr_t)acc)) & 15) == 0 && (((uintptr_t)w_bf16) & 7) == 0,
                  "sgdm: arenas must be 16-byte aligned");
  int grid = bw_grid((n >> 2) + 1, 256, 8);
  sgdm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, g, acc, (__nv_bfloat16*)w_bf16, n, n_decay, lr_dev, momentum,
                                                      nesterov, wd, grad_scale, reg_loss);
  WLSEG_LAUNCH_CHECK();
  return 0;
}
