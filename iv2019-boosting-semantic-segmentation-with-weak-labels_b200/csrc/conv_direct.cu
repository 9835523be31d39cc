// Direct (SIMT) convolution: fprop / dgrad / wgrad for ANY stride / dilation / padding / channel
// count, fp32 or bf16 storage with fp32 accumulation.
//
// This is the fp32 "check mode" path (north star: logits within 1e-4 of the reference in fp32)
// and the path for the few layer shapes the tcgen05 implicit-GEMM kernel does not take
// (see conv_igemm_sm100.cu).  It is hand-written CUDA like everything else in this library, not
// a CPU or library fallback; it is simply not the fast path.
//
// Replaces slim.conv2d / conv2d_same and the TF Conv2DBackpropInput / Conv2DBackpropFilter ops
// (call sites: code/models/resnet50_extended_feature_extractor.py:25-30,39-43;
// code/models/resnet50_extended_model_hierarchical.py:60-64,80).
#include "common.cuh"

namespace wlseg {

constexpr int kKT = 8;  // output channels per thread (fprop), input channels per thread (dgrad)

template <typename T, typename TY>
__global__ void __launch_bounds__(128)
conv_fprop_direct_kernel(const wlseg_conv_params p, const T* __restrict__ x, const T* __restrict__ w,
                         TY* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                         const T* __restrict__ res) {
  const int kgroups = (p.K + kKT - 1) / kKT;
  const int64_t total = (int64_t)p.N * p.P * p.Q * kgroups;
  const bool vec = (p.C % 8 == 0) && (p.x_pitch % 8 == 0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int kg = (int)(i % kgroups);
    int64_t t = i / kgroups;
    const int q = (int)(t % p.Q); t /= p.Q;
    const int pp = (int)(t % p.P);
    const int n = (int)(t / p.P);
    const int k0 = kg * kKT;
    float acc[kKT];
#pragma unroll
    for (int j = 0; j < kKT; ++j) acc[j] = 0.f;
    for (int r = 0; r < p.R; ++r) {
      const int hh = pp * p.stride - p.pad_top + r * p.dilation;
      if (hh < 0 || hh >= p.H) continue;
      for (int s = 0; s < p.S; ++s) {
        const int ww = q * p.stride - p.pad_left + s * p.dilation;
        if (ww < 0 || ww >= p.W) continue;
        const T* xp = x + (((int64_t)n * p.H + hh) * p.W + ww) * p.x_pitch;
        const T* wp = w + (((int64_t)k0 * p.R + r) * p.S + s) * p.C;
        const int64_t wk = (int64_t)p.R * p.S * p.C;  // stride between output channels
        if (vec) {
          for (int c = 0; c < p.C; c += 8) {
            float xv[8];
            Vec8<T> vx;
            vx.load(xp + c);
            vx.unpack(xv);
#pragma unroll
            for (int j = 0; j < kKT; ++j) {
              if (k0 + j < p.K) {
                float wv[8];
                Vec8<T> vw;
                vw.load(wp + j * wk + c);
                vw.unpack(wv);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[j] += xv[e] * wv[e];
              }
            }
          }
        } else {
          for (int c = 0; c < p.C; ++c) {
            const float xv = to_f32<T>(xp[c]);
#pragma unroll
            for (int j = 0; j < kKT; ++j)
              if (k0 + j < p.K) acc[j] += xv * to_f32<T>(wp[j * wk + c]);
          }
        }
      }
    }
    const int64_t opix = ((int64_t)n * p.P + pp) * p.Q + q;
    const T* rp = nullptr;
    if (res != nullptr)
      rp = res + (((int64_t)n * p.res_H + (int64_t)pp * p.res_stride) * p.res_W + (int64_t)q * p.res_stride) * p.res_pitch;
#pragma unroll
    for (int j = 0; j < kKT; ++j) {
      const int k = k0 + j;
      if (k >= p.K) break;
      float v = acc[j];
      if (scale) v *= scale[k];
      if (shift) v += shift[k];
      if (rp) v += to_f32<T>(rp[k]);
      if (p.relu) v = fmaxf(v, 0.f);
      y[opix * p.y_pitch + k] = from_f32<TY>(v);
    }
  }
}

// dx[n,h,w,c] = sum_{r,s,k} dy[n,p,q,k] * w[k,r,s,c] with h = p*stride - pad + r*dil
template <typename T>
__global__ void __launch_bounds__(128)
conv_dgrad_direct_kernel(const wlseg_conv_params p, const T* __restrict__ dy, const T* __restrict__ w,
                         T* __restrict__ dx) {
  const int cgroups = (p.C + kKT - 1) / kKT;
  const int64_t total = (int64_t)p.N * p.H * p.W * cgroups;
  const bool vec = (p.C % 8 == 0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cgroups);
    int64_t t = i / cgroups;
    const int ww = (int)(t % p.W); t /= p.W;
    const int hh = (int)(t % p.H);
    const int n = (int)(t / p.H);
    const int c0 = cg * kKT;
    float acc[kKT];
#pragma unroll
    for (int j = 0; j < kKT; ++j) acc[j] = 0.f;
    for (int r = 0; r < p.R; ++r) {
      const int ph = hh + p.pad_top - r * p.dilation;
      if (ph < 0 || ph % p.stride != 0) continue;
      const int pp = ph / p.stride;
      if (pp >= p.P) continue;
      for (int s = 0; s < p.S; ++s) {
        const int qw = ww + p.pad_left - s * p.dilation;
        if (qw < 0 || qw % p.stride != 0) continue;
        const int q = qw / p.stride;
        if (q >= p.Q) continue;
        const T* dyp = dy + (((int64_t)n * p.P + pp) * p.Q + q) * p.y_pitch;
        for (int k = 0; k < p.K; ++k) {
          const float g = to_f32<T>(dyp[k]);
          const T* wp = w + (((int64_t)k * p.R + r) * p.S + s) * p.C + c0;
          if (vec) {
            float wv[8];
            Vec8<T> vw;
            vw.load(wp);
            vw.unpack(wv);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += g * wv[j];
          } else {
#pragma unroll
            for (int j = 0; j < kKT; ++j)
              if (c0 + j < p.C) acc[j] += g * to_f32<T>(wp[j]);
          }
        }
      }
    }
    T* o = dx + (((int64_t)n * p.H + hh) * p.W + ww) * p.x_pitch + c0;
#pragma unroll
    for (int j = 0; j < kKT; ++j)
      if (c0 + j < p.C) o[j] = from_f32<T>(acc[j]);
  }
}

// dw[k,r,s,c] += sum over a chunk of output pixels of dy[.,k] * x[.,c]; grid.y = pixel chunks
template <typename T>
__global__ void __launch_bounds__(128)
conv_wgrad_direct_kernel(const wlseg_conv_params p, const T* __restrict__ x, const T* __restrict__ dy,
                         float* __restrict__ dw, int pixels_per_chunk) {
  const int64_t total = (int64_t)p.K * p.R * p.S * p.C;
  const int64_t npix = (int64_t)p.N * p.P * p.Q;
  const int64_t pix0 = (int64_t)blockIdx.y * pixels_per_chunk;
  const int64_t pix1 = min(npix, pix0 + pixels_per_chunk);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % p.C);
    int64_t t = i / p.C;
    const int s = (int)(t % p.S); t /= p.S;
    const int r = (int)(t % p.R);
    const int k = (int)(t / p.R);
    float acc = 0.f;
    for (int64_t pix = pix0; pix < pix1; ++pix) {
      const int q = (int)(pix % p.Q);
      const int64_t t2 = pix / p.Q;
      const int pp = (int)(t2 % p.P);
      const int n = (int)(t2 / p.P);
      const int hh = pp * p.stride - p.pad_top + r * p.dilation;
      const int ww = q * p.stride - p.pad_left + s * p.dilation;
      if (hh < 0 || hh >= p.H || ww < 0 || ww >= p.W) continue;
      acc += to_f32<T>(dy[pix * p.y_pitch + k]) * to_f32<T>(x[(((int64_t)n * p.H + hh) * p.W + ww) * p.x_pitch + c]);
    }
    atomicAdd(dw + i, acc);
  }
}

int conv_wgrad_tcgen05(const wlseg_conv_params* p, const void* x, const void* dy, float* dw, cudaStream_t s);
bool conv_wgrad_tcgen05_supported(const wlseg_conv_params* p);

int check_conv_params(const wlseg_conv_params* p) {
  WLSEG_CHECK_ARG(p != nullptr, "conv: params is NULL");
  WLSEG_CHECK_ARG(p->N >= 0 && p->H > 0 && p->W > 0 && p->C > 0 && p->K > 0 && p->R > 0 && p->S > 0 && p->P > 0 &&
                      p->Q > 0,
                  "conv: bad shape N=%d H=%d W=%d C=%d K=%d R=%d S=%d P=%d Q=%d", p->N, p->H, p->W, p->C, p->K, p->R,
                  p->S, p->P, p->Q);
  WLSEG_CHECK_ARG(p->stride > 0 && p->dilation > 0 && p->pad_top >= 0 && p->pad_left >= 0, "conv: bad stride/dilation/pad");
  WLSEG_CHECK_ARG(p->x_pitch >= p->C && p->y_pitch >= p->K, "conv: pitch smaller than channel count");
  WLSEG_CHECK_ARG(p->dtype == WLSEG_F32 || p->dtype == WLSEG_BF16, "conv: bad dtype %d", p->dtype);
  WLSEG_CHECK_ARG(p->y_dtype == WLSEG_F32 || p->y_dtype == p->dtype, "conv: y_dtype must be fp32 or equal dtype");
  // the last output must start inside the (padded) input
  WLSEG_CHECK_ARG((p->P - 1) * p->stride - p->pad_top < p->H && (p->Q - 1) * p->stride - p->pad_left < p->W,
                  "conv: output size inconsistent with input size");
  return 0;
}

int conv_fprop_direct(const wlseg_conv_params* p, const void* x, const void* w, void* y, const float* scale,
                      const float* shift, const void* residual, cudaStream_t s) {
  const int kgroups = (p->K + kKT - 1) / kKT;
  const int64_t total = (int64_t)p->N * p->P * p->Q * kgroups;
  const int grid = bw_grid(total, 128, 16);
  if (p->dtype == WLSEG_BF16 && p->y_dtype == WLSEG_BF16)
    conv_fprop_direct_kernel<<<grid, 128, 0, s>>>(*p, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w,
                                                  (__nv_bfloat16*)y, scale, shift, (const __nv_bfloat16*)residual);
  else if (p->dtype == WLSEG_BF16)
    conv_fprop_direct_kernel<<<grid, 128, 0, s>>>(*p, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (float*)y,
                                                  scale, shift, (const __nv_bfloat16*)residual);
  else
    conv_fprop_direct_kernel<<<grid, 128, 0, s>>>(*p, (const float*)x, (const float*)w, (float*)y, scale, shift,
                                                  (const float*)residual);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

}  // namespace wlseg

using namespace wlseg;

extern "C" int wlseg_conv2d_dgrad(const wlseg_conv_params* p, const void* dy, const void* w, void* dx,
                                  wlseg_stream_t stream) {
  if (int e = check_conv_params(p)) return e;
  if (p->N == 0) return 0;
  WLSEG_CHECK_ARG(dy && w && dx, "conv_dgrad: null pointer");
  WLSEG_CHECK_ARG(p->y_dtype == p->dtype, "conv_dgrad: dy must be stored in dtype");
  const int cgroups = (p->C + kKT - 1) / kKT;
  const int64_t total = (int64_t)p->N * p->H * p->W * cgroups;
  const int grid = bw_grid(total, 128, 16);
  if (p->dtype == WLSEG_BF16)
    conv_dgrad_direct_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(*p, (const __nv_bfloat16*)dy,
                                                                     (const __nv_bfloat16*)w, (__nv_bfloat16*)dx);
  else
    conv_dgrad_direct_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(*p, (const float*)dy, (const float*)w, (float*)dx);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_conv2d_wgrad(const wlseg_conv_params* p, const void* x, const void* dy, float* dw,
                                  wlseg_stream_t stream) {
  if (int e = check_conv_params(p)) return e;
  WLSEG_CHECK_ARG(dw != nullptr, "conv_wgrad: dw is NULL");
  const int64_t total = (int64_t)p->K * p->R * p->S * p->C;
  if (!p->accumulate) WLSEG_CUDA(cudaMemsetAsync(dw, 0, total * sizeof(float), (cudaStream_t)stream));
  if (p->N == 0) return 0;
  WLSEG_CHECK_ARG(x && dy, "conv_wgrad: null pointer");
  WLSEG_CHECK_ARG(p->y_dtype == p->dtype, "conv_wgrad: dy must be stored in dtype");
  {
    int algo = p->algo;
    if (algo == WLSEG_ALGO_AUTO) algo = conv_wgrad_tcgen05_supported(p) ? WLSEG_ALGO_TCGEN05 : WLSEG_ALGO_DIRECT;
    if (algo == WLSEG_ALGO_TCGEN05) {
      WLSEG_CHECK_ARG(conv_wgrad_tcgen05_supported(p), "conv_wgrad: configuration not covered by the tcgen05 kernel");
      return conv_wgrad_tcgen05(p, x, dy, dw, (cudaStream_t)stream);
    }
  }
  const int64_t npix = (int64_t)p->N * p->P * p->Q;
  int gx = bw_grid(total, 128, 4);
  // enough pixel chunks to fill the machine a few times over
  int chunks = (int)ceil_div((int64_t)kNumSMs * 16, gx);
  if (chunks > npix) chunks = (int)npix;
  if (chunks > 65535) chunks = 65535;
  if (chunks < 1) chunks = 1;
  int per = (int)ceil_div(npix, chunks);
  chunks = (int)ceil_div(npix, per);
  dim3 grid(gx, chunks);
  if (p->dtype == WLSEG_BF16)
    conv_wgrad_direct_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(*p, (const __nv_bfloat16*)x,
                                                                     (const __nv_bfloat16*)dy, dw, per);
  else
    conv_wgrad_direct_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(*p, (const float*)x, (const float*)dy, dw, per);
  WLSEG_LAUNCH_CHECK();
  return 0;
}
