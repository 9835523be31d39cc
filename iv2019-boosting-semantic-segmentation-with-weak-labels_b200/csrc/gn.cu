// Group normalisation, `--norm_layer group` (code/models/resnet50_extended_model_hierarchical.py:314-333:
// normalizer_fn = tf.contrib.layers.group_norm with groups = 32, or 1 for the logits layers (:75-77), epsilon
// 1e-5, scale and centre on, applied by slim.conv2d between the convolution and its activation).
//
// [TF-1.12] group_norm over NHWC: per sample n and group g the mean / biased variance of the H*W*(C/G) values
// (nn.moments), y = (x - mean) * rsqrt(var + eps) * gamma[c] + beta[c].  There are no moving statistics: the
// layer does the same thing in training and inference.
//
// Division of labour with bn.cu (the per-sample passes ARE batch-norm passes over one sample's H*W rows):
//   per-(n, c) sums / sums of squares          wlseg_bn_stats        on the sample's slice
//   -> per-(n, c) scale / shift / mean / invstd  wlseg_gn_finalize     (here)
//   y = relu?(z * scale + shift (+ residual))     wlseg_bn_apply        on the sample's slice
//   per-(n, c) sum g * xhat, sum g               wlseg_bn_bwd_reduce   on the sample's slice
//   -> per-(n, c) A, c1, c0 and dgamma, dbeta      wlseg_gn_bwd_finalize (here)
//   dz = A * g + c1 * z + c0                       wlseg_gn_bwd_apply    (here, whole batch)
// This is the non-default normaliser; the kernels here are simple element-wise / tiny-reduction code.
#include "common.cuh"

namespace wlseg {

// one thread per (sample, group)
__global__ void gn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sqsum, int N, int C, int G,
                                   int64_t hw, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean,
                                   float* __restrict__ invstd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * G) return;
  const int n = i / G, g = i % G;
  const int gs = C / G;
  double s = 0.0, q = 0.0;
  for (int c = g * gs; c < (g + 1) * gs; ++c) { s += sum[n * C + c]; q += sqsum[n * C + c]; }
  const double m = (double)hw * gs;
  const double mu = s / m;
  double var = q / m - mu * mu;
  if (var < 0.0) var = 0.0;
  const float muf = (float)mu;
  const float is = rsqrtf((float)var + eps);
  for (int c = g * gs; c < (g + 1) * gs; ++c) {
    const float sc = gamma[c] * is;
    scale[n * C + c] = sc;
    shift[n * C + c] = beta[c] - muf * sc;
    mean[n * C + c] = muf;
    invstd[n * C + c] = is;
  }
}

// y = gamma * xhat + beta, xhat = (z - mu_g) * is_g.  With gh = gamma_c * g (g = dL/dy after the ReLU mask) and
// M1 = mean_group(gh), M2 = mean_group(gh * xhat) over the group's H*W*(C/G) values:
//   dz = is * (gh - M1 - xhat * M2) = A * g + c1 * z + c0,  A = gamma_c * is, c1 = -is^2 * M2, c0 = -is * M1 - c1 * mu
// The sums over the pixels come per channel from bn_bwd_reduce (sum g * xhat -> dgam_nc, sum g -> dbet_nc).
// Threads [0, N*G) do the groups; threads [0, C) then add the per-sample sums into the parameter gradients.
__global__ void gn_bwd_finalize_kernel(const double* __restrict__ dgam_nc, const double* __restrict__ dbet_nc, int N, int C,
                                       int G, int64_t hw, const float* __restrict__ gamma, const float* __restrict__ mean,
                                       const float* __restrict__ invstd, float* __restrict__ cA, float* __restrict__ c1,
                                       float* __restrict__ c0, double* __restrict__ dgamma, double* __restrict__ dbeta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N * G) {
    const int n = i / G, g = i % G;
    const int gs = C / G;
    double m1 = 0.0, m2 = 0.0;
    for (int c = g * gs; c < (g + 1) * gs; ++c) {
      m1 += (double)gamma[c] * dbet_nc[n * C + c];
      m2 += (double)gamma[c] * dgam_nc[n * C + c];
    }
    const double m = (double)hw * gs;
    m1 /= m;
    m2 /= m;
    for (int c = g * gs; c < (g + 1) * gs; ++c) {
      const float is = invstd[n * C + c], mu = mean[n * C + c];
      const float k1 = -is * is * (float)m2;
      cA[n * C + c] = gamma[c] * is;
      c1[n * C + c] = k1;
      c0[n * C + c] = -is * (float)m1 - k1 * mu;
    }
  }
  if (i < C) {
    double a = 0.0, b = 0.0;
    for (int n = 0; n < N; ++n) { a += dgam_nc[n * C + i]; b += dbet_nc[n * C + i]; }
    dgamma[i] += a;
    dbeta[i] += b;
  }
}

// one thread per element (any C); the ReLU mask comes from yact or, when it is NULL, from the sign of
// fmaf(z, scale, shift) - the exact fp32 value the forward pass rounded to y
template <typename T>
__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ yact, const T* __restrict__ z,
                    const float* __restrict__ cA, const float* __restrict__ c1, const float* __restrict__ c0,
                    const float* __restrict__ scale, const float* __restrict__ shift, int64_t total, int64_t hw, int C,
                    int relu, T* __restrict__ dz, T* __restrict__ dres) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int n = (int)(i / ((int64_t)C * hw));
    const int k = n * C + c;
    float g = to_f32<T>(dy[i]);
    const float zz = to_f32<T>(z[i]);
    if (relu) {
      const float yy = yact != nullptr ? to_f32<T>(yact[i]) : fmaf(zz, scale[k], shift[k]);
      if (!(yy > 0.f)) g = 0.f;
    }
    if (dres != nullptr) dres[i] = from_f32<T>(g);
    dz[i] = from_f32<T>(fmaf(cA[k], g, fmaf(c1[k], zz, c0[k])));
  }
}

}  // namespace wlseg

using namespace wlseg;

extern "C" int wlseg_gn_finalize(const double* sum, const double* sqsum, int32_t N, int32_t C, int32_t groups, int64_t hw,
                                 const float* gamma, const float* beta, float eps, float* scale, float* shift, float* mean,
                                 float* invstd, wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(N > 0 && C > 0 && groups > 0 && C % groups == 0 && hw > 0,
                  "gn_finalize: bad shape (N %d, C %d, groups %d)", N, C, groups);
  WLSEG_CHECK_ARG(sum && sqsum && gamma && beta && scale && shift && mean && invstd, "gn_finalize: null pointer");
  const int n = N * groups;
  gn_finalize_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sum, sqsum, N, C, groups, hw, gamma, beta, eps, scale,
                                                                        shift, mean, invstd);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_gn_bwd_finalize(const double* dgamma_nc, const double* dbeta_nc, int32_t N, int32_t C, int32_t groups,
                                     int64_t hw, const float* gamma, const float* mean, const float* invstd, float* cA,
                                     float* c1, float* c0, double* dgamma, double* dbeta, wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(N > 0 && C > 0 && groups > 0 && C % groups == 0 && hw > 0,
                  "gn_bwd_finalize: bad shape (N %d, C %d, groups %d)", N, C, groups);
  WLSEG_CHECK_ARG(dgamma_nc && dbeta_nc && gamma && mean && invstd && cA && c1 && c0 && dgamma && dbeta,
                  "gn_bwd_finalize: null pointer");
  const int n = N * groups > C ? N * groups : C;
  gn_bwd_finalize_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dgamma_nc, dbeta_nc, N, C, groups, hw, gamma, mean,
                                                                            invstd, cA, c1, c0, dgamma, dbeta);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_gn_bwd_apply(const void* dy, const void* y, const void* z, const float* cA, const float* c1,
                                  const float* c0, const float* scale, const float* shift, int32_t N, int64_t hw, int32_t C,
                                  int32_t relu, int32_t dtype, void* dz, void* dres, wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(N >= 0 && hw > 0 && C > 0, "gn_bwd_apply: bad shape");
  if (N == 0) return 0;
  WLSEG_CHECK_ARG(dy && z && cA && c1 && c0 && dz && (!relu || y || (scale && shift)), "gn_bwd_apply: null pointer");
  const int64_t total = (int64_t)N * hw * C;
  const int grid = bw_grid(total, 256, 8);
  if (dtype == WLSEG_BF16)
    gn_bwd_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y,
                                                                (const __nv_bfloat16*)z, cA, c1, c0, scale, shift, total, hw, C,
                                                                relu, (__nv_bfloat16*)dz, (__nv_bfloat16*)dres);
  else if (dtype == WLSEG_F32)
    gn_bwd_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)dy, (const float*)y, (const float*)z, cA, c1, c0,
                                                                scale, shift, total, hw, C, relu, (float*)dz, (float*)dres);
  else
    WLSEG_CHECK_ARG(false, "gn_bwd_apply: bad dtype %d", dtype);
  WLSEG_LAUNCH_CHECK();
  return 0;
}
