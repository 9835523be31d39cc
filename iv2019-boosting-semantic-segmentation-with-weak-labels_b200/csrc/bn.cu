// Training-mode batch normalisation around the tensor-core convolutions.
//
// Replaces tf.contrib.layers.batch_norm (FusedBatchNorm / FusedBatchNormGrad) configured by
// code/models/resnet50_extended_model_hierarchical.py:298-312,325: batch mean / biased variance
// over N*H*W, eps 1e-5, moving statistics updated with `decay` using the UNBIASED variance.
// In inference mode BN is folded into the convolution epilogue (scale/shift) and none of these
// kernels run.
//
// All kernels are HBM-bound passes over [count x C] NHWC activations, 8 channels per thread.
// Statistics are reduced per thread in fp32 over a few rows, per CTA in shared memory, and
// across CTAs with fp64 global atomics (sums of up to 10^5..10^6 values per channel).
#include <cstdlib>

#include "common.cuh"

namespace wlseg {

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

constexpr int kBnThreads = 256;
constexpr int kBnUnroll = 4;

// thread (tx, ty): tx = channel-vector index within C/8 (<= 256), ty = row lane
template <typename T, bool kBackward>
__global__ void __launch_bounds__(kBnThreads)
bn_reduce_kernel(const T* __restrict__ a, const T* __restrict__ yact, const T* __restrict__ z,
                 const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ scale,
                 const float* __restrict__ shift, int64_t count, int C, int pitch, int relu,
                 double* __restrict__ out0, double* __restrict__ out1, int cspan, int rev) {
  // forward (kBackward=false): a = z;  out0 += sum z, out1 += sum z^2
  // backward: a = dy; g = dy*(y>0 if relu); out0 += sum g*(z-mean)*invstd (dgamma), out1 += sum g (dbeta)
  //   the ReLU mask comes from yact, or - when yact is NULL (layers without a residual input) - from the
  //   sign of fmaf(z, scale, shift), the exact fp32 value the forward pass rounded to y
  extern __shared__ float part[];  // [lanes][2][C]
  pdl_launch_dependents();
  pdl_wait();   // mean / invstd / dy all come from earlier kernels of the chain
  // channel split: blockIdx.y owns the channels [blockIdx.y * cspan, + cspan) of every row it visits, so a
  // channel's fp64 accumulator is hit by gridDim.x CTAs instead of by the whole grid (same-address L2 atomics
  // serialise); a warp still reads 16 * cspan / 8 >= 512 contiguous bytes of a row
  if (cspan < C) {
    const int cb = blockIdx.y * cspan;
    a += cb;
    if (yact != nullptr) yact += cb;
    if (z != nullptr) z += cb;
    if (kBackward) { mean += cb; invstd += cb; if (scale != nullptr) { scale += cb; shift += cb; } }
    out0 += cb;
    out1 += cb;
    C = cspan;
  }
  const int cv = C / 8;
  const int lanes = kBnThreads / cv;
  const int tx = threadIdx.x % cv, ty = threadIdx.x / cv;
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s0[j] = 0.f; s1[j] = 0.f; }
  float mu[8], is[8], sc[8], sh[8];
  const bool zmask = kBackward && relu && yact == nullptr;
  if (kBackward) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { mu[j] = mean[tx * 8 + j]; is[j] = invstd[tx * 8 + j]; }
    if (zmask) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { sc[j] = scale[tx * 8 + j]; sh[j] = shift[tx * 8 + j]; }
    }
  }
  if (ty < lanes) {
    // kBnUnroll independent rows per iteration: all their 16-byte loads are issued before any is used,
    // so a thread keeps up to 3 x kBnUnroll requests in flight (one row at a time left the kernel
    // latency-bound at 2.9 TB/s on the widest layers)
    const int64_t step = (int64_t)gridDim.x * lanes;
    for (int64_t row0 = (int64_t)blockIdx.x * lanes + ty; row0 < count; row0 += step * kBnUnroll) {
      Vec8<T> va[kBnUnroll], vzz[kBnUnroll], vyy[kBnUnroll];
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int64_t row = row0 + u * step;
        if (row < count) {
          // rev: walk the tensor from its END - the producer (dgrad, ascending tile order) has just written the
          // last rows, they are the ones still in L2
          const int64_t r = rev ? count - 1 - row : row;
          va[u].load(a + r * pitch + tx * 8);
          if (kBackward) {
            vzz[u].load(z + r * pitch + tx * 8);
            if (relu && !zmask) vyy[u].load(yact + r * pitch + tx * 8);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int64_t row = row0 + u * step;
        if (row >= count) break;
        float f[8];
        va[u].unpack(f);
        if (!kBackward) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { s0[j] += f[j]; s1[j] += f[j] * f[j]; }
        } else {
          float zz[8];
          vzz[u].unpack(zz);
          if (zmask) {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaf(zz[j], sc[j], sh[j]) > 0.f ? f[j] : 0.f;
          } else if (relu) {
            float yy[8];
            vyy[u].unpack(yy);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = yy[j] > 0.f ? f[j] : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) { s0[j] += f[j] * (zz[j] - mu[j]) * is[j]; s1[j] += f[j]; }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      part[(ty * 2 + 0) * C + tx * 8 + j] = s0[j];
      part[(ty * 2 + 1) * C + tx * 8 + j] = s1[j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kBnThreads) {
    double t0 = 0.0, t1 = 0.0;
    for (int l = 0; l < lanes; ++l) { t0 += (double)part[(l * 2 + 0) * C + c]; t1 += (double)part[(l * 2 + 1) * C + c]; }
    atomicAdd(out0 + c, t0);
    atomicAdd(out1 + c, t1);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sqsum, int64_t count,
                                   int C, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float decay, float moving_var_factor, float* __restrict__ moving_mean,
                                   float* __restrict__ moving_var, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ saved_mean, float* __restrict__ saved_invstd) {
  pdl_launch_dependents();
  pdl_wait();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double n = (double)count;
  double m = sum[c] / n;
  double var = sqsum[c] / n - m * m;
  if (var < 0.0) var = 0.0;
  float mf = (float)m, vf = (float)var;
  float inv = rsqrtf(vf + eps);
  float sc = gamma[c] * inv;
  if (scale) scale[c] = sc;
  if (shift) shift[c] = beta[c] - mf * sc;
  if (saved_mean) saved_mean[c] = mf;
  if (saved_invstd) saved_invstd[c] = inv;
  if (moving_mean) {
    // TF: moving <- moving - (1 - decay) * (moving - stat); variance with Bessel's correction, or - under
    // --cross_replica_norm - the biased GLOBAL variance times (n_local - 1) / n_local
    // (utils/cross_replica_batch_normalization.py:452-459), passed in as moving_var_factor >= 0
    float unbiased = moving_var_factor >= 0.f ? vf * moving_var_factor : (count > 1 ? (float)(var * (n / (n - 1.0))) : vf);
    moving_mean[c] -= (1.0f - decay) * (moving_mean[c] - mf);
    moving_var[c] -= (1.0f - decay) * (moving_var[c] - unbiased);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const T* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                const T* __restrict__ res, T* __restrict__ y, int64_t count, int C, int relu) {
  const int cv = C / 8;
  const int64_t total = count * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c0 = (int)(i % cv) * 8;
    float f[8];
    Vec8<T> v;
    v.load(z + i * 8);
    v.unpack(f);
    float4 sa = *reinterpret_cast<const float4*>(scale + c0), sb = *reinterpret_cast<const float4*>(scale + c0 + 4);
    float4 ha = *reinterpret_cast<const float4*>(shift + c0), hb = *reinterpret_cast<const float4*>(shift + c0 + 4);
    const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
    const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = f[j] * sc[j] + sh[j];
    if (res != nullptr) {
      float r[8];
      Vec8<T> vr;
      vr.load(res + i * 8);
      vr.unpack(r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += r[j];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    Vec8<T> o;
    o.pack(f);
    o.store(y + i * 8);
  }
}

// bn_finalize fused into bn_apply: every CTA derives scale / shift of all C channels from the fp64
// sums into shared memory (C <= 2048: a few flops per thread); CTA 0 also publishes scale / shift /
// saved mean / inverse std for the backward pass and updates the moving statistics.  One launch per
// layer less, and the apply loop reads its constants from shared memory.
template <typename T>
__global__ void __launch_bounds__(256)
bn_finalize_apply_kernel(const double* __restrict__ sum, const double* __restrict__ sqsum, int64_t count, int C,
                         const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float decay,
                         float* __restrict__ moving_mean, float* __restrict__ moving_var, float* __restrict__ scale,
                         float* __restrict__ shift, float* __restrict__ saved_mean, float* __restrict__ saved_invstd,
                         const T* __restrict__ z, const T* __restrict__ res, T* __restrict__ y, int relu) {
  extern __shared__ float sconst[];  // [2][C]: scale | shift
  float* ssc = sconst;
  float* ssh = sconst + C;
  const double n = (double)count;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double m = sum[c] / n;
    double var = sqsum[c] / n - m * m;
    if (var < 0.0) var = 0.0;
    const float mf = (float)m, vf = (float)var;
    const float inv = rsqrtf(vf + eps);
    const float sc = gamma[c] * inv;
    const float sh = beta[c] - mf * sc;
    ssc[c] = sc;
    ssh[c] = sh;
    if (blockIdx.x == 0) {
      scale[c] = sc;
      shift[c] = sh;
      saved_mean[c] = mf;
      saved_invstd[c] = inv;
      if (moving_mean != nullptr) {
        const float unbiased = count > 1 ? (float)(var * (n / (n - 1.0))) : vf;
        moving_mean[c] -= (1.0f - decay) * (moving_mean[c] - mf);
        moving_var[c] -= (1.0f - decay) * (moving_var[c] - unbiased);
      }
    }
  }
  __syncthreads();
  const int cv = C / 8;
  const int64_t total = count * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cv) * 8;
    float f[8];
    Vec8<T> v;
    v.load(z + i * 8);
    v.unpack(f);
    const float4 sa = *reinterpret_cast<const float4*>(ssc + c0), sb = *reinterpret_cast<const float4*>(ssc + c0 + 4);
    const float4 ha = *reinterpret_cast<const float4*>(ssh + c0), hb = *reinterpret_cast<const float4*>(ssh + c0 + 4);
    const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
    const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = f[j] * sc[j] + sh[j];
    if (res != nullptr) {
      float r[8];
      Vec8<T> vr;
      vr.load(res + i * 8);
      vr.unpack(r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += r[j];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    Vec8<T> o;
    o.pack(f);
    o.store(y + i * 8);
  }
}

// dz = gamma*invstd*(g - dbeta/n - zhat*dgamma/n) = A*g + c1*z + c0 with per-channel constants
//   A = gamma*invstd, c1 = -A*invstd*dgamma/n, c0 = -A*dbeta/n - c1*mean
// computed once per CTA into shared memory (the fp64 sums are touched C times, not count*C times).
template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ yact, const T* __restrict__ z,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                    const float* __restrict__ scale, const float* __restrict__ shift,
                    const double* __restrict__ dgamma, const double* __restrict__ dbeta, int64_t count,
                    int64_t stat_count, int C, int pitch, int relu, T* __restrict__ dz, T* __restrict__ dres) {
  extern __shared__ float bconst[];  // [5][C]: A | c1 | c0 | scale | shift
  float* sA = bconst;
  float* s1 = bconst + C;
  float* s0 = bconst + 2 * C;
  float* ssc = bconst + 3 * C;
  float* ssh = bconst + 4 * C;
  const bool zmask = relu && yact == nullptr;
  const double invn = 1.0 / (double)stat_count;   // pixels the sums were taken over (all replicas under sync BN)
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float is = invstd[c];
    const float A = gamma[c] * is;
    const float c1 = -A * is * (float)(dgamma[c] * invn);
    sA[c] = A;
    s1[c] = c1;
    s0[c] = -A * (float)(dbeta[c] * invn) - c1 * mean[c];
    if (zmask) { ssc[c] = scale[c]; ssh[c] = shift[c]; }
  }
  __syncthreads();
  const int cv = C / 8;
  const int64_t total = count * cv;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 1;  // U = 2 (6 loads in flight per thread) measured 6 % SLOWER: occupancy matters more here
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += step * U) {
    Vec8<T> v[U], vz[U], vy[U];
    int64_t e[U];
    int c0[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * step;
      if (i < total) {
        c0[u] = (int)(i % cv) * 8;
        e[u] = (i / cv) * pitch + c0[u];   // element offset: row * pitch + channel
        v[u].load(dy + e[u]);
        vz[u].load(z + e[u]);
        if (relu && !zmask) vy[u].load(yact + e[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u * step >= total) break;
      float g[8], zz[8];
      v[u].unpack(g);
      vz[u].unpack(zz);
      if (zmask) {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = fmaf(zz[j], ssc[c0[u] + j], ssh[c0[u] + j]) > 0.f ? g[j] : 0.f;
      } else if (relu) {
        float yy[8];
        vy[u].unpack(yy);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = yy[j] > 0.f ? g[j] : 0.f;
      }
      if (dres != nullptr) {
        Vec8<T> o;
        o.pack(g);
        o.store(dres + e[u]);
      }
      const float4 a0 = *reinterpret_cast<const float4*>(sA + c0[u]), a1 = *reinterpret_cast<const float4*>(sA + c0[u] + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(s1 + c0[u]), b1 = *reinterpret_cast<const float4*>(s1 + c0[u] + 4);
      const float4 d0 = *reinterpret_cast<const float4*>(s0 + c0[u]), d1 = *reinterpret_cast<const float4*>(s0 + c0[u] + 4);
      const float A[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float B[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const float D[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
      float out[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) out[j] = fmaf(A[j], g[j], fmaf(B[j], zz[j], D[j]));
      Vec8<T> o;
      o.pack(out);
      o.store(dz + e[u]);
    }
  }
}

// ---- row-mapped variants of bn_apply / bn_bwd_apply (C % 8 == 0, C <= 2048) ----------------------------
// thread (tx, ty) keeps ONE channel vector for its whole life: tx = vector index within the row (C / 8 of
// them), ty = row lane; a CTA walks `lanes` = 256 / (C / 8) consecutive rows per iteration.  The per-channel
// constants therefore live in registers (no shared-memory staging, no __syncthreads) and the element index
// is a multiply-add instead of the 64-bit divide + modulo per 16-byte vector of the flat-index kernels
// above; kRowsU independent rows are in flight per thread.  A warp still touches 512 contiguous bytes
// (C >= 256) or 32 * 16 contiguous bytes spanning consecutive rows (C < 256, dense rows).
// bn_finalize folded into the row-mapped apply kernel: a thread derives scale / shift of ITS 8 channels from the fp64
// sums (same arithmetic as bn_finalize_kernel) instead of reading them; the first row lane of CTA 0 publishes scale /
// shift / saved mean / inverse std for the backward pass and updates the moving statistics.  One launch less per
// layer (64 per training step at ~3 us each) without the per-CTA shared-memory staging of bn_finalize_apply_kernel.
struct BnFin {
  const double* sum;      // NULL: scale / shift are inputs (plain wlseg_bn_apply)
  const double* sqsum;
  const float* gamma;
  const float* beta;
  float* moving_mean;
  float* moving_var;
  float* scale_out;
  float* shift_out;
  float* saved_mean;
  float* saved_invstd;
  int64_t count;
  float eps, decay;
};

template <typename T, int kRowsU>
__global__ void __launch_bounds__(256)
bn_apply_rows_kernel(const T* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                     const T* __restrict__ res, T* __restrict__ y, int64_t count, int C, int relu, int rev,
                     uint8_t* __restrict__ relu_mask, const BnFin fin) {
  pdl_launch_dependents();
  pdl_wait();   // scale / shift (or the sums) were written by the kernel right before this one
  const int cv = C / 8;
  const int lanes = 256 / cv;
  const int tx = threadIdx.x % cv, ty = threadIdx.x / cv;
  if (ty >= lanes) return;
  const int c0 = tx * 8;
  float sc[8], sh[8];
  if (fin.sum != nullptr) {
    // same fp64 arithmetic as bn_finalize_kernel, bit for bit (multiplying by 1 / n instead was measured too: the 16
    // divisions cost ~7 us per launch, but the fused launch loses to two launches even without them)
    const double n = (double)fin.count;
    const bool publish = blockIdx.x == 0 && ty == 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      const double m = fin.sum[c] / n;
      double var = fin.sqsum[c] / n - m * m;
      if (var < 0.0) var = 0.0;
      const float mf = (float)m, vf = (float)var;
      const float inv = rsqrtf(vf + fin.eps);
      sc[j] = fin.gamma[c] * inv;
      sh[j] = fin.beta[c] - mf * sc[j];
      if (publish) {
        fin.scale_out[c] = sc[j];
        fin.shift_out[c] = sh[j];
        fin.saved_mean[c] = mf;
        fin.saved_invstd[c] = inv;
        if (fin.moving_mean != nullptr) {
          const float unbiased = fin.count > 1 ? (float)(var * (n / (n - 1.0))) : vf;
          fin.moving_mean[c] -= (1.0f - fin.decay) * (fin.moving_mean[c] - mf);
          fin.moving_var[c] -= (1.0f - fin.decay) * (fin.moving_var[c] - unbiased);
        }
      }
    }
  } else {
    const float4 sa = *reinterpret_cast<const float4*>(scale + c0), sb = *reinterpret_cast<const float4*>(scale + c0 + 4);
    const float4 ha = *reinterpret_cast<const float4*>(shift + c0), hb = *reinterpret_cast<const float4*>(shift + c0 + 4);
    sc[0] = sa.x; sc[1] = sa.y; sc[2] = sa.z; sc[3] = sa.w; sc[4] = sb.x; sc[5] = sb.y; sc[6] = sb.z; sc[7] = sb.w;
    sh[0] = ha.x; sh[1] = ha.y; sh[2] = ha.z; sh[3] = ha.w; sh[4] = hb.x; sh[5] = hb.y; sh[6] = hb.z; sh[7] = hb.w;
  }
  const int64_t step = (int64_t)gridDim.x * lanes;
  for (int64_t row0 = (int64_t)blockIdx.x * lanes + ty; row0 < count; row0 += step * kRowsU) {
    Vec8<T> v[kRowsU], vr[kRowsU];
#pragma unroll
    for (int u = 0; u < kRowsU; ++u) {
      const int64_t row = row0 + u * step;
      if (row < count) {
        // rev: start at the END of z (the convolution's last tiles, still in L2) and finish at the START of y
        // (the rows the next convolution reads first)
        const int64_t r = rev ? count - 1 - row : row;
        v[u].load(z + r * C + c0);
        if (res != nullptr) vr[u].load(res + r * C + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < kRowsU; ++u) {
      int64_t row = row0 + u * step;
      if (row >= count) break;
      if (rev) row = count - 1 - row;
      float f[8];
      v[u].unpack(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = f[j] * sc[j] + sh[j];
      if (res != nullptr) {
        float r[8];
        vr[u].unpack(r);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += r[j];
      }
      if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
      }
      Vec8<T> o;
      o.pack(f);
      o.store(y + row * C + c0);
      if (relu_mask != nullptr) {
        // one bit per element: y > 0 (wlseg_bn_apply_mask); byte tx of the row = channels [8 tx, 8 tx + 8)
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) m |= (f[j] > 0.f ? 1u : 0u) << j;
        relu_mask[row * cv + tx] = (uint8_t)m;
      }
    }
  }
}

template <typename T, int kBwdRowsU>
__global__ void __launch_bounds__(256)
bn_bwd_apply_rows_kernel(const T* __restrict__ dy, const T* __restrict__ yact, const T* __restrict__ z,
                         const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ gamma, const float* __restrict__ scale,
                         const float* __restrict__ shift, const double* __restrict__ dgamma,
                         const double* __restrict__ dbeta, int64_t count, int64_t stat_count, int C, int pitch, int relu,
                         T* __restrict__ dz, T* __restrict__ dres) {
  pdl_launch_dependents();
  pdl_wait();   // dgamma / dbeta were accumulated by bn_reduce, the kernel right before this one
  const int cv = C / 8;
  const int lanes = 256 / cv;
  const int tx = threadIdx.x % cv, ty = threadIdx.x / cv;
  if (ty >= lanes) return;
  const int c0 = tx * 8;
  const bool zmask = relu && yact == nullptr;
  const double invn = 1.0 / (double)stat_count;
  // dz = A*g + c1*z + c0 (see bn_bwd_apply_kernel); this thread's 8 channels only
  float A[8], B[8], D[8], sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    const float is = invstd[c];
    A[j] = gamma[c] * is;
    B[j] = -A[j] * is * (float)(dgamma[c] * invn);
    D[j] = -A[j] * (float)(dbeta[c] * invn) - B[j] * mean[c];
    sc[j] = zmask ? scale[c] : 0.f;
    sh[j] = zmask ? shift[c] : 0.f;
  }
  const int64_t step = (int64_t)gridDim.x * lanes;
  for (int64_t row0 = (int64_t)blockIdx.x * lanes + ty; row0 < count; row0 += step * kBwdRowsU) {
    Vec8<T> v[kBwdRowsU], vz[kBwdRowsU], vy[kBwdRowsU];
#pragma unroll
    for (int u = 0; u < kBwdRowsU; ++u) {
      const int64_t row = row0 + u * step;
      if (row < count) {
        const int64_t e = row * pitch + c0;
        v[u].load(dy + e);
        vz[u].load(z + e);
        if (relu && !zmask) vy[u].load(yact + e);
      }
    }
#pragma unroll
    for (int u = 0; u < kBwdRowsU; ++u) {
      const int64_t row = row0 + u * step;
      if (row >= count) break;
      const int64_t e = row * pitch + c0;
      float g[8], zz[8];
      v[u].unpack(g);
      vz[u].unpack(zz);
      if (zmask) {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = fmaf(zz[j], sc[j], sh[j]) > 0.f ? g[j] : 0.f;
      } else if (relu) {
        float yy[8];
        vy[u].unpack(yy);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = yy[j] > 0.f ? g[j] : 0.f;
      }
      if (dres != nullptr) {
        Vec8<T> o;
        o.pack(g);
        o.store(dres + e);
      }
      float out[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) out[j] = fmaf(A[j], g[j], fmaf(B[j], zz[j], D[j]));
      Vec8<T> o;
      o.pack(out);
      o.store(dz + e);
    }
  }
}

// WLSEG_BN_FLAT=1 selects the flat-index kernels (A/B measurements; both forms are bit-identical)
static bool bn_rows_enabled() { return getenv("WLSEG_BN_FLAT") == nullptr; }
// tuning knobs of the row-mapped kernels (tools/bn_sweep.py): rows in flight per thread and CTAs per SM
static int bn_apply_u() { return env_int("WLSEG_BN_APPLY_U", 2); }
static int bn_apply_ctas() { return env_int("WLSEG_BN_APPLY_CTAS", 4); }
static int bn_bwd_u() { return env_int("WLSEG_BN_BWD_U", 2); }
static int bn_bwd_ctas() { return env_int("WLSEG_BN_BWD_CTAS", 2); }

template <typename T>
static void launch_apply_rows(const void* z, const float* scale, const float* shift, const void* res, void* y, int64_t count,
                              int C, int relu, cudaStream_t s, uint8_t* mask = nullptr, const BnFin* finp = nullptr) {
  BnFin fin = {};
  if (finp != nullptr) fin = *finp;
  const int lanes = 256 / (C / 8);
  // measured (tools/bn_sweep.py, graph-timed, HBM-cold): without a residual one row per thread at full occupancy
  // wins (871 vs 902 us per step-equivalent); with a residual stream two rows at 4 CTAs / SM do
  const int U = res != nullptr ? bn_apply_u() : env_int("WLSEG_BN_APPLY_U_PLAIN", 1);
  const int g = bw_grid(ceil_div(count, (int64_t)lanes * U) * 256, 256,
                        res != nullptr ? bn_apply_ctas() : env_int("WLSEG_BN_APPLY_CTAS_PLAIN", 8));
  const int rev = env_int("WLSEG_BN_APPLY_REV", 1);
  cudaError_t e;
  if (U == 1) e = launch_pdl(bn_apply_rows_kernel<T, 1>, dim3(g), dim3(256), 0, s, (const T*)z, scale, shift, (const T*)res, (T*)y, count, C, relu, rev, mask, fin);
  else if (U == 4) e = launch_pdl(bn_apply_rows_kernel<T, 4>, dim3(g), dim3(256), 0, s, (const T*)z, scale, shift, (const T*)res, (T*)y, count, C, relu, rev, mask, fin);
  else e = launch_pdl(bn_apply_rows_kernel<T, 2>, dim3(g), dim3(256), 0, s, (const T*)z, scale, shift, (const T*)res, (T*)y, count, C, relu, rev, mask, fin);
  (void)e;   // reported by the caller's WLSEG_LAUNCH_CHECK (cudaGetLastError)
}

template <typename T>
static void launch_bwd_apply_rows(const void* dy, const void* y, const void* z, const float* mean, const float* invstd,
                                  const float* gamma, const float* scale, const float* shift, const double* dgamma,
                                  const double* dbeta, int64_t count, int64_t stat_count, int C, int pitch, int relu,
                                  void* dz, void* dres, cudaStream_t s) {
  const int lanes = 256 / (C / 8);
  const int U = bn_bwd_u();
  const int g = bw_grid(ceil_div(count, (int64_t)lanes * U) * 256, 256, bn_bwd_ctas());
#define WLSEG_BWD_ROWS(UU)                                                                                          \
  (void)launch_pdl(bn_bwd_apply_rows_kernel<T, UU>, dim3(g), dim3(256), 0, s, (const T*)dy, (const T*)y, (const T*)z,  \
                   mean, invstd, gamma, scale, shift, dgamma, dbeta, count, stat_count, C, pitch, relu, (T*)dz,        \
                   (T*)dres)
  if (U == 1) WLSEG_BWD_ROWS(1);
  else if (U == 4) WLSEG_BWD_ROWS(4);
  else WLSEG_BWD_ROWS(2);
#undef WLSEG_BWD_ROWS
}

// ---- generic path for channel counts that are not a multiple of 8 (the logits layers: 14/7/3 and
// 53/12/5 channels).  Tiny tensors; one thread per (row lane, channel), fp64 partial sums.
constexpr int kSmallMaxC = 64;

template <typename T, bool kBackward>
__global__ void __launch_bounds__(kBnThreads)
bn_reduce_small_kernel(const T* __restrict__ a, const T* __restrict__ yact, const T* __restrict__ z,
                       const float* __restrict__ mean, const float* __restrict__ invstd, int64_t count, int C,
                       int pitch, int relu, double* __restrict__ out0, double* __restrict__ out1) {
  __shared__ double part[2][kBnThreads];
  const int lanes = kBnThreads / C;
  const int c = threadIdx.x % C, ty = threadIdx.x / C;
  double s0 = 0.0, s1 = 0.0;
  if (ty < lanes) {
    const float mu = kBackward ? mean[c] : 0.f, is = kBackward ? invstd[c] : 0.f;
    for (int64_t row = (int64_t)blockIdx.x * lanes + ty; row < count; row += (int64_t)gridDim.x * lanes) {
      float f = to_f32<T>(a[row * pitch + c]);
      if (!kBackward) {
        s0 += f;
        s1 += (double)f * f;
      } else {
        if (relu && !(to_f32<T>(yact[row * pitch + c]) > 0.f)) f = 0.f;
        s0 += f * (to_f32<T>(z[row * pitch + c]) - mu) * is;
        s1 += f;
      }
    }
  }
  part[0][threadIdx.x] = s0;
  part[1][threadIdx.x] = s1;
  __syncthreads();
  if (threadIdx.x < C) {
    double t0 = 0.0, t1 = 0.0;
    for (int l = 0; l < lanes; ++l) { t0 += part[0][l * C + threadIdx.x]; t1 += part[1][l * C + threadIdx.x]; }
    atomicAdd(out0 + threadIdx.x, t0);
    atomicAdd(out1 + threadIdx.x, t1);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_apply_small_kernel(const T* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                      const T* __restrict__ res, T* __restrict__ y, int64_t total, int C, int relu) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float f = to_f32<T>(z[i]) * scale[c] + shift[c];
    if (res != nullptr) f += to_f32<T>(res[i]);
    if (relu) f = fmaxf(f, 0.f);
    y[i] = from_f32<T>(f);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_small_kernel(const T* __restrict__ dy, const T* __restrict__ yact, const T* __restrict__ z,
                          const float* __restrict__ mean, const float* __restrict__ invstd,
                          const float* __restrict__ gamma, const double* __restrict__ dgamma,
                          const double* __restrict__ dbeta, int64_t total, int64_t stat_count, int C, int relu,
                          T* __restrict__ dz, T* __restrict__ dres) {
  const float invn = 1.0f / (float)stat_count;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float g = to_f32<T>(dy[i]);
    if (relu && !(to_f32<T>(yact[i]) > 0.f)) g = 0.f;
    if (dres != nullptr) dres[i] = from_f32<T>(g);
    const float is = invstd[c];
    const float zh = (to_f32<T>(z[i]) - mean[c]) * is;
    dz[i] = from_f32<T>(gamma[c] * is * (g - (float)dbeta[c] * invn - zh * (float)dgamma[c] * invn));
  }
}

static int check_bn_shape(int64_t count, int C, const char* who) {
  WLSEG_CHECK_ARG(count >= 0 && C > 0 && ((C % 8 == 0 && C <= 2048) || C <= kSmallMaxC),
                  "%s: C (%d) must be a multiple of 8 and <= 2048, or <= %d", who, C, kSmallMaxC);
  return 0;
}

template <typename T, bool kBackward>
static int launch_reduce(const void* a, const void* y, const void* z, const float* mean, const float* invstd,
                         const float* scale, const float* shift, int64_t count, int C, int pitch, int relu,
                         double* o0, double* o1, cudaStream_t s) {
  if (C % 8 != 0) {
    int grid = bw_grid(count * C, kBnThreads, 4);
    bn_reduce_small_kernel<T, kBackward><<<grid, kBnThreads, 0, s>>>((const T*)a, (const T*)y, (const T*)z, mean,
                                                                      invstd, count, C, pitch, relu, o0, o1);
    WLSEG_LAUNCH_CHECK();
    return 0;
  }
  // 2 CTAs per SM: with kBnUnroll rows in flight per thread that saturates HBM, and it halves the number
  // of CTAs queueing on the same C fp64 accumulators at the end
  const int per_sm = env_int("WLSEG_BN_RED_CTAS", 2);
  int cspan = env_int("WLSEG_BN_RED_SPAN", 128);          // channels per CTA (0 = all of them); tools/bn_sweep.py
  if (cspan <= 0 || cspan >= C || C % cspan != 0 || cspan % 8 != 0) cspan = C;
  const int cv = cspan / 8;
  const int lanes = kBnThreads / cv;
  size_t smem = (size_t)lanes * 2 * cspan * sizeof(float);
  const int ny = C / cspan;
  int gx = bw_grid(ceil_div(count * cv, kBnUnroll), kBnThreads, per_sm) / ny;
  if (gx < 1) gx = 1;
  WLSEG_CUDA(launch_pdl(bn_reduce_kernel<T, kBackward>, dim3(gx, ny), dim3(kBnThreads), smem, s, (const T*)a, (const T*)y,
                        (const T*)z, mean, invstd, scale, shift, count, C, pitch, relu, o0, o1, cspan,
                        kBackward ? env_int("WLSEG_BN_RED_REV", 1) : 0));
  return 0;
}

}  // namespace wlseg

using namespace wlseg;

extern "C" int wlseg_bn_stats(const void* z, int64_t count, int32_t C, int32_t pitch, int32_t dtype, double* sum,
                              double* sqsum, wlseg_stream_t stream) {
  if (int e = check_bn_shape(count, C, "bn_stats")) return e;
  WLSEG_CHECK_ARG(pitch >= C && (pitch % 8 == 0 || C % 8 != 0), "bn_stats: bad pitch %d", pitch);
  if (count == 0) return 0;
  WLSEG_CHECK_ARG(z && sum && sqsum, "bn_stats: null pointer");
  if (dtype == WLSEG_BF16)
    return launch_reduce<__nv_bfloat16, false>(z, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, count, C, pitch,
                                               0, sum, sqsum, (cudaStream_t)stream);
  if (dtype == WLSEG_F32)
    return launch_reduce<float, false>(z, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, count, C, pitch, 0, sum,
                                       sqsum, (cudaStream_t)stream);
  WLSEG_CHECK_ARG(false, "bn_stats: bad dtype %d", dtype);
}

extern "C" int wlseg_bn_finalize(const double* sum, const double* sqsum, int64_t count, int32_t C, const float* gamma,
                                 const float* beta, float eps, float decay, float moving_var_factor,
                                 float* moving_mean, float* moving_var, float* scale, float* shift, float* saved_mean,
                                 float* saved_invstd, wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(sum && sqsum && gamma && beta && count > 0 && C > 0, "bn_finalize: bad args");
  WLSEG_CHECK_ARG((moving_mean == nullptr) == (moving_var == nullptr), "bn_finalize: moving stats must come in pairs");
  WLSEG_CUDA(launch_pdl(bn_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, (cudaStream_t)stream, sum, sqsum, count, C,
                        gamma, beta, eps, decay, moving_var_factor, moving_mean, moving_var, scale, shift, saved_mean,
                        saved_invstd));
  return 0;
}

extern "C" int wlseg_bn_finalize_apply(const double* sum, const double* sqsum, int64_t count, int32_t C,
                                       const float* gamma, const float* beta, float eps, float decay,
                                       float* moving_mean, float* moving_var, float* scale, float* shift,
                                       float* saved_mean, float* saved_invstd, const void* z, const void* residual,
                                       void* y, uint8_t* relu_mask, int32_t relu, int32_t dtype, wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(count > 0 && C > 0 && C % 8 == 0 && C <= 2048, "bn_finalize_apply: C (%d) must be a multiple of 8, <= 2048", C);
  WLSEG_CHECK_ARG(sum && sqsum && gamma && beta && scale && shift && saved_mean && saved_invstd && z && y,
                  "bn_finalize_apply: null pointer");
  WLSEG_CHECK_ARG((moving_mean == nullptr) == (moving_var == nullptr), "bn_finalize_apply: moving stats must come in pairs");
  WLSEG_CHECK_ARG(relu_mask == nullptr || (relu && C % 32 == 0 && bn_rows_enabled()), "bn_finalize_apply: the ReLU mask needs relu, C %% 32 == 0");
  if (bn_rows_enabled()) {
    BnFin fin = {sum, sqsum, gamma, beta, moving_mean, moving_var, scale, shift, saved_mean, saved_invstd, count, eps, decay};
    if (dtype == WLSEG_BF16) launch_apply_rows<__nv_bfloat16>(z, nullptr, nullptr, residual, y, count, C, relu, (cudaStream_t)stream, relu_mask, &fin);
    else if (dtype == WLSEG_F32) launch_apply_rows<float>(z, nullptr, nullptr, residual, y, count, C, relu, (cudaStream_t)stream, relu_mask, &fin);
    else WLSEG_CHECK_ARG(false, "bn_finalize_apply: bad dtype %d", dtype);
    WLSEG_LAUNCH_CHECK();
    return 0;
  }
  const int grid = bw_grid(count * (C / 8), 256, 8);
  const size_t smem = 2 * (size_t)C * sizeof(float);
  if (dtype == WLSEG_BF16)
    bn_finalize_apply_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(
        sum, sqsum, count, C, gamma, beta, eps, decay, moving_mean, moving_var, scale, shift, saved_mean, saved_invstd,
        (const __nv_bfloat16*)z, (const __nv_bfloat16*)residual, (__nv_bfloat16*)y, relu);
  else if (dtype == WLSEG_F32)
    bn_finalize_apply_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(
        sum, sqsum, count, C, gamma, beta, eps, decay, moving_mean, moving_var, scale, shift, saved_mean, saved_invstd,
        (const float*)z, (const float*)residual, (float*)y, relu);
  else
    WLSEG_CHECK_ARG(false, "bn_finalize_apply: bad dtype %d", dtype);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_bn_apply(const void* z, const float* scale, const float* shift, const void* residual, void* y,
                              int64_t count, int32_t C, int32_t relu, int32_t dtype, wlseg_stream_t stream) {
  if (int e = check_bn_shape(count, C, "bn_apply")) return e;
  if (count == 0) return 0;
  WLSEG_CHECK_ARG(z && scale && shift && y, "bn_apply: null pointer");
  if (C % 8 != 0) {
    const int64_t total = count * C;
    int g = bw_grid(total, 256, 8);
    if (dtype == WLSEG_BF16)
      bn_apply_small_kernel<<<g, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)z, scale, shift,
                                                                 (const __nv_bfloat16*)residual, (__nv_bfloat16*)y, total, C, relu);
    else if (dtype == WLSEG_F32)
      bn_apply_small_kernel<<<g, 256, 0, (cudaStream_t)stream>>>((const float*)z, scale, shift, (const float*)residual,
                                                                 (float*)y, total, C, relu);
    else
      WLSEG_CHECK_ARG(false, "bn_apply: bad dtype %d", dtype);
    WLSEG_LAUNCH_CHECK();
    return 0;
  }
  if (C <= 2048 && bn_rows_enabled()) {
    if (dtype == WLSEG_BF16) launch_apply_rows<__nv_bfloat16>(z, scale, shift, residual, y, count, C, relu, (cudaStream_t)stream);
    else if (dtype == WLSEG_F32) launch_apply_rows<float>(z, scale, shift, residual, y, count, C, relu, (cudaStream_t)stream);
    else WLSEG_CHECK_ARG(false, "bn_apply: bad dtype %d", dtype);
    WLSEG_LAUNCH_CHECK();
    return 0;
  }
  int grid = bw_grid(count * (C / 8), 256, 8);
  if (dtype == WLSEG_BF16)
    bn_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)z, scale, shift,
                                                            (const __nv_bfloat16*)residual, (__nv_bfloat16*)y, count, C, relu);
  else if (dtype == WLSEG_F32)
    bn_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)z, scale, shift, (const float*)residual,
                                                            (float*)y, count, C, relu);
  else
    WLSEG_CHECK_ARG(false, "bn_apply: bad dtype %d", dtype);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_bn_apply_mask(const void* z, const float* scale, const float* shift, const void* residual, void* y,
                                   uint8_t* relu_mask, int64_t count, int32_t C, int32_t dtype, wlseg_stream_t stream) {
  if (int e = check_bn_shape(count, C, "bn_apply_mask")) return e;
  if (count == 0) return 0;
  WLSEG_CHECK_ARG(z && scale && shift && y && relu_mask, "bn_apply_mask: null pointer");
  WLSEG_CHECK_ARG(C % 32 == 0 && C <= 2048, "bn_apply_mask: C must be a multiple of 32 and <= 2048 (got %d)", C);
  if (dtype == WLSEG_BF16) launch_apply_rows<__nv_bfloat16>(z, scale, shift, residual, y, count, C, 1, (cudaStream_t)stream, relu_mask);
  else if (dtype == WLSEG_F32) launch_apply_rows<float>(z, scale, shift, residual, y, count, C, 1, (cudaStream_t)stream, relu_mask);
  else WLSEG_CHECK_ARG(false, "bn_apply_mask: bad dtype %d", dtype);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_bn_bwd_reduce(const void* dy, const void* y, const void* z, const float* mean,
                                   const float* invstd, const float* scale, const float* shift, int64_t count,
                                   int32_t C, int32_t pitch, int32_t relu, int32_t dtype, double* dgamma,
                                   double* dbeta, wlseg_stream_t stream) {
  if (int e = check_bn_shape(count, C, "bn_bwd_reduce")) return e;
  if (count == 0) return 0;
  WLSEG_CHECK_ARG(dy && z && mean && invstd && dgamma && dbeta && (!relu || y || (scale && shift)),
                  "bn_bwd_reduce: null pointer");
  WLSEG_CHECK_ARG(pitch >= C && (C % 8 != 0 ? pitch == C : pitch % 8 == 0), "bn_bwd_reduce: bad pitch %d", pitch);
  WLSEG_CHECK_ARG(C % 8 == 0 || !relu || y, "bn_bwd_reduce: the small-C path needs y for the ReLU mask");
  if (dtype == WLSEG_BF16)
    return launch_reduce<__nv_bfloat16, true>(dy, y, z, mean, invstd, scale, shift, count, C, pitch, relu, dgamma, dbeta,
                                              (cudaStream_t)stream);
  if (dtype == WLSEG_F32)
    return launch_reduce<float, true>(dy, y, z, mean, invstd, scale, shift, count, C, pitch, relu, dgamma, dbeta,
                                      (cudaStream_t)stream);
  WLSEG_CHECK_ARG(false, "bn_bwd_reduce: bad dtype %d", dtype);
}

extern "C" int wlseg_bn_bwd_apply(const void* dy, const void* y, const void* z, const float* mean, const float* invstd,
                                  const float* gamma, const float* scale, const float* shift, const double* dgamma,
                                  const double* dbeta, int64_t count, int64_t stat_count, int32_t C, int32_t pitch,
                                  int32_t relu, int32_t dtype, void* dz, void* dres, wlseg_stream_t stream) {
  if (int e = check_bn_shape(count, C, "bn_bwd_apply")) return e;
  if (count == 0) return 0;
  WLSEG_CHECK_ARG(stat_count >= count, "bn_bwd_apply: stat_count (%lld) < count (%lld)", (long long)stat_count,
                  (long long)count);
  WLSEG_CHECK_ARG(dy && z && mean && invstd && gamma && dgamma && dbeta && dz && (!relu || y || (scale && shift)),
                  "bn_bwd_apply: null pointer");
  WLSEG_CHECK_ARG(pitch >= C && (C % 8 != 0 ? pitch == C : pitch % 8 == 0), "bn_bwd_apply: bad pitch %d", pitch);
  if (C % 8 != 0) {
    WLSEG_CHECK_ARG(!relu || y, "bn_bwd_apply: the small-C path needs y for the ReLU mask");
    const int64_t total = count * C;
    int g = bw_grid(total, 256, 8);
    if (dtype == WLSEG_BF16)
      bn_bwd_apply_small_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(
          (const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, (const __nv_bfloat16*)z, mean, invstd, gamma, dgamma, dbeta,
          total, stat_count, C, relu, (__nv_bfloat16*)dz, (__nv_bfloat16*)dres);
    else if (dtype == WLSEG_F32)
      bn_bwd_apply_small_kernel<<<g, 256, 0, (cudaStream_t)stream>>>((const float*)dy, (const float*)y, (const float*)z,
                                                                     mean, invstd, gamma, dgamma, dbeta, total, stat_count,
                                                                     C, relu, (float*)dz, (float*)dres);
    else
      WLSEG_CHECK_ARG(false, "bn_bwd_apply: bad dtype %d", dtype);
    WLSEG_LAUNCH_CHECK();
    return 0;
  }
  if (C <= 2048 && bn_rows_enabled()) {
    if (dtype == WLSEG_BF16)
      launch_bwd_apply_rows<__nv_bfloat16>(dy, y, z, mean, invstd, gamma, scale, shift, dgamma, dbeta, count, stat_count, C,
                                           pitch, relu, dz, dres, (cudaStream_t)stream);
    else if (dtype == WLSEG_F32)
      launch_bwd_apply_rows<float>(dy, y, z, mean, invstd, gamma, scale, shift, dgamma, dbeta, count, stat_count, C, pitch,
                                   relu, dz, dres, (cudaStream_t)stream);
    else
      WLSEG_CHECK_ARG(false, "bn_bwd_apply: bad dtype %d", dtype);
    WLSEG_LAUNCH_CHECK();
    return 0;
  }
  int grid = bw_grid(count * (C / 8), 256, 8);
  const size_t smem = 5 * (size_t)C * sizeof(float);
  if (dtype == WLSEG_BF16)
    bn_bwd_apply_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, (const __nv_bfloat16*)z, mean, invstd, gamma, scale, shift,
        dgamma, dbeta, count, stat_count, C, pitch, relu, (__nv_bfloat16*)dz, (__nv_bfloat16*)dres);
  else if (dtype == WLSEG_F32)
    bn_bwd_apply_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const float*)dy, (const float*)y, (const float*)z,
                                                                   mean, invstd, gamma, scale, shift, dgamma, dbeta, count,
                                                                   stat_count, C, pitch, relu, (float*)dz, (float*)dres);
  else
    WLSEG_CHECK_ARG(false, "bn_bwd_apply: bad dtype %d", dtype);
  WLSEG_LAUNCH_CHECK();
  return 0;
}
