// Max pooling with TensorFlow 'SAME' padding, NHWC, forward and backward.
//
// Replaces slim.max_pool2d under the arg scope of
// code/models/resnet50_extended_model_hierarchical.py:351-353: resnet pool1 (3x3, stride 2) and
// the 1x1 stride-2 `resnet_utils.subsample` on the identity shortcut of block1/unit_3.
// SAME: out = ceil(in/stride); pad_total = max((out-1)*stride + k - in, 0); pad_before =
// pad_total/2 (the odd cell goes bottom/right); padded cells never win.
// Backward: every window's gradient goes to its FIRST maximum in row-major scan order (TF
// MaxPoolGrad); implemented as a gather per input element so no atomics are needed.
//
// HBM-bound: 8 channels (16 B for bf16) per thread, channels innermost => coalesced.
#include <cstdlib>

#include "common.cuh"

namespace wlseg {

struct PoolGeom {
  int N, H, W, C, P, Q, k, stride, pad_t, pad_l;
};

template <typename T>
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, uint8_t* __restrict__ argmax, PoolGeom g) {
  const int cv = g.C / 8;
  const int64_t total = (int64_t)g.N * g.P * g.Q * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c8 = (int)(i % cv);
    int64_t t = i / cv;
    int q = (int)(t % g.Q); t /= g.Q;
    int p = (int)(t % g.P);
    int n = (int)(t / g.P);
    float best[8];
    uint32_t arg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; arg[j] = 255u; }
    for (int r = 0; r < g.k; ++r) {
      int hh = p * g.stride - g.pad_t + r;
      if (hh < 0 || hh >= g.H) continue;
      for (int s = 0; s < g.k; ++s) {
        int ww = q * g.stride - g.pad_l + s;
        if (ww < 0 || ww >= g.W) continue;
        Vec8<T> v;
        v.load(x + (((int64_t)n * g.H + hh) * g.W + ww) * g.C + c8 * 8);
        float f[8];
        v.unpack(f);
        const uint32_t pos = (uint32_t)(r * g.k + s);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // strict '>' keeps the FIRST maximum in row-major scan order (TF MaxPoolGrad routing)
          if (f[j] > best[j] || arg[j] == 255u) { best[j] = f[j]; arg[j] = pos; }
        }
      }
    }
    Vec8<T> o;
    o.pack(best);
    const int64_t oi = (((int64_t)n * g.P + p) * g.Q + q) * g.C + c8 * 8;
    o.store(y + oi);
    if (argmax != nullptr) {
      uint2 packed;
      packed.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
      packed.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
      *reinterpret_cast<uint2*>(argmax + oi) = packed;
    }
  }
}

// pool1 (3x3, stride 2) specialisation: a thread owns one (row p, channel vector) and walks kPoolRun consecutive
// output columns, so the window column shared by neighbouring outputs (input column 2q + 2 - pad) is loaded once
// and carried in registers; the three row loads of a column are issued together.  6 + 6 per run instead of 9
// loads per output, compile-time tap loops, no 64-bit divisions per output.  Same comparison order as the
// generic kernel (row-major scan, strict '>'), hence the same arg-max map bit for bit.
constexpr int kPoolRun = 4;

template <typename T>
__global__ void __launch_bounds__(256)
maxpool3x3s2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, uint8_t* __restrict__ argmax, PoolGeom g) {
  const int cv = g.C / 8;
  const int runs = (g.Q + kPoolRun - 1) / kPoolRun;
  const int64_t total = (int64_t)g.N * g.P * runs * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    int64_t t = i / cv;
    const int run = (int)(t % runs); t /= runs;
    const int p = (int)(t % g.P);
    const int n = (int)(t / g.P);
    const int q0 = run * kPoolRun;
    const int h0 = p * 2 - g.pad_t;
    const T* base = x + (int64_t)n * g.H * g.W * g.C + c8 * 8;
    // column ww of the three window rows -> col[r][8]; invalid cells are flagged, never compared
    Vec8<T> col[3][3];     // [window column s][row r], kept packed (4 registers per bf16 vector)
    bool okc[3], okr[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) okr[r] = (h0 + r >= 0) && (h0 + r < g.H);
    auto load_col = [&](int ww, int slot) {
      okc[slot] = ww >= 0 && ww < g.W;
#pragma unroll
      for (int r = 0; r < 3; ++r)
        if (okc[slot] && okr[r]) col[slot][r].load(base + ((int64_t)(h0 + r) * g.W + ww) * g.C);
    };
    const int w0 = q0 * 2 - g.pad_l;
    load_col(w0, 0);
#pragma unroll
    for (int u = 0; u < kPoolRun; ++u) {
      const int q = q0 + u;
      if (q >= g.Q) break;
      const int ww = q * 2 - g.pad_l;
      load_col(ww + 1, 1);
      load_col(ww + 2, 2);
      float best[8];
      uint32_t arg[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; arg[j] = 255u; }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int s_ = 0; s_ < 3; ++s_) {
          if (!(okr[r] && okc[s_])) continue;
          const uint32_t pos = (uint32_t)(r * 3 + s_);
          float f[8];
          col[s_][r].unpack(f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (f[j] > best[j] || arg[j] == 255u) { best[j] = f[j]; arg[j] = pos; }
          }
        }
      }
      Vec8<T> o;
      o.pack(best);
      const int64_t oi = (((int64_t)n * g.P + p) * g.Q + q) * g.C + c8 * 8;
      o.store(y + oi);
      if (argmax != nullptr) {
        uint2 packed;
        packed.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
        packed.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
        *reinterpret_cast<uint2*>(argmax + oi) = packed;
      }
      // the last window column becomes the first one of the next output
      okc[0] = okc[2];
#pragma unroll
      for (int r = 0; r < 3; ++r) col[0][r] = col[2][r];
    }
  }
}

// dx[n,h,w,c] = sum over windows (p,q) containing (h,w) whose first maximum is (h,w) of dy[n,p,q,c]
template <typename T>
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, PoolGeom g) {
  const int cv = g.C / 8;
  const int64_t total = (int64_t)g.N * g.H * g.W * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c8 = (int)(i % cv);
    int64_t t = i / cv;
    int w = (int)(t % g.W); t /= g.W;
    int h = (int)(t % g.H);
    int n = (int)(t / g.H);
    float me[8], acc[8];
    {
      Vec8<T> v;
      v.load(x + (((int64_t)n * g.H + h) * g.W + w) * g.C + c8 * 8);
      v.unpack(me);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // windows p with p*stride - pad_t <= h <= p*stride - pad_t + k - 1
    int p_lo = (h + g.pad_t - g.k + 1 + g.stride - 1);
    p_lo = p_lo <= 0 ? 0 : p_lo / g.stride;
    int p_hi = min((h + g.pad_t) / g.stride, g.P - 1);
    int q_lo = (w + g.pad_l - g.k + 1 + g.stride - 1);
    q_lo = q_lo <= 0 ? 0 : q_lo / g.stride;
    int q_hi = min((w + g.pad_l) / g.stride, g.Q - 1);
    for (int p = p_lo; p <= p_hi; ++p) {
      for (int q = q_lo; q <= q_hi; ++q) {
        // am I the first maximum of window (p,q)?  earlier cells must be strictly smaller,
        // later cells must not be larger
        bool win[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) win[j] = true;
        for (int r = 0; r < g.k; ++r) {
          int hh = p * g.stride - g.pad_t + r;
          if (hh < 0 || hh >= g.H) continue;
          for (int s = 0; s < g.k; ++s) {
            int ww = q * g.stride - g.pad_l + s;
            if (ww < 0 || ww >= g.W) continue;
            if (hh == h && ww == w) continue;
            const bool earlier = (hh < h) || (hh == h && ww < w);
            Vec8<T> v;
            v.load(x + (((int64_t)n * g.H + hh) * g.W + ww) * g.C + c8 * 8);
            float f[8];
            v.unpack(f);
#pragma unroll
            for (int j = 0; j < 8; ++j) win[j] = win[j] && (earlier ? (f[j] < me[j]) : (f[j] <= me[j]));
          }
        }
        Vec8<T> gv;
        gv.load(dy + (((int64_t)n * g.P + p) * g.Q + q) * g.C + c8 * 8);
        float gf[8];
        gv.unpack(gf);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += win[j] ? gf[j] : 0.f;
      }
    }
    Vec8<T> o;
    o.pack(acc);
    o.store(dx + (((int64_t)n * g.H + h) * g.W + w) * g.C + c8 * 8);
  }
}

// backward with the forward's argmax map (uint8 window position per output element): every input
// cell looks at the <= ceil(k/stride)^2 windows containing it and takes dy where it was the winner.
template <typename T>
__global__ void __launch_bounds__(256)
maxpool_bwd_argmax_kernel(const uint8_t* __restrict__ argmax, const T* __restrict__ dy, T* __restrict__ dx,
                          PoolGeom g) {
  const int cv = g.C / 8;
  const int64_t total = (int64_t)g.N * g.H * g.W * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c8 = (int)(i % cv);
    int64_t t = i / cv;
    int w = (int)(t % g.W); t /= g.W;
    int h = (int)(t % g.H);
    int n = (int)(t / g.H);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    int p_lo = (h + g.pad_t - g.k + 1 + g.stride - 1);
    p_lo = p_lo <= 0 ? 0 : p_lo / g.stride;
    int p_hi = min((h + g.pad_t) / g.stride, g.P - 1);
    int q_lo = (w + g.pad_l - g.k + 1 + g.stride - 1);
    q_lo = q_lo <= 0 ? 0 : q_lo / g.stride;
    int q_hi = min((w + g.pad_l) / g.stride, g.Q - 1);
    for (int p = p_lo; p <= p_hi; ++p) {
      const int r = h - (p * g.stride - g.pad_t);
      for (int q = q_lo; q <= q_hi; ++q) {
        const int s = w - (q * g.stride - g.pad_l);
        const uint32_t pos = (uint32_t)(r * g.k + s);
        const int64_t oi = (((int64_t)n * g.P + p) * g.Q + q) * g.C + c8 * 8;
        const uint2 am = *reinterpret_cast<const uint2*>(argmax + oi);
        Vec8<T> gv;
        gv.load(dy + oi);
        float gf[8];
        gv.unpack(gf);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[j] += (((am.x >> (8 * j)) & 255u) == pos) ? gf[j] : 0.f;
          acc[4 + j] += (((am.y >> (8 * j)) & 255u) == pos) ? gf[4 + j] : 0.f;
        }
      }
    }
    Vec8<T> o;
    o.pack(acc);
    o.store(dx + i * 8);
  }
}

// Round 2: row-mapped forms of the two backward kernels the training step launches.  The flat-index kernels above
// decode (n, h, w, c8) with three 64-bit divisions per 16-byte vector and loop over runtime window bounds: 105 us
// (3x3 / 2 arg-max backward, 4 x 384 x 384 x 64) and 68 us (the two stride-2 identity shortcuts) against ~16 / ~22 us
// of bytes.  Here blockIdx.z = image, blockIdx.y = input row, x = (column, channel vector): 32-bit arithmetic only.
template <typename T>
__global__ void __launch_bounds__(256)
maxpool_bwd_argmax_rows_kernel(const uint8_t* __restrict__ argmax, const T* __restrict__ dy, T* __restrict__ dx,
                               PoolGeom g) {
  const int cv = g.C / 8;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.W * cv) return;
  const int w = idx / cv, c8 = idx - w * cv;
  const int h = blockIdx.y, n = blockIdx.z;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  int p_lo = h + g.pad_t - g.k + g.stride;
  p_lo = p_lo <= 0 ? 0 : p_lo / g.stride;
  const int p_hi = min((h + g.pad_t) / g.stride, g.P - 1);
  int q_lo = w + g.pad_l - g.k + g.stride;
  q_lo = q_lo <= 0 ? 0 : q_lo / g.stride;
  const int q_hi = min((w + g.pad_l) / g.stride, g.Q - 1);
  for (int p = p_lo; p <= p_hi; ++p) {
    const int r = h - (p * g.stride - g.pad_t);
    const int64_t row = ((int64_t)n * g.P + p) * g.Q;
    for (int q = q_lo; q <= q_hi; ++q) {
      const int s = w - (q * g.stride - g.pad_l);
      const uint32_t pos = (uint32_t)(r * g.k + s);
      const int64_t oi = (row + q) * g.C + c8 * 8;
      const uint2 am = *reinterpret_cast<const uint2*>(argmax + oi);
      Vec8<T> gv;
      gv.load(dy + oi);
      float gf[8];
      gv.unpack(gf);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j] += (((am.x >> (8 * j)) & 255u) == pos) ? gf[j] : 0.f;
        acc[4 + j] += (((am.y >> (8 * j)) & 255u) == pos) ? gf[4 + j] : 0.f;
      }
    }
  }
  Vec8<T> o;
  o.pack(acc);
  o.store(dx + (((int64_t)n * g.H + h) * g.W + w) * g.C + c8 * 8);
}

// k = 1 (the stride-s subsampling of an identity shortcut): dx is dy scattered onto the sampled cells, zero elsewhere
template <typename T>
__global__ void __launch_bounds__(256)
subsample_bwd_rows_kernel(const T* __restrict__ dy, T* __restrict__ dx, PoolGeom g) {
  const int cv = g.C / 8;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.W * cv) return;
  const int w = idx / cv, c8 = idx - w * cv;
  const int h = blockIdx.y, n = blockIdx.z;
  const int hh = h + g.pad_t, ww = w + g.pad_l;
  const int p = hh / g.stride, q = ww / g.stride;
  Vec8<T> v;
  const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  v.pack(z);
  if (p * g.stride == hh && q * g.stride == ww && p < g.P && q < g.Q)
    v.load(dy + (((int64_t)n * g.P + p) * g.Q + q) * g.C + c8 * 8);
  v.store(dx + (((int64_t)n * g.H + h) * g.W + w) * g.C + c8 * 8);
}

static int make_geom(PoolGeom& g, int N, int H, int W, int C, int k, int stride) {
  WLSEG_CHECK_ARG(N >= 0 && H > 0 && W > 0 && C > 0 && k > 0 && stride > 0, "maxpool: bad shape");
  WLSEG_CHECK_ARG(C % 8 == 0, "maxpool: C (%d) must be a multiple of 8", C);
  g.N = N; g.H = H; g.W = W; g.C = C; g.k = k; g.stride = stride;
  g.P = (H + stride - 1) / stride;
  g.Q = (W + stride - 1) / stride;
  int tot_h = (g.P - 1) * stride + k - H; if (tot_h < 0) tot_h = 0;
  int tot_w = (g.Q - 1) * stride + k - W; if (tot_w < 0) tot_w = 0;
  g.pad_t = tot_h / 2;
  g.pad_l = tot_w / 2;
  return 0;
}

}  // namespace wlseg

using namespace wlseg;

extern "C" int wlseg_maxpool_same_fwd(const void* x, void* y, uint8_t* argmax, int32_t N, int32_t H, int32_t W,
                                      int32_t C, int32_t ksize, int32_t stride, int32_t dtype,
                                      wlseg_stream_t stream) {
  PoolGeom g;
  if (int e = make_geom(g, N, H, W, C, ksize, stride)) return e;
  if (N == 0) return 0;
  WLSEG_CHECK_ARG(x && y, "maxpool_fwd: null pointer");
  WLSEG_CHECK_ARG(argmax == nullptr || ksize * ksize < 255, "maxpool_fwd: window too large for the uint8 argmax map");
  if (ksize == 3 && stride == 2 && getenv("WLSEG_POOL_GENERIC") == nullptr) {
    const int64_t items3 = (int64_t)N * g.P * ((g.Q + kPoolRun - 1) / kPoolRun) * (C / 8);
    const int grid3 = bw_grid(items3, 256, 2);
    if (dtype == WLSEG_BF16)
      maxpool3x3s2_fwd_kernel<<<grid3, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, argmax, g);
    else if (dtype == WLSEG_F32)
      maxpool3x3s2_fwd_kernel<<<grid3, 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, argmax, g);
    else
      WLSEG_CHECK_ARG(false, "maxpool_fwd: bad dtype %d", dtype);
    WLSEG_LAUNCH_CHECK();
    return 0;
  }
  int64_t items = (int64_t)N * g.P * g.Q * (C / 8);
  int grid = bw_grid(items, 256, 8);
  if (dtype == WLSEG_BF16)
    maxpool_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, argmax, g);
  else if (dtype == WLSEG_F32)
    maxpool_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, argmax, g);
  else
    WLSEG_CHECK_ARG(false, "maxpool_fwd: bad dtype %d", dtype);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_maxpool_same_bwd(const void* x, const uint8_t* argmax, const void* dy, void* dx, int32_t N,
                                      int32_t H, int32_t W, int32_t C, int32_t ksize, int32_t stride,
                                      int32_t dtype, wlseg_stream_t stream) {
  PoolGeom g;
  if (int e = make_geom(g, N, H, W, C, ksize, stride)) return e;
  if (N == 0) return 0;
  WLSEG_CHECK_ARG((x || argmax) && dy && dx, "maxpool_bwd: null pointer");
  int64_t items = (int64_t)N * H * W * (C / 8);
  int grid = bw_grid(items, 256, 8);
  const bool rows_ok = H <= 65535 && N <= 65535 && getenv("WLSEG_POOL_FLAT") == nullptr;
  const dim3 rgrid((unsigned)ceil_div((int64_t)W * (C / 8), 256), (unsigned)H, (unsigned)N);
  if (rows_ok && ksize == 1) {
    // dx depends on dy alone (every window is one cell)
    if (dtype == WLSEG_BF16)
      subsample_bwd_rows_kernel<<<rgrid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, g);
    else if (dtype == WLSEG_F32)
      subsample_bwd_rows_kernel<<<rgrid, 256, 0, (cudaStream_t)stream>>>((const float*)dy, (float*)dx, g);
    else
      WLSEG_CHECK_ARG(false, "maxpool_bwd: bad dtype %d", dtype);
    WLSEG_LAUNCH_CHECK();
    return 0;
  }
  if (rows_ok && argmax != nullptr) {
    if (dtype == WLSEG_BF16)
      maxpool_bwd_argmax_rows_kernel<<<rgrid, 256, 0, (cudaStream_t)stream>>>(argmax, (const __nv_bfloat16*)dy,
                                                                              (__nv_bfloat16*)dx, g);
    else if (dtype == WLSEG_F32)
      maxpool_bwd_argmax_rows_kernel<<<rgrid, 256, 0, (cudaStream_t)stream>>>(argmax, (const float*)dy, (float*)dx, g);
    else
      WLSEG_CHECK_ARG(false, "maxpool_bwd: bad dtype %d", dtype);
    WLSEG_LAUNCH_CHECK();
    return 0;
  }
  if (argmax != nullptr) {
    if (dtype == WLSEG_BF16)
      maxpool_bwd_argmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(argmax, (const __nv_bfloat16*)dy,
                                                                        (__nv_bfloat16*)dx, g);
    else if (dtype == WLSEG_F32)
      maxpool_bwd_argmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(argmax, (const float*)dy, (float*)dx, g);
    else
      WLSEG_CHECK_ARG(false, "maxpool_bwd: bad dtype %d", dtype);
    WLSEG_LAUNCH_CHECK();
    return 0;
  }
  if (dtype == WLSEG_BF16)
    maxpool_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy,
                                                               (__nv_bfloat16*)dx, g);
  else if (dtype == WLSEG_F32)
    maxpool_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (const float*)dy, (float*)dx, g);
  else
    WLSEG_CHECK_ARG(false, "maxpool_bwd: bad dtype %d", dtype);
  WLSEG_LAUNCH_CHECK();
  return 0;
}
