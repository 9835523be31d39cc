// Kernels of the optional pyramid (PSP) module, code/models/resnet50_extended_model_hierarchical.py:186-207
// (`--psp_module`, used by the checkpoint the reference's README publishes):
//   slim.layers.avg_pool2d(bottom, k, stride=k)   VALID average pooling into 1 / 2 / 3 / 6 bins
//   tf.image.resize_images(conv, size(bottom), align_corners=True) of the pooled branch back to h x w
// and their gradients (AvgPoolGrad, ResizeBilinearGrad).  The 1x1 convolutions and batch norms between
// them are the ordinary layers of this library.
//
// All four are small bandwidth / latency kernels on an [N, h, w, 256] feature map (19 MB at 4 x 96 x 96):
// they are written for full coalescing (8 channels = 16 bytes per thread, channel-fastest) and for one
// CTA per output cell where a reduction is involved (shared-memory tree, fixed summation order, no
// atomics), and take pixel pitches so that the upsampled branches are written straight into / read
// straight out of their channel slice of the 1280-channel concatenation.
#include "common.cuh"

namespace wlseg {

float resize_scale(int in, int out);

constexpr int kPspThreads = 256;

// ---- average pooling, VALID -------------------------------------------------------------------------
// one CTA per output cell (n, p, q): thread = (channel vector tx, lane ty); lanes stride over the window
template <typename T>
__global__ void __launch_bounds__(kPspThreads)
avgpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int H, int W, int C, int P, int Q, int kh, int kw,
                   int sh, int sw) {
  extern __shared__ float part[];  // [lanes][C]
  const int cv = C / 8, lanes = kPspThreads / cv;
  const int tx = threadIdx.x % cv, ty = threadIdx.x / cv;
  int cell = blockIdx.x;
  const int q = cell % Q; cell /= Q;
  const int p = cell % P;
  const int n = cell / P;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (ty < lanes) {
    const int win = kh * kw;
    for (int i = ty; i < win; i += lanes) {
      const int hh = p * sh + i / kw, ww = q * sw + i % kw;
      Vec8<T> v;
      v.load(x + (((int64_t)n * H + hh) * W + ww) * C + tx * 8);
      float f[8];
      v.unpack(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) part[ty * C + tx * 8 + j] = acc[j];
  }
  __syncthreads();
  if (ty == 0) {
    float out[8];
    const float inv = 1.0f / (float)(kh * kw);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += part[l * C + tx * 8 + j];
      out[j] = s * inv;
    }
    Vec8<T> o;
    o.pack(out);
    o.store(y + (((int64_t)n * P + p) * Q + q) * C + tx * 8);
  }
}

// dx[n, h, w, :] (+)= dy[n, h / sh, w / sw, :] / (kh * kw) inside the pooled region (kernel == stride:
// every input pixel belongs to at most one window), 0 outside it
template <typename T>
__global__ void __launch_bounds__(256)
avgpool_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int C, int P, int Q, int kh,
                   int kw, int accumulate) {
  const int cv = C / 8;
  const int64_t total = (int64_t)N * H * W * cv;
  const float inv = 1.0f / (float)(kh * kw);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    int64_t t = i / cv;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    const int p = h / kh, q = w / kw;
    float g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = 0.f;
    if (p < P && q < Q) {
      Vec8<T> v;
      v.load(dy + (((int64_t)n * P + p) * Q + q) * C + c8 * 8);
      v.unpack(g);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] *= inv;
    }
    if (accumulate) {
      Vec8<T> old;
      old.load(dx + i * 8);
      float f[8];
      old.unpack(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += f[j];
    }
    Vec8<T> o;
    o.pack(g);
    o.store(dx + i * 8);
  }
}

// ---- bilinear resize, align_corners = True ------------------------------------------------------------
// forward: thread = (output pixel, channel vector); the source is tiny (<= 6 x 6 cells) and cache resident
template <typename T>
__global__ void __launch_bounds__(256)
resize_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int h, int w, int C, int H, int W, int y_pitch,
                  float sy, float sx) {
  const int cv = C / 8;
  const int64_t total = (int64_t)N * H * W * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    int64_t t = i / cv;
    const int X = (int)(t % W); t /= W;
    const int Y = (int)(t % H);
    const int n = (int)(t / H);
    const float fy = Y * sy, fx = X * sx;
    const int yl = (int)floorf(fy), xl = (int)floorf(fx);
    const int yh = min(yl + 1, h - 1), xh = min(xl + 1, w - 1);
    const float ly = fy - (float)yl, lx = fx - (float)xl;
    const T* base = x + (int64_t)n * h * w * C + c8 * 8;
    Vec8<T> a, b, c, d;
    a.load(base + ((int64_t)yl * w + xl) * C);
    b.load(base + ((int64_t)yl * w + xh) * C);
    c.load(base + ((int64_t)yh * w + xl) * C);
    d.load(base + ((int64_t)yh * w + xh) * C);
    float tl[8], tr[8], bl[8], br[8], out[8];
    a.unpack(tl); b.unpack(tr); c.unpack(bl); d.unpack(br);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // TF ResizeBilinear: top = tl + (tr - tl) * x_lerp ; out = top + (bottom - top) * y_lerp
      const float top = tl[j] + (tr[j] - tl[j]) * lx;
      const float bot = bl[j] + (br[j] - bl[j]) * lx;
      out[j] = top + (bot - top) * ly;
    }
    Vec8<T> o;
    o.pack(out);
    o.store(y + (((int64_t)n * H + Y) * W + X) * y_pitch + c8 * 8);
  }
}

// backward = exact transpose: one CTA per SOURCE cell (n, i, j) gathers every output pixel that read it
// (rows with floor(Y * sy) in {i - 1, i}, columns likewise) with the forward's own weights
template <typename T>
__global__ void __launch_bounds__(kPspThreads)
resize_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int h, int w, int C, int H, int W, int dy_pitch,
                  float sy, float sx) {
  extern __shared__ float part[];  // [lanes][C]
  const int cv = C / 8, lanes = kPspThreads / cv;
  const int tx = threadIdx.x % cv, ty = threadIdx.x / cv;
  int cell = blockIdx.x;
  const int j0 = cell % w; cell /= w;
  const int i0 = cell % h;
  const int n = cell / h;
  // candidate output range (a superset; the exact membership test is the weight below)
  int Y0 = 0, Y1 = H - 1, X0 = 0, X1 = W - 1;
  if (sy > 0.f) { Y0 = max(0, (int)floorf((i0 - 1) / sy) - 1); Y1 = min(H - 1, (int)ceilf((i0 + 1) / sy) + 1); }
  if (sx > 0.f) { X0 = max(0, (int)floorf((j0 - 1) / sx) - 1); X1 = min(W - 1, (int)ceilf((j0 + 1) / sx) + 1); }
  const int nx = X1 - X0 + 1;
  const int cand = (Y1 - Y0 + 1) * nx;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (ty < lanes) {
    for (int k = ty; k < cand; k += lanes) {
      const int Y = Y0 + k / nx, X = X0 + k % nx;
      const float fy = Y * sy, fx = X * sx;
      const int yl = (int)floorf(fy), xl = (int)floorf(fx);
      const int yh = min(yl + 1, h - 1), xh = min(xl + 1, w - 1);
      const float ly = fy - (float)yl, lx = fx - (float)xl;
      const float wy = (yl == i0 ? 1.0f - ly : 0.f) + (yh == i0 ? ly : 0.f);
      const float wx = (xl == j0 ? 1.0f - lx : 0.f) + (xh == j0 ? lx : 0.f);
      const float wgt = wy * wx;
      if (wgt == 0.f) continue;
      Vec8<T> v;
      v.load(dy + (((int64_t)n * H + Y) * W + X) * dy_pitch + tx * 8);
      float g[8];
      v.unpack(g);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(wgt, g[j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) part[ty * C + tx * 8 + j] = acc[j];
  }
  __syncthreads();
  if (ty == 0) {
    float out[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += part[l * C + tx * 8 + j];
      out[j] = s;
    }
    Vec8<T> o;
    o.pack(out);
    o.store(dx + (((int64_t)n * h + i0) * w + j0) * C + tx * 8);
  }
}

static int check_c(int C, const char* who) {
  WLSEG_CHECK_ARG(C > 0 && C % 8 == 0 && C <= 2048, "%s: C (%d) must be a multiple of 8 and <= 2048", who, C);
  return 0;
}

}  // namespace wlseg

using namespace wlseg;

extern "C" int wlseg_avgpool_valid_fwd(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, int32_t kh,
                                       int32_t kw, int32_t sh, int32_t sw, int32_t dtype, wlseg_stream_t stream) {
  if (int e = check_c(C, "avgpool_fwd")) return e;
  WLSEG_CHECK_ARG(N >= 0 && H > 0 && W > 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0 && kh <= H && kw <= W,
                  "avgpool_fwd: bad geometry");
  if (N == 0) return 0;
  WLSEG_CHECK_ARG(x && y, "avgpool_fwd: null pointer");
  const int P = (H - kh) / sh + 1, Q = (W - kw) / sw + 1;
  const int lanes = kPspThreads / (C / 8);
  const size_t smem = (size_t)lanes * C * sizeof(float);
  const int grid = N * P * Q;
  if (dtype == WLSEG_BF16)
    avgpool_fwd_kernel<<<grid, kPspThreads, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, H, W,
                                                                         C, P, Q, kh, kw, sh, sw);
  else if (dtype == WLSEG_F32)
    avgpool_fwd_kernel<<<grid, kPspThreads, smem, (cudaStream_t)stream>>>((const float*)x, (float*)y, H, W, C, P, Q, kh,
                                                                         kw, sh, sw);
  else
    WLSEG_CHECK_ARG(false, "avgpool_fwd: bad dtype %d", dtype);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_avgpool_valid_bwd(const void* dy, void* dx, int32_t N, int32_t H, int32_t W, int32_t C, int32_t kh,
                                       int32_t kw, int32_t accumulate, int32_t dtype, wlseg_stream_t stream) {
  if (int e = check_c(C, "avgpool_bwd")) return e;
  WLSEG_CHECK_ARG(N >= 0 && H > 0 && W > 0 && kh > 0 && kw > 0 && kh <= H && kw <= W, "avgpool_bwd: bad geometry");
  if (N == 0) return 0;
  WLSEG_CHECK_ARG(dy && dx, "avgpool_bwd: null pointer");
  const int P = (H - kh) / kh + 1, Q = (W - kw) / kw + 1;   // kernel == stride (the PSP bins)
  const int grid = bw_grid((int64_t)N * H * W * (C / 8), 256, 8);
  if (dtype == WLSEG_BF16)
    avgpool_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, N, H, W, C, P, Q,
                                                              kh, kw, accumulate);
  else if (dtype == WLSEG_F32)
    avgpool_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)dy, (float*)dx, N, H, W, C, P, Q, kh, kw,
                                                              accumulate);
  else
    WLSEG_CHECK_ARG(false, "avgpool_bwd: bad dtype %d", dtype);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_resize_bilinear_fwd(const void* x, void* y, int32_t N, int32_t h, int32_t w, int32_t C, int32_t H,
                                         int32_t W, int32_t y_pitch, int32_t dtype, wlseg_stream_t stream) {
  if (int e = check_c(C, "resize_bilinear_fwd")) return e;
  WLSEG_CHECK_ARG(N >= 0 && h > 0 && w > 0 && H > 0 && W > 0 && y_pitch >= C && y_pitch % 8 == 0,
                  "resize_bilinear_fwd: bad geometry");
  if (N == 0) return 0;
  WLSEG_CHECK_ARG(x && y, "resize_bilinear_fwd: null pointer");
  const float sy = resize_scale(h, H), sx = resize_scale(w, W);
  const int grid = bw_grid((int64_t)N * H * W * (C / 8), 256, 8);
  if (dtype == WLSEG_BF16)
    resize_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, h, w, C, H, W,
                                                             y_pitch, sy, sx);
  else if (dtype == WLSEG_F32)
    resize_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, N, h, w, C, H, W, y_pitch, sy, sx);
  else
    WLSEG_CHECK_ARG(false, "resize_bilinear_fwd: bad dtype %d", dtype);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_resize_bilinear_bwd(const void* dy, void* dx, int32_t N, int32_t h, int32_t w, int32_t C, int32_t H,
                                         int32_t W, int32_t dy_pitch, int32_t dtype, wlseg_stream_t stream) {
  if (int e = check_c(C, "resize_bilinear_bwd")) return e;
  WLSEG_CHECK_ARG(N >= 0 && h > 0 && w > 0 && H > 0 && W > 0 && dy_pitch >= C && dy_pitch % 8 == 0,
                  "resize_bilinear_bwd: bad geometry");
  if (N == 0) return 0;
  WLSEG_CHECK_ARG(dy && dx, "resize_bilinear_bwd: null pointer");
  const float sy = resize_scale(h, H), sx = resize_scale(w, W);
  const int lanes = kPspThreads / (C / 8);
  const size_t smem = (size_t)lanes * C * sizeof(float);
  const int grid = N * h * w;
  if (dtype == WLSEG_BF16)
    resize_bwd_kernel<<<grid, kPspThreads, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, h, w, C,
                                                                        H, W, dy_pitch, sy, sx);
  else if (dtype == WLSEG_F32)
    resize_bwd_kernel<<<grid, kPspThreads, smem, (cudaStream_t)stream>>>((const float*)dy, (float*)dx, h, w, C, H, W,
                                                                        dy_pitch, sy, sx);
  else
    WLSEG_CHECK_ARG(false, "resize_bilinear_bwd: bad dtype %d", dtype);
  WLSEG_LAUNCH_CHECK();
  return 0;
}
