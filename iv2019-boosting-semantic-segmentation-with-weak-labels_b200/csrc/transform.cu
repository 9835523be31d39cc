// Input packing for the ResNet root convolution (conv1: 7x7, stride 2, 3 -> 64 channels,
// slim `conv2d_same`, reached from code/models/resnet50_extended_feature_extractor.py:25-30).
//
// A 3-channel NHWC image cannot feed TMA / tcgen05 directly (6-byte pixels).  The stride-2 7x7
// convolution is rewritten exactly as a stride-1 convolution over a space-to-depth(2) image:
//   out(p,q) = sum_{r,s<7} x[2p-3+r, 2q-3+s] w[r,s];  r = 2a+i-1, s = 2b+j-1 (a,b in [0,4), i,j in {0,1},
//   the r = -1 / s = -1 taps carry zero weights)  =>  a 4x4 convolution over S2D[y, x, (i,j,c)].
// This kernel additionally unrolls the 4 horizontal taps b into channels ("im2col along W"):
//   X2[n, y, x, b*16 + (i*2+j)*3 + c] = img[n, 2y+i, 2(x-2+b)+j, c]      (12 of 16 slots used)
// so conv1 becomes an R=4, S=1, C=64 convolution that the implicit-GEMM kernel runs with four
// 64-channel K blocks per tile.  Bandwidth kernel: reads the image once, writes 128 B per pixel.
#include "common.cuh"

namespace wlseg {

// One CTA = kPackTX consecutive packed pixels of one packed row (n, y): the two image rows 2y, 2y + 1 that
// feed it are staged ONCE in shared memory with fully coalesced loads (2 * (2 * kPackTX + 6) * 3 values,
// zero outside the image), then every thread composes 16-byte output vectors from the staged rows
// (compile-time tap tables, no global gathers, no divisions in the inner loop).  The flat-index form it
// replaces issued 8 scalar global loads with 64-bit index arithmetic per output vector and ran at
// 1.8 TB/s; output traffic dominates (64 bf16 channels per packed pixel vs 12 image values).
constexpr int kPackTX = 128;
constexpr int kPackCols = 2 * kPackTX + 6;   // image columns 2 * (x0 - 2) .. 2 * (x0 + kPackTX + 1) - 1

template <typename T>
__global__ void __launch_bounds__(256)
conv1_pack_kernel(const T* __restrict__ img, __nv_bfloat16* __restrict__ out, int N, int H, int W, int Hs, int Ws) {
  __shared__ float rows[2][kPackCols * 3];
  const int x0 = blockIdx.x * kPackTX;
  const int y = blockIdx.y;
  const int n = blockIdx.z;
  const int wbase = 2 * (x0 - 2);
  for (int i = threadIdx.x; i < 2 * kPackCols * 3; i += 256) {
    const int ii = i / (kPackCols * 3);
    const int k = i - ii * (kPackCols * 3);
    const int ww = wbase + k / 3;
    const int hh = 2 * y + ii;
    float v = 0.f;
    if (hh < H && ww >= 0 && ww < W) v = to_f32<T>(img[(((int64_t)n * H + hh) * W + wbase) * 3 + k]);
    rows[ii][k] = v;
  }
  __syncthreads();
  // output vector (x, g): g = 8-channel group of the packed pixel; channel = b * 16 + slot,
  // slot = (ii * 2 + jj) * 3 + c for slot < 12 (zero above), b = horizontal tap, image column 2 * (x - 2 + b) + jj
  const int nx = min(kPackTX, Ws - x0);
  __nv_bfloat16* orow = out + (((int64_t)n * Hs + y) * Ws + x0) * 64;
  for (int v = threadIdx.x; v < nx * 8; v += 256) {
    const int g = v & 7, xl = v >> 3;
    const int b = g >> 1;
    const float* r0 = &rows[0][(2 * (xl + b)) * 3];   // column 2 * (x - 2 + b) - wbase = 2 * (xl + b)
    const float* r1 = &rows[1][(2 * (xl + b)) * 3];
    float f[8];
    if ((g & 1) == 0) {
      // slots 0..7: (ii=0: jj=0 c0..2, jj=1 c0..2) = r0[0..5], then (ii=1, jj=0, c0..1) = r1[0..1]
#pragma unroll
      for (int e = 0; e < 6; ++e) f[e] = r0[e];
      f[6] = r1[0];
      f[7] = r1[1];
    } else {
      // slots 8..15: (ii=1, jj=0, c2) = r1[2], (ii=1, jj=1, c0..2) = r1[3..5], slots 12..15 zero
      f[0] = r1[2];
      f[1] = r1[3];
      f[2] = r1[4];
      f[3] = r1[5];
      f[4] = f[5] = f[6] = f[7] = 0.f;
    }
    Vec8<__nv_bfloat16> o;
    o.pack(f);
    o.store(orow + (int64_t)v * 8);
  }
}

// dst[n, h, w, :] = (h % s == 0 && w % s == 0 && h/s < P && w/s < Q) ? src[n, h/s, w/s, :] : 0
// The gradient of a stride-s convolution wrt its input is a stride-1 convolution over this
// zero-inserted dy, which the tensor-core fprop kernel runs (3 of 4 MMAs multiply zeros: the only
// strided layer with an input gradient is block1/unit_3/conv2, 0.6 % of the network's FLOPs).
template <typename T>
__global__ void __launch_bounds__(256)
zero_insert_kernel(const T* __restrict__ src, T* __restrict__ dst, int N, int P, int Q, int C, int s, int Hu, int Wu) {
  const int cv = C / 8;
  const int64_t total = (int64_t)N * Hu * Wu * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    int64_t t = i / cv;
    const int w = (int)(t % Wu); t /= Wu;
    const int h = (int)(t % Hu);
    const int n = (int)(t / Hu);
    Vec8<T> v;
    const int p = h / s, q = w / s;
    if (h - p * s == 0 && w - q * s == 0 && p < P && q < Q) {
      v.load(src + (((int64_t)n * P + p) * Q + q) * C + c8 * 8);
    } else {
      const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      v.pack(z);
    }
    v.store(dst + i * 8);
  }
}

// All dgrad filter banks of the network in ONE launch (table row per layer: {src_off, dst_off, K, R, S, C}),
// dst[c][R-1-r][S-1-s][k] = src[k][r][s][c]: per tap a K x C -> C x K transpose, done in 64 x 64 tiles through shared
// memory (reads coalesced along c, writes along k).  Round 2: ONE flat list of tiles over all layers, walked by a
// persistent 1-D grid (the first version gave every layer 48 CTAs: the three 512 x 3 x 3 x 512 banks were the long pole
// while the CTAs of the small layers had nothing to do), and 4-byte accesses for 2-byte elements: 124 us -> see
// profiles/r2_timeline_train.txt.
constexpr int kFlipMaxLayers = 512;

template <typename T>
__global__ void __launch_bounds__(256)
transpose_flip_batched_kernel(const T* __restrict__ src, T* __restrict__ dst, const int32_t* __restrict__ table,
                              int n_layers) {
  __shared__ T tile[64][66];
  __shared__ int pref[kFlipMaxLayers + 1];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int l = 0; l < n_layers; ++l) {
      const int32_t* e = table + 6 * l;
      pref[l] = acc;
      acc += e[3] * e[4] * ((e[2] + 63) / 64) * ((e[5] + 63) / 64);
    }
    pref[n_layers] = acc;
  }
  __syncthreads();
  const int total = pref[n_layers];
  int layer = 0;
  for (int u = blockIdx.x; u < total; u += gridDim.x) {
    while (u >= pref[layer + 1]) ++layer;   // u only grows
    const int32_t* e = table + 6 * layer;
    const int64_t so = e[0], dofs = e[1];
    const int K = e[2], R = e[3], S = e[4], C = e[5];
    const int RS = R * S;
    const int kt = (K + 63) / 64, ct = (C + 63) / 64;
    const int lu = u - pref[layer];
    const int tap = lu / (kt * ct);
    const int rem = lu - tap * (kt * ct);
    const int k0 = (rem / ct) * 64, c0 = (rem % ct) * 64;
    const int tap2 = RS - 1 - tap;  // (R-1-r)*S + (S-1-s)
    const bool vec2 = sizeof(T) == 2 && ((K | C) & 1) == 0 && ((so | dofs) & 1) == 0;
    if (vec2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = threadIdx.x + i * 256;
        const int kk = idx >> 5, cc = (idx & 31) * 2;
        if (k0 + kk < K && c0 + cc < C)
          *reinterpret_cast<uint32_t*>(&tile[kk][cc]) =
              *reinterpret_cast<const uint32_t*>(src + so + ((int64_t)(k0 + kk) * RS + tap) * C + c0 + cc);
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = threadIdx.x + i * 256;
        const int cc = idx >> 5, kk = (idx & 31) * 2;
        if (k0 + kk < K && c0 + cc < C) {
          const uint32_t lo = *reinterpret_cast<const uint16_t*>(&tile[kk][cc]);
          const uint32_t hi = *reinterpret_cast<const uint16_t*>(&tile[kk + 1][cc]);
          *reinterpret_cast<uint32_t*>(dst + dofs + ((int64_t)(c0 + cc) * RS + tap2) * K + k0 + kk) = lo | (hi << 16);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int idx = threadIdx.x + i * 256;
        const int kk = idx >> 6, cc = idx & 63;
        if (k0 + kk < K && c0 + cc < C) tile[kk][cc] = src[so + ((int64_t)(k0 + kk) * RS + tap) * C + c0 + cc];
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int idx = threadIdx.x + i * 256;
        const int cc = idx >> 6, kk = idx & 63;
        if (k0 + kk < K && c0 + cc < C) dst[dofs + ((int64_t)(c0 + cc) * RS + tap2) * K + k0 + kk] = tile[kk][cc];
      }
    }
    __syncthreads();
  }
}

}  // namespace wlseg

using namespace wlseg;

extern "C" int wlseg_zero_insert(const void* src, void* dst, int32_t N, int32_t P, int32_t Q, int32_t C,
                                 int32_t stride, int32_t Hu, int32_t Wu, int32_t dtype, wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(N >= 0 && P > 0 && Q > 0 && C > 0 && stride > 0 && Hu > 0 && Wu > 0, "zero_insert: bad shape");
  WLSEG_CHECK_ARG(C % 8 == 0, "zero_insert: C (%d) must be a multiple of 8", C);
  if (N == 0) return 0;
  WLSEG_CHECK_ARG(src && dst, "zero_insert: null pointer");
  const int64_t total = (int64_t)N * Hu * Wu * (C / 8);
  const int grid = bw_grid(total, 256, 8);
  if (dtype == WLSEG_BF16)
    zero_insert_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, N, P, Q, C,
                                                               stride, Hu, Wu);
  else if (dtype == WLSEG_F32)
    zero_insert_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src, (float*)dst, N, P, Q, C, stride, Hu, Wu);
  else
    WLSEG_CHECK_ARG(false, "zero_insert: bad dtype %d", dtype);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_weights_transpose_flip_batched(const void* src_arena, void* dst_arena, const int32_t* table,
                                                    int32_t n_layers, int32_t dtype, wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(n_layers >= 0, "transpose_flip_batched: bad layer count");
  if (n_layers == 0) return 0;
  WLSEG_CHECK_ARG(src_arena && dst_arena && table, "transpose_flip_batched: null pointer");
  WLSEG_CHECK_ARG(n_layers <= kFlipMaxLayers, "transpose_flip_batched: too many layers");
  dim3 grid(8 * kNumSMs);
  if (dtype == WLSEG_BF16)
    transpose_flip_batched_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src_arena,
                                                                          (__nv_bfloat16*)dst_arena, table, n_layers);
  else if (dtype == WLSEG_F32)
    transpose_flip_batched_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src_arena, (float*)dst_arena, table, n_layers);
  else
    WLSEG_CHECK_ARG(false, "transpose_flip_batched: bad dtype %d", dtype);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_conv1_pack(const void* img, int32_t dtype, int32_t N, int32_t H, int32_t W, void* out,
                                wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(N >= 0 && H > 0 && W > 0, "conv1_pack: bad shape");
  if (N == 0) return 0;
  WLSEG_CHECK_ARG(img && out, "conv1_pack: null pointer");
  const int Hs = (H + 1) / 2, Ws = (W + 1) / 2;
  WLSEG_CHECK_ARG(Hs <= 65535 && N <= 65535, "conv1_pack: image too large for the launch grid");
  const dim3 grid((unsigned)ceil_div(Ws, kPackTX), (unsigned)Hs, (unsigned)N);
  if (dtype == WLSEG_F32)
    conv1_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)img, (__nv_bfloat16*)out, N, H, W, Hs, Ws);
  else if (dtype == WLSEG_BF16)
    conv1_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)img, (__nv_bfloat16*)out, N, H, W,
                                                              Hs, Ws);
  else
    WLSEG_CHECK_ARG(false, "conv1_pack: bad dtype %d", dtype);
  WLSEG_LAUNCH_CHECK();
  return 0;
}
