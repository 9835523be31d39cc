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

template <typename T>
__global__ void __launch_bounds__(256)
conv1_pack_kernel(const T* __restrict__ img, __nv_bfloat16* __restrict__ out, int N, int H, int W, int Hs, int Ws) {
  const int64_t total = (int64_t)N * Hs * Ws * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i & 7);
    int64_t t = i >> 3;
    const int x = (int)(t % Ws); t /= Ws;
    const int y = (int)(t % Hs);
    const int n = (int)(t / Hs);
    const int b = g >> 1;
    const int slot0 = (g & 1) * 8;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int slot = slot0 + e;
      float v = 0.f;
      if (slot < 12) {
        const int ii = slot / 6, jj = (slot / 3) & 1, c = slot % 3;
        const int hh = 2 * y + ii;
        const int ww = 2 * (x - 2 + b) + jj;
        if (hh < H && ww >= 0 && ww < W) v = to_f32<T>(img[(((int64_t)n * H + hh) * W + ww) * 3 + c]);
      }
      f[e] = v;
    }
    Vec8<__nv_bfloat16> o;
    o.pack(f);
    o.store(out + i * 8);
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

// All dgrad filter banks of the network in ONE launch: blockIdx.y = layer (table row
// {src_off, dst_off, K, R, S, C}), dst[c][R-1-r][S-1-s][k] = src[k][r][s][c]: per tap a K x C -> C x K
// transpose, done in 64 x 64 tiles through shared memory (reads coalesced along c, writes along k).
template <typename T>
__global__ void __launch_bounds__(256)
transpose_flip_batched_kernel(const T* __restrict__ src, T* __restrict__ dst, const int32_t* __restrict__ table) {
  __shared__ T tile[64][66];
  const int32_t* e = table + 6 * blockIdx.y;
  const int64_t so = e[0], dofs = e[1];
  const int K = e[2], R = e[3], S = e[4], C = e[5];
  const int RS = R * S;
  const int kt = (K + 63) / 64, ct = (C + 63) / 64;
  const int units = RS * kt * ct;
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int tap = u / (kt * ct);
    const int rem = u - tap * (kt * ct);
    const int k0 = (rem / ct) * 64, c0 = (rem % ct) * 64;
    const int tap2 = RS - 1 - tap;  // (R-1-r)*S + (S-1-s)
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
  WLSEG_CHECK_ARG(n_layers <= 65535, "transpose_flip_batched: too many layers");
  dim3 grid(48, n_layers);
  if (dtype == WLSEG_BF16)
    transpose_flip_batched_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src_arena,
                                                                          (__nv_bfloat16*)dst_arena, table);
  else if (dtype == WLSEG_F32)
    transpose_flip_batched_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src_arena, (float*)dst_arena, table);
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
  const int64_t total = (int64_t)N * Hs * Ws * 8;
  const int grid = bw_grid(total, 256, 8);
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
