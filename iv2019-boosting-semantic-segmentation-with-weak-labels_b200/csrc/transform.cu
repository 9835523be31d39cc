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

}  // namespace wlseg

using namespace wlseg;

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
