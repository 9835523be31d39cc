// Input-side resize + crop of the training pipelines, code/input_pipelines/utils.py:181-247
// `resize_images_and_labels` over code/utils/utils.py:540-605 `resize_images_or_labels`:
//   images  tf.image.resize_images(BILINEAR),          align_corners=False  (TF-1.12 legacy mapping src = dst * in/out)
//   labels  tf.image.resize_images(NEAREST_NEIGHBOR),  align_corners=False  (src = min(floor(dst * in/out), in - 1))
//   --preserve_aspect_ratio: both are resized so that the target size fits tightly (mode 'max') and the SAME random
//   window of the target size is cut out of them.
// One pass: the crop window is read straight out of the source - the resized full-size tensor never exists.
// Bandwidth kernel, one thread per output element (consecutive threads = consecutive addresses).
#include "common.cuh"

namespace wlseg {

// kind 0: fp32 bilinear, 1: fp32 nearest (dense 15-way weak labels), 2: int32 nearest (class-id labels)
template <int kKind, typename T>
__global__ void __launch_bounds__(256)
resize_crop_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int oy, int ox, int TH, int TW,
                   float sy, float sx) {
  const int64_t total = (int64_t)N * TH * TW * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t t = i / C;
    const int X = (int)(t % TW); t /= TW;
    const int Y = (int)(t % TH);
    const int n = (int)(t / TH);
    const float fy = (float)(Y + oy) * sy, fx = (float)(X + ox) * sx;
    const T* base = x + (int64_t)n * H * W * C + c;
    if (kKind == 0) {
      const int yl = min((int)floorf(fy), H - 1), xl = min((int)floorf(fx), W - 1);
      const int yh = min(yl + 1, H - 1), xh = min(xl + 1, W - 1);
      const float ly = fy - floorf(fy), lx = fx - floorf(fx);
      const float tl = (float)__ldg(base + ((int64_t)yl * W + xl) * C), tr = (float)__ldg(base + ((int64_t)yl * W + xh) * C);
      const float bl = (float)__ldg(base + ((int64_t)yh * W + xl) * C), br = (float)__ldg(base + ((int64_t)yh * W + xh) * C);
      const float top = tl + (tr - tl) * lx;
      const float bot = bl + (br - bl) * lx;
      y[i] = (T)(top + (bot - top) * ly);
    } else {
      const int yi = min((int)floorf(fy), H - 1), xi = min((int)floorf(fx), W - 1);
      y[i] = __ldg(base + ((int64_t)yi * W + xi) * C);
    }
  }
}

}  // namespace wlseg

using namespace wlseg;

extern "C" int wlseg_resize_crop(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, int32_t RH,
                                 int32_t RW, int32_t oy, int32_t ox, int32_t TH, int32_t TW, int32_t kind,
                                 wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(N >= 0 && H > 0 && W > 0 && C > 0 && RH > 0 && RW > 0 && TH > 0 && TW > 0, "resize_crop: bad geometry");
  WLSEG_CHECK_ARG(oy >= 0 && ox >= 0 && oy + TH <= RH && ox + TW <= RW, "resize_crop: the crop window [%d+%d, %d+%d] leaves the resized tensor %d x %d",
                  oy, TH, ox, TW, RH, RW);
  WLSEG_CHECK_ARG(kind >= 0 && kind <= 2, "resize_crop: kind must be 0 (fp32 bilinear), 1 (fp32 nearest) or 2 (int32 nearest)");
  if (N == 0) return 0;
  WLSEG_CHECK_ARG(x && y, "resize_crop: null pointer");
  // [TF-1.12] CalculateResizeScale without align_corners: in / static_cast<float>(out)
  const float sy = (float)H / (float)RH, sx = (float)W / (float)RW;
  const int grid = bw_grid((int64_t)N * TH * TW * C, 256, 8);
  cudaStream_t s = (cudaStream_t)stream;
  if (kind == 0) resize_crop_kernel<0, float><<<grid, 256, 0, s>>>((const float*)x, (float*)y, N, H, W, C, oy, ox, TH, TW, sy, sx);
  else if (kind == 1) resize_crop_kernel<1, float><<<grid, 256, 0, s>>>((const float*)x, (float*)y, N, H, W, C, oy, ox, TH, TW, sy, sx);
  else resize_crop_kernel<2, int32_t><<<grid, 256, 0, s>>>((const int32_t*)x, (int32_t*)y, N, H, W, C, oy, ox, TH, TW, sy, sx);
  WLSEG_LAUNCH_CHECK();
  return 0;
}
