// Prediction post-processing of the EVAL / PREDICT branches, code/estimator/define_estimator_hierarchical.py:
//   :530-571  _resize_predictions  probabilities: tf.image.resize_images(bilinear, align_corners=True)
//                                  decisions:     NEAREST_NEIGHBOR, align_corners=True (roundf)
//   :573-630  _replace_voids       void decisions -> runner-up of tf.nn.top_k(probs, 2)
// Both are pure bandwidth work on full-resolution maps.  Layout: probabilities fp32 [N, H, W, C] with
// C = 14 / 7 / 3 (53 / 12 / 5): a pixel is 56 / 28 / 12 bytes, so the resize maps one thread to one output
// ELEMENT (consecutive threads = consecutive addresses: every store instruction of a warp is one
// contiguous 128-byte span) and reads its four source elements through L1/L2 (the source is re-read
// ~(H/h)*(W/w) times from cache, once from HBM).  Algorithmic bytes: 4*C per output pixel (+ the source
// once); nearest: 4 per output pixel.
#include "common.cuh"

namespace wlseg {

float resize_scale(int in, int out);
int check_hierarchy(const wlseg_hierarchy* hier);

__global__ void __launch_bounds__(256)
resize_probs_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int h, int w, int C, int H, int W, float sy,
                    float sx) {
  const int64_t total = (int64_t)N * H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t t = i / C;
    const int X = (int)(t % W); t /= W;
    const int Y = (int)(t % H);
    const int n = (int)(t / H);
    const float fy = Y * sy, fx = X * sx;
    const int yl = (int)floorf(fy), xl = (int)floorf(fx);
    const int yh = min(yl + 1, h - 1), xh = min(xl + 1, w - 1);
    const float ly = fy - (float)yl, lx = fx - (float)xl;
    const float* base = x + (int64_t)n * h * w * C + c;
    const float tl = __ldg(base + ((int64_t)yl * w + xl) * C), tr = __ldg(base + ((int64_t)yl * w + xh) * C);
    const float bl = __ldg(base + ((int64_t)yh * w + xl) * C), br = __ldg(base + ((int64_t)yh * w + xh) * C);
    // TF ResizeBilinear: top = tl + (tr - tl) * x_lerp ; out = top + (bottom - top) * y_lerp
    const float top = tl + (tr - tl) * lx;
    const float bot = bl + (br - bl) * lx;
    y[i] = top + (bot - top) * ly;
  }
}

// TF-1.12 ResizeNearestNeighbor, align_corners: in = min(roundf(out * scale), in_size - 1)
__global__ void __launch_bounds__(256)
resize_nearest_kernel(const int32_t* __restrict__ x, int32_t* __restrict__ y, int N, int h, int w, int H, int W, float sy,
                      float sx) {
  const int64_t total = (int64_t)N * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % W);
    int64_t t = i / W;
    const int Y = (int)(t % H);
    const int n = (int)(t / H);
    const int yi = min((int)roundf(Y * sy), h - 1), xi = min((int)roundf(X * sx), w - 1);
    y[i] = __ldg(x + ((int64_t)n * h + yi) * w + xi);
  }
}

// arg-max of p[0, C - 1): the head's classes without its void channel (the last one); first maximum wins
__device__ __forceinline__ int argmax_nonvoid(const float* __restrict__ p, int C) {
  int arg = 0;
  float best = p[0];
  for (int c = 1; c < C - 1; ++c) {
    const float v = p[c];
    if (v > best) { best = v; arg = c; }
  }
  return arg;
}

// _replace_voids for the hierarchical classifier.  The reference's rule is "where the decision is void take
// indices[..., 1] of top_k(probs, 2)" on a flat classifier whose last channel is void; on this model it stops
// at its key-set assert (:589-592).  The same rule applied per head: wherever the composed decision is the
// void id, every head on the decision path that chose its void channel (always its last one) takes its
// runner-up instead, i.e. its best non-void class, and the decision is composed again.
__global__ void __launch_bounds__(256)
replace_voids_kernel(const __grid_constant__ wlseg_hierarchy hier, const float* __restrict__ p1,
                     const float* __restrict__ pv, const float* __restrict__ ph, int32_t* __restrict__ decisions,
                     int64_t n, int void_cid) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (decisions[i] != void_cid) continue;
    const int d1 = argmax_nonvoid(p1 + i * hier.C1, hier.C1);
    int dec;
    if (d1 == hier.cid_l1_vehicle) dec = hier.veh_to_common[argmax_nonvoid(pv + i * hier.Cv, hier.Cv)];
    else if (d1 == hier.cid_l1_human) dec = hier.hum_to_common[argmax_nonvoid(ph + i * hier.Ch, hier.Ch)];
    else dec = hier.l1_to_common[d1];
    decisions[i] = dec;
  }
}

}  // namespace wlseg

using namespace wlseg;

extern "C" int wlseg_resize_probabilities(const float* x, float* y, int32_t N, int32_t h, int32_t w, int32_t C, int32_t H,
                                          int32_t W, wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(N >= 0 && h > 0 && w > 0 && C > 0 && H > 0 && W > 0, "resize_probabilities: bad geometry");
  if (N == 0) return 0;
  WLSEG_CHECK_ARG(x && y, "resize_probabilities: null pointer");
  const int grid = bw_grid((int64_t)N * H * W * C, 256, 8);
  resize_probs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, N, h, w, C, H, W, resize_scale(h, H), resize_scale(w, W));
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_resize_decisions(const int32_t* x, int32_t* y, int32_t N, int32_t h, int32_t w, int32_t H, int32_t W,
                                      wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(N >= 0 && h > 0 && w > 0 && H > 0 && W > 0, "resize_decisions: bad geometry");
  if (N == 0) return 0;
  WLSEG_CHECK_ARG(x && y, "resize_decisions: null pointer");
  const int grid = bw_grid((int64_t)N * H * W, 256, 8);
  resize_nearest_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, N, h, w, H, W, resize_scale(h, H), resize_scale(w, W));
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_replace_voids(const wlseg_hierarchy* hier, const float* l1_probs, const float* l2v_probs,
                                   const float* l2h_probs, int32_t* decisions, int64_t n_pixels, int32_t void_cid,
                                   wlseg_stream_t stream) {
  if (int e = check_hierarchy(hier)) return e;
  WLSEG_CHECK_ARG(n_pixels >= 0, "replace_voids: negative pixel count");
  WLSEG_CHECK_ARG(hier->C1 >= 2 && hier->Cv >= 2 && hier->Ch >= 2, "replace_voids: every head needs a non-void class");
  if (n_pixels == 0) return 0;
  WLSEG_CHECK_ARG(l1_probs && l2v_probs && l2h_probs && decisions, "replace_voids: null pointer");
  const int grid = bw_grid(n_pixels, 256, 8);
  replace_voids_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*hier, l1_probs, l2v_probs, l2h_probs, decisions, n_pixels,
                                                              void_cid);
  WLSEG_LAUNCH_CHECK();
  return 0;
}
