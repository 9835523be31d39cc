// Hierarchical head, forward: bilinear x8 upsample (align_corners) of the three low-resolution
// logit maps fused with softmax, argmax and the hierarchical decision composition.
//
// Replaces, from code/models/resnet50_extended_model_hierarchical.py:
//   :84-86,:143-184  _create_upsampler  -> tf.image.resize_images(bilinear, align_corners=True)
//   :88-93           tf.nn.softmax x3, tf.argmax x3 (lowest index on ties)
//   :95-117          tf.gather / tf.where composition into common class ids
// so that the 24- (70-) channel full-resolution logits are never materialised unless asked for.
//
// Bandwidth kernel.  A CTA owns a 64 x 16 output tile, stages the low-resolution patch that
// supports it in shared memory (read once from L2/HBM), and every thread interpolates its
// pixels from the patch.  Algorithmic HBM bytes per output pixel: 4 (decisions) + 4*channels
// of each requested probability map + 4*(C1+Cv+Ch)/64 (low-res logits).
#include <cstdlib>

#include "common.cuh"

namespace wlseg {

constexpr int kHeadTX = 64;
constexpr int kHeadTY = 16;
constexpr int kHeadThreads = 256;

struct HeadArgs {
  const float* logits;
  int N, h, w, H, W;
  int cp;        // channel pitch of the low-res logits (>= Ct)
  float sy, sx;  // (h-1)/(H-1), (w-1)/(W-1) in fp32, as TF's CalculateResizeScale
  int ph, pw;    // patch capacity (rows, cols)
  int32_t* decisions;
  int32_t* l1_dec;
  int32_t* l2v_dec;
  int32_t* l2h_dec;
  float* l1_probs;
  float* l2v_probs;
  float* l2h_probs;
  float* full_logits;
};

struct Interp {
  const float* p00;
  const float* p01;
  const float* p10;
  const float* p11;
  float tx, ty;
  // TF ResizeBilinear: top = tl + (tr - tl)*x_lerp ; out = top + (bottom - top)*y_lerp
  __device__ __forceinline__ float at(int c) const {
    float tl = p00[c], tr = p01[c], bl = p10[c], br = p11[c];
    float top = tl + (tr - tl) * tx;
    float bot = bl + (br - bl) * tx;
    return top + (bot - top) * ty;
  }
};

__device__ __forceinline__ int head_argmax(const Interp& it, int c0, int C, float& best) {
  int arg = 0;
  best = it.at(c0);
  for (int c = 1; c < C; ++c) {
    float v = it.at(c0 + c);
    if (v > best) { best = v; arg = c; }  // strict: first maximum wins, as tf.argmax
  }
  return arg;
}

__device__ __forceinline__ void head_probs(const Interp& it, int c0, int C, float mx, float* out) {
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(it.at(c0 + c) - mx);
  float inv = 1.0f / s;
  for (int c = 0; c < C; ++c) out[c] = expf(it.at(c0 + c) - mx) * inv;
}

__global__ void __launch_bounds__(kHeadThreads)
head_fwd_kernel(const __grid_constant__ wlseg_hierarchy hier, const HeadArgs a) {
  extern __shared__ float patch[];  // [ph][pw][Ct]
  const int Ct = hier.C1 + hier.Cv + hier.Ch;
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * kHeadTY, x0 = blockIdx.x * kHeadTX;
  const int yl0 = (int)floorf(y0 * a.sy), xl0 = (int)floorf(x0 * a.sx);

  // stage the low-res patch (clamped at the borders; clamped cells are only read with weight 0
  // or as the duplicated hi neighbour, exactly as TF's min(lo+1, in-1))
  const int cells = a.ph * a.pw;
  const float* src = a.logits + (int64_t)n * a.h * a.w * a.cp;
  for (int i = threadIdx.x; i < cells * Ct; i += kHeadThreads) {
    int c = i % Ct;
    int cell = i / Ct;
    int px = cell % a.pw, py = cell / a.pw;
    int yy = min(yl0 + py, a.h - 1), xx = min(xl0 + px, a.w - 1);
    patch[i] = __ldg(src + ((int64_t)yy * a.w + xx) * a.cp + c);
  }
  __syncthreads();

  const int tx = threadIdx.x % kHeadTX;
  const int x = x0 + tx;
  if (x >= a.W) return;
  const float fx = x * a.sx;
  const int xl = (int)floorf(fx);
  const int xh = min(xl + 1, a.w - 1);
  const float lx = fx - (float)xl;

  for (int ry = threadIdx.x / kHeadTX; ry < kHeadTY; ry += kHeadThreads / kHeadTX) {
    const int y = y0 + ry;
    if (y >= a.H) break;
    const float fy = y * a.sy;
    const int yl = (int)floorf(fy);
    const int yh = min(yl + 1, a.h - 1);
    Interp it;
    it.tx = lx;
    it.ty = fy - (float)yl;
    it.p00 = patch + ((yl - yl0) * a.pw + (xl - xl0)) * Ct;
    it.p01 = patch + ((yl - yl0) * a.pw + (xh - xl0)) * Ct;
    it.p10 = patch + ((yh - yl0) * a.pw + (xl - xl0)) * Ct;
    it.p11 = patch + ((yh - yl0) * a.pw + (xh - xl0)) * Ct;

    float m1, mv, mh;
    const int d1 = head_argmax(it, 0, hier.C1, m1);
    const int dv = head_argmax(it, hier.C1, hier.Cv, mv);
    const int dh = head_argmax(it, hier.C1 + hier.Cv, hier.Ch, mh);
    int dec;
    if (d1 == hier.cid_l1_vehicle) dec = hier.veh_to_common[dv];
    else if (d1 == hier.cid_l1_human) dec = hier.hum_to_common[dh];
    else dec = hier.l1_to_common[d1];

    const int64_t pix = ((int64_t)n * a.H + y) * a.W + x;
    if (a.decisions) a.decisions[pix] = dec;
    if (a.l1_dec) a.l1_dec[pix] = d1;
    if (a.l2v_dec) a.l2v_dec[pix] = dv;
    if (a.l2h_dec) a.l2h_dec[pix] = dh;
    if (a.l1_probs) head_probs(it, 0, hier.C1, m1, a.l1_probs + pix * hier.C1);
    if (a.l2v_probs) head_probs(it, hier.C1, hier.Cv, mv, a.l2v_probs + pix * hier.Cv);
    if (a.l2h_probs) head_probs(it, hier.C1 + hier.Cv, hier.Ch, mh, a.l2h_probs + pix * hier.Ch);
    if (a.full_logits) {
      float* o = a.full_logits + pix * Ct;
      for (int c = 0; c < Ct; ++c) o[c] = it.at(c);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Column-walking variant (the product path for the known head widths 14/7/3 and 53/12/5):
// thread = one output column, walking kColsTY rows.  The x-interpolated low-res rows `top` / `bot`
// live in registers and are refreshed only when the source row changes (every ~8 output rows at x8),
// so a pixel costs one lerp + one compare per channel instead of four shared-memory loads and three
// lerps: the eval launch is bound by its 4 B/pixel of decisions + the fp32 pipe, not by shared memory.
// Probability / logit maps leave through a per-warp shared-memory transpose: 32 pixels x C floats
// are one contiguous span of the NHWC output, written with fully coalesced stores.
constexpr int kColsTX = 128;
constexpr int kColsTY = 32;

// arg-max over v[LO, HI) as a balanced tournament (depth log2 n instead of a serial chain of n - 1
// dependent compares).  The left operand always holds the lower indices and loses only to a
// STRICTLY larger value, so the first maximum wins exactly as in the serial scan / tf.argmax.
template <int LO, int HI, int CT>
__device__ __forceinline__ void tree_argmax(const float (&v)[CT], float& m, int& d) {
  if constexpr (HI - LO == 1) {
    m = v[LO];
    d = LO;
  } else {
    constexpr int MID = LO + (HI - LO + 1) / 2;
    float ml, mr;
    int dl, dr;
    tree_argmax<LO, MID, CT>(v, ml, dl);
    tree_argmax<MID, HI, CT>(v, mr, dr);
    const bool right = mr > ml;
    m = right ? mr : ml;
    d = right ? dr : dl;
  }
}

template <int C>
__device__ __forceinline__ void warp_store_rows(float* __restrict__ dst, float* __restrict__ stage, const float (&v)[C],
                                                int lane, int n_valid) {
  // stage[lane][c] (row pitch C + 1: conflict-free), then linear coalesced copy of n_valid * C floats
  __syncwarp();
#pragma unroll
  for (int c = 0; c < C; ++c) stage[lane * (C + 1) + c] = v[c];
  __syncwarp();
  const int total = n_valid * C;
  for (int i = lane; i < total; i += 32) dst[i] = stage[(i / C) * (C + 1) + (i % C)];
}

template <int C1, int CV, int CH>
__global__ void __launch_bounds__(kColsTX)
head_fwd_cols_kernel(const __grid_constant__ wlseg_hierarchy hier, const HeadArgs a) {
  constexpr int CT = C1 + CV + CH;
  constexpr int CMAX = C1 > CT ? C1 : CT;   // staging row width (the full-logits map is the widest)
  extern __shared__ float patch[];  // [ph][pw][CT] then per-warp staging [4][32][CMAX + 1]
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * kColsTY, x0 = blockIdx.x * kColsTX;
  const int yl0 = (int)floorf(y0 * a.sy), xl0 = (int)floorf(x0 * a.sx);
  const int cells = a.ph * a.pw;
  const float* src = a.logits + (int64_t)n * a.h * a.w * a.cp;
  for (int i = threadIdx.x; i < cells * CT; i += kColsTX) {
    const int c = i % CT;
    const int cell = i / CT;
    const int px = cell % a.pw, py = cell / a.pw;
    const int yy = min(yl0 + py, a.h - 1), xx = min(xl0 + px, a.w - 1);
    patch[i] = __ldg(src + ((int64_t)yy * a.w + xx) * a.cp + c);
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* stage = patch + cells * CT + warp * 32 * (CMAX + 1);
  const int x = x0 + threadIdx.x;
  const bool live = x < a.W;
  const int xc = live ? x : a.W - 1;
  const float fx = xc * a.sx;
  const int xl = (int)floorf(fx);
  const int xh = min(xl + 1, a.w - 1);
  const float lx = fx - (float)xl;
  const int o0 = (xl - xl0) * CT, o1 = (xh - xl0) * CT;
  const int warp_x0 = x0 + warp * 32;
  const int n_valid = min(32, a.W - warp_x0);   // <= 0: the whole warp is outside the image
  const bool want_maps = a.l1_probs || a.l2v_probs || a.l2h_probs || a.full_logits;

  float top[CT], bot[CT];
  int trow = -1, brow = -1;
  auto load_row = [&](int r, float (&dst)[CT]) {
    const float* base = patch + (r - yl0) * a.pw * CT;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      const float tl = base[o0 + c], tr = base[o1 + c];
      dst[c] = tl + (tr - tl) * lx;   // TF ResizeBilinear: top = tl + (tr - tl) * x_lerp
    }
  };

  for (int ry = 0; ry < kColsTY; ++ry) {
    const int y = y0 + ry;
    if (y >= a.H) break;
    const float fy = y * a.sy;
    const int yl = (int)floorf(fy);
    const int yh = min(yl + 1, a.h - 1);
    const float ly = fy - (float)yl;
    if (yl != trow) {
      if (yl == brow) {
#pragma unroll
        for (int c = 0; c < CT; ++c) top[c] = bot[c];
      } else {
        load_row(yl, top);
      }
      trow = yl;
    }
    if (yh != brow) {
      load_row(yh, bot);
      brow = yh;
    }
    float v[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) v[c] = top[c] + (bot[c] - top[c]) * ly;
    // three arg-maxima, strict '>' : the first maximum wins, as tf.argmax
    int d1, dv, dh;
    float m1, mv, mh;
    tree_argmax<0, C1, CT>(v, m1, d1);
    tree_argmax<C1, C1 + CV, CT>(v, mv, dv);
    tree_argmax<C1 + CV, CT, CT>(v, mh, dh);
    dv -= C1;
    dh -= C1 + CV;
    int dec;
    if (d1 == hier.cid_l1_vehicle) dec = hier.veh_to_common[dv];
    else if (d1 == hier.cid_l1_human) dec = hier.hum_to_common[dh];
    else dec = hier.l1_to_common[d1];
    const int64_t pix = ((int64_t)n * a.H + y) * a.W + x;
    if (live) {
      if (a.decisions) a.decisions[pix] = dec;
      if (a.l1_dec) a.l1_dec[pix] = d1;
      if (a.l2v_dec) a.l2v_dec[pix] = dv;
      if (a.l2h_dec) a.l2h_dec[pix] = dh;
    }
    if (want_maps && n_valid > 0) {
      const int64_t wpix = ((int64_t)n * a.H + y) * a.W + warp_x0;   // first pixel of this warp's span
      if (a.full_logits) warp_store_rows<CT>(a.full_logits + wpix * CT, stage, v, lane, n_valid);
      if (a.l1_probs) {
        float e[C1];
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < C1; ++c) { e[c] = expf(v[c] - m1); s += e[c]; }
        const float inv = 1.0f / s;
#pragma unroll
        for (int c = 0; c < C1; ++c) e[c] *= inv;
        warp_store_rows<C1>(a.l1_probs + wpix * C1, stage, e, lane, n_valid);
      }
      if (a.l2v_probs) {
        float e[CV];
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < CV; ++c) { e[c] = expf(v[C1 + c] - mv); s += e[c]; }
        const float inv = 1.0f / s;
#pragma unroll
        for (int c = 0; c < CV; ++c) e[c] *= inv;
        warp_store_rows<CV>(a.l2v_probs + wpix * CV, stage, e, lane, n_valid);
      }
      if (a.l2h_probs) {
        float e[CH];
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) { e[c] = expf(v[C1 + CV + c] - mh); s += e[c]; }
        const float inv = 1.0f / s;
#pragma unroll
        for (int c = 0; c < CH; ++c) e[c] *= inv;
        warp_store_rows<CH>(a.l2h_probs + wpix * CH, stage, e, lane, n_valid);
      }
    }
  }
}

template <int C1, int CV, int CH>
static int launch_head_cols(const wlseg_hierarchy* hier, HeadArgs& a, cudaStream_t stream) {
  constexpr int CT = C1 + CV + CH;
  a.ph = (int)fminf((float)a.h, floorf(kColsTY * a.sy) + 3.f);
  a.pw = (int)fminf((float)a.w, floorf(kColsTX * a.sx) + 3.f);
  const size_t smem = ((size_t)a.ph * a.pw * CT + (size_t)(kColsTX / 32) * 32 * (CT + 1)) * sizeof(float);
  if (smem > 200 * 1024) return 1;  // not covered: the caller falls back to the generic kernel
  static bool configured = false;
  if (!configured) {
    WLSEG_CUDA(cudaFuncSetAttribute(head_fwd_cols_kernel<C1, CV, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(a.W, kColsTX), (unsigned)ceil_div(a.H, kColsTY), (unsigned)a.N);
  head_fwd_cols_kernel<C1, CV, CH><<<grid, kColsTX, smem, stream>>>(*hier, a);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Evaluation head: decisions and / or the confusion matrix, nothing else (define_estimator EVAL,
// code/estimator/define_estimator_hierarchical.py:161-202: argmax composition -> _streaming_confusion_matrix).
// Same column walk and the same arithmetic as head_fwd_cols_kernel (bit-identical decisions), trimmed for the
// evaluation step - ncu on the general kernel (profiles/r2_head_loss_ncu.md): 247 instructions per pixel, only
// ~110 of them the interpolation and the three arg-maxima:
//   * the y-interpolation keeps `bot - top` in registers (one FFMA per channel and pixel instead of FADD + FFMA);
//   * the L2 heads are interpolated and arg-maxed only in rows where some lane of the warp decided the L1 vehicle /
//     human super-class (the composition reads them nowhere else);
//   * one output pointer advanced per row, no per-map branches;
//   * the confusion matrix is fused: label and decision meet in registers, a thread merges its vertical run of equal
//     (label, decision) pairs and commits it to the CTA's shared int32 histogram once per run; the histogram is
//     flushed with 64-bit global atomics (integer, order independent: bit-exact).  The 8 B/pixel round trip of the
//     decisions through HBM (4 written here + 4 read by confmat_kernel) disappears when the caller asks for the
//     matrix only.
constexpr int kEvalTX = 128;
constexpr int kEvalTY = 32;

struct HeadEvalArgs {
  const float* logits;
  int N, h, w, H, W, cp;
  float sy, sx;
  int ph, pw;
  int32_t* decisions;          // may be NULL
  const int32_t* labels;       // may be NULL (no confusion matrix)
  int num_classes;
  const int32_t* lut;          // decisions -> evaluation class ids, may be NULL
  int lut_size;
  unsigned long long* cm;
  unsigned long long* invalid;
};

template <int LO, int C, int CT>
__device__ __forceinline__ void eval_load_rows(const float* __restrict__ rt, const float* __restrict__ rb, int o0, int o1,
                                               float lx, float (&top)[CT], float (&dlt)[CT]) {
#pragma unroll
  for (int c = LO; c < LO + C; ++c) {
    const float tl = rt[o0 + c], tr = rt[o1 + c];
    const float bl = rb[o0 + c], br = rb[o1 + c];
    const float t = tl + (tr - tl) * lx;   // TF ResizeBilinear: top = tl + (tr - tl) * x_lerp
    const float b = bl + (br - bl) * lx;
    top[c] = t;
    dlt[c] = b - t;
  }
}

// kCm: accumulate the confusion matrix; kLut: decisions pass through a.lut first; kDec: store the decisions
template <int C1, int CV, int CH, bool kCm, bool kLut, bool kDec>
__global__ void __launch_bounds__(kEvalTX, (C1 + CV + CH <= 32) ? 4 : 1)
head_eval_kernel(const __grid_constant__ wlseg_hierarchy hier, const HeadEvalArgs a) {
  constexpr int CT = C1 + CV + CH;
  extern __shared__ float patch[];   // [ph][pw][CT], then the int32 histogram [num_classes^2] and the label tile [TY][TX]
  const int cells = a.ph * a.pw;
  int32_t* hist = reinterpret_cast<int32_t*>(patch + cells * CT);
  int32_t* slab = hist + a.num_classes * a.num_classes;
  __shared__ unsigned s_bad;
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * kEvalTY, x0 = blockIdx.x * kEvalTX;
  const int yl0 = (int)floorf(y0 * a.sy), xl0 = (int)floorf(x0 * a.sx);
  const float* src = a.logits + (int64_t)n * a.h * a.w * a.cp;
  for (int i = threadIdx.x; i < cells * CT; i += kEvalTX) {
    const int c = i % CT;
    const int cell = i / CT;
    const int px = cell % a.pw, py = cell / a.pw;
    const int yy = min(yl0 + py, a.h - 1), xx = min(xl0 + px, a.w - 1);
    patch[i] = __ldg(src + ((int64_t)yy * a.w + xx) * a.cp + c);
  }
  constexpr bool want_cm = kCm;
  const int bins = a.num_classes * a.num_classes;
  if (want_cm) {
    for (int i = threadIdx.x; i < bins; i += kEvalTX) hist[i] = 0;
    if (threadIdx.x == 0) s_bad = 0u;
    // the CTA's label tile, all rows in flight at once (a thread walking its column would otherwise wait one HBM
    // latency per row: measured +32 us on a 67 us launch)
    const int32_t* lsrc = a.labels + ((int64_t)n * a.H + y0) * a.W + x0;
    const bool in_x = x0 + (int)threadIdx.x < a.W;
#pragma unroll 8
    for (int r = 0; r < kEvalTY; ++r)
      slab[r * kEvalTX + threadIdx.x] = (in_x && y0 + r < a.H) ? __ldg(lsrc + (int64_t)r * a.W + threadIdx.x) : 0;
  }
  __syncthreads();

  const int x = x0 + threadIdx.x;
  const bool live = x < a.W;
  const int xc = live ? x : a.W - 1;
  const float fx = xc * a.sx;
  const int xl = (int)floorf(fx);
  const int xh = min(xl + 1, a.w - 1);
  const float lx = fx - (float)xl;
  const int o0 = (xl - xl0) * CT, o1 = (xh - xl0) * CT;

  float top[CT], dlt[CT];
  int row = -1;               // source row pair (row, min(row + 1, h - 1)) held in top / dlt
  bool l2_loaded = false;     // ... including the L2 channels
  int run_bin = -2, run_len = 0;
  unsigned bad = 0;
  const int y_end = min(y0 + kEvalTY, a.H);
  int32_t* dptr = kDec ? a.decisions + ((int64_t)n * a.H + y0) * a.W + x : nullptr;
  const int32_t* lptr = slab + threadIdx.x;
  const float* rt = patch;
  const float* rb = patch;
  const int C = a.num_classes;
  for (int y = y0; y < y_end; ++y) {
    const float fy = y * a.sy;
    const int yl = (int)floorf(fy);
    const float ly = fy - (float)yl;
    if (yl != row) {          // uniform over the CTA
      rt = patch + (yl - yl0) * a.pw * CT;
      rb = patch + (min(yl + 1, a.h - 1) - yl0) * a.pw * CT;
      eval_load_rows<0, C1, CT>(rt, rb, o0, o1, lx, top, dlt);
      row = yl;
      l2_loaded = false;
    }
    float v[C1];
#pragma unroll
    for (int c = 0; c < C1; ++c) v[c] = top[c] + dlt[c] * ly;   // = top + (bot - top) * y_lerp
    int d1;
    float m1;
    tree_argmax<0, C1, C1>(v, m1, d1);
    const bool is_v = d1 == hier.cid_l1_vehicle, is_h = d1 == hier.cid_l1_human;
    int dec = hier.l1_to_common[d1];
    if (__any_sync(0xffffffffu, is_v || is_h)) {    // warp-uniform
      if (!l2_loaded) {
        eval_load_rows<C1, CV + CH, CT>(rt, rb, o0, o1, lx, top, dlt);
        l2_loaded = true;
      }
      float u[CV + CH];
#pragma unroll
      for (int c = 0; c < CV + CH; ++c) u[c] = top[C1 + c] + dlt[C1 + c] * ly;
      int dv, dh;
      float mv, mh;
      tree_argmax<0, CV, CV + CH>(u, mv, dv);
      tree_argmax<CV, CV + CH, CV + CH>(u, mh, dh);
      dh -= CV;
      if (is_v) dec = hier.veh_to_common[dv];
      else if (is_h) dec = hier.hum_to_common[dh];
    }
    if (kDec) {
      if (live) *dptr = dec;
      dptr += a.W;
    }
    if (want_cm) {
      const int32_t l = *lptr;
      lptr += kEvalTX;
      int d = dec;
      bool ok = true;
      if (kLut) {
        ok = (unsigned)d < (unsigned)a.lut_size;
        d = ok ? __ldg(a.lut + d) : 0;
      }
      ok = ok && (unsigned)l < (unsigned)C && (unsigned)d < (unsigned)C;
      const int bin = !live ? -2 : (ok ? l * C + d : -1);   // -2: idle lane; -1: out-of-range pair, skipped and counted
      if (bin != run_bin) {
        if (run_bin >= 0) atomicAdd(&hist[run_bin], run_len);
        else if (run_bin == -1) bad += (unsigned)run_len;
        run_bin = bin;
        run_len = 0;
      }
      ++run_len;
    }
  }
  if (want_cm) {
    if (run_bin >= 0) atomicAdd(&hist[run_bin], run_len);
    else if (run_bin == -1) bad += (unsigned)run_len;
    if (bad) atomicAdd(&s_bad, bad);
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += kEvalTX) {
      const int32_t c = hist[i];
      if (c) atomicAdd(a.cm + i, (unsigned long long)c);
    }
    if (threadIdx.x == 0 && a.invalid != nullptr && s_bad) atomicAdd(a.invalid, (unsigned long long)s_bad);
  }
}

template <int C1, int CV, int CH, bool kCm, bool kLut, bool kDec>
static int launch_head_eval_v(const wlseg_hierarchy* hier, HeadEvalArgs& a, size_t smem, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    WLSEG_CUDA(cudaFuncSetAttribute(head_eval_kernel<C1, CV, CH, kCm, kLut, kDec>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    200 * 1024));
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(a.W, kEvalTX), (unsigned)ceil_div(a.H, kEvalTY), (unsigned)a.N);
  head_eval_kernel<C1, CV, CH, kCm, kLut, kDec><<<grid, kEvalTX, smem, stream>>>(*hier, a);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

template <int C1, int CV, int CH>
static int launch_head_eval(const wlseg_hierarchy* hier, HeadEvalArgs& a, cudaStream_t stream) {
  constexpr int CT = C1 + CV + CH;
  a.ph = (int)fminf((float)a.h, floorf(kEvalTY * a.sy) + 3.f);
  a.pw = (int)fminf((float)a.w, floorf(kEvalTX * a.sx) + 3.f);
  const bool cm = a.labels != nullptr, lut = a.lut != nullptr, dec = a.decisions != nullptr;
  const size_t smem = (size_t)a.ph * a.pw * CT * sizeof(float) +
                      (cm ? ((size_t)a.num_classes * a.num_classes + kEvalTX * kEvalTY) * sizeof(int32_t) : 0);
  if (smem > 200 * 1024) return 1;
  if (!cm) return launch_head_eval_v<C1, CV, CH, false, false, true>(hier, a, smem, stream);
  if (lut) return dec ? launch_head_eval_v<C1, CV, CH, true, true, true>(hier, a, smem, stream)
                      : launch_head_eval_v<C1, CV, CH, true, true, false>(hier, a, smem, stream);
  return dec ? launch_head_eval_v<C1, CV, CH, true, false, true>(hier, a, smem, stream)
             : launch_head_eval_v<C1, CV, CH, true, false, false>(hier, a, smem, stream);
}

int check_hierarchy(const wlseg_hierarchy* hier) {
  WLSEG_CHECK_ARG(hier != nullptr, "hierarchy is NULL");
  WLSEG_CHECK_ARG(hier->C1 > 0 && hier->C1 <= 64 && hier->Cv > 0 && hier->Cv <= 16 && hier->Ch > 0 && hier->Ch <= 8,
                  "hierarchy head widths (%d, %d, %d) out of range", hier->C1, hier->Cv, hier->Ch);
  WLSEG_CHECK_ARG(hier->cid_l1_vehicle >= 0 && hier->cid_l1_vehicle < hier->C1 && hier->cid_l1_human >= 0 &&
                      hier->cid_l1_human < hier->C1,
                  "hierarchy: l1 super-class ids out of range");
  return 0;
}

// scale exactly as TF: (in - 1) / float(out - 1) when out > 1 (align_corners), else in/out
float resize_scale(int in, int out) {
  return (out > 1) ? (float)(in - 1) / (float)(out - 1) : (float)in / (float)out;
}

}  // namespace wlseg

using namespace wlseg;

extern "C" int wlseg_head_fwd(const wlseg_hierarchy* hier, const float* logits, int32_t logits_pitch, int32_t N,
                              int32_t h, int32_t w, int32_t H, int32_t W, int32_t* decisions, int32_t* l1_decisions,
                              int32_t* l2v_decisions, int32_t* l2h_decisions, float* l1_probs, float* l2v_probs,
                              float* l2h_probs, float* fullres_logits, wlseg_stream_t stream) {
  if (int e = check_hierarchy(hier)) return e;
  WLSEG_CHECK_ARG(N >= 0 && h > 0 && w > 0 && H > 0 && W > 0, "head_fwd: bad shape");
  if (N == 0) return 0;
  WLSEG_CHECK_ARG(logits != nullptr, "head_fwd: logits is NULL");
  WLSEG_CHECK_ARG(N <= 65535, "head_fwd: N too large");
  WLSEG_CHECK_ARG(logits_pitch >= hier->C1 + hier->Cv + hier->Ch, "head_fwd: logits_pitch %d < channels", logits_pitch);
  HeadArgs a;
  a.logits = logits;
  a.N = N; a.h = h; a.w = w; a.H = H; a.W = W;
  a.cp = logits_pitch;
  a.sy = resize_scale(h, H);
  a.sx = resize_scale(w, W);
  a.decisions = decisions; a.l1_dec = l1_decisions; a.l2v_dec = l2v_decisions; a.l2h_dec = l2h_decisions;
  a.l1_probs = l1_probs; a.l2v_probs = l2v_probs; a.l2h_probs = l2h_probs; a.full_logits = fullres_logits;
  // the column-walking kernel is instantiated for the two label hierarchies the reference ships
  // (cityscapes 14/7/3, vistas 53/12/5) when the upsampling factor is >= 2; anything else takes the
  // generic kernel below
  if (a.sy <= 0.5f && a.sx <= 0.5f && getenv("WLSEG_HEAD_GENERIC") == nullptr) {
    int rc = 1;
    if (hier->C1 == 14 && hier->Cv == 7 && hier->Ch == 3) rc = launch_head_cols<14, 7, 3>(hier, a, (cudaStream_t)stream);
    else if (hier->C1 == 53 && hier->Cv == 12 && hier->Ch == 5) rc = launch_head_cols<53, 12, 5>(hier, a, (cudaStream_t)stream);
    if (rc != 1) return rc;
  }
  a.ph = (int)fminf((float)h, floorf(kHeadTY * a.sy) + 3.f);
  a.pw = (int)fminf((float)w, floorf(kHeadTX * a.sx) + 3.f);
  const int Ct = hier->C1 + hier->Cv + hier->Ch;
  size_t smem = (size_t)a.ph * a.pw * Ct * sizeof(float);
  WLSEG_CHECK_ARG(smem <= 200 * 1024, "head_fwd: low-res patch (%zu B) does not fit shared memory; "
                  "downsampling by more than ~8x is not supported", smem);
  if (smem > 48 * 1024)
    WLSEG_CUDA(cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(W, kHeadTX), (unsigned)ceil_div(H, kHeadTY), (unsigned)N);
  head_fwd_kernel<<<grid, kHeadThreads, smem, (cudaStream_t)stream>>>(*hier, a);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_confmat_accumulate(const int32_t* labels, const int32_t* decisions, int64_t n, int32_t num_classes,
                                        const int32_t* lut, int32_t lut_size, int64_t* cm, int64_t* invalid,
                                        wlseg_stream_t stream);

extern "C" int wlseg_head_confmat(const wlseg_hierarchy* hier, const float* logits, int32_t logits_pitch, int32_t N,
                                  int32_t h, int32_t w, int32_t H, int32_t W, const int32_t* labels, int32_t num_classes,
                                  const int32_t* lut, int32_t lut_size, int64_t* cm, int64_t* invalid, int32_t* decisions,
                                  wlseg_stream_t stream) {
  if (int e = check_hierarchy(hier)) return e;
  WLSEG_CHECK_ARG(N >= 0 && h > 0 && w > 0 && H > 0 && W > 0, "head_confmat: bad shape");
  if (N == 0) return 0;
  WLSEG_CHECK_ARG(logits != nullptr, "head_confmat: logits is NULL");
  WLSEG_CHECK_ARG(N <= 65535, "head_confmat: N too large");
  WLSEG_CHECK_ARG(logits_pitch >= hier->C1 + hier->Cv + hier->Ch, "head_confmat: logits_pitch %d < channels", logits_pitch);
  WLSEG_CHECK_ARG(labels != nullptr || decisions != nullptr, "head_confmat: neither labels nor a decisions buffer given");
  if (labels != nullptr) {
    WLSEG_CHECK_ARG(cm != nullptr, "head_confmat: cm is NULL");
    WLSEG_CHECK_ARG(num_classes > 0 && num_classes <= 104, "head_confmat: num_classes %d out of (0, 104]", num_classes);
    WLSEG_CHECK_ARG(lut == nullptr || lut_size > 0, "head_confmat: lut given with lut_size %d", lut_size);
  }
  HeadEvalArgs a;
  a.logits = logits;
  a.N = N; a.h = h; a.w = w; a.H = H; a.W = W; a.cp = logits_pitch;
  a.sy = resize_scale(h, H);
  a.sx = resize_scale(w, W);
  a.decisions = decisions; a.labels = labels; a.num_classes = num_classes; a.lut = lut; a.lut_size = lut_size;
  a.cm = reinterpret_cast<unsigned long long*>(cm);
  a.invalid = reinterpret_cast<unsigned long long*>(invalid);
  int rc = 1;
  if (a.sy <= 0.5f && a.sx <= 0.5f && getenv("WLSEG_HEAD_GENERIC") == nullptr) {
    if (hier->C1 == 14 && hier->Cv == 7 && hier->Ch == 3) rc = launch_head_eval<14, 7, 3>(hier, a, (cudaStream_t)stream);
    else if (hier->C1 == 53 && hier->Cv == 12 && hier->Ch == 5) rc = launch_head_eval<53, 12, 5>(hier, a, (cudaStream_t)stream);
  }
  if (rc != 1) return rc;
  // other hierarchies / upsampling factors below 2: the general head kernel into a decisions buffer, then the
  // histogram kernel (the caller must provide the buffer in that case)
  WLSEG_CHECK_ARG(decisions != nullptr, "head_confmat: this configuration needs a decisions buffer (general kernels)");
  if (int e = wlseg_head_fwd(hier, logits, logits_pitch, N, h, w, H, W, decisions, nullptr, nullptr, nullptr, nullptr,
                             nullptr, nullptr, nullptr, stream))
    return e;
  if (labels == nullptr) return 0;
  return wlseg_confmat_accumulate(labels, decisions, (int64_t)N * H * W, num_classes, lut, lut_size, cm, invalid, stream);
}
