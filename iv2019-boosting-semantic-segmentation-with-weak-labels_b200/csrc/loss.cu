// Hierarchical strong + weak-label masked softmax cross-entropy, forward AND backward, fused with
// the x8 bilinear upsample (align_corners) of the low-resolution logits and with its transpose.
//
// Replaces code/estimator/define_losses_hierarchical.py:97-203 (gather / one_hot / _segment_sum
// targets, sparse + dense softmax CE, the weak-label weights that depend on the current L1
// argmax, compute_weighted_loss) plus the TF gradients of those ops and ResizeBilinearGrad for
// _create_upsampler (code/models/resnet50_extended_model_hierarchical.py:167).
//
// Bandwidth kernel: per full-resolution pixel it reads 4 B (strong label) or 60 B (weak 15-way
// label) and nothing else from HBM; the low-resolution logits patch of a tile is staged once in
// shared memory and the low-resolution gradient patch is accumulated in shared memory and
// flushed once.  A CTA owns a 128 x 16 pixel tile and walks it two rows at a time:
//   phase A  thread = pixel: interpolate the Ct logits, three softmaxes, targets, weights,
//            loss partials; write w*(softmax - target) for the pixel to G[row][x][c]
//   phase B  thread = (low-res column j, channel c): reduce G along x with the bilinear column
//            weights, then add the two row-weighted shares into the owned D[.][j][c] entries
// so shared memory needs no atomics and the in-tile summation order is fixed; only the flush of
// the tile's D patch into global dlogits uses (fp32) atomics, where patches of neighbouring tiles
// overlap by one low-res row / column.
#include <cstdlib>

#include "common.cuh"

namespace wlseg {

int check_hierarchy(const wlseg_hierarchy* hier);
float resize_scale(int in, int out);

constexpr int kLossTX = 128;
constexpr int kLossTY = 16;
constexpr int kLossThreads = 256;
constexpr int kNumWeak = 15;

struct LossArgs {
  const float* logits;  // [B, h, w, Ct]
  int n_strong, n_bbox, n_image, h, w, H, W;
  int cp;  // channel pitch of logits / dlogits in global memory (>= Ct)
  float sy, sx;
  int ph, pw;
  int gs;  // row stride of G in floats (odd)
  const int32_t* strong;
  const float* bbox;
  const float* image;
  double* sums;
  double* counts;
  float* dlogits;
  // weak labels in their COMPACT form (wlseg_loss_fwd_bwd_lists): the kernel builds each pixel's 15-way multinomial
  // itself instead of reading 60 B of it - input_subset_bboxes_v2.py:74-98 (`_generate_rla`) and
  // input_subset_image_labels.py:73-107 evaluated in registers.  Used when bbox / image are NULL.
  const float* box_coords;    // [n_bbox][max_boxes][4] = xmin, xmax, ymin, ymax, normalised
  const int32_t* box_cids;    // [n_bbox][max_boxes], outside [0, 14] = padding / unknown label
  int max_boxes;
  const float* image_vec;     // [n_image][15]
  int first;                  // generic tile kernel: first image of its range (blockIdx.z + first)
};

constexpr int kLossMaxBoxes = 516;   // MAX_N_BBOXES, input_subset_bboxes_v2.py:33
struct LossBox { int x0, x1, y0, y1, cid; };

__device__ __forceinline__ void softmax_stats(const float* g, int C, float& mx, int& arg, float& lse_minus_max) {
  mx = g[0];
  arg = 0;
  for (int c = 1; c < C; ++c) {
    float v = g[c];
    if (v > mx) { mx = v; arg = c; }
  }
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(g[c] - mx);
  lse_minus_max = logf(s);
}

// dense-target head (L2): g[0..C) holds logits on entry, w*(softmax - t) on exit.
// Returns ce*w through `loss`, w through the return value.
template <bool kWeak>
__device__ __forceinline__ float l2_head(float* g, int C, const int32_t* __restrict__ bb_map, int strong_idx,
                                         const float (&wl)[kNumWeak], bool l1_ok, float& loss) {
  float mx, lse;
  int arg;
  softmax_stats(g, C, mx, arg, lse);
  const float inv = expf(-lse);  // 1 / sum exp(g - mx)
  float w, ce = 0.f;
  if (!kWeak) {
    w = (strong_idx != C - 1) ? 1.f : 0.f;
    ce = (lse + mx) - g[strong_idx];
    for (int k = 0; k < C; ++k) {
      float p = expf(g[k] - mx) * inv;
      g[k] = w * (p - (k == strong_idx ? 1.f : 0.f));
    }
  } else {
    // targets: t[k] = sum_{c: bb_map[c]==k} wl[c]   (_segment_sum, ascending c)
    float t_last = 0.f, t_max = 0.f;
    for (int k = 0; k < C; ++k) {
      float t = 0.f;
#pragma unroll
      for (int c = 0; c < kNumWeak; ++c) t += (bb_map[c] == k) ? wl[c] : 0.f;
      if (k == C - 1) t_last = t; else t_max = fmaxf(t_max, t);
      ce += t * ((lse + mx) - g[k]);
    }
    const bool on = ((1.0f - t_last) > 0.01f) && l1_ok && (t_max >= 0.01f);
    w = on ? 1.f : 0.f;
    for (int k = 0; k < C; ++k) {
      float t = 0.f;
#pragma unroll
      for (int c = 0; c < kNumWeak; ++c) t += (bb_map[c] == k) ? wl[c] : 0.f;
      float p = expf(g[k] - mx) * inv;
      g[k] = w * (p - t);
    }
  }
  loss = ce * w;
  return w;
}

__global__ void __launch_bounds__(kLossThreads)
loss_fwd_bwd_kernel(const __grid_constant__ wlseg_hierarchy hier, const LossArgs a) {
  extern __shared__ float smem[];
  const int C1 = hier.C1, Cv = hier.Cv, Ch = hier.Ch;
  const int Ct = C1 + Cv + Ch;
  const int cells = a.ph * a.pw;
  float* patch = smem;                       // [ph][pw][Ct] logits
  float* D = patch + cells * Ct;             // [ph][pw][Ct] gradient accumulators
  float* G = D + cells * Ct;                 // [2][TX][gs]
  int* xlo = reinterpret_cast<int*>(G + 2 * kLossTX * a.gs);  // [TX]
  float* xt = reinterpret_cast<float*>(xlo + kLossTX);         // [TX]
  int* xstart = reinterpret_cast<int*>(xt + kLossTX);          // [pw + 1]
  __shared__ double red[3][kLossThreads / 32];
  __shared__ double redc[3][kLossThreads / 32];

  const int b = blockIdx.z + a.first;
  const int y0 = blockIdx.y * kLossTY, x0 = blockIdx.x * kLossTX;
  const int yl0 = (int)floorf(y0 * a.sy), xl0 = (int)floorf(x0 * a.sx);
  const int tile_w = min(kLossTX, a.W - x0);

  const float* src = a.logits + (int64_t)b * a.h * a.w * a.cp;
  for (int i = threadIdx.x; i < cells * Ct; i += kLossThreads) {
    int c = i % Ct;
    int cell = i / Ct;
    int px = cell % a.pw, py = cell / a.pw;
    int yy = min(yl0 + py, a.h - 1), xx = min(xl0 + px, a.w - 1);
    patch[i] = __ldg(src + ((int64_t)yy * a.w + xx) * a.cp + c);
    D[i] = 0.f;
  }
  for (int i = threadIdx.x; i <= a.pw; i += kLossThreads) xstart[i] = tile_w;
  if (threadIdx.x < kLossTX) {
    int x = x0 + threadIdx.x;
    float fx = x * a.sx;
    int xl = (int)floorf(fx);
    xlo[threadIdx.x] = xl;
    xt[threadIdx.x] = fx - (float)xl;
  }
  __syncthreads();
  if (threadIdx.x < tile_w) {
    int xl = xlo[threadIdx.x];
    if (threadIdx.x == 0 || xlo[threadIdx.x - 1] != xl) xstart[xl - xl0] = threadIdx.x;
  }
  __syncthreads();
  // columns without pixels in this tile (only past the last one): make runs empty
  if (threadIdx.x == 0) {
    for (int j = a.pw - 1; j >= 0; --j)
      if (xstart[j] > xstart[j + 1]) xstart[j] = xstart[j + 1];
  }
  __syncthreads();

  const int kind = (b < a.n_strong) ? 0 : (b < a.n_strong + a.n_bbox ? 1 : 2);
  const int tx = threadIdx.x % kLossTX;
  const int trow = threadIdx.x / kLossTX;
  const int x = x0 + tx;
  double acc_loss[3] = {0.0, 0.0, 0.0};
  double acc_cnt[3] = {0.0, 0.0, 0.0};

  for (int ry = 0; ry < kLossTY; ry += 2) {
    // ---------------- phase A ----------------
    const int y = y0 + ry + trow;
    float* g = G + (trow * kLossTX + tx) * a.gs;
    if (x < a.W && y < a.H) {
      const float fy = y * a.sy;
      const int yl = (int)floorf(fy);
      const int yh = min(yl + 1, a.h - 1);
      const float ly = fy - (float)yl;
      const int xl = xlo[tx];
      const int xh = min(xl + 1, a.w - 1);
      const float lx = xt[tx];
      const float* p00 = patch + ((yl - yl0) * a.pw + (xl - xl0)) * Ct;
      const float* p01 = patch + ((yl - yl0) * a.pw + (xh - xl0)) * Ct;
      const float* p10 = patch + ((yh - yl0) * a.pw + (xl - xl0)) * Ct;
      const float* p11 = patch + ((yh - yl0) * a.pw + (xh - xl0)) * Ct;
      for (int c = 0; c < Ct; ++c) {
        float tl = p00[c], tr = p01[c], bl = p10[c], br = p11[c];
        float top = tl + (tr - tl) * lx;
        float bot = bl + (br - bl) * lx;
        g[c] = top + (bot - top) * ly;
      }
      const int64_t pix = (int64_t)y * a.W + x;
      float wl[kNumWeak];
      float lv, lh;
      if (kind == 0) {
        int label = __ldg(a.strong + (int64_t)b * a.H * a.W + pix);
        if ((unsigned)label >= (unsigned)hier.num_classes) {
          for (int c = 0; c < Ct; ++c) g[c] = 0.f;  // malformed label: contributes nothing
        } else {
          // L1: sparse CE on strong pixels, weight drops the L1 void class
          const int y1 = hier.pp_to_l1[label];
          float mx, lse;
          int arg;
          softmax_stats(g, C1, mx, arg, lse);
          const float w1 = (y1 <= C1 - 2) ? 1.f : 0.f;
          const float ce1 = (lse + mx) - g[y1];
          const float inv = expf(-lse);
          for (int k = 0; k < C1; ++k) {
            float p = expf(g[k] - mx) * inv;
            g[k] = w1 * (p - (k == y1 ? 1.f : 0.f));
          }
          acc_loss[0] += (double)(ce1 * w1);
          acc_cnt[0] += (double)w1;
#pragma unroll
          for (int c = 0; c < kNumWeak; ++c) wl[c] = 0.f;
          float wv = l2_head<false>(g + C1, Cv, hier.bb_to_veh, hier.pp_to_veh[label], wl, true, lv);
          float wh = l2_head<false>(g + C1 + Cv, Ch, hier.bb_to_hum, hier.pp_to_hum[label], wl, true, lh);
          acc_loss[1] += (double)lv; acc_cnt[1] += (double)wv;
          acc_loss[2] += (double)lh; acc_cnt[2] += (double)wh;
        }
      } else {
        const float* lab = (kind == 1)
            ? a.bbox + ((int64_t)(b - a.n_strong) * a.H * a.W + pix) * kNumWeak
            : a.image + ((int64_t)(b - a.n_strong - a.n_bbox) * a.H * a.W + pix) * kNumWeak;
#pragma unroll
        for (int c = 0; c < kNumWeak; ++c) wl[c] = __ldg(lab + c);
        // weak images: no L1 loss, but the current L1 argmax gates the L2 weights
        int d1 = 0;
        float best = g[0];
        for (int k = 1; k < C1; ++k) {
          float v = g[k];
          if (v > best) { best = v; d1 = k; }
        }
        for (int k = 0; k < C1; ++k) g[k] = 0.f;
        float wv = l2_head<true>(g + C1, Cv, hier.bb_to_veh, 0, wl, d1 == hier.cid_l1_vehicle, lv);
        float wh = l2_head<true>(g + C1 + Cv, Ch, hier.bb_to_hum, 0, wl, d1 == hier.cid_l1_human, lh);
        acc_loss[1] += (double)lv; acc_cnt[1] += (double)wv;
        acc_loss[2] += (double)lh; acc_cnt[2] += (double)wh;
      }
    } else if (tx < kLossTX) {
      for (int c = 0; c < Ct; ++c) g[c] = 0.f;
    }
    __syncthreads();
    // ---------------- phase B ----------------
    for (int item = threadIdx.x; item < a.pw * Ct; item += kLossThreads) {
      const int j = item / Ct, c = item % Ct;
      const int xl = xl0 + j;
      if (xl > a.w - 1) continue;
      const int s0 = xstart[j], s1 = xstart[j + 1];
      const int sp = (j > 0) ? xstart[j - 1] : s0;  // run of the previous column: its hi share
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int yy = y0 + ry + r;
        if (yy >= a.H) break;
        const float* gr = G + (r * kLossTX) * a.gs + c;
        float sum = 0.f;
        for (int xx = s0; xx < s1; ++xx) {
          float t = xt[xx];
          // hi neighbour clamped onto the same column at the right border: weight (1-t)+t
          float wgt = (xl == a.w - 1) ? 1.0f : (1.0f - t);
          sum += wgt * gr[xx * a.gs];
        }
        for (int xx = sp; xx < s0; ++xx) sum += xt[xx] * gr[xx * a.gs];
        const float fy = yy * a.sy;
        const int yl = (int)floorf(fy);
        const int yh = min(yl + 1, a.h - 1);
        const float ly = fy - (float)yl;
        D[((yl - yl0) * a.pw + j) * Ct + c] += (1.0f - ly) * sum;
        D[((yh - yl0) * a.pw + j) * Ct + c] += ly * sum;
      }
    }
    __syncthreads();
  }

  // flush the gradient patch (neighbouring tiles share border cells -> atomics)
  float* dst = a.dlogits + (int64_t)b * a.h * a.w * a.cp;
  for (int i = threadIdx.x; i < cells * Ct; i += kLossThreads) {
    float v = D[i];
    if (v != 0.f) {
      int c = i % Ct;
      int cell = i / Ct;
      int px = cell % a.pw, py = cell / a.pw;
      int yy = yl0 + py, xx = xl0 + px;
      if (yy < a.h && xx < a.w) atomicAdd(dst + ((int64_t)yy * a.w + xx) * a.cp + c, v);
    }
  }
  // loss / count partials: warp shuffle -> shared -> one double atomic per CTA and head
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double l = warp_sum(acc_loss[k]);
    double n = warp_sum(acc_cnt[k]);
    if (lane == 0) { red[k][wid] = l; redc[k][wid] = n; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double l = 0.0, n = 0.0;
    for (int i = 0; i < kLossThreads / 32; ++i) { l += red[threadIdx.x][i]; n += redc[threadIdx.x][i]; }
    if (n != 0.0 || l != 0.0) {
      atomicAdd(a.sums + threadIdx.x, l);
      atomicAdd(a.counts + threadIdx.x, n);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Column-walking variant (product path for the 14/7/3 hierarchy): thread = one output column walking
// kColTY rows.  Forward: the x-interpolated low-res rows `top` / `bot` live in registers and are
// refreshed only when the source row changes.  Backward (the transpose of the upsample): the pixel
// gradients are accumulated in registers per source row (accT / accB, weights 1-ly / ly) and leave for
// the CTA's shared gradient patch - x weights 1-lx / lx, shared-memory atomics - only when the row
// changes: ~1/8 of the scatter traffic of a per-pixel scatter, no full-resolution gradient staging.
// exp() is evaluated once per class.  Weak labels (60 B/pixel) are staged per warp with coalesced loads.
constexpr int kColTX = 128;

template <int LO, int HI, int CT>
__device__ __forceinline__ void tree_argmax_l(const float (&v)[CT], float& m, int& d) {
  if constexpr (HI - LO == 1) {
    m = v[LO];
    d = LO;
  } else {
    constexpr int MID = LO + (HI - LO + 1) / 2;
    float ml, mr;
    int dl, dr;
    tree_argmax_l<LO, MID, CT>(v, ml, dl);
    tree_argmax_l<MID, HI, CT>(v, mr, dr);
    const bool right = mr > ml;   // strict: the first maximum wins (tf.argmax)
    m = right ? mr : ml;
    d = right ? dr : dl;
  }
}

// One head over v[LO, LO + C): on exit v[LO + k] = w * (softmax_k - t_k); returns w, ce * w in `loss`.
// kDense = false: sparse target `idx` (relative to LO), weight = idx_ok;  kDense = true: soft targets t[].
template <int LO, int C, int CT, bool kDense>
__device__ __forceinline__ float head_ce(float (&v)[CT], int idx, bool idx_ok, const float (&t)[C], bool dense_on,
                                         float& loss) {
  float mx;
  int arg;
  tree_argmax_l<LO, LO + C, CT>(v, mx, arg);
  float e[C];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < C; ++k) { e[k] = __expf(v[LO + k] - mx); s += e[k]; }   // ex2.approx: 2 ulp, inputs <= 0
  const float lse = __logf(s) + mx;
  const float inv = __fdividef(1.0f, s);
  float ce = 0.f, w;
  if (!kDense) {
    float vy = 0.f;
#pragma unroll
    for (int k = 0; k < C; ++k) vy = (k == idx) ? v[LO + k] : vy;
    ce = lse - vy;
    w = idx_ok ? 1.f : 0.f;
#pragma unroll
    for (int k = 0; k < C; ++k) v[LO + k] = w * (e[k] * inv - (k == idx ? 1.f : 0.f));
  } else {
#pragma unroll
    for (int k = 0; k < C; ++k) ce += t[k] * (lse - v[LO + k]);
    w = dense_on ? 1.f : 0.f;
#pragma unroll
    for (int k = 0; k < C; ++k) v[LO + k] = w * (e[k] * inv - t[k]);
  }
  loss = ce * w;
  return w;
}

// soft targets of an L2 head from the 15-way weak label: t[k] = sum_{c: map[c]==k} wl[c] (_segment_sum)
// `sel` = 0/1 matrix [15][C] in shared memory (sel[c][k] = bb_map[c] == k): one FFMA per term, ascending c
template <int C>
__device__ __forceinline__ bool weak_targets(const float* __restrict__ sel, const float (&wl)[kNumWeak], float (&t)[C]) {
#pragma unroll
  for (int k = 0; k < C; ++k) {
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < kNumWeak; ++c) acc = fmaf(sel[c * C + k], wl[c], acc);
    t[k] = acc;
  }
  float t_max = 0.f;
#pragma unroll
  for (int k = 0; k < C - 1; ++k) t_max = fmaxf(t_max, t[k]);
  return ((1.0f - t[C - 1]) > 0.01f) && (t_max >= 0.01f);
}

// kWeak = false: the strong images [0, n_strong); kWeak = true: the bbox + image-level images after them
template <int C1, int CV, int CH, bool kWeak, int kColTY>
__global__ void __launch_bounds__(kColTX)
loss_cols_kernel(const __grid_constant__ wlseg_hierarchy hier, const LossArgs a) {
  constexpr int CT = C1 + CV + CH;
  extern __shared__ float smem[];
  const int cells = a.ph * a.pw;
  float* patch = smem;                 // [ph][pw][CT] logits
  float* D = patch + cells * CT;       // [ph][pw][CT] gradient accumulators (shared-memory atomics)
  float* wstage = D + cells * CT;      // [4 warps][32 * 15] weak-label staging
  __shared__ float red[3][kColTX / 32];
  __shared__ float redc[3][kColTX / 32];
  __shared__ float selv[kNumWeak * CV];
  __shared__ float selh[kNumWeak * CH];
  __shared__ LossBox sbox[kWeak ? kLossMaxBoxes : 1];
  __shared__ int snbox;
  __shared__ float simg[kNumWeak];
  if (kWeak) {
    for (int i = threadIdx.x; i < kNumWeak * CV; i += kColTX) selv[i] = (hier.bb_to_veh[i / CV] == i % CV) ? 1.f : 0.f;
    for (int i = threadIdx.x; i < kNumWeak * CH; i += kColTX) selh[i] = (hier.bb_to_hum[i / CH] == i % CH) ? 1.f : 0.f;
  }

  const int b = blockIdx.z + (kWeak ? a.n_strong : 0);   // the weak launch covers the images after the strong ones
  const int y0 = blockIdx.y * kColTY, x0 = blockIdx.x * kColTX;
  const int yl0 = (int)floorf(y0 * a.sy), xl0 = (int)floorf(x0 * a.sx);
  const float* src = a.logits + (int64_t)b * a.h * a.w * a.cp;
  for (int i = threadIdx.x; i < cells * CT; i += kColTX) {
    const int c = i % CT;
    const int cell = i / CT;
    const int px = cell % a.pw, py = cell / a.pw;
    const int yy = min(yl0 + py, a.h - 1), xx = min(xl0 + px, a.w - 1);
    patch[i] = __ldg(src + ((int64_t)yy * a.w + xx) * a.cp + c);
    D[i] = 0.f;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* wl_s = wstage + warp * (32 * kNumWeak);
  const int kind = (b < a.n_strong) ? 0 : (b < a.n_strong + a.n_bbox ? 1 : 2);
  // compact weak labels: the boxes of this image that touch this CTA's tile, as integer pixel bounds - the same
  // arithmetic as wlseg_rasterize_bbox_labels (int(coord * size) in double, python slice [min : max + 1] clipped)
  const bool from_boxes = kWeak && kind == 1 && a.bbox == nullptr;
  const bool from_vec = kWeak && kind == 2 && a.image == nullptr;
  if constexpr (kWeak) {
    if (threadIdx.x == 0) snbox = 0;
    __syncthreads();
    if (from_boxes) {
      const int bi = b - a.n_strong;
      for (int k = threadIdx.x; k < a.max_boxes; k += kColTX) {
        const int cid = a.box_cids[(int64_t)bi * a.max_boxes + k];
        if (cid < 0 || cid >= kNumWeak) continue;
        const float* c = a.box_coords + ((int64_t)bi * a.max_boxes + k) * 4;
        LossBox bx;
        bx.x0 = (int)((double)c[0] * (double)a.W);
        bx.x1 = (int)((double)c[1] * (double)a.W);
        bx.y0 = (int)((double)c[2] * (double)a.H);
        bx.y1 = (int)((double)c[3] * (double)a.H);
        bx.cid = cid;
        if (bx.x0 < 0) bx.x0 = 0;
        if (bx.y0 < 0) bx.y0 = 0;
        if (bx.x1 > a.W - 1) bx.x1 = a.W - 1;
        if (bx.y1 > a.H - 1) bx.y1 = a.H - 1;
        if (bx.x0 > bx.x1 || bx.y0 > bx.y1) continue;
        if (bx.x1 < x0 || bx.x0 >= x0 + kColTX || bx.y1 < y0 || bx.y0 >= y0 + kColTY) continue;   // misses this tile
        const int slot = atomicAdd(&snbox, 1);
        if (slot < kLossMaxBoxes) sbox[slot] = bx;
      }
    }
    if (from_vec && threadIdx.x < kNumWeak)
      simg[threadIdx.x] = a.image_vec[(int64_t)(b - a.n_strong - a.n_bbox) * kNumWeak + threadIdx.x];
    __syncthreads();
  }
  const int nbox_tile = kWeak ? min(snbox, kLossMaxBoxes) : 0;
  const int x = x0 + threadIdx.x;
  const bool live = x < a.W;
  const int xc = live ? x : a.W - 1;
  const float fx = xc * a.sx;
  const int xl = (int)floorf(fx);
  const int xh = min(xl + 1, a.w - 1);
  const float lx = fx - (float)xl;
  const int o0 = (xl - xl0) * CT, o1 = (xh - xl0) * CT;
  const int warp_x0 = x0 + warp * 32;
  const int n_valid = min(32, a.W - warp_x0);

  float top[CT], bot[CT], accT[CT], accB[CT];
  int trow = -1, brow = -1;
  auto load_row = [&](int r, float (&dst)[CT]) {
    const float* base = patch + (r - yl0) * a.pw * CT;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      const float tl = base[o0 + c], tr = base[o1 + c];
      dst[c] = tl + (tr - tl) * lx;
    }
  };
  // (Round 2 tried the staged warp transpose of loss_strong_kernel here instead of shared atomics: SLOWER for the weak
  // images, 567 -> 735 us per 4 + 8 + 4 batch - their gradients are sparse (an L2 weight needs the L1 decision AND a
  // box), so the `g != 0` guards below skip most of the atomics, while the staged form always pays in full.)
  auto flush_row = [&](int r, const float (&acc)[CT]) {
    float* base = D + (r - yl0) * a.pw * CT;
    const float w0 = 1.0f - lx;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      const float g = acc[c];
      if (g != 0.f) {
        atomicAdd(base + o0 + c, w0 * g);
        atomicAdd(base + o1 + c, lx * g);
      }
    }
  };
  // per-thread partials over <= kColTY pixels stay in fp32; fp64 from the warp reduction on
  float acc_loss[3] = {0.f, 0.f, 0.f};
  float acc_cnt[3] = {0.f, 0.f, 0.f};

  for (int ry = 0; ry < kColTY; ++ry) {
    const int y = y0 + ry;
    if (y >= a.H) break;
    const float fy = y * a.sy;
    const int yl = (int)floorf(fy);
    const int yh = min(yl + 1, a.h - 1);
    const float ly = fy - (float)yl;
    if (yl != trow) {
      if (trow >= 0) flush_row(trow, accT);
      if (yl == brow) {
#pragma unroll
        for (int c = 0; c < CT; ++c) { top[c] = bot[c]; accT[c] = accB[c]; }
        brow = -1;  // moved up: the bottom row is reloaded (and its accumulator restarted) below
      } else {
        load_row(yl, top);
#pragma unroll
        for (int c = 0; c < CT; ++c) accT[c] = 0.f;
      }
      trow = yl;
    }
    if (yh != brow) {
      if (brow >= 0) flush_row(brow, accB);
      load_row(yh, bot);
#pragma unroll
      for (int c = 0; c < CT; ++c) accB[c] = 0.f;
      brow = yh;
    }
    float v[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) v[c] = top[c] + (bot[c] - top[c]) * ly;
    const int64_t pix = (int64_t)y * a.W + x;
    bool contributes = live;
    if constexpr (!kWeak) {
      int label = live ? __ldg(a.strong + (int64_t)b * a.H * a.W + pix) : -1;
      if ((unsigned)label >= (unsigned)hier.num_classes) {
        contributes = false;  // outside the image or malformed label: contributes nothing
        label = 0;
      }
      const int y1 = hier.pp_to_l1[label];
      const float none1[C1] = {};
      const float nonev[CV] = {};
      const float noneh[CH] = {};
      float l1, lv, lh;
      const float w1 = head_ce<0, C1, CT, false>(v, y1, y1 <= C1 - 2, none1, false, l1);
      const int yv = hier.pp_to_veh[label], yh2 = hier.pp_to_hum[label];
      const float wv = head_ce<C1, CV, CT, false>(v, yv, yv != CV - 1, nonev, false, lv);
      const float wh = head_ce<C1 + CV, CH, CT, false>(v, yh2, yh2 != CH - 1, noneh, false, lh);
      if (contributes) {
        acc_loss[0] += l1; acc_cnt[0] += w1;
        acc_loss[1] += lv; acc_cnt[1] += wv;
        acc_loss[2] += lh; acc_cnt[2] += wh;
      }
    } else {
      float wl[kNumWeak];
      if (from_boxes) {
        // _generate_rla for this pixel: one count per class over the boxes that contain it, normalised to a
        // multinomial (one IEEE division per channel), void where no box
#pragma unroll
        for (int c = 0; c < kNumWeak; ++c) wl[c] = 0.f;
        for (int k = 0; k < nbox_tile; ++k) {
          const LossBox bx = sbox[k];
          const bool in = (xc >= bx.x0) & (xc <= bx.x1) & (y >= bx.y0) & (y <= bx.y1);
#pragma unroll
          for (int c = 0; c < kNumWeak; ++c) wl[c] += (in && bx.cid == c) ? 1.f : 0.f;
        }
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < kNumWeak; ++c) s += wl[c];
        if (s > 0.5f) {
#pragma unroll
          for (int c = 0; c < kNumWeak; ++c) wl[c] = __fdiv_rn(wl[c], s);
        } else {
#pragma unroll
          for (int c = 0; c < kNumWeak; ++c) wl[c] = (c == kNumWeak - 1) ? 1.f : 0.f;
        }
        if (!live) {
#pragma unroll
          for (int c = 0; c < kNumWeak; ++c) wl[c] = 0.f;
        }
      } else if (from_vec) {
#pragma unroll
        for (int c = 0; c < kNumWeak; ++c) wl[c] = live ? simg[c] : 0.f;
      } else {
        // dense weak labels: stage the warp's 32 x 15 label floats with coalesced loads
        const float* lab = (kind == 1)
            ? a.bbox + ((int64_t)(b - a.n_strong) * a.H * a.W + (int64_t)y * a.W + warp_x0) * kNumWeak
            : a.image + ((int64_t)(b - a.n_strong - a.n_bbox) * a.H * a.W + (int64_t)y * a.W + warp_x0) * kNumWeak;
        __syncwarp();
        if (n_valid > 0) {
          const int total = n_valid * kNumWeak;
          for (int i = lane; i < total; i += 32) wl_s[i] = __ldg(lab + i);
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < kNumWeak; ++c) wl[c] = live ? wl_s[lane * kNumWeak + c] : 0.f;
      }
      // no L1 loss, but the current L1 argmax gates the L2 weights
      float best;
      int d1;
      tree_argmax_l<0, C1, CT>(v, best, d1);
#pragma unroll
      for (int k = 0; k < C1; ++k) v[k] = 0.f;
      float tv[CV], th[CH];
      const bool onv = weak_targets<CV>(selv, wl, tv) && (d1 == hier.cid_l1_vehicle);
      const bool onh = weak_targets<CH>(selh, wl, th) && (d1 == hier.cid_l1_human);
      float lv, lh;
      const float wv = head_ce<C1, CV, CT, true>(v, 0, false, tv, onv, lv);
      const float wh = head_ce<C1 + CV, CH, CT, true>(v, 0, false, th, onh, lh);
      if (contributes) {
        acc_loss[1] += lv; acc_cnt[1] += wv;
        acc_loss[2] += lh; acc_cnt[2] += wh;
      }
    }
    if (contributes) {
      const float wt = 1.0f - ly;
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        accT[c] = fmaf(wt, v[c], accT[c]);
        accB[c] = fmaf(ly, v[c], accB[c]);
      }
    }
  }
  if (trow >= 0) flush_row(trow, accT);
  if (brow >= 0) flush_row(brow, accB);
  __syncthreads();

  // flush the gradient patch (neighbouring tiles share border cells -> atomics)
  float* dst = a.dlogits + (int64_t)b * a.h * a.w * a.cp;
  for (int i = threadIdx.x; i < cells * CT; i += kColTX) {
    const float g = D[i];
    if (g != 0.f) {
      const int c = i % CT;
      const int cell = i / CT;
      const int px = cell % a.pw, py = cell / a.pw;
      const int yy = yl0 + py, xx = xl0 + px;
      if (yy < a.h && xx < a.w) atomicAdd(dst + ((int64_t)yy * a.w + xx) * a.cp + c, g);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float l = warp_sum(acc_loss[k]);
    const float n = warp_sum(acc_cnt[k]);
    if (lane == 0) { red[k][warp] = l; redc[k][warp] = n; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double l = 0.0, n = 0.0;
    for (int i = 0; i < kColTX / 32; ++i) { l += (double)red[threadIdx.x][i]; n += (double)redc[threadIdx.x][i]; }
    if (n != 0.0 || l != 0.0) {
      atomicAdd(a.sums + threadIdx.x, l);
      atomicAdd(a.counts + threadIdx.x, n);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Strong-label images, round 2 (the product path for the 14/7/3 hierarchy): the column walk of loss_cols_kernel with
// the instruction count cut to what the arithmetic needs.  ncu on the kernels above (profiles/r2_head_loss_ncu.md):
// the tile kernel retires 2370 instructions per pixel, the column kernel 845 at 25 % issue utilisation (12 warps per
// SM, 166 registers) - against ~170 for the arithmetic itself.  Here:
//   * a CTA is 128 columns x kLsTY rows and TWO warp groups: warps 0-3 walk the L1 head (14 channels), warps 4-7 the two
//     L2 heads (7 + 3) of the same columns - half the registers per thread, twice the warps, and the L2 group skips
//     every row in which no lane of its warp has a vehicle / human label (the weights are zero there: 70-90 % of a
//     street scene), which the reference computes and multiplies by zero;
//   * no arg-max (the loss needs the maximum only: FMNMX tree), exp2 with the max folded into one FFMA per class,
//     one reciprocal per head, the softmax scale folded into the two gradient accumulator weights;
//   * the one-hot part of the gradient never becomes a per-class select chain: a thread merges its vertical run of
//     equal target classes into two scalars and subtracts them from the CTA's gradient patch with four shared atomics
//     per run; the target logit is re-interpolated from the shared patch (10 instructions, bit-identical);
//   * the label tile is staged once per CTA (all rows in flight), the label -> class tables sit in shared memory.
constexpr int kLsTX = 128;
constexpr int kLsThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int LO, int HI, int CT>
__device__ __forceinline__ float tree_max(const float (&v)[CT]) {
  if constexpr (HI - LO == 1) {
    return v[LO];
  } else {
    constexpr int MID = LO + (HI - LO + 1) / 2;
    return fmaxf(tree_max<LO, MID, CT>(v), tree_max<MID, HI, CT>(v));
  }
}

// State of one head (channels [LO, LO + C) of the logits) along one output column.  The gradient of the source-row
// pair (row, row + 1) accumulates in accT / accB; when the walk moves to the pair (row + 1, row + 2), accT is complete:
// it leaves through the warp's shared staging area G (transpose of the x-interpolation, see flush_top), accB carries
// over as the new accT.
template <int LO, int C, int CT>
struct StrongHead {
  static constexpr int kPitch = 2 * C + 1;   // floats per lane in G (odd: conflict-free row writes)
  float top[C], dlt[C], accT[C], accB[C];
  float runT, runB;   // one-hot part of the current run of equal targets: sums of (1 - ly) * w and ly * w
  int run_idx;
  float loss, cnt;

  __device__ __forceinline__ void init() {
#pragma unroll
    for (int c = 0; c < C; ++c) accT[c] = accB[c] = 0.f;
    runT = runB = 0.f;
    run_idx = -1;
    loss = cnt = 0.f;
  }
  __device__ __forceinline__ void load(const float* __restrict__ rt, const float* __restrict__ rb, int o0, int o1, float lx) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float tl = rt[o0 + LO + c], tr = rt[o1 + LO + c];
      const float bl = rb[o0 + LO + c], br = rb[o1 + LO + c];
      const float t = tl + (tr - tl) * lx;   // TF ResizeBilinear: top = tl + (tr - tl) * x_lerp
      const float b = bl + (br - bl) * lx;
      top[c] = t;
      dlt[c] = b - t;
    }
  }
  // the finished run's one-hot gradient (-w on class run_idx) joins the accumulators: one select chain per RUN
  __device__ __forceinline__ void fold_run() {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const bool hit = c == run_idx;
      accT[c] -= hit ? runT : 0.f;
      accB[c] -= hit ? runB : 0.f;
    }
    runT = runB = 0.f;
  }
  // accT (gradient of source row `grow`, complete) -> global dlogits through the transpose of the x-interpolation:
  // every lane stages (1 - lx) * accT and lx * accT in G, then lane = (source column j, class c) sums the lanes of
  // run j (their left neighbour is j) and of run j - 1 (their right neighbour is j) and issues ONE global reduction.
  // xs[j] = first lane whose left neighbour is column j of the warp's span (xs[jw] = 32).  No shared atomics: a
  // float atomicAdd on shared memory is a compare-and-swap loop, and eight lanes share every source column.
  __device__ __forceinline__ void flush_top(float* __restrict__ G, const int* __restrict__ xs, int jw, int lane, float w0,
                                            float w1, float* __restrict__ gdst /* dlogits of (grow, first column), or NULL */,
                                            int cp, int cols_left) {
    fold_run_top();
    float* mine = G + lane * kPitch;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      mine[c] = w0 * accT[c];
      mine[C + c] = w1 * accT[c];
    }
    __syncwarp();
    if (gdst != nullptr) {
      for (int it = lane; it < jw * C; it += 32) {
        const int j = it / C, c = it - j * C;
        float sum = 0.f;
        for (int l = xs[j]; l < xs[j + 1]; ++l) sum += G[l * kPitch + c];
        if (j > 0)
          for (int l = xs[j - 1]; l < xs[j]; ++l) sum += G[l * kPitch + C + c];
        if (sum != 0.f && j < cols_left) atomicAdd(gdst + j * cp + LO + c, sum);
      }
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < C; ++c) { accT[c] = accB[c]; accB[c] = 0.f; }
    runT = runB;
    runB = 0.f;
  }
  __device__ __forceinline__ void fold_run_top() {
    // only the top half of the run leaves with accT; the bottom half stays with the carried accumulator
#pragma unroll
    for (int c = 0; c < C; ++c) accT[c] -= (c == run_idx) ? runT : 0.f;
    runT = 0.f;
  }
  // one pixel: idx = target class (relative to LO), w = 0 / 1
  __device__ __forceinline__ void pixel(int idx, float w, float ly, const float* __restrict__ rt, const float* __restrict__ rb,
                                        int o0, int o1, float lx) {
    float v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = top[c] + dlt[c] * ly;
    const float m = tree_max<0, C, C>(v);
    const float mb = -m * kLog2e;
    float e[C];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { e[c] = ex2_approx(fmaf(v[c], kLog2e, mb)); s += e[c]; }
    // target logit, re-interpolated (the same operations as v[idx])
    const float tl = rt[o0 + LO + idx], tr = rt[o1 + LO + idx];
    const float bl = rb[o0 + LO + idx], br = rb[o1 + LO + idx];
    const float t = tl + (tr - tl) * lx;
    const float b = bl + (br - bl) * lx;
    const float vy = t + (b - t) * ly;
    const float lse = fmaf(lg2_approx(s), kLn2, m);
    loss += w * (lse - vy);
    cnt += w;
    const float inv = __fdividef(w, s);
    const float wt = 1.0f - ly;
    const float ga = wt * inv, gb = ly * inv;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      accT[c] = fmaf(ga, e[c], accT[c]);
      accB[c] = fmaf(gb, e[c], accB[c]);
    }
    if (idx != run_idx) {
      fold_run();
      run_idx = idx;
    }
    runT = fmaf(wt, w, runT);
    runB = fmaf(ly, w, runB);
  }
};

constexpr int kLsMaxJw = 8;
constexpr int kColMaxJwWide = 8;   // source columns under one warp's 32 output columns (+ the right neighbour)

template <int C1, int CV, int CH, int kLsTY>
__global__ void __launch_bounds__(kLsThreads, 2)
loss_strong_kernel(const __grid_constant__ wlseg_hierarchy hier, const LossArgs a) {
  constexpr int CT = C1 + CV + CH;
  constexpr int CL2 = CV > CH ? CV : CH;
  constexpr int CG = C1 > CL2 ? C1 : CL2;
  constexpr int kGFloats = 32 * (2 * CG + 1);
  extern __shared__ float smem[];
  const int cells = a.ph * a.pw;
  float* patch = smem;                                          // [ph][pw][CT] logits
  float* Gall = patch + cells * CT;                             // [8 warps][32][2 * CG + 1] staging
  int32_t* slab = reinterpret_cast<int32_t*>(Gall + (kLsThreads / 32) * kGFloats);   // [kLsTY][kLsTX] labels
  int32_t* map1 = slab + kLsTY * kLsTX;                         // [80] x 3: label -> target class per head
  int32_t* mapv = map1 + 80;
  int32_t* maph = mapv + 80;
  int32_t* xsall = maph + 80;                                   // [8 warps][kLsMaxJw + 2] run starts
  __shared__ float red[3][kLsThreads / 32];
  __shared__ float redc[3][kLsThreads / 32];
  __shared__ int s_done;

  const int b = blockIdx.z;
  const int y0 = blockIdx.y * kLsTY, x0 = blockIdx.x * kLsTX;
  const int yl0 = (int)floorf(y0 * a.sy), xl0 = (int)floorf(x0 * a.sx);
  const float* src = a.logits + (int64_t)b * a.h * a.w * a.cp;
  for (int i = threadIdx.x; i < cells * CT; i += kLsThreads) {
    const int c = i % CT;
    const int cell = i / CT;
    const int px = cell % a.pw, py = cell / a.pw;
    const int yy = min(yl0 + py, a.h - 1), xx = min(xl0 + px, a.w - 1);
    patch[i] = __ldg(src + ((int64_t)yy * a.w + xx) * a.cp + c);
  }
  for (int i = threadIdx.x; i < 80; i += kLsThreads) {
    map1[i] = hier.pp_to_l1[i];
    mapv[i] = hier.pp_to_veh[i];
    maph[i] = hier.pp_to_hum[i];
  }
  if (threadIdx.x == 0) s_done = 0;
  {
    // label tile: every row in flight at once; -1 (malformed: contributes nothing) outside the image
    const int32_t* lsrc = a.strong + ((int64_t)b * a.H + y0) * a.W + x0;
    for (int i = threadIdx.x; i < kLsTY * kLsTX; i += kLsThreads) {
      const int r = i / kLsTX, cx = i % kLsTX;
      slab[i] = (x0 + cx < a.W && y0 + r < a.H) ? __ldg(lsrc + (int64_t)r * a.W + cx) : -1;
    }
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = threadIdx.x & (kLsTX - 1);
  const int role = threadIdx.x / kLsTX;          // 0: L1 head, 1: the two L2 heads (warp-uniform)
  const int x = min(x0 + col, a.W - 1);
  const float fx = x * a.sx;
  const int xl = (int)floorf(fx);
  const int xh = min(xl + 1, a.w - 1);
  const float lx = fx - (float)xl;
  const int o0 = (xl - xl0) * CT, o1 = (xh - xl0) * CT;
  // x weights of the transpose; at the right border both neighbours are the last column
  const float w0 = xh == xl ? 1.0f : 1.0f - lx, w1 = xh == xl ? 0.0f : lx;
  // the warp's span of source columns and the first lane of every run of equal left neighbours
  const int xlw = __shfl_sync(0xffffffffu, xl, 0);
  const int jw = __shfl_sync(0xffffffffu, xl, 31) - xlw + 2;   // + the right neighbour of the last run
  int* xs = xsall + warp * (kLsMaxJw + 2);
  for (int j = 0; j <= jw; ++j) {
    const int first = __popc(__ballot_sync(0xffffffffu, xl - xlw < j));
    if (lane == 0) xs[j] = first;
  }
  __syncwarp();
  float* G = Gall + warp * kGFloats;
  const int y_end = min(y0 + kLsTY, a.H);
  const int32_t* lptr = slab + col;
  const int nc = hier.num_classes;
  float* dimg = a.dlogits + (int64_t)b * a.h * a.w * a.cp;
  const int cols_left = a.w - xlw;
  float out_loss[3] = {0.f, 0.f, 0.f}, out_cnt[3] = {0.f, 0.f, 0.f};
  if (jw > kLsMaxJw) __trap();   // the host checks the upsampling factor

  auto grad_row = [&](int r) -> float* { return dimg + ((int64_t)r * a.w + xlw) * a.cp; };

  if (role == 0) {
    StrongHead<0, C1, CT> hd;
    hd.init();
    int row = -1;
    const float* rt = patch; const float* rb = patch;
    for (int y = y0; y < y_end; ++y, lptr += kLsTX) {
      const float fy = y * a.sy;
      const int yl = (int)floorf(fy);
      const float ly = fy - (float)yl;
      if (yl != row) {
        if (row >= 0) hd.flush_top(G, xs, jw, lane, w0, w1, grad_row(row), a.cp, cols_left);
        rt = patch + (yl - yl0) * a.pw * CT;
        rb = patch + (min(yl + 1, a.h - 1) - yl0) * a.pw * CT;
        hd.load(rt, rb, o0, o1, lx);
        row = yl;
      }
      const int32_t label = *lptr;
      const bool valid = (unsigned)label < (unsigned)nc;
      const int idx = valid ? map1[label] : 0;
      const float w = (valid && idx <= C1 - 2) ? 1.f : 0.f;     // the L1 void class carries no weight
      if (!__any_sync(0xffffffffu, w != 0.f)) continue;
      hd.pixel(idx, w, ly, rt, rb, o0, o1, lx);
    }
    if (row >= 0) {
      const int rowh = min(row + 1, a.h - 1);
      hd.flush_top(G, xs, jw, lane, w0, w1, grad_row(row), a.cp, cols_left);
      hd.flush_top(G, xs, jw, lane, w0, w1, grad_row(rowh), a.cp, cols_left);   // the carried bottom half
    }
    out_loss[0] = hd.loss; out_cnt[0] = hd.cnt;
  } else {
    StrongHead<C1, CV, CT> hv;
    StrongHead<C1 + CV, CH, CT> hh;
    hv.init(); hh.init();
    int row = -1;
    bool loaded_v = false, loaded_h = false;   // this source-row pair's logits are in registers
    bool dirty_v = false, dirty_h = false;     // the accumulators hold something (also carried bottoms)
    const float* rt = patch; const float* rb = patch;
    for (int y = y0; y < y_end; ++y, lptr += kLsTX) {
      const float fy = y * a.sy;
      const int yl = (int)floorf(fy);
      const float ly = fy - (float)yl;
      if (yl != row) {
        if (row >= 0) {
          // warp-uniform flags: a head that saw no weighted pixel in the last TWO source-row pairs has nothing to flush
          if (dirty_v) hv.flush_top(G, xs, jw, lane, w0, w1, grad_row(row), a.cp, cols_left);
          if (dirty_h) hh.flush_top(G, xs, jw, lane, w0, w1, grad_row(row), a.cp, cols_left);
          dirty_v = loaded_v;   // what was accumulated in this pair's bottom half is carried
          dirty_h = loaded_h;
        }
        rt = patch + (yl - yl0) * a.pw * CT;
        rb = patch + (min(yl + 1, a.h - 1) - yl0) * a.pw * CT;
        row = yl;
        loaded_v = loaded_h = false;
      }
      const int32_t label = *lptr;
      const bool valid = (unsigned)label < (unsigned)nc;
      const int iv = valid ? mapv[label] : CV - 1;
      const int ih = valid ? maph[label] : CH - 1;
      const float wv = iv != CV - 1 ? 1.f : 0.f;
      const float wh = ih != CH - 1 ? 1.f : 0.f;
      if (__any_sync(0xffffffffu, wv != 0.f)) {       // warp-uniform: rows without a vehicle pixel cost nothing
        if (!loaded_v) { hv.load(rt, rb, o0, o1, lx); loaded_v = true; dirty_v = true; }
        hv.pixel(iv, wv, ly, rt, rb, o0, o1, lx);
      }
      if (__any_sync(0xffffffffu, wh != 0.f)) {
        if (!loaded_h) { hh.load(rt, rb, o0, o1, lx); loaded_h = true; dirty_h = true; }
        hh.pixel(ih, wh, ly, rt, rb, o0, o1, lx);
      }
    }
    if (row >= 0) {
      const int rowh = min(row + 1, a.h - 1);
      if (dirty_v) {
        hv.flush_top(G, xs, jw, lane, w0, w1, grad_row(row), a.cp, cols_left);
        if (loaded_v) hv.flush_top(G, xs, jw, lane, w0, w1, grad_row(rowh), a.cp, cols_left);
      }
      if (dirty_h) {
        hh.flush_top(G, xs, jw, lane, w0, w1, grad_row(row), a.cp, cols_left);
        if (loaded_h) hh.flush_top(G, xs, jw, lane, w0, w1, grad_row(rowh), a.cp, cols_left);
      }
    }
    out_loss[1] = hv.loss; out_cnt[1] = hv.cnt;
    out_loss[2] = hh.loss; out_cnt[2] = hh.cnt;
  }

  // loss / count partials: warp shuffle -> shared; the LAST warp of the CTA to arrive adds the CTA's sums to the global
  // fp64 accumulators (no CTA-wide barrier: the two warp groups finish at different times)
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float l = warp_sum(out_loss[k]);
    const float n = warp_sum(out_cnt[k]);
    if (lane == 0) { red[k][warp] = l; redc[k][warp] = n; }
  }
  __syncwarp();
  int last = 0;
  if (lane == 0) {
    __threadfence_block();
    last = atomicAdd(&s_done, 1) == kLsThreads / 32 - 1;
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (last && lane < 3) {
    __threadfence_block();
    double l = 0.0, n = 0.0;
    for (int i = 0; i < kLsThreads / 32; ++i) { l += (double)red[lane][i]; n += (double)redc[lane][i]; }
    if (n != 0.0 || l != 0.0) {
      atomicAdd(a.sums + lane, l);
      atomicAdd(a.counts + lane, n);
    }
  }
}

template <int C1, int CV, int CH, int kLsTY>
static int launch_loss_strong(const wlseg_hierarchy* hier, LossArgs& a, cudaStream_t stream) {
  constexpr int CT = C1 + CV + CH;
  constexpr int CL2 = CV > CH ? CV : CH;
  constexpr int CG = C1 > CL2 ? C1 : CL2;
  a.ph = (int)fminf((float)a.h, floorf(kLsTY * a.sy) + 3.f);
  a.pw = (int)fminf((float)a.w, floorf(kLsTX * a.sx) + 3.f);
  if ((int)floorf(32 * a.sx) + 3 > kLsMaxJw) return 1;
  const size_t smem = ((size_t)a.ph * a.pw * CT + (size_t)(kLsThreads / 32) * 32 * (2 * CG + 1)) * sizeof(float) +
                      (size_t)(kLsTY * kLsTX + 3 * 80 + (kLsThreads / 32) * (kLsMaxJw + 2)) * sizeof(int32_t);
  if (smem > 200 * 1024) return 1;
  static bool configured = false;
  if (!configured) {
    WLSEG_CUDA(cudaFuncSetAttribute(loss_strong_kernel<C1, CV, CH, kLsTY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(a.W, kLsTX), (unsigned)ceil_div(a.H, kLsTY), (unsigned)a.n_strong);
  loss_strong_kernel<C1, CV, CH, kLsTY><<<grid, kLsThreads, smem, stream>>>(*hier, a);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Wide hierarchies (Vistas 53 / 12 / 5): the same column walk with a head SPLIT INTO CHUNKS of <= kChunk classes, one
// warp per chunk over the same 32 output columns, so that no thread holds more than 5 x kChunk gradient / logit
// registers.  The softmax of a split head needs the maximum and the sum over all of its chunks: every chunk warp
// publishes its local (max, sum of exp) per pixel in shared memory, one named barrier per row joins the warps of the
// head, and each rescales its own exponentials - exp(m_own - M) / S.  Only the chunk that holds the target class adds
// the pixel's loss and the one-hot gradient.  A CTA is 32 columns x kLwTY rows x (number of chunks) warps.
constexpr int kChunk = 11;
constexpr int kLwTY = 32;
constexpr int kLwMaxWarps = 8;

__device__ __forceinline__ void named_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct ChunkRole { int head_lo, head_c, lo, c, head_id, peers, peer0; };   // peer0: first warp of the head

// one chunk [LO, LO + C) of a head along one output column; see StrongHead for the accumulator protocol
template <int C>
struct ChunkWalk {
  float top[C], dlt[C], accT[C], accB[C];
  float runT, runB;
  int run_idx;     // target class relative to the chunk, -1 = not in this chunk
  float loss, cnt;
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int c = 0; c < C; ++c) top[c] = dlt[c] = accT[c] = accB[c] = 0.f;
    runT = runB = loss = cnt = 0.f;
    run_idx = -1;
  }
};

template <int C>
__device__ __forceinline__ void chunk_load(ChunkWalk<C>& k, int nc, const float* __restrict__ rt, const float* __restrict__ rb,
                                           int o0, int o1, int lo, float lx) {
#pragma unroll
  for (int c = 0; c < C; ++c) {
    if (c < nc) {
      const float tl = rt[o0 + lo + c], tr = rt[o1 + lo + c];
      const float bl = rb[o0 + lo + c], br = rb[o1 + lo + c];
      const float t = tl + (tr - tl) * lx;
      const float b = bl + (br - bl) * lx;
      k.top[c] = t;
      k.dlt[c] = b - t;
    }
  }
}

template <int C>
__device__ __forceinline__ void chunk_flush_top(ChunkWalk<C>& k, int nc, float* __restrict__ G, const int* __restrict__ xs, int jw,
                                                int lane, float w0, float w1, float* __restrict__ gdst, int cp, int cols_left,
                                                int lo) {
  constexpr int kPitch = 2 * C + 1;
#pragma unroll
  for (int c = 0; c < C; ++c) k.accT[c] -= (c == k.run_idx) ? k.runT : 0.f;
  k.runT = 0.f;
  float* mine = G + lane * kPitch;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    mine[c] = w0 * k.accT[c];
    mine[C + c] = w1 * k.accT[c];
  }
  __syncwarp();
  for (int it = lane; it < jw * nc; it += 32) {
    const int j = it / nc, c = it - j * nc;
    float sum = 0.f;
    for (int l = xs[j]; l < xs[j + 1]; ++l) sum += G[l * kPitch + c];
    if (j > 0)
      for (int l = xs[j - 1]; l < xs[j]; ++l) sum += G[l * kPitch + C + c];
    if (sum != 0.f && j < cols_left) atomicAdd(gdst + j * cp + lo + c, sum);
  }
  __syncwarp();
#pragma unroll
  for (int c = 0; c < C; ++c) { k.accT[c] = k.accB[c]; k.accB[c] = 0.f; }
  k.runT = k.runB;
  k.runB = 0.f;
}

__global__ void __launch_bounds__(32 * kLwMaxWarps, 2)
loss_strong_wide_kernel(const __grid_constant__ wlseg_hierarchy hier, const LossArgs a, int n_warps) {
  extern __shared__ float smem[];
  const int CT = hier.C1 + hier.Cv + hier.Ch;
  const int cells = a.ph * a.pw;
  constexpr int kGFloats = 32 * (2 * kChunk + 1);
  float* patch = smem;                                          // [ph][pw][CT] logits
  float* Gall = patch + cells * CT;                             // [n_warps][32][2 * kChunk + 1] staging
  float* exch = Gall + n_warps * kGFloats;                      // [2 row parities][n_warps][32][2]: local max, sum of exp
  int32_t* slab = reinterpret_cast<int32_t*>(exch + 2 * n_warps * 64);   // [kLwTY][32] labels
  int32_t* map1 = slab + kLwTY * 32;
  int32_t* mapv = map1 + 80;
  int32_t* maph = mapv + 80;
  __shared__ int xs[kColMaxJwWide + 2];
  __shared__ float red[3][kLwMaxWarps];
  __shared__ float redc[3][kLwMaxWarps];
  __shared__ int s_done;
  __shared__ ChunkRole roles[kLwMaxWarps];

  const int nthreads = 32 * n_warps;
  const int b = blockIdx.z;
  const int y0 = blockIdx.y * kLwTY, x0 = blockIdx.x * 32;
  const int yl0 = (int)floorf(y0 * a.sy), xl0 = (int)floorf(x0 * a.sx);
  const float* src = a.logits + (int64_t)b * a.h * a.w * a.cp;
  for (int i = threadIdx.x; i < cells * CT; i += nthreads) {
    const int c = i % CT;
    const int cell = i / CT;
    const int px = cell % a.pw, py = cell / a.pw;
    const int yy = min(yl0 + py, a.h - 1), xx = min(xl0 + px, a.w - 1);
    patch[i] = __ldg(src + ((int64_t)yy * a.w + xx) * a.cp + c);
  }
  for (int i = threadIdx.x; i < 80; i += nthreads) {
    map1[i] = hier.pp_to_l1[i];
    mapv[i] = hier.pp_to_veh[i];
    maph[i] = hier.pp_to_hum[i];
  }
  if (threadIdx.x == 0) {
    s_done = 0;
    // warp -> chunk of a head, in head order
    int wi = 0;
    const int hlo[3] = {0, hier.C1, hier.C1 + hier.Cv}, hc[3] = {hier.C1, hier.Cv, hier.Ch};
    for (int hd = 0; hd < 3; ++hd) {
      const int nch = (hc[hd] + kChunk - 1) / kChunk;
      const int per = (hc[hd] + nch - 1) / nch;   // balanced chunks
      for (int q = 0; q < nch; ++q, ++wi) {
        roles[wi].head_lo = hlo[hd]; roles[wi].head_c = hc[hd]; roles[wi].head_id = hd;
        roles[wi].lo = q * per; roles[wi].c = min(per, hc[hd] - q * per);
        roles[wi].peers = nch; roles[wi].peer0 = wi - q;
      }
    }
  }
  {
    const int32_t* lsrc = a.strong + ((int64_t)b * a.H + y0) * a.W + x0;
    for (int i = threadIdx.x; i < kLwTY * 32; i += nthreads) {
      const int r = i >> 5, cx = i & 31;
      slab[i] = (x0 + cx < a.W && y0 + r < a.H) ? __ldg(lsrc + (int64_t)r * a.W + cx) : -1;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = min(x0 + lane, a.W - 1);
  const float fx = x * a.sx;
  const int xl = (int)floorf(fx);
  const int xh = min(xl + 1, a.w - 1);
  const float lx = fx - (float)xl;
  const float w0 = xh == xl ? 1.0f : 1.0f - lx, w1 = xh == xl ? 0.0f : lx;
  const int xlw = __shfl_sync(0xffffffffu, xl, 0);
  const int jw = __shfl_sync(0xffffffffu, xl, 31) - xlw + 2;
  if (warp == 0) {
    for (int j = 0; j <= jw && j <= kColMaxJwWide + 1; ++j) {
      const int first = __popc(__ballot_sync(0xffffffffu, xl - xlw < j));
      if (lane == 0) xs[j] = first;
    }
  }
  __syncthreads();
  if (jw > kColMaxJwWide) __trap();

  const ChunkRole role = roles[warp];
  const int lo = role.head_lo + role.lo;     // first channel of this chunk in the logits
  const int nc = role.c;
  const int o0 = (xl - xl0) * CT, o1 = (xh - xl0) * CT;
  const int32_t* mp = role.head_id == 0 ? map1 : (role.head_id == 1 ? mapv : maph);
  const int void_idx = role.head_c - 1;      // every head's last class is its void / "other" class: no weight
  float* G = Gall + warp * kGFloats;
  float* dimg = a.dlogits + (int64_t)b * a.h * a.w * a.cp;
  const int cols_left = a.w - xlw;
  const int ncls = hier.num_classes;
  const int y_end = min(y0 + kLwTY, a.H);
  const int32_t* lptr = slab + lane;
  auto grad_row = [&](int r) -> float* { return dimg + ((int64_t)r * a.w + xlw) * a.cp; };

  ChunkWalk<kChunk> k;
  k.init();
  int row = -1;
  bool loaded = false, dirty = false;
  int parity = 0;
  const float* rt = patch; const float* rb = patch;
  for (int y = y0; y < y_end; ++y, lptr += 32) {
    const float fy = y * a.sy;
    const int yl = (int)floorf(fy);
    const float ly = fy - (float)yl;
    if (yl != row) {
      if (row >= 0) {
        if (dirty) chunk_flush_top(k, nc, G, xs, jw, lane, w0, w1, grad_row(row), a.cp, cols_left, lo);
        dirty = loaded;
      }
      rt = patch + (yl - yl0) * a.pw * CT;
      rb = patch + (min(yl + 1, a.h - 1) - yl0) * a.pw * CT;
      row = yl;
      loaded = false;
    }
    const int32_t label = *lptr;
    const bool valid = (unsigned)label < (unsigned)ncls;
    const int idx = valid ? mp[label] : void_idx;          // target class of this head
    const float w = idx != void_idx ? 1.f : 0.f;
    if (!__any_sync(0xffffffffu, w != 0.f)) continue;     // identical in every chunk warp of the head: same labels
    if (!loaded) { chunk_load(k, nc, rt, rb, o0, o1, lo, lx); loaded = true; dirty = true; }
    float v[kChunk];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < kChunk; ++c) {
      v[c] = k.top[c] + k.dlt[c] * ly;
      if (c < nc) m = fmaxf(m, v[c]);
    }
    const float mb = -m * kLog2e;
    float e[kChunk];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kChunk; ++c) {
      e[c] = c < nc ? ex2_approx(fmaf(v[c], kLog2e, mb)) : 0.f;
      s += e[c];
    }
    float M = m, S = s, f = 1.0f;
    if (role.peers > 1) {
      float* ex = exch + ((parity * n_warps + warp) * 32 + lane) * 2;
      ex[0] = m; ex[1] = s;
      named_barrier(1 + role.head_id, 32 * role.peers);
      S = 0.f;
      for (int q = 0; q < role.peers; ++q) M = fmaxf(M, exch[((parity * n_warps + role.peer0 + q) * 32 + lane) * 2]);
      for (int q = 0; q < role.peers; ++q) {
        const float* o = exch + ((parity * n_warps + role.peer0 + q) * 32 + lane) * 2;
        S = fmaf(o[1], ex2_approx((o[0] - M) * kLog2e), S);
      }
      f = ex2_approx((m - M) * kLog2e);
      parity ^= 1;
    }
    const int rel = idx - role.lo;                         // target class relative to this chunk
    const bool mine = rel >= 0 && rel < nc;
    const float wm = mine ? w : 0.f;
    {
      // target logit, re-interpolated from the shared patch (any chunk could; the owner does)
      const int ch = role.head_lo + idx;
      const float tl = rt[o0 + ch], tr = rt[o1 + ch];
      const float bl = rb[o0 + ch], br = rb[o1 + ch];
      const float t = tl + (tr - tl) * lx;
      const float bb = bl + (br - bl) * lx;
      const float vy = t + (bb - t) * ly;
      const float lse = fmaf(lg2_approx(S), kLn2, M);
      k.loss += wm * (lse - vy);
      k.cnt += wm;
    }
    const float inv = __fdividef(w * f, S);
    const float wt = 1.0f - ly;
    const float ga = wt * inv, gb = ly * inv;
#pragma unroll
    for (int c = 0; c < kChunk; ++c) {
      k.accT[c] = fmaf(ga, e[c], k.accT[c]);
      k.accB[c] = fmaf(gb, e[c], k.accB[c]);
    }
    const int ridx = mine ? rel : -1;
    if (ridx != k.run_idx) {
#pragma unroll
      for (int c = 0; c < kChunk; ++c) {
        const bool hit = c == k.run_idx;
        k.accT[c] -= hit ? k.runT : 0.f;
        k.accB[c] -= hit ? k.runB : 0.f;
      }
      k.runT = k.runB = 0.f;
      k.run_idx = ridx;
    }
    k.runT = fmaf(wt, wm, k.runT);
    k.runB = fmaf(ly, wm, k.runB);
  }
  if (row >= 0 && dirty) {
    const int rowh = min(row + 1, a.h - 1);
    chunk_flush_top(k, nc, G, xs, jw, lane, w0, w1, grad_row(row), a.cp, cols_left, lo);
    if (loaded) chunk_flush_top(k, nc, G, xs, jw, lane, w0, w1, grad_row(rowh), a.cp, cols_left, lo);
  }
  {
    const float l = warp_sum(k.loss);
    const float n = warp_sum(k.cnt);
    if (lane == 0) { red[role.head_id][warp] = l; redc[role.head_id][warp] = n; }
  }
  __syncwarp();
  int last = 0;
  if (lane == 0) {
    __threadfence_block();
    last = atomicAdd(&s_done, 1) == n_warps - 1;
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (last && lane < 3) {
    __threadfence_block();
    double l = 0.0, n = 0.0;
    for (int i = 0; i < n_warps; ++i)
      if (roles[i].head_id == lane) { l += (double)red[lane][i]; n += (double)redc[lane][i]; }
    if (n != 0.0 || l != 0.0) {
      atomicAdd(a.sums + lane, l);
      atomicAdd(a.counts + lane, n);
    }
  }
}

static int launch_loss_strong_wide(const wlseg_hierarchy* hier, LossArgs& a, cudaStream_t stream) {
  const int CT = hier->C1 + hier->Cv + hier->Ch;
  int n_warps = 0;
  for (int c : {hier->C1, hier->Cv, hier->Ch}) n_warps += (c + kChunk - 1) / kChunk;
  if (n_warps > kLwMaxWarps) return 1;
  a.ph = (int)fminf((float)a.h, floorf(kLwTY * a.sy) + 3.f);
  a.pw = (int)fminf((float)a.w, floorf(32 * a.sx) + 3.f);
  if ((int)floorf(32 * a.sx) + 3 > kColMaxJwWide) return 1;
  const size_t smem = ((size_t)a.ph * a.pw * CT + (size_t)n_warps * 32 * (2 * kChunk + 1) + (size_t)2 * n_warps * 64) * sizeof(float) +
                      (size_t)(kLwTY * 32 + 3 * 80) * sizeof(int32_t);
  if (smem > 200 * 1024) return 1;
  static bool configured = false;
  if (!configured) {
    WLSEG_CUDA(cudaFuncSetAttribute(loss_strong_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(a.W, 32), (unsigned)ceil_div(a.H, kLwTY), (unsigned)a.n_strong);
  loss_strong_wide_kernel<<<grid, 32 * n_warps, smem, stream>>>(*hier, a, n_warps);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

template <int C1, int CV, int CH, int kColTY>
static int launch_loss_cols(const wlseg_hierarchy* hier, LossArgs& a, int first, int count, cudaStream_t stream) {
  // images [first, first + count) of the batch; first < n_strong selects the strong-label instantiation
  constexpr int CT = C1 + CV + CH;
  a.ph = (int)fminf((float)a.h, floorf(kColTY * a.sy) + 3.f);
  a.pw = (int)fminf((float)a.w, floorf(kColTX * a.sx) + 3.f);
  const size_t smem = ((size_t)2 * a.ph * a.pw * CT + (kColTX / 32) * 32 * kNumWeak) * sizeof(float);
  if (smem > 200 * 1024) return 1;
  static bool configured = false;
  if (!configured) {
    WLSEG_CUDA(cudaFuncSetAttribute(loss_cols_kernel<C1, CV, CH, false, kColTY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    WLSEG_CUDA(cudaFuncSetAttribute(loss_cols_kernel<C1, CV, CH, true, kColTY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(a.W, kColTX), (unsigned)ceil_div(a.H, kColTY), (unsigned)count);
  if (first < a.n_strong)
    loss_cols_kernel<C1, CV, CH, false, kColTY><<<grid, kColTX, smem, stream>>>(*hier, a);
  else
    loss_cols_kernel<C1, CV, CH, true, kColTY><<<grid, kColTX, smem, stream>>>(*hier, a);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

__global__ void loss_finalize_kernel(int C1, int Cv, int Ch, int cp, const double* __restrict__ sums,
                                     const double* __restrict__ counts, float l2_coef, float grad_scale,
                                     float* __restrict__ dlogits, int64_t n_pix, float* __restrict__ losses) {
  // SUM_BY_NONZERO_WEIGHTS with safe-div (tf.losses.compute_weighted_loss)
  float inv[3], l[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double n = counts[k];
    inv[k] = n > 0.0 ? (float)(1.0 / n) : 0.f;
    l[k] = n > 0.0 ? (float)(sums[k] / n) : 0.f;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && losses != nullptr) {
    losses[0] = l[0];
    losses[1] = l[1];
    losses[2] = l[2];
    losses[3] = l[0] + l2_coef * (l[1] + l[2]);
  }
  if (dlogits == nullptr) return;
  const int Ct = C1 + Cv + Ch;
  const float s1 = grad_scale * inv[0], sv = grad_scale * l2_coef * inv[1], sh = grad_scale * l2_coef * inv[2];
  const int64_t total = n_pix * cp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % cp);
    float s = c < C1 ? s1 : (c < C1 + Cv ? sv : (c < Ct ? sh : 0.f));
    dlogits[i] *= s;
  }
}

}  // namespace wlseg

using namespace wlseg;

static int env_ty() {
  const char* e = getenv("WLSEG_LOSS_TY");   // rows per CTA strip of loss_strong_kernel: 16 | 32
  return e != nullptr ? atoi(e) : 32;
}

static int loss_impl(const wlseg_hierarchy* hier, const float* logits, int32_t logits_pitch, int32_t n_strong, int32_t n_bbox,
                     int32_t n_image, int32_t h, int32_t w, int32_t H, int32_t W, const int32_t* strong_labels,
                     const float* bbox_labels, const float* image_labels, const float* box_coords, const int32_t* box_cids,
                     int32_t max_boxes, const float* image_vectors, double* sums, double* counts, float* dlogits,
                     wlseg_stream_t stream) {
  const bool lists = box_coords != nullptr || image_vectors != nullptr;
  if (int e = check_hierarchy(hier)) return e;
  WLSEG_CHECK_ARG(n_strong >= 0 && n_bbox >= 0 && n_image >= 0, "loss: negative batch part");
  WLSEG_CHECK_ARG(h > 0 && w > 0 && H >= h && W >= w, "loss: expects upsampling (h,w)=(%d,%d) -> (H,W)=(%d,%d)", h, w, H, W);
  WLSEG_CHECK_ARG(hier->num_classes > 0 && hier->num_classes <= 80, "loss: num_classes out of range");
  const int B = n_strong + n_bbox + n_image;
  if (B == 0) return 0;
  WLSEG_CHECK_ARG(logits && sums && counts && dlogits, "loss: null pointer");
  WLSEG_CHECK_ARG(n_strong == 0 || strong_labels, "loss: strong labels missing");
  WLSEG_CHECK_ARG(n_bbox == 0 || bbox_labels || (box_coords && box_cids && max_boxes >= 0), "loss: bbox labels missing");
  WLSEG_CHECK_ARG(n_image == 0 || image_labels || image_vectors, "loss: image labels missing");
  WLSEG_CHECK_ARG(max_boxes <= kLossMaxBoxes, "loss: more than %d boxes per image", kLossMaxBoxes);
  WLSEG_CHECK_ARG(B <= 65535, "loss: batch too large");
  WLSEG_CHECK_ARG(logits_pitch >= hier->C1 + hier->Cv + hier->Ch, "loss: logits_pitch %d < channels", logits_pitch);
  LossArgs a;
  a.logits = logits;
  a.n_strong = n_strong; a.n_bbox = n_bbox; a.n_image = n_image;
  a.h = h; a.w = w; a.H = H; a.W = W;
  a.cp = logits_pitch;
  a.sy = resize_scale(h, H);
  a.sx = resize_scale(w, W);
  const int Ct = hier->C1 + hier->Cv + hier->Ch;
  a.gs = Ct | 1;
  a.strong = strong_labels; a.bbox = bbox_labels; a.image = image_labels;
  a.sums = sums; a.counts = counts; a.dlogits = dlogits;
  a.box_coords = box_coords; a.box_cids = box_cids; a.max_boxes = max_boxes; a.image_vec = image_vectors;
  // Cityscapes hierarchy (14/7/3), upsampling >= 2x: the strong images take loss_strong_kernel, the WEAK images
  // (60 B/pixel of labels, targets by segment sums) the register-resident column-walking loss_cols_kernel (measured
  // 1.3x faster than the generic tile kernel there).  The 70-channel Vistas hierarchy takes the generic kernel.
  int g0 = 0, g1 = B;   // images [g0, g1) are left to the generic tile kernel
  a.first = 0;
  if (a.sy <= 0.5f && a.sx <= 0.5f && hier->C1 == 14 && hier->Cv == 7 && hier->Ch == 3) {
    // WLSEG_LOSS_COLS (experiments): "all" = strong images on loss_cols_kernel too, "none" = everything on the tile
    // kernel, "old" = round-1 split (strong on the tile kernel, weak on loss_cols_kernel), "ty16" = 16-row weak strips
    const char* mode = getenv("WLSEG_LOSS_COLS");
    const bool all = mode != nullptr && mode[0] == 'a';
    const bool none = mode != nullptr && mode[0] == 'n';
    const bool old = mode != nullptr && mode[0] == 'o';
    const bool ty16 = mode != nullptr && mode[0] == 't';
    if (!none) {
      if (n_strong > 0 && !old) {
        const int rc = all ? launch_loss_cols<14, 7, 3, 32>(hier, a, 0, n_strong, (cudaStream_t)stream)
                           : (env_ty() == 16 ? launch_loss_strong<14, 7, 3, 16>(hier, a, (cudaStream_t)stream)
                                             : launch_loss_strong<14, 7, 3, 32>(hier, a, (cudaStream_t)stream));
        if (rc > 1 || rc < 0) return rc;
        if (rc == 0) g0 = n_strong;
      }
      if (B - n_strong > 0) {
        const int rc = ty16 ? launch_loss_cols<14, 7, 3, 16>(hier, a, n_strong, B - n_strong, (cudaStream_t)stream)
                            : launch_loss_cols<14, 7, 3, 32>(hier, a, n_strong, B - n_strong, (cudaStream_t)stream);
        if (rc > 1 || rc < 0) return rc;
        if (rc == 0) g1 = n_strong;
      }
    }
  }
  else if (a.sy <= 0.5f && a.sx <= 0.5f && n_strong > 0 && getenv("WLSEG_LOSS_WIDE_OFF") == nullptr) {
    // wider hierarchies (Vistas 53 / 12 / 5): the strong images take the chunked column walk, weak images the tile kernel
    const int rc = launch_loss_strong_wide(hier, a, (cudaStream_t)stream);
    if (rc > 1 || rc < 0) return rc;
    if (rc == 0) g0 = n_strong;
  }
  // the compact weak labels exist in the column-walking kernel only
  WLSEG_CHECK_ARG(!lists || g1 <= n_strong,
                  "loss(lists): box / class lists need the 14/7/3 hierarchy at >= 2x upsampling; rasterise them "
                  "(wlseg_rasterize_bbox_labels, wlseg_tile_image_labels) and call wlseg_loss_fwd_bwd");
  if (g0 >= g1) return 0;
  a.first = g0;
  const int n_generic = g1 - g0;
  a.ph = (int)fminf((float)h, floorf(kLossTY * a.sy) + 3.f);
  a.pw = (int)fminf((float)w, floorf(kLossTX * a.sx) + 3.f);
  size_t smem = (size_t)(2 * a.ph * a.pw * Ct + 2 * kLossTX * a.gs + 2 * kLossTX + a.pw + 1) * sizeof(float);
  WLSEG_CHECK_ARG(smem <= 200 * 1024, "loss: tile state (%zu B) does not fit shared memory", smem);
  if (smem > 48 * 1024)
    WLSEG_CUDA(cudaFuncSetAttribute(loss_fwd_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(W, kLossTX), (unsigned)ceil_div(H, kLossTY), (unsigned)n_generic);
  loss_fwd_bwd_kernel<<<grid, kLossThreads, smem, (cudaStream_t)stream>>>(*hier, a);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int wlseg_loss_fwd_bwd(const wlseg_hierarchy* hier, const float* logits, int32_t logits_pitch,
                                  int32_t n_strong, int32_t n_bbox, int32_t n_image, int32_t h, int32_t w, int32_t H, int32_t W,
                                  const int32_t* strong_labels, const float* bbox_labels, const float* image_labels,
                                  double* sums, double* counts, float* dlogits, wlseg_stream_t stream) {
  return loss_impl(hier, logits, logits_pitch, n_strong, n_bbox, n_image, h, w, H, W, strong_labels, bbox_labels, image_labels,
                   nullptr, nullptr, 0, nullptr, sums, counts, dlogits, stream);
}

extern "C" int wlseg_loss_fwd_bwd_lists(const wlseg_hierarchy* hier, const float* logits, int32_t logits_pitch,
                                        int32_t n_strong, int32_t n_bbox, int32_t n_image, int32_t h, int32_t w, int32_t H,
                                        int32_t W, const int32_t* strong_labels, const float* box_coords,
                                        const int32_t* box_cids, int32_t max_boxes, const float* image_vectors,
                                        double* sums, double* counts, float* dlogits, wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(n_bbox == 0 || (box_coords && box_cids), "loss(lists): box lists missing");
  WLSEG_CHECK_ARG(n_image == 0 || image_vectors, "loss(lists): image-level class vectors missing");
  return loss_impl(hier, logits, logits_pitch, n_strong, n_bbox, n_image, h, w, H, W, strong_labels, nullptr, nullptr,
                   box_coords, box_cids, max_boxes, image_vectors, sums, counts, dlogits, stream);
}

extern "C" int wlseg_loss_finalize(const wlseg_hierarchy* hier, const double* sums, const double* counts,
                                   float l2_coef, float grad_scale, float* dlogits, int32_t logits_pitch,
                                   int64_t n_lowres_pixels, float* losses, wlseg_stream_t stream) {
  if (int e = check_hierarchy(hier)) return e;
  WLSEG_CHECK_ARG(sums && counts, "loss_finalize: null sums / counts");
  WLSEG_CHECK_ARG(n_lowres_pixels >= 0, "loss_finalize: negative size");
  const int Ct = hier->C1 + hier->Cv + hier->Ch;
  WLSEG_CHECK_ARG(logits_pitch >= Ct, "loss_finalize: logits_pitch %d < channels", logits_pitch);
  int grid = dlogits ? bw_grid(n_lowres_pixels * logits_pitch, 256, 4) : 1;
  loss_finalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(hier->C1, hier->Cv, hier->Ch, logits_pitch, sums, counts, l2_coef,
                                                              grad_scale, dlogits, n_lowres_pixels, losses);
  WLSEG_LAUNCH_CHECK();
  return 0;
}
