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
};

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

  const int b = blockIdx.z;
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

extern "C" int wlseg_loss_fwd_bwd(const wlseg_hierarchy* hier, const float* logits, int32_t logits_pitch,
                                  int32_t n_strong, int32_t n_bbox, int32_t n_image, int32_t h, int32_t w, int32_t H, int32_t W,
                                  const int32_t* strong_labels, const float* bbox_labels, const float* image_labels,
                                  double* sums, double* counts, float* dlogits, wlseg_stream_t stream) {
  if (int e = check_hierarchy(hier)) return e;
  WLSEG_CHECK_ARG(n_strong >= 0 && n_bbox >= 0 && n_image >= 0, "loss: negative batch part");
  WLSEG_CHECK_ARG(h > 0 && w > 0 && H >= h && W >= w, "loss: expects upsampling (h,w)=(%d,%d) -> (H,W)=(%d,%d)", h, w, H, W);
  WLSEG_CHECK_ARG(hier->num_classes > 0 && hier->num_classes <= 80, "loss: num_classes out of range");
  const int B = n_strong + n_bbox + n_image;
  if (B == 0) return 0;
  WLSEG_CHECK_ARG(logits && sums && counts && dlogits, "loss: null pointer");
  WLSEG_CHECK_ARG(n_strong == 0 || strong_labels, "loss: strong labels missing");
  WLSEG_CHECK_ARG(n_bbox == 0 || bbox_labels, "loss: bbox labels missing");
  WLSEG_CHECK_ARG(n_image == 0 || image_labels, "loss: image labels missing");
  WLSEG_CHECK_ARG(B <= 65535, "loss: batch too large");
  WLSEG_CHECK_ARG(logits_pitch >= hier->C1 + hier->Cv + hier->Ch, "loss: logits_pitch %d < channels", logits_pitch);
  LossArgs a;
  a.logits = logits;
  a.n_strong = n_strong; a.n_bbox = n_bbox; a.n_image = n_image;
  a.h = h; a.w = w; a.H = H; a.W = W;
  a.cp = logits_pitch;
  a.sy = resize_scale(h, H);
  a.sx = resize_scale(w, W);
  a.ph = (int)fminf((float)h, floorf(kLossTY * a.sy) + 3.f);
  a.pw = (int)fminf((float)w, floorf(kLossTX * a.sx) + 3.f);
  const int Ct = hier->C1 + hier->Cv + hier->Ch;
  a.gs = Ct | 1;
  a.strong = strong_labels; a.bbox = bbox_labels; a.image = image_labels;
  a.sums = sums; a.counts = counts; a.dlogits = dlogits;
  size_t smem = (size_t)(2 * a.ph * a.pw * Ct + 2 * kLossTX * a.gs + 2 * kLossTX + a.pw + 1) * sizeof(float);
  WLSEG_CHECK_ARG(smem <= 200 * 1024, "loss: tile state (%zu B) does not fit shared memory", smem);
  if (smem > 48 * 1024)
    WLSEG_CUDA(cudaFuncSetAttribute(loss_fwd_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(W, kLossTX), (unsigned)ceil_div(H, kLossTY), (unsigned)B);
  loss_fwd_bwd_kernel<<<grid, kLossThreads, smem, (cudaStream_t)stream>>>(*hier, a);
  WLSEG_LAUNCH_CHECK();
  return 0;
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
