// Confusion-matrix histogram: cm[label, decision] += 1 (int64, bit-exact).
//
// Replaces [TF-1.12] metrics_impl._streaming_confusion_matrix reached from
// code/estimator/define_estimator_hierarchical.py:185-194 and tf.confusion_matrix in
// code/estimator/define_metrics.py:10-12; the optional LUT is the decisions part of
// _map_predictions_to_new_cids (define_estimator_hierarchical.py:511-514).
//
// HBM-bound integer work: 8 B/pixel read (int32 label + int32 decision), 8*C^2 B written.
// Each CTA keeps a private int32 C x C histogram in shared memory; lanes of a warp that hit
// the same bin are merged with match.any so a segmentation map (long runs of one class) costs
// one shared atomic per distinct bin per warp instead of 32 serialised ones.  The private
// histograms are flushed with 64-bit global atomics (integer => order independent => exact).
#include "common.cuh"

namespace wlseg {

constexpr int kConfmatThreads = 256;

__device__ __forceinline__ void confmat_vote(int bin, int32_t* hist, unsigned& bad) {
  // bin < 0: out-of-range pair, skipped and counted
  unsigned peers = __match_any_sync(0xffffffffu, bin);
  int leader = __ffs(peers) - 1;
  if ((int)(threadIdx.x & 31) == leader) {
    if (bin >= 0) atomicAdd(&hist[bin], __popc(peers));
    else if (bin == -1) bad += __popc(peers);
  }
}

__device__ __forceinline__ int confmat_bin(int32_t l, int32_t d, int C, const int32_t* __restrict__ lut,
                                           int lut_size) {
  if (lut != nullptr) {
    if ((unsigned)d >= (unsigned)lut_size) return -1;
    d = __ldg(lut + d);
  }
  if ((unsigned)l >= (unsigned)C || (unsigned)d >= (unsigned)C) return -1;
  return l * C + d;
}

template <bool kVec4>
__global__ void __launch_bounds__(kConfmatThreads)
confmat_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ decisions, int64_t n, int C,
               const int32_t* __restrict__ lut, int lut_size, unsigned long long* __restrict__ cm,
               unsigned long long* __restrict__ invalid) {
  extern __shared__ int32_t hist[];
  const int bins = C * C;
  for (int i = threadIdx.x; i < bins; i += blockDim.x) hist[i] = 0;
  __syncthreads();

  unsigned bad = 0;
  const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;

  if (kVec4) {
    const int64_t nvec = n >> 2;
    const int4* l4 = reinterpret_cast<const int4*>(labels);
    const int4* d4 = reinterpret_cast<const int4*>(decisions);
    // warp-uniform trip count so that match.any always sees the full warp; -2 = idle lane
    for (int64_t base = gtid - lane; base < nvec; base += gstride) {
      int64_t v = base + lane;
      int b0 = -2, b1 = -2, b2 = -2, b3 = -2;
      if (v < nvec) {
        int4 l = __ldg(l4 + v);
        int4 d = __ldg(d4 + v);
        b0 = confmat_bin(l.x, d.x, C, lut, lut_size);
        b1 = confmat_bin(l.y, d.y, C, lut, lut_size);
        b2 = confmat_bin(l.z, d.z, C, lut, lut_size);
        b3 = confmat_bin(l.w, d.w, C, lut, lut_size);
      }
      confmat_vote(b0, hist, bad);
      confmat_vote(b1, hist, bad);
      confmat_vote(b2, hist, bad);
      confmat_vote(b3, hist, bad);
    }
    // tail (< 4 elements), handled by lane-0..2 of the first warp of CTA 0
    if (blockIdx.x == 0 && threadIdx.x < 32) {
      int64_t i = (nvec << 2) + lane;
      int b = -2;
      if (i < n) b = confmat_bin(labels[i], decisions[i], C, lut, lut_size);
      confmat_vote(b, hist, bad);
    }
  } else {
    for (int64_t base = gtid - lane; base < n; base += gstride) {
      int64_t i = base + lane;
      int b = -2;
      if (i < n) b = confmat_bin(__ldg(labels + i), __ldg(decisions + i), C, lut, lut_size);
      confmat_vote(b, hist, bad);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += blockDim.x) {
    int32_t c = hist[i];
    if (c) atomicAdd(cm + i, (unsigned long long)c);
  }
  if (invalid != nullptr && bad) atomicAdd(invalid, (unsigned long long)bad);
}

}  // namespace wlseg

using namespace wlseg;

extern "C" int wlseg_confmat_accumulate(const int32_t* labels, const int32_t* decisions, int64_t n,
                                        int32_t num_classes, const int32_t* lut, int32_t lut_size,
                                        int64_t* cm, int64_t* invalid, wlseg_stream_t stream) {
  WLSEG_CHECK_ARG(cm != nullptr, "confmat: cm is NULL");
  WLSEG_CHECK_ARG(num_classes > 0 && num_classes <= 104, "confmat: num_classes %d out of (0, 104]", num_classes);
  WLSEG_CHECK_ARG(n >= 0, "confmat: negative n");
  WLSEG_CHECK_ARG(lut == nullptr || lut_size > 0, "confmat: lut given with lut_size %d", lut_size);
  if (n == 0) return 0;  // empty batch: nothing to add
  WLSEG_CHECK_ARG(labels && decisions, "confmat: null labels / decisions");
  const size_t smem = (size_t)num_classes * num_classes * sizeof(int32_t);
  const bool vec = ((((uintptr_t)labels) | ((uintptr_t)decisions)) & 15) == 0;
  const int64_t items = vec ? (n >> 2) + 1 : n;
  int grid = bw_grid(items, kConfmatThreads, 4);
  // int32 private bins cannot overflow: one CTA sees at most ceil(n / grid) * 4 pixels
  WLSEG_CHECK_ARG(n / grid < (int64_t)1 << 30, "confmat: n too large for one call");
  cudaStream_t s = (cudaStream_t)stream;
  if (vec)
    confmat_kernel<true><<<grid, kConfmatThreads, smem, s>>>(labels, decisions, n, num_classes, lut, lut_size,
                                                             (unsigned long long*)cm, (unsigned long long*)invalid);
  else
    confmat_kernel<false><<<grid, kConfmatThreads, smem, s>>>(labels, decisions, n, num_classes, lut, lut_size,
                                                              (unsigned long long*)cm, (unsigned long long*)invalid);
  WLSEG_LAUNCH_CHECK();
  return 0;
}
