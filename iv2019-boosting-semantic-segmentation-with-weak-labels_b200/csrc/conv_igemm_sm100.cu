// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a.
//
// Replaces the cuDNN convolutions TF-1.12 ran for slim.conv2d / conv2d_same in
//   code/models/resnet50_extended_feature_extractor.py:25-30,39-43   (ResNet-50 OS8 + 1x1 reduce)
//   code/models/resnet50_extended_model_hierarchical.py:60-64,80     (adaptation units, logits)
// with the inference batch-norm (:298-312), ReLU and the bottleneck shortcut add fused into the
// epilogue, and (training) the per-channel sum / sum-of-squares of the raw accumulators.
//
// GEMM view:  D[m, k] = sum_{tap, c} A[m, (tap, c)] * W[k, (tap, c)]
//   m   = output pixel; an M tile is a TH x TW = 128-pixel spatial patch of one image
//   A   = NHWC bf16 activations, fetched per (tap, 64-channel chunk) by ONE 4-D TMA box
//         {64 ch, TW, TH, 1} whose start is shifted by the tap offset (r*dil - pad, s*dil - pad);
//         out-of-image rows/cols are zero-filled by TMA, which IS the convolution padding;
//         stride-2 convolutions use the tensor map's element strides {1,2,2,1}
//   W   = KRSC bf16 filters seen as a 3-D tensor {C, R*S, K}, box {64, 1, BN}
//   both land in shared memory as K-major SWIZZLE_128B tiles (rows of 128 B), exactly the
//   canonical layout tcgen05.mma reads through shared-memory descriptors
//   D   = fp32 accumulators in TMEM, 128 lanes x BN columns, double buffered (2*BN <= 512 cols)
//
// One persistent CTA per SM, warp specialised:
//   warp 0  TMA producer   (one elected lane; ring of kStages {A,B} stages, full/empty mbarriers)
//   warp 1  MMA issuer     (one elected lane; 4 x tcgen05.mma K=16 per stage; tcgen05.commit
//                           releases the stage and finally publishes the accumulator)
//   warp 2  TMEM allocator
//   warps 4-7 epilogue     (tcgen05.ld 32 lanes x 32 columns -> scale/shift/residual/ReLU -> bf16;
//                           overlaps the next tile's MMAs).  bf16 outputs stream through shared
//                           memory: the residual tile arrives by TMA into a swizzled staging
//                           buffer, results are written to a second swizzled buffer and leave by
//                           TMA store (cp.async.bulk.tensor ... bulk_group), 64 channels at a time,
//                           double buffered - the epilogue threads never touch global memory.
//                           fp32 outputs (the logits layers) use plain per-thread stores.
//
// Roofline: tensor-bound for the block3/block4 3x3 and wide 1x1 layers
// (flops = 2*N*P*Q*R*S*C*K); HBM-bound for block1 and the 64-channel 1x1 layers.
#include <cuda.h>

#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "tc_sm100.cuh"

namespace wlseg {

int check_conv_params(const wlseg_conv_params* p);
int conv_fprop_direct(const wlseg_conv_params* p, const void* x, const void* w, void* y, const float* scale,
                      const float* shift, const void* residual, cudaStream_t s);

constexpr int kBM = 128;          // pixels per tile (UMMA M)
constexpr int kBK = 64;           // channels per pipeline stage (one 128-byte swizzle row)
constexpr int kUmmaK = 16;        // bf16 MMA K
constexpr int kEpiWarp0 = 4;      // first epilogue warp
constexpr int kABytes = kBM * kBK * 2;  // 16 KB
constexpr int kSmemBudget = 224 * 1024;

struct IgemmParams {
  CUtensorMap map_a;  // activations {C, W, H, N}
  CUtensorMap map_b;  // filters {C, R*S, K}
  CUtensorMap map_y;  // output {K, Q, P, N}                         (TMA epilogue only)
  CUtensorMap map_r;  // residual {K, res_W, res_H, N}, strided box  (TMA epilogue only)
  void* y;
  const float* scale;
  const float* shift;
  const void* res;
  double* bn_sum;
  double* bn_sqsum;
  const uint32_t* out_mask;   // optional bit mask [N*P*Q][K / 32] applied to the output (wlseg_conv2d_fprop_masked)
  // kBnb (wlseg_conv2d_fprop_bnbwd): the output is the gradient of a tensor a = relu(bn(z)); `res` / map_r describe z,
  // the epilogue multiplies by the ReLU derivative (z * bnb_scale + bnb_shift > 0) and accumulates the BN backward sums
  // sum g * (z - mean) * invstd -> bn_sum (dgamma), sum g -> bn_sqsum (dbeta)
  const float* bnb_scale;
  const float* bnb_shift;
  const float* bnb_mean;
  const float* bnb_invstd;
  // wlseg_conv2d_fprop_bn: the LAST CTA to commit its statistics finalises the layer's batch norm (fin.counter != NULL)
  wlseg_bn_finalize_args fin;
  // kHalo: ONE activation box per tile {64 ch, 16 px, TH + R - 1 rows} (map_ah) and the whole filter bank resident
  CUtensorMap map_ah;
  int halo_rows, halo_bytes, halo_sbo;
  int epi_split;      // BN = 64 (one sub-tile per tile): the two epilogue column groups take alternate TILES
  int N, P, Q, K, C;
  int R, S, stride, dilation, pad_top, pad_left;
  int y_pitch, res_pitch, res_stride, res_H, res_W;
  int relu;
  int tw_log2;        // tile is TH x TW pixels with TW = 1 << tw_log2, TH = 128 / TW
  int tiles_w, tiles_h, n_tiles, total_tiles;
  int cchunks;        // ceil(C / 64)
  int num_kb;         // R * S * cchunks
  int stages;         // depth of the {A,B} operand ring
  int epi_bufs;       // 4 KB per-warp epilogue staging buffers (0: direct epilogue, else EW or 2 * EW)
  int res_mid;        // residual layers: issue the next residual box in the MIDDLE of a step (see the epilogue)
  int epi_db;         // layers without a residual: two staging buffers per warp (else one)
  int reverse;        // walk the M tiles from the last to the first (wlseg_conv_params::reverse)
  FastDiv fd_n_tiles, fd_tiles_w, fd_tiles_h;   // tile decode without runtime divisions (common.cuh)
  int m_tiles;        // N * tiles_h * tiles_w
  int units;          // work units of the persistent loop: tiles, or (CTA pair) pairs of M tiles x N tiles
};

constexpr int kSubW = 64;                       // epilogue sub-tile: 64 channels = one 128-byte row
constexpr int kWarpBufBytes = 32 * kSubW * 2;   // 4 KB: one epilogue warp's 32 rows x 64 channels
constexpr int kMaxStages = 8;
constexpr int kMaxEpiWarps = 16;
constexpr int kBarBytes = 512;                  // full[8] empty[8] tfull[2] tempty[2] rfull[32] + TMEM slot
constexpr int kSmemMax = 227 * 1024;            // opt-in dynamic shared memory per CTA on sm_100

template <int BN, bool kPair = false>
struct IgemmCfg {
  static constexpr int kBRows = kPair ? BN / 2 : BN;      // filter rows this CTA stages
  static constexpr int kBBytes = kBRows * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
};

// 32 x 32 transpose-reduce: on entry every lane holds 32 column values of ITS row; on exit
// v[0] of lane L is the sum over the warp's 32 rows of column L.  31 shuffles instead of 160.
__device__ __forceinline__ void warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

// Sequence of 64-channel output sub-tiles a CTA walks through (persistent tile loop x sub-tiles).
struct SubTileCursor {
  int tile, st;     // current tile id and sub-tile index inside it
  int n, p0, q0, k0;  // decoded tile origin
};

// kPair: a unit is TWO consecutive M tiles on one N tile, CTA `rank` of the pair takes M tile 2 * (unit / n_tiles) + rank.
// An odd M-tile count leaves the last pair's second CTA a phantom tile: image index N, so that TMA zero-fills every
// load and clips every store.
template <int BN, bool kPair>
__device__ __forceinline__ void decode_tile(const IgemmParams& prm, int tile, int rank, int& n, int& p0, int& q0, int& k0) {
  uint32_t umt, unt;
  prm.fd_n_tiles.divmod((uint32_t)tile, umt, unt);
  int mt = (int)umt;
  if (kPair) mt = 2 * mt + rank;
  k0 = (int)unt * BN;
  if (kPair && mt >= prm.m_tiles) {
    n = prm.N; p0 = 0; q0 = 0;
    return;
  }
  // reverse: same N tile (the fused statistics keep one N tile per CTA), M tiles from the end of the tensor
  if (prm.reverse) mt = prm.m_tiles - 1 - mt;
  uint32_t rest, twi, un, thi;
  prm.fd_tiles_w.divmod((uint32_t)mt, rest, twi);
  prm.fd_tiles_h.divmod(rest, un, thi);
  n = (int)un;
  q0 = (int)twi << prm.tw_log2;
  p0 = (int)thi * (kBM >> prm.tw_log2);
}

template <int BN>
__device__ __forceinline__ int num_subtiles(const IgemmParams& prm, int k0) {
  const int left = prm.K - k0;
  const int full = BN / kSubW;
  const int need = (left + kSubW - 1) / kSubW;
  return need < full ? need : full;
}

// ----------------------------------------------------------------------------- kernel
// EW epilogue warps (8 or 16): warp e serves TMEM lane quarter e % 4 (the hardware restriction:
// a warp reads the lanes 32*(warpid % 4) ..) and column group e / 4 of every 64-channel sub-tile.
// Two to four epilogue warps per scheduler hide each other's ALU / shared-memory latencies - with
// one warp per scheduler the epilogue, not HBM, bounded every bandwidth-bound layer (profiles/).
//
// kPair: the CTA-pair form (cluster of 2, tcgen05 cta_group::2).  One MMA of M = 256 spans two M tiles, each CTA stages
// its own activation tile and only HALF of the filter tile (rows [rank * BN/2, +BN/2)): the per-SM L2 -> SM ingress of
// the bandwidth-bound 1x1 / conv3 + residual layers - two thirds of which were filters (profiles/r1_eval_igemm256_full_
// summary.txt: 1.07 GB through the fabric against 0.56 GB of DRAM traffic) - drops by a third, and a stage shrinks from
// 48 to 32 KB.  The leader CTA (rank 0) issues the MMAs; both CTAs run their own producer and their own epilogue.
//
// kBnb (staged epilogue only): the BN-backward form of a data gradient, see IgemmParams::bnb_* - three staging buffers
// per warp: the z tile of the NEXT step arrives by TMA in one of two while this step reads the other, the masked
// gradient leaves from the third; the sums are read back column-wise from the staged z and g tiles.
//
// kHalo (C = 64, K <= 64, stride 1, dilation 1, S in {1, 3}; single CTA): the narrow 3x3 layers of block1 and the packed
// root convolution were bound by the L2 -> SM fabric, not by HBM or the tensor pipe - every tap re-fetched the same
// activations (9 x 16 KB per 128-pixel tile) and the 8 KB filter slice of the tap (profiles/r1_layers_eval.txt: 0.27-0.49
// of their byte bound).  Here the activation patch arrives ONCE per tile with its halo - a 4-D box {64 ch, 16 px, TH + R - 1
// rows} whose pixel rows are 2048 bytes apart in shared memory - and a tap is a shared-memory DESCRIPTOR shifted by
// r * 2048 + s * 128 bytes (8-row groups `halo_sbo` apart; the swizzle follows the absolute address bits);
// the R * S filter slices (8 KB each) are loaded once per CTA and stay.  Two patch buffers: the TMA of tile i + 1 runs
// under the MMAs of tile i.
template <int BN, typename TY, bool kTmaEpi, int EW, bool kPair, bool kBnb = false, bool kHalo = false>
__global__ void __launch_bounds__(128 + 32 * EW, 1)
conv_igemm_kernel(const __grid_constant__ IgemmParams prm) {
  using Cfg = IgemmCfg<BN, kPair>;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const int unit0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int CG = EW / 4;             // warps per TMEM lane quarter: each takes every CG-th sub-tile
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_launch_dependents();   // the next kernel may stage its CTAs; it waits for this grid before touching memory
  // SWIZZLE_128B operands need 1024-byte aligned stages (the kernel has no static shared memory,
  // so the dynamic window starts at offset 0 of the CTA's shared space)
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int stages = prm.stages;
  // kHalo: [2 patch buffers][num_kb filter slices][epilogue buffers]
  uint8_t* halo_b = smem + 2 * prm.halo_bytes;
  uint8_t* epi_smem = kHalo ? halo_b + prm.num_kb * Cfg::kBBytes : smem + stages * Cfg::kStageBytes;   // [buf0 .. buf(epi_bufs-1)]
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + prm.epi_bufs * kWarpBufBytes);
  uint64_t* full_bar = bars;                             // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;               // [kMaxStages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;           // [2]
  uint64_t* tempty_bar = bars + 2 * kMaxStages + 2;      // [2]
  uint64_t* rfull_bar = bars + 2 * kMaxStages + 4;       // [2 * EW] residual box landed (per warp, per buffer)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4 + 2 * kMaxEpiWarps);
  uint64_t* afull_bar = bars + 2 * kMaxStages + 4 + 2 * kMaxEpiWarps + 1;   // [2] kHalo: patch landed
  uint64_t* aempty_bar = afull_bar + 2;                                     // [2] kHalo: patch consumed
  uint64_t* bfull_bar = afull_bar + 4;                                      // kHalo: filter bank landed
  float* s_stat = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kBarBytes);  // [2][BN] (kTmaEpi)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(kHalo ? &prm.map_ah : &prm.map_a);
    tma_prefetch_desc(&prm.map_b);
    if (kTmaEpi) {
      tma_prefetch_desc(&prm.map_y);
      if (prm.res != nullptr) tma_prefetch_desc(&prm.map_r);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(full_bar + s), 1);
      mbar_init(smem_u32(empty_bar + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(tfull_bar + a), 1);
      mbar_init(smem_u32(tempty_bar + a), kPair ? 2 * EW : EW);  // one arrival per epilogue warp (of both CTAs)
    }
    for (int a = 0; a < 2 * EW; ++a) mbar_init(smem_u32(rfull_bar + a), 1);
    if (kHalo) {
      for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(afull_bar + a), 1); mbar_init(smem_u32(aempty_bar + a), 1); }
      mbar_init(smem_u32(bfull_bar), 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (kPair) tmem_alloc_pair(smem_u32(tmem_slot), Cfg::kTmemCols);
    else tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
  }
  if (kTmaEpi && warp == 3 && prm.bn_sum != nullptr)
    for (int j = lane; j < 2 * BN; j += 32) s_stat[j] = 0.f;
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync();   // the peer's barriers are initialised before anything of this CTA signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above touched only this CTA's shared memory / TMEM and the kernel parameters; from here on
  // the previous kernel's output is read (and buffers it may still be reading are overwritten)
  pdl_wait();

  const int TW = 1 << prm.tw_log2;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (kHalo && lane == 0) {
      // the filter bank once (one N tile: k0 = 0), then one patch per tile
      mbar_arrive_expect_tx(smem_u32(bfull_bar), prm.num_kb * Cfg::kBBytes);
      for (int kb = 0; kb < prm.num_kb; ++kb)
        tma_load_3d(smem_u32(halo_b + kb * Cfg::kBBytes), &prm.map_b, smem_u32(bfull_bar), 0, kb, 0);
      int it = 0;
      for (int tile = unit0; tile < prm.units; tile += unit_step, ++it) {
        int n, p0, q0, k0;
        decode_tile<BN, kPair>(prm, tile, rank, n, p0, q0, k0);
        const int ab = it & 1;
        mbar_wait(smem_u32(aempty_bar + ab), ((uint32_t)(it >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(smem_u32(afull_bar + ab), prm.halo_bytes);
        tma_load_4d(smem_u32(smem + ab * prm.halo_bytes), &prm.map_ah, smem_u32(afull_bar + ab), 0, q0 - prm.pad_left,
                    p0 - prm.pad_top, n);
      }
    } else if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit0; tile < prm.units; tile += unit_step) {
        int n, p0, q0, k0;
        decode_tile<BN, kPair>(prm, tile, rank, n, p0, q0, k0);
        // (tap, channel chunk) walked with counters: no division in the producer's issue loop
        int tap = 0, cc = 0, r = 0, s = 0;
        for (int kb = 0; kb < prm.num_kb; ++kb) {
          mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
          const uint32_t a_dst = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_dst = a_dst + kABytes;
          const uint32_t bar = smem_u32(full_bar + stage);
          if constexpr (kPair) {
            // the leader's barrier collects the bytes of BOTH CTAs' boxes (a peer box may complete before the leader
            // has posted its expectation: the pending arrival keeps the phase open)
            if (rank == 0) mbar_arrive_expect_tx(bar, 2 * Cfg::kStageBytes);
            const uint32_t lbar = mapa_shared(bar, 0);
            tma_load_4d_pair(a_dst, &prm.map_a, lbar, cc * kBK, q0 * prm.stride - prm.pad_left + s * prm.dilation,
                             p0 * prm.stride - prm.pad_top + r * prm.dilation, n);
            tma_load_3d_pair(b_dst, &prm.map_b, lbar, cc * kBK, tap, k0 + (int)rank * Cfg::kBRows);
          } else {
            mbar_arrive_expect_tx(bar, Cfg::kStageBytes);
            tma_load_4d(a_dst, &prm.map_a, bar, cc * kBK, q0 * prm.stride - prm.pad_left + s * prm.dilation,
                        p0 * prm.stride - prm.pad_top + r * prm.dilation, n);
            tma_load_3d(b_dst, &prm.map_b, bar, cc * kBK, tap, k0);
          }
          if (++stage == stages) { stage = 0; phase ^= 1; }
          if (++cc == prm.cchunks) {
            cc = 0; ++tap;
            if (++s == prm.S) { s = 0; ++r; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (kHalo && lane == 0) {
      constexpr uint32_t idesc = make_idesc(kBM, BN);
      mbar_wait(smem_u32(bfull_bar), 0);
      int iter = 0;
      for (int tile = unit0; tile < prm.units; tile += unit_step, ++iter) {
        const int acc = iter & 1, ab = iter & 1;
        const uint32_t ph = (iter >> 1) & 1;
        mbar_wait(smem_u32(tempty_bar + acc), ph ^ 1);   // epilogue drained this accumulator
        mbar_wait(smem_u32(afull_bar + ab), ph);          // this tile's patch has landed
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const uint32_t a_base = smem_u32(smem + ab * prm.halo_bytes);
        int r = 0, sx = 0;
        for (int kb = 0; kb < prm.num_kb; ++kb) {
          const uint64_t adesc = make_smem_desc_shifted(a_base + (uint32_t)(r * 2048 + sx * 128), (uint32_t)prm.halo_sbo);
          const uint64_t bdesc = make_smem_desc(smem_u32(halo_b + kb * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k)
            umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          if (++sx == prm.S) { sx = 0; ++r; }
        }
        umma_commit(smem_u32(aempty_bar + ab));   // the patch buffer is free once these MMAs retire
        umma_commit(smem_u32(tfull_bar + acc));
      }
    } else if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc(kPair ? 2 * kBM : kBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int tile = unit0; tile < prm.units; tile += unit_step, ++iter) {
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        mbar_wait(smem_u32(tempty_bar + acc), acc_phase ^ 1);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < prm.num_kb; ++kb) {
          mbar_wait(smem_u32(full_bar + stage), phase);          // TMA bytes have landed
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t adesc = make_smem_desc(a_addr);
          const uint64_t bdesc = make_smem_desc(a_addr + kABytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            // advance 16 elements = 32 bytes along K inside the 128-byte swizzle row
            if (kPair) umma_bf16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            else umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          // stage free once these MMAs retire (pair: in both CTAs)
          if (kPair) umma_commit_pair(smem_u32(empty_bar + stage), 3); else umma_commit(smem_u32(empty_bar + stage));
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete (pair: each CTA's epilogue waits on its own barrier)
        if (kPair) umma_commit_pair(smem_u32(tfull_bar + acc), 3); else umma_commit(smem_u32(tfull_bar + acc));
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue =====================
    const int ew = warp - kEpiWarp0;
    const int quarter = ew & 3;               // TMEM lanes [32*quarter, 32*quarter + 32)
    const int cgrp = ew >> 2;                 // column group inside a 64-channel sub-tile
    const int et = threadIdx.x - kEpiWarp0 * 32;
    const int row = quarter * 32 + lane;      // tile row = pixel within the patch
    const int dy_ = row >> prm.tw_log2, dx_ = row & (TW - 1);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    // the accumulator goes back to the MMA warp - of the leader CTA in the pair form
    auto release_acc = [&](int a) {
      if (kPair) mbar_arrive_cluster(mapa_shared(smem_u32(tempty_bar + a), 0));
      else mbar_arrive(smem_u32(tempty_bar + a));
    };
    int iter = 0;
    if constexpr (kTmaEpi) {
      // ---- staged epilogue, one INDEPENDENT pipeline per warp: a warp owns the 32 rows of its TMEM
      // lane quarter and every CG-th 64-channel sub-tile; TMEM -> registers -> its own swizzled 4 KB
      // staging buffer -> its own TMA store (box {64 ch, 32 pixels}).  The residual box is TMA-loaded
      // INTO the staging buffer it will leave from (read-modify-write in place, double buffered per
      // warp).  No CTA-wide barrier: the warps drift apart and hide each other's latencies.
      const bool has_res = prm.res != nullptr;
      const bool has_stat = prm.bn_sum != nullptr;
      const uint32_t sw = (uint32_t)(lane & 7);           // row & 7 of this thread's staging row
      const bool dbuf = has_res || prm.epi_db;            // two staging buffers per warp
      // kHalo (BN = 64: only the warps of column group 0 ever stage anything) keeps buffers for those four warps only
      // split (BN = 64): a tile is ONE 64-channel sub-tile, so the second column group would idle - and one warp per
      // TMEM lane quarter walking the tiles serially bounded the narrow layers (root convolution: 110 tiles per CTA x
      // ~1.3 us of epilogue step latency = the whole 154 us launch).  The groups take alternate tiles instead:
      // group g drains accumulator buffer g, two tiles' epilogues are in flight per quarter.
      const bool split = BN == kSubW && prm.epi_split != 0;
      uint8_t* wbuf = epi_smem + ((kHalo && !split) ? quarter : ew) * (kBnb ? 3 : (dbuf ? 2 : 1)) * kWarpBufBytes;
      const int st0 = split ? 0 : cgrp;            // this warp's first sub-tile of a tile, its sub-tile stride,
      const int st_step = split ? BN / kSubW : CG;
      const int tile0 = unit0 + (split ? cgrp * unit_step : 0);   // its first tile and its tile stride
      const int tile_step = split ? 2 * unit_step : unit_step;
      uint64_t* my_rfull = rfull_bar + 2 * ew;
      const int r0 = quarter * 32;                        // first tile row of this warp
      const int dy0 = r0 >> prm.tw_log2, dx0 = r0 & (TW - 1);
      // residual cursors (lane 0) over this warp's (tile, sub-tile) sequence: `ld` stages the box of the
      // NEXT step into the warp's other buffer, `pf` runs kResPf steps ahead and only pulls the box into
      // L2 (cp.async.bulk.prefetch), so that the staging load is an L2 hit - the 4 KB-per-warp staging
      // buffers alone keep far too few bytes in flight to cover HBM latency
      constexpr int kResPf = 4;
      struct ResCursor { int tile, st, cnt; };
      ResCursor ld = {tile0, st0, 0}, pf = {tile0, st0, 0};
      auto next_residual = [&](ResCursor& c, bool stage) {
        while (c.tile < prm.units) {
          int n, p0, q0, k0;
          decode_tile<BN, kPair>(prm, c.tile, rank, n, p0, q0, k0);
          if (c.st < num_subtiles<BN>(prm, k0)) {
            const int cx = (q0 + dx0) * prm.res_stride, cy = (p0 + dy0) * prm.res_stride;
            if (stage) {
              const uint32_t bar = smem_u32(my_rfull + (c.cnt & 1));
              mbar_arrive_expect_tx(bar, kWarpBufBytes);
              tma_load_4d(smem_u32(wbuf + (c.cnt & 1) * kWarpBufBytes), &prm.map_r, bar, k0 + c.st * kSubW, cx, cy, n);
            } else {
              tma_prefetch_4d(&prm.map_r, k0 + c.st * kSubW, cx, cy, n);
            }
            ++c.cnt;
            c.st += st_step;
            return;
          }
          c.st = st0;
          c.tile += tile_step;
        }
      };
      if (has_res && lane == 0) {
        for (int i = 0; i < kResPf; ++i) next_residual(pf, false);
        next_residual(ld, true);   // step 0 -> buffer 0
      }
      // BN statistic partials of this warp's rows, kept in registers for the whole kernel (the host
      // sizes the grid as a multiple of the N-tile count, so a CTA never changes its N tile):
      // lane = channel pair, slot = which of the warp's sub-tiles
      constexpr int kSlots = (BN / kSubW + CG - 1) / CG;
      float a1x[kSlots], a1y[kSlots], a2x[kSlots], a2y[kSlots];
#pragma unroll
      for (int i = 0; i < kSlots; ++i) a1x[i] = a1y[i] = a2x[i] = a2y[i] = 0.f;
      int stat_k0 = -1;
      int cnt = 0;
      // kBnb: saved mean / inverse std of this lane's channel pair, per slot (the CTA never changes its N tile)
      float bmx[kSlots], bmy[kSlots], bix[kSlots], biy[kSlots];
#pragma unroll
      for (int i = 0; i < kSlots; ++i) bmx[i] = bmy[i] = bix[i] = biy[i] = 0.f;
      // output mask (wlseg_conv2d_fprop_masked: a data gradient leaving through the ReLU of the tensor it belongs to):
      // 64 bits per pixel and sub-tile, fetched ONE TILE AHEAD - a load issued inside the step sat on its critical
      // path with a full HBM latency (measured: +18 us on a 40 us dgrad)
      uint2 mcur[kSlots], mnext[kSlots];
      auto load_mask = [&](int tile_, uint2 (&m)[kSlots]) {
#pragma unroll
        for (int i = 0; i < kSlots; ++i) m[i] = make_uint2(0xffffffffu, 0xffffffffu);
        if (prm.out_mask == nullptr || tile_ >= prm.units) return;
        int n_, p0_, q0_, k0_;
        decode_tile<BN, kPair>(prm, tile_, rank, n_, p0_, q0_, k0_);
        const bool ok = (p0_ + dy_ < prm.P) && (q0_ + dx_ < prm.Q) && (n_ < prm.N);
        const int nsub_ = num_subtiles<BN>(prm, k0_);
        const uint2* row = reinterpret_cast<const uint2*>(
            prm.out_mask + (((int64_t)n_ * prm.P + p0_ + dy_) * prm.Q + q0_ + dx_) * (prm.K >> 5) + (k0_ >> 5));
#pragma unroll
        for (int i = 0; i < kSlots; ++i) {
          const int st_ = st0 + i * st_step;
          m[i] = (ok && st_ < nsub_) ? __ldg(row + st_) : make_uint2(0u, 0u);
        }
      };
      load_mask(unit0, mnext);
      for (int tile = unit0; tile < prm.units; tile += unit_step, ++iter) {
        int n, p0, q0, k0;
        decode_tile<BN, kPair>(prm, tile, rank, n, p0, q0, k0);
        const bool valid = (p0 + dy_ < prm.P) && (q0 + dx_ < prm.Q) && (n < prm.N);
#pragma unroll
        for (int i = 0; i < kSlots; ++i) mcur[i] = mnext[i];
        load_mask(tile + unit_step, mnext);
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        const int nsub = num_subtiles<BN>(prm, k0);
        if (has_stat) {
          if (stat_k0 >= 0 && k0 != stat_k0) __trap();  // grid not a multiple of the N-tile count
          if constexpr (kBnb) {
            if (stat_k0 < 0) {
#pragma unroll
              for (int i = 0; i < kSlots; ++i) {
                const int c = k0 + (st0 + i * st_step) * kSubW + 2 * lane;
                if (c + 1 < prm.K) {
                  bmx[i] = __ldg(prm.bnb_mean + c); bmy[i] = __ldg(prm.bnb_mean + c + 1);
                  bix[i] = __ldg(prm.bnb_invstd + c); biy[i] = __ldg(prm.bnb_invstd + c + 1);
                }
              }
            }
          }
          stat_k0 = k0;
        }
        mbar_wait(smem_u32(tfull_bar + acc), acc_phase);
        tc_fence_after();
        const bool idle = split ? ((iter & 1) != cgrp) : (cgrp >= nsub);
        if (idle) {
          // nothing to do in this tile: still one arrival per warp and tile
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc(acc);
        }
#pragma unroll
        for (int slot = 0; slot < kSlots; ++slot) {
          const int st = st0 + slot * st_step;
          if (st >= nsub || idle) break;
          if constexpr (kBnb) {
            uint8_t* zb = wbuf + (cnt & 1) * kWarpBufBytes;   // z tile of this step (TMA)
            uint8_t* ob = wbuf + 2 * kWarpBufBytes;           // masked gradient, leaves by TMA store
            if (lane == 0) {
              bulk_wait_read<0>();        // the previous step's store has read `ob`
              next_residual(ld, true);    // z of step cnt + 1 -> the other z buffer (its readers finished in step cnt - 1)
              next_residual(pf, false);
            }
            __syncwarp();
            mbar_wait(smem_u32(my_rfull + (cnt & 1)), (uint32_t)((cnt >> 1) & 1));
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int col = st * kSubW + half * 32;
              uint32_t v[32];
              tmem_ld<32>(lane_addr + (uint32_t)(acc * BN + col), v);
              tmem_ld_wait();
              if (half == 1 && st + CG >= nsub) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) release_acc(acc);
              }
              float f[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = valid ? __uint_as_float(v[j]) : 0.f;
              uint4 raw[4];
#pragma unroll
              for (int g = 0; g < 4; ++g)
                raw[g] = *reinterpret_cast<const uint4*>(zb + lane * 128 + ((((uint32_t)(half * 4 + g)) ^ sw) << 4));
              const float4* sc4 = reinterpret_cast<const float4*>(prm.bnb_scale + k0 + col);
              const float4* sh4 = reinterpret_cast<const float4*>(prm.bnb_shift + k0 + col);
              uint4 outv[4];
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const __nv_bfloat162* hz = reinterpret_cast<const __nv_bfloat162*>(&raw[g]);
                __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&outv[g]);
                const float4 sca = __ldg(sc4 + 2 * g), scb = __ldg(sc4 + 2 * g + 1);
                const float4 sha = __ldg(sh4 + 2 * g), shb = __ldg(sh4 + 2 * g + 1);
                const float scs[8] = {sca.x, sca.y, sca.z, sca.w, scb.x, scb.y, scb.z, scb.w};
                const float shs[8] = {sha.x, sha.y, sha.z, sha.w, shb.x, shb.y, shb.z, shb.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 zz = __bfloat1622float2(hz[e]);
                  // the ReLU derivative from the sign of the exact fp32 value the forward pass rounded (bn_reduce_kernel)
                  const float gx = fmaf(zz.x, scs[2 * e], shs[2 * e]) > 0.f ? f[g * 8 + 2 * e] : 0.f;
                  const float gy = fmaf(zz.y, scs[2 * e + 1], shs[2 * e + 1]) > 0.f ? f[g * 8 + 2 * e + 1] : 0.f;
                  ho[e] = __floats2bfloat162_rn(gx, gy);
                }
              }
#pragma unroll
              for (int g = 0; g < 4; ++g)
                *reinterpret_cast<uint4*>(ob + lane * 128 + ((((uint32_t)(half * 4 + g)) ^ sw) << 4)) = outv[g];
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&prm.map_y, smem_u32(ob), k0 + st * kSubW, q0 + dx0, p0 + dy0, n);
              bulk_commit();
            }
            {
              // dgamma / dbeta partials of the STORED (bf16) gradient: lane = channel pair, column-wise over the 32 rows
              const uint32_t coff = (uint32_t)((lane & 3) << 2);
              float s0x = 0.f, s0y = 0.f, s1x = 0.f, s1y = 0.f;
              const float mx = bmx[slot], my = bmy[slot], ix = bix[slot], iy = biy[slot];
#pragma unroll
              for (int r = 0; r < 32; ++r) {
                const uint32_t o = (uint32_t)(r * 128) + ((((uint32_t)(lane >> 2)) ^ (uint32_t)(r & 7)) << 4) + coff;
                const float2 g2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ob + o));
                const float2 z2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(zb + o));
                s0x += g2.x * (z2.x - mx) * ix; s0y += g2.y * (z2.y - my) * iy;
                s1x += g2.x; s1y += g2.y;
              }
              a1x[slot] += s0x; a1y[slot] += s0y; a2x[slot] += s1x; a2y[slot] += s1y;
            }
            __syncwarp();   // every lane is done with `zb` before lane 0 lets the next z box land there (next step)
            ++cnt;
            continue;
          }
          uint8_t* buf = wbuf + (dbuf ? (cnt & 1) * kWarpBufBytes : 0);
          // the store that last left from the buffer about to be (re)written must have read it:
          // without a residual that is this step's buffer, with one it is the NEXT step's
          // res_mid: the wait for the previous step's store (it must have READ the other buffer before the next
          // residual box may land there) and the issue of that box move from the start of the step to its middle,
          // and the wait for THIS step's residual moves behind the first TMEM load + scale/shift: both latencies
          // (store read ~0.5 us, L2-hit box load ~0.7 us) then run under this step's own arithmetic instead of
          // in front of it.  Measured: 512 -> 2048 + residual 369 -> 347 us, 256 -> 1024 + residual 155 -> 152 us,
          // the narrower ones unchanged (WLSEG_RES_MID=0 restores the old schedule)
          const bool mid = has_res && prm.res_mid;
          if (lane == 0 && !mid) {
            // without a residual and with two buffers only the store from TWO steps back (this buffer's last
            // tenant) must have been read: the previous step's store stays in flight under this step
            if (!has_res && dbuf) bulk_wait_read<1>(); else bulk_wait_read<0>();
            if (has_res) {
              next_residual(ld, true);    // step cnt + 1 -> the other buffer
              next_residual(pf, false);   // step cnt + 1 + kResPf -> L2
            }
          }
          __syncwarp();
          if (has_res && !mid) mbar_wait(smem_u32(my_rfull + (cnt & 1)), (uint32_t)((cnt >> 1) & 1));
          uint8_t* myrow = buf + lane * 128;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int col = st * kSubW + half * 32;   // column inside the BN-wide tile
            if (mid && half == 1) {
              if (lane == 0) {
                bulk_wait_read<0>();
                next_residual(ld, true);    // step cnt + 1 -> the other buffer
                next_residual(pf, false);   // step cnt + 1 + kResPf -> L2
              }
              __syncwarp();
            }
            const uint32_t mword = half == 0 ? mcur[slot].x : mcur[slot].y;
            uint32_t v[32];
            tmem_ld<32>(lane_addr + (uint32_t)(acc * BN + col), v);
            tmem_ld_wait();
            if (half == 1 && st + CG >= nsub) {
              // this warp's last TMEM read of the tile: hand the accumulator back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) release_acc(acc);
            }
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (prm.scale != nullptr) {
              // folded batch norm; warp-uniform addresses (broadcast, L1 resident)
              const float4* sc4 = reinterpret_cast<const float4*>(prm.scale + k0 + col);
              const float4* sh4 = reinterpret_cast<const float4*>(prm.shift + k0 + col);
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 sc = __ldg(sc4 + g);
                const float4 sh = __ldg(sh4 + g);
                f[g * 4 + 0] = fmaf(f[g * 4 + 0], sc.x, sh.x);
                f[g * 4 + 1] = fmaf(f[g * 4 + 1], sc.y, sh.y);
                f[g * 4 + 2] = fmaf(f[g * 4 + 2], sc.z, sh.z);
                f[g * 4 + 3] = fmaf(f[g * 4 + 3], sc.w, sh.w);
              }
            }
            uint4* slot4[4];
#pragma unroll
            for (int g = 0; g < 4; ++g)
              slot4[g] = reinterpret_cast<uint4*>(myrow + ((((uint32_t)(half * 4 + g)) ^ sw) << 4));
            if (mid && half == 0) mbar_wait(smem_u32(my_rfull + (cnt & 1)), (uint32_t)((cnt >> 1) & 1));
            if (has_res) {
              // all 16-byte chunks are loaded before any is rewritten (no false aliasing stalls)
              uint4 raw[4];
#pragma unroll
              for (int g = 0; g < 4; ++g) raw[g] = *slot4[g];
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[g]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 t = __bfloat1622float2(h[e]);
                  f[g * 8 + 2 * e] += t.x;
                  f[g * 8 + 2 * e + 1] += t.y;
                }
              }
            }
            if (prm.out_mask != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = ((mword >> j) & 1u) ? f[j] : 0.f;
            }
            if (has_stat && !valid) {
              // pixels outside the image must not reach the statistics (TMA clips them from the store)
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = 0.f;
            }
            uint4 outv[4];
            if (prm.relu) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&outv[g]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  // cvt.rn.relu: clamp and round in ONE instruction (rounding is monotone and keeps zero: the same bits
                  // as fmaxf followed by the conversion, 64 FMNMX per step less)
                  uint32_t packed;
                  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(packed) : "f"(f[g * 8 + 2 * e + 1]), "f"(f[g * 8 + 2 * e]));
                  reinterpret_cast<uint32_t*>(ho)[e] = packed;
                }
              }
            } else {
              // training forward (raw z) and every data gradient: no clamp - 64 FMNMX of a ~650-instruction step that
              // ncu shows issue-bound on the shallow layers (profiles/r2_epilogue_issue_bound.md)
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&outv[g]);
#pragma unroll
                for (int e = 0; e < 4; ++e) ho[e] = __floats2bfloat162_rn(f[g * 8 + 2 * e], f[g * 8 + 2 * e + 1]);
              }
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) *slot4[g] = outv[g];
          }
          fence_async_smem();   // my generic-proxy writes -> visible to the TMA store
          __syncwarp();         // the warp's 32 rows are written
          if (lane == 0) {
            tma_store_4d(&prm.map_y, smem_u32(buf), k0 + st * kSubW, q0 + dx0, p0 + dy0, n);
            bulk_commit();
          }
          if (has_stat) {
            // training-mode BN statistics of the STORED (bf16) activations, read back column-wise from
            // the staging buffer: lane = channel pair; conflict-free under the swizzle
            const uint8_t* base = buf + ((lane & 3) << 2);
            float s1x = 0.f, s1y = 0.f, s2x = 0.f, s2y = 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(
                  base + r * 128 + ((((uint32_t)(lane >> 2)) ^ (uint32_t)(r & 7)) << 4));
              const float2 t = __bfloat1622float2(h);
              s1x += t.x; s1y += t.y;
              s2x = fmaf(t.x, t.x, s2x); s2y = fmaf(t.y, t.y, s2y);
            }
            a1x[slot] += s1x; a1y[slot] += s1y; a2x[slot] += s2x; a2y[slot] += s2y;
          }
          ++cnt;
        }
      }
      if (has_stat) {
        // combine the warps' partials in a FIXED order, then ONE fp64 atomic per channel and CTA (same-address L2
        // atomics serialise: per-warp flushes cost more than the layer's math).  Every warp parks its partial sums in
        // its own staging buffer - float atomics on shared memory commit in a varying order, and a 1e-7 wobble of a
        // batch-norm sum is amplified ~1e5 by a train-mode ResNet at random init (run-to-run 1.8e-2 on the gradients)
        constexpr int kBufsPerWarp = kBnb ? 3 : 0;   // 0: (dbuf ? 2 : 1), resolved below
        const int nbuf = kBufsPerWarp ? kBufsPerWarp : (dbuf ? 2 : 1);
        if (lane == 0) bulk_wait_read<0>();          // the last stores have read the staging buffers
        __syncwarp();
        float* park = reinterpret_cast<float*>(wbuf);
#pragma unroll
        for (int i = 0; i < kSlots; ++i) {
          if (kHalo && !split && cgrp != 0) break;   // no buffer of its own, and nothing to add
          park[(i * 2 + 0) * kSubW + 2 * lane] = a1x[i]; park[(i * 2 + 0) * kSubW + 2 * lane + 1] = a1y[i];
          park[(i * 2 + 1) * kSubW + 2 * lane] = a2x[i]; park[(i * 2 + 1) * kSubW + 2 * lane + 1] = a2y[i];
        }
        epi_barrier<32 * EW>();
        if (stat_k0 >= 0) {
          for (int j = et; j < BN; j += 32 * EW) {
            if (stat_k0 + j < prm.K) {
              const int st = j / kSubW, ch = j % kSubW;
              const int cg = split ? 0 : st % CG, i = split ? 0 : st / CG;
              float t1 = 0.f, t2 = 0.f;
              for (int g2 = 0; g2 < (split ? CG : 1); ++g2) {     // split: both column groups hold partials of sub-tile 0
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                  const float* src = reinterpret_cast<const float*>(epi_smem + (q4 + 4 * (cg + g2)) * nbuf * kWarpBufBytes);
                  t1 += src[(i * 2 + 0) * kSubW + ch];
                  t2 += src[(i * 2 + 1) * kSubW + ch];
                }
              }
              atomicAdd(prm.bn_sum + stat_k0 + j, (double)t1);
              atomicAdd(prm.bn_sqsum + stat_k0 + j, (double)t2);
            }
          }
        }
      }
      if constexpr (!kBnb) {
        if (has_stat && prm.fin.counter != nullptr) {
          // bn_finalize_kernel's arithmetic, run by the last CTA to arrive: a ~3 us launch per layer less.  Every CTA
          // of the grid (also one without tiles) passes here exactly once; the counter is left at zero for the next
          // layer (kernels of one stream do not overlap: the fused form is not combined with programmatic launches).
          uint32_t* s_last = tmem_slot + 1;
          __threadfence();                      // this CTA's fp64 atomics before its ticket
          epi_barrier<32 * EW>();
          if (et == 0) *s_last = (atomicAdd(prm.fin.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
          epi_barrier<32 * EW>();
          if (*s_last != 0u) {
            __threadfence();
            const double nn = (double)prm.fin.count;
            for (int c = et; c < prm.K; c += 32 * EW) {
              const double m = __ldcg(prm.bn_sum + c) / nn;
              double var = __ldcg(prm.bn_sqsum + c) / nn - m * m;
              if (var < 0.0) var = 0.0;
              const float mf = (float)m, vf = (float)var;
              const float inv = rsqrtf(vf + prm.fin.eps);
              const float sc = prm.fin.gamma[c] * inv;
              prm.fin.scale[c] = sc;
              prm.fin.shift[c] = prm.fin.beta[c] - mf * sc;
              prm.fin.saved_mean[c] = mf;
              prm.fin.saved_invstd[c] = inv;
              if (prm.fin.moving_mean != nullptr) {
                const float unbiased = prm.fin.moving_var_factor >= 0.f
                                           ? vf * prm.fin.moving_var_factor
                                           : (prm.fin.count > 1 ? (float)(var * (nn / (nn - 1.0))) : vf);
                prm.fin.moving_mean[c] -= (1.0f - prm.fin.decay) * (prm.fin.moving_mean[c] - mf);
                prm.fin.moving_var[c] -= (1.0f - prm.fin.decay) * (prm.fin.moving_var[c] - unbiased);
              }
            }
            if (et == 0) *prm.fin.counter = 0u;
          }
        }
      }
      if (lane == 0) bulk_wait_all();  // outstanding stores complete before the CTA retires
    } else {
      // ---- direct epilogue: per-thread global stores (fp32 logits, odd channel counts); 32-column
      // chunks are dealt round-robin to the column groups
      for (int tile = unit0; tile < prm.units; tile += unit_step, ++iter) {
        int n, p0, q0, k0;
        decode_tile<BN, kPair>(prm, tile, rank, n, p0, q0, k0);
        const int q = q0 + dx_;
        const int p = p0 + dy_;
        const bool valid = (p < prm.P) && (q < prm.Q) && (n < prm.N);
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        mbar_wait(smem_u32(tfull_bar + acc), acc_phase);
        tc_fence_after();
        const int64_t opix = ((int64_t)n * prm.P + p) * prm.Q + q;
        TY* yrow = reinterpret_cast<TY*>(prm.y) + opix * prm.y_pitch;
        const __nv_bfloat16* rrow = nullptr;
        if (prm.res != nullptr && valid)
          rrow = reinterpret_cast<const __nv_bfloat16*>(prm.res) +
                 (((int64_t)n * prm.res_H + (int64_t)p * prm.res_stride) * prm.res_W + (int64_t)q * prm.res_stride) *
                     prm.res_pitch;
#pragma unroll 1
        for (int ch = cgrp; ch < BN / 32; ch += CG) {
          const int kbase = k0 + ch * 32;
          if (kbase >= prm.K) break;  // warp-uniform
          uint32_t v[32];
          tmem_ld<32>(lane_addr + (uint32_t)(acc * BN + ch * 32), v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);

          if (prm.bn_sum != nullptr) {
            float s1[32], s2[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              s1[j] = valid ? f[j] : 0.f;
              s2[j] = s1[j] * s1[j];
            }
            warp_column_sums(s1, lane);
            warp_column_sums(s2, lane);
            if (kbase + lane < prm.K) {
              atomicAdd(prm.bn_sum + kbase + lane, (double)s1[0]);
              atomicAdd(prm.bn_sqsum + kbase + lane, (double)s2[0]);
            }
          }
          if (valid) {
            const bool full = (kbase + 32 <= prm.K);
            if (prm.scale != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (full || kbase + j < prm.K) f[j] = f[j] * __ldg(prm.scale + kbase + j) + __ldg(prm.shift + kbase + j);
            }
            if (rrow != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (full || kbase + j < prm.K) f[j] += __bfloat162float(rrow[kbase + j]);
            }
            if (prm.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (sizeof(TY) == 2) {
              __nv_bfloat16* yo = reinterpret_cast<__nv_bfloat16*>(yrow) + kbase;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (full || kbase + j < prm.K) yo[j] = __float2bfloat16_rn(f[j]);
            } else {
              float* yo = reinterpret_cast<float*>(yrow) + kbase;
              if (full && (prm.y_pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(yo) & 15) == 0)) {
#pragma unroll
                for (int g = 0; g < 8; ++g)
                  reinterpret_cast<float4*>(yo)[g] = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (kbase + j < prm.K) yo[j] = f[j];
              }
            }
          }
        }
        // this warp no longer reads the accumulator: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(acc);
      }
    }
  }

  // teardown: everyone done with TMEM before it is freed
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_relaxed();   // the peer no longer reads this CTA's shared memory / signals its barriers
  if (warp == 2) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ----------------------------------------------------------------------------- host side
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    // libcuda is not linked at build time (the build box has no driver): resolve at run time
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// the only shared mutable state of the library: a mutex-guarded tensor-map cache
typedef std::vector<uint64_t> MapKeyV;
static std::mutex g_map_mutex;
static std::map<MapKeyV, CUtensorMap> g_map_cache;

// Encodes (or fetches from the cache) a SWIZZLE_128B tiled tensor map.  elem_bytes 2 = bf16, 4 = fp32.
int encode_tensor_map(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estr, int kind) {
  MapKeyV key;
  key.reserve(4 + 4 * rank);
  key.push_back((uint64_t)(uintptr_t)base);
  key.push_back((uint64_t)elem_bytes);
  key.push_back((uint64_t)rank);
  key.push_back((uint64_t)kind);
  for (int i = 0; i < rank; ++i) key.push_back(dims[i]);
  for (int i = 0; i + 1 < rank; ++i) key.push_back(strides_bytes[i]);
  for (int i = 0; i < rank; ++i) key.push_back(box[i]);
  for (int i = 0; i < rank; ++i) key.push_back(estr[i]);
  std::lock_guard<std::mutex> lock(g_map_mutex);
  auto it = g_map_cache.find(key);
  if (it != g_map_cache.end()) {
    *out = it->second;
    return 0;
  }
  EncodeTiledFn fn = get_encode_fn();
  WLSEG_CHECK_ARG(fn != nullptr, "tensor map: cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t d[5], st[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = estr[i]; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                  (cuuint32_t)rank, const_cast<void*>(base), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  WLSEG_CHECK_ARG(r == CUDA_SUCCESS, "tensor map: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  if (g_map_cache.size() > 8192) g_map_cache.clear();
  g_map_cache[key] = *out;
  return 0;
}

static int pick_bn(int K) { return K > 128 ? 256 : (K > 64 ? 128 : (K > 32 ? 64 : 32)); }

static bool igemm_supported(const wlseg_conv_params* p) {
  if (p->dtype != WLSEG_BF16) return false;
  if (p->stride != 1 && p->stride != 2) return false;
  if (p->C % 8 != 0 || p->x_pitch % 8 != 0) return false;
  if (p->R * p->S > 64 || p->dilation > 64) return false;
  if (p->N > 65535) return false;
  return true;
}

// CTA-pair launches: cluster of 2 along x; the number of co-resident clusters is asked from the driver once per kernel
// (a GPC with an odd number of free SMs cannot host a pair on its last SM)
template <typename K>
static int max_active_pairs(K kernel, int threads, int smem_bytes) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * kNumSMs);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_pair(void (*kernel)(KArgs...), int clusters, int threads, size_t smem, cudaStream_t stream,
                               Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// shared memory of the halo form: two patch buffers, the resident filter bank, staging buffers of the four working
// epilogue warps (BN = 64), barriers + statistics
static int halo_smem_bytes(const IgemmParams& prm, int bn, int bufs_per_warp, int warps = 4) {
  return 2 * prm.halo_bytes + prm.num_kb * bn * kBK * 2 + warps * bufs_per_warp * kWarpBufBytes + kBarBytes + 2 * bn * 4;
}
// BN = 64: the two epilogue column groups take alternate tiles (conv_igemm_kernel, `split`); WLSEG_EPI_SPLIT=0 disables
static bool epi_split_enabled() { return env_int("WLSEG_EPI_SPLIT", 1) != 0; }

template <int BN, typename TY, bool kTmaEpi, int EW, bool kPair = false, bool kBnb = false, bool kHalo = false>
static int launch_igemm_ew(IgemmParams& prm, cudaStream_t s) {
  using Cfg = IgemmCfg<BN, kPair>;
  static_assert(!kBnb || kTmaEpi, "the BN-backward form lives in the staged epilogue");
  static_assert(!kHalo || (kTmaEpi && !kPair && BN == 64), "the halo form: staged epilogue, single CTA, BN = 64");
  static bool configured = false;
  if (!configured) {
    WLSEG_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BN, TY, kTmaEpi, EW, kPair, kBnb, kHalo>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    configured = true;
  }
  if constexpr (kHalo) {
    prm.epi_db = 0;
    prm.res_mid = getenv("WLSEG_RES_MID") != nullptr ? atoi(getenv("WLSEG_RES_MID")) : 1;
    const int bufs = kBnb ? 3 : (prm.res != nullptr ? 2 : 1);
    // staging buffers for all eight epilogue warps when they fit (then the column groups alternate tiles), else for
    // the four warps of group 0
    prm.epi_split = (epi_split_enabled() && halo_smem_bytes(prm, BN, bufs, 8) <= kSmemMax) ? 1 : 0;
    const int ewarps = prm.epi_split ? 8 : 4;
    prm.epi_bufs = ewarps * bufs;
    prm.stages = 0;
    const int smem_bytes = halo_smem_bytes(prm, BN, bufs, ewarps);
    WLSEG_CHECK_ARG(smem_bytes <= kSmemMax, "conv(tcgen05, halo): shared memory plan does not fit");
    prm.units = prm.total_tiles;
    int grid = prm.total_tiles < conv_sms() ? prm.total_tiles : conv_sms();
    WLSEG_CUDA(launch_pdl(conv_igemm_kernel<BN, TY, kTmaEpi, EW, false, kBnb, true>, dim3(grid), dim3(128 + 32 * EW), smem_bytes,
                          s, prm));
    return 0;
  }
  // shared memory plan: [stages x {A,B}] [epi_bufs x 4 KB] [barriers]
  // layers without a residual: a second staging buffer per warp for the 1x1 layers, so that the previous step's
  // store stays in flight under the current one.  OFF by default - measured on B200 (4 x 128 x 256 eval shapes): it
  // costs the fourth operand stage, which the wide-C layers need more (1024 -> 2048: 439 -> 469 us, 2048 -> 512:
  // 216 -> 250 us) than the narrow ones gain (256 -> 768: 83 -> 78 us); eval step 8.75 -> 8.95 ms.  The exposed
  // store-read wait is therefore NOT what holds the bandwidth-bound layers at 0.5-0.8 of their byte bound.
  // Round 2: ON for the SHALLOW 1x1 layers (<= 8 k-blocks per tile: the expansions 64 -> 256, 128 -> 512, 256 -> 1024, 256 -> 768):
  // their tiles finish their MMAs long before the epilogue has drained the previous one, the operand ring is never
  // deeper than a tile, and with one buffer every step waits for the previous store to have READ it.
  // WLSEG_EPI_DB = 0 / 1 forces it off / on for every 1x1 layer without a residual.
  const int db_env = getenv("WLSEG_EPI_DB") != nullptr ? atoi(getenv("WLSEG_EPI_DB")) : -1;
  prm.epi_db = (kTmaEpi && !kBnb && prm.res == nullptr && prm.R * prm.S == 1 &&
                (db_env >= 0 ? db_env != 0 : prm.num_kb <= 8)) ? 1 : 0;
  prm.epi_bufs = kTmaEpi ? (kBnb ? 3 * EW : ((prm.res != nullptr || prm.epi_db) ? 2 * EW : EW)) : 0;
  prm.epi_split = (kTmaEpi && BN == kSubW && epi_split_enabled()) ? 1 : 0;
  prm.res_mid = getenv("WLSEG_RES_MID") != nullptr ? atoi(getenv("WLSEG_RES_MID")) : 1;
  const int fixed = prm.epi_bufs * kWarpBufBytes + kBarBytes + (kTmaEpi ? 2 * BN * 4 : 0);
  int stages = (kSmemMax - fixed) / Cfg::kStageBytes;
  if (stages > kMaxStages) stages = kMaxStages;
  WLSEG_CHECK_ARG(stages >= 2, "conv(tcgen05): shared memory plan leaves fewer than 2 pipeline stages");
  prm.stages = stages;
  const int smem_bytes = stages * Cfg::kStageBytes + fixed;
  if constexpr (kPair) {
    static int max_pairs = -1;
    if (max_pairs < 0) max_pairs = max_active_pairs(conv_igemm_kernel<BN, TY, kTmaEpi, EW, true, kBnb>, 128 + 32 * EW, kSmemMax);
    WLSEG_CHECK_ARG(max_pairs > 0, "conv(tcgen05): the device cannot host a CTA pair of this kernel");
    prm.units = (int)ceil_div(prm.m_tiles, 2) * prm.n_tiles;
    int clusters = prm.units < max_pairs ? prm.units : max_pairs;
    if (clusters > conv_sms() / 2) clusters = conv_sms() / 2;
    if (prm.bn_sum != nullptr && clusters % prm.n_tiles != 0) clusters -= clusters % prm.n_tiles;
    WLSEG_CHECK_ARG(clusters > 0, "conv(tcgen05): no CTA pair fits the N-tile constraint of the fused statistics");
    WLSEG_CUDA(launch_pair(conv_igemm_kernel<BN, TY, kTmaEpi, EW, true, kBnb>, clusters, 128 + 32 * EW, smem_bytes, s, prm));
    return 0;
  }
  prm.units = prm.total_tiles;
  int grid = prm.total_tiles < conv_sms() ? prm.total_tiles : conv_sms();
  // fused BN statistics live in registers across tiles: every CTA must stay on one N tile
  if (kTmaEpi && prm.bn_sum != nullptr && grid % prm.n_tiles != 0) grid -= grid % prm.n_tiles;
  WLSEG_CHECK_ARG(grid > 0, "conv(tcgen05): no CTA fits the N-tile constraint of the fused statistics");
  WLSEG_CUDA(launch_pdl(conv_igemm_kernel<BN, TY, kTmaEpi, EW, false, kBnb>, dim3(grid), dim3(128 + 32 * EW), smem_bytes, s, prm));
  return 0;
}

template <int BN, typename TY, bool kTmaEpi>
static int launch_igemm(IgemmParams& prm, cudaStream_t s) {
  return launch_igemm_ew<BN, TY, kTmaEpi, 8>(prm, s);
}

// CTA pair (cta_group::2) for the 256-wide tiles with the staged epilogue.  WLSEG_PAIR: 0 = never, 1 = every such
// layer, unset = the layers that measured faster with it (see pair_default below)
static int pair_mode() {
  const char* e = getenv("WLSEG_PAIR");   // read per call: the parity tests force both forms in one process
  return e != nullptr ? atoi(e) : -1;
}

// BN-backward form of a data gradient (wlseg_conv2d_fprop_bnbwd): z travels in the residual slot
struct BnbArgs { const float* scale; const float* shift; const float* mean; const float* invstd; };

static int conv_fprop_igemm(const wlseg_conv_params* p, const void* x, const void* w, void* y, const float* scale,
                            const float* shift, const void* residual, double* bn_sum, double* bn_sqsum,
                            cudaStream_t s, const uint32_t* out_mask = nullptr, const BnbArgs* bnb = nullptr,
                            const wlseg_bn_finalize_args* fin = nullptr) {
  WLSEG_CHECK_ARG((((uintptr_t)x) & 15) == 0 && (((uintptr_t)w) & 15) == 0, "conv(tcgen05): x / w must be 16-byte aligned");
  IgemmParams prm;
  const int BN = pick_bn(p->K);
  // spatial patch: TW x TH = 128 output pixels, TW a power of two no wider than needed
  int tw_log2 = 4;
  if (p->Q <= 8) tw_log2 = 3;
  if (p->Q <= 4) tw_log2 = 2;
  if (p->P == 1) tw_log2 = 7;
  // halo form (see conv_igemm_kernel): C = 64, K = 64 (one N tile), stride 1, dilation 1, S in {1, 3}, 1 < R * S <= 9;
  // its tile is 8 x 16 pixels (S = 1) or 16 x 8 (S = 3)
  const bool halo_shape = BN == 64 && p->C == 64 && p->K == 64 && p->stride == 1 && p->dilation == 1 &&
                          (p->S == 1 || p->S == 3) && p->R * p->S <= 9 && p->R * p->S > 1 && p->W >= 16 && p->Q >= 16 &&
                          p->P > 1 && p->y_dtype != WLSEG_F32 && env_int("WLSEG_HALO", 1) != 0;
  if (halo_shape) tw_log2 = p->S == 1 ? 4 : 3;
  const int TW = 1 << tw_log2, TH = kBM / TW;
  WLSEG_CHECK_ARG(TW * p->stride <= 256 && TH * p->stride <= 256, "conv(tcgen05): TMA box too large");
  {
    uint64_t dims[4] = {(uint64_t)p->C, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->N};
    uint64_t strides[3] = {(uint64_t)p->x_pitch * 2, (uint64_t)p->x_pitch * 2 * p->W,
                           (uint64_t)p->x_pitch * 2 * p->W * p->H};
    uint32_t box[4] = {(uint32_t)kBK, (uint32_t)(TW * p->stride), (uint32_t)(TH * p->stride), 1};
    uint32_t estr[4] = {1, (uint32_t)p->stride, (uint32_t)p->stride, 1};
    if (int e = encode_tensor_map(&prm.map_a, x, 2, 4, dims, strides, box, estr, 0)) return e;
  }
  const bool f32out = (p->y_dtype == WLSEG_F32);
  // staged (TMA) epilogue: bf16 output whose pixel pitch and base keep every 64-channel row 16-byte aligned
  // (whole 64-channel sub-tiles only; scale / shift are read as float4)
  bool tma_epi = !f32out && BN >= kSubW && (p->K % kSubW == 0) && (p->y_pitch % 8 == 0) &&
                 ((((uintptr_t)y) & 15) == 0) && ((((uintptr_t)scale) & 15) == 0) && ((((uintptr_t)shift) & 15) == 0);
  // one epilogue warp stores (and loads the residual of) its 32 tile rows: bw x bh pixels
  const int bw = TW < 32 ? TW : 32, bh = 32 / bw;
  if (residual != nullptr && ((p->res_pitch % 8 != 0) || ((((uintptr_t)residual) & 15) != 0) ||
                              bw * p->res_stride > 256 || bh * p->res_stride > 256))
    tma_epi = false;
  // CTA pair (cta_group::2): each CTA stages half of the filter tile
  // Measured on B200, eval shapes 4 x 128 x 256 (gpurun_out -> profiles/r2_layers_eval_pair{0,1}.txt): the pair form wins
  // 5-15 % wherever a tile runs >= 8 k-blocks (3x3 layers, C >= 512) and on the 1x1 layers without a residual; the
  // shallow conv3 + residual layers (C <= 256: <= 4 k-blocks per 64 KB residual tile) lose 2-9 %: their epilogues are
  // the bottleneck and the pair couples two of them per accumulator hand-over.
  bool use_pair = false;
  if (BN == 256 && tma_epi) {
    const int mode = pair_mode();
    const int kblocks = p->R * p->S * (int)ceil_div(p->C, kBK);
    const bool pair_default = residual != nullptr ? kblocks >= 8 : kblocks >= 2;
    use_pair = (mode == 1) || (mode == -1 && pair_default);
    // the BN-backward form keeps three staging buffers per warp: only the pair's 32 KB stages leave a deep enough ring
    if (bnb != nullptr) use_pair = true;
  }
  WLSEG_CHECK_ARG(bnb == nullptr || (tma_epi && residual != nullptr && BN >= kSubW && p->res_stride == 1 &&
                                     bn_sum != nullptr && bn_sqsum != nullptr && scale == nullptr && p->relu == 0),
                  "conv_fprop_bnbwd: needs the staged bf16 epilogue (K %% 64 == 0, 16-byte aligned y / z, pitches %% 8 == 0)");
  bool use_halo = false;
  if (halo_shape && tma_epi) {
    prm.halo_rows = TH + p->R - 1;
    prm.halo_bytes = prm.halo_rows * 2048;
    prm.halo_sbo = p->S == 1 ? 1024 : 2048;
    prm.num_kb = p->R * p->S;
    const int bufs = bnb != nullptr ? 3 : (residual != nullptr ? 2 : 1);
    if (halo_smem_bytes(prm, BN, bufs) <= kSmemMax) {
      use_halo = true;
      use_pair = false;
      uint64_t dims[4] = {(uint64_t)p->C, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->N};
      uint64_t strides[3] = {(uint64_t)p->x_pitch * 2, (uint64_t)p->x_pitch * 2 * p->W, (uint64_t)p->x_pitch * 2 * p->W * p->H};
      uint32_t box[4] = {(uint32_t)kBK, 16, (uint32_t)prm.halo_rows, 1};
      uint32_t estr[4] = {1, 1, 1, 1};
      if (int e = encode_tensor_map(&prm.map_ah, x, 2, 4, dims, strides, box, estr, 4)) return e;
    }
  }
  {
    uint64_t dims[3] = {(uint64_t)p->C, (uint64_t)(p->R * p->S), (uint64_t)p->K};
    uint64_t strides[2] = {(uint64_t)p->C * 2, (uint64_t)p->C * 2 * p->R * p->S};
    uint32_t box[3] = {(uint32_t)kBK, 1, (uint32_t)(use_pair ? BN / 2 : BN)};
    uint32_t estr[3] = {1, 1, 1};
    if (int e = encode_tensor_map(&prm.map_b, w, 2, 3, dims, strides, box, estr, 1)) return e;
  }
  if (tma_epi) {
    {
      uint64_t dims[4] = {(uint64_t)p->K, (uint64_t)p->Q, (uint64_t)p->P, (uint64_t)p->N};
      uint64_t strides[3] = {(uint64_t)p->y_pitch * 2, (uint64_t)p->y_pitch * 2 * p->Q,
                             (uint64_t)p->y_pitch * 2 * p->Q * p->P};
      uint32_t box[4] = {(uint32_t)kSubW, (uint32_t)bw, (uint32_t)bh, 1};
      uint32_t estr[4] = {1, 1, 1, 1};
      if (int e = encode_tensor_map(&prm.map_y, y, 2, 4, dims, strides, box, estr, 2)) return e;
    }
    if (residual != nullptr) {
      const int rs = p->res_stride;
      uint64_t dims[4] = {(uint64_t)p->K, (uint64_t)p->res_W, (uint64_t)p->res_H, (uint64_t)p->N};
      uint64_t strides[3] = {(uint64_t)p->res_pitch * 2, (uint64_t)p->res_pitch * 2 * p->res_W,
                             (uint64_t)p->res_pitch * 2 * p->res_W * p->res_H};
      uint32_t box[4] = {(uint32_t)kSubW, (uint32_t)(bw * rs), (uint32_t)(bh * rs), 1};
      uint32_t estr[4] = {1, (uint32_t)rs, (uint32_t)rs, 1};
      if (int e = encode_tensor_map(&prm.map_r, residual, 2, 4, dims, strides, box, estr, 3)) return e;
    }
  }
  prm.y = y; prm.scale = scale; prm.shift = shift; prm.res = residual;
  prm.bn_sum = bn_sum; prm.bn_sqsum = bn_sqsum;
  prm.out_mask = out_mask;
  prm.bnb_scale = prm.bnb_shift = prm.bnb_mean = prm.bnb_invstd = nullptr;
  prm.fin = wlseg_bn_finalize_args{};
  if (fin != nullptr) {
    WLSEG_CHECK_ARG(tma_epi && bn_sum != nullptr && bnb == nullptr && !pdl_enabled(),
                    "conv_fprop_bn: needs the staged bf16 epilogue (K %% 64 == 0, aligned y) and no programmatic launches");
    prm.fin = *fin;
  }
  if (bnb != nullptr) {
    prm.bnb_scale = bnb->scale; prm.bnb_shift = bnb->shift; prm.bnb_mean = bnb->mean; prm.bnb_invstd = bnb->invstd;
  }
  WLSEG_CHECK_ARG(out_mask == nullptr || (tma_epi && p->K % 32 == 0 && (((uintptr_t)out_mask) & 3) == 0),
                  "conv_fprop_masked: needs the staged bf16 epilogue (K %% 64 == 0, aligned tensors)");
  prm.N = p->N; prm.P = p->P; prm.Q = p->Q; prm.K = p->K; prm.C = p->C;
  prm.R = p->R; prm.S = p->S; prm.stride = p->stride; prm.dilation = p->dilation;
  prm.pad_top = p->pad_top; prm.pad_left = p->pad_left;
  prm.y_pitch = p->y_pitch; prm.res_pitch = p->res_pitch; prm.res_stride = p->res_stride;
  prm.res_H = p->res_H; prm.res_W = p->res_W;
  prm.relu = p->relu;
  prm.reverse = p->reverse ? 1 : 0;
  prm.tw_log2 = tw_log2;
  prm.tiles_w = (int)ceil_div(p->Q, TW);
  prm.tiles_h = (int)ceil_div(p->P, TH);
  prm.n_tiles = (int)ceil_div(p->K, BN);
  const int64_t total = (int64_t)p->N * prm.tiles_h * prm.tiles_w * prm.n_tiles;
  WLSEG_CHECK_ARG(total < ((int64_t)1 << 31), "conv(tcgen05): too many tiles");
  prm.total_tiles = (int)total;
  prm.m_tiles = (int)(total / prm.n_tiles);
  prm.fd_n_tiles = make_fastdiv((uint32_t)prm.n_tiles);
  prm.fd_tiles_w = make_fastdiv((uint32_t)prm.tiles_w);
  prm.fd_tiles_h = make_fastdiv((uint32_t)prm.tiles_h);
  prm.units = prm.total_tiles;
  prm.cchunks = (int)ceil_div(p->C, kBK);
  prm.num_kb = p->R * p->S * prm.cchunks;
#define WLSEG_IGEMM_CASE(bn)                                                          \
  case bn:                                                                            \
    if (f32out) return launch_igemm<bn, float, false>(prm, s);                        \
    if (bn >= kSubW && tma_epi) return launch_igemm<bn, __nv_bfloat16, (bn >= kSubW)>(prm, s); \
    return launch_igemm<bn, __nv_bfloat16, false>(prm, s);
  if (use_halo) {
    if (bnb != nullptr) return launch_igemm_ew<64, __nv_bfloat16, true, 8, false, true, true>(prm, s);
    return launch_igemm_ew<64, __nv_bfloat16, true, 8, false, false, true>(prm, s);
  }
  if (bnb != nullptr) {
    if (BN == 256) return launch_igemm_ew<256, __nv_bfloat16, true, 8, true, true>(prm, s);
    if (BN == 128) return launch_igemm_ew<128, __nv_bfloat16, true, 8, false, true>(prm, s);
    return launch_igemm_ew<64, __nv_bfloat16, true, 8, false, true>(prm, s);
  }
  if (use_pair) return launch_igemm_ew<256, __nv_bfloat16, true, 8, true>(prm, s);
  switch (BN) {
    WLSEG_IGEMM_CASE(32)
    WLSEG_IGEMM_CASE(64)
    WLSEG_IGEMM_CASE(128)
    WLSEG_IGEMM_CASE(256)
  }
#undef WLSEG_IGEMM_CASE
  WLSEG_CHECK_ARG(false, "conv(tcgen05): internal tile selection error");
}

}  // namespace wlseg

using namespace wlseg;

extern "C" int wlseg_conv2d_tcgen05_supported(const wlseg_conv_params* p) {
  if (p == nullptr) return 0;
  return igemm_supported(p) ? 1 : 0;
}

extern "C" int wlseg_conv2d_fprop_masked(const wlseg_conv_params* p, const void* x, const void* w, void* y,
                                         const void* residual, const uint8_t* out_mask, wlseg_stream_t stream) {
  if (int e = check_conv_params(p)) return e;
  if (p->N == 0) return 0;
  WLSEG_CHECK_ARG(x && w && y && out_mask, "conv_fprop_masked: null pointer");
  WLSEG_CHECK_ARG(igemm_supported(p) && p->y_dtype == WLSEG_BF16 && p->relu == 0,
                  "conv_fprop_masked: tcgen05 bf16 configurations without ReLU only");
  if (residual != nullptr)
    WLSEG_CHECK_ARG(p->res_stride > 0 && p->res_pitch >= p->K && (p->P - 1) * p->res_stride < p->res_H &&
                        (p->Q - 1) * p->res_stride < p->res_W,
                    "conv_fprop_masked: residual geometry inconsistent");
  return conv_fprop_igemm(p, x, w, y, nullptr, nullptr, residual, nullptr, nullptr, (cudaStream_t)stream,
                          reinterpret_cast<const uint32_t*>(out_mask));
}

extern "C" int wlseg_conv2d_fprop_bnbwd(const wlseg_conv_params* p, const void* x, const void* w, void* y, const void* z,
                                        const float* scale, const float* shift, const float* mean, const float* invstd,
                                        double* dgamma, double* dbeta, wlseg_stream_t stream) {
  if (int e = check_conv_params(p)) return e;
  if (p->N == 0) return 0;
  WLSEG_CHECK_ARG(x && w && y && z && scale && shift && mean && invstd && dgamma && dbeta, "conv_fprop_bnbwd: null pointer");
  WLSEG_CHECK_ARG(igemm_supported(p) && p->y_dtype == WLSEG_BF16 && p->relu == 0 && p->K % 64 == 0,
                  "conv_fprop_bnbwd: tcgen05 bf16 configurations with K %% 64 == 0 and without ReLU only");
  WLSEG_CHECK_ARG(p->res_stride == 1 && p->res_pitch >= p->K && p->res_H == p->P && p->res_W == p->Q,
                  "conv_fprop_bnbwd: z must have the output's shape (res_* fields of the parameters describe it)");
  WLSEG_CHECK_ARG(((((uintptr_t)scale) | ((uintptr_t)shift)) & 15) == 0, "conv_fprop_bnbwd: scale / shift must be 16-byte aligned");
  const BnbArgs bnb = {scale, shift, mean, invstd};
  return conv_fprop_igemm(p, x, w, y, nullptr, nullptr, z, dgamma, dbeta, (cudaStream_t)stream, nullptr, &bnb);
}

extern "C" int wlseg_conv2d_fprop_bn(const wlseg_conv_params* p, const void* x, const void* w, void* y, double* bn_sum,
                                     double* bn_sqsum, const wlseg_bn_finalize_args* fin, wlseg_stream_t stream) {
  if (int e = check_conv_params(p)) return e;
  if (p->N == 0) return 0;
  WLSEG_CHECK_ARG(x && w && y && bn_sum && bn_sqsum && fin, "conv_fprop_bn: null pointer");
  WLSEG_CHECK_ARG(fin->count > 0 && fin->gamma && fin->beta && fin->scale && fin->shift && fin->saved_mean &&
                      fin->saved_invstd && fin->counter && ((fin->moving_mean == nullptr) == (fin->moving_var == nullptr)),
                  "conv_fprop_bn: incomplete wlseg_bn_finalize_args");
  WLSEG_CHECK_ARG(igemm_supported(p) && p->y_dtype == WLSEG_BF16 && p->relu == 0 && p->K % 64 == 0,
                  "conv_fprop_bn: tcgen05 bf16 configurations with K %% 64 == 0 and without ReLU only");
  return conv_fprop_igemm(p, x, w, y, nullptr, nullptr, nullptr, bn_sum, bn_sqsum, (cudaStream_t)stream, nullptr, nullptr, fin);
}

extern "C" int wlseg_conv2d_fprop(const wlseg_conv_params* p, const void* x, const void* w, void* y,
                                  const float* scale, const float* shift, const void* residual, double* bn_sum,
                                  double* bn_sqsum, wlseg_stream_t stream) {
  if (int e = check_conv_params(p)) return e;
  if (p->N == 0) return 0;
  WLSEG_CHECK_ARG(x && w && y, "conv_fprop: null pointer");
  WLSEG_CHECK_ARG((scale == nullptr) == (shift == nullptr), "conv_fprop: scale and shift must come together");
  WLSEG_CHECK_ARG((bn_sum == nullptr) == (bn_sqsum == nullptr), "conv_fprop: bn_sum and bn_sqsum must come together");
  if (residual != nullptr)
    WLSEG_CHECK_ARG(p->res_stride > 0 && p->res_pitch >= p->K && (p->P - 1) * p->res_stride < p->res_H &&
                        (p->Q - 1) * p->res_stride < p->res_W,
                    "conv_fprop: residual geometry inconsistent");
  int algo = p->algo;
  if (algo == WLSEG_ALGO_AUTO) algo = igemm_supported(p) ? WLSEG_ALGO_TCGEN05 : WLSEG_ALGO_DIRECT;
  if (algo == WLSEG_ALGO_TCGEN05) {
    WLSEG_CHECK_ARG(igemm_supported(p), "conv_fprop: configuration not covered by the tcgen05 kernel");
    return conv_fprop_igemm(p, x, w, y, scale, shift, residual, bn_sum, bn_sqsum, (cudaStream_t)stream);
  }
  WLSEG_CHECK_ARG(algo == WLSEG_ALGO_DIRECT, "conv_fprop: unknown algo %d", algo);
  WLSEG_CHECK_ARG(bn_sum == nullptr, "conv_fprop(direct): fused BN statistics are not available; call wlseg_bn_stats");
  return conv_fprop_direct(p, x, w, y, scale, shift, residual, (cudaStream_t)stream);
}
