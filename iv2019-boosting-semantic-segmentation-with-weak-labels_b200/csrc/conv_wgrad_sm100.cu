// Filter gradient (wgrad) of the convolutions on the 5th-generation tensor cores, sm_100a.
//
// Replaces TF Conv2DBackpropFilter, reached from create_train_op
// (code/estimator/define_estimator_hierarchical.py:120-129) for every slim.conv2d of
// code/models/resnet50_extended_feature_extractor.py:25-43 and
// code/models/resnet50_extended_model_hierarchical.py:60-64,80.
//
//   dw[k, r, s, c] = sum_{n,p,q} dy[n, p, q, k] * x[n, p*stride - pad + r*dil, q*stride - pad + s*dil, c]
//
// GEMM view (the reduction runs over PIXELS, so both operands are "MN-major": the GEMM row /
// column index - a channel - is the contiguous one in memory, which tcgen05 shared-memory
// descriptors express directly; nothing is transposed in HBM):
//   D[k, (tap, c)] += sum_pix A[k, pix] * B[(tap, c), pix]
//   A = dy, one TMA box {64 ch, TW, TH, 1} per 64 output channels  -> smem [64 pix][64 ch] (128 B rows)
//   B = x shifted by the tap offset (TMA zero fill = padding), one box per (tap, 64-channel chunk);
//       an N tile is 4 such chunks, so a 3x3 kernel over 64 channels still fills N = 256
//   one pipeline stage = one patch of 64 pixels = 4 MMAs (K = 16 pixels each)
//   D = fp32 128 x 256 in TMEM, double buffered
// Work unit = (output tile, pixel split): the pixel range is split so that ~2 x 148 units exist;
// partial tiles are combined by the TMA itself: every epilogue warp stages its 32 x 32 fp32 sub-tiles
// in swizzled shared memory and issues cp.reduce.async.bulk.tensor (.add) into dw, which the caller
// (this file's entry point) zeroes first.
//
// Warp roles as in conv_igemm_sm100.cu: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM
// allocator, warps 4-11 epilogue (independent per-warp drains).
#include <cuda.h>

#include "common.cuh"
#include "tc_sm100.cuh"

namespace wlseg {

int check_conv_params(const wlseg_conv_params* p);
int encode_tensor_map(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estr, int kind);

constexpr int kWgEpiWarps = 8;
constexpr int kWgThreads = 128 + 32 * kWgEpiWarps;
constexpr int kWgEpiWarp0 = 4;
constexpr int kWgWarpBufBytes = 32 * 128;   // epilogue staging per warp: 32 rows x 32 fp32
constexpr int kPix = 64;                    // pixels per stage (reduction depth of one stage)
constexpr int kChunkBytes = kPix * 128;     // one {64 ch x 64 pix} bf16 box = 8 KB
constexpr int kWgABytes = 2 * kChunkBytes;  // M = 128 output channels
constexpr int kWgMaxStages = 8;
constexpr int kWgSmemMax = 227 * 1024;

struct WgradParams {
  CUtensorMap map_dy;  // bf16 {K, Q, P, N}
  CUtensorMap map_x;   // bf16 {C, W, H, N}
  CUtensorMap map_dw;  // fp32 {C, R*S, K}
  int K, C, S;
  int stride, dilation, pad_top, pad_left;
  int tw_log2, th;                 // pixel patch: (1 << tw_log2) x th = 64 pixels
  int patches_w, patches_h, patches;
  int cchunks, chunks_total;       // ceil(C / 64), R*S*cchunks
  int m_tiles, n_tiles, tiles;
  int splits, patches_per_split;
  int units;
  int stages, epi_bufs;
  int swap_offsets;                // debug: exchange LBO / SBO in the MN-major descriptors
  // every runtime divisor of the kernel as multiply-shift constants (common.cuh FastDiv): the producer used to spend
  // three ~150-cycle division chains per 64-pixel stage, next to 512 cycles of tensor time for that stage
  FastDiv fd_tiles, fd_n_tiles, fd_cchunks, fd_S, fd_patches_w, fd_patches_h;
};

// MN-major SWIZZLE_128B shared-memory descriptor: 64 contiguous MN elements (128 B) per row, 8
// reduction rows per 1024-byte swizzle atom.  LBO = bytes between 64-element MN chunks,
// SBO = bytes between 8-row groups along the reduction dimension.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor with both operands MN-major (bits 15 and 16)
__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// kPair (BN = 256 only): CTA pair, tcgen05 cta_group::2 - one MMA of M = 256 output channels; CTA `rank` stages ITS 128
// channels of dy and HALF of the x chunks of the N tile (chunks 2 * rank, 2 * rank + 1): 32 KB per stage and CTA instead
// of 48 KB for the same 4.2 MFLOP per SM - the kernel was bound by the L2 -> SM ingress.
template <int BN, bool kPair>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ WgradParams prm) {
  static_assert(!kPair || BN == 256, "the pair form is built for 256-wide N tiles");
  constexpr int CH = BN / 64;                              // B chunks per N tile
  constexpr int CHL = kPair ? CH / 2 : CH;                 // B chunks this CTA stages
  constexpr int kStageBytes = kWgABytes + CHL * kChunkBytes;
  constexpr int kTmemCols = 2 * BN;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const int unit0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_launch_dependents();
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int stages = prm.stages;
  uint8_t* epi_smem = smem + stages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + kWgEpiWarps * kWgWarpBufBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kWgMaxStages;
  uint64_t* tfull_bar = bars + 2 * kWgMaxStages;
  uint64_t* tempty_bar = bars + 2 * kWgMaxStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWgMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&prm.map_dy);
    tma_prefetch_desc(&prm.map_x);
    tma_prefetch_desc(&prm.map_dw);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(full_bar + s), 1);
      mbar_init(smem_u32(empty_bar + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(tfull_bar + a), 1);
      mbar_init(smem_u32(tempty_bar + a), kPair ? 2 * kWgEpiWarps : kWgEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (kPair) tmem_alloc_pair(smem_u32(tmem_slot), kTmemCols);
    else tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // prologue done; dz / x of the previous kernels are read from here on

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = unit0; unit < prm.units; unit += unit_step) {
        uint32_t tile, split, nt, mt;
        prm.fd_tiles.divmod((uint32_t)unit, split, tile);
        prm.fd_n_tiles.divmod(tile, mt, nt);
        const int k0 = kPair ? ((int)mt * 2 + (int)rank) * 128 : (int)mt * 128;
        int nch = prm.chunks_total - (int)nt * CH;
        if (nch > CH) nch = CH;
        if (kPair) nch = CHL;   // the host takes the pair form only when every N tile is complete
        // tap offsets / channel origins of this tile's B chunks
        int coff[CHL], xoff[CHL], yoff[CHL];
#pragma unroll
        for (int j = 0; j < CHL; ++j) {
          const int id = (int)nt * CH + (kPair ? (int)rank * CHL : 0) + j;
          const int tap = (int)prm.fd_cchunks.div((uint32_t)id);
          const int r = (int)prm.fd_S.div((uint32_t)tap), s = tap - r * prm.S;
          coff[j] = (id - tap * prm.cchunks) * 64;
          xoff[j] = s * prm.dilation - prm.pad_left;
          yoff[j] = r * prm.dilation - prm.pad_top;
        }
        const int pp0 = (int)split * prm.patches_per_split;
        int pp1 = pp0 + prm.patches_per_split;
        if (pp1 > prm.patches) pp1 = prm.patches;
        // patch coordinates: decoded once per unit, then walked with counters
        uint32_t rest, upwi, un, uphi;
        prm.fd_patches_w.divmod((uint32_t)pp0, rest, upwi);
        prm.fd_patches_h.divmod(rest, un, uphi);
        int pwi = (int)upwi, phi = (int)uphi, n = (int)un;
        for (int pp = pp0; pp < pp1; ++pp) {
          const int q0 = pwi << prm.tw_log2, p0 = phi * prm.th;
          mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
          const uint32_t a_dst = smem_u32(smem + stage * kStageBytes);
          const uint32_t b_dst = a_dst + kWgABytes;
          const uint32_t bar = smem_u32(full_bar + stage);
          if constexpr (kPair) {
            if (rank == 0) mbar_arrive_expect_tx(bar, (uint32_t)(2 * kStageBytes));   // both CTAs' boxes
            const uint32_t lbar = mapa_shared(bar, 0);
            tma_load_4d_pair(a_dst, &prm.map_dy, lbar, k0, q0, p0, n);
            tma_load_4d_pair(a_dst + kChunkBytes, &prm.map_dy, lbar, k0 + 64, q0, p0, n);
#pragma unroll
            for (int j = 0; j < CHL; ++j)
              tma_load_4d_pair(b_dst + j * kChunkBytes, &prm.map_x, lbar, coff[j], q0 * prm.stride + xoff[j],
                               p0 * prm.stride + yoff[j], n);
          } else {
            mbar_arrive_expect_tx(bar, (uint32_t)((2 + nch) * kChunkBytes));
            tma_load_4d(a_dst, &prm.map_dy, bar, k0, q0, p0, n);
            tma_load_4d(a_dst + kChunkBytes, &prm.map_dy, bar, k0 + 64, q0, p0, n);
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (j < nch)
                tma_load_4d(b_dst + j * kChunkBytes, &prm.map_x, bar, coff[j], q0 * prm.stride + xoff[j],
                            p0 * prm.stride + yoff[j], n);
          }
          if (++stage == stages) { stage = 0; phase ^= 1; }
          if (++pwi == prm.patches_w) {
            pwi = 0;
            if (++phi == prm.patches_h) { phi = 0; ++n; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_mn(kPair ? 256 : 128, BN);
      const uint32_t lbo = prm.swap_offsets ? 1024u : (uint32_t)kChunkBytes;
      const uint32_t sbo = prm.swap_offsets ? (uint32_t)kChunkBytes : 1024u;
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int unit = unit0; unit < prm.units; unit += unit_step, ++iter) {
        const int split = (int)prm.fd_tiles.div((uint32_t)unit);
        const int pp0 = split * prm.patches_per_split;
        int pp1 = pp0 + prm.patches_per_split;
        if (pp1 > prm.patches) pp1 = prm.patches;
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        mbar_wait(smem_u32(tempty_bar + acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int pp = pp0; pp < pp1; ++pp) {
          mbar_wait(smem_u32(full_bar + stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
          const uint32_t b_addr = a_addr + kWgABytes;
#pragma unroll
          for (int k = 0; k < kPix / 16; ++k) {
            // 16 pixels = two 8-row swizzle atoms = 2048 bytes further down both operands
            const uint64_t adesc = make_smem_desc_mn(a_addr + k * 2048, lbo, sbo);
            const uint64_t bdesc = make_smem_desc_mn(b_addr + k * 2048, lbo, sbo);
            if (kPair) umma_bf16_pair(d_tmem, adesc, bdesc, idesc, (uint32_t)((pp != pp0) | (k != 0)));
            else umma_bf16(d_tmem, adesc, bdesc, idesc, (uint32_t)((pp != pp0) | (k != 0)));
          }
          if (kPair) umma_commit_pair(smem_u32(empty_bar + stage), 3); else umma_commit(smem_u32(empty_bar + stage));
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        if (kPair) umma_commit_pair(smem_u32(tfull_bar + acc), 3); else umma_commit(smem_u32(tfull_bar + acc));
      }
    }
  } else if (warp >= kWgEpiWarp0) {
    // ===================== epilogue: TMEM -> swizzled smem -> TMA reduce-add into dw =====================
    // one independent pipeline per warp (as in conv_igemm_sm100.cu): warp e owns the 32 output channels
    // of TMEM lane quarter e % 4 and every second 32-column sub-tile; its own 4 KB staging buffer and
    // its own cp.reduce.async.bulk.tensor - no CTA-wide barrier in the drain
    const int ew = warp - kWgEpiWarp0;
    const int quarter = ew & 3, cgrp = ew >> 2;
    constexpr int CG = kWgEpiWarps / 4;
    const uint32_t sw = (uint32_t)(lane & 7);
    uint8_t* buf = epi_smem + ew * kWgWarpBufBytes;
    uint8_t* myrow = buf + lane * 128;
    auto release_acc = [&](int a) {
      if (kPair) mbar_arrive_cluster(mapa_shared(smem_u32(tempty_bar + a), 0));
      else mbar_arrive(smem_u32(tempty_bar + a));
    };
    int iter = 0;
    for (int unit = unit0; unit < prm.units; unit += unit_step, ++iter) {
      uint32_t tile, split_, unt, umt;
      prm.fd_tiles.divmod((uint32_t)unit, split_, tile);
      prm.fd_n_tiles.divmod(tile, umt, unt);
      const int nt = (int)unt, k0 = kPair ? ((int)umt * 2 + (int)rank) * 128 : (int)umt * 128;
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      mbar_wait(smem_u32(tfull_bar + acc), acc_phase);
      tc_fence_after();
      // sub-tiles that exist: chunk < chunks_total and channel origin < C
      int nsub = 0;
      for (int j = 0; j < 2 * CH; ++j) {
        const int id = nt * CH + (j >> 1);
        if (id >= prm.chunks_total) break;
        const int tap = (int)prm.fd_cchunks.div((uint32_t)id);
        if ((id - tap * prm.cchunks) * 64 + (j & 1) * 32 < prm.C) nsub = j + 1;
      }
      // this warp's last sub-tile index (its final TMEM read of the unit)
      int last = -1;
      for (int j = cgrp; j < nsub; j += CG) last = j;
      if (last < 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(acc);
      }
      for (int j = cgrp; j < nsub; j += CG) {
        const int id = nt * CH + (j >> 1);
        const int tap = (int)prm.fd_cchunks.div((uint32_t)id);
        const int c0 = (id - tap * prm.cchunks) * 64 + (j & 1) * 32;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + j * 32), v);
        tmem_ld_wait();
        if (j == last) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc(acc);
        }
        if (c0 >= prm.C) continue;  // warp-uniform (odd C chunk): nothing to store
        if (lane == 0) bulk_wait_read<0>();   // the reduce that last left from this buffer has read it
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<uint4*>(myrow + ((((uint32_t)g) ^ sw) << 4)) =
              make_uint4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_3d(&prm.map_dw, smem_u32(buf), c0, tap, k0 + quarter * 32);
          bulk_commit();
        }
      }
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_relaxed();
  if (warp == 2) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

static bool wgrad_tc_supported(const wlseg_conv_params* p) {
  if (p->dtype != WLSEG_BF16 || p->y_dtype != WLSEG_BF16) return false;
  if (p->stride != 1 && p->stride != 2) return false;
  if (p->C % 8 != 0 || p->x_pitch % 8 != 0 || p->y_pitch % 8 != 0) return false;
  if (p->R * p->S > 64 || p->dilation > 64 || p->N > 65535) return false;
  return true;
}

template <int BN, bool kPair = false>
static int launch_wgrad(WgradParams& prm, cudaStream_t s) {
  constexpr int kStageBytes = kWgABytes + (BN / 64 / (kPair ? 2 : 1)) * kChunkBytes;
  static bool configured = false;
  if (!configured) {
    WLSEG_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<BN, kPair>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemMax));
    configured = true;
  }
  prm.epi_bufs = kWgEpiWarps;
  const int fixed = kWgEpiWarps * kWgWarpBufBytes + 256;
  int stages = (kWgSmemMax - fixed) / kStageBytes;
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  prm.stages = stages;
  const int smem_bytes = stages * kStageBytes + fixed;
  if constexpr (kPair) {
    const int clusters = prm.units < conv_sms() / 2 ? prm.units : conv_sms() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kWgThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    WLSEG_CUDA(cudaLaunchKernelEx(&cfg, conv_wgrad_kernel<BN, true>, prm));
    return 0;
  }
  const int grid = prm.units < conv_sms() ? prm.units : conv_sms();
  WLSEG_CUDA(launch_pdl(conv_wgrad_kernel<BN, false>, dim3(grid), dim3(kWgThreads), smem_bytes, s, prm));
  return 0;
}

int conv_wgrad_tcgen05(const wlseg_conv_params* p, const void* x, const void* dy, float* dw, cudaStream_t s) {
  WLSEG_CHECK_ARG((((uintptr_t)x) & 15) == 0 && (((uintptr_t)dy) & 15) == 0 && (((uintptr_t)dw) & 15) == 0,
                  "conv_wgrad(tcgen05): x / dy / dw must be 16-byte aligned");
  WLSEG_CHECK_ARG(p->C % 4 == 0, "conv_wgrad(tcgen05): C must be a multiple of 4");
  WgradParams prm;
  int tw_log2 = 3;  // 8 x 8 pixel patches
  if (p->Q <= 4) tw_log2 = 2;
  if (p->P == 1) tw_log2 = 6;
  if (p->P <= 4 && p->P > 1 && p->Q > 8) tw_log2 = 4;
  const int TW = 1 << tw_log2, TH = kPix / TW;
  WLSEG_CHECK_ARG(TW * p->stride <= 256 && TH * p->stride <= 256, "conv_wgrad(tcgen05): TMA box too large");
  {
    uint64_t dims[4] = {(uint64_t)p->K, (uint64_t)p->Q, (uint64_t)p->P, (uint64_t)p->N};
    uint64_t strides[3] = {(uint64_t)p->y_pitch * 2, (uint64_t)p->y_pitch * 2 * p->Q,
                           (uint64_t)p->y_pitch * 2 * p->Q * p->P};
    uint32_t box[4] = {64, (uint32_t)TW, (uint32_t)TH, 1};
    uint32_t estr[4] = {1, 1, 1, 1};
    if (int e = encode_tensor_map(&prm.map_dy, dy, 2, 4, dims, strides, box, estr, 10)) return e;
  }
  {
    uint64_t dims[4] = {(uint64_t)p->C, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->N};
    uint64_t strides[3] = {(uint64_t)p->x_pitch * 2, (uint64_t)p->x_pitch * 2 * p->W,
                           (uint64_t)p->x_pitch * 2 * p->W * p->H};
    uint32_t box[4] = {64, (uint32_t)(TW * p->stride), (uint32_t)(TH * p->stride), 1};
    uint32_t estr[4] = {1, (uint32_t)p->stride, (uint32_t)p->stride, 1};
    if (int e = encode_tensor_map(&prm.map_x, x, 2, 4, dims, strides, box, estr, 11)) return e;
  }
  {
    uint64_t dims[3] = {(uint64_t)p->C, (uint64_t)(p->R * p->S), (uint64_t)p->K};
    uint64_t strides[2] = {(uint64_t)p->C * 4, (uint64_t)p->C * 4 * p->R * p->S};
    uint32_t box[3] = {32, 1, 32};   // one epilogue warp's 32 output channels x 32 fp32
    uint32_t estr[3] = {1, 1, 1};
    if (int e = encode_tensor_map(&prm.map_dw, dw, 4, 3, dims, strides, box, estr, 12)) return e;
  }
  prm.K = p->K; prm.C = p->C; prm.S = p->S;
  prm.stride = p->stride; prm.dilation = p->dilation; prm.pad_top = p->pad_top; prm.pad_left = p->pad_left;
  prm.tw_log2 = tw_log2; prm.th = TH;
  prm.patches_w = (int)ceil_div(p->Q, TW);
  prm.patches_h = (int)ceil_div(p->P, TH);
  const int64_t patches = (int64_t)p->N * prm.patches_h * prm.patches_w;
  WLSEG_CHECK_ARG(patches < ((int64_t)1 << 30), "conv_wgrad(tcgen05): too many pixel patches");
  prm.patches = (int)patches;
  prm.cchunks = (int)ceil_div(p->C, 64);
  prm.chunks_total = p->R * p->S * prm.cchunks;
  const int BN = prm.chunks_total >= 4 ? 256 : (prm.chunks_total >= 2 ? 128 : 64);
  const int CH = BN / 64;
  prm.m_tiles = (int)ceil_div(p->K, 128);
  prm.n_tiles = (int)ceil_div(prm.chunks_total, CH);
  // CTA pair: two M tiles per unit; whole 256-channel output pairs and complete N tiles only.  WLSEG_WGRAD_PAIR=0/1
  const char* pair_e = getenv("WLSEG_WGRAD_PAIR");   // read per call: the parity tests force both forms
  const int pair_env = pair_e != nullptr ? atoi(pair_e) : -1;
  const bool use_pair = BN == 256 && (p->K % 256 == 0) && (prm.chunks_total % 4 == 0) && (p->C % 64 == 0) && pair_env != 0;
  if (use_pair) prm.m_tiles /= 2;
  prm.tiles = prm.m_tiles * prm.n_tiles;
  // pixel splits: about one wave of work units, at least 4 patches per split
  // work units per SM (pair) the pixel splits aim at.  Round 1 took two waves; every split adds one fp32 reduce-add
  // pass over dw through L2 (a 1 MB 256 <-> 1024 gradient was summed 37 times: those launches sat at 0.46 of their byte
  // bound), and one wave measured faster on the whole step: 10.84 -> 10.68 / 10.75 ms (three waves: 10.88 / 10.97).
  static const int waves = [] {
    const char* e = getenv("WLSEG_WGRAD_WAVES");
    return e != nullptr && atoi(e) > 0 ? atoi(e) : 1;
  }();
  int splits = (waves * (use_pair ? conv_sms() / 2 : conv_sms())) / prm.tiles;
  const int max_splits = prm.patches / 4 > 0 ? prm.patches / 4 : 1;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  prm.patches_per_split = (int)ceil_div(prm.patches, splits);
  prm.splits = (int)ceil_div(prm.patches, prm.patches_per_split);
  prm.units = prm.tiles * prm.splits;
  prm.fd_tiles = make_fastdiv((uint32_t)prm.tiles);
  prm.fd_n_tiles = make_fastdiv((uint32_t)prm.n_tiles);
  prm.fd_cchunks = make_fastdiv((uint32_t)prm.cchunks);
  prm.fd_S = make_fastdiv((uint32_t)prm.S);
  prm.fd_patches_w = make_fastdiv((uint32_t)prm.patches_w);
  prm.fd_patches_h = make_fastdiv((uint32_t)prm.patches_h);
  static const int swap = [] {
    const char* e = getenv("WLSEG_WGRAD_SWAP_OFFSETS");
    return (e != nullptr && e[0] == '1') ? 1 : 0;
  }();
  prm.swap_offsets = swap;
  if (use_pair) return launch_wgrad<256, true>(prm, s);
  switch (BN) {
    case 64: return launch_wgrad<64>(prm, s);
    case 128: return launch_wgrad<128>(prm, s);
    default: return launch_wgrad<256>(prm, s);
  }
}

bool conv_wgrad_tcgen05_supported(const wlseg_conv_params* p) { return wgrad_tc_supported(p); }

}  // namespace wlseg
