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
//   warps 4-7 epilogue     (tcgen05.ld 32 lanes x 32 columns -> scale/shift/residual/ReLU ->
//                           bf16 (or fp32) NHWC stores; overlaps the next tile's MMAs)
//
// Roofline: tensor-bound for the block3/block4 3x3 and wide 1x1 layers
// (flops = 2*N*P*Q*R*S*C*K); HBM-bound for block1 and the 64-channel 1x1 layers.
#include <cuda.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace wlseg {

int check_conv_params(const wlseg_conv_params* p);
int conv_fprop_direct(const wlseg_conv_params* p, const void* x, const void* w, void* y, const float* scale,
                      const float* shift, const void* residual, cudaStream_t s);

constexpr int kBM = 128;          // pixels per tile (UMMA M)
constexpr int kBK = 64;           // channels per pipeline stage (one 128-byte swizzle row)
constexpr int kUmmaK = 16;        // bf16 MMA K
constexpr int kIgemmThreads = 256;
constexpr int kEpiWarp0 = 4;      // first epilogue warp
constexpr int kABytes = kBM * kBK * 2;  // 16 KB
constexpr int kSmemBudget = 200 * 1024;

struct IgemmParams {
  CUtensorMap map_a;  // activations {C, W, H, N}
  CUtensorMap map_b;  // filters {C, R*S, K}
  void* y;
  const float* scale;
  const float* shift;
  const void* res;
  double* bn_sum;
  double* bn_sqsum;
  int N, P, Q, K, C;
  int R, S, stride, dilation, pad_top, pad_left;
  int y_pitch, res_pitch, res_stride, res_H, res_W;
  int relu;
  int tw_log2;        // tile is TH x TW pixels with TW = 1 << tw_log2, TH = 128 / TW
  int tiles_w, tiles_h, n_tiles, total_tiles;
  int cchunks;        // ceil(C / 64)
  int num_kb;         // R * S * cchunks
};

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, M = 128, N from the instruction descriptor
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory operand descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (= 1024 B:
//   8 rows x 128 B per swizzle atom) | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 [4,6)=1, A bf16 [7,10)=1,
// B bf16 [10,13)=1, A/B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int BN>
struct IgemmCfg {
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagesRaw = (kSmemBudget - 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// ----------------------------------------------------------------------------- kernel
template <int BN, typename TY>
__global__ void __launch_bounds__(kIgemmThreads, 1)
conv_igemm_kernel(const __grid_constant__ IgemmParams prm) {
  using Cfg = IgemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned stages
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                         // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;         // [kStages]
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;     // [2]
  uint64_t* tempty_bar = bars + 2 * Cfg::kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&prm.map_a);
    tma_prefetch_desc(&prm.map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(smem_u32(full_bar + s), 1);
      mbar_init(smem_u32(empty_bar + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(tfull_bar + a), 1);
      mbar_init(smem_u32(tempty_bar + a), 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int TW = 1 << prm.tw_log2;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x) {
        const int nt = tile % prm.n_tiles;
        int mt = tile / prm.n_tiles;
        const int twi = mt % prm.tiles_w; mt /= prm.tiles_w;
        const int thi = mt % prm.tiles_h;
        const int n = mt / prm.tiles_h;
        const int q0 = twi << prm.tw_log2;
        const int p0 = thi * (kBM >> prm.tw_log2);
        const int k0 = nt * BN;
        for (int kb = 0; kb < prm.num_kb; ++kb) {
          const int tap = kb / prm.cchunks;
          const int cc = kb - tap * prm.cchunks;
          const int r = tap / prm.S;
          const int s = tap - r * prm.S;
          mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
          const uint32_t a_dst = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_dst = a_dst + kABytes;
          const uint32_t bar = smem_u32(full_bar + stage);
          mbar_arrive_expect_tx(bar, Cfg::kStageBytes);
          tma_load_4d(a_dst, &prm.map_a, bar, cc * kBK, q0 * prm.stride - prm.pad_left + s * prm.dilation,
                      p0 * prm.stride - prm.pad_top + r * prm.dilation, n);
          tma_load_3d(b_dst, &prm.map_b, bar, cc * kBK, tap, k0);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x, ++iter) {
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
            umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          umma_commit(smem_u32(empty_bar + stage));              // stage free once these MMAs retire
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(tfull_bar + acc));                  // accumulator complete
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue =====================
    const int ew = warp - kEpiWarp0;          // TMEM lanes [32*ew, 32*ew + 32)
    const int row = ew * 32 + lane;           // tile row = pixel within the patch
    const int dy_ = row >> prm.tw_log2, dx_ = row & (TW - 1);
    int iter = 0;
    for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x, ++iter) {
      const int nt = tile % prm.n_tiles;
      int mt = tile / prm.n_tiles;
      const int twi = mt % prm.tiles_w; mt /= prm.tiles_w;
      const int thi = mt % prm.tiles_h;
      const int n = mt / prm.tiles_h;
      const int q = (twi << prm.tw_log2) + dx_;
      const int p = thi * (kBM >> prm.tw_log2) + dy_;
      const bool valid = (p < prm.P) && (q < prm.Q);
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
      for (int ch = 0; ch < BN / 32; ++ch) {
        const int kbase = nt * BN + ch * 32;
        if (kbase >= prm.K) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN + ch * 32), v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);

        if (prm.bn_sum != nullptr) {
          // training-mode BN statistics of the raw accumulators (rows outside the image masked)
          float mys = 0.f, mysq = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float a = valid ? f[j] : 0.f;
            float s1 = warp_sum(a);
            float s2 = warp_sum(a * a);
            if (lane == j) { mys = s1; mysq = s2; }
          }
          if (kbase + lane < prm.K) {
            atomicAdd(prm.bn_sum + kbase + lane, (double)mys);
            atomicAdd(prm.bn_sqsum + kbase + lane, (double)mysq);
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
            if (full && (prm.res_pitch % 8 == 0)) {
              const uint4* rp = reinterpret_cast<const uint4*>(rrow + kbase);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint4 raw = __ldg(rp + g);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float2 t = __bfloat1622float2(h[e]);
                  f[g * 8 + 2 * e] += t.x;
                  f[g * 8 + 2 * e + 1] += t.y;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (kbase + j < prm.K) f[j] += __bfloat162float(rrow[kbase + j]);
            }
          }
          if (prm.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (sizeof(TY) == 2) {
            __nv_bfloat16* yo = reinterpret_cast<__nv_bfloat16*>(yrow) + kbase;
            if (full && (prm.y_pitch % 8 == 0)) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint4 raw;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
                for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[g * 8 + 2 * e], f[g * 8 + 2 * e + 1]);
                reinterpret_cast<uint4*>(yo)[g] = raw;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (kbase + j < prm.K) yo[j] = __float2bfloat16_rn(f[j]);
            }
          } else {
            float* yo = reinterpret_cast<float*>(yrow) + kbase;
            if (full && (prm.y_pitch % 4 == 0)) {
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
      if (lane == 0) mbar_arrive(smem_u32(tempty_bar + acc));
    }
  }

  // teardown: everyone done with TMEM before it is freed
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ----------------------------------------------------------------------------- host side
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
typedef std::tuple<const void*, int, int, int, int, int, int, int, int, int> MapKey;
static std::mutex g_map_mutex;
static std::map<MapKey, CUtensorMap> g_map_cache;

static int encode_cached(const MapKey& key, CUtensorMap* out, int rank, const void* base, const cuuint64_t* dims,
                         const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr) {
  std::lock_guard<std::mutex> lock(g_map_mutex);
  auto it = g_map_cache.find(key);
  if (it != g_map_cache.end()) {
    *out = it->second;
    return 0;
  }
  EncodeTiledFn fn = get_encode_fn();
  WLSEG_CHECK_ARG(fn != nullptr, "conv(tcgen05): cuTensorMapEncodeTiled not available from the driver");
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  WLSEG_CHECK_ARG(r == CUDA_SUCCESS, "conv(tcgen05): cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
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

template <int BN, typename TY>
static int launch_igemm(const IgemmParams& prm, cudaStream_t s) {
  using Cfg = IgemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    WLSEG_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BN, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::kSmemBytes));
    configured = true;
  }
  int grid = prm.total_tiles < kNumSMs ? prm.total_tiles : kNumSMs;
  conv_igemm_kernel<BN, TY><<<grid, kIgemmThreads, Cfg::kSmemBytes, s>>>(prm);
  WLSEG_LAUNCH_CHECK();
  return 0;
}

static int conv_fprop_igemm(const wlseg_conv_params* p, const void* x, const void* w, void* y, const float* scale,
                            const float* shift, const void* residual, double* bn_sum, double* bn_sqsum,
                            cudaStream_t s) {
  WLSEG_CHECK_ARG((((uintptr_t)x) & 15) == 0 && (((uintptr_t)w) & 15) == 0, "conv(tcgen05): x / w must be 16-byte aligned");
  IgemmParams prm;
  const int BN = pick_bn(p->K);
  // spatial patch: TW x TH = 128 output pixels, TW a power of two no wider than needed
  int tw_log2 = 4;
  if (p->Q <= 8) tw_log2 = 3;
  if (p->Q <= 4) tw_log2 = 2;
  if (p->P == 1) tw_log2 = 7;
  const int TW = 1 << tw_log2, TH = kBM / TW;
  WLSEG_CHECK_ARG(TW * p->stride <= 256 && TH * p->stride <= 256, "conv(tcgen05): TMA box too large");
  {
    cuuint64_t dims[4] = {(cuuint64_t)p->C, (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->N};
    cuuint64_t strides[3] = {(cuuint64_t)p->x_pitch * 2, (cuuint64_t)p->x_pitch * 2 * p->W,
                             (cuuint64_t)p->x_pitch * 2 * p->W * p->H};
    cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)(TW * p->stride), (cuuint32_t)(TH * p->stride), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)p->stride, (cuuint32_t)p->stride, 1};
    MapKey key(x, 4, p->C, p->W, p->H, p->N, p->x_pitch, TW, p->stride, 0);
    if (int e = encode_cached(key, &prm.map_a, 4, x, dims, strides, box, estr)) return e;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)p->C, (cuuint64_t)(p->R * p->S), (cuuint64_t)p->K};
    cuuint64_t strides[2] = {(cuuint64_t)p->C * 2, (cuuint64_t)p->C * 2 * p->R * p->S};
    cuuint32_t box[3] = {(cuuint32_t)kBK, 1, (cuuint32_t)BN};
    cuuint32_t estr[3] = {1, 1, 1};
    MapKey key(w, 3, p->C, p->R * p->S, p->K, BN, 0, 0, 0, 1);
    if (int e = encode_cached(key, &prm.map_b, 3, w, dims, strides, box, estr)) return e;
  }
  prm.y = y; prm.scale = scale; prm.shift = shift; prm.res = residual;
  prm.bn_sum = bn_sum; prm.bn_sqsum = bn_sqsum;
  prm.N = p->N; prm.P = p->P; prm.Q = p->Q; prm.K = p->K; prm.C = p->C;
  prm.R = p->R; prm.S = p->S; prm.stride = p->stride; prm.dilation = p->dilation;
  prm.pad_top = p->pad_top; prm.pad_left = p->pad_left;
  prm.y_pitch = p->y_pitch; prm.res_pitch = p->res_pitch; prm.res_stride = p->res_stride;
  prm.res_H = p->res_H; prm.res_W = p->res_W;
  prm.relu = p->relu;
  prm.tw_log2 = tw_log2;
  prm.tiles_w = (int)ceil_div(p->Q, TW);
  prm.tiles_h = (int)ceil_div(p->P, TH);
  prm.n_tiles = (int)ceil_div(p->K, BN);
  const int64_t total = (int64_t)p->N * prm.tiles_h * prm.tiles_w * prm.n_tiles;
  WLSEG_CHECK_ARG(total < ((int64_t)1 << 31), "conv(tcgen05): too many tiles");
  prm.total_tiles = (int)total;
  prm.cchunks = (int)ceil_div(p->C, kBK);
  prm.num_kb = p->R * p->S * prm.cchunks;
  const bool f32out = (p->y_dtype == WLSEG_F32);
#define WLSEG_IGEMM_CASE(bn)                                                   \
  case bn:                                                                     \
    return f32out ? launch_igemm<bn, float>(prm, s) : launch_igemm<bn, __nv_bfloat16>(prm, s);
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
