// K4: bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
//   C[m][n] = sum_k A[m][k] * W[n][k]      A: activations (K-major), W: nn.Linear / Conv1d weight (K-major)
//
// This is every dense contraction of the reference's exported graphs
// (/root/reference/model_convert/export_onnx.py: conv1/conv2 :158-159 as implicit GEMMs over overlapping
// time windows, attention / MLP projections of the encoder blocks :177-178, cross K/V projections :205-210,
// every Linear of the decoder step :245-247,:228-230,:298 and the tied logits product :378-385).
//
// This file is the single-CTA kernel: the decoder-step GEMMs (M = micro-batch <= 128 rows, BLOCK_N = 32, one tile per CTA,
// 192 threads x 72 registers so that a CTA fits next to the resident cross-attention CTAs of the other micro-batch), the
// logits GEMM (BLOCK_N = 128, fused arg-max) and the comparator / fallback for the encoder, whose large GEMMs run on the
// CTA-pair kernel of gemm2cta_tcgen05.cu.
// Structure (persistent over output tiles, 2 + 4 warps for BLOCK_N < 128, 2 + 8 otherwise):
//   warp 0  (1 lane)  TMA producer: A tile 128x64 and W tile BLOCK_Nx64 per k-block, SWIZZLE_128B, kStages ring; the W
//                     halves of the first stages are issued before the programmatic-dependent-launch wait
//   warp 1  (1 lane)  MMA issuer: 4 x tcgen05.mma (M=128, N=BLOCK_N, K=16) per k-block into one of two TMEM
//                     accumulator stages; tcgen05.commit releases smem slots / publishes the accumulator
//   other warps       epilogue: tcgen05.ld 32 columns at a time (thread = output row; with 8 warps two warps per TMEM lane
//                     group split the columns), software-pipelined against the next chunk's TMEM read, bias staged in
//                     smem; fused bias / GELU / positional embedding / head-major scatter / arg-max with direct 16-byte
//                     stores, residual update as a vector reduction at L2 (red.global.add.v4.f32)
// Tile order is n-fastest so the CTAs in flight share A row-blocks and the whole W through L2.
#include <cfloat>
#include <mutex>

#include "common.cuh"
#include "gemm_common.cuh"
#include "kernels.h"

namespace b200w {

using namespace gemm_detail;

void gemm2cta_launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap* to, const CUtensorMap* to2, int epilogue, int n_batch,
                     int n_taps, int kb_per_tap, const int* a_c0, const int* a_row, const int* w_k0, const GemmParams& p, cudaStream_t stream);
void gemm2cta_set_attributes();

namespace {

constexpr int kSmemBudget = 196608;  // 192 KB of operand stages

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kStageBytesA = BLOCK_M * BLOCK_K * 2;
  static constexpr int kStageBytesB = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = kSmemBudget / kStageBytes > 8 ? 8 : kSmemBudget / kStageBytes;  // 256 -> 4, 128 -> 6, 64 -> 8, 32 -> 8
  static constexpr int kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static constexpr int kEpiWarps = BLOCK_N >= 128 ? 8 : 4;  // 8: two warps per TMEM lane group, each taking half of the columns
  // register cap: the BLOCK_N = 32 kernels of a decoder step must fit next to the resident cross-attention CTAs of the
  // other micro-batch (72 registers x 192 threads = 13.8 K: what one retiring cross-attention CTA frees on an SM plus the 4 K spare)
  static constexpr int kMaxRegs = BLOCK_N >= 128 ? 168 : (BLOCK_N == 32 ? 72 : 128);
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + 2 * BLOCK_N * 4 /*bias*/;
};

__device__ __forceinline__ TileCoord tile_coord(int t, const GemmGeom& g) {
  TileCoord c;
  c.n_blk = t % g.n_tiles;
  int mt = t / g.n_tiles;
  c.m_blk = mt % g.m_tiles_per_batch;
  c.batch = mt / g.m_tiles_per_batch;
  return c;
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(GemmCfg<BLOCK_N>::kThreads) __maxnreg__(GemmCfg<BLOCK_N>::kMaxRegs)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmGeom g,
                    const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ unsigned char smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], Cfg::kEpiWarps);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_base_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) may overlap the previous kernel (PDL), and so may
  // the first weight tiles: W is constant, only the activations depend on the predecessor.  The producer fills the W half
  // of the first pipeline stages now and adds the A half after the dependency wait (a decoder-step GEMM is latency-bound:
  // this takes the HBM round trip of the weights off the dependent chain).
  int pre_issued = 0;
  if (warp == 0 && lane == 0 && blockIdx.x < g.total_tiles) {
    const TileCoord c = tile_coord(blockIdx.x, g);
    const int n_pre = g.num_k_blocks < kStages ? g.num_k_blocks : kStages;
    for (int ks = 0; ks < n_pre; ++ks) {
      const int tap = ks / g.kb_per_tap, kb = ks - tap * g.kb_per_tap;
      mbar_arrive_expect_tx(&full_bar[ks], Cfg::kStageBytesB + g.a_stage_bytes);
      tma_load_2d(smem + ks * Cfg::kStageBytes + Cfg::kStageBytesA, &tmap_b, &full_bar[ks], g.w_k0[tap] + kb * BLOCK_K, c.n_blk * BLOCK_N);
    }
    pre_issued = n_pre;
  }
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
        const TileCoord c = tile_coord(t, g);
        for (int tap = 0; tap < g.n_taps; ++tap) {
          for (int kb = 0; kb < g.kb_per_tap; ++kb) {
            unsigned char* sa = smem + stage * Cfg::kStageBytes;
            unsigned char* sb = sa + Cfg::kStageBytesA;
            if (pre_issued > 0) {  // first stages of the first tile: the slot is fresh and its W half is already on its way
              --pre_issued;
            } else {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytesB + g.a_stage_bytes);
              tma_load_2d(sb, &tmap_b, &full_bar[stage], g.w_k0[tap] + kb * BLOCK_K, c.n_blk * BLOCK_N);
            }
            tma_load_3d(sa, &tmap_a, &full_bar[stage], g.a_c0[tap] + kb * BLOCK_K, c.m_blk * BLOCK_M + g.a_row[tap] + p.a_row_offset, c.batch + p.a_batch_offset);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc_stage = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc_stage], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc_stage * BLOCK_N;
        for (int kb = 0; kb < g.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t da = umma_desc_kmajor_sw128(sa);
          const uint64_t db = umma_desc_kmajor_sw128(sa + Cfg::kStageBytesA);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advancing 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the 16-byte address field
            umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[acc_stage]);  // accumulator complete
        if (++acc_stage == 2) {
          acc_stage = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===== epilogue warps: TMEM lane group = warp % 4; with 8 warps each lane group's columns are split in two halves =====
    constexpr int kEpiWarps = Cfg::kEpiWarps;
    float* s_bias = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes + 256);  // [2][BLOCK_N]
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
      const TileCoord c = tile_coord(t, g);
      epilogue_tile<BLOCK_N, EPI, kEpiWarps>(p, c, tmem_base + acc_stage * BLOCK_N, s_bias + acc_stage * BLOCK_N, &tmem_full_bar[acc_stage],
                                             acc_phase, warp - 2, warp & 3, lane);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc_stage]);
      if (++acc_stage == 2) {
        acc_stage = 0;
        acc_phase ^= 1;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || p == nullptr) throw CudaError("cuTensorMapEncodeTiled not available from the driver");
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

}  // namespace

CUtensorMap make_tmap_sw128(const void* base, bool f32, int rank, const uint64_t* dims, const uint64_t* pitches_bytes, const uint32_t* box);
CUtensorMap make_tmap_bf16_sw128(const void* base, int rank, const uint64_t* dims, const uint64_t* pitches_bytes, const uint32_t* box) {
  return make_tmap_sw128(base, false, rank, dims, pitches_bytes, box);
}
CUtensorMap make_tmap_sw128(const void* base, bool f32, int rank, const uint64_t* dims, const uint64_t* pitches_bytes, const uint32_t* box) {
  CUtensorMap m;
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t bdim[3], estr[3];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = pitches_bytes[i];
  CUresult r = get_encode_fn()(&m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu)", (int)r, rank,
             (unsigned long long)dims[0], (unsigned long long)dims[1]);
    throw CudaError(buf);
  }
  return m;
}

namespace {
inline CUtensorMap make_tmap(const void* base, int rank, const uint64_t* dims, const uint64_t* pitches_bytes, const uint32_t* box) {
  return make_tmap_bf16_sw128(base, rank, dims, pitches_bytes, box);
}

template <int BLOCK_N, int EPI>
void launch_one(const CUtensorMap& ta, const CUtensorMap& tb, const GemmGeom& g, const GemmParams& p, int grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  launch_k(p.use_pdl != 0, gemm_tcgen05_kernel<BLOCK_N, EPI>, dim3(grid), dim3(Cfg::kThreads), (size_t)Cfg::kSmemBytes, stream, ta, tb, g, p);
}

template <int EPI>
void launch_bn(int block_n, const CUtensorMap& ta, const CUtensorMap& tb, const GemmGeom& g, const GemmParams& p, int grid,
               cudaStream_t stream) {
  switch (block_n) {
    case 256: launch_one<256, EPI>(ta, tb, g, p, grid, stream); break;
    case 128: launch_one<128, EPI>(ta, tb, g, p, grid, stream); break;
    case 64: launch_one<64, EPI>(ta, tb, g, p, grid, stream); break;
    case 32: launch_one<32, EPI>(ta, tb, g, p, grid, stream); break;
    default: throw CudaError("gemm: unsupported BLOCK_N");
  }
}

template <int BLOCK_N, int EPI>
void set_attr_one() {
  CUDA_CHECK(cudaFuncSetAttribute(gemm_tcgen05_kernel<BLOCK_N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BLOCK_N>::kSmemBytes));
}
template <int EPI>
void set_attr_epi() {
  set_attr_one<256, EPI>();
  set_attr_one<128, EPI>();
  set_attr_one<64, EPI>();
  set_attr_one<32, EPI>();
}

}  // namespace

void gemm_set_attributes() {
  gemm2cta_set_attributes();
  set_attr_epi<EPI_BIAS_BF16>();
  set_attr_epi<EPI_BIAS_GELU_BF16>();
  set_attr_epi<EPI_BIAS_F32>();
  set_attr_epi<EPI_BIAS_RESID_F32>();
  set_attr_epi<EPI_GELU_POS_F32>();
  set_attr_epi<EPI_CROSSKV_BF16>();
  set_attr_epi<EPI_ARGMAX>();
}

struct GemmPlan {
  CUtensorMap tmap_a, tmap_b;
  CUtensorMap tmap_a_thin[2];  // A boxes of 32 and 64 rows (1-CTA kernel, single row block of <= 32 / 64 valid rows)
  GemmGeom geom;
  int block_n, epilogue, grid;
  bool two_cta;
  bool has_out;            // two_cta only: outputs leave through TMA stores / reduce-adds
  CUtensorMap tmap_out, tmap_out2;
};



GemmPlan* gemm_plan_create(const GemmOperandA& a, const __nv_bfloat16* w, int n_rows_w, int block_n, int epilogue, bool two_cta,
                           const GemmTmaOut* out) {
  if ((a.row_pitch * 2) % 16 != 0 || (a.batch_pitch * 2) % 16 != 0) throw CudaError("gemm: operand pitches must be multiples of 16 bytes");
  GemmPlan* pl = new GemmPlan();
  const uint64_t adims[3] = {(uint64_t)a.K, (uint64_t)a.rows, (uint64_t)a.n_batch};
  const uint64_t apitch[2] = {(uint64_t)a.row_pitch * 2, (uint64_t)(a.n_batch > 1 ? a.batch_pitch : (long)a.rows * a.row_pitch) * 2};
  const uint32_t abox[3] = {BLOCK_K, BLOCK_M, 1};
  pl->tmap_a = make_tmap(a.ptr, 3, adims, apitch, abox);
  for (int i = 0; i < 2; ++i) {
    const uint32_t tbox[3] = {BLOCK_K, i == 0 ? 32u : 64u, 1};
    pl->tmap_a_thin[i] = (uint64_t)a.rows >= tbox[1] ? make_tmap(a.ptr, 3, adims, apitch, tbox) : pl->tmap_a;
  }
  const int n_taps = a.n_taps > 0 ? a.n_taps : 1;
  const int k_per_tap = a.n_taps > 0 ? a.k_per_tap : a.K;
  const long w_k = (long)n_taps * k_per_tap;  // W is [n_rows_w][n_taps * k_per_tap], row-major
  if ((w_k * 2) % 16 != 0) throw CudaError("gemm: weight row pitch must be a multiple of 16 bytes");
  const uint64_t bdims[2] = {(uint64_t)w_k, (uint64_t)n_rows_w};
  const uint64_t bpitch[1] = {(uint64_t)w_k * 2};
  if (two_cta) block_n = 256;  // pair tile 256 x 256; each CTA of the pair stages 128 rows of W
  const uint32_t bbox[2] = {BLOCK_K, (uint32_t)(two_cta ? 128 : block_n)};
  pl->tmap_b = make_tmap(w, 2, bdims, bpitch, bbox);
  pl->block_n = block_n;
  pl->epilogue = epilogue;
  pl->two_cta = two_cta;
  pl->has_out = false;
  if (two_cta && out != nullptr) {
    pl->has_out = true;
    if (epilogue == EPI_CROSSKV_BF16) {
      // cross K / V [L*B*H][T][64] bf16: one 64-column box = one head; rows past T are clipped by TMA
      const uint64_t od[3] = {64, (uint64_t)out->rows, (uint64_t)out->n_slices};
      const uint64_t op[2] = {128, (uint64_t)out->rows * 128};
      const uint32_t ob[3] = {64, 128, 1};
      pl->tmap_out = make_tmap_sw128(out->ptr, false, 3, od, op, ob);
      pl->tmap_out2 = make_tmap_sw128(out->ptr2, false, 3, od, op, ob);
    } else {
      const bool f32 = epilogue == EPI_BIAS_RESID_F32;
      const uint64_t od[2] = {(uint64_t)out->cols, (uint64_t)out->rows};
      const uint64_t op[1] = {(uint64_t)out->ld * (f32 ? 4 : 2)};
      const uint32_t ob[2] = {f32 ? 32u : 64u, 128};
      pl->tmap_out = make_tmap_sw128(out->ptr, f32, 2, od, op, ob);
      pl->tmap_out2 = pl->tmap_out;
    }
  }
  pl->geom.n_batch = a.n_batch;
  pl->geom.n_taps = n_taps;
  pl->geom.kb_per_tap = (k_per_tap + BLOCK_K - 1) / BLOCK_K;
  pl->geom.num_k_blocks = n_taps * pl->geom.kb_per_tap;
  for (int t = 0; t < 3; ++t) {
    pl->geom.a_c0[t] = a.n_taps > 0 ? a.tap_c0[t] : 0;
    pl->geom.a_row[t] = a.n_taps > 0 ? a.tap_row[t] : 0;
    pl->geom.w_k0[t] = t * k_per_tap;
  }
  pl->geom.m_tiles_per_batch = 0;  // set at launch (depends on rows_valid)
  pl->geom.n_tiles = 0;
  return pl;
}

void gemm_plan_destroy(GemmPlan* p) { delete p; }

void gemm_launch(const GemmPlan* plan, const GemmParams& p, cudaStream_t stream) {
  if (plan->two_cta) {
    gemm2cta_launch(plan->tmap_a, plan->tmap_b, plan->has_out ? &plan->tmap_out : nullptr, plan->has_out ? &plan->tmap_out2 : nullptr,
                    plan->epilogue, plan->geom.n_batch, plan->geom.n_taps, plan->geom.kb_per_tap, plan->geom.a_c0, plan->geom.a_row,
                    plan->geom.w_k0, p, stream);
    return;
  }
  GemmGeom g = plan->geom;
  g.m_tiles_per_batch = (p.rows_valid + BLOCK_M - 1) / BLOCK_M;
  g.n_tiles = (p.N + plan->block_n - 1) / plan->block_n;
  if (p.n_batch > 0) g.n_batch = p.n_batch;
  g.total_tiles = g.n_batch * g.m_tiles_per_batch * g.n_tiles;
  if (g.total_tiles <= 0) return;
  const int grid = g.total_tiles < kNumSMs ? g.total_tiles : kNumSMs;
  // thin A operand for the decoder-step shapes (one row block with few valid rows): a 32- or 64-row TMA box
  static const bool thin_ok = getenv("B200W_NO_THIN_A") == nullptr;
  const CUtensorMap* ta = &plan->tmap_a;
  g.a_stage_bytes = BLOCK_M * BLOCK_K * 2;
  if (thin_ok && g.m_tiles_per_batch == 1 && g.n_batch == 1 && g.n_taps == 1 && p.rows_valid <= 64) {
    const int rows = p.rows_valid <= 32 ? 32 : 64;
    ta = &plan->tmap_a_thin[rows == 32 ? 0 : 1];
    g.a_stage_bytes = rows * BLOCK_K * 2;
  }
  switch (plan->epilogue) {
    case EPI_BIAS_BF16: launch_bn<EPI_BIAS_BF16>(plan->block_n, *ta, plan->tmap_b, g, p, grid, stream); break;
    case EPI_BIAS_GELU_BF16: launch_bn<EPI_BIAS_GELU_BF16>(plan->block_n, *ta, plan->tmap_b, g, p, grid, stream); break;
    case EPI_BIAS_F32: launch_bn<EPI_BIAS_F32>(plan->block_n, *ta, plan->tmap_b, g, p, grid, stream); break;
    case EPI_BIAS_RESID_F32: launch_bn<EPI_BIAS_RESID_F32>(plan->block_n, *ta, plan->tmap_b, g, p, grid, stream); break;
    case EPI_GELU_POS_F32: launch_bn<EPI_GELU_POS_F32>(plan->block_n, *ta, plan->tmap_b, g, p, grid, stream); break;
    case EPI_CROSSKV_BF16: launch_bn<EPI_CROSSKV_BF16>(plan->block_n, *ta, plan->tmap_b, g, p, grid, stream); break;
    case EPI_ARGMAX: launch_bn<EPI_ARGMAX>(plan->block_n, *ta, plan->tmap_b, g, p, grid, stream); break;
    default: throw CudaError("gemm: unknown epilogue");
  }
}

// ---- SIMT comparator (self-tests only) ------------------------------------------------------------------
namespace {
__global__ void gemm_reference_simt_kernel(const __nv_bfloat16* __restrict__ a, long lda, const __nv_bfloat16* __restrict__ w, long ldw,
                                           const float* __restrict__ bias, float* __restrict__ out, long ldo, int M, int N, int K) {
  __shared__ float sa[16][17], sw[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    const int am = blockIdx.y * 16 + ty, wn = blockIdx.x * 16 + ty;
    sa[ty][tx] = (am < M && k0 + tx < K) ? __bfloat162float(a[am * lda + k0 + tx]) : 0.f;
    sw[ty][tx] = (wn < N && k0 + tx < K) ? __bfloat162float(w[wn * ldw + k0 + tx]) : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(sa[ty][k], sw[tx][k], acc);
    __syncthreads();
  }
  if (m < M && n < N) out[m * ldo + n] = acc + (bias ? bias[n] : 0.f);
}
}  // namespace

void gemm_reference_simt(const __nv_bfloat16* a, long lda, const __nv_bfloat16* w, long ldw, const float* bias, float* out, long ldo,
                         int M, int N, int K, cudaStream_t stream) {
  dim3 grid((N + 15) / 16, (M + 15) / 16), block(16, 16);
  gemm_reference_simt_kernel<<<grid, block, 0, stream>>>(a, lda, w, ldw, bias, out, ldo, M, N, K);
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace b200w
