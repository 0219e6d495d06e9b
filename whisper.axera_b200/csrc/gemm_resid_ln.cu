// Decoder-step GEMM with the residual update AND the LayerNorm that follows it in one kernel:
//
//     x[m][:] += A[m][:] . W^T + bias          (out-projection / cross out-projection / fc2 of a decoder block,
//     h[m][:]  = LayerNorm(x[m][:]) * g + b     /root/reference/model_convert/export_onnx.py:286-298; whisper/model.py LayerNorm eps 1e-5)
//
// Why: a decoder step on a mid-size batch is a chain of ~11 short dependent kernels per layer, and the chain -- not HBM -- sets
// the step time (profiles/r02_decode_timeline_*.txt: ~3.6 us per kernel boundary).  Three of the eleven are LayerNorms that
// only re-read what the preceding GEMM has just written.  Here the GEMM keeps the updated rows on chip and normalises them
// itself: 11 -> 8 kernels per layer.
//
// A LayerNorm needs whole rows, a latency-bound GEMM wants its N columns spread over many CTAs.  Both: the N tiles of one
// 128-row block form ONE thread-block cluster (N / 64 CTAs, <= 16); each CTA computes its 128 x 64 tile with tcgen05.mma (TMA
// operands, accumulator in TMEM), adds residual and bias, keeps the fp32 tile in shared memory, and the per-row sums travel
// through distributed shared memory: two exchanges (mean, then centred sum of squares: the same two-pass statistics as
// layernorm_kernel), cluster-rank order, so the result does not depend on timing or batch size.
//
// Structure (192 threads, one tile per CTA): warp 0 = TMA producer (W tiles requested before the PDL dependency wait),
// warp 1 = MMA issuer, warps 2-5 = epilogue (thread = row).  <= 72 registers so that a CTA still fits next to the resident
// cross-attention stream CTAs of the other micro-batch.
#include <cfloat>

#include "common.cuh"
#include "gemm_common.cuh"
#include "kernels.h"

namespace b200w {

using namespace gemm_detail;

namespace {

constexpr int kBN = 64;
constexpr int kStages = 6;
constexpr int kStageBytesA = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int kStageBytesB = kBN * BLOCK_K * 2;      // 8 KB
constexpr int kStageBytes = kStageBytesA + kStageBytesB;
constexpr int kTileLd = kBN + 4;                     // fp32 tile row pitch: 16-byte aligned rows, conflict-free float4 access by row
constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;
// stages | barriers (256) | bias, gamma, beta [64] each | row statistics [2][128] | fp32 tile [128][68]
constexpr int kOffBars = kStages * kStageBytes;
constexpr int kOffVec = kOffBars + 256;
constexpr int kOffStat = kOffVec + 3 * kBN * 4;
constexpr int kOffTile = kOffStat + 2 * BLOCK_M * 4;
constexpr int kSmemBytes = kOffTile + BLOCK_M * kTileLd * 4 + 1024 /*align slack*/;

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}

__global__ void __maxnreg__(72)
gemm_resid_ln_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const int num_k_blocks,
                     const int n_tiles, const GemmParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_bias = reinterpret_cast<float*>(smem + kOffVec);
  float* s_gamma = s_bias + kBN;
  float* s_beta = s_gamma + kBN;
  float* s_stat = reinterpret_cast<float*>(smem + kOffStat);  // [2][128]
  float* s_tile = reinterpret_cast<float*>(smem + kOffTile);  // [128][kTileLd]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blk = blockIdx.x % n_tiles;  // == cluster rank: the cluster spans the N tiles of one row block
  const int m_blk = blockIdx.x / n_tiles;
  const int n0 = n_blk * kBN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_base_slot, kBN);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  // W is constant: its first pipeline stages are requested before the dependency wait (the A halves follow after it)
  int pre_issued = 0;
  if (warp == 0 && lane == 0) {
    const int n_pre = num_k_blocks < kStages ? num_k_blocks : kStages;
    for (int ks = 0; ks < n_pre; ++ks) {
      mbar_arrive_expect_tx(&full_bar[ks], kStageBytes);
      tma_load_2d(smem + ks * kStageBytes + kStageBytesA, &tmap_b, &full_bar[ks], ks * BLOCK_K, n0);
    }
    pre_issued = n_pre;
  }
  if (warp >= 2) {  // LayerNorm parameters and bias do not depend on the predecessor either
    const int et = threadIdx.x - 64;
    if (et < kBN) {
      s_bias[et] = p.bias != nullptr ? p.bias[n0 + et] : 0.f;
      s_gamma[et] = p.ln_gamma[n0 + et];
      s_beta[et] = p.ln_beta[n0 + et];
    }
  }
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        unsigned char* sa = smem + stage * kStageBytes;
        if (pre_issued > 0) {
          --pre_issued;
        } else {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
          tma_load_2d(sa + kStageBytesA, &tmap_b, &full_bar[stage], kb * BLOCK_K, n0);
        }
        tma_load_3d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M + p.a_row_offset, p.a_batch_offset);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, kBN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + stage * kStageBytes);
        const uint64_t da = umma_desc_kmajor_sw128(sa);
        const uint64_t db = umma_desc_kmajor_sw128(sa + kStageBytesA);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma_bf16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tmem_full_bar);
    }
    __syncwarp();
  } else {
    // ===== epilogue warps 2..5: TMEM lane group = warp % 4, thread = row =====
    const int et = threadIdx.x - 64;  // 0..127
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    float* xg = reinterpret_cast<float*>(p.out);
    // the residual tile x[128][64] -> shared memory while the MMAs run (coalesced: 16 consecutive threads cover one 256-byte row segment)
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
      const int i = et + kEpiThreads * k;
      const int rr = i >> 4, piece = i & 15;
      const int gr = m_blk * BLOCK_M + rr;
      cp_async16_zfill(s_tile + rr * kTileLd + piece * 4, xg + (long)(gr + p.out_row_offset) * p.ldo + n0 + piece * 4, gr < p.rows_valid);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // tile + bias / gamma / beta visible to all epilogue threads
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    // pass 1: x_new = x_old + acc + bias (kept in shared memory), row sum of this CTA's 64 columns
    float* trow = s_tile + row * kTileLd;
    float sum = 0.f;
    const uint32_t tbase = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
#pragma unroll
    for (int ch = 0; ch < kBN / 32; ++ch) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tbase + ch * 32, v);
      tcgen05_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 x = *reinterpret_cast<float4*>(trow + ch * 32 + j);
        const float4 b = *reinterpret_cast<const float4*>(s_bias + ch * 32 + j);
        x.x += __uint_as_float(v[j]) + b.x, x.y += __uint_as_float(v[j + 1]) + b.y;
        x.z += __uint_as_float(v[j + 2]) + b.z, x.w += __uint_as_float(v[j + 3]) + b.w;
        *reinterpret_cast<float4*>(trow + ch * 32 + j) = x;
        sum += (x.x + x.y) + (x.z + x.w);
      }
    }
    s_stat[row] = sum;
  }
  // ---- exchange 1: row means over the whole cluster (every thread of every CTA takes part in a cluster barrier) ----
  cluster_sync_all();
  float mean = 0.f, sq = 0.f;
  const int row = (warp & 3) * 32 + lane;
  float* trow = s_tile + row * kTileLd;
  if (warp >= 2) {
    for (int rk = 0; rk < n_tiles; ++rk) mean += dsmem_ld_f32(dsmem_addr(&s_stat[row], rk));  // rank order: deterministic
    mean /= (float)p.N;
#pragma unroll
    for (int j = 0; j < kBN; j += 4) {
      const float4 x = *reinterpret_cast<const float4*>(trow + j);
      const float a = x.x - mean, b = x.y - mean, c = x.z - mean, e = x.w - mean;
      sq += (a * a + b * b) + (c * c + e * e);
    }
    s_stat[BLOCK_M + row] = sq;
    // the updated residual rows go back to global memory in the meantime (coalesced, from shared memory; every epilogue
    // thread finished pass 1 before the cluster barrier above)
    const int et = threadIdx.x - 64;
    float* xg = reinterpret_cast<float*>(p.out);
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
      const int i = et + kEpiThreads * k;
      const int rr = i >> 4, piece = i & 15;
      const int gr = m_blk * BLOCK_M + rr;
      if (gr < p.rows_valid)
        *reinterpret_cast<float4*>(xg + (long)(gr + p.out_row_offset) * p.ldo + n0 + piece * 4) = *reinterpret_cast<const float4*>(s_tile + rr * kTileLd + piece * 4);
    }
  }
  // ---- exchange 2: centred sums of squares ----
  cluster_sync_all();
  if (warp >= 2) {
    float var = 0.f;
    for (int rk = 0; rk < n_tiles; ++rk) var += dsmem_ld_f32(dsmem_addr(&s_stat[BLOCK_M + row], rk));
    const float rstd = rsqrtf(var / (float)p.N + 1e-5f);
    const int r = m_blk * BLOCK_M + row;
    if (r < p.rows_valid) {
      __nv_bfloat16* ho = p.ln_out + (long)(r + p.out_row_offset) * p.ln_ldo + n0;
#pragma unroll
      for (int j = 0; j < kBN; j += 8) {
        const float4 x0 = *reinterpret_cast<const float4*>(trow + j), x1 = *reinterpret_cast<const float4*>(trow + j + 4);
        const float4 g0 = *reinterpret_cast<const float4*>(s_gamma + j), g1 = *reinterpret_cast<const float4*>(s_gamma + j + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(s_beta + j), b1 = *reinterpret_cast<const float4*>(s_beta + j + 4);
        uint4 o;
        o.x = pack_bf16x2((x0.x - mean) * rstd * g0.x + b0.x, (x0.y - mean) * rstd * g0.y + b0.y);
        o.y = pack_bf16x2((x0.z - mean) * rstd * g0.z + b0.z, (x0.w - mean) * rstd * g0.w + b0.w);
        o.z = pack_bf16x2((x1.x - mean) * rstd * g1.x + b1.x, (x1.y - mean) * rstd * g1.y + b1.y);
        o.w = pack_bf16x2((x1.z - mean) * rstd * g1.z + b1.z, (x1.w - mean) * rstd * g1.w + b1.w);
        *reinterpret_cast<uint4*>(ho + j) = o;
      }
    }
  }
  // nobody leaves (and releases its shared memory) before every peer has read the statistics
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kBN);
  }
}

}  // namespace

void gemm_resid_ln_set_attributes() {
  CUDA_CHECK(cudaFuncSetAttribute(gemm_resid_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  CUDA_CHECK(cudaFuncSetAttribute(gemm_resid_ln_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
}

bool gemm_resid_ln_supported(int d) {
  static const bool off = getenv("B200W_NO_FUSED_LN") != nullptr;
  if (off || d % kBN != 0 || d / kBN > 16 || d / kBN < 1) return false;
  // can the device co-schedule a cluster of d / 64 such CTAs?
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(d / kBN), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = kSmemBytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)(d / kBN), attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  int n = 0;
  gemm_resid_ln_set_attributes();
  if (cudaOccupancyMaxActiveClusters(&n, gemm_resid_ln_kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return n >= 1;
}

// tmap_b must have been built with a 64-row box (gemm_plan_create(..., block_n = 64, EPI_RESID_LN_F32))
void gemm_resid_ln_launch(const CUtensorMap& ta, const CUtensorMap& tb, int num_k_blocks, const GemmParams& p, cudaStream_t stream) {
  if (p.N % kBN != 0 || p.N / kBN > 16) throw CudaError("gemm_resid_ln: N must be a multiple of 64 and at most 1024");
  if (p.ln_gamma == nullptr || p.ln_beta == nullptr || p.ln_out == nullptr || p.out == nullptr) throw CudaError("gemm_resid_ln: missing LayerNorm operands");
  const int n_tiles = p.N / kBN;
  const int m_tiles = (p.rows_valid + BLOCK_M - 1) / BLOCK_M;
  if (m_tiles <= 0) return;
  launch_kc(p.use_pdl != 0, n_tiles, gemm_resid_ln_kernel, dim3(n_tiles * m_tiles), dim3(kThreads), (size_t)kSmemBytes, stream, ta, tb, num_k_blocks,
            n_tiles, p);
}

}  // namespace b200w
