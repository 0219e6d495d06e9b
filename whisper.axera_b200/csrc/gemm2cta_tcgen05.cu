// K4 (large-M variant): the same bf16 GEMM as gemm_tcgen05.cu, issued as tcgen05.mma.cta_group::2 by a CTA PAIR.
//
// Why: with 128 x 256 tiles per CTA every k-block needs 48 KB of operands per SM (A 16 KB + W 32 KB); on B200 the
// L2 -> shared-memory feed (~75 GB/s per SM with ~190 KB in flight) then caps the tensor pipe at ~40 % (measured, see
// profiles/).  A CTA pair computes a 256 x 256 tile: each CTA stages only its 128 rows of A and HALF of the W tile
// (32 KB per k-block per SM, one third less), the pair's tensor cores read both halves of W across the two SMs, and the
// smaller stage gives a 6-deep ring instead of 4.
//
//   cluster (2,1,1), persistent over 256 x 256 tiles, 320 threads per CTA
//   warp 0 (1 lane, both CTAs)   TMA producer: its A rows + its W half; completion bytes go to the LEADER's mbarrier
//   warp 1 (1 lane, leader CTA)  tcgen05.mma.cta_group::2 (M 256, N 256, K 16); tcgen05.commit multicast frees the smem
//                                slot in both CTAs and publishes the accumulator to both epilogues
//   warps 2-9 (both CTAs)        epilogue of the CTA's own 128 accumulator rows (shared code: gemm_common.cuh); arrives on
//                                the leader's accumulator-free barrier through the cluster address space
#include "common.cuh"
#include "gemm_common.cuh"
#include "kernels.h"

namespace b200w {

using namespace gemm_detail;

namespace {

constexpr int kPairN = 256;                       // N of the pair tile
constexpr int kStageBytesA2 = BLOCK_M * BLOCK_K * 2;        // 16 KB: this CTA's 128 rows of A
constexpr int kStageBytesB2 = (kPairN / 2) * BLOCK_K * 2;   // 16 KB: this CTA's half of the W tile
constexpr int kStageBytes2 = kStageBytesA2 + kStageBytesB2;
constexpr int kStages2 = 6;
constexpr int kEpiWarps2 = 8;
constexpr int kThreads2 = 64 + 32 * kEpiWarps2;
constexpr int kSmemBytes2 = kStages2 * kStageBytes2 + 1024 + 256 + 2 * kPairN * 4;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;    // clears the CTA-rank bit of a shared::cluster address -> leader CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* smem_dst, const void* desc, uint64_t* leader_bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* smem_dst, const void* desc, uint64_t* leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all earlier MMAs of this thread are complete) on the barrier at the same smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_multicast(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(0));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
gemm2cta_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmGeom g,
                        const GemmParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages2 * kStageBytes2);
  uint64_t* empty_bar = full_bar + kStages2;
  uint64_t* tmem_full_bar = empty_bar + kStages2;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(smem + kStages2 * kStageBytes2 + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < kStages2; ++i) {
      mbar_init(&full_bar[i], 1);   // the leader's arrive.expect_tx; both CTAs' TMA bytes complete on the leader's barrier
      mbar_init(&empty_bar[i], 1);  // one multicast commit from the leader's MMA thread
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 2 * kEpiWarps2);  // epilogue warps of both CTAs (only the leader's copy is used)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / multicast commit / TMA completion
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  auto tile_of = [&](int t) {
    TileCoord c;
    c.n_blk = t % g.n_tiles;
    const int mp = t / g.n_tiles;
    c.m_blk = (mp % g.m_tiles_per_batch) * 2 + (int)rank;  // m_tiles_per_batch counts PAIRS of 128-row tiles here
    c.batch = mp / g.m_tiles_per_batch;
    return c;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < g.total_tiles; t += n_clusters) {
        const TileCoord c = tile_of(t);
        for (int tap = 0; tap < g.n_taps; ++tap) {
          for (int kb = 0; kb < g.kb_per_tap; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            unsigned char* sa = smem + stage * kStageBytes2;
            unsigned char* sb = sa + kStageBytesA2;
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes2);
            tma2_load_3d(sa, &tmap_a, &full_bar[stage], g.a_c0[tap] + kb * BLOCK_K, c.m_blk * BLOCK_M + g.a_row[tap] + p.a_row_offset,
                         c.batch + p.a_batch_offset);
            tma2_load_2d(sb, &tmap_b, &full_bar[stage], g.w_k0[tap] + kb * BLOCK_K, c.n_blk * kPairN + (int)rank * (kPairN / 2));
            if (++stage == kStages2) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BLOCK_M, kPairN);
      int stage = 0;
      uint32_t phase = 0;
      int acc_stage = 0;
      uint32_t acc_phase = 0;
      for (int t = cluster_id; t < g.total_tiles; t += n_clusters) {
        mbar_wait(&tmem_empty_bar[acc_stage], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc_stage * kPairN;
        for (int kb = 0; kb < g.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes2);
          const uint64_t da = umma_desc_kmajor_sw128(sa);
          const uint64_t db = umma_desc_kmajor_sw128(sa + kStageBytesA2);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma2_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma2_commit_multicast(&empty_bar[stage]);
          if (++stage == kStages2) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma2_commit_multicast(&tmem_full_bar[acc_stage]);
        if (++acc_stage == 2) {
          acc_stage = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    for (int t = cluster_id; t < g.total_tiles; t += n_clusters) {
      const TileCoord c = tile_of(t);
      epilogue_tile<kPairN, EPI, kEpiWarps2>(p, c, tmem_base + acc_stage * kPairN, s_bias + acc_stage * kPairN, &tmem_full_bar[acc_stage], acc_phase,
                                             warp - 2, warp & 3, lane);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[acc_stage]);
      if (++acc_stage == 2) {
        acc_stage = 0;
        acc_phase ^= 1;
      }
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();  // the peer's smem / TMEM stay alive until both CTAs are done
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <int EPI>
void launch2(const CUtensorMap& ta, const CUtensorMap& tb, const GemmGeom& g, const GemmParams& p, int grid, cudaStream_t stream) {
  gemm2cta_tcgen05_kernel<EPI><<<grid, kThreads2, kSmemBytes2, stream>>>(ta, tb, g, p);
  CUDA_CHECK(cudaGetLastError());
}
template <int EPI>
void set_attr2() {
  CUDA_CHECK(cudaFuncSetAttribute(gemm2cta_tcgen05_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes2));
}

}  // namespace

void gemm2cta_set_attributes() {
  set_attr2<EPI_BIAS_BF16>();
  set_attr2<EPI_BIAS_GELU_BF16>();
  set_attr2<EPI_BIAS_F32>();
  set_attr2<EPI_BIAS_RESID_F32>();
  set_attr2<EPI_GELU_POS_F32>();
  set_attr2<EPI_CROSSKV_BF16>();
}

// geometry / launch for a plan created with two_cta = true (see gemm_plan_create)
void gemm2cta_launch(const CUtensorMap& ta, const CUtensorMap& tb, int epilogue, int n_batch, int n_taps, int kb_per_tap, const int* a_c0,
                     const int* a_row, const int* w_k0, const GemmParams& p, cudaStream_t stream) {
  GemmGeom g{};
  g.n_batch = p.n_batch > 0 ? p.n_batch : n_batch;
  g.n_taps = n_taps, g.kb_per_tap = kb_per_tap, g.num_k_blocks = n_taps * kb_per_tap;
  for (int t = 0; t < 3; ++t) g.a_c0[t] = a_c0[t], g.a_row[t] = a_row[t], g.w_k0[t] = w_k0[t];
  const int m_tiles = (p.rows_valid + BLOCK_M - 1) / BLOCK_M;
  g.m_tiles_per_batch = (m_tiles + 1) / 2;  // pairs
  g.n_tiles = (p.N + kPairN - 1) / kPairN;
  g.total_tiles = g.n_batch * g.m_tiles_per_batch * g.n_tiles;
  if (g.total_tiles <= 0) return;
  const int n_clusters = g.total_tiles < kNumSMs / 2 ? g.total_tiles : kNumSMs / 2;
  const int grid = 2 * n_clusters;
  switch (epilogue) {
    case EPI_BIAS_BF16: launch2<EPI_BIAS_BF16>(ta, tb, g, p, grid, stream); break;
    case EPI_BIAS_GELU_BF16: launch2<EPI_BIAS_GELU_BF16>(ta, tb, g, p, grid, stream); break;
    case EPI_BIAS_F32: launch2<EPI_BIAS_F32>(ta, tb, g, p, grid, stream); break;
    case EPI_BIAS_RESID_F32: launch2<EPI_BIAS_RESID_F32>(ta, tb, g, p, grid, stream); break;
    case EPI_GELU_POS_F32: launch2<EPI_GELU_POS_F32>(ta, tb, g, p, grid, stream); break;
    case EPI_CROSSKV_BF16: launch2<EPI_CROSSKV_BF16>(ta, tb, g, p, grid, stream); break;
    default: throw CudaError("gemm2cta: unsupported epilogue");
  }
}

}  // namespace b200w
