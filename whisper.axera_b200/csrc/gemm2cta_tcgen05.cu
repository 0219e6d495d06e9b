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
constexpr int kStagingBytes = 2 * 16384;  // output staging for TMA stores: 2 boxes of 128 rows x 128 bytes
constexpr int kSmemBytes2 = kStages2 * kStageBytes2 + kStagingBytes + 256 + kPairN * 4 + 1024;
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

__device__ __forceinline__ void tma_store_2d(const void* desc, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(desc)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* desc, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(desc)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const void* desc, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(desc)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void epi_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kEpiWarps2 * 32) : "memory"); }

// Epilogue with coalesced output: every thread owns one accumulator row, so direct stores touch 32 different rows per
// warp instruction (measured: the LSU, not the tensor pipe, set the tile time).  Here the 8 epilogue warps stage one
// 128-row box at a time in shared memory (SWIZZLE_128B, conflict-free) and one thread hands it to the TMA unit:
//   bf16 outputs (bias / bias+GELU / cross-K/V): 4 boxes of 64 columns per tile, double-buffered;
//   fp32 residual stream: x += acc + bias as a TMA REDUCE-ADD (cp.reduce.async.bulk.tensor .add) -- the residual is never
//   read into the SM; 4 steps of two 32-column boxes per tile.
// warp (lane group lg, half) handles chunk 2*i + half of the tile's eight 32-column chunks at step i.
template <int EPI>
__device__ __forceinline__ void epilogue_tile_tma(const GemmParams& p, const CUtensorMap* tmap_out, const CUtensorMap* tmap_out2, const TileCoord& c,
                                                  uint32_t tmem_acc, float* sb, unsigned char* staging, uint64_t* tmem_full_bar, uint32_t acc_phase,
                                                  uint64_t* tmem_empty_bar_leader, int epi_warp, int lg, int lane) {
  constexpr bool kF32 = (EPI == EPI_BIAS_RESID_F32);
  const int half = epi_warp >> 2;
  const int etid = epi_warp * 32 + lane;
  const int row = lg * 32 + lane;
  const bool issuer = etid == 0;
  for (int i = etid; i < kPairN; i += kEpiWarps2 * 32) {
    const int n = c.n_blk * kPairN + i;
    sb[i] = (p.bias != nullptr && n < p.N) ? p.bias[n] : 0.f;
  }
  epi_bar(1);
  mbar_wait(tmem_full_bar, acc_phase);
  tcgen05_fence_after();
  const uint32_t tbase = tmem_acc + (static_cast<uint32_t>(lg * 32) << 16);
  uint32_t v[2][32];
  tmem_ld_32x32b_x32(tbase + half * 32, v[0]);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ch = 2 * i + half;
    tcgen05_wait_ld();
    if (i + 1 < 4) {
      tmem_ld_32x32b_x32(tbase + (ch + 2) * 32, v[(i + 1) & 1]);
    } else {
      // all TMEM reads of this tile are done: hand the accumulator stage back to the MMA issuer before the stores
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tmem_empty_bar_leader);
    }
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bv = *reinterpret_cast<const float4*>(sb + ch * 32 + j);
      f[j] = __uint_as_float(v[i & 1][j]) + bv.x, f[j + 1] = __uint_as_float(v[i & 1][j + 1]) + bv.y;
      f[j + 2] = __uint_as_float(v[i & 1][j + 2]) + bv.z, f[j + 3] = __uint_as_float(v[i & 1][j + 3]) + bv.w;
    }
    if constexpr (EPI == EPI_BIAS_GELU_BF16) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = gelu_fast(f[j]);
    }
    epi_bar(2);  // the staging buffer written below is no longer being read by an earlier TMA store (issuer waited)
    if constexpr (kF32) {
      unsigned char* dst = staging + half * 16384 + row * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(dst + ((q ^ (row & 7)) << 4)) = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
    } else {
      unsigned char* dst = staging + (i & 1) * 16384 + row * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u;
        u.x = pack_bf16x2(f[8 * q], f[8 * q + 1]), u.y = pack_bf16x2(f[8 * q + 2], f[8 * q + 3]);
        u.z = pack_bf16x2(f[8 * q + 4], f[8 * q + 5]), u.w = pack_bf16x2(f[8 * q + 6], f[8 * q + 7]);
        *reinterpret_cast<uint4*>(dst + (((half * 4 + q) ^ (row & 7)) << 4)) = u;
      }
    }
    fence_proxy_async();
    epi_bar(3);
    if (issuer) {
      const int r0 = c.m_blk * BLOCK_M;
      if constexpr (kF32) {
        const int n0 = c.n_blk * kPairN + i * 64;
        if (n0 < p.N) tma_reduce_add_2d(tmap_out, staging, n0, r0);
        if (n0 + 32 < p.N) tma_reduce_add_2d(tmap_out, staging + 16384, n0 + 32, r0);
        bulk_commit();
        bulk_wait_read<0>();
      } else if constexpr (EPI == EPI_CROSSKV_BF16) {
        const int n0 = c.n_blk * kPairN + i * 64;  // one 64-column box = one (layer, k|v, head)
        if (n0 < p.N) {
          const int which = n0 / p.d_model, hh = (n0 - which * p.d_model) >> 6, layer = which >> 1;
          tma_store_3d((which & 1) ? tmap_out2 : tmap_out, staging + (i & 1) * 16384, 0, r0,
                       (layer * p.kv_batch + c.batch + p.kv_batch_offset) * p.n_head + hh);
        }
        bulk_commit();
        bulk_wait_read<1>();
      } else {
        const int n0 = c.n_blk * kPairN + i * 64;
        if (n0 < p.N) tma_store_2d(tmap_out, staging + (i & 1) * 16384, n0, r0);
        bulk_commit();
        bulk_wait_read<1>();
      }
    }
  }
}

template <int EPI, bool kTmaOut>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
gemm2cta_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                        const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_out2, const GemmGeom g,
                        const GemmParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages2 * kStageBytes2 + kStagingBytes);
  uint64_t* empty_bar = full_bar + kStages2;
  uint64_t* tmem_full_bar = empty_bar + kStages2;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  unsigned char* staging = smem + kStages2 * kStageBytes2;  // 1024-aligned (stage sizes are multiples of 1024)
  float* s_bias = reinterpret_cast<float*>(staging + kStagingBytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (kTmaOut) tma_prefetch_desc(&tmap_out);
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
      if constexpr (kTmaOut) {
        epilogue_tile_tma<EPI>(p, &tmap_out, &tmap_out2, c, tmem_base + acc_stage * kPairN, s_bias, staging, &tmem_full_bar[acc_stage], acc_phase,
                               &tmem_empty_bar[acc_stage], warp - 2, warp & 3, lane);
      } else {
        // direct-store epilogue (conv stems: padded / offset outputs); single bias buffer -> barrier before it is rewritten
        epi_bar(2);
        epilogue_tile<kPairN, EPI, kEpiWarps2>(p, c, tmem_base + acc_stage * kPairN, s_bias, &tmem_full_bar[acc_stage], acc_phase, warp - 2,
                                               warp & 3, lane);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[acc_stage]);
      }
      if (++acc_stage == 2) {
        acc_stage = 0;
        acc_phase ^= 1;
      }
    }
  }

  if (kTmaOut && threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // outstanding TMA stores of the issuer thread
  tcgen05_fence_before();
  cluster_sync_all();  // the peer's smem / TMEM stay alive until both CTAs are done
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <int EPI, bool kTmaOut>
void launch2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& to2, const GemmGeom& g, const GemmParams& p,
             int grid, cudaStream_t stream) {
  gemm2cta_tcgen05_kernel<EPI, kTmaOut><<<grid, kThreads2, kSmemBytes2, stream>>>(ta, tb, to, to2, g, p);
  CUDA_CHECK(cudaGetLastError());
}
template <int EPI, bool kTmaOut>
void set_attr2() {
  CUDA_CHECK(cudaFuncSetAttribute(gemm2cta_tcgen05_kernel<EPI, kTmaOut>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes2));
}

}  // namespace

void gemm2cta_set_attributes() {
  set_attr2<EPI_BIAS_BF16, true>();
  set_attr2<EPI_BIAS_GELU_BF16, true>();
  set_attr2<EPI_BIAS_RESID_F32, true>();
  set_attr2<EPI_CROSSKV_BF16, true>();
  set_attr2<EPI_BIAS_BF16, false>();
  set_attr2<EPI_BIAS_GELU_BF16, false>();
  set_attr2<EPI_BIAS_F32, false>();
  set_attr2<EPI_BIAS_RESID_F32, false>();
  set_attr2<EPI_GELU_POS_F32, false>();
  set_attr2<EPI_CROSSKV_BF16, false>();
}

// geometry / launch for a plan created with two_cta = true (see gemm_plan_create)
void gemm2cta_launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap* to, const CUtensorMap* to2, int epilogue, int n_batch,
                     int n_taps, int kb_per_tap, const int* a_c0, const int* a_row, const int* w_k0, const GemmParams& p, cudaStream_t stream) {
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
  if (to != nullptr) {
    switch (epilogue) {
      case EPI_BIAS_BF16: launch2<EPI_BIAS_BF16, true>(ta, tb, *to, *to2, g, p, grid, stream); break;
      case EPI_BIAS_GELU_BF16: launch2<EPI_BIAS_GELU_BF16, true>(ta, tb, *to, *to2, g, p, grid, stream); break;
      case EPI_BIAS_RESID_F32: launch2<EPI_BIAS_RESID_F32, true>(ta, tb, *to, *to2, g, p, grid, stream); break;
      case EPI_CROSSKV_BF16: launch2<EPI_CROSSKV_BF16, true>(ta, tb, *to, *to2, g, p, grid, stream); break;
      default: throw CudaError("gemm2cta: this epilogue has no TMA-store variant");
    }
    return;
  }
  switch (epilogue) {
    case EPI_BIAS_BF16: launch2<EPI_BIAS_BF16, false>(ta, tb, ta, ta, g, p, grid, stream); break;
    case EPI_BIAS_GELU_BF16: launch2<EPI_BIAS_GELU_BF16, false>(ta, tb, ta, ta, g, p, grid, stream); break;
    case EPI_BIAS_F32: launch2<EPI_BIAS_F32, false>(ta, tb, ta, ta, g, p, grid, stream); break;
    case EPI_BIAS_RESID_F32: launch2<EPI_BIAS_RESID_F32, false>(ta, tb, ta, ta, g, p, grid, stream); break;
    case EPI_GELU_POS_F32: launch2<EPI_GELU_POS_F32, false>(ta, tb, ta, ta, g, p, grid, stream); break;
    case EPI_CROSSKV_BF16: launch2<EPI_CROSSKV_BF16, false>(ta, tb, ta, ta, g, p, grid, stream); break;
    default: throw CudaError("gemm2cta: unsupported epilogue");
  }
}

}  // namespace b200w
