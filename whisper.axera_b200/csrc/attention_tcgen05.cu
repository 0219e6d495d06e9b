// K5: non-causal multi-head attention of the audio encoder on tcgen05 tensor cores (head_dim 64, fp32 softmax).
//
// Math: MultiHeadAttention.qkv_attention of openai-whisper 20240930 as used by the reference's encoder graph
// (/root/reference/model_convert/export_onnx.py:153-181): w = softmax((q * s)(k * s)^T) with s = 64^-0.25, out = w v.
//
// One CTA = 256 query rows (two 128-row tiles Q0, Q1) of one (chunk, head); keys/values stream through in 64-key tiles
// shared by both query tiles.  320 threads:
//   warp 0 (1 lane)  TMA producer: Q0/Q1 once, then K/V tiles through a 5-slot ring in the order the MMA warp consumes
//                    them (K0 K1 V0 K2 V1 ...), straight out of the fused QKV activation [B][T][3d] (3-D tensor map,
//                    SWIZZLE_128B; rows past T are zero-filled by TMA)
//   warp 1 (1 lane)  MMA issuer, per key tile j and query tile t: S_t(j+1) = Q_t K_{j+1}^T (M128 N64 K64) as soon as the
//                    softmax warps have pulled S_t(j) into registers, then O_t += P_t(j) V_j (A = P from smem, B = V tile
//                    as loaded = MN-major).  O_t accumulates in TMEM over all key tiles.
//   warps 2-5 / 6-9  softmax of Q0 / Q1: one thread per query row (= TMEM lane), 64 keys per tile: tcgen05.ld, row max,
//                    exp2, row sum (packed fp32x2 math), P_t as bf16 into the swizzled K-major smem tile.  The running
//                    maximum is only raised when it grows by more than 2^8 (P stays <= 256, exact in the final division),
//                    so the rescale of O_t in TMEM (tcgen05.ld / st by the same threads) happens on the first tiles only.
// TMEM: S0, S1 at columns 0 / 64, O0, O1 at 128 / 192 -> 256 columns, two CTAs per SM (~107 KB smem each) interleave.
#include <cfloat>

#include "common.cuh"
#include "kernels.h"

namespace b200w {
namespace {

constexpr int kQ = 128;        // rows per query tile
constexpr int kQTiles = 2;     // query tiles per CTA
constexpr int kKV = 64;
constexpr int kRing = 5;
constexpr int kAttThreads = 320;  // TMA warp, MMA warp, 2 x 4 softmax warps
constexpr int kQBytes = kQ * 64 * 2;         // 16 KB per query tile
constexpr int kKVBytes = kKV * 64 * 2;       // 8 KB
constexpr int kPBytes = kQ * kKV * 2;        // 16 KB per query tile

constexpr float kScaleLog2 = 0.125f * 1.4426950408889634f;  // (64^-0.25)^2 * log2(e)
constexpr float kRescaleThreshold = 8.f;                      // log2 units the stale running maximum may lag

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// explicit shared-space stores (32-bit addresses): the softmax loop is issue-bound, generic 64-bit address math costs slots
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// MN-major bf16 operand tile [rows = K index][64 elements of N, 128 bytes] written by TMA with SWIZZLE_128B:
// 8-row groups 1024 bytes apart (stride byte offset); a single 64-element block in the N direction.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>(1024 >> 4) << 16;  // LBO (unused: N = 64 is one swizzle atom wide)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;  // SBO
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

struct Bars {
  uint64_t q_full;
  uint64_t ring_full[kRing], ring_free[kRing];
  uint64_t s_full[kQTiles], s_free[kQTiles], p_ready[kQTiles], pv_done[kQTiles];
  uint32_t tmem_slot;
};
constexpr int kSmemBytes = kQTiles * kQBytes + kRing * kKVBytes + kQTiles * kPBytes + 256 + 1024;
static_assert(sizeof(Bars) <= 256, "barrier block");
static_assert(2 * (kSmemBytes + 1024) <= 228 * 1024, "two CTAs per SM");

__global__ void __launch_bounds__(kAttThreads, 2)
encoder_attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                                 __nv_bfloat16* __restrict__ out, int T, int d) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;
  unsigned char* sRing = sQ + kQTiles * kQBytes;
  unsigned char* sP = sRing + kRing * kKVBytes;
  Bars& bar = *reinterpret_cast<Bars*>(sP + kQTiles * kPBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (kQ * kQTiles), h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (T + kKV - 1) / kKV;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    mbar_init(&bar.q_full, 1);
    for (int i = 0; i < kRing; ++i) mbar_init(&bar.ring_full[i], 1), mbar_init(&bar.ring_free[i], 1);
    for (int i = 0; i < kQTiles; ++i) {
      mbar_init(&bar.s_full[i], 1), mbar_init(&bar.s_free[i], 4);
      mbar_init(&bar.p_ready[i], 4), mbar_init(&bar.pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&bar.tmem_slot, 256);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = bar.tmem_slot;
  // TMEM columns: S_t at t * 64, O_t at 128 + t * 64

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bar.q_full, kQTiles * kQBytes);
      for (int t = 0; t < kQTiles; ++t) tma_load_3d(sQ + t * kQBytes, &tmap_q, &bar.q_full, h * 64, q0 + t * kQ, b);
      int i = 0;
      auto load_tile = [&](int col, int tile) {
        const int slot = i % kRing;
        mbar_wait(&bar.ring_free[slot], ((i / kRing) & 1) ^ 1);
        mbar_arrive_expect_tx(&bar.ring_full[slot], kKVBytes);
        tma_load_3d(sRing + slot * kKVBytes, &tmap_kv, &bar.ring_full[slot], col, tile * kKV, b);
        ++i;
      };
      const int kcol = d + h * 64, vcol = 2 * d + h * 64;
      load_tile(kcol, 0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) load_tile(kcol, j + 1);
        load_tile(vcol, j);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(kQ, kKV);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(kQ, 64) | (1u << 16);  // B operand (V tile) is MN-major
      uint64_t desc_q[kQTiles], desc_p[kQTiles];
      for (int t = 0; t < kQTiles; ++t) {
        desc_q[t] = umma_desc_kmajor_sw128(smem_u32(sQ + t * kQBytes));
        desc_p[t] = umma_desc_kmajor_sw128(smem_u32(sP + t * kPBytes));
      }
      auto issue_s = [&](int t, uint64_t desc_k) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + t * 64, desc_q[t] + 2 * k, desc_k + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&bar.s_full[t]);
      };
      mbar_wait(&bar.q_full, 0);
      int i = 0;
      {
        mbar_wait(&bar.ring_full[0], 0);
        tcgen05_fence_after();
        const uint64_t desc_k = umma_desc_kmajor_sw128(smem_u32(sRing));
        for (int t = 0; t < kQTiles; ++t) issue_s(t, desc_k);
        umma_commit(&bar.ring_free[0]);
        ++i;
      }
      for (int j = 0; j < n_tiles; ++j) {
        const bool has_next = j + 1 < n_tiles;
        int slot_k = 0;
        uint64_t desc_k = 0;
        if (has_next) {
          slot_k = i % kRing;
          mbar_wait(&bar.ring_full[slot_k], (i / kRing) & 1);
          desc_k = umma_desc_kmajor_sw128(smem_u32(sRing + slot_k * kKVBytes));
          ++i;
        }
        const int slot_v = i % kRing;
        const uint32_t phase_v = (i / kRing) & 1;
        const uint64_t desc_v = umma_desc_mnmajor_sw128(smem_u32(sRing + slot_v * kKVBytes));
        ++i;
        for (int t = 0; t < kQTiles; ++t) {
          if (has_next) {
            mbar_wait(&bar.s_free[t], j & 1);  // S_t(j) sits in the softmax warps' registers
            tcgen05_fence_after();
            issue_s(t, desc_k);
            if (t == kQTiles - 1) umma_commit(&bar.ring_free[slot_k]);
          }
          mbar_wait(&bar.p_ready[t], j & 1);  // P_t(j) is in smem and O_t has been rescaled if needed
          if (t == 0) mbar_wait(&bar.ring_full[slot_v], phase_v);
          tcgen05_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 16 keys per MMA: +32 bytes along P's rows, +16 rows (2048 bytes) down the V tile
            umma_bf16_ss(tmem_base + 128 + t * 64, desc_p[t] + 2 * k, desc_v + (2048 >> 4) * k, idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
          umma_commit(&bar.pv_done[t]);
        }
        umma_commit(&bar.ring_free[slot_v]);
      }
    }
  } else {
    const int t = (warp - 2) >> 2;  // query tile of this warp
    const int lg = warp & 3;        // TMEM lane group this warp may access
    const int row = lg * 32 + lane;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    const uint32_t s_col = tlane + t * 64, o_col = tlane + 128 + t * 64;
    // shared-space addresses of the eight 16-byte chunks of this thread's P row (K-major SWIZZLE_128B: chunk q of row r
    // sits at position q ^ (r & 7))
    const uint32_t p_row = smem_u32(sP + t * kPBytes) + row * 128;
    const uint32_t swz = static_cast<uint32_t>(row & 7) << 4;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&bar.s_full[t], j & 1);
      tcgen05_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32b_x32(s_col, s0);
      tmem_ld_32x32b_x32(s_col + 32, s1);
      tcgen05_wait_ld();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar.s_free[t]);  // S_t(j) is in registers: the tensor core may write S_t(j+1)
      const int nvalid = T - j * kKV;              // keys of this tile that exist
      if (nvalid < kKV) {                          // last tile only: mask the keys past T (warp-uniform branch)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= nvalid) s0[i] = 0xff800000u;  // -inf
          if (i + 32 >= nvalid) s1[i] = 0xff800000u;
        }
      }
      float mx = fmaxf(__uint_as_float(s0[0]), __uint_as_float(s1[0]));
#pragma unroll
      for (int i = 1; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(s0[i]), __uint_as_float(s1[i])));
      // P_t(j-1) has been consumed and O_t holds every product up to tile j-1
      if (j > 0) mbar_wait(&bar.pv_done[t], (j - 1) & 1);
      const bool grow = (mx - m_run) * kScaleLog2 > kRescaleThreshold;  // always true on the first tile (m_run = -inf)
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? mx : m_run;
        const float alpha = fast_exp2((m_run - m_new) * kScaleLog2);  // 1 for the rows that keep their maximum, 0 on tile 0
        l_run *= alpha;
        m_run = m_new;
        if (j > 0) {
          tcgen05_fence_after();
#pragma unroll 1
          for (int c8 = 0; c8 < 64; c8 += 8) {  // rare path: 8 columns at a time keeps the register footprint small
            uint32_t o[8];
            tmem_ld_32x32b_x8(o_col + c8, o);
            tcgen05_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x8(o_col + c8, o);
          }
          tcgen05_wait_st();
        }
      }
      // packed fp32x2 math (sm_100 FFMA2 / FADD2): half the issue slots for the exp arguments and the row sum
      const float msc = m_run * kScaleLog2;
      const float2 sc2 = make_float2(kScaleLog2, kScaleLog2), nm2 = make_float2(-msc, -msc);
      float2 rs2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 8; ++q) {  // 16-byte chunk q = keys 8q .. 8q+7
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = (q & 3) * 8 + 2 * i;
          const uint32_t lo = q < 4 ? s0[k] : s1[k], hi = q < 4 ? s0[k + 1] : s1[k + 1];
          const float2 arg = __ffma2_rn(make_float2(__uint_as_float(lo), __uint_as_float(hi)), sc2, nm2);
          const float2 e = make_float2(fast_exp2(arg.x), fast_exp2(arg.y));  // exp2(-inf) = 0 for masked keys
          rs2 = __fadd2_rn(rs2, e);
          pk[i] = pack_bf16x2(e.x, e.y);
        }
        st_shared_v4(p_row + ((static_cast<uint32_t>(q) << 4) ^ swz), pk[0], pk[1], pk[2], pk[3]);
      }
      l_run += rs2.x + rs2.y;
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar.p_ready[t]);
    }
    // epilogue: O_t / l
    mbar_wait(&bar.pv_done[t], (n_tiles - 1) & 1);
    tcgen05_fence_after();
    const float inv = 1.f / l_run;
    const int q_row = q0 + t * kQ + row;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(o_col + half * 32, o);
      tcgen05_wait_ld();
      if (q_row < T) {
        __nv_bfloat16* dst = out + ((long)b * T + q_row) * d + h * 64 + half * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + i) = u;
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

void attention_tcgen05_set_attributes() {
  CUDA_CHECK(cudaFuncSetAttribute(encoder_attention_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  CUDA_CHECK(cudaFuncSetAttribute(encoder_attention_tcgen05_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
}

void launch_encoder_attention_tcgen05(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int n_head, cudaStream_t stream) {
  const int d = n_head * 64;
  const uint64_t dims[3] = {(uint64_t)3 * d, (uint64_t)T, (uint64_t)B};
  const uint64_t pitches[2] = {(uint64_t)3 * d * 2, (uint64_t)T * 3 * d * 2};
  const uint32_t box_q[3] = {64, kQ, 1}, box_kv[3] = {64, kKV, 1};
  const CUtensorMap tq = make_tmap_bf16_sw128(qkv, 3, dims, pitches, box_q);
  const CUtensorMap tkv = make_tmap_bf16_sw128(qkv, 3, dims, pitches, box_kv);
  dim3 grid((T + kQ * kQTiles - 1) / (kQ * kQTiles), n_head, B);
  encoder_attention_tcgen05_kernel<<<grid, kAttThreads, kSmemBytes, stream>>>(tq, tkv, out, T, d);
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace b200w
