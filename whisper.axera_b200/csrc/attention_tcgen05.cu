// K5: non-causal multi-head attention of the audio encoder on tcgen05 tensor cores (head_dim 64, fp32 softmax).
//
// Math: MultiHeadAttention.qkv_attention of openai-whisper 20240930 as used by the reference's encoder graph
// (/root/reference/model_convert/export_onnx.py:153-181): w = softmax((q * s)(k * s)^T) with s = 64^-0.25, out = w v.
//
// One CTA = 128 query rows of one (chunk, head); keys/values stream through in 64-key tiles.  320 threads:
//   warp 0 (1 lane)  TMA producer: Q tile once, then K/V tiles through a 4-slot ring in the order the MMA warp
//                    consumes them (K0 K1 V0 K2 V1 ...), straight out of the fused QKV activation [B][T][3d] (3-D tensor
//                    map, SWIZZLE_128B; rows past T are zero-filled by TMA)
//   warp 1 (1 lane)  MMA issuer: S_j = Q K_j^T (M128 N64 K64, both operands K-major) into one of two TMEM S buffers,
//                    then PV_j = P_j V_j (A = P from smem K-major, B = V tile as loaded = MN-major) into one of two
//                    64-column TMEM buffers; tcgen05.commit signals the softmax warps and frees smem slots
//   warps 2-9        softmax: two threads per query row (TMEM lane), 32 keys and 32 output columns each.  Read S_j with
//                    tcgen05.ld, online max (halves exchanged through smem) / exp2 / sum in fp32, write P_j as bf16 into
//                    the swizzled K-major smem tile, then fold the previous tile's PV product into the fp32 output they
//                    keep in registers: O = (O + PV_{j-1}) * alpha_j (rescale skipped while no maximum changes).
// S_{j+1} is issued before PV_j, so the tensor core works on the next scores while the softmax warps are in their exp
// phase; two CTAs are resident per SM (256 TMEM columns and ~82 KB smem each) and interleave as well.
#include <cfloat>

#include "common.cuh"
#include "kernels.h"

namespace b200w {
namespace {

constexpr int kQ = 128;
constexpr int kKV = 64;
constexpr int kRing = 4;
constexpr int kAttThreads = 320;  // TMA warp, MMA warp, 8 softmax warps
constexpr int kQBytes = kQ * 64 * 2;         // 16 KB
constexpr int kKVBytes = kKV * 64 * 2;       // 8 KB
constexpr int kPBytes = kQ * kKV * 2;        // 16 KB

constexpr float kScaleLog2 = 0.125f * 1.4426950408889634f;  // (64^-0.25)^2 * log2(e)

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// explicit shared-space accesses (32-bit addresses): the softmax loop is issue-bound, generic 64-bit address math costs slots
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
// named barrier over the two warps that share a TMEM lane group (ids 1..4)
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

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
  uint64_t s_full[2], s_free[2], p_ready[2], p_free[2], o_full[2], o_free[2];
  uint32_t tmem_slot;
};
constexpr int kXchBytes = 2 * 2 * kQ * 4;  // row-max exchange between the two threads of a row: [tile parity][half][row]
constexpr int kSmemBytes = kQBytes + kRing * kKVBytes + 2 * kPBytes + 256 + kXchBytes + 1024;

__global__ void __launch_bounds__(kAttThreads, 2)
encoder_attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                                 __nv_bfloat16* __restrict__ out, int T, int d) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;
  unsigned char* sRing = sQ + kQBytes;
  unsigned char* sP = sRing + kRing * kKVBytes;
  Bars& bar = *reinterpret_cast<Bars*>(sP + 2 * kPBytes);
  float(*xch)[2][kQ] = reinterpret_cast<float(*)[2][kQ]>(sP + 2 * kPBytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQ, h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (T + kKV - 1) / kKV;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    mbar_init(&bar.q_full, 1);
    for (int i = 0; i < kRing; ++i) mbar_init(&bar.ring_full[i], 1), mbar_init(&bar.ring_free[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar.s_full[i], 1), mbar_init(&bar.s_free[i], 8);
      mbar_init(&bar.p_ready[i], 8), mbar_init(&bar.p_free[i], 1);
      mbar_init(&bar.o_full[i], 1), mbar_init(&bar.o_free[i], 8);
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
  // TMEM columns: S buffers at 0 and 64, PV buffers at 128 and 192

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bar.q_full, kQBytes);
      tma_load_3d(sQ, &tmap_q, &bar.q_full, h * 64, q0, b);
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
      const uint64_t desc_q = umma_desc_kmajor_sw128(smem_u32(sQ));
      mbar_wait(&bar.q_full, 0);
      int i = 0;
      auto issue_s = [&](int j) {
        const int slot = i % kRing;
        mbar_wait(&bar.ring_full[slot], (i / kRing) & 1);
        tcgen05_fence_after();
        const uint64_t desc_k = umma_desc_kmajor_sw128(smem_u32(sRing + slot * kKVBytes));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + (j & 1) * 64, desc_q + 2 * k, desc_k + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&bar.ring_free[slot]);
        umma_commit(&bar.s_full[j & 1]);
        ++i;
      };
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) {
          mbar_wait(&bar.s_free[(j + 1) & 1], (((j + 1) >> 1) & 1) ^ 1);
          tcgen05_fence_after();
          issue_s(j + 1);
        }
        mbar_wait(&bar.p_ready[j & 1], (j >> 1) & 1);
        mbar_wait(&bar.o_free[j & 1], ((j >> 1) & 1) ^ 1);
        const int slot = i % kRing;
        mbar_wait(&bar.ring_full[slot], (i / kRing) & 1);
        tcgen05_fence_after();
        const uint64_t desc_p = umma_desc_kmajor_sw128(smem_u32(sP + (j & 1) * kPBytes));
        const uint64_t desc_v = umma_desc_mnmajor_sw128(smem_u32(sRing + slot * kKVBytes));
#pragma unroll
        for (int k = 0; k < 4; ++k)  // 16 keys per MMA: +32 bytes along P's rows, +16 rows (2048 bytes) down the V tile
          umma_bf16_ss(tmem_base + 128 + (j & 1) * 64, desc_p + 2 * k, desc_v + (2048 >> 4) * k, idesc_pv, k > 0 ? 1u : 0u);
        umma_commit(&bar.ring_free[slot]);
        umma_commit(&bar.p_free[j & 1]);
        umma_commit(&bar.o_full[j & 1]);
        ++i;
      }
    }
  } else {
    // 8 softmax warps: two warps share a TMEM lane group (= 32 query rows); each takes 32 of the 64 keys of a tile and
    // 32 of the 64 output columns, so a row's work is split over two threads (more warps in flight per scheduler)
    const int lg = warp & 3;
    const int ch = (warp - 2) >> 2;  // column half
    const int row = lg * 32 + lane;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    float m_run = -INFINITY, l_run = 0.f;
    // shared-space addresses of this thread's exchange slots and of its four 16-byte chunks of a P row
    const uint32_t xch_mine = smem_u32(&xch[0][ch][row]), xch_other = smem_u32(&xch[0][ch ^ 1][row]);
    constexpr uint32_t kXchParity = 2 * kQ * 4;
    const uint32_t p_row = smem_u32(sP) + row * 128;
    uint32_t p_chunk[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) p_chunk[q] = p_row + (((ch * 4 + q) ^ (row & 7)) << 4);
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int sb = j & 1;
      mbar_wait(&bar.s_full[sb], (j >> 1) & 1);
      tcgen05_fence_after();
      uint32_t s0[32];
      tmem_ld_32x32b_x32(tlane + sb * 64 + ch * 32, s0);
      tcgen05_wait_ld();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar.s_free[sb]);  // S_j is in registers: the tensor core may overwrite the buffer
      const int nvalid = T - j * kKV - ch * 32;     // keys of this half-tile that exist
      if (nvalid < 32) {                            // last tile only: mask the keys past T (warp-uniform branch)
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i >= nvalid) s0[i] = 0xff800000u;     // -inf
      }
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(s0[i]));
      // the row maximum needs the other half's keys: exchange through smem (double-buffered by tile parity)
      st_shared_f32(xch_mine + sb * kXchParity, mx);
      pair_barrier(1 + lg);
      mx = fmaxf(mx, ld_shared_f32(xch_other + sb * kXchParity));
      const float m_new = fmaxf(m_run, mx);
      const float alpha = fast_exp2((m_run - m_new) * kScaleLog2);
      const float msc = m_new * kScaleLog2;
      // packed fp32x2 math (sm_100 FFMA2 / FADD2): half the issue slots for the exp arguments and the row sum
      const float2 sc2 = make_float2(kScaleLog2, kScaleLog2), nm2 = make_float2(-msc, -msc);
      float2 rs2 = make_float2(0.f, 0.f);
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float2 arg = __ffma2_rn(make_float2(__uint_as_float(s0[i]), __uint_as_float(s0[i + 1])), sc2, nm2);
        const float2 e = make_float2(fast_exp2(arg.x), fast_exp2(arg.y));  // exp2(-inf) = 0 for masked keys
        rs2 = __fadd2_rn(rs2, e);
        pk[i >> 1] = pack_bf16x2(e.x, e.y);
      }
      const float rs = rs2.x + rs2.y;
      l_run = fmaf(l_run, alpha, rs);
      // P_j -> smem, K-major SWIZZLE_128B: row r at r*128 bytes, 16-byte chunk q (keys 8q..8q+7) at position q ^ (r & 7)
      mbar_wait(&bar.p_free[sb], ((j >> 1) & 1) ^ 1);
#pragma unroll
      for (int q = 0; q < 4; ++q) st_shared_v4(p_chunk[q] + sb * kPBytes, pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar.p_ready[sb]);
      // fold in the previous tile's P V product and rescale to the new running maximum
      if (j > 0) {
        const int ob = (j - 1) & 1;
        mbar_wait(&bar.o_full[ob], ((j - 1) >> 1) & 1);
        tcgen05_fence_after();
        uint32_t t0[32];
        tmem_ld_32x32b_x32(tlane + 128 + ob * 64 + ch * 32, t0);
        tcgen05_wait_ld();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar.o_free[ob]);
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
          const float2 al2 = make_float2(alpha, alpha);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 r = __fmul2_rn(__fadd2_rn(make_float2(o[i], o[i + 1]), make_float2(__uint_as_float(t0[i]), __uint_as_float(t0[i + 1]))), al2);
            o[i] = r.x, o[i + 1] = r.y;
          }
        } else {  // the running maximum of every row of this warp is unchanged (the common case after the first tiles)
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 r = __fadd2_rn(make_float2(o[i], o[i + 1]), make_float2(__uint_as_float(t0[i]), __uint_as_float(t0[i + 1])));
            o[i] = r.x, o[i + 1] = r.y;
          }
        }
      }
      m_run = m_new;
    }
    {
      const int ob = (n_tiles - 1) & 1;
      mbar_wait(&bar.o_full[ob], ((n_tiles - 1) >> 1) & 1);
      tcgen05_fence_after();
      uint32_t t0[32];
      tmem_ld_32x32b_x32(tlane + 128 + ob * 64 + ch * 32, t0);
      tcgen05_wait_ld();
      // the row sum is split over the two threads of the row
      // every exchange of the tile loop has been consumed by both threads of the row (they met at the last pair barrier
      // after reading), parity slot of the next tile index is free
      st_shared_f32(xch_mine + (n_tiles & 1) * kXchParity, l_run);
      pair_barrier(1 + lg);
      const float inv = 1.f / (l_run + ld_shared_f32(xch_other + (n_tiles & 1) * kXchParity));
      if (q0 + row < T) {
        __nv_bfloat16* dst = out + ((long)b * T + q0 + row) * d + h * 64 + ch * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u;
          u.x = pack_bf16x2((o[i] + __uint_as_float(t0[i])) * inv, (o[i + 1] + __uint_as_float(t0[i + 1])) * inv);
          u.y = pack_bf16x2((o[i + 2] + __uint_as_float(t0[i + 2])) * inv, (o[i + 3] + __uint_as_float(t0[i + 3])) * inv);
          u.z = pack_bf16x2((o[i + 4] + __uint_as_float(t0[i + 4])) * inv, (o[i + 5] + __uint_as_float(t0[i + 5])) * inv);
          u.w = pack_bf16x2((o[i + 6] + __uint_as_float(t0[i + 6])) * inv, (o[i + 7] + __uint_as_float(t0[i + 7])) * inv);
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
  dim3 grid((T + kQ - 1) / kQ, n_head, B);
  encoder_attention_tcgen05_kernel<<<grid, kAttThreads, kSmemBytes, stream>>>(tq, tkv, out, T, d);
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace b200w
