// Shared pieces of the tcgen05 GEMM kernels (1-CTA persistent kernel in gemm_tcgen05.cu, 2-CTA cluster kernel in
// gemm2cta_tcgen05.cu): tile bookkeeping and the fused epilogue of one 128 x BLOCK_N accumulator tile.
#pragma once
#include <cfloat>

#include "common.cuh"
#include "kernels.h"

namespace b200w {
namespace gemm_detail {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;

struct TileCoord {
  int batch, m_blk, n_blk;
};

struct GemmGeom {
  int m_tiles_per_batch, n_tiles, n_batch, num_k_blocks, total_tiles;
  // implicit-GEMM convolution: the k loop runs over n_taps shifted views of A (tap t reads A columns
  // a_c0[t] + k, rows row + a_row[t]) against W columns w_k0[t] + k.  A plain GEMM has one tap.
  int n_taps, kb_per_tap;
  int a_c0[3], a_row[3], w_k0[3];
  // bytes TMA delivers per A stage: 128 rows x 128 B normally; a decoder-step GEMM over M <= 32 / 64 rows loads only a 32- / 64-row
  // box (the tensor core still multiplies 128 rows: the rest of the stage is stale shared memory whose accumulator rows are
  // never stored).  Every CTA of such a GEMM reads the whole A from L2, so this cuts its operand traffic up to 4x.
  int a_stage_bytes;
};

// Epilogue of one accumulator tile, executed by all kEpiWarps epilogue warps of a CTA (epi_warp = 0..kEpiWarps-1, its
// TMEM lane group is tmem_lane_group = (hardware warp id) % 4).  tmem_acc = TMEM column base of the accumulator stage.
// Waits on tmem_full_bar itself (after staging the bias and prefetching the first positional-embedding chunk).
template <int BLOCK_N, int EPI, int kEpiWarps>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, const TileCoord& c, uint32_t tmem_acc, float* sb, uint64_t* tmem_full_bar,
                                              uint32_t acc_phase, int epi_warp, int lg, int lane) {
  constexpr int kEpiThreads = kEpiWarps * 32;
  constexpr int CH_PER_WARP = (BLOCK_N / 32) / (kEpiWarps / 4);
  // the residual update x += acc + bias is a vector reduction at L2 (red.global.add.v4.f32): no read of x in the epilogue,
  // no registers holding it; one add per element per kernel, so the result is the same as load-add-store
  constexpr bool kHasExtra = (EPI == EPI_GELU_POS_F32);
  const int half = epi_warp >> 2;
  const int etid = epi_warp * 32 + lane;
  const int row_in_tile = lg * 32 + lane;
  const int r = c.m_blk * BLOCK_M + row_in_tile;  // row within the batch
  const bool row_ok = r < p.rows_valid;
  const long orow = (long)c.batch * p.out_batch_pitch + (long)(r + p.out_row_offset) * p.ldo;
  const int ch0 = half * CH_PER_WARP;
  // while the MMAs of this tile run: stage the tile's bias slice in smem, prefetch the first chunk's residual / pos rows
  for (int i = etid; i < BLOCK_N; i += kEpiThreads) {
    const int n = c.n_blk * BLOCK_N + i;
    sb[i] = (p.bias != nullptr && n < p.N) ? p.bias[n] : 0.f;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
  float4 extra[2][8];
  auto load_extra = [&](float4(&dst)[8], int ch) {
    if constexpr (kHasExtra) {
      const int n0 = c.n_blk * BLOCK_N + ch * 32;
      const float* src = p.pos + (long)r * p.N + n0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        dst[j] = (row_ok && n0 + j * 4 + 4 <= p.N) ? *reinterpret_cast<const float4*>(src + j * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  load_extra(extra[0], ch0);
  mbar_wait(tmem_full_bar, acc_phase);
  tcgen05_fence_after();
  const uint32_t tbase = tmem_acc + (static_cast<uint32_t>(lg * 32) << 16);
  uint32_t v[2][32];
  tmem_ld_32x32b_x32(tbase + ch0 * 32, v[0]);
  float best = -FLT_MAX;
  int best_idx = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < CH_PER_WARP; ++i) {
    const int ch = ch0 + i;
    tcgen05_wait_ld();
    if (i + 1 < CH_PER_WARP) {  // next chunk's TMEM read and residual loads fly while this chunk is processed
      tmem_ld_32x32b_x32(tbase + (ch + 1) * 32, v[(i + 1) & 1]);
      load_extra(extra[(i + 1) & 1], ch + 1);
    }
    const int n0 = c.n_blk * BLOCK_N + ch * 32;
    if (!row_ok || n0 >= p.N) continue;
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bv = *reinterpret_cast<const float4*>(sb + ch * 32 + j);
      f[j] = __uint_as_float(v[i & 1][j]) + bv.x, f[j + 1] = __uint_as_float(v[i & 1][j + 1]) + bv.y;
      f[j + 2] = __uint_as_float(v[i & 1][j + 2]) + bv.z, f[j + 3] = __uint_as_float(v[i & 1][j + 3]) + bv.w;
    }
    const bool full = n0 + 32 <= p.N;
    if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16) {
      if constexpr (EPI == EPI_BIAS_GELU_BF16) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = gelu_fast(f[j]);
      }
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + orow + n0;
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 u;
          u.x = pack_bf16x2(f[j], f[j + 1]), u.y = pack_bf16x2(f[j + 2], f[j + 3]);
          u.z = pack_bf16x2(f[j + 4], f[j + 5]), u.w = pack_bf16x2(f[j + 6], f[j + 7]);
          *reinterpret_cast<uint4*>(o + j) = u;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < p.N) o[j] = __float2bfloat16_rn(f[j]);
      }
    } else if constexpr (EPI == EPI_BIAS_F32 || EPI == EPI_BIAS_RESID_F32 || EPI == EPI_GELU_POS_F32) {
      float* o = reinterpret_cast<float*>(p.out) + orow + n0;
      if constexpr (EPI == EPI_GELU_POS_F32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = gelu_fast(f[j]);
      }
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 u = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          if constexpr (kHasExtra) {
            const float4 rv = extra[i & 1][j >> 2];
            u.x += rv.x, u.y += rv.y, u.z += rv.z, u.w += rv.w;
          }
          if constexpr (EPI == EPI_BIAS_RESID_F32) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(u.x), "f"(u.y), "f"(u.z), "f"(u.w) : "memory");
          } else {
            *reinterpret_cast<float4*>(o + j) = u;
          }
        }
      } else {
        const float* xs = (EPI == EPI_GELU_POS_F32) ? p.pos + (long)r * p.N + n0 : o;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < p.N) o[j] = ((kHasExtra || EPI == EPI_BIAS_RESID_F32) ? xs[j] : 0.f) + f[j];
      }
    } else if constexpr (EPI == EPI_CROSSKV_BF16) {
      // n0 is 32-aligned, so the chunk stays inside one (layer, k|v, head) slice of 64 columns
      const int which = n0 / p.d_model;  // layer * 2 + kv
      const int within = n0 - which * p.d_model;
      const int h = within >> 6, dh = within & 63;
      const int layer = which >> 1;
      __nv_bfloat16* base = (which & 1) ? p.cross_v : p.cross_k;
      const long off = ((((long)layer * p.kv_batch + (c.batch + p.kv_batch_offset)) * p.n_head + h) * p.n_ctx_kv + r) * 64 + dh;
      __nv_bfloat16* o = base + off;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 u;
        u.x = pack_bf16x2(f[j], f[j + 1]), u.y = pack_bf16x2(f[j + 2], f[j + 3]);
        u.z = pack_bf16x2(f[j + 4], f[j + 5]), u.w = pack_bf16x2(f[j + 6], f[j + 7]);
        *reinterpret_cast<uint4*>(o + j) = u;
      }
    } else if constexpr (EPI == EPI_ARGMAX) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (n0 + j < p.N && f[j] > best) {  // strict '>' keeps the first maximum (std::max_element, Whisper.cpp:42-45)
          best = f[j];
          best_idx = n0 + j;
        }
      }
      if (p.out != nullptr) {
        float* o = reinterpret_cast<float*>(p.out) + orow + n0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < p.N) o[j] = f[j];
      }
    }
  }
  if constexpr (EPI == EPI_ARGMAX) {
    if (row_ok) {
      const long pi = ((long)c.batch * p.rows_valid + r) * p.part_ld + c.n_blk * (kEpiWarps / 4) + half;
      p.part_val[pi] = best;
      p.part_idx[pi] = best_idx;
    }
  }
}

}  // namespace gemm_detail
}  // namespace b200w
