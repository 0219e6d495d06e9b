// Encoder-side kernels that are not GEMMs: LayerNorm (K3) and non-causal multi-head attention (K5).
//
// Arithmetic restated from openai-whisper 20240930 whisper/model.py as used by the reference's encoder
// wrapper (/root/reference/model_convert/export_onnx.py:153-181): LayerNorm eps 1e-5 with fp32 statistics;
// MultiHeadAttention.qkv_attention with scale (d/H)^-0.25 on both q and k (= 1/8 on the product for
// head_dim 64), softmax in fp32, no mask for the audio encoder.
#include <cfloat>

#include "common.cuh"
#include "kernels.h"

namespace b200w {
namespace {

// ---------------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row kept in registers (d <= 1280 -> <= 10 float4 per lane)
// ---------------------------------------------------------------------------------------------------------
constexpr int kLnMaxVec = 10;

__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, int rows, int d) {
  const int lane = threadIdx.x & 31;
  // gamma / beta do not depend on the predecessor: pull their lines towards L2 while waiting for it (PDL).  Decoder-step
  // launches only: with the encoder's 192 k rows millions of prefetches of the same few lines serialise on one L2 slice
  // (measured: 130 -> 635 us per launch).
  if (rows <= 1024 && lane * 32 < d) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(gamma + lane * 32));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(beta + lane * 32));
  }
  pdl_wait();
  pdl_launch_dependents();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= rows) return;
  const int nvec = d >> 7;  // float4 per lane (d is a multiple of 128)
  const float4* xr = reinterpret_cast<const float4*>(x + (long)warp * d);
  float4 v[kLnMaxVec];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    if (i < nvec) {
      v[i] = xr[i * 32 + lane];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    if (i < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      q += (a * a + b * b) + (c * c + e * e);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  uint2* yr = reinterpret_cast<uint2*>(y + (long)warp * d);
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    if (i < nvec) {
      const float4 g = g4[i * 32 + lane], bb = b4[i * 32 + lane];
      uint2 o;
      o.x = pack_bf16x2((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y);
      o.y = pack_bf16x2((v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w);
      yr[i * 32 + lane] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Flash attention (online softmax), head_dim 64, bf16 mma.sync m16n8k16 with fp32 accumulation.
// One CTA = 128 query rows of one (batch, head); 8 warps x 16 rows; K/V streamed in 64-key tiles (cp.async,
// double buffered).  smem tiles are [rows][64] bf16 with the 16-byte chunk index XOR-swizzled by (row & 7).
// ---------------------------------------------------------------------------------------------------------
constexpr int kAttThreads = 256;
constexpr int kQTile = 128;
constexpr int kKvTile = 64;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// address of 16-byte chunk `chunk` (0..7) of row `row` in a swizzled [rows][64] bf16 tile
__device__ __forceinline__ __nv_bfloat16* swz(__nv_bfloat16* tile, int row, int chunk) {
  return tile + row * 64 + ((chunk ^ (row & 7)) << 3);
}

__global__ void __launch_bounds__(kAttThreads) encoder_attention_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                       __nv_bfloat16* __restrict__ out, int T, int d) {
  extern __shared__ __align__(128) unsigned char att_smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(att_smem);  // [128][64]
  __nv_bfloat16* sK = sQ + kQTile * 64;                            // [2][64][64]
  __nv_bfloat16* sV = sK + 2 * kKvTile * 64;                       // [2][64][64]

  const int q0 = blockIdx.x * kQTile;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long ld = 3L * d;
  const __nv_bfloat16* base = qkv + (long)b * T * ld;
  const __nv_bfloat16* gq = base + h * 64;
  const __nv_bfloat16* gk = base + d + h * 64;
  const __nv_bfloat16* gv = base + 2 * d + h * 64;

  // Q tile: 128 rows x 8 chunks
  for (int i = tid; i < kQTile * 8; i += kAttThreads) {
    const int r = i >> 3, c = i & 7;
    const bool ok = q0 + r < T;
    cp_async16(swz(sQ, r, c), gq + (long)(ok ? q0 + r : 0) * ld + c * 8, ok);
  }
  auto load_kv = [&](int tile, int buf) {
    const int k0 = tile * kKvTile;
    for (int i = tid; i < kKvTile * 8; i += kAttThreads) {
      const int r = i >> 3, c = i & 7;
      const bool ok = k0 + r < T;
      const long row = ok ? k0 + r : 0;
      cp_async16(swz(sK + buf * kKvTile * 64, r, c), gk + row * ld + c * 8, ok);
      cp_async16(swz(sV + buf * kKvTile * 64, r, c), gv + row * ld + c * 8, ok);
    }
  };
  load_kv(0, 0);
  cp_async_commit();

  const int n_tiles = (T + kKvTile - 1) / kKvTile;
  const float c_log2 = 0.125f * 1.4426950408889634f;  // (64^-0.25)^2 * log2(e)

  uint32_t qf[4][4];
  float o_acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) o_acc[j][e] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  for (int tile = 0; tile < n_tiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < n_tiles) {
      load_kv(tile + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (tile == 0) {
      // Q fragments: 16 rows of this warp x 64 dh -> 4 k-steps
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = ks * 2 + (lane >> 4);
        ldmatrix_x4(qf[ks], swz(sQ, r, c));
      }
    }
    const __nv_bfloat16* tK = sK + buf * kKvTile * 64;
    const __nv_bfloat16* tV = sV + buf * kKvTile * 64;

    // S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {  // pairs of n-tiles (16 keys)
        uint32_t kb[4];
        const int mi = lane >> 3;
        const int key = jp * 16 + (lane & 7) + (mi >> 1) * 8;
        const int c = ks * 2 + (mi & 1);
        ldmatrix_x4(kb, swz(const_cast<__nv_bfloat16*>(tK), key, c));
        mma_bf16_16816(s[jp * 2], qf[ks], kb[0], kb[1]);
        mma_bf16_16816(s[jp * 2 + 1], qf[ks], kb[2], kb[3]);
      }
    }
    // mask keys beyond T (only the last tile)
    const int k0 = tile * kKvTile;
    if (k0 + kKvTile > T) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = k0 + j * 8 + (lane & 3) * 2;
        if (key >= T) s[j][0] = -INFINITY, s[j][2] = -INFINITY;
        if (key + 1 >= T) s[j][1] = -INFINITY, s[j][3] = -INFINITY;
      }
    }
    // online softmax; rows g = lane/4 (elements 0,1) and g+8 (elements 2,3)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float alpha[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float m_new = fmaxf(m_run[r], mx[r]);
      alpha[r] = exp2f((m_run[r] - m_new) * c_log2);
      m_run[r] = m_new;
      msc[r] = m_new * c_log2;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = exp2f(s[j][0] * c_log2 - msc[0]);
      s[j][1] = exp2f(s[j][1] * c_log2 - msc[0]);
      s[j][2] = exp2f(s[j][2] * c_log2 - msc[1]);
      s[j][3] = exp2f(s[j][3] * c_log2 - msc[1]);
      rs[0] += s[j][0] + s[j][1];
      rs[1] += s[j][2] + s[j][3];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * alpha[r] + rs[r];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o_acc[j][0] *= alpha[0], o_acc[j][1] *= alpha[0];
      o_acc[j][2] *= alpha[1], o_acc[j][3] *= alpha[1];
    }
    // O += P V   (k = 64 keys in 4 steps of 16; n = 64 dh in 8 tiles)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of dh n-tiles
        uint32_t vb[4];
        const int mi = lane >> 3;
        const int key = kk * 16 + (lane & 7) + (mi & 1) * 8;
        const int c = np * 2 + (mi >> 1);
        ldmatrix_x4_trans(vb, swz(const_cast<__nv_bfloat16*>(tV), key, c));
        mma_bf16_16816(o_acc[np * 2], pa, vb[0], vb[1]);
        mma_bf16_16816(o_acc[np * 2 + 1], pa, vb[2], vb[3]);
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled two iterations later
  }

  // finalise: row sums across the quad, normalise, store bf16
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int row0 = q0 + warp * 16 + (lane >> 2);
  __nv_bfloat16* ob = out + (long)b * T * d + h * 64 + (lane & 3) * 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (row0 < T) *reinterpret_cast<uint32_t*>(ob + (long)row0 * d + j * 8) = pack_bf16x2(o_acc[j][0] * inv0, o_acc[j][1] * inv0);
    if (row0 + 8 < T) *reinterpret_cast<uint32_t*>(ob + (long)(row0 + 8) * d + j * 8) = pack_bf16x2(o_acc[j][2] * inv1, o_acc[j][3] * inv1);
  }
}

}  // namespace

void launch_layernorm(const float* x, const float* gamma, const float* beta, __nv_bfloat16* y, int rows, int d, cudaStream_t stream) {
  if (d % 128 != 0 || d > 128 * kLnMaxVec) throw CudaError("layernorm: d must be a multiple of 128 and <= 1280");
  // a decoder step normalises only B rows: spread them over more CTAs and let the launch overlap its predecessor (PDL)
  const int warps_per_cta = rows <= 1024 ? 2 : 8;
  const dim3 grid((rows + warps_per_cta - 1) / warps_per_cta), block(warps_per_cta * 32);
  if (rows <= 1024) {
    launch_pdl(layernorm_kernel, grid, block, 0, stream, x, gamma, beta, y, rows, d);
  } else {
    layernorm_kernel<<<grid, block, 0, stream>>>(x, gamma, beta, y, rows, d);
    CUDA_CHECK(cudaGetLastError());
  }
}

void encoder_ops_set_attributes() {
  CUDA_CHECK(cudaFuncSetAttribute(encoder_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (kQTile * 64 + 4 * kKvTile * 64) * 2));
}

void launch_encoder_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int n_head, cudaStream_t stream) {
  const int smem = (kQTile * 64 + 4 * kKvTile * 64) * 2;  // 48 KB
  dim3 grid((T + kQTile - 1) / kQTile, n_head, B);
  encoder_attention_kernel<<<grid, kAttThreads, smem, stream>>>(qkv, out, T, n_head * 64);
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace b200w
