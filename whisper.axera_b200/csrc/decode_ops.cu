// K7: the HBM-bound kernels of one greedy decoder step (everything of the step that is not a GEMM).
//
// One step = TextDecoderTensorCache.forward of the reference (/root/reference/model_convert/export_onnx.py:312-387)
// plus the host glue of Whisper::run_decoder (/root/reference/cpp/src/Whisper.cpp:290-346), for B sequences at once:
//   embed            token_embedding[token] + positional_embedding[offset]              export_onnx.py:334-336
//   self attention   static 448-slot cache + current token, mask == "positions < offset"  :103-147 (mask value -60000
//                    underflows to exactly 0 in the fp32 softmax, so attending to positions 0..offset is identical);
//                    the row append the reference does on the host (Whisper.cpp:328-342) happens in the kernel
//   cross attention  1500 encoder keys, no mask                                           :216-230
//   argmax           first maximum (std::max_element, Whisper.cpp:42-45), EOT / loop bookkeeping (:214-222)
// K/V live in HBM as bf16, head-major [B][H][n][64]: one (sequence, head) is a contiguous 128-byte-per-key stream,
// read with 16-byte loads (8 lanes per key, 4 keys per warp instruction), fp32 scores / softmax / accumulation.
#include <algorithm>
#include <cfloat>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace b200w {
namespace {

__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// streaming load that marks its L2 lines evict-first: the 0.6 GB a cross-attention launch pulls through L2 is never reused,
// while the decoder weights the other micro-batch read ~100 us ago are needed again by this one
__device__ __forceinline__ uint64_t l2_evict_first_policy(bool evict_first) {
  uint64_t pol, normal;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(normal));
  return evict_first ? pol : normal;
}
__device__ __forceinline__ uint4 ld_stream16_ef(const void* p, uint64_t pol) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ float dot8(const uint4& kv, const float (&q)[8]) {
  float a = bf16lo_to_f32(kv.x) * q[0];
  a = fmaf(bf16hi_to_f32(kv.x), q[1], a);
  a = fmaf(bf16lo_to_f32(kv.y), q[2], a);
  a = fmaf(bf16hi_to_f32(kv.y), q[3], a);
  a = fmaf(bf16lo_to_f32(kv.z), q[4], a);
  a = fmaf(bf16hi_to_f32(kv.z), q[5], a);
  a = fmaf(bf16lo_to_f32(kv.w), q[6], a);
  a = fmaf(bf16hi_to_f32(kv.w), q[7], a);
  return a;
}
__device__ __forceinline__ void axpy8(float (&acc)[8], float p, const uint4& v) {
  acc[0] = fmaf(p, bf16lo_to_f32(v.x), acc[0]);
  acc[1] = fmaf(p, bf16hi_to_f32(v.x), acc[1]);
  acc[2] = fmaf(p, bf16lo_to_f32(v.y), acc[2]);
  acc[3] = fmaf(p, bf16hi_to_f32(v.y), acc[3]);
  acc[4] = fmaf(p, bf16lo_to_f32(v.z), acc[4]);
  acc[5] = fmaf(p, bf16hi_to_f32(v.z), acc[5]);
  acc[6] = fmaf(p, bf16lo_to_f32(v.w), acc[6]);
  acc[7] = fmaf(p, bf16hi_to_f32(v.w), acc[7]);
}

constexpr float kScoreScaleLog2 = 0.125f * 1.4426950408889634f;  // (64^-0.25)^2 * log2(e)

// ---- embedding -------------------------------------------------------------------------------------------
__global__ void embed_kernel(const int* __restrict__ step_ptr, const int* __restrict__ tokens, const int* __restrict__ slot_seq,
                             const float* __restrict__ tok_emb, const float* __restrict__ pos_emb, float* __restrict__ x, int d, int n_text_ctx) {
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.x;  // slot
  const int step = *step_ptr;
  const int tok = tokens[(long)slot_seq[b] * n_text_ctx + step];
  const float4* te = reinterpret_cast<const float4*>(tok_emb + (long)tok * d);
  const float4* pe = reinterpret_cast<const float4*>(pos_emb + (long)step * d);
  float4* xo = reinterpret_cast<float4*>(x + (long)b * d);
  for (int i = threadIdx.x; i < d / 4; i += blockDim.x) {
    const float4 a = te[i], p = pe[i];
    xo[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
  }
}

// ---- self attention ----------------------------------------------------------------------------------------
// One warp per (sequence, head): the cache of a head is at most 447 keys, so a whole CTA with block-wide reductions is
// all latency.  8 lanes share a key (16 bytes each), 4 keys per warp instruction; every key slot runs its own online
// softmax (max, sum, 8-wide accumulator per lane) and the four slots plus the current token (fp32 k1/v1) are merged
// with shuffles at the end.  No shared memory, no block barrier.
constexpr int kSelfWarps = 4;
__global__ void __launch_bounds__(kSelfWarps * 32) self_attention_decode_kernel(const float* __restrict__ qkv, __nv_bfloat16* __restrict__ k_cache,
                                                                               __nv_bfloat16* __restrict__ v_cache, const int* __restrict__ step_ptr,
                                                                               const int* __restrict__ slot_seq, __nv_bfloat16* __restrict__ out,
                                                                               int n_pairs, int n_head, int n_ctx) {
  pdl_wait();
  pdl_launch_dependents();
  const int pair = blockIdx.x * kSelfWarps + (threadIdx.x >> 5);  // b * n_head + h
  if (pair >= n_pairs) return;
  const int lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7;
  const int b = pair / n_head, h = pair - b * n_head;
  const int d = n_head * 64;
  const int pos = *step_ptr;  // cached positions 0..pos-1, current token at pos
  const float* qg = qkv + (long)b * 3 * d + h * 64 + sub * 8;
  const long cpair = (long)slot_seq[b] * n_head + h;  // the cache belongs to the sequence, the activations to the slot
  __nv_bfloat16* Kc = k_cache + cpair * n_ctx * 64;
  __nv_bfloat16* Vc = v_cache + cpair * n_ctx * 64;
  float q[8], k1[8], v1[8];
  {
    const float4 a = *reinterpret_cast<const float4*>(qg), c = *reinterpret_cast<const float4*>(qg + 4);
    q[0] = a.x, q[1] = a.y, q[2] = a.z, q[3] = a.w, q[4] = c.x, q[5] = c.y, q[6] = c.z, q[7] = c.w;
    const float4 ka = *reinterpret_cast<const float4*>(qg + d), kb = *reinterpret_cast<const float4*>(qg + d + 4);
    k1[0] = ka.x, k1[1] = ka.y, k1[2] = ka.z, k1[3] = ka.w, k1[4] = kb.x, k1[5] = kb.y, k1[6] = kb.z, k1[7] = kb.w;
    const float4 va = *reinterpret_cast<const float4*>(qg + 2 * d), vb = *reinterpret_cast<const float4*>(qg + 2 * d + 4);
    v1[0] = va.x, v1[1] = va.y, v1[2] = va.z, v1[3] = va.w, v1[4] = vb.x, v1[5] = vb.y, v1[6] = vb.z, v1[7] = vb.w;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] *= kScoreScaleLog2;
  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  constexpr int U = 4;
  for (int jb = 0; jb < pos; jb += 4 * U) {  // warp-uniform bounds
    uint4 kk[U], vv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = jb + u * 4 + grp;
      const bool ok = j < pos;
      kk[u] = ok ? ld_stream16(Kc + (long)j * 64 + sub * 8) : make_uint4(0, 0, 0, 0);
      vv[u] = ok ? ld_stream16(Vc + (long)j * 64 + sub * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = jb + u * 4 + grp;
      float sc = dot8(kk[u], q);
      sc += __shfl_xor_sync(0xffffffffu, sc, 1);
      sc += __shfl_xor_sync(0xffffffffu, sc, 2);
      sc += __shfl_xor_sync(0xffffffffu, sc, 4);
      if (j < pos) {
        const float m_new = fmaxf(m, sc);
        const float corr = exp2f(m - m_new), pw = exp2f(sc - m_new);
        l = fmaf(l, corr, pw);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] *= corr;
        axpy8(acc, pw, vv[u]);
        m = m_new;
      }
    }
  }
  // current token (fp32): its score, then merge the four key slots and the current token
  float s_cur = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s_cur = fmaf(q[i], k1[i], s_cur);
  s_cur += __shfl_xor_sync(0xffffffffu, s_cur, 1);
  s_cur += __shfl_xor_sync(0xffffffffu, s_cur, 2);
  s_cur += __shfl_xor_sync(0xffffffffu, s_cur, 4);
  float m_tot = fmaxf(m, s_cur);
  m_tot = fmaxf(m_tot, __shfl_xor_sync(0xffffffffu, m_tot, 8));
  m_tot = fmaxf(m_tot, __shfl_xor_sync(0xffffffffu, m_tot, 16));
  const float corr = exp2f(m - m_tot);  // exp2(-inf) = 0 for an empty slot
  l *= corr;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] *= corr;
  l += __shfl_xor_sync(0xffffffffu, l, 8);
  l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  const float pc = exp2f(s_cur - m_tot);
  l += pc;
  if (grp == 0) {
    const float inv = 1.f / l;
    uint4 o, kq, vq;
    o.x = pack_bf16x2(fmaf(pc, v1[0], acc[0]) * inv, fmaf(pc, v1[1], acc[1]) * inv);
    o.y = pack_bf16x2(fmaf(pc, v1[2], acc[2]) * inv, fmaf(pc, v1[3], acc[3]) * inv);
    o.z = pack_bf16x2(fmaf(pc, v1[4], acc[4]) * inv, fmaf(pc, v1[5], acc[5]) * inv);
    o.w = pack_bf16x2(fmaf(pc, v1[6], acc[6]) * inv, fmaf(pc, v1[7], acc[7]) * inv);
    *reinterpret_cast<uint4*>(out + (long)b * d + h * 64 + sub * 8) = o;
    // append this token's key / value (the reference does this on the host after the step, Whisper.cpp:328-342)
    kq.x = pack_bf16x2(k1[0], k1[1]), kq.y = pack_bf16x2(k1[2], k1[3]), kq.z = pack_bf16x2(k1[4], k1[5]), kq.w = pack_bf16x2(k1[6], k1[7]);
    vq.x = pack_bf16x2(v1[0], v1[1]), vq.y = pack_bf16x2(v1[2], v1[3]), vq.z = pack_bf16x2(v1[4], v1[5]), vq.w = pack_bf16x2(v1[6], v1[7]);
    *reinterpret_cast<uint4*>(Kc + (long)pos * 64 + sub * 8) = kq;
    *reinterpret_cast<uint4*>(Vc + (long)pos * 64 + sub * 8) = vq;
  }
}

// ---- cross attention ---------------------------------------------------------------------------------------
// Both kernels below evaluate one (sequence, head) item with EXACTLY the same arithmetic, so that a sequence's tokens do not
// depend on how many sequences share its batch (VERDICT r01 "output depends on the shard size"):
//   * 256 threads = 32 key classes (warp w, group g: keys = 4w + g mod 32) x 8 lanes (16 bytes of the 128-byte key each);
//   * the T keys are cut into SEGMENTS of 256 consecutive keys; inside a segment a thread owns 8 keys (slots u = 0..7, key
//     256 s + 32 u + 4 w + g) and accumulates p.v over them in slot order starting from zero;
//   * the softmax maximum is the exact maximum over all keys (order-free), p = exp2(score - max);
//   * a thread's total is the sum of its segment partials in segment order; totals are reduced over the 4 groups by an
//     xor-shuffle tree (8, 16), then over the 8 warps in warp order, then divided by the sum reduced the same way.
// The streaming kernel walks the segments of an item one after the other in one CTA; the split kernel gives each CTA of a
// thread-block cluster a contiguous range of segments, exchanges the maximum through distributed shared memory and lets
// cluster rank 0 add the per-thread segment partials in the same order -- bit-identical results, any batch size.
constexpr int kCrossThreads = 256;
constexpr int kXU = 8;                                        // key slots per thread per segment
constexpr int kXKeysPerStep = (kCrossThreads / 32) * 4 * kXU;  // 256 keys per segment
constexpr int kXMaxT = 1536;
constexpr int kXMaxSegPerCta = 3;                              // split kernel: segments per CTA (cluster of >= 2 CTAs)

// Small batches: cluster of n_split CTAs per (sequence, head), n_split dividing the segment count.
__global__ void __launch_bounds__(kCrossThreads)
cross_attention_split_kernel(const float* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
                             const int* __restrict__ slot_seq, __nv_bfloat16* __restrict__ out, int T, int n_head, int evict_first) {
  constexpr int NW = kCrossThreads / 32;
  __shared__ float s_scores[kXMaxSegPerCta * kXKeysPerStep];
  __shared__ __align__(16) float s_part[kXMaxSegPerCta][8][kCrossThreads];  // per-thread segment partials of p.v (read by rank 0)
  __shared__ float s_lpart[kXMaxSegPerCta][kCrossThreads / 8];              // per key-class segment partials of sum p
  __shared__ float s_acc[NW * 64];
  __shared__ float s_redm[NW], s_redl[NW];
  __shared__ float s_max;                                                   // this CTA's maximum (read by every rank)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 3, sub = lane & 7;
  const int nseg = (T + kXKeysPerStep - 1) / kXKeysPerStep;
  const uint32_t n_split = cluster_nctarank(), rank = cluster_ctarank();
  const int seg_per = nseg / (int)n_split;           // the launcher picks n_split dividing nseg
  const int seg0 = (int)rank * seg_per;
  const int item = blockIdx.x / n_split;             // (slot, head): q / out index
  const int slot = item / n_head;
  const long citem = (long)slot_seq[slot] * n_head + (item - slot * n_head);  // (sequence, head): cache index
  const int key0 = warp * 4 + grp;
  const uint64_t pol = l2_evict_first_policy(evict_first != 0);
  const __nv_bfloat16* kb = k + citem * T * 64 + sub * 8;
  const __nv_bfloat16* vb = v + citem * T * 64 + sub * 8;

  // K of the first segment is requested before the dependency wait: the cross K/V cache was written by the encoder long ago,
  // only q depends on the predecessor kernel
  uint4 kv[kXU];
  {
    const int j0 = seg0 * kXKeysPerStep + key0;
#pragma unroll
    for (int u = 0; u < kXU; ++u) kv[u] = (j0 + 32 * u < T) ? ld_stream16_ef(kb + (long)(j0 + 32 * u) * 64, pol) : make_uint4(0, 0, 0, 0);
  }
  pdl_wait();
  float qr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qr[i] = q[(long)item * 64 + sub * 8 + i] * kScoreScaleLog2;

  // ---- phase 1: scores of this CTA's segments, local maximum ----
  float mx = -INFINITY;
  for (int sl = 0; sl < seg_per; ++sl) {
    const int j0 = (seg0 + sl) * kXKeysPerStep + key0;
    const bool last = sl + 1 == seg_per;
    // next loads: the following K segment, or the first V segment
    const __nv_bfloat16* nb = last ? vb : kb;
    const int nj = (last ? seg0 : seg0 + sl + 1) * kXKeysPerStep + key0;
#pragma unroll
    for (int u = 0; u < kXU; ++u) {
      const int j = j0 + 32 * u;
      float sc = dot8(kv[u], qr);
      kv[u] = (nj + 32 * u < T) ? ld_stream16_ef(nb + (long)(nj + 32 * u) * 64, pol) : make_uint4(0, 0, 0, 0);
      sc += __shfl_xor_sync(0xffffffffu, sc, 1);
      sc += __shfl_xor_sync(0xffffffffu, sc, 2);
      sc += __shfl_xor_sync(0xffffffffu, sc, 4);
      if (j < T) {
        if (sub == 0) s_scores[sl * kXKeysPerStep + 32 * u + key0] = sc;
        mx = fmaxf(mx, sc);
      }
    }
  }
  mx = warp_max(mx);
  if (lane == 0) s_redm[warp] = mx;
  __syncthreads();
  if (tid == 0) {
    float m = s_redm[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) m = fmaxf(m, s_redm[w]);
    s_max = m;
  }
  cluster_sync_all();  // every rank's s_max is written (also orders s_scores / s_redm within the CTA)
  float m = -INFINITY;
  for (uint32_t r = 0; r < n_split; ++r) m = fmaxf(m, dsmem_ld_f32(dsmem_addr(&s_max, r)));  // exact, order-free

  // ---- phase 2: softmax weights and P.V, one partial per segment ----
  for (int sl = 0; sl < seg_per; ++sl) {
    const int j0 = (seg0 + sl) * kXKeysPerStep + key0;
    const bool last = sl + 1 == seg_per;
    const int nj = (seg0 + sl + 1) * kXKeysPerStep + key0;
    float seg[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) seg[i] = 0.f;
    float lseg = 0.f;
#pragma unroll
    for (int u = 0; u < kXU; ++u) {
      const int j = j0 + 32 * u;
      const float pw = j < T ? exp2f(s_scores[sl * kXKeysPerStep + 32 * u + key0] - m) : 0.f;
      axpy8(seg, pw, kv[u]);
      kv[u] = (!last && nj + 32 * u < T) ? ld_stream16_ef(vb + (long)(nj + 32 * u) * 64, pol) : make_uint4(0, 0, 0, 0);
      if (sub == 0) lseg += pw;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s_part[sl][i][tid] = seg[i];
    if (sub == 0) s_lpart[sl][tid >> 3] = lseg;
  }
  pdl_launch_dependents();
  cluster_sync_all();  // all partials of all ranks are in shared memory
  if (rank == 0) {
    // thread totals in segment order (rank r holds segments r * seg_per ..), exactly what the streaming kernel accumulates
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    float lsum = 0.f;
    for (uint32_t r = 0; r < n_split; ++r) {
      const uint32_t pa = dsmem_addr(&s_part[0][0][0], r), la = dsmem_addr(&s_lpart[0][0], r);
      for (int sl = 0; sl < seg_per; ++sl) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += dsmem_ld_f32(pa + (uint32_t)(((sl * 8 + i) * kCrossThreads + tid) * 4));
        if (sub == 0) lsum += dsmem_ld_f32(la + (uint32_t)((sl * (kCrossThreads / 8) + (tid >> 3)) * 4));
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
      acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
    }
    lsum = warp_sum(lsum);
    if (grp == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s_acc[warp * 64 + sub * 8 + i] = acc[i];
    }
    if (lane == 0) s_redl[warp] = lsum;
    __syncthreads();
    if (tid < 64) {
      float o = 0.f, l = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) o += s_acc[w * 64 + tid], l += s_redl[w];
      out[(long)item * 64 + tid] = __float2bfloat16_rn(o / l);  // out is [B][H*64]
    }
  }
  cluster_sync_all();  // nobody leaves (and frees its shared memory) before rank 0 has read the partials
}

// Streaming variant for large batches.  A single resident wave of CTAs (at most three per SM, 64 registers) each works
// through several (sequence, head) items back to back and keeps its 16-byte loads rolling across the phase and item
// boundaries: a slot is refilled with the load of the NEXT step (next K block, first V block, first K block of the next
// item) the moment it has been consumed, so the reductions between the phases never drain the memory pipeline.  Because
// the whole grid is dispatched at once and leaves ~16 K registers and most of the shared memory of every SM free, the
// short kernels of the other micro-batch (192-thread tcgen05 GEMM CTAs, LayerNorm, self attention) are placed next to
// it immediately instead of queueing behind undispatched CTAs.
__global__ void __maxnreg__(64)
cross_attention_stream_kernel(const float* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
                              const int* __restrict__ slot_seq, __nv_bfloat16* __restrict__ out, int T, int n_head, int n_items,
                              int* __restrict__ work /*[2]: next item, CTAs done*/, int evict_first) {
  constexpr int NW = kCrossThreads / 32;
  __shared__ float s_scores[kXMaxT];
  __shared__ float s_acc[NW * 64];
  __shared__ float s_redm[NW], s_redl[NW];
  __shared__ float s_q[2][64];
  __shared__ int s_item[2], s_citem[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 3, sub = lane & 7;  // 4 keys per warp instruction, 8 lanes (16 B each) per key
  const int nk = (T + kXKeysPerStep - 1) / kXKeysPerStep;  // steps per phase
  const int spi = 2 * nk;                                  // steps per item
  const int key0 = warp * 4 + grp;                         // key of slot 0 within a step; slot u adds 32 u
  const uint64_t pol = l2_evict_first_policy(evict_first != 0);
  pdl_wait();
  // items ((sequence, head) pairs in the cache's own order) are handed out by an atomic counter: every SM keeps exactly
  // the CTAs it was given busy until the work runs out, whatever B * H is
  // items are (slot, head) pairs (q / out index); the cache is indexed by (sequence, head): resolved once per claim by the
  // claiming thread and published next to the item (no dependent global load on the streaming path)
  auto cache_item = [&](int it) {
    if (it >= n_items) return 0;
    const int slot = it / n_head;
    return slot_seq[slot] * n_head + (it - slot * n_head);
  };
  if (tid == 0) {
    s_item[0] = atomicAdd(&work[0], 1);
    s_citem[0] = cache_item(s_item[0]);
  }
  __syncthreads();
  int item = s_item[0];
  int citem = s_citem[0];
  int par = 0;

  // slot u of step (item, st): tensor K for st < nk else V, key (st % nk) * 256 + 32 u + key0
  auto step_ptr = [&](int cit, int st) {
    const __nv_bfloat16* base = (st < nk ? k : v) + (long)cit * T * 64;
    const int kb = (st < nk ? st : st - nk) * kXKeysPerStep + key0;
    return base + (long)kb * 64 + sub * 8;
  };
  auto step_key = [&](int st) { return (st < nk ? st : st - nk) * kXKeysPerStep + key0; };

  if (item < n_items) {
    uint4 kv[kXU];
    {
      const __nv_bfloat16* p0 = step_ptr(citem, 0);
      const int j0 = step_key(0);
#pragma unroll
      for (int u = 0; u < kXU; ++u) kv[u] = (j0 + 32 * u < T) ? ld_stream16_ef(p0 + u * 32 * 64, pol) : make_uint4(0, 0, 0, 0);
    }
    float qr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) qr[i] = q[(long)item * 64 + sub * 8 + i] * kScoreScaleLog2;  // q is [B][H*64]: item * 64

    int next_item = n_items;
    for (;;) {
      // ---- K phase: scores (log2 domain) into smem, running maximum.  The step after each one -- next K segment, then the
      // first V segment -- is requested slot by slot while the current one is consumed ----
      float mx = -INFINITY;
      for (int st = 0; st < nk; ++st) {
        const __nv_bfloat16* np = step_ptr(citem, st + 1);
        const int nj = step_key(st + 1);
        const int j0 = step_key(st);
#pragma unroll
        for (int u = 0; u < kXU; ++u) {
          const int j = j0 + 32 * u;
          float sc = dot8(kv[u], qr);
          kv[u] = (nj + 32 * u < T) ? ld_stream16_ef(np + u * 32 * 64, pol) : make_uint4(0, 0, 0, 0);
          sc += __shfl_xor_sync(0xffffffffu, sc, 1);
          sc += __shfl_xor_sync(0xffffffffu, sc, 2);
          sc += __shfl_xor_sync(0xffffffffu, sc, 4);
          if (j < T) {
            if (sub == 0) s_scores[j] = sc;
            mx = fmaxf(mx, sc);
          }
        }
      }
      // phase boundary: block maximum (the first V loads are already in flight)
      mx = warp_max(mx);
      if (lane == 0) s_redm[warp] = mx;
      if (tid == 0) {  // claim the next item while V streams
        s_item[par ^ 1] = atomicAdd(&work[0], 1);
        s_citem[par ^ 1] = cache_item(s_item[par ^ 1]);
      }
      __syncthreads();
      float m = s_redm[0];
#pragma unroll
      for (int w = 1; w < NW; ++w) m = fmaxf(m, s_redm[w]);
      next_item = s_item[par ^ 1];
      // the next item's query goes to shared memory asynchronously while V streams (the registers hold the segment partials
      // now); it is picked up when that item's K phase starts
      if (next_item < n_items && tid < 64) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&s_q[par ^ 1][tid])), "l"(q + (long)next_item * 64 + tid) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      // ---- V phase: softmax weights and P.V; every 256-key segment is ONE partial added to the thread's total (the canonical
      // summation order shared with cross_attention_split_kernel) ----
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      float lsum = 0.f;
      for (int st = nk; st < spi; ++st) {
        const bool same = st + 1 < spi;
        const bool has_next = same || next_item < n_items;
        const __nv_bfloat16* np = has_next ? step_ptr(same ? citem : s_citem[par ^ 1], same ? st + 1 : 0) : k;
        const int nj = has_next ? step_key(same ? st + 1 : 0) : T;  // T: every slot predicated off
        const int j0 = step_key(st);
        float seg[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) seg[i] = 0.f;
        float lseg = 0.f;
#pragma unroll
        for (int u = 0; u < kXU; ++u) {
          const int j = j0 + 32 * u;
          const float pw = j < T ? exp2f(s_scores[j] - m) : 0.f;
          axpy8(seg, pw, kv[u]);
          kv[u] = (nj + 32 * u < T) ? ld_stream16_ef(np + u * 32 * 64, pol) : make_uint4(0, 0, 0, 0);
          if (sub == 0) lseg += pw;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += seg[i];
        lsum += lseg;
      }
      // item done: reduce over key groups and warps, write the head's output (the next item's K loads are in flight)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
      }
      lsum = warp_sum(lsum);
      if (grp == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s_acc[warp * 64 + sub * 8 + i] = acc[i];
      }
      if (lane == 0) s_redl[warp] = lsum;
      asm volatile("cp.async.wait_all;" ::: "memory");  // this thread's piece of the next query has landed
      __syncthreads();
      if (tid < 64) {
        float o = 0.f, l = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) o += s_acc[w * 64 + tid], l += s_redl[w];
        out[(long)item * 64 + tid] = __float2bfloat16_rn(o / l);  // out is [B][H*64]
      }
      if (next_item < n_items) {
#pragma unroll
        for (int i = 0; i < 8; ++i) qr[i] = s_q[par ^ 1][sub * 8 + i] * kScoreScaleLog2;
      }
      __syncthreads();  // s_acc / s_redl / s_scores are rewritten by the next item
      if (next_item >= n_items) break;
      item = next_item;
      par ^= 1;
      citem = s_citem[par];
    }
  }
  pdl_launch_dependents();
  // the last CTA to leave re-arms the counters for the next launch (every CTA has made its final claim by then)
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&work[1], 1) == (int)gridDim.x - 1) {
      work[0] = 0;
      work[1] = 0;
      __threadfence();
    }
  }
}

// ---- argmax finalize + loop bookkeeping --------------------------------------------------------------------
__global__ void __launch_bounds__(128) argmax_finalize_kernel(DecodeState st, const float* __restrict__ part_val, const int* __restrict__ part_idx,
                                                              int n_tiles, int part_ld, int n_text_ctx, int eot, int honor_eot, int sot_len) {
  __shared__ float s_v[4];
  __shared__ int s_i[4];
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.x;
  float best = -FLT_MAX;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n_tiles; i += blockDim.x) {
    const float v = part_val[(long)b * part_ld + i];
    const int ix = part_idx[(long)b * part_ld + i];
    if (v > best || (v == best && ix < bi)) best = v, bi = ix;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
  }
  if ((threadIdx.x & 31) == 0) s_v[threadIdx.x >> 5] = best, s_i[threadIdx.x >> 5] = bi;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 4; ++w)
      if (s_v[w] > best || (s_v[w] == best && s_i[w] < bi)) best = s_v[w], bi = s_i[w];
    const int pos = *st.step;
    const long seq = st.slot_seq[b];
    st.out_tokens[seq * n_text_ctx + pos] = bi;
    if (pos + 1 >= sot_len && pos + 1 < n_text_ctx) {
      const int f = st.forced ? st.forced[seq * n_text_ctx + pos + 1] : -1;
      st.tokens[seq * n_text_ctx + pos + 1] = f >= 0 ? f : bi;
    }
    if (honor_eot && pos >= sot_len - 1 && bi == eot) st.finished[seq] = 1;
  }
}

// ---- K/V cache import / export (model-ABI boundary: the reference's decoder takes and returns f32 [B][T][d] caches) ----
// f32 [n_seq][T_src][d] (token-major, the reference layout) <-> bf16 head-major [n_seq][H][T_dst][64] (resident layout);
// rows [0, n_rows) of every sequence are converted.  One thread per 8 consecutive features of one (sequence, row, head).
__global__ void kv_import_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int n_seq, int n_rows, int T_src, int T_dst, int n_head) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;  // ((seq * n_rows + t) * n_head + h) * 8 + part
  const long total = (long)n_seq * n_rows * n_head * 8;
  if (i >= total) return;
  const int part = (int)(i & 7);
  long r = i >> 3;
  const int h = (int)(r % n_head);
  r /= n_head;
  const int t = (int)(r % n_rows), b = (int)(r / n_rows);
  const float* s = src + ((long)b * T_src + t) * n_head * 64 + h * 64 + part * 8;
  const float4 a = *reinterpret_cast<const float4*>(s), c = *reinterpret_cast<const float4*>(s + 4);
  uint4 o;
  o.x = pack_bf16x2(a.x, a.y), o.y = pack_bf16x2(a.z, a.w), o.z = pack_bf16x2(c.x, c.y), o.w = pack_bf16x2(c.z, c.w);
  *reinterpret_cast<uint4*>(dst + (((long)b * n_head + h) * T_dst + t) * 64 + part * 8) = o;
}
__global__ void kv_export_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int n_seq, int n_rows, int T_src, int T_dst, int n_head) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)n_seq * n_rows * n_head * 8;
  if (i >= total) return;
  const int part = (int)(i & 7);
  long r = i >> 3;
  const int h = (int)(r % n_head);
  r /= n_head;
  const int t = (int)(r % n_rows), b = (int)(r / n_rows);
  const uint4 u = *reinterpret_cast<const uint4*>(src + (((long)b * n_head + h) * T_src + t) * 64 + part * 8);
  float* o = dst + ((long)b * T_dst + t) * n_head * 64 + h * 64 + part * 8;
  *reinterpret_cast<float4*>(o) = make_float4(bf16lo_to_f32(u.x), bf16hi_to_f32(u.x), bf16lo_to_f32(u.y), bf16hi_to_f32(u.y));
  *reinterpret_cast<float4*>(o + 4) = make_float4(bf16lo_to_f32(u.z), bf16hi_to_f32(u.z), bf16lo_to_f32(u.w), bf16hi_to_f32(u.w));
}

// ---- step boundary: argmax finalize + loop bookkeeping + NEXT step's embedding + first LayerNorm + step counter ----------
// One kernel instead of four dependent ones (argmax_finalize, advance_step, embed, layernorm): at mid-size batches the decoder
// step is bound by its chain of short kernels, not by HBM.  One CTA per slot; the last CTA to finish advances the step counter
// (every CTA has read it by then).
constexpr int kBoundaryThreads = 128;
constexpr int kBoundaryMaxVec = 3;  // float4 per thread: d <= 1536
__global__ void __launch_bounds__(kBoundaryThreads)
step_boundary_kernel(DecodeState st, const float* __restrict__ part_val, const int* __restrict__ part_idx, int n_tiles, int part_ld, int n_text_ctx,
                     int eot, int honor_eot, int sot_len, const float* __restrict__ tok_emb, const float* __restrict__ pos_emb,
                     const float* __restrict__ ln_g, const float* __restrict__ ln_b, float* __restrict__ x, __nv_bfloat16* __restrict__ h, int d,
                     int* __restrict__ ticket) {
  __shared__ float s_v[4];
  __shared__ int s_i[4];
  __shared__ int s_next[2];
  __shared__ float s_red[4];
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float best = -FLT_MAX;
  int bi = 0x7fffffff;
  for (int i = tid; i < n_tiles; i += kBoundaryThreads) {
    const float v = part_val[(long)b * part_ld + i];
    const int ix = part_idx[(long)b * part_ld + i];
    if (v > best || (v == best && ix < bi)) best = v, bi = ix;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
  }
  if (lane == 0) s_v[warp] = best, s_i[warp] = bi;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 4; ++w)
      if (s_v[w] > best || (s_v[w] == best && s_i[w] < bi)) best = s_v[w], bi = s_i[w];
    const int pos = *st.step;
    const long seq = st.slot_seq[b];
    st.out_tokens[seq * n_text_ctx + pos] = bi;
    int next = -1;
    if (pos + 1 < n_text_ctx) {
      if (pos + 1 >= sot_len) {
        const int f = st.forced ? st.forced[seq * n_text_ctx + pos + 1] : -1;
        next = f >= 0 ? f : bi;
        st.tokens[seq * n_text_ctx + pos + 1] = next;
      } else {
        next = st.tokens[seq * n_text_ctx + pos + 1];  // still inside the SOT prefix
      }
    }
    if (honor_eot && pos >= sot_len - 1 && bi == eot) st.finished[seq] = 1;
    s_next[0] = next;
    s_next[1] = pos + 1;
  }
  __syncthreads();
  const int tok = s_next[0], npos = s_next[1];
  if (tok >= 0) {  // block-uniform
    // x = token_embedding[tok] + positional_embedding[npos]; h = LayerNorm(x) (two-pass statistics like layernorm_kernel)
    const int nvec = d >> 2;
    const float4* te = reinterpret_cast<const float4*>(tok_emb + (long)tok * d);
    const float4* pe = reinterpret_cast<const float4*>(pos_emb + (long)npos * d);
    float4* xo = reinterpret_cast<float4*>(x + (long)b * d);
    float4 v[kBoundaryMaxVec];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kBoundaryMaxVec; ++k) {
      const int i = tid + k * kBoundaryThreads;
      if (i < nvec) {
        const float4 a = te[i], p = pe[i];
        v[k] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
        xo[i] = v[k];
        s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
      }
    }
    s = warp_sum(s);
    if (lane == 0) s_red[warp] = s;
    __syncthreads();
    const float mean = ((s_red[0] + s_red[1]) + (s_red[2] + s_red[3])) / (float)d;
    __syncthreads();
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < kBoundaryMaxVec; ++k) {
      const int i = tid + k * kBoundaryThreads;
      if (i < nvec) {
        const float a = v[k].x - mean, c = v[k].y - mean, e = v[k].z - mean, f = v[k].w - mean;
        q += (a * a + c * c) + (e * e + f * f);
      }
    }
    q = warp_sum(q);
    if (lane == 0) s_red[warp] = q;
    __syncthreads();
    const float rstd = rsqrtf(((s_red[0] + s_red[1]) + (s_red[2] + s_red[3])) / (float)d + 1e-5f);
    const float4* g4 = reinterpret_cast<const float4*>(ln_g);
    const float4* b4 = reinterpret_cast<const float4*>(ln_b);
    uint2* ho = reinterpret_cast<uint2*>(h + (long)b * d);
#pragma unroll
    for (int k = 0; k < kBoundaryMaxVec; ++k) {
      const int i = tid + k * kBoundaryThreads;
      if (i < nvec) {
        const float4 g = g4[i], bb = b4[i];
        uint2 o;
        o.x = pack_bf16x2((v[k].x - mean) * rstd * g.x + bb.x, (v[k].y - mean) * rstd * g.y + bb.y);
        o.y = pack_bf16x2((v[k].z - mean) * rstd * g.z + bb.z, (v[k].w - mean) * rstd * g.w + bb.w);
        ho[i] = o;
      }
    }
  }
  // the last CTA advances the step counter: every CTA took its ticket after reading the counter
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {
      *st.step = npos;
      *ticket = 0;
      __threadfence();
    }
  }
}

__global__ void advance_step_kernel(int* step) {
  pdl_wait();
  pdl_launch_dependents();
  *step += 1;
}

}  // namespace

bool pdl_enabled() {
  static const bool on = getenv("B200W_NO_PDL") == nullptr;
  return on;
}

int& launch_priority() {
  static thread_local int prio = 0;
  return prio;
}

void launch_embed(const DecodeState& st, const float* tok_emb, const float* pos_emb, float* x, int B, int d, int n_text_ctx,
                  cudaStream_t stream, bool pdl) {
  launch_k(pdl, embed_kernel, dim3(B), dim3(128), 0, stream, st.step, st.tokens, st.slot_seq, tok_emb, pos_emb, x, d, n_text_ctx);
}

void launch_self_attention_decode(const float* qkv, __nv_bfloat16* k_cache, __nv_bfloat16* v_cache, const int* step, const int* slot_seq,
                                  __nv_bfloat16* out, int B, int n_head, int n_ctx, cudaStream_t stream) {
  const int n_pairs = B * n_head;
  launch_pdl(self_attention_decode_kernel, dim3((n_pairs + kSelfWarps - 1) / kSelfWarps), dim3(kSelfWarps * 32), 0, stream, qkv, k_cache,
             v_cache, step, slot_seq, out, n_pairs, n_head, n_ctx);
}

int cross_attention_pick_split(int B, int n_head, int T) {
  // 0 = streaming kernel (one resident wave, items claimed dynamically); n > 0 = cluster of n CTAs per item: the smallest
  // split that puts at least two CTAs on every SM (measured: 2 beats 3 and 6 at 128-384 items), the largest one when even
  // that does not fill the machine (a handful of sequences).  The arithmetic is the same for every choice.
  static const int forced = getenv("B200W_CROSS_SPLIT") ? atoi(getenv("B200W_CROSS_SPLIT")) : -1;
  static const bool no_stream = getenv("B200W_NO_CROSS_STREAM") != nullptr;
  const int nseg = (T + kXKeysPerStep - 1) / kXKeysPerStep;
  const int n_items = B * n_head;
  if (T > kXMaxT) throw CudaError("cross attention: more keys than the kernels are sized for");
  auto valid = [&](int n) { return n >= 1 && n <= 8 && nseg % n == 0 && nseg / n <= kXMaxSegPerCta; };  // portable cluster sizes only
  if (forced == 0) return 0;
  if (forced > 0 && valid(forced)) return forced;
  if (!no_stream && n_items >= 2 * kNumSMs) return 0;
  int best = 0;
  for (int n = 1; n <= 8; ++n) {
    if (!valid(n)) continue;
    best = n;
    if (n_items * n >= 2 * kNumSMs) break;
  }
  return best;  // 0 only if no valid split exists (then the streaming kernel runs whatever the batch)
}

void launch_cross_attention_decode(const float* q, const __nv_bfloat16* k, const __nv_bfloat16* v, const int* slot_seq, __nv_bfloat16* out,
                                   int B, int n_head, int T, int n_split, cudaStream_t stream, bool pdl, int* work) {
  const int n_items = B * n_head;
  static const int evict_first = getenv("B200W_NO_EVICT_FIRST") == nullptr;
  if (n_split <= 0) {
    if (work == nullptr) throw CudaError("cross attention: the streaming kernel needs its work counters");
    // single resident wave: three CTAs on every SM, items claimed dynamically
    static const int ctas_per_sm = getenv("B200W_CROSS_CTAS_PER_SM") ? atoi(getenv("B200W_CROSS_CTAS_PER_SM")) : 3;
    const int grid = std::min(n_items, ctas_per_sm * kNumSMs);
    launch_k(pdl, cross_attention_stream_kernel, dim3(grid), dim3(kCrossThreads), 0, stream, q, k, v, slot_seq, out, T, n_head, n_items, work, evict_first);
    return;
  }
  launch_kc(pdl, n_split, cross_attention_split_kernel, dim3(n_items * n_split), dim3(kCrossThreads), 0, stream, q, k, v, slot_seq, out, T,
            n_head, n_split == 1 ? 0 : evict_first);
}

void launch_argmax_finalize(const DecodeState& st, const float* part_val, const int* part_idx, int n_tiles, int part_ld, int B,
                            int n_text_ctx, int eot, int honor_eot, int sot_len, cudaStream_t stream) {
  launch_pdl(argmax_finalize_kernel, dim3(B), dim3(128), 0, stream, st, part_val, part_idx, n_tiles, part_ld, n_text_ctx, eot, honor_eot, sot_len);
}

void launch_kv_import(const float* src, __nv_bfloat16* dst, int n_seq, int n_rows, int T_src, int T_dst, int n_head, cudaStream_t stream) {
  const long total = (long)n_seq * n_rows * n_head * 8;
  if (total <= 0) return;
  kv_import_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src, dst, n_seq, n_rows, T_src, T_dst, n_head);
  CUDA_CHECK(cudaGetLastError());
}
void launch_kv_export(const __nv_bfloat16* src, float* dst, int n_seq, int n_rows, int T_src, int T_dst, int n_head, cudaStream_t stream) {
  const long total = (long)n_seq * n_rows * n_head * 8;
  if (total <= 0) return;
  kv_export_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src, dst, n_seq, n_rows, T_src, T_dst, n_head);
  CUDA_CHECK(cudaGetLastError());
}

void launch_step_boundary(const DecodeState& st, const float* part_val, const int* part_idx, int n_tiles, int part_ld, int B, int n_text_ctx,
                          int eot, int honor_eot, int sot_len, const float* tok_emb, const float* pos_emb, const float* ln_g, const float* ln_b,
                          float* x, __nv_bfloat16* h, int d, int* ticket, cudaStream_t stream) {
  if (d % 4 != 0 || d / 4 > kBoundaryThreads * kBoundaryMaxVec) throw CudaError("step boundary: unsupported model width");
  launch_pdl(step_boundary_kernel, dim3(B), dim3(kBoundaryThreads), 0, stream, st, part_val, part_idx, n_tiles, part_ld, n_text_ctx, eot, honor_eot,
             sot_len, tok_emb, pos_emb, ln_g, ln_b, x, h, d, ticket);
}

void launch_advance_step(int* step, cudaStream_t stream, bool pdl) { launch_k(pdl, advance_step_kernel, dim3(1), dim3(1), 0, stream, step); }

}  // namespace b200w
