// Host-side launchers of the sm_100a kernels (definitions in *.cu).  Internal to the engine; the public
// boundary is include/ax_whisper_api.h and include/b200w_model_abi.h.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace b200w {

// one-time per-process kernel attribute setup (dynamic shared memory opt-in); call before any launch / graph capture
void gemm_set_attributes();
void logmel_set_attributes();
void encoder_ops_set_attributes();
void attention_tcgen05_set_attributes();
inline void kernels_set_attributes() {
  gemm_set_attributes();
  logmel_set_attributes();
  encoder_ops_set_attributes();
  attention_tcgen05_set_attributes();
}

// ---- K1 log-mel (logmel.cu) -----------------------------------------------------------------------
void logmel_upload_tables();
size_t logmel_mel_table_copy(int n_mels, float* dense_bank, float* window400);
// pcm [B][pcm_stride] f32 (device), n_samples_dev [B]; out_mel [B][n_mels][3000] f32; out_tm bf16
// [B][3002][n_mels] (may be null); utt_state [2 B + 1] int scratch (per utterance: running maximum key, arrivals; + the block-ticket counter).  One memset + ONE kernel.
void launch_logmel(const float* pcm, long pcm_stride, const int* n_samples_dev, int max_samples, int B, int n_mels,
                   float* out_mel, __nv_bfloat16* out_tm, int* utt_state, cudaStream_t stream);
void launch_mel_to_timemajor(const float* mel, int B, int n_mels, __nv_bfloat16* out_tm, cudaStream_t stream);

// ---- K4 tcgen05 GEMM (gemm_tcgen05.cu) ------------------------------------------------------------
// C[m][n] = sum_k A[m][k] * W[n][k]   (A and W bf16, K-major; fp32 accumulation in TMEM)
enum GemmEpilogue : int {
  EPI_BIAS_BF16 = 0,       // out bf16 = acc + bias
  EPI_BIAS_GELU_BF16 = 1,  // out bf16 = gelu(acc + bias)
  EPI_BIAS_F32 = 2,        // out f32  = acc + bias
  EPI_BIAS_RESID_F32 = 3,  // out f32 += acc + bias         (residual stream, in place)
  EPI_GELU_POS_F32 = 4,    // out f32  = gelu(acc + bias) + pos[row_in_batch][n]   (conv2 + positional embedding)
  EPI_CROSSKV_BF16 = 5,    // head-major bf16 scatter of the stacked cross-attention K/V projections
  EPI_ARGMAX = 6,          // per-row (max, first index) over this tile's columns -> partials (+ optional f32 logits)
};

struct GemmOperandA {  // activation operand, viewed as [n_batch][rows][K] with arbitrary (16 B-multiple) pitches
  const __nv_bfloat16* ptr;
  int K;                 // inner extent (elements)
  int rows;              // rows per batch visible to TMA (>= rows_valid; rows beyond the tensor are zero-filled)
  int n_batch;
  long row_pitch;        // elements between consecutive rows
  long batch_pitch;      // elements between batches
  // implicit-GEMM convolution taps (n_taps == 0: plain GEMM over K). Tap t multiplies A[row + tap_row[t]]
  // [tap_c0[t] .. tap_c0[t] + k_per_tap) with W[:, t * k_per_tap ..).
  int n_taps;
  int k_per_tap;
  int tap_c0[3];
  int tap_row[3];
};

struct GemmParams {
  int rows_valid;        // rows per batch that are real (stores are masked beyond)
  int N;                 // valid output columns (stores are masked beyond)
  int n_rows_w;          // rows of W visible to TMA (>= N)
  void* out;             // bf16 or f32 depending on the epilogue
  long ldo;              // output leading dimension (elements)
  long out_batch_pitch;  // output elements between batches
  int out_row_offset;    // added to the row index within a batch (conv1 writes into a padded buffer)
  int a_batch_offset;    // added to the batch coordinate of A (sub-batches of a larger resident tensor)
  int a_row_offset;      // added to the row coordinate of A (a launch over rows [a_row_offset, a_row_offset + rows_valid))
  int n_batch;           // batches in this launch (<= the plan's n_batch)
  int use_pdl;           // launch with programmatic dependent launch (decoder-step GEMMs)
  const float* bias;     // [N] or null
  const float* pos;      // EPI_GELU_POS_F32: [rows_valid][N] f32
  // EPI_CROSSKV_BF16: n = (layer*2 + kv)*d + h*64 + dh ; row = (batch b, t)
  __nv_bfloat16* cross_k;  // [L][B][H][T][64]
  __nv_bfloat16* cross_v;
  int d_model, n_head, n_ctx_kv, kv_batch, kv_batch_offset;
  // EPI_ARGMAX
  float* part_val;       // [rows][n_tiles]
  int* part_idx;
  int part_ld;
};

struct GemmPlan;  // opaque: tensor maps + launch geometry, built once per (operand, shape)
// Output tensor of a two-CTA GEMM whose epilogue leaves through TMA (plain [rows][cols] tensor; for EPI_CROSSKV_BF16 the two
// head-major caches viewed as [n_slices][rows = T][64]).
struct GemmTmaOut {
  void* ptr;
  void* ptr2;
  long rows, cols, ld;
  long n_slices;
};
// two_cta: 256 x 256 tiles computed by CTA pairs (tcgen05.mma.cta_group::2, gemm2cta_tcgen05.cu); for the large-M encoder GEMMs
GemmPlan* gemm_plan_create(const GemmOperandA& a, const __nv_bfloat16* w, int n_rows_w, int block_n, int epilogue, bool two_cta = false,
                           const GemmTmaOut* out = nullptr);
void gemm_plan_destroy(GemmPlan*);
void gemm_launch(const GemmPlan* plan, const GemmParams& p, cudaStream_t stream);
// plain SIMT comparator used by the self-tests only (same operand conventions, f32 output = acc + bias)
void gemm_reference_simt(const __nv_bfloat16* a, long lda, const __nv_bfloat16* w, long ldw, const float* bias, float* out,
                         long ldo, int M, int N, int K, cudaStream_t stream);

// bf16 tiled tensor map with SWIZZLE_128B (rank 2 or 3; dims / box innermost first; pitches in bytes for dims 1..)
CUtensorMap make_tmap_bf16_sw128(const void* base, int rank, const uint64_t* dims, const uint64_t* pitches_bytes, const uint32_t* box);

// ---- encoder ops (encoder_ops.cu) -----------------------------------------------------------------
// LayerNorm (eps 1e-5, fp32 statistics): x f32 [rows][d] -> y bf16 [rows][d]
void launch_layernorm(const float* x, const float* gamma, const float* beta, __nv_bfloat16* y, int rows, int d, cudaStream_t stream);
// non-causal multi-head attention over T keys, head_dim 64: qkv bf16 [B*T][3d] -> out bf16 [B*T][d]
void launch_encoder_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int n_head, cudaStream_t stream);  // mma.sync comparator
// same contract on tcgen05 tensor cores (attention_tcgen05.cu); this is the product path
void launch_encoder_attention_tcgen05(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int n_head, cudaStream_t stream);

// ---- decode ops (decode_ops.cu) -------------------------------------------------------------------
struct DecodeState {
  // device-resident control block shared by the per-step kernels (one CUDA graph serves every step)
  int* step;          // [1] current offset (position of the token being consumed)
  int* tokens;        // [B][n_text_ctx] token fed at each position (SOT prefix + generated)
  int* forced;        // [B][n_text_ctx] teacher-forcing tokens or -1
  int* finished;      // [B] 1 once EOT was produced (honoured only when honor_eot)
  int* out_tokens;    // [B][n_text_ctx] generated tokens (argmax results)
  // Decoder rows ("slots") are decoupled from sequences: slot i of the step works on sequence slot_seq[i].  Activations are
  // per slot, everything that persists across steps (token tables, finished flags, self / cross K/V caches) is per sequence,
  // so dropping finished sequences (EOT) is just a shorter slot list -- nothing is moved in HBM.
  const int* slot_seq;  // [n_slots]
};
// x[b] = tok_emb[token[b][step]] + pos_emb[step]      (f32)
void launch_embed(const DecodeState& st, const float* tok_emb, const float* pos_emb, float* x, int B, int d, int n_text_ctx,
                  cudaStream_t stream, bool pdl = true);
// self attention for one new token per slot over a bf16 head-major cache [n_seq][H][n_ctx][64] (sequence = slot_seq[slot]);
// qkv f32 [B][3d] (q | k | v of the current token). Appends k,v at position *step, writes out bf16 [B][d].
void launch_self_attention_decode(const float* qkv, __nv_bfloat16* k_cache, __nv_bfloat16* v_cache, const int* step, const int* slot_seq,
                                  __nv_bfloat16* out, int B, int n_head, int n_ctx, cudaStream_t stream);
// cross attention: q f32 [B][d] (per slot) over bf16 head-major K/V [n_seq][H][T][64] (sequence = slot_seq[slot]); out bf16 [B][d].
// n_split from cross_attention_pick_split(): 0 = streaming kernel (needs `work`: two zero-initialised ints owned by this launch
// site), n > 0 = thread-block cluster of n CTAs per (sequence, head).  Every variant computes bit-identical results.
void launch_cross_attention_decode(const float* q, const __nv_bfloat16* k, const __nv_bfloat16* v, const int* slot_seq, __nv_bfloat16* out,
                                   int B, int n_head, int T, int n_split, cudaStream_t stream, bool pdl = true, int* work = nullptr);
// reduces the argmax partials, applies teacher forcing / EOT bookkeeping, stores the next token
void launch_advance_step(int* step, cudaStream_t stream, bool pdl = true);  // *step += 1 (once per decoder step)
void launch_argmax_finalize(const DecodeState& st, const float* part_val, const int* part_idx, int n_tiles, int part_ld, int B,
                            int n_text_ctx, int eot, int honor_eot, int sot_len, cudaStream_t stream);
// step boundary = argmax_finalize + advance_step + embed of the NEXT position + its first LayerNorm in one kernel (one CTA per slot;
// ticket: one zero-initialised int owned by this launch site).  x f32 [B][d], h bf16 [B][d] = LayerNorm(x) with ln_g / ln_b.
void launch_step_boundary(const DecodeState& st, const float* part_val, const int* part_idx, int n_tiles, int part_ld, int B, int n_text_ctx,
                          int eot, int honor_eot, int sot_len, const float* tok_emb, const float* pos_emb, const float* ln_g, const float* ln_b,
                          float* x, __nv_bfloat16* h, int d, int* ticket, cudaStream_t stream);
int cross_attention_pick_split(int B, int n_head, int T);
// K/V cache import / export at the model-ABI boundary: f32 token-major [n_seq][T][H*64] (the reference's tensors) <-> the
// resident bf16 head-major [n_seq][H][T][64]; rows [0, n_rows) of every sequence.
void launch_kv_import(const float* src, __nv_bfloat16* dst, int n_seq, int n_rows, int T_src, int T_dst, int n_head, cudaStream_t stream);
void launch_kv_export(const __nv_bfloat16* src, float* dst, int n_seq, int n_rows, int T_src, int T_dst, int n_head, cudaStream_t stream);

}  // namespace b200w
