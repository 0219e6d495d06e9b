/*
 * b200w_model_abi.h -- model-level C ABI of libax_whisper.so (inner boundary).
 *
 * This is what the reference reaches through AxModelRunner::{set_input, run, get_output}
 * (/root/reference/cpp/src/ax_model_runner/ax_model_runner.hpp:29-61, ax_model_runner.cpp:95-171) with tensor
 * indices, for its two NPU graphs (/root/reference/model_convert/export_onnx.py:587-588 encoder I/O names,
 * :668-670 decoder I/O names; call sites /root/reference/cpp/src/Whisper.cpp:190-198, :299-345):
 *
 *   encoder        in0 mel[B,n_mels,3000] f32                  -> out0 cross_k, out1 cross_v [L,B,1500,d] f32
 *   decoder        in0 tokens[B] i32, in1/2 self_k/v cache, in3/4 cross_k/v, in5 offset i32, in6 mask[448]
 *                                                               -> out0 logits[B,V] f32, out1/2 this_self_k/v [L,B,d]
 * BASELINE.json's north_star names the older three-graph split (encoder, decoder_main, decoder_loop); both
 * decoder entry points are defined in terms of the snapshot's single-step graph:
 * decoder_main == the 4 sequential SOT steps of Whisper.cpp:214-217, decoder_loop == one step of :219-222.
 *
 * On the B200 the KV caches never leave HBM: cross_k/v produced by b200w_encoder and the self-attention cache
 * stay resident in the engine (bf16, head-major), so decoder_main / decoder_loop take only (tokens, offset); the mask of the
 * reference graph is implied by offset (mask[j] = j >= offset, export_onnx.py:59-68).  The reference's stateless form --
 * every cache an input, new rows an output -- is b200w_decoder_step together with b200w_{set,get}_{cross,self}_kv below.  Every pointer is a plain
 * HOST pointer unless its name ends in _dev; every function returns 0 on success and -1 on error
 * (message via b200w_last_error()).  No function falls back to the CPU.
 */
#ifndef B200W_MODEL_ABI_H_
#define B200W_MODEL_ABI_H_

#ifdef __cplusplus
extern "C" {
#endif

#define B200W_API __attribute__((visibility("default")))

typedef struct b200w_engine b200w_engine;

typedef struct b200w_dims {
  int n_mels, n_vocab, d_model, n_head, n_audio_layer, n_text_layer, n_audio_ctx, n_text_ctx;
  int sot, eot, transcribe, no_timestamps;
} b200w_dims;

typedef struct b200w_times {  /* per-stage device times of the last transcribe call, CUDA events on the engine stream */
  float h2d_ms, mel_ms, encoder_ms, decode_ms, total_ms;
  int decode_steps;
  long long kernel_launches;
} b200w_times;

B200W_API const char* b200w_last_error(void);

/* Model directory convention as AX_WHISPER_Init.  max_batch sizes the resident workspaces (grown on demand). */
B200W_API int b200w_engine_create(const char* model_path, const char* model_type, int device, int max_batch, b200w_engine** out);
B200W_API void b200w_engine_destroy(b200w_engine* e);
B200W_API int b200w_get_dims(const b200w_engine* e, b200w_dims* out);
/* {sot, language, transcribe, no_timestamps}; unknown language falls back to "zh" (Whisper.cpp:139, :241-251) */
B200W_API int b200w_sot_sequence(const b200w_engine* e, const char* language, int out_tokens[4]);

/* K1, replaces librosa::Feature::melspectrogram + Whisper::preprocess (Whisper.cpp:151-184).
 * pcm [B][pcm_stride] f32, n_samples [B] (each >= 201) -> mel_out [B][n_mels][3000] f32 (may be NULL). The result
 * also stays resident as the encoder's input. */
B200W_API int b200w_logmel(b200w_engine* e, const float* pcm, long pcm_stride, const int* n_samples, int B, float* mel_out);

/* encoder graph (export_onnx.py:187-213).  mel [B][n_mels][3000] f32, or NULL to use the resident result of
 * b200w_logmel.  cross_k / cross_v [L][B][1500][d] f32 receive a copy of the resident cache (either may be NULL). */
B200W_API int b200w_encoder(b200w_engine* e, const float* mel, int B, float* cross_k, float* cross_v);

/* decoder_main: resets the self-attention cache and runs the SOT prefix (n_tokens steps, normally 4) for B
 * sequences sharing the prefix.  logits [B][n_vocab] of the last step, this_self_k/v [L][B][n_tokens][d]. */
B200W_API int b200w_decoder_main(b200w_engine* e, const int* sot_tokens, int n_tokens, int B, float* logits, float* this_self_k,
                                 float* this_self_v);

/* decoder_loop: one step (TextDecoderTensorCache.forward, export_onnx.py:312-387 + cache row update of
 * Whisper.cpp:328-342).  tokens [B], offset = position of these tokens. logits [B][n_vocab], this_self_k/v [L][B][d]. */
B200W_API int b200w_decoder_loop(b200w_engine* e, const int* tokens, int offset, int B, float* logits, float* this_self_k,
                                 float* this_self_v);

/* ---- the caches as tensors (the reference graph's in1..in4 / out1..out2) ------------------------------------------------
 * The reference's decoder is stateless: self_k/v [L][B][448][d], cross_k/v [L][B][1500][d], offset and mask[448] go IN at every
 * step and this_self_k/v [L][B][d] come OUT (export_onnx.py:668-670; AxModelRunner::set_input / get_output,
 * ax_model_runner.cpp:110-171; call site Whisper.cpp:306-326, cache row update :328-342).  Here the caches stay resident, and
 * these entry points move them across the boundary in the reference's f32 token-major layouts, so a caller can supply an
 * encoder output of its own, fork or resume a decode, or read the updated cache back. */
/* resident cross K/V of sequences [b0, b0 + nb) -> cross_k / cross_v [L][nb][1500][d] f32 (either may be NULL) */
B200W_API int b200w_get_cross_kv(b200w_engine* e, int b0, int nb, float* cross_k, float* cross_v);
/* cross_k / cross_v [L][B][1500][d] f32 -> resident cache (replaces b200w_encoder's result for sequences 0..B-1) */
B200W_API int b200w_set_cross_kv(b200w_engine* e, const float* cross_k, const float* cross_v, int B);
/* self_k / self_v [L][B][448][d] f32, rows [0, n_valid) of every sequence -> resident self-attention cache */
B200W_API int b200w_set_self_kv(b200w_engine* e, const float* self_k, const float* self_v, int n_valid, int B);
/* resident self-attention cache rows [0, n_rows) -> self_k / self_v [L][B][n_rows][d] f32 */
B200W_API int b200w_get_self_kv(b200w_engine* e, float* self_k, float* self_v, int n_rows, int B);
/* One decoder run with the reference graph's full input list.  Any of self_k/self_v/cross_k/cross_v may be NULL = keep the
 * resident tensor; mask (may be NULL) must be the causal mask the reference host builds, mask[j] = (j >= offset)
 * (Whisper.cpp:201,:253-258), anything else is rejected.  logits [B][n_vocab], this_self_k/v [L][B][d] (row `offset`). */
B200W_API int b200w_decoder_step(b200w_engine* e, const int* tokens, const float* self_k, const float* self_v, const float* cross_k,
                                 const float* cross_v, int offset, const int* mask, int B, float* logits, float* this_self_k,
                                 float* this_self_v);
/* EOT handling of the greedy loop: finished sequences are dropped from the decoder's slot list at the 16-step polls (the
 * remaining steps cost what the still-running sequences cost).  *compactions = how often that happened since the engine was
 * created, *last_active = sequences still decoding when the last greedy loop stopped. */
B200W_API int b200w_decode_stats(const b200w_engine* e, long* compactions, int* last_active);
/* With logits_out in b200w_greedy: copy only these sequences' logits, logits_out becomes [max_new_tokens][n][n_vocab]
 * (n = 0 restores "all sequences"). */
B200W_API int b200w_set_logit_rows(b200w_engine* e, const int* rows, int n);

/* Greedy loop of Whisper::run (Whisper.cpp:200-222) on the resident cross K/V of the last b200w_encoder call.
 * forced_tokens [B][forced_len] (or NULL): teacher forcing.  logits_out [max_new_tokens][B][n_vocab] or NULL.
 * tokens_out [B][max_tokens], n_tokens_out [B]. */
B200W_API int b200w_greedy(b200w_engine* e, int B, const char* language, int max_new_tokens, int honor_eot, const int* forced_tokens,
                           int forced_len, float* logits_out, int* tokens_out, int max_tokens, int* n_tokens_out);

/* Whole path PCM -> token ids with host buffers (H2D / D2H inside). pcm [B][pcm_stride]. */
B200W_API int b200w_transcribe(b200w_engine* e, const float* pcm, long pcm_stride, const int* n_samples, int B, const char* language,
                               int max_new_tokens, int honor_eot, int* tokens_out, int max_tokens, int* n_tokens_out, b200w_times* times);
/* Upload once, then run the whole path on HBM-resident PCM (bench "value" leg). */
B200W_API int b200w_upload_pcm(b200w_engine* e, const float* pcm, long pcm_stride, const int* n_samples, int B);
B200W_API int b200w_transcribe_resident(b200w_engine* e, int B, const char* language, int max_new_tokens, int honor_eot, int* tokens_out,
                                        int max_tokens, int* n_tokens_out, b200w_times* times);

/* Stage runners on resident data without host traffic (benchmarks / profiling): each enqueues on the engine stream and
 * returns the CUDA-event time of `iters` back-to-back runs in *ms. stage: 0 = log-mel, 1 = encoder, 2 = decode (n_steps),
 * 3 = only the cross-attention decode kernel, one launch per decoder layer. */
B200W_API int b200w_time_stage(b200w_engine* e, int stage, int B, int iters, int n_steps, float* ms);

/* Test hooks: tcgen05 GEMM against the SIMT comparator on random data (returns max abs difference), constant tables. */
B200W_API int b200w_selftest_gemm(int M, int N, int K, int block_n, int epilogue, unsigned seed, float* max_abs_diff, float* max_abs_ref);
B200W_API int b200w_selftest_attention(int B, int T, int n_head, unsigned seed, float* max_abs_diff, float* max_abs_ref);
/* decode cross attention: streaming kernel (B * n_head >= 296) vs the per-(sequence, head) kernel, same random inputs */
B200W_API int b200w_selftest_cross_attention(int B, int n_head, int T, unsigned seed, float* max_abs_diff, float* max_abs_ref);
B200W_API int b200w_mel_tables(int n_mels, float* bank /*[n_mels][201]*/, float* window /*[400]*/);
/* Host-logic test hooks (no GPU needed): config parser (Whisper.cpp:93-101), WAV reader (AudioFile.h equivalent; out is
 * [frame][channel]), base64 token decoder (base64.cpp equivalent; returns the decoded length). */
B200W_API int b200w_test_parse_config(const char* model_path, const char* model_type, b200w_dims* out, int* n_languages);
B200W_API int b200w_test_load_wav(const char* path, float* out, int capacity_frames, int* n_frames, int* n_channels, int* sample_rate);
B200W_API int b200w_test_base64(const char* in, unsigned char* out, int capacity);
/* token table loader + detokeniser of the outer API (Whisper.cpp:115-127, :224-229) on a {type}-tokens.txt file: the bytes of
 * ids[0..n) concatenated into out (may contain NUL); returns the full length, -1 on error. */
B200W_API int b200w_test_detokenize(const char* tokens_file, const int* ids, int n, unsigned char* out, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* B200W_MODEL_ABI_H_ */
