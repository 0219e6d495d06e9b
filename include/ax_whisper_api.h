/*
 * ax_whisper_api.h -- C interface of libax_whisper.so, B200 (sm_100a) build.
 *
 * Drop-in for the reference's public header /root/reference/cpp/src/api/ax_whisper_api.h: the four
 * AX_WHISPER_* entry points below keep its exact names, signatures, ownership and error conventions
 * (header :22,:54,:67,:81-83,:98-101; behaviour from ax_whisper_api.cpp:48-163).  Callers in the reference:
 * cpp/whisper_cli.cpp:82,95,107 and cpp/src/WhisperHTTPServer.hpp:17,21,78.
 *
 * Differences a caller can observe:
 *   - model directory: {model_path}/{model_type}/{model_type}-encoder.b200w and -decoder.b200w replace the two
 *     .axmodel blobs; {model_type}-tokens.txt and {model_type}_config.json are unchanged (Whisper.cpp:87-90).
 *   - device initialisation (the executables' AX_SYS_Init / AX_ENGINE_Init, whisper_cli.cpp:37-61) happens
 *     inside AX_WHISPER_Init: CUDA device selection (env B200W_DEVICE, default 0), weight upload, workspaces.
 *   - the handle is internally locked, so concurrent Run* calls from several threads serialise instead of racing
 *     (the reference's HTTP server calls an unlocked handle, WhisperHTTPServer.hpp:78).
 *   - audio shorter than 201 samples returns -1 (the reference reads out of bounds, librosa.h:51-56).
 *   - there is no CPU fallback: without a B200 the Init call fails and returns NULL.
 * The *Batch / *Tokens entry points are additive extensions for data-parallel use.
 */
#ifndef _AX_WHISPER_API_H_
#define _AX_WHISPER_API_H_

#ifdef __cplusplus
extern "C" {
#endif

#define AX_WHISPER_API __attribute__((visibility("default")))

/* Opaque handle (reference: ax_whisper_api.h:22) */
typedef void* AX_WHISPER_HANDLE;

/* Load {model_path}/{model_type}/..., pick the language token (unknown language -> "zh", Whisper.cpp:241-251).
 * Returns NULL on failure (reference: ax_whisper_api.h:54, ax_whisper_api.cpp:48-57). */
AX_WHISPER_API AX_WHISPER_HANDLE AX_WHISPER_Init(const char* model_type, const char* model_path, const char* language);

/* Release everything; Uninit(NULL) is a no-op (reference: ax_whisper_api.h:67, ax_whisper_api.cpp:69-74). */
AX_WHISPER_API void AX_WHISPER_Uninit(AX_WHISPER_HANDLE handle);

/* Transcribe the first 30 s of a WAV file (16 kHz; stereo is averaged to mono, ax_whisper_api.cpp:109-113).
 * *result is malloc'ed (strdup) and must be free()d by the caller.  0 on success, -1 on error
 * (reference: ax_whisper_api.h:81-83, ax_whisper_api.cpp:88-124). */
AX_WHISPER_API int AX_WHISPER_RunFile(AX_WHISPER_HANDLE handle, const char* wav_file, char** result);

/* Transcribe 16 kHz mono f32 PCM in [-1, 1]; the buffer is copied, the caller keeps ownership
 * (reference: ax_whisper_api.h:98-101, ax_whisper_api.cpp:139-163). */
AX_WHISPER_API int AX_WHISPER_RunPCM(AX_WHISPER_HANDLE handle, float* pcm_data, int num_samples, char** result);

/* ---- extensions (not in the reference) ------------------------------------------------------------------ */

/* Transcribe `batch` independent utterances in one pass.  results[i] is malloc'ed per utterance; the array
 * `results` itself is supplied by the caller (batch entries).  0 on success, -1 on error (no results set). */
AX_WHISPER_API int AX_WHISPER_RunPCMBatch(AX_WHISPER_HANDLE handle, const float* const* pcm_data, const int* num_samples, int batch,
                                          char** results);

/* Same, returning token ids instead of text: tokens is [batch][max_tokens], n_tokens [batch].
 * max_new_tokens <= 0 means "until EOT or the 448-token context is full" like the reference loop
 * (Whisper.cpp:219); honor_eot = 0 keeps generating past EOT (throughput measurements). */
AX_WHISPER_API int AX_WHISPER_RunPCMTokens(AX_WHISPER_HANDLE handle, const float* const* pcm_data, const int* num_samples, int batch,
                                           int max_new_tokens, int honor_eot, int* tokens, int max_tokens, int* n_tokens);

/* Long-form extension (BASELINE.json configs[4]): the reference transcribes only the first 30 s (Whisper.cpp:172); this
 * entry point cuts the audio into consecutive 30 s windows (the last one may be shorter; a tail under 201 samples is
 * dropped), transcribes them as one data-parallel batch (window_batch windows per pass, <= 0 for all at once) and
 * concatenates the window texts in order.  Windows are independent: no timestamps, no previous-text conditioning. */
AX_WHISPER_API int AX_WHISPER_RunPCMLong(AX_WHISPER_HANDLE handle, const float* pcm_data, long num_samples, int window_batch, char** result);

/* Concurrency: unlike the reference (whose handle is not re-entrant although whisper_srv calls it from a thread pool,
 * WhisperHTTPServer.hpp:78), every entry point may be called from several threads on one handle.  AX_WHISPER_RunPCM /
 * RunFile calls that arrive while a GPU pass is running are coalesced: the next pass transcribes all of them as one batch
 * (at most B200W_COALESCE_MAX = 64; B200W_COALESCE_WAIT_US adds a fixed wait for company, default 0).
 * AX_WHISPER_GetStats reports how many single-utterance requests were served in how many GPU passes. */
AX_WHISPER_API int AX_WHISPER_GetStats(AX_WHISPER_HANDLE handle, long* n_requests, long* n_gpu_passes);

/* Text of the last error on this thread ("" if none). */
AX_WHISPER_API const char* AX_WHISPER_LastError(void);

#ifdef __cplusplus
}
#endif

#endif /* _AX_WHISPER_API_H_ */
