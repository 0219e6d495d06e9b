// C ABI of the model-level boundary (include/b200w_model_abi.h) over the Engine.
#include "../../include/b200w_model_abi.h"

#include <cmath>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "common.cuh"
#include "engine.h"
#include "host_utils.h"

using namespace b200w;

struct b200w_engine {
  Engine* eng;
  int last_max_samples = kChunkSamples;
  std::vector<int> logit_rows;  // b200w_set_logit_rows
};

namespace {
thread_local std::string g_err;

template <class F>
int guarded(F&& f) {
  try {
    g_err.clear();
    f();
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    fprintf(stderr, "[b200w] error: %s\n", ex.what());
    return -1;
  } catch (...) {
    g_err = "unknown error";
    return -1;
  }
}

void fill_tokens(const std::vector<std::vector<int>>& toks, int B, int* tokens_out, int max_tokens, int* n_tokens_out) {
  for (int b = 0; b < B; ++b) {
    const int n = std::min<int>((int)toks[b].size(), max_tokens);
    if (tokens_out)
      for (int i = 0; i < n; ++i) tokens_out[(size_t)b * max_tokens + i] = toks[b][i];
    if (n_tokens_out) n_tokens_out[b] = n;
  }
}
void fill_times(const StageTimes& t, b200w_times* out) {
  if (!out) return;
  out->h2d_ms = t.h2d_ms, out->mel_ms = t.mel_ms, out->encoder_ms = t.encoder_ms, out->decode_ms = t.decode_ms;
  out->total_ms = t.total_ms, out->decode_steps = t.decode_steps, out->kernel_launches = t.kernel_launches;
}
DecodeOptions make_opts(int max_new_tokens, int honor_eot) {
  DecodeOptions o;
  if (max_new_tokens > 0) o.max_new_tokens = std::min(max_new_tokens, kTextCtx - kSotLen);
  o.honor_eot = honor_eot != 0;
  return o;
}
}  // namespace

extern "C" {

const char* b200w_last_error(void) { return g_err.c_str(); }

int b200w_engine_create(const char* model_path, const char* model_type, int device, int max_batch, b200w_engine** out) {
  if (!model_path || !model_type || !out) return -1;
  *out = nullptr;
  return guarded([&] {
    Engine* e = new Engine(model_path, model_type, device, max_batch);
    *out = new b200w_engine{e};
  });
}
void b200w_engine_destroy(b200w_engine* e) {
  if (!e) return;
  delete e->eng;
  delete e;
}
int b200w_get_dims(const b200w_engine* e, b200w_dims* out) {
  if (!e || !out) return -1;
  const ModelConfig& c = e->eng->config();
  out->n_mels = c.n_mels, out->n_vocab = c.n_vocab, out->d_model = c.d, out->n_head = c.n_head;
  out->n_audio_layer = c.l_enc, out->n_text_layer = c.l_dec, out->n_audio_ctx = c.n_audio_ctx, out->n_text_ctx = c.n_text_ctx;
  out->sot = c.sot, out->eot = c.eot, out->transcribe = c.transcribe, out->no_timestamps = c.no_timestamps;
  return 0;
}
int b200w_sot_sequence(const b200w_engine* e, const char* language, int out_tokens[4]) {
  if (!e || !out_tokens) return -1;
  return guarded([&] {
    auto s = e->eng->sot_sequence(language ? language : "zh");
    for (int i = 0; i < 4; ++i) out_tokens[i] = s[i];
  });
}

int b200w_upload_pcm(b200w_engine* e, const float* pcm, long pcm_stride, const int* n_samples, int B) {
  if (!e || !pcm || !n_samples || B <= 0) return -1;
  return guarded([&] {
    Engine& E = *e->eng;
    int max_samples = 0;
    for (int b = 0; b < B; ++b) {
      if (n_samples[b] < 201 || n_samples[b] > pcm_stride) throw std::runtime_error("n_samples must be in [201, pcm_stride]");
      max_samples = std::max(max_samples, n_samples[b]);
    }
    E.ensure_capacity(B, max_samples);
    CUDA_CHECK(cudaSetDevice(E.device()));
    CUDA_CHECK(cudaMemcpy2DAsync(E.pcm_dev(), (size_t)E.pcm_stride() * 4, pcm, (size_t)pcm_stride * 4, (size_t)max_samples * 4, B,
                                 cudaMemcpyHostToDevice, E.stream()));
    CUDA_CHECK(cudaMemcpyAsync(E.n_samples_dev(), n_samples, sizeof(int) * B, cudaMemcpyHostToDevice, E.stream()));
    CUDA_CHECK(cudaStreamSynchronize(E.stream()));
    e->last_max_samples = max_samples;
  });
}

int b200w_logmel(b200w_engine* e, const float* pcm, long pcm_stride, const int* n_samples, int B, float* mel_out) {
  if (b200w_upload_pcm(e, pcm, pcm_stride, n_samples, B) != 0) return -1;
  return guarded([&] {
    Engine& E = *e->eng;
    E.run_logmel(B, e->last_max_samples);
    if (mel_out)
      CUDA_CHECK(cudaMemcpyAsync(mel_out, E.mel_dev(), (size_t)B * E.config().n_mels * kMelFrames * 4, cudaMemcpyDeviceToHost, E.stream()));
    CUDA_CHECK(cudaStreamSynchronize(E.stream()));
  });
}

int b200w_encoder(b200w_engine* e, const float* mel, int B, float* cross_k, float* cross_v) {
  if (!e || B <= 0) return -1;
  return guarded([&] {
    Engine& E = *e->eng;
    E.ensure_capacity(B);
    CUDA_CHECK(cudaSetDevice(E.device()));
    if (mel) {
      CUDA_CHECK(cudaMemcpyAsync(E.mel_dev(), mel, (size_t)B * E.config().n_mels * kMelFrames * 4, cudaMemcpyHostToDevice, E.stream()));
      E.run_mel_convert(B);
    }
    E.run_encoder(B);
    CUDA_CHECK(cudaStreamSynchronize(E.stream()));
    if (cross_k || cross_v) E.read_cross_kv(0, B, cross_k, cross_v);
  });
}

int b200w_decoder_main(b200w_engine* e, const int* sot_tokens, int n_tokens, int B, float* logits, float* this_self_k, float* this_self_v) {
  if (!e || !sot_tokens || n_tokens <= 0 || B <= 0) return -1;
  return guarded([&] {
    Engine& E = *e->eng;
    const int d = E.config().d, L = E.config().l_dec;
    E.decode_reset(B);
    std::vector<int> tok(B);
    std::vector<float> k1((size_t)L * B * d), v1((size_t)L * B * d);
    for (int i = 0; i < n_tokens; ++i) {
      for (int b = 0; b < B; ++b) tok[b] = sot_tokens[i];
      const bool last = i == n_tokens - 1;
      const bool want_kv = this_self_k || this_self_v;
      E.decode_step_tokens(B, tok.data(), i, last ? logits : nullptr, want_kv ? k1.data() : nullptr, want_kv ? v1.data() : nullptr);
      if (want_kv)
        for (int l = 0; l < L; ++l)
          for (int b = 0; b < B; ++b) {
            if (this_self_k) memcpy(this_self_k + (((size_t)l * B + b) * n_tokens + i) * d, k1.data() + ((size_t)l * B + b) * d, d * 4);
            if (this_self_v) memcpy(this_self_v + (((size_t)l * B + b) * n_tokens + i) * d, v1.data() + ((size_t)l * B + b) * d, d * 4);
          }
    }
  });
}

int b200w_decoder_loop(b200w_engine* e, const int* tokens, int offset, int B, float* logits, float* this_self_k, float* this_self_v) {
  if (!e || !tokens || B <= 0) return -1;
  return guarded([&] { e->eng->decode_step_tokens(B, tokens, offset, logits, this_self_k, this_self_v); });
}

int b200w_get_cross_kv(b200w_engine* e, int b0, int nb, float* cross_k, float* cross_v) {
  if (!e || nb <= 0) return -1;
  return guarded([&] { e->eng->read_cross_kv(b0, nb, cross_k, cross_v); });
}
int b200w_set_cross_kv(b200w_engine* e, const float* cross_k, const float* cross_v, int B) {
  if (!e || B <= 0 || (!cross_k && !cross_v)) return -1;
  return guarded([&] { e->eng->load_cross_kv(B, cross_k, cross_v); });
}
int b200w_set_self_kv(b200w_engine* e, const float* self_k, const float* self_v, int n_valid, int B) {
  if (!e || B <= 0 || (!self_k && !self_v)) return -1;
  return guarded([&] { e->eng->load_self_kv(B, n_valid, self_k, self_v); });
}
int b200w_get_self_kv(b200w_engine* e, float* self_k, float* self_v, int n_rows, int B) {
  if (!e || B <= 0) return -1;
  return guarded([&] {
    if (B > e->eng->capacity()) throw std::runtime_error("get_self_kv: batch exceeds the resident capacity");
    e->eng->read_self_kv(B, n_rows, self_k, self_v);
  });
}

// The reference's decoder graph, stateless form (export_onnx.py:668-670; the call in Whisper.cpp:306-326): every input tensor is
// supplied by the caller, the new cache rows are returned.  NULL for a cache means "the resident one".
int b200w_decoder_step(b200w_engine* e, const int* tokens, const float* self_k, const float* self_v, const float* cross_k,
                       const float* cross_v, int offset, const int* mask, int B, float* logits, float* this_self_k, float* this_self_v) {
  if (!e || !tokens || B <= 0) return -1;
  return guarded([&] {
    Engine& E = *e->eng;
    if (offset < 0 || offset >= kTextCtx) throw std::runtime_error("decoder_step: offset out of range");
    if (mask != nullptr) {
      // the graph's mask input (1 = masked, export_onnx.py:59-68,:130) is always the causal one the host builds
      // (Whisper.cpp:201,:253-258: all ones, then mask[offset - 1] = 0 before each step): slots < offset visible
      for (int j = 0; j < kTextCtx; ++j)
        if ((mask[j] != 0) != (j >= offset))
          throw std::runtime_error("decoder_step: only the causal mask of Whisper.cpp:253-258 (mask[j] = j >= offset) is supported; mask[" +
                                   std::to_string(j) + "] = " + std::to_string(mask[j]) + " at offset " + std::to_string(offset));
    }
    if (cross_k || cross_v) E.load_cross_kv(B, cross_k, cross_v);
    if (self_k || self_v) E.load_self_kv(B, offset, self_k, self_v);
    E.decode_step_tokens(B, tokens, offset, logits, this_self_k, this_self_v);
  });
}

int b200w_decode_stats(const b200w_engine* e, long* compactions, int* last_active) {
  if (!e) return -1;
  if (compactions) *compactions = e->eng->compactions();
  if (last_active) *last_active = e->eng->last_active();
  return 0;
}

int b200w_set_logit_rows(b200w_engine* e, const int* rows, int n) {
  if (!e || n < 0 || (n > 0 && !rows)) return -1;
  e->logit_rows.assign(rows, rows + n);
  return 0;
}

int b200w_greedy(b200w_engine* e, int B, const char* language, int max_new_tokens, int honor_eot, const int* forced_tokens, int forced_len,
                 float* logits_out, int* tokens_out, int max_tokens, int* n_tokens_out) {
  if (!e || B <= 0) return -1;
  return guarded([&] {
    Engine& E = *e->eng;
    DecodeOptions o = make_opts(max_new_tokens, honor_eot);
    o.forced_tokens = forced_tokens;
    o.forced_len = forced_tokens ? forced_len : 0;
    o.logits_out = logits_out;
    if (logits_out && !e->logit_rows.empty()) o.logit_rows = e->logit_rows.data(), o.n_logit_rows = (int)e->logit_rows.size();
    std::vector<std::vector<int>> toks;
    E.run_decode(B, E.sot_sequence(language ? language : "zh"), o, &toks);
    CUDA_CHECK(cudaStreamSynchronize(E.stream()));
    fill_tokens(toks, B, tokens_out, max_tokens, n_tokens_out);
  });
}

int b200w_transcribe(b200w_engine* e, const float* pcm, long pcm_stride, const int* n_samples, int B, const char* language,
                     int max_new_tokens, int honor_eot, int* tokens_out, int max_tokens, int* n_tokens_out, b200w_times* times) {
  if (!e || !pcm || !n_samples || B <= 0) return -1;
  return guarded([&] {
    Engine& E = *e->eng;
    std::vector<const float*> ptrs(B);
    for (int b = 0; b < B; ++b) ptrs[b] = pcm + (size_t)b * pcm_stride;
    std::vector<std::vector<int>> toks;
    StageTimes t;
    E.transcribe(ptrs.data(), n_samples, B, language ? language : "zh", make_opts(max_new_tokens, honor_eot), &toks, &t);
    fill_tokens(toks, B, tokens_out, max_tokens, n_tokens_out);
    fill_times(t, times);
  });
}

int b200w_transcribe_resident(b200w_engine* e, int B, const char* language, int max_new_tokens, int honor_eot, int* tokens_out,
                              int max_tokens, int* n_tokens_out, b200w_times* times) {
  if (!e || B <= 0) return -1;
  return guarded([&] {
    Engine& E = *e->eng;
    if (B > E.capacity()) throw std::runtime_error("transcribe_resident: upload PCM first");
    std::vector<std::vector<int>> toks;
    StageTimes t;
    E.transcribe_resident(B, e->last_max_samples, language ? language : "zh", make_opts(max_new_tokens, honor_eot), &toks, &t);
    fill_tokens(toks, B, tokens_out, max_tokens, n_tokens_out);
    fill_times(t, times);
  });
}

int b200w_time_stage(b200w_engine* e, int stage, int B, int iters, int n_steps, float* ms) {
  if (!e || !ms || B <= 0 || iters <= 0) return -1;
  return guarded([&] {
    Engine& E = *e->eng;
    if (B > E.capacity()) throw std::runtime_error("time_stage: batch exceeds the resident capacity");
    CUDA_CHECK(cudaSetDevice(E.device()));
    cudaEvent_t a, b;
    CUDA_CHECK(cudaEventCreate(&a));
    CUDA_CHECK(cudaEventCreate(&b));
    CUDA_CHECK(cudaStreamSynchronize(E.stream()));
    CUDA_CHECK(cudaEventRecord(a, E.stream()));
    for (int i = 0; i < iters; ++i) {
      if (stage == 0) {
        E.run_logmel(B, e->last_max_samples);
      } else if (stage == 1) {
        E.run_encoder(B);
      } else if (stage == 2) {
        DecodeOptions o;
        o.honor_eot = false;
        o.max_new_tokens = std::max(1, std::min(n_steps, kTextCtx) - kSotLen);
        E.run_decode(B, E.sot_sequence("zh"), o, nullptr);
      } else if (stage == 3) {
        E.run_cross_attention_only(B);
      } else {
        throw std::runtime_error("time_stage: unknown stage");
      }
    }
    CUDA_CHECK(cudaEventRecord(b, E.stream()));
    CUDA_CHECK(cudaStreamSynchronize(E.stream()));
    CUDA_CHECK(cudaEventElapsedTime(ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
  });
}

// ---- host-logic test hooks (no GPU needed) ----
int b200w_test_parse_config(const char* model_path, const char* model_type, b200w_dims* out, int* n_languages) {
  if (!model_path || !model_type || !out) return -1;
  return guarded([&] {
    const ModelConfig c = load_model_config(model_path, model_type);
    out->n_mels = c.n_mels, out->n_vocab = c.n_vocab, out->d_model = c.d, out->n_head = c.n_head;
    out->n_audio_layer = c.l_enc, out->n_text_layer = c.l_dec, out->n_audio_ctx = c.n_audio_ctx, out->n_text_ctx = c.n_text_ctx;
    out->sot = c.sot, out->eot = c.eot, out->transcribe = c.transcribe, out->no_timestamps = c.no_timestamps;
    if (n_languages) *n_languages = (int)c.lang_codes.size();
  });
}
int b200w_test_load_wav(const char* path, float* out, int capacity_frames, int* n_frames, int* n_channels, int* sample_rate) {
  if (!path || !n_frames || !n_channels || !sample_rate) return -1;
  return guarded([&] {
    WavData w;
    std::string err;
    if (!load_wav(path, &w, &err)) throw std::runtime_error(err);
    *n_channels = (int)w.channels.size();
    *n_frames = (int)w.channels[0].size();
    *sample_rate = w.sample_rate;
    // channel-interleaved copy [frame][channel]
    if (out)
      for (int i = 0; i < *n_frames && i < capacity_frames; ++i)
        for (int c = 0; c < *n_channels; ++c) out[(size_t)i * *n_channels + c] = w.channels[c][i];
  });
}
int b200w_test_base64(const char* in, unsigned char* out, int capacity) {
  if (!in || !out) return -1;
  const std::string s = base64_decode(in);
  const int n = (int)s.size() < capacity ? (int)s.size() : capacity;
  memcpy(out, s.data(), n);
  return (int)s.size();
}

int b200w_test_detokenize(const char* tokens_file, const int* ids, int n, unsigned char* out, int capacity) {
  if (!tokens_file || !ids || n < 0 || !out) return -1;
  TokenTable t;
  std::string err;
  if (!t.load(tokens_file, &err)) {
    g_err = err;
    return -1;
  }
  const std::string s = t.detokenize(ids, (size_t)n);
  memcpy(out, s.data(), std::min<size_t>(s.size(), (size_t)std::max(capacity, 0)));
  return (int)s.size();
}

int b200w_mel_tables(int n_mels, float* bank, float* window) {
  if (n_mels != 80 && n_mels != 128) return -1;
  logmel_mel_table_copy(n_mels, bank, window);
  return 0;
}

// tcgen05 attention against the mma.sync comparator kernel (every element) AND an fp64 host evaluation (sample rows) on random
// bf16 q/k/v ([B*T][3d], head_dim 64).
int b200w_selftest_attention(int B, int T, int n_head, unsigned seed, float* max_abs_diff, float* max_abs_ref) {
  if (!max_abs_diff || !max_abs_ref) return -1;
  return guarded([&] {
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) throw CudaError("no CUDA device");
    kernels_set_attributes();
    const int d = n_head * 64;
    const size_t n_in = (size_t)B * T * 3 * d, n_out = (size_t)B * T * d;
    std::mt19937 rng(seed);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<__nv_bfloat16> h(n_in);
    for (size_t i = 0; i < n_in; ++i) h[i] = __float2bfloat16(nd(rng) * ((i / d) % 3 == 2 ? 1.0f : 1.5f));
    __nv_bfloat16 *dq, *o1, *o2;
    CUDA_CHECK(cudaMalloc(&dq, n_in * 2));
    CUDA_CHECK(cudaMalloc(&o1, n_out * 2));
    CUDA_CHECK(cudaMalloc(&o2, n_out * 2));
    CUDA_CHECK(cudaMemcpy(dq, h.data(), n_in * 2, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemset(o1, 0, n_out * 2));
    CUDA_CHECK(cudaMemset(o2, 0, n_out * 2));
    cudaStream_t s;
    CUDA_CHECK(cudaStreamCreate(&s));
    launch_encoder_attention(dq, o1, B, T, n_head, s);
    launch_encoder_attention_tcgen05(dq, o2, B, T, n_head, s);
    CUDA_CHECK(cudaStreamSynchronize(s));
    std::vector<__nv_bfloat16> a(n_out), b(n_out);
    CUDA_CHECK(cudaMemcpy(a.data(), o1, n_out * 2, cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(b.data(), o2, n_out * 2, cudaMemcpyDeviceToHost));
    double md = 0, mr = 0;
    for (size_t i = 0; i < n_out; ++i) {
      const float x = __bfloat162float(a[i]), y = __bfloat162float(b[i]);
      if (!(y == y)) md = 1e9;  // NaN
      md = std::max(md, (double)fabsf(x - y));
      mr = std::max(mr, (double)fabsf(x));
    }
    // independent of any CUDA code: fp64 host softmax(q k^T / 8) v for sample (chunk, head, query) rows of the tcgen05 output
    for (int smp = 0; smp < 12; ++smp) {
      const int bb = (int)(rng() % (unsigned)B), hh = (int)(rng() % (unsigned)n_head);
      const int qi = smp == 0 ? 0 : smp == 1 ? T - 1 : (int)(rng() % (unsigned)T);
      auto at = [&](int t, int which, int i) { return (double)__bfloat162float(h[((size_t)bb * T + t) * 3 * d + (size_t)which * d + hh * 64 + i]); };
      std::vector<double> sc(T);
      double mx = -1e300;
      for (int t = 0; t < T; ++t) {
        double acc = 0;
        for (int i = 0; i < 64; ++i) acc += at(qi, 0, i) * at(t, 1, i);
        sc[t] = acc * 0.125;
        mx = std::max(mx, sc[t]);
      }
      double l = 0;
      std::vector<double> o(64, 0.0);
      for (int t = 0; t < T; ++t) {
        const double p = std::exp(sc[t] - mx);
        l += p;
        for (int i = 0; i < 64; ++i) o[i] += p * at(t, 2, i);
      }
      for (int i = 0; i < 64; ++i) {
        const double ref = o[i] / l;
        md = std::max(md, fabs((double)__bfloat162float(b[((size_t)bb * T + qi) * d + hh * 64 + i]) - ref));
        mr = std::max(mr, fabs(ref));
      }
    }
    *max_abs_diff = (float)md;
    *max_abs_ref = (float)mr;
    cudaStreamDestroy(s);
    cudaFree(dq), cudaFree(o1), cudaFree(o2);
  });
}

// Decode cross attention on random q / K / V: the streaming kernel and every valid cluster split must agree BIT FOR BIT (the
// canonical summation order of decode_ops.cu: a sequence's result does not depend on the batch it is in), and all of them
// must match an fp32 host evaluation of softmax(q K^T / 8) V within bf16 output rounding.
// *max_abs_diff = largest difference between kernel variants (0 expected), *max_abs_ref_err = largest error against the host.
int b200w_selftest_cross_attention(int B, int n_head, int T, unsigned seed, float* max_abs_diff, float* max_abs_ref_err) {
  if (!max_abs_diff || !max_abs_ref_err) return -1;
  return guarded([&] {
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) throw CudaError("no CUDA device");
    kernels_set_attributes();
    const int d = n_head * 64;
    const size_t n_kv = (size_t)B * n_head * T * 64, n_q = (size_t)B * d;
    std::mt19937 rng(seed);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<__nv_bfloat16> hk(n_kv), hv(n_kv);
    std::vector<float> hq(n_q);
    for (size_t i = 0; i < n_kv; ++i) hk[i] = __float2bfloat16(nd(rng)), hv[i] = __float2bfloat16(nd(rng));
    for (size_t i = 0; i < n_q; ++i) hq[i] = nd(rng) * 1.5f;
    __nv_bfloat16 *dk, *dv, *o;
    float* dq;
    int *work, *dmap;
    // slot -> sequence map of the decoder (EOT compaction): a reversal here, so that q / out (per slot) and K / V (per
    // sequence) are really indexed differently
    std::vector<int> hmap(B);
    for (int i = 0; i < B; ++i) hmap[i] = B - 1 - i;
    CUDA_CHECK(cudaMalloc(&dmap, B * sizeof(int)));
    CUDA_CHECK(cudaMemcpy(dmap, hmap.data(), B * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMalloc(&dk, n_kv * 2));
    CUDA_CHECK(cudaMalloc(&dv, n_kv * 2));
    CUDA_CHECK(cudaMalloc(&dq, n_q * 4));
    CUDA_CHECK(cudaMalloc(&o, n_q * 2));
    CUDA_CHECK(cudaMalloc(&work, 2 * sizeof(int)));
    CUDA_CHECK(cudaMemcpy(dk, hk.data(), n_kv * 2, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(dv, hv.data(), n_kv * 2, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(dq, hq.data(), n_q * 4, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemset(work, 0, 2 * sizeof(int)));
    cudaStream_t s;
    CUDA_CHECK(cudaStreamCreate(&s));
    // host reference for the first and the last (sequence, head) pairs (fp64 accumulation)
    const int n_items = B * n_head;
    std::vector<int> ref_items;
    for (int it = 0; it < n_items; it += std::max(1, n_items / 7)) ref_items.push_back(it);
    ref_items.push_back(n_items - 1);
    std::vector<std::vector<double>> ref(ref_items.size(), std::vector<double>(64));
    for (size_t r = 0; r < ref_items.size(); ++r) {
      const int it = ref_items[r];                                               // (slot, head): q / out index
      const size_t cit = (size_t)hmap[it / n_head] * n_head + (it % n_head);     // (sequence, head): cache index
      std::vector<double> sc(T);
      double mx = -1e300;
      for (int t = 0; t < T; ++t) {
        double a = 0;
        for (int i = 0; i < 64; ++i) a += (double)hq[(size_t)it * 64 + i] * (double)__bfloat162float(hk[(cit * T + t) * 64 + i]);
        sc[t] = a * 0.125;
        mx = std::max(mx, sc[t]);
      }
      double l = 0;
      for (int t = 0; t < T; ++t) {
        const double p = std::exp(sc[t] - mx);
        l += p;
        for (int i = 0; i < 64; ++i) ref[r][i] += p * (double)__bfloat162float(hv[(cit * T + t) * 64 + i]);
      }
      for (int i = 0; i < 64; ++i) ref[r][i] /= l;
    }
    std::vector<std::vector<__nv_bfloat16>> results;
    const int nseg = (T + 255) / 256;
    for (int n_split = 0; n_split <= 8; ++n_split) {
      if (n_split > 0 && (nseg % n_split != 0 || nseg / n_split > 3)) continue;
      CUDA_CHECK(cudaMemsetAsync(o, 0xff, n_q * 2, s));
      for (int rep = 0; rep < (n_split == 0 ? 2 : 1); ++rep)  // streaming kernel twice: the second launch runs on re-armed counters
        launch_cross_attention_decode(dq, dk, dv, dmap, o, B, n_head, T, n_split, s, false, work);
      CUDA_CHECK(cudaStreamSynchronize(s));
      results.emplace_back(n_q);
      CUDA_CHECK(cudaMemcpy(results.back().data(), o, n_q * 2, cudaMemcpyDeviceToHost));
    }
    double md = 0, me = 0;
    for (size_t v = 0; v < results.size(); ++v) {
      for (size_t i = 0; i < n_q; ++i) {
        const float x = __bfloat162float(results[v][i]), y = __bfloat162float(results[0][i]);
        if (!(x == x)) md = 1e9;  // NaN (also the 0xff fill of an element a kernel never wrote)
        if (memcmp(&results[v][i], &results[0][i], 2) != 0) md = std::max(md, std::max(1e-30, (double)fabsf(x - y)));
      }
      for (size_t r = 0; r < ref_items.size(); ++r)
        for (int i = 0; i < 64; ++i)
          me = std::max(me, fabs((double)__bfloat162float(results[v][(size_t)ref_items[r] * 64 + i]) - ref[r][i]));
    }
    if (results.size() < 2) md = 1e9;  // nothing was cross-checked
    *max_abs_diff = (float)md;
    *max_abs_ref_err = (float)me;
    cudaStreamDestroy(s);
    cudaFree(dk), cudaFree(dv), cudaFree(dq), cudaFree(o), cudaFree(work), cudaFree(dmap);
  });
}

// tcgen05 GEMM vs the SIMT comparator (every element) and vs an fp64 host evaluation (sample rows) on random bf16 data.  Epilogues covered here: EPI_BIAS_F32 (2), EPI_BIAS_BF16 (0),
// EPI_BIAS_GELU_BF16 (1), EPI_BIAS_RESID_F32 (3), EPI_ARGMAX (6).
int b200w_selftest_gemm(int M, int N, int K, int block_n, int epilogue, unsigned seed, float* max_abs_diff, float* max_abs_ref) {
  if (!max_abs_diff || !max_abs_ref) return -1;
  return guarded([&] {
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) throw CudaError("no CUDA device");
    gemm_set_attributes();
    std::mt19937 rng(seed);
    std::normal_distribution<float> nd(0.f, 1.f);
    const int Mpad = (M + 255) / 256 * 256, Npad = (N + 255) / 256 * 256;
    std::vector<__nv_bfloat16> ha((size_t)Mpad * K), hw((size_t)Npad * K);
    for (auto& v : ha) v = __float2bfloat16(0.f);
    for (auto& v : hw) v = __float2bfloat16(0.f);
    for (int m = 0; m < M; ++m)
      for (int k = 0; k < K; ++k) ha[(size_t)m * K + k] = __float2bfloat16(nd(rng));
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) hw[(size_t)n * K + k] = __float2bfloat16(nd(rng) * 0.1f);
    std::vector<float> hb(Npad, 0.f), hres((size_t)M * N);
    for (int n = 0; n < N; ++n) hb[n] = nd(rng);
    for (auto& v : hres) v = nd(rng);
    __nv_bfloat16 *da, *dw, *dout_bf;
    float *db, *dref, *dout;
    float* dpv;
    int* dpi;
    CUDA_CHECK(cudaMalloc(&da, ha.size() * 2));
    CUDA_CHECK(cudaMalloc(&dw, hw.size() * 2));
    CUDA_CHECK(cudaMalloc(&db, hb.size() * 4));
    CUDA_CHECK(cudaMalloc(&dref, (size_t)M * N * 4));
    CUDA_CHECK(cudaMalloc(&dout, (size_t)M * N * 4));
    CUDA_CHECK(cudaMalloc(&dout_bf, (size_t)M * N * 2));
    const int n_tiles = (N + block_n - 1) / block_n * (block_n >= 128 ? 2 : 1);  // arg-max partials per row
    CUDA_CHECK(cudaMalloc(&dpv, (size_t)M * n_tiles * 4));
    CUDA_CHECK(cudaMalloc(&dpi, (size_t)M * n_tiles * 4));
    CUDA_CHECK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(dout, hres.data(), hres.size() * 4, cudaMemcpyHostToDevice));
    cudaStream_t s;
    CUDA_CHECK(cudaStreamCreate(&s));
    gemm_reference_simt(da, K, dw, K, db, dref, N, M, N, K, s);
    GemmOperandA a{};
    a.ptr = da, a.K = K, a.rows = Mpad, a.n_batch = 1, a.row_pitch = K, a.batch_pitch = (long)Mpad * K, a.n_taps = 0;
    const bool two_cta = block_n == 512;  // test-hook convention: block_n 512 selects the CTA-pair kernel (256 x 256 tiles)
    GemmPlan* plan = gemm_plan_create(a, dw, Npad, two_cta ? 256 : block_n, epilogue, two_cta);
    GemmParams p{};
    p.rows_valid = M, p.N = N, p.ldo = N, p.bias = db, p.n_batch = 1;
    const bool bf_out = epilogue == EPI_BIAS_BF16 || epilogue == EPI_BIAS_GELU_BF16;
    p.out = bf_out ? (void*)dout_bf : (void*)dout;
    p.part_val = dpv, p.part_idx = dpi, p.part_ld = n_tiles;
    gemm_launch(plan, p, s);
    CUDA_CHECK(cudaStreamSynchronize(s));
    std::vector<float> ref((size_t)M * N), got((size_t)M * N);
    CUDA_CHECK(cudaMemcpy(ref.data(), dref, ref.size() * 4, cudaMemcpyDeviceToHost));
    if (bf_out) {
      std::vector<__nv_bfloat16> gb((size_t)M * N);
      CUDA_CHECK(cudaMemcpy(gb.data(), dout_bf, gb.size() * 2, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < gb.size(); ++i) got[i] = __bfloat162float(gb[i]);
    } else {
      CUDA_CHECK(cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost));
    }
    double md = 0, mr = 0;
    if (epilogue == EPI_ARGMAX) {
      std::vector<float> pv((size_t)M * n_tiles);
      std::vector<int> pi((size_t)M * n_tiles);
      CUDA_CHECK(cudaMemcpy(pv.data(), dpv, pv.size() * 4, cudaMemcpyDeviceToHost));
      CUDA_CHECK(cudaMemcpy(pi.data(), dpi, pi.size() * 4, cudaMemcpyDeviceToHost));
      for (int m = 0; m < M; ++m) {
        // logits are also stored (p.out != null): compare them, and the partial argmax against the stored row
        float best = -3.4e38f;
        int bi = -1;
        for (int n = 0; n < N; ++n) {
          const float g = got[(size_t)m * N + n];
          md = std::max(md, (double)fabsf(g - ref[(size_t)m * N + n]));
          mr = std::max(mr, (double)fabsf(ref[(size_t)m * N + n]));
          if (g > best) best = g, bi = n;
        }
        float pb = -3.4e38f;
        int pbi = 0x7fffffff;
        for (int t = 0; t < n_tiles; ++t) {
          const float v = pv[(size_t)m * n_tiles + t];
          const int ix = pi[(size_t)m * n_tiles + t];
          if (v > pb || (v == pb && ix < pbi)) pb = v, pbi = ix;
        }
        if (pbi != bi) md = std::max(md, 1e9);  // flag an argmax mismatch loudly
      }
    } else {
      for (size_t i = 0; i < ref.size(); ++i) {
        float r = ref[i];
        if (epilogue == EPI_BIAS_GELU_BF16) r = 0.5f * r * (1.f + erff(r * 0.70710678f));
        if (epilogue == EPI_BIAS_RESID_F32) r += hres[i];
        md = std::max(md, (double)fabsf(got[i] - r));
        mr = std::max(mr, (double)fabsf(r));
      }
    }
    // independent of any CUDA code: fp64 host evaluation of sample rows from the same bf16 inputs (the SIMT comparator above
    // covers every element; this pins both kernels to the definition C = A W^T + bias)
    {
      std::vector<int> rows = {0, M / 3, (2 * M) / 3, M - 1};
      for (int i = 0; i < 4; ++i) rows.push_back((int)(rng() % (unsigned)M));
      for (int m : rows) {
        for (int n = 0; n < N; ++n) {
          double acc = hb[n];
          for (int k = 0; k < K; ++k) acc += (double)__bfloat162float(ha[(size_t)m * K + k]) * (double)__bfloat162float(hw[(size_t)n * K + k]);
          if (epilogue == EPI_BIAS_GELU_BF16) acc = 0.5 * acc * (1.0 + erf(acc * 0.7071067811865476));
          if (epilogue == EPI_BIAS_RESID_F32) acc += hres[(size_t)m * N + n];
          const double g = got[(size_t)m * N + n];
          const double tol_bf16 = (epilogue == EPI_BIAS_BF16 || epilogue == EPI_BIAS_GELU_BF16) ? fabs(acc) * 0.00390625 : 0.0;  // 2^-8: bf16 output
          md = std::max(md, std::max(0.0, fabs(g - acc) - tol_bf16));
        }
      }
    }
    *max_abs_diff = (float)md;
    *max_abs_ref = (float)mr;
    gemm_plan_destroy(plan);
    cudaStreamDestroy(s);
    cudaFree(da), cudaFree(dw), cudaFree(db), cudaFree(dref), cudaFree(dout), cudaFree(dout_bf), cudaFree(dpv), cudaFree(dpi);
  });
}

}  // extern "C"
