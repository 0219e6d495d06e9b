// libax_whisper.so outer boundary: the reference's four AX_WHISPER_* entry points
// (/root/reference/cpp/src/api/ax_whisper_api.cpp:48-163) over the B200 engine, plus batched extensions.
// Also holds the small host pieces the reference takes from third-party headers: a WAV reader (the reference uses
// AudioFile.h, GPLv3 -- re-written, not copied), the token table loader (Whisper.cpp:115-127) and a length-safe
// base64 decoder (the reference's writes into char[32], Whisper.cpp:226-228).
#include "../../include/ax_whisper_api.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "engine.h"
#include "host_utils.h"

using namespace b200w;

namespace {

thread_local std::string g_api_err;

// One pending AX_WHISPER_RunPCM / RunFile call (the reference serves these one at a time and its handle is not re-entrant,
// SURVEY.md section 8b "Threading"; its HTTP server calls it from a thread pool regardless).
struct PendingRequest {
  const float* pcm = nullptr;
  int n_samples = 0;
  std::vector<int> tokens;
  int rc = 0;
  std::string err;
  bool filled = false;  // rc / tokens / err are final
  bool done = false;
};

struct WhisperHandle {
  std::unique_ptr<Engine> engine;                // device 0 of the handle (config, SOT sequence)
  std::vector<std::unique_ptr<Engine>> extra;    // further devices (B200W_DEVICES): own weights, contexts, streams
  std::vector<std::unique_ptr<std::mutex>> extra_mu;
  TokenTable token_table;                // base64 text per id (line index = id, Whisper.cpp:123-126)
  std::string lang;
  std::mutex mu;                         // the engine (one GPU pass at a time)
  // request coalescing: calls that arrive while a pass is running are transcribed together in the next pass
  std::mutex qmu;
  std::condition_variable qcv;
  std::deque<PendingRequest*> queue;
  bool leader_active = false;
  int coalesce_max = 64;                 // B200W_COALESCE_MAX
  int coalesce_wait_us = 0;              // B200W_COALESCE_WAIT_US: extra time a leader waits for company
  long n_requests = 0, n_passes = 0;
};

void set_err(const std::string& s) {
  g_api_err = s;
  fprintf(stderr, "[ax_whisper] %s\n", s.c_str());
}

std::string detokenize(const WhisperHandle& h, const std::vector<int>& ids) {
  // The reference converts Traditional -> Simplified Chinese with OpenCC when lang == "zh" (Whisper.cpp:231-236); its
  // prebuilt OpenCC is AArch64-only and the conversion is text cosmetics, not arithmetic: identity here.
  return h.token_table.detokenize(ids.data(), ids.size());
}

int run_on(Engine* eng, std::mutex* mu, const std::string& lang, const float* const* pcm, const int* n_samples, int B,
           const DecodeOptions& opt, std::vector<std::vector<int>>* toks, std::string* err) {
  try {
    std::lock_guard<std::mutex> lock(*mu);
    eng->transcribe(pcm, n_samples, B, lang, opt, toks, nullptr);
    return 0;
  } catch (const std::exception& ex) {
    *err = std::string("run whisper failed: ") + ex.what();
    return -1;
  } catch (...) {
    *err = "run whisper failed: unknown error";
    return -1;
  }
}

// One pass over `B` utterances.  With several devices on the handle the utterances are split into contiguous shards, one
// host thread per GPU, no data-path collective; token lists come back in utterance order.
int run_batch(WhisperHandle* h, const float* const* pcm, const int* n_samples, int B, const DecodeOptions& opt,
              std::vector<std::vector<int>>* toks) {
  const int n_dev = 1 + (int)h->extra.size();
  const int G = std::min(n_dev, B);
  std::string err;
  if (G <= 1) {
    const int rc = run_on(h->engine.get(), &h->mu, h->lang, pcm, n_samples, B, opt, toks, &err);
    if (rc != 0) set_err(err);
    return rc;
  }
  std::vector<std::vector<std::vector<int>>> part(G);
  std::vector<std::string> errs(G);
  std::vector<int> rcs(G, -1);
  auto shard = [&](int g) { return (int)((long)B * g / G); };
  {
    // joins whatever was started, also when a later std::thread constructor throws (a joinable thread that is destroyed
    // would call std::terminate)
    struct Joiner {
      std::vector<std::thread> t;
      ~Joiner() {
        for (auto& w : t)
          if (w.joinable()) w.join();
      }
    } workers;
    workers.t.reserve(G);
    try {
      for (int g = 0; g < G; ++g) {
        workers.t.emplace_back([&, g] {
          Engine* eng = g == 0 ? h->engine.get() : h->extra[g - 1].get();
          std::mutex* mu = g == 0 ? &h->mu : h->extra_mu[g - 1].get();
          const int b0 = shard(g), b1 = shard(g + 1);
          rcs[g] = run_on(eng, mu, h->lang, pcm + b0, n_samples + b0, b1 - b0, opt, &part[g], &errs[g]);
        });
      }
    } catch (const std::exception& ex) {
      errs[0] = std::string("run whisper failed: cannot start a worker thread: ") + ex.what();
      rcs[0] = -1;
    }
  }
  toks->clear();
  for (int g = 0; g < G; ++g) {
    if (rcs[g] != 0) {
      set_err(errs[g].empty() ? "run whisper failed: worker did not run" : errs[g]);
      return -1;
    }
  }
  for (int g = 0; g < G; ++g)
    for (auto& t : part[g]) toks->push_back(std::move(t));
  return 0;
}

}  // namespace

namespace {

// Single-utterance entry: enqueue, then either lead (take everything queued, run one batched pass) or wait for a leader.
int run_coalesced(WhisperHandle* h, const float* pcm, int n_samples, std::vector<int>* toks) {
  PendingRequest req;
  req.pcm = pcm, req.n_samples = n_samples;
  std::unique_lock<std::mutex> lk(h->qmu);
  h->queue.push_back(&req);
  ++h->n_requests;
  while (!req.done) {
    if (h->leader_active) {
      h->qcv.wait(lk);
      continue;
    }
    h->leader_active = true;
    if (h->coalesce_wait_us > 0) {
      lk.unlock();
      std::this_thread::sleep_for(std::chrono::microseconds(h->coalesce_wait_us));
      lk.lock();
    }
    std::vector<PendingRequest*> batch;
    while (!h->queue.empty() && (int)batch.size() < h->coalesce_max) {
      batch.push_back(h->queue.front());
      h->queue.pop_front();
    }
    ++h->n_passes;
    lk.unlock();
    // Nothing below may leave this scope without marking the batch done and handing the leadership back: a waiter that is
    // never woken (or a stuck leader_active flag) would block every later call on this handle for ever.
    try {
      // requests that cannot be transcribed (too short) fail alone, not the whole pass
      std::vector<const float*> ptrs;
      std::vector<int> lens;
      std::vector<PendingRequest*> ok;
      for (PendingRequest* r : batch) {
        if (r->n_samples < 201) {
          r->rc = -1, r->err = "run whisper failed: audio shorter than 201 samples (reflect padding needs n_fft/2 + 1)";
          r->filled = true;
        } else {
          ptrs.push_back(r->pcm), lens.push_back(r->n_samples), ok.push_back(r);
        }
      }
      if (!ok.empty()) {
        std::vector<std::vector<int>> out;
        const int rc = run_batch(h, ptrs.data(), lens.data(), (int)ok.size(), DecodeOptions(), &out);
        for (size_t i = 0; i < ok.size(); ++i) {
          ok[i]->rc = rc;
          if (rc == 0) ok[i]->tokens = std::move(out[i]);
          else ok[i]->err = g_api_err;
          ok[i]->filled = true;
        }
      }
    } catch (const std::exception& ex) {
      for (PendingRequest* r : batch)
        if (!r->filled) r->rc = -1, r->err = std::string("run whisper failed: ") + ex.what();
    } catch (...) {
      for (PendingRequest* r : batch)
        if (!r->filled) r->rc = -1, r->err = "run whisper failed: unknown error";
    }
    lk.lock();
    for (PendingRequest* r : batch) r->done = true;
    h->leader_active = false;
    h->qcv.notify_all();
  }
  lk.unlock();
  if (req.rc != 0) {
    if (!req.err.empty()) set_err(req.err);
    return -1;
  }
  *toks = std::move(req.tokens);
  return 0;
}

}  // namespace

extern "C" {

AX_WHISPER_API const char* AX_WHISPER_LastError(void) { return g_api_err.c_str(); }

AX_WHISPER_API AX_WHISPER_HANDLE AX_WHISPER_Init(const char* model_type, const char* model_path, const char* language) {
  if (!model_type || !model_path || !language) return nullptr;
  try {
    g_api_err.clear();
    std::unique_ptr<WhisperHandle> h(new WhisperHandle());
    const char* dev_env = getenv("B200W_DEVICE");
    const char* mb_env = getenv("B200W_MAX_BATCH");
    // B200W_DEVICES = "all" or "0,2,3": one engine (weights, context, streams) per listed GPU; batches are sharded over them
    std::vector<int> devices;
    if (const char* devs = getenv("B200W_DEVICES")) {
      if (std::string(devs) == "all") {
        const int n = device_count();
        for (int i = 0; i < n; ++i) devices.push_back(i);
      } else {
        std::string cur;
        for (const char* c = devs;; ++c) {
          if (*c == ',' || *c == 0) {
            if (!cur.empty()) devices.push_back(atoi(cur.c_str()));
            cur.clear();
            if (*c == 0) break;
          } else {
            cur += *c;
          }
        }
      }
    }
    if (devices.empty()) devices.push_back(dev_env ? atoi(dev_env) : 0);
    const int max_batch = mb_env ? atoi(mb_env) : 1;
    h->engine.reset(new Engine(model_path, model_type, devices[0], max_batch));
    for (size_t i = 1; i < devices.size(); ++i) {
      h->extra.emplace_back(new Engine(model_path, model_type, devices[i], max_batch));
      h->extra_mu.emplace_back(new std::mutex());
    }
    const std::string token_path = std::string(model_path) + "/" + model_type + "/" + model_type + "-tokens.txt";
    std::string terr;
    if (!h->token_table.load(token_path, &terr)) {
      set_err(terr);
      return nullptr;
    }
    h->engine->sot_sequence(language, &h->lang);  // resolves the "unknown language -> zh" fallback once
    if (const char* e = getenv("B200W_COALESCE_MAX")) h->coalesce_max = std::max(1, atoi(e));
    if (const char* e = getenv("B200W_COALESCE_WAIT_US")) h->coalesce_wait_us = std::max(0, atoi(e));
    return static_cast<AX_WHISPER_HANDLE>(h.release());
  } catch (const std::exception& ex) {
    set_err(std::string("load models failed: ") + ex.what());
    return nullptr;  // unlike the reference (ax_whisper_api.cpp:49-53) nothing is leaked on failure
  } catch (...) {
    set_err("load models failed: unknown error");
    return nullptr;
  }
}

AX_WHISPER_API void AX_WHISPER_Uninit(AX_WHISPER_HANDLE handle) {
  if (handle) delete static_cast<WhisperHandle*>(handle);
}

AX_WHISPER_API int AX_WHISPER_RunPCM(AX_WHISPER_HANDLE handle, float* pcm_data, int num_samples, char** result) {
  if (!handle || !pcm_data || !result) return -1;
  *result = nullptr;
  WhisperHandle* h = static_cast<WhisperHandle*>(handle);
  std::vector<int> toks;
  if (run_coalesced(h, pcm_data, num_samples, &toks) != 0) return -1;
  *result = strdup(detokenize(*h, toks).c_str());
  return *result ? 0 : -1;
}

AX_WHISPER_API int AX_WHISPER_RunFile(AX_WHISPER_HANDLE handle, const char* wav_file, char** result) {
  if (!handle || !wav_file || !result) return -1;
  *result = nullptr;
  WavData wav;
  std::string err;
  if (!load_wav(wav_file, &wav, &err)) {
    set_err("load wav failed: " + err);
    return -1;
  }
  if (wav.sample_rate != 16000)
    fprintf(stderr, "[ax_whisper] warning: %s is %d Hz; it is treated as 16 kHz like the reference does\n", wav_file, wav.sample_rate);
  // mono mix only for stereo; with more channels the first one is used (ax_whisper_api.cpp:105-113)
  std::vector<float>& s = wav.channels[0];
  if (wav.channels.size() == 2)
    for (size_t i = 0; i < s.size(); ++i) s[i] = (s[i] + wav.channels[1][i]) / 2;
  return AX_WHISPER_RunPCM(handle, s.data(), (int)s.size(), result);
}

AX_WHISPER_API int AX_WHISPER_RunPCMBatch(AX_WHISPER_HANDLE handle, const float* const* pcm_data, const int* num_samples, int batch,
                                          char** results) {
  if (!handle || !pcm_data || !num_samples || !results || batch <= 0) return -1;
  for (int i = 0; i < batch; ++i) results[i] = nullptr;
  WhisperHandle* h = static_cast<WhisperHandle*>(handle);
  std::vector<std::vector<int>> toks;
  if (run_batch(h, pcm_data, num_samples, batch, DecodeOptions(), &toks) != 0) return -1;
  for (int i = 0; i < batch; ++i) results[i] = strdup(detokenize(*h, toks[i]).c_str());
  return 0;
}

AX_WHISPER_API int AX_WHISPER_RunPCMLong(AX_WHISPER_HANDLE handle, const float* pcm_data, long num_samples, int window_batch, char** result) {
  if (!handle || !pcm_data || !result) return -1;
  *result = nullptr;
  WhisperHandle* h = static_cast<WhisperHandle*>(handle);
  std::vector<const float*> ptrs;
  std::vector<int> lens;
  for (long off = 0; off < num_samples; off += kChunkSamples) {
    const long n = std::min<long>(kChunkSamples, num_samples - off);
    if (n < 201) break;
    ptrs.push_back(pcm_data + off);
    lens.push_back((int)n);
  }
  if (ptrs.empty()) {
    set_err("audio shorter than 201 samples");
    return -1;
  }
  const int total = (int)ptrs.size();
  const int step = window_batch > 0 ? window_batch : total;
  std::string text;
  for (int i = 0; i < total; i += step) {
    const int nb = std::min(step, total - i);
    std::vector<std::vector<int>> toks;
    if (run_batch(h, ptrs.data() + i, lens.data() + i, nb, DecodeOptions(), &toks) != 0) return -1;
    for (int b = 0; b < nb; ++b) text += detokenize(*h, toks[b]);
  }
  *result = strdup(text.c_str());
  return *result ? 0 : -1;
}

AX_WHISPER_API int AX_WHISPER_GetStats(AX_WHISPER_HANDLE handle, long* n_requests, long* n_gpu_passes) {
  if (!handle) return -1;
  WhisperHandle* h = static_cast<WhisperHandle*>(handle);
  std::lock_guard<std::mutex> lk(h->qmu);
  if (n_requests) *n_requests = h->n_requests;
  if (n_gpu_passes) *n_gpu_passes = h->n_passes;
  return 0;
}

AX_WHISPER_API int AX_WHISPER_RunPCMTokens(AX_WHISPER_HANDLE handle, const float* const* pcm_data, const int* num_samples, int batch,
                                           int max_new_tokens, int honor_eot, int* tokens, int max_tokens, int* n_tokens) {
  if (!handle || !pcm_data || !num_samples || !tokens || !n_tokens || batch <= 0 || max_tokens <= 0) return -1;
  WhisperHandle* h = static_cast<WhisperHandle*>(handle);
  DecodeOptions opt;
  if (max_new_tokens > 0) opt.max_new_tokens = std::min(max_new_tokens, kTextCtx - kSotLen);
  opt.honor_eot = honor_eot != 0;
  std::vector<std::vector<int>> toks;
  if (run_batch(h, pcm_data, num_samples, batch, opt, &toks) != 0) return -1;
  for (int b = 0; b < batch; ++b) {
    const int n = std::min<int>((int)toks[b].size(), max_tokens);
    for (int i = 0; i < n; ++i) tokens[(size_t)b * max_tokens + i] = toks[b][i];
    n_tokens[b] = n;
  }
  return 0;
}

}  // extern "C"
