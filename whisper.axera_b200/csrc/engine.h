// The B200 Whisper engine: the device-side equivalent of the reference's `Whisper` class
// (/root/reference/cpp/src/Whisper.hpp:28-59, Whisper.cpp) with its two AxModelRunner members
// (/root/reference/cpp/src/ax_model_runner/ax_model_runner.hpp:29-61) replaced by CUDA kernels on one stream.
// Everything is batched over B independent utterances / 30 s windows; B = 1 reproduces the reference call.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "kernels.h"

namespace b200w {

constexpr int kAudioCtx = 1500;   // n_audio_ctx
constexpr int kTextCtx = 448;     // n_text_ctx
constexpr int kMelFrames = 3000;
constexpr int kChunkSamples = 480000;
constexpr int kSotLen = 4;        // {sot, language, transcribe, no_timestamps}, Whisper.cpp:139

struct ModelConfig {              // the keys Whisper::load_models reads (Whisper.cpp:93-137) + the dims our kernels need
  int n_mels = 0, n_vocab = 0, d = 0, n_head = 0, l_enc = 0, l_dec = 0;
  int n_text_ctx = kTextCtx, n_audio_ctx = kAudioCtx;
  int sot = 0, eot = 0, transcribe = 0, no_timestamps = 0;
  std::vector<int> lang_tokens;
  std::vector<std::string> lang_codes;
};

// parses {root}/{type}/{type}_config.json (keys of /root/reference/model_convert/export_onnx.py:592-625) and validates it
ModelConfig load_model_config(const std::string& model_root, const std::string& model_type);

struct HostTensor {
  std::vector<size_t> dims;
  const float* data = nullptr;  // points into the mmapped / loaded file image
  size_t numel() const {
    size_t n = 1;
    for (size_t d : dims) n *= d;
    return n;
  }
};
// flat weight file written by tools/make_model.py ("B200W001")
struct WeightFile {
  std::vector<unsigned char> blob;
  std::map<std::string, HostTensor> tensors;
  void load(const std::string& path);
  const HostTensor& get(const std::string& name) const;
};

struct DecodeOptions {
  int max_new_tokens = kTextCtx - kSotLen;  // reference: offset < n_text_ctx (Whisper.cpp:219)
  bool honor_eot = true;
  const int* forced_tokens = nullptr;       // [B][forced_len] teacher forcing (parity tests) or null
  int forced_len = 0;
  float* logits_out = nullptr;              // host [n_steps][B][n_vocab] (parity tests) or null
  const int* logit_rows = nullptr;          // with logits_out: only these sequences are copied, logits_out is [n_steps][n_logit_rows][n_vocab]
  int n_logit_rows = 0;
  bool use_graph = true;
};

struct StageTimes {
  float h2d_ms = 0, mel_ms = 0, encoder_ms = 0, decode_ms = 0, d2h_ms = 0, total_ms = 0;
  int decode_steps = 0;
  long kernel_launches = 0;
};

// number of visible CUDA devices (0 when the driver reports none)
int device_count();

class Engine {
 public:
  Engine(const std::string& model_root, const std::string& model_type, int device, int max_batch);
  ~Engine();
  Engine(const Engine&) = delete;

  const ModelConfig& config() const { return cfg_; }
  int device() const { return device_; }
  cudaStream_t stream() const { return stream_; }
  int capacity() const { return cap_; }
  void ensure_capacity(int B, long max_samples = kChunkSamples);

  std::vector<int> sot_sequence(const std::string& lang, std::string* resolved_lang = nullptr) const;

  // ---- stages on device-resident buffers (all enqueue on stream(), no host sync) ----
  float* pcm_dev() { return pcm_; }                // [cap][pcm_stride]
  long pcm_stride() const { return pcm_stride_; }
  int* n_samples_dev() { return n_samples_; }      // [cap]
  float* mel_dev() { return mel_; }                // [cap][n_mels][3000]
  void run_logmel(int B, int max_samples);         // pcm_ -> mel_ (+ bf16 time-major copy for conv1)
  void run_mel_convert(int B);                     // mel_ (f32, caller supplied) -> bf16 time-major copy
  void run_encoder(int B);                         // -> cross K/V cache
  void run_logmel_range(int b0, int nb, int max_samples);
  void run_encoder_range(int b0, int nb);          // one encoder sub-batch (nb <= enc_sub)
  // greedy loop; returns number of decoder steps executed. Tokens land in host vector per sequence.
  int run_decode(int B, const std::vector<int>& sot, const DecodeOptions& opt, std::vector<std::vector<int>>* tokens);

  // ---- model-ABI style single step on the resident caches (parity tests / b200w_decoder_loop) ----
  void decode_reset(int B);
  void decode_step_tokens(int B, const int* tokens_host, int offset, float* logits_host /*[B][n_vocab]*/, float* this_k /*[L][B][d]*/,
                          float* this_v);
  // caches across the model-ABI boundary, f32 in the reference's layouts (export_onnx.py:587-588, :668-670)
  void read_cross_kv(int b0, int nb, float* cross_k /*[L][nb][1500][d]*/, float* cross_v) const;  // sequences [b0, b0 + nb)
  void load_cross_kv(int B, const float* cross_k /*[L][B][1500][d]*/, const float* cross_v);
  void read_self_kv(int B, int n_rows, float* self_k /*[L][B][n_rows][d]*/, float* self_v) const;
  void load_self_kv(int B, int n_valid, const float* self_k /*[L][B][448][d]*/, const float* self_v);  // rows [0, n_valid)
  void read_encoder_hidden(int B, float* out /*[B][1500][d]*/) const;  // ln_post input (residual stream) for diagnostics

  // ---- whole pipeline with host buffers (copies inside) ----
  void transcribe(const float* const* pcm, const int* n_samples, int B, const std::string& lang, const DecodeOptions& opt,
                  std::vector<std::vector<int>>* tokens, StageTimes* times);
  // same, PCM already resident in pcm_dev()/n_samples_dev() (bench "value" leg)
  void transcribe_resident(int B, int max_samples, const std::string& lang, const DecodeOptions& opt,
                           std::vector<std::vector<int>>* tokens, StageTimes* times);

  // profiling helper: only the cross-attention decode kernel, once per decoder layer, on the resident K/V
  void run_cross_attention_only(int B);

  long launches() const { return launches_; }
  long compactions() const { return compactions_; }   // EOT compactions of the slot list so far
  int last_active() const { return last_active_; }    // sequences still decoding when the last run_decode stopped

 private:
  struct LayerEnc;
  struct LayerDec;
  void load_weights(const std::string& dir, const std::string& type);
  void free_workspace();
  void check_batch(int B, const char* what) const;
  void reset_slot_map();
  void export_cache(const __nv_bfloat16* cache, int T, int b0, int nb, int n_rows, float* out) const;
  void import_cache(__nv_bfloat16* cache, int T, int B, int n_rows, const float* in);
  void allocate_workspace(int new_cap, long new_stride);
  void build_plans();
  void enqueue_decode_step(int B, bool want_logits, bool finalize, int honor_eot, int n_fused = 1, bool need_embed = true);
  void enqueue_embed_ln(int B);

  ModelConfig cfg_;
  int device_ = 0;
  cudaStream_t stream_ = nullptr;
  cudaStream_t stream2_ = nullptr;          // second micro-batch of a decoder step
  int* step_ctr_ = nullptr;                 // [4] decoder position of each micro-batch
  int* boundary_ticket_ = nullptr;          // [4] arrival counters of the step-boundary kernels
  int graph_steps_ = 8;                     // decoder steps captured per CUDA graph (B200W_GRAPH_STEPS)
  int* cross_work_ = nullptr;               // [l_dec][4 micro-batches][2] work counters (+ one pair for time_stage)
  int prio_high_ = 0;                       // most urgent launch priority of the device (cudaDeviceGetStreamPriorityRange)
  std::vector<cudaEvent_t> step_events_;    // fork / join / per-layer cross-attention hand-over
  std::vector<cudaEvent_t> copy_events_;    // per encoder sub-batch: its PCM has been copied (pipelined transcribe())
  bool micro_batch_ = true;
  int n_micro_batch_ = 2;                   // micro-batches of a decoder step (B200W_N_MICROBATCH, 1..4)
  bool cross_chain_forced_ = false;         // B200W_CROSS_CHAIN: hand over regardless of the launch size
  bool cross_chain_ = true;                 // hand the cross-attention kernels over micro-batch to micro-batch with events
  cudaStream_t mb_streams_[2] = {nullptr, nullptr};  // streams of micro-batches 2 and 3
  int cap_ = 0;
  int enc_sub_ = 0;
  bool attn_mma_sync_ = false;
  long pcm_stride_ = 0;
  long launches_ = 0;

  // weights (device)
  std::vector<void*> owned_;  // every cudaMalloc'ed weight pointer
  __nv_bfloat16 *w_conv1_ = nullptr, *w_conv2_ = nullptr, *w_crosskv_ = nullptr, *w_emb_bf16_ = nullptr;
  float *b_conv1_ = nullptr, *b_conv2_ = nullptr, *b_crosskv_ = nullptr, *pos_audio_ = nullptr;
  float *ln_post_g_ = nullptr, *ln_post_b_ = nullptr, *dec_ln_g_ = nullptr, *dec_ln_b_ = nullptr;
  float *emb_f32_ = nullptr, *pos_text_ = nullptr;
  int vocab_pad_ = 0;
  std::vector<LayerEnc> enc_;
  std::vector<LayerDec> dec_;

  // workspace (device), sized for cap_
  float *pcm_ = nullptr, *mel_ = nullptr;
  int *n_samples_ = nullptr, *utt_state_ = nullptr;  // utt_state_: [cap][2] per-utterance scratch of the log-mel kernel
  __nv_bfloat16 *mel_tm_ = nullptr, *conv1_out_ = nullptr;
  float* x_enc_ = nullptr;
  __nv_bfloat16 *h_enc_ = nullptr, *qkv_enc_ = nullptr, *attn_enc_ = nullptr, *mlp_enc_ = nullptr;
  __nv_bfloat16 *cross_k_ = nullptr, *cross_v_ = nullptr, *self_k_ = nullptr, *self_v_ = nullptr;
  float *x_dec_ = nullptr, *qkv_dec_ = nullptr, *q_dec_ = nullptr, *logits_ = nullptr;
  __nv_bfloat16 *h_dec_ = nullptr, *attn_dec_ = nullptr, *mlp_dec_ = nullptr;
  float* part_val_ = nullptr;
  int* part_idx_ = nullptr;
  DecodeState st_{};
  int dec_rows_pad_ = 0;
  int logits_tiles_ = 0;
  std::vector<void*> ws_owned_;
  std::vector<GemmPlan*> plans_;
  // plans
  GemmPlan *p_conv1_ = nullptr, *p_conv2_ = nullptr, *p_crosskv_ = nullptr, *p_logits_ = nullptr;
  struct EncPlans {
    GemmPlan *qkv, *out, *fc1, *fc2;
  };
  struct DecPlans {
    GemmPlan *qkv, *out, *cq, *co, *fc1, *fc2;
  };
  std::vector<EncPlans> enc_plans_;
  std::vector<DecPlans> dec_plans_;
  // decode graph cache (keyed by batch size)
  std::map<int, cudaGraphExec_t> graphs_;
  std::map<int, long> per_step_launches_;
  static constexpr int kMaxPolled = 4096;   // EOT polling / compaction covers batches up to this size
  int* pinned_flags_ = nullptr;             // [kMaxPolled] finished flags, [kMaxPolled] new slot map (pinned host)
  int* pinned_map_ = nullptr;
  int* slot_seq_ = nullptr;                 // device [cap]: sequence of each decoder slot
  bool slot_map_identity_ = false;
  long compactions_ = 0;
  int last_active_ = 0;
};

}  // namespace b200w
