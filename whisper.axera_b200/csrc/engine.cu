// Engine implementation: model loading, HBM layout, stage orchestration, CUDA-graphed greedy loop.
// Reference anchors: Whisper::load_models (/root/reference/cpp/src/Whisper.cpp:86-149), Whisper::run (:186-239),
// Whisper::run_decoder (:290-346); graph semantics from /root/reference/model_convert/export_onnx.py:153-387.
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <fstream>
#include <sstream>

#include "common.cuh"
#include "json_min.h"

namespace b200w {

// ---------------------------------------------------------------------------------------------------------
// weight file
// ---------------------------------------------------------------------------------------------------------
void WeightFile::load(const std::string& path) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) throw std::runtime_error("cannot open weight file: " + path);
  const size_t size = (size_t)f.tellg();
  f.seekg(0);
  blob.resize(size);
  f.read(reinterpret_cast<char*>(blob.data()), (std::streamsize)size);
  if (size < 12 || memcmp(blob.data(), "B200W001", 8) != 0) throw std::runtime_error("bad weight file magic: " + path);
  size_t pos = 8;
  auto rd32 = [&]() {
    if (pos + 4 > size) throw std::runtime_error("truncated weight file: " + path);
    uint32_t v;
    memcpy(&v, blob.data() + pos, 4);
    pos += 4;
    return v;
  };
  auto rd64 = [&]() {
    if (pos + 8 > size) throw std::runtime_error("truncated weight file: " + path);
    uint64_t v;
    memcpy(&v, blob.data() + pos, 8);
    pos += 8;
    return v;
  };
  const uint32_t n = rd32();
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t nl = rd32();
    if (pos + nl > size) throw std::runtime_error("truncated weight file: " + path);
    std::string name(reinterpret_cast<const char*>(blob.data() + pos), nl);
    pos += nl;
    const uint32_t dtype = rd32(), ndim = rd32();
    if (dtype != 0 || ndim > 4) throw std::runtime_error("unsupported tensor in weight file: " + name);
    HostTensor t;
    for (uint32_t k = 0; k < ndim; ++k) t.dims.push_back((size_t)rd64());
    const uint64_t off = rd64(), nbytes = rd64();
    if (off + nbytes > size || nbytes != t.numel() * 4) throw std::runtime_error("bad tensor extent in weight file: " + name);
    t.data = reinterpret_cast<const float*>(blob.data() + off);
    tensors[name] = t;
  }
}
const HostTensor& WeightFile::get(const std::string& name) const {
  auto it = tensors.find(name);
  if (it == tensors.end()) throw std::runtime_error("missing tensor in weight file: " + name);
  return it->second;
}

// ---------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------
namespace {

inline uint16_t f32_to_bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);  // inf / nan
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

template <class T>
T* dev_alloc(std::vector<void*>& owner, size_t n, bool zero = true) {
  void* p = nullptr;
  CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  if (zero) CUDA_CHECK(cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(T)));
  owner.push_back(p);
  return reinterpret_cast<T*>(p);
}

float* upload_f32(std::vector<void*>& owner, const float* src, size_t n) {
  float* d = dev_alloc<float>(owner, n, false);
  CUDA_CHECK(cudaMemcpy(d, src, n * sizeof(float), cudaMemcpyHostToDevice));
  return d;
}
__nv_bfloat16* upload_bf16(std::vector<void*>& owner, const std::vector<float>& src, size_t pad_to = 0) {
  const size_t n = std::max(src.size(), pad_to);
  std::vector<uint16_t> h(n, 0);
  for (size_t i = 0; i < src.size(); ++i) h[i] = f32_to_bf16_rne(src[i]);
  __nv_bfloat16* d = dev_alloc<__nv_bfloat16>(owner, n, false);
  CUDA_CHECK(cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice));
  return d;
}
std::vector<float> to_vec(const HostTensor& t) { return std::vector<float>(t.data, t.data + t.numel()); }
void append(std::vector<float>& dst, const HostTensor& t) { dst.insert(dst.end(), t.data, t.data + t.numel()); }

std::vector<std::string> split_csv(const std::string& s) {
  std::vector<std::string> out;
  std::stringstream ss(s);
  std::string tok;
  while (std::getline(ss, tok, ',')) out.push_back(tok);
  return out;
}

}  // namespace

struct Engine::LayerEnc {
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  __nv_bfloat16 *w_qkv, *w_out, *w_fc1, *w_fc2;
  float *b_qkv, *b_out, *b_fc1, *b_fc2;
};
struct Engine::LayerDec {
  float *ln1_g, *ln1_b, *lnx_g, *lnx_b, *ln2_g, *ln2_b;
  __nv_bfloat16 *w_qkv, *w_out, *w_cq, *w_co, *w_fc1, *w_fc2;
  float *b_qkv, *b_out, *b_cq, *b_co, *b_fc1, *b_fc2;
};

// ---------------------------------------------------------------------------------------------------------
// construction / loading
// ---------------------------------------------------------------------------------------------------------
ModelConfig load_model_config(const std::string& model_root, const std::string& model_type) {
  ModelConfig c;
  const std::string cfg_path = model_root + "/" + model_type + "/" + model_type + "_config.json";
  std::ifstream cf(cfg_path);
  if (!cf) throw std::runtime_error("Cannot open config file: " + cfg_path);
  std::stringstream ss;
  ss << cf.rdbuf();
  const JsonFlat j = parse_flat_json(ss.str());
  c.n_mels = j.get_int("n_mels");
  c.n_vocab = j.get_int("n_vocab");
  c.d = j.get_int("n_text_state");
  c.n_text_ctx = j.get_int("n_text_ctx");
  c.l_dec = j.get_int("n_text_layer");
  c.l_enc = j.get_int("n_audio_layer");
  c.n_head = j.get_int("n_text_head");
  c.n_audio_ctx = j.has("n_audio_ctx") ? j.get_int("n_audio_ctx") : kAudioCtx;
  c.sot = j.get_int("sot");
  c.eot = j.get_int("eot");
  c.transcribe = j.get_int("transcribe");
  c.no_timestamps = j.get_int("no_timestamps");
  for (const std::string& t : split_csv(j.get_str("all_language_tokens"))) c.lang_tokens.push_back(std::stoi(t));
  c.lang_codes = split_csv(j.get_str("all_language_codes"));
  if (c.lang_tokens.size() != c.lang_codes.size()) throw std::runtime_error("config: language token / code lists differ in length");
  if (j.get_int("n_audio_state") != c.d || j.get_int("n_audio_head") != c.n_head)
    throw std::runtime_error("config: encoder and decoder widths differ (unsupported)");
  if (c.d != c.n_head * 64) throw std::runtime_error("config: head_dim must be 64");
  if (c.n_text_ctx != kTextCtx || c.n_audio_ctx != kAudioCtx) throw std::runtime_error("config: unexpected context sizes");
  if (c.n_mels != 80 && c.n_mels != 128) throw std::runtime_error("config: n_mels must be 80 or 128");
  return c;
}

Engine::Engine(const std::string& model_root, const std::string& model_type, int device, int max_batch) : device_(device) {
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    throw CudaError("no CUDA device available: this engine has no CPU path (cudaGetDeviceCount: " + std::string(cudaGetErrorString(e)) + ")");
  if (device < 0 || device >= n_dev) throw CudaError("invalid CUDA device index");
  CUDA_CHECK(cudaSetDevice(device_));
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device_));
  if (prop.major != 10) throw CudaError(std::string("this build contains sm_100a code only; device is ") + prop.name);
  CUDA_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  CUDA_CHECK(cudaStreamCreateWithFlags(&stream2_, cudaStreamNonBlocking));
  {
    int least = 0, greatest = 0;
    CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    prio_high_ = greatest;
  }
  micro_batch_ = getenv("B200W_NO_MICROBATCH") == nullptr;
  if (const char* e = getenv("B200W_N_MICROBATCH")) n_micro_batch_ = std::max(1, std::min(4, atoi(e)));
  cross_chain_ = getenv("B200W_NO_CROSS_CHAIN") == nullptr;
  cross_chain_forced_ = getenv("B200W_CROSS_CHAIN") != nullptr;
  if (const char* e = getenv("B200W_GRAPH_STEPS")) graph_steps_ = std::max(1, std::min(16, atoi(e)));
  for (auto& st : mb_streams_) CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&pinned_flags_), 2 * kMaxPolled * sizeof(int)));
  pinned_map_ = pinned_flags_ + kMaxPolled;

  const std::string dir = model_root + "/" + model_type;  // {root}/{type}/{type}-*  (Whisper.cpp:87-90)
  cfg_ = load_model_config(model_root, model_type);

  const char* attn_env = getenv("B200W_ATTN");
  attn_mma_sync_ = attn_env != nullptr && std::string(attn_env) == "mma_sync";
  logmel_upload_tables();
  kernels_set_attributes();
  load_weights(dir, model_type);
  step_events_.resize(8 + 4 * cfg_.l_dec);
  for (auto& e : step_events_) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  ensure_capacity(std::max(1, max_batch));
}

int device_count() {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

Engine::~Engine() {
  cudaSetDevice(device_);
  cudaStreamSynchronize(stream_);
  free_workspace();  // also destroys the captured decode graphs
  for (void* p : owned_) cudaFree(p);
  if (pinned_flags_) cudaFreeHost(pinned_flags_);
  for (auto& e : step_events_) cudaEventDestroy(e);
  for (auto& e : copy_events_) cudaEventDestroy(e);
  for (auto& st : mb_streams_)
    if (st) cudaStreamDestroy(st);
  if (stream2_) cudaStreamDestroy(stream2_);
  if (stream_) cudaStreamDestroy(stream_);
}

void Engine::load_weights(const std::string& dir, const std::string& type) {
  WeightFile fe, fd;
  fe.load(dir + "/" + type + "-encoder.b200w");
  fd.load(dir + "/" + type + "-decoder.b200w");
  const int d = cfg_.d, n_mels = cfg_.n_mels;
  auto check = [](const HostTensor& t, std::initializer_list<size_t> dims, const char* name) {
    if (t.dims != std::vector<size_t>(dims)) throw std::runtime_error(std::string("unexpected shape for ") + name);
  };
  // the weight files must describe exactly the architecture of the config: no blocks beyond n_audio_layer / n_text_layer
  // (a deeper checkpoint converted under a shallower config would otherwise load and produce garbage), every matrix of the
  // expected shape
  for (const WeightFile* f : {&fe, &fd})
    for (const auto& kv : f->tensors) {
      const std::string& n = kv.first;
      const bool enc = n.rfind("encoder.blocks.", 0) == 0, dec = n.rfind("decoder.blocks.", 0) == 0;
      if (enc || dec) {
        const int idx = atoi(n.c_str() + 15);
        if (idx >= (enc ? cfg_.l_enc : cfg_.l_dec))
          throw std::runtime_error("weight file holds " + n + " but the config declares only " + std::to_string(enc ? cfg_.l_enc : cfg_.l_dec) +
                                   " such blocks");
      }
      const size_t ud = (size_t)d;
      const auto ends = [&](const char* suf) { const size_t l = strlen(suf); return n.size() >= l && n.compare(n.size() - l, l, suf) == 0; };
      if (enc || dec) {
        if (ends(".query.weight") || ends(".key.weight") || ends(".value.weight") || ends(".out.weight")) check(kv.second, {ud, ud}, n.c_str());
        else if (ends(".mlp.0.weight")) check(kv.second, {4 * ud, ud}, n.c_str());
        else if (ends(".mlp.2.weight")) check(kv.second, {ud, 4 * ud}, n.c_str());
        else if (ends(".mlp.0.bias")) check(kv.second, {4 * ud}, n.c_str());
        else if (ends(".bias") || ends("_ln.weight")) check(kv.second, {ud}, n.c_str());
      }
    }
  // conv weights [d][C][3] -> [d][3][C] so that a tap is a contiguous K slice
  auto conv_reorder = [&](const HostTensor& t, int C) {
    std::vector<float> out((size_t)d * 3 * C);
    for (int o = 0; o < d; ++o)
      for (int c = 0; c < C; ++c)
        for (int k = 0; k < 3; ++k) out[((size_t)o * 3 + k) * C + c] = t.data[((size_t)o * C + c) * 3 + k];
    return out;
  };
  {
    const HostTensor& w1 = fe.get("encoder.conv1.weight");
    check(w1, {(size_t)d, (size_t)n_mels, 3}, "encoder.conv1.weight");
    w_conv1_ = upload_bf16(owned_, conv_reorder(w1, n_mels));
    b_conv1_ = upload_f32(owned_, fe.get("encoder.conv1.bias").data, d);
    const HostTensor& w2 = fe.get("encoder.conv2.weight");
    check(w2, {(size_t)d, (size_t)d, 3}, "encoder.conv2.weight");
    w_conv2_ = upload_bf16(owned_, conv_reorder(w2, d));
    b_conv2_ = upload_f32(owned_, fe.get("encoder.conv2.bias").data, d);
  }
  {  // sinusoids(1500, d), whisper/model.py
    std::vector<float> pos((size_t)kAudioCtx * d);
    const int half = d / 2;
    const float inc = logf(10000.0f) / (float)(half - 1);
    for (int t = 0; t < kAudioCtx; ++t)
      for (int i = 0; i < half; ++i) {
        const float inv = expf(-inc * (float)i);
        const float a = (float)t * inv;
        pos[(size_t)t * d + i] = sinf(a);
        pos[(size_t)t * d + half + i] = cosf(a);
      }
    pos_audio_ = upload_f32(owned_, pos.data(), pos.size());
  }
  auto qkv_pack = [&](const WeightFile& f, const std::string& p, __nv_bfloat16** w, float** b) {
    std::vector<float> wq;
    append(wq, f.get(p + ".query.weight"));
    append(wq, f.get(p + ".key.weight"));
    append(wq, f.get(p + ".value.weight"));
    std::vector<float> bq(3 * (size_t)d, 0.f);
    memcpy(bq.data(), f.get(p + ".query.bias").data, d * 4);
    memcpy(bq.data() + 2 * d, f.get(p + ".value.bias").data, d * 4);  // key has no bias
    *w = upload_bf16(owned_, wq);
    *b = upload_f32(owned_, bq.data(), bq.size());
  };
  enc_.resize(cfg_.l_enc);
  for (int i = 0; i < cfg_.l_enc; ++i) {
    const std::string p = "encoder.blocks." + std::to_string(i);
    LayerEnc& L = enc_[i];
    L.ln1_g = upload_f32(owned_, fe.get(p + ".attn_ln.weight").data, d);
    L.ln1_b = upload_f32(owned_, fe.get(p + ".attn_ln.bias").data, d);
    qkv_pack(fe, p + ".attn", &L.w_qkv, &L.b_qkv);
    L.w_out = upload_bf16(owned_, to_vec(fe.get(p + ".attn.out.weight")));
    L.b_out = upload_f32(owned_, fe.get(p + ".attn.out.bias").data, d);
    L.ln2_g = upload_f32(owned_, fe.get(p + ".mlp_ln.weight").data, d);
    L.ln2_b = upload_f32(owned_, fe.get(p + ".mlp_ln.bias").data, d);
    L.w_fc1 = upload_bf16(owned_, to_vec(fe.get(p + ".mlp.0.weight")));
    L.b_fc1 = upload_f32(owned_, fe.get(p + ".mlp.0.bias").data, 4 * d);
    L.w_fc2 = upload_bf16(owned_, to_vec(fe.get(p + ".mlp.2.weight")));
    L.b_fc2 = upload_f32(owned_, fe.get(p + ".mlp.2.bias").data, d);
  }
  ln_post_g_ = upload_f32(owned_, fe.get("encoder.ln_post.weight").data, d);
  ln_post_b_ = upload_f32(owned_, fe.get("encoder.ln_post.bias").data, d);
  {  // stacked cross K/V projections: row (layer*2 + kv)*d + out_feature
    std::vector<float> w, b((size_t)cfg_.l_dec * 2 * d, 0.f);
    for (int l = 0; l < cfg_.l_dec; ++l) {
      const std::string p = "decoder.blocks." + std::to_string(l) + ".cross_attn";
      append(w, fe.get(p + ".key.weight"));
      append(w, fe.get(p + ".value.weight"));
      memcpy(b.data() + ((size_t)l * 2 + 1) * d, fe.get(p + ".value.bias").data, d * 4);
    }
    w_crosskv_ = upload_bf16(owned_, w);
    b_crosskv_ = upload_f32(owned_, b.data(), b.size());
  }
  dec_.resize(cfg_.l_dec);
  for (int i = 0; i < cfg_.l_dec; ++i) {
    const std::string p = "decoder.blocks." + std::to_string(i);
    LayerDec& L = dec_[i];
    L.ln1_g = upload_f32(owned_, fd.get(p + ".attn_ln.weight").data, d);
    L.ln1_b = upload_f32(owned_, fd.get(p + ".attn_ln.bias").data, d);
    qkv_pack(fd, p + ".attn", &L.w_qkv, &L.b_qkv);
    L.w_out = upload_bf16(owned_, to_vec(fd.get(p + ".attn.out.weight")));
    L.b_out = upload_f32(owned_, fd.get(p + ".attn.out.bias").data, d);
    L.lnx_g = upload_f32(owned_, fd.get(p + ".cross_attn_ln.weight").data, d);
    L.lnx_b = upload_f32(owned_, fd.get(p + ".cross_attn_ln.bias").data, d);
    L.w_cq = upload_bf16(owned_, to_vec(fd.get(p + ".cross_attn.query.weight")));
    L.b_cq = upload_f32(owned_, fd.get(p + ".cross_attn.query.bias").data, d);
    L.w_co = upload_bf16(owned_, to_vec(fd.get(p + ".cross_attn.out.weight")));
    L.b_co = upload_f32(owned_, fd.get(p + ".cross_attn.out.bias").data, d);
    L.ln2_g = upload_f32(owned_, fd.get(p + ".mlp_ln.weight").data, d);
    L.ln2_b = upload_f32(owned_, fd.get(p + ".mlp_ln.bias").data, d);
    L.w_fc1 = upload_bf16(owned_, to_vec(fd.get(p + ".mlp.0.weight")));
    L.b_fc1 = upload_f32(owned_, fd.get(p + ".mlp.0.bias").data, 4 * d);
    L.w_fc2 = upload_bf16(owned_, to_vec(fd.get(p + ".mlp.2.weight")));
    L.b_fc2 = upload_f32(owned_, fd.get(p + ".mlp.2.bias").data, d);
  }
  dec_ln_g_ = upload_f32(owned_, fd.get("decoder.ln.weight").data, d);
  dec_ln_b_ = upload_f32(owned_, fd.get("decoder.ln.bias").data, d);
  const HostTensor& emb = fd.get("decoder.token_embedding.weight");
  check(emb, {(size_t)cfg_.n_vocab, (size_t)d}, "decoder.token_embedding.weight");
  vocab_pad_ = (cfg_.n_vocab + 255) / 256 * 256;  // zero rows so every W tile is in bounds
  emb_f32_ = upload_f32(owned_, emb.data, emb.numel());
  w_emb_bf16_ = upload_bf16(owned_, to_vec(emb), (size_t)vocab_pad_ * d);
  const HostTensor& pe = fd.get("decoder.positional_embedding");
  check(pe, {(size_t)kTextCtx, (size_t)d}, "decoder.positional_embedding");
  pos_text_ = upload_f32(owned_, pe.data, pe.numel());
}

std::vector<int> Engine::sot_sequence(const std::string& lang, std::string* resolved) const {
  // Whisper::get_lang_token, Whisper.cpp:241-251: unknown language falls back to DEFAULT_LANG "zh"
  std::string l = lang;
  auto it = std::find(cfg_.lang_codes.begin(), cfg_.lang_codes.end(), l);
  if (it == cfg_.lang_codes.end()) {
    l = "zh";
    it = std::find(cfg_.lang_codes.begin(), cfg_.lang_codes.end(), l);
    if (it == cfg_.lang_codes.end()) throw std::runtime_error("config has no language 'zh' to fall back to");
  }
  if (resolved) *resolved = l;
  return {cfg_.sot, cfg_.lang_tokens[it - cfg_.lang_codes.begin()], cfg_.transcribe, cfg_.no_timestamps};
}

// ---------------------------------------------------------------------------------------------------------
// workspace + plans
// ---------------------------------------------------------------------------------------------------------
void Engine::free_workspace() {
  for (auto& kv : graphs_) cudaGraphExecDestroy(kv.second);
  graphs_.clear();
  per_step_launches_.clear();
  for (GemmPlan* p : plans_) gemm_plan_destroy(p);
  plans_.clear();
  enc_plans_.clear();
  dec_plans_.clear();
  for (void* p : ws_owned_) cudaFree(p);
  ws_owned_.clear();
  // the engine is EMPTY from here until ensure_capacity() has rebuilt everything: a failed re-allocation must not leave
  // capacities that describe freed memory
  cap_ = 0, pcm_stride_ = 0, enc_sub_ = 0, dec_rows_pad_ = 0;
  pcm_ = mel_ = x_enc_ = x_dec_ = qkv_dec_ = q_dec_ = logits_ = part_val_ = nullptr;
  n_samples_ = part_idx_ = cross_work_ = step_ctr_ = slot_seq_ = boundary_ticket_ = utt_state_ = nullptr;
  mel_tm_ = conv1_out_ = h_enc_ = qkv_enc_ = attn_enc_ = mlp_enc_ = cross_k_ = cross_v_ = self_k_ = self_v_ = nullptr;
  h_dec_ = attn_dec_ = mlp_dec_ = nullptr;
  st_ = DecodeState{};
}

void Engine::ensure_capacity(int B, long max_samples) {
  const long stride = std::max<long>(kChunkSamples, (max_samples + 7) / 8 * 8);
  if (B <= cap_ && stride <= pcm_stride_) return;
  if (B <= 0) throw std::runtime_error("ensure_capacity: batch must be positive");
  CUDA_CHECK(cudaSetDevice(device_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  const int new_cap = std::max(B, cap_);
  const long new_stride = std::max(stride, pcm_stride_);
  free_workspace();  // resets cap_ / pcm_stride_ / pointers: the engine is empty until the allocations below succeed
  try {
    allocate_workspace(new_cap, new_stride);
  } catch (...) {
    free_workspace();  // back to the empty state (capacity 0): the next, smaller request re-allocates cleanly
    throw;
  }
}

void Engine::allocate_workspace(int new_cap, long new_stride) {
  cap_ = new_cap;
  pcm_stride_ = new_stride;
  const int d = cfg_.d, H = cfg_.n_head, L = cfg_.l_dec;
  const char* sub_env = getenv("B200W_ENC_SUB_BATCH");
  enc_sub_ = std::min(cap_, sub_env ? std::max(1, atoi(sub_env)) : 128);  // measured: larger sub-batches are slightly faster
  const size_t rows_sub = (size_t)enc_sub_ * kAudioCtx;
  auto& o = ws_owned_;
  pcm_ = dev_alloc<float>(o, (size_t)cap_ * pcm_stride_, false);
  n_samples_ = dev_alloc<int>(o, cap_);
  utt_state_ = dev_alloc<int>(o, 2 * (size_t)cap_ + 1);
  mel_ = dev_alloc<float>(o, (size_t)cap_ * cfg_.n_mels * kMelFrames);
  mel_tm_ = dev_alloc<__nv_bfloat16>(o, (size_t)cap_ * (kMelFrames + 2) * cfg_.n_mels);
  conv1_out_ = dev_alloc<__nv_bfloat16>(o, (size_t)enc_sub_ * (kMelFrames + 2) * d);
  x_enc_ = dev_alloc<float>(o, rows_sub * d);
  h_enc_ = dev_alloc<__nv_bfloat16>(o, rows_sub * d);
  qkv_enc_ = dev_alloc<__nv_bfloat16>(o, rows_sub * 3 * d);
  attn_enc_ = dev_alloc<__nv_bfloat16>(o, rows_sub * d);
  mlp_enc_ = dev_alloc<__nv_bfloat16>(o, rows_sub * 4 * d);
  const size_t ckv = (size_t)L * cap_ * H * kAudioCtx * 64;
  cross_k_ = dev_alloc<__nv_bfloat16>(o, ckv, false);
  cross_v_ = dev_alloc<__nv_bfloat16>(o, ckv, false);
  const size_t skv = (size_t)L * cap_ * H * kTextCtx * 64;
  self_k_ = dev_alloc<__nv_bfloat16>(o, skv);
  self_v_ = dev_alloc<__nv_bfloat16>(o, skv);
  dec_rows_pad_ = (cap_ + 127) / 128 * 128;
  x_dec_ = dev_alloc<float>(o, (size_t)dec_rows_pad_ * d);
  qkv_dec_ = dev_alloc<float>(o, (size_t)dec_rows_pad_ * 3 * d);
  q_dec_ = dev_alloc<float>(o, (size_t)dec_rows_pad_ * d);
  h_dec_ = dev_alloc<__nv_bfloat16>(o, (size_t)dec_rows_pad_ * d);
  attn_dec_ = dev_alloc<__nv_bfloat16>(o, (size_t)dec_rows_pad_ * d);
  mlp_dec_ = dev_alloc<__nv_bfloat16>(o, (size_t)dec_rows_pad_ * 4 * d);
  logits_ = dev_alloc<float>(o, (size_t)cap_ * vocab_pad_);
  logits_tiles_ = vocab_pad_ / 128 * 2;  // arg-max partials: two column halves per 128-wide logits tile (gemm epilogue)
  part_val_ = dev_alloc<float>(o, (size_t)cap_ * logits_tiles_);
  part_idx_ = dev_alloc<int>(o, (size_t)cap_ * logits_tiles_);
  cross_work_ = dev_alloc<int>(o, (size_t)cfg_.l_dec * 4 * 2 + 2);  // item counters of the streaming cross-attention launches
  step_ctr_ = dev_alloc<int>(o, 4);  // one step counter per micro-batch (they advance independently inside a graph)
  boundary_ticket_ = dev_alloc<int>(o, 4);  // arrival counters of the step-boundary kernel, one per micro-batch
  st_.step = step_ctr_;
  st_.tokens = dev_alloc<int>(o, (size_t)cap_ * kTextCtx);
  st_.forced = dev_alloc<int>(o, (size_t)cap_ * kTextCtx);
  st_.finished = dev_alloc<int>(o, cap_);
  st_.out_tokens = dev_alloc<int>(o, (size_t)cap_ * kTextCtx);
  slot_seq_ = dev_alloc<int>(o, cap_);
  st_.slot_seq = slot_seq_;
  slot_map_identity_ = false;
  reset_slot_map();
  build_plans();
}

void Engine::build_plans() {
  const int d = cfg_.d, n_mels = cfg_.n_mels;
  const bool two = getenv("B200W_GEMM_1CTA") == nullptr;  // encoder GEMMs: CTA-pair kernel (bring-up switch to the 1-CTA kernel)
  const bool tma_out = two && getenv("B200W_GEMM_DIRECT_STORE") == nullptr;  // epilogues through TMA stores / reduce-adds
  auto keep = [&](GemmPlan* p) {
    plans_.push_back(p);
    return p;
  };
  auto flat = [&](const __nv_bfloat16* ptr, int K, int rows) {
    GemmOperandA a{};
    a.ptr = ptr, a.K = K, a.rows = rows, a.n_batch = 1, a.row_pitch = K, a.batch_pitch = (long)rows * K, a.n_taps = 0;
    return a;
  };
  {  // conv1: x_pad [cap][3002][n_mels], taps = rows +0,+1,+2 (padded row index = t + tap)
    GemmOperandA a{};
    a.ptr = mel_tm_, a.K = n_mels, a.rows = kMelFrames + 2, a.n_batch = cap_, a.row_pitch = n_mels;
    a.batch_pitch = (long)(kMelFrames + 2) * n_mels;
    a.n_taps = 3, a.k_per_tap = n_mels;
    for (int t = 0; t < 3; ++t) a.tap_c0[t] = 0, a.tap_row[t] = t;
    p_conv1_ = keep(gemm_plan_create(a, w_conv1_, d, 128, EPI_BIAS_GELU_BF16, two));
  }
  {  // conv2 (stride 2): conv1_out viewed as [sub][1501][2d]; out t reads padded rows 2t, 2t+1, 2t+2
    GemmOperandA a{};
    a.ptr = conv1_out_, a.K = 2 * d, a.rows = (kMelFrames + 2) / 2, a.n_batch = enc_sub_, a.row_pitch = 2L * d;
    a.batch_pitch = (long)(kMelFrames + 2) * d;
    a.n_taps = 3, a.k_per_tap = d;
    a.tap_c0[0] = 0, a.tap_row[0] = 0;
    a.tap_c0[1] = d, a.tap_row[1] = 0;
    a.tap_c0[2] = 0, a.tap_row[2] = 1;
    p_conv2_ = keep(gemm_plan_create(a, w_conv2_, d, 128, EPI_GELU_POS_F32, two));
  }
  const int rows_sub = enc_sub_ * kAudioCtx;
  enc_plans_.resize(cfg_.l_enc);
  for (int i = 0; i < cfg_.l_enc; ++i) {
    const LayerEnc& L = enc_[i];
    const GemmTmaOut o_qkv{qkv_enc_, nullptr, rows_sub, 3L * d, 3L * d, 0}, o_x{x_enc_, nullptr, rows_sub, d, d, 0};
    const GemmTmaOut o_mlp{mlp_enc_, nullptr, rows_sub, 4L * d, 4L * d, 0};
    enc_plans_[i].qkv = keep(gemm_plan_create(flat(h_enc_, d, rows_sub), L.w_qkv, 3 * d, 256, EPI_BIAS_BF16, two, tma_out ? &o_qkv : nullptr));
    enc_plans_[i].out = keep(gemm_plan_create(flat(attn_enc_, d, rows_sub), L.w_out, d, 128, EPI_BIAS_RESID_F32, two, tma_out ? &o_x : nullptr));
    enc_plans_[i].fc1 = keep(gemm_plan_create(flat(h_enc_, d, rows_sub), L.w_fc1, 4 * d, 256, EPI_BIAS_GELU_BF16, two, tma_out ? &o_mlp : nullptr));
    enc_plans_[i].fc2 = keep(gemm_plan_create(flat(mlp_enc_, 4 * d, rows_sub), L.w_fc2, d, 128, EPI_BIAS_RESID_F32, two, tma_out ? &o_x : nullptr));
  }
  {  // cross K/V: rows are (chunk, t) so the epilogue can scatter head-major per chunk
    GemmOperandA a{};
    a.ptr = h_enc_, a.K = d, a.rows = kAudioCtx, a.n_batch = enc_sub_, a.row_pitch = d, a.batch_pitch = (long)kAudioCtx * d, a.n_taps = 0;
    const GemmTmaOut o_kv{cross_k_, cross_v_, kAudioCtx, 64, 64, (long)cfg_.l_dec * cap_ * cfg_.n_head};
    p_crosskv_ = keep(gemm_plan_create(a, w_crosskv_, 2 * cfg_.l_dec * d, 256, EPI_CROSSKV_BF16, two, tma_out ? &o_kv : nullptr));
  }
  dec_plans_.resize(cfg_.l_dec);
  const char* bn_env = getenv("B200W_DEC_BN");
  const int bn = bn_env ? atoi(bn_env) : 32;  // N tile of the decoder-step GEMMs: narrow tiles = more CTAs streaming W (measured best)
  for (int i = 0; i < cfg_.l_dec; ++i) {
    const LayerDec& L = dec_[i];
    dec_plans_[i].qkv = keep(gemm_plan_create(flat(h_dec_, d, dec_rows_pad_), L.w_qkv, 3 * d, bn, EPI_BIAS_F32));
    dec_plans_[i].out = keep(gemm_plan_create(flat(attn_dec_, d, dec_rows_pad_), L.w_out, d, bn, EPI_BIAS_RESID_F32));
    dec_plans_[i].cq = keep(gemm_plan_create(flat(h_dec_, d, dec_rows_pad_), L.w_cq, d, bn, EPI_BIAS_F32));
    dec_plans_[i].co = keep(gemm_plan_create(flat(attn_dec_, d, dec_rows_pad_), L.w_co, d, bn, EPI_BIAS_RESID_F32));
    dec_plans_[i].fc1 = keep(gemm_plan_create(flat(h_dec_, d, dec_rows_pad_), L.w_fc1, 4 * d, bn, EPI_BIAS_GELU_BF16));
    dec_plans_[i].fc2 = keep(gemm_plan_create(flat(mlp_dec_, 4 * d, dec_rows_pad_), L.w_fc2, d, bn, EPI_BIAS_RESID_F32));
  }
  p_logits_ = keep(gemm_plan_create(flat(h_dec_, d, dec_rows_pad_), w_emb_bf16_, vocab_pad_, 128, EPI_ARGMAX));
}

// ---------------------------------------------------------------------------------------------------------
// stages
// ---------------------------------------------------------------------------------------------------------
void Engine::run_logmel(int B, int max_samples) { run_logmel_range(0, B, max_samples); }
void Engine::run_logmel_range(int b0, int nb, int max_samples) {
  launch_logmel(pcm_ + (size_t)b0 * pcm_stride_, pcm_stride_, n_samples_ + b0, max_samples, nb, cfg_.n_mels,
                mel_ + (size_t)b0 * cfg_.n_mels * kMelFrames, mel_tm_ + (size_t)b0 * (kMelFrames + 2) * cfg_.n_mels, utt_state_ + 2 * (size_t)b0, stream_);
  launches_ += 1;
}
void Engine::run_mel_convert(int B) {
  launch_mel_to_timemajor(mel_, B, cfg_.n_mels, mel_tm_, stream_);
  launches_ += 1;
}

void Engine::run_encoder(int B) {
  for (int b0 = 0; b0 < B; b0 += enc_sub_) run_encoder_range(b0, std::min(enc_sub_, B - b0));
}

void Engine::run_encoder_range(int b0, int nb) {
  const int d = cfg_.d, H = cfg_.n_head;
  if (nb > enc_sub_) throw std::runtime_error("encoder sub-batch exceeds its workspace");
  {
    const int rows = nb * kAudioCtx;
    GemmParams p{};
    // conv1 + GELU -> padded bf16 [nb][3002][d] (row t+1)
    p = GemmParams{};
    p.rows_valid = kMelFrames, p.N = d, p.out = conv1_out_, p.ldo = d, p.out_batch_pitch = (long)(kMelFrames + 2) * d;
    p.out_row_offset = 1, p.a_batch_offset = b0, p.n_batch = nb, p.bias = b_conv1_;
    gemm_launch(p_conv1_, p, stream_);
    // conv2 (stride 2) + GELU + positional embedding -> residual stream f32 [nb*1500][d]
    p = GemmParams{};
    p.rows_valid = kAudioCtx, p.N = d, p.out = x_enc_, p.ldo = d, p.out_batch_pitch = (long)kAudioCtx * d;
    p.n_batch = nb, p.bias = b_conv2_, p.pos = pos_audio_;
    gemm_launch(p_conv2_, p, stream_);
    launches_ += 2;
    for (int i = 0; i < cfg_.l_enc; ++i) {
      const LayerEnc& L = enc_[i];
      launch_layernorm(x_enc_, L.ln1_g, L.ln1_b, h_enc_, rows, d, stream_);
      p = GemmParams{};
      p.rows_valid = rows, p.N = 3 * d, p.out = qkv_enc_, p.ldo = 3 * d, p.bias = L.b_qkv, p.n_batch = 1;
      gemm_launch(enc_plans_[i].qkv, p, stream_);
      if (attn_mma_sync_)
        launch_encoder_attention(qkv_enc_, attn_enc_, nb, kAudioCtx, H, stream_);  // bring-up comparator (B200W_ATTN=mma_sync)
      else
        launch_encoder_attention_tcgen05(qkv_enc_, attn_enc_, nb, kAudioCtx, H, stream_);
      p = GemmParams{};
      p.rows_valid = rows, p.N = d, p.out = x_enc_, p.ldo = d, p.bias = L.b_out, p.n_batch = 1;
      gemm_launch(enc_plans_[i].out, p, stream_);
      launch_layernorm(x_enc_, L.ln2_g, L.ln2_b, h_enc_, rows, d, stream_);
      p = GemmParams{};
      p.rows_valid = rows, p.N = 4 * d, p.out = mlp_enc_, p.ldo = 4 * d, p.bias = L.b_fc1, p.n_batch = 1;
      gemm_launch(enc_plans_[i].fc1, p, stream_);
      p = GemmParams{};
      p.rows_valid = rows, p.N = d, p.out = x_enc_, p.ldo = d, p.bias = L.b_fc2, p.n_batch = 1;
      gemm_launch(enc_plans_[i].fc2, p, stream_);
      launches_ += 7;
    }
    launch_layernorm(x_enc_, ln_post_g_, ln_post_b_, h_enc_, rows, d, stream_);
    p = GemmParams{};
    p.rows_valid = kAudioCtx, p.N = 2 * cfg_.l_dec * d, p.n_batch = nb, p.bias = b_crosskv_;
    p.cross_k = cross_k_, p.cross_v = cross_v_, p.d_model = d, p.n_head = H, p.n_ctx_kv = kAudioCtx, p.kv_batch = cap_, p.kv_batch_offset = b0;
    gemm_launch(p_crosskv_, p, stream_);
    launches_ += 2;
  }
}

// n_fused decoder steps for B sequences.  With B >= 32 the batch is split into two micro-batches on two streams: while one
// micro-batch streams its cross-attention K/V (HBM-bound; a single resident wave of 3 CTAs per SM that leaves registers and
// shared memory for others), the other runs its chain of short latency-bound kernels (LayerNorm, M <= 128 GEMMs, self
// attention) on the same SMs.  Events hand the HBM "token" back and forth so the cross-attention kernels alternate instead
// of competing; sequences are independent, so results do not change.
void Engine::enqueue_decode_step(int B, bool want_logits, bool finalize, int honor_eot, int n_fused, bool need_embed) {
  const int d = cfg_.d, H = cfg_.n_head, Ld = cfg_.l_dec;
  int n_mb = (micro_batch_ && B >= 32) ? n_micro_batch_ : 1;
  while (n_mb > 1 && B / n_mb < 16) --n_mb;
  struct MB {
    int b0, nb;
    cudaStream_t s;
  } mb[4];
  cudaStream_t mb_stream[4] = {stream_, stream2_, mb_streams_[0], mb_streams_[1]};
  for (int i = 0, b0 = 0; i < n_mb; ++i) {
    const int nb = (B - b0 + (n_mb - i) - 1) / (n_mb - i);
    mb[i] = {b0, nb, mb_stream[i]};
    b0 += nb;
  }
  // events: [0] fork, [1..3] joins, [8 + 4 l + i] end of micro-batch i's cross attention of layer l
  cudaEvent_t ev_fork = step_events_[0];
  auto cross_event = [&](int l, int i) { return step_events_[8 + 4 * l + i]; };
  // With several micro-batches the short kernels are launched urgent and the long cross-attention kernels at default priority:
  // the block scheduler then slots another micro-batch's dependent chain in between the cross-attention CTAs.
  static const bool use_prio = getenv("B200W_NO_PRIORITY") == nullptr;
  const int prio_small = (n_mb > 1 && use_prio) ? prio_high_ : 0;
  ScopedLaunchPriority prio_scope(prio_small);
  if (n_mb > 1) {
    CUDA_CHECK(cudaEventRecord(ev_fork, stream_));
    for (int i = 1; i < n_mb; ++i) CUDA_CHECK(cudaStreamWaitEvent(mb[i].s, ev_fork, 0));
  }
  // hand-over events between the cross-attention kernels only pay off when a launch is long against the chain it has to
  // cover (measured: small B=256, 590 MB per launch: +2 % with; base B=64, 98 MB: -5 % with; turbo B=128 / base B=256: neutral);
  // decided once per step from the largest micro-batch so that every micro-batch takes part or none does
  const bool chain = n_mb > 1 && cross_chain_ && (cross_chain_forced_ || (size_t)mb[0].nb * kAudioCtx * d * 4 >= ((size_t)512 << 20));
  auto gp = [&](const MB& m, void* out, size_t elem, long ldo, int N, const float* bias) {
    GemmParams q{};
    q.rows_valid = m.nb, q.N = N, q.out = static_cast<char*>(out) + (size_t)m.b0 * ldo * elem, q.ldo = ldo, q.bias = bias, q.n_batch = 1;
    q.use_pdl = 1, q.a_row_offset = m.b0;
    return q;
  };
  // n_fused consecutive steps in one enqueue (= one CUDA graph): every micro-batch has its own step counter and moves on to
  // its next step without waiting for the others, so one micro-batch's logits / arg-max / embedding run underneath the other's
  // cross attention instead of leaving HBM idle at every step boundary
  static const bool fuse_boundary = getenv("B200W_NO_STEP_BOUNDARY") == nullptr;
  for (int fs = 0; fs < n_fused; ++fs) {
    // x = embedding of the token at this position, h = LayerNorm 1 of the first block: produced by the previous step's boundary
    // kernel, except for the first step of a decode (or after the slot list changed), where the caller asks for them here
    const bool boundary_made_it = fuse_boundary && finalize && (fs > 0 || !need_embed);
    if (!boundary_made_it) {
      for (int i = 0; i < n_mb; ++i) {
        DecodeState st = st_;
        st.step = step_ctr_ + i;
        st.slot_seq = slot_seq_ + mb[i].b0;  // token tables and caches are per sequence, rows per slot
        launch_embed(st, emb_f32_, pos_text_, x_dec_ + (size_t)mb[i].b0 * d, mb[i].nb, d, kTextCtx, mb[i].s, /*pdl=*/n_mb == 1 || fs > 0);
        launch_layernorm(x_dec_ + (size_t)mb[i].b0 * d, dec_[0].ln1_g, dec_[0].ln1_b, h_dec_ + (size_t)mb[i].b0 * d, mb[i].nb, d, mb[i].s);
      }
      launches_ += 2 * n_mb;
    }
    for (int l = 0; l < Ld; ++l) {
      const LayerDec& L = dec_[l];
      const DecPlans& P = dec_plans_[l];
      for (int i = 0; i < n_mb; ++i) {
        const MB& m = mb[i];
        const size_t skv_off = (size_t)l * cap_ * H * kTextCtx * 64;  // layer base; the kernel adds the sequence of each slot
        float* x = x_dec_ + (size_t)m.b0 * d;
        __nv_bfloat16* h = h_dec_ + (size_t)m.b0 * d;
        if (l > 0) launch_layernorm(x, L.ln1_g, L.ln1_b, h, m.nb, d, m.s);  // l == 0: made by the embed / step-boundary kernel
        gemm_launch(P.qkv, gp(m, qkv_dec_, 4, 3 * d, 3 * d, L.b_qkv), m.s);
        launch_self_attention_decode(qkv_dec_ + (size_t)m.b0 * 3 * d, self_k_ + skv_off, self_v_ + skv_off, step_ctr_ + i, slot_seq_ + m.b0,
                                     attn_dec_ + (size_t)m.b0 * d, m.nb, H, kTextCtx, m.s);
        gemm_launch(P.out, gp(m, x_dec_, 4, d, d, L.b_out), m.s);
        launch_layernorm(x, L.lnx_g, L.lnx_b, h, m.nb, d, m.s);
        gemm_launch(P.cq, gp(m, q_dec_, 4, d, d, L.b_cq), m.s);
      }
      for (int i = 0; i < n_mb; ++i) {
        const MB& m = mb[i];
        const int n_split = cross_attention_pick_split(m.nb, H, kAudioCtx);
        const size_t ckv_off = (size_t)l * cap_ * H * kAudioCtx * 64;
        if (chain) {
          // the cross-attention kernels take turns on HBM: micro-batch i waits for micro-batch i-1's kernel of this layer,
          // micro-batch 0 for the last micro-batch's kernel of the previous layer
          if (i == 0 && (l > 0 || fs > 0)) CUDA_CHECK(cudaStreamWaitEvent(m.s, cross_event(l > 0 ? l - 1 : Ld - 1, n_mb - 1), 0));
          if (i > 0) CUDA_CHECK(cudaStreamWaitEvent(m.s, cross_event(l, i - 1), 0));
        }
        {
          ScopedLaunchPriority low(0);
          launch_cross_attention_decode(q_dec_ + (size_t)m.b0 * d, cross_k_ + ckv_off, cross_v_ + ckv_off, slot_seq_ + m.b0,
                                        attn_dec_ + (size_t)m.b0 * d, m.nb, H, kAudioCtx, n_split, m.s, /*pdl=*/!chain,
                                        cross_work_ + (l * 4 + i) * 2);
        }
        if (chain) CUDA_CHECK(cudaEventRecord(cross_event(l, i), m.s));
        launches_ += 1;
      }
      for (int i = 0; i < n_mb; ++i) {
        const MB& m = mb[i];
        float* x = x_dec_ + (size_t)m.b0 * d;
        __nv_bfloat16* h = h_dec_ + (size_t)m.b0 * d;
        gemm_launch(P.co, gp(m, x_dec_, 4, d, d, L.b_co), m.s);
        launch_layernorm(x, L.ln2_g, L.ln2_b, h, m.nb, d, m.s);
        gemm_launch(P.fc1, gp(m, mlp_dec_, 2, 4 * d, 4 * d, L.b_fc1), m.s);
        gemm_launch(P.fc2, gp(m, x_dec_, 4, d, d, L.b_fc2), m.s);
      }
      launches_ += (l == 0 ? 9 : 10) * n_mb;
    }
    for (int i = 0; i < n_mb; ++i) {
      const MB& m = mb[i];
      launch_layernorm(x_dec_ + (size_t)m.b0 * d, dec_ln_g_, dec_ln_b_, h_dec_ + (size_t)m.b0 * d, m.nb, d, m.s);
      GemmParams p = gp(m, want_logits ? logits_ : nullptr, 4, vocab_pad_, cfg_.n_vocab, nullptr);
      if (!want_logits) p.out = nullptr;
      p.part_val = part_val_ + (size_t)m.b0 * logits_tiles_, p.part_idx = part_idx_ + (size_t)m.b0 * logits_tiles_, p.part_ld = logits_tiles_;
      gemm_launch(p_logits_, p, m.s);
      launches_ += 2;
      if (finalize) {
        DecodeState st = st_;
        st.step = step_ctr_ + i;
        st.slot_seq = slot_seq_ + m.b0;
        if (fuse_boundary) {
          launch_step_boundary(st, p.part_val, p.part_idx, logits_tiles_, logits_tiles_, m.nb, kTextCtx, cfg_.eot, honor_eot, kSotLen, emb_f32_,
                               pos_text_, dec_[0].ln1_g, dec_[0].ln1_b, x_dec_ + (size_t)m.b0 * d, h_dec_ + (size_t)m.b0 * d, d,
                               boundary_ticket_ + i, m.s);
          launches_ += 1;
        } else {
          launch_argmax_finalize(st, p.part_val, p.part_idx, logits_tiles_, logits_tiles_, m.nb, kTextCtx, cfg_.eot, honor_eot, kSotLen, m.s);
          launch_advance_step(step_ctr_ + i, m.s, /*pdl=*/true);  // this micro-batch's own step counter
          launches_ += 2;
        }
      }
    }
  }  // fused steps
  for (int i = 1; i < n_mb; ++i) {
    CUDA_CHECK(cudaEventRecord(step_events_[i], mb[i].s));
    CUDA_CHECK(cudaStreamWaitEvent(stream_, step_events_[i], 0));
  }
}

// embedding of the current position + LayerNorm 1 of the first block for slots [0, B) (start of a decode, or after an EOT
// compaction; every other step gets them from the previous step's boundary kernel)
void Engine::enqueue_embed_ln(int B) {
  DecodeState st = st_;  // all micro-batch step counters agree between steps: counter 0 serves the whole slot list
  launch_embed(st, emb_f32_, pos_text_, x_dec_, B, cfg_.d, kTextCtx, stream_, /*pdl=*/false);
  launch_layernorm(x_dec_, dec_[0].ln1_g, dec_[0].ln1_b, h_dec_, B, cfg_.d, stream_);
  launches_ += 2;
}

void Engine::run_cross_attention_only(int B) {
  check_batch(B, "run_cross_attention_only");
  const int H = cfg_.n_head;
  const int n_split = cross_attention_pick_split(B, H, kAudioCtx);
  for (int l = 0; l < cfg_.l_dec; ++l) {
    const size_t ckv_off = (size_t)l * cap_ * H * kAudioCtx * 64;
    launch_cross_attention_decode(q_dec_, cross_k_ + ckv_off, cross_v_ + ckv_off, slot_seq_, attn_dec_, B, H, kAudioCtx, n_split, stream_, true,
                                  cross_work_ + (size_t)cfg_.l_dec * 4 * 2);
    launches_ += 1;
  }
}

void Engine::check_batch(int B, const char* what) const {
  if (B <= 0 || B > cap_)
    throw std::runtime_error(std::string(what) + ": batch " + std::to_string(B) + " exceeds the resident capacity " + std::to_string(cap_) +
                             " (run the encoder / upload PCM for this batch first)");
}

void Engine::decode_reset(int B) {
  check_batch(B, "decode_reset");
  CUDA_CHECK(cudaSetDevice(device_));
  CUDA_CHECK(cudaMemsetAsync(step_ctr_, 0, 4 * sizeof(int), stream_));
  CUDA_CHECK(cudaMemsetAsync(st_.finished, 0, sizeof(int) * B, stream_));
  CUDA_CHECK(cudaMemsetAsync(st_.forced, 0xff, sizeof(int) * (size_t)B * kTextCtx, stream_));  // -1
  reset_slot_map();
}

// slot i works on sequence i: the state outside run_decode's EOT compaction
void Engine::reset_slot_map() {
  if (slot_map_identity_) return;
  std::vector<int> id(cap_);
  for (int i = 0; i < cap_; ++i) id[i] = i;
  CUDA_CHECK(cudaMemcpyAsync(slot_seq_, id.data(), sizeof(int) * cap_, cudaMemcpyHostToDevice, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));  // `id` is a stack-owned host vector
  slot_map_identity_ = true;
}

int Engine::run_decode(int B, const std::vector<int>& sot, const DecodeOptions& opt, std::vector<std::vector<int>>* tokens) {
  if ((int)sot.size() != kSotLen) throw std::runtime_error("sot sequence must have 4 tokens");
  check_batch(B, "run_decode");
  for (int r = 0; r < opt.n_logit_rows; ++r)
    if (opt.logit_rows[r] < 0 || opt.logit_rows[r] >= B) throw std::runtime_error("run_decode: logit row outside the batch");
  CUDA_CHECK(cudaSetDevice(device_));
  decode_reset(B);
  // prefill token / forcing tables
  std::vector<int> tok((size_t)B * kTextCtx, 0), forced((size_t)B * kTextCtx, -1);
  for (int b = 0; b < B; ++b) {
    for (int i = 0; i < kSotLen; ++i) tok[(size_t)b * kTextCtx + i] = sot[i];
    for (int i = 0; i < opt.forced_len && kSotLen + i < kTextCtx; ++i)
      forced[(size_t)b * kTextCtx + kSotLen + i] = opt.forced_tokens[(size_t)b * opt.forced_len + i];
  }
  CUDA_CHECK(cudaMemcpyAsync(st_.tokens, tok.data(), tok.size() * sizeof(int), cudaMemcpyHostToDevice, stream_));
  if (opt.forced_tokens)
    CUDA_CHECK(cudaMemcpyAsync(st_.forced, forced.data(), forced.size() * sizeof(int), cudaMemcpyHostToDevice, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));  // tok/forced are stack-owned host vectors

  // number of decoder runs: 4 SOT steps + one per generated token (each generated token is fed back, Whisper.cpp:219-222)
  const int max_new = std::max(1, std::min(opt.max_new_tokens, kTextCtx - kSotLen));
  const int n_steps = kSotLen + max_new;  // like the reference, the last token is fed back too and its result discarded
  const bool want_logits = opt.logits_out != nullptr;
  const bool graph = opt.use_graph && !want_logits && getenv("B200W_NO_GRAPH") == nullptr;
  // graphs of `k` consecutive decoder steps (k = graph_steps_ and, for the remainder, 1), cached per (B, honor_eot, k)
  auto get_graph = [&](int B, int k) {  // B = number of active slots
    const int key = (B * 2 + (opt.honor_eot ? 1 : 0)) * 64 + k;
    auto it = graphs_.find(key);
    if (it != graphs_.end()) return std::make_pair(it->second, per_step_launches_[key]);
    cudaGraph_t g;
    cudaGraphExec_t exec = nullptr;
    CUDA_CHECK(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
    const long before = launches_;
    enqueue_decode_step(B, false, true, opt.honor_eot ? 1 : 0, k, /*need_embed=*/false);
    per_step_launches_[key] = launches_ - before;
    launches_ = before;  // capture does not launch
    CUDA_CHECK(cudaStreamEndCapture(stream_, &g));
    if (getenv("B200W_DIAG_GRAPH")) {  // diagnostic: launch priorities the captured kernel nodes carry
      size_t n_nodes = 0;
      CUDA_CHECK(cudaGraphGetNodes(g, nullptr, &n_nodes));
      std::vector<cudaGraphNode_t> nodes(n_nodes);
      CUDA_CHECK(cudaGraphGetNodes(g, nodes.data(), &n_nodes));
      std::map<int, int> hist;
      for (cudaGraphNode_t nd : nodes) {
        cudaGraphNodeType ty;
        CUDA_CHECK(cudaGraphNodeGetType(nd, &ty));
        if (ty != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeAttrValue v{};
        CUDA_CHECK(cudaGraphKernelNodeGetAttribute(nd, cudaKernelNodeAttributePriority, &v));
        hist[v.priority]++;
      }
      for (auto& kv : hist) fprintf(stderr, "[b200w] decode graph (%d steps): %d kernel nodes with priority %d\n", k, kv.second, kv.first);
    }
    CUDA_CHECK(cudaGraphInstantiate(&exec, g, 0));
    CUDA_CHECK(cudaGraphDestroy(g));
    graphs_[key] = exec;
    return std::make_pair(exec, per_step_launches_[key]);
  };
  int steps_done = 0;
  std::vector<int> active(B);  // sequences still decoding, in slot order
  for (int b = 0; b < B; ++b) active[b] = b;
  static const bool compact = getenv("B200W_NO_COMPACT") == nullptr;
  bool need_embed = true;  // first step, and the first one after the slot list changed: nobody has produced x / h for it yet
  while (steps_done < n_steps) {
    int k = 1;
    const int Ba = (int)active.size();
    if (graph) {
      k = (n_steps - steps_done >= graph_steps_) ? graph_steps_ : 1;
      if (need_embed) enqueue_embed_ln(Ba);  // the graphs start from a ready x / h (their boundary kernels keep it so)
      need_embed = false;
      const auto ge = get_graph(Ba, k);
      CUDA_CHECK(cudaGraphLaunch(ge.first, stream_));
      launches_ += ge.second;
    } else {
      const int s = steps_done;
      enqueue_decode_step(Ba, want_logits, true, opt.honor_eot ? 1 : 0, 1, need_embed);
      need_embed = false;
      if (want_logits && s >= kSotLen - 1 && s - (kSotLen - 1) < max_new) {
        // logits after consuming position s = prediction of generated token (s - 3)
        const size_t step_i = (size_t)(s - (kSotLen - 1));
        if (opt.logit_rows != nullptr) {
          for (int r = 0; r < opt.n_logit_rows; ++r)
            CUDA_CHECK(cudaMemcpyAsync(opt.logits_out + (step_i * opt.n_logit_rows + r) * cfg_.n_vocab, logits_ + (size_t)opt.logit_rows[r] * vocab_pad_,
                                       (size_t)cfg_.n_vocab * 4, cudaMemcpyDeviceToHost, stream_));
        } else {
          CUDA_CHECK(cudaMemcpy2DAsync(opt.logits_out + step_i * B * cfg_.n_vocab, (size_t)cfg_.n_vocab * 4, logits_, (size_t)vocab_pad_ * 4,
                                       (size_t)cfg_.n_vocab * 4, B, cudaMemcpyDeviceToHost, stream_));
        }
      }
    }
    steps_done += k;
    if (opt.honor_eot && (steps_done % 16 == 0) && steps_done < n_steps && B <= kMaxPolled) {
      // EOT bookkeeping every 16 steps: stop when every sequence has finished; otherwise drop the finished ones from the slot
      // list so that the remaining steps cost what the still-running sequences cost (nothing moves in HBM: caches and token
      // tables are per sequence, kernels reach them through slot_seq).  Results do not change: a sequence's arithmetic is
      // independent of the batch it runs in (decode_ops.cu).
      CUDA_CHECK(cudaMemcpyAsync(pinned_flags_, st_.finished, sizeof(int) * B, cudaMemcpyDeviceToHost, stream_));
      CUDA_CHECK(cudaStreamSynchronize(stream_));
      std::vector<int> still;
      still.reserve(active.size());
      for (int s : active)
        if (pinned_flags_[s] == 0) still.push_back(s);
      if (still.empty()) break;
      if (compact && !want_logits && (int)still.size() <= (int)active.size() - std::max(1, (int)active.size() / 8)) {
        for (size_t i = 0; i < still.size(); ++i) pinned_map_[i] = still[i];
        CUDA_CHECK(cudaMemcpyAsync(slot_seq_, pinned_map_, sizeof(int) * still.size(), cudaMemcpyHostToDevice, stream_));
        slot_map_identity_ = false;
        active.swap(still);
        need_embed = true;  // rows are per slot: the boundary kernel's x / h belong to the old slot list
        ++compactions_;
        if (graphs_.size() > 48) {  // many distinct batch sizes over time: start the graph cache over
          CUDA_CHECK(cudaStreamSynchronize(stream_));
          for (auto& kv : graphs_) cudaGraphExecDestroy(kv.second);
          graphs_.clear();
          per_step_launches_.clear();
        }
      }
    }
  }
  last_active_ = (int)active.size();
  reset_slot_map();  // later stateful calls (decoder_loop, a new decode) see slot i == sequence i again
  if (tokens) {
    std::vector<int> out((size_t)B * kTextCtx);
    CUDA_CHECK(cudaMemcpyAsync(out.data(), st_.out_tokens, out.size() * sizeof(int), cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    tokens->assign(B, {});
    for (int b = 0; b < B; ++b) {
      // generated token i is the argmax after consuming position 3 + i (Whisper.cpp:214-222)
      for (int i = 0; i < max_new && kSotLen - 1 + i < steps_done; ++i) {
        const int t = out[(size_t)b * kTextCtx + kSotLen - 1 + i];
        if (opt.honor_eot && t == cfg_.eot) break;
        (*tokens)[b].push_back(t);
      }
    }
  }
  return steps_done;
}

void Engine::decode_step_tokens(int B, const int* tokens_host, int offset, float* logits_host, float* this_k, float* this_v) {
  // one decoder run on the resident caches, driven like the reference's run_decoder(token, offset)
  check_batch(B, "decode_step_tokens");
  CUDA_CHECK(cudaSetDevice(device_));
  const int d = cfg_.d, H = cfg_.n_head, L = cfg_.l_dec;
  if (offset < 0 || offset >= kTextCtx) throw std::runtime_error("decoder offset out of range");
  std::vector<int> col(B);
  for (int b = 0; b < B; ++b) col[b] = tokens_host[b];
  CUDA_CHECK(cudaMemcpy2DAsync(st_.tokens + offset, kTextCtx * sizeof(int), col.data(), sizeof(int), sizeof(int), B, cudaMemcpyHostToDevice, stream_));
  const int offs4[4] = {offset, offset, offset, offset};
  CUDA_CHECK(cudaMemcpyAsync(step_ctr_, offs4, sizeof(offs4), cudaMemcpyHostToDevice, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  enqueue_decode_step(B, true, false, 0);
  if (logits_host)
    CUDA_CHECK(cudaMemcpy2DAsync(logits_host, (size_t)cfg_.n_vocab * 4, logits_, (size_t)vocab_pad_ * 4, (size_t)cfg_.n_vocab * 4, B,
                                 cudaMemcpyDeviceToHost, stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  if (this_k || this_v) {
    // rows just appended to the bf16 caches, returned as f32 [L][B][d]
    float* tmp = nullptr;
    CUDA_CHECK(cudaMalloc(&tmp, (size_t)B * d * 4));
    for (int kv = 0; kv < 2; ++kv) {
      float* dst = kv == 0 ? this_k : this_v;
      if (!dst) continue;
      const __nv_bfloat16* cache = kv == 0 ? self_k_ : self_v_;
      for (int l = 0; l < L; ++l) {
        launch_kv_export(cache + (size_t)l * cap_ * H * kTextCtx * 64 + (size_t)offset * 64, tmp, B, 1, kTextCtx, 1, H, stream_);
        CUDA_CHECK(cudaMemcpyAsync(dst + (size_t)l * B * d, tmp, (size_t)B * d * 4, cudaMemcpyDeviceToHost, stream_));
      }
    }
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    cudaFree(tmp);
  }
}

// ---- K/V caches across the model-ABI boundary (the reference's decoder graph takes self_k/v and cross_k/v as INPUTS and
// returns the new cache rows, export_onnx.py:668-670, Whisper.cpp:306-313; here they are resident, so a caller that wants to
// supply or inspect them goes through these converters: f32 token-major <-> bf16 head-major) ----
void Engine::export_cache(const __nv_bfloat16* cache, int T, int b0, int nb, int n_rows, float* out) const {
  // out [L][nb][n_rows][d] f32
  const int d = cfg_.d, H = cfg_.n_head, L = cfg_.l_dec;
  if (b0 < 0 || nb <= 0 || b0 + nb > cap_) throw std::runtime_error("cache export: sequence range exceeds the resident capacity");
  if (n_rows <= 0 || n_rows > T) throw std::runtime_error("cache export: row count out of range");
  CUDA_CHECK(cudaSetDevice(device_));
  float* tmp = nullptr;
  const size_t per = (size_t)nb * n_rows * d;
  CUDA_CHECK(cudaMalloc(&tmp, per * 4));
  for (int l = 0; l < L; ++l) {
    launch_kv_export(cache + ((size_t)l * cap_ + b0) * H * T * 64, tmp, nb, n_rows, T, n_rows, H, stream_);
    CUDA_CHECK(cudaMemcpyAsync(out + (size_t)l * per, tmp, per * 4, cudaMemcpyDeviceToHost, stream_));
  }
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  cudaFree(tmp);
}

void Engine::import_cache(__nv_bfloat16* cache, int T, int B, int n_rows, const float* in) {
  // in [L][B][T][d] f32 (the reference's full-size tensors); rows [0, n_rows) are loaded
  const int d = cfg_.d, H = cfg_.n_head, L = cfg_.l_dec;
  check_batch(B, "cache import");
  if (n_rows < 0 || n_rows > T) throw std::runtime_error("cache import: row count out of range");
  if (n_rows == 0) return;
  CUDA_CHECK(cudaSetDevice(device_));
  float* tmp = nullptr;
  const size_t per = (size_t)B * T * d;
  CUDA_CHECK(cudaMalloc(&tmp, per * 4));
  for (int l = 0; l < L; ++l) {
    CUDA_CHECK(cudaMemcpyAsync(tmp, in + (size_t)l * per, per * 4, cudaMemcpyHostToDevice, stream_));
    launch_kv_import(tmp, cache + (size_t)l * cap_ * H * T * 64, B, n_rows, T, T, H, stream_);
  }
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  cudaFree(tmp);
}

void Engine::read_cross_kv(int b0, int nb, float* cross_k, float* cross_v) const {
  if (cross_k) export_cache(cross_k_, kAudioCtx, b0, nb, kAudioCtx, cross_k);
  if (cross_v) export_cache(cross_v_, kAudioCtx, b0, nb, kAudioCtx, cross_v);
}
void Engine::load_cross_kv(int B, const float* cross_k, const float* cross_v) {
  ensure_capacity(B);
  if (cross_k) import_cache(cross_k_, kAudioCtx, B, kAudioCtx, cross_k);
  if (cross_v) import_cache(cross_v_, kAudioCtx, B, kAudioCtx, cross_v);
}
void Engine::read_self_kv(int B, int n_rows, float* self_k, float* self_v) const {
  if (self_k) export_cache(self_k_, kTextCtx, 0, B, n_rows, self_k);
  if (self_v) export_cache(self_v_, kTextCtx, 0, B, n_rows, self_v);
}
void Engine::load_self_kv(int B, int n_valid, const float* self_k, const float* self_v) {
  ensure_capacity(B);
  if (self_k) import_cache(self_k_, kTextCtx, B, n_valid, self_k);
  if (self_v) import_cache(self_v_, kTextCtx, B, n_valid, self_v);
}

void Engine::read_encoder_hidden(int B, float* out) const {
  if (B > enc_sub_) throw std::runtime_error("read_encoder_hidden: batch exceeds the encoder sub-batch");
  CUDA_CHECK(cudaMemcpy(out, x_enc_, (size_t)B * kAudioCtx * cfg_.d * 4, cudaMemcpyDeviceToHost));
}

// ---------------------------------------------------------------------------------------------------------
// whole pipeline
// ---------------------------------------------------------------------------------------------------------
void Engine::transcribe_resident(int B, int max_samples, const std::string& lang, const DecodeOptions& opt,
                                 std::vector<std::vector<int>>* tokens, StageTimes* times) {
  CUDA_CHECK(cudaSetDevice(device_));
  cudaEvent_t ev[4];
  for (auto& e : ev) CUDA_CHECK(cudaEventCreate(&e));
  const long l0 = launches_;
  CUDA_CHECK(cudaEventRecord(ev[0], stream_));
  run_logmel(B, max_samples);
  CUDA_CHECK(cudaEventRecord(ev[1], stream_));
  run_encoder(B);
  CUDA_CHECK(cudaEventRecord(ev[2], stream_));
  const int steps = run_decode(B, sot_sequence(lang), opt, tokens);
  CUDA_CHECK(cudaEventRecord(ev[3], stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  if (times) {
    CUDA_CHECK(cudaEventElapsedTime(&times->mel_ms, ev[0], ev[1]));
    CUDA_CHECK(cudaEventElapsedTime(&times->encoder_ms, ev[1], ev[2]));
    CUDA_CHECK(cudaEventElapsedTime(&times->decode_ms, ev[2], ev[3]));
    CUDA_CHECK(cudaEventElapsedTime(&times->total_ms, ev[0], ev[3]));
    times->decode_steps = steps;
    times->kernel_launches = launches_ - l0;
  }
  for (auto& e : ev) cudaEventDestroy(e);
}

void Engine::transcribe(const float* const* pcm, const int* n_samples, int B, const std::string& lang, const DecodeOptions& opt,
                        std::vector<std::vector<int>>* tokens, StageTimes* times) {
  CUDA_CHECK(cudaSetDevice(device_));
  int max_samples = 0;
  for (int b = 0; b < B; ++b) {
    if (n_samples[b] < 201) throw std::runtime_error("audio shorter than 201 samples (reflect padding needs n_fft/2 + 1)");
    max_samples = std::max(max_samples, n_samples[b]);
  }
  ensure_capacity(B, max_samples);
  // Front end pipelined against the host->device copies: the PCM of encoder sub-batch s+1 is copied on the second stream
  // while sub-batch s runs log-mel + encoder on the main stream.
  cudaEvent_t ev[5];
  for (auto& e : ev) CUDA_CHECK(cudaEventCreate(&e));
  const long l0 = launches_;
  CUDA_CHECK(cudaEventRecord(ev[0], stream_));
  CUDA_CHECK(cudaMemcpyAsync(n_samples_, n_samples, sizeof(int) * B, cudaMemcpyHostToDevice, stream_));
  CUDA_CHECK(cudaStreamWaitEvent(stream2_, ev[0], 0));  // the copy stream starts after everything queued so far
  const int n_sub = (B + enc_sub_ - 1) / enc_sub_;
  if ((int)copy_events_.size() < n_sub) {
    const size_t old = copy_events_.size();
    copy_events_.resize(n_sub);
    for (size_t i = old; i < copy_events_.size(); ++i) CUDA_CHECK(cudaEventCreateWithFlags(&copy_events_[i], cudaEventDisableTiming));
  }
  for (int s = 0; s < n_sub; ++s) {
    const int b0 = s * enc_sub_, nb = std::min(enc_sub_, B - b0);
    for (int b = b0; b < b0 + nb; ++b)
      CUDA_CHECK(cudaMemcpyAsync(pcm_ + (size_t)b * pcm_stride_, pcm[b], (size_t)n_samples[b] * 4, cudaMemcpyHostToDevice, stream2_));
    CUDA_CHECK(cudaEventRecord(copy_events_[s], stream2_));
  }
  for (int s = 0; s < n_sub; ++s) {
    const int b0 = s * enc_sub_, nb = std::min(enc_sub_, B - b0);
    CUDA_CHECK(cudaStreamWaitEvent(stream_, copy_events_[s], 0));
    if (s == 0) CUDA_CHECK(cudaEventRecord(ev[1], stream_));  // first PCM has landed: exposed part of the copy
    run_logmel_range(b0, nb, max_samples);
    run_encoder_range(b0, nb);
  }
  CUDA_CHECK(cudaEventRecord(ev[2], stream_));
  const int steps = run_decode(B, sot_sequence(lang), opt, tokens);
  CUDA_CHECK(cudaEventRecord(ev[3], stream_));
  CUDA_CHECK(cudaStreamSynchronize(stream_));
  if (times) {
    CUDA_CHECK(cudaEventElapsedTime(&times->h2d_ms, ev[0], ev[1]));
    float front = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&front, ev[1], ev[2]));
    times->mel_ms = 0.f;           // interleaved with the encoder per sub-batch in this path
    times->encoder_ms = front;     // log-mel + encoder (+ the copies they overlap)
    CUDA_CHECK(cudaEventElapsedTime(&times->decode_ms, ev[2], ev[3]));
    CUDA_CHECK(cudaEventElapsedTime(&times->total_ms, ev[0], ev[3]));
    times->decode_steps = steps;
    times->kernel_launches = launches_ - l0;
  }
  for (auto& e : ev) cudaEventDestroy(e);
}

}  // namespace b200w
