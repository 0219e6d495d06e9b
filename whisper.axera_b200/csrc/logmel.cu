// K1: fused log-mel frontend for sm_100a.
//
// Replaces the reference's host frontend: librosa::Feature::melspectrogram
// (/root/reference/cpp/src/librosa/librosa.h:218-229 -> pad :46-77, stft :79-96, spectrogram :98-100,
// melfilter :102-144, melspectrogram :146-155) and Whisper::preprocess
// (/root/reference/cpp/src/Whisper.cpp:151-184).
//
// Numerics: the reference's STFT is Eigen's kissfft in fp32 (400-pt real FFT = 200-pt complex FFT with
// radices 4,2,5,5 plus a real-input recombination, eigen3/unsupported/Eigen/src/FFT/ei_kissfft_impl.h:
// work :62-97, bfly2 :100-107, bfly4 :110-131, bfly5 :164-227, real fwd :305-334).  For tonal audio the
// STFT cancels to ~1e-4 of the input amplitude, so any fp32 implementation with a different operation
// order differs from the reference by more than the 1e-4 parity gate near the (max - 8) floor.  This
// kernel therefore executes exactly the same butterflies in exactly the same order with round-to-nearest
// mul/add intrinsics (no FMA contraction) and the reference build's own window / twiddle constants
// (mel_tables.inc): the complex spectrum is bit-identical to the reference's.  The remaining steps
// (|X|^2, sparse mel triangles, log10) are sums of positive terms and agree to ~1e-6.
//
// Layout: PCM [B][pcm_stride] f32 -> pass 1 writes log10 mel power to out [B][n_mels][3000] f32 and an
// utterance max (all frames of the supplied audio, including frame >= 3000: Whisper.cpp:157-167);
// pass 2 applies max(., mmax-8), (x+4)/4, zero-fills frames the audio does not cover (:171-172), in place,
// and also emits the bf16 time-major copy [B][3002][n_mels] (one zero row either side) that feeds conv1.
#include <cfloat>
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace b200w {

#include "mel_tables.inc"

namespace {

constexpr int kFramesPerCta = 8;  // 32 KB of smem per CTA -> 7 CTAs per SM (32 frames: 2 CTAs, 2.42 ms per 256 chunks; 8: 1.59 ms)
constexpr int kThreads = 256;
constexpr int kNfft = 400;
constexpr int kHop = 160;
constexpr int kBins = 201;
constexpr int kPcmTile = kHop * (kFramesPerCta - 1) + kNfft;  // samples covered by the CTA's frames
constexpr int kOutFrames = 3000;

struct MelTablesDev {
  float window[400];
  float2 tw[200];
  float2 rtw[100];
  float weights[2][400];
  unsigned short start[2][128];
  unsigned short count[2][128];
  unsigned short woff[2][128];
};
__device__ MelTablesDev g_mel_tables;

struct C32 {
  float r, i;
};
__device__ __forceinline__ C32 cmul(C32 a, C32 b) {
  C32 o;
  o.r = __fsub_rn(__fmul_rn(a.r, b.r), __fmul_rn(a.i, b.i));
  o.i = __fadd_rn(__fmul_rn(a.r, b.i), __fmul_rn(a.i, b.r));
  return o;
}
__device__ __forceinline__ C32 cadd(C32 a, C32 b) { return {__fadd_rn(a.r, b.r), __fadd_rn(a.i, b.i)}; }
__device__ __forceinline__ C32 csub(C32 a, C32 b) { return {__fsub_rn(a.r, b.r), __fsub_rn(a.i, b.i)}; }
__device__ __forceinline__ C32 ldc(const float2* p) {
  float2 v = *p;
  return {v.x, v.y};
}
__device__ __forceinline__ void stc(float2* p, C32 v) { *p = make_float2(v.r, v.i); }

// kissfft radix-5 butterfly (ei_kissfft_impl.h:164-227) on elements F[0], F[m], .., F[4m] with twiddle step fs*u
__device__ __forceinline__ void bfly5(float2* F, int m, int u, int fs, const float2* tw, C32 ya, C32 yb) {
  C32 s0 = ldc(F);
  C32 s1 = cmul(ldc(F + m), ldc(tw + u * fs));
  C32 s2 = cmul(ldc(F + 2 * m), ldc(tw + 2 * u * fs));
  C32 s3 = cmul(ldc(F + 3 * m), ldc(tw + 3 * u * fs));
  C32 s4 = cmul(ldc(F + 4 * m), ldc(tw + 4 * u * fs));
  C32 s7 = cadd(s1, s4), s10 = csub(s1, s4), s8 = cadd(s2, s3), s9 = csub(s2, s3);
  C32 f0 = cadd(cadd(s0, s7), s8);
  C32 s5 = cadd(s0, C32{__fadd_rn(__fmul_rn(s7.r, ya.r), __fmul_rn(s8.r, yb.r)), __fadd_rn(__fmul_rn(s7.i, ya.r), __fmul_rn(s8.i, yb.r))});
  C32 s6 = C32{__fadd_rn(__fmul_rn(s10.i, ya.i), __fmul_rn(s9.i, yb.i)), __fsub_rn(-__fmul_rn(s10.r, ya.i), __fmul_rn(s9.r, yb.i))};
  C32 s11 = cadd(s0, C32{__fadd_rn(__fmul_rn(s7.r, yb.r), __fmul_rn(s8.r, ya.r)), __fadd_rn(__fmul_rn(s7.i, yb.r), __fmul_rn(s8.i, ya.r))});
  C32 s12 = C32{__fadd_rn(-__fmul_rn(s10.i, yb.i), __fmul_rn(s9.i, ya.i)), __fsub_rn(__fmul_rn(s10.r, yb.i), __fmul_rn(s9.r, ya.i))};
  stc(F, f0);
  stc(F + m, csub(s5, s6));
  stc(F + 4 * m, cadd(s5, s6));
  stc(F + 2 * m, cadd(s11, s12));
  stc(F + 3 * m, csub(s11, s12));
}

struct __align__(16) LogmelSmem {
  float pcm[kPcmTile + 8];
  float2 fft[kFramesPerCta][200];
  float pw[kFramesPerCta][kBins];
  float window[400];
  float2 tw[200];
  float2 rtw[100];
  float weights[400];
  unsigned short start[128], count[128], woff[128];
  float red[kThreads / 32];
};

// Per-utterance state of one launch, two ints: [0] running maximum as an order-preserving integer key, [1] arrivals of its frame
// blocks; one more int after the last utterance hands out the block tickets.
// Both start from the byte pattern 0x80 (one cudaMemsetAsync): as a key that is -3.4e38, below every log-mel value.
constexpr int kStateInit = (int)0x80808080;
__device__ __forceinline__ int float_key(float v) {
  const int k = __float_as_int(v);
  return k >= 0 ? k : k ^ 0x7fffffff;  // monotonic in v, an involution
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

constexpr int kNormFrames = 32;  // frames per tile of the normalisation pass: 128-byte rows of the [n_mels][3000] tensor
constexpr int kNormCtas = 12;    // normalisation CTAs per utterance (94 tiles of 32 frames between them)

// Normalisation CTA `part` (0 .. kNormCtas-1) of utterance b: waits until every frame CTA of the utterance has arrived, then
// normalises its share of the 3000 output frames against the utterance maximum and writes the bf16 time-major copy.
// These CTAs sit at the END of the utterance's blockIdx.x range: blocks are dispatched in order, so by the time one of them
// runs, every frame CTA it waits for has been dispatched (they never wait on anything) -- the wait cannot deadlock.
// (max(L, float(mmax - 8.0)) + 4.0) / 4.0 is evaluated in double by the reference (Whisper.cpp:171); L + 4 is exact in double
// and /4 is a power-of-two scaling, so the float result equals fl32(L + 4) * 0.25 exactly.
__device__ __forceinline__ void logmel_normalize_part(int b, int part, int n_frames, int n_mels, float* __restrict__ out,
                                                      int* __restrict__ utt_state, __nv_bfloat16* __restrict__ out_tm,
                                                      unsigned char* smem_raw) {
  const int tid = threadIdx.x;
  const int n_ctas = (n_frames + kFramesPerCta - 1) / kFramesPerCta;  // frame CTAs of this utterance that do arrive
  if (tid == 0) {
    const volatile int* tickets = utt_state + 2 * b + 1;
    while (*tickets - kStateInit < n_ctas) __nanosleep(200);
    __threadfence();
  }
  __syncthreads();
  const float mmax = key_float(*reinterpret_cast<volatile int*>(&utt_state[2 * b]));
  const float floor_v = __fadd_rn(mmax, -8.0f);
  float(*tile)[kNormFrames + 1] = reinterpret_cast<float(*)[kNormFrames + 1]>(smem_raw);  // [128][33]
  constexpr int kTiles = (kOutFrames + kNormFrames - 1) / kNormFrames;                    // 94
  for (int t = part; t < kTiles; t += kNormCtas) {
    const int t0 = t * kNormFrames;
    __syncthreads();  // the previous tile's transposed reads are done
    for (int it = tid; it < n_mels * kNormFrames; it += kThreads) {
      const int fr = it % kNormFrames, mel = it / kNormFrames;
      const int f = t0 + fr;
      float v = 0.f;
      if (f < kOutFrames) {
        float* p = out + ((long)b * n_mels + mel) * kOutFrames + f;
        if (f < n_frames) v = __fmul_rn(__fadd_rn(fmaxf(__ldcg(p), floor_v), 4.0f), 0.25f);
        *p = v;  // frames past the audio are zero-filled AFTER normalisation (Whisper.cpp:172)
      }
      tile[mel][fr] = v;
    }
    if (out_tm == nullptr) continue;
    __syncthreads();
    for (int it = tid; it < n_mels * kNormFrames; it += kThreads) {
      const int mel = it % n_mels, fr = it / n_mels;
      const int f = t0 + fr;
      if (f < kOutFrames) out_tm[((long)b * (kOutFrames + 2) + f + 1) * n_mels + mel] = __float2bfloat16_rn(tile[mel][fr]);
    }
  }
}

// The whole frontend in ONE launch.  One CTA = kFramesPerCta consecutive STFT frames of one utterance: reflect-padded PCM tile ->
// Hann window -> 400-point real FFT (the reference's kissfft butterfly order) -> |X|^2 -> mel triangles -> log10 -> global
// memory + the utterance's running maximum.  Whisper::preprocess normalises against the maximum over ALL frames of the
// utterance (Whisper.cpp:157-171), which only exists once every frame CTA of the utterance is done: the last kNormCtas blocks of
// every utterance's blockIdx.x range wait for that (arrival counter) and then do the second pass -- values mostly still in L2 --
// and write the bf16 time-major copy the first convolution reads.  (Round 1 used three launches: init, frames, normalise.)
__global__ void __launch_bounds__(kThreads) logmel_fused_kernel(const float* __restrict__ pcm, long pcm_stride,
                                                               const int* __restrict__ n_samples_arr, int n_mels, int bank,
                                                               float* __restrict__ out, int* __restrict__ utt_state /* [B][2] + 1 */,
                                                               __nv_bfloat16* __restrict__ out_tm /* [B][3002][n_mels] or null */, int n_utt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LogmelSmem& S = *reinterpret_cast<LogmelSmem*>(smem_raw);
  const int tid = threadIdx.x;
  // Work is handed out by an arrival ticket, not by blockIdx: a normalisation block only ever waits for blocks with LOWER
  // tickets, i.e. blocks that are already running -- independent of the order in which the hardware dispatches blocks.
  __shared__ int s_vid;
  if (tid == 0) s_vid = atomicAdd(&utt_state[2 * n_utt], 1) - kStateInit;
  __syncthreads();
  const int per_utt = gridDim.x;  // frame blocks first, then the utterance's kNormCtas normalisation blocks
  const int b = s_vid / per_utt, bx = s_vid - b * per_utt;
  const int n = n_samples_arr[b];
  const int n_frames = 1 + n / kHop;  // librosa.h:87 with centre padding of n_fft/2 each side
  const int n_frame_ctas = per_utt - kNormCtas;
  if (bx >= n_frame_ctas) {
    logmel_normalize_part(b, bx - n_frame_ctas, n_frames, n_mels, out, utt_state, out_tm, smem_raw);
    return;
  }
  const int f0 = bx * kFramesPerCta;
  if (f0 >= n_frames) return;
  const int nf = min(kFramesPerCta, n_frames - f0);
  const float* x = pcm + (long)b * pcm_stride;

  // tables -> smem
  for (int i = tid; i < 400; i += kThreads) {
    S.window[i] = g_mel_tables.window[i];
    S.weights[i] = g_mel_tables.weights[bank][i];
  }
  for (int i = tid; i < 200; i += kThreads) S.tw[i] = g_mel_tables.tw[i];
  for (int i = tid; i < 100; i += kThreads) S.rtw[i] = g_mel_tables.rtw[i];
  for (int i = tid; i < 128; i += kThreads) {
    S.start[i] = g_mel_tables.start[bank][i];
    S.count[i] = g_mel_tables.count[bank][i];
    S.woff[i] = g_mel_tables.woff[bank][i];
  }
  // reflect-padded PCM tile (librosa.h:50-56): padded index p -> x[200-p] | x[p-200] | x[n-2-(p-200-n)]
  const int p0 = f0 * kHop;
  const int tile_len = kHop * (nf - 1) + kNfft;
  for (int i = tid; i < tile_len; i += kThreads) {
    int p = p0 + i;
    int src = p < 200 ? 200 - p : (p < 200 + n ? p - 200 : n - 2 - (p - 200 - n));
    S.pcm[i] = x[src];
  }
  __syncthreads();

  // stage A: leaf gather + radix-5 (m=1, fstride=40). out o = q0*50+q1*25+q2*5+q3 <- in i = q0+4q1+8q2+40q3
  const C32 ya = ldc(&S.tw[40]), yb = ldc(&S.tw[80]);
  for (int it = tid; it < nf * 40; it += kThreads) {
    const int fr = it / 40, g = it % 40;
    const int q0 = g / 10, q1 = (g / 5) % 2, q2 = g % 5;
    const int ibase = q0 + 4 * q1 + 8 * q2;
    const float* xs = S.pcm + fr * kHop;
    float2* F = &S.fft[fr][g * 5];
#pragma unroll
    for (int q3 = 0; q3 < 5; ++q3) {
      const int i = ibase + 40 * q3;  // complex sample i = (x[2i], x[2i+1]) of the windowed frame (librosa.h:92)
      F[q3] = make_float2(__fmul_rn(S.window[2 * i], xs[2 * i]), __fmul_rn(S.window[2 * i + 1], xs[2 * i + 1]));
    }
    bfly5(F, 1, 0, 40, S.tw, ya, yb);
  }
  __syncthreads();
  // stage B: radix-5, m=5, fstride=8, 8 groups of 25
  for (int it = tid; it < nf * 40; it += kThreads) {
    const int fr = it / 40, r = it % 40;
    const int grp = r / 5, u = r % 5;
    bfly5(&S.fft[fr][grp * 25 + u], 5, u, 8, S.tw, ya, yb);
  }
  __syncthreads();
  // stage C: radix-2, m=25, fstride=4, 4 groups of 50 (ei_kissfft_impl.h:100-107)
  for (int it = tid; it < nf * 100; it += kThreads) {
    const int fr = it / 100, r = it % 100;
    const int grp = r / 25, k = r % 25;
    float2* F = &S.fft[fr][grp * 50];
    C32 t = cmul(ldc(F + 25 + k), ldc(&S.tw[k * 4]));
    C32 a = ldc(F + k);
    stc(F + 25 + k, csub(a, t));
    stc(F + k, cadd(a, t));
  }
  __syncthreads();
  // stage D: radix-4, m=50, fstride=1 (ei_kissfft_impl.h:110-131, forward transform)
  for (int it = tid; it < nf * 50; it += kThreads) {
    const int fr = it / 50, k = it % 50;
    float2* F = &S.fft[fr][0];
    const int m = 50;
    C32 s0 = cmul(ldc(F + k + m), ldc(&S.tw[k]));
    C32 s1 = cmul(ldc(F + k + 2 * m), ldc(&S.tw[2 * k]));
    C32 s2 = cmul(ldc(F + k + 3 * m), ldc(&S.tw[3 * k]));
    C32 fk = ldc(F + k);
    C32 s5 = csub(fk, s1);
    fk = cadd(fk, s1);
    C32 s3 = cadd(s0, s2);
    C32 s4 = csub(s0, s2);
    s4 = C32{s4.i, -s4.r};
    stc(F + k + 2 * m, csub(fk, s3));
    stc(F + k, cadd(fk, s3));
    stc(F + k + m, cadd(s5, s4));
    stc(F + k + 3 * m, csub(s5, s4));
  }
  __syncthreads();
  // real-input recombination (ei_kissfft_impl.h:305-334) + power spectrum (librosa.h:98-100)
  for (int it = tid; it < nf * 101; it += kThreads) {
    const int fr = it / 101, k = it % 101;
    const float2* F = &S.fft[fr][0];
    float* P = S.pw[fr];
    if (k == 0) {
      C32 f = ldc(F);
      float dc = __fadd_rn(f.r, f.i), ny = __fsub_rn(f.r, f.i);
      P[0] = dc * dc;
      P[200] = ny * ny;
    } else {
      C32 fpk = ldc(F + k);
      C32 fq = ldc(F + 200 - k);
      C32 fpnk = C32{fq.r, -fq.i};
      C32 f1k = cadd(fpk, fpnk), f2k = csub(fpk, fpnk);
      C32 t = cmul(f2k, ldc(&S.rtw[k - 1]));
      C32 hi = csub(f1k, t);  // X[200-k] = conj(f1k - tw) * .5
      float hr = __fmul_rn(hi.r, 0.5f), hi_i = __fmul_rn(-hi.i, 0.5f);
      P[200 - k] = hr * hr + hi_i * hi_i;
      if (k != 100) {  // for k == 100 the reference's second store overwrites the first
        C32 lo = cadd(f1k, t);  // X[k] = (f1k + tw) * .5
        float lr = __fmul_rn(lo.r, 0.5f), li = __fmul_rn(lo.i, 0.5f);
        P[k] = lr * lr + li * li;
      }
    }
  }
  __syncthreads();
  // mel triangles (librosa.h:153) + log10(max(., 1e-10)) (Whisper.cpp:161) + running max (:163-165)
  float vmax = -FLT_MAX;
  for (int it = tid; it < n_mels * kFramesPerCta; it += kThreads) {
    const int fr = it % kFramesPerCta, mel = it / kFramesPerCta;
    if (fr >= nf) continue;
    const float* P = S.pw[fr] + S.start[mel];
    const float* w = S.weights + S.woff[mel];
    const int cnt = S.count[mel];
    float acc = 0.f;
    for (int j = 0; j < cnt; ++j) acc = fmaf(w[j], P[j], acc);
    const float L = log10f(fmaxf(acc, 1e-10f));
    vmax = fmaxf(vmax, L);
    const int f = f0 + fr;
    if (f < kOutFrames) out[((long)b * n_mels + mel) * kOutFrames + f] = L;
  }
  vmax = warp_max(vmax);
  if ((tid & 31) == 0) S.red[tid >> 5] = vmax;
  __threadfence();  // this thread's log-mel values are visible device-wide before the arrival below
  __syncthreads();
  if (tid == 0) {
    float v = S.red[0];
    for (int i = 1; i < kThreads / 32; ++i) v = fmaxf(v, S.red[i]);
    atomicMax(&utt_state[2 * b], float_key(v));
    __threadfence();
    atomicAdd(&utt_state[2 * b + 1], 1);  // arrival: the normalisation CTAs of this utterance wait for all of them
  }
}

// mel f32 [B][n_mels][3000] -> bf16 time-major [B][3002][n_mels] (rows 0 and 3001 stay zero); used when the
// encoder is entered with a caller-supplied mel tensor (model ABI) instead of through the fused frontend.
__global__ void __launch_bounds__(kThreads) mel_to_timemajor_kernel(const float* __restrict__ mel, int n_mels,
                                                                   __nv_bfloat16* __restrict__ out_tm) {
  __shared__ float tile[128][kFramesPerCta + 1];
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * kFramesPerCta;
  const int tid = threadIdx.x;
  for (int it = tid; it < n_mels * kFramesPerCta; it += kThreads) {
    const int fr = it % kFramesPerCta, mel_i = it / kFramesPerCta;
    const int f = f0 + fr;
    tile[mel_i][fr] = f < kOutFrames ? mel[((long)b * n_mels + mel_i) * kOutFrames + f] : 0.f;
  }
  __syncthreads();
  for (int it = tid; it < n_mels * kFramesPerCta; it += kThreads) {
    const int mel_i = it % n_mels, fr = it / n_mels;
    const int f = f0 + fr;
    if (f < kOutFrames) out_tm[((long)b * (kOutFrames + 2) + f + 1) * n_mels + mel_i] = __float2bfloat16_rn(tile[mel_i][fr]);
  }
}

}  // namespace

void logmel_upload_tables() {
  static MelTablesDev h;  // ~6 KB, built once
  auto bits = [](uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
  };
  for (int i = 0; i < 400; ++i) h.window[i] = bits(kHannWindowBits[i]);
  for (int i = 0; i < 200; ++i) h.tw[i] = make_float2(bits(kFftTwiddleBits[2 * i]), bits(kFftTwiddleBits[2 * i + 1]));
  for (int i = 0; i < 100; ++i) h.rtw[i] = make_float2(bits(kRealTwiddleBits[2 * i]), bits(kRealTwiddleBits[2 * i + 1]));
  for (int bank = 0; bank < 2; ++bank) {
    const int n_mels = bank == 0 ? 80 : 128;
    const uint16_t* st = bank == 0 ? kMelStart80 : kMelStart128;
    const uint16_t* ct = bank == 0 ? kMelCount80 : kMelCount128;
    const uint32_t* wb = bank == 0 ? kMelWeightBits80 : kMelWeightBits128;
    int off = 0;
    for (int m = 0; m < n_mels; ++m) {
      h.start[bank][m] = st[m];
      h.count[bank][m] = ct[m];
      h.woff[bank][m] = (unsigned short)off;
      for (int j = 0; j < ct[m]; ++j) h.weights[bank][off + j] = bits(wb[off + j]);
      off += ct[m];
    }
  }
  CUDA_CHECK(cudaMemcpyToSymbol(g_mel_tables, &h, sizeof(h)));
}

size_t logmel_mel_table_copy(int n_mels, float* dense_bank /* [n_mels][201] host, may be null */, float* window400) {
  // host-side view of the constant tables (for tests / introspection through the C ABI)
  auto bits = [](uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
  };
  if (window400)
    for (int i = 0; i < 400; ++i) window400[i] = bits(kHannWindowBits[i]);
  if (dense_bank) {
    const uint16_t* st = n_mels == 80 ? kMelStart80 : kMelStart128;
    const uint16_t* ct = n_mels == 80 ? kMelCount80 : kMelCount128;
    const uint32_t* wb = n_mels == 80 ? kMelWeightBits80 : kMelWeightBits128;
    memset(dense_bank, 0, sizeof(float) * n_mels * 201);
    int off = 0;
    for (int m = 0; m < n_mels; ++m) {
      for (int j = 0; j < ct[m]; ++j) dense_bank[m * 201 + st[m] + j] = bits(wb[off + j]);
      off += ct[m];
    }
  }
  return sizeof(MelTablesDev);
}

void launch_logmel(const float* pcm, long pcm_stride, const int* n_samples_dev, int max_samples, int B, int n_mels,
                   float* out_mel, __nv_bfloat16* out_tm, int* utt_state, cudaStream_t stream) {
  if (n_mels != 80 && n_mels != 128) throw CudaError("logmel: n_mels must be 80 or 128");
  const int bank = n_mels == 80 ? 0 : 1;
  const int max_frames = 1 + max_samples / kHop;
  CUDA_CHECK(cudaMemsetAsync(utt_state, 0x80, sizeof(int) * (2 * (size_t)B + 1), stream));  // kStateInit: maxima, arrivals, block tickets
  dim3 g1((max_frames + kFramesPerCta - 1) / kFramesPerCta + kNormCtas, B);  // frame CTAs, then the utterance's normalisation CTAs
  logmel_fused_kernel<<<g1, kThreads, sizeof(LogmelSmem), stream>>>(pcm, pcm_stride, n_samples_dev, n_mels, bank, out_mel, utt_state, out_tm, B);
  CUDA_CHECK(cudaGetLastError());
}

void logmel_set_attributes() {
  static_assert(sizeof(LogmelSmem) >= sizeof(float) * 128 * (kNormFrames + 1), "the normalisation tile aliases the FFT staging");
  CUDA_CHECK(cudaFuncSetAttribute(logmel_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LogmelSmem)));
}

void launch_mel_to_timemajor(const float* mel, int B, int n_mels, __nv_bfloat16* out_tm, cudaStream_t stream) {
  dim3 g((kOutFrames + kFramesPerCta - 1) / kFramesPerCta, B);
  mel_to_timemajor_kernel<<<g, kThreads, 0, stream>>>(mel, n_mels, out_tm);
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace b200w
