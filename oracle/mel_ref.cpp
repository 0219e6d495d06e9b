// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or executed from the product path.
//
// oracle/_ref/libmel_ref.so: the reference's OWN log-mel frontend, compiled from the sources where
// they lie under /root/reference (include path only; nothing is copied into this repo).
//
//   * librosa::Feature::melspectrogram      <- /root/reference/cpp/src/librosa/librosa.h:218-229
//     (reflect pad :46-77, Hann STFT over Eigen kissfft :79-96, |X|^2 :98-100, Slaney bank :102-144)
//   * the log10 / global-max / clamp / (x+4)/4 / resize(3000) tail is a restatement of
//     Whisper::preprocess, /root/reference/cpp/src/Whisper.cpp:151-184 (that TU cannot be compiled
//     here: it includes the AXera BSP and OpenCC headers), with the same float/double typing:
//       log10f(max(x,1e-10f)); running max over ALL mel rows and ALL frames (:157-167);
//       (double(max(L, float(double(mmax) - 8.0))) + 4.0) / 4.0 rounded to float (:171);
//       rows cropped / zero-filled to 3000 AFTER normalisation (:172).
//     The reference's out-of-bounds write at :169-172 is not reproduced; its
//     result (first 3000 frames kept) is.
//
// Extra exports dump the constants the reference computes on the fly (window, mel bank, FFT
// twiddles) so that tools/gen_mel_tables.py can pin the product's constant tables to the exact
// bits of the reference build, and tests can assert they still agree.
#include <librosa/librosa.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <limits>
#include <vector>

extern "C" {

// returns number of STFT frames (1 + n/160 for centred reflect padding), <0 on error
__attribute__((visibility("default")))
int melref_preprocess(const float* pcm, int n_samples, int n_mels, float* out /* [n_mels*3000] */) {
  if (!pcm || !out || n_samples < 201) return -1;  // reflect pad needs x[200]
  std::vector<float> audio(pcm, pcm + n_samples);
  auto mel = librosa::Feature::melspectrogram(audio, 16000, 400, 160, "hann", true, "reflect", 2.0f,
                                              n_mels, 0.0f, 16000 / 2.0f);
  const int n_len = (int)mel[0].size();
  float mmax = -std::numeric_limits<float>::max();
  for (int i = 0; i < n_mels; ++i)
    for (int n = 0; n < n_len; ++n) {
      float v = std::log10(std::max(mel[i][n], 1e-10f));
      mel[i][n] = v;
      if (v > mmax) mmax = v;
    }
  const float floor_v = (float)(mmax - 8.0);
  for (int i = 0; i < n_mels; ++i) {
    float* dst = out + (size_t)i * 3000;
    const int keep = std::min(n_len, 3000);
    for (int n = 0; n < keep; ++n) dst[n] = (float)((std::max(mel[i][n], floor_v) + 4.0) / 4.0);
    for (int n = keep; n < 3000; ++n) dst[n] = 0.0f;
  }
  return n_len;
}

// raw power-mel (before log), [n_mels][n_frames]; returns n_frames
__attribute__((visibility("default")))
int melref_melspectrogram(const float* pcm, int n_samples, int n_mels, float* out, int out_capacity_frames) {
  if (!pcm || !out || n_samples < 201) return -1;
  std::vector<float> audio(pcm, pcm + n_samples);
  auto mel = librosa::Feature::melspectrogram(audio, 16000, 400, 160, "hann", true, "reflect", 2.0f,
                                              n_mels, 0.0f, 16000 / 2.0f);
  const int n_len = (int)mel[0].size();
  if (n_len > out_capacity_frames) return -2;
  for (int i = 0; i < n_mels; ++i) std::memcpy(out + (size_t)i * n_len, mel[i].data(), sizeof(float) * n_len);
  return n_len;
}

// complex STFT, [n_frames][201][2]; returns n_frames
__attribute__((visibility("default")))
int melref_stft(const float* pcm, int n_samples, float* out, int out_capacity_frames) {
  if (!pcm || !out || n_samples < 201) return -1;
  std::vector<float> audio(pcm, pcm + n_samples);
  auto X = librosa::Feature::stft(audio, 400, 160, "hann", true, "reflect");
  const int nf = (int)X.size();
  if (nf > out_capacity_frames) return -2;
  for (int f = 0; f < nf; ++f)
    for (int k = 0; k < 201; ++k) {
      out[((size_t)f * 201 + k) * 2 + 0] = X[f][k].real();
      out[((size_t)f * 201 + k) * 2 + 1] = X[f][k].imag();
    }
  return nf;
}

// the Hann window exactly as librosa.h:81 evaluates it (Eigen float expression, vectorised cos)
__attribute__((visibility("default")))
void melref_window(float* out400) {
  const int n_fft = 400;
  librosa::Vectorf window =
      0.5 * (1.f - (librosa::Vectorf::LinSpaced(n_fft, 0.f, static_cast<float>(n_fft - 1)) * 2.f * M_PI / n_fft).array().cos());
  for (int i = 0; i < n_fft; ++i) out400[i] = window[i];
}

// the Slaney mel bank exactly as librosa.h:102-144 evaluates it; out is [n_mels][201]
__attribute__((visibility("default")))
void melref_bank(int n_mels, float* out) {
  librosa::Matrixf w = librosa::internal::melfilter(16000, 400, n_mels, 0, 8000);
  for (int i = 0; i < n_mels; ++i)
    for (int k = 0; k < 201; ++k) out[(size_t)i * 201 + k] = w(i, k);
}

// kissfft twiddles as ei_kissfft_impl.h:29-37 (200-pt complex plan) and :388-401 (real recombine)
__attribute__((visibility("default")))
void melref_twiddles(float* tw200x2, float* rtw100x2) {
  typedef std::complex<float> C;
  const int nfft = 200;
  float phinc = -2 * std::acos((float)-1) / nfft;
  for (int i = 0; i < nfft; ++i) {
    C t = std::exp(C(0, i * phinc));
    tw200x2[2 * i] = t.real();
    tw200x2[2 * i + 1] = t.imag();
  }
  const int ncfft2 = 100, ncfft = 200;
  float pi = std::acos(float(-1));
  for (int k = 1; k <= ncfft2; ++k) {
    C t = std::exp(C(0, -pi * (float(k) / ncfft + float(.5))));
    rtw100x2[2 * (k - 1)] = t.real();
    rtw100x2[2 * (k - 1) + 1] = t.imag();
  }
}

}  // extern "C"
