// Host-side helpers around the boundary: WAV reading and base64 token decoding.
#pragma once
#include <string>
#include <vector>

namespace b200w {

struct WavData {
  int sample_rate = 0;
  int bits_per_sample = 0;
  std::vector<std::vector<float>> channels;  // [channel][frame], floats in [-1, 1)
};
// RIFF/WAVE: PCM 8/16/24/32-bit and IEEE float32 (incl. WAVE_FORMAT_EXTENSIBLE).  The reference reads WAV through
// AudioFile<float>::load (/root/reference/cpp/src/AudioFile.h:450,:501-640); sample scaling follows it:
// 8-bit (u - 128)/128, 16-bit s/32768 (:1241-1243), 24-bit s/8388608, 32-bit s/2147483648.
// FORM/AIFF and FORM/AIFC files (big-endian PCM 8/16/24/32, AIFC fl32) are read too, like AudioFile<float>::load does
// (AudioFile.h:490,:643-770).
bool load_wav(const std::string& path, WavData* out, std::string* err);

// Length-safe base64 (standard alphabet, '=' padding); invalid characters end the decode.
// Replaces /root/reference/cpp/src/base64.cpp:84-120, which writes into a caller-sized char buffer.
std::string base64_decode(const std::string& in);

// {type}-tokens.txt: one "<base64(token bytes)> <rank>" line per id, line index = id (export_onnx.py:391-417; loaded by
// Whisper.cpp:115-127).  detokenize() is Whisper.cpp:224-229 without its 32-byte stack buffer: ids outside the table
// (specials >= 50257) carry no text and are skipped, like python/whisper.py:258-260.
struct TokenTable {
  std::vector<std::string> b64;  // as read; decoded on use
  bool load(const std::string& path, std::string* err);
  std::string detokenize(const int* ids, size_t n) const;
};

}  // namespace b200w
