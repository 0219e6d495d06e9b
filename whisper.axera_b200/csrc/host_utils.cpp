#include "host_utils.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>

namespace b200w {

namespace {
uint32_t rd_u32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t rd_u16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
}  // namespace

namespace {
uint32_t be_u32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }
uint16_t be_u16(const unsigned char* p) { return (uint16_t)((p[0] << 8) | p[1]); }

// IEEE 754 80-bit extended (big endian) -> double: the COMM chunk stores the sample rate this way
double be_extended80(const unsigned char* p) {
  const int sign = p[0] >> 7;
  const int exp = ((p[0] & 0x7f) << 8) | p[1];
  uint64_t mant = 0;
  for (int i = 0; i < 8; ++i) mant = (mant << 8) | p[2 + i];
  if (exp == 0 && mant == 0) return 0.0;
  const double v = std::ldexp((double)mant, exp - 16383 - 63);
  return sign ? -v : v;
}

// FORM/AIFF and FORM/AIFC (uncompressed big-endian PCM 8/16/24/32; AIFC 32-bit samples are IEEE floats).  Sample scaling
// as the reference's reader: s8/128, s16/32768, s24/8388608, s32/(2^31 - 1) (/root/reference/cpp/src/AudioFile.h:643-770).
bool decode_aiff(const std::vector<unsigned char>& buf, const std::string& path, WavData* out, std::string* err) {
  auto fail = [&](const char* m) {
    if (err) *err = path + ": " + m;
    return false;
  };
  const size_t size = buf.size();
  const bool aifc = memcmp(buf.data() + 8, "AIFC", 4) == 0;
  int n_ch = 0, bits = 0;
  uint32_t n_frames = 0;
  const unsigned char* data = nullptr;
  size_t data_len = 0;
  bool have_comm = false;
  size_t pos = 12;
  while (pos + 8 <= size) {
    const uint32_t len = be_u32(buf.data() + pos + 4);
    const unsigned char* body = buf.data() + pos + 8;
    const size_t avail = size - (pos + 8);
    if (memcmp(buf.data() + pos, "COMM", 4) == 0) {
      if (len < 18 || avail < 18) return fail("short COMM chunk");
      n_ch = (int16_t)be_u16(body);
      n_frames = be_u32(body + 2);
      bits = (int16_t)be_u16(body + 6);
      out->sample_rate = (int)std::lround(be_extended80(body + 8));
      if (aifc && len >= 22 && avail >= 22) {
        const bool pcm = memcmp(body + 18, "NONE", 4) == 0 || memcmp(body + 18, "twos", 4) == 0;
        const bool f32 = memcmp(body + 18, "fl32", 4) == 0 || memcmp(body + 18, "FL32", 4) == 0;
        if (!pcm && !f32) return fail("unsupported AIFC compression (NONE / twos / fl32 expected)");
        if (f32 && bits != 32) return fail("AIFC fl32 with a bit depth other than 32");
      }
      have_comm = true;
    } else if (memcmp(buf.data() + pos, "SSND", 4) == 0) {
      if (len < 8 || avail < 8) return fail("short SSND chunk");
      const uint32_t offset = be_u32(body);
      const size_t body_len = std::min<size_t>(len, avail);
      if (8 + (size_t)offset > body_len) return fail("SSND offset beyond the chunk");
      data = body + 8 + offset;
      data_len = body_len - 8 - offset;
    }
    pos += 8 + (size_t)len + (len & 1);
  }
  if (!have_comm || !data) return fail("missing COMM or SSND chunk");
  if (out->sample_rate <= 0) return fail("unsupported AIFF sample rate");
  if (n_ch < 1) return fail("AIFF without channels");
  if (bits != 8 && bits != 16 && bits != 24 && bits != 32) return fail("unsupported AIFF bit depth (8/16/24/32 expected)");
  const int bps = bits / 8;
  const size_t frame_bytes = (size_t)bps * n_ch;
  const size_t frames = std::min<size_t>(n_frames, data_len / frame_bytes);  // tolerate a truncated file
  out->bits_per_sample = bits;
  out->channels.assign(n_ch, std::vector<float>(frames));
  for (size_t i = 0; i < frames; ++i)
    for (int c = 0; c < n_ch; ++c) {
      const unsigned char* p = data + i * frame_bytes + (size_t)c * bps;
      float v;
      if (bits == 8) {
        v = (float)(int8_t)p[0] / 128.0f;
      } else if (bits == 16) {
        v = (float)(int16_t)be_u16(p) / 32768.0f;
      } else if (bits == 24) {
        int32_t s = (int32_t)(((uint32_t)p[0] << 16) | ((uint32_t)p[1] << 8) | (uint32_t)p[2]);
        if (s & 0x800000) s |= ~0xFFFFFF;
        v = (float)s / 8388608.0f;
      } else if (aifc) {
        const uint32_t u = be_u32(p);
        memcpy(&v, &u, 4);
      } else {
        v = (float)(int32_t)be_u32(p) / 2147483647.0f;
      }
      out->channels[c][i] = v;
    }
  return true;
}
}  // namespace

bool load_wav(const std::string& path, WavData* out, std::string* err) {
  auto fail = [&](const char* m) {
    if (err) *err = path + ": " + m;
    return false;
  };
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) return fail("cannot open file");
  const size_t size = (size_t)f.tellg();
  f.seekg(0);
  std::vector<unsigned char> buf(size);
  f.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)size);
  if (size >= 12 && memcmp(buf.data(), "FORM", 4) == 0 && (memcmp(buf.data() + 8, "AIFF", 4) == 0 || memcmp(buf.data() + 8, "AIFC", 4) == 0))
    return decode_aiff(buf, path, out, err);
  if (size < 12 || memcmp(buf.data(), "RIFF", 4) != 0 || memcmp(buf.data() + 8, "WAVE", 4) != 0)
    return fail("not a RIFF/WAVE or FORM/AIFF file");
  int fmt_tag = 0, n_ch = 0, bits = 0, block_align = 0;
  const unsigned char* data = nullptr;
  size_t data_len = 0;
  size_t pos = 12;
  while (pos + 8 <= size) {
    const uint32_t len = rd_u32(buf.data() + pos + 4);
    const unsigned char* body = buf.data() + pos + 8;
    const size_t avail = size - (pos + 8);
    if (memcmp(buf.data() + pos, "fmt ", 4) == 0) {
      if (len < 16 || avail < 16) return fail("short fmt chunk");
      fmt_tag = rd_u16(body);
      n_ch = rd_u16(body + 2);
      out->sample_rate = (int)rd_u32(body + 4);
      block_align = rd_u16(body + 12);
      bits = rd_u16(body + 14);
      if (fmt_tag == 0xFFFE && len >= 26 && avail >= 26) fmt_tag = rd_u16(body + 24);  // WAVE_FORMAT_EXTENSIBLE sub-format
    } else if (memcmp(buf.data() + pos, "data", 4) == 0) {
      data = body;
      data_len = len <= avail ? len : avail;  // tolerate a wrong length field (streamed files)
      break;
    }
    pos += 8 + (size_t)len + (len & 1);
  }
  if (!data || n_ch <= 0) return fail("missing fmt or data chunk");
  if (!((fmt_tag == 1 && (bits == 8 || bits == 16 || bits == 24 || bits == 32)) || (fmt_tag == 3 && bits == 32)))
    return fail("unsupported sample format (PCM 8/16/24/32 or float32 expected)");
  const int bps = bits / 8;
  if (block_align < bps * n_ch) block_align = bps * n_ch;
  const size_t n_frames = data_len / (size_t)block_align;
  out->bits_per_sample = bits;
  out->channels.assign(n_ch, std::vector<float>(n_frames));
  for (size_t i = 0; i < n_frames; ++i)
    for (int c = 0; c < n_ch; ++c) {
      const unsigned char* p = data + i * block_align + (size_t)c * bps;
      float v;
      if (fmt_tag == 3) {
        memcpy(&v, p, 4);
      } else if (bits == 8) {
        v = ((int)p[0] - 128) / 128.0f;
      } else if (bits == 16) {
        v = (int16_t)rd_u16(p) / 32768.0f;
      } else if (bits == 24) {
        int32_t s = (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16));
        if (s & 0x800000) s |= ~0xFFFFFF;
        v = s / 8388608.0f;
      } else {
        v = (int32_t)rd_u32(p) / 2147483648.0f;
      }
      out->channels[c][i] = v;
    }
  return true;
}

std::string base64_decode(const std::string& in) {
  // function-local static with a lambda initialiser: built exactly once, thread-safe (concurrent Run* calls detokenise
  // outside every lock)
  static const std::array<int8_t, 256> lut = [] {
    std::array<int8_t, 256> t;
    t.fill(-1);
    const char* abc = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    for (int i = 0; i < 64; ++i) t[(unsigned char)abc[i]] = (int8_t)i;
    return t;
  }();
  std::string out;
  uint32_t acc = 0;
  int nbits = 0;
  for (unsigned char c : in) {
    if (c == '=') break;
    const int v = lut[c];
    if (v < 0) break;
    acc = (acc << 6) | (uint32_t)v;
    nbits += 6;
    if (nbits >= 8) {
      nbits -= 8;
      out.push_back((char)((acc >> nbits) & 0xFF));
    }
  }
  return out;
}

bool TokenTable::load(const std::string& path, std::string* err) {
  std::ifstream fs(path);
  if (!fs.is_open()) {
    if (err) *err = "Can NOT open " + path;
    return false;
  }
  b64.clear();
  std::string line;
  while (std::getline(fs, line)) b64.push_back(line.substr(0, line.find(' ')));
  return true;
}

std::string TokenTable::detokenize(const int* ids, size_t n) const {
  std::string s;
  for (size_t i = 0; i < n; ++i) {
    const int id = ids[i];
    if (id < 0 || (size_t)id >= b64.size()) continue;  // specials carry no text; the reference indexes out of bounds here
    s += base64_decode(b64[id]);
  }
  return s;
}

}  // namespace b200w
