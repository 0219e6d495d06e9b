#include "host_utils.h"

#include <cstdint>
#include <cstring>
#include <fstream>

namespace b200w {

namespace {
uint32_t rd_u32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t rd_u16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
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
  if (size < 12 || memcmp(buf.data(), "RIFF", 4) != 0 || memcmp(buf.data() + 8, "WAVE", 4) != 0) return fail("not a RIFF/WAVE file");
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
  static int8_t lut[256];
  static bool init = false;
  if (!init) {
    memset(lut, -1, sizeof(lut));
    const char* abc = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    for (int i = 0; i < 64; ++i) lut[(unsigned char)abc[i]] = (int8_t)i;
    init = true;
  }
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

}  // namespace b200w
