// whisper_cli for the B200 build: same flags and stdout lines as the reference's
// /root/reference/cpp/whisper_cli.cpp:19-110 (--wav/-w, --model_type/-t default "turbo", --model_path/-p,
// --language default "zh"; prints "Result: ..." and "RTF: ..." where RTF = wall(RunFile) / audio duration).
// Device initialisation (the reference's AX_SYS_Init / AX_ENGINE_Init, :37-61) lives inside AX_WHISPER_Init.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ax_whisper_api.h"
#include "host_utils.h"

static void usage(const char* argv0) {
  fprintf(stderr,
          "usage: %s --wav=string [options] ...\n"
          "options:\n"
          "  -w, --wav           wav file (string)\n"
          "  -t, --model_type    tiny, base, small, turbo, large (string [=turbo])\n"
          "  -p, --model_path    model path which contains tiny/ base/ small/ turbo/ (string [=../models-b200])\n"
          "      --language      en, zh (string [=zh])\n"
          "      --long          transcribe the whole file in 30 s windows (extension; the reference stops after 30 s)\n"
          "  -?, --help          print this message\n",
          argv0);
}

int main(int argc, char** argv) {
  std::string wav_file, model_type = "turbo", model_path = "../models-b200", language = "zh";
  bool long_form = false;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--long") {
      long_form = true;
      continue;
    }
    auto value = [&](const char* long_name, const char* short_name, std::string* dst) {
      const std::string ln = std::string("--") + long_name;
      if (a.rfind(ln + "=", 0) == 0) {
        *dst = a.substr(ln.size() + 1);
        return true;
      }
      if (a == ln || (short_name && a == short_name)) {
        if (i + 1 >= argc) {
          fprintf(stderr, "option needs value: %s\n", a.c_str());
          exit(1);
        }
        *dst = argv[++i];
        return true;
      }
      return false;
    };
    if (value("wav", "-w", &wav_file) || value("model_type", "-t", &model_type) || value("model_path", "-p", &model_path) ||
        value("language", nullptr, &language))
      continue;
    usage(argv[0]);
    return (a == "-?" || a == "--help") ? 0 : 1;
  }
  if (wav_file.empty()) {
    fprintf(stderr, "need option: --wav\n");
    usage(argv[0]);
    return 1;
  }
  printf("wav_file: %s\n", wav_file.c_str());
  printf("model_path: %s\n", model_path.c_str());
  printf("model_type: %s\n", model_type.c_str());
  printf("language: %s\n", language.c_str());

  b200w::WavData wav;
  std::string err;
  if (!b200w::load_wav(wav_file, &wav, &err)) {
    printf("load wav failed!\n");
    return -1;
  }
  const float duration = wav.channels[0].size() * 1.f / 16000;

  auto t0 = std::chrono::steady_clock::now();
  AX_WHISPER_HANDLE handle = AX_WHISPER_Init(model_type.c_str(), model_path.c_str(), language.c_str());
  auto t1 = std::chrono::steady_clock::now();
  if (!handle) {
    printf("AX_WHISPER_Init failed!\n");
    return -1;
  }
  printf("Init whisper success, take %.4fseconds\n", std::chrono::duration<double>(t1 - t0).count());

  t0 = std::chrono::steady_clock::now();
  char* result = nullptr;
  int rc;
  if (long_form) {  // same channel handling as RunFile (ax_whisper_api.cpp:105-113), every 30 s window instead of the first
    std::vector<float>& s = wav.channels[0];
    if (wav.channels.size() == 2)
      for (size_t i = 0; i < s.size(); ++i) s[i] = (s[i] + wav.channels[1][i]) / 2;
    rc = AX_WHISPER_RunPCMLong(handle, s.data(), (long)s.size(), 0, &result);
  } else {
    rc = AX_WHISPER_RunFile(handle, wav_file.c_str(), &result);
  }
  if (0 != rc) {
    printf("AX_WHISPER_Run failed!\n");
    AX_WHISPER_Uninit(handle);
    return -1;
  }
  t1 = std::chrono::steady_clock::now();
  printf("Result: %s\n", result);
  printf("RTF: %.4f\n", std::chrono::duration<double>(t1 - t0).count() / duration);
  free(result);
  AX_WHISPER_Uninit(handle);
  return 0;
}
