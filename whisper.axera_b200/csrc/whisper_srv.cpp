// whisper_srv -- HTTP front-end over the C API, the B200 build of the reference's server executable
// (/root/reference/cpp/whisper_srv.cpp:10-70 + /root/reference/cpp/src/WhisperHTTPServer.hpp:39-100).
//
// Same wire contract:  POST /asr, Content-Type application/octet-stream, body = raw little-endian float32 PCM (16 kHz mono)
//   200  {"success": true, "text": "..."}                              (pretty-printed with 2 spaces like nlohmann dump(2))
//   400  {"error": "Content-Type must be application/octet-stream"}     WhisperHTTPServer.hpp:49-54
//   400  {"error": "Request body is empty"}                             :57-61
//   400  {"error": "Data size must be multiple of 4 bytes"}             :64-70
//   400  {"error": "Run model failed!"}                                 :76-81
//   500  {"error": "Internal server error", "message": ...}             :91-98 (e.g. text that is not valid UTF-8, which
//                                                                        nlohmann::json::dump refuses as well)
//   CORS headers on every /asr response (:112-117); any other path / method: 404.
// Same command line (--port, -t/--model_type, -p/--model_path, -l/--language) and start-up lines.
//
// What is different: no httplib / nlohmann / cmdline dependency (a ~300-line HTTP/1.1 server on POSIX sockets, one thread per
// connection, keep-alive, Expect: 100-continue, chunked request bodies); the result string is free()d (the reference leaks it,
// WhisperHTTPServer.hpp:77-88); and concurrent posts do not race on the handle: AX_WHISPER_RunPCM is re-entrant here and
// requests that arrive while a GPU pass is running are coalesced into ONE batched pass (ax_whisper_api.cpp: run_coalesced).
// Extension: GET /stats -> {"requests": n, "gpu_passes": m} (AX_WHISPER_GetStats), --coalesce_wait_us, --max_batch, --devices.
// --no-model serves a stub transcriber for protocol tests on machines without a GPU (it says so loudly); there is no CPU path.
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <signal.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ax_whisper_api.h"

namespace {

constexpr size_t kMaxHeaderBytes = 64 * 1024;
constexpr size_t kMaxBodyBytes = (size_t)1 << 30;  // 1 GiB of f32 PCM = 4.6 h of audio; larger posts are refused with 413

std::atomic<bool> g_stop{false};
int g_listen_fd = -1;
AX_WHISPER_HANDLE g_model = nullptr;
bool g_no_model = false;
std::atomic<long> g_stub_requests{0};

struct Request {
  std::string method, path, version;
  std::map<std::string, std::string> headers;  // lower-cased names
  std::string body;
  bool keep_alive = true;
};
struct Response {
  int status = 200;
  std::string content_type = "application/json";
  std::string body;
  std::vector<std::pair<std::string, std::string>> headers;
};

const char* reason(int status) {
  switch (status) {
    case 200: return "OK";
    case 400: return "Bad Request";
    case 404: return "Not Found";
    case 411: return "Length Required";
    case 413: return "Payload Too Large";
    case 500: return "Internal Server Error";
    default: return "Error";
  }
}

bool send_all(int fd, const char* p, size_t n) {
  while (n > 0) {
    const ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
    if (k <= 0) {
      if (k < 0 && errno == EINTR) continue;
      return false;
    }
    p += k, n -= (size_t)k;
  }
  return true;
}

// buffered reader over one connection
struct Conn {
  int fd;
  std::string buf;
  size_t pos = 0;
  bool fill() {
    char tmp[65536];
    for (;;) {
      const ssize_t k = ::recv(fd, tmp, sizeof(tmp), 0);
      if (k > 0) {
        if (pos > 0 && pos == buf.size()) buf.clear(), pos = 0;
        buf.append(tmp, (size_t)k);
        return true;
      }
      if (k < 0 && errno == EINTR) continue;
      return false;
    }
  }
  // reads up to and including "\r\n"; false on EOF / oversize
  bool read_line(std::string* line, size_t limit) {
    for (;;) {
      const size_t e = buf.find("\r\n", pos);
      if (e != std::string::npos) {
        line->assign(buf, pos, e - pos);
        pos = e + 2;
        return true;
      }
      if (buf.size() - pos > limit) return false;
      if (!fill()) return false;
    }
  }
  bool read_exact(std::string* out, size_t n) {
    out->reserve(out->size() + std::min<size_t>(n, (size_t)16 << 20));  // a declared length is not trusted with memory up front
    while (n > 0) {
      if (pos == buf.size() && !fill()) return false;
      const size_t k = std::min(n, buf.size() - pos);
      out->append(buf, pos, k);
      pos += k, n -= k;
    }
    return true;
  }
};

std::string lower(std::string s) {
  for (char& c : s) c = (char)tolower((unsigned char)c);
  return s;
}
std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) ++a;
  while (b > a && isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}

// 0 = request parsed, 1 = clean EOF before a request, otherwise an HTTP status to answer with before closing
int read_request(Conn& c, Request* rq) {
  std::string line;
  do {
    if (!c.read_line(&line, kMaxHeaderBytes)) return 1;
  } while (line.empty());  // tolerate blank lines between pipelined requests
  {
    const size_t a = line.find(' '), b = line.rfind(' ');
    if (a == std::string::npos || b == a) return 400;
    rq->method = line.substr(0, a), rq->path = trim(line.substr(a + 1, b - a - 1)), rq->version = line.substr(b + 1);
    const size_t qm = rq->path.find('?');
    if (qm != std::string::npos) rq->path.resize(qm);
  }
  size_t header_bytes = 0;
  for (;;) {
    if (!c.read_line(&line, kMaxHeaderBytes)) return 400;
    if (line.empty()) break;
    header_bytes += line.size();
    if (header_bytes > kMaxHeaderBytes) return 400;
    const size_t colon = line.find(':');
    if (colon == std::string::npos) return 400;
    rq->headers[lower(trim(line.substr(0, colon)))] = trim(line.substr(colon + 1));
  }
  const std::string conn = lower(rq->headers.count("connection") ? rq->headers["connection"] : "");
  rq->keep_alive = rq->version == "HTTP/1.1" ? conn.find("close") == std::string::npos : conn.find("keep-alive") != std::string::npos;
  const bool chunked = rq->headers.count("transfer-encoding") && lower(rq->headers["transfer-encoding"]).find("chunked") != std::string::npos;
  const bool has_len = rq->headers.count("content-length") != 0;
  if (rq->headers.count("expect") && lower(rq->headers["expect"]).find("100-continue") != std::string::npos) {
    // curl sends this for bodies above 1 KiB and waits up to a second for the go-ahead
    static const char kContinue[] = "HTTP/1.1 100 Continue\r\n\r\n";
    if (!send_all(c.fd, kContinue, sizeof(kContinue) - 1)) return 1;
  }
  if (chunked) {
    for (;;) {
      if (!c.read_line(&line, 1024)) return 400;
      const size_t n = (size_t)strtoull(line.c_str(), nullptr, 16);
      if (n == 0) {
        while (c.read_line(&line, kMaxHeaderBytes) && !line.empty()) {
        }  // trailers
        break;
      }
      if (rq->body.size() + n > kMaxBodyBytes) return 413;
      if (!c.read_exact(&rq->body, n)) return 400;
      if (!c.read_line(&line, 16)) return 400;  // CRLF after the chunk
    }
  } else if (has_len) {
    const unsigned long long n = strtoull(rq->headers["content-length"].c_str(), nullptr, 10);
    if (n > kMaxBodyBytes) return 413;
    if (!c.read_exact(&rq->body, (size_t)n)) return 400;
  } else if (rq->method == "POST" || rq->method == "PUT") {
    return 411;
  }
  return 0;
}

bool write_response(int fd, const Response& rs, bool keep_alive) {
  std::string h = "HTTP/1.1 " + std::to_string(rs.status) + " " + reason(rs.status) + "\r\n";
  for (const auto& kv : rs.headers) h += kv.first + ": " + kv.second + "\r\n";
  h += "Content-Type: " + rs.content_type + "\r\n";
  h += "Content-Length: " + std::to_string(rs.body.size()) + "\r\n";
  h += keep_alive ? "Connection: keep-alive\r\n" : "Connection: close\r\n";
  h += "\r\n";
  return send_all(fd, h.data(), h.size()) && send_all(fd, rs.body.data(), rs.body.size());
}

// ---- JSON output (what nlohmann::json::dump(2) produces for a flat object of strings / booleans / integers) ----
// Returns false (and the offending index / byte) if `s` is not valid UTF-8: nlohmann throws type_error.316 there, which the
// reference turns into its 500 answer.
bool json_escape(const std::string& s, std::string* out, size_t* bad_index, unsigned* bad_byte) {
  out->clear();
  out->push_back('"');
  for (size_t i = 0; i < s.size();) {
    const unsigned char c = (unsigned char)s[i];
    if (c < 0x80) {
      switch (c) {
        case '"': *out += "\\\""; break;
        case '\\': *out += "\\\\"; break;
        case '\b': *out += "\\b"; break;
        case '\f': *out += "\\f"; break;
        case '\n': *out += "\\n"; break;
        case '\r': *out += "\\r"; break;
        case '\t': *out += "\\t"; break;
        default:
          if (c < 0x20 || c == 0x7f) {
            char u[8];
            snprintf(u, sizeof(u), "\\u%04x", c);
            *out += u;
          } else {
            out->push_back((char)c);
          }
      }
      ++i;
      continue;
    }
    int n = 0;
    unsigned cp = 0;
    if ((c & 0xE0) == 0xC0) n = 1, cp = c & 0x1F;
    else if ((c & 0xF0) == 0xE0) n = 2, cp = c & 0x0F;
    else if ((c & 0xF8) == 0xF0) n = 3, cp = c & 0x07;
    bool ok = n > 0 && i + (size_t)n < s.size();  // the continuation bytes must exist
    for (int k = 1; ok && k <= n; ++k) {
      const unsigned char cc = (unsigned char)s[i + k];
      if ((cc & 0xC0) != 0x80) ok = false;
      cp = (cp << 6) | (cc & 0x3F);
    }
    if (ok) {
      static const unsigned kMin[4] = {0, 0x80, 0x800, 0x10000};
      if (cp < kMin[n] || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) ok = false;  // overlong / out of range / surrogate
    }
    if (!ok) {
      *bad_index = i;
      *bad_byte = c;
      return false;
    }
    out->append(s, i, (size_t)n + 1);
    i += (size_t)n + 1;
  }
  out->push_back('"');
  return true;
}

void set_cors(Response* rs) {  // WhisperHTTPServer.hpp:112-117
  rs->headers.emplace_back("Access-Control-Allow-Origin", "*");
  rs->headers.emplace_back("Access-Control-Allow-Methods", "POST, GET, OPTIONS");
  rs->headers.emplace_back("Access-Control-Allow-Headers", "Content-Type, X-Array-Name, X-Array-Description, X-Array-Size");
}

void handle_asr(const Request& rq, Response* rs) {
  set_cors(rs);
  const auto ct = rq.headers.find("content-type");
  if (ct == rq.headers.end() || ct->second.find("application/octet-stream") == std::string::npos) {
    rs->status = 400;
    rs->body = R"({"error": "Content-Type must be application/octet-stream"})";
    return;
  }
  if (rq.body.empty()) {
    rs->status = 400;
    rs->body = R"({"error": "Request body is empty"})";
    return;
  }
  if (rq.body.size() % sizeof(float) != 0) {
    rs->status = 400;
    rs->body = R"({"error": "Data size must be multiple of 4 bytes"})";
    return;
  }
  std::vector<float> audio(rq.body.size() / sizeof(float));
  memcpy(audio.data(), rq.body.data(), rq.body.size());
  std::string text;
  if (g_no_model) {
    ++g_stub_requests;
    text = "stub: " + std::to_string(audio.size()) + " samples";
    if (audio.size() < 201) {
      rs->status = 400;
      rs->body = R"({"error": "Run model failed!"})";
      return;
    }
  } else {
    char* result = nullptr;
    if (0 != AX_WHISPER_RunPCM(g_model, audio.data(), (int)std::min<size_t>(audio.size(), 0x7fffffff), &result)) {
      fprintf(stderr, "run whisper failed!\n");
      rs->status = 400;
      rs->body = R"({"error": "Run model failed!"})";
      return;
    }
    text = result ? result : "";
    free(result);  // the caller owns *result (ax_whisper_api.h); the reference's server never frees it
  }
  std::string esc;
  size_t bad_i = 0;
  unsigned bad_b = 0;
  if (!json_escape(text, &esc, &bad_i, &bad_b)) {
    char msg[160];
    snprintf(msg, sizeof(msg), "invalid UTF-8 byte at index %zu: 0x%02X", bad_i, bad_b);
    rs->status = 500;
    rs->body = std::string("{\n  \"error\": \"Internal server error\",\n  \"message\": \"") + msg + "\"\n}";
    fprintf(stderr, "Error: %s\n", msg);
    return;
  }
  rs->body = "{\n  \"success\": true,\n  \"text\": " + esc + "\n}";
}

void handle_stats(Response* rs) {
  long n = 0, p = 0;
  if (g_no_model) n = p = g_stub_requests.load();
  else AX_WHISPER_GetStats(g_model, &n, &p);
  rs->body = "{\n  \"gpu_passes\": " + std::to_string(p) + ",\n  \"requests\": " + std::to_string(n) + "\n}";
}

void serve_connection(int fd) {
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
  Conn c{fd};
  for (;;) {
    Request rq;
    const int st = read_request(c, &rq);
    if (st == 1) break;
    Response rs;
    bool keep = false;
    if (st != 0) {
      rs.status = st;
      rs.body = std::string("{\"error\": \"") + reason(st) + "\"}";
    } else {
      keep = rq.keep_alive && !g_stop.load();
      try {
        if (rq.method == "POST" && rq.path == "/asr") handle_asr(rq, &rs);
        else if (rq.method == "GET" && rq.path == "/stats") handle_stats(&rs);
        else rs.status = 404, rs.content_type = "text/plain", rs.body = "";
      } catch (const std::exception& e) {  // WhisperHTTPServer.hpp:91-98
        rs = Response();
        set_cors(&rs);
        rs.status = 500;
        std::string esc;
        size_t bi;
        unsigned bb;
        if (!json_escape(e.what(), &esc, &bi, &bb)) esc = "\"?\"";
        rs.body = "{\n  \"error\": \"Internal server error\",\n  \"message\": " + esc + "\n}";
        fprintf(stderr, "Error: %s\n", e.what());
      }
    }
    if (!write_response(fd, rs, keep) || !keep) break;
  }
  ::shutdown(fd, SHUT_RDWR);
  ::close(fd);
}

void on_signal(int) {
  g_stop.store(true);
  if (g_listen_fd >= 0) ::shutdown(g_listen_fd, SHUT_RDWR);  // wakes accept()
}

struct Args {
  int port = 8080;
  std::string model_type = "turbo", model_path = "../models-b200", language = "zh", host = "0.0.0.0", port_file;
};

bool parse_args(int argc, char** argv, Args* a) {
  auto usage = [&] {
    fprintf(stderr,
            "usage: %s [options] ...\noptions:\n"
            "      --port                http port (int [=8080])\n"
            "  -t, --model_type          tiny, base, small, turbo, large (string [=turbo])\n"
            "  -p, --model_path          model path which contains tiny/ base/ small/ turbo/ (string [=../models-b200])\n"
            "  -l, --language            en, zh (string [=zh])\n"
            "      --host                listen address (string [=0.0.0.0])\n"
            "      --coalesce_wait_us    time a request waits for company before a GPU pass starts (int [=0])\n"
            "      --coalesce_max        most requests per GPU pass (int [=64])\n"
            "      --max_batch           initial batch capacity of the engine (int [=coalesce_max])\n"
            "      --devices             GPUs of the handle: all or 0,1,... (string [=B200W_DEVICE or 0])\n"
            "      --port_file           write the bound port here once listening (for --port 0)\n"
            "      --no-model            protocol test mode without a GPU: a stub answers instead of the model\n"
            "  -?, --help                print this message\n",
            argv[0]);
  };
  for (int i = 1; i < argc; ++i) {
    std::string k = argv[i], v;
    const size_t eq = k.find('=');
    bool has_v = false;
    if (k.rfind("--", 0) == 0 && eq != std::string::npos) v = k.substr(eq + 1), k = k.substr(0, eq), has_v = true;
    auto value = [&]() -> std::string {
      if (has_v) return v;
      if (i + 1 >= argc) {
        fprintf(stderr, "option needs value: %s\n", k.c_str());
        usage();
        exit(1);
      }
      return argv[++i];
    };
    if (k == "--port") a->port = atoi(value().c_str());
    else if (k == "-t" || k == "--model_type") a->model_type = value();
    else if (k == "-p" || k == "--model_path") a->model_path = value();
    else if (k == "-l" || k == "--language") a->language = value();
    else if (k == "--host") a->host = value();
    else if (k == "--coalesce_wait_us") setenv("B200W_COALESCE_WAIT_US", value().c_str(), 1);
    else if (k == "--coalesce_max") setenv("B200W_COALESCE_MAX", value().c_str(), 1);
    else if (k == "--max_batch") setenv("B200W_MAX_BATCH", value().c_str(), 1);
    else if (k == "--devices") setenv("B200W_DEVICES", value().c_str(), 1);
    else if (k == "--port_file") a->port_file = value();
    else if (k == "--no-model") g_no_model = true;
    else if (k == "-?" || k == "--help") {
      usage();
      exit(0);
    } else {
      fprintf(stderr, "undefined option: %s\n", k.c_str());
      usage();
      return false;
    }
  }
  return true;
}

}  // namespace

int main(int argc, char** argv) {
  Args args;
  if (!parse_args(argc, argv, &args)) return 1;
  // Device initialisation (the reference's AX_SYS_Init / AX_ENGINE_Init, whisper_srv.cpp:28-52) happens inside AX_WHISPER_Init.
  printf("port: %d\n", args.port);
  printf("model_path: %s\n", args.model_path.c_str());
  printf("model_type: %s\n", args.model_type.c_str());
  printf("language: %s\n", args.language.c_str());
  fflush(stdout);
  if (g_no_model) {
    fprintf(stderr, "[I] NO MODEL LOADED: --no-model serves a stub for protocol tests; nothing is transcribed\n");
  } else {
    if (!getenv("B200W_MAX_BATCH")) setenv("B200W_MAX_BATCH", getenv("B200W_COALESCE_MAX") ? getenv("B200W_COALESCE_MAX") : "64", 1);
    fprintf(stderr, "[I] Initializing server...\n");
    g_model = AX_WHISPER_Init(args.model_type.c_str(), args.model_path.c_str(), args.language.c_str());
    if (!g_model) {
      fprintf(stderr, "[E] whisper load models failed!\n");
      printf("init server failed!\n");
      return -1;
    }
    fprintf(stderr, "[I] Init server success\n");
  }
  g_listen_fd = ::socket(AF_INET, SOCK_STREAM, 0);
  if (g_listen_fd < 0) {
    perror("socket");
    return -1;
  }
  int one = 1;
  setsockopt(g_listen_fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  sockaddr_in addr{};
  addr.sin_family = AF_INET;
  addr.sin_port = htons((uint16_t)args.port);
  if (inet_pton(AF_INET, args.host.c_str(), &addr.sin_addr) != 1) {
    fprintf(stderr, "bad listen address %s\n", args.host.c_str());
    return -1;
  }
  if (::bind(g_listen_fd, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0 || ::listen(g_listen_fd, 256) != 0) {
    perror("bind/listen");
    return -1;
  }
  socklen_t alen = sizeof(addr);
  getsockname(g_listen_fd, reinterpret_cast<sockaddr*>(&addr), &alen);
  const int port = ntohs(addr.sin_port);
  struct sigaction sa {};
  sa.sa_handler = on_signal;
  sigaction(SIGINT, &sa, nullptr);
  sigaction(SIGTERM, &sa, nullptr);
  signal(SIGPIPE, SIG_IGN);
  fprintf(stderr, "[I] Start server at port %d, POST binary stream to IP:%d/asr\n", port, port);
  if (!args.port_file.empty()) {
    FILE* f = fopen((args.port_file + ".tmp").c_str(), "w");
    if (f) {
      fprintf(f, "%d\n", port);
      fclose(f);
      rename((args.port_file + ".tmp").c_str(), args.port_file.c_str());
    }
  }
  std::atomic<int> live{0};
  while (!g_stop.load()) {
    const int fd = ::accept(g_listen_fd, nullptr, nullptr);
    if (fd < 0) {
      if (errno == EINTR) continue;
      break;
    }
    ++live;
    std::thread([fd, &live] {
      serve_connection(fd);
      --live;
    }).detach();
  }
  ::close(g_listen_fd);
  for (int i = 0; i < 300 && live.load() > 0; ++i) usleep(10000);  // let answers in flight finish (3 s at most)
  if (g_model) AX_WHISPER_Uninit(g_model);
  fprintf(stderr, "[I] server stopped\n");
  return 0;
}
