// Host-only stress of the C ABI pieces that never touch the GPU — the endpoint group / CHWBL router (router.cc), the BPE
// tokenizer and its incremental detokeniser (tokenizer.cc) and the JSON parser they share (hostutil.h) — built twice by
// tests/test_host_sanitizers.py: with AddressSanitizer + UndefinedBehaviorSanitizer (argv[1] = "asan") and with
// ThreadSanitizer (argv[1] = "tsan").  Exit code 0 and a silent sanitizer are the assertion.
//   host_sanitize asan <tokenizer.json>     tokenizer + JSON fuzz on one thread
//   host_sanitize tsan                       router under concurrent reconcile / pick / done
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200engine.h"
#include "../../kubeai_b200/csrc/errors.h"
#include "../../kubeai_b200/csrc/hostutil.h"

// the pieces of abi_ops.cu the host files link against
namespace b200 {
static thread_local char g_err[512];
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int require_device() { return B200_ERR_NO_DEVICE; }
int cuda_fail(const char*, int rc) { return rc; }
}  // namespace b200
extern "C" const char* b200_last_error(void) { return b200::g_err; }

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() {
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 7;
  rng_state ^= rng_state << 17;
  return static_cast<uint32_t>(rng_state >> 16);
}

#define CHECK(c)                                                        \
  do {                                                                  \
    if (!(c)) {                                                         \
      fprintf(stderr, "CHECK failed at line %d: %s\n", __LINE__, #c);   \
      exit(1);                                                          \
    }                                                                   \
  } while (0)

static bool valid_utf8(const std::string& s) {
  const size_t n = s.size();
  size_t i = 0;
  while (i < n) {
    const unsigned char c = static_cast<unsigned char>(s[i]);
    size_t need;
    if (c < 0x80) need = 0;
    else if (c >= 0xC2 && c <= 0xDF) need = 1;
    else if (c >= 0xE0 && c <= 0xEF) need = 2;
    else if (c >= 0xF0 && c <= 0xF4) need = 3;
    else return false;
    if (i + need >= n + (need == 0 ? 1 : 0) && need > 0) return false;
    for (size_t k = 1; k <= need; ++k)
      if ((static_cast<unsigned char>(s[i + k]) & 0xC0) != 0x80) return false;
    i += need + 1;
  }
  return true;
}

static int run_asan(const char* tok_path) {
  b200_tokenizer* t = nullptr;
  CHECK(b200_tokenizer_load(tok_path, &t) == 0);
  const int V = b200_tokenizer_vocab_size(t);
  CHECK(V > 256);
  std::vector<int32_t> ids(1 << 16);
  std::vector<char> text(1 << 18);
  // 1. encode arbitrary bytes (valid and invalid UTF-8, specials spelled in the text), decode them back
  const char* frags[] = {"hello", " world", "'s", "\n\n", "   ", "<|eot_id|>", "<|begin_of_text|>", "12345", "\xe4\xbd\xa0\xe5\xa5\xbd", "\xf0\x9f\x99\x82",
                         "\xc3", "\xe2\x82", "\xff\xfe", "\xed\xa0\x80", "\t", "\r\n", "\xc2\xa0", "'RE"};
  for (int it = 0; it < 4000; ++it) {
    std::string s;
    const int parts = static_cast<int>(rnd() % 12);
    for (int k = 0; k < parts; ++k) {
      if (rnd() % 4 == 0) s.push_back(static_cast<char>(rnd() & 0xFF));
      else s += frags[rnd() % (sizeof(frags) / sizeof(frags[0]))];
    }
    for (int special = 0; special < 2; ++special) {
      const int64_t n = b200_tokenizer_encode(t, s.data(), s.size(), special, ids.data(), ids.size());
      CHECK(n >= 0 && n <= static_cast<int64_t>(ids.size()));
      for (int64_t i = 0; i < n; ++i) CHECK(ids[i] >= 0 && ids[i] < V);
      const int64_t m = b200_tokenizer_decode(t, ids.data(), static_cast<size_t>(n), 0, text.data(), text.size());
      CHECK(m >= 0 && m < static_cast<int64_t>(text.size()));
      if (!special) CHECK(std::string(text.data(), static_cast<size_t>(m)) == s);   // byte-level BPE is lossless on any bytes
    }
    // tiny output buffers: the length is still reported, nothing is written past cap
    int32_t one[1];
    CHECK(b200_tokenizer_encode(t, s.data(), s.size(), 1, one, 1) >= 0);
    char two[2];
    CHECK(b200_tokenizer_decode(t, ids.data(), 1, 0, two, sizeof(two)) >= 0);
  }
  // 2. incremental detokenisation of random id soup (out-of-range ids included): every push is complete UTF-8
  for (int it = 0; it < 500; ++it) {
    b200_detok_stream* d = b200_tokenizer_stream_new(t, it & 1);
    std::string all;
    char buf[512];
    for (int k = 0; k < 80; ++k) {
      const int32_t id = (rnd() % 50 == 0) ? V + static_cast<int32_t>(rnd() % 5) : static_cast<int32_t>(rnd() % static_cast<uint32_t>(V));
      const int64_t n = b200_tokenizer_stream_push(d, id, buf, sizeof(buf));
      CHECK(n >= 0 && n < static_cast<int64_t>(sizeof(buf)));
      const std::string piece(buf, static_cast<size_t>(n));
      CHECK(valid_utf8(piece));
      all += piece;
    }
    const int64_t n = b200_tokenizer_stream_push(d, -1, buf, sizeof(buf));
    CHECK(n >= 0);
    b200_tokenizer_stream_free(d);
  }
  // 3. chat framing with odd inputs
  {
    const char* roles[] = {"system", "user", "", "assistant"};
    const char* contents[] = {"  \n", "hi <|eot_id|> there", "\xff", ""};
    CHECK(b200_tokenizer_chat_llama3(t, roles, contents, 4, 1, ids.data(), ids.size()) > 8);
    CHECK(b200_tokenizer_chat_llama3(t, nullptr, nullptr, 0, 0, ids.data(), ids.size()) == 1);
  }
  b200_tokenizer_destroy(t);
  // 4. truncated / mutated tokenizer.json: load must fail (or succeed) without touching freed or foreign memory
  {
    std::string file;
    FILE* f = fopen(tok_path, "rb");
    CHECK(f != nullptr);
    char buf[1 << 16];
    size_t k;
    while ((k = fread(buf, 1, sizeof(buf), f)) > 0) file.append(buf, k);
    fclose(f);
    const std::string tmp = std::string(tok_path) + ".mut";
    for (int it = 0; it < 60; ++it) {
      std::string m = file;
      if (it % 2 == 0) m.resize(rnd() % m.size());
      else
        for (int j = 0; j < 8; ++j) m[rnd() % m.size()] = static_cast<char>(rnd() & 0x7F);
      FILE* o = fopen(tmp.c_str(), "wb");
      CHECK(o != nullptr);
      fwrite(m.data(), 1, m.size(), o);
      fclose(o);
      b200_tokenizer* bad = nullptr;
      if (b200_tokenizer_load(tmp.c_str(), &bad) == 0) b200_tokenizer_destroy(bad);
    }
    remove(tmp.c_str());
  }
  // 5. the JSON parser on random mutations of a small document
  {
    const std::string doc = "{\"a\":[1,2.5,-3e10,true,false,null],\"s\":\"x\\u00e9\\ud83d\\ude42\\n\",\"o\":{\"k\":{\"deep\":[[[[]]]]}}}";
    for (int it = 0; it < 20000; ++it) {
      std::string m = doc;
      const int muts = 1 + static_cast<int>(rnd() % 4);
      for (int j = 0; j < muts; ++j) {
        const size_t p = rnd() % m.size();
        switch (rnd() % 3) {
          case 0: m[p] = static_cast<char>(rnd() & 0xFF); break;
          case 1: m.erase(p, 1 + rnd() % 3); break;
          default: m.insert(p, 1, "{}[]\",:\\u"[rnd() % 10]); break;
        }
        if (m.empty()) m = "0";
      }
      b200::JVal v;
      std::string err;
      b200::JParser(m.data(), m.size()).parse(&v, &err);
    }
  }
  printf("asan leg ok\n");
  return 0;
}

static int run_tsan() {
  b200_router* r = nullptr;
  CHECK(b200_router_create(64, &r) == 0);
  std::atomic<bool> stop{false};
  std::atomic<long> picks{0};
  // reconciler: the endpoint set keeps changing (pods come and go, group.go:108-137)
  std::thread reconciler([&] {
    const char* names[] = {"pod-a", "pod-b", "pod-c", "pod-d", "pod-e"};
    const char* addrs[] = {"gpu:0", "gpu:1", "gpu:2", "gpu:3", "gpu:4"};
    const char* adapters[] = {"lora1", nullptr, "lora1,lora2", nullptr, "lora2"};
    uint64_t s = 1;
    while (!stop) {
      s = s * 6364136223846793005ull + 1442695040888963407ull;
      const int n = 1 + static_cast<int>((s >> 33) % 5);
      CHECK(b200_router_set_endpoints(r, names, addrs, adapters, n) == 0);
      std::this_thread::yield();
    }
  });
  std::vector<std::thread> clients;
  for (int c = 0; c < 6; ++c) {
    clients.emplace_back([&, c] {
      char addr[64];
      char prefix[32];
      for (int i = 0; i < 4000; ++i) {
        snprintf(prefix, sizeof(prefix), "conversation-%d-%d", c, i % 17);
        uint64_t tok = 0;
        const int strategy = (i & 1) ? B200_LB_PREFIX_HASH : B200_LB_LEAST_LOAD;
        const char* adapter = (i % 5 == 0) ? "lora1" : "";
        const int rc = b200_router_pick(r, strategy, adapter, prefix, static_cast<int>(strlen(prefix)), 125, 2000, addr, sizeof(addr), &tok);
        if (rc == 0) {
          picks.fetch_add(1);
          int64_t a = 0, b = 0;
          b200_router_inflight(r, "pod-a", &a, &b);
          CHECK(b200_router_done(r, tok) == 0);
        }
      }
    });
  }
  for (auto& t : clients) t.join();
  stop = true;
  reconciler.join();
  char buf[1 << 15];
  CHECK(b200_router_metrics(r, buf, sizeof(buf)) > 0);
  b200_router_destroy(r);
  CHECK(picks.load() > 1000);
  printf("tsan leg ok (%ld picks)\n", picks.load());
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 3 && !strcmp(argv[1], "asan")) return run_asan(argv[2]);
  if (argc >= 2 && !strcmp(argv[1], "tsan")) return run_tsan();
  fprintf(stderr, "usage: host_sanitize asan <tokenizer.json> | tsan\n");
  return 2;
}
