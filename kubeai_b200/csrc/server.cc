// Serving shell: the reference's request path restated in C++ around the in-process engine.
//   internal/openaiserver/handler.go:20-49     route table under /openai (R1)
//   internal/modelproxy/handler.go:57-159      ServeHTTP / proxyHTTP: parse -> gauge -> pick -> serve, <=3 retries (R2, R5)
//   internal/modelproxy/request.go:45-63       error bodies {"error":"..."}\n, 5xx text hidden
//   internal/apiutils/request.go:64-225        ParseRequest: JSON body, model_adapter split, prefix (R3)
//   api/openai/v1/chat_completions.go:525-543  Prefix(n): first user message, first n runes (R4)
//   internal/metrics/metrics.go:16-27          kubeai_inference_requests_active gauge (R11)
// and the backend half the reference leaves to vLLM: tokenise + chat template, stream tokens as
// OpenAI SSE chunks in vLLM's framing (api/openai/v1/reference/example-requests.vllm.output:139-160;
// vllm/entrypoints/openai/chat_completion/serving.py) (K12, K13).
// The HTTP hop of proxyHTTP (handler.go:116-158) is replaced by b200_submit/b200_wait/b200_poll.
#include <arpa/inet.h>
#include <errno.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <signal.h>
#include <sys/socket.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <map>
#include <mutex>
#include <random>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200engine.h"
#include "errors.h"
#include "hostutil.h"
#include "tokenizer.h"

namespace b200 {

struct Server {
  std::string model;
  std::set<std::string> adapters;
  int strategy = B200_LB_LEAST_LOAD;
  int mean_load_pct = 125, replication = 256, prefix_chars = 100;
  int max_retries = 3;
  int default_max_tokens = 256;
  std::vector<b200_engine*> replicas;
  std::vector<std::string> addrs;  // "gpu:<i>"
  b200_router* router = nullptr;
  Tokenizer tok;
  int max_model_len = 2048;
  // optional real tokenizer (b200_server_set_tokenizer): a checkpoint's tokenizer.json instead of the synthetic byte/word ids
  const b200_tokenizer* real_tok = nullptr;

  std::mutex mmu;
  std::map<std::string, int64_t> active;  // kubeai_inference_requests_active{request_model=...}
  std::atomic<int64_t> requests_total{0}, retries_total{0}, errors_total{0};
  std::vector<std::unique_ptr<std::atomic<int>>> faults;  // fault injection: fail the next N submits on replica i

  // HTTP listener
  int listen_fd = -1;
  std::thread acceptor;
  std::atomic<bool> stopping{false};
  std::atomic<int> live_conns{0};
  std::mutex conn_mu;
  std::set<int> conn_fds;          // open client sockets: shut down by the destructor so no thread outlives the Server
  static constexpr int kMaxConns = 4096;
  static constexpr int kIdleTimeoutS = 60;  // keep-alive read timeout (SO_RCVTIMEO)

  // replicas whose engine reported a fatal error are taken out of the endpoint set, as reconcileEndpoints drops
  // endpoints that vanished (internal/loadbalancer/group.go:119-131)
  std::mutex dead_mu;
  std::vector<char> dead;
  std::string adapters_csv;
  void drop_replica(int i);

  ~Server() {
    stopping = true;
    if (listen_fd >= 0) {
      shutdown(listen_fd, SHUT_RDWR);
      close(listen_fd);
    }
    if (acceptor.joinable()) acceptor.join();
    {
      // wake every connection thread blocked in recv()/send(); each closes its own fd and leaves
      std::lock_guard<std::mutex> lk(conn_mu);
      for (int fd : conn_fds) shutdown(fd, SHUT_RDWR);
    }
    while (live_conns.load() > 0) usleep(1000);
    if (router) b200_router_destroy(router);
  }
};

void Server::drop_replica(int i) {
  std::lock_guard<std::mutex> lk(dead_mu);
  if (i < 0 || i >= static_cast<int>(dead.size()) || dead[i]) return;
  dead[i] = 1;
  std::vector<std::string> names;
  std::vector<const char*> np, ap, dp;
  for (size_t k = 0; k < replicas.size(); ++k) {
    if (dead[k]) continue;
    names.push_back("gpu-" + std::to_string(k));
  }
  size_t j = 0;
  for (size_t k = 0; k < replicas.size(); ++k) {
    if (dead[k]) continue;
    np.push_back(names[j++].c_str());
    ap.push_back(addrs[k].c_str());
    dp.push_back(adapters_csv.c_str());
  }
  b200_router_set_endpoints(router, np.data(), ap.data(), dp.data(), static_cast<int32_t>(np.size()));
}

namespace {

std::string hex_id(int n) {
  static thread_local std::mt19937_64 rng{std::random_device{}()};
  static const char* h = "0123456789abcdef";
  std::string s;
  for (int i = 0; i < n; ++i) s += h[rng() & 15];
  return s;
}

struct Writer {
  const b200_response_writer* w;
  bool begun = false;
  bool gone = false;
  void begin(int status, const char* ctype) {
    if (begun) return;
    begun = true;
    if (w->begin && w->begin(w->ud, status, ctype)) gone = true;
  }
  void write(const std::string& s) {
    if (gone || !w->write) return;
    if (w->write(w->ud, s.data(), s.size())) gone = true;
  }
};

// modelproxy/request.go:45-63 sendErrorResponse
int send_error(Writer& w, int status, const std::string& msg) {
  static const std::map<int, const char*> text = {{500, "Internal Server Error"}, {502, "Bad Gateway"},
                                                  {503, "Service Unavailable"}, {504, "Gateway Timeout"}};
  std::string m = msg;
  if (status >= 500) {
    auto it = text.find(status);
    m = it == text.end() ? "Internal Server Error" : it->second;
  }
  w.begin(status, "application/json");
  w.write("{\"error\":" + json_str(m) + "}\n");
  return status;
}

struct ParsedRequest {
  bool chat = true;
  std::string requested_model, model, adapter, prefix;
  std::vector<std::pair<std::string, std::string>> messages;
  std::vector<int32_t> prompt_ids;  // completions with token-id prompt
  std::string prompt_text;
  bool prompt_is_ids = false;
  int max_tokens = -1;
  bool stream = false, include_usage = false, ignore_eos = false;
  double temperature = 0.0;
  bool has_temperature = false;
  std::vector<int32_t> stop_ids;
};

std::string content_text(const JVal& c) {
  if (c.type == JVal::Str) return c.str;
  std::string s;
  if (c.type == JVal::Arr)
    for (auto& part : c.arr) {
      const JVal* t = part.get("text");
      if (t && t->type == JVal::Str) s += t->str;
    }
  return s;  // content:null -> "" (the reference nil-derefs here, chat_completions.go:532; not copied)
}

// apiutils.ParseRequest for the JSON branch.  Returns 0 or an HTTP status with *err set.
int parse_request(Server& sv, const std::string& path, const std::string& ctype, const char* body, size_t len,
                  ParsedRequest* pr, std::string* err) {
  std::string media = ctype.substr(0, ctype.find(';'));
  while (!media.empty() && media.back() == ' ') media.pop_back();
  for (auto& ch : media) ch = static_cast<char>(tolower(ch));
  if (media == "multipart/form-data") {
    *err = "bad request: reading multipart form data: not supported by this engine (speech routes are out of scope)";
    return 400;
  }
  if (path == "/v1/chat/completions") pr->chat = true;
  else if (path == "/v1/completions") pr->chat = false;
  else {
    *err = "bad request: reading model from body: unknown path: \"" + path + "\"";
    return 400;
  }
  JVal root;
  std::string jerr;
  if (!JParser(body, len).parse(&root, &jerr) || root.type != JVal::Obj) {
    *err = "bad request: reading model from body: decoding: " + (jerr.empty() ? std::string("expected JSON object") : jerr);
    return 400;
  }
  const JVal* m = root.get("model");
  if (!m || m->type != JVal::Str || m->str.empty()) {
    *err = "bad request: reading model from body: missing 'model' field";
    return 400;
  }
  pr->requested_model = m->str;
  // apiutils/model.go:23-29 SplitModelAdapter: first "_"
  size_t us = m->str.find('_');
  pr->model = m->str.substr(0, us);
  pr->adapter = us == std::string::npos ? "" : m->str.substr(us + 1);
  // LookupModel (modelclient/client.go:27-64): nil -> ErrModelNotFound
  if (pr->model != sv.model || (!pr->adapter.empty() && !sv.adapters.count(pr->adapter))) {
    *err = "model not found: \"" + pr->requested_model + "\"";
    return 404;
  }
  if (pr->chat) {
    const JVal* msgs = root.get("messages");
    if (msgs && msgs->type == JVal::Arr) {
      for (auto& mm : msgs->arr) {
        const JVal* role = mm.get("role");
        const JVal* content = mm.get("content");
        pr->messages.emplace_back(role && role->type == JVal::Str ? role->str : "", content ? content_text(*content) : "");
      }
    }
    for (auto& mm : pr->messages)
      if (mm.first == "user") {
        pr->prefix = mm.second;
        break;
      }
  } else {
    const JVal* p = root.get("prompt");
    if (p && p->type == JVal::Str) {
      pr->prompt_text = p->str;
      pr->prefix = p->str;
    } else if (p && p->type == JVal::Arr && !p->arr.empty()) {
      const JVal& first = p->arr[0];
      if (first.type == JVal::Str) {
        pr->prompt_text = first.str;
        pr->prefix = first.str;
      } else if (first.type == JVal::Num) {
        pr->prompt_is_ids = true;
        for (auto& e : p->arr) pr->prompt_ids.push_back(static_cast<int32_t>(e.num));
      } else if (first.type == JVal::Arr) {
        pr->prompt_is_ids = true;
        for (auto& e : first.arr) pr->prompt_ids.push_back(static_cast<int32_t>(e.num));
      }
    }
  }
  // r.Prefix only under PrefixHash (apiutils/request.go:219-223)
  pr->prefix = sv.strategy == B200_LB_PREFIX_HASH ? first_n_runes(pr->prefix, sv.prefix_chars) : "";  // firstNChars (utils.go:5-8)
  if (const JVal* v = root.get("max_completion_tokens"); v && v->type == JVal::Num) pr->max_tokens = static_cast<int>(v->num);
  if (const JVal* v = root.get("max_tokens"); v && v->type == JVal::Num) pr->max_tokens = static_cast<int>(v->num);
  if (const JVal* v = root.get("stream"); v && v->type == JVal::Bool) pr->stream = v->b;
  if (const JVal* so = root.get("stream_options"))
    if (const JVal* iu = so->get("include_usage"); iu && iu->type == JVal::Bool) pr->include_usage = iu->b;
  if (const JVal* v = root.get("ignore_eos"); v && v->type == JVal::Bool) pr->ignore_eos = v->b;
  if (const JVal* v = root.get("temperature"); v && v->type == JVal::Num) {
    pr->temperature = v->num;
    pr->has_temperature = true;
  }
  if (const JVal* v = root.get("stop_token_ids"); v && v->type == JVal::Arr)
    for (auto& e : v->arr)
      if (e.type == JVal::Num) pr->stop_ids.push_back(static_cast<int32_t>(e.num));
  // Fields of the reference's request schema (api/openai/v1/chat_completions.go:361-470, completions.go:19-130) whose
  // semantics this engine does not implement are refused, never silently dropped.
  auto num_of = [&](const char* k, double def) { const JVal* v = root.get(k); return v && v->type == JVal::Num ? v->num : def; };
  auto truthy = [&](const char* k) {
    const JVal* v = root.get(k);
    return v && ((v->type == JVal::Bool && v->b) || (v->type == JVal::Num && v->num > 0));
  };
  auto nonempty = [&](const char* k) {
    const JVal* v = root.get(k);
    return v && ((v->type == JVal::Str && !v->str.empty()) || (v->type == JVal::Arr && !v->arr.empty()) ||
                 (v->type == JVal::Obj && !v->obj.empty()));
  };
  const char* bad = nullptr;
  if (num_of("n", 1) > 1) bad = "n > 1";
  else if (num_of("best_of", 1) > 1) bad = "best_of > 1";
  else if (truthy("logprobs") || truthy("top_logprobs") || truthy("prompt_logprobs")) bad = "logprobs";
  else if (nonempty("stop")) bad = "stop strings (use stop_token_ids)";
  else if (truthy("echo")) bad = "echo";
  else if (nonempty("suffix")) bad = "suffix";
  else if (num_of("presence_penalty", 0) != 0 || num_of("frequency_penalty", 0) != 0 || num_of("repetition_penalty", 1) != 1) bad = "sampling penalties";
  else if (nonempty("logit_bias")) bad = "logit_bias";
  else if (nonempty("tools") || nonempty("functions")) bad = "tool calling";
  else if (truthy("min_tokens")) bad = "min_tokens";
  else if (const JVal* rf = root.get("response_format"); rf && rf->type == JVal::Obj) {
    const JVal* ty = rf->get("type");
    if (ty && ty->type == JVal::Str && ty->str != "text") bad = "response_format other than text";
  }
  if (bad) {
    *err = std::string("bad request: ") + bad + " is not supported by this engine";
    return 400;
  }
  return 0;
}

std::string usage_json(const b200_usage& u) {
  char b[256];
  snprintf(b, sizeof(b),
           "{\"prompt_tokens\":%d,\"total_tokens\":%d,\"completion_tokens\":%d,\"prompt_tokens_details\":{\"cached_tokens\":%d}}",
           u.prompt_tokens, u.prompt_tokens + u.completion_tokens, u.completion_tokens, u.cached_tokens);
  return b;
}

const char* finish_str(int code) { return code == B200_FINISH_STOP ? "stop" : code == B200_FINISH_LENGTH ? "length" : "abort"; }

}  // namespace

// modelproxy.Handler.ServeHTTP + proxyHTTP with the engine in place of the reverse proxy.
// K13: the ids the backend would see for this request.  With a real tokenizer attached: the Llama-3 chat framing for chat
// requests, <|begin_of_text|> + the encoded text for completions (what the reference's vLLM pod does with the checkpoint's
// tokenizer.json and chat_template); otherwise the synthetic tokenizer and ChatML of hostutil.h.
static void render_prompt(const Server& sv, const ParsedRequest& pr, std::vector<int32_t>* ids) {
  if (pr.prompt_is_ids && !pr.chat) {
    *ids = pr.prompt_ids;
  } else if (sv.real_tok) {
    std::vector<int32_t> buf(1024);
    int64_t n = 0;
    if (pr.chat) {
      std::vector<const char*> roles, contents;
      for (auto& m : pr.messages) { roles.push_back(m.first.c_str()); contents.push_back(m.second.c_str()); }
      const bool gen = pr.messages.empty() || pr.messages.back().first != "assistant";
      n = b200_tokenizer_chat_llama3(sv.real_tok, roles.data(), contents.data(), static_cast<int>(roles.size()), gen ? 1 : 0, buf.data(), buf.size());
      if (n > static_cast<int64_t>(buf.size())) {
        buf.resize(static_cast<size_t>(n));
        n = b200_tokenizer_chat_llama3(sv.real_tok, roles.data(), contents.data(), static_cast<int>(roles.size()), gen ? 1 : 0, buf.data(), buf.size());
      }
      if (n > 0) ids->assign(buf.begin(), buf.begin() + n);
    } else {
      if (tokenizer_bos(sv.real_tok) >= 0) ids->push_back(tokenizer_bos(sv.real_tok));
      n = b200_tokenizer_encode(sv.real_tok, pr.prompt_text.data(), pr.prompt_text.size(), 1, buf.data(), buf.size());
      if (n > static_cast<int64_t>(buf.size())) {
        buf.resize(static_cast<size_t>(n));
        n = b200_tokenizer_encode(sv.real_tok, pr.prompt_text.data(), pr.prompt_text.size(), 1, buf.data(), buf.size());
      }
      if (n > 0) ids->insert(ids->end(), buf.begin(), buf.begin() + n);
    }
  } else if (pr.chat) {
    sv.tok.chat_prompt(pr.messages, ids);
  } else {
    sv.tok.encode(pr.prompt_text, ids);
  }
  if (ids->empty()) ids->push_back(sv.real_tok && tokenizer_bos(sv.real_tok) >= 0 ? tokenizer_bos(sv.real_tok) : sv.tok.im_start());
}

static int serve_inference(Server& sv, const std::string& path, const std::string& ctype, const char* body,
                           size_t len, Writer& w) {
  ParsedRequest pr;
  std::string err;
  if (int st = parse_request(sv, path, ctype, body, len, &pr, &err)) return send_error(w, st, err);
  if (pr.has_temperature && pr.temperature >= 1e-5) {
    // greedy only (north star); refuse rather than silently change semantics
    return send_error(w, 400, "bad request: only greedy sampling is implemented: set \"temperature\": 0");
  }
  // tokenise (K13)
  std::vector<int32_t> ids;
  render_prompt(sv, pr, &ids);
  for (auto t : ids)
    if (t < 0 || t >= sv.tok.vocab) return send_error(w, 400, "bad request: prompt token id out of range");
  int max_tokens = pr.max_tokens > 0 ? pr.max_tokens : sv.default_max_tokens;
  if (static_cast<int>(ids.size()) + 1 > sv.max_model_len)
    return send_error(w, 400, "bad request: prompt is longer than the model's context length");
  max_tokens = std::min(max_tokens, sv.max_model_len - static_cast<int>(ids.size()));

  // metrics.InferenceRequestsActive +1 / defer -1 (handler.go:76-81)
  {
    std::lock_guard<std::mutex> lk(sv.mmu);
    ++sv.active[pr.requested_model];
  }
  sv.requests_total.fetch_add(1);
  struct Dec {
    Server& s;
    std::string m;
    ~Dec() {
      std::lock_guard<std::mutex> lk(s.mmu);
      --s.active[m];
    }
  } dec{sv, pr.requested_model};

  const std::string id = (pr.chat ? "chatcmpl-" : "cmpl-") + hex_id(32);
  const long created = static_cast<long>(time(nullptr));
  const std::string head = "{\"id\":\"" + id + "\",\"object\":\"" + (pr.chat ? "chat.completion.chunk" : "text_completion") +
                           "\",\"created\":" + std::to_string(created) + ",\"model\":" + json_str(pr.requested_model);

  // proxyHTTP (handler.go:96-159): pick an endpoint, serve, retry with a fresh pick on failure;
  // the in-flight count of every attempt is released when the request ends.
  std::vector<uint64_t> tokens_held;
  auto release_all = [&] {
    for (auto t : tokens_held) b200_router_done(sv.router, t);
    tokens_held.clear();
  };
  for (int attempt = 0;; ++attempt) {
    char addr[64];
    uint64_t ep_token = 0;
    int rc = b200_router_pick(sv.router, sv.strategy, pr.adapter.c_str(), pr.prefix.data(), static_cast<int>(pr.prefix.size()),
                              sv.mean_load_pct, 30ll * 1000 * 1000, addr, sizeof(addr), &ep_token);
    if (rc) {
      release_all();
      sv.errors_total.fetch_add(1);
      return send_error(w, 504, "request timeout while finding host");
    }
    tokens_held.push_back(ep_token);
    const int replica = atoi(addr + 4);  // "gpu:<i>"
    b200_engine* eng = sv.replicas[replica];
    b200_sampling sp;
    sp.max_tokens = max_tokens;
    sp.temperature = 0.f;
    sp.ignore_eos = pr.ignore_eos ? 1 : 0;
    std::vector<int32_t> stops = pr.stop_ids;
    if (sv.real_tok && !pr.ignore_eos) {   // the checkpoint's end-of-turn / end-of-text tokens end the generation
      for (const char* name : {"<|eot_id|>", "<|end_of_text|>"}) {
        const int32_t t = b200_tokenizer_token_id(sv.real_tok, name);
        if (t >= 0) stops.push_back(t);
      }
    }
    sp.num_stop_ids = static_cast<int>(stops.size());
    sp.stop_ids = stops.empty() ? nullptr : stops.data();
    uint64_t rid = 0;
    bool failed = false;
    if (sv.faults[replica]->load() < 0) failed = true;  // permanent fault: stands for an engine in the failed state
    else if (sv.faults[replica]->load() > 0 && sv.faults[replica]->fetch_sub(1) > 0) failed = true;
    if (!failed && b200_submit(eng, ids.data(), static_cast<int>(ids.size()), &sp, &rid)) failed = true;

    std::vector<int32_t> all;
    DetokStream detok;
    detok.tok = sv.real_tok;
    int fin = 0;
    b200_usage usage{};
    bool sent_any = false;
    if (!failed) {
      int32_t buf[256];
      while (!fin) {
        b200_wait(eng, rid, 200000);
        int n = 0;
        if (b200_poll(eng, rid, buf, 256, &n, &fin, &usage)) { failed = true; break; }
        if (fin == B200_FINISH_ERROR && !sent_any) { failed = true; break; }
        if (pr.stream && n > 0) {
          if (!sent_any) {
            w.begin(200, "text/event-stream; charset=utf-8");
            if (pr.chat)
              w.write("data: " + head + ",\"choices\":[{\"index\":0,\"delta\":{\"role\":\"assistant\",\"content\":\"\"},\"logprobs\":null,\"finish_reason\":null}]}\n\n");
            sent_any = true;
          }
          for (int i = 0; i < n; ++i) {
            const bool last = fin && i == n - 1;
            const std::string fr = last ? std::string("\"") + finish_str(fin) + "\"" : "null";
            // one chunk per token; with a real tokenizer a chunk carries the text that token completed (possibly none: a
            // token may end inside a UTF-8 sequence), and the last chunk also what was still held
            const std::string piece = json_str(sv.real_tok ? detok.push(buf[i]) + (last ? detok.flush() : std::string()) : sv.tok.piece(buf[i]));
            if (pr.chat)
              w.write("data: " + head + ",\"choices\":[{\"index\":0,\"delta\":{\"content\":" + piece + "},\"logprobs\":null,\"finish_reason\":" + fr +
                      (last ? ",\"stop_reason\":null" : "") + "}]}\n\n");
            else
              w.write("data: " + head + ",\"choices\":[{\"index\":0,\"text\":" + piece + ",\"logprobs\":null,\"finish_reason\":" + fr + ",\"stop_reason\":null}],\"usage\":null}\n\n");
          }
        }
        all.insert(all.end(), buf, buf + n);
        if (w.gone) {  // client disconnected: abort the sequence and free its KV blocks
          b200_abort(eng, rid);
          break;
        }
      }
      b200_release(eng, rid);
    }
    if (failed) {
      if (b200_engine_is_failed(eng) || sv.faults[replica]->load() < 0) sv.drop_replica(replica);
      if (attempt < sv.max_retries) {
        sv.retries_total.fetch_add(1);
        continue;
      }
      release_all();
      sv.errors_total.fetch_add(1);
      return send_error(w, 502, "proxy: exceeded retries");
    }
    release_all();
    if (w.gone) return 499;
    if (pr.stream) {
      if (!sent_any) w.begin(200, "text/event-stream; charset=utf-8");
      if (pr.include_usage) w.write("data: " + head + ",\"choices\":[],\"usage\":" + usage_json(usage) + "}\n\n");
      w.write("data: [DONE]\n\n");
    } else {
      std::string text;
      if (sv.real_tok) {
        DetokStream whole;
        whole.tok = sv.real_tok;
        for (auto t : all) text += whole.push(t);
        text += whole.flush();
      } else {
        for (auto t : all) text += sv.tok.piece(t);
      }
      std::string out = "{\"id\":\"" + id + "\",\"object\":\"" + (pr.chat ? "chat.completion" : "text_completion") +
                        "\",\"created\":" + std::to_string(created) + ",\"model\":" + json_str(pr.requested_model) + ",\"choices\":[{\"index\":0,";
      if (pr.chat) out += "\"message\":{\"role\":\"assistant\",\"content\":" + json_str(text) + "},";
      else out += "\"text\":" + json_str(text) + ",";
      out += std::string("\"logprobs\":null,\"finish_reason\":\"") + finish_str(fin) + "\",\"stop_reason\":null}],\"usage\":" + usage_json(usage) + "}";
      w.begin(200, "application/json");
      w.write(out);
    }
    return 200;
  }
}

static std::string metrics_text(Server& sv) {
  std::string o;
  o += "# HELP kubeai_inference_requests_active The number of active requests by model\n# TYPE kubeai_inference_requests_active gauge\n";
  {
    std::lock_guard<std::mutex> lk(sv.mmu);
    for (auto& kv : sv.active)
      o += "kubeai_inference_requests_active{request_model=" + json_str(kv.first) + ",request_type=\"http\"} " + std::to_string(kv.second) + "\n";
  }
  {
    const int64_t n = b200_router_metrics(sv.router, nullptr, 0);
    if (n > 0) {
      std::string rm(static_cast<size_t>(n) + 1, '\0');
      b200_router_metrics(sv.router, &rm[0], rm.size());
      rm.resize(static_cast<size_t>(n));
      o += rm;
    }
  }
  o += "# TYPE b200_requests_total counter\nb200_requests_total " + std::to_string(sv.requests_total.load()) + "\n";
  o += "# TYPE b200_request_retries_total counter\nb200_request_retries_total " + std::to_string(sv.retries_total.load()) + "\n";
  for (size_t i = 0; i < sv.replicas.size(); ++i) {
    b200_stats st;
    if (b200_stats_get(sv.replicas[i], &st)) continue;
    char b[1024];
    snprintf(b, sizeof(b),
             "b200_engine_steps_total{replica=\"%zu\"} %lld\nb200_engine_running{replica=\"%zu\"} %d\nb200_engine_waiting{replica=\"%zu\"} %d\n"
             "b200_engine_kv_blocks_free{replica=\"%zu\"} %lld\nb200_engine_kv_blocks_total{replica=\"%zu\"} %lld\n"
             "b200_engine_prompt_tokens_total{replica=\"%zu\"} %lld\nb200_engine_cached_prompt_tokens_total{replica=\"%zu\"} %lld\n"
             "b200_engine_generated_tokens_total{replica=\"%zu\"} %lld\nb200_engine_preemptions_total{replica=\"%zu\"} %lld\n"
             "b200_engine_last_step_device_us{replica=\"%zu\"} %.1f\n",
             i, (long long)st.steps, i, st.running, i, st.waiting, i, (long long)st.kv_blocks_free, i, (long long)st.kv_blocks_total, i,
             (long long)st.prompt_tokens, i, (long long)st.cached_prompt_tokens, i, (long long)st.generated_tokens, i,
             (long long)st.preemptions, i, st.last_step_device_us);
    o += b;
  }
  return o;
}

// openaiserver.NewHandler route table (handler.go:20-49) + /metrics + /healthz.
static int handle(Server& sv, const std::string& method, const std::string& full_path, const std::string& ctype,
                  const char* body, size_t len, Writer& w) {
  std::string path = full_path.substr(0, full_path.find('?'));
  if (path == "/healthz" || path == "/readyz") {
    w.begin(200, "text/plain");
    w.write("ok\n");
    return 200;
  }
  if (path == "/metrics") {
    w.begin(200, "text/plain; version=0.0.4");
    w.write(metrics_text(sv));
    return 200;
  }
  if (path.rfind("/openai/", 0) != 0) {
    w.begin(404, "text/plain");
    w.write("404 page not found\n");
    return 404;
  }
  path = path.substr(7);  // http.StripPrefix("/openai", ...)
  if (path == "/v1/models" && method == "GET") {
    // openaiserver/models.go:13-77 (static: one model + adapters)
    std::string o = "{\"object\":\"list\",\"data\":[{\"id\":" + json_str(sv.model) + ",\"object\":\"model\",\"created\":0,\"owned_by\":\"kubeai-b200\",\"features\":[\"TextGeneration\"]}";
    for (auto& a : sv.adapters)
      o += ",{\"id\":" + json_str(sv.model + "_" + a) + ",\"object\":\"model\",\"created\":0,\"owned_by\":\"kubeai-b200\",\"features\":[\"TextGeneration\"]}";
    o += "]}";
    w.begin(200, "application/json");
    w.write(o);
    return 200;
  }
  // "X-Proxy: lingo" (handler.go:60) is added by the HTTP layer below / by the Go shim
  if (path == "/v1/chat/completions" || path == "/v1/completions") return serve_inference(sv, path, ctype, body, len, w);
  if (path == "/v1/embeddings" || path == "/v1/rerank" || path == "/v1/audio/transcriptions")
    return send_error(w, 404, "model not found: this engine serves TextGeneration only");
  w.begin(404, "text/plain");
  w.write("404 page not found\n");
  return 404;
}

// ------------------------------------------------------------------ minimal HTTP/1.1 front (thread per connection)
namespace {

struct ConnWriter {
  int fd;
  bool chunked = false;
  bool head_sent = false;
  bool failed = false;
  static bool send_all(int fd, const char* p, size_t n) {
    while (n) {
      ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
      if (k <= 0) {
        if (k < 0 && errno == EINTR) continue;
        return false;
      }
      p += k;
      n -= static_cast<size_t>(k);
    }
    return true;
  }
};

int conn_begin(void* ud, int status, const char* ctype) {
  ConnWriter* c = static_cast<ConnWriter*>(ud);
  if (status == 0) return 0;
  static const std::map<int, const char*> text = {{200, "OK"}, {400, "Bad Request"}, {404, "Not Found"}, {499, "Client Closed Request"},
                                                  {500, "Internal Server Error"}, {502, "Bad Gateway"}, {504, "Gateway Timeout"}};
  auto it = text.find(status);
  char h[512];
  c->chunked = true;
  int n = snprintf(h, sizeof(h), "HTTP/1.1 %d %s\r\nContent-Type: %s\r\nX-Proxy: lingo\r\nCache-Control: no-cache\r\nTransfer-Encoding: chunked\r\nConnection: keep-alive\r\n\r\n",
                   status, it == text.end() ? "Status" : it->second, ctype);
  c->head_sent = true;
  if (!ConnWriter::send_all(c->fd, h, static_cast<size_t>(n))) c->failed = true;
  return c->failed ? 1 : 0;
}

int conn_write(void* ud, const char* data, size_t len) {
  ConnWriter* c = static_cast<ConnWriter*>(ud);
  if (c->failed || len == 0) return c->failed ? 1 : 0;
  char h[32];
  int n = snprintf(h, sizeof(h), "%zx\r\n", len);
  std::string frame(h, static_cast<size_t>(n));
  frame.append(data, len);
  frame += "\r\n";
  if (!ConnWriter::send_all(c->fd, frame.data(), frame.size())) c->failed = true;
  return c->failed ? 1 : 0;
}

void serve_conn(Server* sv, int fd) {
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
  timeval tv{Server::kIdleTimeoutS, 0};
  setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
  setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
  std::string buf;
  char tmp[16384];
  for (;;) {
    // read headers
    size_t hend;
    while ((hend = buf.find("\r\n\r\n")) == std::string::npos) {
      ssize_t k = ::recv(fd, tmp, sizeof(tmp), 0);
      if (k <= 0) goto done;
      buf.append(tmp, static_cast<size_t>(k));
      if (buf.size() > (1 << 20)) goto done;
    }
    {
      std::string headers = buf.substr(0, hend);
      size_t body_off = hend + 4;
      size_t le = headers.find("\r\n");
      std::string reqline = headers.substr(0, le);
      std::string method = reqline.substr(0, reqline.find(' '));
      size_t p1 = reqline.find(' '), p2 = reqline.rfind(' ');
      std::string path = p1 != std::string::npos && p2 > p1 ? reqline.substr(p1 + 1, p2 - p1 - 1) : "/";
      size_t clen = 0;
      std::string ctype;
      bool close_after = false;
      size_t pos = le == std::string::npos ? headers.size() : le + 2;
      while (pos < headers.size()) {
        size_t e = headers.find("\r\n", pos);
        if (e == std::string::npos) e = headers.size();
        std::string line = headers.substr(pos, e - pos);
        pos = e + 2;
        size_t c = line.find(':');
        if (c == std::string::npos) continue;
        std::string k = line.substr(0, c), v = line.substr(c + 1);
        while (!v.empty() && v.front() == ' ') v.erase(v.begin());
        for (auto& ch : k) ch = static_cast<char>(tolower(ch));
        if (k == "content-length") clen = static_cast<size_t>(strtoull(v.c_str(), nullptr, 10));
        else if (k == "content-type") ctype = v;
        else if (k == "connection") { for (auto& ch : v) ch = static_cast<char>(tolower(ch)); close_after = v == "close"; }
      }
      if (clen > (64u << 20)) goto done;
      while (buf.size() < body_off + clen) {
        ssize_t k = ::recv(fd, tmp, sizeof(tmp), 0);
        if (k <= 0) goto done;
        buf.append(tmp, static_cast<size_t>(k));
      }
      ConnWriter cw{fd};
      b200_response_writer rw{&cw, conn_begin, conn_write};
      Writer w{&rw};
      handle(*sv, method, path, ctype, buf.data() + body_off, clen, w);
      if (cw.head_sent && !cw.failed) ConnWriter::send_all(fd, "0\r\n\r\n", 5);
      buf.erase(0, body_off + clen);
      if (cw.failed || close_after || sv->stopping) goto done;
    }
  }
done:
  {
    std::lock_guard<std::mutex> lk(sv->conn_mu);
    sv->conn_fds.erase(fd);
  }
  close(fd);
  sv->live_conns.fetch_sub(1);  // last touch of *sv: the destructor waits for this count to reach zero
}

}  // namespace
}  // namespace b200

using namespace b200;

struct b200_server {
  Server impl;
};

extern "C" {

int b200_server_create(const b200_server_config* cfg, b200_engine* const* replicas, int32_t n, b200_server** out) {
  if (!cfg || !replicas || n <= 0 || !out || !cfg->model) { set_error("b200_server_create: bad arguments"); return B200_ERR_INVALID; }
  b200_server* s = new (std::nothrow) b200_server();
  if (!s) { set_error("host OOM"); return B200_ERR_OOM; }
  Server& sv = s->impl;
  sv.model = cfg->model;
  if (cfg->adapters) {
    std::string a(cfg->adapters);
    size_t p = 0;
    while (p <= a.size()) {
      size_t q = a.find(',', p);
      if (q == std::string::npos) q = a.size();
      if (q > p) sv.adapters.insert(a.substr(p, q - p));
      p = q + 1;
    }
  }
  sv.strategy = cfg->strategy;
  // api/k8s/v1/model_types.go:173-209 defaults
  sv.mean_load_pct = cfg->mean_load_pct > 0 ? cfg->mean_load_pct : 125;
  sv.replication = cfg->replication > 0 ? cfg->replication : 256;
  sv.prefix_chars = cfg->prefix_char_length > 0 ? cfg->prefix_char_length : 100;
  sv.max_retries = cfg->max_retries >= 0 ? cfg->max_retries : 3;
  sv.default_max_tokens = cfg->default_max_tokens > 0 ? cfg->default_max_tokens : 256;
  sv.tok.vocab = cfg->vocab > 258 ? cfg->vocab : 128256;
  sv.max_model_len = cfg->max_model_len > 0 ? cfg->max_model_len : 2048;
  if (b200_router_create(sv.replication, &sv.router)) { delete s; return B200_ERR_INVALID; }
  std::vector<std::string> names;
  std::string ad = cfg->adapters ? cfg->adapters : "";
  sv.adapters_csv = ad;
  sv.dead.assign(static_cast<size_t>(n), 0);
  for (int i = 0; i < n; ++i) {
    sv.replicas.push_back(replicas[i]);
    names.push_back("gpu-" + std::to_string(i));
    sv.addrs.push_back("gpu:" + std::to_string(i));
    sv.faults.emplace_back(new std::atomic<int>(0));
  }
  std::vector<const char*> np, ap, dp;
  for (int i = 0; i < n; ++i) { np.push_back(names[i].c_str()); ap.push_back(sv.addrs[i].c_str()); dp.push_back(ad.c_str()); }
  if (int rc = b200_router_set_endpoints(sv.router, np.data(), ap.data(), dp.data(), n)) { delete s; return rc; }
  *out = s;
  return 0;
}

void b200_server_destroy(b200_server* s) { delete s; }

int b200_server_handle(b200_server* s, const char* method, const char* path, const char* content_type, const char* body,
                       size_t body_len, const b200_response_writer* writer) {
  if (!s || !method || !path || !writer) { set_error("b200_server_handle: bad arguments"); return B200_ERR_INVALID; }
  Writer w{writer};
  return handle(s->impl, method, path, content_type ? content_type : "", body ? body : "", body ? body_len : 0, w);
}

/* apiutils.ParseRequest on its own (for parity tests against internal/apiutils/request_test.go and the Prefix tables):
 * returns 0 or the HTTP status of the error; out_json gets {"model","adapter","requested_model","prefix"} or {"error"}. */
int b200_server_parse_request(b200_server* s, const char* path, const char* content_type, const char* body, size_t body_len,
                              int32_t prefix_chars, char* out_json, size_t cap) {
  if (!s || !path || !out_json || cap == 0) { set_error("b200_server_parse_request: bad arguments"); return B200_ERR_INVALID; }
  Server& sv = s->impl;
  ParsedRequest pr;
  std::string err;
  const int saved = sv.prefix_chars;
  if (prefix_chars >= 0) sv.prefix_chars = prefix_chars;
  std::string p(path);
  if (p.rfind("/openai", 0) == 0) p = p.substr(7);
  const int st = parse_request(sv, p, content_type ? content_type : "", body ? body : "", body ? body_len : 0, &pr, &err);
  sv.prefix_chars = saved;
  std::string o = st ? "{\"error\":" + json_str(err) + "}"
                     : "{\"model\":" + json_str(pr.model) + ",\"adapter\":" + json_str(pr.adapter) + ",\"requested_model\":" +
                           json_str(pr.requested_model) + ",\"prefix\":" + json_str(pr.prefix) + "}";
  snprintf(out_json, cap, "%s", o.c_str());
  return st;
}

int b200_server_listen(b200_server* s, const char* host, int32_t port, int32_t* bound_port) {
  if (!s) { set_error("null server"); return B200_ERR_INVALID; }
  Server& sv = s->impl;
  int fd = socket(AF_INET, SOCK_STREAM, 0);
  if (fd < 0) { set_error("socket: %s", strerror(errno)); return B200_ERR_INVALID; }
  int one = 1;
  setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  sockaddr_in a{};
  a.sin_family = AF_INET;
  a.sin_port = htons(static_cast<uint16_t>(port));
  inet_pton(AF_INET, host && *host ? host : "127.0.0.1", &a.sin_addr);
  if (bind(fd, reinterpret_cast<sockaddr*>(&a), sizeof(a)) || listen(fd, 1024)) {
    set_error("bind/listen: %s", strerror(errno));
    close(fd);
    return B200_ERR_INVALID;
  }
  socklen_t al = sizeof(a);
  getsockname(fd, reinterpret_cast<sockaddr*>(&a), &al);
  if (bound_port) *bound_port = ntohs(a.sin_port);
  sv.listen_fd = fd;
  sv.acceptor = std::thread([&sv, fd] {
    while (!sv.stopping) {
      int c = accept(fd, nullptr, nullptr);
      if (c < 0) {
        if (sv.stopping) break;
        if (errno == EINTR) continue;
        break;
      }
      if (sv.live_conns.load() >= Server::kMaxConns) {
        static const char busy[] = "HTTP/1.1 503 Service Unavailable\r\nContent-Length: 0\r\nConnection: close\r\n\r\n";
        (void)!::send(c, busy, sizeof(busy) - 1, MSG_NOSIGNAL);
        close(c);
        continue;
      }
      {
        std::lock_guard<std::mutex> lk(sv.conn_mu);
        if (sv.stopping) { close(c); break; }
        sv.conn_fds.insert(c);
      }
      sv.live_conns.fetch_add(1);
      std::thread(serve_conn, &sv, c).detach();
    }
  });
  return 0;
}

int b200_server_metrics(b200_server* s, char* buf, size_t cap) {
  if (!s || !buf || cap == 0) { set_error("bad arguments"); return B200_ERR_INVALID; }
  std::string m = metrics_text(s->impl);
  snprintf(buf, cap, "%s", m.c_str());
  return static_cast<int>(m.size());
}

int b200_server_set_tokenizer(b200_server* s, const b200_tokenizer* t) {
  if (!s) { set_error("b200_server_set_tokenizer: bad arguments"); return B200_ERR_INVALID; }
  if (t && b200_tokenizer_vocab_size(t) > s->impl.tok.vocab) {
    set_error("the tokenizer has %d ids, the model's vocabulary %d", b200_tokenizer_vocab_size(t), s->impl.tok.vocab);
    return B200_ERR_INVALID;
  }
  s->impl.real_tok = t;
  return 0;
}

int64_t b200_server_render_prompt(b200_server* s, const char* path, const char* content_type, const char* body, size_t len, int32_t* ids,
                                  size_t cap) {
  if (!s || !path || (!body && len)) { set_error("b200_server_render_prompt: bad arguments"); return -1; }
  ParsedRequest pr;
  std::string err;
  if (int st = parse_request(s->impl, path, content_type ? content_type : "application/json", body, len, &pr, &err)) {
    set_error("%d %s", st, err.c_str());
    return -1;
  }
  std::vector<int32_t> out;
  render_prompt(s->impl, pr, &out);
  for (size_t i = 0; i < out.size() && i < cap; ++i) ids[i] = out[i];
  return static_cast<int64_t>(out.size());
}

int b200_server_inject_fault(b200_server* s, int32_t replica, int32_t count) {
  if (!s || replica < 0 || replica >= static_cast<int>(s->impl.replicas.size())) { set_error("bad replica"); return B200_ERR_INVALID; }
  s->impl.faults[replica]->store(count);
  return 0;
}

int b200_tokenize(int32_t vocab, const char* text, size_t len, int32_t* out, int32_t cap) {
  Tokenizer t;
  t.vocab = vocab;
  std::vector<int32_t> ids;
  t.encode(std::string(text ? text : "", text ? len : 0), &ids);
  const int n = static_cast<int>(ids.size());
  if (out) memcpy(out, ids.data(), static_cast<size_t>(std::min(n, cap)) * 4);
  return n;
}

int b200_detokenize(int32_t vocab, const int32_t* ids, int32_t n, char* out, size_t cap) {
  Tokenizer t;
  t.vocab = vocab;
  std::string s;
  for (int i = 0; i < n; ++i) s += t.piece(ids[i]);
  if (out && cap) snprintf(out, cap, "%s", s.c_str());
  return static_cast<int>(s.size());
}

}  // extern "C"
