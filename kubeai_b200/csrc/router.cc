// Request router over GPU replicas: C++ restatement of the reference's endpoint group
//   internal/loadbalancer/group.go:25-150          (group, getBestAddr, reconcileEndpoints, broadcast)
//   internal/loadbalancer/balance_chwbl.go:14-162  (CHWBL: XXH64 ring, bounded-load walk)
//   internal/loadbalancer/balance_least_load.go:3-23
// Data-parallel strategy of the reference: independent replicas, no collective.
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/b200engine.h"
#include "errors.h"
#include "xxh64.h"

namespace b200 {

struct RouterEndpoint {
  std::string address;
  std::set<std::string> adapters;
  std::shared_ptr<std::atomic<int64_t>> in_flight;
  uint64_t token = 0;
};

struct Router {
  std::shared_mutex mtx;  // group.mtx
  std::map<std::string, RouterEndpoint> endpoints;
  std::atomic<int64_t> total_in_flight{0};
  int replication = 256;
  std::unordered_map<uint64_t, std::string> chwbl_hashes;  // hash -> endpoint name
  std::vector<uint64_t> chwbl_sorted;

  std::mutex bmtx;  // group.bmtx + bcast channel == generation counter + condvar
  std::condition_variable bcv;
  uint64_t generation = 0;

  std::mutex tmtx;
  std::vector<std::shared_ptr<std::atomic<int64_t>>> counters;  // token -> in-flight counter (outlives endpoints)

  // metrics (internal/metrics/metrics.go:16-76): iterations histogram (buckets 1..1024) and per-endpoint
  // initial / final / default counters, emitted where balance_chwbl.go:24-26,58-61,74-80 emits them
  std::atomic<int64_t> lookups{0}, lookup_iterations{0}, lookup_defaults{0};
  std::atomic<int64_t> iter_hist[12] = {};  // le 1,2,4,...,1024,+Inf (non-cumulative)
  std::mutex emtx;
  std::map<std::string, std::array<int64_t, 3>> ep_metrics;  // endpoint -> {initial, final, default}
  void count_ep(const std::string& name, int which) {
    std::lock_guard<std::mutex> lk(emtx);
    ++ep_metrics[name][which];
  }
  void observe_iterations(int64_t n) {
    lookup_iterations.fetch_add(n, std::memory_order_relaxed);
    int b = 0;
    while (b < 11 && n > (1ll << b)) ++b;
    iter_hist[b].fetch_add(1, std::memory_order_relaxed);
  }

  static uint64_t hash(const std::string& s) { return xxh64(s.data(), s.size(), 0); }
  // chwblEndpointReplicaHashInput: fmt.Sprintf("%s%d", name, replica)
  static std::string replica_input(const std::string& name, int i) { return name + std::to_string(i); }

  void chwbl_add(const std::string& name) {
    for (int i = 0; i < replication; ++i) {
      uint64_t h = hash(replica_input(name, i));
      chwbl_hashes[h] = name;  // collisions overwrite, as in the reference
      chwbl_sorted.push_back(h);
    }
    std::sort(chwbl_sorted.begin(), chwbl_sorted.end());
  }
  void chwbl_remove(const std::string& name) {
    for (int i = 0; i < replication; ++i) {
      uint64_t h = hash(replica_input(name, i));
      chwbl_hashes.erase(h);
      auto it = std::lower_bound(chwbl_sorted.begin(), chwbl_sorted.end(), h);
      if (it != chwbl_sorted.end() && *it == h) chwbl_sorted.erase(it);
    }
  }
  static bool load_ok(int64_t load, int64_t total, int n, double factor) {
    if (total == 0) return true;
    const double avg = static_cast<double>(total + 1) / static_cast<double>(n);
    return static_cast<double>(load) <= avg * factor;
  }
  const RouterEndpoint* chwbl_get(const std::string& key, double factor, const std::string& adapter) {
    if (chwbl_sorted.empty()) return nullptr;
    const uint64_t h = hash(key);
    size_t i = std::lower_bound(chwbl_sorted.begin(), chwbl_sorted.end(), h) - chwbl_sorted.begin();
    if (i >= chwbl_sorted.size()) i = 0;
    lookups.fetch_add(1, std::memory_order_relaxed);
    count_ep(chwbl_hashes[chwbl_sorted[i]], 0);
    const RouterEndpoint* def = nullptr;
    std::string def_name;
    for (size_t n = 0; n < chwbl_sorted.size(); ++n) {
      const std::string& name = chwbl_hashes[chwbl_sorted[i]];
      auto it = endpoints.find(name);
      if (it == endpoints.end()) {
        fprintf(stderr, "[b200router] endpoints corrupted, %s should be in map\n", name.c_str());
        abort();
      }
      const RouterEndpoint& ep = it->second;
      const bool match = adapter.empty() || ep.adapters.count(adapter);
      if (match) {
        if (!def) {
          def = &ep;
          def_name = name;
        }
        if (load_ok(ep.in_flight->load(), total_in_flight.load(), static_cast<int>(endpoints.size()), factor)) {
          observe_iterations(static_cast<int64_t>(n + 1));
          count_ep(name, 1);
          return &ep;
        }
      }
      if (++i >= chwbl_sorted.size()) i = 0;
    }
    if (def) {
      observe_iterations(static_cast<int64_t>(chwbl_sorted.size()));
      lookup_defaults.fetch_add(1, std::memory_order_relaxed);
      count_ep(def_name, 1);
      count_ep(def_name, 2);
    }
    return def;
  }
  const RouterEndpoint* least_load(const std::string& adapter) {
    const RouterEndpoint* best = nullptr;
    int64_t min_in_flight = 0;
    for (auto& kv : endpoints) {
      const RouterEndpoint& ep = kv.second;
      if (!adapter.empty() && !ep.adapters.count(adapter)) continue;
      const int64_t f = ep.in_flight->load();
      if (!best || f < min_in_flight) {
        best = &ep;
        min_in_flight = f;
      }
    }
    return best;
  }
};

}  // namespace b200

using namespace b200;

struct b200_router {
  Router impl;
};

extern "C" {

uint64_t b200_xxh64(const void* data, size_t len) { return xxh64(data, len, 0); }

int b200_router_create(int32_t replication, b200_router** out) {
  if (!out || replication < 0) { set_error("b200_router_create: bad arguments"); return B200_ERR_INVALID; }
  b200_router* r = new (std::nothrow) b200_router();
  if (!r) { set_error("host OOM"); return B200_ERR_OOM; }
  r->impl.replication = replication;
  *out = r;
  return 0;
}

void b200_router_destroy(b200_router* r) { delete r; }

int b200_router_set_endpoints(b200_router* r, const char* const* names, const char* const* addresses,
                              const char* const* adapters, int32_t n) {
  if (!r || n < 0 || (n > 0 && (!names || !addresses))) { set_error("b200_router_set_endpoints: bad arguments"); return B200_ERR_INVALID; }
  Router& g = r->impl;
  std::map<std::string, std::pair<std::string, std::set<std::string>>> observed;
  for (int i = 0; i < n; ++i) {
    std::set<std::string> ad;
    if (adapters && adapters[i]) {
      std::string s(adapters[i]);
      size_t p = 0;
      while (p <= s.size()) {
        size_t q = s.find(',', p);
        if (q == std::string::npos) q = s.size();
        if (q > p) ad.insert(s.substr(p, q - p));
        p = q + 1;
      }
    }
    observed[names[i]] = {addresses[i], ad};
  }
  {
    std::unique_lock<std::shared_mutex> lk(g.mtx);
    for (auto& kv : observed) {
      auto it = g.endpoints.find(kv.first);
      if (it != g.endpoints.end()) {
        it->second.adapters = kv.second.second;
      } else {
        RouterEndpoint ep;
        ep.address = kv.second.first;
        ep.adapters = kv.second.second;
        ep.in_flight = std::make_shared<std::atomic<int64_t>>(0);
        {
          std::lock_guard<std::mutex> tl(g.tmtx);
          ep.token = g.counters.size();
          g.counters.push_back(ep.in_flight);
        }
        g.endpoints[kv.first] = ep;
        g.chwbl_add(kv.first);
      }
    }
    for (auto it = g.endpoints.begin(); it != g.endpoints.end();) {
      if (!observed.count(it->first)) {
        // in-flight counts of a vanished endpoint drain when its requests complete (group.go:123-130)
        g.chwbl_remove(it->first);
        it = g.endpoints.erase(it);
      } else {
        ++it;
      }
    }
  }
  if (n > 0) {
    std::lock_guard<std::mutex> bl(g.bmtx);
    ++g.generation;
    g.bcv.notify_all();
  }
  return 0;
}

int b200_router_pick(b200_router* r, int32_t strategy, const char* adapter, const char* prefix, int32_t prefix_len,
                     int32_t mean_load_pct, int64_t timeout_us, char* addr_out, int32_t addr_cap,
                     uint64_t* endpoint_token) {
  if (!r || !addr_out || addr_cap <= 0 || !endpoint_token) { set_error("b200_router_pick: bad arguments"); return B200_ERR_INVALID; }
  if (strategy != B200_LB_LEAST_LOAD && strategy != B200_LB_PREFIX_HASH) {
    set_error("unknown load balancing strategy: %d", strategy);
    return B200_ERR_INVALID;
  }
  Router& g = r->impl;
  const std::string ad = adapter ? adapter : "";
  std::string key = ad;
  if (prefix && prefix_len > 0) key.append(prefix, static_cast<size_t>(prefix_len));
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(timeout_us < 0 ? 0 : timeout_us);
  bool await_change = false;
  for (;;) {
    uint64_t gen;
    {
      std::lock_guard<std::mutex> bl(g.bmtx);
      gen = g.generation;
    }
    {
      std::shared_lock<std::shared_mutex> lk(g.mtx);
      if (!await_change && !g.endpoints.empty()) {
        const RouterEndpoint* ep = strategy == B200_LB_PREFIX_HASH
                                       ? g.chwbl_get(key, static_cast<double>(mean_load_pct) / 100.0, ad)
                                       : g.least_load(ad);
        if (ep) {
          g.total_in_flight.fetch_add(1);
          ep->in_flight->fetch_add(1);
          snprintf(addr_out, static_cast<size_t>(addr_cap), "%s", ep->address.c_str());
          *endpoint_token = ep->token;
          return 0;
        }
      }
    }
    // block until the endpoint set changes (group.go:56-64, 77-80)
    std::unique_lock<std::mutex> bl(g.bmtx);
    auto changed = [&] { return g.generation != gen; };
    if (timeout_us < 0) {
      g.bcv.wait(bl, changed);
    } else if (timeout_us == 0 || !g.bcv.wait_until(bl, deadline, changed)) {
      set_error("context deadline exceeded");
      return B200_ERR_TIMEOUT;
    }
    await_change = false;
  }
}

int b200_router_done(b200_router* r, uint64_t endpoint_token) {
  if (!r) { set_error("null router"); return B200_ERR_INVALID; }
  Router& g = r->impl;
  std::shared_ptr<std::atomic<int64_t>> c;
  {
    std::lock_guard<std::mutex> tl(g.tmtx);
    if (endpoint_token >= g.counters.size()) { set_error("bad endpoint token"); return B200_ERR_NOT_FOUND; }
    c = g.counters[endpoint_token];
  }
  g.total_in_flight.fetch_sub(1);
  c->fetch_sub(1);
  return 0;
}

/* test hook: group.addInFlight (group.go:147-150) on a named endpoint */
int b200_router_add_inflight(b200_router* r, const char* name, int64_t delta) {
  if (!r || !name) { set_error("bad arguments"); return B200_ERR_INVALID; }
  Router& g = r->impl;
  std::shared_lock<std::shared_mutex> lk(g.mtx);
  auto it = g.endpoints.find(name);
  if (it == g.endpoints.end()) { set_error("unknown endpoint %s", name); return B200_ERR_NOT_FOUND; }
  g.total_in_flight.fetch_add(delta);
  it->second.in_flight->fetch_add(delta);
  return 0;
}

/* Prometheus text of the hash-lookup instruments (names as OtelNameToPromName would render them). */
int64_t b200_router_metrics(b200_router* r, char* buf, size_t cap) {
  if (!r) { set_error("null router"); return B200_ERR_INVALID; }
  Router& g = r->impl;
  std::string o = "# TYPE kubeai_inference_requests_hash_lookup_iterations histogram\n";
  int64_t cum = 0;
  for (int b = 0; b < 12; ++b) {
    cum += g.iter_hist[b].load();
    o += "kubeai_inference_requests_hash_lookup_iterations_bucket{le=\"" + (b < 11 ? std::to_string(1 << b) : std::string("+Inf")) + "\"} " + std::to_string(cum) + "\n";
  }
  o += "kubeai_inference_requests_hash_lookup_iterations_sum " + std::to_string(g.lookup_iterations.load()) + "\n";
  o += "kubeai_inference_requests_hash_lookup_iterations_count " + std::to_string(cum) + "\n";
  static const char* kinds[3] = {"initial", "final", "default"};
  std::lock_guard<std::mutex> lk(g.emtx);
  for (int k = 0; k < 3; ++k) {
    o += std::string("# TYPE kubeai_inference_requests_hash_lookup_") + kinds[k] + " counter\n";
    for (auto& kv : g.ep_metrics)
      if (kv.second[k]) o += std::string("kubeai_inference_requests_hash_lookup_") + kinds[k] + "{endpoint=\"" + kv.first + "\"} " + std::to_string(kv.second[k]) + "\n";
  }
  if (buf && cap > o.size()) memcpy(buf, o.c_str(), o.size() + 1);
  return static_cast<int64_t>(o.size());
}

int b200_router_inflight(b200_router* r, const char* name, int64_t* endpoint_inflight, int64_t* total_inflight) {
  if (!r) { set_error("null router"); return B200_ERR_INVALID; }
  Router& g = r->impl;
  std::shared_lock<std::shared_mutex> lk(g.mtx);
  if (total_inflight) *total_inflight = g.total_in_flight.load();
  if (name && endpoint_inflight) {
    auto it = g.endpoints.find(name);
    if (it == g.endpoints.end()) { set_error("unknown endpoint %s", name); return B200_ERR_NOT_FOUND; }
    *endpoint_inflight = it->second.in_flight->load();
  }
  return 0;
}

}  // extern "C"
