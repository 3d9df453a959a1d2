// HF-layout checkpoint loading (SURVEY.md §8f-2 "next": real weights).  The reference mounts a model
// directory into the backend pod (internal/modelcontroller/model_source.go:231-287, hf:// / pvc:// sources)
// and vLLM reads its *.safetensors; the in-process engine reads the same files itself:
//   <dir>/config.json                       -> b200_config (b200_config_from_hf)
//   <dir>/model.safetensors[.index.json]    -> engine tensors (b200_engine_load_safetensors)
// q/k/v and gate/up projections are fused on load into the engine's wqkv / wgu layout (same row order as
// vLLM's merged qkv_proj / gate_up_proj, model_executor/models/llama.py:81-121,157-161).
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <math.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/b200engine.h"
#include "errors.h"
#include "hostutil.h"

namespace b200 {
namespace {

struct StTensor {
  std::string dtype;
  std::vector<int64_t> shape;
  size_t begin = 0, end = 0;  // byte offsets inside the data section
};

struct StFile {
  int fd = -1;
  const uint8_t* map = nullptr;
  size_t size = 0, data_off = 0;
  std::map<std::string, StTensor> tensors;
  ~StFile() {
    if (map) munmap(const_cast<uint8_t*>(map), size);
    if (fd >= 0) close(fd);
  }
  bool open_file(const std::string& path, std::string* err) {
    fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) { *err = "cannot open " + path; return false; }
    struct stat st;
    if (fstat(fd, &st) || st.st_size < 8) { *err = "bad safetensors file " + path; return false; }
    size = static_cast<size_t>(st.st_size);
    void* p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (p == MAP_FAILED) { *err = "mmap failed for " + path; return false; }
    map = static_cast<const uint8_t*>(p);
    uint64_t hlen = 0;
    memcpy(&hlen, map, 8);
    if (hlen > size - 8) { *err = "safetensors header length out of range in " + path; return false; }
    data_off = 8 + static_cast<size_t>(hlen);
    JVal root;
    std::string jerr;
    if (!JParser(reinterpret_cast<const char*>(map + 8), static_cast<size_t>(hlen)).parse(&root, &jerr) || root.type != JVal::Obj) {
      *err = "safetensors header of " + path + ": " + jerr;
      return false;
    }
    for (auto& kv : root.obj) {
      if (kv.first == "__metadata__") continue;
      StTensor t;
      const JVal* dt = kv.second.get("dtype");
      const JVal* sh = kv.second.get("shape");
      const JVal* off = kv.second.get("data_offsets");
      if (!dt || dt->type != JVal::Str || !sh || sh->type != JVal::Arr || !off || off->type != JVal::Arr || off->arr.size() != 2) {
        *err = "malformed entry " + kv.first + " in " + path;
        return false;
      }
      t.dtype = dt->str;
      // checkpoints are untrusted input: every number must be a non-negative integer below 2^53, offsets must lie inside
      // the data section (compared without forming data_off + end, which could wrap), shape products must not overflow
      auto as_index = [](const JVal& v, uint64_t* out) {
        if (v.type != JVal::Num || !(v.num >= 0.0) || v.num >= 9007199254740992.0 || v.num != floor(v.num)) return false;
        *out = static_cast<uint64_t>(v.num);
        return true;
      };
      uint64_t b = 0, e = 0, elems = 1;
      bool ok = as_index(off->arr[0], &b) && as_index(off->arr[1], &e);
      for (auto& d : sh->arr) {
        uint64_t dim = 0;
        if (!ok || !as_index(d, &dim) || (dim != 0 && elems > (1ull << 53) / dim)) { ok = false; break; }
        elems *= dim;
        t.shape.push_back(static_cast<int64_t>(dim));
      }
      const uint64_t data_len = static_cast<uint64_t>(size - data_off);
      if (!ok || e < b || e > data_len) { *err = "shape or data_offsets out of range for " + kv.first; return false; }
      t.begin = static_cast<size_t>(b);
      t.end = static_cast<size_t>(e);
      tensors[kv.first] = t;
    }
    return true;
  }
};

bool read_text(const std::string& path, std::string* out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  char buf[65536];
  size_t n;
  out->clear();
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out->append(buf, n);
  fclose(f);
  return true;
}

bool is_dir(const std::string& p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

size_t dtype_size(const std::string& d) { return d == "BF16" || d == "F16" ? 2 : d == "F32" ? 4 : 0; }

inline uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
inline float f16_to_f32(uint16_t h) {
  const uint32_t s = (h & 0x8000u) << 16, e = (h >> 10) & 0x1f, m = h & 0x3ffu;
  uint32_t u;
  if (e == 0) {
    if (m == 0) u = s;
    else {
      int sh = 0;
      uint32_t mm = m;
      while (!(mm & 0x400u)) { mm <<= 1; ++sh; }
      u = s | ((113 - sh) << 23) | ((mm & 0x3ffu) << 13);
    }
  } else if (e == 31) u = s | 0x7f800000u | (m << 13);
  else u = s | ((e + 112) << 23) | (m << 13);
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// All shards of a checkpoint: tensor name -> (file, entry)
struct Checkpoint {
  std::vector<std::unique_ptr<StFile>> files;
  std::map<std::string, std::pair<StFile*, const StTensor*>> index;
  bool open_path(const std::string& path, std::string* err) {
    std::vector<std::string> shard_paths;
    if (is_dir(path)) {
      std::string idx;
      if (read_text(path + "/model.safetensors.index.json", &idx)) {
        JVal root;
        std::string jerr;
        if (!JParser(idx.data(), idx.size()).parse(&root, &jerr)) { *err = "index json: " + jerr; return false; }
        const JVal* wm = root.get("weight_map");
        if (!wm || wm->type != JVal::Obj) { *err = "index json has no weight_map"; return false; }
        std::map<std::string, bool> seen;
        for (auto& kv : wm->obj)
          if (kv.second.type == JVal::Str && !seen[kv.second.str]) {
            seen[kv.second.str] = true;
            shard_paths.push_back(path + "/" + kv.second.str);
          }
      } else {
        shard_paths.push_back(path + "/model.safetensors");
      }
    } else {
      shard_paths.push_back(path);
    }
    for (auto& sp : shard_paths) {
      files.emplace_back(new StFile());
      if (!files.back()->open_file(sp, err)) return false;
      for (auto& kv : files.back()->tensors) index[kv.first] = {files.back().get(), &kv.second};
    }
    return true;
  }
};

// Copy a [rows, cols] checkpoint tensor into rows [row0, row0+rows) of a bf16 device matrix with `cols` columns.
int copy_rows(const Checkpoint& ck, const std::string& name, void* dev_base, size_t dev_bytes, int64_t row0, int64_t rows,
              int64_t cols, std::string* err) {
  auto it = ck.index.find(name);
  if (it == ck.index.end()) { *err = "missing tensor " + name; return B200_ERR_NOT_FOUND; }
  const StTensor& t = *it->second.second;
  const StFile& f = *it->second.first;
  int64_t n = 1;
  for (auto d : t.shape) n *= d;
  const int64_t trows = t.shape.size() == 2 ? t.shape[0] : 1, tcols = t.shape.size() == 2 ? t.shape[1] : n;
  if (trows != rows || tcols != cols) {
    *err = name + ": shape [" + std::to_string(trows) + "," + std::to_string(tcols) + "] != expected [" + std::to_string(rows) + "," + std::to_string(cols) + "]";
    return B200_ERR_INVALID;
  }
  const size_t es = dtype_size(t.dtype);
  if (!es || t.end - t.begin != static_cast<size_t>(n) * es) { *err = name + ": unsupported dtype " + t.dtype; return B200_ERR_INVALID; }
  const size_t dst_off = static_cast<size_t>(row0) * cols * 2;
  if (dst_off + static_cast<size_t>(n) * 2 > dev_bytes) { *err = name + ": does not fit the engine tensor"; return B200_ERR_INVALID; }
  const uint8_t* src = f.map + f.data_off + t.begin;
  uint8_t* dst = static_cast<uint8_t*>(dev_base) + dst_off;
  cudaError_t ce;
  if (t.dtype == "BF16") {
    ce = cudaMemcpy(dst, src, static_cast<size_t>(n) * 2, cudaMemcpyHostToDevice);
  } else {
    std::vector<uint16_t> tmp(static_cast<size_t>(n));
    if (t.dtype == "F32") {
      const float* p = reinterpret_cast<const float*>(src);
      for (int64_t i = 0; i < n; ++i) tmp[i] = f32_to_bf16(p[i]);
    } else {
      const uint16_t* p = reinterpret_cast<const uint16_t*>(src);
      for (int64_t i = 0; i < n; ++i) tmp[i] = f32_to_bf16(f16_to_f32(p[i]));
    }
    ce = cudaMemcpy(dst, tmp.data(), static_cast<size_t>(n) * 2, cudaMemcpyHostToDevice);
  }
  if (ce != cudaSuccess) { *err = std::string("cudaMemcpy: ") + cudaGetErrorString(ce); return B200_ERR_CUDA; }
  return 0;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

// JSON listing {"name": {"dtype": "...", "shape": [...], "bytes": n}, ...} of a checkpoint (dir, index or single file).
int64_t b200_safetensors_list(const char* path, char* buf, size_t cap) {
  if (!path) { set_error("null path"); return B200_ERR_INVALID; }
  Checkpoint ck;
  std::string err;
  if (!ck.open_path(path, &err)) { set_error("%s", err.c_str()); return B200_ERR_INVALID; }
  std::string o = "{";
  bool first = true;
  for (auto& kv : ck.index) {
    if (!first) o += ",";
    first = false;
    o += json_str(kv.first) + ":{\"dtype\":" + json_str(kv.second.second->dtype) + ",\"shape\":[";
    for (size_t i = 0; i < kv.second.second->shape.size(); ++i) o += (i ? "," : "") + std::to_string(kv.second.second->shape[i]);
    o += "],\"bytes\":" + std::to_string(kv.second.second->end - kv.second.second->begin) + "}";
  }
  o += "}";
  if (buf && cap > o.size()) memcpy(buf, o.c_str(), o.size() + 1);
  return static_cast<int64_t>(o.size());
}

// Fill the architecture fields of cfg from <dir>/config.json (HF LlamaConfig keys); other fields keep their values.
int b200_config_from_hf(const char* dir, b200_config* cfg) {
  if (!dir || !cfg) { set_error("null argument"); return B200_ERR_INVALID; }
  std::string txt;
  if (!read_text(std::string(dir) + "/config.json", &txt)) { set_error("cannot read %s/config.json", dir); return B200_ERR_NOT_FOUND; }
  JVal root;
  std::string jerr;
  if (!JParser(txt.data(), txt.size()).parse(&root, &jerr) || root.type != JVal::Obj) { set_error("config.json: %s", jerr.c_str()); return B200_ERR_INVALID; }
  auto num = [&](const char* k, double def) { const JVal* v = root.get(k); return v && v->type == JVal::Num ? v->num : def; };
  cfg->num_layers = static_cast<int>(num("num_hidden_layers", cfg->num_layers));
  cfg->hidden = static_cast<int>(num("hidden_size", cfg->hidden));
  cfg->q_heads = static_cast<int>(num("num_attention_heads", cfg->q_heads));
  cfg->kv_heads = static_cast<int>(num("num_key_value_heads", cfg->q_heads));
  cfg->intermediate = static_cast<int>(num("intermediate_size", cfg->intermediate));
  cfg->vocab = static_cast<int>(num("vocab_size", cfg->vocab));
  cfg->rms_eps = static_cast<float>(num("rms_norm_eps", cfg->rms_eps));
  double theta = num("rope_theta", cfg->rope_theta);
  if (const JVal* rp = root.get("rope_parameters")) if (const JVal* t = rp->get("rope_theta"); t && t->type == JVal::Num) theta = t->num;
  cfg->rope_theta = static_cast<float>(theta);
  // RoPE frequency scaling: HF "rope_scaling" (transformers < 5) or "rope_parameters" (>= 5); Llama-3.1/3.2 ship
  // rope_type "llama3" with factor 8 — loading those with an unscaled table would silently diverge from vLLM/HF
  cfg->rope_scaling_type = 0;
  const JVal* rs = root.get("rope_scaling");
  if (!rs || rs->type != JVal::Obj) rs = root.get("rope_parameters");
  if (rs && rs->type == JVal::Obj) {
    std::string ty = "default";
    if (const JVal* t = rs->get("rope_type"); t && t->type == JVal::Str) ty = t->str;
    else if (const JVal* t2 = rs->get("type"); t2 && t2->type == JVal::Str) ty = t2->str;
    auto rnum = [&](const char* k, double def) { const JVal* v = rs->get(k); return v && v->type == JVal::Num ? v->num : def; };
    if (ty == "default") {
    } else if (ty == "linear") {
      cfg->rope_scaling_type = 1;
      cfg->rope_factor = static_cast<float>(rnum("factor", 1.0));
    } else if (ty == "llama3") {
      cfg->rope_scaling_type = 2;
      cfg->rope_factor = static_cast<float>(rnum("factor", 8.0));
      cfg->rope_low_freq_factor = static_cast<float>(rnum("low_freq_factor", 1.0));
      cfg->rope_high_freq_factor = static_cast<float>(rnum("high_freq_factor", 4.0));
      cfg->rope_original_max_pos = static_cast<int>(rnum("original_max_position_embeddings", 8192));
    } else {
      set_error("rope scaling type \"%s\" is not supported (default, linear, llama3)", ty.c_str());
      return B200_ERR_INVALID;
    }
  }
  // architecture switches the kernels do not implement are refused, not ignored
  if (const JVal* mt = root.get("model_type"); mt && mt->type == JVal::Str && mt->str != "llama") {
    set_error("model_type \"%s\" is not supported (llama)", mt->str.c_str());
    return B200_ERR_INVALID;
  }
  if (const JVal* ha = root.get("hidden_act"); ha && ha->type == JVal::Str && ha->str != "silu") {
    set_error("hidden_act \"%s\" is not supported (silu)", ha->str.c_str());
    return B200_ERR_INVALID;
  }
  for (const char* k : {"attention_bias", "mlp_bias"})
    if (const JVal* bflag = root.get(k); bflag && bflag->type == JVal::Bool && bflag->b) {
      set_error("%s = true is not supported", k);
      return B200_ERR_INVALID;
    }
  if (cfg->hidden % 256 || cfg->hidden > 8192 || cfg->vocab % 8 || cfg->intermediate % 64) {
    set_error("shape not supported by the kernels (hidden %% 256 == 0 and <= 8192, vocab %% 8 == 0, intermediate %% 64 == 0)");
    return B200_ERR_INVALID;
  }
  const int head_dim = static_cast<int>(num("head_dim", cfg->q_heads ? cfg->hidden / cfg->q_heads : 0));
  if (head_dim != 128) { set_error("head_dim %d is not supported (kernels are specialised for 128)", head_dim); return B200_ERR_INVALID; }
  if (cfg->q_heads != 4 * cfg->kv_heads) { set_error("GQA ratio %d:%d is not supported (kernels are specialised for 4:1)", cfg->q_heads, cfg->kv_heads); return B200_ERR_INVALID; }
  if (const JVal* e = root.get("eos_token_id"); e && e->type == JVal::Num) cfg->eos_token_id = static_cast<int>(e->num);
  return 0;
}

// Replace the engine's (seeded random) weights with an HF Llama checkpoint.  The engine must be idle.
int b200_engine_load_safetensors(b200_engine* e, const char* path) {
  if (!e || !path) { set_error("null argument"); return B200_ERR_INVALID; }
  Checkpoint ck;
  std::string err;
  if (!ck.open_path(path, &err)) { set_error("%s", err.c_str()); return B200_ERR_INVALID; }
  auto dev = [&](const std::string& name, void** p, uint64_t* nb) { return b200_engine_tensor_info(e, name.c_str(), nb, p); };
  void* p = nullptr;
  uint64_t nb = 0;
  // shapes are recovered from the engine tensors themselves
  if (dev("final_norm", &p, &nb)) return B200_ERR_NOT_FOUND;
  const int64_t H = static_cast<int64_t>(nb / 2);
  if (dev("embed", &p, &nb)) return B200_ERR_NOT_FOUND;
  const int64_t V = static_cast<int64_t>(nb / 2) / H;
  int rc;
#define LD(hfname, row0, rows, cols) if ((rc = copy_rows(ck, hfname, p, nb, row0, rows, cols, &err))) { set_error("%s", err.c_str()); return rc; }
  LD("model.embed_tokens.weight", 0, V, H);
  if (dev("final_norm", &p, &nb)) return B200_ERR_NOT_FOUND;
  LD("model.norm.weight", 0, 1, H);
  if (dev("lm_head", &p, &nb)) return B200_ERR_NOT_FOUND;
  {
    // tie_word_embeddings: checkpoints without lm_head.weight reuse the embedding matrix
    const std::string head = ck.index.count("lm_head.weight") ? "lm_head.weight" : "model.embed_tokens.weight";
    LD(head, 0, V, H);
  }
  for (int l = 0;; ++l) {
    const std::string ep = "layers." + std::to_string(l) + ".", hp = "model.layers." + std::to_string(l) + ".";
    if (dev(ep + "wqkv", &p, &nb)) break;  // no more layers in the engine
    auto qit = ck.index.find(hp + "self_attn.q_proj.weight"), kit = ck.index.find(hp + "self_attn.k_proj.weight");
    if (qit == ck.index.end() || kit == ck.index.end() || qit->second.second->shape.size() != 2) { set_error("missing %sself_attn.{q,k}_proj.weight", hp.c_str()); return B200_ERR_NOT_FOUND; }
    const int64_t qrows = qit->second.second->shape[0], krows = kit->second.second->shape[0];
    LD(hp + "self_attn.q_proj.weight", 0, qrows, H);
    LD(hp + "self_attn.k_proj.weight", qrows, krows, H);
    LD(hp + "self_attn.v_proj.weight", qrows + krows, krows, H);
    if (dev(ep + "wo", &p, &nb)) return B200_ERR_NOT_FOUND;
    LD(hp + "self_attn.o_proj.weight", 0, H, qrows);
    if (dev(ep + "wgu", &p, &nb)) return B200_ERR_NOT_FOUND;
    const int64_t I = static_cast<int64_t>(nb / 2) / H / 2;
    LD(hp + "mlp.gate_proj.weight", 0, I, H);
    LD(hp + "mlp.up_proj.weight", I, I, H);
    if ((rc = engine_relayout_gate_up(e, (ep + "wgu").c_str()))) return rc;   // logical [gate | up] -> interleaved 64-row blocks
    if (dev(ep + "wdown", &p, &nb)) return B200_ERR_NOT_FOUND;
    LD(hp + "mlp.down_proj.weight", 0, H, I);
    if (dev(ep + "norm1", &p, &nb)) return B200_ERR_NOT_FOUND;
    LD(hp + "input_layernorm.weight", 0, 1, H);
    if (dev(ep + "norm2", &p, &nb)) return B200_ERR_NOT_FOUND;
    LD(hp + "post_attention_layernorm.weight", 0, 1, H);
  }
#undef LD
  return 0;
}

}  // extern "C"
