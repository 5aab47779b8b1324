// capi_experiment.cu — host-only entry points of the C ABI: the seeded Fuzzer (fuzzing/Fuzzer.scala:24-194) and the
// flat experiment directory that replaces the reference's Java-serialized *.bin files (Serialization.scala:57-74,
// :176-254).  No device code; lives in the library so that a JVM host gets them through the same JNI shim.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <sys/stat.h>
#include "engine.hpp"

namespace {
// java.util.Random (Java SE specification); scala.util.Random delegates to it (Fuzzer.scala:68)
struct JavaRandom {
  uint64_t s;
  explicit JavaRandom(int64_t seed) : s(((uint64_t)seed ^ 0x5DEECE66Dull) & ((1ull << 48) - 1)) {}
  int32_t next(int bits) { s = (s * 0x5DEECE66Dull + 0xBull) & ((1ull << 48) - 1); return (int32_t)(int64_t)(s >> (48 - bits)); }
  int32_t nextInt(int32_t bound) {
    int32_t r = next(31), m = bound - 1;
    if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
    for (int32_t u = r; u - (r = u % bound) + m < 0; u = next(31)) {}
    return r;
  }
  double nextDouble() { return (double)(((int64_t)next(26) << 27) + next(27)) * (1.0 / (double)(1ll << 53)); }
};
// RandomizedHashSet (schedulers/Util.scala:110-185) over small integer payloads
struct RandSet {
  std::vector<uint32_t> arr; JavaRandom rand;
  explicit RandSet(int64_t seed) : rand(seed) {}
  void insert(uint32_t v) { arr.push_back(v); }
  bool empty() const { return arr.empty(); }
  uint32_t removeRandomElement() { const int32_t i = rand.nextInt((int32_t)arr.size()); const uint32_t v = arr[i]; arr[i] = arr.back(); arr.pop_back(); return v; }
  uint32_t getRandomElement() { return arr[rand.nextInt((int32_t)arr.size())]; }
};
}  // namespace

// Fuzzer.generateFuzzTest (Fuzzer.scala:123-174) with generateNextEvent (:84-121), FuzzerWeights (:24-58) and reset
// (:176-193).  The reference seeds the fuzzer and its three sets from the wall clock; all four take `seed` here.
extern "C" int32_t demi_fuzzer_generate(const demi_fuzzer_config* cfg, int64_t seed,
                                        const demi_ext_event* prefix, uint32_t n_prefix,
                                        const demi_ext_event* postfix, uint32_t n_postfix,
                                        demi_ext_event* out, uint32_t cap, uint32_t* n_out) {
  if (!cfg || !out || !n_out || (!prefix && n_prefix) || (!postfix && n_postfix))
    return fail(nullptr, DEMI_ERR_INVALID, "demi_fuzzer_generate: null argument");
  const double weights[4] = {cfg->kill, cfg->send, cfg->partition, cfg->unpartition};     // allWeights, in this order (:40)
  const double total = cfg->kill + cfg->send + cfg->partition + cfg->unpartition + cfg->wait_quiescence;
  std::vector<uint32_t> nodes;
  uint32_t next_id = 0;
  for (uint32_t i = 0; i < n_prefix; i++) { if (prefix[i].kind == DEMI_EXT_START) nodes.push_back(prefix[i].a); next_id = std::max(next_id, prefix[i].id); }
  for (uint32_t i = 0; i < n_postfix; i++) next_id = std::max(next_id, postfix[i].id);
  JavaRandom rand(seed);
  RandSet alive(seed), partitioned(seed), unpartitioned(seed);
  for (uint32_t n : nodes) alive.insert(n);
  for (size_t i = 0; i < nodes.size(); i++) for (size_t j = i + 1; j < nodes.size(); j++) unpartitioned.insert(nodes[i] | (nodes[j] << 8));
  std::vector<demi_ext_event> test(prefix, prefix + n_prefix);
  uint32_t counter = 0;
  auto mk = [&](uint8_t kind, uint32_t a, uint32_t b, uint32_t type, uint32_t p0) {
    demi_ext_event e{}; e.kind = kind; e.a = (uint8_t)a; e.b = (uint8_t)b; e.type = (uint8_t)type; e.p0 = p0; e.id = ++next_id; return e;
  };
  // returns false for "no more events" (Kill with nobody alive, :92-95)
  auto next_event = [&](demi_ext_event& ev) -> bool {
    for (;;) {
      const double scaled = rand.nextDouble() * total;
      int t = -1; double cur = 0.0;
      for (int k = 0; k < 4; k++) { cur += weights[k]; if (scaled < cur) { t = k; break; } }
      if (t < 0) { ev = mk(DEMI_EXT_WAIT_QUIESCENCE, 0, 0, 0, 0); return true; }
      if (t == 0) { if (alive.empty()) return false; ev = mk(DEMI_EXT_KILL, alive.removeRandomElement(), 0, 0, 0); return true; }
      if (t == 1) { ++counter; ev = mk(DEMI_EXT_SEND, alive.getRandomElement(), 0, cfg->send_type, counter); return true; }
      if (t == 2) {
        if (unpartitioned.empty()) continue;                                                // "Try again..."
        const uint32_t p = unpartitioned.removeRandomElement(); partitioned.insert(p);
        ev = mk(DEMI_EXT_PARTITION, p & 0xFF, p >> 8, 0, 0); return true;
      }
      if (partitioned.empty()) continue;
      const uint32_t p = partitioned.removeRandomElement(); unpartitioned.insert(p);
      ev = mk(DEMI_EXT_UNPARTITION, p & 0xFF, p >> 8, 0, 0); return true;
    }
  };
  bool just_wq = !test.empty() && test.back().kind == DEMI_EXT_WAIT_QUIESCENCE;
  bool ran_out = false;
  for (uint32_t i = 0; i < cfg->num_events && !ran_out; i++) {
    demi_ext_event ev{};
    bool ok = next_event(ev);
    while (ok && ev.kind == DEMI_EXT_WAIT_QUIESCENCE && just_wq) ok = next_event(ev);       // no two WaitQuiescence in a row
    if (!ok) { ran_out = true; break; }
    just_wq = ev.kind == DEMI_EXT_WAIT_QUIESCENCE;
    test.push_back(ev);
  }
  if (!ran_out) {
    test.insert(test.end(), postfix, postfix + n_postfix);
    if (!test.empty() && test.back().kind != DEMI_EXT_WAIT_QUIESCENCE) test.push_back(mk(DEMI_EXT_WAIT_QUIESCENCE, 0, 0, 0, 0));
  }
  *n_out = (uint32_t)test.size();
  if (test.size() > cap) return fail(nullptr, DEMI_ERR_CAPACITY, "demi_fuzzer_generate: %zu events, room for %u", test.size(), cap);
  std::copy(test.begin(), test.end(), out);
  return DEMI_OK;
}

// ---- flat experiment directory: raw little-endian arrays of the C-ABI records + meta.json
namespace {
bool write_file(const std::string& p, const void* data, size_t bytes) {
  FILE* f = fopen(p.c_str(), "wb"); if (!f) return false;
  const bool ok = !bytes || fwrite(data, 1, bytes, f) == bytes;
  return fclose(f) == 0 && ok;
}
long file_size(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 ? (long)st.st_size : -1; }
bool read_file(const std::string& p, void* data, size_t bytes) {
  FILE* f = fopen(p.c_str(), "rb"); if (!f) return false;
  const bool ok = !bytes || fread(data, 1, bytes, f) == bytes;
  fclose(f); return ok;
}
long json_int(const std::string& txt, const char* key, long dflt) {
  const std::string k = std::string("\"") + key + "\"";
  size_t p = txt.find(k); if (p == std::string::npos) return dflt;
  p = txt.find(':', p); if (p == std::string::npos) return dflt;
  return strtol(txt.c_str() + p + 1, nullptr, 10);
}
}  // namespace

extern "C" int32_t demi_experiment_save(const char* dir, const demi_experiment* e) {
  if (!dir || !e || !e->externals || !e->events) return fail(nullptr, DEMI_ERR_INVALID, "demi_experiment_save: null argument");
  mkdir(dir, 0777);
  const std::string d(dir);
  bool ok = write_file(d + "/externals.bin", e->externals, (size_t)e->n_externals * sizeof(demi_ext_event)) &&
            write_file(d + "/event_trace.bin", e->events, (size_t)e->n_events * sizeof(demi_event));
  if (ok && e->dep_parent) ok = write_file(d + "/dep_parent.bin", e->dep_parent, (size_t)e->n_nodes * 2);
  if (ok && e->mcs_mask) ok = write_file(d + "/mcs.bin", e->mcs_mask, (size_t)e->mask_words * 8);
  char meta[256];
  snprintf(meta, sizeof(meta), "{\n \"format\": \"demi_b200/1\",\n \"model\": %d,\n \"model_flags\": %u,\n \"violation\": %u\n}\n",
           e->model, e->model_flags, e->violation);
  if (ok) ok = write_file(d + "/meta.json", meta, strlen(meta));
  if (!ok) return fail(nullptr, DEMI_ERR_INVALID, "demi_experiment_save: cannot write under %s", dir);
  return DEMI_OK;
}

extern "C" int32_t demi_experiment_load(const char* dir, demi_experiment* e) {
  if (!dir || !e) return fail(nullptr, DEMI_ERR_INVALID, "demi_experiment_load: null argument");
  const std::string d(dir);
  const long b_ext = file_size(d + "/externals.bin"), b_ev = file_size(d + "/event_trace.bin"),
             b_par = file_size(d + "/dep_parent.bin"), b_mcs = file_size(d + "/mcs.bin"), b_meta = file_size(d + "/meta.json");
  if (b_ext < 0 || b_ev < 0 || b_meta < 0 || b_ext % 16 || b_ev % 16)
    return fail(nullptr, DEMI_ERR_INVALID, "demi_experiment_load: %s is not an experiment directory", dir);
  e->n_externals = (uint32_t)(b_ext / 16); e->n_events = (uint32_t)(b_ev / 16);
  e->n_nodes = b_par > 0 ? (uint32_t)(b_par / 2) : 0; e->mask_words = b_mcs > 0 ? (uint32_t)(b_mcs / 8) : 0;
  std::string meta((size_t)b_meta, '\0');
  if (!read_file(d + "/meta.json", &meta[0], (size_t)b_meta)) return fail(nullptr, DEMI_ERR_INVALID, "demi_experiment_load: unreadable meta.json");
  e->model = (int32_t)json_int(meta, "model", 0); e->model_flags = (uint32_t)json_int(meta, "model_flags", 0);
  e->violation = (uint32_t)json_int(meta, "violation", 0);
  // sizes only (a first call with null / too small buffers tells the caller what to allocate)
  if (!e->externals || !e->events || e->cap_externals < e->n_externals || e->cap_events < e->n_events ||
      (e->dep_parent && e->cap_nodes < e->n_nodes) || (e->mcs_mask && e->cap_mask_words < e->mask_words))
    return fail(nullptr, DEMI_ERR_CAPACITY, "demi_experiment_load: %u externals, %u events, %u nodes, %u mask words", e->n_externals,
                e->n_events, e->n_nodes, e->mask_words);
  bool ok = read_file(d + "/externals.bin", e->externals, (size_t)b_ext) && read_file(d + "/event_trace.bin", e->events, (size_t)b_ev);
  if (ok && e->dep_parent && e->n_nodes) ok = read_file(d + "/dep_parent.bin", e->dep_parent, (size_t)e->n_nodes * 2);
  if (ok && e->mcs_mask && e->mask_words) ok = read_file(d + "/mcs.bin", e->mcs_mask, (size_t)e->mask_words * 8);
  if (!ok) return fail(nullptr, DEMI_ERR_INVALID, "demi_experiment_load: short read under %s", dir);
  return DEMI_OK;
}
