// ddmin_driver.hpp — DDMin (minification/DeltaDebugging.scala:27-109) as a sequential commit loop over a
// memo of test results that is filled speculatively in batches.  The decisions, the MCS and the counters are
// the sequential ones; what changes is how many tests are evaluated and when.  The test oracle is a hook:
// STSSched replays (capi_replay.cu) or resumable DPOR instances (capi_dpor.cu).
#pragma once
#include <algorithm>
#include <map>
#include <vector>
#include "engine.hpp"

namespace demi {

typedef std::vector<uint64_t> Mask;

struct Atom { uint32_t first, second; };

struct DDMinDriver {
  demi_handle* h;
  uint32_t looking_for, flags, mw, n_ext;
  const demi_ext_event* ext;
  std::map<Mask, bool> memo;           // mask -> violation reproduced?
  uint32_t original_num_events = 0, total_inputs_pruned = 0, total_replays = 0;
  std::vector<uint32_t> iteration_sizes;
  uint32_t replays_executed = 0, batches = 0;
  int32_t error = DEMI_OK;
  bool malformed = false;
  // interference siblings waiting on the recursion stack: (dag, remainder)
  std::vector<std::pair<Mask, Mask>> pending_siblings;
  static constexpr size_t BATCH_TARGET = 4096;
  // wide speculation (STSSched replays are cheap and a test's latency, not its cost, is what a minimisation waits
  // for): on a miss, the closure of the decision tree below the current frame is evaluated level by level up to
  // `wide_cap` tests — about one full wave of the replay kernel — so a whole DDMin run is one or two launches
  size_t wide_cap = 0;
  struct HalfCache { Mask hv[2]; size_t na; bool ok; };
  std::map<Mask, HalfCache> half_cache;        // halves() depends on the dag alone; a run meets ~2n distinct dags
  const HalfCache& halves_of(const Mask& dag) {
    auto it = half_cache.find(dag);
    if (it != half_cache.end()) return it->second;
    HalfCache c; c.ok = halves(dag, c.hv, c.na);
    return half_cache.emplace(dag, std::move(c)).first->second;
  }
  void expand_wide(const std::vector<std::pair<Mask, Mask>>& roots, std::vector<Mask>& want) {
    std::vector<std::pair<Mask, Mask>> cur = roots, next;
    while (!cur.empty() && want.size() < wide_cap) {
      next.clear();
      for (const auto& f : cur) {
        const HalfCache& c = halves_of(f.first);
        if (!c.ok || c.na <= 1) continue;
        Mask t0 = unite(c.hv[0], f.second), t1 = unite(c.hv[1], f.second);
        if (!memo.count(t0)) want.push_back(t0);
        if (!memo.count(t1)) want.push_back(t1);
        next.emplace_back(c.hv[0], f.second);                      // left half violates
        next.emplace_back(c.hv[1], f.second);                      // right half violates
        next.emplace_back(c.hv[0], t1);                            // interference: remainder grows by the sibling
        next.emplace_back(c.hv[1], t0);
      }
      if (want.size() + 2 * next.size() > 2 * wide_cap) break;      // the next level would not fit a wave
      cur.swap(next);
    }
  }

  static bool bit(const Mask& m, uint32_t i) { return (m[i >> 6] >> (i & 63)) & 1ull; }
  static void setbit(Mask& m, uint32_t i) { m[i >> 6] |= 1ull << (i & 63); }
  static uint32_t popcount(const Mask& m) { uint32_t c = 0; for (uint64_t w : m) c += (uint32_t)__builtin_popcountll(w); return c; }
  static Mask unite(const Mask& a, const Mask& b) { Mask r(a.size()); for (size_t i = 0; i < a.size(); i++) r[i] = a[i] | b[i]; return r; }

  // UnmodifiedEventDag.get_atomic_events (minification/Util.scala:197-265)
  // UnmodifiedEventDag.conjoinAtoms (minification/Util.scala:167-178): partner index per external, -1 = none
  const std::vector<int32_t>* conjoined = nullptr;
  bool is_conjoined(uint32_t i) const { return conjoined && i < conjoined->size() && (*conjoined)[i] >= 0; }
  bool atomic_events(const Mask& dag, std::vector<Atom>& atoms) const {
    std::vector<int32_t> last_start(DEMI_MAX_ACTORS, -1);
    std::map<std::pair<int, int>, int32_t> last_part;
    atoms.clear();
    // "First deal with explicitly conjoined atoms" (:210-219); both halves must be present (:211).  The reference
    // takes HashSet.head as the atom's first element; canonically the lower original index.
    for (uint32_t i = 0; i < n_ext; i++) {
      if (!bit(dag, i) || !is_conjoined(i)) continue;
      const uint32_t p = (uint32_t)(*conjoined)[i];
      if (p >= n_ext || !bit(dag, p)) return false;
      if (p > i) atoms.push_back({i, p});
    }
    for (uint32_t i = 0; i < n_ext; i++) {
      if (!bit(dag, i)) continue;
      if (is_conjoined(i)) continue;                              // filterNot(_conjoinedAtoms contains ...) :222
      const demi_ext_event& e = ext[i];
      switch (e.kind) {
        case DEMI_EXT_KILL:
          if (last_start[e.a] < 0) return false;                  // "Kill without preceding Start"
          atoms.push_back({(uint32_t)last_start[e.a], i}); last_start[e.a] = -1; break;
        case DEMI_EXT_START: last_start[e.a] = (int32_t)i; break;
        case DEMI_EXT_PARTITION: last_part[{e.a, e.b}] = (int32_t)i; break;
        case DEMI_EXT_UNPARTITION: {
          auto it = last_part.find({e.a, e.b});
          if (it == last_part.end() || it->second < 0) return false;   // "UnPartition without preceding Partition"
          atoms.push_back({(uint32_t)it->second, i}); it->second = -1; break;
        }
        default: atoms.push_back({i, 0xFFFFFFFFu}); break;
      }
    }
    for (int a = 0; a < DEMI_MAX_ACTORS; a++) if (last_start[a] >= 0) atoms.push_back({(uint32_t)last_start[a], 0xFFFFFFFFu});
    for (auto& kv : last_part) if (kv.second >= 0) atoms.push_back({(uint32_t)kv.second, 0xFFFFFFFFu});
    std::stable_sort(atoms.begin(), atoms.end(), [](const Atom& x, const Atom& y) { return x.first < y.first; });
    return true;
  }
  // MinificationUtil.split_list(atoms, 2) (minification/Util.scala:9-37) + remove_events:
  // halves[0] = events of the first chunk, halves[1] = events of the second.
  bool halves(const Mask& dag, Mask out[2], size_t& n_atoms) const {
    std::vector<Atom> atoms;
    if (!atomic_events(dag, atoms)) return false;
    n_atoms = atoms.size();
    size_t n0 = atoms.size() / 2 + (atoms.size() % 2 ? 1 : 0);
    out[0].assign(mw, 0); out[1].assign(mw, 0);
    for (size_t i = 0; i < atoms.size(); i++) {
      Mask& s = out[i < n0 ? 0 : 1];
      setbit(s, atoms[i].first);
      if (atoms[i].second != 0xFFFFFFFFu) setbit(s, atoms[i].second);
    }
    return true;
  }

  // speculative closure of the tests ddmin2(dag, rem) may issue, `depth` levels deep
  void expand(const Mask& dag, const Mask& rem, int depth, std::vector<Mask>& want) {
    if (want.size() >= 4 * BATCH_TARGET) return;
    Mask hv[2]; size_t na;
    if (!halves(dag, hv, na) || na <= 1) return;
    for (int k = 0; k < 2; k++) {
      Mask t = unite(hv[k], rem);
      if (!memo.count(t)) want.push_back(t);
    }
    if (depth <= 0) return;
    expand(hv[0], rem, depth - 1, want);                       // left half violates
    expand(hv[1], rem, depth - 1, want);                       // right half violates
    expand(hv[0], unite(hv[1], rem), depth - 1, want);         // interference
    expand(hv[1], unite(hv[0], rem), depth - 1, want);
  }

  // the oracle: fill `out[i]` = "test(want[i]) reproduces the violation" for the whole batch
  virtual int32_t evaluate_batch(const std::vector<Mask>& want, std::vector<char>& out) = 0;
  // the sequential walk consumed the result of test(m)
  virtual void consumed(const Mask&) {}
  virtual ~DDMinDriver() {}

  void evaluate(std::vector<Mask>& want) {
    std::sort(want.begin(), want.end());
    want.erase(std::unique(want.begin(), want.end()), want.end());
    if (want.empty()) return;
    std::vector<char> res(want.size(), 0);
    int32_t rc = evaluate_batch(want, res);
    if (rc != DEMI_OK) { error = rc; return; }
    for (size_t i = 0; i < want.size(); i++) memo[want[i]] = res[i] != 0;
    replays_executed += (uint32_t)want.size();
    batches++;
  }

  // TestOracle.test for the sequential walk
  bool test(const Mask& m, const Mask& cur_dag, const Mask& cur_rem) {
    auto it = memo.find(m);
    if (it == memo.end()) {
      std::vector<Mask> want;
      want.push_back(m);
      if (wide_cap) {
        std::vector<std::pair<Mask, Mask>> roots;
        roots.emplace_back(cur_dag, cur_rem);
        for (auto it2 = pending_siblings.rbegin(); it2 != pending_siblings.rend(); ++it2) roots.push_back(*it2);
        expand_wide(roots, want);
      } else {
      // depth: 4^d frames * 2 tests; stay near BATCH_TARGET
      int depth = 5;
      expand(cur_dag, cur_rem, depth, want);
      for (auto it2 = pending_siblings.rbegin(); it2 != pending_siblings.rend() && want.size() < BATCH_TARGET; ++it2)
        expand(it2->first, it2->second, 3, want);
      }
      evaluate(want);
      if (error != DEMI_OK) return false;
      it = memo.find(m);
    }
    const bool r = it->second;
    consumed(m);
    return r;
  }

  // DDMin.ddmin2 (DeltaDebugging.scala:73-109)
  Mask ddmin2(const Mask& dag, const Mask& rem) {
    if (error != DEMI_OK) return dag;
    Mask hv[2]; size_t na;
    if (!halves(dag, hv, na)) { malformed = true; error = fail(h, DEMI_ERR_INVALID, "demi_ddmin: Kill/UnPartition without preceding Start/Partition"); return dag; }
    if (na <= 1) return dag;                                             // base case :74-77
    const uint32_t dag_len = popcount(dag);
    for (int k = 0; k < 2; k++) {                                        // :88-102
      Mask t = unite(hv[k], rem);
      bool violates = test(t, dag, rem);
      if (error != DEMI_OK) return dag;
      total_replays++;
      iteration_sizes.push_back(original_num_events - total_inputs_pruned);
      if (violates) {
        total_inputs_pruned += dag_len - popcount(hv[k]);
        return ddmin2(hv[k], rem);
      }
    }
    // interference :104-108 — the right-hand recursion does not depend on the left-hand result
    Mask rem_l = unite(hv[1], rem), rem_r = unite(hv[0], rem);
    pending_siblings.push_back({hv[1], rem_r});
    Mask left = ddmin2(hv[0], rem_l);
    pending_siblings.pop_back();
    Mask right = ddmin2(hv[1], rem_r);
    return unite(left, right);
  }
};


}  // namespace demi
