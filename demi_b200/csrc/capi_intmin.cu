// capi_intmin.cu — internal-event minimization: STSSchedMinimizer
// (minification/internal_minimization/ScheduleCheckers.scala:19-107) with the
// LeftToRightOneAtATime removal strategy (OneAtATimeRemoval.scala:17-137).
//
// The reference tries ONE delivery at a time: drop it from the last failing trace,
// ask STSSched whether the violation still shows, and on success continue from the
// trace STSSched recorded.  While removals keep failing, the candidates all refer to
// the same base trace, so they are independent: here the whole remaining left-to-right
// candidate list of the current base trace is evaluated in one K2 launch
// (demi_replay_batch_ex with skip_events) and then committed in the reference's order
// up to the first success; the successful one is re-run in recording mode to obtain
// the new base trace.  Decisions, counters and the final trace are the sequential ones.
#include <map>
#include <tuple>
#include "engine.hpp"

namespace {
typedef std::tuple<uint32_t, uint32_t, uint32_t, uint32_t, uint32_t> Key;   // (snd, rcv, type, p0, p1) == (snd, rcv, fingerprint)
typedef std::map<Key, uint32_t> MultiSet;
Key key_of(const demi_event& e) { return Key(e.src, e.dst, e.type, e.p0, e.p1); }
uint32_t count_of(const MultiSet& m, const Key& k) { auto it = m.find(k); return it == m.end() ? 0u : it->second; }
}

extern "C" int32_t demi_internal_minimize(demi_handle* h, uint32_t looking_for, uint32_t flags,
                                          demi_event* out_trace, uint32_t cap_events,
                                          uint32_t* internal_sizes, uint32_t cap_sizes, demi_intmin_out* out) {
  if (!h) return DEMI_ERR_INVALID;
  if (!out) return fail(h, DEMI_ERR_INVALID, "demi_internal_minimize: null output");
  if (h->trace_host.empty()) return fail(h, DEMI_ERR_STATE, "demi_set_trace has not been called");
  memset(out, 0, sizeof(*out));
  const uint32_t ext_mask = demi_external_type_mask(h->cfg.model);
  std::vector<demi_event> cur = h->trace_host;
  const std::vector<demi_ext_event> ext = h->trace_ext_host;
  const uint32_t mw = std::max<uint32_t>(1, ((uint32_t)ext.size() + 63) / 64);
  MultiSet tried, pruned;
  // OneAtATimeStrategy.init (:27-48): external deliveries are never ignored
  for (const demi_event& e : cur)
    if (e.kind == DEMI_EV_MSG_EVENT && ((ext_mask >> (e.type & 31)) & 1u)) tried[key_of(e)]++;
  for (auto& kv : tried) out->unignorable += kv.second;
  uint32_t last_size = 0;
  for (const demi_event& e : cur) last_size += e.kind == DEMI_EV_MSG_EVENT;
  out->deliveries_before = last_size;
  std::vector<uint32_t> sizes;
  std::vector<demi_event> rec(65536);

  for (;;) {
    // The candidate list getNextTrace would produce on this base trace if every test failed (:57-124)
    std::vector<uint32_t> cand;
    {
      MultiSet t = tried;
      for (;;) {
        MultiSet keys = pruned;                                       // keysThisIteration ++= alreadyRemoved
        int found = -1;
        for (uint32_t i = 0; i < cur.size() && found < 0; i++) {
          if (cur[i].kind != DEMI_EV_MSG_EVENT) continue;
          Key k = key_of(cur[i]);
          uint32_t c = ++keys[k];
          if (c > count_of(t, k)) { t[k]++; found = (int)i; }         // choiceFilter == true
        }
        if (found < 0) break;
        cand.push_back((uint32_t)found);
      }
    }
    if (cand.empty()) break;
    std::vector<demi_replay_result> res(cand.size());
    int32_t rc = demi_replay_batch_ex(h, nullptr, cand.data(), (uint32_t)cand.size(), mw, looking_for, flags, res.data());
    if (rc != DEMI_OK) return rc;
    out->replays_executed += (uint32_t)cand.size();
    out->batches++;
    bool advanced = false;
    for (size_t c = 0; c < cand.size(); c++) {
      if (res[c].status) return fail(h, DEMI_ERR_CAPACITY, "demi_internal_minimize: a replay reported status %u", (unsigned)res[c].status);
      out->total_replays++;                                           // stats.increment_replays
      tried[key_of(cur[cand[c]])]++;                                  // triedIgnoring += key (:84)
      if (!res[c].violation) { sizes.push_back(last_size); continue; }   // "Ignoring didn't work."
      // success: the trace STSSched recorded becomes lastFailingTrace (:57-91)
      uint32_t n_rec = 0; demi_replay_result r1;
      rc = demi_replay_trace(h, nullptr, mw, cand[c], looking_for, flags, rec.data(), (uint32_t)rec.size(), &n_rec, &r1);
      if (rc != DEMI_OK) return rc;
      if (r1.status || !r1.violation) return fail(h, DEMI_ERR_STATE, "demi_internal_minimize: recorded re-run disagrees with the batch");
      MultiSet prior, fresh;
      for (const demi_event& e : cur) if (e.kind == DEMI_EV_MSG_EVENT) prior[key_of(e)]++;
      uint32_t new_size = 0;
      for (uint32_t i = 0; i < n_rec; i++) if (rec[i].kind == DEMI_EV_MSG_EVENT) { fresh[key_of(rec[i])]++; new_size++; }
      for (auto& kv : prior) {                                        // MultiSet.setDifference (schedulers/Util.scala:93-106)
        uint32_t c2 = count_of(fresh, kv.first);
        if (kv.second > c2) pruned[kv.first] += kv.second - c2;
      }
      cur.assign(rec.begin(), rec.begin() + n_rec);
      last_size = new_size;
      sizes.push_back(last_size);
      rc = demi_set_trace(h, cur.data(), (uint32_t)cur.size(), ext.data(), (uint32_t)ext.size());
      if (rc != DEMI_OK) return rc;
      advanced = true;
      break;
    }
    if (!advanced) break;          // every remaining candidate was tried and failed
  }
  out->n_events = (uint32_t)cur.size();
  out->deliveries_after = last_size;
  out->n_internal_sizes = (uint32_t)sizes.size();
  if (out_trace) std::copy(cur.begin(), cur.begin() + std::min<size_t>(cur.size(), cap_events), out_trace);
  if (internal_sizes) for (uint32_t i = 0; i < cap_sizes && i < sizes.size(); i++) internal_sizes[i] = sizes[i];
  return DEMI_OK;
}
