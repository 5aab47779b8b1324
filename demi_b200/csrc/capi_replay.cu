// capi_replay.cu — C ABI for STSSched replay batches (K2) and DDMin.
//
// DDMin (minification/DeltaDebugging.scala:27-109) is sequential by definition:
// every test depends on the outcome of the previous one.  Here the recursion is
// still walked in exactly that order (so the decisions, the MCS and the
// MinimizationStats counters are the sequential ones), but test outcomes come
// from a memo that is filled speculatively: whenever the walk needs a test that
// has not been evaluated, the possible futures of the current ddmin2 frame (and of
// the interference siblings waiting on the recursion stack) are expanded a few
// levels deep and evaluated in ONE kernel launch.
#include <map>
#include <deque>
#include "replay_kernel.cuh"
#include "engine.hpp"
#include "models/model_ir.cuh"

using namespace demi;

typedef void (*replay_fn)(const ReplayArgs);
struct ReplayVariant { int model; int bd; replay_fn fn; replay_fn fn_rec; size_t smem; int n_actors; };
template <class MODEL, int BD>
static ReplayVariant make_rv() {
  using M = ReplayMachine<MODEL, BD>;
  return ReplayVariant{MODEL::ID, BD, replay_lane_kernel<MODEL, BD, false>, replay_lane_kernel<MODEL, BD, true>,
                       (size_t)M::WORDS * BD * sizeof(uint32_t), MODEL::N_ACTORS};
}
static const ReplayVariant* pick_rv(int model) {
  static const std::vector<ReplayVariant> v = {
    make_rv<PingPong3, 256>(), make_rv<Raft5, 256>(), make_rv<Bcast32, 128>(), make_rv<IrModel, 128>(),
  };
  for (const ReplayVariant& r : v) if (r.model == model) return &r;
  return nullptr;
}

void demi_replay_free(demi_handle* h) {
  cudaFree(h->trace_dev); cudaFree(h->trace_ext_dev); cudaFree(h->ev_ordinal_dev); cudaFree(h->send_ext_index_dev);
  cudaFree(h->rp_table); cudaFree(h->rp_tosend); cudaFree(h->rp_pruned); cudaFree(h->rp_masks); cudaFree(h->rp_results);
  cudaFree(h->rp_counters);
}

extern "C" int32_t demi_set_trace(demi_handle* h, const demi_event* events, uint32_t n_events,
                                  const demi_ext_event* externals, uint32_t n_externals) {
  if (!h) return DEMI_ERR_INVALID;
  if (!events || !n_events || !externals || !n_externals)
    return fail(h, DEMI_ERR_INVALID, "demi_set_trace: empty trace or externals");   // assume(!original_trace.isEmpty)
  if (n_externals > 4096) return fail(h, DEMI_ERR_INVALID, "demi_set_trace: more than 4096 external events");
  if (n_events >= (1u << 31)) return fail(h, DEMI_ERR_INVALID, "demi_set_trace: trace too long");
  { int32_t vrc = demi_check_events(h, "demi_set_trace", events, n_events, 0); if (vrc != DEMI_OK) return vrc;
    vrc = demi_check_externals(h, "demi_set_trace", externals, n_externals); if (vrc != DEMI_OK) return vrc; }   // (rejects HardKill)
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  { int32_t mrc = demi_need_model(h); if (mrc != DEMI_OK) return mrc; }
  const uint32_t ext_mask = demi_ext_type_mask(h);
  // FIFO correspondence of external MsgSends and original Sends (EventTrace.scala:385-386, :419-436)
  std::vector<uint16_t> ordinal(n_events, 0xFFFF);
  std::map<uint32_t, uint16_t> uniq_to_ord;
  uint32_t ord = 0, max_uniq = 0, n_send_events = 0;
  for (uint32_t i = 0; i < n_events; i++) {
    const demi_event& e = events[i];
    if (e.kind < DEMI_EV_MSG_SEND || e.kind > DEMI_EV_QUIESCENCE)
      return fail(h, DEMI_ERR_INVALID, "demi_set_trace: event %u has unknown kind %u", i, e.kind);
    if (e.kind == DEMI_EV_MSG_SEND) {
      n_send_events++;
      if (e.uniq > max_uniq) max_uniq = e.uniq;
      if ((ext_mask >> (e.type & 31)) & 1u) {
        if (ord >= 0xFFFF) return fail(h, DEMI_ERR_INVALID, "demi_set_trace: too many external sends");
        ordinal[i] = (uint16_t)ord; uniq_to_ord[e.uniq] = (uint16_t)ord; ord++;
      }
    }
  }
  for (uint32_t i = 0; i < n_events; i++)
    if (events[i].kind == DEMI_EV_MSG_EVENT) {
      auto it = uniq_to_ord.find(events[i].uniq);
      if (it != uniq_to_ord.end()) ordinal[i] = it->second;
      if (events[i].uniq > max_uniq) max_uniq = events[i].uniq;
    }
  std::vector<uint16_t> sidx;
  for (uint32_t i = 0; i < n_externals; i++) if (externals[i].kind == DEMI_EXT_SEND) sidx.push_back((uint16_t)i);
  h->trace_host.assign(events, events + n_events);
  h->trace_ext_host.assign(externals, externals + n_externals);
  h->send_ext_index_host = sidx;
  h->conjoined.assign(n_externals, -1);                       // a new EventDag has no conjoined atoms
  h->trace_n_uniq = max_uniq + 1;
  h->trace_n_send_events = n_send_events;
  h->trace_n_ext_sends = (uint32_t)sidx.size();
  cudaFree(h->trace_dev); cudaFree(h->trace_ext_dev); cudaFree(h->ev_ordinal_dev); cudaFree(h->send_ext_index_dev);
  h->trace_dev = h->trace_ext_dev = nullptr; h->ev_ordinal_dev = h->send_ext_index_dev = nullptr;
  CUDA_TRY(h, cudaMalloc(&h->trace_dev, n_events * sizeof(demi_event)));
  CUDA_TRY(h, cudaMalloc(&h->trace_ext_dev, n_externals * sizeof(demi_ext_event)));
  CUDA_TRY(h, cudaMalloc(&h->ev_ordinal_dev, n_events * sizeof(uint16_t)));
  CUDA_TRY(h, cudaMalloc(&h->send_ext_index_dev, std::max<size_t>(sidx.size(), 1) * sizeof(uint16_t)));
  CUDA_TRY(h, cudaMemcpy(h->trace_dev, events, n_events * sizeof(demi_event), cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(h->trace_ext_dev, externals, n_externals * sizeof(demi_ext_event), cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(h->ev_ordinal_dev, ordinal.data(), n_events * sizeof(uint16_t), cudaMemcpyHostToDevice));
  if (!sidx.empty())
    CUDA_TRY(h, cudaMemcpy(h->send_ext_index_dev, sidx.data(), sidx.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  if (!h->rp_counters) CUDA_TRY(h, cudaMalloc(&h->rp_counters, 2 * sizeof(unsigned long long)));
  return DEMI_OK;
}

static int32_t launch_replay(demi_handle* h, const void* masks_dev, const void* skips_dev, uint32_t n_masks, uint32_t mask_words,
                             uint32_t looking_for, uint32_t flags, void* out_dev, void* stream,
                             demi_event* rec_dev, uint32_t rec_cap, uint32_t* rec_count_dev);

extern "C" int32_t demi_conjoin_atoms(demi_handle* h, uint32_t e1, uint32_t e2) {
  if (!h) return DEMI_ERR_INVALID;
  if (h->trace_ext_host.empty()) return fail(h, DEMI_ERR_STATE, "demi_set_trace has not been called");
  const uint32_t n = (uint32_t)h->trace_ext_host.size();
  if (e1 >= n || e2 >= n) return fail(h, DEMI_ERR_INVALID, "No such external event: %u", e1 >= n ? e1 : e2);   // :168-173
  if (e1 == e2) return fail(h, DEMI_ERR_INVALID, "demi_conjoin_atoms: an event cannot be conjoined with itself");
  if (h->conjoined[e1] >= 0 || h->conjoined[e2] >= 0)                          // the asserts at :174-175
    return fail(h, DEMI_ERR_STATE, "demi_conjoin_atoms: external %u is already conjoined", h->conjoined[e1] >= 0 ? e1 : e2);
  h->conjoined[e1] = (int32_t)e2; h->conjoined[e2] = (int32_t)e1;
  return DEMI_OK;
}

extern "C" int32_t demi_replay_batch_dev(demi_handle* h, const void* masks_dev, uint32_t n_masks, uint32_t mask_words,
                                         uint32_t looking_for, uint32_t flags, void* out_dev, void* stream) {
  if (!h) return DEMI_ERR_INVALID;
  if (!masks_dev) return fail(h, DEMI_ERR_INVALID, "demi_replay_batch_dev: null buffer");
  return launch_replay(h, masks_dev, nullptr, n_masks, mask_words, looking_for, flags, out_dev, stream, nullptr, 0, nullptr);
}

static int32_t launch_replay(demi_handle* h, const void* masks_dev, const void* skips_dev, uint32_t n_masks, uint32_t mask_words,
                             uint32_t looking_for, uint32_t flags, void* out_dev, void* stream,
                             demi_event* rec_dev, uint32_t rec_cap, uint32_t* rec_count_dev) {
  if (!h) return DEMI_ERR_INVALID;
  if (h->trace_host.empty()) return fail(h, DEMI_ERR_STATE, "demi_set_trace has not been called");
  if (!out_dev) return fail(h, DEMI_ERR_INVALID, "replay: null output buffer");
  const bool record = rec_dev != nullptr;
  const uint32_t n_ext = (uint32_t)h->trace_ext_host.size();
  if (masks_dev && mask_words * 64 < n_ext) return fail(h, DEMI_ERR_INVALID, "mask_words %u too small for %u externals", mask_words, n_ext);
  if (n_masks == 0) return DEMI_OK;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const ReplayVariant* rv = pick_rv(h->cfg.model);
  if (!rv) return fail(h, DEMI_ERR_INVALID, "no replay kernel for model %d", h->cfg.model);
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_TRY(h, cudaFuncSetAttribute(rv->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rv->smem));
  int bps = 0;
  CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, rv->fn, rv->bd, rv->smem));
  if (bps < 1) return fail(h, DEMI_ERR_CAPACITY, "replay kernel does not fit on an SM");
  uint64_t want = ((uint64_t)n_masks + rv->bd - 1) / rv->bd;
  int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)h->sm_count * bps));
  const uint64_t warps = (uint64_t)grid * (rv->bd / 32);

  ReplayArgs a{};
  a.model_flags = h->cfg.model_flags; a.blocked_mask = h->cfg.blocked_mask; a.ignore_timers = h->cfg.ignore_timers;
  a.looking_for = looking_for; a.flags = flags;
  a.events = (const uint4*)h->trace_dev; a.n_events = (uint32_t)h->trace_host.size();
  a.ev_ordinal = h->ev_ordinal_dev;
  a.ext = (const uint4*)h->trace_ext_dev; a.n_ext = n_ext;
  a.send_ext_index = h->send_ext_index_dev; a.n_sends = h->trace_n_ext_sends;
  a.external_type_mask = demi_ext_type_mask(h);
  a.n_uniq_words = (h->trace_n_uniq + 31) / 32;
  a.masks = (const uint64_t*)masks_dev; a.n_masks = n_masks; a.mask_words = mask_words;
  a.skips = (const uint32_t*)skips_dev;
  a.rec_events = rec_dev; a.rec_cap = rec_cap; a.rec_count = rec_count_dev;
  a.results = (demi_replay_result*)out_dev;
  a.pending_cap = demi_replay_pending_cap(h->trace_n_send_events);
  a.tosend_cap = demi_tosend_cap(h->trace_n_ext_sends);
  a.table_slots = demi_pow2_at_least(2 * a.pending_cap, 256, 1u << 16);
  int32_t rc;
  const size_t table_bytes = warps * a.table_slots * 32 * sizeof(uint4);
  const bool fresh_table = h->rp_table_bytes < table_bytes;
  if ((rc = ensure_bytes(h, &h->rp_table, &h->rp_table_bytes, table_bytes)) != DEMI_OK) return rc;
  if ((rc = ensure_bytes(h, &h->rp_tosend, &h->rp_tosend_bytes, warps * a.tosend_cap * 32 * sizeof(uint32_t))) != DEMI_OK) return rc;
  if (flags & DEMI_RF_FILTER_KNOWN_ABSENTS)
    if ((rc = ensure_bytes(h, &h->rp_pruned, &h->rp_pruned_bytes,
                           warps * (a.n_uniq_words + rv->n_actors) * 32 * sizeof(uint32_t))) != DEMI_OK) return rc;
  // Generation stamps: each launch gets a fresh range [gen_base, gen_base + tests per thread]; the table is
  // cleared only when it was (re)allocated, its geometry changed, or the 16-bit range is exhausted.
  const uint32_t per_thread = (uint32_t)(((uint64_t)n_masks + (uint64_t)grid * rv->bd - 1) / ((uint64_t)grid * rv->bd)) + 1;
  const uint64_t geometry = ((uint64_t)a.table_slots << 32) ^ warps;
  if (fresh_table || h->rp_gen_next == 0 || h->rp_table_geometry != geometry || h->rp_gen_next + per_thread >= 0xFFF0u || record) {
    CUDA_TRY(h, cudaMemsetAsync(h->rp_table, 0, table_bytes, s));
    h->rp_gen_next = 1;
    h->rp_table_geometry = geometry;
  }
  a.gen_base = h->rp_gen_next;
  h->rp_gen_next = record ? 0 : h->rp_gen_next + per_thread;   // the recording variant leaves list entries behind: clear next time
  CUDA_TRY(h, cudaMemsetAsync(h->rp_counters, 0, 2 * sizeof(unsigned long long), s));
  a.table = (uint4*)h->rp_table; a.tosend = (uint32_t*)h->rp_tosend;
  a.pruned = (flags & DEMI_RF_FILTER_KNOWN_ABSENTS) ? (uint32_t*)h->rp_pruned : nullptr;
  a.counters = h->rp_counters;
  if (record) {
    CUDA_TRY(h, cudaFuncSetAttribute(rv->fn_rec, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rv->smem));
    if (h->cfg.model == DEMI_MODEL_IR) CUDA_TRY(h, ir_bind(h->ir_dev, s));
    rv->fn_rec<<<1, rv->bd, rv->smem, s>>>(a);
  } else {
    if (h->cfg.model == DEMI_MODEL_IR) CUDA_TRY(h, ir_bind(h->ir_dev, s));
    rv->fn<<<grid, rv->bd, rv->smem, s>>>(a);
  }
  CUDA_TRY(h, cudaGetLastError());
  h->perf.kernel_launches = 1;
  h->perf.prefixes = n_masks;
  return DEMI_OK;
}

extern "C" int32_t demi_replay_batch_ex(demi_handle* h, const uint64_t* masks, const uint32_t* skip_events, uint32_t n_tests,
                                        uint32_t mask_words, uint32_t looking_for, uint32_t flags, demi_replay_result* out_host) {
  if (!h) return DEMI_ERR_INVALID;
  if (!out_host) return fail(h, DEMI_ERR_INVALID, "demi_replay_batch_ex: null output");
  if (n_tests == 0) return DEMI_OK;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  int32_t rc;
  const size_t mbytes = masks ? (size_t)n_tests * mask_words * sizeof(uint64_t) : 0, rbytes = (size_t)n_tests * sizeof(demi_replay_result);
  const size_t sbytes = skip_events ? (size_t)n_tests * sizeof(uint32_t) : 0;
  if ((rc = ensure_bytes(h, &h->rp_masks, &h->rp_masks_bytes, std::max<size_t>(mbytes + sbytes, 8))) != DEMI_OK) return rc;
  if ((rc = ensure_bytes(h, &h->rp_results, &h->rp_results_bytes, rbytes)) != DEMI_OK) return rc;
  char* base = (char*)h->rp_masks;
  if (masks) CUDA_TRY(h, cudaMemcpyAsync(base, masks, mbytes, cudaMemcpyHostToDevice, h->stream));
  if (skip_events) CUDA_TRY(h, cudaMemcpyAsync(base + mbytes, skip_events, sbytes, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
  rc = launch_replay(h, masks ? base : nullptr, skip_events ? base + mbytes : nullptr, n_tests, mask_words, looking_for, flags,
                     h->rp_results, h->stream, nullptr, 0, nullptr);
  if (rc != DEMI_OK) return rc;
  CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(out_host, h->rp_results, rbytes, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  float ms = 0;
  CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->perf.kernel_ms = ms; h->perf.h2d_bytes = mbytes + sbytes; h->perf.d2h_bytes = rbytes;
  return DEMI_OK;
}

extern "C" int32_t demi_replay_trace(demi_handle* h, const uint64_t* mask, uint32_t mask_words, uint32_t skip_event,
                                     uint32_t looking_for, uint32_t flags,
                                     demi_event* events, uint32_t cap_events, uint32_t* n_events, demi_replay_result* result) {
  if (!h) return DEMI_ERR_INVALID;
  if (!events || !cap_events || !result) return fail(h, DEMI_ERR_INVALID, "demi_replay_trace: null buffer");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  void *d_mask = 0, *d_skip = 0, *d_rec = 0, *d_cnt = 0, *d_res = 0;
  cudaError_t e = cudaMalloc(&d_rec, (size_t)cap_events * sizeof(demi_event));
  if (e == cudaSuccess) e = cudaMalloc(&d_cnt, 4);
  if (e == cudaSuccess) e = cudaMalloc(&d_res, sizeof(demi_replay_result));
  if (e == cudaSuccess) e = cudaMalloc(&d_skip, 4);
  if (e == cudaSuccess && mask) e = cudaMalloc(&d_mask, (size_t)mask_words * 8);
  if (e == cudaSuccess && mask) e = cudaMemcpyAsync(d_mask, mask, (size_t)mask_words * 8, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_skip, &skip_event, 4, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt, 0, 4, h->stream);
  int32_t rc = DEMI_OK;
  if (e == cudaSuccess) rc = launch_replay(h, d_mask, d_skip, 1, mask_words, looking_for, flags, d_res, h->stream,
                                           (demi_event*)d_rec, cap_events, (uint32_t*)d_cnt);
  uint32_t cnt = 0;
  if (e == cudaSuccess && rc == DEMI_OK) e = cudaMemcpyAsync(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && rc == DEMI_OK) e = cudaMemcpyAsync(result, d_res, sizeof(demi_replay_result), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && rc == DEMI_OK) e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess && rc == DEMI_OK && cnt) e = cudaMemcpy(events, d_rec, (size_t)std::min(cnt, cap_events) * sizeof(demi_event), cudaMemcpyDeviceToHost);
  cudaFree(d_mask); cudaFree(d_skip); cudaFree(d_rec); cudaFree(d_cnt); cudaFree(d_res);
  if (e != cudaSuccess) return fail(h, DEMI_ERR_CUDA, "demi_replay_trace: %s", cudaGetErrorString(e));
  if (rc != DEMI_OK) return rc;
  if (n_events) *n_events = cnt;
  return DEMI_OK;
}

extern "C" int32_t demi_replay_batch(demi_handle* h, const uint64_t* masks, uint32_t n_masks, uint32_t mask_words,
                                     uint32_t looking_for, uint32_t flags, demi_replay_result* out_host) {
  if (!h) return DEMI_ERR_INVALID;
  if (!masks || !out_host) return fail(h, DEMI_ERR_INVALID, "demi_replay_batch: null buffer");
  if (n_masks == 0) return DEMI_OK;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  int32_t rc;
  const size_t mbytes = (size_t)n_masks * mask_words * sizeof(uint64_t), rbytes = (size_t)n_masks * sizeof(demi_replay_result);
  if ((rc = ensure_bytes(h, &h->rp_masks, &h->rp_masks_bytes, mbytes)) != DEMI_OK) return rc;
  if ((rc = ensure_bytes(h, &h->rp_results, &h->rp_results_bytes, rbytes)) != DEMI_OK) return rc;
  CUDA_TRY(h, cudaMemcpyAsync(h->rp_masks, masks, mbytes, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
  rc = demi_replay_batch_dev(h, h->rp_masks, n_masks, mask_words, looking_for, flags, h->rp_results, h->stream);
  if (rc != DEMI_OK) return rc;
  CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(out_host, h->rp_results, rbytes, cudaMemcpyDeviceToHost, h->stream));
  unsigned long long c[2] = {0, 0};
  CUDA_TRY(h, cudaMemcpyAsync(c, h->rp_counters, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  float ms = 0;
  CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->perf.kernel_ms = ms; h->perf.violations = c[0]; h->perf.deliveries = c[1];
  h->perf.h2d_bytes = mbytes; h->perf.d2h_bytes = rbytes + sizeof(c);
  return DEMI_OK;
}

// ---------------------------------------------------------------------- DDMin
#include "ddmin_driver.hpp"
namespace {

// STSSched as DDMin's TestOracle: a batch of masks = one replay launch
struct StsDDMinDriver : DDMinDriver {
  int32_t evaluate_flat(const uint64_t* masks, size_t n, signed char* out) override {   // the arena slice goes to the kernel as is
    std::vector<demi_replay_result> res(n);
    int32_t rc = demi_replay_batch(h, masks, (uint32_t)n, mw, looking_for, flags, res.data());
    if (rc != DEMI_OK) return rc;
    for (size_t i = 0; i < n; i++) {
      if (res[i].status != 0) return fail(h, DEMI_ERR_CAPACITY, "demi_ddmin: a replay reported status %u", (unsigned)res[i].status);
      out[i] = res[i].violation != 0;
    }
    return DEMI_OK;
  }
  int32_t evaluate_batch(const std::vector<Mask>& want, std::vector<char>& out) override {
    std::vector<uint64_t> flat(want.size() * mw);
    for (size_t i = 0; i < want.size(); i++) std::copy(want[i].begin(), want[i].end(), flat.begin() + i * mw);
    std::vector<demi_replay_result> res(want.size());
    int32_t rc = demi_replay_batch(h, flat.data(), (uint32_t)want.size(), mw, looking_for, flags, res.data());
    if (rc != DEMI_OK) return rc;
    for (size_t i = 0; i < want.size(); i++) {
      if (res[i].status != 0) return fail(h, DEMI_ERR_CAPACITY, "demi_ddmin: a replay reported status %u", (unsigned)res[i].status);
      out[i] = res[i].violation != 0;
    }
    return DEMI_OK;
  }
};

}  // namespace

extern "C" int32_t demi_ddmin(demi_handle* h, uint32_t looking_for, uint32_t flags, int32_t check_unmodified,
                              uint64_t* mcs_mask, uint32_t mask_words,
                              uint32_t* iteration_sizes, uint32_t cap_iterations, demi_ddmin_out* out) {
  if (!h) return DEMI_ERR_INVALID;
  if (!mcs_mask || !out) return fail(h, DEMI_ERR_INVALID, "demi_ddmin: null output");
  if (h->trace_host.empty()) return fail(h, DEMI_ERR_STATE, "demi_set_trace has not been called");
  const uint32_t n_ext = (uint32_t)h->trace_ext_host.size();
  if (mask_words * 64 < n_ext) return fail(h, DEMI_ERR_INVALID, "mask_words too small");
  memset(out, 0, sizeof(*out));
  StsDDMinDriver d;
  d.h = h; d.looking_for = looking_for; d.flags = flags; d.mw = mask_words; d.n_ext = n_ext;
  d.ext = h->trace_ext_host.data();
  d.conjoined = &h->conjoined;
  // tests per speculative batch: a launch costs ~3 ms of latency whatever its size (one test walks the whole trace),
  // building a batch ~0.2 us per test on the host; ~6000 tests = five to six levels of the decision tree per launch
  d.wide_cap = getenv("DEMI_DDMIN_WIDE") ? (size_t)atol(getenv("DEMI_DDMIN_WIDE")) : 6000;
  // STSSched ignores WaitQuiescence: drop them from the DAG (RunnerUtils.scala:678-684)
  Mask dag(mask_words, 0), zero(mask_words, 0);
  for (uint32_t i = 0; i < n_ext; i++) if (d.ext[i].kind != DEMI_EXT_WAIT_QUIESCENCE) DDMinDriver::setbit(dag, i);
  if (check_unmodified) {                                                 // DeltaDebugging.scala:41-47
    bool v = d.test(dag, dag, zero);
    if (d.error != DEMI_OK) return d.error;
    if (!v) return fail(h, DEMI_ERR_INVALID, "Unmodified trace does not trigger violation");
  }
  d.original_num_events = DDMinDriver::popcount(dag);
  Mask mcs = d.ddmin2(dag, zero);
  if (d.error != DEMI_OK) return d.error;
  d.iteration_sizes.push_back(d.original_num_events - d.total_inputs_pruned);     // fencepost :60
  // assert(original_num_events - total_inputs_pruned == mcs.length) :58
  if (d.original_num_events - d.total_inputs_pruned != DDMinDriver::popcount(mcs))
    return fail(h, DEMI_ERR_STATE, "demi_ddmin: bookkeeping assertion failed (DeltaDebugging.scala:58)");
  std::copy(mcs.begin(), mcs.end(), mcs_mask);
  // verify_mcs (DeltaDebugging.scala:64-71; RunnerUtils.scala:689-701)
  bool verified = d.test(mcs, mcs, zero);
  if (d.error != DEMI_OK) return d.error;
  out->mcs_size = DDMinDriver::popcount(mcs);
  out->reserved[0] = (uint32_t)d.spec_us; out->reserved[1] = (uint32_t)d.eval_us;    // host microseconds: building batches / evaluating them
  out->total_replays = d.total_replays;
  out->n_iterations = (uint32_t)d.iteration_sizes.size();
  out->replays_executed = d.replays_executed;
  out->batches = d.batches;
  out->verified = verified ? 1u : 0u;
  if (iteration_sizes)
    for (uint32_t i = 0; i < cap_iterations && i < out->n_iterations; i++) iteration_sizes[i] = d.iteration_sizes[i];
  return DEMI_OK;
}
