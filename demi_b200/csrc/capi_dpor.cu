// capi_dpor.cu — C ABI for batched DPORwHeuristics searches (K3).
#include "dpor_kernel.cuh"
#include "engine.hpp"

using namespace demi;

typedef void (*dpor_fn)(const DporArgs);
struct DporVariant { int model; int bd; dpor_fn fn; size_t smem; int nq; };
template <class MODEL, int BD>
static DporVariant make_dv() {
  using M = DporMachine<MODEL, BD>;
  return DporVariant{MODEL::ID, BD, dpor_kernel<MODEL, BD>, (size_t)M::WORDS * BD * sizeof(uint32_t), M::NQ};
}
static const DporVariant* pick_dv(int model) {
  static const std::vector<DporVariant> v = { make_dv<PingPong3, 32>(), make_dv<Raft5, 32>(), make_dv<Bcast32, 32>() };
  for (const DporVariant& d : v) if (d.model == model) return &d;
  return nullptr;
}

extern "C" int32_t demi_dpor_batch(demi_handle* h, const demi_ext_event* ext, const uint32_t* ext_offsets, uint32_t n_searches,
                                   const demi_dpor_params* params, demi_dpor_result* results,
                                   demi_dpor_violation* viol, uint32_t cap_viol, uint64_t* hashes, uint32_t cap_hashes) {
  if (!h) return DEMI_ERR_INVALID;
  if (!ext || !ext_offsets || !params || !results) return fail(h, DEMI_ERR_INVALID, "demi_dpor_batch: null argument");
  if (n_searches == 0) return DEMI_OK;
  const demi_dpor_params& P = *params;
  if (P.node_cap < 2 || P.node_cap > (1u << 20)) return fail(h, DEMI_ERR_INVALID, "node_cap must be in [2, 2^20]");
  if (P.explored_slots < 2 || (P.explored_slots & (P.explored_slots - 1))) return fail(h, DEMI_ERR_INVALID, "explored_slots must be a power of two");
  if (P.max_messages < 0 || P.max_messages > 1022) return fail(h, DEMI_ERR_INVALID, "max_messages must be in [0, 1022] (setMaxMessagesToSchedule)");
  if (!P.heap_cap || !P.max_interleavings) return fail(h, DEMI_ERR_INVALID, "heap_cap / max_interleavings must be positive");
  if (P.max_interleavings >= (1u << 20)) return fail(h, DEMI_ERR_INVALID, "max_interleavings must be below 2^20 (backtrack key packing)");
  if (P.max_interleavings >= (1u << 20)) return fail(h, DEMI_ERR_INVALID, "max_interleavings must be below 2^20");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const DporVariant* dv = pick_dv(h->cfg.model);
  if (!dv) return fail(h, DEMI_ERR_INVALID, "no DPOR kernel for model %d", h->cfg.model);
  const uint32_t n_ext = ext_offsets[n_searches];
  const int n_actors = h->cfg.model == DEMI_MODEL_PINGPONG3 ? 3 : h->cfg.model == DEMI_MODEL_RAFT5 ? 5 : 32;
  for (uint32_t i = 0; i < n_ext; i++) {
    // "unsuported external event" (DPORwHeuristics.scala:710)
    if (ext[i].kind != DEMI_EXT_START && ext[i].kind != DEMI_EXT_SEND)
      return fail(h, DEMI_ERR_INVALID, "demi_dpor_batch: external %u: DPOR accepts Start and Send only", i);
    if (ext[i].a >= n_actors) return fail(h, DEMI_ERR_INVALID, "demi_dpor_batch: external %u names an unknown actor", i);
  }
  DporArgs a{};
  a.model_flags = h->cfg.model_flags; a.blocked_mask = h->cfg.blocked_mask; a.ignore_timers = h->cfg.ignore_timers;
  a.P = P; a.n_searches = n_searches; a.cap_viol = cap_viol; a.cap_hashes = cap_hashes;
  a.T1 = (uint32_t)P.max_messages + 2;
  a.child_slots = demi_pow2_at_least(2 * P.node_cap, 4, 1u << 22);
  const size_t S = n_searches, NI = (size_t)P.max_interleavings + 1;
  struct Buf { void** p; size_t bytes; int fill; };
  void *d_ext = 0, *d_off = 0, *d_res = 0, *d_viol = 0, *d_hash = 0, *d_nodes = 0, *d_child = 0, *d_q = 0, *d_ex = 0, *d_heap = 0,
       *d_tr = 0, *d_tl = 0, *d_cur = 0, *d_next = 0, *d_npos = 0, *d_scan = 0;
  Buf bufs[] = {
    {&d_ext, n_ext * sizeof(demi_ext_event), -1}, {&d_off, (S + 1) * sizeof(uint32_t), -1},
    {&d_res, S * sizeof(demi_dpor_result), 0}, {&d_viol, std::max<size_t>(S * cap_viol, 1) * sizeof(demi_dpor_violation), 0},
    {&d_hash, std::max<size_t>(S * cap_hashes, 1) * sizeof(uint64_t), 0},
    {&d_nodes, S * P.node_cap * sizeof(uint4), -1}, {&d_child, S * a.child_slots * sizeof(uint32_t), 0},
    {&d_q, S * dv->nq * DPOR_QCAP * sizeof(uint32_t), -1}, {&d_ex, S * P.explored_slots * sizeof(uint64_t), 0xFF},
    {&d_heap, S * P.heap_cap * sizeof(DporKey), -1}, {&d_tr, S * NI * a.T1 * sizeof(uint32_t), -1},
    {&d_tl, S * NI * sizeof(uint32_t), -1}, {&d_cur, S * a.T1 * sizeof(uint32_t), -1}, {&d_next, S * a.T1 * sizeof(uint32_t), -1},
    {&d_npos, S * P.node_cap * sizeof(uint32_t), 0}, {&d_scan, S * a.T1 * sizeof(uint32_t), -1},
  };
  cudaError_t e = cudaSuccess;
  size_t total = 0;
  for (Buf& b : bufs) {
    if (e != cudaSuccess) break;
    e = cudaMalloc(b.p, b.bytes); total += b.bytes;
    if (e == cudaSuccess && b.fill >= 0) e = cudaMemsetAsync(*b.p, b.fill, b.bytes, h->stream);
  }
  auto cleanup = [&]() { for (Buf& b : bufs) cudaFree(*b.p); };
  if (e != cudaSuccess) { cleanup(); return fail(h, DEMI_ERR_CUDA, "demi_dpor_batch: %s (%.1f MB of search state)", cudaGetErrorString(e), total / 1e6); }
  e = cudaMemcpyAsync(d_ext, ext, n_ext * sizeof(demi_ext_event), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_off, ext_offsets, (S + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream);
  a.ext = (const uint4*)d_ext; a.ext_offsets = (const uint32_t*)d_off; a.results = (demi_dpor_result*)d_res;
  a.viol = viol ? (demi_dpor_violation*)d_viol : nullptr; a.hashes = hashes ? (uint64_t*)d_hash : nullptr;
  a.nodes = (uint4*)d_nodes; a.child_hash = (uint32_t*)d_child; a.queues = (uint32_t*)d_q; a.explored = (uint64_t*)d_ex;
  a.heap = (DporKey*)d_heap; a.traces = (uint32_t*)d_tr; a.trace_len = (uint32_t*)d_tl;
  a.cur_trace = (uint32_t*)d_cur; a.next_trace = (uint32_t*)d_next;
  a.node_pos = (uint32_t*)d_npos; a.scan = (uint32_t*)d_scan;
  if (e == cudaSuccess) e = cudaFuncSetAttribute(dv->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dv->smem);
  if (e == cudaSuccess) e = cudaEventRecord(h->ev0, h->stream);
  if (e == cudaSuccess) {
    dv->fn<<<(n_searches + dv->bd - 1) / dv->bd, dv->bd, dv->smem, h->stream>>>(a);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaEventRecord(h->ev1, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(results, d_res, S * sizeof(demi_dpor_result), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && viol && cap_viol) e = cudaMemcpyAsync(viol, d_viol, S * cap_viol * sizeof(demi_dpor_violation), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && hashes && cap_hashes) e = cudaMemcpyAsync(hashes, d_hash, S * cap_hashes * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  float ms = 0;
  if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  cleanup();
  if (e != cudaSuccess) return fail(h, DEMI_ERR_CUDA, "demi_dpor_batch: %s", cudaGetErrorString(e));
  h->perf.kernel_ms = ms; h->perf.kernel_launches = 1;
  uint64_t il = 0, del = 0, vi = 0;
  for (size_t s = 0; s < S; s++) { il += results[s].interleavings; del += results[s].deliveries; vi += results[s].violations; }
  h->perf.prefixes = il; h->perf.deliveries = del; h->perf.violations = vi;
  return DEMI_OK;
}
