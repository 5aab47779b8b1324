// provenance_kernel.cuh — ProvenanceTracker.pruneConcurrentEvents (schedulers/Util.scala:267-376) for a
// batch of recorded executions, one warp per execution.
//
// The reference materialises the whole happens-before relation as a set of pairs (and notes that it runs
// out of memory doing so).  The filter only ever asks, for the <= 32 "last delivery on an affected node"
// events o_a, whether u reaches o_a and whether o_a reaches u.  Both questions are answered for every
// vertex with one 32-bit mask each and two linear sweeps over the delivery order, because every first-order
// edge points forward in that order:
//   edges into u:  the previous delivery on u's receiver (earlier ones follow by transitivity) and u's
//                  parent in the DepTracker tree (the delivery that created the message);
//   D[u] (which o_a reach u)  = own bit | D[previous on receiver] | D[parent]      forward sweep
//   A[u] (which o_a u reaches) = own bit | A[later on receiver] | A[children]      reverse sweep, pushed
//                                                                                   into parent/previous
//   keep(u) = (A[u] & ~D[u]) != 0          == !(forall o: concurrent(o,u) || happensBefore(o,u))
// A Unique delivered twice with another delivery on the same receiver in between makes the relation
// cyclic, which is Util.topologicalSort's sys.error (Util.scala:506): reported as DEMI_PV_CYCLE.
#pragma once
#include "machine.cuh"

namespace demi {

struct ProvArgs {
  const demi_event* events; uint32_t ev_stride;     // [n][ev_stride] recorded EventTraces
  const uint16_t* parent; uint32_t par_stride;       // [n][par_stride] DepTracker trees (parent per node)
  const uint32_t* counts;                            // [n][4] {n_events, n_nodes, affectedNodes, violation}
  const demi_fuzz_result* results;                   // [n] or null: executions with a status are skipped
  uint32_t* scratch; uint32_t scratch_stride;        // [warps][T_cap + 3 * par_stride] words
  uint32_t t_cap;                                    // = mask_words * 64
  uint64_t* keep; uint32_t mask_words;               // [n][mask_words]
  demi_provenance_out* out;                          // [n]
  uint32_t n;
};

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
provenance_kernel(const __grid_constant__ ProvArgs args) {
  __shared__ uint32_t s_run[WARPS][36];              // per receiver (0..31 actors, 32 = "null"): running mask
  __shared__ uint32_t s_last[WARPS][36];             // per affected receiver: node id + 1 of its last delivery
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t gw = (uint64_t)blockIdx.x * WARPS + warp;
  const uint64_t total = (uint64_t)gridDim.x * WARPS;
  uint32_t* tr = args.scratch + gw * args.scratch_stride;           // position -> node | receiver << 16
  uint32_t* Am = tr + args.t_cap;
  uint32_t* Dm = Am + args.par_stride;
  uint32_t* Pm = Dm + args.par_stride;                              // node -> last position delivered + 1

  for (uint64_t it = gw; it < args.n; it += total) {
    const uint32_t* c = args.counts + it * 4;
    const uint32_t n_events = c[0], n_nodes = c[1], affected = c[2];
    const demi_event* ev = args.events + it * (uint64_t)args.ev_stride;
    const uint16_t* par = args.parent + it * (uint64_t)args.par_stride;
    uint64_t* keep = args.keep + it * (uint64_t)args.mask_words;
    uint32_t status = DEMI_PV_OK;
    if (args.results && args.results[it].status) status = DEMI_PV_PREFIX_FAILED;
    if (n_nodes > args.par_stride || n_events > args.ev_stride) status = status ? status : DEMI_PV_OVERFLOW;

    // DepTracker.initialTrace: root, then the delivered Uniques in order (warp-wide stream compaction)
    uint32_t T = 1;
    if (lane == 0) tr[0] = 0u | (32u << 16);
    if (!status) {
      for (uint32_t base = 0; base < n_events; base += 32) {
        const uint32_t i = base + lane;
        uint4 e = make_uint4(0, 0, 0, 0);
        if (i < n_events) e = __ldg(reinterpret_cast<const uint4*>(ev) + i);
        const bool is_delivery = i < n_events && (e.x & 0xFFu) == DEMI_EV_MSG_EVENT;
        const uint32_t b = __ballot_sync(FULL_MASK, is_delivery);
        const uint32_t pos = T + __popc(b & ((1u << lane) - 1u));
        const uint32_t node = e.w >> 16, rcv = (e.x >> 16) & 0xFFu;
        if (is_delivery) {
          if (pos < args.t_cap && node < n_nodes && rcv < 32u) tr[pos] = node | (rcv << 16);
          else status = DEMI_PV_OVERFLOW;
        }
        T += __popc(b);
      }
      status = __reduce_max_sync(FULL_MASK, status);
    }
    for (uint32_t i = lane; i < n_nodes && i < args.par_stride; i += 32) { Am[i] = 0; Dm[i] = 0; Pm[i] = 0; }
    for (uint32_t i = lane; i < 36; i += 32) { s_run[warp][i] = 0; s_last[warp][i] = 0; }
    __syncwarp();

    uint32_t n_kept = 0;
    if (!status) {
      // the sweeps are chains of dependent loads over a few hundred positions: lane 0 walks them
      if (lane == 0) {
        uint32_t seen = 0;
        // reverse sweep: the first delivery met on an affected receiver is findLastEventForNode (:357-363)
        for (uint32_t t = T; t-- > 0;) {
          const uint32_t x = tr[t], u = x & 0xFFFFu, r = x >> 16;
          uint32_t a = Am[u] | s_run[warp][r];
          if (r < 32u && ((affected >> r) & 1u) && !((seen >> r) & 1u)) { seen |= 1u << r; a |= 1u << r; s_last[warp][r] = u + 1; }
          Am[u] = a;
          s_run[warp][r] = a;
          if (u) Am[par[u]] |= a;
        }
        // forward sweep; a vertex is o_r from its first delivery on (a Unique may be delivered twice)
        uint32_t last_on[33];
#pragma unroll 1
        for (uint32_t r = 0; r < 33; r++) { s_run[warp][r] = 0; last_on[r] = 0; }
        for (uint32_t t = 0; t < T && !status; t++) {
          const uint32_t x = tr[t], u = x & 0xFFFFu, r = x >> 16;
          if (Pm[u] && last_on[r] != Pm[u]) status = DEMI_PV_CYCLE;     // another delivery on r since u's last
          uint32_t d = Dm[u] | s_run[warp][r] | (u ? Dm[par[u]] : 0u);
          if (r < 32u && s_last[warp][r] == u + 1) d |= 1u << r;
          Dm[u] = d;
          s_run[warp][r] = d;
          Pm[u] = t + 1; last_on[r] = t + 1;
        }
      }
      status = __shfl_sync(FULL_MASK, status, 0);
      __syncwarp();
    }
    // keep bits, 32 positions per ballot
    for (uint32_t base = 0; base < args.mask_words * 64; base += 32) {
      const uint32_t t = base + lane;
      bool k = false;
      if (!status && t < T) { const uint32_t u = tr[t] & 0xFFFFu; k = (Am[u] & ~Dm[u]) != 0u; }
      const uint32_t b = __ballot_sync(FULL_MASK, k);
      if (lane == 0) reinterpret_cast<uint32_t*>(keep)[base >> 5] = b;
      n_kept += __popc(b);
    }
    if (lane == 0) {
      demi_provenance_out o;
      o.status = status; o.violation = c[3]; o.affected_mask = affected;
      o.n_trace = status == DEMI_PV_OVERFLOW || status == DEMI_PV_PREFIX_FAILED ? 0u : T;
      o.n_kept = n_kept; o.reserved[0] = o.reserved[1] = o.reserved[2] = 0;
      args.out[it] = o;
    }
    __syncwarp();
  }
}

}  // namespace demi
