/*
 * ORACLE — test infrastructure only (see machine.h).  PARITY UNPINNED against a running reference.
 *
 * ProvenanceTracker (schedulers/Util.scala:267-376), restated literally: build the first-order
 * happens-before pairs, topologically sort them, close the relation transitively with the
 * reverse-topological sweep the reference uses, then filter the trace.  The relation is a dense bit
 * matrix over Unique ids; the GPU path computes the same answer with a different algorithm
 * (per-vertex reachability masks, provenance_kernel.cuh), which is what makes the comparison a test.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "machine.h"
#include "provenance.h"

#define BIT(m, w, r, c) (((m)[(size_t)(r) * (w) + ((c) >> 6)] >> ((c) & 63)) & 1ull)
#define SET(m, w, r, c) ((m)[(size_t)(r) * (w) + ((c) >> 6)] |= 1ull << ((c) & 63))

/* DepTracker.initialTrace (DepTracker.scala:63, :126-129): the root, then every delivered Unique */
static uint32_t initial_trace(const demi_event* ev, uint32_t n_events, uint16_t* nodes, uint8_t* rcv, uint32_t cap) {
  uint32_t t = 0;
  if (cap) { nodes[0] = 0; rcv[0] = 0xFF; t = 1; }        /* MsgEvent("null","null",null), id 0 */
  for (uint32_t i = 0; i < n_events; i++)
    if (ev[i].kind == DEMI_EV_MSG_EVENT) {
      if (t >= cap) return 0xFFFFFFFFu;
      nodes[t] = ev[i].node; rcv[t] = ev[i].dst; t++;
    }
  return t;
}

int oracle_provenance(const demi_event* events, uint32_t n_events, const uint16_t* dep_parent, uint32_t n_nodes,
                      uint32_t affected_mask, uint64_t* keep_mask, uint32_t mask_words, demi_provenance_out* out) {
  memset(out, 0, sizeof(*out));
  memset(keep_mask, 0, (size_t)mask_words * 8);
  out->affected_mask = affected_mask;
  const uint32_t cap = mask_words * 64;
  uint16_t* tr = (uint16_t*)malloc((size_t)(cap + 1) * 2);
  uint8_t* rc = (uint8_t*)malloc(cap + 1);
  const uint32_t T = initial_trace(events, n_events, tr, rc, cap);
  if (T == 0xFFFFFFFFu || !T) { out->status = DEMI_PV_OVERFLOW; free(tr); free(rc); return 0; }
  out->n_trace = T;
  const uint32_t V = n_nodes, W = (V + 63) / 64;
  uint64_t* hb = (uint64_t*)calloc((size_t)V * W, 8);     /* happensBefore: bit (a, b) */
  for (uint32_t t = 0; t < T; t++) if (tr[t] >= V) { out->status = DEMI_PV_OVERFLOW; goto done; }

  /* first-order pairs (:289-304): every prior receive on the same machine INCLUDING the event itself
   * (`priorReceives += u` precedes the foreach), and every message sent as a result of the receive
   * (depGraph.get(u).inNeighbors: edges run child ~> parent, DepTracker.scala:111-116) */
  for (uint32_t t = 0; t < T; t++) {
    for (uint32_t p = 0; p <= t; p++) if (rc[p] == rc[t]) SET(hb, W, tr[p], tr[t]);
    for (uint32_t s = 1; s < V; s++) if (dep_parent[s] == tr[t]) SET(hb, W, tr[t], s);
  }
  {
    /* Util.topologicalSort on the pairs with u1 != u2 (:315, :501-517): repeatedly peel the vertices
     * with no remaining predecessor; a non-empty remainder is `sys.error` */
    uint8_t* in_graph = (uint8_t*)calloc(V, 1);
    uint32_t* npred = (uint32_t*)calloc(V, 4);
    uint32_t* order = (uint32_t*)malloc((size_t)V * 4);
    for (uint32_t a = 0; a < V; a++) for (uint32_t b = 0; b < V; b++)
      if (a != b && BIT(hb, W, a, b)) { in_graph[a] = in_graph[b] = 1; npred[b]++; }
    uint32_t n_sorted = 0, total = 0;
    for (uint32_t v = 0; v < V; v++) total += in_graph[v];
    uint8_t* done = (uint8_t*)calloc(V, 1);
    for (;;) {
      uint32_t first = n_sorted;
      for (uint32_t v = 0; v < V; v++) if (in_graph[v] && !done[v] && npred[v] == 0) order[n_sorted++] = v;
      if (n_sorted == first) break;
      for (uint32_t k = first; k < n_sorted; k++) {
        uint32_t v = order[k]; done[v] = 1;
        for (uint32_t b = 0; b < V; b++) if (b != v && BIT(hb, W, v, b)) npred[b]--;
      }
    }
    if (n_sorted != total) out->status = DEMI_PV_CYCLE;
    else {
      /* transitive closure (:325-347).  node2parents is built from the first-order pairs only */
      uint64_t* first_order = (uint64_t*)malloc((size_t)V * W * 8);
      memcpy(first_order, hb, (size_t)V * W * 8);
      uint64_t* succ = (uint64_t*)calloc((size_t)V * W, 8);
      for (uint32_t k = n_sorted; k-- > 0;) {
        uint32_t u = order[k];
        SET(succ, W, u, u);
        for (uint32_t p = 0; p < V; p++)
          if (BIT(first_order, W, p, u))
            for (uint32_t w = 0; w < W; w++) succ[(size_t)p * W + w] |= succ[(size_t)u * W + w];
        for (uint32_t w = 0; w < W; w++) hb[(size_t)u * W + w] |= succ[(size_t)u * W + w];
      }
      free(first_order); free(succ);
    }
    free(in_graph); free(npred); free(order); free(done);
  }
  if (!out->status) {
    /* pruneConcurrentEvents (:354-375) */
    uint32_t last[DEMI_MAX_ACTORS], n_last = 0;
    for (uint32_t a = 0; a < DEMI_MAX_ACTORS; a++) {
      if (!((affected_mask >> a) & 1u)) continue;
      for (uint32_t t = T; t-- > 0;) if (rc[t] == a) { last[n_last++] = tr[t]; break; }   /* findLastEventForNode */
    }
    for (uint32_t t = 0; t < T; t++) {
      int all = 1;                                        /* concurrentOrAfterAllLastEvents */
      for (uint32_t k = 0; k < n_last && all; k++) {
        int o_u = (int)BIT(hb, W, last[k], tr[t]), u_o = (int)BIT(hb, W, tr[t], last[k]);
        int concurrent = !(o_u || u_o);
        if (!(concurrent || o_u)) all = 0;
      }
      if (!all) { keep_mask[t >> 6] |= 1ull << (t & 63); out->n_kept++; }
    }
  }
done:
  free(hb); free(tr); free(rc);
  return 0;
}

int oracle_fuzz_provenance(const demi_config* cfg, const demi_ext_event* ext, uint32_t n_ext,
                           const demi_fuzz_params* p, int64_t seed, uint64_t* keep_mask, uint32_t mask_words,
                           demi_provenance_out* out) {
  om_machine* m = (om_machine*)malloc(sizeof(om_machine));
  demi_event* ev = (demi_event*)malloc((size_t)OM_MAX_EVENTS * sizeof(demi_event));
  uint16_t* par = (uint16_t*)malloc((size_t)OM_MAX_NODES * 2);
  demi_fuzz_result r;
  oracle_run_prefix(cfg, ext, n_ext, p, seed, &r, ev, OM_MAX_EVENTS, par, OM_MAX_NODES, m);
  if (r.status) {
    memset(out, 0, sizeof(*out)); memset(keep_mask, 0, (size_t)mask_words * 8);
    out->status = DEMI_PV_PREFIX_FAILED;
  } else {
    uint32_t affected = r.violation ? m->model->affected(m->states, m->model_flags, r.violation) : 0;
    oracle_provenance(ev, r.n_events, par, r.n_nodes, affected, keep_mask, mask_words, out);
    out->violation = r.violation;
  }
  free(m); free(ev); free(par);
  return 0;
}
