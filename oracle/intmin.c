/*
 * ORACLE — test infrastructure only (see machine.h header).  PARITY UNPINNED
 * against a running reference; pinned by derivable KATs only.
 *
 * Internal-event minimization: STSSchedMinimizer
 * (minification/internal_minimization/ScheduleCheckers.scala:19-107) driving
 * LeftToRightOneAtATime (OneAtATimeRemoval.scala:17-137) or SrcDstFIFORemoval (:139-251), strictly sequential.
 * Each step removes ONE delivery (UniqueMsgEvent) from the last failing trace and
 * asks STSSched whether the violation still shows (RunnerUtils.testWithStsSched,
 * RunnerUtils.scala:913-943); on success the trace STSSched recorded becomes the
 * new last failing trace.
 */
#include <stdlib.h>
#include <string.h>
#include "machine.h"
uint32_t oracle_ir_external_mask(void);
#include "sts.h"

typedef struct { uint8_t src, dst, type; uint32_t p0, p1; uint32_t count; } mkey;
typedef struct { mkey* k; uint32_t n, cap; } mset;

static mkey* mset_find(mset* s, const demi_event* e) {
  for (uint32_t i = 0; i < s->n; i++)
    if (s->k[i].src == e->src && s->k[i].dst == e->dst && s->k[i].type == e->type && s->k[i].p0 == e->p0 && s->k[i].p1 == e->p1)
      return &s->k[i];
  return 0;
}
static uint32_t mset_count(mset* s, const demi_event* e) { mkey* k = mset_find(s, e); return k ? k->count : 0; }
static void mset_add(mset* s, const demi_event* e, uint32_t c) {
  mkey* k = mset_find(s, e);
  if (k) { k->count += c; return; }
  if (s->n == s->cap) { s->cap = s->cap ? s->cap * 2 : 64; s->k = (mkey*)realloc(s->k, sizeof(mkey) * s->cap); }
  mkey* q = &s->k[s->n++];
  q->src = e->src; q->dst = e->dst; q->type = e->type; q->p0 = e->p0; q->p1 = e->p1; q->count = c;
}

/* SrcDstFIFORemoval (OneAtATimeRemoval.scala:139-251): per (snd,rcv) pair the fingerprints of its deliveries in
 * the verified trace; only the LAST one of a pair is eligible, timers always are */
#define FP_PAIRS (DEMI_MAX_ACTORS * DEMI_MAX_ACTORS)
typedef struct { uint8_t type; uint32_t p0, p1; } fprint;
typedef struct {
  int enabled;
  fprint* msgs[FP_PAIRS]; uint32_t len[FP_PAIRS]; uint8_t present[FP_PAIRS];   /* srcDstToMessages */
  int32_t cur_idx[FP_PAIRS];                                                     /* srcDstToCurrentIdx */
  int32_t prev;                                                                  /* previouslyChosenSrcDst, -1 = None */
  const demi_event* verified; uint32_t n_verified;
} fifo_strategy;

static void fifo_push(fifo_strategy* f, uint32_t pair, const demi_event* e, int prepend) {
  f->msgs[pair] = (fprint*)realloc(f->msgs[pair], sizeof(fprint) * (f->len[pair] + 1));
  fprint q; q.type = e->type; q.p0 = e->p0; q.p1 = e->p1;
  if (prepend) { memmove(f->msgs[pair] + 1, f->msgs[pair], sizeof(fprint) * f->len[pair]); f->msgs[pair][0] = q; }
  else f->msgs[pair][f->len[pair]] = q;
  f->len[pair]++; f->present[pair] = 1;
}
static void fifo_init(fifo_strategy* f, const demi_event* verified, uint32_t n) {          /* :152-159 */
  memset(f, 0, sizeof(*f));
  f->enabled = 1; f->prev = -1; f->verified = verified; f->n_verified = n;
  for (uint32_t i = 0; i < n; i++)
    if (verified[i].kind == DEMI_EV_MSG_EVENT && verified[i].src < DEMI_MAX_ACTORS)
      fifo_push(f, (uint32_t)verified[i].src * DEMI_MAX_ACTORS + verified[i].dst, &verified[i], 0);
}
static void fifo_free(fifo_strategy* f) { for (int i = 0; i < FP_PAIRS; i++) free(f->msgs[i]); }
/* choiceFilter (:178-203) */
static int fifo_choice(fifo_strategy* f, const demi_event* e) {
  if (e->src < DEMI_MAX_ACTORS) {
    uint32_t pair = (uint32_t)e->src * DEMI_MAX_ACTORS + e->dst;
    if (f->present[pair]) {
      int32_t idx = ++f->cur_idx[pair];
      if (idx == (int32_t)f->len[pair] - 1) {
        f->len[pair]--;                                                  /* dropRight(1) */
        if (!f->len[pair]) f->present[pair] = 0;
        f->prev = (int32_t)pair;
        return 1;
      }
    }
  }
  f->prev = -1;
  return e->src >= DEMI_MAX_ACTORS;                                      /* snd == "deadLetters": a timer */
}
/* the part of SrcDstFIFORemoval.getNextTrace that precedes super.getNextTrace (:209-249) */
static void fifo_before_next(fifo_strategy* f, const mset* already_removed, int triggered) {
  if (!triggered && f->prev >= 0) { f->present[f->prev] = 0; f->len[f->prev] = 0; }          /* "this src,dst is done" */
  if (triggered) {
    for (int i = 0; i < FP_PAIRS; i++) { f->present[i] = 0; f->len[i] = 0; }
    mset copy = {0, 0, 0};
    for (uint32_t i = 0; i < already_removed->n; i++) {
      demi_event t; t.src = already_removed->k[i].src; t.dst = already_removed->k[i].dst; t.type = already_removed->k[i].type;
      t.p0 = already_removed->k[i].p0; t.p1 = already_removed->k[i].p1;
      mset_add(&copy, &t, already_removed->k[i].count);
    }
    for (uint32_t i = f->n_verified; i-- > 0;) {                          /* reverse order, prepend */
      const demi_event* e = &f->verified[i];
      if (e->kind != DEMI_EV_MSG_EVENT || e->src >= DEMI_MAX_ACTORS) continue;
      mkey* k = mset_find(&copy, e);
      if (k && k->count) { k->count--; continue; }
      fifo_push(f, (uint32_t)e->src * DEMI_MAX_ACTORS + e->dst, e, 1);
    }
    free(copy.k);
  }
  for (int i = 0; i < FP_PAIRS; i++) f->cur_idx[i] = -1;                  /* resetSrcDstToCurrentIdx */
}

/* OneAtATimeStrategy.getNextTrace (OneAtATimeRemoval.scala:57-124): index of the delivery to drop next, or -1 */
static int next_to_ignore(const demi_event* ev, uint32_t n, mset* tried, const mset* already_removed,
                          fifo_strategy* fifo, int triggered) {
  if (fifo && fifo->enabled) fifo_before_next(fifo, already_removed, triggered);
  mset keys = {0, 0, 0};
  for (uint32_t i = 0; i < already_removed->n; i++) {
    demi_event t; t.src = already_removed->k[i].src; t.dst = already_removed->k[i].dst; t.type = already_removed->k[i].type;
    t.p0 = already_removed->k[i].p0; t.p1 = already_removed->k[i].p1;
    mset_add(&keys, &t, already_removed->k[i].count);                 /* keysThisIteration ++= alreadyRemoved :66 */
  }
  int found = -1;
  for (uint32_t i = 0; i < n && found < 0; i++) {
    if (ev[i].kind != DEMI_EV_MSG_EVENT) continue;
    mset_add(&keys, &ev[i], 1);                                       /* checkDelivery :71-93 */
    if (mset_count(&keys, &ev[i]) > mset_count(tried, &ev[i]) &&
        (!(fifo && fifo->enabled) || fifo_choice(fifo, &ev[i]))) {    /* choiceFilter: LeftToRightOneAtATime = true (:131-137) */
      mset_add(tried, &ev[i], 1);
      found = (int)i;
    }
  }
  free(keys.k);
  return found;
}

int oracle_internal_minimize(const demi_config* cfg, const demi_event* verified, uint32_t n_verified,
                             const demi_ext_event* mcs_ext, uint32_t n_ext, uint32_t looking_for, uint32_t flags_in,
                             demi_event* out_trace, uint32_t cap_out, uint32_t* n_out,
                             uint32_t* total_replays, uint32_t* internal_sizes, uint32_t cap_sizes, uint32_t* n_sizes,
                             uint32_t* unignorable) {
  const uint32_t ext_mask = cfg->model == 100 ? oracle_ir_external_mask() : demi_external_type_mask(cfg->model);
  const uint32_t flags = flags_in & ~DEMI_IM_SRC_DST_FIFO;
  fifo_strategy* fifo = 0;
  if (flags_in & DEMI_IM_SRC_DST_FIFO) { fifo = (fifo_strategy*)malloc(sizeof(fifo_strategy)); fifo_init(fifo, verified, n_verified); }
  int triggered = 0;                                                  /* violationTriggered (ScheduleCheckers.scala:48) */
  demi_event* cur = (demi_event*)malloc(sizeof(demi_event) * (n_verified + 1));
  demi_event* rec = (demi_event*)malloc(sizeof(demi_event) * 65536);
  memcpy(cur, verified, sizeof(demi_event) * n_verified);
  uint32_t n_cur = n_verified;
  mset tried = {0, 0, 0}, pruned = {0, 0, 0};
  /* OneAtATimeStrategy.init (:27-48): external deliveries are never ignored */
  for (uint32_t i = 0; i < n_verified; i++)
    if (verified[i].kind == DEMI_EV_MSG_EVENT && ((ext_mask >> (verified[i].type & 31)) & 1u)) mset_add(&tried, &verified[i], 1);
  uint32_t unig = 0;
  for (uint32_t i = 0; i < tried.n; i++) unig += tried.k[i].count;
  uint64_t full[64]; memset(full, 0, sizeof(full));
  for (uint32_t i = 0; i < n_ext; i++) full[i >> 6] |= 1ull << (i & 63);
  void* scratch = malloc(oracle_sts_scratch_size());
  uint32_t replays = 0, ns = 0, last_size = 0;
  for (uint32_t i = 0; i < n_cur; i++) last_size += cur[i].kind == DEMI_EV_MSG_EVENT;
  int rc = 0;
  for (;;) {
    int skip = next_to_ignore(cur, n_cur, &tried, &pruned, fifo, triggered);
    if (skip < 0) break;
    /* nextTrace = cur minus that one delivery; STSScheduler(nextTrace).test(mcs) */
    demi_replay_input in;
    uint32_t n_send_ev = 0, n_ext_sends = 0;
    for (uint32_t i = 0; i < n_cur; i++) n_send_ev += cur[i].kind == DEMI_EV_MSG_SEND;
    for (uint32_t i = 0; i < n_ext; i++) n_ext_sends += mcs_ext[i].kind == DEMI_EXT_SEND;
    in.events = cur; in.n_events = n_cur; in.externals = mcs_ext; in.n_externals = n_ext; in.external_type_mask = ext_mask;
    in.pending_cap = demi_replay_pending_cap(n_send_ev); in.tosend_cap = demi_tosend_cap(n_ext_sends);
    demi_replay_result r; uint32_t n_rec = 0;
    oracle_sts_replay_ex(cfg, &in, full, looking_for, flags, (uint32_t)skip, &r, rec, 65536, &n_rec, scratch);
    replays++;                                                         /* stats.increment_replays (STSScheduler.scala:213-215) */
    if (r.status) { rc = -1; break; }
    triggered = r.violation != 0;
    if (r.violation) {
      /* prunedThisRun = deliveries(lastFailingTrace) - deliveries(new trace) (:69-84) */
      mset prior = {0, 0, 0}, fresh = {0, 0, 0};
      for (uint32_t i = 0; i < n_cur; i++) if (cur[i].kind == DEMI_EV_MSG_EVENT) mset_add(&prior, &cur[i], 1);
      uint32_t new_size = 0;
      for (uint32_t i = 0; i < n_rec; i++) if (rec[i].kind == DEMI_EV_MSG_EVENT) { mset_add(&fresh, &rec[i], 1); new_size++; }
      for (uint32_t i = 0; i < prior.n; i++) {                          /* MultiSet.setDifference (schedulers/Util.scala:93-106) */
        demi_event t; t.src = prior.k[i].src; t.dst = prior.k[i].dst; t.type = prior.k[i].type; t.p0 = prior.k[i].p0; t.p1 = prior.k[i].p1;
        uint32_t c2 = mset_count(&fresh, &t);
        if (prior.k[i].count > c2) mset_add(&pruned, &t, prior.k[i].count - c2);
      }
      free(prior.k); free(fresh.k);
      cur = (demi_event*)realloc(cur, sizeof(demi_event) * (n_rec + 1));
      memcpy(cur, rec, sizeof(demi_event) * n_rec);
      n_cur = n_rec;
      last_size = new_size;
    }
    if (internal_sizes && ns < cap_sizes) internal_sizes[ns] = last_size;   /* record_internal_size (:90, :95) */
    ns++;
  }
  if (n_out) *n_out = n_cur;
  if (out_trace) memcpy(out_trace, cur, sizeof(demi_event) * (n_cur < cap_out ? n_cur : cap_out));
  if (total_replays) *total_replays = replays;
  if (n_sizes) *n_sizes = ns;
  if (unignorable) *unignorable = unig;
  free(cur); free(rec); free(tried.k); free(pruned.k); free(scratch);
  if (fifo) { fifo_free(fifo); free(fifo); }
  return rc;
}

/* single recorded replay (verify_mcs's returned trace etc.) */
int oracle_sts_replay_trace(const demi_config* cfg, const demi_replay_input* in, const uint64_t* mask, uint32_t looking_for,
                            uint32_t flags, uint32_t skip_event, demi_replay_result* out,
                            demi_event* rec, uint32_t cap_rec, uint32_t* n_rec) {
  oracle_sts_replay_ex(cfg, in, mask, looking_for, flags, skip_event, out, rec, cap_rec, n_rec, 0);
  return out->status ? -1 : 0;
}
