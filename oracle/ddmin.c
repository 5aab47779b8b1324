/*
 * ORACLE — test infrastructure only (see machine.h header).  PARITY UNPINNED
 * against a running reference; pinned by derivable KATs only.
 *
 * DDMin (minification/DeltaDebugging.scala:27-109) over EventDag masks
 * (minification/Util.scala:69-304), strictly sequential, with either the STS
 * oracle (sts.c) or a synthetic "violates iff mask ⊇ K" oracle as TestOracle.
 * A subsequence is a bitmask over the positions of original_externals.
 */
#include <stdlib.h>
#include <string.h>
#include "machine.h"
#include "sts.h"

#define MW_MAX 64   /* up to 4096 external events */

typedef struct { uint32_t first, second; } atom_t;   /* second == UINT32_MAX: single event */

typedef int (*ddmin_test_fn)(void* ctx, const uint64_t* mask);

typedef struct {
  const demi_ext_event* ext; uint32_t n_ext; uint32_t mw;
  ddmin_test_fn test; void* ctx;
  uint32_t original_num_events, total_inputs_pruned, total_replays;
  uint32_t* iter_sizes; uint32_t cap_iter, n_iter;
  /* optional log of every tested mask, in order */
  uint64_t* test_log; uint32_t cap_log, n_log;
  int error;
} ddmin_t;

static int bit(const uint64_t* m, uint32_t i) { return (int)((m[i >> 6] >> (i & 63)) & 1ull); }
static void setbit(uint64_t* m, uint32_t i) { m[i >> 6] |= 1ull << (i & 63); }
static uint32_t popcount_mask(const uint64_t* m, uint32_t mw) {
  uint32_t c = 0; for (uint32_t i = 0; i < mw; i++) c += (uint32_t)__builtin_popcountll(m[i]); return c;
}

/* MinificationUtil.split_list(l, 2) (minification/Util.scala:9-37): returns the
 * size of the first chunk; the first `remainder` chunks get one extra element. */
uint32_t oracle_split_first_len(uint32_t len, uint32_t ways, uint32_t which) {
  uint32_t interval = len / ways, rem = len % ways, start = 0, size = 0;
  for (uint32_t k = 0; k <= which; k++) {
    size = interval + (rem > 0 ? 1u : 0u);
    if (rem > 0) rem--;
    if (k < which) start += size;
  }
  (void)start;
  return size;
}

/* UnmodifiedEventDag.get_atomic_events (minification/Util.scala:197-265): pair
 * Start..Kill and Partition..UnPartition, everything else single; sorted by the
 * original index of the first element.  Returns -1 on "Kill without preceding
 * Start" / "UnPartition without preceding Partition". */
/* UnmodifiedEventDag.conjoinAtoms (minification/Util.scala:167-178): partner index per external, -1 = none.
 * Thread-local test state: set before a minimization, cleared (n = 0) after it. */
static __thread const int32_t* g_conj = 0;
static __thread uint32_t g_n_conj = 0;
void oracle_set_conjoined(const int32_t* partner, uint32_t n) { g_conj = n ? partner : 0; g_n_conj = n; }

int oracle_atomic_events(const demi_ext_event* ext, uint32_t n_ext, const uint64_t* mask, atom_t* atoms) {
  int32_t last_start[DEMI_MAX_ACTORS];
  int32_t last_part[DEMI_MAX_ACTORS][DEMI_MAX_ACTORS];
  for (int a = 0; a < DEMI_MAX_ACTORS; a++) { last_start[a] = -1; for (int b = 0; b < DEMI_MAX_ACTORS; b++) last_part[a][b] = -1; }
  uint32_t n = 0;
  /* "First deal with explicitly conjoined atoms" (:210-219): both halves must be present (the assert at :211);
   * canonical head = the lower original index (the reference takes HashSet.head) */
  for (uint32_t i = 0; i < n_ext && i < g_n_conj; i++) {
    if (!bit(mask, i) || g_conj[i] < 0) continue;
    if ((uint32_t)g_conj[i] >= n_ext || !bit(mask, (uint32_t)g_conj[i])) return -1;
    if ((uint32_t)g_conj[i] > i) { atoms[n].first = i; atoms[n].second = (uint32_t)g_conj[i]; n++; }
  }
  for (uint32_t i = 0; i < n_ext; i++) {
    if (!bit(mask, i)) continue;
    if (i < g_n_conj && g_conj[i] >= 0) continue;               /* filterNot(_conjoinedAtoms contains ...) :222 */
    const demi_ext_event* e = &ext[i];
    switch (e->kind) {
      case DEMI_EXT_KILL:
        if (last_start[e->a] < 0) return -1;
        atoms[n].first = (uint32_t)last_start[e->a]; atoms[n].second = i; n++;
        last_start[e->a] = -1;
        break;
      case DEMI_EXT_START: last_start[e->a] = (int32_t)i; break;
      case DEMI_EXT_PARTITION: last_part[e->a][e->b] = (int32_t)i; break;
      case DEMI_EXT_UNPARTITION:
        if (last_part[e->a][e->b] < 0) return -1;
        atoms[n].first = (uint32_t)last_part[e->a][e->b]; atoms[n].second = i; n++;
        last_part[e->a][e->b] = -1;
        break;
      default: atoms[n].first = i; atoms[n].second = 0xFFFFFFFFu; n++; break;
    }
  }
  for (int a = 0; a < DEMI_MAX_ACTORS; a++) {
    if (last_start[a] >= 0) { atoms[n].first = (uint32_t)last_start[a]; atoms[n].second = 0xFFFFFFFFu; n++; }
    for (int b = 0; b < DEMI_MAX_ACTORS; b++)
      if (last_part[a][b] >= 0) { atoms[n].first = (uint32_t)last_part[a][b]; atoms[n].second = 0xFFFFFFFFu; n++; }
  }
  /* sortBy first index (insertion sort: stable, n is small) */
  for (uint32_t i = 1; i < n; i++) {
    atom_t k = atoms[i]; uint32_t j = i;
    while (j > 0 && atoms[j - 1].first > k.first) { atoms[j] = atoms[j - 1]; j--; }
    atoms[j] = k;
  }
  return (int)n;
}

static int run_test(ddmin_t* d, const uint64_t* mask) {
  d->total_replays++;                                   /* stats.increment_replays (STSScheduler.scala:213-215) */
  if (d->test_log && d->n_log < d->cap_log) { memcpy(d->test_log + (size_t)d->n_log * d->mw, mask, d->mw * 8); }
  d->n_log++;
  return d->test(d->ctx, mask);
}
static void record_iteration(ddmin_t* d) {
  if (d->iter_sizes && d->n_iter < d->cap_iter) d->iter_sizes[d->n_iter] = d->original_num_events - d->total_inputs_pruned;
  d->n_iter++;
}

/* DDMin.ddmin2 (DeltaDebugging.scala:73-109) */
static void ddmin2(ddmin_t* d, const uint64_t* dag, const uint64_t* remainder, uint64_t* result) {
  const uint32_t mw = d->mw;
  atom_t* atoms = (atom_t*)malloc(sizeof(atom_t) * (d->n_ext + 1));
  int na = oracle_atomic_events(d->ext, d->n_ext, dag, atoms);
  if (na < 0) { d->error = 1; memcpy(result, dag, mw * 8); free(atoms); return; }
  if (na <= 1) { memcpy(result, dag, mw * 8); free(atoms); return; }       /* base case :74-77 */
  uint32_t n0 = oracle_split_first_len((uint32_t)na, 2, 0);
  uint64_t split[2][MW_MAX];
  memset(split, 0, sizeof(split));
  /* splits = [dag - chunk0, dag - chunk1].reverse = [events of chunk0, events of chunk1] (:82-85) */
  for (uint32_t i = 0; i < (uint32_t)na; i++) {
    uint64_t* s = split[i < n0 ? 0 : 1];
    setbit(s, atoms[i].first);
    if (atoms[i].second != 0xFFFFFFFFu) setbit(s, atoms[i].second);
  }
  free(atoms);
  uint32_t dag_len = popcount_mask(dag, mw);
  for (int k = 0; k < 2; k++) {                                                /* :88-102 */
    uint64_t un[MW_MAX];
    for (uint32_t w = 0; w < mw; w++) un[w] = split[k][w] | remainder[w];      /* EventDagView.union (Util.scala:287-293) */
    int violates = run_test(d, un);
    record_iteration(d);
    if (violates) {
      d->total_inputs_pruned += dag_len - popcount_mask(split[k], mw);
      ddmin2(d, split[k], remainder, result);
      return;
    }
  }
  /* interference :104-108 */
  uint64_t rem_l[MW_MAX], rem_r[MW_MAX], left[MW_MAX], right[MW_MAX];
  for (uint32_t w = 0; w < mw; w++) { rem_l[w] = split[1][w] | remainder[w]; rem_r[w] = split[0][w] | remainder[w]; }
  ddmin2(d, split[0], rem_l, left);
  ddmin2(d, split[1], rem_r, right);
  for (uint32_t w = 0; w < mw; w++) result[w] = left[w] | right[w];
}

/* DDMin.minimize (DeltaDebugging.scala:27-62).  Returns 0, -1 on
 * "Unmodified trace does not trigger violation", -2 on malformed atoms. */
static int ddmin_minimize(ddmin_t* d, const uint64_t* dag, int check_unmodified, uint64_t* mcs) {
  uint64_t zero[MW_MAX];
  memset(zero, 0, sizeof(zero));
  if (check_unmodified) {                                                     /* :41-47 */
    int v = d->test(d->ctx, dag);
    if (!v) return -1;
  }
  d->total_replays = 0; d->n_iter = 0; d->n_log = 0;                           /* _stats.reset() */
  d->original_num_events = popcount_mask(dag, d->mw);
  d->total_inputs_pruned = 0;
  ddmin2(d, dag, zero, mcs);
  if (d->error) return -2;
  record_iteration(d);                                                        /* fencepost :60 */
  return 0;
}

/* ------------------------------------------------------------ test oracles */
typedef struct { const uint64_t* K; uint32_t mw; } superset_ctx;
static int superset_test(void* ctx, const uint64_t* mask) {
  superset_ctx* c = (superset_ctx*)ctx;
  for (uint32_t w = 0; w < c->mw; w++) if ((mask[w] & c->K[w]) != c->K[w]) return 0;
  return 1;
}

typedef struct { const demi_config* cfg; const demi_replay_input* in; uint32_t looking_for, flags; void* scratch; } sts_ctx;
static int sts_test(void* ctx, const uint64_t* mask) {
  sts_ctx* c = (sts_ctx*)ctx;
  demi_replay_result r;
  oracle_sts_replay(c->cfg, c->in, mask, c->looking_for, c->flags, &r, c->scratch);
  return r.status == 0 && r.violation != 0;
}

/* the DAG DDMin starts from: all externals except WaitQuiescence (RunnerUtils.scala:678-684) */
static void initial_dag(const demi_ext_event* ext, uint32_t n_ext, uint64_t* dag, uint32_t mw) {
  memset(dag, 0, mw * 8);
  for (uint32_t i = 0; i < n_ext; i++) if (ext[i].kind != DEMI_EXT_WAIT_QUIESCENCE) setbit(dag, i);
}

int oracle_ddmin_superset(const demi_ext_event* ext, uint32_t n_ext, const uint64_t* K, uint32_t mw,
                          uint64_t* mcs, uint32_t* total_replays, uint32_t* iter_sizes, uint32_t cap_iter,
                          uint32_t* n_iter, uint64_t* test_log, uint32_t cap_log, uint32_t* n_log) {
  if (mw > MW_MAX) return -3;
  superset_ctx c = { K, mw };
  ddmin_t d; memset(&d, 0, sizeof(d));
  d.ext = ext; d.n_ext = n_ext; d.mw = mw; d.test = superset_test; d.ctx = &c;
  d.iter_sizes = iter_sizes; d.cap_iter = cap_iter; d.test_log = test_log; d.cap_log = cap_log;
  uint64_t dag[MW_MAX];
  initial_dag(ext, n_ext, dag, mw);
  int rc = ddmin_minimize(&d, dag, 1, mcs);
  if (total_replays) *total_replays = d.total_replays;
  if (n_iter) *n_iter = d.n_iter;
  if (n_log) *n_log = d.n_log;
  return rc;
}

int oracle_ddmin_sts(const demi_config* cfg, const demi_replay_input* in, uint32_t looking_for, uint32_t flags,
                     int check_unmodified, uint64_t* mcs, uint32_t mw, uint32_t* total_replays,
                     uint32_t* iter_sizes, uint32_t cap_iter, uint32_t* n_iter, int* verified) {
  if (mw > MW_MAX) return -3;
  sts_ctx c = { cfg, in, looking_for, flags, malloc(oracle_sts_scratch_size()) };
  ddmin_t d; memset(&d, 0, sizeof(d));
  d.ext = in->externals; d.n_ext = in->n_externals; d.mw = mw; d.test = sts_test; d.ctx = &c;
  d.iter_sizes = iter_sizes; d.cap_iter = cap_iter;
  uint64_t dag[MW_MAX];
  initial_dag(in->externals, in->n_externals, dag, mw);
  int rc = ddmin_minimize(&d, dag, check_unmodified, mcs);
  if (total_replays) *total_replays = d.total_replays;
  if (n_iter) *n_iter = d.n_iter;
  if (verified) *verified = (rc == 0) ? sts_test(&c, mcs) : 0;                /* verify_mcs (DeltaDebugging.scala:64-71) */
  free(c.scratch);
  return rc;
}

/* ------------------------------------------------ IncrementalDDMin over ResumableDPOR */
/* ResumableDPOR (IncrementalDeltaDebugging.scala:90-122): one DPORwHeuristics per external subsequence */
#include "dpor.h"
typedef struct {
  const demi_config* cfg; const demi_ext_event* ext; uint32_t n_ext, mw;
  const demi_dpor_params* P; const oracle_dpor_opts* opts;
  int32_t max_distance;
  uint64_t* keys; void** inst; uint32_t n_inst, cap_inst;
  uint64_t interleavings; uint32_t dpor_tests; int error;
} rdpor_ctx;
static int rdpor_test(void* ctx, const uint64_t* mask) {
  rdpor_ctx* c = (rdpor_ctx*)ctx;
  uint32_t slot = c->n_inst;
  for (uint32_t i = 0; i < c->n_inst; i++) if (!memcmp(c->keys + (size_t)i * c->mw, mask, c->mw * 8)) { slot = i; break; }
  if (slot == c->n_inst) {                                                    /* subseqToDPOR(events) = ctor() :110-113 */
    if (c->n_inst >= c->cap_inst) { c->error = 1; return 0; }
    demi_ext_event* sub = (demi_ext_event*)malloc(sizeof(demi_ext_event) * (c->n_ext + 1));
    uint32_t n = 0;
    for (uint32_t i = 0; i < c->n_ext; i++) if (bit(mask, i)) sub[n++] = c->ext[i];
    memcpy(c->keys + (size_t)slot * c->mw, mask, c->mw * 8);
    c->inst[slot] = oracle_dpor_open(c->cfg, sub, n, c->P, c->opts);
    free(sub);
    if (!c->inst[slot]) { c->error = 1; return 0; }
    c->n_inst++;
  }
  demi_dpor_result r;
  oracle_dpor_test(c->inst[slot], c->max_distance, &r, 0, 0, 0, 0);              /* setMaxDistance; test :115-116 */
  if (r.status) c->error = 1;
  c->interleavings += r.interleavings; c->dpor_tests++;
  return r.violations > 0;
}

/* IncrementalDDMin.minimize (IncrementalDeltaDebugging.scala:49-77).  `dag` = the externals to minimise
 * (Start/Send only).  iteration sizes are merged as mergeStats (:33-41) does: keyed by replay count. */
int oracle_incremental_ddmin(const demi_config* cfg, const demi_ext_event* ext, uint32_t n_ext,
                             const demi_dpor_params* P, const oracle_dpor_opts* opts,
                             int32_t max_max_distance, uint32_t stop_at_size, uint32_t cap_instances,
                             uint64_t* mcs, uint32_t mw, uint32_t* total_replays, uint32_t* rounds,
                             uint64_t* total_interleavings, uint32_t* n_instances, uint32_t* mcs_sizes, uint32_t cap_sizes) {
  if (mw > MW_MAX) return -3;
  rdpor_ctx c; memset(&c, 0, sizeof(c));
  c.cfg = cfg; c.ext = ext; c.n_ext = n_ext; c.mw = mw; c.P = P; c.opts = opts;
  c.cap_inst = cap_instances; c.keys = (uint64_t*)calloc((size_t)cap_instances * mw, 8); c.inst = (void**)calloc(cap_instances, sizeof(void*));
  uint64_t cur[MW_MAX], next[MW_MAX];
  initial_dag(ext, n_ext, cur, mw);
  int32_t dist = 0; uint32_t total = 0, nr = 0; int rc = 0;
  c.max_distance = dist;                                                        /* oracle.setMaxDistance(currentDistance) :53 */
  while (dist < max_max_distance && popcount_mask(cur, mw) > stop_at_size) {    /* :66 */
    ddmin_t d; memset(&d, 0, sizeof(d));
    d.ext = ext; d.n_ext = n_ext; d.mw = mw; d.test = rdpor_test; d.ctx = &c;
    rc = ddmin_minimize(&d, cur, 0, next);                                      /* new DDMin(oracle, checkUnmodifed=false) :68-69 */
    if (rc || c.error) break;
    memcpy(cur, next, mw * 8);
    total += d.total_replays;                                                   /* mergeStats :33-41 */
    if (mcs_sizes && nr < cap_sizes) mcs_sizes[nr] = popcount_mask(cur, mw);
    nr++;
    dist = dist == 0 ? 2 : dist << 1;                                           /* :72 */
    c.max_distance = dist;
  }
  memcpy(mcs, cur, mw * 8);
  if (total_replays) *total_replays = total;
  if (rounds) *rounds = nr;
  if (total_interleavings) *total_interleavings = c.interleavings;
  if (n_instances) *n_instances = c.n_inst;
  for (uint32_t i = 0; i < c.n_inst; i++) oracle_dpor_close(c.inst[i]);
  free(c.keys); free(c.inst);
  if (c.error) return -4;
  return rc;
}

/* batch of independent STS tests (CPU baseline for the replay workload) */
#include <pthread.h>
typedef struct { const demi_config* cfg; const demi_replay_input* in; const uint64_t* masks; uint32_t mw;
                 uint32_t looking_for, flags; demi_replay_result* out; uint32_t lo, hi; } rb_job;
static void* rb_worker(void* a) {
  rb_job* j = (rb_job*)a;
  void* scratch = malloc(oracle_sts_scratch_size());
  for (uint32_t i = j->lo; i < j->hi; i++)
    oracle_sts_replay(j->cfg, j->in, j->masks + (size_t)i * j->mw, j->looking_for, j->flags, &j->out[i], scratch);
  free(scratch);
  return 0;
}
int oracle_replay_batch(const demi_config* cfg, const demi_replay_input* in, const uint64_t* masks, uint32_t n_masks,
                        uint32_t mw, uint32_t looking_for, uint32_t flags, demi_replay_result* out, int threads) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  pthread_t tid[256]; rb_job jobs[256];
  uint32_t per = (n_masks + (uint32_t)threads - 1) / (uint32_t)threads;
  int started = 0;
  for (int t = 0; t < threads; t++) {
    uint32_t lo = per * (uint32_t)t, hi = lo + per;
    if (lo >= n_masks) break;
    if (hi > n_masks) hi = n_masks;
    jobs[t] = (rb_job){ cfg, in, masks, mw, looking_for, flags, out, lo, hi };
    pthread_create(&tid[t], 0, rb_worker, &jobs[t]);
    started++;
  }
  for (int t = 0; t < started; t++) pthread_join(tid[t], 0);
  return 0;
}
