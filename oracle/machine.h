/*
 * ORACLE — test infrastructure only.  Never linked into the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * Parity status: PARITY UNPINNED against a running reference.  DEMi has no
 * tests, fixtures or golden vectors and cannot be compiled
 * here (no JVM/sbt/Akka).  This file restates the reference algorithm with a
 * file:line citation on every function; it is pinned only by the derivable
 * known-answer checks of SURVEY.md §8c (java.util.Random spec values,
 * RandomizedHashSet swap-remove order, find_non_blocked_message re-append
 * order, split_list arithmetic, ddmin2 test order) in tests/test_oracle_kat.py.
 *
 * The scalar "DEMi machine": one RandomScheduler execution
 * (RandomScheduler.explore, schedulers/RandomScheduler.scala:234-272) over a
 * data-only actor model.
 */
#ifndef ORACLE_MACHINE_H
#define ORACLE_MACHINE_H

#include <stdint.h>
#include "../include/demi_b200.h"
#include "../include/demi_limits.h"
#include "jrandom.h"

#define OM_MAX_PENDING 8192
#define OM_MAX_TOSEND  1024
#define OM_MAX_NODES   65536
#define OM_MAX_EVENTS  65536
#define OM_MAX_STATE_WORDS 16

typedef struct om_machine om_machine;

typedef struct oracle_model {
  int id;
  int n_actors;
  int state_words;                      /* u32 words of state per actor */
  void (*init)(uint32_t* states, uint32_t flags);
  /* the application's receive(): state x message -> state', sends, timers */
  void (*receive)(om_machine* m, int self, uint32_t* st, const demi_msg* msg);
  /* TestOracle.Invariant (minification/TestOracle.scala:27) on actor states */
  uint32_t (*invariant)(const uint32_t* states, uint32_t flags);
  /* ViolationFingerprint.affectedNodes (TestOracle.scala:9-18) as an actor bitmask */
  uint32_t (*affected)(const uint32_t* states, uint32_t flags, uint32_t code);
} oracle_model;

const oracle_model* oracle_get_model(int id);

typedef struct { uint8_t dst, type; uint32_t p0, p1; } om_timer_key;

typedef struct { demi_msg msg; uint16_t uniq, node; } om_pending;

struct om_machine {
  const oracle_model* model;
  uint32_t model_flags;
  uint32_t blocked_mask;
  int ignore_timers;
  /* config */
  int32_t max_messages, interval;
  uint32_t looking_for;
  uint32_t pending_cap, tosend_cap, node_cap, event_cap;
  /* rng: FullyRandom(seed) -> RandomizedHashSet.rand (Util.scala:115) */
  jrandom rng;
  /* actor states */
  uint32_t states[DEMI_MAX_ACTORS * OM_MAX_STATE_WORDS];
  /* EventOrchestrator network state (EventOrchestrator.scala:51-59) */
  uint32_t inaccessible, killed;
  uint32_t dead;                            /* hard-killed and not started again: Instrumenter.receiverIsAlive is false */
  uint32_t partitioned[DEMI_MAX_ACTORS];   /* ordered pairs: bit b of row a */
  /* RandomizedHashSet.arr (Util.scala:112) */
  om_pending pending[OM_MAX_PENDING];
  uint32_t n_pending, max_pending;
  /* SrcDstFIFO (RandomScheduler.scala:702-909): `pending` then holds timersAndExternals (:712);
   * srcDsts (:704) is `pairs`, srcDstToMessages (:706) the per-pair FIFO lists below */
  int strategy;
  jrandom rng_pairs;                        /* SrcDstFIFO.rand (:705) */
  uint16_t pairs[1056]; uint32_t n_pairs;   /* pair code = src * 32 + dst */
  om_pending fifo_pool[OM_MAX_PENDING]; uint16_t fifo_next[OM_MAX_PENDING];
  uint16_t fifo_head[1056], fifo_tail[1056]; uint16_t fifo_free; uint32_t n_queued;
  /* ExternalEventInjector.messagesToSend (ExternalEventInjector.scala:109) */
  demi_msg tosend[OM_MAX_TOSEND];
  uint32_t n_tosend;
  /* RandomScheduler.justScheduledTimers / timersToResend (RandomScheduler.scala:109-113) */
  om_timer_key just[DEMI_TIMERSET_CAP]; uint32_t n_just;
  om_timer_key resend[DEMI_TIMERSET_CAP]; uint32_t n_resend;
  /* Instrumenter.timerToCancellable restricted to ongoing timers (Instrumenter.scala:136-142) */
  om_timer_key registry[DEMI_TIMERSET_CAP]; uint32_t n_registry;
  /* Instrumenter.timersCancelledThisStep (Instrumenter.scala:144) */
  om_timer_key cancelled[DEMI_TIMERSET_CAP]; uint32_t n_cancelled;
  /* DepTracker (DepTracker.scala:27-135): parent-pointer tree */
  demi_msg node_msg[OM_MAX_NODES];
  uint16_t node_parent[OM_MAX_NODES];
  /* children of each node in creation (= id) order: what `inNeighbors` of the parent holds (DepTracker.scala:94-101) */
  uint16_t node_first_child[OM_MAX_NODES], node_last_child[OM_MAX_NODES], node_next_sibling[OM_MAX_NODES];
  uint32_t n_nodes;
  uint32_t parent_event;       /* DepTracker.parentEvent */
  /* EventTrace.events (EventTrace.scala:20) */
  demi_event* events;          /* optional recording buffer (event_cap entries) */
  uint32_t n_events;
  uint64_t trace_hash;
  /* counters */
  uint32_t n_uniq;
  int32_t nsched;              /* messagesScheduledSoFar */
  uint32_t ext_idx;            /* EventOrchestrator.traceIdx */
  const demi_ext_event* ext; uint32_t n_ext;
  uint32_t violation;          /* violationFound */
  uint16_t status;
};

/* callbacks for model receive() */
void om_send(om_machine* m, int src, int dst, uint8_t type, uint32_t p0, uint32_t p1);
void om_schedule_once(om_machine* m, int self, uint8_t type, uint32_t p0, uint32_t p1);
void om_schedule_repeating(om_machine* m, int self, uint8_t type, uint32_t p0, uint32_t p1);
void om_cancel_timer(om_machine* m, int self, uint8_t type, uint32_t p0, uint32_t p1);

/* FullyRandom's userDefinedFilter as rules (include/demi_b200.h); thread-local test state, n = 0 clears */
void oracle_set_user_filter(const demi_filter_rule* rules, uint32_t n);

/* building blocks exposed for the known-answer tests */
void     om_pending_insert(om_machine* m, const om_pending* e);
om_pending om_pending_remove_random(om_machine* m);
int      om_find_non_blocked(om_machine* m, om_pending* out);

/* One full execution.  `events`/`dep_parent` may be NULL. */
void oracle_run_prefix(const demi_config* cfg, const demi_ext_event* ext, uint32_t n_ext,
                       const demi_fuzz_params* p, int64_t seed,
                       demi_fuzz_result* out,
                       demi_event* events, uint32_t cap_events,
                       uint16_t* dep_parent, uint32_t cap_nodes,
                       om_machine* scratch /* may be NULL: allocates */);

/* Batch driver over `threads` host threads (static block partition). */
int oracle_fuzz_batch(const demi_config* cfg, const demi_ext_event* ext, uint32_t n_ext,
                      const demi_fuzz_params* p, demi_fuzz_result* out, int threads);

#endif
