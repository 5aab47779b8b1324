/*
 * ORACLE — test infrastructure only (see machine.h header).  PARITY UNPINNED
 * against a running reference; pinned by derivable KATs only.
 *
 * STSScheduler.test (schedulers/STSScheduler.scala:199-310) restated:
 *   1. project the original EventTrace onto an external-event subsequence:
 *      EventTrace.subsequenceIntersection (EventTrace.scala:290-380) ->
 *      filterSends (:382-452) -> filterKnownAbsentInternals (:458-534);
 *   2. replay it: advanceReplay (STSScheduler.scala:405-559) skips expected
 *      deliveries that are not pending, schedule_new_message (:643-776) delivers
 *      the expected ones, event_produced (:561-623) collects what actors send;
 *   3. test the invariant on the final state (:278-300).
 * ReplayScheduler's strict mode (schedulers/ReplayScheduler.scala:256-342): the
 * same walk, but an expected delivery that is not pending is a divergence.
 */
#include <stdlib.h>
#include <string.h>
#include "machine.h"
#include "sts.h"

/* ------------------------------------------------------------ projection */
int oracle_sts_project(const demi_replay_input* in, const uint64_t* mask, int filter_known_absents,
                       uint8_t* keep /* n_events */) {
  const demi_event* ev = in->events;
  const demi_ext_event* ext = in->externals;
  const uint32_t n_ev = in->n_events, n_ext = in->n_externals;
#define IN_MASK(i) ((mask[(i) >> 6] >> ((i) & 63)) & 1ull)
  /* remaining = subseq minus Sends, in order (EventTrace.scala:299-302) */
  uint32_t* rem = (uint32_t*)malloc(sizeof(uint32_t) * (n_ext + 1));
  uint32_t n_rem = 0, rp = 0;
  for (uint32_t i = 0; i < n_ext; i++)
    if (IN_MASK(i) && ext[i].kind != DEMI_EXT_SEND) rem[n_rem++] = i;
  /* missing_indices over original Sends (EventTrace.scala:398-416) */
  uint8_t* missing = (uint8_t*)calloc(n_ext + 1, 1);
  uint32_t n_sends = 0;
  for (uint32_t i = 0; i < n_ext; i++)
    if (ext[i].kind == DEMI_EXT_SEND) { missing[n_sends] = IN_MASK(i) ? 0 : 1; n_sends++; }
  uint8_t* pruned_ids = (uint8_t*)calloc(65536, 1);       /* filterSends: pruned_msg_ids */
  uint8_t* pruned_sends = (uint8_t*)calloc(65536, 1);     /* filterKnownAbsentInternals: prunedMessageSends */
  uint32_t alive = 0;                                     /* actorToAlive; deadLetters/Timer always alive */
  uint32_t partflag[DEMI_MAX_ACTORS];                     /* actorsToPartitioned((a,b)) == true bits */
  memset(partflag, 0, sizeof(partflag));
  int32_t msg_send_idx = -1;
  uint32_t kept = 0;

  for (uint32_t i = 0; i < n_ev; i++) {
    const demi_event* e = &ev[i];
    int is_msg = e->kind == DEMI_EV_MSG_SEND || e->kind == DEMI_EV_MSG_EVENT;
    int k = 0;
    /* ---- pass 1: subsequenceIntersection main loop (:307-374) */
    if (rp >= n_rem) {
      /* remaining.isEmpty: keep message events and non-external events (:308-314) */
      if (is_msg) k = 1;
      else k = !(e->kind == DEMI_EV_KILL || e->kind == DEMI_EV_SPAWN ||
                 e->kind == DEMI_EV_PARTITION || e->kind == DEMI_EV_UNPARTITION);
    } else {
      const demi_ext_event* h = &ext[rem[rp]];
      switch (e->kind) {
        case DEMI_EV_KILL:
          if (h->kind == DEMI_EXT_KILL && h->a == e->dst) { k = 1; rp++; }
          break;
        case DEMI_EV_PARTITION:
          if (h->kind == DEMI_EXT_PARTITION && h->a == e->src && h->b == e->dst) { k = 1; rp++; }
          break;
        case DEMI_EV_UNPARTITION:
          if (h->kind == DEMI_EXT_UNPARTITION && h->a == e->src && h->b == e->dst) { k = 1; rp++; }
          break;
        case DEMI_EV_SPAWN:
          if (h->kind == DEMI_EXT_START && h->a == e->dst) { k = 1; rp++; }
          break;
        default: k = 1; break;                     /* "Always include all other internal events" */
      }
    }
    /* ---- pass 2: filterSends (:425-446) */
    if (k && e->kind == DEMI_EV_MSG_SEND) {
      if (in->external_type_mask >> (e->type & 31) & 1u) {     /* EventTypes.isExternal */
        msg_send_idx++;
        if (msg_send_idx < (int32_t)n_sends && missing[msg_send_idx]) { k = 0; pruned_ids[e->uniq] = 1; }
      }
    } else if (k && e->kind == DEMI_EV_MSG_EVENT) {
      if (pruned_ids[e->uniq]) k = 0;
    }
    /* ---- pass 3: filterKnownAbsentInternals (:501-532), as written: a
     * PartitionEvent stores false and an UnPartitionEvent stores true */
    if (k && filter_known_absents) {
      switch (e->kind) {
        case DEMI_EV_MSG_SEND: {
          int snd_alive = e->src >= DEMI_MAX_ACTORS ? 1 : (int)((alive >> e->src) & 1u);
          int parted = e->src < DEMI_MAX_ACTORS ? (int)((partflag[e->src] >> e->dst) & 1u) : 0;
          if (!(snd_alive && !parted)) { k = 0; pruned_sends[e->uniq] = 1; }
          break;
        }
        case DEMI_EV_MSG_EVENT: {
          int rcv_alive = (int)((alive >> e->dst) & 1u);
          int parted = e->src < DEMI_MAX_ACTORS ? (int)((partflag[e->src] >> e->dst) & 1u) : 0;
          if (!(rcv_alive && !parted && !pruned_sends[e->uniq])) k = 0;
          break;
        }
        case DEMI_EV_SPAWN: alive |= 1u << e->dst; break;
        case DEMI_EV_KILL: alive &= ~(1u << e->dst); break;
        case DEMI_EV_PARTITION: partflag[e->src] &= ~(1u << e->dst); break;
        case DEMI_EV_UNPARTITION: partflag[e->src] |= 1u << e->dst; break;
        default: break;
      }
    }
    keep[i] = (uint8_t)k;
    kept += (uint32_t)k;
  }
  free(rem); free(missing); free(pruned_ids); free(pruned_sends);
  return (int)kept;
#undef IN_MASK
}

/* ----------------------------------------------------------- STS machine */
/* Reuses om_machine for actor states, network state, messagesToSend, the timer
 * registry and the cancelled set; the pending set is a multiset here
 * ((snd,rcv) -> fingerprint -> FIFO of indistinguishable entries,
 * STSScheduler.scala:112-114). */
typedef struct { om_machine m; uint32_t delivered, ignored; uint64_t rhash; } sts_machine;

/* Recording (the EventTrace STSScheduler.test returns, STSScheduler.scala:292-297): trigger_start/kill/
 * partition add their events (EventOrchestrator.scala:219-332), Quiescence markers are copied (:530-534),
 * every event_produced appends a MsgSend with a fresh Uniq (:570, :622) and every delivery a MsgEvent (:749).
 * With recording on, equal pending messages are dequeued oldest-first (Queue.dequeue, :729) so that the Uniq
 * ids pair up exactly as in the reference. */
static __thread demi_event* g_rec = 0;
static __thread uint32_t g_rec_cap = 0, g_rec_n = 0, g_rec_uniq = 0;
static __thread int g_rec_ovf = 0;
static void rec_push(uint8_t kind, uint8_t src, uint8_t dst, uint8_t type, uint32_t p0, uint32_t p1, uint16_t uniq) {
  if (!g_rec) return;
  if (g_rec_n >= g_rec_cap) { g_rec_ovf = 1; return; }
  demi_event* e = &g_rec[g_rec_n++];
  e->kind = kind; e->src = src; e->dst = dst; e->type = type; e->p0 = p0; e->p1 = p1; e->uniq = uniq; e->node = 0;
}


static int sts_find(const om_machine* m, int src, int dst, uint8_t type, uint32_t p0, uint32_t p1) {
  for (uint32_t i = 0; i < m->n_pending; i++) {
    const demi_msg* q = &m->pending[i].msg;
    if (q->src == src && q->dst == dst && q->type == type && q->p0 == p0 && q->p1 == p1) return (int)i;
  }
  return -1;
}
static void sts_remove_at(om_machine* m, uint32_t i) {
  if (g_rec) {                       /* keep insertion order: sts_find then returns the oldest equal entry */
    for (uint32_t j = i; j + 1 < m->n_pending; j++) m->pending[j] = m->pending[j + 1];
    m->n_pending--;
    return;
  }
  m->pending[i] = m->pending[m->n_pending - 1];
  m->n_pending--;
}
static int sts_crosses(const om_machine* m, int snd, int rcv) {
  int snd_actor = snd < DEMI_MAX_ACTORS;
  if (snd == rcv && !((m->killed >> snd) & 1u)) return 0;
  if (snd_actor && ((m->partitioned[snd] >> rcv) & 1u)) return 1;
  if (snd_actor && ((m->partitioned[rcv] >> snd) & 1u)) return 1;
  if ((m->inaccessible >> rcv) & 1u) return 1;
  if (snd_actor && ((m->inaccessible >> snd) & 1u)) return 1;
  return 0;
}
static int key_find(const om_timer_key* a, uint32_t n, int dst, uint8_t type, uint32_t p0, uint32_t p1) {
  for (uint32_t i = 0; i < n; i++)
    if (a[i].dst == dst && a[i].type == type && a[i].p0 == p0 && a[i].p1 == p1) return (int)i;
  return -1;
}
static void key_remove(om_timer_key* a, uint32_t* n, int i) {
  for (uint32_t j = (uint32_t)i; j + 1 < *n; j++) a[j] = a[j + 1];
  (*n)--;
}
static void key_push(om_machine* m, om_timer_key* a, uint32_t* n, int dst, uint8_t type, uint32_t p0, uint32_t p1) {
  if (*n >= DEMI_TIMERSET_CAP) { m->status = DEMI_PS_QUEUE_OVF; return; }
  a[*n].dst = (uint8_t)dst; a[*n].type = type; a[*n].p0 = p0; a[*n].p1 = p1; (*n)++;
}

/* STSScheduler.event_produced (STSScheduler.scala:561-623) after the
 * cancelled-timer drop of Instrumenter.aroundDispatch (Instrumenter.scala:1090-1096) */
static void sts_event_produced(om_machine* m, const demi_msg* msg) {
  if (m->status) return;
  int ci = key_find(m->cancelled, m->n_cancelled, msg->dst, msg->type, msg->p0, msg->p1);
  if (ci >= 0) { key_remove(m->cancelled, &m->n_cancelled, ci); return; }
  uint16_t uniq = (uint16_t)(++g_rec_uniq);                                /* Uniq(...) :570 */
  int is_timer = !(msg->flags & DEMI_MF_EXTERNAL) && msg->src == DEMI_DEADLETTERS;
  rec_push(DEMI_EV_MSG_SEND, is_timer ? DEMI_TIMER_SND : msg->src, msg->dst, msg->type, msg->p0, msg->p1, uniq);   /* :620-622 */
  if (!(msg->flags & DEMI_MF_EXTERNAL) && sts_crosses(m, msg->src, msg->dst)) return;
  if (m->n_pending >= m->pending_cap) { m->status = DEMI_PS_PENDING_OVF; return; }
  m->pending[m->n_pending].msg = *msg;
  m->pending[m->n_pending].uniq = uniq; m->pending[m->n_pending].node = 0;
  m->n_pending++;
  if (m->n_pending > m->max_pending) m->max_pending = m->n_pending;
}
static void sts_tosend_push(om_machine* m, const demi_msg* msg) {
  if (m->n_tosend >= m->tosend_cap) { m->status = DEMI_PS_QUEUE_OVF; return; }
  m->tosend[m->n_tosend++] = *msg;
}
static void sts_flush(om_machine* m) {
  for (uint32_t i = 0; i < m->n_tosend && !m->status; i++) sts_event_produced(m, &m->tosend[i]);
  m->n_tosend = 0;
}
/* STSScheduler.enqueue_timer == handle_timer (STSScheduler.scala:870) */
static void sts_handle_timer(om_machine* m, int rcv, uint8_t type, uint32_t p0, uint32_t p1) {
  if (m->ignore_timers) return;
  demi_msg t; t.src = DEMI_DEADLETTERS; t.dst = (uint8_t)rcv; t.type = type; t.flags = DEMI_MF_TIMER; t.p0 = p0; t.p1 = p1;
  sts_tosend_push(m, &t);
}

/* model callbacks are shared with the fuzz machine through a mode switch */
static __thread int g_sts_mode = 0;
int oracle_in_sts_mode(void) { return g_sts_mode; }

void sts_om_send(om_machine* m, int src, int dst, uint8_t type, uint32_t p0, uint32_t p1) {
  demi_msg msg; msg.src = (uint8_t)src; msg.dst = (uint8_t)dst; msg.type = type; msg.flags = 0; msg.p0 = p0; msg.p1 = p1;
  sts_event_produced(m, &msg);
}
void sts_om_schedule(om_machine* m, int self, uint8_t type, uint32_t p0, uint32_t p1, int repeating) {
  if (m->status) return;
  if (key_find(m->registry, m->n_registry, self, type, p0, p1) >= 0) return;
  if (repeating) { key_push(m, m->registry, &m->n_registry, self, type, p0, p1); if (m->status) return; }
  sts_handle_timer(m, self, type, p0, p1);
}
/* STSScheduler.notify_timer_cancel (STSScheduler.scala:846-868) */
void sts_om_cancel(om_machine* m, int self, uint8_t type, uint32_t p0, uint32_t p1) {
  if (m->status) return;
  if (key_find(m->cancelled, m->n_cancelled, self, type, p0, p1) < 0)
    key_push(m, m->cancelled, &m->n_cancelled, self, type, p0, p1);
  int ri = key_find(m->registry, m->n_registry, self, type, p0, p1);
  if (ri >= 0) key_remove(m->registry, &m->n_registry, ri);
  for (uint32_t i = 0; i < m->n_tosend; i++) {
    const demi_msg* q = &m->tosend[i];
    if (q->dst == self && q->type == type && q->p0 == p0 && q->p1 == p1) {
      for (uint32_t j = i; j + 1 < m->n_tosend; j++) m->tosend[j] = m->tosend[j + 1];
      m->n_tosend--;
      return;
    }
  }
  int pi = sts_find(m, DEMI_DEADLETTERS, self, type, p0, p1);
  if (pi >= 0) sts_remove_at(m, (uint32_t)pi);
}

void oracle_sts_replay(const demi_config* cfg, const demi_replay_input* in, const uint64_t* mask,
                       uint32_t looking_for, uint32_t flags, demi_replay_result* out, void* scratch) {
  oracle_sts_replay_ex(cfg, in, mask, looking_for, flags, 0xFFFFFFFFu, out, 0, 0, 0, scratch);
}

void oracle_sts_replay_ex(const demi_config* cfg, const demi_replay_input* in, const uint64_t* mask,
                          uint32_t looking_for, uint32_t flags, uint32_t skip_event, demi_replay_result* out,
                          demi_event* rec, uint32_t cap_rec, uint32_t* n_rec, void* scratch) {
  sts_machine* S = scratch ? (sts_machine*)scratch : (sts_machine*)malloc(sizeof(sts_machine));
  g_rec = rec; g_rec_cap = cap_rec; g_rec_n = 0; g_rec_uniq = 0; g_rec_ovf = 0;
  om_machine* m = &S->m;
  const oracle_model* model = oracle_get_model(cfg->model);
  memset(out, 0, sizeof(*out));
  uint8_t* keep = (uint8_t*)malloc(in->n_events + 1);
  oracle_sts_project(in, mask, (flags & DEMI_RF_FILTER_KNOWN_ABSENTS) != 0, keep);

  m->model = model; m->model_flags = cfg->model_flags; m->blocked_mask = cfg->blocked_mask;
  m->ignore_timers = cfg->ignore_timers; m->looking_for = looking_for;
  m->pending_cap = in->pending_cap; m->tosend_cap = in->tosend_cap;
  memset(m->states, 0, sizeof(m->states));
  model->init(m->states, cfg->model_flags);
  m->inaccessible = model->n_actors >= 32 ? 0xFFFFFFFFu : ((1u << model->n_actors) - 1u);
  m->killed = 0; memset(m->partitioned, 0, sizeof(m->partitioned));
  m->n_pending = m->max_pending = m->n_tosend = 0;
  m->n_registry = m->n_cancelled = m->n_just = m->n_resend = 0;
  m->status = 0; m->violation = 0;
  S->delivered = S->ignored = 0; S->rhash = 0;
  const int strict = (flags & DEMI_RF_STRICT) != 0;
  int diverged = 0;
  g_sts_mode = 1;

  uint32_t idx = 0;
  const uint32_t n = in->n_events;
  for (;;) {
    sts_flush(m);                                                   /* schedule_new_message :655 */
    /* advanceReplay (:405-559) */
    int found = 0;
    while (idx < n && !m->status) {
      const demi_event* e = &in->events[idx];
      if (keep[idx] && idx != skip_event) {
        switch (e->kind) {
          case DEMI_EV_SPAWN:                                       /* trigger_start */
            rec_push(DEMI_EV_SPAWN, e->src, e->dst, 0, 0, 0, 0);
            m->inaccessible &= ~(1u << e->dst); m->killed &= ~(1u << e->dst); break;
          case DEMI_EV_KILL:
            rec_push(DEMI_EV_KILL, e->src, e->dst, 0, 0, 0, 0);
            m->killed |= 1u << e->dst; m->inaccessible |= 1u << e->dst; break;
          case DEMI_EV_PARTITION: rec_push(DEMI_EV_PARTITION, e->src, e->dst, 0, 0, 0, 0); m->partitioned[e->src] |= 1u << e->dst; break;
          case DEMI_EV_UNPARTITION: rec_push(DEMI_EV_UNPARTITION, e->src, e->dst, 0, 0, 0, 0); m->partitioned[e->src] &= ~(1u << e->dst); break;
          case DEMI_EV_QUIESCENCE: case DEMI_EV_BEGIN_WAIT_QUIESCENCE:
            rec_push(e->kind, e->src, e->dst, 0, 0, 0, 0); break;
          case DEMI_EV_MSG_SEND:
            if ((in->external_type_mask >> (e->type & 31)) & 1u) {  /* :469-470 enqueue_message */
              demi_msg s; s.src = DEMI_DEADLETTERS; s.dst = e->dst; s.type = e->type; s.flags = DEMI_MF_EXTERNAL;
              s.p0 = e->p0; s.p1 = e->p1;
              sts_tosend_push(m, &s);
            }
            break;
          case DEMI_EV_MSG_EVENT: {
            sts_flush(m);                                           /* messagePending :381-403 */
            int pi = sts_find(m, e->src, e->dst, e->type, e->p0, e->p1);
            int enabled = pi >= 0 && !((m->blocked_mask >> e->dst) & 1u);
            if (enabled) { found = 1; }
            else if (strict) { diverged = 1; }                      /* ReplayScheduler: ReplayException */
            else S->ignored++;                                      /* "Ignoring message" :528-529 */
            break;
          }
          default: break;                                           /* Quiescence / BeginWaitQuiescence */
        }
      }
      if (found || diverged) break;
      idx++;
    }
    if (m->status || diverged) break;
    sts_flush(m);                                                   /* :682 */
    if (idx >= n) break;                                            /* trace_finished :685-689 */
    /* deliver the expected message (:696-772) */
    const demi_event* e = &in->events[idx];
    int pi = sts_find(m, e->src, e->dst, e->type, e->p0, e->p1);
    rec_push(DEMI_EV_MSG_EVENT, e->src, e->dst, e->type, e->p0, e->p1, m->pending[pi].uniq);   /* appendMsgEvent :749 */
    sts_remove_at(m, (uint32_t)pi);
    idx++;
    S->rhash += demi_event_term((uint32_t)e->src | ((uint32_t)e->dst << 8) | ((uint32_t)e->type << 16),
                                e->p0, e->p1, S->delivered, 0, 0);
    S->delivered++;
    /* Instrumenter.dispatch_new_message: re-arm a repeating timer, then receive() */
    if (key_find(m->registry, m->n_registry, e->dst, e->type, e->p0, e->p1) >= 0)
      sts_handle_timer(m, e->dst, e->type, e->p0, e->p1);
    if (m->status) break;
    demi_msg msg; msg.src = e->src; msg.dst = e->dst; msg.type = e->type; msg.flags = 0; msg.p0 = e->p0; msg.p1 = e->p1;
    model->receive(m, e->dst, &m->states[e->dst * model->state_words], &msg);
    if (m->status) break;
  }
  g_sts_mode = 0;

  if (m->status) { out->status = m->status; }
  else if (diverged) { out->status = DEMI_RS_DIVERGED; out->delivered = (uint16_t)S->delivered; }
  else {
    uint32_t v = model->invariant(m->states, m->model_flags);       /* :283-289 */
    out->violation = (looking_for ? (v == looking_for ? v : 0) : v);
    out->delivered = (uint16_t)S->delivered;
    out->ignored = (uint16_t)S->ignored;
    uint64_t sh = 0;
    uint32_t nw = (uint32_t)(model->n_actors * model->state_words);
    for (uint32_t i = 0; i < nw; i++) sh += demi_state_term(m->states[i], i);
    out->state_hash = sh + S->rhash;
  }
  if (n_rec) *n_rec = g_rec_n;
  if (g_rec_ovf && !out->status) out->status = DEMI_PS_EVENT_OVF;
  g_rec = 0;
  free(keep);
  if (!scratch) free(S);
}

size_t oracle_sts_scratch_size(void) { return sizeof(sts_machine); }
