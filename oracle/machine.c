/*
 * ORACLE — test infrastructure only (see machine.h header).  PARITY UNPINNED
 * against a running reference; pinned by derivable KATs only.
 *
 * Scalar restatement of one RandomScheduler execution.  Every function cites
 * the reference lines it follows (paths relative to
 * /root/reference/src/main/scala/verification/).
 */
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include "machine.h"
#include "sts.h"
#include "dpor.h"
#include "dpor_frontier.h"

uint32_t oracle_ir_fanout(void); uint32_t oracle_ir_external_mask(void);
int oracle_model_key(int model) { return model == 100 ? (100 | (int)(oracle_ir_fanout() << 8)) : model; }

/* ------------------------------------------------------------------ helpers */
static int key_eq(const om_timer_key* k, int dst, uint8_t type, uint32_t p0, uint32_t p1) {
  return k->dst == dst && k->type == type && k->p0 == p0 && k->p1 == p1;
}
static int set_find(const om_timer_key* a, uint32_t n, int dst, uint8_t type, uint32_t p0, uint32_t p1) {
  for (uint32_t i = 0; i < n; i++) if (key_eq(&a[i], dst, type, p0, p1)) return (int)i;
  return -1;
}
static void set_remove_at(om_timer_key* a, uint32_t* n, int i) {
  for (uint32_t j = (uint32_t)i; j + 1 < *n; j++) a[j] = a[j + 1];
  (*n)--;
}
static int set_push(om_machine* m, om_timer_key* a, uint32_t* n, int dst, uint8_t type, uint32_t p0, uint32_t p1) {
  if (*n >= DEMI_TIMERSET_CAP) { m->status = DEMI_PS_QUEUE_OVF; return 0; }
  a[*n].dst = (uint8_t)dst; a[*n].type = type; a[*n].p0 = p0; a[*n].p1 = p1;
  (*n)++;
  return 1;
}

/* EventTrace.+= / appendMsgSend / appendMsgEvent (EventTrace.scala:88-110) */
static void record_event(om_machine* m, uint8_t kind, uint8_t src, uint8_t dst, uint8_t type,
                         uint32_t p0, uint32_t p1, uint16_t uniq, uint16_t node, uint32_t parent) {
  uint32_t w0 = (uint32_t)kind | ((uint32_t)src << 8) | ((uint32_t)dst << 16) | ((uint32_t)type << 24);
  uint32_t w3 = (uint32_t)uniq | ((uint32_t)node << 16);
  m->trace_hash += demi_event_term(w0, p0, p1, w3, m->n_events, parent);
  if (m->events) {
    if (m->n_events >= m->event_cap) { m->status = DEMI_PS_EVENT_OVF; return; }
    demi_event* e = &m->events[m->n_events];
    e->kind = kind; e->src = src; e->dst = dst; e->type = type;
    e->p0 = p0; e->p1 = p1; e->uniq = uniq; e->node = node;
  }
  m->n_events++;
}

/* ------------------------------------------------- RandomizedHashSet (a1) */
#define FIFO_NIL 0xFFFFu
/* SrcDstFIFO.+= (RandomScheduler.scala:786-805) */
static void fifo_insert(om_machine* m, const om_pending* e) {
  if (m->n_pending + m->n_queued >= m->pending_cap) { m->status = DEMI_PS_PENDING_OVF; return; }
  if (e->msg.src == DEMI_DEADLETTERS) {
    m->pending[m->n_pending++] = *e;                       /* timersAndExternals += */
  } else {
    uint32_t pair = (uint32_t)e->msg.src * 32u + e->msg.dst;
    uint16_t slot = m->fifo_free;
    m->fifo_free = m->fifo_next[slot];
    m->fifo_pool[slot] = *e; m->fifo_next[slot] = FIFO_NIL;
    if (m->fifo_head[pair] == FIFO_NIL) {                  /* no queue yet: srcDsts.add */
      m->pairs[m->n_pairs++] = (uint16_t)pair;
      m->fifo_head[pair] = m->fifo_tail[pair] = slot;
    } else {
      m->fifo_next[m->fifo_tail[pair]] = slot;
      m->fifo_tail[pair] = slot;
    }
    m->n_queued++;
  }
  if (m->n_pending + m->n_queued > m->max_pending) m->max_pending = m->n_pending + m->n_queued;
}
/* RandomizedHashSet.insert: append (schedulers/Util.scala:126-136) */
void om_pending_insert(om_machine* m, const om_pending* e) {
  if (m->strategy == DEMI_RS_SRC_DST_FIFO) { fifo_insert(m, e); return; }
  if (m->n_pending >= m->pending_cap) { m->status = DEMI_PS_PENDING_OVF; return; }
  m->pending[m->n_pending++] = *e;
  if (m->n_pending > m->max_pending) m->max_pending = m->n_pending;
}
/* RandomizedHashSet.remove: A[i] = A[last]; shrink (schedulers/Util.scala:146-163) */
static om_pending pending_remove_at(om_machine* m, uint32_t i) {
  om_pending v = m->pending[i];
  m->pending[i] = m->pending[m->n_pending - 1];
  m->n_pending--;
  return v;
}
/* RandomizedHashSet.removeRandomElement (schedulers/Util.scala:171-176);
 * FullyRandom.removeRandomElement with the default (always-true) filter
 * (RandomScheduler.scala:666-684) reduces to exactly one draw. */
static demi_filter_rule g_filter[DEMI_MAX_FILTER_RULES];      /* process-wide: the batch driver's worker threads read it */
static uint32_t g_n_filter = 0;
void oracle_set_user_filter(const demi_filter_rule* rules, uint32_t n) {
  g_n_filter = n > DEMI_MAX_FILTER_RULES ? DEMI_MAX_FILTER_RULES : n;
  for (uint32_t i = 0; i < g_n_filter; i++) g_filter[i] = rules[i];
}
/* !userDefinedFilter(snd, rcv, msg) */
static int filter_rejects(const demi_msg* c) {
  for (uint32_t i = 0; i < g_n_filter; i++) {
    const demi_filter_rule* r = &g_filter[i];
    int src_ok = c->src < DEMI_MAX_ACTORS ? (int)((r->src_mask >> c->src) & 1u) : (int)(r->flags & DEMI_FRULE_DEADLETTERS);
    if (src_ok && ((r->dst_mask >> c->dst) & 1u) && ((r->type_mask >> (c->type & 31)) & 1u)) return 1;
  }
  return 0;
}
/* FullyRandom.removeRandomElement (RandomScheduler.scala:666-684), as written: redraw while the filter rejects and
 * more than one element is left; the rejected draws go back (appended, in draw order) after the loop */
om_pending om_pending_remove_random(om_machine* m) {
  int32_t idx = jr_next_int_bound(&m->rng, (int32_t)m->n_pending);
  om_pending ret = pending_remove_at(m, (uint32_t)idx);
  if (!g_n_filter) return ret;
  om_pending rejected[OM_MAX_PENDING]; uint32_t nr = 0;
  while (m->n_pending > 1 && filter_rejects(&ret.msg)) {
    rejected[nr++] = ret;
    idx = jr_next_int_bound(&m->rng, (int32_t)m->n_pending);
    ret = pending_remove_at(m, (uint32_t)idx);
  }
  for (uint32_t i = 0; i < nr; i++) m->pending[m->n_pending++] = rejected[i];
  return ret;
}
/* Util.find_non_blocked_message (schedulers/Util.scala:470-489): draw until
 * the receiver is not blocked; rejected draws are re-appended in draw order. */
int om_find_non_blocked(om_machine* m, om_pending* out) {
  if (m->n_pending == 0) return 0;
  static __thread om_pending blocked[OM_MAX_PENDING];
  uint32_t nb = 0;
  om_pending e = om_pending_remove_random(m);
  while (e.msg.dst < 32 && ((m->blocked_mask >> e.msg.dst) & 1u)) {
    blocked[nb++] = e;
    if (m->n_pending == 0) {
      for (uint32_t i = 0; i < nb; i++) m->pending[m->n_pending++] = blocked[i];
      return 0;
    }
    e = om_pending_remove_random(m);
  }
  for (uint32_t i = 0; i < nb; i++) m->pending[m->n_pending++] = blocked[i];
  *out = e;
  return 1;
}

/* SrcDstFIFO.getNonBlockedMessage (RandomScheduler.scala:716-756) + dequeue (:758-768) */
static int fifo_get_non_blocked(om_machine* m, om_pending* out) {
  int any = 0;
  for (uint32_t i = 0; i < m->n_pairs; i++) if (!((m->blocked_mask >> (m->pairs[i] & 31u)) & 1u)) { any = 1; break; }
  if (!any) return om_find_non_blocked(m, out);                            /* only timers left :717-728 */
  if ((uint32_t)jr_next_int_bound(&m->rng_pairs, (int32_t)(m->n_pending + m->n_queued)) < m->n_pending)   /* :732 */
    if (om_find_non_blocked(m, out)) return 1;
  int32_t idx = jr_next_int_bound(&m->rng_pairs, (int32_t)m->n_pairs);      /* :750-753 */
  while ((m->blocked_mask >> (m->pairs[idx] & 31u)) & 1u) idx = jr_next_int_bound(&m->rng_pairs, (int32_t)m->n_pairs);
  uint32_t pair = m->pairs[idx];
  uint16_t slot = m->fifo_head[pair];
  *out = m->fifo_pool[slot];
  m->fifo_head[pair] = m->fifo_next[slot];
  m->fifo_next[slot] = m->fifo_free; m->fifo_free = slot;
  if (m->fifo_head[pair] == FIFO_NIL) {                                    /* srcDsts.remove(idx): order preserving */
    for (uint32_t j = (uint32_t)idx; j + 1 < m->n_pairs; j++) m->pairs[j] = m->pairs[j + 1];
    m->n_pairs--;
  }
  m->n_queued--;
  return 1;
}

/* ------------------------------------------------------- DepTracker (a7) */
/* DepTracker.getMessage + addNodeAndEdge (DepTracker.scala:82-116): reuse the
 * child of parentEvent with equal (snd, rcv, fingerprint), else allocate a new
 * Unique.  Canonical order among equal children: lowest id. */
static uint16_t dep_report_newly_enabled(om_machine* m, const demi_msg* msg) {
  /* only the parent's children are compared (`parent.inNeighbors.find`, DepTracker.scala:94-101) */
  for (uint32_t i = m->node_first_child[m->parent_event]; i; i = m->node_next_sibling[i]) {
    const demi_msg* c = &m->node_msg[i];
    if (c->src == msg->src && c->dst == msg->dst && c->type == msg->type &&
        c->p0 == msg->p0 && c->p1 == msg->p1) return (uint16_t)i;
  }
  if (m->n_nodes >= m->node_cap) { m->status = DEMI_PS_NODE_OVF; return 0; }
  uint32_t id = m->n_nodes++;
  m->node_msg[id] = *msg;
  m->node_msg[id].flags = 0;
  m->node_parent[id] = (uint16_t)m->parent_event;
  m->node_first_child[id] = m->node_last_child[id] = m->node_next_sibling[id] = 0;
  if (m->node_last_child[m->parent_event]) m->node_next_sibling[m->node_last_child[m->parent_event]] = (uint16_t)id;
  else m->node_first_child[m->parent_event] = (uint16_t)id;
  m->node_last_child[m->parent_event] = (uint16_t)id;
  return (uint16_t)id;
}

/* ------------------------------------------------ EventOrchestrator (a6) */
/* EventOrchestrator.crosses_partition (EventOrchestrator.scala:345-351) */
static int crosses_partition(const om_machine* m, int snd, int rcv) {
  int snd_actor = snd < DEMI_MAX_ACTORS;
  if (snd == rcv && !((m->killed >> snd) & 1u)) return 0;
  if (snd_actor && ((m->partitioned[snd] >> rcv) & 1u)) return 1;
  if (snd_actor && ((m->partitioned[rcv] >> snd) & 1u)) return 1;
  if ((m->inaccessible >> rcv) & 1u) return 1;
  if (snd_actor && ((m->inaccessible >> snd) & 1u)) return 1;
  return 0;
}

/* RandomScheduler.event_produced(cell, envelope) (RandomScheduler.scala:274-321)
 * preceded by Instrumenter.aroundDispatch's cancelled-timer drop
 * (Instrumenter.scala:1090-1096). */
static void event_produced(om_machine* m, const demi_msg* msg) {
  if (m->status) return;
  int ci = set_find(m->cancelled, m->n_cancelled, msg->dst, msg->type, msg->p0, msg->p1);
  if (ci >= 0) { set_remove_at(m->cancelled, &m->n_cancelled, ci); return; }

  uint16_t uniq = (uint16_t)(++m->n_uniq);                 /* Uniq(...) :283 */
  int is_timer = 0;
  uint16_t node;
  if (msg->flags & DEMI_MF_EXTERNAL) {
    /* ExternalMessage branch :298-307 -> reportNewlyEnabledExternal
     * (DepTracker.scala:119-122): parentEvent = lastQuiescence (= root, since
     * noopWaitQuiescence defaults to true, DepTracker.scala:28,139-150). */
    m->parent_event = 0;
    node = dep_report_newly_enabled(m, msg);
    if (m->status) return;
    om_pending e; e.msg = *msg; e.uniq = uniq; e.node = node;
    om_pending_insert(m, &e);
  } else {
    /* InternalMessage branch :287-297 */
    if (msg->src == DEMI_DEADLETTERS) is_timer = 1;
    node = dep_report_newly_enabled(m, msg);
    if (m->status) return;
    if (!crosses_partition(m, msg->src, msg->dst)) {
      om_pending e; e.msg = *msg; e.uniq = uniq; e.node = node;
      om_pending_insert(m, &e);
    }
  }
  /* :319-320 record the MsgSend; timers are recorded with snd "Timer" */
  record_event(m, DEMI_EV_MSG_SEND, is_timer ? DEMI_TIMER_SND : msg->src, msg->dst, msg->type,
               msg->p0, msg->p1, uniq, node, m->node_parent[node]);
}

/* ------------------------------------- ExternalEventInjector (a10, timers) */
static void tosend_push(om_machine* m, const demi_msg* msg) {
  if (m->n_tosend >= m->tosend_cap) { m->status = DEMI_PS_QUEUE_OVF; return; }
  m->tosend[m->n_tosend++] = *msg;
}
/* ExternalEventInjector.handle_timer (ExternalEventInjector.scala:282-297) */
static void handle_timer(om_machine* m, int rcv, uint8_t type, uint32_t p0, uint32_t p1) {
  if (m->ignore_timers) return;
  demi_msg t; t.src = DEMI_DEADLETTERS; t.dst = (uint8_t)rcv; t.type = type;
  t.flags = DEMI_MF_TIMER; t.p0 = p0; t.p1 = p1;
  tosend_push(m, &t);
}
/* RandomScheduler.enqueue_timer (RandomScheduler.scala:549-559) */
static void enqueue_timer(om_machine* m, int rcv, uint8_t type, uint32_t p0, uint32_t p1) {
  if (set_find(m->just, m->n_just, rcv, type, p0, p1) >= 0) {
    set_push(m, m->resend, &m->n_resend, rcv, type, p0, p1);
    return;
  }
  handle_timer(m, rcv, type, p0, p1);
}
/* ExternalEventInjector.send_external_messages (ExternalEventInjector.scala:306-365):
 * drain messagesToSend in queue order; each one reaches event_produced. */
static void send_external_messages(om_machine* m) {
  for (uint32_t i = 0; i < m->n_tosend && !m->status; i++) {
    if ((m->dead >> m->tosend[i].dst) & 1u) continue;         /* "Dropping message to non-existent receiver" :343-346 */
    event_produced(m, &m->tosend[i]);
  }
  m->n_tosend = 0;
}

/* ---------------------------------------------- model-facing API (a8) */
/* `!` inside receive(): Instrumenter.tell -> aroundDispatch -> event_produced,
 * synchronously and in program order (Instrumenter.scala:1098-1108). */
void om_send(om_machine* m, int src, int dst, uint8_t type, uint32_t p0, uint32_t p1) {
  if (oracle_in_frontier_mode()) { front_om_send(m, src, dst, type, p0, p1); return; }
  if (oracle_in_dpor_mode()) { dpor_om_send(m, src, dst, type, p0, p1); return; }
  if (oracle_in_sts_mode()) { sts_om_send(m, src, dst, type, p0, p1); return; }
  demi_msg msg; msg.src = (uint8_t)src; msg.dst = (uint8_t)dst; msg.type = type; msg.flags = 0;
  msg.p0 = p0; msg.p1 = p1;
  event_produced(m, &msg);
}
/* scheduler.scheduleOnce: WeaveActor.aj:240-255 -> Instrumenter.registerCancellable
 * (ongoing=false) -> handleTick -> enqueue_timer -> removeCancellable
 * (Instrumenter.scala:1145-1200). */
void om_schedule_once(om_machine* m, int self, uint8_t type, uint32_t p0, uint32_t p1) {
  if (oracle_in_frontier_mode()) { front_om_schedule(m, self, type, p0, p1, 0); return; }
  if (oracle_in_dpor_mode()) { dpor_om_schedule(m, self, type, p0, p1, 0); return; }
  if (oracle_in_sts_mode()) { sts_om_schedule(m, self, type, p0, p1, 0); return; }
  if (m->status) return;
  if (set_find(m->registry, m->n_registry, self, type, p0, p1) >= 0) return; /* "Non-unique timer" :1154-1157 */
  enqueue_timer(m, self, type, p0, p1);
}
/* scheduler.schedule (repeating): WeaveActor.aj:264-279, ongoing=true. */
void om_schedule_repeating(om_machine* m, int self, uint8_t type, uint32_t p0, uint32_t p1) {
  if (oracle_in_frontier_mode()) { front_om_schedule(m, self, type, p0, p1, 1); return; }
  if (oracle_in_dpor_mode()) { dpor_om_schedule(m, self, type, p0, p1, 1); return; }
  if (oracle_in_sts_mode()) { sts_om_schedule(m, self, type, p0, p1, 1); return; }
  if (m->status) return;
  if (set_find(m->registry, m->n_registry, self, type, p0, p1) >= 0) return;
  if (!set_push(m, m->registry, &m->n_registry, self, type, p0, p1)) return;
  enqueue_timer(m, self, type, p0, p1);
}
/* Cancellable.cancel(): Instrumenter.cancelTimer (Instrumenter.scala:159-168)
 * -> RandomScheduler.notify_timer_cancel (RandomScheduler.scala:525-534):
 * first messagesToSend (ExternalEventInjector.scala:601-610), else the first
 * matching ("deadLetters", rcv, msg) in pendingEvents.arr order
 * (FullyRandom.remove, RandomScheduler.scala:653-664). */
void om_cancel_timer(om_machine* m, int self, uint8_t type, uint32_t p0, uint32_t p1) {
  if (oracle_in_frontier_mode()) { front_om_cancel(m, self, type, p0, p1); return; }
  if (oracle_in_dpor_mode()) { dpor_om_cancel(m, self, type, p0, p1); return; }
  if (oracle_in_sts_mode()) { sts_om_cancel(m, self, type, p0, p1); return; }
  if (m->status) return;
  if (set_find(m->cancelled, m->n_cancelled, self, type, p0, p1) < 0)
    set_push(m, m->cancelled, &m->n_cancelled, self, type, p0, p1);
  int ri = set_find(m->registry, m->n_registry, self, type, p0, p1);
  if (ri >= 0) set_remove_at(m->registry, &m->n_registry, ri);
  for (uint32_t i = 0; i < m->n_tosend; i++) {
    const demi_msg* q = &m->tosend[i];
    if (q->dst == self && q->type == type && q->p0 == p0 && q->p1 == p1) {
      for (uint32_t j = i; j + 1 < m->n_tosend; j++) m->tosend[j] = m->tosend[j + 1];
      m->n_tosend--;
      return;
    }
  }
  for (uint32_t i = 0; i < m->n_pending; i++) {
    const demi_msg* q = &m->pending[i].msg;
    if (q->src == DEMI_DEADLETTERS && q->dst == self && q->type == type && q->p0 == p0 && q->p1 == p1) {
      pending_remove_at(m, i);
      return;
    }
  }
}

/* ------------------------------------------------------ invariants (a9) */
/* RandomScheduler.violationMatches (RandomScheduler.scala:138-154) */
static uint32_t violation_matches(const om_machine* m, uint32_t v) {
  if (!m->looking_for) return v;
  if (!v) return 0;
  return v == m->looking_for ? m->looking_for : 0;
}

/* ------------------------------------------- EventOrchestrator externals */
/* EventOrchestrator.inject_until_quiescence (EventOrchestrator.scala:132-189) */
static void inject_until_quiescence(om_machine* m) {
  int loop = 1;
  while (loop && m->ext_idx < m->n_ext && !m->status) {
    const demi_ext_event* e = &m->ext[m->ext_idx];
    switch (e->kind) {
      case DEMI_EXT_START:       /* trigger_start :219-231 */
        record_event(m, DEMI_EV_SPAWN, DEMI_DEADLETTERS, e->a, 0, 0, 0, 0, 0, 0);
        m->inaccessible &= ~(1u << e->a);
        m->killed &= ~(1u << e->a);
        m->dead &= ~(1u << e->a);
        break;
      case DEMI_EXT_HARD_KILL: { /* trigger_hard_kill :243-310 */
        record_event(m, DEMI_EV_HARD_KILL, DEMI_DEADLETTERS, e->a, 0, 0, 0, 0, 0, 0);
        /* scheduler.actorTerminated(name) -> FullyRandom.removeAll (:686-696): every element of the array as it was
         * when the loop started is visited once, in position order; a match is swap-removed at its CURRENT index */
        if (m->strategy != DEMI_RS_FULLY_RANDOM) { m->status = DEMI_PS_QUEUE_OVF; break; }
        {
          om_pending snap[OM_MAX_PENDING]; uint32_t ns = m->n_pending;
          memcpy(snap, m->pending, sizeof(om_pending) * ns);
          for (uint32_t k = 0; k < ns; k++) {
            if (snap[k].msg.dst != e->a) continue;
            for (uint32_t j = 0; j < m->n_pending; j++)
              if (m->pending[j].uniq == snap[k].uniq) { pending_remove_at(m, j); break; }
          }
        }
        m->blocked_mask &= ~(1u << e->a);                       /* blockedActors - name :280 */
        for (uint32_t k = 0; k < m->n_registry;)                /* removeCancellable for its timers :281-287 */
          if (m->registry[k].dst == e->a) set_remove_at(m->registry, &m->n_registry, (int)k); else k++;
        m->killed |= 1u << e->a; m->inaccessible |= 1u << e->a; m->dead |= 1u << e->a;
        {                                                       /* the stopped instance is gone: a later Start is a fresh actor */
          uint32_t fresh[DEMI_MAX_ACTORS * OM_MAX_STATE_WORDS];
          memset(fresh, 0, sizeof(fresh));
          m->model->init(fresh, m->model_flags);
          memcpy(&m->states[e->a * m->model->state_words], &fresh[e->a * m->model->state_words], 4u * (uint32_t)m->model->state_words);
        }
        break;
      }
      case DEMI_EXT_KILL:        /* trigger_kill :233-241 */
        record_event(m, DEMI_EV_KILL, DEMI_DEADLETTERS, e->a, 0, 0, 0, 0, 0, 0);
        m->killed |= 1u << e->a;
        m->inaccessible |= 1u << e->a;
        break;
      case DEMI_EXT_SEND: {      /* :160-161 -> enqueue_message (ExternalEventInjector.scala:258-268) */
        demi_msg s; s.src = DEMI_DEADLETTERS; s.dst = e->a; s.type = e->type;
        s.flags = DEMI_MF_EXTERNAL; s.p0 = e->p0; s.p1 = e->p1;
        tosend_push(m, &s);
        break;
      }
      case DEMI_EXT_PARTITION:   /* trigger_partition :314-322 */
        record_event(m, DEMI_EV_PARTITION, e->a, e->b, 0, 0, 0, 0, 0, 0);
        m->partitioned[e->a] |= 1u << e->b;
        break;
      case DEMI_EXT_UNPARTITION: /* trigger_unpartition :324-332: removes the ordered pair only */
        record_event(m, DEMI_EV_UNPARTITION, e->a, e->b, 0, 0, 0, 0, 0, 0);
        m->partitioned[e->a] &= ~(1u << e->b);
        break;
      case DEMI_EXT_WAIT_QUIESCENCE: /* :182-184 */
        record_event(m, DEMI_EV_BEGIN_WAIT_QUIESCENCE, DEMI_DEADLETTERS, DEMI_DEADLETTERS, 0, 0, 0, 0, 0, 0);
        loop = 0;
        break;
      default: break;
    }
    m->ext_idx++;
  }
}

/* --------------------------------- RandomScheduler.schedule_new_message */
/* RandomScheduler.scala:352-485.  Returns 1 and the chosen entry, or 0 (None). */
static int schedule_new_message(om_machine* m, om_pending* out) {
  if (m->status) return 0;
  if (m->violation) return 0;                                   /* :354-360 */
  if (m->nsched > m->max_messages) {                            /* :369-373 */
    m->ext_idx = m->n_ext;                                      /* finish_early */
    return 0;
  }
  /* :376-401 with checkpointing disabled: lastCheckpoint stays 0 */
  if (m->interval > 0 && (m->nsched % m->interval) == 0 && m->nsched != 0) {
    uint32_t v = m->model->invariant(m->states, m->model_flags);
    m->violation = violation_matches(m, v);
    if (m->violation) return 0;
  }
  send_external_messages(m);                                    /* :424 */
  if (m->status) return 0;
  /* :426-439 pendingSystemMessages: always empty here (no FD / checkpoint actors) */
  om_pending pick;
  if (m->strategy == DEMI_RS_SRC_DST_FIFO) {                    /* :446-449 */
    if (!fifo_get_non_blocked(m, &pick)) return 0;
  } else if (!om_find_non_blocked(m, &pick)) return 0;          /* :451-457 */
  if (m->status) return 0;
  m->nsched++;                                                  /* :462 */
  if (m->nsched == INT_MAX) m->nsched = 1;
  /* :467 appendMsgEvent, :468 depTracker.reportNewlyDelivered (DepTracker.scala:132-135) */
  record_event(m, DEMI_EV_MSG_EVENT, pick.msg.src, pick.msg.dst, pick.msg.type,
               pick.msg.p0, pick.msg.p1, pick.uniq, pick.node, 0);
  m->parent_event = pick.node;
  /* updateRepeatingTimer :405-421 */
  if (set_find(m->registry, m->n_registry, pick.msg.dst, pick.msg.type, pick.msg.p0, pick.msg.p1) >= 0) {
    if (set_find(m->just, m->n_just, pick.msg.dst, pick.msg.type, pick.msg.p0, pick.msg.p1) < 0)
      set_push(m, m->just, &m->n_just, pick.msg.dst, pick.msg.type, pick.msg.p0, pick.msg.p1);
  } else {
    for (uint32_t i = 0; i < m->n_resend; i++)
      handle_timer(m, m->resend[i].dst, m->resend[i].type, m->resend[i].p0, m->resend[i].p1);
    m->n_resend = 0;
    m->n_just = 0;
  }
  *out = pick;
  return 1;
}

/* Instrumenter.dispatch_new_message (Instrumenter.scala:913-1017): hand the
 * message to the actor; a repeating timer is re-armed right after the hand-off
 * (:1008-1016), i.e. before the actor's receive() has run on its own thread. */
static void dispatch_new_message(om_machine* m, const om_pending* pick) {
  const demi_msg* msg = &pick->msg;
  if (set_find(m->registry, m->n_registry, msg->dst, msg->type, msg->p0, msg->p1) >= 0)
    enqueue_timer(m, msg->dst, msg->type, msg->p0, msg->p1);     /* handleTick :1185-1200 */
  if (m->status) return;
  m->model->receive(m, msg->dst, &m->states[msg->dst * m->model->state_words], msg);
}

/* ------------------------------------------------------- one execution */
void oracle_run_prefix(const demi_config* cfg, const demi_ext_event* ext, uint32_t n_ext,
                       const demi_fuzz_params* p, int64_t seed,
                       demi_fuzz_result* out,
                       demi_event* events, uint32_t cap_events,
                       uint16_t* dep_parent, uint32_t cap_nodes,
                       om_machine* scratch) {
  om_machine* m = scratch ? scratch : (om_machine*)malloc(sizeof(om_machine));
  const oracle_model* model = oracle_get_model(cfg->model);
  memset(out, 0, sizeof(*out));
  if (!model) { out->status = 0xFFFF; if (!scratch) free(m); return; }

  uint32_t n_sends = 0;
  for (uint32_t i = 0; i < n_ext; i++) if (ext[i].kind == DEMI_EXT_SEND) n_sends++;

  m->model = model;
  m->model_flags = cfg->model_flags;
  m->blocked_mask = cfg->blocked_mask;
  m->ignore_timers = cfg->ignore_timers;
  m->max_messages = p->max_messages < 0 ? INT_MAX : p->max_messages; /* maxMessages = Int.MaxValue (RandomScheduler.scala:54) */
  m->interval = p->invariant_check_interval;
  m->looking_for = p->looking_for;
  m->pending_cap = demi_pending_cap(oracle_model_key(cfg->model), p->max_messages, n_sends);
  m->tosend_cap = demi_tosend_cap(n_sends);
  m->node_cap = demi_node_cap(m->pending_cap);
  m->event_cap = cap_events;
  jr_seed(&m->rng, seed);
  m->strategy = cfg->strategy;
  jr_seed(&m->rng_pairs, seed);
  m->n_pairs = 0; m->n_queued = 0;
  if (m->strategy == DEMI_RS_SRC_DST_FIFO) {
    memset(m->fifo_head, 0xFF, sizeof(m->fifo_head)); memset(m->fifo_tail, 0xFF, sizeof(m->fifo_tail));
    for (uint32_t i = 0; i < OM_MAX_PENDING; i++) m->fifo_next[i] = (uint16_t)(i + 1 < OM_MAX_PENDING ? i + 1 : 0xFFFFu);
    m->fifo_free = 0;
  }
  memset(m->states, 0, sizeof(m->states));
  model->init(m->states, cfg->model_flags);
  /* populateActorSystem: every actor is created and isolated until its Start
   * (ExternalEventInjector.scala:371-378) */
  m->inaccessible = model->n_actors >= 32 ? 0xFFFFFFFFu : ((1u << model->n_actors) - 1u);
  m->killed = 0; m->dead = 0;
  memset(m->partitioned, 0, sizeof(m->partitioned));
  m->n_pending = 0; m->max_pending = 0; m->n_tosend = 0;
  m->n_just = m->n_resend = m->n_registry = m->n_cancelled = 0;
  m->n_nodes = 1;                       /* DepTracker.root, id 0 (DepTracker.scala:15-17) */
  memset(&m->node_msg[0], 0, sizeof(demi_msg));
  m->node_parent[0] = 0;
  m->node_first_child[0] = m->node_last_child[0] = m->node_next_sibling[0] = 0;
  m->parent_event = 0;
  m->events = events; m->n_events = 0; m->trace_hash = 0;
  m->n_uniq = 0; m->nsched = 0; m->ext_idx = 0; m->ext = ext; m->n_ext = n_ext;
  m->violation = 0; m->status = 0;

  /* execute_trace -> advanceTrace (ExternalEventInjector.scala:382-441), the
   * Instrumenter loop start_dispatch/afterMessageReceive (Instrumenter.scala:
   * 1113-1140, :794-815), notify_quiescence (RandomScheduler.scala:487-500)
   * and handle_quiescence (ExternalEventInjector.scala:541-580). */
  for (;;) {
    inject_until_quiescence(m);
    om_pending pick;
    while (schedule_new_message(m, &pick)) {
      dispatch_new_message(m, &pick);
      if (m->status) break;
    }
    if (m->status) break;
    if (m->violation) break;                        /* "Violation found early. Halting" */
    if (m->ext_idx < m->n_ext) {                    /* !trace_finished */
      record_event(m, DEMI_EV_QUIESCENCE, DEMI_DEADLETTERS, DEMI_DEADLETTERS, 0, 0, 0, 0, 0, 0);
      continue;                                     /* quiescenceCallback is a no-op; advanceTrace */
    }
    break;
  }
  /* explore(): checkIfBugFound only if messagesScheduledSoFar <= maxMessages
   * (RandomScheduler.scala:255-262, :156-180) */
  if (!m->status && m->nsched <= m->max_messages && !m->violation) {
    uint32_t v = model->invariant(m->states, m->model_flags);
    m->violation = violation_matches(m, v);
  }

  if (m->status) {
    out->status = m->status;
  } else {
    out->violation = m->violation;
    out->steps = (uint32_t)m->nsched;
    uint64_t sh = 0;
    uint32_t nw = (uint32_t)(model->n_actors * model->state_words);
    for (uint32_t i = 0; i < nw; i++) sh += demi_state_term(m->states[i], i);
    if (p->flags & DEMI_FF_HASH_PENDING)          /* order-free: a multiset hash of what is still in flight */
    {
      for (uint32_t i = 0; i < m->n_pending; i++) {
        const demi_msg* q = &m->pending[i].msg;
        sh += demi_pending_term((uint32_t)q->src | ((uint32_t)q->dst << 8) | ((uint32_t)q->type << 16), q->p0, q->p1);
      }
      for (uint32_t pi = 0; pi < m->n_pairs; pi++)
        for (uint16_t sl = m->fifo_head[m->pairs[pi]]; sl != 0xFFFFu; sl = m->fifo_next[sl]) {
          const demi_msg* q = &m->fifo_pool[sl].msg;
          sh += demi_pending_term((uint32_t)q->src | ((uint32_t)q->dst << 8) | ((uint32_t)q->type << 16), q->p0, q->p1);
        }
    }
    out->state_hash = sh;
    out->trace_hash = m->trace_hash;
    out->n_nodes = (uint16_t)m->n_nodes;
    out->n_events = (uint16_t)(m->n_events > 65535u ? 65535u : m->n_events);
    out->max_pending = (uint16_t)m->max_pending;
    out->status = 0;
    if (dep_parent) {
      uint32_t n = m->n_nodes < cap_nodes ? m->n_nodes : cap_nodes;
      for (uint32_t i = 0; i < n; i++) dep_parent[i] = m->node_parent[i];
    }
  }
  if (!scratch) free(m);
}
