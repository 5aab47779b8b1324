// replay_kernel.cuh — K2: batched STSScheduler.test (schedulers/STSScheduler.scala:
// 199-310).  One THREAD owns one test, i.e. one external-event subsequence
// (a bitmask over EventTrace.original_externals); the 32 lanes of a warp walk
// the SAME recorded trace in lock-step, so every lane handles the same event at
// the same time and the actor transition that runs for an expected delivery is
// the same message type in all lanes (near-zero divergence).
//
// Per event i (uniform) each lane evaluates, fused in one pass:
//   subsequenceIntersection (EventTrace.scala:290-380)      -> keep / drop
//   filterSends             (EventTrace.scala:382-452)      -> keep / drop
//   filterKnownAbsentInternals (EventTrace.scala:458-534)   -> keep / drop (optional)
// and, for kept events, the STSSched step: advanceReplay (:405-559),
// schedule_new_message (:643-776), event_produced (:561-623).
//
// Data layout: the trace + side tables are shared by the whole batch and stay
// L2/L1 resident (24 KB for 2000 events).  Per test: actor states / outbox /
// partition rows thread-interleaved in shared memory; the pending multiset
// ((snd,rcv) -> fingerprint -> FIFO of indistinguishable entries,
// STSScheduler.scala:112-114) is an open-addressing hash table of
// {hdr,p0,p1,generation|count} in HBM, interleaved per warp.
#pragma once
#include "lane_kernel.cuh"
#include "models/model_traits.cuh"

namespace demi {

struct ReplayArgs {
  uint32_t model_flags, blocked_mask;
  int32_t  ignore_timers;
  uint32_t looking_for, flags;
  // trace (shared by all tests)
  const uint4* events; uint32_t n_events;        // demi_event records
  const uint16_t* ev_ordinal;                    // per event: FIFO ordinal of the external Send it belongs to, 0xFFFF if none
  const uint4* ext; uint32_t n_ext;              // demi_ext_event records (original_externals)
  const uint16_t* send_ext_index; uint32_t n_sends;   // j-th original Send -> its index in ext
  uint32_t external_type_mask;
  uint32_t n_uniq_words;                         // words of the per-test pruned-send bitset
  // tests
  const uint64_t* masks; uint32_t n_masks, mask_words;   // masks == nullptr: every test uses the full subsequence
  const uint32_t* skips;                                   // optional: test i also drops trace event skips[i]
  demi_replay_result* results;
  // recording (single-test launches of the REC variant): the EventTrace STSScheduler.test returns
  demi_event* rec_events; uint32_t rec_cap; uint32_t* rec_count;
  // capacities
  uint32_t pending_cap, tosend_cap, table_slots; // table_slots: power of two >= 2*pending_cap
  uint32_t gen_base;                             // first hash-table generation this launch may use (1..0xFFFE)
  // per-warp scratch in HBM
  uint4*    table;      // [warps][table_slots][32]
  uint32_t* tosend;     // [warps][tosend_cap][32]
  uint32_t* pruned;     // [warps][n_uniq_words + N][32]  (only with DEMI_RF_FILTER_KNOWN_ABSENTS)
  unsigned long long* counters;   // [0] reproduced, [1] delivered
};

// receive()'s view when its operations are applied as they are issued (program order is the reference's order:
// `!`, scheduleOnce and cancel() act synchronously inside receive()); no staging in shared memory.
template <class M>
struct ReplayDirectOutbox {
  M* m; uint32_t self;
  __device__ __forceinline__ void send(uint32_t dst, uint32_t type, uint32_t p0, uint32_t p1) { m->event_produced(self, dst, type, p0, p1, false); }
  __device__ __forceinline__ void schedule_once(uint32_t type, uint32_t p0, uint32_t p1) { m->schedule_timer(OP_SCHED_ONCE, self, type, p0, p1); }
  __device__ __forceinline__ void schedule_repeating(uint32_t type, uint32_t p0, uint32_t p1) { m->schedule_timer(OP_SCHED_REPEAT, self, type, p0, p1); }
  __device__ __forceinline__ void cancel_timer(uint32_t type, uint32_t p0, uint32_t p1) { m->cancel_timer(self, type, p0, p1); }
};

template <class MODEL, int BD, bool REC = false>
struct ReplayMachine {
  static constexpr int N = MODEL::N_ACTORS;
  static constexpr int SW = MODEL::STATE_WORDS;
  static constexpr int OB = MODEL::REPLAY_OUTBOX;
  static constexpr bool DIRECT = model_replay_direct<MODEL>::value;
  static constexpr int OBW = DIRECT ? 0 : OB * 3;      // the outbox is staged in shared memory only when it can overflow
  static constexpr int WORDS = N * SW + OBW + N;

  uint32_t* smw;
  const ReplayArgs* A;
  uint4* table; uint32_t* tosend; uint32_t* pruned;
  const uint64_t* mask;

  uint32_t gen;                      // hash-table generation of this test
  uint32_t n_pending, n_tosend;
  uint32_t registry, cancelled;      // timer-slot bitmasks
  uint32_t inaccessible, killed;
  uint32_t status;
  uint32_t delivered, ignored;
  uint64_t rhash;
  // projection state
  uint32_t rem_cursor;               // next candidate index into ext for `remaining.head`
  uint32_t alive;                    // filterKnownAbsentInternals: actorToAlive

  __device__ __forceinline__ uint32_t& part_row(uint32_t a) { return smw[(N * SW + OBW + a) * BD]; }
  __device__ __forceinline__ LaneState actor(uint32_t a) { return LaneState{smw + a * SW * BD, BD}; }
  // externals are consulted in (nearly) increasing index order: keep the current 64-bit word of the mask in registers
  uint32_t mword_idx; uint64_t mword;
  __device__ __forceinline__ bool in_mask(uint32_t i) {
    if (!mask) return true;
    const uint32_t w = i >> 6;
    if (w != mword_idx) { mword_idx = w; mword = __ldg(mask + w); }
    return (mword >> (i & 63)) & 1ull;
  }
  // ---- recording (REC): trace + Uniq ids; the pending set becomes an insertion-ordered list so that
  // equal messages leave oldest-first (Queue.dequeue, STSScheduler.scala:729) and Uniq ids pair up
  uint32_t n_rec, n_uniq, n_list;
  __device__ __forceinline__ void rec_push(uint32_t kind, uint32_t src, uint32_t dst, uint32_t type,
                                           uint32_t p0, uint32_t p1, uint32_t uniq) {
    if (!REC) return;
    if (n_rec >= A->rec_cap) { status = DEMI_PS_EVENT_OVF; return; }
    reinterpret_cast<uint4*>(A->rec_events)[n_rec++] = make_uint4(kind | (src << 8) | (dst << 16) | (type << 24), p0, p1, uniq);
  }

  // ---- pending multiset
  __device__ __forceinline__ uint32_t hash_key(uint32_t hdr, uint32_t p0, uint32_t p1) const {
    return demi_fmix32((hdr * 0x9E3779B1u) ^ (p0 * 0x85EBCA77u) ^ (p1 * 0xC2B2AE3Du)) & (A->table_slots - 1);
  }
  // returns slot index of the key (live in this generation) or of the first free slot; found tells which
  __device__ __forceinline__ uint32_t probe(uint32_t hdr, uint32_t p0, uint32_t p1, bool& found, uint32_t& count) {
    uint32_t s = hash_key(hdr, p0, p1);
    for (uint32_t tries = 0;; tries++) {
      if (tries >= A->table_slots) { status = DEMI_RS_UNSUPPORTED; found = false; count = 0; return s; }
      uint4 q = table[(size_t)s * 32];
      if ((q.w >> 16) != gen) { found = false; count = 0; return s; }
      if (q.x == hdr && q.y == p0 && q.z == p1) { found = true; count = q.w & 0xFFFFu; return s; }
      s = (s + 1) & (A->table_slots - 1);
    }
  }
  __device__ __forceinline__ void pending_add(uint32_t hdr, uint32_t p0, uint32_t p1, uint32_t uniq = 0) {
    if (n_pending >= A->pending_cap) { status = DEMI_PS_PENDING_OVF; return; }
    if (REC) {
      if (n_list >= A->table_slots) { status = DEMI_RS_UNSUPPORTED; return; }
      table[(size_t)n_list * 32] = make_uint4(hdr, p0, p1, uniq | 0x80000000u);
      n_list++; n_pending++;
      return;
    }
    bool found; uint32_t count;
    uint32_t s = probe(hdr, p0, p1, found, count);
    if (status) return;
    table[(size_t)s * 32] = make_uint4(hdr, p0, p1, (gen << 16) | (count + 1));
    n_pending++;
  }
  __device__ __forceinline__ bool pending_take(uint32_t hdr, uint32_t p0, uint32_t p1, uint32_t* uniq_out = nullptr) {
    if (REC) {
      for (uint32_t i = 0; i < n_list; i++) {
        uint4 q = table[(size_t)i * 32];
        if ((q.w & 0x80000000u) && q.x == hdr && q.y == p0 && q.z == p1) {
          table[(size_t)i * 32] = make_uint4(q.x, q.y, q.z, q.w & 0x7FFFFFFFu);
          if (uniq_out) *uniq_out = q.w & 0xFFFFu;
          n_pending--;
          return true;
        }
      }
      return false;
    }
    bool found; uint32_t count;
    uint32_t s = probe(hdr, p0, p1, found, count);
    if (!found || count == 0) return false;
    table[(size_t)s * 32] = make_uint4(hdr, p0, p1, (gen << 16) | (count - 1));
    n_pending--;
    return true;
  }
  __device__ __forceinline__ bool pending_has(uint32_t hdr, uint32_t p0, uint32_t p1) {
    bool found; uint32_t count;
    probe(hdr, p0, p1, found, count);
    return found && count > 0;
  }

  __device__ __forceinline__ bool crosses_partition(uint32_t snd, uint32_t rcv) {
    bool snd_actor = snd < DEMI_MAX_ACTORS;
    if (snd == rcv && !((killed >> snd) & 1u)) return false;
    if (snd_actor) {
      if ((part_row(snd) >> rcv) & 1u) return true;
      if ((part_row(rcv) >> snd) & 1u) return true;
    }
    if ((inaccessible >> rcv) & 1u) return true;
    if (snd_actor && ((inaccessible >> snd) & 1u)) return true;
    return false;
  }

  // STSScheduler.event_produced (STSScheduler.scala:561-623) after the
  // cancelled-timer drop of Instrumenter.aroundDispatch (Instrumenter.scala:1090-1096)
  __device__ __forceinline__ void event_produced(uint32_t src, uint32_t dst, uint32_t type, uint32_t p0, uint32_t p1,
                                                 bool external) {
    if (status) return;
    int slot = MODEL::timer_slot(dst, type, p0, p1);
    if (cancelled && slot >= 0 && ((cancelled >> slot) & 1u)) { cancelled &= ~(1u << slot); return; }
    uint32_t uniq = 0;
    if (REC) {
      uniq = ++n_uniq;                                              // Uniq(...) :570; appendMsgSend :620-622
      rec_push(DEMI_EV_MSG_SEND, (!external && src == DEMI_DEADLETTERS) ? DEMI_TIMER_SND : src, dst, type, p0, p1, uniq);
    }
    if (!external && crosses_partition(src, dst)) return;
    pending_add(make_hdr(src, dst, type, 0), p0, p1, uniq);
  }
  __device__ __forceinline__ void tosend_push(uint32_t code) {
    if (n_tosend >= A->tosend_cap) { status = DEMI_PS_QUEUE_OVF; return; }
    tosend[(size_t)n_tosend * 32] = code;
    n_tosend++;
  }
  // ExternalEventInjector.send_external_messages: codes are (1<<31 | event index) for
  // externals re-sent from the trace, else a timer slot
  __device__ __forceinline__ void flush() {
    for (uint32_t i = 0; i < n_tosend && !status; i++) {
      uint32_t code = tosend[(size_t)i * 32];
      if (code & 0x80000000u) {
        uint4 e = __ldg(A->events + (code & 0x7FFFFFFFu));
        event_produced(DEMI_DEADLETTERS, (e.x >> 16) & 0xFF, e.x >> 24, e.y, e.z, true);
      } else {
        uint32_t dst, type, p0, p1;
        MODEL::slot_msg(code, dst, type, p0, p1);
        event_produced(DEMI_DEADLETTERS, dst, type, p0, p1, false);
      }
    }
    n_tosend = 0;
  }
  __device__ __forceinline__ void handle_timer(uint32_t slot) {   // STSScheduler.enqueue_timer (:870)
    if (A->ignore_timers) return;
    tosend_push(slot);
  }
  // STSScheduler.notify_timer_cancel (STSScheduler.scala:846-868)
  __device__ __forceinline__ void cancel_timer(uint32_t self, uint32_t type, uint32_t p0, uint32_t p1) {
    if (status) return;
    int slot = MODEL::timer_slot(self, type, p0, p1);
    if (slot < 0) { status = DEMI_RS_UNSUPPORTED; return; }
    uint32_t bit = 1u << slot;
    if (!(cancelled & bit) && __popc(cancelled) >= DEMI_TIMERSET_CAP) { status = DEMI_PS_QUEUE_OVF; return; }
    cancelled |= bit;
    registry &= ~bit;
    for (uint32_t i = 0; i < n_tosend; i++)
      if (tosend[(size_t)i * 32] == (uint32_t)slot) {               // order-preserving removal
        for (uint32_t j = i; j + 1 < n_tosend; j++) tosend[(size_t)j * 32] = tosend[(size_t)(j + 1) * 32];
        n_tosend--;
        return;
      }
    pending_take(make_hdr(DEMI_DEADLETTERS, self, type, 0), p0, p1);
  }

  // Instrumenter.dispatch_new_message: re-arm a repeating timer, then receive()
  __device__ __forceinline__ void deliver(uint32_t src, uint32_t dst, uint32_t type, uint32_t p0, uint32_t p1) {
    rhash += demi_event_term(src | (dst << 8) | (type << 16), p0, p1, delivered, 0, 0);
    delivered++;
    int slot = MODEL::timer_slot(dst, type, p0, p1);
    if (slot >= 0 && ((registry >> slot) & 1u)) handle_timer((uint32_t)slot);
    if (status) return;
    if constexpr (DIRECT) {
      ReplayDirectOutbox<ReplayMachine> direct{this, dst};
      MODEL::receive(direct, dst, actor(dst), src, type, p0, p1, A->model_flags);
    } else {
      LaneOutbox<OB> ob;
      ob.base = smw + N * SW * BD; ob.bd = BD; ob.n = 0; ob.self = dst; ob.overflow = false;
      MODEL::receive(ob, dst, actor(dst), src, type, p0, p1, A->model_flags);
      if (ob.overflow) { status = DEMI_PS_QUEUE_OVF; return; }
      for (uint32_t i = 0; i < ob.n && !status; i++) {
        uint32_t w0 = ob.base[(i * 3) * BD], q0 = ob.base[(i * 3 + 1) * BD], q1 = ob.base[(i * 3 + 2) * BD];
        uint32_t kind = w0 & 0xFF, odst = (w0 >> 8) & 0xFF, otype = (w0 >> 16) & 0xFF;
        if (kind == OP_SEND) event_produced(dst, odst, otype, q0, q1, false);
        else if (kind == OP_CANCEL) cancel_timer(odst, otype, q0, q1);
        else schedule_timer(kind, odst, otype, q0, q1);
      }
    }
    flush();     // schedule_new_message begins with send_external_messages (:655)
  }
  // scheduler.scheduleOnce / schedule (Instrumenter.scala:1126-1190)
  __device__ __forceinline__ void schedule_timer(uint32_t kind, uint32_t self, uint32_t type, uint32_t p0, uint32_t p1) {
    if (status) return;
    const int s2 = MODEL::timer_slot(self, type, p0, p1);
    if (s2 < 0) { status = DEMI_RS_UNSUPPORTED; return; }
    if ((registry >> s2) & 1u) return;
    if (kind == OP_SCHED_REPEAT) {
      if (__popc(registry) >= DEMI_TIMERSET_CAP) { status = DEMI_PS_QUEUE_OVF; return; }
      registry |= 1u << s2;
    }
    handle_timer((uint32_t)s2);
  }

  __device__ __forceinline__ void run(uint32_t test_idx, uint32_t generation, demi_replay_result& out) {
    mask = A->masks ? A->masks + (size_t)test_idx * A->mask_words : nullptr;
    mword_idx = 0xFFFFFFFFu; mword = 0;
    const uint32_t skip = A->skips ? A->skips[test_idx] : 0xFFFFFFFFu;
    n_rec = n_uniq = n_list = 0;
    gen = generation;
    for (uint32_t i = 0; i < N * SW; i++) smw[i * BD] = MODEL::init_word(i, A->model_flags);
    for (uint32_t a = 0; a < N; a++) part_row(a) = 0;
    n_pending = n_tosend = 0; registry = cancelled = 0;
    inaccessible = (N >= 32) ? 0xFFFFFFFFu : ((1u << N) - 1u);
    killed = 0; status = 0; delivered = ignored = 0; rhash = 0;
    rem_cursor = 0; alive = 0;
    const bool fka = (A->flags & DEMI_RF_FILTER_KNOWN_ABSENTS) != 0;
    const bool strict = (A->flags & DEMI_RF_STRICT) != 0;
    // filterKnownAbsentInternals: prunedMessageSends bitset, then N rows of actorsToPartitioned flags
    if (fka) for (uint32_t w = 0; w < A->n_uniq_words + N; w++) pruned[(size_t)w * 32] = 0;
#define PF_ROW(a) pruned[(size_t)(A->n_uniq_words + (a)) * 32]
    bool diverged = false;
    const uint32_t n_ev = A->n_events;

    // the walk is one dependent chain per test; the next event and its ordinal are fetched a step ahead so that their
    // L2 latency overlaps the current event's work
    uint4 e_next = n_ev ? __ldg(A->events) : make_uint4(0, 0, 0, 0);
    uint32_t ord_next = n_ev ? __ldg(A->ev_ordinal) : 0xFFFFu;
    for (uint32_t i = 0; i < n_ev && !status && !diverged; i++) {
      const uint4 e = e_next;
      const uint32_t ord_cur = ord_next;
      if (i + 1 < n_ev) { e_next = __ldg(A->events + i + 1); ord_next = __ldg(A->ev_ordinal + i + 1); }
      const uint32_t kind = e.x & 0xFF, src = (e.x >> 8) & 0xFF, dst = (e.x >> 16) & 0xFF, type = e.x >> 24;
      const uint32_t uniq = e.w & 0xFFFF;
      const bool is_msg = kind == DEMI_EV_MSG_SEND || kind == DEMI_EV_MSG_EVENT;
      bool k = false;
      // ---- pass 1: subsequenceIntersection (EventTrace.scala:307-374)
      if (is_msg || kind == DEMI_EV_QUIESCENCE || kind == DEMI_EV_BEGIN_WAIT_QUIESCENCE) {
        k = true;
      } else {
        // advance to `remaining.head`: next external in the subsequence that is not a Send
        while (rem_cursor < A->n_ext) {
          uint32_t xk = __ldg(A->ext + rem_cursor).x & 0xFF;
          if (xk != DEMI_EXT_SEND && in_mask(rem_cursor)) break;
          rem_cursor++;
        }
        if (rem_cursor < A->n_ext) {
          uint32_t hx = __ldg(A->ext + rem_cursor).x;
          uint32_t hk = hx & 0xFF, ha = (hx >> 8) & 0xFF, hb = (hx >> 16) & 0xFF;
          if (kind == DEMI_EV_KILL) k = (hk == DEMI_EXT_KILL && ha == dst);
          else if (kind == DEMI_EV_SPAWN) k = (hk == DEMI_EXT_START && ha == dst);
          else if (kind == DEMI_EV_PARTITION) k = (hk == DEMI_EXT_PARTITION && ha == src && hb == dst);
          else if (kind == DEMI_EV_UNPARTITION) k = (hk == DEMI_EXT_UNPARTITION && ha == src && hb == dst);
          if (k) rem_cursor++;
        }   // remaining.isEmpty: external-type events are dropped (:308-314)
      }
      // ---- pass 2: filterSends (EventTrace.scala:425-446), FIFO-ordinal form
      if (k && is_msg) {
        const uint32_t ord = ord_cur;
        if (ord < A->n_sends && !in_mask(__ldg(A->send_ext_index + ord))) k = false;
      }
      // ---- pass 3: filterKnownAbsentInternals (EventTrace.scala:501-532), as written
      if (k && fka) {
        if (kind == DEMI_EV_MSG_SEND) {
          bool snd_alive = src >= DEMI_MAX_ACTORS || ((alive >> src) & 1u);
          bool parted = src < DEMI_MAX_ACTORS && ((PF_ROW(src) >> dst) & 1u);
          if (!(snd_alive && !parted)) { k = false; pruned[(size_t)(uniq >> 5) * 32] |= 1u << (uniq & 31); }
        } else if (kind == DEMI_EV_MSG_EVENT) {
          bool rcv_alive = (alive >> dst) & 1u;
          bool parted = src < DEMI_MAX_ACTORS && ((PF_ROW(src) >> dst) & 1u);
          bool ps = (pruned[(size_t)(uniq >> 5) * 32] >> (uniq & 31)) & 1u;
          if (!(rcv_alive && !parted && !ps)) k = false;
        } else if (kind == DEMI_EV_SPAWN) alive |= 1u << dst;
        else if (kind == DEMI_EV_KILL) alive &= ~(1u << dst);
        else if (kind == DEMI_EV_PARTITION) PF_ROW(src) &= ~(1u << dst);
        else if (kind == DEMI_EV_UNPARTITION) PF_ROW(src) |= 1u << dst;
      }
      if (!k || i == skip) continue;
      // ---- STSSched: advanceReplay (:405-559)
      switch (kind) {
        case DEMI_EV_SPAWN: rec_push(kind, src, dst, 0, 0, 0, 0); inaccessible &= ~(1u << dst); killed &= ~(1u << dst); break;
        case DEMI_EV_KILL: rec_push(kind, src, dst, 0, 0, 0, 0); killed |= 1u << dst; inaccessible |= 1u << dst; break;
        case DEMI_EV_PARTITION: rec_push(kind, src, dst, 0, 0, 0, 0); part_row(src) |= 1u << dst; break;
        case DEMI_EV_UNPARTITION: rec_push(kind, src, dst, 0, 0, 0, 0); part_row(src) &= ~(1u << dst); break;
        case DEMI_EV_QUIESCENCE: case DEMI_EV_BEGIN_WAIT_QUIESCENCE: rec_push(kind, src, dst, 0, 0, 0, 0); break;
        case DEMI_EV_MSG_SEND:
          if ((A->external_type_mask >> (type & 31)) & 1u) tosend_push(0x80000000u | i);   // enqueue_message :469-470
          break;
        case DEMI_EV_MSG_EVENT: {
          flush();                                                   // messagePending :381-403
          if (status) break;
          uint32_t hdr = make_hdr(src, dst, type, 0);
          uint32_t duniq = 0;
          bool enabled = !((A->blocked_mask >> (dst & 31)) & 1u) && pending_take(hdr, e.y, e.z, &duniq);
          if (enabled) { rec_push(DEMI_EV_MSG_EVENT, src, dst, type, e.y, e.z, duniq); deliver(src, dst, type, e.y, e.z); }   // :696-772
          else if (strict) diverged = true;                          // ReplayScheduler: ReplayException
          else ignored++;                                            // "Ignoring message" :528-529
          break;
        }
        default: break;
      }
    }
#undef PF_ROW
    if (!status && !diverged) flush();                               // :682 before trace_finished
    out.violation = 0; out.delivered = 0; out.ignored = 0; out.state_hash = 0;
    if (REC && A->rec_count) *A->rec_count = n_rec;
    if (status) { out.status = (uint16_t)status; return; }
    if (diverged) { out.status = DEMI_RS_DIVERGED; out.delivered = (uint16_t)delivered; return; }
    uint32_t v = MODEL::invariant(LaneAll<SW>{smw, BD}, A->model_flags);      // :283-289
    out.violation = (uint16_t)(A->looking_for ? (v == A->looking_for ? v : 0u) : v);
    out.status = 0;
    out.delivered = (uint16_t)delivered; out.ignored = (uint16_t)ignored;
    uint64_t sh = 0;
    for (uint32_t i = 0; i < N * SW; i++) sh += demi_state_term(smw[i * BD], i);
    out.state_hash = sh + rhash;
  }
};

template <class MODEL, int BD, bool REC = false>
__global__ void __launch_bounds__(BD, (model_replay_direct<MODEL>::value && MODEL::N_ACTORS <= 8) ? 1024 / BD : 1)
replay_lane_kernel(const __grid_constant__ ReplayArgs args) {
  using M = ReplayMachine<MODEL, BD, REC>;
  extern __shared__ __align__(16) uint32_t lane_smem[];
  const uint32_t tid = threadIdx.x;
  const uint64_t gthread = (uint64_t)blockIdx.x * BD + tid;
  const uint64_t total = (uint64_t)gridDim.x * BD;
  const uint64_t gwarp = gthread >> 5;
  const uint32_t lane = tid & 31;

  M m;
  m.smw = lane_smem + tid;
  m.A = &args;
  m.table = args.table + gwarp * (uint64_t)args.table_slots * 32 + lane;
  m.tosend = args.tosend + gwarp * (uint64_t)args.tosend_cap * 32 + lane;
  m.pruned = args.pruned ? args.pruned + gwarp * (uint64_t)(args.n_uniq_words + M::N) * 32 + lane : nullptr;

  unsigned long long my_repro = 0, my_deliv = 0;
  // generations continue across launches (the host hands out disjoint ranges), so the table needs no
  // clearing between launches; on wrap-around this thread clears its own slots
  uint32_t generation = args.gen_base - 1;
  for (uint64_t idx = gthread; idx < args.n_masks; idx += total) {
    generation++;
    if (generation >= 0xFFFFu) {
      for (uint32_t s = 0; s < args.table_slots; s++) m.table[(size_t)s * 32] = make_uint4(0, 0, 0, 0);
      generation = 1;
    }
    demi_replay_result r;
    m.run((uint32_t)idx, generation, r);
    uint4* dst = reinterpret_cast<uint4*>(args.results + idx);
    dst[0] = make_uint4((uint32_t)r.violation | ((uint32_t)r.status << 16),
                        (uint32_t)r.delivered | ((uint32_t)r.ignored << 16),
                        (uint32_t)r.state_hash, (uint32_t)(r.state_hash >> 32));
    my_repro += (r.status == 0 && r.violation) ? 1u : 0u;
    my_deliv += r.delivered;
  }
  for (int o = 16; o > 0; o >>= 1) {
    my_repro += __shfl_xor_sync(FULL_MASK, my_repro, o);
    my_deliv += __shfl_xor_sync(FULL_MASK, my_deliv, o);
  }
  if (lane == 0 && args.counters && (my_repro | my_deliv)) {
    atomicAdd(args.counters, my_repro);
    atomicAdd(args.counters + 1, my_deliv);
  }
}

}  // namespace demi
