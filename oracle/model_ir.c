/*
 * ORACLE — test infrastructure only (see machine.h header).
 *
 * The CPU interpreter of the model IR (include/demi_model_ir.h): the oracle's copy of what the engine runs on the
 * device for a model loaded with demi_load_model.  One loaded blob per process (tests load, run, replace).
 */
#include <stdlib.h>
#include <string.h>
#include "machine.h"
#include "../include/demi_model_ir.h"

static uint32_t* g_blob = 0;
static const uint32_t *g_recv = 0, *g_inv = 0, *g_init = 0;
static uint32_t g_recv_len = 0, g_inv_len = 0, g_n_actors = 0, g_state_words = 0, g_ext_mask = 0, g_fanout = 0;

int oracle_load_model(const void* blob, size_t size) {
  const uint32_t* w = (const uint32_t*)blob;
  if (size < DEMI_IR_HEADER_WORDS * 4 || w[0] != DEMI_IR_MAGIC || w[1] != DEMI_IR_VERSION) return DEMI_ERR_INVALID;
  if (w[2] < 1 || w[2] > DEMI_IR_ACTORS || w[3] < 1 || w[3] > DEMI_IR_STATE_WORDS || w[5] > DEMI_IR_MAX_CODE || w[6] > DEMI_IR_MAX_CODE ||
      w[8] > DEMI_IR_OUTBOX) return DEMI_ERR_INVALID;
  const size_t need = (size_t)(DEMI_IR_HEADER_WORDS + w[5] + w[6] + w[2] * w[3]) * 4 + w[9];
  if (size < need) return DEMI_ERR_INVALID;
  free(g_blob);
  g_blob = (uint32_t*)malloc(size); memcpy(g_blob, blob, size);
  g_n_actors = g_blob[2]; g_state_words = g_blob[3]; g_recv_len = g_blob[5]; g_inv_len = g_blob[6]; g_ext_mask = g_blob[7]; g_fanout = g_blob[8];
  g_recv = g_blob + DEMI_IR_HEADER_WORDS; g_inv = g_recv + g_recv_len; g_init = g_inv + g_inv_len;
  return DEMI_OK;
}
uint32_t oracle_ir_external_mask(void) { return g_ext_mask; }
uint32_t oracle_ir_fanout(void) { return g_fanout; }

/* runs `code`; receive: st = own state, all = NULL; invariant: all = every state, st = NULL */
static uint32_t ir_run(const uint32_t* code, uint32_t len, uint32_t* r, om_machine* m, int self, uint32_t* st, const uint32_t* all,
                       uint32_t* affected) {
  uint32_t pc = 0;
  for (uint32_t steps = 0; steps < DEMI_IR_MAX_STEPS && pc < len; steps++) {
    const uint32_t ins = code[pc++];
    const uint32_t op = ins & 0xFF, a = (ins >> 8) & 15, b = (ins >> 16) & 15, c = (ins >> 24) & 15;
    uint32_t imm = 0;
    if (op == DEMI_IR_LDI || (op >= DEMI_IR_JMP && op <= DEMI_IR_JGE)) { if (pc >= len) return 0; imm = code[pc++]; }
    switch (op) {
      case DEMI_IR_HALT: return 0;
      case DEMI_IR_LDI: r[a] = imm; break;
      case DEMI_IR_MOV: r[a] = r[b]; break;
      case DEMI_IR_ADD: r[a] = r[b] + r[c]; break;
      case DEMI_IR_SUB: r[a] = r[b] - r[c]; break;
      case DEMI_IR_MUL: r[a] = r[b] * r[c]; break;
      case DEMI_IR_AND: r[a] = r[b] & r[c]; break;
      case DEMI_IR_OR:  r[a] = r[b] | r[c]; break;
      case DEMI_IR_XOR: r[a] = r[b] ^ r[c]; break;
      case DEMI_IR_SHL: r[a] = r[b] << (r[c] & 31); break;
      case DEMI_IR_SHR: r[a] = r[b] >> (r[c] & 31); break;
      case DEMI_IR_MOD: r[a] = r[c] ? r[b] % r[c] : 0; break;
      case DEMI_IR_LDW: r[a] = (st && r[b] < DEMI_IR_STATE_WORDS) ? st[r[b]] : 0; break;
      case DEMI_IR_STW: if (st && r[a] < DEMI_IR_STATE_WORDS) st[r[a]] = r[b]; break;
      case DEMI_IR_LDA: r[a] = (all && r[b] < DEMI_IR_ACTORS && r[c] < DEMI_IR_STATE_WORDS) ? all[r[b] * DEMI_IR_STATE_WORDS + r[c]] : 0; break;
      case DEMI_IR_JMP: pc = imm; break;
      case DEMI_IR_JEQ: if (r[a] == r[b]) pc = imm; break;
      case DEMI_IR_JNE: if (r[a] != r[b]) pc = imm; break;
      case DEMI_IR_JLT: if (r[a] < r[b]) pc = imm; break;
      case DEMI_IR_JGE: if (r[a] >= r[b]) pc = imm; break;
      case DEMI_IR_SEND: if (m) om_send(m, self, (int)(r[a] & 0xFF), (uint8_t)r[b], r[c], r[(c + 1) & 15]); break;
      case DEMI_IR_SCHED_ONCE: if (m) om_schedule_once(m, self, (uint8_t)r[a], r[b], r[c]); break;
      case DEMI_IR_SCHED_REPEAT: if (m) om_schedule_repeating(m, self, (uint8_t)r[a], r[b], r[c]); break;
      case DEMI_IR_CANCEL: if (m) om_cancel_timer(m, self, (uint8_t)r[a], r[b], r[c]); break;
      case DEMI_IR_RET: if (affected) *affected = r[(a + 1) & 15]; return r[a];
      default: return 0;
    }
  }
  return 0;
}

static void ir_init(uint32_t* states, uint32_t flags) {
  (void)flags;
  for (uint32_t a = 0; a < g_n_actors; a++)
    for (uint32_t w = 0; w < g_state_words; w++) states[a * DEMI_IR_STATE_WORDS + w] = g_init[a * g_state_words + w];
}
static void ir_receive(om_machine* m, int self, uint32_t* st, const demi_msg* msg) {
  uint32_t r[16]; memset(r, 0, sizeof(r));
  r[0] = (uint32_t)self; r[1] = msg->src; r[2] = msg->type; r[3] = msg->p0; r[4] = msg->p1; r[5] = m->model_flags;
  ir_run(g_recv, g_recv_len, r, m, self, st, 0, 0);
}
static uint32_t ir_invariant(const uint32_t* states, uint32_t flags) {
  uint32_t r[16]; memset(r, 0, sizeof(r)); r[5] = flags;
  return ir_run(g_inv, g_inv_len, r, 0, 0, 0, states, 0);
}
static uint32_t ir_affected(const uint32_t* states, uint32_t flags, uint32_t code) {
  uint32_t r[16]; memset(r, 0, sizeof(r)); r[5] = flags;
  uint32_t aff = 0;
  uint32_t got = ir_run(g_inv, g_inv_len, r, 0, 0, 0, states, &aff);
  return got == code ? aff : 0;
}
const oracle_model ORACLE_IR_MODEL = { DEMI_MODEL_IR, DEMI_IR_ACTORS, DEMI_IR_STATE_WORDS, ir_init, ir_receive, ir_invariant, ir_affected };
int oracle_ir_loaded(void) { return g_blob != 0; }
