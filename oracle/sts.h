/* ORACLE — test infrastructure only (see machine.h header). */
#ifndef ORACLE_STS_H
#define ORACLE_STS_H
#include <stdint.h>
#include <stddef.h>
#include "../include/demi_b200.h"

typedef struct demi_replay_input {
  const demi_event* events; uint32_t n_events;            /* the original EventTrace */
  const demi_ext_event* externals; uint32_t n_externals;   /* EventTrace.original_externals */
  uint32_t external_type_mask;                             /* EventTypes.externalMessageFilter by message type */
  uint32_t pending_cap, tosend_cap;
} demi_replay_input;

int  oracle_sts_project(const demi_replay_input* in, const uint64_t* mask, int filter_known_absents, uint8_t* keep);
void oracle_sts_replay(const demi_config* cfg, const demi_replay_input* in, const uint64_t* mask,
                       uint32_t looking_for, uint32_t flags, demi_replay_result* out, void* scratch);
void oracle_sts_replay_ex(const demi_config* cfg, const demi_replay_input* in, const uint64_t* mask,
                          uint32_t looking_for, uint32_t flags, uint32_t skip_event, demi_replay_result* out,
                          demi_event* rec, uint32_t cap_rec, uint32_t* n_rec, void* scratch);
size_t oracle_sts_scratch_size(void);
int  oracle_in_sts_mode(void);

struct om_machine;
void sts_om_send(struct om_machine* m, int src, int dst, uint8_t type, uint32_t p0, uint32_t p1);
void sts_om_schedule(struct om_machine* m, int self, uint8_t type, uint32_t p0, uint32_t p1, int repeating);
void sts_om_cancel(struct om_machine* m, int self, uint8_t type, uint32_t p0, uint32_t p1);
#endif
