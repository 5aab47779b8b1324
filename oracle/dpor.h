/* ORACLE — test infrastructure only (see machine.h header). */
#ifndef ORACLE_DPOR_H
#define ORACLE_DPOR_H
#include "../include/demi_b200.h"
struct om_machine;
int  oracle_in_dpor_mode(void);
void dpor_om_send(struct om_machine* m, int src, int dst, uint8_t type, uint32_t p0, uint32_t p1);
void dpor_om_schedule(struct om_machine* m, int self, uint8_t type, uint32_t p0, uint32_t p1, int repeating);
void dpor_om_cancel(struct om_machine* m, int self, uint8_t type, uint32_t p0, uint32_t p1);
/* configuration of RunnerUtils.editDistanceDporDDMin's DPOR instances (RunnerUtils.scala:822-835) */
typedef struct oracle_dpor_opts {
  const uint32_t* init_nodes; uint32_t n_init_nodes;   /* setInitialDepGraph: {src|dst<<8|type<<16, p0, p1, parent} per node */
  const uint32_t* init_trace; uint32_t n_init_trace;   /* setInitialTrace: node ids, root first */
  uint32_t arvind;                                     /* ArvindDistanceOrdering initialised with init_trace */
  uint32_t prioritize_pending;                         /* prioritizePendingUponDivergence */
} oracle_dpor_opts;
void* oracle_dpor_open(const demi_config* cfg, const demi_ext_event* ext, uint32_t n_ext,
                       const demi_dpor_params* P, const oracle_dpor_opts* opts);
int oracle_dpor_test(void* s, int32_t max_distance, demi_dpor_result* out,
                     demi_dpor_violation* viol, uint32_t cap_viol, uint64_t* interleaving_hashes, uint32_t cap_hashes);
void oracle_dpor_close(void* s);
uint32_t oracle_arvind_distance_of(const int32_t* orig_index_of_path, uint32_t n);
int oracle_dpor_search(const demi_config* cfg, const demi_ext_event* ext, uint32_t n_ext,
                       const demi_dpor_params* P, demi_dpor_result* out,
                       demi_dpor_violation* viol, uint32_t cap_viol, uint64_t* interleaving_hashes, uint32_t cap_hashes);
#endif
