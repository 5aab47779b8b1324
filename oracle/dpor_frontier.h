/* ORACLE — test infrastructure only (see machine.h header). */
#ifndef ORACLE_DPOR_FRONTIER_H
#define ORACLE_DPOR_FRONTIER_H
#include "../include/demi_b200.h"
struct om_machine;
int  oracle_in_frontier_mode(void);
void front_om_send(struct om_machine* m, int src, int dst, uint8_t type, uint32_t p0, uint32_t p1);
void front_om_schedule(struct om_machine* m, int self, uint8_t type, uint32_t p0, uint32_t p1, int repeating);
void front_om_cancel(struct om_machine* m, int self, uint8_t type, uint32_t p0, uint32_t p1);
/* demi_dpor_frontier restated on the CPU; n_ranks > 1 simulates the steal protocol rank by rank.
 * results / viol / hashes are per rank (rank r at index r, r*cap_viol, r*cap_hashes). */
int oracle_dpor_frontier(const demi_config* cfg, const demi_ext_event* ext, uint32_t n_ext,
                         const demi_frontier_params* F, uint32_t n_ranks, demi_frontier_result* results,
                         demi_dpor_violation* viol, uint32_t cap_viol, uint64_t* hashes, uint64_t cap_hashes);
#endif
