/* ORACLE — test infrastructure only (see machine.h header). */
#ifndef ORACLE_DPOR_H
#define ORACLE_DPOR_H
#include "../include/demi_b200.h"
struct om_machine;
int  oracle_in_dpor_mode(void);
void dpor_om_send(struct om_machine* m, int src, int dst, uint8_t type, uint32_t p0, uint32_t p1);
void dpor_om_schedule(struct om_machine* m, int self, uint8_t type, uint32_t p0, uint32_t p1, int repeating);
void dpor_om_cancel(struct om_machine* m, int self, uint8_t type, uint32_t p0, uint32_t p1);
int oracle_dpor_search(const demi_config* cfg, const demi_ext_event* ext, uint32_t n_ext,
                       const demi_dpor_params* P, demi_dpor_result* out,
                       demi_dpor_violation* viol, uint32_t cap_viol, uint64_t* interleaving_hashes, uint32_t cap_hashes);
#endif
