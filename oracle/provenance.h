/* ORACLE — test infrastructure only (see machine.h). */
#ifndef ORACLE_PROVENANCE_H
#define ORACLE_PROVENANCE_H
#include <stdint.h>
#include "../include/demi_b200.h"

/* ProvenanceTracker(initialTrace, depGraph).pruneConcurrentEvents(violation) (schedulers/Util.scala:267-376) */
int oracle_provenance(const demi_event* events, uint32_t n_events, const uint16_t* dep_parent, uint32_t n_nodes,
                      uint32_t affected_mask, uint64_t* keep_mask, uint32_t mask_words, demi_provenance_out* out);
/* RunnerUtils.pruneConcurrentEvents on the execution of one seed (RunnerUtils.scala:138-163) */
int oracle_fuzz_provenance(const demi_config* cfg, const demi_ext_event* ext, uint32_t n_ext,
                           const demi_fuzz_params* p, int64_t seed, uint64_t* keep_mask, uint32_t mask_words,
                           demi_provenance_out* out);
#endif
