/* A host with nothing but a C compiler and dlopen(): the whole pipeline through the C ABI of include/demi_b200.h —
 * RunnerUtils.fuzz (:62-147) -> the violating EventTrace -> DDMin over STSSched (stsSchedDDMin :642-707) -> verify,
 * then one DPORwHeuristics search as a frontier.  This is what the JNI shim (jni/DemiNative.c) forwards to.
 * Without a CUDA device it checks the no-CPU-fallback contract instead. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "demi_b200.h"

#define SYM(name) __typeof__(&name) p_##name = (__typeof__(&name))dlsym(lib, #name); \
  if (!p_##name) { printf("missing symbol %s\n", #name); return 1; }
#define CHECK(c) do { if (!(c)) { printf("CHECK failed: %s (line %d): %s\n", #c, __LINE__, h ? p_demi_last_error(h) : p_demi_last_error(NULL)); return 1; } } while (0)

int main(int argc, char** argv) {
  if (argc < 2) { printf("usage: abi_pipeline <path to libdemi_b200.so>\n"); return 2; }
  void* lib = dlopen(argv[1], RTLD_NOW);
  if (!lib) { printf("dlopen: %s\n", dlerror()); return 1; }
  SYM(demi_version) SYM(demi_last_error) SYM(demi_device_count) SYM(demi_create) SYM(demi_destroy) SYM(demi_set_externals)
  SYM(demi_fuzz_batch) SYM(demi_fuzz_trace) SYM(demi_set_trace) SYM(demi_ddmin) SYM(demi_replay_batch) SYM(demi_dpor_frontier)
  SYM(demi_stats)
  printf("%s\n", p_demi_version());
  demi_handle* h = NULL;
  demi_config cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.model = DEMI_MODEL_RAFT5; cfg.model_flags = 1;
  if (p_demi_device_count() == 0) {
    int32_t rc = p_demi_create(&cfg, &h);
    if (rc != DEMI_ERR_NO_DEVICE || h) { printf("expected DEMI_ERR_NO_DEVICE, got %d\n", rc); return 1; }
    printf("no device: %s\nOK (no-fallback contract)\n", p_demi_last_error(NULL));
    return 0;
  }
  CHECK(p_demi_create(&cfg, &h) == DEMI_OK);
  /* Start x5, bootstrap Send x5, six client commands, WaitQuiescence */
  demi_ext_event ext[17]; memset(ext, 0, sizeof(ext));
  uint32_t n = 0;
  for (int a = 0; a < 5; a++) { ext[n].kind = DEMI_EXT_START; ext[n].a = (uint8_t)a; ext[n].id = n + 1; n++; }
  for (int a = 0; a < 5; a++) { ext[n].kind = DEMI_EXT_SEND; ext[n].a = (uint8_t)a; ext[n].type = 1; ext[n].p0 = 0x1F; ext[n].id = n + 1; n++; }
  for (int k = 0; k < 6; k++) { ext[n].kind = DEMI_EXT_SEND; ext[n].a = (uint8_t)(k % 5); ext[n].type = 2; ext[n].p0 = 1 + k; ext[n].id = n + 1; n++; }
  ext[n].kind = DEMI_EXT_WAIT_QUIESCENCE; ext[n].id = n + 1; n++;
  CHECK(p_demi_set_externals(h, ext, n) == DEMI_OK);
  /* fuzz until a violation */
  enum { N = 20000 };
  demi_fuzz_params fp; memset(&fp, 0, sizeof(fp));
  fp.seed_base = 1; fp.n_prefixes = N; fp.max_messages = 50; fp.invariant_check_interval = 5;
  demi_fuzz_result* res = (demi_fuzz_result*)malloc(sizeof(demi_fuzz_result) * N);
  CHECK(p_demi_fuzz_batch(h, &fp, res) == DEMI_OK);
  int64_t seed = 0; uint32_t code = 0, violating = 0;
  for (uint32_t i = 0; i < N; i++) if (res[i].violation) { violating++; if (!seed) { seed = 1 + i; code = res[i].violation; } }
  CHECK(seed != 0);
  /* the violating execution's EventTrace */
  enum { CAP = 4096 };
  demi_event* ev = (demi_event*)malloc(sizeof(demi_event) * CAP);
  uint16_t* par = (uint16_t*)malloc(sizeof(uint16_t) * CAP);
  uint32_t n_ev = 0, n_nodes = 0; demi_fuzz_result one;
  CHECK(p_demi_fuzz_trace(h, &fp, seed, ev, CAP, &n_ev, par, CAP, &n_nodes, &one) == DEMI_OK);
  CHECK(one.violation == code && n_ev > 0 && n_ev <= CAP);
  printf("fuzz: %u of %d prefixes violate; seed %lld: violation %u, %u events, %u dep-graph nodes\n", violating, (int)N,
         (long long)seed, code, n_ev, n_nodes);
  /* DDMin with STSSched as the oracle, then verify_mcs by replaying the MCS alone */
  CHECK(p_demi_set_trace(h, ev, n_ev, ext, n) == DEMI_OK);
  uint64_t mcs[1] = {0}; uint32_t iters[256]; demi_ddmin_out dd;
  CHECK(p_demi_ddmin(h, code, 0, 1, mcs, 1, iters, 256, &dd) == DEMI_OK);
  CHECK(dd.verified == 1 && dd.mcs_size >= 1 && dd.mcs_size < n);
  demi_replay_result rr;
  CHECK(p_demi_replay_batch(h, mcs, 1, 1, code, 0, &rr) == DEMI_OK);
  CHECK(rr.violation == code);
  printf("ddmin: %u externals -> MCS of %u (mask %#llx), %u sequential tests, %u evaluated in %u batches, verified\n", n - 1,
         dd.mcs_size, (unsigned long long)mcs[0], dd.total_replays, dd.replays_executed, dd.batches);
  /* one DPORwHeuristics search over the MCS's Start/Send events, explored as a frontier */
  demi_ext_event dext[17]; uint32_t nd = 0;
  for (uint32_t i = 0; i < n; i++)
    if (((mcs[0] >> i) & 1u) && (ext[i].kind == DEMI_EXT_START || ext[i].kind == DEMI_EXT_SEND)) dext[nd++] = ext[i];
  demi_frontier_params F; memset(&F, 0, sizeof(F));
  F.max_messages = 40; F.width = 1024; F.max_interleavings = 20000; F.explored_slots = 1u << 18; F.pool_cap = 1u << 20;
  F.trace_cap = 21100; F.rounds_per_exchange = 1; F.steal_max = 0;
  demi_frontier_result fr; demi_dpor_violation viol[16];
  CHECK(p_demi_dpor_frontier(h, dext, nd, &F, &fr, viol, 16, NULL, 0) == DEMI_OK);
  CHECK(fr.status == 0 && fr.interleavings >= 1 && (fr.exhausted || fr.budget_exhausted));
  printf("dpor frontier over the MCS: %llu interleavings in %u rounds, %llu races, %llu violating, %s\n",
         (unsigned long long)fr.interleavings, fr.rounds, (unsigned long long)fr.races, (unsigned long long)fr.violations,
         fr.exhausted ? "backtrack set empty" : "budget reached");
  demi_perf perf; CHECK(p_demi_stats(h, &perf) == DEMI_OK);
  p_demi_destroy(h);
  free(res); free(ev); free(par);
  printf("OK\n");
  return 0;
}
