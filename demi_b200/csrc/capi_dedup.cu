// capi_dedup.cu — C ABI for K5 (state-hash dedup) + K4 (stable compaction).
#include "dedup_kernel.cuh"
#include "engine.hpp"

using namespace demi;

extern "C" int32_t demi_dedup_compact_dev(demi_handle* h, const void* results_dev, uint64_t n, int32_t mode,
                                          void* out_records_dev, void* out_index_dev, void* out_count_dev, void* stream) {
  if (!h) return DEMI_ERR_INVALID;
  if (!results_dev || !out_records_dev || !out_count_dev) return fail(h, DEMI_ERR_INVALID, "demi_dedup_compact_dev: null buffer");
  if (mode != DEMI_DM_UNIQUE && mode != DEMI_DM_VIOLATING) return fail(h, DEMI_ERR_INVALID, "demi_dedup_compact_dev: bad mode");
  if (n > 0xFFFFFFFFull) return fail(h, DEMI_ERR_INVALID, "demi_dedup_compact_dev: n > 2^32-1");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) { CUDA_TRY(h, cudaMemsetAsync(out_count_dev, 0, 8, s)); return DEMI_OK; }
  DedupScratch& S = h->dedup;
  uint64_t slots = 1024;
  while (slots < 2 * n) slots <<= 1;
  const uint32_t n_blocks = (uint32_t)((n + DD_BLOCK - 1) / DD_BLOCK);
  int32_t rc;
  if (mode == DEMI_DM_UNIQUE) {
    if ((rc = ensure_bytes(h, &S.keys, &S.keys_b, slots * sizeof(DDSlot))) != DEMI_OK) return rc;
    CUDA_TRY(h, cudaMemsetAsync(S.keys, 0xFF, slots * sizeof(DDSlot), s));
  }
  if ((rc = ensure_bytes(h, &S.keep, &S.keep_b, n)) != DEMI_OK) return rc;
  if ((rc = ensure_bytes(h, &S.counts, &S.counts_b, (size_t)n_blocks * 4)) != DEMI_OK) return rc;
  const demi_fuzz_result* rec = (const demi_fuzz_result*)results_dev;
  if (mode == DEMI_DM_UNIQUE) {
    int grid = (int)std::min<uint64_t>(n_blocks, (uint64_t)h->sm_count * 8);
    dedup_insert_kernel<<<grid, DD_BLOCK, 0, s>>>(rec, n, (DDSlot*)S.keys, slots);
  }
  dedup_flag_kernel<<<n_blocks, DD_BLOCK, 0, s>>>(rec, n, mode, (const DDSlot*)S.keys, slots,
                                                  (uint8_t*)S.keep, (uint32_t*)S.counts);
  dedup_scan_kernel<<<1, 1024, 0, s>>>((uint32_t*)S.counts, n_blocks, (unsigned long long*)out_count_dev);
  compact_kernel<<<n_blocks, DD_BLOCK, 0, s>>>(rec, n, (const uint8_t*)S.keep, (const uint32_t*)S.counts,
                                               (demi_fuzz_result*)out_records_dev, (uint32_t*)out_index_dev);
  CUDA_TRY(h, cudaGetLastError());
  h->perf.kernel_launches = mode == DEMI_DM_UNIQUE ? 4 : 3;
  return DEMI_OK;
}

extern "C" int32_t demi_dedup_compact(demi_handle* h, const demi_fuzz_result* results, uint64_t n, int32_t mode,
                                      demi_fuzz_result* out_records, uint32_t* out_index, uint64_t* out_count) {
  if (!h) return DEMI_ERR_INVALID;
  if (!results || !out_records || !out_count) return fail(h, DEMI_ERR_INVALID, "demi_dedup_compact: null buffer");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  *out_count = 0;
  if (n == 0) return DEMI_OK;
  void *d_in = 0, *d_out = 0, *d_idx = 0, *d_cnt = 0;
  cudaError_t e = cudaMalloc(&d_in, n * sizeof(demi_fuzz_result));
  if (e == cudaSuccess) e = cudaMalloc(&d_out, n * sizeof(demi_fuzz_result));
  if (e == cudaSuccess) e = cudaMalloc(&d_idx, n * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_cnt, 8);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, results, n * sizeof(demi_fuzz_result), cudaMemcpyHostToDevice, h->stream);
  int32_t rc = DEMI_OK;
  if (e == cudaSuccess) rc = demi_dedup_compact_dev(h, d_in, n, mode, d_out, d_idx, d_cnt, h->stream);
  unsigned long long cnt = 0;
  if (e == cudaSuccess && rc == DEMI_OK) e = cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && rc == DEMI_OK) e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess && rc == DEMI_OK && cnt) {
    e = cudaMemcpy(out_records, d_out, cnt * sizeof(demi_fuzz_result), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && out_index) e = cudaMemcpy(out_index, d_idx, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost);
  }
  cudaFree(d_in); cudaFree(d_out); cudaFree(d_idx); cudaFree(d_cnt);
  if (e != cudaSuccess) return fail(h, DEMI_ERR_CUDA, "demi_dedup_compact: %s", cudaGetErrorString(e));
  if (rc != DEMI_OK) return rc;
  *out_count = cnt;
  return DEMI_OK;
}
