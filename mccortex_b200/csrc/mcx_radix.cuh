// mcx_radix.cuh -- in-tree stable LSD radix sort of (u64 key, u64 value) pairs on the device.
//
// What it is for: the ascending-key order of the .ctx dump.  The reference sorts an array of pointers into its hash
// table with qsort and an indirect compare (HASH_ITERATE_SORTED, src/graph/hash_table.h:115-120,
// src/graph/hash_table.c:362-374; `mccortex sort`: src/commands/ctx_sort.c:117-155); here the (key word, slot index)
// pairs of the occupied slots are radix sorted, 8 bits per pass over the key bits that can differ (2k for the table
// export, all 64 for `sort`).  Round 1 called cub::DeviceRadixSort for this; this file replaces that library call.
//
// One pass over digit d = (key >> shift) & 255, three steps:
//   1. histogram   every block counts the digits of its tile (4096 pairs) -> hist[digit][block]
//   2. scan        exclusive prefix sum over hist in (digit, block) order = where each block's run of each digit starts
//   3. scatter     every block ranks its pairs STABLY (tile order = warp, round, lane; per-warp running counts by
//                  __match_any_sync, then a prefix over the warps) and writes them to their final places
// HBM traffic per pass: keys read twice, values once, both written: 40 B per pair.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MCX_RX_THREADS 256u
#define MCX_RX_WARPS (MCX_RX_THREADS / 32u)
#define MCX_RX_ITEMS 16u                                  /* pairs per thread */
#define MCX_RX_TILE (MCX_RX_THREADS * MCX_RX_ITEMS)       /* pairs per block */
#define MCX_RX_SCAN_CHUNK 4096u

// bytes of scratch mcx_radix_sort_pairs needs for n pairs
static inline size_t mcx_radix_scratch_bytes(uint64_t n)
{
  const uint64_t nblk = (n + MCX_RX_TILE - 1) / MCX_RX_TILE;
  const uint64_t nh = 256u * nblk, nchunks = (nh + MCX_RX_SCAN_CHUNK - 1) / MCX_RX_SCAN_CHUNK;
  return (size_t)(nh + nchunks + 16u) * sizeof(unsigned long long);
}

// Sort the n pairs of (keys, vals) by bits [0, end_bit) of the key, ascending, stable.  The passes ping-pong between
// (keys, vals) and (keys_alt, vals_alt); *out_keys / *out_vals say where the result is.  scratch: device memory of
// mcx_radix_scratch_bytes(n) bytes.  Asynchronous on st.
cudaError_t mcx_radix_sort_pairs(uint64_t *keys, uint64_t *vals, uint64_t *keys_alt, uint64_t *vals_alt, uint64_t n, int end_bit,
                                 void *scratch, uint64_t **out_keys, uint64_t **out_vals, cudaStream_t st);
