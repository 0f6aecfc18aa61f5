// mcx_pcr.cu -- build --remove-pcr on the device.
//
// Replaces seq_reads_are_novel + the remove_pcr_dups branch of build_graph_from_reads_mt
// (src/tools/build_graph.c:35-92,192-231) for a batch of reads resident in device memory
// (LINES layout + line offsets + one mate byte per read).  Three launches, then the normal build:
//   orient : reads flagged MCX_MATE_REVCOMP are reverse-complemented in place (one warp per read)
//   mark   : one thread per read finds its first contig start, find-or-inserts that k-mer in the big
//            table (the reference inserts it too, before it knows whether the read is kept) and
//            atomicMin()s the read's ordinal into first[2 * slot + orient]
//   mask   : one warp per read; a duplicate read (pair) has its bases overwritten with 'N', so the
//            build kernels that follow see a read of the same length without a single k-mer
// mcx_pcr.cuh has the argument for why first[] + compare equals the reference's read-by-read bit test.
#include <cuda_runtime.h>
#include <stdint.h>
#include "mcx_build.h"
#include "mcx_pcr.cuh"

__global__ void __launch_bounds__(256) mcx_pcr_orient_kernel(uint8_t *seq, uint8_t *qual, const uint64_t *__restrict__ off,
                                                             const uint8_t *__restrict__ mate, uint64_t nreads)
{
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  for(uint64_t r = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nreads; r += nwarps) {
    if(!(mate[r] & MCX_MATE_REVCOMP)) continue;
    const uint64_t lo = off[r], len = off[r + 1] - lo - 1;   // the read's terminator is not part of it
    mcx_pcr_revcomp_lanes(seq + lo, qual ? qual + lo : nullptr, len, lane, 32u);
  }
}

template <int W>
__global__ void __launch_bounds__(256) mcx_pcr_mark_kernel(const uint8_t *__restrict__ seq, const uint8_t *__restrict__ qual,
                                                           const uint64_t *__restrict__ off, const uint8_t *__restrict__ mate,
                                                           uint64_t nreads, uint32_t k, uint32_t qcut, uint32_t hp, McxTable t,
                                                           uint32_t *first, uint64_t *node, uint32_t ord_base,
                                                           unsigned long long *counters)
{
  uint32_t n_novel = 0, full = 0;
  for(uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < nreads; r += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t lo = off[r], len = off[r + 1] - lo - 1;
    const uint64_t start = mcx_first_contig_start(seq + lo, qual ? qual + lo : nullptr, len, k, qcut, hp);
    uint64_t nd = MCX_PCR_NONE;
    if(start < len) {
      uint32_t orient;
      const McxKmer<W> key = mcx_kmer_key<W>(mcx_kmer_from_ascii<W>(seq + lo + start, k), k, &orient);
      int novel = 0, isfull = 0;
      uint32_t *s = mcx_table_slot<W>(t, key, true, &novel, &isfull);
      n_novel += (uint32_t)novel; full |= (uint32_t)isfull;
      if(s) {
        nd = 2ull * ((uint64_t)(s - t.slots) / t.stride) + orient;
        atomicMin(&first[nd], ord_base + (uint32_t)mcx_pcr_leader(r, mate[r]));
      }
    }
    node[r] = nd;
  }
  for(int sh = 16; sh > 0; sh >>= 1) {
    n_novel += __shfl_xor_sync(0xFFFFFFFFu, n_novel, sh);
    full |= __shfl_xor_sync(0xFFFFFFFFu, full, sh);
  }
  if((threadIdx.x & 31u) == 0) {
    if(n_novel) atomicAdd(&counters[MCX_CNT_NOVEL], (unsigned long long)n_novel);
    if(full) atomicOr(&counters[MCX_CNT_FULL], 1ull);
  }
}

__global__ void __launch_bounds__(256) mcx_pcr_mask_kernel(uint8_t *seq, const uint64_t *__restrict__ off,
                                                           const uint8_t *__restrict__ mate, uint64_t nreads,
                                                           const uint64_t *__restrict__ node, const uint32_t *__restrict__ first,
                                                           uint32_t ord_base, unsigned long long *counters)
{
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  for(uint64_t r = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nreads; r += nwarps) {
    if(!mcx_pcr_is_dup(r, mate, node, first, ord_base)) continue;   // warp-uniform
    const uint64_t lo = off[r], len = off[r + 1] - lo - 1;
    for(uint64_t i = lane; i < len; i += 32u) seq[lo + i] = 'N';
    const uint32_t kind = mate[r] & MCX_MATE_KIND;
    if(lane == 0 && kind != MCX_MATE_SECOND) atomicAdd(&counters[kind == MCX_MATE_FIRST ? MCX_CNT_DUP_PE : MCX_CNT_DUP_SE], 1ull);
  }
}

static unsigned pcr_grid(uint64_t items_per_block_units)
{
  int dev = 0, sms = 148; cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const uint64_t cap = (uint64_t)sms * 8;
  return (unsigned)(items_per_block_units < 1 ? 1 : (items_per_block_units < cap ? items_per_block_units : cap));
}

cudaError_t mcx_launch_pcr_filter(uint8_t *seq, uint8_t *qual, const uint64_t *off, const uint8_t *mate, uint64_t nreads,
                                  uint32_t k, uint32_t qcut, uint32_t hp, const McxTable &t, uint32_t *first, uint64_t *node,
                                  uint32_t ord_base, unsigned long long *counters, cudaStream_t st)
{
  if(nreads == 0) return cudaSuccess;
  McxTable big = t; big.front = nullptr; big.front_cnt = nullptr; big.front_set_bits = 0;
  const unsigned gw = pcr_grid((nreads + 7) / 8), gt = pcr_grid((nreads + 255) / 256);
  mcx_pcr_orient_kernel<<<gw, 256, 0, st>>>(seq, qual, off, mate, nreads);
  if(k <= 31) mcx_pcr_mark_kernel<1><<<gt, 256, 0, st>>>(seq, qual, off, mate, nreads, k, qcut, hp, big, first, node, ord_base, counters);
  else mcx_pcr_mark_kernel<2><<<gt, 256, 0, st>>>(seq, qual, off, mate, nreads, k, qcut, hp, big, first, node, ord_base, counters);
  mcx_pcr_mask_kernel<<<gw, 256, 0, st>>>(seq, off, mate, nreads, node, first, ord_base, counters);
  return cudaGetLastError();
}
