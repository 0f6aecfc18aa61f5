// mcx_ctxload.cu -- merge the records of a .ctx graph file into the device table.
//
// Replaces (reference, relative to /root/reference):
//   graph_load                       src/graph/graphs_load.c:83-208
//   graph_file_read / _read_reset    src/graph/graph_file_reader.c:389-413   (colour filter applied per record)
//   db_node_add_col_covg             src/graph/db_node.h                     (saturating add)
// One thread per record.  A record is W x u64 key, C_file x u32 covg, C_file x u8 edges, packed
// (8W + 5 C_file bytes, so only byte aligned).  The file filter is a list of (from, into) colour
// pairs; several file colours may land in one graph colour (the reference adds their coverages
// with saturation and ORs their edges before touching the graph -- adding them one after the other
// with a saturating add gives the same result).  A k-mer whose selected colours all have zero
// coverage is skipped (graphs_load.c:121-124).  MCX_LOAD_MUST_EXIST: never insert, k-mers that are
// not in the table are skipped (GraphLoadingPrefs.must_exist_in_graph).  Flag bit 1: the file is an
// intersection graph (edges into isec_edges, no coverage); bit 2: edges are ANDed with isec_edges.
#include <cuda_runtime.h>
#include <stdint.h>
#include "mcx_build.h"

__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t *p)
{
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ uint64_t ld_u64_unaligned(const uint8_t *p)
{
  return (uint64_t)ld_u32_unaligned(p) | ((uint64_t)ld_u32_unaligned(p + 4) << 32);
}

template <int W>
__global__ void __launch_bounds__(256) mcx_load_records_kernel(const uint8_t *__restrict__ recs, uint64_t n, uint32_t file_ncols,
                                                               const uint32_t *__restrict__ from_col, const uint32_t *__restrict__ into_col,
                                                               uint32_t nmap, uint32_t flags, McxTable t, uint8_t *isec_edges,
                                                               unsigned long long *counters)
{
  const uint32_t rec_bytes = 8u * W + 5u * file_ncols;
  uint64_t n_loaded = 0, n_novel = 0; uint32_t full = 0;
  for(uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint8_t *r = recs + i * rec_bytes;
    const uint8_t *cv = r + 8u * W, *ed = cv + 4u * file_ncols;
    uint32_t keep = 0;
    for(uint32_t m = 0; m < nmap; m++) keep |= ld_u32_unaligned(cv + 4u * from_col[m]);
    if(!keep) continue;
    McxKmer<W> key;
#pragma unroll
    for(int w = 0; w < W; w++) key.b[w] = ld_u64_unaligned(r + 8 * w);
    int novel = 0, isfull = 0;
    uint32_t *s = mcx_table_slot<W>(t, key, !(flags & 1u), &novel, &isfull);
    full |= (uint32_t)isfull;
    if(!s) continue;
    n_novel += novel; n_loaded++;
    const uint64_t slot_idx = (uint64_t)(s - t.slots) / t.stride;
    if(flags & 2u) {
      // intersection graph (ctx_build.c:348-361): no coverage, every selected colour's edges into the one
      // edge set the build is intersected with at the end
      uint32_t e = 0;
      for(uint32_t m = 0; m < nmap; m++) e |= ed[from_col[m]];
      uint32_t *w = reinterpret_cast<uint32_t *>(isec_edges + (slot_idx & ~3ull));
      if(e) atomicOr(w, e << (8u * (uint32_t)(slot_idx & 3ull)));
      continue;
    }
    const uint32_t emask = (flags & 4u) ? isec_edges[slot_idx] : 0xFFu;   // GraphLoadingPrefs.must_exist_in_edges
    for(uint32_t m = 0; m < nmap; m++) {
      const uint32_t c = ld_u32_unaligned(cv + 4u * from_col[m]), e = ed[from_col[m]] & emask, into = into_col[m];
      mcx_covg_add(s + 2u * W + into, c, true);
      mcx_edges_or(s, W, t.ncols, into, e, 0, false);
    }
  }
  for(int sh = 16; sh > 0; sh >>= 1) {
    n_loaded += __shfl_xor_sync(0xFFFFFFFFu, n_loaded, sh);
    n_novel += __shfl_xor_sync(0xFFFFFFFFu, n_novel, sh);
    full |= __shfl_xor_sync(0xFFFFFFFFu, full, sh);
  }
  if((threadIdx.x & 31u) == 0) {
    if(n_loaded) atomicAdd(&counters[MCX_CNT_RECS_LOADED], (unsigned long long)n_loaded);
    if(n_novel) atomicAdd(&counters[MCX_CNT_NOVEL], (unsigned long long)n_novel);
    if(full) atomicOr(&counters[MCX_CNT_FULL], 1ull);
  }
}

cudaError_t mcx_launch_load_records(const uint8_t *recs, uint64_t n, uint32_t file_ncols, const uint32_t *from_col,
                                    const uint32_t *into_col, uint32_t nmap, uint32_t flags, uint32_t k, const McxTable &t,
                                    uint8_t *isec_edges, unsigned long long *counters, cudaStream_t st)
{
  if(n == 0 || nmap == 0) return cudaSuccess;
  int dev = 0, sms = 148; cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  uint64_t want = (n + 255) / 256, cap = (uint64_t)sms * 8;
  unsigned grid = (unsigned)(want < cap ? want : cap);
  McxTable big = t; big.front = nullptr; big.front_cnt = nullptr; big.front_set_bits = 0;
  if(k <= 31) mcx_load_records_kernel<1><<<grid, 256, 0, st>>>(recs, n, file_ncols, from_col, into_col, nmap, flags, big, isec_edges, counters);
  else mcx_load_records_kernel<2><<<grid, 256, 0, st>>>(recs, n, file_ncols, from_col, into_col, nmap, flags, big, isec_edges, counters);
  return cudaGetLastError();
}
