// mcx_lookup.cu -- `build --intersect`: reads may only touch k-mers that are already in the graph.
//
// Replaces (reference, relative to /root/reference):
//   build_graph_from_str_mt(..., must_exist_in_graph = true) / _find_or_insert   src/tools/build_graph.c:99-150
//   load_read statistics in that mode                                             src/tools/build_graph.c:173-181
//   db_graph_remove_no_covg_kmers, db_graph_intersect_edges                       src/graph/db_graph.c:632-673
// With must_exist_in_graph a window's k-mer is looked up, never inserted; coverage is added only if it
// is found, and an edge joins two consecutive windows of a contig only if BOTH k-mers were found.  So
// an occurrence's edge mask depends on the lookups of its neighbours: every chunk is done in two
// passes with the slot of each window (or "none") kept in shared memory in between.  Contigs, and the
// bases they span, are counted as without the flag (the reference does the same).
// This path is not the hot path (no front table, no TMA pipeline): plain cooperative loads.
#include <cuda_runtime.h>
#include <stdint.h>
#include "mcx_chunk.cuh"
#include "mcx_table.cuh"
#include "mcx_build.h"

#define LK_THREADS 256
#define LK_NONE 0xFFFFFFFFFFFFFFFFull

struct __align__(16) McxLookupSmem {
  uint8_t raw[MCX_RAW];
  uint8_t qraw[MCX_RAW];      // quality bytes of the same positions (quality cut-off only)
  uint32_t pk[MCX_PKW];
  uint32_t bad[MCX_MSW];
  uint32_t bads[MCX_MSW];     // base cannot be in a window that STARTS a contig (quality cut-off only)
  uint32_t eq[MCX_MSW];
  uint32_t vmask[MCX_VW];
  uint32_t svm[MCX_VW];
  uint64_t slot[MCX_T + 2];   // window -1 .. T of the chunk: slot index | orientation << 63, or LK_NONE
  unsigned long long red[MCX_NCOUNTERS];
};

// in_contig of the window just before `chunk` from the per-chunk carry summaries of pass 1 (same walk as mcx_build.cu)
__device__ __forceinline__ uint32_t lk_carry_in(const uint8_t *summary, uint64_t chunk, uint64_t c_first)
{
  while(chunk > c_first) {
    const uint32_t s = summary[--chunk - c_first];
    if(s == 0u) return 0u;
    if(s == 3u) return 1u;
  }
  return 0u;
}

template <int W, bool QUAL>
__global__ void __launch_bounds__(LK_THREADS) mcx_build_lookup_kernel(McxBuildParams p, McxTable t)
{
  extern __shared__ __align__(16) unsigned char lk_smem[];
  McxLookupSmem &sm = *reinterpret_cast<McxLookupSmem *>(lk_smem);
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  const uint64_t c_first = p.r_begin / MCX_T, c_last = (p.r_end + MCX_T - 1) / MCX_T;
  if(tid < MCX_NCOUNTERS) sm.red[tid] = 0;
  if(tid < 4) {
    sm.pk[MCX_RAW / 16u + tid] = 0; sm.bad[MCX_RAW / 32u + tid] = 0xFFFFFFFFu; sm.eq[MCX_RAW / 32u + tid] = 0;
    sm.bads[MCX_RAW / 32u + tid] = 0xFFFFFFFFu;
  }
  uint64_t n_found = 0, n_notfound = 0, n_contigs = 0, n_reads = 0;
  const uint64_t readable = (p.nbytes + 15ull) & ~15ull;

  for(uint64_t chunk = c_first + blockIdx.x; chunk < c_last; chunk += gridDim.x) {
    const uint64_t cs = chunk * (uint64_t)MCX_T;
    __syncthreads();
    // ---- stage + phase 1: staged byte j <-> buffer offset cs - LB + j
    if(tid < MCX_RAW / 16u) {
      const uint64_t gpos = cs - MCX_LB + tid * 16ull;   // wraps for the look-back of chunk 0
      uint4 v = make_uint4(0, 0, 0, 0);
      if(gpos < readable) v = *reinterpret_cast<const uint4 *>(p.seq + gpos);
      *reinterpret_cast<uint4 *>(&sm.raw[tid * 16u]) = v;
      if(QUAL) {
        uint4 qv = make_uint4(0, 0, 0, 0);
        if(gpos < readable) qv = *reinterpret_cast<const uint4 *>(p.qual + gpos);
        *reinterpret_cast<uint4 *>(&sm.qraw[tid * 16u]) = qv;
      }
    }
    __syncthreads();
    if(tid < MCX_RAW / 16u) {
      const uint4 v = *reinterpret_cast<const uint4 *>(&sm.raw[tid * 16u]);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      const uint32_t prev = tid ? sm.raw[tid * 16u - 1u] : 0u;
      const uint64_t gpos = cs - MCX_LB + tid * 16ull;
      uint32_t pk, b16, e16, n16;
      mcx_convert16(w, prev, gpos, p.nbytes, &pk, &b16, &e16, &n16);
      sm.pk[tid] = pk;
      if(QUAL) {
        const uint4 qv = *reinterpret_cast<const uint4 *>(&sm.qraw[tid * 16u]);
        const uint32_t q[4] = {qv.x, qv.y, qv.z, qv.w};
        uint32_t wk16, st16;
        mcx_qual16(q, p.qcut, &wk16, &st16);
        reinterpret_cast<uint16_t *>(sm.bads)[tid] = (uint16_t)(b16 | st16);
        b16 |= wk16;
      }
      reinterpret_cast<uint16_t *>(sm.bad)[tid] = (uint16_t)b16;
      reinterpret_cast<uint16_t *>(sm.eq)[tid] = (uint16_t)e16;
      if(n16 && tid >= MCX_LB / 16u && tid < (MCX_LB + MCX_T) / 16u)
        for(uint32_t i = 0; i < 16u; i++)
          if(((n16 >> i) & 1u) && gpos + i >= p.r_begin && gpos + i < p.r_end) n_reads++;
    }
    __syncthreads();
    // ---- phase 2a
    if(tid < MCX_VW) {
      const bool live = tid < (MCX_LB + MCX_T + 32u) / 32u;
      sm.vmask[tid] = live ? mcx_valid_word(sm.bad, sm.eq, tid, p.k, p.hp_cutoff) : 0u;
      if(QUAL) sm.svm[tid] = live ? mcx_valid_word(sm.bads, sm.eq, tid, p.k, p.hp_cutoff) : 0u;
    }
    __syncthreads();
    if(QUAL) {
      // vmask holds ev, svm holds sv: in_contig = ev & (sv | in_contig(prev)), carry-in from the summaries of pass 1
      if(tid == 0) {
        const uint32_t cb = MCX_LB - 1u, keep = ~0u << cb;
        const uint32_t cin = lk_carry_in(p.summary, chunk, c_first);
        sm.vmask[0] = (sm.vmask[0] & keep & ~(1u << cb)) | (cin << cb);
        sm.svm[0] = (sm.svm[0] & keep & ~(1u << cb)) | (cin << cb);
        mcx_contig_chain(sm.vmask, sm.svm, MCX_VW, 0u, sm.vmask);
      }
      __syncthreads();
    }
    // ---- pass A: look every window of the chunk, and the one on either side, up
    for(uint32_t idx = tid; idx < MCX_T + 2u; idx += LK_THREADS) {
      const uint32_t q = MCX_LB - 1u + idx;
      uint64_t val = LK_NONE;
      if(mcx_get_bit(sm.vmask, q)) {
        uint32_t orient;
        const McxKmer<W> key = mcx_kmer_key<W>(mcx_kmer_at<W>(sm.pk, q, p.k), p.k, &orient);
        int novel = 0, full = 0;
        const uint32_t *s = mcx_table_slot<W>(t, key, false, &novel, &full);
        if(s) val = ((uint64_t)(s - t.slots) / t.stride) | ((uint64_t)orient << 63);
      }
      sm.slot[idx] = val;
    }
    __syncthreads();
    // ---- pass B: coverage and edges of the windows this launch owns
    for(uint32_t idx = 1u + tid; idx <= MCX_T; idx += LK_THREADS) {
      const uint32_t q = MCX_LB - 1u + idx;
      const uint64_t gpos = cs + (idx - 1u);
      if(gpos < p.r_begin || gpos >= p.r_end || !mcx_get_bit(sm.vmask, q)) continue;
      const bool prev_valid = mcx_get_bit(sm.vmask, q - 1u) != 0, next_valid = mcx_get_bit(sm.vmask, q + 1u) != 0;
      n_contigs += !prev_valid;
      const uint64_t val = sm.slot[idx];
      if(val == LK_NONE) { n_notfound++; continue; }
      n_found++;
      const uint32_t orient = (uint32_t)(val >> 63);
      uint32_t *s = t.slots + (val & ~(1ull << 63)) * (uint64_t)t.stride;
      mcx_covg_add(s + 2u * W + p.colour, 1u, true);
      const bool has_prev = prev_valid && sm.slot[idx - 1u] != LK_NONE, has_next = next_valid && sm.slot[idx + 1u] != LK_NONE;
      const uint32_t emask = mcx_edge_mask(orient, has_prev, mcx_get_base(sm.pk, q - 1u), has_next, mcx_get_base(sm.pk, q + p.k));
      mcx_edges_or(s, W, t.ncols, p.colour, emask, 0, false);
    }
  }
  for(int sh = 16; sh > 0; sh >>= 1) {
    n_found += __shfl_xor_sync(0xFFFFFFFFu, n_found, sh);
    n_notfound += __shfl_xor_sync(0xFFFFFFFFu, n_notfound, sh);
    n_contigs += __shfl_xor_sync(0xFFFFFFFFu, n_contigs, sh);
    n_reads += __shfl_xor_sync(0xFFFFFFFFu, n_reads, sh);
  }
  __syncthreads();
  if(lane == 0) {
    atomicAdd(&sm.red[MCX_CNT_KMERS], (unsigned long long)n_found);
    atomicAdd(&sm.red[MCX_CNT_NOTFOUND], (unsigned long long)n_notfound);
    atomicAdd(&sm.red[MCX_CNT_CONTIGS], (unsigned long long)n_contigs);
    atomicAdd(&sm.red[MCX_CNT_READS], (unsigned long long)n_reads);
  }
  __syncthreads();
  if(tid < MCX_NCOUNTERS && sm.red[tid]) atomicAdd(&p.counters[tid], sm.red[tid]);
}

cudaError_t mcx_launch_build_lookup(const McxBuildParams &p, const McxTable &t, cudaStream_t st)
{
  if(p.r_end <= p.r_begin) return cudaSuccess;
  int dev = 0, sms = 148; cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  uint64_t nch = (p.r_end + MCX_T - 1) / MCX_T - p.r_begin / MCX_T, cap = (uint64_t)sms * 4;
  unsigned grid = (unsigned)(nch < cap ? (nch ? nch : 1) : cap);
  const size_t smem = sizeof(McxLookupSmem);
  McxTable big = t; big.front = nullptr; big.front_cnt = nullptr; big.front_set_bits = 0;
#define LK_LAUNCH(WW, QQ) do { \
    cudaFuncSetAttribute(mcx_build_lookup_kernel<WW, QQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    mcx_build_lookup_kernel<WW, QQ><<<grid, LK_THREADS, smem, st>>>(p, big); } while(0)
  // p.qual set: quality cut-off; p.summary must then hold the carry summaries of mcx_launch_contig_summary
  if(p.k <= 31) { if(p.qual) LK_LAUNCH(1, true); else LK_LAUNCH(1, false); }
  else { if(p.qual) LK_LAUNCH(2, true); else LK_LAUNCH(2, false); }
#undef LK_LAUNCH
  return cudaGetLastError();
}

// end of an intersected build: k-mers that ended up with no coverage in any colour are removed
// (their slot becomes a tombstone: probe chains through it stay intact), every colour's edges are
// ANDed with the intersection graph's edges.  *nkept counts what is left.
__global__ void __launch_bounds__(256) mcx_finish_intersect_kernel(McxTable t, uint32_t W, const uint8_t *__restrict__ isec_edges,
                                                                   unsigned long long *nkept)
{
  uint64_t kept = 0;
  for(uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < t.nslots; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t *s = t.slots + i * (uint64_t)t.stride;
    uint64_t k0 = *reinterpret_cast<uint64_t *>(s);
    if(k0 == 0 || k0 == MCX_KEY_TOMBSTONE) continue;
    uint32_t any = 0;
    for(uint32_t c = 0; c < t.ncols; c++) any |= s[2u * W + c];
    if(!any) {
      *reinterpret_cast<uint64_t *>(s) = MCX_KEY_TOMBSTONE;
      uint8_t *e = reinterpret_cast<uint8_t *>(s + 2u * W + t.ncols);
      for(uint32_t c = 0; c < t.ncols; c++) e[c] = 0;
      continue;
    }
    kept++;
    uint8_t *e = reinterpret_cast<uint8_t *>(s + 2u * W + t.ncols);
    const uint8_t m = isec_edges[i];
    for(uint32_t c = 0; c < t.ncols; c++) e[c] &= m;
  }
  for(int sh = 16; sh > 0; sh >>= 1) kept += __shfl_xor_sync(0xFFFFFFFFu, kept, sh);
  if((threadIdx.x & 31u) == 0 && kept) atomicAdd(nkept, (unsigned long long)kept);
}

cudaError_t mcx_launch_finish_intersect(const McxTable &t, uint32_t W, const uint8_t *isec_edges, unsigned long long *nkept,
                                        cudaStream_t st)
{
  int dev = 0, sms = 148; cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  mcx_finish_intersect_kernel<<<sms * 8, 256, 0, st>>>(t, W, isec_edges, nkept);
  return cudaGetLastError();
}
