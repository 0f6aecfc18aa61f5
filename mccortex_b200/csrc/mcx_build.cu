// mcx_build.cu -- the build hot path on sm_100a.
//
// Kernel A  mcx_build_fused_kernel<W>: reads -> k-mers -> canonical key -> Lookup3 ->
//           find-or-insert -> covg++ -> edge OR, in one pass (single-GPU path; no tuple
//           round trip through HBM).
//           mcx_build_sharded_kernel<W>: the same with owner routing in the parked pass (one shard per GPU).
//           mcx_build_fused_qual_kernel / mcx_build_sharded_qual_kernel + mcx_contig_summary_kernel: quality cut-off.
// Kernel B  mcx_kmer_tuples_kernel<W>: same front end, but emits (key, edge-mask) tuples
//           binned by owning GPU (the NCCL baseline of the multi-GPU path).
// Kernel C  mcx_insert_tuples_kernel<W>: inserts received tuples into the local shard.
//
// Replaces the reference's per-read CPU loop (relative to /root/reference):
//   build_graph_from_reads_mt / load_read / build_graph_from_str_mt  src/tools/build_graph.c:122-231
//   seq_contig_start2 / seq_contig_end2                              src/basic/seq_reader.c:61-172
//   binary_kmer_from_str / left_shift_add / reverse_complement / get_key  src/graph/binary_kmer.{h,c}
//   bklk3_hashlittle                                                 src/kmer/kmer_hash.h:162-211
//   hash_table_find_or_insert_mt, db_graph_update_node_mt, db_graph_add_edge_mt
//
// Front end (shared by all of them; mcx_front_end): persistent CTAs walk chunks of 2048 window positions of the batch byte
// buffer.  One elected thread streams chunk j+4 global->shared with a 1-D TMA bulk copy (cp.async.bulk + mbarrier
// complete_tx) while the CTA, between two barriers, converts chunk j+2 (ASCII -> 2-bit packed words + bad / equal-to-
// previous bit masks, 16 bytes per thread, SWAR), evaluates the contig rules of chunk j+1 word-parallel into a
// valid-window mask, and rolls forward / reverse-complement k-mers over eight consecutive windows per thread of chunk j:
// canonical key, edge mask from the neighbouring windows' valid bits, four keys at a time into the SINK.
// Sinks: FusedSink (kernel A and the sharded kernels: L2 front table in the hot pass, everything else parked in a
// per-CTA queue and drained cooperatively -- front-table claim, Lookup3 + big table, or a tuple for the owning shard),
// TupleSink (kernel B), NullSink (summary pass of the quality cut-off).  DESIGN.md 3.
#include <cuda_runtime.h>
#include <stdint.h>
#include "mcx_chunk.cuh"
#include "mcx_table.cuh"
#include "mcx_build.h"

#define MCX_THREADS MCX_CTA_THREADS
#define MCX_CTAS(n) ((n) * (256u / MCX_CTA_THREADS))   /* resident CTAs per SM, given for 256-thread CTAs */

// Per-CTA staging of the chunk pipeline.  Three chunks are in flight per CTA (phase 1 of chunk
// j+2, phase 2a of chunk j+1, phase 2b of chunk j run in the SAME barrier interval), hence the
// buffer counts: raw x2 (TMA landing zone, read by phase 1), packed bases x3 (written by phase 1
// of j+2 while phase 2b of j still reads its own), base masks x2 (phase 1 -> phase 2a), window
// masks x2 (phase 2a -> phase 2b).
template <bool QUAL> struct __align__(128) McxChunkSmem {
  uint8_t raw[2][MCX_RAW];
  uint8_t qraw[QUAL ? 2 : 1][QUAL ? MCX_RAW : 16]; // quality bytes of the same positions (quality modes only)
  uint32_t pk[3][MCX_PKW];
  uint32_t bad[2][MCX_MSW];     // base cannot be in a window that EXTENDS a contig
  uint32_t bads[QUAL ? 2 : 1][QUAL ? MCX_MSW : 1]; // base cannot be in a window that STARTS a contig (quality modes)
  uint32_t eq[2][MCX_MSW];
  uint32_t vmask[2][MCX_VW];    // in_contig per window
  uint32_t svm[QUAL ? 2 : 1][QUAL ? MCX_VW : 1];   // start-valid windows (quality modes)
  uint32_t carry_in;
  unsigned long long bar[2];
  unsigned long long red[MCX_NCOUNTERS];
};

// ---------------------------------------------------------------- compile-time knobs
// The values are the measured best; scripts/gpu_variants*.sh time alternative builds of the library made
// with -D<knob>=<value>.  "2" = the two-word (k > 31) kernels.
#ifndef MCX_FUSED_MINB
#define MCX_FUSED_MINB 3      /* resident CTAs per SM the fused kernel is built for, k <= 31 */
#endif
#ifndef MCX_FUSED_MINB2
#define MCX_FUSED_MINB2 3     /* ... k > 31 */
#endif
#ifndef MCX_QUAL_MINB
#define MCX_QUAL_MINB 3       /* ... quality cut-off kernels (4 spills and does not fit: 174 ms against 122) */
#endif
#ifndef MCX_FUSED_G
#define MCX_FUSED_G 2         /* front-table probe loads in flight per thread, k <= 31 */
#endif
#ifndef MCX_FUSED_G2
#define MCX_FUSED_G2 2        /* ... k > 31 (1: -4 %, 4: -14 %) */
#endif
#ifndef MCX_QCAP2
#define MCX_QCAP2 (MCX_T + MCX_T / 2u)   /* parked-queue entries, k > 31 */
#endif

// Occurrences that are not a plain front-table hit (first sight of a k-mer, missing edge bit,
// count field filling up, k-mer that lives in the big table) are parked here and handled later,
// one per thread, all lanes busy.  Handling them inline would stall the whole warp on two or
// three dependent memory round trips whenever ANY of its 32 lanes is slow, which is most rounds
// (measured: 100-170 ms instead of 59 ms).  The queue holds two chunks' worth of windows and is
// drained only when the next chunk might not fit (about every tenth chunk on the bench
// workload): draining after every chunk left most warps idle at the barrier (ncu: 27 % of all
// stall samples).
// k > 31: 16-byte keys, so the queue holds one and a half chunks (three CTAs per SM still fit); with its front table it is
// drained every fourth chunk or so, without one (or bypassed) after every chunk.
#define MCX_QCAP(W) ((W) == 1 ? 2u * MCX_T : MCX_QCAP2)
template <int W> struct McxSlowQueue {
  uint64_t key[MCX_QCAP(W) * W];
  uint8_t emask[MCX_QCAP(W)];
  uint32_t n;
  // front-table bypass: queue length at the last chunk barrier, chunks done, and the decision -- on data the
  // front table cannot absorb (a genome much larger than its 8.4 M ways: every k-mer is seen about once per batch) the
  // tag probe, the claim attempt and the re-probe in the parked pass are pure overhead
  uint32_t last_n, epoch, bypass;
  // sharded builds: the drain reserves bin space once per CTA and destination (see FusedSink::drain)
  uint8_t dest[MCX_QCAP(W)];
  uint32_t dcnt[MCX_MAX_PARTS], dfill[MCX_MAX_PARTS];
  unsigned long long dbase[MCX_MAX_PARTS];
};
// the queue lives in dynamic shared memory (static + dynamic exceeds the 48 KB static limit)
extern __shared__ __align__(16) unsigned char mcx_dyn_smem[];
template <int W> __device__ __forceinline__ McxSlowQueue<W> *mcx_queue()
{
  return reinterpret_cast<McxSlowQueue<W> *>(mcx_dyn_smem);
}

// front-end modes
enum { MCX_MODE_PLAIN = 0,   // contig rules are window-local (no quality cut-off)
       MCX_MODE_QUAL = 1,    // quality cut-off: in_contig needs the carry chain + chunk carry-in
       MCX_MODE_QSUM = 2 };  // quality cut-off, pass 1: only the per-chunk carry summary

// ---------------------------------------------------------------- TMA / mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while(!done);
}
// 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <bool QUAL>
__device__ __forceinline__ void issue_chunk_load(McxChunkSmem<QUAL> &sm, const McxBuildParams &p, uint64_t chunk, uint32_t buf)
{
  uint64_t cs = chunk * (uint64_t)MCX_T;
  uint64_t src_off = cs ? cs - MCX_LB : 0;
  uint32_t dst_off = cs ? 0 : MCX_LB;
  uint64_t avail = (p.nbytes - src_off + 15ull) & ~15ull; // p.seq is readable up to nbytes rounded up to 16
  uint32_t want = MCX_RAW - dst_off;
  uint32_t bytes = avail < want ? (uint32_t)avail : want;
  mbar_expect_tx(&sm.bar[buf], QUAL ? 2u * bytes : bytes);
  tma_load_1d(&sm.raw[buf][dst_off], p.seq + src_off, bytes, &sm.bar[buf]);
  if(QUAL) tma_load_1d(&sm.qraw[buf][dst_off], p.qual + src_off, bytes, &sm.bar[buf]);
}

// in_contig of the window just before `chunk`: walk back over the per-chunk summaries
// (bit0 = carry-out if carry-in is 0, bit1 = if it is 1) until one does not depend on its input
__device__ __forceinline__ uint32_t chunk_carry_in(const uint8_t *summary, uint64_t chunk, uint64_t c_first)
{
  while(chunk > c_first) {
    uint32_t s = summary[--chunk - c_first];
    if(s == 0u) return 0u;
    if(s == 3u) return 1u;
  }
  return 0u; // launches start at a read boundary
}

// ---------------------------------------------------------------- sinks
static __host__ __device__ __forceinline__ McxTupleBins mcx_no_bins()
{
  McxTupleBins b;
  for(int i = 0; i < MCX_MAX_PARTS; i++) { b.keys[i] = nullptr; b.meta[i] = nullptr; }
  b.cursor = nullptr; b.cap = 0; b.nparts = 1; b.my_part = 0;
  return b;
}

// what to do with one occurrence.  The table and the bins are the kernel's __grid_constant__ parameters, referenced
// in place (constant bank): the sink itself is a handful of registers.  (The first version carried copies of both and
// lived in local memory: 432 bytes of stack, LDL / STL in the hot loop -- DESIGN.md 8.3 of round 1.)
template <int W, int G, bool SHARDED> struct FusedSink { // G = probe loads kept in flight per thread
  const McxTable &t; const McxTupleBins &bins; // SHARDED: keys owned by another shard leave as tuples
  uint32_t colour; bool may_saturate;
  McxSlowQueue<W> *q;
  unsigned long long *counters;
  bool byp; // this chunk skips the front table (read once per chunk: begin_chunk)
  __device__ __forceinline__ void begin_chunk() { byp = t.front_set_bits && *(volatile uint32_t *)&q->bypass; }

  __device__ __forceinline__ void park(const McxKmer<W> &key, uint32_t emask)
  {
    const uint32_t at = atomicAdd(&q->n, 1u); // < MCX_QCAP(W): a chunk starts with at least MCX_T free entries
#pragma unroll
    for(int w = 0; w < W; w++) q->key[at * W + w] = key.b[w];
    q->emask[at] = (uint8_t)emask;
  }
  __device__ __forceinline__ void consume(const McxKmer<W> keys[MCX_HALF], const uint32_t emasks[MCX_HALF], uint32_t valid,
                                          uint32_t &novel, uint32_t &full)
  {
    (void)novel; (void)full;
    if(byp) {
      // the front table is not absorbing this data: straight to the parked pass, marked so that it skips the front table too
#pragma unroll
      for(uint32_t j = 0; j < MCX_HALF; j++)
        if((valid >> j) & 1u) { McxKmer<W> kk = keys[j]; kk.b[0] |= MCX_KEY_FLAG; park(kk, emasks[j]); }
    } else if(W == 1 && t.front_set_bits) {
      // hot pass: G probe loads in flight per thread, one 32-bit RED per hit; no Lookup3, no
      // big-table access for k-mers that live in the L2-resident front table
      const McxFrontGeom g = mcx_front_geom(t);
#pragma unroll
      for(uint32_t h = 0; h < MCX_HALF; h += G) {
        uint64_t v[G][4]; McxFKey fk[G];
#pragma unroll
        for(uint32_t i = 0; i < G; i++) {
          fk[i] = mcx_fhash(keys[h + i].b[0]);
          if((valid >> (h + i)) & 1u) mcx_ld256(t.front + ((uint64_t)(fk[i].y & g.setmask) << 2), v[i][0], v[i][1], v[i][2], v[i][3]);
        }
#pragma unroll
        for(uint32_t i = 0; i < G; i++) {
          if((valid >> (h + i)) & 1u) {
            if(!mcx_front_hit(g, t.front_cnt + ((uint64_t)(fk[i].y & g.setmask) << 2), fk[i].x, (fk[i].y >> g.S) | g.occ,
                              emasks[h + i] << g.eshift, v[i][0], v[i][1], v[i][2], v[i][3]))
              park(keys[h + i], emasks[h + i]);
          }
        }
      }
    } else if(W == 2 && t.front_set_bits) {
      // the same hot pass over 16-byte tags, two ways per set
      const McxFrontGeom2 g = mcx_front_geom2_bits(t.front_set_bits);
#pragma unroll
      for(uint32_t h = 0; h < MCX_HALF; h += G) {
        uint64_t v[G][4]; McxFKey2 fk[G];
#pragma unroll
        for(uint32_t i = 0; i < G; i++) {
          fk[i] = mcx_fhash2(keys[h + i].b[0], keys[h + i].b[W - 1]);
          if((valid >> (h + i)) & 1u) mcx_ld256(t.front + ((fk[i].y >> g.tshift) << 2), v[i][0], v[i][1], v[i][2], v[i][3]);
        }
#pragma unroll
        for(uint32_t i = 0; i < G; i++) {
          if((valid >> (h + i)) & 1u) {
            if(!mcx_front2_hit(g, t.front_cnt + ((fk[i].y >> g.tshift) << 1), fk[i].x, (fk[i].y & (g.occ - 1ull)) | g.occ,
                               (uint64_t)emasks[h + i] << g.eshift, v[i][0], v[i][1], v[i][2], v[i][3]))
              park(keys[h + i], emasks[h + i]);
          }
        }
      }
    } else {
#pragma unroll
      for(uint32_t j = 0; j < MCX_HALF; j++)
        if((valid >> j) & 1u) park(keys[j], emasks[j]);
    }
  }
  __device__ __forceinline__ void reset() { if(threadIdx.x == 0) { q->n = 0; q->last_n = 0; } }
  // the next chunk may not fit (evaluated per thread just before the step barrier, OR-reduced there)
  __device__ __forceinline__ bool should_drain() const { return q->n > MCX_QCAP(W) - MCX_T; }
  // Thread 0, after the chunk barrier: keep probing the front table, or bypass it for the next chunks?  The signal costs
  // the hot pass nothing: how much the queue grew during the chunk, against the chunk's windows (nwin: in-contig bits of
  // the chunk, counted by warp 0 before the barrier).  Probed, the bench workload parks a tenth of a chunk's windows; data
  // the front table cannot absorb parks all of them.  Above three quarters the next fifteen chunks skip the probe, the
  // sixteenth probes again.  (Other threads may act on the old decision for the first groups of the next chunk: every
  // parked item carries its own mark, so that is harmless.)
  __device__ __forceinline__ void decide(uint32_t nwin)
  {
    if(!t.front_set_bits) return;
    const uint32_t n = q->n, parks = n - q->last_n, e = ++q->epoch;
    q->last_n = n;
    if(q->bypass) { if((e & 15u) == 0u) q->bypass = 0; }
    else q->bypass = (parks * 4u > nwin * 3u && parks > MCX_T / 8u) ? 1u : 0u;
  }
  // one parked occurrence: front table (claim / edge bit), else the big table.  Returns the shard that owns the key if
  // the occurrence has to travel there as a tuple (sharded builds), else MCX_NO_DEST (it has been dealt with).
#define MCX_NO_DEST 0xFFu
  __device__ __forceinline__ uint32_t slow(McxKmer<W> key, uint32_t emask, uint32_t &novel, uint32_t &full)
  {
    bool try_front = t.front_set_bits != 0u;
    if(key.b[0] & MCX_KEY_FLAG) { key.b[0] &= ~MCX_KEY_FLAG; try_front = false; } // parked while the front table was bypassed
    if(try_front && (W == 1 ? mcx_front_add_slow(t, key.b[0], emask) : mcx_front2_add_slow(t, key.b[0], key.b[W - 1], emask)))
      return MCX_NO_DEST; // absorbed by the front table
    uint32_t hb, hc = mcx_lookup3<W>(key, 0u, &hb);
    if(SHARDED) {
      const uint32_t d = mcx_owner(hc, bins.nparts);
      if(d != bins.my_part) return d;
    }
    int r = mcx_table_add<W>(t, key, hc, hb, colour, emask, 1u, may_saturate);
    novel += (r == 1);
    if(r == 2) { full = 1; atomicOr(&counters[MCX_CNT_FULL], 1ull); } // seen at once by every CTA (mcx_front_end stops inserting)
    return MCX_NO_DEST;
  }
  // all threads of the CTA, between two __syncthreads
  __device__ __forceinline__ void drain(uint32_t &novel, uint32_t &full)
  {
    const uint32_t n = q->n;
    if(SHARDED) {
      if(threadIdx.x < MCX_MAX_PARTS) { q->dcnt[threadIdx.x] = 0; q->dfill[threadIdx.x] = 0; }
      __syncthreads();
    }
    for(uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      McxKmer<W> key;
#pragma unroll
      for(int w = 0; w < W; w++) key.b[w] = q->key[i * W + w];
      const uint32_t d = slow(key, q->emask[i], novel, full);
      if(SHARDED) {
        q->dest[i] = (uint8_t)d;
        if(d != MCX_NO_DEST) atomicAdd(&q->dcnt[d], 1u);
      }
    }
    if(SHARDED) {
      // Tuples for other shards.  One reservation per CTA, drain and destination: a cursor bumped once per warp round
      // (the first version) is one L2 address hammered by every warp of the grid -- with two shards the sharded kernel
      // ran 26 % slower than the single-GPU one although it inserts half as many cold k-mers (profiles/r2o_*, r2p_*).
      __syncthreads();
      if(threadIdx.x < bins.nparts && q->dcnt[threadIdx.x])
        q->dbase[threadIdx.x] = atomicAdd(&bins.cursor[threadIdx.x], (unsigned long long)q->dcnt[threadIdx.x]);
      __syncthreads();
      const uint32_t lane = threadIdx.x & 31u;
      for(uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t d = q->dest[i];
        if(d == MCX_NO_DEST) continue;
        // the lanes of a warp that go to the same shard take consecutive slots: their stores coalesce on the wire
        const uint32_t peers = __match_any_sync(__activemask(), d);
        const uint32_t leader = __ffs(peers) - 1u;
        uint32_t first = 0;
        if(lane == leader) first = atomicAdd(&q->dfill[d], (uint32_t)__popc(peers));
        first = __shfl_sync(peers, first, leader);
        const uint64_t at = q->dbase[d] + first + __popc(peers & ((1u << lane) - 1u));
        if(at >= bins.cap) { full = 1; continue; }
        uint64_t *kd = bins.keys[d] + at * W;
#pragma unroll
        for(int w = 0; w < W; w++) kd[w] = (w == 0) ? (q->key[i * W + w] & ~MCX_KEY_FLAG) : q->key[i * W + w];
        bins.meta[d][at] = (1u << 8) | q->emask[i];
      }
    }
  }
};

// tuples binned by owner (kernel B): every occurrence leaves as a tuple
template <int W> struct TupleSink {
  const McxTupleBins &b;
  __device__ __noinline__ void one(McxKmer<W> key, uint32_t emask, uint32_t &full)
  {
    uint32_t hb, hc = mcx_lookup3<W>(key, 0u, &hb);
    mcx_bin_push<W>(b, mcx_owner(hc, b.nparts), key, (1u << 8) | emask, full);
  }
  __device__ __forceinline__ void consume(const McxKmer<W> keys[MCX_HALF], const uint32_t emasks[MCX_HALF], uint32_t valid,
                                          uint32_t &novel, uint32_t &full)
  {
    (void)novel;
#pragma unroll
    for(uint32_t j = 0; j < MCX_HALF; j++)
      if((valid >> j) & 1u) one(keys[j], emasks[j], full);
  }
  __device__ __forceinline__ void drain(uint32_t &, uint32_t &) {}
  __device__ __forceinline__ void reset() {}
  __device__ __forceinline__ bool should_drain() const { return false; }
  __device__ __forceinline__ void decide(uint32_t) {}
  __device__ __forceinline__ void begin_chunk() {}
};

// ---------------------------------------------------------------- contig chain, warp-wide (quality modes)
// The chain x[i] = ev[i] & (sv[i] | x[i-1]) over the MCX_VW mask words of a chunk -- what mcx_contig_chain (mcx_chunk.cuh)
// does serially, word by word.  Each lane owns three consecutive words: it works out its carry-out for carry-in 0 and 1,
// the 32 two-bit functions are composed by a prefix scan, then every lane redoes its words with its real carry-in.  In
// place on vm.  The lane-local pieces are in mcx_chunk.cuh (the CPU emulation runs them too).
__device__ __forceinline__ void mcx_contig_chain_warp(uint32_t *vm, const uint32_t *sv, uint32_t cb, uint32_t cin, uint32_t lane)
{
  uint32_t a[MCX_CHAIN_WPL], b[MCX_CHAIN_WPL], f0, f1;
  mcx_chain_lane_load(vm, sv, cb, cin, lane, a, b);
  mcx_chain_lane_carry(a, b, &f0, &f1);
#pragma unroll
  for(uint32_t d = 1; d < 32u; d <<= 1) { // inclusive scan: (f0, f1) of lanes 0..lane composed
    const uint32_t l0 = __shfl_up_sync(0xFFFFFFFFu, f0, d), l1 = __shfl_up_sync(0xFFFFFFFFu, f1, d);
    if(lane >= d) mcx_chain_compose(l0, l1, &f0, &f1);
  }
  uint32_t c = __shfl_up_sync(0xFFFFFFFFu, f0, 1);
  if(lane == 0) c = 0u;
  mcx_chain_lane_store(vm, a, b, c, lane);
}
// in_contig of window f for carry-in 0 (bit 0) and 1 (bit 1): mcx_summary_lane_* (mcx_chunk.cuh) + a warp max and a vote
__device__ __forceinline__ uint32_t mcx_chunk_summary_warp(const uint32_t *ev, const uint32_t *sv, uint32_t cb, uint32_t f, uint32_t lane)
{
  const int z = __reduce_max_sync(0xFFFFFFFFu, mcx_summary_lane_last_zero(ev, cb, f, lane));
  const uint32_t out0 = __any_sync(0xFFFFFFFFu, mcx_summary_lane_starts(sv, cb, f, z, lane) != 0u) ? 1u : 0u;
  return out0 | ((out0 | (z < 0 ? 1u : 0u)) << 1);
}

// ---------------------------------------------------------------- front end
// One CTA = a persistent worker over chunks c_first + blockIdx.x + j * gridDim.x (j = 0, 1, ...).
// Software pipeline, ONE barrier per chunk: in step s the CTA runs
//   phase 1  of chunk j = s+2  (threads 0 .. RAW/16-1: ASCII -> packed bases + base masks),
//   phase 2a of chunk j = s+1  (the last MCX_VW threads: base masks -> window masks),
//   phase 2b of chunk j = s    (everybody: 8 windows per thread -> sink),
// then __syncthreads, then one thread refills the raw buffer phase 1 has just consumed (TMA for
// chunk s+4).  Steps -2 and -1 fill the pipeline.
template <int W, int MODE, class Sink>
__device__ __forceinline__ void mcx_front_end(const McxBuildParams &p, Sink &sink)
{
  constexpr bool QUAL = MODE != MCX_MODE_PLAIN;
  __shared__ McxChunkSmem<QUAL> sm;
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  const uint64_t c_first = p.r_begin / MCX_T, c_last = (p.r_end + MCX_T - 1) / MCX_T;
  const uint64_t chunk0 = c_first + blockIdx.x, cstride = gridDim.x;

  if(tid == 0) {
    mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if(tid < MCX_NCOUNTERS) sm.red[tid] = 0;
  // over-read padding of the staged arrays must be defined (it is shifted in, then masked off)
  if(tid < 4) {
#pragma unroll
    for(int b = 0; b < 3; b++) sm.pk[b][MCX_RAW / 16u + tid] = 0;
#pragma unroll
    for(int b = 0; b < 2; b++) {
      sm.bad[b][MCX_RAW / 32u + tid] = 0xFFFFFFFFu; sm.eq[b][MCX_RAW / 32u + tid] = 0;
      if(QUAL) sm.bads[b][MCX_RAW / 32u + tid] = 0xFFFFFFFFu;
    }
  }
  __syncthreads();
  if(tid == 0) {
    if(chunk0 < c_last) issue_chunk_load<QUAL>(sm, p, chunk0, 0);
    if(chunk0 + cstride < c_last) issue_chunk_load<QUAL>(sm, p, chunk0 + cstride, 1);
  }

  // per-thread counters: a launch covers < 2^32 positions, so 32 bits are plenty (64-bit ones were spilled: six LDL + six
  // STL per group of eight windows in the first version's hot loop)
  uint32_t n_kmers = 0, n_novel = 0, n_contigs = 0, n_reads = 0;
  uint32_t full = 0;
  uint32_t nwin = 0;    // warp 0: in-contig windows of the chunk just processed
  bool stopped = false; // the table is full (any CTA found out): nothing more is inserted, the launch just runs out

  for(int64_t s = -2;; s++) {
    const uint64_t ch2 = chunk0 + (uint64_t)(s + 2) * cstride;                      // phase 1
    const uint64_t ch1 = chunk0 + (uint64_t)(s + 1) * cstride;                      // phase 2a
    const uint64_t ch0 = chunk0 + (uint64_t)(s < 0 ? 0 : s) * cstride;              // phase 2b
    if(s >= 0 && ch0 >= c_last) break;

    // ---- phase 1 (chunk s+2): ASCII -> packed bases + masks, 16 bytes per thread
    if(ch2 < c_last && tid < MCX_RAW / 16u) {
      const uint32_t j = (uint32_t)(s + 2), rb = j & 1u, pb = j % 3u;
      mbar_wait(&sm.bar[rb], (j >> 1) & 1u);
      const uint64_t cs = ch2 * (uint64_t)MCX_T;
      const uint4 v = *reinterpret_cast<const uint4 *>(&sm.raw[rb][tid * 16u]);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      uint32_t prev = tid ? sm.raw[rb][tid * 16u - 1u] : 0u;
      uint64_t gpos = cs - MCX_LB + tid * 16ull; // wraps for the look-back of chunk 0: handled in convert16
      uint32_t pk, b16, e16, n16;
      mcx_convert16(w, prev, gpos, p.nbytes, &pk, &b16, &e16, &n16);
      sm.pk[pb][tid] = pk;
      if(QUAL) {
        const uint4 qv = *reinterpret_cast<const uint4 *>(&sm.qraw[QUAL ? rb : 0][tid * 16u]);
        const uint32_t q[4] = {qv.x, qv.y, qv.z, qv.w};
        uint32_t wk16, st16;
        mcx_qual16(q, p.qcut, &wk16, &st16);
        reinterpret_cast<uint16_t *>(sm.bads[QUAL ? rb : 0])[tid] = (uint16_t)(b16 | st16);
        b16 |= wk16;
      }
      reinterpret_cast<uint16_t *>(sm.bad[rb])[tid] = (uint16_t)b16;
      reinterpret_cast<uint16_t *>(sm.eq[rb])[tid] = (uint16_t)e16;
      // read terminators owned by this launch and this chunk
      if(MODE != MCX_MODE_QSUM && n16 && tid >= MCX_LB / 16u && tid < (MCX_LB + MCX_T) / 16u) {
        for(uint32_t i = 0; i < 16u; i++)
          if(((n16 >> i) & 1u) && gpos + i >= p.r_begin && gpos + i < p.r_end) n_reads++;
      }
    }

    // ---- phase 2a (chunk s+1): contig rules, 32 windows per thread (word-parallel dilate / erode
    //      of the bad / eq masks); masks are indexed by staged position
    if(s >= -1 && ch1 < c_last && tid >= MCX_THREADS - MCX_VW) {
      const uint32_t wi = tid - (MCX_THREADS - MCX_VW), mb = (uint32_t)(s + 1) & 1u;
      const bool live = wi < (MCX_LB + MCX_T + 32u) / 32u;
      sm.vmask[mb][wi] = live ? mcx_valid_word(sm.bad[mb], sm.eq[mb], wi, p.k, p.hp_cutoff) : 0u;
      if(QUAL) sm.svm[QUAL ? mb : 0][wi] = live ? mcx_valid_word(sm.bads[QUAL ? mb : 0], sm.eq[mb], wi, p.k, p.hp_cutoff) : 0u;
      if(MODE == MCX_MODE_QUAL && wi == 0) sm.carry_in = chunk_carry_in(p.summary, ch1, c_first);
    }

    // ---- phase 2b (chunk s): 8 consecutive windows per thread, rolling k-mers; keys are built
    //      four at a time and handed to the sink, which overlaps its table probes
    if(MODE != MCX_MODE_QSUM && s >= 0) {
      const uint32_t j = (uint32_t)s;
      const uint64_t cs = ch0 * (uint64_t)MCX_T;
      // windows owned by this launch: positions [r_begin, r_end)
      const uint64_t g0 = cs + MCX_WPT * tid;
      const uint32_t lo = p.r_begin > g0 ? (p.r_begin - g0 < MCX_WPT ? (uint32_t)(p.r_begin - g0) : MCX_WPT) : 0u;
      const uint32_t hi = p.r_end > g0 ? (p.r_end - g0 < MCX_WPT ? (uint32_t)(p.r_end - g0) : MCX_WPT) : 0u;
      const uint32_t own = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
      sink.begin_chunk();
      if(tid < 32u) { // in-contig windows of the chunk (its look-back positions included: close enough), for sink.decide
        uint32_t c = 0;
        for(uint32_t w = tid; w < MCX_VW; w += 32u) c += __popc(sm.vmask[j & 1u][w]);
        nwin = __reduce_add_sync(0xFFFFFFFFu, c);
      }
      mcx_thread_occurrences<W>(sm.pk[j % 3u], sm.vmask[j & 1u], tid, p.k,
        [&](const McxKmer<W> *keys, const uint32_t *emasks, uint32_t valid, uint32_t starts, uint32_t j0) {
          valid &= own >> j0;
          if(valid && !stopped) {
            n_kmers += __popc(valid);
            n_contigs += __popc(starts & valid);
            sink.consume(keys, emasks, valid, n_novel, full);
          }
        });
    }
    // the step's only barrier.  It also decides, CTA-uniformly, whether the parked occurrences must
    // be drained now: the last thread to arrive evaluates its predicate after every park of the step
    // (bit 1: some insert has found the table full -- an undersized -n must not turn into hours of full-table scans)
    const int bar_flags = __syncthreads_or((sink.should_drain() ? 1 : 0) |
                                           ((tid == 0 && MODE != MCX_MODE_QSUM && *(volatile unsigned long long *)&p.counters[MCX_CNT_FULL]) ? 2 : 0));
    const int drain_now = bar_flags & 1;
    if(bar_flags & 2) stopped = true;
    // raw[(s+2)&1] has been consumed by phase 1: refill it with chunk s+4
    if(tid == 0) {
      const uint64_t ch4 = chunk0 + (uint64_t)(s + 4) * cstride;
      if(ch4 < c_last) issue_chunk_load<QUAL>(sm, p, ch4, (uint32_t)(s + 4) & 1u);
      sink.decide(nwin);
    }

    if(QUAL && s >= -1 && ch1 < c_last) {
      // vmask holds ev, svm holds sv of chunk s+1: resolve in_contig = ev & (sv | in_contig(prev)).  The window before
      // the chunk (position LB-1) carries the carry-in.  Warp 0 does it (mcx_contig_chain_warp / mcx_chunk_summary_warp):
      // one thread walking the 70 words kept the other 255 at the barrier for ~1.7 k cycles per chunk, twice that in
      // the summary pass.
      if(tid < 32u) {
        const uint32_t mb = (uint32_t)(s + 1) & 1u;
        uint32_t *vm = sm.vmask[mb], *sv = sm.svm[QUAL ? mb : 0];
        if(MODE == MCX_MODE_QUAL) mcx_contig_chain_warp(vm, sv, MCX_LB - 1u, sm.carry_in, tid);
        else {
          // summary of this chunk's own windows: in_contig of its last window for carry-in 0 and 1
          const uint32_t out = mcx_chunk_summary_warp(vm, sv, MCX_LB - 1u, MCX_LB - 1u + MCX_T, tid);
          if(tid == 0) p.summary[ch1 - c_first] = (uint8_t)out;
        }
      }
      __syncthreads();
    }

    // ---- parked (slow) occurrences: only when the queue could overflow during the next chunk
    if(drain_now) {
      if(!stopped) sink.drain(n_novel, full);
      __syncthreads();
      sink.reset();
      __syncthreads();
    }
  }
  __syncthreads();
  if(!stopped) sink.drain(n_novel, full);

  // ---- counters: warp shuffle -> shared -> one global atomic per CTA per counter
  for(int sh = 16; sh > 0; sh >>= 1) {
    n_kmers += __shfl_xor_sync(0xFFFFFFFFu, n_kmers, sh);
    n_novel += __shfl_xor_sync(0xFFFFFFFFu, n_novel, sh);
    n_contigs += __shfl_xor_sync(0xFFFFFFFFu, n_contigs, sh);
    n_reads += __shfl_xor_sync(0xFFFFFFFFu, n_reads, sh);
    full |= __shfl_xor_sync(0xFFFFFFFFu, full, sh);
  }
  if(lane == 0) {
    atomicAdd(&sm.red[MCX_CNT_KMERS], (unsigned long long)n_kmers);
    atomicAdd(&sm.red[MCX_CNT_NOVEL], (unsigned long long)n_novel);
    atomicAdd(&sm.red[MCX_CNT_CONTIGS], (unsigned long long)n_contigs);
    atomicAdd(&sm.red[MCX_CNT_READS], (unsigned long long)n_reads);
    if(full) atomicOr(&sm.red[MCX_CNT_FULL], 1ull);
  }
  __syncthreads();
  if(tid < MCX_NCOUNTERS && sm.red[tid]) {
    if(tid == MCX_CNT_FULL) atomicOr(&p.counters[tid], 1ull);
    else atomicAdd(&p.counters[tid], sm.red[tid]);
  }
}

// k <= 31: 3 CTAs x 256 threads per SM, two probe loads in flight per thread (other occupancies / depths were measured in
// round 1: profiles/r1_exp_occupancy.txt).  k > 31: the same, with the 16-byte-tag front table and a queue of 1.5 chunks.
template <int W>
__global__ void __launch_bounds__(MCX_THREADS, MCX_CTAS(W == 1 ? MCX_FUSED_MINB : MCX_FUSED_MINB2))
mcx_build_fused_kernel(const __grid_constant__ McxBuildParams p, const __grid_constant__ McxTable t, const __grid_constant__ McxTupleBins nobins)
{
  McxSlowQueue<W> *q = mcx_queue<W>();
  if(threadIdx.x == 0) { q->n = 0; q->last_n = q->epoch = q->bypass = 0; }
  FusedSink<W, W == 1 ? MCX_FUSED_G : MCX_FUSED_G2, false> sink{t, nobins, p.colour, p.may_saturate != 0, q, p.counters, false};
  mcx_front_end<W, MCX_MODE_PLAIN>(p, sink);
}

// sharded build (one shard per GPU): the front table absorbs ALL hot k-mers locally whoever owns
// them (they are forwarded, aggregated, at flush); the parked pass inserts owned keys into the
// local big table and bins the others for the exchange
template <int W>
__global__ void __launch_bounds__(MCX_THREADS, MCX_CTAS(3))
mcx_build_sharded_kernel(const __grid_constant__ McxBuildParams p, const __grid_constant__ McxTable t, const __grid_constant__ McxTupleBins b)
{
  McxSlowQueue<W> *q = mcx_queue<W>();
  if(threadIdx.x == 0) { q->n = 0; q->last_n = q->epoch = q->bypass = 0; }
  FusedSink<W, 2, true> sink{t, b, p.colour, p.may_saturate != 0, q, p.counters, false};
  mcx_front_end<W, MCX_MODE_PLAIN>(p, sink);
}

// quality cut-off variants: pass 1 writes the per-chunk carry summaries, pass 2 inserts
struct NullSink {
  __device__ __forceinline__ void consume(const McxKmer<1> *, const uint32_t *, uint32_t, uint32_t &, uint32_t &) {}
  __device__ __forceinline__ void drain(uint32_t &, uint32_t &) {}
  __device__ __forceinline__ void reset() {}
  __device__ __forceinline__ bool should_drain() const { return false; }
  __device__ __forceinline__ void decide(uint32_t) {}
  __device__ __forceinline__ void begin_chunk() {}
};
__global__ void __launch_bounds__(MCX_THREADS, MCX_CTAS(5)) mcx_contig_summary_kernel(const __grid_constant__ McxBuildParams p)
{
  NullSink sink;
  mcx_front_end<1, MCX_MODE_QSUM>(p, sink);
}
template <int W>
__global__ void __launch_bounds__(MCX_THREADS, MCX_CTAS(MCX_QUAL_MINB))
mcx_build_fused_qual_kernel(const __grid_constant__ McxBuildParams p, const __grid_constant__ McxTable t, const __grid_constant__ McxTupleBins nobins)
{
  McxSlowQueue<W> *q = mcx_queue<W>();
  if(threadIdx.x == 0) { q->n = 0; q->last_n = q->epoch = q->bypass = 0; }
  FusedSink<W, 2, false> sink{t, nobins, p.colour, p.may_saturate != 0, q, p.counters, false};
  mcx_front_end<W, MCX_MODE_QUAL>(p, sink);
}

// the same with the sharded sink (multi-GPU builds with --fq-cutoff: what the production pipeline runs)
template <int W>
__global__ void __launch_bounds__(MCX_THREADS, MCX_CTAS(3))
mcx_build_sharded_qual_kernel(const __grid_constant__ McxBuildParams p, const __grid_constant__ McxTable t, const __grid_constant__ McxTupleBins b)
{
  McxSlowQueue<W> *q = mcx_queue<W>();
  if(threadIdx.x == 0) { q->n = 0; q->last_n = q->epoch = q->bypass = 0; }
  FusedSink<W, 2, true> sink{t, b, p.colour, p.may_saturate != 0, q, p.counters, false};
  mcx_front_end<W, MCX_MODE_QUAL>(p, sink);
}

template <int W>
__global__ void __launch_bounds__(MCX_THREADS, MCX_CTAS(4)) mcx_kmer_tuples_kernel(const __grid_constant__ McxBuildParams p, const __grid_constant__ McxTupleBins b)
{
  TupleSink<W> sink{b};
  mcx_front_end<W, MCX_MODE_PLAIN>(p, sink);
}

// ---------------------------------------------------------------- kernel C
// received tuples are (key words, meta = count << 8 | edge mask), already canonical and owned by
// this shard: Lookup3 + big-table find-or-insert + covg += count + edges |= mask, one per thread.
// (They are what the senders' front tables did NOT absorb, or absorbed and aggregated.)
template <int W>
__global__ void __launch_bounds__(MCX_THREADS, 4) mcx_insert_tuples_kernel(const uint64_t *__restrict__ keys,
                                                                           const uint32_t *__restrict__ meta,
                                                                           uint64_t n, const uint64_t *__restrict__ n_dev,
                                                                           McxTable t, uint32_t colour,
                                                                           int may_saturate, unsigned long long *counters)
{
  McxTable big = t; big.front = nullptr; big.front_set_bits = 0;
  uint64_t n_novel = 0, n_kmers = 0; uint32_t full = 0;
  if(n_dev) { const uint64_t nd = *n_dev; if(nd < n) n = nd; } // count written by the sender (exchanged on the stream)
  for(uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    McxKmer<W> key;
#pragma unroll
    for(int w = 0; w < W; w++) key.b[w] = keys[i * W + w];
    const uint32_t m = meta[i];
    if(m == 0u) continue; // nothing to add (no sender emits such a tuple)
    if(full) break;       // this thread has found the table full: the build has failed, stop probing
    uint32_t hb, hc = mcx_lookup3<W>(key, 0u, &hb);
    int r = mcx_table_add<W>(big, key, hc, hb, colour, m & 0xFFu, m >> 8, may_saturate != 0);
    n_novel += (r == 1); full |= (r == 2); n_kmers += m >> 8;
  }
  for(int s = 16; s > 0; s >>= 1) {
    n_novel += __shfl_xor_sync(0xFFFFFFFFu, n_novel, s);
    n_kmers += __shfl_xor_sync(0xFFFFFFFFu, n_kmers, s);
    full |= __shfl_xor_sync(0xFFFFFFFFu, full, s);
  }
  if((threadIdx.x & 31u) == 0) {
    if(n_novel) atomicAdd(&counters[MCX_CNT_NOVEL], (unsigned long long)n_novel);
    if(n_kmers) atomicAdd(&counters[MCX_CNT_INSERTED], (unsigned long long)n_kmers);
    if(full) atomicOr(&counters[MCX_CNT_FULL], 1ull);
  }
}

// ---------------------------------------------------------------- front-table flush
// merge every front entry that has counted something since the last flush into the big table:
// key = inverse hash of (set, tag), covg += count, edges |= edges; its counter restarts at 0
template <int W>
__global__ void __launch_bounds__(MCX_THREADS) mcx_front_flush_kernel(McxTable t, McxTupleBins bins, int may_saturate, unsigned long long *counters)
{
  const McxFrontGeom g = mcx_front_geom(t);
  const McxFrontGeom2 g2 = mcx_front_geom2_bits(t.front_set_bits);
  const uint64_t nways = (W == 1 ? 4ull : 2ull) << t.front_set_bits;
  uint64_t n_novel = 0; uint32_t full = 0;
  for(uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nways; i += (uint64_t)gridDim.x * blockDim.x) {
    McxKmer<W> key; uint32_t edges;
    if(W == 1) {
      const uint64_t v = t.front[i];
      if(v == 0) continue;
      const uint32_t hi = (uint32_t)(v >> 32);
      key.b[0] = mcx_fhash_inv((uint32_t)v, ((hi & (g.occ - 1u)) << g.S) | ((uint32_t)(i >> 2) ^ (hi >> 31))); // displaced: home = the neighbouring set
      edges = (hi >> g.eshift) & 0xFFu;
    } else {
      uint64_t lo, hi;
      mcx_ld128(t.front + 2u * i, lo, hi);
      if(!(hi & g2.occ)) continue;
      const uint64_t set = (i >> 1) ^ (hi >> 63);
      uint64_t kh, kl;
      mcx_fhash2_inv(lo, (set << g2.tshift) | (hi & (g2.occ - 1ull)), &kh, &kl);
      key.b[0] = kh; key.b[W - 1] = kl;
      edges = (uint32_t)(hi >> g2.eshift) & 0xFFu;
    }
    uint32_t count = t.front_cnt[i];
    // The tags stay: the table is still warm after the flush (no second wave of claims for the ~4.6 M hot
    // k-mers).  A record with no new occurrence has nothing new to merge -- its edge bits came with
    // occurrences that an earlier flush merged.
    if(count == 0) continue;
    t.front_cnt[i] = 0;
    uint32_t hb, hc = mcx_lookup3<W>(key, 0u, &hb);
    if(bins.nparts > 1u) {
      // sharded build: an aggregated record of a key owned elsewhere travels as ONE tuple (a tuple
      // carries a 24-bit count: more than that, which takes a k-mer seen > 16 M times, is split)
      const uint32_t d = mcx_owner(hc, bins.nparts);
      if(d != bins.my_part) {
        uint32_t e = edges;
        while(count | e) {
          const uint32_t c = count < 0xFFFFFFu ? count : 0xFFFFFFu;
          mcx_bin_push<W>(bins, d, key, (c << 8) | e, full);
          count -= c; e = 0;
        }
        continue;
      }
    }
    int r = mcx_table_add<W>(t, key, hc, hb, t.front_colour, edges, count, may_saturate != 0);
    n_novel += (r == 1); full |= (r == 2);
    if(full) break; // the build has failed: stop probing
  }
  for(int s = 16; s > 0; s >>= 1) {
    n_novel += __shfl_xor_sync(0xFFFFFFFFu, n_novel, s);
    full |= __shfl_xor_sync(0xFFFFFFFFu, full, s);
  }
  if((threadIdx.x & 31u) == 0) {
    if(n_novel) atomicAdd(&counters[MCX_CNT_NOVEL], (unsigned long long)n_novel);
    if(full) atomicOr(&counters[MCX_CNT_FULL], 1ull);
  }
}

// ---------------------------------------------------------------- repack
// OFFSETS layout (reads abut, offsets[n+1]) -> LINES layout (each read followed by '\n'):
// read r moves from [off[r], off[r+1]) to [off[r]+r, off[r+1]+r), terminator at off[r+1]+r.
__global__ void mcx_repack_lines_kernel(const uint8_t *__restrict__ src, const uint64_t *__restrict__ off, uint64_t nreads,
                                        uint8_t *__restrict__ dst)
{
  const uint32_t lane = threadIdx.x & 31u;
  uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for(uint64_t r = warp; r < nreads; r += nwarps) {
    uint64_t a = off[r], b = off[r + 1];
    for(uint64_t i = a + lane; i < b; i += 32) dst[i + r] = src[i];
    if(lane == 0) dst[b + r] = '\n';
  }
}

// ---------------------------------------------------------------- launchers
static int g_num_sms = 0;
static int num_sms()
{
  if(!g_num_sms) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if(g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// dynamic shared memory of the kernels that own a slow queue; raises the kernel's limit once
template <int W, class K> static size_t queue_smem(K kernel)
{
  const size_t bytes = sizeof(McxSlowQueue<W>);
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return bytes;
}

// CTAs of `kernel` that are resident per SM with `smem` bytes of dynamic shared memory.  The
// build kernels are persistent -- CTA b takes chunks b, b + grid, ... -- so a grid of more CTAs than fit runs a second,
// mostly empty wave (k > 31 was launched 4 per SM when 3 fit: a third of the run at a third of the occupancy).
template <class K> static int resident_ctas(K kernel, size_t smem)
{
  int n = 0;
  if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, (int)MCX_THREADS, smem) != cudaSuccess || n < 1) n = 1;
  return n;
}

static unsigned grid_for_chunks(const McxBuildParams &p, int ctas_per_sm)
{
  uint64_t nch = (p.r_end + MCX_T - 1) / MCX_T - p.r_begin / MCX_T;
  uint64_t g = (uint64_t)num_sms() * MCX_CTAS(ctas_per_sm);
  return (unsigned)(nch < g ? (nch ? nch : 1) : g);
}

cudaError_t mcx_launch_build_fused(const McxBuildParams &p, const McxTable &t, cudaStream_t st)
{
  if(p.r_end <= p.r_begin) return cudaSuccess;
  if(p.k <= 31) mcx_build_fused_kernel<1><<<grid_for_chunks(p, MCX_FUSED_MINB), MCX_THREADS, queue_smem<1>(mcx_build_fused_kernel<1>), st>>>(p, t, mcx_no_bins());
  else {
    const size_t smem = queue_smem<2>(mcx_build_fused_kernel<2>);
    mcx_build_fused_kernel<2><<<grid_for_chunks(p, resident_ctas(mcx_build_fused_kernel<2>, smem)), MCX_THREADS, smem, st>>>(p, t, mcx_no_bins());
  }
  return cudaGetLastError();
}

// quality cut-off: p.qual / p.qcut / p.summary set; the launch must start at a read boundary
cudaError_t mcx_launch_build_fused_qual(const McxBuildParams &p, const McxTable &t, cudaStream_t st)
{
  if(p.r_end <= p.r_begin) return cudaSuccess;
  mcx_contig_summary_kernel<<<grid_for_chunks(p, resident_ctas(mcx_contig_summary_kernel, 0)), MCX_THREADS, 0, st>>>(p);
  if(p.k <= 31) {
    const size_t smem = queue_smem<1>(mcx_build_fused_qual_kernel<1>);
    mcx_build_fused_qual_kernel<1><<<grid_for_chunks(p, resident_ctas(mcx_build_fused_qual_kernel<1>, smem)), MCX_THREADS, smem, st>>>(p, t, mcx_no_bins());
  } else {
    const size_t smem = queue_smem<2>(mcx_build_fused_qual_kernel<2>);
    mcx_build_fused_qual_kernel<2><<<grid_for_chunks(p, resident_ctas(mcx_build_fused_qual_kernel<2>, smem)), MCX_THREADS, smem, st>>>(p, t, mcx_no_bins());
  }
  return cudaGetLastError();
}

// pass 1 of the quality cut-off alone (the must-exist kernel of mcx_lookup.cu runs its own pass 2)
cudaError_t mcx_launch_contig_summary(const McxBuildParams &p, cudaStream_t st)
{
  if(p.r_end <= p.r_begin) return cudaSuccess;
  mcx_contig_summary_kernel<<<grid_for_chunks(p, resident_ctas(mcx_contig_summary_kernel, 0)), MCX_THREADS, 0, st>>>(p);
  return cudaGetLastError();
}

cudaError_t mcx_launch_kmer_tuples(const McxBuildParams &p, const McxTupleBins &b, cudaStream_t st)
{
  if(p.r_end <= p.r_begin) return cudaSuccess;
  unsigned grid = grid_for_chunks(p, 4);
  if(p.k <= 31) mcx_kmer_tuples_kernel<1><<<grid, MCX_THREADS, 0, st>>>(p, b);
  else mcx_kmer_tuples_kernel<2><<<grid, MCX_THREADS, 0, st>>>(p, b);
  return cudaGetLastError();
}

cudaError_t mcx_launch_insert_tuples(const uint64_t *keys, const uint32_t *meta, uint64_t n, const uint64_t *n_dev, uint32_t k,
                                     const McxTable &t, uint32_t colour, int may_saturate, unsigned long long *counters,
                                     cudaStream_t st)
{
  if(n == 0) return cudaSuccess;
  uint64_t want = (n + MCX_THREADS - 1) / MCX_THREADS, cap = (uint64_t)num_sms() * 8;
  unsigned grid = (unsigned)(want < cap ? want : cap);
  if(k <= 31) mcx_insert_tuples_kernel<1><<<grid, MCX_THREADS, 0, st>>>(keys, meta, n, n_dev, t, colour, may_saturate, counters);
  else mcx_insert_tuples_kernel<2><<<grid, MCX_THREADS, 0, st>>>(keys, meta, n, n_dev, t, colour, may_saturate, counters);
  return cudaGetLastError();
}

cudaError_t mcx_launch_repack_lines(const uint8_t *src, const uint64_t *off, uint64_t nreads, uint8_t *dst, cudaStream_t st)
{
  if(nreads == 0) return cudaSuccess;
  uint64_t want = (nreads * 32 + MCX_THREADS - 1) / MCX_THREADS, cap = (uint64_t)num_sms() * 16;
  unsigned grid = (unsigned)(want < cap ? want : cap);
  mcx_repack_lines_kernel<<<grid, MCX_THREADS, 0, st>>>(src, off, nreads, dst);
  return cudaGetLastError();
}

cudaError_t mcx_launch_front_flush(const McxTable &t, int may_saturate, unsigned long long *counters, cudaStream_t st)
{
  if(!t.front_set_bits) return cudaSuccess;
  if(t.front_words == 2u) mcx_front_flush_kernel<2><<<num_sms() * 8, MCX_THREADS, 0, st>>>(t, mcx_no_bins(), may_saturate, counters);
  else mcx_front_flush_kernel<1><<<num_sms() * 8, MCX_THREADS, 0, st>>>(t, mcx_no_bins(), may_saturate, counters);
  return cudaGetLastError(); // (the kernel zeroes the counters it merges; the tags stay claimed)
}

cudaError_t mcx_launch_front_flush_sharded(const McxTable &t, const McxTupleBins &b, int may_saturate,
                                           unsigned long long *counters, cudaStream_t st)
{
  if(!t.front_set_bits) return cudaSuccess;
  if(t.front_words == 2u) mcx_front_flush_kernel<2><<<num_sms() * 8, MCX_THREADS, 0, st>>>(t, b, may_saturate, counters);
  else mcx_front_flush_kernel<1><<<num_sms() * 8, MCX_THREADS, 0, st>>>(t, b, may_saturate, counters);
  return cudaGetLastError();
}

cudaError_t mcx_launch_build_sharded(const McxBuildParams &p, const McxTable &t, const McxTupleBins &b, cudaStream_t st)
{
  if(p.r_end <= p.r_begin) return cudaSuccess;
  unsigned grid = grid_for_chunks(p, 3);
  if(p.qual) { // quality cut-off: summary pass, then the sharded insert pass (the launch starts at a read boundary)
    mcx_contig_summary_kernel<<<grid_for_chunks(p, resident_ctas(mcx_contig_summary_kernel, 0)), MCX_THREADS, 0, st>>>(p);
    if(p.k <= 31) {
      const size_t smem = queue_smem<1>(mcx_build_sharded_qual_kernel<1>);
      mcx_build_sharded_qual_kernel<1><<<grid_for_chunks(p, resident_ctas(mcx_build_sharded_qual_kernel<1>, smem)), MCX_THREADS, smem, st>>>(p, t, b);
    } else {
      const size_t smem = queue_smem<2>(mcx_build_sharded_qual_kernel<2>);
      mcx_build_sharded_qual_kernel<2><<<grid_for_chunks(p, resident_ctas(mcx_build_sharded_qual_kernel<2>, smem)), MCX_THREADS, smem, st>>>(p, t, b);
    }
    return cudaGetLastError();
  }
  if(p.k <= 31) mcx_build_sharded_kernel<1><<<grid, MCX_THREADS, queue_smem<1>(mcx_build_sharded_kernel<1>), st>>>(p, t, b);
  else {
    const size_t smem = queue_smem<2>(mcx_build_sharded_kernel<2>);
    mcx_build_sharded_kernel<2><<<grid_for_chunks(p, resident_ctas(mcx_build_sharded_kernel<2>, smem)), MCX_THREADS, smem, st>>>(p, t, b);
  }
  return cudaGetLastError();
}

