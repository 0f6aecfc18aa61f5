// mcx_build_warp.cu -- kernel A2: the fused build with a WARP-AUTONOMOUS front end (k <= 31, no quality /
// homopolymer cut-off, inserting builds).  Same job as mcx_build_fused_kernel (mcx_build.cu):
//   reads -> 2-bit bases -> contigs -> rolling k-mers -> canonical key -> front table / Lookup3 + big table
// replacing build_graph_from_str_mt (src/tools/build_graph.c:122-150), seq_contig_start2/end2
// (src/basic/seq_reader.c:61-172), binary_kmer_* (src/graph/binary_kmer.{h,c}), bklk3_hashlittle
// (src/kmer/kmer_hash.h:162-211), hash_table_find_or_insert_mt (src/graph/hash_table.c:250-281),
// db_graph_update_node_mt / db_graph_add_edge_mt (src/graph/db_graph.c:101-166) of the reference.
//
// Why a second front end.  ncu of kernel A (profiles/r1e_*, r1i_*): 220-290 warp-instructions per round of
// 32 occurrences, one __syncthreads per 2048 windows with three thread roles per barrier interval, 8 windows
// per thread.  Here a warp owns a run of consecutive 512-position tiles and never talks to another warp:
//   * lane 0 keeps MCX_W_STAGES tiles in flight with 1-D TMA bulk copies (cp.async.bulk + mbarrier
//     complete_tx; SASS UBLKCP) into the warp's own ring; each lane reads its 16-byte piece with one
//     conflict-free LDS.128,
//   * a piece is converted ONCE (packed bases + bad bits), the two following pieces and the base before
//     come from the neighbouring lanes by shuffle,
//   * a lane walks 16 windows with rolling k-mers (set-up amortised over 16 instead of 8),
//   * occurrences the hot pass cannot finish are parked in the warp's own queue with a ballot (no shared
//     atomics: the queue length is a warp-uniform register) and drained by the warp itself.
// No CTA barrier in the loop, no thread roles.  The table side (front table hit = one 32-byte tag load + one
// 32-bit RED; parked pass = claim / Lookup3 + big table) is mcx_table.cuh, unchanged.
//
// Key classes (p.ncls_log2 > 0): the launch handles only keys of class p.cls (top bits of the bijective
// front hash), with a front table of its own per class; the host runs one launch per class over the same
// reads.  Every occurrence belongs to exactly one class, so the classes' counters add up.  This trades a
// second pass over the read stream (1.25 B per occurrence) for a front-table working set that fits the L2.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "mcx_lane.cuh"
#include "mcx_table.cuh"
#include "mcx_build.h"

#define MCX_W_THREADS 256u
#define MCX_W_WARPS (MCX_W_THREADS / 32u)
#define MCX_W_STAGES 4u      /* tiles in flight per warp */
#define MCX_WQ_CAP 768u      /* parked occurrences per warp: MCX_WQ_DRAIN + one tile's worth */
#define MCX_WQ_DRAIN 256u
#define MCX_FULLMASK 0xFFFFFFFFu

struct __align__(128) McxWarpSmem {
  uint8_t tile[MCX_W_STAGES][MCX_TILE + 16u + 16u];  // the piece before the tile, the tile (+ pad to 32 bytes)
  uint64_t qkey[MCX_WQ_CAP];
  uint8_t qmask[MCX_WQ_CAP];
  unsigned long long bar[MCX_W_STAGES];
};
extern __shared__ __align__(128) unsigned char mcx_wdyn[];

// ---------------------------------------------------------------- TMA / mbarrier (per warp)
__device__ __forceinline__ uint32_t w_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void w_mbar_init(unsigned long long *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(w_smem(bar)), "r"(count));
}
__device__ __forceinline__ void w_mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(w_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void w_mbar_wait(unsigned long long *bar, uint32_t parity)
{
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(w_smem(bar)), "r"(parity) : "memory");
  } while(!done);
}
__device__ __forceinline__ void w_tma_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(w_smem(dst)), "l"(src), "r"(bytes), "r"(w_smem(bar)) : "memory");
}

// ---------------------------------------------------------------- sink
template <int G, bool SHARDED> struct WarpSink {
  const McxTable &t; const McxTupleBins &bins;
  McxFrontGeom g; uint32_t colour; bool may_saturate;
  uint32_t cls, cls_shift;      // class of a key = fk.y >> cls_shift (cls_shift = 32: one class)
  uint64_t *qkey; uint8_t *qmask;
  uint32_t qn;                  // parked occurrences in the warp's queue (warp-uniform)
  uint32_t lane_lt;
  uint32_t novel, full, kmers;

  // all 32 lanes, convergent
  __device__ __forceinline__ void park(bool need, uint64_t key, uint32_t emask)
  {
    const uint32_t m = __ballot_sync(MCX_FULLMASK, need);
    if(m) {
      if(need) { const uint32_t at = qn + __popc(m & lane_lt); qkey[at] = key; qmask[at] = (uint8_t)emask; }
      qn += __popc(m);
    }
  }
  // all 32 lanes, convergent; valid = 4-bit mask of the group's windows this lane owns and that are in a contig
  __device__ __forceinline__ void consume(const McxKmer<1> keys[MCX_HALF], const uint32_t emasks[MCX_HALF], uint32_t valid)
  {
    if(t.front_set_bits) {
      McxFKey fk[MCX_HALF];
#pragma unroll
      for(uint32_t h = 0; h < MCX_HALF; h++) {
        fk[h] = mcx_fhash(keys[h].b[0]);
        if(cls_shift < 32u && (fk[h].y >> cls_shift) != cls) valid &= ~(1u << h);   // another launch's key
      }
      kmers += __popc(valid);
#pragma unroll
      for(uint32_t h = 0; h < MCX_HALF; h += G) {
        uint64_t v[G][4];
#pragma unroll
        for(uint32_t i = 0; i < G; i++)
          if((valid >> (h + i)) & 1u)
            mcx_ld256(t.front + ((uint64_t)(fk[h + i].y & g.setmask) << 2), v[i][0], v[i][1], v[i][2], v[i][3]);
#pragma unroll
        for(uint32_t i = 0; i < G; i++) {
          bool need = false;
          if((valid >> (h + i)) & 1u)
            need = !mcx_front_hit(g, t.front_cnt + ((uint64_t)(fk[h + i].y & g.setmask) << 2), fk[h + i].x,
                                  (fk[h + i].y >> g.S) | g.occ, emasks[h + i] << g.eshift, v[i][0], v[i][1], v[i][2], v[i][3]);
          park(need, keys[h + i].b[0], emasks[h + i]);
        }
      }
    } else {
      kmers += __popc(valid);
#pragma unroll
      for(uint32_t h = 0; h < MCX_HALF; h++) park((valid >> h) & 1u, keys[h].b[0], emasks[h]);
    }
  }
  // one parked occurrence: front table (claim / edge bit / displaced entry), else Lookup3 + big table (or its owner's bin)
  __device__ __forceinline__ void slow(uint64_t k0, uint32_t emask)
  {
    if(t.front_set_bits && mcx_front_add_slow(t, k0, emask)) return;
    McxKmer<1> key; key.b[0] = k0;
    uint32_t hb; const uint32_t hc = mcx_lookup3<1>(key, 0u, &hb);
    if(SHARDED) {
      const uint32_t d = mcx_owner(hc, bins.nparts);
      if(d != bins.my_part) { mcx_bin_push<1>(bins, d, key, (1u << 8) | emask, full); return; }
    }
    const int r = mcx_table_add<1>(t, key, hc, hb, colour, emask, 1u, may_saturate);
    novel += (r == 1); full |= (r == 2);
  }
  __device__ __forceinline__ void drain()
  {
    __syncwarp();
    for(uint32_t i = threadIdx.x & 31u; i < qn; i += 32u) slow(qkey[i], qmask[i]);
    __syncwarp();
    qn = 0;
  }
};

// ---------------------------------------------------------------- kernel
template <int G, bool SHARDED>
__global__ void __launch_bounds__(MCX_W_THREADS, 3)
mcx_build_warp_kernel(const __grid_constant__ McxBuildParams p, const __grid_constant__ McxTable t, const __grid_constant__ McxTupleBins bins)
{
  const uint32_t lane = threadIdx.x & 31u, wic = threadIdx.x >> 5;
  McxWarpSmem &sm = reinterpret_cast<McxWarpSmem *>(mcx_wdyn)[wic];
  const uint64_t T0 = p.r_begin / MCX_TILE, T1 = (p.r_end + MCX_TILE - 1u) / MCX_TILE;
  const uint64_t nw = (uint64_t)gridDim.x * MCX_W_WARPS, wid = (uint64_t)blockIdx.x * MCX_W_WARPS + wic;
  // A warp takes RUNS of R consecutive tiles, run q of warp w = tiles T0 + (w + q * nw) * R ...: at any time the grid
  // reads one compact window of the buffer (nw * R tiles), like kernel A's chunk round robin.  (One long run per
  // warp -- 3552 streams 2 MB apart -- cost a third of the speed: every 512-byte bulk copy opened a DRAM row of its own.)
  const uint32_t R = p.run_tiles ? p.run_tiles : 4u;
  const uint64_t stride = nw * R, first_run = T0 + wid * R;

  WarpSink<G, SHARDED> sink{t, bins, mcx_front_geom(t), p.colour, p.may_saturate != 0,
                            p.cls, p.ncls_log2 ? 30u - p.ncls_log2 : 32u, sm.qkey, sm.qmask, 0u, (1u << lane) - 1u, 0u, 0u, 0u};
  uint32_t n_contigs = 0, n_reads = 0;

  if(first_run < T1) {
    constexpr uint32_t nst = MCX_W_STAGES;
    const uint64_t npieces = (p.nbytes + 15u) >> 4;           // readable 16-byte pieces
    // what is loaded for a tile: the piece before it (the base before lane 0's first window) and its own 32 pieces,
    // as far as the buffer goes.  Stage layout: byte 0 = the piece before, byte 16 + 16 l = lane l's piece.
    auto tile_load = [&](uint64_t tile, uint64_t &first_piece) -> uint32_t {
      first_piece = tile ? tile * 32u - 1u : 0u;
      const uint64_t last = tile * 32u + 32u < npieces ? tile * 32u + 32u : npieces;
      return last > first_piece ? (uint32_t)(last - first_piece) * 16u : 0u;
    };
    // items of a run: its own tiles (at most R, fewer at the end of the launch) + the look-ahead tile
    auto run_items = [&](uint64_t run) -> uint32_t { return (uint32_t)(T1 - run < R ? T1 - run : R) + 1u; };
    // producer cursor (only lane 0's copy is used): the next item to issue
    uint64_t p_run = first_run; uint32_t p_i = 0, p_n = 0;
    auto issue_next = [&]() {
      if(p_run >= T1) return;
      const uint64_t tile = p_run + p_i;
      uint64_t first_piece;
      const uint32_t bytes = tile_load(tile, first_piece);
      if(bytes) {
        unsigned long long *bar = &sm.bar[p_n % nst];
        w_mbar_expect_tx(bar, bytes);
        w_tma_load(sm.tile[p_n % nst] + (tile ? 0u : 16u), p.seq + first_piece * 16u, bytes, bar);
      }
      p_n++;
      if(++p_i == run_items(p_run)) { p_i = 0; p_run += stride; }
    };
    if(lane == 0) {
#pragma unroll
      for(uint32_t s = 0; s < MCX_W_STAGES; s++) w_mbar_init(&sm.bar[s], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if(lane == 0) {
      for(uint32_t s = 0; s < nst; s++) issue_next();
    }
    // next item of the sequence = `tile`: this lane's converted piece (and, at a run start, the base before the
    // tile: bits 1:0 code, bit 2 bad); then the stage is refilled with the item STAGES further on
    uint32_t c_n = 0;
    auto fetch = [&](uint64_t tile, const bool want_carry, uint32_t &pk, uint32_t &bad, uint32_t &nl, uint32_t &carry) {
      uint64_t first_piece;
      const uint32_t bytes = tile_load(tile, first_piece);
      uint32_t w[4] = {0u, 0u, 0u, 0u}, wc[4] = {0u, 0u, 0u, 0u};
      if(bytes) {
        const uint8_t *stage = sm.tile[c_n % nst];
        w_mbar_wait(&sm.bar[c_n % nst], (c_n / nst) & 1u);
        if(tile * 32u + lane < npieces) {
          const uint4 v = *reinterpret_cast<const uint4 *>(stage + 16u + lane * 16u);
          w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        }
        if(want_carry && tile) {
          const uint4 v = *reinterpret_cast<const uint4 *>(stage);
          wc[0] = v.x; wc[1] = v.y; wc[2] = v.z; wc[3] = v.w;
        }
      }
      // The stage may be refilled only after EVERY lane's LDS has returned: a warp barrier orders the issue of
      // the loads, not their completion, and on a busy LSU queue the bulk copy of the item STAGES further on can
      // land first (seen on B200: a few windows per million read the wrong tile).  A vote on the loaded data
      // cannot execute before all 32 loads have written their registers.
      const uint32_t seen = __ballot_sync(MCX_FULLMASK, (w[0] ^ wc[0]) == 0x0A0A0A0Au);
      asm volatile("" ::"r"(seen) : "memory");
      c_n++;
      if(lane == 0) issue_next();
      mcx_piece_convert(w, (tile * 32u + lane) * 16u, p.nbytes, &pk, &bad, &nl);
      if(want_carry) {
        carry = 4u;
        if(tile) {
          uint32_t cpk, cbad, cnl;
          mcx_piece_convert(wc, (tile * 32u - 1u) * 16u, p.nbytes, &cpk, &cbad, &cnl);
          carry = (cpk & 3u) | ((cbad >> 15) << 2);
        }
      }
    };

    for(uint64_t run = first_run; run < T1; run += stride) {
      const uint64_t run_end = run + R < T1 ? run + R : T1;
      uint32_t cur_pk, cur_bad, cur_nl, carry, unused;
      fetch(run, true, cur_pk, cur_bad, cur_nl, carry);
      for(uint64_t tt = run; tt < run_end; tt++) {
        uint32_t nx_pk, nx_bad, nx_nl;
        fetch(tt + 1u, false, nx_pk, nx_bad, nx_nl, unused);
        // pieces +1 / +2: the next lanes of this tile, or the first lanes of the next one
        const uint32_t s1 = (lane + 1u) & 31u, s2 = (lane + 2u) & 31u;
        const uint32_t a1 = __shfl_sync(MCX_FULLMASK, cur_pk, s1), b1 = __shfl_sync(MCX_FULLMASK, nx_pk, s1);
        const uint32_t a2 = __shfl_sync(MCX_FULLMASK, cur_pk, s2), b2 = __shfl_sync(MCX_FULLMASK, nx_pk, s2);
        const uint32_t cb = cur_bad | (nx_bad << 16);
        const uint32_t x1 = __shfl_sync(MCX_FULLMASK, cb, s1), x2 = __shfl_sync(MCX_FULLMASK, cb, s2);
        const uint32_t pk1 = lane < 31u ? a1 : b1, pk2 = lane < 30u ? a2 : b2;
        const uint32_t bad1 = lane < 31u ? (x1 & 0xFFFFu) : (x1 >> 16), bad2 = lane < 30u ? (x2 & 0xFFFFu) : (x2 >> 16);
        // the base before: last base of the previous lane's piece
        const uint32_t pb = (cur_pk & 3u) | ((cur_bad >> 15) << 2);
        uint32_t up = __shfl_up_sync(MCX_FULLMASK, pb, 1);
        if(lane == 0) up = carry;
        carry = __shfl_sync(MCX_FULLMASK, pb, 31);

        const uint64_t gpos = (tt * 32u + lane) * 16u;
        const uint32_t own = (tt * MCX_TILE >= p.r_begin && (tt + 1u) * MCX_TILE <= p.r_end) ? 0xFFFFu : mcx_piece_own(gpos, p.r_begin, p.r_end);
        if(p.cls == 0u) n_reads += __popc(cur_nl & own);
        const uint64_t bad48 = (uint64_t)cur_bad | ((uint64_t)bad1 << 16) | ((uint64_t)bad2 << 32);
        const uint32_t vb = mcx_lane_valid(bad48, up >> 2, p.k);
        if(__any_sync(MCX_FULLMASK, ((vb >> 1) & own) != 0u)) {
          mcx_lane_windows(cur_pk, pk1, pk2, vb, up & 3u, p.k,
            [&](const McxKmer<1> *keys, const uint32_t *emasks, uint32_t valid, uint32_t starts, uint32_t j0) {
              valid &= own >> j0;
              n_contigs += __popc(starts & valid);
              sink.consume(keys, emasks, valid);
            });
          if(sink.qn > MCX_WQ_DRAIN) sink.drain();
        }
        cur_pk = nx_pk; cur_bad = nx_bad; cur_nl = nx_nl;
      }
    }
    sink.drain();
  }

  // ---- counters: warp shuffle -> one global atomic per warp and counter
  uint32_t n_kmers = sink.kmers, n_novel = sink.novel, full = sink.full;
  if(p.cls != 0u) n_contigs = 0;
  for(int sh = 16; sh > 0; sh >>= 1) {
    n_kmers += __shfl_xor_sync(MCX_FULLMASK, n_kmers, sh);
    n_novel += __shfl_xor_sync(MCX_FULLMASK, n_novel, sh);
    n_contigs += __shfl_xor_sync(MCX_FULLMASK, n_contigs, sh);
    n_reads += __shfl_xor_sync(MCX_FULLMASK, n_reads, sh);
    full |= __shfl_xor_sync(MCX_FULLMASK, full, sh);
  }
  if(lane == 0) {
    if(n_kmers) atomicAdd(&p.counters[MCX_CNT_KMERS], (unsigned long long)n_kmers);
    if(n_novel) atomicAdd(&p.counters[MCX_CNT_NOVEL], (unsigned long long)n_novel);
    if(n_contigs) atomicAdd(&p.counters[MCX_CNT_CONTIGS], (unsigned long long)n_contigs);
    if(n_reads) atomicAdd(&p.counters[MCX_CNT_READS], (unsigned long long)n_reads);
    if(full) atomicOr(&p.counters[MCX_CNT_FULL], 1ull);
  }
}

// ---------------------------------------------------------------- launchers
static int w_num_sms()
{
  static int n = 0;
  if(!n) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if(n <= 0) n = 148;
  }
  return n;
}
static unsigned w_grid(const McxBuildParams &p)
{
  const uint64_t ntiles = (p.r_end + MCX_TILE - 1u) / MCX_TILE - p.r_begin / MCX_TILE, R = p.run_tiles ? p.run_tiles : 4u;
  uint64_t want = ((ntiles + R - 1u) / R + MCX_W_WARPS - 1u) / MCX_W_WARPS, cap = (uint64_t)w_num_sms() * 3u;
  // MCX_W_GRID=<CTAs>: tests use it to give every warp a long run of tiles on a small input
  if(const char *m = getenv("MCX_W_GRID")) { const long forced = atol(m); if(forced > 0) cap = (uint64_t)forced; }
  return (unsigned)(want < cap ? (want ? want : 1u) : cap);
}
template <class K> static size_t w_smem_bytes(K kernel)
{
  const size_t bytes = sizeof(McxWarpSmem) * MCX_W_WARPS;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return bytes;
}
static McxTupleBins w_no_bins()
{
  McxTupleBins b;
  for(int i = 0; i < MCX_MAX_PARTS; i++) { b.keys[i] = nullptr; b.meta[i] = nullptr; }
  b.cursor = nullptr; b.cap = 0; b.nparts = 1; b.my_part = 0;
  return b;
}

bool mcx_warp_kernel_supports(const McxBuildParams &p)
{
  return p.k <= 31u && p.hp_cutoff == 0u && p.qual == nullptr;
}

cudaError_t mcx_launch_build_warp(const McxBuildParams &p, const McxTable &t, cudaStream_t st)
{
  if(p.r_end <= p.r_begin) return cudaSuccess;
  mcx_build_warp_kernel<2, false><<<w_grid(p), MCX_W_THREADS, w_smem_bytes(mcx_build_warp_kernel<2, false>), st>>>(p, t, w_no_bins());
  return cudaGetLastError();
}

cudaError_t mcx_launch_build_warp_sharded(const McxBuildParams &p, const McxTable &t, const McxTupleBins &b, cudaStream_t st)
{
  if(p.r_end <= p.r_begin) return cudaSuccess;
  mcx_build_warp_kernel<2, true><<<w_grid(p), MCX_W_THREADS, w_smem_bytes(mcx_build_warp_kernel<2, true>), st>>>(p, t, b);
  return cudaGetLastError();
}
