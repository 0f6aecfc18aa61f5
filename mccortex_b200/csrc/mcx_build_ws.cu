// mcx_build_ws.cu -- warp-specialised variant of the fused build kernel (k <= 31, front table present,
// no quality cut-off, single GPU).
//
// Why: in mcx_build_fused_kernel every warp alternates k-mer arithmetic (no load in flight) and table
// probing (no arithmetic); ablation (profiles/r1g_experiments.txt, item 5) puts the arithmetic at 15.7 ms
// and the probe at ~10 ms of a 45.7 ms launch -- they add up instead of overlapping, and a thread cannot
// hold more than two 32-byte probe loads next to the rolling k-mer state in 80 registers (item 16).
// Here a CTA is split in two roles that meet in shared memory:
//   producer warps 0-7  TMA stage, phase 1 (ASCII -> packed bases + masks), phase 2a (contig rules), phase 2b
//                       arithmetic (rolling k-mers, canonical key, edge mask, front-table hash) for the 2048
//                       windows of chunk j+1 -> ring[(j+1) & 1] = (x, y | valid, edge mask) per window
//   probe warps 8-9     read ring[j & 1] (chunk j), keep FOUR tag loads in flight per thread (they carry no
//                       k-mer state), RED the counter on a hit, park the rest
// One CTA-wide barrier per step hands the ring over and decides (OR-reduced) whether the parked queue is
// drained, by all 320 threads.  A named barrier (id 1) separates the producers' own phases.
// (First version: 4 + 4 warps decoupled by full / empty mbarriers -- correct, but with half the warps the
// arithmetic became the bound: 30 G k-mers/s.)  Same device functions as the fused kernel, same results.
#include <cuda_runtime.h>
#include <stdint.h>
#include "mcx_chunk.cuh"
#include "mcx_table.cuh"
#include "mcx_build.h"

#define WS_PROD 256u                      /* producer threads: warps 0-7, one 8-window unit per thread and chunk */
/* probe threads (PROBE: warps 8..) and tag loads in flight per probe thread (G) are template parameters */

struct __align__(128) McxWsSmem {
  uint8_t raw[2][MCX_RAW];
  uint32_t pk[MCX_PKW];
  uint32_t bad[MCX_MSW];
  uint32_t eq[MCX_MSW];
  uint32_t vmask[MCX_VW];
  uint32_t sx[2][MCX_T];                  // ring: entry (window j of producer thread vt) at j * 256 + vt
  uint32_t sy[2][MCX_T];                  //   y (30 bits) | valid << 31
  uint8_t sem[2][MCX_T];
  uint64_t qkey[2][MCX_T];                // parked occurrences of the chunk probed in step s: queue[s & 1]
  uint8_t qem[2][MCX_T];
  uint32_t qn[2];
  unsigned long long tma_bar[2];
  unsigned long long red[MCX_NCOUNTERS];
};

__device__ __forceinline__ uint32_t ws_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ws_mbar_init(unsigned long long *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ws_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ws_mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ws_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ws_mbar_arrive(unsigned long long *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ws_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ws_mbar_wait(unsigned long long *bar, uint32_t parity)
{
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(ws_smem_u32(bar)), "r"(parity) : "memory");
  } while(!done);
}
__device__ __forceinline__ void ws_tma_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(ws_smem_u32(dst)), "l"(src), "r"(bytes), "r"(ws_smem_u32(bar)) : "memory");
}
// named barrier over the producer threads
__device__ __forceinline__ void ws_prod_sync() { asm volatile("bar.sync 1, %0;" ::"r"(WS_PROD) : "memory"); }
__device__ __forceinline__ void ws_issue_chunk(McxWsSmem &sm, const McxBuildParams &p, uint64_t chunk, uint32_t buf)
{
  const uint64_t cs = chunk * (uint64_t)MCX_T;
  const uint64_t src_off = cs ? cs - MCX_LB : 0;
  const uint32_t dst_off = cs ? 0 : MCX_LB;
  const uint64_t avail = (p.nbytes - src_off + 15ull) & ~15ull;
  const uint32_t want = MCX_RAW - dst_off;
  const uint32_t bytes = avail < want ? (uint32_t)avail : want;
  ws_mbar_expect_tx(&sm.tma_bar[buf], bytes);
  ws_tma_load(&sm.raw[buf][dst_off], p.seq + src_off, bytes, &sm.tma_bar[buf]);
}

// parked occurrences, handled by the threads of the probe role (or by everybody at the end)
__device__ __forceinline__ void ws_drain(McxWsSmem &sm, uint32_t qi, const McxTable &t, uint32_t colour, bool may_saturate,
                                         uint32_t me, uint32_t nthreads, uint64_t &novel, uint32_t &full)
{
  const uint32_t n = sm.qn[qi];
  for(uint32_t i = me; i < n; i += nthreads) {
    McxKmer<1> key; key.b[0] = sm.qkey[qi][i];
    const uint32_t em = sm.qem[qi][i];
    if(mcx_front_add_slow(t, key.b[0], em)) continue;
    uint32_t hb, hc = mcx_lookup3<1>(key, 0u, &hb);
    const int r = mcx_table_add<1>(t, key, hc, hb, colour, em, 1u, may_saturate);
    novel += (r == 1);
    full |= (r == 2);
  }
}

template <uint32_t WS_PROBE, uint32_t WS_G, int MINB, uint32_t WS_DRAIN>
__global__ void __launch_bounds__(WS_PROD + WS_PROBE + WS_DRAIN, MINB) mcx_build_ws_kernel(McxBuildParams p, McxTable t)
{
  constexpr uint32_t WS_THREADS = WS_PROD + WS_PROBE + WS_DRAIN;
  extern __shared__ __align__(128) unsigned char ws_dyn[];
  McxWsSmem &sm = *reinterpret_cast<McxWsSmem *>(ws_dyn);
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  const bool producer = tid < WS_PROD;
  const uint64_t c_first = p.r_begin / MCX_T, c_last = (p.r_end + MCX_T - 1) / MCX_T;
  const uint64_t chunk0 = c_first + blockIdx.x, cstride = gridDim.x;
  const McxFrontGeom g = mcx_front_geom(t);

  if(tid == 0) {
    ws_mbar_init(&sm.tma_bar[0], 1); ws_mbar_init(&sm.tma_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    sm.qn[0] = sm.qn[1] = 0;
  }
  if(tid < MCX_NCOUNTERS) sm.red[tid] = 0;
  if(tid < 4) { sm.pk[MCX_RAW / 16u + tid] = 0; sm.bad[MCX_RAW / 32u + tid] = 0xFFFFFFFFu; sm.eq[MCX_RAW / 32u + tid] = 0; }
  __syncthreads();
  if(tid == 0) {
    if(chunk0 < c_last) ws_issue_chunk(sm, p, chunk0, 0);
    if(chunk0 + cstride < c_last) ws_issue_chunk(sm, p, chunk0 + cstride, 1);
  }

  uint64_t n_kmers = 0, n_novel = 0, n_contigs = 0, n_reads = 0;
  uint32_t full = 0;

  // step s: producers stage chunk s+1 into ring[(s+1) & 1], probe warps work through chunk s in ring[s & 1]
  for(int64_t s = -1;; s++) {
    const uint64_t ch_prod = chunk0 + (uint64_t)(s + 1) * cstride;
    const uint64_t ch_probe = chunk0 + (uint64_t)(s < 0 ? 0 : s) * cstride;
    if(s >= 0 && ch_probe >= c_last) break;
    if(producer) {
      if(ch_prod < c_last) {
        const uint32_t j = (uint32_t)(s + 1), b = j & 1u;
        const uint64_t cs = ch_prod * (uint64_t)MCX_T;
        ws_mbar_wait(&sm.tma_bar[b], (j >> 1) & 1u);
        // ---- phase 1
        if(tid < MCX_RAW / 16u) {
          const uint4 v = *reinterpret_cast<const uint4 *>(&sm.raw[b][tid * 16u]);
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
          const uint32_t prev = tid ? sm.raw[b][tid * 16u - 1u] : 0u;
          const uint64_t gpos = cs - MCX_LB + tid * 16ull;
          uint32_t pk, b16, e16, n16;
          mcx_convert16(w, prev, gpos, p.nbytes, &pk, &b16, &e16, &n16);
          sm.pk[tid] = pk;
          reinterpret_cast<uint16_t *>(sm.bad)[tid] = (uint16_t)b16;
          reinterpret_cast<uint16_t *>(sm.eq)[tid] = (uint16_t)e16;
          if(n16 && tid >= MCX_LB / 16u && tid < (MCX_LB + MCX_T) / 16u)
            for(uint32_t i = 0; i < 16u; i++)
              if(((n16 >> i) & 1u) && gpos + i >= p.r_begin && gpos + i < p.r_end) n_reads++;
        }
        ws_prod_sync();
        if(tid == 0 && ch_prod + 2u * cstride < c_last) ws_issue_chunk(sm, p, ch_prod + 2u * cstride, b);
        // ---- phase 2a
        if(tid >= WS_PROD - MCX_VW) {
          const uint32_t wi = tid - (WS_PROD - MCX_VW);
          const bool live = wi < (MCX_LB + MCX_T + 32u) / 32u;
          sm.vmask[wi] = live ? mcx_valid_word(sm.bad, sm.eq, wi, p.k, p.hp_cutoff) : 0u;
        }
        ws_prod_sync();
        // ---- phase 2b arithmetic: 8 windows per thread -> ring
        const uint64_t g0 = cs + MCX_WPT * tid;
        const uint32_t lo = p.r_begin > g0 ? (p.r_begin - g0 < MCX_WPT ? (uint32_t)(p.r_begin - g0) : MCX_WPT) : 0u;
        const uint32_t hi = p.r_end > g0 ? (p.r_end - g0 < MCX_WPT ? (uint32_t)(p.r_end - g0) : MCX_WPT) : 0u;
        const uint32_t own = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
        const bool any = mcx_thread_occurrences2<1>(sm.pk, sm.vmask, tid, p.k, [](uint32_t) {},
          [&](const McxKmer<1> *keys, const uint32_t *emasks, uint32_t valid, uint32_t starts, uint32_t j0) {
            valid &= own >> j0;
            n_kmers += __popc(valid);
            n_contigs += __popc(starts & valid);
#pragma unroll
            for(uint32_t i = 0; i < MCX_HALF; i++) {
              const McxFKey fk = mcx_fhash(keys[i].b[0]);
              const uint32_t e = (j0 + i) * WS_PROD + tid;
              sm.sx[b][e] = fk.x;
              sm.sy[b][e] = fk.y | (((valid >> i) & 1u) << 31);
              sm.sem[b][e] = (uint8_t)emasks[i];
            }
          });
        if(!any) {
#pragma unroll
          for(uint32_t jj = 0; jj < MCX_WPT; jj++) sm.sy[b][jj * WS_PROD + tid] = 0u;
        }
      }
    } else if(tid < WS_PROD + WS_PROBE) {
      if(s >= 0) {
      // ---- probe warps: chunk s, 32 windows per thread, WS_G tag loads in flight
      const uint32_t b = (uint32_t)s & 1u, pt = tid - WS_PROD;
#pragma unroll 1
      for(uint32_t e0 = pt; e0 < MCX_T; e0 += WS_PROBE * WS_G) {
        uint32_t x[WS_G], y[WS_G], em[WS_G]; uint64_t v[WS_G][4];
#pragma unroll
        for(uint32_t i = 0; i < WS_G; i++) {
          const uint32_t e = e0 + i * WS_PROBE;
          y[i] = 0u;
          if(e < MCX_T) { x[i] = sm.sx[b][e]; y[i] = sm.sy[b][e]; em[i] = sm.sem[b][e]; }
          if(y[i] >> 31) mcx_ld256_pol(t.front + ((uint64_t)(y[i] & g.setmask) << 2), t.pol_front, v[i][0], v[i][1], v[i][2], v[i][3]);
        }
#pragma unroll
        for(uint32_t i = 0; i < WS_G; i++) {
          if(y[i] >> 31) {
            const uint32_t yy = y[i] & 0x7FFFFFFFu;
            if(!mcx_front_hit(g, t.front_cnt + ((uint64_t)(yy & g.setmask) << 2), t.pol_cnt, x[i], (yy >> g.S) | g.occ,
                              em[i] << g.eshift, v[i][0], v[i][1], v[i][2], v[i][3])) {
              const uint32_t at = atomicAdd(&sm.qn[b], 1u);   // < MCX_T: one entry per window of the chunk
              sm.qkey[b][at] = mcx_fhash_inv(x[i], yy);
              sm.qem[b][at] = (uint8_t)em[i];
            }
          }
        }
      }
      }
    } else if(s >= 1) {
      // ---- drain warps: what the probe warps parked in the PREVIOUS step (front-table claim / edge bit /
      //      displaced k-mer, else Lookup3 + big table), while they probe this step's chunk
      const uint32_t qi = ((uint32_t)s - 1u) & 1u;
      ws_drain(sm, qi, t, p.colour, p.may_saturate != 0, tid - (WS_PROD + WS_PROBE), WS_DRAIN, n_novel, full);
      asm volatile("bar.sync 2, %0;" ::"r"(WS_DRAIN) : "memory");
      if(tid == WS_PROD + WS_PROBE) sm.qn[qi] = 0;
    }
    // ---- the step's barrier: hands the ring (producers -> probe warps) and the parked queue (probe -> drain warps) over
    __syncthreads();
    if(WS_DRAIN == 0u) {
      // no drain role: everybody drains what this step parked
      const uint32_t qi = (uint32_t)(s < 0 ? 0 : s) & 1u;
      ws_drain(sm, qi, t, p.colour, p.may_saturate != 0, tid, WS_THREADS, n_novel, full);
      __syncthreads();
      if(tid == 0) sm.qn[qi] = 0;
      __syncthreads();
    }
  }
  __syncthreads();
  ws_drain(sm, 0, t, p.colour, p.may_saturate != 0, tid, WS_THREADS, n_novel, full);
  ws_drain(sm, 1, t, p.colour, p.may_saturate != 0, tid, WS_THREADS, n_novel, full);

  // ---- counters
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

static int g_ws_variant = 1;
void mcx_set_ws_variant(int v) { g_ws_variant = v; }

template <uint32_t PROBE, uint32_t G, int MINB, uint32_t DRAIN>
static cudaError_t ws_launch(const McxBuildParams &p, const McxTable &t, cudaStream_t st, int sms)
{
  const uint64_t nch = (p.r_end + MCX_T - 1) / MCX_T - p.r_begin / MCX_T, cap = (uint64_t)sms * MINB;
  const unsigned grid = (unsigned)(nch < cap ? (nch ? nch : 1) : cap);
  const size_t smem = sizeof(McxWsSmem);
  cudaFuncSetAttribute(mcx_build_ws_kernel<PROBE, G, MINB, DRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mcx_build_ws_kernel<PROBE, G, MINB, DRAIN><<<grid, WS_PROD + PROBE + DRAIN, smem, st>>>(p, t);
  return cudaGetLastError();
}

cudaError_t mcx_launch_build_ws(const McxBuildParams &p, const McxTable &t, cudaStream_t st)
{
  if(p.r_end <= p.r_begin) return cudaSuccess;
  int dev = 0, sms = 148; cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  switch(g_ws_variant) {
    case 2: return ws_launch<128u, 4u, 2, 0u>(p, t, st, sms);     // 8 + 4 warps, everybody drains every step
    case 3: return ws_launch<128u, 4u, 2, 128u>(p, t, st, sms);   // 8 + 4 + 4 warps
    case 4: return ws_launch<128u, 4u, 2, 64u>(p, t, st, sms);    // 8 + 4 + 2 warps
    case 5: return ws_launch<192u, 4u, 1, 256u>(p, t, st, sms);   // one CTA per SM: 8 + 6 + 8 warps
    case 6: return ws_launch<64u, 4u, 3, 64u>(p, t, st, sms);     // 8 + 2 + 2 warps, 3 CTAs per SM
    default: return ws_launch<128u, 4u, 2, 128u>(p, t, st, sms);
  }
}
