// part_bench.cu -- the two make-or-break kernels of a radix-partitioned (two-pass) build, in isolation and
// self-checking.  NOT part of the product; it answers, on a B200, the two questions the estimate
// in DESIGN.md leaves open before anyone writes the real thing:
//   A  scatter   : how fast can a CTA that already holds canonical k-mers in registers (what the front end delivers,
//                  153 G k-mers/s on its own) stage 8-byte tuples in shared-memory bins and flush them, sector-aligned,
//                  into 512 partition buffers in HBM?                                  (target: >= 150 G tuples/s)
//   B  aggregate : how fast can one CTA per partition stream such a buffer through a shared-memory hash table
//                  (16 K entries of 8-byte tag + edges, 4-byte counter), with the ~3.5 % of tuples that find no
//                  room (error k-mers) leaving through a block-reserved spill list?    (target: >= 250 G tuples/s)
// The tuple is what the real kernels would use: mcx_fhash (the bijection of the front table) of the key, minus the
// 9 partition bits, plus the 8-bit edge mask: [ valid:1 | - | edges:8 | y>>9 : 21 | x : 32 ].
// Workload: 96.5 % of the occurrences draw from 4.6 M hot keys, the rest are unique (the bench workload's mix).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o part_bench scripts/part_bench.cu && ./part_bench [Mtuples]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../mccortex_b200/csrc/mcx_device.cuh"

#define PB 9u
#define P (1u << PB)
#define BIN 12u              /* tuples per shared-memory bin; flushed in multiples of 4 (= 32-byte sectors) */
#define THREADS_A 256u
#define WPT 8u               /* occurrences per thread per chunk, as in the front end */
#define NE (1u << 14)        /* entries of the aggregation table */
#define THREADS_B 1024u
#define ROUND_B (8u * THREADS_B)
#define HOT 4600000ull

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
  z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
// occurrence i of the synthetic stream -> (62-bit key, edge mask)
__device__ __forceinline__ void occurrence(uint64_t i, uint64_t &key, uint32_t &emask)
{
  const uint64_t r = mix64(i);
  const bool hot = (uint32_t)(r >> 40) % 1000u < 965u;
  key = (hot ? mix64((r % HOT) * 0x51ull + 7ull) : mix64(i ^ 0xC01Dull << 20)) & ((1ull << 62) - 1ull);
  emask = 1u << ((uint32_t)(r >> 8) & 7u);
}
__device__ __forceinline__ uint64_t make_tuple(uint64_t key, uint32_t emask, uint32_t &p)
{
  const McxFKey fk = mcx_fhash(key);
  p = fk.y & (P - 1u);
  return (uint64_t)fk.x | ((uint64_t)(fk.y >> PB) << 32) | ((uint64_t)emask << 53) | (1ull << 63);
}

// ------------------------------------------------------------------------------------------ A: scatter
struct SmemA { unsigned long long bins[P * BIN]; unsigned int cnt[P]; };
__device__ __forceinline__ void flush_bin(SmemA &sm, uint32_t p, bool all, unsigned long long *part, unsigned long long *gcur, uint64_t cap,
                                          unsigned long long *lost)
{
  const uint32_t n = min(sm.cnt[p], BIN), m = all ? n : (n & ~3u);
  if(m) {
    const unsigned long long base = atomicAdd(&gcur[p], (unsigned long long)m);
    for(uint32_t j = 0; j < m; j++) {
      if(base + j < cap) part[(uint64_t)p * cap + base + j] = sm.bins[p * BIN + j];
      else atomicAdd(lost, 1ull);
    }
    for(uint32_t j = 0; j < n - m; j++) sm.bins[p * BIN + j] = sm.bins[p * BIN + m + j];
  }
  sm.cnt[p] = n - m;
}
__global__ void __launch_bounds__(THREADS_A, 3) scatter_kernel(uint64_t n_occ, unsigned long long *part, unsigned long long *gcur, uint64_t cap,
                                                               unsigned long long *lost, unsigned long long *checksum)
{
  extern __shared__ __align__(16) unsigned char dyn[];
  SmemA &sm = *reinterpret_cast<SmemA *>(dyn);
  for(uint32_t p = threadIdx.x; p < P; p += THREADS_A) sm.cnt[p] = 0;
  __syncthreads();
  const uint64_t chunk = (uint64_t)THREADS_A * WPT, nchunks = (n_occ + chunk - 1) / chunk;
  unsigned long long sum = 0;
  for(uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
#pragma unroll
    for(uint32_t j = 0; j < WPT; j++) {
      const uint64_t i = c * chunk + (uint64_t)threadIdx.x * WPT + j;
      if(i >= n_occ) break;
      uint64_t key; uint32_t em, p;
      occurrence(i, key, em);
      const uint64_t t = make_tuple(key, em, p);
      sum += t;
      const uint32_t slot = atomicAdd(&sm.cnt[p], 1u);
      if(slot < BIN) sm.bins[p * BIN + slot] = t;
      else { // the bin is full until the next flush: straight to HBM
        const unsigned long long pos = atomicAdd(&gcur[p], 1ull);
        if(pos < cap) part[(uint64_t)p * cap + pos] = t; else atomicAdd(lost, 1ull);
      }
    }
    __syncthreads();
    for(uint32_t p = threadIdx.x; p < P; p += THREADS_A) if(sm.cnt[p] >= 4u) flush_bin(sm, p, false, part, gcur, cap, lost);
    __syncthreads();
  }
  for(uint32_t p = threadIdx.x; p < P; p += THREADS_A) flush_bin(sm, p, true, part, gcur, cap, lost);
  for(int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, s);
  if((threadIdx.x & 31u) == 0) atomicAdd(checksum, sum);
}
// what the stream looks like without any scatter (generation + hash only): the part of A that is not the scatter
__global__ void __launch_bounds__(THREADS_A, 3) generate_only_kernel(uint64_t n_occ, unsigned long long *checksum)
{
  const uint64_t chunk = (uint64_t)THREADS_A * WPT, nchunks = (n_occ + chunk - 1) / chunk;
  unsigned long long sum = 0;
  for(uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x)
#pragma unroll
    for(uint32_t j = 0; j < WPT; j++) {
      const uint64_t i = c * chunk + (uint64_t)threadIdx.x * WPT + j;
      if(i >= n_occ) break;
      uint64_t key; uint32_t em, p;
      occurrence(i, key, em);
      sum += make_tuple(key, em, p);
    }
  for(int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, s);
  if((threadIdx.x & 31u) == 0) atomicAdd(checksum, sum);
}

// ------------------------------------------------------------------------------------------ B: aggregate
#define TAGMASK ((1ull << 53) - 1ull)
#define OCC (1ull << 63)
#define QCAP 2048u            /* spill staging per round (a round expects ~3.5 % of 8192 tuples); beyond it: one atomic per tuple */
struct SmemB { unsigned long long tab[NE]; unsigned int cnt[NE]; unsigned long long queue[QCAP]; unsigned int qn, qdirect; unsigned long long qbase; };
__global__ void __launch_bounds__(THREADS_B, 1) aggregate_kernel(const unsigned long long *part, const unsigned long long *gcur, uint64_t cap,
                                                                 unsigned long long *spill, unsigned long long *spill_cur, uint64_t spill_cap,
                                                                 unsigned long long *totals /* [0] counted, [1] distinct entries, [2] spilled */)
{
  extern __shared__ __align__(16) unsigned char dyn[];
  SmemB &sm = *reinterpret_cast<SmemB *>(dyn);
  unsigned long long counted = 0, distinct = 0, spilled = 0;
  for(uint32_t p = blockIdx.x; p < P; p += gridDim.x) {
    for(uint32_t e = threadIdx.x; e < NE; e += THREADS_B) { sm.tab[e] = 0; sm.cnt[e] = 0; }
    if(threadIdx.x == 0) { sm.qn = 0; sm.qdirect = 0; }
    __syncthreads();
    const uint64_t n = min((uint64_t)gcur[p], cap);
    const unsigned long long *src = part + (uint64_t)p * cap;
    for(uint64_t r0 = 0; r0 < n; r0 += ROUND_B) {
#pragma unroll 2
      for(uint32_t j = 0; j < ROUND_B / THREADS_B; j++) {
        const uint64_t i = r0 + (uint64_t)j * THREADS_B + threadIdx.x;
        if(i >= n) break;
        const unsigned long long t = src[i];
        const unsigned long long tag = t & TAGMASK, eb = t & (0xFFull << 53);
        uint32_t e = ((uint32_t)t * 0x9E3779B1u) >> (32u - 14u);
        bool done = false;
        for(uint32_t probe = 0; probe < 8u && !done; probe++, e = (e + 1u) & (NE - 1u)) {
          unsigned long long cur = sm.tab[e];
          if(cur == 0) {
            const unsigned long long old = atomicCAS(&sm.tab[e], 0ull, tag | eb | OCC);
            cur = old ? old : (tag | eb | OCC);
          }
          if(((cur ^ tag) & TAGMASK) == 0) {
            if((cur & eb) != eb) atomicOr(&sm.tab[e], eb);
            atomicAdd(&sm.cnt[e], 1u);
            done = true;
          }
        }
        if(!done) { // no room within 8 probes (a cold k-mer, mostly): leaves as it is
          const uint32_t q = atomicAdd(&sm.qn, 1u);
          if(q < QCAP) sm.queue[q] = t;
          else { const unsigned long long at = atomicAdd(spill_cur, 1ull); if(at < spill_cap) spill[at] = t; atomicAdd(&sm.qdirect, 1u); }
        }
      }
      __syncthreads();
      const uint32_t qn = min(sm.qn, QCAP);
      if(threadIdx.x == 0) { spilled += sm.qdirect; }
      if(qn) { // one reservation per round, not one atomic per tuple
        if(threadIdx.x == 0) sm.qbase = atomicAdd(spill_cur, (unsigned long long)qn);
        __syncthreads();
        for(uint32_t q = threadIdx.x; q < qn; q += THREADS_B) if(sm.qbase + q < spill_cap) spill[sm.qbase + q] = sm.queue[q];
        spilled += (threadIdx.x == 0) ? qn : 0;
      }
      __syncthreads();
      if(threadIdx.x == 0) { sm.qn = 0; sm.qdirect = 0; }
      __syncthreads();
    }
    __syncthreads();
    // (the real kernel would merge every entry into the partition's slice of the big table here)
    for(uint32_t e = threadIdx.x; e < NE; e += THREADS_B) if(sm.tab[e]) { counted += sm.cnt[e]; distinct++; }
    __syncthreads();
  }
  for(int s = 16; s > 0; s >>= 1) {
    counted += __shfl_xor_sync(0xFFFFFFFFu, counted, s); distinct += __shfl_xor_sync(0xFFFFFFFFu, distinct, s); spilled += __shfl_xor_sync(0xFFFFFFFFu, spilled, s);
  }
  if((threadIdx.x & 31u) == 0) { atomicAdd(&totals[0], counted); atomicAdd(&totals[1], distinct); atomicAdd(&totals[2], spilled); }
}

#define CK(x) do { cudaError_t e_ = (x); if(e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while(0)
int main(int argc, char **argv)
{
  const uint64_t n_occ = (argc > 1 ? strtoull(argv[1], NULL, 10) : 1200ull) * 1000000ull;
  const uint64_t cap = (n_occ / P) * 5 / 4 + 4096;
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long *part, *gcur, *scal, *spill;
  const uint64_t spill_cap = n_occ / 8;
  CK(cudaMalloc(&part, (size_t)P * cap * 8)); CK(cudaMalloc(&gcur, P * 8)); CK(cudaMalloc(&scal, 8 * 8)); CK(cudaMalloc(&spill, spill_cap * 8));
  CK(cudaFuncSetAttribute(scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemA)));
  CK(cudaFuncSetAttribute(aggregate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemB)));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms_gen = 0, ms_a = 0, ms_b = 0;
  unsigned long long h[8];
  for(int rep = 0; rep < 2; rep++) { // first repetition = warm-up
    CK(cudaMemset(scal, 0, 64)); CK(cudaMemset(gcur, 0, P * 8));
    cudaEventRecord(e0);
    generate_only_kernel<<<sms * 3, THREADS_A>>>(n_occ, scal + 1);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms_gen, e0, e1);
    cudaEventRecord(e0);
    scatter_kernel<<<sms * 3, THREADS_A, sizeof(SmemA)>>>(n_occ, part, gcur, cap, scal + 0, scal + 2);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms_a, e0, e1);
    cudaEventRecord(e0);
    aggregate_kernel<<<sms, THREADS_B, sizeof(SmemB)>>>(part, gcur, cap, spill, scal + 3, spill_cap, scal + 4);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms_b, e0, e1);
    CK(cudaGetLastError());
  }
  CK(cudaMemcpy(h, scal, 64, cudaMemcpyDeviceToHost));
  unsigned long long *hc = (unsigned long long *)malloc(P * 8), total = 0, mx = 0;
  CK(cudaMemcpy(hc, gcur, P * 8, cudaMemcpyDeviceToHost));
  for(uint32_t p = 0; p < P; p++) { total += hc[p]; if(hc[p] > mx) mx = hc[p]; }
  printf("tuples %llu, partitions %u (largest %llu, capacity %llu), lost %llu, checksum %s\n", (unsigned long long)n_occ, P, mx,
         (unsigned long long)cap, h[0], h[1] == h[2] ? "ok" : "MISMATCH");
  printf("scatter self-check   : %s (sum of cursors %llu)\n", total == n_occ && h[0] == 0 ? "ok" : "FAILED", total);
  printf("aggregate self-check : %s (counted %llu + spilled %llu; %llu table entries, %.2f %% spilled)\n",
         h[4] + h[6] == n_occ && h[6] == h[3] ? "ok" : "FAILED", h[4], h[6], h[5], 100.0 * (double)h[6] / (double)n_occ);
  printf("generate + hash only : %7.2f ms  %7.1f G tuples/s\n", ms_gen, n_occ / ms_gen / 1e6);
  printf("A scatter (incl. gen): %7.2f ms  %7.1f G tuples/s   (%.1f GB written, %.0f GB/s)\n", ms_a, n_occ / ms_a / 1e6, n_occ * 8 / 1e9, n_occ * 8 / ms_a / 1e6);
  printf("B aggregate          : %7.2f ms  %7.1f G tuples/s   (%.0f GB/s read)\n", ms_b, n_occ / ms_b / 1e6, n_occ * 8 / ms_b / 1e6);
  printf("A - gen + B per 2.4 G occurrences: %.1f ms (the fused kernel spends ~30 ms of its 45.7 ms outside the front end)\n",
         (ms_a - ms_gen + ms_b) * 2.4e9 / (double)n_occ);
  return 0;
}
