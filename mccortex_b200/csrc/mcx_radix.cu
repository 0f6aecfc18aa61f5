// mcx_radix.cu -- stable LSD radix sort of (u64 key, u64 value) pairs, 8 bits per pass (see mcx_radix.cuh).
#include "mcx_radix.cuh"

// ---- 1. histogram: hist[digit * nblk + block] = pairs of the block's tile with that digit
__global__ void __launch_bounds__(MCX_RX_THREADS) mcx_rx_hist_kernel(const uint64_t *__restrict__ keys, uint64_t n, uint32_t shift, uint64_t nblk,
                                                                     unsigned long long *__restrict__ hist)
{
  __shared__ unsigned int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = blockIdx.x * (uint64_t)MCX_RX_TILE;
#pragma unroll
  for(uint32_t i = 0; i < MCX_RX_ITEMS; i++) {
    const uint64_t at = base + i * MCX_RX_THREADS + threadIdx.x;
    if(at < n) atomicAdd(&h[(uint32_t)(keys[at] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(uint64_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// ---- 2. exclusive scan of the flat hist array (digit-major): chunks of 4096 entries, then the chunk totals
__global__ void __launch_bounds__(MCX_RX_THREADS) mcx_rx_scan_local_kernel(unsigned long long *__restrict__ a, uint64_t n, unsigned long long *__restrict__ totals)
{
  __shared__ unsigned long long wsum[MCX_RX_WARPS];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint64_t base = blockIdx.x * (uint64_t)MCX_RX_SCAN_CHUNK + (uint64_t)threadIdx.x * (MCX_RX_SCAN_CHUNK / MCX_RX_THREADS);
  unsigned long long v[MCX_RX_SCAN_CHUNK / MCX_RX_THREADS], sum = 0;
#pragma unroll
  for(uint32_t i = 0; i < MCX_RX_SCAN_CHUNK / MCX_RX_THREADS; i++) { v[i] = base + i < n ? a[base + i] : 0ull; sum += v[i]; }
  unsigned long long incl = sum;   // inclusive scan of the thread sums inside the warp
  for(uint32_t s = 1; s < 32u; s <<= 1) { const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, s); if(lane >= s) incl += t; }
  if(lane == 31u) wsum[warp] = incl;
  __syncthreads();
  unsigned long long off = incl - sum;
  for(uint32_t w = 0; w < warp; w++) off += wsum[w];
#pragma unroll
  for(uint32_t i = 0; i < MCX_RX_SCAN_CHUNK / MCX_RX_THREADS; i++) { if(base + i < n) a[base + i] = off; off += v[i]; }
  if(threadIdx.x == MCX_RX_THREADS - 1u) totals[blockIdx.x] = off;
}
__global__ void __launch_bounds__(MCX_RX_THREADS) mcx_rx_scan_totals_kernel(unsigned long long *__restrict__ totals, uint64_t nchunks)
{
  __shared__ unsigned long long wsum[MCX_RX_WARPS];
  __shared__ unsigned long long carry_s;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if(threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for(uint64_t b = 0; b < nchunks; b += MCX_RX_THREADS) {
    const uint64_t i = b + threadIdx.x;
    const unsigned long long v = i < nchunks ? totals[i] : 0ull;
    unsigned long long incl = v;
    for(uint32_t s = 1; s < 32u; s <<= 1) { const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, s); if(lane >= s) incl += t; }
    if(lane == 31u) wsum[warp] = incl;
    __syncthreads();
    unsigned long long off = carry_s + incl - v;
    for(uint32_t w = 0; w < warp; w++) off += wsum[w];
    if(i < nchunks) totals[i] = off;
    __syncthreads();
    if(threadIdx.x == MCX_RX_THREADS - 1u) carry_s = off + v;
    __syncthreads();
  }
}

// ---- 3. stable scatter
__global__ void __launch_bounds__(MCX_RX_THREADS) mcx_rx_scatter_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ vals, uint64_t n,
                                                                        uint32_t shift, uint64_t nblk, const unsigned long long *__restrict__ hist,
                                                                        const unsigned long long *__restrict__ totals,
                                                                        uint64_t *__restrict__ okeys, uint64_t *__restrict__ ovals)
{
  __shared__ unsigned int wcnt[MCX_RX_WARPS][256];   // per warp: pairs of each digit seen so far, then the warp's offset in the block's run
  __shared__ unsigned long long gbase[256];           // where the block's run of each digit starts in the output
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, lt = (1u << lane) - 1u;
  for(uint32_t i = threadIdx.x; i < MCX_RX_WARPS * 256u; i += MCX_RX_THREADS) (&wcnt[0][0])[i] = 0u;
  { const uint64_t e = (uint64_t)threadIdx.x * nblk + blockIdx.x; gbase[threadIdx.x] = hist[e] + totals[e / MCX_RX_SCAN_CHUNK]; }
  __syncthreads();
  // tile order: warp w owns pairs [512 w, 512 w + 512) of the tile; round r = 32 consecutive pairs, one per lane
  const uint64_t wbase = blockIdx.x * (uint64_t)MCX_RX_TILE + warp * (MCX_RX_ITEMS * 32u);
  uint64_t k[MCX_RX_ITEMS]; uint32_t rank[MCX_RX_ITEMS];
#pragma unroll
  for(uint32_t r = 0; r < MCX_RX_ITEMS; r++) {
    const uint64_t at = wbase + r * 32u + lane;
    const bool valid = at < n;
    k[r] = valid ? keys[at] : 0ull;
    const uint32_t d = (uint32_t)(k[r] >> shift) & 255u;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, valid ? d : (256u | lane));   // a lane past the end matches nobody
    const uint32_t seen = wcnt[warp][d];
    rank[r] = seen + __popc(peers & lt);
    __syncwarp();
    if(valid && (peers & lt) == 0u) wcnt[warp][d] = seen + __popc(peers);   // the group's first lane
    __syncwarp();
  }
  __syncthreads();
  { // exclusive prefix over the warps, per digit (thread = digit)
    unsigned int run = 0;
#pragma unroll
    for(uint32_t w = 0; w < MCX_RX_WARPS; w++) { const unsigned int c = wcnt[w][threadIdx.x]; wcnt[w][threadIdx.x] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for(uint32_t r = 0; r < MCX_RX_ITEMS; r++) {
    const uint64_t at = wbase + r * 32u + lane;
    if(at < n) {
      const uint32_t d = (uint32_t)(k[r] >> shift) & 255u;
      const uint64_t pos = gbase[d] + wcnt[warp][d] + rank[r];
      okeys[pos] = k[r];
      ovals[pos] = vals[at];
    }
  }
}

cudaError_t mcx_radix_sort_pairs(uint64_t *keys, uint64_t *vals, uint64_t *keys_alt, uint64_t *vals_alt, uint64_t n, int end_bit,
                                 void *scratch, uint64_t **out_keys, uint64_t **out_vals, cudaStream_t st)
{
  *out_keys = keys; *out_vals = vals;
  if(n == 0 || end_bit <= 0) return cudaSuccess;
  const uint64_t nblk = (n + MCX_RX_TILE - 1) / MCX_RX_TILE, nh = 256u * nblk, nchunks = (nh + MCX_RX_SCAN_CHUNK - 1) / MCX_RX_SCAN_CHUNK;
  if(nblk > 0x7FFFFFFFull) return cudaErrorInvalidValue;
  unsigned long long *hist = (unsigned long long *)scratch, *totals = hist + nh;
  for(int shift = 0; shift < end_bit; shift += 8) {
    mcx_rx_hist_kernel<<<(unsigned)nblk, MCX_RX_THREADS, 0, st>>>(keys, n, (uint32_t)shift, nblk, hist);
    mcx_rx_scan_local_kernel<<<(unsigned)nchunks, MCX_RX_THREADS, 0, st>>>(hist, nh, totals);
    mcx_rx_scan_totals_kernel<<<1, MCX_RX_THREADS, 0, st>>>(totals, nchunks);
    mcx_rx_scatter_kernel<<<(unsigned)nblk, MCX_RX_THREADS, 0, st>>>(keys, vals, n, (uint32_t)shift, nblk, hist, totals, keys_alt, vals_alt);
    cudaError_t e = cudaGetLastError();
    if(e != cudaSuccess) return e;
    uint64_t *t = keys; keys = keys_alt; keys_alt = t;
    t = vals; vals = vals_alt; vals_alt = t;
  }
  *out_keys = keys; *out_vals = vals;
  return cudaSuccess;
}
