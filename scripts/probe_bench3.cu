// microbenchmark 3: split layout -- tags in a read-mostly region, counters in a write-only region
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
// MODE 0: ld32B(tags) + RED32(counters[set*4+way])   1: ld32B(tags) + RED32 on packed 16-bit counters
// MODE 2: ATOM64 only on tags (returns the slot)      3: ld16B(tags: 2-way of 8B) + RED32(counters)
// MODE 4: ld32B(tags) only                            5: RED32(counters) only
template <int MODE, int ILP>
__global__ void __launch_bounds__(256) k(unsigned long long *tags, unsigned int *cnt, uint32_t setmask, uint32_t iters, unsigned long long *sink)
{
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
  for(uint32_t it = 0; it < iters; it += ILP) {
    uint32_t s[ILP]; uint64_t v[ILP][4];
#pragma unroll
    for(int i = 0; i < ILP; i++) {
      s[i] = mix(t * 2654435761u + (it + i) * 40503u + 12345u) & setmask;
      const unsigned long long *p = tags + 4ull * s[i];
      if(MODE == 0 || MODE == 1 || MODE == 4) asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[i][0]), "=l"(v[i][1]), "=l"(v[i][2]), "=l"(v[i][3]) : "l"(p));
      else if(MODE == 3) { asm volatile("ld.global.v2.u64 {%0,%1}, [%2];" : "=l"(v[i][0]), "=l"(v[i][1]) : "l"(p)); v[i][2] = v[i][3] = 0; }
      else if(MODE == 2) { v[i][0] = atomicAdd(tags + 4ull * s[i] + (s[i] >> 19), 1ull << 50); v[i][1] = v[i][2] = v[i][3] = 0; }
      else { v[i][0] = s[i] * 0x9E3779B97F4A7C15ull; v[i][1] = v[i][2] = v[i][3] = 0; }
    }
#pragma unroll
    for(int i = 0; i < ILP; i++) {
      uint64_t x = v[i][0] ^ v[i][1] ^ v[i][2] ^ v[i][3]; uint32_t way = (uint32_t)(x >> 61) & 3u; acc += (uint32_t)x;
      if(MODE == 0 || MODE == 3 || MODE == 5) atomicAdd(cnt + 4u * s[i] + way, 1u);
      if(MODE == 1) atomicAdd(cnt + 2u * s[i] + (way >> 1), (way & 1u) ? 0x10000u : 1u);
    }
  }
  if(acc == 0x12345678u) sink[0] = acc;
}
template <int MODE, int ILP> void run(const char *name, unsigned long long *tags, unsigned int *cnt, unsigned long long *sink, uint32_t setbits = 21, int ctas_per_sm = 8)
{
  int sms = 148; uint32_t iters = 2048;
  dim3 grid(sms * ctas_per_sm), block(256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, ILP><<<grid, block>>>(tags, cnt, (1u << setbits) - 1u, 256, sink);
  cudaEventRecord(e0);
  k<MODE, ILP><<<grid, block>>>(tags, cnt, (1u << setbits) - 1u, iters, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double n = (double)grid.x * 256 * iters;
  printf("%-60s sets 2^%u ILP %d ctas/SM %d : %7.1f G probes/s\n", name, setbits, ILP, ctas_per_sm, n / ms / 1e6);
}
int main()
{
  unsigned long long *tags, *sink; unsigned int *cnt;
  cudaMalloc(&tags, 32ull << 23); cudaMemset(tags, 0, 32ull << 23); cudaMalloc(&cnt, 16ull << 23); cudaMemset(cnt, 0, 16ull << 23); cudaMalloc(&sink, 8);
  for(uint32_t sb : {21u, 20u, 22u}) {
    run<4, 2>("ld32B tags only", tags, cnt, sink, sb);
    run<5, 2>("RED32 counters only", tags, cnt, sink, sb);
    run<0, 1>("ld32B tags (read-only) + RED32 counters (4B each)", tags, cnt, sink, sb);
    run<0, 2>("ld32B tags (read-only) + RED32 counters (4B each)", tags, cnt, sink, sb);
    run<0, 4>("ld32B tags (read-only) + RED32 counters (4B each)", tags, cnt, sink, sb);
    run<0, 2>("ld32B tags (read-only) + RED32 counters (4B each)", tags, cnt, sink, sb, 4);
    run<1, 2>("ld32B tags (read-only) + RED32 on packed 16-bit counters", tags, cnt, sink, sb);
    run<3, 2>("ld16B tags (read-only) + RED32 counters", tags, cnt, sink, sb);
    run<2, 2>("ATOM64 (returns slot) only", tags, cnt, sink, sb);
    run<2, 4>("ATOM64 (returns slot) only", tags, cnt, sink, sb);
    cudaMemset(tags, 0, 32ull << 23);
  }
  return 0;
}
