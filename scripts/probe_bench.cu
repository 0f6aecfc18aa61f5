// microbenchmark: how fast can an SM-resident kernel do "random 32-byte set load + 32-bit RED" on an
// L2-resident table?  (the access pattern of the front-table hot pass, with no k-mer math)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
template <int MODE, int ILP, int LDB>
__global__ void __launch_bounds__(256) k(unsigned long long *tab, uint32_t setmask, uint32_t iters, unsigned long long *sink)
{
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
  for(uint32_t it = 0; it < iters; it += ILP) {
    uint32_t s[ILP]; uint64_t v[ILP][4];
#pragma unroll
    for(int i = 0; i < ILP; i++) {
      s[i] = mix(t * 2654435761u + (it + i) * 40503u + 12345u) & setmask;
      if(MODE & 1) {
        if(LDB == 32) asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[i][0]), "=l"(v[i][1]), "=l"(v[i][2]), "=l"(v[i][3]) : "l"(tab + 4ull * s[i]));
        else if(LDB == 16) { asm volatile("ld.global.v2.u64 {%0,%1}, [%2];" : "=l"(v[i][0]), "=l"(v[i][1]) : "l"(tab + 4ull * s[i])); v[i][2] = v[i][3] = 0; }
        else { asm volatile("ld.global.u64 %0, [%1];" : "=l"(v[i][0]) : "l"(tab + 4ull * s[i])); v[i][1] = v[i][2] = v[i][3] = 0; }
      }
    }
#pragma unroll
    for(int i = 0; i < ILP; i++) {
      uint32_t way = 0;
      if(MODE & 1) { uint64_t x = v[i][0] ^ v[i][1] ^ v[i][2] ^ v[i][3]; way = (uint32_t)(x >> 61) & 3u; acc += (uint32_t)x; }
      else way = s[i] & 3u;
      if(MODE & 2) atomicAdd(reinterpret_cast<unsigned int *>(tab + 4ull * s[i] + way) + 1, 1u << 18);
    }
  }
  if(acc == 0x12345678u) sink[0] = acc;
}
template <int MODE, int ILP, int LDB> void run(const char *name, unsigned long long *tab, uint32_t setbits, int ctas_per_sm, unsigned long long *sink)
{
  int sms = 148; uint32_t iters = 4096;
  dim3 grid(sms * ctas_per_sm), block(256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, ILP, LDB><<<grid, block>>>(tab, (1u << setbits) - 1u, 256, sink);
  cudaEventRecord(e0);
  k<MODE, ILP, LDB><<<grid, block>>>(tab, (1u << setbits) - 1u, iters, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double n = (double)grid.x * 256 * iters;
  printf("%-28s sets 2^%u (%4u MB) ctas/SM %d ILP %d ld %2dB : %7.1f G probes/s\n", name, setbits, (32u << setbits) >> 20, ctas_per_sm, ILP, LDB, n / ms / 1e6);
}
int main()
{
  unsigned long long *tab, *sink; size_t bytes = 32ull << 24;
  cudaMalloc(&tab, bytes); cudaMemset(tab, 0, bytes); cudaMalloc(&sink, 8);
  for(uint32_t sb : {21u, 20u, 23u}) {
    run<1, 1, 32>("load only", tab, sb, 8, sink);
    run<1, 2, 32>("load only", tab, sb, 8, sink);
    run<1, 4, 32>("load only", tab, sb, 8, sink);
    run<1, 4, 16>("load only", tab, sb, 8, sink);
    run<1, 4, 8>("load only", tab, sb, 8, sink);
    run<2, 1, 32>("RED only", tab, sb, 8, sink);
    run<2, 4, 32>("RED only", tab, sb, 8, sink);
    run<3, 1, 32>("load+RED", tab, sb, 4, sink);
    run<3, 1, 32>("load+RED", tab, sb, 8, sink);
    run<3, 2, 32>("load+RED", tab, sb, 4, sink);
    run<3, 2, 32>("load+RED", tab, sb, 8, sink);
    run<3, 4, 32>("load+RED", tab, sb, 4, sink);
    run<3, 4, 32>("load+RED", tab, sb, 8, sink);
    run<3, 4, 8>("load+RED", tab, sb, 8, sink);
    cudaMemset(tab, 0, bytes);
  }
  return 0;
}
