// microbenchmark 2: why is "set load, then RED into the same sector" 4x slower than either alone?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
enum { LD_PLAIN, LD_CG, LD_CV, LD_NOALLOC, LD_NONE };
enum { OP_RED32, OP_RED64, OP_ATOM32, OP_RED_OTHER_SECTOR, OP_RED_OTHER_LINE, OP_NONE, OP_ST32 };
template <int LD> __device__ __forceinline__ void ld32B(const unsigned long long *p, uint64_t v[4])
{
  if(LD == LD_PLAIN) asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
  if(LD == LD_CG) asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
  if(LD == LD_CV) asm volatile("ld.global.cv.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
  if(LD == LD_NOALLOC) asm volatile("ld.global.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
  if(LD == LD_NONE) { v[0] = v[1] = v[2] = v[3] = (uint64_t)(uintptr_t)p * 0x9E3779B97F4A7C15ull; }
}
template <int LD, int OP, int ILP, int LAG>
__global__ void __launch_bounds__(256) k(unsigned long long *tab, uint32_t setmask, uint32_t iters, unsigned long long *sink)
{
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
  unsigned int *lagp[LAG > 0 ? LAG : 1]; 
#pragma unroll
  for(int i = 0; i < (LAG > 0 ? LAG : 1); i++) lagp[i] = nullptr;
  for(uint32_t it = 0; it < iters; it += ILP) {
    uint32_t s[ILP]; uint64_t v[ILP][4];
#pragma unroll
    for(int i = 0; i < ILP; i++) {
      s[i] = mix(t * 2654435761u + (it + i) * 40503u + 12345u) & setmask;
      ld32B<LD>(tab + 4ull * s[i], v[i]);
    }
#pragma unroll
    for(int i = 0; i < ILP; i++) {
      uint64_t x = v[i][0] ^ v[i][1] ^ v[i][2] ^ v[i][3]; uint32_t way = (uint32_t)(x >> 61) & 3u; acc += (uint32_t)x;
      uint32_t ts = s[i];
      if(OP == OP_RED_OTHER_SECTOR) ts ^= 1u;        // neighbouring 32-byte sector of the same 128-byte line
      if(OP == OP_RED_OTHER_LINE) ts ^= 0x15554u;    // some other line
      unsigned int *p = reinterpret_cast<unsigned int *>(tab + 4ull * ts + way) + 1;
      if(LAG > 0) { unsigned int *q = lagp[0];
#pragma unroll
        for(int j = 0; j + 1 < LAG; j++) lagp[j] = lagp[j + 1];
        lagp[LAG - 1] = p; p = q; if(!p) continue; }
      if(OP == OP_RED32 || OP == OP_RED_OTHER_SECTOR || OP == OP_RED_OTHER_LINE) atomicAdd(p, 1u << 18);
      if(OP == OP_RED64) atomicAdd(reinterpret_cast<unsigned long long *>(p - 1), 1ull << 50);
      if(OP == OP_ATOM32) acc += atomicAdd(p, 1u << 18);
      if(OP == OP_ST32) *reinterpret_cast<volatile unsigned int *>(p) = acc;
    }
  }
  if(acc == 0x12345678u) sink[0] = acc;
}
template <int LD, int OP, int ILP, int LAG> void run(const char *name, unsigned long long *tab, unsigned long long *sink, uint32_t setbits = 21, int ctas_per_sm = 8)
{
  int sms = 148; uint32_t iters = 2048;
  dim3 grid(sms * ctas_per_sm), block(256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<LD, OP, ILP, LAG><<<grid, block>>>(tab, (1u << setbits) - 1u, 256, sink);
  cudaEventRecord(e0);
  k<LD, OP, ILP, LAG><<<grid, block>>>(tab, (1u << setbits) - 1u, iters, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double n = (double)grid.x * 256 * iters;
  printf("%-44s ILP %d lag %d : %7.1f G probes/s\n", name, ILP, LAG, n / ms / 1e6);
}
int main()
{
  unsigned long long *tab, *sink; size_t bytes = 32ull << 24;
  cudaMalloc(&tab, bytes); cudaMemset(tab, 0, bytes); cudaMalloc(&sink, 8);
  run<LD_PLAIN, OP_NONE, 2, 0>("ld plain only", tab, sink);
  run<LD_NONE, OP_RED32, 2, 0>("RED32 only", tab, sink);
  run<LD_NONE, OP_ATOM32, 2, 0>("ATOM32 (return used) only", tab, sink);
  run<LD_NONE, OP_RED64, 2, 0>("RED64 only", tab, sink);
  run<LD_NONE, OP_ST32, 2, 0>("ST32 only", tab, sink);
  run<LD_PLAIN, OP_RED32, 2, 0>("ld plain + RED32 same sector", tab, sink);
  run<LD_CG, OP_RED32, 2, 0>("ld.cg + RED32 same sector", tab, sink);
  run<LD_CV, OP_RED32, 2, 0>("ld.cv + RED32 same sector", tab, sink);
  run<LD_NOALLOC, OP_RED32, 2, 0>("ld no_allocate + RED32 same sector", tab, sink);
  run<LD_PLAIN, OP_RED_OTHER_SECTOR, 2, 0>("ld plain + RED32 other sector same line", tab, sink);
  run<LD_PLAIN, OP_RED_OTHER_LINE, 2, 0>("ld plain + RED32 other line", tab, sink);
  run<LD_PLAIN, OP_RED64, 2, 0>("ld plain + RED64 same sector", tab, sink);
  run<LD_PLAIN, OP_ATOM32, 2, 0>("ld plain + ATOM32 same sector", tab, sink);
  run<LD_PLAIN, OP_ST32, 2, 0>("ld plain + ST32 same sector", tab, sink);
  run<LD_PLAIN, OP_RED32, 2, 2>("ld plain + RED32 same sector, lagged", tab, sink);
  run<LD_PLAIN, OP_RED32, 2, 8>("ld plain + RED32 same sector, lagged", tab, sink);
  run<LD_CG, OP_RED32, 4, 0>("ld.cg + RED32 same sector", tab, sink);
  run<LD_CG, OP_RED32, 1, 0>("ld.cg + RED32 same sector", tab, sink);
  return 0;
}
