// mcx_export.cu -- dump the device table as .ctx v6 records in ascending key order.
//
// Replaces (reference, relative to /root/reference):
//   HASH_ITERATE_SORTED / hash_table_sorted   src/graph/hash_table.h:115-120, hash_table.c:362-374
//   graph_write_kmer                          src/graph/graph_writer.c:116-127
// The reference qsorts an array of pointers with an indirect compare and fwrites three
// fields per k-mer.  Here: compact occupied slots -> LSD radix sort of (key word, slot
// index) pairs (mcx_radix.cu, in-tree; it only touches the 2k key bits) -> one kernel that
// gathers key/covg/edges of each slot and writes the packed 8W+5C byte records through
// shared memory so global stores stay coalesced.
#include <cuda_runtime.h>
#include <stdint.h>
#include "mcx_build.h"
#include "mcx_radix.cuh"

#define CK(x) do { cudaError_t e_ = (x); if(e_ != cudaSuccess) { err = e_; goto fail; } } while(0)

// scratch and output come from the device's stream-ordered pool (cudaMallocAsync): freed blocks stay in the pool (its
// release threshold is raised by pool_keep()), so the second export of a process does not pay for cudaMalloc / cudaFree of
// gigabytes again (config 2: 0.6 s of a 0.78 s job before).  mcx_pool_trim() gives the memory back when a table
// allocation would otherwise fail.
static void pool_keep(int dev)
{
  static bool done[64];
  if(dev < 0 || dev >= 64 || done[dev]) return;
  cudaMemPool_t pool;
  if(cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  cudaGetLastError();
  done[dev] = true;
}
void mcx_pool_trim(int dev)
{
  cudaMemPool_t pool;
  if(cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
  cudaGetLastError();
}

// number of occupied slots (the compaction buffers are then sized exactly)
__global__ void mcx_count_kernel(McxTable t, unsigned long long *count)
{
  unsigned long long n = 0;
  for(uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < t.nslots; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k0 = *reinterpret_cast<const uint64_t *>(t.slots + i * (uint64_t)t.stride);
    n += (k0 != 0 && k0 != MCX_KEY_TOMBSTONE);
  }
  for(int s = 16; s > 0; s >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, s);
  if((threadIdx.x & 31u) == 0 && n) atomicAdd(count, n);
}

// occupied slots -> (low key word, slot index) pairs
template <int W>
__global__ void mcx_compact_kernel(McxTable t, uint64_t *__restrict__ keys, uint64_t *__restrict__ slots,
                                   unsigned long long *cursor)
{
  const uint32_t lane = threadIdx.x & 31u;
  uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  uint64_t rounds = (t.nslots + nthreads - 1) / nthreads;
  for(uint64_t r = 0; r < rounds; r++) {
    uint64_t i = r * nthreads + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t k0 = 0;
    if(i < t.nslots) k0 = *reinterpret_cast<const uint64_t *>(t.slots + i * (uint64_t)t.stride);
    bool occ = k0 != 0 && k0 != MCX_KEY_TOMBSTONE; // (a k-mer removed by an intersected build)
    uint32_t m = __ballot_sync(0xFFFFFFFFu, occ);
    if(!m) continue;
    unsigned long long base = 0;
    if(lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if(occ) {
      uint64_t at = base + __popc(m & ((1u << lane) - 1u));
      // sort key of the first pass = least significant key word
      keys[at] = (W == 1) ? (k0 & ~MCX_KEY_FLAG)
                          : *reinterpret_cast<const uint64_t *>(t.slots + i * (uint64_t)t.stride + 2);
      slots[at] = i;
    }
  }
}

// second pass key for W == 2: the most significant word of each (already b[1]-sorted) slot
__global__ void mcx_gather_hi_kernel(McxTable t, const uint64_t *__restrict__ slots, uint64_t n, uint64_t *__restrict__ keys)
{
  for(uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    keys[i] = *reinterpret_cast<const uint64_t *>(t.slots + slots[i] * (uint64_t)t.stride) & ~MCX_KEY_FLAG;
}

// record i <- slot order[i] : W x u64 key (flag cleared), C x u32 covg, C x u8 edges
// A block formats rpb records at a time in shared memory (rpb = as many as fit the shared-memory budget, at most
// MCX_EXP_THREADS: a 4096-colour record is 20 KB), word by word across the block's threads, then writes them out
// with coalesced stores.
#define MCX_EXP_THREADS 128
#define MCX_EXP_SMEM (96u * 1024u)
__global__ void __launch_bounds__(MCX_EXP_THREADS) mcx_format_kernel(McxTable t, uint32_t W, const uint64_t *__restrict__ order,
                                                                     uint64_t n, uint32_t rec_bytes, uint32_t rpb, uint8_t *__restrict__ out)
{
  extern __shared__ __align__(16) uint8_t tile[];
  const uint32_t C = t.ncols, nw = 2u * W + C;
  for(uint64_t blk = blockIdx.x; blk * rpb < n; blk += gridDim.x) {
    const uint64_t first = blk * rpb;
    const uint32_t cnt = (uint32_t)((n - first < rpb) ? (n - first) : rpb);
    for(uint32_t idx = threadIdx.x; idx < cnt * nw; idx += blockDim.x) {
      const uint32_t r = idx / nw, w = idx - r * nw;
      uint32_t v = t.slots[order[first + r] * (uint64_t)t.stride + w];
      if(w == 1u) v &= 0x7FFFFFFFu; // MCX_KEY_FLAG lives in the top bit of key word b[0]
      uint8_t *d = tile + r * rec_bytes + 4u * w;
      d[0] = (uint8_t)v; d[1] = (uint8_t)(v >> 8); d[2] = (uint8_t)(v >> 16); d[3] = (uint8_t)(v >> 24);
    }
    for(uint32_t idx = threadIdx.x; idx < cnt * C; idx += blockDim.x) {
      const uint32_t r = idx / C, c = idx - r * C;
      tile[r * rec_bytes + 4u * nw + c] = (uint8_t)(t.slots[order[first + r] * (uint64_t)t.stride + nw + (c >> 2)] >> (8u * (c & 3u)));
    }
    __syncthreads();
    const uint64_t obase = first * rec_bytes;
    const uint32_t nbytes = cnt * rec_bytes;
    if(((obase | nbytes) & 3u) == 0) {
      uint32_t *o32 = reinterpret_cast<uint32_t *>(out + obase);
      const uint32_t *t32 = reinterpret_cast<const uint32_t *>(tile);
      for(uint32_t i = threadIdx.x; i < nbytes / 4u; i += blockDim.x) o32[i] = t32[i];
    } else {
      for(uint32_t i = threadIdx.x; i < nbytes; i += blockDim.x) out[obase + i] = tile[i];
    }
    __syncthreads();
  }
}

// one sort of the (key word, index) pairs by bits [0, end_bit): the in-tree radix sort (mcx_radix.cu); leaves the sorted
// data in (*keys, *vals)
static cudaError_t radix_pass(uint64_t **keys, uint64_t **vals, uint64_t **keys_alt, uint64_t **vals_alt, uint64_t n,
                              int end_bit, void **tmp, size_t *tmp_bytes, cudaStream_t st)
{
  const size_t need = mcx_radix_scratch_bytes(n);
  cudaError_t e;
  if(need > *tmp_bytes) {
    if(*tmp) cudaFreeAsync(*tmp, st);
    *tmp = nullptr; *tmp_bytes = 0;
    e = cudaMallocAsync(tmp, need, st);
    if(e != cudaSuccess) return e;
    *tmp_bytes = need;
  }
  uint64_t *ok, *ov;
  e = mcx_radix_sort_pairs(*keys, *vals, *keys_alt, *vals_alt, n, end_bit, *tmp, &ok, &ov, st);
  if(e != cudaSuccess) return e;
  if(ok != *keys) { uint64_t *x = *keys; *keys = *keys_alt; *keys_alt = x; }
  if(ov != *vals) { uint64_t *x = *vals; *vals = *vals_alt; *vals_alt = x; }
  return cudaSuccess;
}

cudaError_t mcx_export_build(const McxTable &t, uint32_t k, bool sorted, McxExport *out, cudaStream_t st)
{
  cudaError_t err = cudaSuccess;
  const uint32_t W = (k + 31u) / 32u;
  unsigned long long *cursor = nullptr;
  uint64_t *keys = nullptr, *vals = nullptr, *keys_alt = nullptr, *vals_alt = nullptr;
  void *tmp = nullptr; size_t tmp_bytes = 0;
  unsigned long long n = 0, n2 = 0;
  int sms = 148, dev = 0;
  out->records = nullptr; out->nrec = 0; out->rec_bytes = 8u * W + 5u * t.ncols; out->stream = st;

  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  pool_keep(dev);

  CK(cudaMallocAsync(&cursor, 2 * sizeof(*cursor), st));
  CK(cudaMemsetAsync(cursor, 0, 2 * sizeof(*cursor), st));
  mcx_count_kernel<<<sms * 8, 256, 0, st>>>(t, cursor + 1);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(&n, cursor + 1, sizeof(n), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if(n) {
    CK(cudaMallocAsync(&keys, n * sizeof(uint64_t), st));
    CK(cudaMallocAsync(&vals, n * sizeof(uint64_t), st));
    if(W == 1) mcx_compact_kernel<1><<<sms * 8, 256, 0, st>>>(t, keys, vals, cursor);
    else mcx_compact_kernel<2><<<sms * 8, 256, 0, st>>>(t, keys, vals, cursor);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&n2, cursor, sizeof(n2), cudaMemcpyDeviceToHost, st));   // (nothing inserts during an export)
    if(sorted) {
      CK(cudaMallocAsync(&keys_alt, n * sizeof(uint64_t), st));
      CK(cudaMallocAsync(&vals_alt, n * sizeof(uint64_t), st));
      if(W == 1) {
        CK(radix_pass(&keys, &vals, &keys_alt, &vals_alt, n, (int)(2u * k), &tmp, &tmp_bytes, st));
      } else {
        CK(radix_pass(&keys, &vals, &keys_alt, &vals_alt, n, 64, &tmp, &tmp_bytes, st));
        mcx_gather_hi_kernel<<<sms * 8, 256, 0, st>>>(t, vals, n, keys);
        CK(cudaGetLastError());
        CK(radix_pass(&keys, &vals, &keys_alt, &vals_alt, n, (int)(2u * (k - 32u)), &tmp, &tmp_bytes, st));
      }
    }
    CK(cudaMallocAsync(&out->records, n * (uint64_t)out->rec_bytes + 16, st));
    // records per block: what fits the shared-memory budget (a multiple of 4 keeps the vector stores aligned)
    uint32_t rpb = MCX_EXP_SMEM / out->rec_bytes;
    if(rpb > MCX_EXP_THREADS) rpb = MCX_EXP_THREADS;
    if(rpb >= 4u) rpb &= ~3u;
    if(rpb == 0u) rpb = 1u; // (rec_bytes <= 8 * 2 + 5 * 4096 < MCX_EXP_SMEM: never)
    uint64_t nblk = (n + rpb - 1) / rpb, cap = (uint64_t)sms * 16;
    size_t smem = (size_t)rpb * out->rec_bytes;
    if(smem > 48 * 1024) CK(cudaFuncSetAttribute(mcx_format_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mcx_format_kernel<<<(unsigned)(nblk < cap ? nblk : cap), MCX_EXP_THREADS, smem, st>>>(t, W, vals, n, out->rec_bytes, rpb, out->records);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    if(n2 != n) { err = cudaErrorUnknown; goto fail; }
  }
  out->nrec = n;
fail:
  if(cursor) cudaFreeAsync(cursor, st);
  if(keys) cudaFreeAsync(keys, st);
  if(vals) cudaFreeAsync(vals, st);
  if(keys_alt) cudaFreeAsync(keys_alt, st);
  if(vals_alt) cudaFreeAsync(vals_alt, st);
  if(tmp) cudaFreeAsync(tmp, st);
  if(err != cudaSuccess && out->records) { cudaFreeAsync(out->records, st); out->records = nullptr; }
  return err;
}

// ---------------------------------------------------------------- sort of a graph FILE's records
// Replaces ctx_sort (src/commands/ctx_sort.c:38-160): the records of a .ctx file (W x u64 key,
// C x u32 covg, C x u8 edges, packed) are ordered by key.  The reference qsorts pointers with an
// unaligned compare; here the keys are pulled out of the packed records, (key word, index) pairs
// are radix sorted (stable; two passes for k > 31) and the records gathered in that order.
__global__ void mcx_rec_keys_kernel(const uint8_t *__restrict__ recs, uint64_t n, uint32_t rec_bytes, uint32_t word,
                                    const uint64_t *__restrict__ order, uint64_t *__restrict__ keys, uint64_t *__restrict__ idx)
{
  for(uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = order ? order[i] : i;
    const uint8_t *p = recs + r * rec_bytes + 8u * word;
    uint64_t v = 0;
#pragma unroll
    for(int b = 7; b >= 0; b--) v = (v << 8) | p[b];
    keys[i] = v;
    if(!order) idx[i] = i;
  }
}
__global__ void mcx_rec_gather_kernel(const uint8_t *__restrict__ recs, const uint64_t *__restrict__ order, uint64_t n,
                                      uint32_t rec_bytes, uint8_t *__restrict__ out)
{
  // one warp per record: lanes copy its bytes
  const uint32_t lane = threadIdx.x & 31u;
  uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for(uint64_t i = warp; i < n; i += nwarps) {
    const uint8_t *s = recs + order[i] * rec_bytes;
    uint8_t *d = out + i * rec_bytes;
    for(uint32_t b = lane; b < rec_bytes; b += 32u) d[b] = s[b];
  }
}

cudaError_t mcx_sort_records_device(const uint8_t *d_in, uint64_t n, uint32_t k, uint32_t ncols, uint8_t *d_out, cudaStream_t st)
{
  cudaError_t err = cudaSuccess;
  const uint32_t W = (k + 31u) / 32u, rec_bytes = 8u * W + 5u * ncols;
  uint64_t *keys = nullptr, *vals = nullptr, *keys_alt = nullptr, *vals_alt = nullptr;
  void *tmp = nullptr; size_t tmp_bytes = 0;
  int sms = 148, dev = 0;
  if(n == 0) return cudaSuccess;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  CK(cudaMallocAsync(&keys, n * sizeof(uint64_t), st));
  CK(cudaMallocAsync(&vals, n * sizeof(uint64_t), st));
  CK(cudaMallocAsync(&keys_alt, n * sizeof(uint64_t), st));
  CK(cudaMallocAsync(&vals_alt, n * sizeof(uint64_t), st));
  // least significant key word first
  mcx_rec_keys_kernel<<<sms * 8, 256, 0, st>>>(d_in, n, rec_bytes, W - 1u, nullptr, keys, vals);
  CK(cudaGetLastError());
  // all 64 bits of every word: the reference's `sort` compares whole words whatever k the header claims
  // (ctx_sort.c:117-155, binary_kmer.h:79-94), so a file with bits above 2k comes out in the same order
  CK(radix_pass(&keys, &vals, &keys_alt, &vals_alt, n, 64, &tmp, &tmp_bytes, st));
  if(W == 2) {
    mcx_rec_keys_kernel<<<sms * 8, 256, 0, st>>>(d_in, n, rec_bytes, 0u, vals, keys, nullptr);
    CK(cudaGetLastError());
    CK(radix_pass(&keys, &vals, &keys_alt, &vals_alt, n, 64, &tmp, &tmp_bytes, st));
  }
  mcx_rec_gather_kernel<<<sms * 16, 256, 0, st>>>(d_in, vals, n, rec_bytes, d_out);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
fail:
  if(keys) cudaFreeAsync(keys, st);
  if(vals) cudaFreeAsync(vals, st);
  if(keys_alt) cudaFreeAsync(keys_alt, st);
  if(vals_alt) cudaFreeAsync(vals_alt, st);
  if(tmp) cudaFreeAsync(tmp, st);
  return err;
}

void mcx_export_free(McxExport *e)
{
  if(e && e->records) { cudaFreeAsync(e->records, e->stream); e->records = nullptr; }
  if(e) e->nrec = 0;
}
