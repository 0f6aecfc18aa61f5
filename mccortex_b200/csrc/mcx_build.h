// mcx_build.h -- internal interface between the kernels (mcx_build.cu, mcx_export.cu)
// and the C-ABI layer (mcx_abi.cu).  Not part of the public ABI (include/mcx_gpu.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "mcx_table.cuh"

enum {
  MCX_CNT_KMERS = 0,   // k-mer occurrences loaded            (SeqLoadingStats.num_kmers_loaded)
  MCX_CNT_NOVEL,       // slots claimed                        (num_kmers_novel / ht.num_kmers)
  MCX_CNT_CONTIGS,     // contigs                              (contigs_parsed)
  MCX_CNT_READS,       // read terminators seen                (num_se_reads)
  MCX_CNT_FULL,        // non-zero: table (or tuple bin) overflowed
  MCX_CNT_INSERTED,    // tuples inserted by kernel C
  MCX_CNT_RECS_LOADED, // graph-file records merged (mcx_ctxload.cu)
  MCX_CNT_NOTFOUND,    // must-exist builds: windows of a contig whose k-mer is not in the graph
  MCX_NCOUNTERS = 8,   // counters the build kernels reduce in shared memory
  MCX_CNT_DUP_SE = 8,  // --remove-pcr: single-end reads dropped       (num_dup_se_reads)
  MCX_CNT_DUP_PE,      // --remove-pcr: read pairs dropped              (num_dup_pe_pairs)
  MCX_NCOUNTERS_ALL = 10 // size of the device counter block
};

struct McxBuildParams {
  const uint8_t *seq;   // LINES layout, 16-byte aligned, readable up to nbytes rounded up to 16
  uint64_t nbytes;      // bytes of data in seq
  uint64_t r_begin;     // positions [r_begin, r_end) are owned by this launch (window starts and terminators)
  uint64_t r_end;
  uint32_t k;
  uint32_t hp_cutoff;
  uint32_t colour;
  int may_saturate;
  unsigned long long *counters; // MCX_NCOUNTERS u64, device
  // quality cut-off (all NULL/0 when off)
  const uint8_t *qual;  // quality bytes parallel to seq (same layout; terminator bytes ignored)
  uint32_t qcut;        // cut-off including the FASTQ ASCII offset
  uint8_t *summary;     // one byte per chunk of [r_begin, r_end): carry summaries (pass 1 -> pass 2)
  // key classes (kernel A2 only): this launch handles the keys whose front hash has its top ncls_log2 bits == cls
  uint32_t cls, ncls_log2;
  uint32_t run_tiles;   // kernel A2: consecutive tiles a warp takes at a time (0 = default)
};

// tuples for other shards: bin d holds up to cap tuples for shard d
//   key  : W x u64 canonical key words
//   meta : u32 = (count << 8) | edge mask     (count 1 for a single occurrence, larger for a
//          record aggregated in the sender's front table)
#define MCX_MAX_PARTS 16
struct McxTupleBins {
  uint64_t *keys[MCX_MAX_PARTS]; // bin of shard d: cap * W words.  Local memory, or memory of GPU d mapped
  uint32_t *meta[MCX_MAX_PARTS]; // over NVLink (CUDA IPC): then the kernel's stores ARE the all-to-all
  unsigned long long *cursor;    // nparts, local
  uint64_t cap;                  // tuples per destination
  uint32_t nparts;
  uint32_t my_part;              // sharded kernels: tuples owned by my_part are inserted locally instead
};

#if defined(__CUDACC__)
// append one tuple to the bin of shard d (warp-aggregated cursor bump)
template <int W>
__device__ __forceinline__ void mcx_bin_push(const McxTupleBins &b, uint32_t d, const McxKmer<W> &key, uint32_t meta, uint32_t &full)
{
  uint32_t peers = __match_any_sync(__activemask(), d);
  uint32_t leader = __ffs(peers) - 1u, lane = threadIdx.x & 31u;
  unsigned long long base = 0;
  if(lane == leader) base = atomicAdd(&b.cursor[d], (unsigned long long)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  uint64_t at = base + __popc(peers & ((1u << lane) - 1u));
  if(at >= b.cap) { full = 1; return; }
  uint64_t *kd = b.keys[d] + at * W;
#pragma unroll
  for(int w = 0; w < W; w++) kd[w] = key.b[w];
  b.meta[d][at] = meta;
}
#endif

cudaError_t mcx_launch_build_fused(const McxBuildParams &p, const McxTable &t, cudaStream_t st);
// kernel A2 (mcx_build_warp.cu, only in `make EXPERIMENTS=1` builds: it is slower than kernel A on the bench workload,
// DESIGN.md 4): warp-autonomous front end; k <= 31, no quality / homopolymer cut-off
bool mcx_warp_kernel_supports(const McxBuildParams &p);
cudaError_t mcx_launch_build_warp(const McxBuildParams &p, const McxTable &t, cudaStream_t st);
cudaError_t mcx_launch_build_warp_sharded(const McxBuildParams &p, const McxTable &t, const McxTupleBins &b, cudaStream_t st);
cudaError_t mcx_launch_build_fused_qual(const McxBuildParams &p, const McxTable &t, cudaStream_t st);
cudaError_t mcx_launch_kmer_tuples(const McxBuildParams &p, const McxTupleBins &b, cudaStream_t st);
// n_dev (may be NULL): device word holding the tuple count, read by the kernel (min(*n_dev, n) tuples)
cudaError_t mcx_launch_insert_tuples(const uint64_t *keys, const uint32_t *meta, uint64_t n, const uint64_t *n_dev, uint32_t k,
                                     const McxTable &t, uint32_t colour, int may_saturate, unsigned long long *counters,
                                     cudaStream_t st);
// sharded build: fused local front table, big-table inserts for owned keys, tuples for the rest
cudaError_t mcx_launch_build_sharded(const McxBuildParams &p, const McxTable &t, const McxTupleBins &b, cudaStream_t st);
cudaError_t mcx_launch_front_flush_sharded(const McxTable &t, const McxTupleBins &b, int may_saturate,
                                           unsigned long long *counters, cudaStream_t st);
cudaError_t mcx_launch_repack_lines(const uint8_t *src, const uint64_t *off, uint64_t nreads, uint8_t *dst, cudaStream_t st);

cudaError_t mcx_launch_front_flush(const McxTable &t, int may_saturate, unsigned long long *counters, cudaStream_t st);

// graph files (mcx_ctxload.cu): from_col / into_col are device arrays of nmap colour pairs; flags bit 0 = must exist
cudaError_t mcx_launch_load_records(const uint8_t *recs, uint64_t n, uint32_t file_ncols, const uint32_t *from_col,
                                    const uint32_t *into_col, uint32_t nmap, uint32_t flags, uint32_t k, const McxTable &t,
                                    uint8_t *isec_edges, unsigned long long *counters, cudaStream_t st);
// build --intersect (mcx_lookup.cu): reads update only k-mers that are already in the table; the final pass
// removes k-mers without coverage and ANDs every colour's edges with isec_edges
cudaError_t mcx_launch_build_lookup(const McxBuildParams &p, const McxTable &t, cudaStream_t st);
cudaError_t mcx_launch_contig_summary(const McxBuildParams &p, cudaStream_t st);
cudaError_t mcx_launch_finish_intersect(const McxTable &t, uint32_t W, const uint8_t *isec_edges, unsigned long long *nkept,
                                        cudaStream_t st);
// build --remove-pcr (mcx_pcr.cu): first = 2 * nslots u32 (MCX_PCR_UNSET when no read has started there), node = nreads
// u64 scratch; seq / qual are modified in place (reads re-oriented, duplicates overwritten with 'N')
cudaError_t mcx_launch_pcr_filter(uint8_t *seq, uint8_t *qual, const uint64_t *off, const uint8_t *mate, uint64_t nreads,
                                  uint32_t k, uint32_t qcut, uint32_t hp, const McxTable &t, uint32_t *first, uint64_t *node,
                                  uint32_t ord_base, unsigned long long *counters, cudaStream_t st);
#define MCX_KEY_TOMBSTONE (MCX_KEY_FLAG | (1ULL << 62))  /* slot of a removed k-mer: never equal to a key, never empty */

// export (mcx_export.cu)
struct McxExport {
  uint8_t *records;   // nrec * rec_bytes, .ctx record layout, ascending key order if sorted
  uint64_t nrec;
  uint32_t rec_bytes;
  cudaStream_t stream; // the records come from the stream-ordered pool: freed on this stream
};
void mcx_pool_trim(int dev); // give the pool's cached blocks back to the device (before a large cudaMalloc is retried)
cudaError_t mcx_export_build(const McxTable &t, uint32_t kmer_size, bool sorted, McxExport *out, cudaStream_t st);
void mcx_export_free(McxExport *e);
cudaError_t mcx_sort_records_device(const uint8_t *d_in, uint64_t n, uint32_t k, uint32_t ncols, uint8_t *d_out, cudaStream_t st);
