/*
 * mcx_gpu.h -- C ABI of libmcxgpu.so: the B200 (sm_100a) implementation of the
 * `mccortex build` hot path.  Plain C, plain pointers and sizes; no CUDA or
 * torch types.  Every entry point returns an int status (MCX_OK == 0) and
 * never exits the process; the host driver maps non-zero to the reference's
 * die() texts ("Hash table is full", ...).
 *
 * The reference (mcveanlab/mccortex) has no FFI; the boundary below is the
 * in-process C interface that src/commands/ctx_build.c calls.  Each function
 * cites the reference interface it replaces (paths relative to the reference
 * root).  INTEGRATION.md shows the binding a maintainer would add.
 *
 * Threading: one host thread per mcx_graph at a time (calls are asynchronous
 * with respect to the GPU; mcx_graph_sync joins).  Several graphs / devices may
 * be driven from different threads.
 */
#ifndef MCX_GPU_H_
#define MCX_GPU_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------ */
#define MCX_OK               0
#define MCX_ERR_BAD_ARG      1
#define MCX_ERR_CUDA         2  /* see mcx_last_error() */
#define MCX_ERR_TABLE_FULL   3  /* reference: die("Hash table is full"), src/graph/hash_table.c:119-123,280 */
#define MCX_ERR_NOMEM        4  /* reference: ctx_malloc dies, src/global/ctx_alloc.h:24-27 */
#define MCX_ERR_UNSUPPORTED  5
#define MCX_ERR_NO_DEVICE    6  /* no CUDA device / driver: there is NO CPU fallback */

/* ---- batch description -------------------------------------------------- */
/* layout of mcx_read_batch.seq */
#define MCX_LAYOUT_LINES    0  /* every read is followed by exactly one '\n' ("plain" one-read-per-line text) */
#define MCX_LAYOUT_OFFSETS  1  /* reads abut; offsets[nreads+1] gives their boundaries */
/* where seq / qual / offsets live */
#define MCX_MEM_HOST        0  /* pageable or pinned host memory (pinned = no staging copy; see mcx_host_alloc) */
#define MCX_MEM_DEVICE      1  /* device memory of the graph's GPU: 16-byte aligned, readable up to nbytes rounded up to 16 */

/* Mirrors what build_graph_from_reads_mt() receives per read (src/tools/build_graph.h:57-61)
 * plus the SeqLoadingPrefs it is called with (src/tools/build_graph.h:18-24), batched. */
typedef struct {
  const char     *seq;       /* bases, ASCII; any byte outside ACGTacgt breaks contigs like the reference's LUT (src/basic/dna.c:8-25) */
  const char     *qual;      /* NULL, or quality bytes parallel to seq: byte i is the quality of base i (bytes at
                                terminator positions are ignored).  Where a read has no / too short a quality
                                string the caller stores 0x7F (the reference does not filter those positions,
                                src/basic/seq_reader.c:82,149).  Only consulted when fq_cutoff != 0. */
  const uint64_t *offsets;   /* MCX_LAYOUT_OFFSETS: nreads+1 byte offsets into seq; else NULL */
  uint64_t        nreads;    /* MCX_LAYOUT_OFFSETS only (LINES: counted from the terminators) */
  uint64_t        nbytes;    /* bytes in seq (LINES: including the terminators) */
  uint32_t        layout;    /* MCX_LAYOUT_* */
  uint32_t        mem;       /* MCX_MEM_* */
  uint32_t        colour;    /* SeqLoadingPrefs.colour */
  uint8_t         fq_cutoff; /* 0 = off; else quality threshold ALREADY including the FASTQ ASCII offset
                                (build_graph.c:202-207), < 127.  A contig starts at a k-mer whose bases all have
                                qual > cutoff and extends while qual >= cutoff (seq_reader.c:84,149). */
  uint8_t         hp_cutoff; /* 0 = off; else break contigs at homopolymer runs >= hp_cutoff (2 <= hp_cutoff <= k) */
  uint8_t         must_exist; /* SeqLoadingPrefs.must_exist_in_graph (build --intersect): k-mers are looked up, never
                                 inserted; coverage and edges only where the k-mers are in the graph
                                 (src/tools/build_graph.c:99-150). */
  uint8_t         reserved;
} mcx_read_batch;

/* Counters of SeqLoadingStats (src/basic/seq_loading_stats.h:5-14) that feed the
 * .ctx header and the per-task log (src/tools/build_graph.c:173-188,352-386). */
typedef struct {
  uint64_t total_bases_read;
  uint64_t total_bases_loaded;
  uint64_t contigs_parsed;
  uint64_t num_kmers_loaded;
  uint64_t num_kmers_novel;
  uint64_t num_se_reads;
  uint64_t num_pe_reads;
  uint64_t num_good_reads;   /* UINT64_MAX when not computed */
  uint64_t num_bad_reads;    /* UINT64_MAX when not computed */
  uint64_t num_dup_se_reads; /* mcx_graph_add_reads_pcr: single-end reads / read pairs dropped as duplicates */
  uint64_t num_dup_pe_pairs;
} mcx_load_stats;

typedef struct mcx_graph mcx_graph;

/* ---- library ------------------------------------------------------------ */
/* Number of usable CUDA devices (0 => nothing in this library can run). */
int mcx_device_count(void);
/* Text of the last CUDA error seen by the calling thread's last failing call. */
const char *mcx_last_error(void);
/* Pinned host memory for read batches (H2D without a staging copy). */
int mcx_host_alloc(void **ptr, size_t bytes);
int mcx_host_free(void *ptr);

/* ---- graph life cycle --------------------------------------------------- */
/* replaces db_graph_alloc(&g, k, ncols, ncols, capacity, EDGES|COVGS|BKTLOCKS)
 * (src/graph/db_graph.h:62-64, src/commands/ctx_build.c:335-339) + hash_table_alloc
 * (src/graph/hash_table.c:16-52).  capacity = number of k-mer slots (what
 * cmd_get_kmers_in_hash returns, src/graph/cmd_mem.c:38-130); 3 <= k <= 63, k odd. */
#define MCX_GRAPH_INTERSECT 1u  /* flags: also allocate the intersection edge set (Edges *isec_edges, ctx_build.c:341-343) */
#define MCX_GRAPH_READSTRT  2u  /* flags: also allocate the read-start marks of build --remove-pcr (DBG_ALLOC_READSTRT,
                                   ctx_build.c:336; two u32 per k-mer slot here, two bits in the reference) */
int mcx_graph_create(uint32_t kmer_size, uint32_t ncols, uint64_t capacity, int device, uint32_t flags, mcx_graph **out);
/* replaces db_graph_dealloc (src/graph/db_graph.h:67) */
int mcx_graph_destroy(mcx_graph *g);
/* forget all k-mers (table zeroed), keep the allocation */
int mcx_graph_clear(mcx_graph *g);
/* run all subsequent device work of this graph on an existing CUDA stream
 * (a cudaStream_t / CUstream passed as void*; NULL = the graph's own streams) */
int mcx_graph_set_stream(mcx_graph *g, void *cuda_stream);

/* ---- the hot path ------------------------------------------------------- */
/* replaces build_graph(&g, tasks, n, nthreads) -> build_graph_from_reads_mt per read
 * (src/tools/build_graph.c:192-301).  Asynchronous: returns once the batch is
 * queued (host buffers may be reused after return unless they are pinned, in which
 * case they must stay valid until mcx_graph_sync). */
int mcx_graph_add_reads(mcx_graph *g, const mcx_read_batch *batch);
/* optional: allocate now the staging ring that host batches go through (3 x 32 MB of pinned host memory and as much
 * device memory; otherwise the first mcx_graph_add_reads with MCX_MEM_HOST does it).  The reference has no
 * counterpart: its workers read the hosts's read_t buffers in place (src/basic/async_read_io.c:118-141). */
int mcx_graph_prepare_host(mcx_graph *g);
/* replaces build_graph_from_str_mt(&g, colour, seq, len, false) (src/tools/build_graph.h:77-79):
 * one contig-to-be, synchronous. */
int mcx_graph_add_str(mcx_graph *g, uint32_t colour, const char *seq, size_t len);
/* join all queued work; *stats (may be NULL) receives the counters accumulated since the
 * previous sync (the reference merges per-thread stats the same way,
 * src/tools/build_graph.c:285-288).  Returns MCX_ERR_TABLE_FULL if any insert overflowed. */
int mcx_graph_sync(mcx_graph *g, mcx_load_stats *stats);
/* queue (asynchronously, on the graph's stream) the merge of everything the L2-resident front
 * table has aggregated into the big table; mcx_graph_sync / export do this implicitly.  After it
 * completes the big table alone holds the graph, as the reference's hash table does after
 * build_graph() returns (src/tools/build_graph.c:283-300). */
int mcx_graph_flush(mcx_graph *g);

/* replaces graph_load(file, prefs, stats) (src/graph/graphs_load.c:83-208) for records the caller has
 * read from a .ctx file: nrecords packed records of the FILE's layout (W x u64 key, file_ncols x u32
 * covg, file_ncols x u8 edges).  (from_col[i], into_col[i]), i < nmap, is the reference's FileFilter
 * (src/basic/file_filter.h): coverage of file colour from is added (saturating) to graph colour into,
 * edges are ORed; a k-mer whose selected colours all have zero coverage is skipped.
 * flags: MCX_LOAD_MUST_EXIST = only k-mers already in the graph (GraphLoadingPrefs.must_exist_in_graph).
 * Synchronous.  *nkmers_loaded / *nkmers_novel (may be NULL) = GraphLoadingStats of this call.
 * The novel k-mers are also part of what the next mcx_graph_sync reports as num_kmers_novel. */
#define MCX_LOAD_MUST_EXIST 1u
#define MCX_LOAD_INTO_ISEC  2u  /* the file is an intersection graph (ctx_build.c:348-361): k-mers are inserted without
                                   coverage, the edges of all selected colours go into the intersection edge set */
#define MCX_LOAD_MASK_ISEC  4u  /* GraphLoadingPrefs.must_exist_in_edges: edges are ANDed with the intersection edge set */
int mcx_graph_load_records(mcx_graph *g, const void *records, uint64_t nrecords, uint32_t file_ncols, uint32_t mem,
                           const uint32_t *from_col, const uint32_t *into_col, uint32_t nmap, uint32_t flags,
                           uint64_t *nkmers_loaded, uint64_t *nkmers_novel);

/* replaces db_graph_remove_no_covg_kmers + db_graph_intersect_edges (src/graph/db_graph.c:632-673,
 * ctx_build.c:409-413) at the end of a build --intersect: k-mers without coverage in any colour leave the
 * graph, every colour's edges are ANDed with the intersection edge set.  *nkmers = k-mers left. */
int mcx_graph_finish_intersect(mcx_graph *g, uint64_t *nkmers);

/* ---- build --remove-pcr ---------------------------------------------------- */
/* replaces build_graph_from_reads_mt with SeqLoadingPrefs.remove_pcr_dups (src/tools/build_graph.c:192-231), i.e.
 * seq_reads_are_novel (:35-92) in front of load_read, for one batch.  batch: MCX_LAYOUT_LINES, host or device memory
 * (device buffers are MODIFIED: reads re-oriented, duplicates overwritten with 'N'); read_off[nreads + 1] = byte offset
 * of every read's first base (read_off[nreads] = nbytes), mate[nreads] = MCX_MATE_* per read, both in the same memory
 * as the batch; the two reads of a pair are consecutive and in the same batch.  A read (pair) is dropped when each of its mates that has a k-mer starts, in the same orientation,
 * on a k-mer where a mate of an EARLIER read (pair) of this colour started -- "earlier" in the order of the calls and
 * of the reads inside a batch, which is what the reference does with one worker thread and one input task (with more
 * threads its result depends on scheduling).  MCX_MATE_REVCOMP is seq_reader_orient_mp_FF (seq_reader.c:506-510): the
 * caller sets it on read 1 when matedir & 2 and on read 2 when matedir & 1; the read is loaded in that orientation.
 * Synchronous.  mcx_graph_sync reports the dropped reads / pairs in num_dup_se_reads / num_dup_pe_pairs. */
#define MCX_MATE_SINGLE  0u
#define MCX_MATE_FIRST   1u  /* first read of a pair: the next read is its mate */
#define MCX_MATE_SECOND  2u
#define MCX_MATE_REVCOMP 4u  /* OR-ed in: reverse-complement the read and reverse its qualities first */
int mcx_graph_add_reads_pcr(mcx_graph *g, const mcx_read_batch *batch, const uint64_t *read_off, const uint8_t *mate,
                            uint64_t nreads);
/* forget all read starts: the memset of db_graph.readstrt when the colour changes (ctx_build.c:392-395) */
int mcx_graph_pcr_reset(mcx_graph *g);

/* replaces hash_table_print_stats inputs (src/graph/hash_table.h:73): occupancy */
int mcx_graph_stats(mcx_graph *g, uint64_t *nkmers, uint64_t *capacity);

/* ---- dump --------------------------------------------------------------- */
/* replaces graph_write_all_kmers_direct / HASH_ITERATE[_SORTED] + graph_write_kmer
 * (src/graph/graph_writer.c:116-127,182-193): builds, on the device, the .ctx v6 record
 * stream (W x u64 key, ncols x u32 covg, ncols x u8 edges per k-mer), in ascending key order
 * if sorted != 0.  The host writes the header (it owns GraphInfo) and streams the records. */
int mcx_graph_export_begin(mcx_graph *g, int sorted, uint64_t *nrecords, uint32_t *record_bytes);
int mcx_graph_export_read(mcx_graph *g, uint64_t first_record, uint64_t nrecords, void *host_dst);
int mcx_graph_export_end(mcx_graph *g);

/* replaces the body of `mccortex sort` (src/commands/ctx_sort.c:117-155: read every record, qsort pointers by
 * key, write them back): nrecords packed .ctx records (W x u64 key, ncols x u32 covg, ncols x u8 edges) in
 * host memory are ordered by ascending key on the GPU; records_out may be records_in. */
int mcx_sort_records(int device, uint32_t kmer_size, uint32_t ncols, const void *records_in, uint64_t nrecords,
                     void *records_out);

/* ---- multi-GPU pieces (one graph shard per GPU, one process per GPU) ------------------------
 * Ownership: owner(key) = (c * nparts) >> 32 with c = Lookup3(key) (the hash the reference uses
 * for its bucket choice, src/graph/hash_table.c:259).  A tuple is
 *     key  : W x u64 canonical key words            (W = 1 for k <= 31, 2 for k <= 63)
 *     meta : u32 = (count << 8) | edge mask         (count occurrences of the key, edges to OR)
 * Bins: keys_out holds nparts bins of cap_per_part tuples (x W u64), meta_out nparts x
 * cap_per_part u32, counts_out nparts u64 (zeroed by the call, filled on the stream).  All
 * pointers are device memory of g's GPU; batches must be MCX_MEM_DEVICE / MCX_LAYOUT_LINES.
 * A bin overflow is reported by the next mcx_graph_sync as MCX_ERR_TABLE_FULL.
 *
 * Sharded build of one step (what bench.py --gpus N drives):
 *   per batch : mcx_graph_add_reads_sharded  -> exchange bins -> mcx_graph_insert_tuples
 *   at the end: mcx_graph_flush_sharded      -> exchange bins -> mcx_graph_insert_tuples -> mcx_graph_sync
 * Hot k-mers are counted in the LOCAL front table whoever owns them and cross the wire once,
 * aggregated, at the flush; only what the front table could not absorb travels per occurrence. */
int mcx_graph_add_reads_sharded(mcx_graph *g, const mcx_read_batch *batch, uint32_t nparts, uint32_t my_part,
                                uint64_t cap_per_part, uint64_t *keys_out, uint32_t *meta_out, uint64_t *counts_out);
int mcx_graph_flush_sharded(mcx_graph *g, uint32_t nparts, uint32_t my_part, uint64_t cap_per_part,
                            uint64_t *keys_out, uint32_t *meta_out, uint64_t *counts_out);
/* Kernel C: insert n received tuples (device pointers) into this shard. */
int mcx_graph_insert_tuples(mcx_graph *g, const uint64_t *keys, const uint32_t *meta, uint64_t n, uint32_t colour);
/* same, with the tuple count in DEVICE memory (min(*n_dev, n_max) tuples are inserted): the count a
 * sender wrote travels on the stream, the host never waits for it */
int mcx_graph_insert_tuples_n(mcx_graph *g, const uint64_t *keys, const uint32_t *meta, const uint64_t *n_dev,
                              uint64_t n_max, uint32_t colour);
/* same, on a CUDA stream of the caller's (cudaStream_t as void*; NULL = the graph's stream): inserting
 * what arrived for batch b can then overlap the sharded kernel of batch b+1 (table updates commute;
 * the caller orders the stream against the counter exchange and the ring reuse, and joins it before
 * mcx_graph_sync) */
int mcx_graph_insert_tuples_on(mcx_graph *g, void *cuda_stream, const uint64_t *keys, const uint32_t *meta,
                               const uint64_t *n_dev, uint64_t n_max, uint32_t colour);

/* Routed variants: ONE kernel does the compute step and the all-to-all.  keys_dst[d] / meta_dst[d]
 * (host arrays of nparts device pointers; entry my_part is ignored) are where tuples for shard d
 * are appended: normally memory of GPU d mapped into this process (mcx_ipc_open), so the kernel's
 * stores cross NVLink while it is still hashing the rest of the batch, and no copy or collective
 * follows -- only the nparts counters (counts_out, local) are exchanged afterwards.  Each sender
 * owns its own region on the destination (cap_per_part tuples), so there are no remote atomics. */
int mcx_graph_add_reads_routed(mcx_graph *g, const mcx_read_batch *batch, uint32_t nparts, uint32_t my_part,
                               uint64_t cap_per_part, uint64_t *const *keys_dst, uint32_t *const *meta_dst,
                               uint64_t *counts_out);
int mcx_graph_flush_routed(mcx_graph *g, uint32_t nparts, uint32_t my_part, uint64_t cap_per_part,
                           uint64_t *const *keys_dst, uint32_t *const *meta_dst, uint64_t *counts_out);
/* device buffers a peer process can map: cudaMalloc'ed (zeroed) memory, its 64-byte CUDA IPC handle,
 * and the peer side (opens with lazy peer access: NVLink P2P between the GPUs of one box) */
int mcx_device_alloc(int device, size_t bytes, void **dptr);
int mcx_device_free(int device, void *dptr);
int mcx_ipc_export(const void *dptr, unsigned char handle[64]);
int mcx_ipc_open(int device, const unsigned char handle[64], void **dptr);
int mcx_ipc_close(int device, void *dptr);
/* ---- the sharded build from ONE process (the C driver's `build -D 0,1,.. --shard`) ----------------------------
 * One table shard per device; what ctx_build.c:389-425 does with one shared-memory table -- build_graph() over all
 * inputs, hash_table_print_stats, graph_writer_save_mkhdr -- over P device shards: host batches are dealt to the
 * devices piece by piece, every device runs the sharded kernel (tuples for keys it does not own are stored straight
 * into the owner's ring over NVLink: cudaDeviceEnablePeerAccess), owners insert what arrives, the dump is the P-way
 * merge of the shards' sorted runs.  capacity = total k-mer slots (what cmd_get_kmers_in_hash returns); batches are
 * MCX_MEM_HOST / MCX_LAYOUT_LINES without quality / homopolymer cut-off.  One host thread per shard set. */
typedef struct mcx_shardset mcx_shardset;
int mcx_shardset_create(uint32_t kmer_size, uint32_t ncols, uint64_t capacity, const int *devices, uint32_t ndevices, mcx_shardset **out);
int mcx_shardset_destroy(mcx_shardset *s);
int mcx_shardset_add_reads(mcx_shardset *s, const mcx_read_batch *batch);  /* asynchronous; pinned buffers must stay valid until sync */
int mcx_shardset_sync(mcx_shardset *s, mcx_load_stats *stats);
int mcx_shardset_stats(mcx_shardset *s, uint64_t *nkmers, uint64_t *capacity);
int mcx_shardset_export_begin(mcx_shardset *s, int sorted, uint64_t *nrecords, uint32_t *record_bytes);
/* the next (at most max_records) records of the dump, in file order, into host_dst; *got == 0 at the end */
int mcx_shardset_export_next(mcx_shardset *s, void *host_dst, uint64_t max_records, uint64_t *got);
int mcx_shardset_export_end(mcx_shardset *s);

/* Kernel B alone: reads -> one tuple per occurrence (count 1), binned by owner; nothing is
 * inserted locally.  Kept as the unaggregated baseline of the exchange and for tests. */
int mcx_kmer_tuples(mcx_graph *g, const mcx_read_batch *batch, uint32_t nparts, uint64_t cap_per_part,
                    uint64_t *keys_out, uint32_t *meta_out, uint64_t *counts_out);
/* owner of a key, for tests: same function the kernels use */
uint32_t mcx_key_owner(const uint64_t *key_words, uint32_t kmer_size, uint32_t nparts);

#ifdef __cplusplus
}
#endif
#endif /* MCX_GPU_H_ */
