/*
 * mcx_host.h -- C host side of `mccortex-b200 build` (everything above the C ABI).
 *
 * Mirrors, for the build path only, the reference modules named on each
 * declaration (paths relative to the reference root).  Plain C99.
 */
#ifndef MCX_HOST_H_
#define MCX_HOST_H_

#include <stdint.h>
#include <stddef.h>
#include <stdbool.h>
#include <stdio.h>
#include "mcx_gpu.h"

/* ---- output conventions: src/global/ctx_output.{c,h} ---------------------- */
extern FILE *mcx_msg_out;                 /* NULL when -q/--quiet */
void mcx_status(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
void mcx_warn(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
void mcx_phase(const char *what); /* MCX_TIMING=1: phase wall clock on stderr */
void mcx_die(const char *fmt, ...) __attribute__((noreturn)) __attribute__((format(printf, 1, 2)));
void mcx_print_usage(const char *usage, const char *errfmt, ...) __attribute__((noreturn))
    __attribute__((format(printf, 2, 3)));
void mcx_ulong_to_str(uint64_t n, char *out);          /* 1,234,567  (util.c ulong_to_str) */
void mcx_bytes_to_str(uint64_t bytes, char *out);      /* 1.5GB      (util.c bytes_to_str, 1 decimal) */
bool mcx_mem_to_integer(const char *arg, size_t *bytes); /* src/global/util.c:206-222 */

/* ---- table sizing: src/basic/hash_mem.{c,h}, src/graph/cmd_mem.c ----------- */
#define MCX_IDEAL_OCCUPANCY 0.75f
#define MCX_WARN_OCCUPANCY  0.9f
#define MCX_MAX_BUCKET_SIZE 48
#define MCX_DEFAULT_MEM     (1UL << 29)
#define MCX_DEFAULT_NKMERS  (1UL << 22)
size_t mcx_hash_table_cap(uint64_t nkmers, uint64_t *num_bkts, uint8_t *bkt_size);
size_t mcx_hash_table_mem(uint64_t nkmers, size_t entrybits, uint64_t *nkmers_out);
size_t mcx_hash_table_mem_limit(size_t memlimit, size_t entrybits, uint64_t *nkmers_out);
size_t mcx_get_kmers_in_hash(size_t mem_to_use, bool mem_to_use_set, size_t num_kmers, bool num_kmers_set,
                             size_t entry_bits, int64_t min_num_kmer_req, int64_t max_num_kmers_req,
                             bool use_mem_limit, size_t *graph_mem_ptr);

/* ---- header metadata: src/basic/graph_info.{c,h}, src/graph/graph_writer.c -- */
typedef struct {              /* ErrorCleaning, src/basic/graph_info.h */
  uint8_t cleaned_tips, cleaned_unitigs, cleaned_kmers, is_graph_intersection; /* the bytes as stored in a header: merges OR them as they are, like the reference's `bool |= bool` on bytes that are not 0 / 1 */
  uint32_t clean_unitigs_thresh, clean_kmers_thresh;
  char *intersection_name;    /* malloc'd, "undefined" by default */
} McxCleaning;
typedef struct {
  uint32_t mean_read_length;
  uint64_t total_sequence;
  long double seq_err;
  char *sample_name;          /* malloc'd */
  McxCleaning cleaning;
} McxGInfo;
void mcx_ginfo_init(McxGInfo *g);
void mcx_ginfo_free(McxGInfo *g);
void mcx_ginfo_set_name(McxGInfo *g, const char *name);
void mcx_ginfo_update_contigs(McxGInfo *g, uint64_t added_seq, uint64_t num_contigs);
void mcx_ginfo_merge(McxGInfo *dst, const McxGInfo *src);
/* writes the .ctx v6 header for ncols colours after merging every colour into a fresh
 * header entry, exactly as graph_writer_mkhdr does; returns bytes written */
size_t mcx_write_ctx_header(FILE *fh, uint32_t kmer_size, uint32_t ncols, const McxGInfo *ginfo);
/* graph_write_header of the colours as they stand (no merge into a fresh header): `join` */
size_t mcx_write_ctx_header_as_is(FILE *fh, uint32_t kmer_size, uint32_t ncols, const McxGInfo *ginfo, uint32_t nbitfields /* 0: the minimum for k */);

/* ---- graph files: src/graph/graph_file_reader.{c,h}, src/basic/file_filter.{c,h}, src/basic/range.c,
 *      src/graph/graphs_load.c ------------------------------------------------------------- */
typedef struct {
  char *input, *path;           /* "[into:]path[:from]" as given, and the bare path */
  FILE *fh;
  uint32_t version, kmer_size, num_of_bitfields, num_of_cols;
  McxGInfo *ginfo;              /* num_of_cols entries */
  unsigned char (*seq_err_raw)[16]; /* the 16 bytes of each colour's long double as stored (padding included) */
  unsigned char (*clean_flags_raw)[4]; /* the four cleaning flag bytes of each colour as stored */
  size_t hdr_size;
  int64_t file_size, num_of_kmers;   /* -1 if unknown */
  uint32_t nfilter, *from_col, *into_col;   /* FileFilter, sorted by into */
  uint32_t into_ncols;          /* max into + 1 */
} McxCtxFile;
/* words per k-mer in the records: what the reference binary for this k is compiled with (sizeof(BinaryKmer)), whatever a
 * damaged header's bitfield count says (its checks, graph_file_reader.c:118-129, wrap at 2^32) */
#define MCX_CTX_W(f) (((f)->kmer_size + 31u) / 32u)
/* graph_file_open2(file, input, "r", true, into_offset): dies on a malformed file like the reference */
McxCtxFile *mcx_ctx_open(const char *input, size_t into_offset);
void mcx_ctx_close(McxCtxFile *f);
/* graph_load(): merge the header's colours into ginfo[] (graph_load_ginfo) and the records into g.
 * ginfo == NULL: the header metadata is not merged (intersection graphs: ctx_build.c:359-360 resets it).
 * load_flags: MCX_LOAD_* of mcx_gpu.h.  Returns 0 or an MCX_ERR_*. */
int mcx_ctx_load(mcx_graph *g, McxCtxFile *f, McxGInfo *ginfo, size_t graph_ncols, uint32_t load_flags,
                 uint64_t *nkmers_read, uint64_t *nkmers_loaded, uint64_t *nkmers_novel);
void mcx_ctx_flatten(McxCtxFile *f, uint32_t intocol);
void mcx_ctx_check_records(const McxCtxFile *f, const unsigned char *recs, size_t n); /* dies on an oversized k-mer */
/* graph_write_header(fh, &file->hdr) (src/graph/graph_writer.c:62-110): the header as parsed, not merged */
size_t mcx_ctx_write_header_raw(FILE *fh, const McxCtxFile *f);
bool mcx_ctx_filter_is_direct(const McxCtxFile *f);     /* file_filter_from_direct */

/* `join` command: src/commands/ctx_join.c (without --intersect) */
int mcx_cmd_join(int argc, char **argv);
/* `sort` command: src/commands/ctx_sort.c */
int mcx_cmd_sort(int argc, char **argv);   /* file_filter_flatten, src/basic/file_filter.c:218-226 */

/* ---- sequence input: libs/seq_file/seq_file.h, src/basic/seq_reader.c -------- */
typedef struct McxSeqFile McxSeqFile;
McxSeqFile *mcx_seq_open(const char *path);            /* gz or plain, "-" = stdin */
void mcx_seq_close(McxSeqFile *sf);
const char *mcx_seq_path(const McxSeqFile *sf);
int64_t mcx_seq_file_size(const McxSeqFile *sf);       /* -1 if unknown (stdin) */

typedef struct {
  uint8_t fq_cutoff, hp_cutoff, fq_offset;             /* fq_offset 0 = auto-detect */
  bool must_exist;                                     /* SeqLoadingPrefs.must_exist_in_graph (--intersect) */
  bool remove_pcr;                                     /* SeqLoadingPrefs.remove_pcr_dups */
  uint8_t matedir;                                     /* ReadMateDir: 0 FF, 1 FR (default), 2 RF, 3 RR */
  uint32_t colour;
} McxLoadPrefs;

/* The loaders below accept g == NULL when mcx_graph_source is set: they parse ahead while the graph is being
 * created elsewhere (ready() polls, wait() blocks and returns it) -- see seq_ingest.c */
typedef struct {
  mcx_graph *(*wait)(void *ctx);
  bool (*ready)(void *ctx);
  void *ctx;
} McxGraphSource;
extern McxGraphSource mcx_graph_source;

/* several files loaded at once (see seq_ingest.c): submissions are serialised, the caller syncs after the last file */
#include <pthread.h>
typedef struct {
  bool concurrent; pthread_mutex_t lock; int nfiles;
  /* several devices (build -D 0,1,...): every batch goes to the replica route() picks, sync() joins all of them */
  mcx_graph *(*route)(mcx_graph *g);
  int (*sync)(mcx_graph *g, mcx_load_stats *st);
  /* sharded build (build -D 0,1,... --shard): batches go to the shard set instead of a graph */
  int (*submit)(mcx_graph *g, const mcx_read_batch *b);
} McxIngestShared;
extern McxIngestShared mcx_ingest;
int mcx_submit_reads(mcx_graph *g, const mcx_read_batch *b);
int mcx_sync_reads(mcx_graph *g, mcx_load_stats *st);
void mcx_add_load_stats(mcx_load_stats *stats, const mcx_load_stats *st);

/* Parse every read of sf (FASTA / FASTQ / plain, sniffed from the first byte like
 * seq_file.h:311-323) and feed the graph in LINES batches.  Stats of this file are ADDED
 * to *stats.  Returns 0, or the MCX_ERR_* that stopped the load. */
int mcx_load_seq_file(mcx_graph *g, McxSeqFile *sf, const McxLoadPrefs *prefs, mcx_load_stats *stats);

/* seq_ingest_par.c: the same for a regular, uncompressed FASTA / FASTQ / one-read-per-line file of >= 32 MB, parsed by
 * several threads (MCX_PARSE_THREADS, default min(cores, 16); 1 = off).  Returns false when the file is not
 * eligible (then nothing was done); else *rc is what mcx_load_seq_file would have returned. */
typedef struct { int qmin, qmax; size_t qcount, bcount; } McxQStat; /* quality range of the first reads (seq_file.h:636-682) */
uint8_t mcx_guess_fq_offset(const McxQStat *qs);                    /* seq_guess_fastq_format + FASTQ_OFFSET */
typedef struct {
  bool resume; size_t offset;   /* FASTQ that stops being strict: the sequential reader continues at this file offset */
  mcx_graph *g;                 /* the graph, if it was obtained from mcx_graph_source meanwhile */
  McxQStat qs; bool any_qual, offset_known; uint8_t fq_offset;
  uint64_t nreads;
} McxParResume;
bool mcx_load_seq_file_par(mcx_graph *g, McxSeqFile *sf, const McxLoadPrefs *prefs, mcx_load_stats *stats, int *rc,
                           McxParResume *resume);

/* One task with --remove-pcr in force (build_graph_from_reads_mt with remove_pcr_dups): sf2 != NULL for a --seq2
 * pair, interleaved for --seqi (consecutive reads whose names match are a pair), else single-end.  Reads and
 * pairs go through mcx_graph_add_reads_pcr in file order. */
int mcx_load_seq_pcr(mcx_graph *g, McxSeqFile *sf1, McxSeqFile *sf2, bool interleaved, const McxLoadPrefs *prefs,
                     mcx_load_stats *stats);

#endif /* MCX_HOST_H_ */
