/* tests/emul/abi_shim.c -- TEST INFRASTRUCTURE, not product code, never shipped or linked into the product.
 *
 * The subset of include/mcx_gpu.h that the host driver (mccortex_b200/host/) calls, implemented on top of the
 * oracle (oracle/mcx_oracle.c).  tests/conftest.py links the driver's C sources against this file instead of
 * libmcxgpu.so into tests/emul/hostcheck, so that the CPU-only test run (-m "not gpu") can exercise everything ABOVE
 * the C ABI -- argument parsing, table sizing, FASTA / FASTQ / plain / gzip ingest (sequential, multi-threaded,
 * concurrent files), read pairing for --remove-pcr, the .ctx reader and colour filters, header arithmetic, the
 * pipelined writer, `sort` -- against the golden files and the compiled reference.  It says nothing about the
 * CUDA path; that is what the -m gpu tests do with the real library.
 */
#include "../../include/mcx_gpu.h"
#include <stdlib.h>
#include <string.h>

/* ---- the oracle's C interface (oracle/mcx_oracle.c) */
typedef struct OrcGraph OrcGraph;
typedef struct {
  uint64_t total_bases_read, total_bases_loaded, contigs_parsed;
  uint64_t num_kmers_loaded, num_kmers_novel;
  uint64_t num_se_reads, num_pe_reads, num_good_reads, num_bad_reads;
  uint64_t num_dup_se_reads, num_dup_pe_pairs;
} OrcStats;
OrcGraph *orc_graph_new(size_t k, size_t ncols, uint64_t capacity);
void orc_graph_free(OrcGraph *g);
uint64_t orc_graph_nkmers(const OrcGraph *g);
int orc_graph_is_full(const OrcGraph *g);
void orc_graph_add_read(OrcGraph *g, const char *seq, size_t seqlen, const char *qual, size_t quallen, size_t colour,
                        uint8_t fq_cutoff, uint8_t fq_offset, uint8_t hp_cutoff, OrcStats *st);
void orc_graph_wipe_readstrt(OrcGraph *g);
void orc_graph_add_reads_pcr(OrcGraph *g, const char *seq1, size_t sl1, const char *qual1, size_t ql1, const char *seq2,
                             size_t sl2, const char *qual2, size_t ql2, size_t colour, uint8_t fq_cutoff,
                             uint8_t fq_offset1, uint8_t fq_offset2, uint8_t hp_cutoff, int matedir, OrcStats *st);
uint64_t orc_graph_load_records(OrcGraph *g, const uint8_t *recs, uint64_t n, uint32_t file_ncols, const uint32_t *from,
                                const uint32_t *into, uint32_t nmap, int flags, uint64_t *novel_out);
void orc_graph_set_intersect(OrcGraph *g, int must_exist_reads);
void orc_graph_finish_intersect(OrcGraph *g);
size_t orc_graph_write_header(const OrcGraph *g, uint8_t *buf);
size_t orc_graph_dump_sorted(const OrcGraph *g, uint8_t *buf);

struct mcx_graph {
  OrcGraph *o; OrcStats st;
  uint32_t k, ncols, flags; uint64_t capacity;
  uint64_t novel_from_files;
  uint8_t *exp; size_t exp_hdr; uint64_t exp_n; uint32_t rec_bytes;
};

int mcx_device_count(void) { return 1; }
const char *mcx_last_error(void) { return "(host-check shim)"; }
int mcx_host_alloc(void **ptr, size_t bytes) { *ptr = malloc(bytes ? bytes : 1); return *ptr ? MCX_OK : MCX_ERR_NOMEM; }
int mcx_host_free(void *ptr) { free(ptr); return MCX_OK; }

int mcx_graph_create(uint32_t k, uint32_t ncols, uint64_t capacity, int device, uint32_t flags, mcx_graph **out)
{
  if(!out || k < 3 || k > 63 || !(k & 1u) || ncols == 0 || capacity == 0 || device < 0 || device >= 16) return MCX_ERR_BAD_ARG; /* (pretends to have 16 devices) */
  mcx_graph *g = calloc(1, sizeof(*g));
  g->o = orc_graph_new(k, ncols, capacity);
  g->k = k; g->ncols = ncols; g->flags = flags; g->capacity = capacity;
  if(flags & MCX_GRAPH_INTERSECT) orc_graph_set_intersect(g->o, 0);
  if(flags & MCX_GRAPH_READSTRT) orc_graph_wipe_readstrt(g->o);
  *out = g;
  return MCX_OK;
}
int mcx_graph_destroy(mcx_graph *g) { if(g) { orc_graph_free(g->o); free(g->exp); free(g); } return MCX_OK; }
int mcx_graph_prepare_host(mcx_graph *g) { return g ? MCX_OK : MCX_ERR_BAD_ARG; }

int mcx_graph_add_reads(mcx_graph *g, const mcx_read_batch *b)
{
  if(!g || !b || b->colour >= g->ncols || b->layout != MCX_LAYOUT_LINES || b->mem != MCX_MEM_HOST) return MCX_ERR_BAD_ARG;
  if(b->hp_cutoff == 1 || b->hp_cutoff > g->k || b->fq_cutoff >= 127) return MCX_ERR_UNSUPPORTED;
  if(b->nbytes && b->seq[b->nbytes - 1] != '\n') return MCX_ERR_BAD_ARG;
  if(g->flags & MCX_GRAPH_INTERSECT) orc_graph_set_intersect(g->o, b->must_exist);
  else if(b->must_exist) return MCX_ERR_BAD_ARG;
  const uint8_t cut = b->qual ? b->fq_cutoff : 0;
  for(uint64_t p = 0; p < b->nbytes;) {
    const char *nl = memchr(b->seq + p, '\n', b->nbytes - p);
    const size_t len = (size_t)(nl - (b->seq + p));
    /* the batch's cut-off already includes the ASCII offset; 0x7F padding is never below it */
    orc_graph_add_read(g->o, b->seq + p, len, cut ? b->qual + p : NULL, cut ? len : 0, b->colour, cut, 0, b->hp_cutoff, &g->st);
    p += len + 1;
  }
  return MCX_OK;
}

int mcx_graph_add_reads_pcr(mcx_graph *g, const mcx_read_batch *b, const uint64_t *off, const uint8_t *mate, uint64_t n)
{
  if(!g || !b || b->colour >= g->ncols || b->layout != MCX_LAYOUT_LINES || b->mem != MCX_MEM_HOST) return MCX_ERR_BAD_ARG;
  if(!(g->flags & MCX_GRAPH_READSTRT) || b->must_exist) return MCX_ERR_BAD_ARG;
  if(n == 0 || b->nbytes == 0) return MCX_OK;
  if(off[0] != 0 || off[n] != b->nbytes) return MCX_ERR_BAD_ARG;
  const uint8_t cut = b->qual ? b->fq_cutoff : 0;
  OrcStats before = g->st;
  for(uint64_t r = 0; r < n; r++) {
    const size_t l1 = (size_t)(off[r + 1] - off[r] - 1);
    const char *s1 = b->seq + off[r], *q1 = cut ? b->qual + off[r] : NULL;
    if((mate[r] & 3u) == MCX_MATE_FIRST) {
      if(r + 1 >= n || (mate[r + 1] & 3u) != MCX_MATE_SECOND) return MCX_ERR_BAD_ARG;
      const size_t l2 = (size_t)(off[r + 2] - off[r + 1] - 1);
      const int matedir = ((mate[r] & MCX_MATE_REVCOMP) ? 2 : 0) | ((mate[r + 1] & MCX_MATE_REVCOMP) ? 1 : 0);
      orc_graph_add_reads_pcr(g->o, s1, l1, q1, cut ? l1 : 0, b->seq + off[r + 1], l2, cut ? b->qual + off[r + 1] : NULL,
                              cut ? l2 : 0, b->colour, cut, 0, 0, b->hp_cutoff, matedir, &g->st);
      r++;
    } else if((mate[r] & 3u) == MCX_MATE_SINGLE) {
      orc_graph_add_reads_pcr(g->o, s1, l1, q1, cut ? l1 : 0, NULL, 0, NULL, 0, b->colour, cut, 0, 0, b->hp_cutoff,
                              (mate[r] & MCX_MATE_REVCOMP) ? 2 : 0, &g->st);
    } else return MCX_ERR_BAD_ARG;
  }
  /* the library counts reads from the terminators; the driver overrides SE / PE itself (mcx_load_seq_pcr) */
  g->st.num_se_reads = before.num_se_reads; g->st.num_pe_reads = before.num_pe_reads;
  return MCX_OK;
}
int mcx_graph_pcr_reset(mcx_graph *g)
{
  if(!g || !(g->flags & MCX_GRAPH_READSTRT)) return MCX_ERR_BAD_ARG;
  orc_graph_wipe_readstrt(g->o);
  return MCX_OK;
}

int mcx_graph_sync(mcx_graph *g, mcx_load_stats *s)
{
  if(!g) return MCX_ERR_BAD_ARG;
  if(s) {
    memset(s, 0, sizeof(*s));
    s->total_bases_read = g->st.total_bases_read; s->total_bases_loaded = g->st.total_bases_loaded;
    s->contigs_parsed = g->st.contigs_parsed; s->num_kmers_loaded = g->st.num_kmers_loaded;
    s->num_kmers_novel = g->st.num_kmers_novel + g->novel_from_files; s->num_se_reads = g->st.num_se_reads;
    s->num_good_reads = s->num_bad_reads = UINT64_MAX;
    s->num_dup_se_reads = g->st.num_dup_se_reads; s->num_dup_pe_pairs = g->st.num_dup_pe_pairs;
  }
  memset(&g->st, 0, sizeof(g->st)); g->novel_from_files = 0;
  return orc_graph_is_full(g->o) ? MCX_ERR_TABLE_FULL : MCX_OK;
}
int mcx_graph_stats(mcx_graph *g, uint64_t *nkmers, uint64_t *capacity)
{
  if(!g) return MCX_ERR_BAD_ARG;
  if(nkmers) *nkmers = orc_graph_nkmers(g->o);
  if(capacity) *capacity = g->capacity;
  return MCX_OK;
}

int mcx_graph_load_records(mcx_graph *g, const void *records, uint64_t n, uint32_t file_ncols, uint32_t mem,
                           const uint32_t *from_col, const uint32_t *into_col, uint32_t nmap, uint32_t flags,
                           uint64_t *nkmers_loaded, uint64_t *nkmers_novel)
{
  if(!g || mem != MCX_MEM_HOST || (n && !records)) return MCX_ERR_BAD_ARG;
  uint64_t novel = 0, loaded = orc_graph_load_records(g->o, records, n, file_ncols, from_col, into_col, nmap, (int)flags, &novel);
  g->novel_from_files += novel;
  if(nkmers_loaded) *nkmers_loaded = loaded;
  if(nkmers_novel) *nkmers_novel = novel;
  return orc_graph_is_full(g->o) ? MCX_ERR_TABLE_FULL : MCX_OK;
}
int mcx_graph_finish_intersect(mcx_graph *g, uint64_t *nkmers)
{
  if(!g || !(g->flags & MCX_GRAPH_INTERSECT)) return MCX_ERR_BAD_ARG;
  orc_graph_finish_intersect(g->o);
  if(nkmers) *nkmers = orc_graph_nkmers(g->o);
  return MCX_OK;
}

/* an unsorted dump has no defined order: the shim hands it out in DESCENDING key order, so that tests which sort it
 * afterwards have something to do */
int mcx_graph_export_begin(mcx_graph *g, int sorted, uint64_t *nrecords, uint32_t *record_bytes)
{
  if(!g) return MCX_ERR_BAD_ARG;
  free(g->exp);
  const size_t total = orc_graph_dump_sorted(g->o, NULL);
  g->exp = malloc(total + 1);
  orc_graph_dump_sorted(g->o, g->exp);
  g->exp_hdr = orc_graph_write_header(g->o, NULL);
  g->rec_bytes = 8u * ((g->k + 31u) / 32u) + 5u * g->ncols;
  g->exp_n = (total - g->exp_hdr) / g->rec_bytes;
  if(!sorted) {
    uint8_t tmp[8 * 2 + 5 * 4096], *r = g->exp + g->exp_hdr;
    for(uint64_t i = 0, j = g->exp_n; j && i < j - 1; i++, j--) {
      memcpy(tmp, r + i * g->rec_bytes, g->rec_bytes);
      memcpy(r + i * g->rec_bytes, r + (j - 1) * g->rec_bytes, g->rec_bytes);
      memcpy(r + (j - 1) * g->rec_bytes, tmp, g->rec_bytes);
    }
  }
  if(nrecords) *nrecords = g->exp_n;
  if(record_bytes) *record_bytes = g->rec_bytes;
  return MCX_OK;
}
int mcx_graph_export_read(mcx_graph *g, uint64_t first, uint64_t n, void *dst)
{
  if(!g || !g->exp || first + n > g->exp_n) return MCX_ERR_BAD_ARG;
  memcpy(dst, g->exp + g->exp_hdr + first * g->rec_bytes, n * g->rec_bytes);
  return MCX_OK;
}
int mcx_graph_export_end(mcx_graph *g) { if(g) { free(g->exp); g->exp = NULL; } return MCX_OK; }

static uint32_t sort_W, sort_rb;
static int cmp_rec(const void *a, const void *b)
{
  uint64_t x, y;
  for(uint32_t w = 0; w < sort_W; w++) {
    memcpy(&x, (const char *)a + 8 * w, 8); memcpy(&y, (const char *)b + 8 * w, 8);
    if(x != y) return x < y ? -1 : 1;
  }
  return 0;
}
int mcx_sort_records(int device, uint32_t k, uint32_t ncols, const void *in, uint64_t n, void *out)
{
  if(device != 0) return MCX_ERR_BAD_ARG;
  sort_W = (k + 31u) / 32u; sort_rb = 8u * sort_W + 5u * ncols;
  if(out != in) memmove(out, in, (size_t)n * sort_rb);
  qsort(out, n, sort_rb, cmp_rec);
  return MCX_OK;
}

/* the shard set: a sharded build gives the graph one table gives, so the stand-in is ONE oracle graph behind the same
 * entry points (what the host driver's --shard mode needs on a box without GPUs: arguments, refusals, the dump loop) */
struct mcx_shardset { mcx_graph *g; uint64_t at, n; uint32_t rb; };
int mcx_shardset_create(uint32_t k, uint32_t ncols, uint64_t capacity, const int *devices, uint32_t ndevices, mcx_shardset **out)
{
  if(!out || !devices || ndevices < 2 || ndevices > 16) return MCX_ERR_BAD_ARG;
  mcx_shardset *s = calloc(1, sizeof(*s));
  if(!s) return MCX_ERR_NOMEM;
  int r = mcx_graph_create(k, ncols, capacity + capacity / 10 + 1024 * ndevices, 0, 0, &s->g);
  if(r) { free(s); return r; }
  *out = s;
  return MCX_OK;
}
int mcx_shardset_destroy(mcx_shardset *s) { if(s) { mcx_graph_destroy(s->g); free(s); } return MCX_OK; }
int mcx_shardset_add_reads(mcx_shardset *s, const mcx_read_batch *b)
{
  if(!s || !b) return MCX_ERR_BAD_ARG;
  if(b->must_exist) return MCX_ERR_UNSUPPORTED;
  return mcx_graph_add_reads(s->g, b);
}
int mcx_shardset_sync(mcx_shardset *s, mcx_load_stats *st) { mcx_load_stats tmp; return s ? mcx_graph_sync(s->g, st ? st : &tmp) : MCX_ERR_BAD_ARG; }
int mcx_shardset_stats(mcx_shardset *s, uint64_t *nkmers, uint64_t *capacity) { return s ? mcx_graph_stats(s->g, nkmers, capacity) : MCX_ERR_BAD_ARG; }
int mcx_shardset_export_begin(mcx_shardset *s, int sorted, uint64_t *nrecords, uint32_t *record_bytes)
{
  if(!s) return MCX_ERR_BAD_ARG;
  int r = mcx_graph_export_begin(s->g, sorted, &s->n, &s->rb);
  s->at = 0;
  if(nrecords) *nrecords = s->n;
  if(record_bytes) *record_bytes = s->rb;
  return r;
}
int mcx_shardset_export_next(mcx_shardset *s, void *dst, uint64_t max_records, uint64_t *got)
{
  if(!s || !got) return MCX_ERR_BAD_ARG;
  uint64_t n = s->n - s->at < max_records ? s->n - s->at : max_records;
  int r = n ? mcx_graph_export_read(s->g, s->at, n, dst) : MCX_OK;
  s->at += n; *got = n;
  return r;
}
int mcx_shardset_export_end(mcx_shardset *s) { return s ? mcx_graph_export_end(s->g) : MCX_ERR_BAD_ARG; }
