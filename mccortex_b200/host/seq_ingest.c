/* seq_ingest.c -- FASTA / FASTQ / plain (optionally gzip) -> LINES batches -> the GPU.
 *
 * Replaces, for the build path: the record readers of libs/seq_file/seq_file.h:245-323
 * (format sniffed from the first non-space byte; multi-line FASTA/FASTQ records are
 * concatenated, '\r' and '\n' stripped), the single-end parse loop of
 * src/basic/seq_reader.c:421-462, and the reader->queue->worker hand-off of
 * src/basic/async_read_io.c (msg-pool of 2048 reads): here one host thread inflates and
 * parses straight into a LINES buffer (read + '\n'), and ships it through
 * mcx_graph_add_reads() every MCX_BATCH_BYTES; the library overlaps H2D and kernels.
 *
 * SAM/BAM/CRAM input is out of scope (needs htslib).
 */
#include "mcx_host.h"
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <zlib.h>

#define MCX_BATCH_BYTES (96u << 20)
#define MCX_IN_BYTES (4u << 20)

struct McxSeqFile {
  char *path;
  gzFile gz;
  unsigned char *in; size_t in_len, in_pos; bool eof;
};

McxSeqFile *mcx_seq_open(const char *path)
{
  McxSeqFile *sf = calloc(1, sizeof(*sf));
  sf->path = strdup(path);
  sf->gz = strcmp(path, "-") == 0 ? gzdopen(0, "r") : gzopen(path, "r");
  if(!sf->gz) { free(sf->path); free(sf); return NULL; }
  gzbuffer(sf->gz, 1u << 20);
  sf->in = malloc(MCX_IN_BYTES);
  return sf;
}

void mcx_seq_close(McxSeqFile *sf)
{
  if(!sf) return;
  if(sf->gz) gzclose(sf->gz);
  free(sf->in); free(sf->path); free(sf);
}

const char *mcx_seq_path(const McxSeqFile *sf) { return sf->path; }

int64_t mcx_seq_file_size(const McxSeqFile *sf)
{
  struct stat st;
  if(strcmp(sf->path, "-") == 0 || stat(sf->path, &st) != 0) return -1;
  return (int64_t)st.st_size;
}

/* ---- byte stream with getc / ungetc / "rest of line" ------------------------- */
static bool refill(McxSeqFile *sf)
{
  if(sf->eof) return false;
  int n = gzread(sf->gz, sf->in, MCX_IN_BYTES);
  if(n <= 0) { sf->eof = true; sf->in_len = sf->in_pos = 0; return false; }
  sf->in_len = (size_t)n; sf->in_pos = 0;
  return true;
}
static inline int sgetc(McxSeqFile *sf)
{
  if(sf->in_pos >= sf->in_len && !refill(sf)) return -1;
  return sf->in[sf->in_pos++];
}
static inline int speek(McxSeqFile *sf)
{
  if(sf->in_pos >= sf->in_len && !refill(sf)) return -1;
  return sf->in[sf->in_pos];
}

typedef struct { char *b; size_t len, cap; } Buf;
static inline void buf_reserve(Buf *b, size_t extra)
{
  if(b->len + extra + 1 > b->cap) {
    b->cap = (b->len + extra + 1) * 2;
    b->b = realloc(b->b, b->cap);
    if(!b->b) mcx_die("Out of memory");
  }
}
static inline void buf_push(Buf *b, char c) { buf_reserve(b, 1); b->b[b->len++] = c; }

/* append the rest of the current line (without its '\n') to dst (dst may be NULL = skip);
 * returns the number of bytes consumed including the newline (0 at EOF) */
static size_t sreadline(McxSeqFile *sf, Buf *dst)
{
  size_t total = 0;
  for(;;) {
    if(sf->in_pos >= sf->in_len && !refill(sf)) return total;
    unsigned char *s = sf->in + sf->in_pos, *e = memchr(s, '\n', sf->in_len - sf->in_pos);
    size_t n = e ? (size_t)(e - s) : sf->in_len - sf->in_pos;
    if(dst) { buf_reserve(dst, n); memcpy(dst->b + dst->len, s, n); dst->len += n; }
    total += n; sf->in_pos += n;
    if(e) { sf->in_pos++; return total + 1; }
  }
}
/* the reference chomps '\r' and '\n' off the end of what it has accumulated */
static inline void chomp_from(Buf *b, size_t floor_len)
{
  while(b->len > floor_len && (b->b[b->len - 1] == '\n' || b->b[b->len - 1] == '\r')) b->len--;
}
static inline bool is_space(int c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

/* ---- loader state ---------------------------------------------------------------- */
typedef struct {
  mcx_graph *g; const McxLoadPrefs *prefs; mcx_load_stats *stats;
  Buf lines;          /* LINES batch under construction */
  Buf qlines;         /* quality bytes parallel to `lines` (only when a quality cut-off is set) */
  Buf qual;           /* quality string of the current FASTQ record */
  uint64_t nreads_total;
  /* FASTQ offset auto-detection (seq_file.h:636-682): min/max of the first <= 1000 quals */
  int qmin, qmax; size_t qcount, bcount; bool saw_qual;
  bool any_qual, offset_known; uint8_t fq_offset;
  int err;
} Loader;

/* FASTQ ASCII offset from the quality range of the first reads, exactly the decision list of
 * seq_guess_fastq_format (libs/seq_file/seq_file.h:666-682) + FASTQ_OFFSET (:127) */
static uint8_t guess_fq_offset(const Loader *L)
{
  static const int OFFS[6] = {33, 33, 64, 64, 64, 33};
  int fmt, mn = L->qmin, mx = L->qmax;
  if(L->qcount == 0) return 0; /* no qualities seen: offset stays 0 (seq_reader.c:436-441) */
  if(mn >= 33 && mx <= 73) fmt = 1;
  else if(mn >= 33 && mx <= 75) fmt = 5;
  else if(mn >= 67 && mx <= 105) fmt = 4;
  else if(mn >= 64 && mx <= 105) fmt = 3;
  else if(mn >= 59 && mx <= 105) fmt = 2;
  else fmt = 0;
  return (uint8_t)OFFS[fmt];
}

static void flush_batch(Loader *L)
{
  if(L->lines.len == 0 || L->err) { L->lines.len = 0; L->qlines.len = 0; return; }
  mcx_read_batch b; memset(&b, 0, sizeof(b));
  b.seq = L->lines.b; b.nbytes = L->lines.len;
  b.layout = MCX_LAYOUT_LINES; b.mem = MCX_MEM_HOST;
  b.colour = L->prefs->colour; b.hp_cutoff = L->prefs->hp_cutoff; b.must_exist = L->prefs->must_exist;
  if(L->prefs->fq_cutoff && L->any_qual) {
    /* build_graph.c:202-207: the ASCII offset is added only when a cut-off is set */
    if(!L->offset_known) {
      L->fq_offset = L->prefs->fq_offset ? L->prefs->fq_offset : guess_fq_offset(L);
      L->offset_known = true;
      if(L->fq_offset + L->prefs->fq_cutoff >= 127) { L->err = MCX_ERR_UNSUPPORTED; L->lines.len = L->qlines.len = 0; return; }
    }
    b.qual = L->qlines.b;
    b.fq_cutoff = (uint8_t)(L->prefs->fq_cutoff + L->fq_offset);
  }
  L->qlines.len = 0;
  int r = mcx_graph_add_reads(L->g, &b);
  if(r != MCX_OK) L->err = r;
  L->lines.len = 0;
}

/* a record's sequence now sits at lines[start .. len): terminate it, maybe ship the batch */
static void end_read(Loader *L, size_t start)
{
  size_t seqlen = L->lines.len - start;
  if(L->qual.len && L->bcount < 1000) {
    size_t lim = 1000 - L->qcount, n = L->qual.len < lim ? L->qual.len : lim, i;
    for(i = 0; i < n; i++) {
      int q = (signed char)L->qual.b[i];
      if(q > L->qmax) L->qmax = q;
      if(q < L->qmin) L->qmin = q;
    }
    L->bcount += seqlen; L->qcount += L->qual.len; L->saw_qual = true;
  } else if(L->bcount < 1000) L->bcount += seqlen;
  if(L->prefs->fq_cutoff) {
    /* quality bytes parallel to the sequence; 0x7F where the read has none (not filtered there) */
    size_t have = L->qual.len < seqlen ? L->qual.len : seqlen;
    buf_reserve(&L->qlines, seqlen + 1);
    memcpy(L->qlines.b + L->qlines.len, L->qual.b, have);
    memset(L->qlines.b + L->qlines.len + have, 0x7F, seqlen - have + 1);
    L->qlines.len += seqlen + 1;
    if(L->qual.len) L->any_qual = true;
  }
  buf_push(&L->lines, '\n');
  L->nreads_total++;
  if(L->lines.len >= MCX_BATCH_BYTES) flush_batch(L);
}

/* libs/seq_file/seq_file.h:274-295 */
static int read_fasta(McxSeqFile *sf, Loader *L)
{
  int c = sgetc(sf);
  if(c == -1) return 0;
  if(c != '>' || sreadline(sf, NULL) == 0) return -1;
  size_t start = L->lines.len;
  L->qual.len = 0;
  while((c = speek(sf)) != '>') {
    if(c == -1) break;
    sf->in_pos++;
    if(c != '\r' && c != '\n') {
      buf_push(&L->lines, (char)c);
      size_t nread = sreadline(sf, &L->lines);
      chomp_from(&L->lines, start);
      if(nread == 0) break;
    }
  }
  end_read(L, start);
  return 1;
}

/* libs/seq_file/seq_file.h:245-272 */
static int read_fastq(McxSeqFile *sf, Loader *L)
{
  int c = sgetc(sf);
  if(c == -1) return 0;
  if(c != '@' || sreadline(sf, NULL) == 0) return -1;
  size_t start = L->lines.len;
  L->qual.len = 0;
  while((c = sgetc(sf)) != '+') {
    if(c == -1) { L->lines.len = start; return -1; }
    if(c != '\r' && c != '\n') {
      buf_push(&L->lines, (char)c);
      if(sreadline(sf, &L->lines) == 0) { L->lines.len = start; return -1; }
      chomp_from(&L->lines, start);
    }
  }
  while((c = sgetc(sf)) != -1 && c != '\n') {}
  if(c == -1) { L->lines.len = start; return -1; }
  size_t seqlen = L->lines.len - start;
  bool eof_in_qual = false;
  do {
    if(sreadline(sf, &L->qual) > 0) chomp_from(&L->qual, 0);
    else { eof_in_qual = true; break; }
  } while(L->qual.len < seqlen);
  if(!eof_in_qual) {
    while((c = speek(sf)) != -1 && c != '@') sf->in_pos++;
  }
  end_read(L, start);
  return 1;
}

/* libs/seq_file/seq_file.h:298-309 */
static int read_plain(McxSeqFile *sf, Loader *L)
{
  int c;
  while((c = sgetc(sf)) != -1 && is_space(c)) if(c != '\n') sreadline(sf, NULL);
  if(c == -1) return 0;
  size_t start = L->lines.len;
  L->qual.len = 0;
  buf_push(&L->lines, (char)c);
  sreadline(sf, &L->lines);
  chomp_from(&L->lines, start);
  end_read(L, start);
  return 1;
}

int mcx_load_seq_file(mcx_graph *g, McxSeqFile *sf, const McxLoadPrefs *prefs, mcx_load_stats *stats)
{
  Loader L; memset(&L, 0, sizeof(L));
  L.g = g; L.prefs = prefs; L.stats = stats; L.qmin = 0x7fffffff; L.qmax = 0;
  buf_reserve(&L.lines, MCX_BATCH_BYTES + (1u << 20));
  buf_reserve(&L.qual, 1u << 16);

  mcx_status("[seq] Parsing sequence file %s", sf->path);

  /* format sniff, seq_file.h:311-323 */
  int c, s = 0;
  int (*reader)(McxSeqFile *, Loader *) = NULL;
  while((c = sgetc(sf)) != -1 && is_space(c)) if(c != '\n') sreadline(sf, NULL);
  if(c != -1) {
    reader = c == '@' ? read_fastq : (c == '>' ? read_fasta : read_plain);
    sf->in_pos--; /* ungetc: the byte came from the current buffer */
    while((s = reader(sf, &L)) > 0 && !L.err) {}
  }
  flush_batch(&L);
  if(s < 0 && !L.err) mcx_warn("Input error: %s\n", sf->path);

  mcx_load_stats st;
  int r = mcx_graph_sync(g, &st);
  if(L.err) r = L.err;
  stats->total_bases_read += st.total_bases_read;
  stats->total_bases_loaded += st.total_bases_loaded;
  stats->contigs_parsed += st.contigs_parsed;
  stats->num_kmers_loaded += st.num_kmers_loaded;
  stats->num_kmers_novel += st.num_kmers_novel;
  stats->num_se_reads += st.num_se_reads;
  if(st.num_good_reads != UINT64_MAX) { stats->num_good_reads += st.num_good_reads; stats->num_bad_reads += st.num_bad_reads; }
  else { stats->num_good_reads = stats->num_bad_reads = UINT64_MAX; }

  char n1[64]; mcx_ulong_to_str(L.nreads_total, n1);
  mcx_status("[seq] Loaded %s reads and 0 reads pairs (file: %s)", n1, sf->path);
  free(L.lines.b); free(L.qual.b); free(L.qlines.b);
  return r;
}
