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
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <zlib.h>

/* Where the graph comes from when a loader is called with g == NULL: the driver creates the device table on a
 * second thread (CUDA start-up + context take 0.5-1.5 s, as long as parsing a GB of FASTA) and the loader parses
 * ahead meanwhile, keeping up to MCX_RUNAHEAD_BYTES of finished batches until wait() hands the graph over. */
McxGraphSource mcx_graph_source = {NULL, NULL, NULL};

/* Several files of one colour loaded at the same time (the reference runs one reader thread per file of a
 * build_graph() call, src/basic/async_read_io.c): the library wants one host thread per graph at a time, so
 * every submission takes mcx_ingest.lock; the loaders do not sync (the counters are the graph's, not a file's) --
 * the caller syncs once after the last file and gets the totals of the whole call. */
McxIngestShared mcx_ingest = {false, PTHREAD_MUTEX_INITIALIZER, 0, NULL, NULL, NULL};
int mcx_submit_reads(mcx_graph *g, const mcx_read_batch *b)
{
  if(!mcx_ingest.concurrent) return mcx_ingest.submit ? mcx_ingest.submit(g, b) : mcx_graph_add_reads(mcx_ingest.route ? mcx_ingest.route(g) : g, b);
  pthread_mutex_lock(&mcx_ingest.lock);
  int r = mcx_ingest.submit ? mcx_ingest.submit(g, b) : mcx_graph_add_reads(mcx_ingest.route ? mcx_ingest.route(g) : g, b);
  pthread_mutex_unlock(&mcx_ingest.lock);
  return r;
}
int mcx_sync_reads(mcx_graph *g, mcx_load_stats *st) { return mcx_ingest.sync ? mcx_ingest.sync(g, st) : mcx_graph_sync(g, st); }
#define MCX_RUNAHEAD_BYTES (2048ull << 20)

#define MCX_BATCH_BYTES_DEFAULT (96u << 20)
/* bytes per batch; MCX_BATCH_BYTES=<n> in the environment overrides it (tests use tiny batches) */
static size_t batch_bytes_v;
static void batch_bytes_init(void)
{
  const char *e = getenv("MCX_BATCH_BYTES");
  batch_bytes_v = e && atol(e) > 0 ? (size_t)atol(e) : MCX_BATCH_BYTES_DEFAULT;
}
static size_t batch_bytes(void)
{
  static pthread_once_t once = PTHREAD_ONCE_INIT; /* several files may be loading at once */
  pthread_once(&once, batch_bytes_init);
  return batch_bytes_v;
}
#define MCX_BATCH_BYTES batch_bytes()
#define MCX_IN_BYTES (4u << 20)

/* The stream is read (and, for .gz, inflated: ~0.3 GB/s, the slowest stage of a gzip'd build) by a thread of its own
 * into two buffers that the parser and the reader swap: inflate and parse overlap instead of alternating. */
struct Loader;
struct McxSeqFile {
  char *path;
  gzFile gz;
  unsigned char *in; size_t in_len, in_pos; bool eof;   /* the buffer the parser is consuming */
  size_t in_raw;                                        /* bytes the reader put there (in_len shrinks when Q9 deletes) */
  unsigned char *buf[2]; int nread[2]; int full[2];     /* full[i]: buf[i] holds nread[i] bytes (<= 0: end / error) for the parser */
  int cur;                                              /* index of the parser's buffer, -1 before the first refill */
  /* quirk Q9 (see q9_apply): lines to delete from the stream at q9_at */
  uint64_t pd_off;                                      /* offset of in[0] in the stream AFTER deletions (what the parser has been given so far) */
  uint64_t q9_at;                                       /* where (same coordinates) the pending deletions start */
  uint32_t q9_lines; bool q9_mid;                       /* lines still to delete; the deletion continues at in[0] of the next buffer */
  /* how the next record is read (see next_record) */
  bool unknown, lookahead; uint64_t la_bases;
  int (*rec_fn)(McxSeqFile *, struct Loader *);
  bool started, stop;
  pthread_t thread; pthread_mutex_t mu; pthread_cond_t cv;
};

static void *seq_reader_main(void *arg)
{
  McxSeqFile *sf = arg;
  for(int i = 0;; i ^= 1) {
    pthread_mutex_lock(&sf->mu);
    while(sf->full[i] && !sf->stop) pthread_cond_wait(&sf->cv, &sf->mu);
    const bool stop = sf->stop;
    pthread_mutex_unlock(&sf->mu);
    if(stop) return NULL;
    const int n = gzread(sf->gz, sf->buf[i], MCX_IN_BYTES);
    pthread_mutex_lock(&sf->mu);
    sf->nread[i] = n; sf->full[i] = 1;
    pthread_cond_broadcast(&sf->cv);
    pthread_mutex_unlock(&sf->mu);
    if(n <= 0) return NULL;
  }
}

McxSeqFile *mcx_seq_open(const char *path)
{
  McxSeqFile *sf = calloc(1, sizeof(*sf));
  sf->path = strdup(path);
  sf->gz = strcmp(path, "-") == 0 ? gzdopen(0, "r") : gzopen(path, "r");
  if(!sf->gz) { free(sf->path); free(sf); return NULL; }
  gzbuffer(sf->gz, 1u << 20);
  sf->cur = -1; sf->unknown = true; sf->lookahead = true;
  pthread_mutex_init(&sf->mu, NULL); pthread_cond_init(&sf->cv, NULL);
  return sf;
}

void mcx_seq_close(McxSeqFile *sf)
{
  if(!sf) return;
  if(sf->started) {
    pthread_mutex_lock(&sf->mu); sf->stop = true; pthread_cond_broadcast(&sf->cv); pthread_mutex_unlock(&sf->mu);
    pthread_join(sf->thread, NULL);
  }
  if(sf->gz) gzclose(sf->gz);
  pthread_mutex_destroy(&sf->mu); pthread_cond_destroy(&sf->cv);
  free(sf->buf[0]); free(sf->buf[1]); free(sf->path); free(sf);
}

const char *mcx_seq_path(const McxSeqFile *sf) { return sf->path; }

int64_t mcx_seq_file_size(const McxSeqFile *sf)
{
  struct stat st;
  if(strcmp(sf->path, "-") == 0 || stat(sf->path, &st) != 0) return -1;
  return (int64_t)st.st_size;
}

/* ---- byte stream with getc / ungetc / "rest of line" ------------------------- */
/* Quirk Q9 [probed against the compiled reference, found by tests/test_host_fuzz.py].  The reference reads a record
 * through _read_unknown (libs/seq_file/seq_file.h:311-323) -- white space in front of the record is skipped, the first
 * other byte picks the format -- not only for the first record: seq_get_qual_limits buffers reads worth 1000 bases with
 * it before anything is parsed (seq_file.h:359-377,636-660; src/basic/seq_reader.c:436), and with an explicit
 * --fq-offset nothing ever replaces it (seq_file.h:97,337,437).  In that function a white-space byte other than '\n' is
 * meant to skip the rest of its line, but the buffered readers are instantiated with the UNBUFFERED skipline
 * (seq_file.h:426-427), which reads from the file behind the 1 MB stream buffer (DEFAULT_BUFSIZE, :129,:580).  So the
 * current line is NOT skipped (only the byte is), and the stream loses, per such byte, the bytes from the end of the
 * buffered megabyte through the next '\n' -- nothing, if the file ends inside that megabyte.  Every fill reads exactly
 * 2^20 bytes of what is left of the stream, so in the coordinates of the stream after deletions the place is the next
 * multiple of 2^20 behind the byte just read.  Deleted here in place, in the part of the buffer the parser has not
 * seen yet, or when the buffer that holds the place arrives. */
static void q9_apply(McxSeqFile *sf)
{
  size_t at;
  if(sf->q9_mid) at = 0;
  else if(sf->pd_off + sf->in_len <= sf->q9_at) return;         /* the place is not in this buffer yet */
  else at = sf->q9_at <= sf->pd_off ? 0 : (size_t)(sf->q9_at - sf->pd_off);
  if(at < sf->in_pos) at = sf->in_pos;
  while((sf->q9_lines || sf->q9_mid) && at < sf->in_len) {
    if(!sf->q9_mid) { sf->q9_lines--; sf->q9_mid = true; }      /* start deleting one more line */
    unsigned char *nl = memchr(sf->in + at, '\n', sf->in_len - at);
    const size_t end = nl ? (size_t)(nl - sf->in) + 1 : sf->in_len;
    memmove(sf->in + at, sf->in + end, sf->in_len - end);
    sf->in_len -= end - at;
    if(nl) sf->q9_mid = false;
  }
}
/* the white space in front of the first record (seq_file.h:316), with quirk Q9 */
static int sniff_skip_space(McxSeqFile *sf);

static bool refill(McxSeqFile *sf)
{
  if(sf->eof) return false;
  if(!sf->started) {
    sf->buf[0] = malloc(MCX_IN_BYTES); sf->buf[1] = malloc(MCX_IN_BYTES);
    if(!sf->buf[0] || !sf->buf[1]) mcx_die("Out of memory");
    if(pthread_create(&sf->thread, NULL, seq_reader_main, sf) != 0) mcx_die("Cannot start a thread");
    sf->started = true;
  }
  pthread_mutex_lock(&sf->mu);
  if(sf->cur >= 0) { sf->full[sf->cur] = 0; pthread_cond_broadcast(&sf->cv); } /* consumed: the reader may refill it */
  sf->cur = sf->cur < 0 ? 0 : sf->cur ^ 1;
  while(!sf->full[sf->cur]) pthread_cond_wait(&sf->cv, &sf->mu);
  const int n = sf->nread[sf->cur];
  pthread_mutex_unlock(&sf->mu);
  if(n <= 0) { sf->eof = true; sf->in_len = sf->in_pos = 0; return false; }
  sf->pd_off += sf->in_len;
  sf->in = sf->buf[sf->cur]; sf->in_len = sf->in_raw = (size_t)n; sf->in_pos = 0;
  if(sf->q9_lines || sf->q9_mid) q9_apply(sf);
  if(sf->in_len == 0) return refill(sf); /* (everything of this buffer was deleted) */
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
static int sniff_skip_space(McxSeqFile *sf)
{
  int c;
  while((c = sgetc(sf)) != -1 && is_space(c))
    if(c != '\n') { /* not the rest of this line: a line at the end of the reference's stream buffer (Q9) */
      const uint64_t p = sf->pd_off + sf->in_pos;   /* bytes given to the parser so far, this one included */
      if(!sf->q9_lines && !sf->q9_mid) sf->q9_at = (((p - 1) >> 20) + 1) << 20;
      sf->q9_lines++;
      q9_apply(sf);
    }
  return c;
}

/* ---- loader state ---------------------------------------------------------------- */
typedef McxQStat QStat;
typedef struct { mcx_read_batch b; char *seq, *qual; } PendingBatch; /* parsed before the graph existed */
typedef struct Loader {
  mcx_graph *g; const McxLoadPrefs *prefs; mcx_load_stats *stats;
  size_t last_seqlen;  /* length of the record just read (look-ahead accounting of next_record) */
  PendingBatch *pend; size_t npend, pend_cap; uint64_t pend_bytes;
  Buf lines;          /* LINES batch under construction */
  Buf qlines;         /* quality bytes parallel to `lines` (only when a quality cut-off is set) */
  Buf qual;           /* quality string of the current FASTQ record */
  uint64_t nreads_total;
  /* FASTQ offset auto-detection (seq_file.h:636-682): min/max of the first <= 1000 quals */
  QStat qs[2]; int cur;   /* per input file (two with --seq2 under --remove-pcr) */
  bool any_qual, offset_known; uint8_t fq_offset;
  int err;
  /* --remove-pcr: one entry per read of the batch under construction */
  bool pcr, want_names;
  Buf name;           /* name line of the record just parsed (want_names) */
  uint64_t *read_off; uint8_t *mate; size_t nreads, reads_cap;
  uint8_t next_flip;  /* MCX_MATE_REVCOMP if the read being parsed will be reverse-complemented on the device */
} Loader;

/* FASTQ ASCII offset from the quality range of the first reads, exactly the decision list of
 * seq_guess_fastq_format (libs/seq_file/seq_file.h:666-682) + FASTQ_OFFSET (:127) */
uint8_t mcx_guess_fq_offset(const McxQStat *L)
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

/* the call for the batch under construction; false: nothing to send */
static bool prepare_batch(Loader *L, mcx_read_batch *b)
{
  if(L->lines.len == 0 || L->err) { L->lines.len = 0; L->qlines.len = 0; return false; }
  memset(b, 0, sizeof(*b));
  b->seq = L->lines.b; b->nbytes = L->lines.len;
  b->layout = MCX_LAYOUT_LINES; b->mem = MCX_MEM_HOST;
  b->colour = L->prefs->colour; b->hp_cutoff = L->prefs->hp_cutoff; b->must_exist = L->prefs->must_exist;
  if(L->prefs->fq_cutoff && L->any_qual) {
    /* build_graph.c:202-207: the ASCII offset is added only when a cut-off is set */
    if(!L->offset_known) {
      L->fq_offset = L->prefs->fq_offset ? L->prefs->fq_offset : mcx_guess_fq_offset(&L->qs[0]);
      L->offset_known = true;
      if(L->fq_offset + L->prefs->fq_cutoff >= 127) { L->err = MCX_ERR_UNSUPPORTED; L->lines.len = L->qlines.len = 0; return false; }
    }
    b->qual = L->qlines.b;
    b->fq_cutoff = (uint8_t)(L->prefs->fq_cutoff + L->fq_offset);
  }
  return true;
}

/* block until the graph exists, then send, in order, the batches parsed meanwhile */
static void loader_need_graph(Loader *L)
{
  if(L->g || !mcx_graph_source.wait) return; /* (g == NULL without a source: the CPU-only test harness) */
  L->g = mcx_graph_source.wait(mcx_graph_source.ctx);
  for(size_t i = 0; i < L->npend; i++) {
    if(!L->err) { int r = mcx_submit_reads(L->g, &L->pend[i].b); if(r != MCX_OK) L->err = r; }
    free(L->pend[i].seq); free(L->pend[i].qual);
  }
  free(L->pend); L->pend = NULL; L->npend = L->pend_cap = 0; L->pend_bytes = 0;
}

static void flush_batch(Loader *L)
{
  mcx_read_batch b;
  if(!prepare_batch(L, &b)) return;
  if(!L->g && mcx_graph_source.wait && mcx_graph_source.ready && !mcx_graph_source.ready(mcx_graph_source.ctx) &&
     L->pend_bytes + b.nbytes < MCX_RUNAHEAD_BYTES / (uint64_t)(mcx_ingest.nfiles > 1 ? mcx_ingest.nfiles : 1)) {
    /* the device is still starting up: keep the batch, go on parsing into fresh buffers */
    if(L->npend == L->pend_cap) {
      L->pend_cap = L->pend_cap ? 2 * L->pend_cap : 16;
      L->pend = realloc(L->pend, L->pend_cap * sizeof(*L->pend));
      if(!L->pend) mcx_die("Out of memory");
    }
    PendingBatch *pb = &L->pend[L->npend++];
    pb->b = b; pb->seq = L->lines.b; pb->qual = L->qlines.b;
    L->pend_bytes += b.nbytes;
    memset(&L->lines, 0, sizeof(L->lines)); memset(&L->qlines, 0, sizeof(L->qlines));
    buf_reserve(&L->lines, MCX_BATCH_BYTES + (1u << 20));
    return;
  }
  loader_need_graph(L);
  if(!L->err) { int r = mcx_submit_reads(L->g, &b); if(r != MCX_OK) L->err = r; }
  L->lines.len = 0; L->qlines.len = 0;
}

/* a record's sequence now sits at lines[start .. len): terminate it, maybe ship the batch */
static void end_read(Loader *L, size_t start)
{
  size_t seqlen = L->lines.len - start;
  L->last_seqlen = seqlen;
  QStat *qs = &L->qs[L->cur];
  if(L->qual.len && qs->bcount < 1000) {
    size_t lim = 1000 - qs->qcount, n = L->qual.len < lim ? L->qual.len : lim, i;
    for(i = 0; i < n; i++) {
      int q = (signed char)L->qual.b[i];
      if(q > qs->qmax) qs->qmax = q;
      if(q < qs->qmin) qs->qmin = q;
    }
    qs->bcount += seqlen; qs->qcount += L->qual.len;
  } else if(qs->bcount < 1000) qs->bcount += seqlen;
  if(L->prefs->fq_cutoff) {
    /* quality bytes parallel to the sequence; 0x7F where the read has none (not filtered there) */
    size_t have = L->qual.len < seqlen ? L->qual.len : seqlen;
    buf_reserve(&L->qlines, seqlen + 1);
    memcpy(L->qlines.b + L->qlines.len, L->qual.b, have);
    memset(L->qlines.b + L->qlines.len + have, 0x7F, seqlen - have + 1);
    /* a read that gets reverse-complemented has a too short quality string padded to the read's length
     * first (seq_file.h:726-733,763; '.' is what that code means to pad with) */
    if(L->next_flip && have > 0 && have < seqlen) memset(L->qlines.b + L->qlines.len + have, '.', seqlen - have);
    L->qlines.len += seqlen + 1;
    if(L->qual.len) L->any_qual = true;
  }
  buf_push(&L->lines, '\n');
  L->nreads_total++;
  if(L->pcr) {
    if(L->nreads + 2 > L->reads_cap) {
      L->reads_cap = L->reads_cap ? L->reads_cap * 2 : 1u << 16;
      L->read_off = realloc(L->read_off, L->reads_cap * sizeof(uint64_t));
      L->mate = realloc(L->mate, L->reads_cap);
      if(!L->read_off || !L->mate) mcx_die("Out of memory");
    }
    L->read_off[L->nreads] = start; L->mate[L->nreads] = L->next_flip; L->nreads++;
    return; /* the pair logic decides when a batch is complete */
  }
  if(L->lines.len >= MCX_BATCH_BYTES) flush_batch(L);
}

/* libs/seq_file/seq_file.h:274-295 */
static int read_fasta(McxSeqFile *sf, Loader *L)
{
  int c = sgetc(sf);
  if(c == -1) return 0;
  L->name.len = 0;
  if(c != '>' || sreadline(sf, L->want_names ? &L->name : NULL) == 0) return -1;
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
  L->name.len = 0;
  if(c != '@' || sreadline(sf, L->want_names ? &L->name : NULL) == 0) return -1;
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
  L->qual.len = 0; L->name.len = 0;
  buf_push(&L->lines, (char)c);
  sreadline(sf, &L->lines);
  chomp_from(&L->lines, start);
  end_read(L, start);
  return 1;
}

/* One record, the way the reference gets it (quirk Q9 above): while the file is in `unknown` mode -- the reads worth
 * the first 1000 bases, or every read when --fq-offset was given -- white space in front of the record is skipped byte
 * by byte and the record's own first byte picks the reader; afterwards the reader of the last such record serves the
 * rest of the file.  Returns what the readers return: 1, 0 at the end, -1 on a truncated record. */
static int next_record(McxSeqFile *sf, Loader *L)
{
  if(sf->unknown) {
    const int c = sniff_skip_space(sf);
    if(c == -1) return 0;
    sf->in_pos--; /* ungetc: the byte came from the current buffer */
    sf->rec_fn = c == '@' ? read_fastq : (c == '>' ? read_fasta : read_plain);
  }
  const int s = sf->rec_fn(sf, L);
  if(s > 0 && sf->unknown && sf->lookahead) {
    sf->la_bases += L->last_seqlen;
    if(sf->la_bases >= 1000) sf->unknown = false; /* _seq_buffer_reads stops here; the format reader takes over */
  }
  return s;
}

void mcx_add_load_stats(mcx_load_stats *stats, const mcx_load_stats *st)
{
  stats->total_bases_read += st->total_bases_read;
  stats->total_bases_loaded += st->total_bases_loaded;
  stats->contigs_parsed += st->contigs_parsed;
  stats->num_kmers_loaded += st->num_kmers_loaded;
  stats->num_kmers_novel += st->num_kmers_novel;
  stats->num_se_reads += st->num_se_reads;
  if(st->num_good_reads != UINT64_MAX && stats->num_good_reads != UINT64_MAX) { stats->num_good_reads += st->num_good_reads; stats->num_bad_reads += st->num_bad_reads; }
  else { stats->num_good_reads = stats->num_bad_reads = UINT64_MAX; }
}

int mcx_load_seq_file(mcx_graph *g, McxSeqFile *sf, const McxLoadPrefs *prefs, mcx_load_stats *stats)
{
  /* big uncompressed files: several threads (seq_ingest_par.c); a FASTQ file that stops being strictly
   * four lines per record comes back with `resume` set: the rest of it is read here */
  McxParResume pr;
  { int rc = 0; if(mcx_load_seq_file_par(g, sf, prefs, stats, &rc, &pr)) return rc; }
  Loader L; memset(&L, 0, sizeof(L));
  L.g = g; L.prefs = prefs; L.stats = stats; L.qs[0].qmin = L.qs[1].qmin = 0x7fffffff;
  buf_reserve(&L.lines, MCX_BATCH_BYTES + (1u << 20));
  buf_reserve(&L.qual, 1u << 16);

  if(!pr.resume) mcx_status("[seq] Parsing sequence file %s", sf->path);

  int s = 0;
  sf->lookahead = prefs->fq_offset == 0; /* seq_reader.c:436: the look-ahead only runs when the offset is to be guessed */
  if(pr.resume) {
    /* the multi-threaded reader has taken the strict FASTQ records up to pr.offset (>= 1000 bases: format known) */
    if(pr.g) L.g = pr.g;
    L.qs[0] = pr.qs; L.any_qual = pr.any_qual; L.offset_known = pr.offset_known; L.fq_offset = pr.fq_offset;
    L.nreads_total = pr.nreads;
    if(gzseek(sf->gz, (z_off_t)pr.offset, SEEK_SET) < 0) mcx_die("Cannot seek in %s", sf->path);
    sf->in_len = sf->in_pos = 0; sf->eof = false;
    if(sf->lookahead) { sf->unknown = false; sf->rec_fn = read_fastq; }
  }
  while((s = next_record(sf, &L)) > 0 && !L.err) {}
  flush_batch(&L);
  if(s < 0 && !L.err) mcx_warn("Input error: %s\n", sf->path);
  mcx_phase("  parsed");
  loader_need_graph(&L);
  mcx_phase("  submitted");

  int r = L.err;
  if(!mcx_ingest.concurrent) {
    mcx_load_stats st;
    r = mcx_sync_reads(L.g, &st);
    if(L.err) r = L.err;
    mcx_add_load_stats(stats, &st);
  }
  char n1[64]; mcx_ulong_to_str(L.nreads_total, n1);
  mcx_status("[seq] Loaded %s reads and 0 reads pairs (file: %s)", n1, sf->path);
  free(L.lines.b); free(L.qual.b); free(L.qlines.b);
  return r;
}

/* ---- --remove-pcr -------------------------------------------------------------------------
 * The reads of one task go to the device together with their offsets and a mate byte each
 * (mcx_graph_add_reads_pcr); which reads form a pair is decided here exactly like the reference's
 * parse loops: --seq every read single (seq_reader.c:421-462), --seq2 read i of file 1 with read i of
 * file 2 until either ends (:357-419), --seqi consecutive reads whose names match (:289-355,
 * seq_file.h:783-800).  Re-orienting the mates (seq_reader.c:506-510) and the duplicate test itself
 * run on the GPU. */
static bool name_end(unsigned char c) { return !c || is_space(c); }
static int names_cmp(const char *aa, const char *bb)
{
  const unsigned char *a = (const unsigned char *)aa, *b = (const unsigned char *)bb, *a0 = a, *b0 = b;
  while(*a && *b && *a == *b && !is_space(*a)) { a++; b++; }
  if(a > a0 && b > b0 && a[-1] == '/' && b[-1] == '/' && ((*a == '1' && *b == '2') || (*a == '2' && *b == '1')) &&
     name_end(a[1]) && name_end(b[1])) return 0;
  return name_end(*a) && name_end(*b) ? 0 : (int)*a - (int)*b;
}

/* ship the first n reads of the batch; later reads (at most one: a read still waiting for its mate) move to the front */
static void flush_batch_pcr(Loader *L, size_t n)
{
  if(n == 0) return;
  loader_need_graph(L);
  size_t cut = n < L->nreads ? L->read_off[n] : L->lines.len, i;
  if(!L->err) {
    mcx_read_batch b; memset(&b, 0, sizeof(b));
    b.seq = L->lines.b; b.nbytes = cut;
    b.layout = MCX_LAYOUT_LINES; b.mem = MCX_MEM_HOST;
    b.colour = L->prefs->colour; b.hp_cutoff = L->prefs->hp_cutoff;
    if(L->prefs->fq_cutoff && L->any_qual) {
      if(!L->offset_known) {
        uint8_t o1 = L->prefs->fq_offset ? L->prefs->fq_offset : mcx_guess_fq_offset(&L->qs[0]), o2 = o1;
        if(!L->prefs->fq_offset && L->qs[1].bcount) o2 = mcx_guess_fq_offset(&L->qs[1]);
        /* one threshold per batch: files of one pair with different ASCII offsets are not supported */
        if(o1 != o2 && o1 && o2) mcx_die("Paired files with different FASTQ offsets (%u, %u) are not supported", o1, o2);
        L->fq_offset = o1 ? o1 : o2;
        L->offset_known = true;
        if(L->fq_offset + L->prefs->fq_cutoff >= 127) L->err = MCX_ERR_UNSUPPORTED;
      }
      b.qual = L->qlines.b;
      b.fq_cutoff = (uint8_t)(L->prefs->fq_cutoff + L->fq_offset);
    }
    if(!L->err) {
      L->read_off[n] = cut; /* read_off[nreads] = nbytes for the call (already so when a read stays behind) */
      int r = mcx_graph_add_reads_pcr(L->g, &b, L->read_off, L->mate, n);
      if(r != MCX_OK) L->err = r;
    }
  }
  memmove(L->lines.b, L->lines.b + cut, L->lines.len - cut);
  L->lines.len -= cut;
  if(L->prefs->fq_cutoff) { memmove(L->qlines.b, L->qlines.b + cut, L->qlines.len - cut); L->qlines.len -= cut; }
  for(i = n; i < L->nreads; i++) { L->read_off[i - n] = L->read_off[i] - cut; L->mate[i - n] = L->mate[i]; }
  L->nreads -= n;
}


int mcx_load_seq_pcr(mcx_graph *g, McxSeqFile *sf1, McxSeqFile *sf2, bool interleaved, const McxLoadPrefs *prefs,
                     mcx_load_stats *stats)
{
  Loader L; memset(&L, 0, sizeof(L));
  L.g = g; L.prefs = prefs; L.stats = stats; L.qs[0].qmin = L.qs[1].qmin = 0x7fffffff;
  L.pcr = true; L.want_names = interleaved;
  buf_reserve(&L.lines, MCX_BATCH_BYTES + (1u << 20));
  buf_reserve(&L.qual, 1u << 16);
  buf_reserve(&L.name, 1u << 10);
  const uint8_t flip1 = (prefs->matedir & 2) ? MCX_MATE_REVCOMP : 0, flip2 = (prefs->matedir & 1) ? MCX_MATE_REVCOMP : 0;
  uint64_t num_se = 0, num_pairs = 0;
  int s1 = 0, s2 = 0;

  if(sf2) mcx_status("[seq] Parsing sequence files %s %s\n", sf1->path, sf2->path);
  else if(interleaved) mcx_status("[seq] Reading a (possibly) interleaved file (expect both S.E. & P.E. reads)");
  else mcx_status("[seq] Parsing sequence file %s", sf1->path);

  sf1->lookahead = prefs->fq_offset == 0;
  if(sf2) sf2->lookahead = prefs->fq_offset == 0;
  if(sf2) {
    for(;;) {
      const size_t n_before = L.nreads, len_before = L.lines.len, qlen_before = L.qlines.len;
      const uint64_t total_before = L.nreads_total;
      L.cur = 0; L.next_flip = flip1;
      s1 = next_record(sf1, &L);
      L.cur = 1; L.next_flip = flip2;
      s2 = next_record(sf2, &L);
      if(s1 < 0) mcx_warn("input error: %s", sf1->path);
      if(s2 < 0) mcx_warn("input error: %s", sf2->path);
      if((s1 > 0) != (s2 > 0) && !(s1 < 0 || s2 < 0)) mcx_warn("Different number of reads in pe files [%s; %s]\n", sf1->path, sf2->path);
      if(s1 <= 0 || s2 <= 0) {
        /* drop a read that has no mate (the reference never hands it to the graph) */
        L.nreads = n_before; L.lines.len = len_before; L.qlines.len = qlen_before; L.nreads_total = total_before;
        break;
      }
      L.mate[L.nreads - 2] |= MCX_MATE_FIRST; L.mate[L.nreads - 1] |= MCX_MATE_SECOND;
      num_pairs++;
      if(L.lines.len >= MCX_BATCH_BYTES) flush_batch_pcr(&L, L.nreads);
      if(L.err) break;
    }
  } else if(interleaved) {
    Buf prev_name = {0, 0, 0};
    bool pending = false; /* the batch's last read is waiting to see whether the next one is its mate */
    buf_reserve(&prev_name, 1u << 10);
    L.cur = 0;
    for(;;) {
      L.next_flip = pending ? flip2 : flip1;
      size_t before = L.nreads;
      s1 = next_record(sf1, &L);
      if(s1 <= 0) break;
      buf_push(&L.name, '\0'); L.name.len--;
      if(pending && names_cmp(prev_name.b, L.name.b) == 0) {
        L.mate[before - 1] |= MCX_MATE_FIRST; L.mate[before] |= MCX_MATE_SECOND;
        num_pairs++; pending = false;
      } else {
        if(pending) num_se++;
        if(pending && flip1 != flip2) {
          /* parsed as a possible second mate, it is a first one: its quality padding rule and flip flag follow read 1 */
          L.mate[before] = flip1;
        }
        prev_name.len = 0; buf_reserve(&prev_name, L.name.len + 1);
        memcpy(prev_name.b, L.name.b, L.name.len + 1); prev_name.len = L.name.len;
        pending = true;
      }
      if(L.lines.len >= MCX_BATCH_BYTES) flush_batch_pcr(&L, L.nreads - (pending ? 1 : 0));
      if(L.err) break;
    }
    if(pending) num_se++;
    if(s1 < 0) mcx_warn("Input error: %s\n", sf1->path);
    free(prev_name.b);
  } else {
    L.cur = 0; L.next_flip = flip1;
    while((s1 = next_record(sf1, &L)) > 0 && !L.err) {
      num_se++;
      if(L.lines.len >= MCX_BATCH_BYTES) flush_batch_pcr(&L, L.nreads);
    }
    if(s1 < 0) mcx_warn("Input error: %s\n", sf1->path);
  }
  flush_batch_pcr(&L, L.nreads);
  loader_need_graph(&L);

  mcx_load_stats st;
  int r = mcx_graph_sync(L.g, &st);
  if(L.err) r = L.err;
  stats->total_bases_read += st.total_bases_read;
  stats->total_bases_loaded += st.total_bases_loaded;
  stats->contigs_parsed += st.contigs_parsed;
  stats->num_kmers_loaded += st.num_kmers_loaded;
  stats->num_kmers_novel += st.num_kmers_novel;
  stats->num_se_reads += num_se;
  stats->num_pe_reads += 2 * num_pairs;
  stats->num_dup_se_reads += st.num_dup_se_reads;
  stats->num_dup_pe_pairs += st.num_dup_pe_pairs;
  stats->num_good_reads = stats->num_bad_reads = UINT64_MAX;

  char n1[64], n2[64]; mcx_ulong_to_str(num_se, n1); mcx_ulong_to_str(num_pairs, n2);
  if(sf2) mcx_status("[seq] Loaded %s read pairs (files: %s, %s)", n2, sf1->path, sf2->path);
  else mcx_status("[seq] Loaded %s reads and %s reads pairs (file: %s)", n1, n2, sf1->path);
  free(L.lines.b); free(L.qual.b); free(L.qlines.b); free(L.name.b); free(L.read_off); free(L.mate);
  return r;
}
