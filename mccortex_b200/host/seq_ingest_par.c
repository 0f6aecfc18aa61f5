/* seq_ingest_par.c -- uncompressed FASTA / plain files parsed by several host threads.
 *
 * The GPU side of `build` takes host batches at tens of GB/s (PCIe-bound); one thread running the record
 * reader of seq_ingest.c delivers ~2 GB/s.  A file that is a regular, uncompressed FASTA or one-read-per-line
 * file is therefore mapped and cut into segments at record starts; worker threads turn segments into LINES
 * buffers (same rules as read_fasta / read_plain of seq_ingest.c, i.e. libs/seq_file/seq_file.h:274-309) and the
 * calling thread hands the buffers to mcx_graph_add_reads in file order.  The reference reads one file per
 * thread (src/basic/async_read_io.c:118-141, one async_io_reader each); a build's table updates commute, so the
 * order in which reads arrive does not change the graph.
 *
 * Where a cut may be made:
 *   FASTA  at a '>' that follows a '\n'.  The sequential reader only ever looks for '>' at the start of a line,
 *          header lines are consumed whole, so such a byte is a record start whatever precedes it.
 *   plain  after any '\n' (every line is a record or is skipped on its own).
 *   FASTQ  candidates are '@' lines whose next-but-one line starts with '+', but '@' may also start a quality line
 *          and the sequential reader accepts multi-line records, so a candidate is only a guess.  It is made
 *          certain by induction: workers parse STRICT records only ('@' line, one sequence line, '+' line, one
 *          quality line at least as long as the sequence, next byte '@' or the end) -- on those the sequential reader
 *          (read_fastq, seq_file.h:245-272) does exactly the same -- and the calling thread accepts segment s only
 *          if every earlier segment ended on its cut with a complete record.  The first record that is not strict
 *          (or a segment that does not end on its cut) hands the REST OF THE FILE, from that record on, to the
 *          sequential reader.
 * gzip, stdin, --remove-pcr (order matters) and small files keep the sequential reader.
 */
#include "mcx_host.h"
#include <fcntl.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#define PAR_MIN_BYTES (32u << 20)          /* smaller files: not worth the threads */
#define PAR_SEG_BYTES_DEFAULT (16u << 20)  /* raw bytes per segment (MCX_PARSE_SEG_BYTES overrides; tests use tiny ones) */
#define PAR_MAX_THREADS 32

typedef struct {
  char *b; size_t len, cap;   /* LINES bytes of one segment */
  char *q;                    /* FASTQ with a quality cut-off: quality bytes parallel to b (cap bytes too) */
  uint64_t nreads;
  size_t dev_at;              /* FASTQ: file offset of the first record that is not strict, or SIZE_MAX */
  int state;                  /* 0 free, 1 being filled, 2 full */
  size_t seg;                 /* which segment it holds */
} SegBuf;

enum { FMT_FASTA, FMT_PLAIN, FMT_FASTQ };
typedef struct {
  const unsigned char *data; size_t size;
  int fmt; bool want_qual;
  size_t nseg; size_t *cut;   /* segment i = [cut[i], cut[i+1]) */
  size_t next_seg;            /* next segment to hand to a worker */
  SegBuf *bufs; size_t nbufs;
  pthread_mutex_t mu; pthread_cond_t cv;
} ParState;

static inline bool par_is_space(int c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

/* read_fasta of seq_ingest.c over memory [p, end): end is the end of the file or a record start */
static void parse_fasta_range(const unsigned char *d, size_t p, size_t end, SegBuf *o)
{
  while(p < end) {
    /* d[p] == '>' : header line, dropped */
    const unsigned char *e = memchr(d + p + 1, '\n', end - p - 1);
    if(!e) {
      if(end - p - 1 == 0) return; /* '>' then EOF: the sequential reader reports an input error and stops */
      p = end;
    } else p = (size_t)(e - d) + 1;
    const size_t start = o->len;
    while(p < end && d[p] != '>') {
      const unsigned char c = d[p++];
      if(c != '\r' && c != '\n') {
        const unsigned char *nl = memchr(d + p, '\n', end - p);
        const size_t n = nl ? (size_t)(nl - (d + p)) : end - p;
        o->b[o->len++] = (char)c;
        memcpy(o->b + o->len, d + p, n); o->len += n;
        p += n + (nl ? 1 : 0);
        while(o->len > start && (o->b[o->len - 1] == '\n' || o->b[o->len - 1] == '\r')) o->len--;
        if(!nl && n == 0) break;
      }
    }
    o->b[o->len++] = '\n';
    o->nreads++;
  }
}

/* read_plain of seq_ingest.c over memory [p, end): both are line starts (or the end of the file) */
static void parse_plain_range(const unsigned char *d, size_t p, size_t end, SegBuf *o)
{
  while(p < end) {
    const unsigned char c = d[p++];
    if(par_is_space(c)) {
      if(c != '\n') { const unsigned char *nl = memchr(d + p, '\n', end - p); p = nl ? (size_t)(nl - d) + 1 : end; }
      continue;
    }
    const size_t start = o->len;
    const unsigned char *nl = memchr(d + p, '\n', end - p);
    const size_t n = nl ? (size_t)(nl - (d + p)) : end - p;
    o->b[o->len++] = (char)c;
    memcpy(o->b + o->len, d + p, n); o->len += n;
    p += n + (nl ? 1 : 0);
    while(o->len > start && (o->b[o->len - 1] == '\n' || o->b[o->len - 1] == '\r')) o->len--;
    o->b[o->len++] = '\n';
    o->nreads++;
  }
}

/* one STRICT FASTQ record at *p (see the header of this file); end = the segment's cut (or the file's end, then
 * at_eof).  Returns false, leaving *p alone, if the record is not strict or does not end at or before `end`. */
typedef struct { const unsigned char *seq, *qual; size_t seqlen, quallen; } FqRec;
static bool fq_next(const unsigned char *d, size_t *pp, size_t end, bool at_eof, FqRec *r)
{
  size_t p = *pp;
  if(d[p] != '@') return false;
  const unsigned char *nl = memchr(d + p + 1, '\n', end - p - 1);
  if(!nl) return false;
  p = (size_t)(nl - d) + 1;
  if(p >= end || d[p] == '+' || d[p] == '\r' || d[p] == '\n') return false; /* empty line / no sequence: not strict */
  nl = memchr(d + p, '\n', end - p);
  if(!nl) return false;
  r->seq = d + p; r->seqlen = (size_t)(nl - (d + p));
  while(r->seqlen && r->seq[r->seqlen - 1] == '\r') r->seqlen--; /* (the first byte is not a CR: seqlen >= 1) */
  p = (size_t)(nl - d) + 1;
  if(p >= end || d[p] != '+') return false;
  nl = memchr(d + p, '\n', end - p);
  if(!nl) return false;
  p = (size_t)(nl - d) + 1;
  if(p >= end) return false;             /* no quality line */
  nl = memchr(d + p, '\n', end - p);
  if(!nl && !at_eof) return false;       /* the quality line crosses the cut: the cut was not a record start */
  r->qual = d + p; r->quallen = nl ? (size_t)(nl - (d + p)) : end - p;
  while(r->quallen && (r->qual[r->quallen - 1] == '\r')) r->quallen--;
  if(r->quallen < r->seqlen) return false; /* the sequential reader would go on reading quality lines */
  p = nl ? (size_t)(nl - d) + 1 : end;
  if(p < end && d[p] != '@') return false; /* junk between records */
  *pp = p;
  return true;
}

static void parse_fastq_range(const unsigned char *d, size_t p, size_t end, bool at_eof, bool want_qual, SegBuf *o)
{
  FqRec r;
  while(p < end) {
    const size_t rec_at = p;
    if(!fq_next(d, &p, end, at_eof, &r)) { o->dev_at = rec_at; return; }
    memcpy(o->b + o->len, r.seq, r.seqlen);
    if(want_qual) { memcpy(o->q + o->len, r.qual, r.seqlen); o->q[o->len + r.seqlen] = 0x7F; }
    o->len += r.seqlen;
    o->b[o->len++] = '\n';
    o->nreads++;
  }
}

static void *par_worker(void *arg)
{
  ParState *ps = arg;
  for(;;) {
    pthread_mutex_lock(&ps->mu);
    SegBuf *o = NULL;
    while(ps->next_seg < ps->nseg) {
      /* segments are consumed in order: take one only if a buffer is free */
      for(size_t i = 0; i < ps->nbufs; i++) if(ps->bufs[i].state == 0) { o = &ps->bufs[i]; break; }
      if(o) break;
      pthread_cond_wait(&ps->cv, &ps->mu);
    }
    if(!o) { pthread_mutex_unlock(&ps->mu); return NULL; }
    const size_t seg = ps->next_seg++;
    o->state = 1; o->seg = seg; o->len = 0; o->nreads = 0; o->dev_at = SIZE_MAX;
    pthread_mutex_unlock(&ps->mu);

    const size_t a = ps->cut[seg], b = ps->cut[seg + 1];
    if(o->cap < b - a + 2) {
      free(o->b); free(o->q); o->q = NULL;
      o->cap = b - a + 2 + (1u << 16);
      o->b = malloc(o->cap);
      if(ps->want_qual) o->q = malloc(o->cap);
      if(!o->b || (ps->want_qual && !o->q)) mcx_die("Out of memory");
    }
    if(ps->fmt == FMT_FASTA) parse_fasta_range(ps->data, a, b, o);
    else if(ps->fmt == FMT_PLAIN) parse_plain_range(ps->data, a, b, o);
    else parse_fastq_range(ps->data, a, b, b == ps->size, ps->want_qual, o);

    pthread_mutex_lock(&ps->mu);
    o->state = 2;
    pthread_cond_broadcast(&ps->cv);
    pthread_mutex_unlock(&ps->mu);
  }
}

bool mcx_load_seq_file_par(mcx_graph *g, McxSeqFile *sf, const McxLoadPrefs *prefs, mcx_load_stats *stats, int *rc,
                           McxParResume *resume)
{
  memset(resume, 0, sizeof(*resume));
  resume->qs.qmin = 0x7fffffff;
  const char *path = mcx_seq_path(sf);
  long nthreads = sysconf(_SC_NPROCESSORS_ONLN);
  if(nthreads > 16) nthreads = 16;
  if(mcx_ingest.concurrent && mcx_ingest.nfiles > 1) { nthreads /= mcx_ingest.nfiles; if(nthreads < 2) nthreads = 2; } /* files share the cores */
  if(getenv("MCX_PARSE_THREADS")) nthreads = atol(getenv("MCX_PARSE_THREADS"));
  if(nthreads > PAR_MAX_THREADS) nthreads = PAR_MAX_THREADS;
  size_t seg_bytes = PAR_SEG_BYTES_DEFAULT, min_bytes = PAR_MIN_BYTES;
  if(getenv("MCX_PARSE_SEG_BYTES") && atol(getenv("MCX_PARSE_SEG_BYTES")) > 0) { seg_bytes = (size_t)atol(getenv("MCX_PARSE_SEG_BYTES")); min_bytes = 0; }
  if(nthreads < 2 || prefs->remove_pcr || strcmp(path, "-") == 0) return false;

  struct stat st;
  int fd = open(path, O_RDONLY);
  if(fd < 0) return false;
  if(fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || (size_t)st.st_size < min_bytes || st.st_size < 2) { close(fd); return false; }
  const size_t size = (size_t)st.st_size;
  const unsigned char *d = mmap(NULL, size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if(d == MAP_FAILED) return false;
  /* gzip goes through zlib; the format is the first byte that is not white space (seq_file.h:311-323) */
  size_t p0 = 0;
  bool ok = !(d[0] == 0x1f && d[1] == 0x8b);
  if(ok) {
    /* leading blank lines are skipped; a file that begins with other white space is left to the sequential reader
     * (quirk Q9 of seq_ingest.c: the reference then drops lines at file offset 2^20) */
    while(p0 < size && d[p0] == '\n') p0++;
    ok = p0 < size && !par_is_space(d[p0]);
  }
  if(!ok) { munmap((void *)d, size); return false; }

  ParState ps; memset(&ps, 0, sizeof(ps));
  ps.data = d; ps.size = size; ps.fmt = d[p0] == '>' ? FMT_FASTA : (d[p0] == '@' ? FMT_FASTQ : FMT_PLAIN);
  ps.want_qual = ps.fmt == FMT_FASTQ && prefs->fq_cutoff != 0;
  McxQStat qs = resume->qs;
  if(ps.fmt == FMT_PLAIN) {
    /* The reference picks the reader per record while it is in its look-ahead (the reads worth the first 1000 bases;
     * every read when --fq-offset is given): see next_record / quirk Q9 in seq_ingest.c.  For FASTA and FASTQ that is the
     * same thing -- their readers stop on the next record's '>' / '@' -- but a line of a plain file that starts with
     * '>', '@' or white space is read differently there.  Such files are left to the sequential reader. */
    bool plain_ok = prefs->fq_offset == 0;
    size_t p = p0, bases = 0;
    while(plain_ok && p < size && bases < 1000) {
      const unsigned char c = d[p];
      const unsigned char *nl = memchr(d + p, '\n', size - p);
      size_t end = nl ? (size_t)(nl - d) : size, len;
      if(c == '\n') { p++; continue; }
      if(c == '>' || c == '@' || par_is_space(c)) plain_ok = false;
      for(len = end - p; len && (d[p + len - 1] == '\r'); len--) {}
      bases += len;
      p = nl ? end + 1 : size;
    }
    if(!plain_ok) { munmap((void *)d, size); return false; }
  }
  if(ps.fmt == FMT_FASTQ) {
    /* quality range of the first reads (until 1000 bases have been seen), exactly as the sequential reader collects
     * it (end_read, seq_ingest.c) -- it decides the ASCII offset at the first batch.  If the file stops being strict
     * before that, it is not for this path at all. */
    size_t p = p0; FqRec r;
    while(p < size && qs.bcount < 1000) {
      if(!fq_next(d, &p, size, true, &r)) { munmap((void *)d, size); return false; }
      if(r.quallen) {
        size_t lim = 1000 - qs.qcount, n = r.quallen < lim ? r.quallen : lim;
        for(size_t i = 0; i < n; i++) { int q = (signed char)r.qual[i]; if(q > qs.qmax) qs.qmax = q; if(q < qs.qmin) qs.qmin = q; }
        qs.bcount += r.seqlen; qs.qcount += r.quallen;
      } else qs.bcount += r.seqlen;
    }
  }
  /* cuts */
  size_t max_seg = (size - p0) / seg_bytes + 2;
  ps.cut = malloc((max_seg + 1) * sizeof(size_t));
  if(!ps.cut) mcx_die("Out of memory");
  ps.cut[0] = p0; ps.nseg = 0;
  for(size_t at = p0;;) {
    size_t q = at + seg_bytes;
    if(q >= size) { ps.cut[++ps.nseg] = size; break; }
    /* first legal cut at or after q */
    for(;;) {
      const unsigned char *nl = memchr(d + q - 1, '\n', size - (q - 1));
      if(!nl) { q = size; break; }
      q = (size_t)(nl - d) + 1;
      if(q >= size || ps.fmt == FMT_PLAIN || (ps.fmt == FMT_FASTA && d[q] == '>')) break;
      if(ps.fmt == FMT_FASTQ && d[q] == '@') {
        /* a record start if the next-but-one line starts with '+' (a guess: verified when the segments are consumed) */
        const unsigned char *l1 = memchr(d + q, '\n', size - q), *l2 = l1 ? memchr(l1 + 1, '\n', size - (size_t)(l1 + 1 - d)) : NULL;
        if(l2 && (size_t)(l2 + 1 - d) < size && l2[1] == '+') break;
      }
      q++; /* not a place to cut: keep looking */
    }
    ps.cut[++ps.nseg] = q;
    if(q >= size) break;
    at = q;
  }

  mcx_status("[seq] Parsing sequence file %s", path);
  ps.nbufs = (size_t)nthreads * 2 < ps.nseg ? (size_t)nthreads * 2 : ps.nseg;
  ps.bufs = calloc(ps.nbufs, sizeof(SegBuf));
  pthread_mutex_init(&ps.mu, NULL); pthread_cond_init(&ps.cv, NULL);
  pthread_t th[PAR_MAX_THREADS];
  size_t nth = (size_t)nthreads < ps.nseg ? (size_t)nthreads : ps.nseg;
  for(size_t i = 0; i < nth; i++) if(pthread_create(&th[i], NULL, par_worker, &ps) != 0) mcx_die("Cannot start a thread");

  int err = 0; uint64_t nreads_total = 0;
  bool any_qual = false, offset_known = false; uint8_t fq_offset = 0;
  size_t resume_at = SIZE_MAX;
  for(size_t seg = 0; seg < ps.nseg; seg++) {
    SegBuf *o = NULL;
    pthread_mutex_lock(&ps.mu);
    for(;;) {
      for(size_t i = 0; i < ps.nbufs; i++) if(ps.bufs[i].state == 2 && ps.bufs[i].seg == seg) { o = &ps.bufs[i]; break; }
      if(o) break;
      pthread_cond_wait(&ps.cv, &ps.mu);
    }
    pthread_mutex_unlock(&ps.mu);
    if(seg == 0) mcx_phase("  first segment parsed");
    if(!g && mcx_graph_source.wait) g = mcx_graph_source.wait(mcx_graph_source.ctx);
    if(!err && o->len) {
      mcx_read_batch b; memset(&b, 0, sizeof(b));
      b.seq = o->b; b.nbytes = o->len; b.layout = MCX_LAYOUT_LINES; b.mem = MCX_MEM_HOST;
      b.colour = prefs->colour; b.hp_cutoff = prefs->hp_cutoff; b.must_exist = prefs->must_exist;
      if(ps.fmt == FMT_FASTQ) any_qual = true; /* a strict record has a quality string */
      if(ps.want_qual) {
        /* build_graph.c:202-207 + seq_file.h:636-682, as prepare_batch of seq_ingest.c */
        if(!offset_known) {
          fq_offset = prefs->fq_offset ? prefs->fq_offset : mcx_guess_fq_offset(&qs);
          offset_known = true;
          if(fq_offset + prefs->fq_cutoff >= 127) err = MCX_ERR_UNSUPPORTED;
        }
        b.qual = o->q; b.fq_cutoff = (uint8_t)(prefs->fq_cutoff + fq_offset);
      }
      if(!err) { int r = mcx_submit_reads(g, &b); if(r != MCX_OK) err = r; }
    }
    nreads_total += o->nreads;
    const size_t dev_at = o->dev_at;
    pthread_mutex_lock(&ps.mu);
    o->state = 0;
    if(dev_at != SIZE_MAX) ps.next_seg = ps.nseg; /* no more segments: the sequential reader takes over */
    pthread_cond_broadcast(&ps.cv);
    pthread_mutex_unlock(&ps.mu);
    if(dev_at != SIZE_MAX) { resume_at = dev_at; break; }
  }
  for(size_t i = 0; i < nth; i++) pthread_join(th[i], NULL);
  for(size_t i = 0; i < ps.nbufs; i++) { free(ps.bufs[i].b); free(ps.bufs[i].q); }
  free(ps.bufs); free(ps.cut);
  pthread_mutex_destroy(&ps.mu); pthread_cond_destroy(&ps.cv);
  munmap((void *)d, size);
  if(resume_at != SIZE_MAX && !err) {
    resume->resume = true; resume->offset = resume_at; resume->g = g;
    resume->qs = qs; resume->any_qual = any_qual; resume->offset_known = offset_known; resume->fq_offset = fq_offset;
    resume->nreads = nreads_total;
    return false;
  }
  mcx_phase("  parsed + submitted");
  if(!g && mcx_graph_source.wait) g = mcx_graph_source.wait(mcx_graph_source.ctx);

  int r = err;
  if(!mcx_ingest.concurrent) {
    mcx_load_stats s;
    r = mcx_sync_reads(g, &s);
    if(err) r = err;
    mcx_add_load_stats(stats, &s);
  }
  char n1[64]; mcx_ulong_to_str(nreads_total, n1);
  mcx_status("[seq] Loaded %s reads and 0 reads pairs (file: %s)", n1, path);
  *rc = r;
  return true;
}
