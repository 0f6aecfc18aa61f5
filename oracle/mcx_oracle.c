/*
 * oracle/mcx_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded CPU restatement of the `mccortex build` hot path
 * (read -> contigs -> k-mer -> canonical key -> Lookup3 -> find-or-insert ->
 * coverage / edges -> sorted .ctx v6 bytes).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load this library; the product
 * (mccortex_b200/, include/) never links, imports or executes it.
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * it restates.  Parity pinned: tests/test_oracle_vs_ref.py compares the bytes
 * this file produces with the compiled, unmodified reference (oracle/_ref/,
 * built by oracle/Makefile) and with the golden vectors in tests/golden/.
 *
 * The hash set below is deliberately NOT the reference's bucketed layout: the
 * reference's unsorted iteration order is seed/thread dependent, so
 * the contract is the sorted dump; any exact set with the same update rules
 * gives the same bytes.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <zlib.h>

#define ORC_MAXW 2

typedef struct { uint64_t b[ORC_MAXW]; } OrcKmer; /* b[0] = most significant word (binary_kmer.h:7,39-49) */

/* ------------------------------------------------------------------ row A */
/* src/basic/dna.c:8-25 : A/a=0 C/c=1 G/g=2 T/t=3, N/n=4, everything else 8 */
uint8_t orc_char_to_nuc(char c)
{
  switch(c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    case 'N': case 'n': return 4;
    default: return 8;
  }
}
/* src/basic/dna.h:15-16 */
static inline int orc_is_acgt(char c) { return orc_char_to_nuc(c) < 4; }

/* ------------------------------------------------------------------ row B */
/* src/basic/seq_reader.c:61-117 (seq_contig_start2) */
size_t orc_contig_start(const char *seq, size_t seqlen, const char *qual, size_t quallen,
                        size_t offset, size_t k, uint8_t qual_cutoff, uint8_t hp_cutoff)
{
  if(!qual || !quallen) { qual = NULL; quallen = 0; }
  size_t kmerend, pos = offset;
  while((kmerend = pos + k) <= seqlen)
  {
    size_t i = kmerend;
    while(i > pos && orc_is_acgt(seq[i-1])) i--;
    if(i > pos) { pos = i; continue; }

    if(qual && qual_cutoff > 0) {
      i = kmerend < quallen ? kmerend : quallen;
      /* `char` vs uint8_t compare after integer promotion, as in the reference */
      while(i > pos && qual[i-1] > qual_cutoff) i--;
      if(i > pos) { pos = i; continue; }
    }

    if(hp_cutoff > 0) {
      size_t run_length = 1;
      for(i = kmerend-1; i > pos; i--) {
        if(seq[i-1] == seq[i]) { run_length++; if(run_length == (size_t)hp_cutoff) break; }
        else run_length = 1;
      }
      if(i > pos) { pos = i; continue; }
    }
    return pos;
  }
  return seqlen;
}

/* src/basic/seq_reader.c:127-172 (seq_contig_end2) */
size_t orc_contig_end(const char *seq, size_t seqlen, const char *qual, size_t quallen,
                      size_t contig_start, size_t k, uint8_t qual_cutoff, uint8_t hp_cutoff,
                      size_t *search_start)
{
  if(!qual || !quallen) { qual = NULL; quallen = 0; }
  size_t contig_end = contig_start + k;
  size_t hp_run = 1;
  if(hp_cutoff > 0) {
    while(hp_run < contig_end && seq[contig_end-1-hp_run] == seq[contig_end-1]) hp_run++;
  }
  for(; contig_end < seqlen; contig_end++)
  {
    if(!orc_is_acgt(seq[contig_end]) ||
       (contig_end < quallen && qual[contig_end] < qual_cutoff)) break;
    if(hp_cutoff > 0) {
      if(seq[contig_end] == seq[contig_end-1]) { hp_run++; if(hp_run >= (size_t)hp_cutoff) break; }
      else hp_run = 1;
    }
  }
  if(hp_cutoff > 0 && hp_run >= (size_t)hp_cutoff) *search_start = contig_end - (size_t)hp_cutoff + 1;
  else *search_start = contig_end;
  return contig_end;
}

/* ------------------------------------------------------------------ row C */
static inline int orc_nwords(size_t k) { return (int)((k + 31) / 32); }       /* binary_kmer.h:10-11 */
static inline unsigned orc_top_bases(size_t k) { return (unsigned)(k & 31); } /* bases in b[0]; k odd => 1..31 */

/* src/graph/binary_kmer.h:139-146,160-167 + binary_kmer.c:80-97 */
static void orc_left_shift_add(OrcKmer *bk, size_t k, uint8_t nuc)
{
  int W = orc_nwords(k), i;
  for(i = 0; i + 1 < W; i++) bk->b[i] = (bk->b[i] << 2) | (bk->b[i+1] >> 62);
  bk->b[W-1] = (bk->b[W-1] << 2) | nuc;
  bk->b[0] &= (UINT64_MAX >> (64 - 2*orc_top_bases(k)));
}

/* src/graph/binary_kmer.c:156-186 */
OrcKmer orc_bkmer_from_str(const char *seq, size_t k)
{
  OrcKmer bk; size_t i;
  memset(&bk, 0, sizeof(bk));
  for(i = 0; i < k; i++) orc_left_shift_add(&bk, k, orc_char_to_nuc(seq[i]));
  return bk;
}

/* ------------------------------------------------------------------ row D */
/* src/graph/binary_kmer.c:102-133 : per word byte-swap, swap the 2-bit fields
 * inside each byte, NOT, reverse word order, then shift the whole W-word value
 * right by 64 - 2*(k&31) bits.  Written here base by base (the slow, obviously
 * correct way) so that it is independent of the bit tricks the GPU code uses. */
OrcKmer orc_revcomp(OrcKmer bk, size_t k)
{
  int W = orc_nwords(k);
  OrcKmer rc; size_t i;
  memset(&rc, 0, sizeof(rc));
  for(i = 0; i < k; i++) {
    /* base i counted from the LAST base of bk (least significant 2 bits) */
    size_t word = (size_t)(W - 1) - i / 32, sh = 2 * (i % 32);
    uint8_t nuc = (uint8_t)((bk.b[word] >> sh) & 3);
    orc_left_shift_add(&rc, k, (uint8_t)(~nuc & 3)); /* dna.h:22 complement */
  }
  return rc;
}

static inline int orc_kmer_lt(const OrcKmer *a, const OrcKmer *b, int W) /* binary_kmer.h:79-94 */
{
  int i;
  for(i = 0; i < W; i++) if(a->b[i] != b->b[i]) return a->b[i] < b->b[i];
  return 0;
}
static inline int orc_kmer_eq(const OrcKmer *a, const OrcKmer *b, int W)
{
  int i;
  for(i = 0; i < W; i++) if(a->b[i] != b->b[i]) return 0;
  return 1;
}

/* src/graph/binary_kmer.c:43-57 ; orientation db_node.h:109-110 (0 FORWARD, 1 REVERSE) */
OrcKmer orc_get_key(OrcKmer bk, size_t k, int *orient)
{
  int W = orc_nwords(k);
  OrcKmer rc = orc_revcomp(bk, k);
  if(orc_kmer_lt(&rc, &bk, W)) { if(orient) *orient = 1; return rc; }
  if(orient) *orient = 0;
  return bk;
}

/* ------------------------------------------------------------------ row E */
#define ORC_ROT(x,k) (((x)<<(k)) | ((x)>>(32-(k))))
/* src/kmer/kmer_hash.h:89-97 */
#define ORC_MIX(a,b,c) { \
  a -= c;  a ^= ORC_ROT(c, 4);  c += b; \
  b -= a;  b ^= ORC_ROT(a, 6);  a += c; \
  c -= b;  c ^= ORC_ROT(b, 8);  b += a; \
  a -= c;  a ^= ORC_ROT(c,16);  c += b; \
  b -= a;  b ^= ORC_ROT(a,19);  a += c; \
  c -= b;  c ^= ORC_ROT(b, 4);  b += a; }
/* src/kmer/kmer_hash.h:124-133 */
#define ORC_FINAL(a,b,c) { \
  c ^= b; c -= ORC_ROT(b,14); \
  a ^= c; a -= ORC_ROT(c,11); \
  b ^= a; b -= ORC_ROT(a,25); \
  c ^= b; c -= ORC_ROT(b,16); \
  a ^= c; a -= ORC_ROT(c,4);  \
  b ^= a; b -= ORC_ROT(a,14); \
  c ^= b; c -= ORC_ROT(b,24); }

/* src/kmer/kmer_hash.h:162-211 specialised to 8*W key bytes read as
 * little-endian u32 from the in-memory BinaryKmer {b[0], b[1]}.
 * If b_out != NULL it also returns the second 32-bit lane (`b` after final),
 * which lookup3's hashlittle2 exposes; the reference only uses `c`. */
uint32_t orc_lookup3(const uint64_t *kb, int W, uint32_t initval, uint32_t *b_out)
{
  uint32_t a, b, c;
  a = b = c = 0xdeadbeefu + (uint32_t)(8*W) + initval;
  if(W == 1) {
    a += (uint32_t)kb[0]; b += (uint32_t)(kb[0] >> 32);
  } else {
    a += (uint32_t)kb[0]; b += (uint32_t)(kb[0] >> 32); c += (uint32_t)kb[1];
    ORC_MIX(a, b, c);
    a += (uint32_t)(kb[1] >> 32);
  }
  ORC_FINAL(a, b, c);
  if(b_out) *b_out = b;
  return c;
}

/* XOR of hashes of {b[0]=i} for i in [0,n): what `mccortex31 hashtest -F n`
 * prints (src/commands/ctx_exp_hashtest.c:40-69 fast-hash loop). */
uint32_t orc_hashtest_xor(uint64_t n)
{
  uint64_t i; uint32_t x = 0;
  for(i = 0; i < n; i++) { uint64_t kb[1] = { i }; x ^= orc_lookup3(kb, 1, 0, NULL); }
  return x;
}

/* --------------------------------------------------------------- rows F, G */
#define ORC_FLAG (1ULL << 63) /* BKMER_SET_FLAG, src/graph/hash_table.h:14-15 */

typedef struct {
  uint64_t total_bases_read, total_bases_loaded, contigs_parsed;
  uint64_t num_kmers_loaded, num_kmers_novel;
  uint64_t num_se_reads, num_pe_reads, num_good_reads, num_bad_reads;
  uint64_t num_dup_se_reads, num_dup_pe_pairs;
} OrcStats; /* subset of src/basic/seq_loading_stats.h:5-14 */

typedef struct {
  uint32_t mean_read_length;
  uint64_t total_sequence;
  long double seq_err;
  char name[1024];
  /* ErrorCleaning, src/basic/graph_info.h: all-default unless a graph file is loaded (build --graph) */
  uint8_t cleaned_tips, cleaned_unitigs, cleaned_kmers, is_graph_intersection;
  uint32_t clean_unitigs_thresh, clean_kmers_thresh;
  char isec_name[1024];
} OrcGInfo; /* src/basic/graph_info.h:20-27 */

typedef struct {
  size_t k, ncols; int W;
  uint64_t cap, mask, nkmers, limit; /* limit = "capacity" the caller asked for */
  uint64_t *keys;  /* cap * W, word 0 carries ORC_FLAG when assigned */
  uint32_t *covgs; /* cap * ncols   (db_graph.h:39, db_node.h:284-285) */
  uint8_t  *edges; /* cap * ncols   (db_graph.h:40) */
  OrcGInfo *ginfo;
  int full;
  /* build --intersect (src/commands/ctx_build.c:293-303,341-361,409-413) */
  int must_exist;        /* SeqLoadingPrefs.must_exist_in_graph for every read */
  uint8_t *isec_edges;   /* cap bytes, or NULL */
  /* build --remove-pcr: "a read started here" per node and orientation (db_graph.h readstrt,
   * src/tools/build_graph.c:29-32), one byte per bit here */
  uint8_t *readstrt;     /* cap * 2 bytes, or NULL */
} OrcGraph;

static void orc_ginfo_init(OrcGInfo *g) /* src/basic/graph_info.c:60-67 */
{
  strcpy(g->name, "undefined");
  g->total_sequence = 0; g->mean_read_length = 0; g->seq_err = 0.01;
  g->cleaned_tips = g->cleaned_unitigs = g->cleaned_kmers = g->is_graph_intersection = 0;   /* graph_info.c:4-10 */
  g->clean_unitigs_thresh = g->clean_kmers_thresh = 0;
  strcpy(g->isec_name, "undefined");
}

OrcGraph *orc_graph_new(size_t k, size_t ncols, uint64_t capacity)
{
  OrcGraph *g = (OrcGraph*)calloc(1, sizeof(*g));
  size_t i;
  g->k = k; g->ncols = ncols; g->W = orc_nwords(k);
  g->limit = capacity;
  g->cap = 1024; while(g->cap < capacity * 2) g->cap <<= 1;
  g->mask = g->cap - 1;
  g->keys  = (uint64_t*)calloc(g->cap * (size_t)g->W, 8);
  g->covgs = (uint32_t*)calloc(g->cap * ncols, 4);
  g->edges = (uint8_t*)calloc(g->cap * ncols, 1);
  g->ginfo = (OrcGInfo*)calloc(ncols, sizeof(OrcGInfo));
  for(i = 0; i < ncols; i++) orc_ginfo_init(&g->ginfo[i]);
  return g;
}

void orc_graph_free(OrcGraph *g)
{
  if(!g) return;
  free(g->keys); free(g->covgs); free(g->edges); free(g->ginfo); free(g->isec_edges); free(g->readstrt); free(g);
}

uint64_t orc_graph_nkmers(const OrcGraph *g) { return g->nkmers; }
int orc_graph_is_full(const OrcGraph *g) { return g->full; }
void orc_graph_set_name(OrcGraph *g, size_t col, const char *name)
{
  strncpy(g->ginfo[col].name, name, sizeof(g->ginfo[col].name)-1);
}

/* find-or-insert semantics of src/graph/hash_table.c:250-281: returns slot, sets *found.
 * "Hash table is full" (hash_table.c:119-123) becomes g->full once the caller's
 * requested capacity is exhausted. */
static uint64_t orc_find_or_insert(OrcGraph *g, const OrcKmer *key, int *found)
{
  int W = g->W, i;
  uint64_t h = orc_lookup3(key->b, W, 0, NULL) & g->mask;
  /* (a graph already flagged full is discarded by every caller: stop before the array itself has no empty slot left) */
  if(g->full && g->nkmers + 2 >= g->cap) { *found = 1; return h; }
  for(;; h = (h + 1) & g->mask) {
    uint64_t *s = g->keys + h * (uint64_t)W;
    if(s[0] == 0) {
      if(g->nkmers >= g->limit) { g->full = 1; }
      s[0] = key->b[0] | ORC_FLAG;
      for(i = 1; i < W; i++) s[i] = key->b[i];
      g->nkmers++;
      *found = 0;
      return h;
    }
    if(s[0] == (key->b[0] | ORC_FLAG)) {
      for(i = 1; i < W && s[i] == key->b[i]; i++) {}
      if(i == W) { *found = 1; return h; }
    }
  }
}

typedef struct { uint64_t slot; int orient; } OrcNode;

/* src/tools/build_graph.c:99-117 (_find_or_insert, must_exist_in_graph=false)
 * + src/graph/db_graph.c:101-105,126-134 + src/graph/db_node.c:139-144 */
#define ORC_NOT_FOUND UINT64_MAX
/* hash_table_find */
static uint64_t orc_find(const OrcGraph *g, const OrcKmer *key)
{
  size_t W = (size_t)g->W, w;
  uint64_t h = orc_lookup3(key->b, (int)W, 0, NULL) & g->mask;
  for(;; h = (h + 1) & g->mask) {
    const uint64_t *sl = g->keys + h * W;
    if(sl[0] == 0) return ORC_NOT_FOUND;
    if(sl[0] == (key->b[0] | ORC_FLAG)) { for(w = 1; w < W && sl[w] == key->b[w]; w++) {} if(w == W) return h; }
  }
}

/* _find_or_insert, src/tools/build_graph.c:99-118: with must_exist_in_graph the k-mer is only looked up
 * (slot ORC_NOT_FOUND if absent) and coverage is added only when it is there */
static OrcNode orc_add_kmer(OrcGraph *g, OrcKmer bk, size_t colour, int *found)
{
  OrcNode n; int o;
  OrcKmer key = orc_get_key(bk, g->k, &o);
  n.orient = o;
  if(g->must_exist) {
    n.slot = orc_find(g, &key);
    *found = (n.slot != ORC_NOT_FOUND);
    if(!*found) return n;
  } else n.slot = orc_find_or_insert(g, &key, found);
  uint32_t *cv = &g->covgs[n.slot * g->ncols + colour];
  if(*cv < UINT32_MAX) (*cv)++; /* saturating */
  return n;
}

/* src/tools/build_graph.c:122-150 (build_graph_from_str_mt) with
 * src/graph/db_graph.c:152-166 (db_graph_add_edge_mt) and db_node.h:180,273-274.
 * Returns the number of non-novel k-mers. */
size_t orc_graph_add_contig(OrcGraph *g, size_t colour, const char *seq, size_t len)
{
  size_t k = g->k, i, nonnovel = 0;
  int found;
  OrcKmer bk = orc_bkmer_from_str(seq, k);
  OrcNode prev = orc_add_kmer(g, bk, colour, &found), curr;
  nonnovel += (size_t)found;
  for(i = k; i < len; i++, prev = curr) {
    uint8_t nuc = orc_char_to_nuc(seq[i]);
    orc_left_shift_add(&bk, k, nuc);
    curr = orc_add_kmer(g, bk, colour, &found);
    /* lhs = first base of prev as read, rhs = last base of curr as read */
    uint8_t lhs = orc_char_to_nuc(seq[i - k]), rhs = nuc;
    uint8_t lhs_rev = (uint8_t)(~lhs & 3);
    if(prev.slot != ORC_NOT_FOUND && curr.slot != ORC_NOT_FOUND) { /* build_graph.c:144-145 */
      g->edges[prev.slot * g->ncols + colour] |= (uint8_t)(1u << (rhs + 4 * prev.orient));
      g->edges[curr.slot * g->ncols + colour] |= (uint8_t)(1u << (lhs_rev + 4 * (!curr.orient)));
    }
    nonnovel += (size_t)found;
  }
  return nonnovel;
}

/* src/tools/build_graph.c:154-189 (load_read) + :192-231 (build_graph_from_reads_mt,
 * single-end, no PCR-duplicate removal).  fq_cutoff is the raw -Q value;
 * fq_offset is added only when fq_cutoff != 0 (build_graph.c:202-207). */
void orc_graph_add_read(OrcGraph *g, const char *seq, size_t seqlen, const char *qual, size_t quallen,
                        size_t colour, uint8_t fq_cutoff, uint8_t fq_offset, uint8_t hp_cutoff,
                        OrcStats *st)
{
  size_t k = g->k, cs, ce, search = 0, ncontigs = 0;
  uint8_t qcut = fq_cutoff ? (uint8_t)(fq_cutoff + fq_offset) : 0;
  st->total_bases_read += seqlen;
  st->num_se_reads += 1;
  while((cs = orc_contig_start(seq, seqlen, qual, quallen, search, k, qcut, hp_cutoff)) < seqlen) {
    ce = orc_contig_end(seq, seqlen, qual, quallen, cs, k, qcut, hp_cutoff, &search);
    size_t clen = ce - cs;
    size_t nonnovel = orc_graph_add_contig(g, colour, seq + cs, clen);
    size_t ck = clen + 1 - k;
    st->total_bases_loaded += clen;
    if(g->must_exist) st->num_kmers_loaded += nonnovel;   /* build_graph.c:176-181 */
    else { st->num_kmers_loaded += ck; st->num_kmers_novel += ck - nonnovel; }
    ncontigs++;
  }
  st->contigs_parsed += ncontigs;
  st->num_good_reads += (ncontigs > 0);
  st->num_bad_reads += (ncontigs == 0);
}

/* ---- build --remove-pcr (row N3) ------------------------------------------------------------ */
/* DBG_ALLOC_READSTRT (ctx_build.c:336) / the wipe between colours (ctx_build.c:392-395) */
void orc_graph_wipe_readstrt(OrcGraph *g)
{
  if(!g->readstrt) g->readstrt = (uint8_t*)malloc(g->cap * 2);
  memset(g->readstrt, 0, g->cap * 2);
}

/* seq_read_reverse_complement, libs/seq_file/seq_file.h:758-777 (complement of seq_file.h:715-723: only
 * ACGTacgt change).  A quality string longer than the read is cut to the read's length first; for one
 * that is SHORTER the reference calls cbuf_capacity with the length field in place of the capacity
 * (seq_file.h:726-733), which leaves the missing qualities undefined -- the oracle pads with '.', the
 * value that code was written to pad with, and the tests do not depend on it. */
static char orc_complement(char c)
{
  switch(c) {
    case 'a': return 't'; case 'A': return 'T'; case 'c': return 'g'; case 'C': return 'G';
    case 'g': return 'c'; case 'G': return 'C'; case 't': return 'a'; case 'T': return 'A';
    default: return c;
  }
}
static void orc_read_revcomp(char *seq, size_t sl, char *qual, size_t *ql)
{
  size_t i, j;
  if(*ql > 0) { for(i = *ql; i < sl; i++) qual[i] = '.'; *ql = sl; }
  for(i = 0; i < sl / 2; i++) {
    char a = seq[i], b = seq[sl - 1 - i];
    seq[i] = orc_complement(b); seq[sl - 1 - i] = orc_complement(a);
  }
  if(sl & 1) seq[sl / 2] = orc_complement(seq[sl / 2]);
  if(*ql > 0) for(i = 0, j = *ql - 1; i < j; i++, j--) { char t = qual[i]; qual[i] = qual[j]; qual[j] = t; }
}

/* build_graph_from_reads_mt (src/tools/build_graph.c:192-231) with prefs->remove_pcr_dups set, i.e.
 * including seq_reads_are_novel (:35-92).  seq2 == NULL: a single-end read.  matedir: READPAIR_FF 0,
 * FR 1, RF 2, RR 3 (cortex_types.h:17-25); read 1 is reverse-complemented when matedir & 2, read 2 when
 * matedir & 1 (seq_reader.c:506-510) -- the reads stay that way when they are loaded.  fq_offset1/2: the
 * ASCII offsets of the two files.  Reads are taken in the order of the calls (the reference with one
 * worker thread and one input task). */
void orc_graph_add_reads_pcr(OrcGraph *g, const char *seq1, size_t sl1, const char *qual1, size_t ql1,
                             const char *seq2, size_t sl2, const char *qual2, size_t ql2,
                             size_t colour, uint8_t fq_cutoff, uint8_t fq_offset1, uint8_t fq_offset2,
                             uint8_t hp_cutoff, int matedir, OrcStats *st)
{
  size_t k = g->k;
  uint8_t qc1 = fq_cutoff ? (uint8_t)(fq_cutoff + fq_offset1) : 0, qc2 = fq_cutoff ? (uint8_t)(fq_cutoff + fq_offset2) : 0;
  char *s1 = (char*)malloc(sl1 + 1), *q1 = (char*)malloc((ql1 > sl1 ? ql1 : sl1) + 1), *s2 = NULL, *q2 = NULL;
  memcpy(s1, seq1, sl1); if(ql1) memcpy(q1, qual1, ql1);
  if(seq2) {
    s2 = (char*)malloc(sl2 + 1); q2 = (char*)malloc((ql2 > sl2 ? ql2 : sl2) + 1);
    memcpy(s2, seq2, sl2); if(ql2) memcpy(q2, qual2, ql2);
  }
  if(!g->readstrt) orc_graph_wipe_readstrt(g);
  st->total_bases_read += sl1 + (seq2 ? sl2 : 0);
  if(seq2) st->num_pe_reads += 2; else st->num_se_reads += 1;

  /* seq_reads_are_novel */
  if(matedir & 2) orc_read_revcomp(s1, sl1, q1, &ql1);
  if(seq2 && (matedir & 1)) orc_read_revcomp(s2, sl2, q2, &ql2);
  size_t start1 = orc_contig_start(s1, sl1, q1, ql1, 0, k, qc1, hp_cutoff), start2 = 0;
  int got1 = start1 < sl1, got2 = 0, found1 = 0, found2 = 0, o1 = 0, o2 = 0;
  uint64_t n1 = 0, n2 = 0;
  if(seq2) { start2 = orc_contig_start(s2, sl2, q2, ql2, 0, k, qc2, hp_cutoff); got2 = start2 < sl2; }
  if(got1) { OrcKmer key = orc_get_key(orc_bkmer_from_str(s1 + start1, k), k, &o1); n1 = orc_find_or_insert(g, &key, &found1); }
  if(got2) { OrcKmer key = orc_get_key(orc_bkmer_from_str(s2 + start2, k), k, &o2); n2 = orc_find_or_insert(g, &key, &found2); }
  st->num_kmers_novel += (uint64_t)(!found1 + !found2); /* build_graph.c:75-76: counts reads without a k-mer too */
  int novel = !((!got1 || g->readstrt[2 * n1 + (uint64_t)o1]) && (!got2 || g->readstrt[2 * n2 + (uint64_t)o2]));
  if(novel) {
    if(got1) g->readstrt[2 * n1 + (uint64_t)o1] = 1;
    if(got2) g->readstrt[2 * n2 + (uint64_t)o2] = 1;
  }

  if(!novel) { if(seq2) st->num_dup_pe_pairs++; else st->num_dup_se_reads++; }
  else {
    /* load_read of each mate; orc_graph_add_read also counts the read and its bases: undo that part */
    OrcStats t = *st;
    orc_graph_add_read(g, s1, sl1, ql1 ? q1 : NULL, ql1, colour, fq_cutoff, fq_offset1, hp_cutoff, st);
    if(seq2) orc_graph_add_read(g, s2, sl2, ql2 ? q2 : NULL, ql2, colour, fq_cutoff, fq_offset2, hp_cutoff, st);
    st->total_bases_read = t.total_bases_read; st->num_se_reads = t.num_se_reads;
  }
  free(s1); free(q1); free(s2); free(q2);
}

/* Per-window view of one read, for tuple-level checks of the GPU kernel:
 * for every window start p in [0, seqlen-k] writes in_contig[p] (0/1) and, if 1,
 * key words, orientation and the edge bits this occurrence contributes
 * (the two ORs of db_graph_add_edge_mt that land on THIS node: the edge to the
 * next window and the edge from the previous one).  Derived from the same
 * sequential contig walk as orc_graph_add_read. */
void orc_read_windows(const char *seq, size_t seqlen, const char *qual, size_t quallen, size_t k,
                      uint8_t qcut, uint8_t hp_cutoff,
                      uint8_t *in_contig, uint64_t *keys /* [n*W] */, uint8_t *orient, uint8_t *emask)
{
  int W = orc_nwords(k), j;
  size_t n = seqlen >= k ? seqlen - k + 1 : 0, cs, ce, search = 0, p;
  memset(in_contig, 0, n);
  while((cs = orc_contig_start(seq, seqlen, qual, quallen, search, k, qcut, hp_cutoff)) < seqlen) {
    ce = orc_contig_end(seq, seqlen, qual, quallen, cs, k, qcut, hp_cutoff, &search);
    OrcKmer bk = orc_bkmer_from_str(seq + cs, k);
    for(p = cs; p + k <= ce; p++) {
      if(p > cs) orc_left_shift_add(&bk, k, orc_char_to_nuc(seq[p + k - 1]));
      int o; OrcKmer key = orc_get_key(bk, k, &o);
      uint8_t m = 0;
      if(p + k < ce) m |= (uint8_t)(1u << (orc_char_to_nuc(seq[p + k]) + 4 * o));
      if(p > cs)     m |= (uint8_t)(1u << ((~orc_char_to_nuc(seq[p - 1]) & 3) + 4 * (!o)));
      in_contig[p] = 1; orient[p] = (uint8_t)o; emask[p] = m;
      for(j = 0; j < W; j++) keys[p * (size_t)W + (size_t)j] = key.b[j];
    }
  }
}

/* ------------------------------------------------------------------ row H */
/* src/basic/graph_info.c:116-133 */
static void orc_ginfo_update_contigs(OrcGInfo *gi, uint64_t added_seq, uint64_t num_contigs)
{
  if(!added_seq && !num_contigs) return;
  size_t ginfo_num_contigs = 0;
  if(gi->total_sequence && gi->mean_read_length)
    ginfo_num_contigs = ((double)gi->total_sequence / gi->mean_read_length) + 0.5;
  if(ginfo_num_contigs + num_contigs > 0) {
    gi->mean_read_length = (uint32_t)((double)(gi->total_sequence + added_seq) /
                                      (ginfo_num_contigs + num_contigs));
  }
  gi->total_sequence += added_seq;
}

/* src/tools/build_graph.c:295-300 -> graph_info_update_stats (graph_info.c:172-175).
 * Called once per build_graph() batch with the stats the reference credits to
 * the batch's FIRST task (quirk Q1, build_graph.c:242). */
void orc_graph_update_ginfo(OrcGraph *g, size_t colour, const OrcStats *st)
{
  orc_ginfo_update_contigs(&g->ginfo[colour], st->total_bases_loaded, st->contigs_parsed);
}

/* src/basic/graph_info.c:135-170 with dst freshly initialised (graph_writer.c:11-30) */
static void orc_ginfo_merge(OrcGInfo *dst, const OrcGInfo *src)
{
  if(strcmp(src->name, "undefined") != 0) {
    if(strcmp(dst->name, "undefined") == 0) strcpy(dst->name, src->name);
    else { strcat(dst->name, ","); strcat(dst->name, src->name); }
  }
  uint64_t total_sequence = dst->total_sequence + src->total_sequence;
  if(total_sequence > 0) {
    dst->seq_err = (dst->seq_err * dst->total_sequence + src->seq_err * src->total_sequence) / total_sequence;
    size_t src_num_contigs = 0;
    if(src->total_sequence && src->mean_read_length)
      src_num_contigs = ((double)src->total_sequence / src->mean_read_length) + 0.5;
    orc_ginfo_update_contigs(dst, src->total_sequence, src_num_contigs);
  }
  /* error_cleaning_merge, src/basic/graph_info.c:34-58 (+ graph_info_append_intersect :88-101) */
  dst->cleaned_tips |= src->cleaned_tips;
  dst->cleaned_unitigs |= src->cleaned_unitigs;
  dst->cleaned_kmers |= src->cleaned_kmers;
  if(src->clean_unitigs_thresh > 0 && (dst->clean_unitigs_thresh == 0 || src->clean_unitigs_thresh < dst->clean_unitigs_thresh))
    dst->clean_unitigs_thresh = src->clean_unitigs_thresh;
  if(src->clean_kmers_thresh > 0 && (dst->clean_kmers_thresh == 0 || src->clean_kmers_thresh < dst->clean_kmers_thresh))
    dst->clean_kmers_thresh = src->clean_kmers_thresh;
  if(src->is_graph_intersection) {
    if(!dst->is_graph_intersection) strcpy(dst->isec_name, src->isec_name);
    else { strcat(dst->isec_name, ","); strcat(dst->isec_name, src->isec_name); }
    dst->is_graph_intersection = 1;
  }
  dst->total_sequence = total_sequence;
}

static size_t orc_put(uint8_t *buf, size_t off, const void *src, size_t n)
{
  if(buf) memcpy(buf + off, src, n);
  return off + n;
}

/* src/graph/graph_writer.c:62-110 (graph_write_header) + :33-60 (error cleaning object).
 * buf may be NULL to size the header. */
size_t orc_graph_write_header(const OrcGraph *g, uint8_t *buf)
{
  size_t off = 0, i, ncols = g->ncols;
  uint32_t version = 6, k32 = (uint32_t)g->k, W32 = (uint32_t)g->W, c32 = (uint32_t)ncols;
  OrcGInfo *h = (OrcGInfo*)calloc(ncols, sizeof(OrcGInfo));
  for(i = 0; i < ncols; i++) { orc_ginfo_init(&h[i]); orc_ginfo_merge(&h[i], &g->ginfo[i]); }
  off = orc_put(buf, off, "CORTEX", 6);
  off = orc_put(buf, off, &version, 4);
  off = orc_put(buf, off, &k32, 4);
  off = orc_put(buf, off, &W32, 4);
  off = orc_put(buf, off, &c32, 4);
  for(i = 0; i < ncols; i++) off = orc_put(buf, off, &h[i].mean_read_length, 4);
  for(i = 0; i < ncols; i++) off = orc_put(buf, off, &h[i].total_sequence, 8);
  for(i = 0; i < ncols; i++) {
    uint32_t len = (uint32_t)strlen(h[i].name);
    off = orc_put(buf, off, &len, 4);
    off = orc_put(buf, off, h[i].name, len);
  }
  for(i = 0; i < ncols; i++) {
    /* raw x87 long double: 10 value bytes + 6 padding bytes.  The reference
     * writes whatever is in the padding; it is zero there because the struct
     * comes from calloc and x87 stores only touch 10 bytes.  We zero it. */
    uint8_t ld[16]; memset(ld, 0, 16); memcpy(ld, &h[i].seq_err, 10);
    off = orc_put(buf, off, ld, sizeof(long double));
  }
  for(i = 0; i < ncols; i++) {
    uint8_t flags[4] = {h[i].cleaned_tips, h[i].cleaned_unitigs, h[i].cleaned_kmers, h[i].is_graph_intersection};
    uint32_t tu = h[i].cleaned_unitigs ? h[i].clean_unitigs_thresh : 0, tk = h[i].cleaned_kmers ? h[i].clean_kmers_thresh : 0;
    uint32_t len = (uint32_t)strlen(h[i].isec_name);
    off = orc_put(buf, off, flags, 4);
    off = orc_put(buf, off, &tu, 4);
    off = orc_put(buf, off, &tk, 4);
    off = orc_put(buf, off, &len, 4);
    off = orc_put(buf, off, h[i].isec_name, len);
  }
  off = orc_put(buf, off, "CORTEX", 6);
  free(h);
  return off;
}

/* ------------------------------------------------------------- build --graph
 * graph_load(), src/graph/graphs_load.c:83-208, on the records of a .ctx file that the caller has
 * read; the colour filter is the list of (from, into) pairs of src/basic/file_filter.c.
 * Per record (graph_file_read, graph_file_reader.c:389-404): covgs[into] = SAFE_ADD(covgs[into],
 * file covg[from]); edges[into] |= file edges[from]; skipped if every covgs[into] is zero
 * (graphs_load.c:121-124); find-or-insert (or find only: must_exist_in_graph); then
 * db_node_add_col_covg (saturating) and col_edges[into] |= edges (edge_mask = 0xff without --intersect). */
static uint32_t orc_safe_add_covg(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; return s > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)s; }
uint64_t orc_graph_load_records(OrcGraph *g, const uint8_t *recs, uint64_t n, uint32_t file_ncols,
                                const uint32_t *from, const uint32_t *into, uint32_t nmap, int flags, uint64_t *novel_out)
{
  const int must_exist = flags & 1;
  const size_t W = (size_t)g->W, C = g->ncols, rec = 8*W + 5*(size_t)file_ncols;
  uint32_t *cv = (uint32_t*)calloc(C, 4); uint8_t *ed = (uint8_t*)calloc(C, 1);
  uint64_t i, loaded = 0, novel = 0; uint32_t m; size_t c;
  for(i = 0; i < n; i++) {
    const uint8_t *r = recs + i * rec;
    OrcKmer key; memset(&key, 0, sizeof(key));
    memcpy(key.b, r, 8*W);
    memset(cv, 0, 4*C); memset(ed, 0, C);
    for(m = 0; m < nmap; m++) {
      uint32_t fc; memcpy(&fc, r + 8*W + 4*(size_t)from[m], 4);
      cv[into[m]] = orc_safe_add_covg(cv[into[m]], fc);
      ed[into[m]] |= r[8*W + 4*(size_t)file_ncols + from[m]];
    }
    uint32_t keep = 0;
    for(c = 0; c < C; c++) keep |= cv[c];
    if(!keep) continue;
    uint64_t slot; int found = 0;
    if(must_exist) {
      slot = orc_find(g, &key);
      if(slot == ORC_NOT_FOUND) continue;
    } else {
      slot = orc_find_or_insert(g, &key, &found);
      novel += !found;
    }
    if(flags & 2) {
      /* intersection graph: col_covgs is NULL and col_edges is the single isec edge set (ctx_build.c:350-358) */
      for(c = 0; c < C; c++) g->isec_edges[slot] |= ed[c];
    } else {
      uint8_t edge_mask = (flags & 4) ? g->isec_edges[slot] : 0xff;   /* prefs.must_exist_in_edges */
      for(c = 0; c < C; c++) {
        g->covgs[slot*C + c] = orc_safe_add_covg(g->covgs[slot*C + c], cv[c]);
        g->edges[slot*C + c] |= (uint8_t)(ed[c] & edge_mask);
      }
    }
    loaded++;
  }
  free(cv); free(ed);
  if(novel_out) *novel_out = novel;
  return loaded;
}

/* graph_load_ginfo (graphs_load.c:45-76): graph_info_merge(ginfo + into, file header colour).
 * seq_err16: the 16 bytes of the x87 long double as stored in the file. */
void orc_graph_merge_file_ginfo(OrcGraph *g, size_t into, uint32_t mean_read_length, uint64_t total_sequence,
                                const char *name, const uint8_t *seq_err16, const uint8_t *flags4,
                                uint32_t thr_unitigs, uint32_t thr_kmers, const char *isec_name)
{
  OrcGInfo src; orc_ginfo_init(&src);
  src.mean_read_length = mean_read_length; src.total_sequence = total_sequence;
  strncpy(src.name, name, sizeof(src.name) - 1); src.name[sizeof(src.name) - 1] = 0;
  memcpy(&src.seq_err, seq_err16, 10);
  src.cleaned_tips = flags4[0]; src.cleaned_unitigs = flags4[1]; src.cleaned_kmers = flags4[2]; src.is_graph_intersection = flags4[3];
  /* graph_file_reader.c:221-243: thresholds without the matching flag are dropped */
  src.clean_unitigs_thresh = src.cleaned_unitigs ? thr_unitigs : 0;
  src.clean_kmers_thresh = src.cleaned_kmers ? thr_kmers : 0;
  strncpy(src.isec_name, isec_name, sizeof(src.isec_name) - 1); src.isec_name[sizeof(src.isec_name) - 1] = 0;
  orc_ginfo_merge(&g->ginfo[into], &src);
}

/* build --intersect set-up and tear-down.
 * orc_graph_set_intersect: allocate isec_edges (ctx_build.c:341-343); reads from now on must exist.
 * orc_graph_finish_intersect: db_graph_remove_no_covg_kmers + db_graph_intersect_edges
 * (src/graph/db_graph.c:632-673): a deleted k-mer simply is not dumped (hash_table_delete). */
void orc_graph_set_intersect(OrcGraph *g, int must_exist_reads)
{
  if(!g->isec_edges) g->isec_edges = (uint8_t*)calloc(g->cap, 1);
  g->must_exist = must_exist_reads;
}
#define ORC_TOMBSTONE (ORC_FLAG | (1ULL << 62))
void orc_graph_finish_intersect(OrcGraph *g)
{
  uint64_t h; size_t c, C = g->ncols, W = (size_t)g->W;
  for(h = 0; h < g->cap; h++) {
    uint64_t *sl = g->keys + h * W;
    if(sl[0] == 0 || sl[0] == ORC_TOMBSTONE) continue;
    uint32_t any = 0;
    for(c = 0; c < C; c++) any |= g->covgs[h*C + c];
    if(!any) { sl[0] = ORC_TOMBSTONE; g->nkmers--; continue; }
    for(c = 0; c < C; c++) g->edges[h*C + c] &= g->isec_edges[h];
  }
}

static int orc_W_for_sort;
static int orc_cmp_keys(const void *a, const void *b)
{
  const uint64_t *x = *(const uint64_t * const *)a, *y = *(const uint64_t * const *)b;
  int i;
  uint64_t x0 = x[0] & ~ORC_FLAG, y0 = y[0] & ~ORC_FLAG;
  if(x0 != y0) return x0 < y0 ? -1 : 1;
  for(i = 1; i < orc_W_for_sort; i++) if(x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
  return 0;
}

/* src/graph/graph_writer.c:116-127,182-193 + hash_table.c:362-374: header, then one
 * record per k-mer in ascending key order: W x u64 key (flag cleared), u32 covg[C], u8 edges[C].
 * buf may be NULL to size the output.  Returns total bytes. */
size_t orc_graph_dump_sorted(const OrcGraph *g, uint8_t *buf)
{
  size_t off = orc_graph_write_header(g, buf), W = (size_t)g->W, C = g->ncols, i, n = 0;
  if(!buf) return off + g->nkmers * (8*W + 5*C);
  const uint64_t **ptrs = (const uint64_t**)malloc((g->nkmers + 1) * sizeof(*ptrs));
  for(i = 0; i < g->cap; i++) if(g->keys[i*W] && g->keys[i*W] != (ORC_FLAG | (1ULL << 62))) ptrs[n++] = &g->keys[i*W]; /* (not the removed ones) */
  orc_W_for_sort = g->W;
  qsort(ptrs, n, sizeof(*ptrs), orc_cmp_keys);
  for(i = 0; i < n; i++) {
    size_t slot = (size_t)(ptrs[i] - g->keys) / W, j;
    uint64_t w0 = ptrs[i][0] & ~ORC_FLAG;
    off = orc_put(buf, off, &w0, 8);
    for(j = 1; j < W; j++) off = orc_put(buf, off, &ptrs[i][j], 8);
    off = orc_put(buf, off, &g->covgs[slot*C], 4*C);
    off = orc_put(buf, off, &g->edges[slot*C], C);
  }
  free(ptrs);
  return off;
}

/* ------------------------------------------------------------------ row I */
/* src/basic/hash_mem.c:5-15 */
uint64_t orc_hash_table_cap(uint64_t nkmers, uint64_t *nbkts, uint8_t *bktsize)
{
  uint64_t num_of_buckets, bucket_size, num_of_bits = 10;
  while(nkmers / (1UL << num_of_bits) > 48) num_of_bits++;
  num_of_buckets = 1UL << num_of_bits;
  bucket_size = (nkmers + num_of_buckets - 1) / num_of_buckets;
  if(bucket_size < 1) bucket_size = 1;
  if(nbkts) *nbkts = num_of_buckets;
  if(bktsize) *bktsize = (uint8_t)bucket_size;
  return num_of_buckets * bucket_size;
}
static size_t orc_ht_mem(size_t bktsize, size_t nbkts, size_t nbits) { return (bktsize*nbkts*nbits)/8 + nbkts*2; } /* hash_mem.h:12-14 */
/* src/basic/hash_mem.c:27-51 */
uint64_t orc_hash_table_mem_limit(size_t memlimit, size_t entrybits, uint64_t *nkmers_ptr)
{
  size_t bktsize, num_of_bits = 10, num_of_buckets = 1UL << num_of_bits, num_of_kmers;
  while(orc_ht_mem(48, num_of_buckets, entrybits) < memlimit) { num_of_bits++; num_of_buckets = 1UL << num_of_bits; }
  bktsize = (memlimit - num_of_buckets*2) / ((num_of_buckets * entrybits) / 8);
  if(bktsize == 0) {
    num_of_bits--; num_of_buckets = 1UL << num_of_bits;
    num_of_kmers = bktsize * num_of_buckets;
    bktsize = num_of_kmers / num_of_buckets; if(bktsize < 1) bktsize = 1;
  }
  if(bktsize > 48) bktsize = 48;
  if(nkmers_ptr) *nkmers_ptr = num_of_buckets * bktsize;
  return orc_ht_mem(bktsize, num_of_buckets, entrybits);
}

/* ------------------------------------------------------------------ row K */
/* In-memory restatement of libs/seq_file/seq_file.h:245-323 (FASTQ / FASTA /
 * plain record readers and the first-byte format sniff) over a whole file
 * that has been inflated by zlib's gzread (seq_file.h:512,573-581: gzopen
 * reads plain and gzip transparently).  Calls cb(seq,len,qual,qlen,ctx) per read. */
typedef void (*orc_read_cb)(const char *seq, size_t seqlen, const char *qual, size_t quallen, void *ctx);

static const char *orc_cur_name = ""; /* name line of the record being delivered (without '@' / '>') */
typedef struct { const char *p, *end; } OrcCur;
static int orc_getc(OrcCur *c) { return c->p < c->end ? (unsigned char)*c->p++ : -1; }
/* append the rest of the current line (including '\n') to dst; returns bytes read */
static size_t orc_readline(OrcCur *c, char **dst, size_t *len, size_t *cap)
{
  const char *s = c->p;
  while(c->p < c->end && *c->p != '\n') c->p++;
  if(c->p < c->end) c->p++;
  size_t n = (size_t)(c->p - s);
  if(*len + n + 1 > *cap) { *cap = (*len + n + 1) * 2; *dst = (char*)realloc(*dst, *cap); }
  memcpy(*dst + *len, s, n); *len += n; (*dst)[*len] = 0;
  return n;
}
static void orc_chomp(char *b, size_t *len) { while(*len && (b[*len-1] == '\n' || b[*len-1] == '\r')) (*len)--; b[*len] = 0; }
static void orc_pushc(char **dst, size_t *len, size_t *cap, char ch)
{
  if(*len + 2 > *cap) { *cap = (*len + 2) * 2; *dst = (char*)realloc(*dst, *cap); }
  (*dst)[(*len)++] = ch; (*dst)[*len] = 0;
}
static int orc_isspace(int c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

/* returns number of reads, or -(reads+1) on a malformed record (the reference
 * stops at the first malformed record, seq_reader.c:445-452).
 *
 * How the reference really reads records [probed against the compiled reference]: through _read_unknown
 * (seq_file.h:311-323) -- skip white space, let the first other byte pick FASTQ / FASTA / plain for THIS record -- as
 * long as sf->readfunc is that function: for the reads worth the first 1000 bases that seq_get_qual_limits buffers when
 * the FASTQ offset is to be guessed (seq_file.h:359-377,636-660, seq_reader.c:436; afterwards _seq_read_pop installs the
 * reader of the last record's format, :337), and for every read when an offset was given (`lookahead` = 0).  And in
 * _read_unknown the buffered readers are instantiated with the UNBUFFERED skipline (seq_file.h:426-427): after a
 * white-space byte other than '\n' it is not the rest of the current line that goes but a line of the FILE, at the
 * position the 1 MB stream buffer (DEFAULT_BUFSIZE) has read up to -- in the coordinates of the stream after such
 * deletions, the next multiple of 2^20 behind the byte; nothing if the stream ends before that. */
long orc_parse_buffer_la(const char *buf, size_t n, int lookahead, orc_read_cb cb, void *ctx)
{
  char *w = (char*)malloc(n + 1);
  size_t wn = n, la_bases = 0;
  memcpy(w, buf, n);
  OrcCur cur = { w, w + wn };
  char *name = NULL, *seq = NULL, *qual = NULL;
  size_t nl = 0, nc = 0, sl = 0, sc = 0, ql = 0, qc = 0;
  long nreads = 0; int c, fmt = 0, bad = 0, unknown = 1; /* fmt: 1 fastq 2 fasta 3 plain */
  orc_pushc(&name, &nl, &nc, 0); orc_pushc(&seq, &sl, &sc, 0); orc_pushc(&qual, &ql, &qc, 0);

  for(;;) {
    nl = sl = ql = 0; name[0] = seq[0] = qual[0] = 0;
    if(unknown) {
      while((c = orc_getc(&cur)) != -1 && orc_isspace(c)) if(c != '\n') {
        size_t p = (size_t)(cur.p - w), at = (((p - 1) >> 20) + 1) << 20;
        if(at < wn) {
          const char *e = (const char*)memchr(w + at, '\n', wn - at);
          size_t to = e ? (size_t)(e - w) + 1 : wn;
          memmove(w + at, w + to, wn - to);
          wn -= to - at; cur.end = w + wn;
        }
      }
      if(c == -1) break;
      fmt = c == '@' ? 1 : (c == '>' ? 2 : 3);
      cur.p--;
    }
    if(fmt == 1) { /* seq_file.h:245-272 */
      c = orc_getc(&cur);
      if(c == -1) break;
      if(c != '@' || orc_readline(&cur, &name, &nl, &nc) == 0) { bad = 1; break; }
      orc_chomp(name, &nl);
      while((c = orc_getc(&cur)) != '+') {
        if(c == -1) { bad = 1; break; }
        if(c != '\r' && c != '\n') {
          orc_pushc(&seq, &sl, &sc, (char)c);
          if(orc_readline(&cur, &seq, &sl, &sc) == 0) { bad = 1; break; }
          orc_chomp(seq, &sl);
        }
      }
      if(bad) break;
      while((c = orc_getc(&cur)) != -1 && c != '\n') {}
      if(c == -1) { bad = 1; break; }
      int eof_in_qual = 0;
      do {
        if(orc_readline(&cur, &qual, &ql, &qc) > 0) orc_chomp(qual, &ql);
        else { eof_in_qual = 1; break; }
      } while(ql < sl);
      if(!eof_in_qual) { while((c = orc_getc(&cur)) != -1 && c != '@') {} if(c != -1) cur.p--; }
    } else if(fmt == 2) { /* seq_file.h:274-295 */
      c = orc_getc(&cur);
      if(c == -1) break;
      if(c != '>' || orc_readline(&cur, &name, &nl, &nc) == 0) { bad = 1; break; }
      orc_chomp(name, &nl);
      while((c = orc_getc(&cur)) != '>') {
        if(c == -1) break;
        if(c != '\r' && c != '\n') {
          orc_pushc(&seq, &sl, &sc, (char)c);
          size_t nread = orc_readline(&cur, &seq, &sl, &sc);
          orc_chomp(seq, &sl);
          if(nread == 0) break;
        }
      }
      if(c == '>') cur.p--;
    } else { /* plain, seq_file.h:298-309 */
      while((c = orc_getc(&cur)) != -1 && orc_isspace(c)) if(c != '\n') { while((c = orc_getc(&cur)) != -1 && c != '\n') {} }
      if(c == -1) break;
      orc_pushc(&seq, &sl, &sc, (char)c);
      orc_readline(&cur, &seq, &sl, &sc);
      orc_chomp(seq, &sl);
    }
    orc_cur_name = name;
    cb(seq, sl, ql ? qual : NULL, ql, ctx);
    nreads++;
    if(unknown && lookahead) { la_bases += sl; if(la_bases >= 1000) unknown = 0; }
  }
  free(name); free(seq); free(qual); free(w);
  return bad ? -(nreads + 1) : nreads;
}
long orc_parse_buffer(const char *buf, size_t n, orc_read_cb cb, void *ctx) { return orc_parse_buffer_la(buf, n, 1, cb, ctx); }

typedef struct { OrcGraph *g; size_t colour; uint8_t fq_cutoff, fq_offset, hp_cutoff; OrcStats *st; } OrcLoadCtx;
static void orc_load_cb(const char *seq, size_t sl, const char *qual, size_t ql, void *ctx)
{
  OrcLoadCtx *c = (OrcLoadCtx*)ctx;
  orc_graph_add_read(c->g, seq, sl, qual, ql, c->colour, c->fq_cutoff, c->fq_offset, c->hp_cutoff, c->st);
}

/* libs/seq_file/seq_file.h:636-682 restated over the first reads of a parsed
 * file: min/max over the first <=1000 quality bytes seen while seq bases < 1000. */
typedef struct { int min, max; size_t count, qcount; } OrcQLim;
static void orc_qlim_cb(const char *seq, size_t sl, const char *qual, size_t ql, void *ctx)
{
  OrcQLim *q = (OrcQLim*)ctx; size_t limit = 1000, len, i;
  (void)seq;
  if(q->count >= limit) return;
  len = ql < limit - q->qcount ? ql : limit - q->qcount;
  for(i = 0; i < len; i++) { if(qual[i] > q->max) q->max = qual[i]; if(qual[i] < q->min) q->min = qual[i]; }
  q->count += sl; q->qcount += ql;
}
/* returns FASTQ ascii offset (33/64) or 0 if the file has no qualities */
int orc_guess_fq_offset(const char *buf, size_t n)
{
  static const int OFFS[6] = {33, 33, 64, 64, 64, 33}; /* seq_file.h:127 */
  OrcQLim q = { 0x7fffffff, 0, 0, 0 };
  int fmt;
  orc_parse_buffer(buf, n, orc_qlim_cb, &q);
  if(q.qcount == 0) return 0;
  if(q.min >= 33 && q.max <= 73) fmt = 1;
  else if(q.min >= 33 && q.max <= 75) fmt = 5;
  else if(q.min >= 67 && q.max <= 105) fmt = 4;
  else if(q.min >= 64 && q.max <= 105) fmt = 3;
  else if(q.min >= 59 && q.max <= 105) fmt = 2;
  else fmt = 0;
  return OFFS[fmt];
}

static char *orc_slurp(const char *path, size_t *n)
{
  gzFile gz = gzopen(path, "r");
  size_t cap = 1 << 20, len = 0; int r;
  char *buf;
  if(!gz) return NULL;
  buf = (char*)malloc(cap);
  while((r = gzread(gz, buf + len, (unsigned)(cap - len))) > 0) {
    len += (size_t)r;
    if(len == cap) { cap *= 2; buf = (char*)realloc(buf, cap); }
  }
  gzclose(gz);
  *n = len;
  return buf;
}

/* One `--seq <file>` task (src/basic/seq_reader.c:421-462 seq_parse_se_sf ->
 * build_graph.c:233-254).  fq_offset 0 = auto-detect.  Stats accumulate into *st. */
long orc_graph_load_file(OrcGraph *g, const char *path, size_t colour,
                         uint8_t fq_cutoff, uint8_t fq_offset, uint8_t hp_cutoff, OrcStats *st)
{
  size_t n; long r;
  char *buf = orc_slurp(path, &n);
  if(!buf) return -1000000000L;
  const int lookahead = fq_offset == 0; /* reads are only buffered for the offset guess when none was given */
  if(fq_offset == 0) fq_offset = (uint8_t)orc_guess_fq_offset(buf, n);
  OrcLoadCtx c = { g, colour, fq_cutoff, fq_offset, hp_cutoff, st };
  r = orc_parse_buffer_la(buf, n, lookahead, orc_load_cb, &c);
  free(buf);
  return r;
}

/* ---- build --remove-pcr over files (row N3) ------------------------------------------------ */
typedef struct { char *name, *seq, *qual; size_t sl, ql; } OrcRead;
typedef struct { OrcRead *r; size_t n, cap; } OrcReadVec;
static void orc_collect_cb(const char *seq, size_t sl, const char *qual, size_t ql, void *ctx)
{
  OrcReadVec *v = (OrcReadVec*)ctx;
  if(v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 1024; v->r = (OrcRead*)realloc(v->r, v->cap * sizeof(OrcRead)); }
  OrcRead *r = &v->r[v->n++];
  r->name = (char*)malloc(strlen(orc_cur_name) + 1); strcpy(r->name, orc_cur_name);
  r->seq = (char*)malloc(sl + 1); memcpy(r->seq, seq, sl); r->seq[sl] = 0; r->sl = sl;
  r->qual = (char*)malloc(ql + 1); if(ql) memcpy(r->qual, qual, ql); r->qual[ql] = 0; r->ql = ql;
}
static void orc_readvec_free(OrcReadVec *v)
{
  size_t i;
  for(i = 0; i < v->n; i++) { free(v->r[i].name); free(v->r[i].seq); free(v->r[i].qual); }
  free(v->r);
}
static int orc_name_end(unsigned char c) { return !c || orc_isspace(c); }
/* seq_read_names_cmp, libs/seq_file/seq_file.h:783-800: 0 when the names agree up to the first white
 * space, or differ only in a trailing /1 vs /2 */
static int orc_names_cmp(const char *aa, const char *bb)
{
  const unsigned char *a = (const unsigned char*)aa, *b = (const unsigned char*)bb, *a0 = a, *b0 = b;
  while(*a && *b && *a == *b && !orc_isspace(*a)) { a++; b++; }
  if(a > a0 && b > b0 && a[-1] == '/' && b[-1] == '/' && ((*a == '1' && *b == '2') || (*a == '2' && *b == '1')) &&
     orc_name_end(a[1]) && orc_name_end(b[1])) return 0;
  return orc_name_end(*a) && orc_name_end(*b) ? 0 : (int)*a - (int)*b;
}

/* One --seq / --seq2 / --seqi task with --remove-pcr in force: seq_parse_se_sf (seq_reader.c:421-462),
 * seq_parse_pe_sf (:357-419: pairs until either file ends) or seq_parse_interleaved_sf (:289-355:
 * consecutive reads whose names match are a pair, the rest single-end), every read / pair through
 * orc_graph_add_reads_pcr in file order.  fq_offset 0 = auto-detect per file.  mode: 0 se, 1 pe (path2), 2 interleaved.
 * Returns the number of reads consumed or -1000000000 if a file cannot be opened. */
long orc_graph_load_pcr(OrcGraph *g, const char *path1, const char *path2, int mode, size_t colour,
                        uint8_t fq_cutoff, uint8_t fq_offset, uint8_t hp_cutoff, int matedir, OrcStats *st)
{
  size_t n1 = 0, n2 = 0, i; long used = 0;
  char *b1 = orc_slurp(path1, &n1), *b2 = NULL;
  OrcReadVec v1 = {0, 0, 0}, v2 = {0, 0, 0};
  uint8_t off1 = fq_offset, off2 = fq_offset;
  if(!b1) return -1000000000L;
  if(mode == 1) { b2 = orc_slurp(path2, &n2); if(!b2) { free(b1); return -1000000000L; } }
  if(fq_offset == 0) { off1 = (uint8_t)orc_guess_fq_offset(b1, n1); if(b2) off2 = (uint8_t)orc_guess_fq_offset(b2, n2); }
  orc_parse_buffer_la(b1, n1, fq_offset == 0, orc_collect_cb, &v1);
  if(b2) orc_parse_buffer_la(b2, n2, fq_offset == 0, orc_collect_cb, &v2);
  if(mode == 0) {
    for(i = 0; i < v1.n; i++, used++)
      orc_graph_add_reads_pcr(g, v1.r[i].seq, v1.r[i].sl, v1.r[i].qual, v1.r[i].ql, NULL, 0, NULL, 0,
                              colour, fq_cutoff, off1, 0, hp_cutoff, matedir, st);
  } else if(mode == 1) {
    for(i = 0; i < v1.n && i < v2.n; i++, used += 2)
      orc_graph_add_reads_pcr(g, v1.r[i].seq, v1.r[i].sl, v1.r[i].qual, v1.r[i].ql, v2.r[i].seq, v2.r[i].sl, v2.r[i].qual, v2.r[i].ql,
                              colour, fq_cutoff, off1, off2, hp_cutoff, matedir, st);
  } else {
    for(i = 0; i < v1.n; ) {
      OrcRead *a = &v1.r[i], *b = i + 1 < v1.n ? &v1.r[i + 1] : NULL;
      if(b && orc_names_cmp(a->name, b->name) == 0) {
        orc_graph_add_reads_pcr(g, a->seq, a->sl, a->qual, a->ql, b->seq, b->sl, b->qual, b->ql,
                                colour, fq_cutoff, off1, off1, hp_cutoff, matedir, st);
        i += 2; used += 2;
      } else {
        orc_graph_add_reads_pcr(g, a->seq, a->sl, a->qual, a->ql, NULL, 0, NULL, 0, colour, fq_cutoff, off1, 0, hp_cutoff, matedir, st);
        i += 1; used += 1;
      }
    }
  }
  orc_readvec_free(&v1); orc_readvec_free(&v2); free(b1); free(b2);
  return used;
}
