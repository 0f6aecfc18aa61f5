/* ctx_reader.c -- read a CORTEX graph file (.ctx, versions 4-7) and merge it into the device graph.
 *
 * Replaces, for `build --graph` (and --intersect), relative to the reference root:
 *   graph_file_open2 / graph_file_read_header      src/graph/graph_file_reader.c:78-335
 *   file_filter_open / file_filter_set_cols        src/basic/file_filter.c:9-153   ("[into:]path[:from]")
 *   range_get_num / range_parse_array[_fill]       src/basic/range.c
 *   graph_load_ginfo / graph_load                  src/graph/graphs_load.c:45-208
 * The records themselves are merged on the GPU (mcx_graph_load_records); this file parses the
 * header, applies the colour filter to the header metadata and streams the records.
 */
#include "mcx_host.h"
#include <errno.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#define LOAD_CHUNK_BYTES (64u << 20)

/* ---- range.c ------------------------------------------------------------------------- */
static int range_parse(const char *range_str, size_t *start, size_t *end, size_t range_max)
{
  char *endptr; const char *str = range_str; unsigned long from, to;
  if(*str == '*') { *start = 0; *end = range_max; return 1; }
  from = strtoul(str, &endptr, 10); to = from;
  if(endptr == str) return -1;
  if(*endptr == '-') { str = endptr + 1; to = strtoul(str, &endptr, 10); if(endptr == str) return -1; }
  if(from > range_max || to > range_max) return -1;
  *start = from; *end = to;
  return (int)(endptr - range_str);
}
static int range_get_num(const char *str, size_t range_max)
{
  const char *ptr = str; size_t start, end, n = 0; int bytes;
  while(*ptr != '\0') {
    if((bytes = range_parse(ptr, &start, &end, range_max)) == -1) return -1;
    ptr += bytes;
    n += (start > end ? start - end : end - start) + 1;
    if(*ptr == ',') ptr++;
  }
  return n == 0 ? (int)(range_max + 1) : (int)n;
}
/* A backward range a-b (a > b) does not stop at b in the reference: its loop is `for(j = start; j <= start; j--)`
 * (src/basic/range.c:67-68), so it emits a, a-1, ..., 0 -- more entries than range_get_num counted, which shifts what
 * follows (a colour filter uses the first range_get_num() entries, file_filter.c:119-122; an `into` list of the wrong
 * length is an error, range.c:107-109) and overruns the reference's scratch array.  Same entries here, in an array
 * that is large enough (range_emitted).  [found by tests/test_host_fuzz.py against the compiled reference] */
static size_t range_emitted(const char *str, size_t range_max)
{
  const char *ptr = str; size_t n = 0, start, end; int bytes;
  while(*ptr != '\0') {
    if((bytes = range_parse(ptr, &start, &end, range_max)) == -1) return n;
    ptr += bytes;
    if(*ptr == ',') ptr++;
    n += start <= end ? end - start + 1 : start + 1;
  }
  return n;
}
static int range_parse_array(const char *str, size_t *arr, size_t range_max)
{
  const char *ptr = str; size_t n = 0, j, start, end; int bytes;
  while(*ptr != '\0') {
    if((bytes = range_parse(ptr, &start, &end, range_max)) == -1) return -1;
    ptr += bytes;
    if(*ptr == ',') ptr++;
    if(start <= end) for(j = start; j <= end; j++) arr[n++] = j;
    else for(j = start; j <= start; j--) arr[n++] = j;
  }
  if(ptr > str && *(ptr - 1) == ',') return -1;
  if(n == 0) for(n = 0; n <= range_max; n++) arr[n] = n;
  return (int)n;
}
static int range_parse_array_fill(const char *str, size_t *arr, size_t range_max, size_t num_entries)
{
  size_t i; int r = range_parse_array(str, arr, range_max);
  if(r < 0) return -1;
  else if(r == 0) for(i = 0; i < num_entries; i++) arr[i] = i;
  else if(r == 1) for(i = 1; i < num_entries; i++) arr[i] = arr[0];
  else if((size_t)r != num_entries) return -1;
  return (int)num_entries;
}

/* ---- file_filter.c --------------------------------------------------------------------- */
#define is_range_char(c) (((c) >= '0' && (c) <= '9') || (c) == '-' || (c) == ',')
static void deconstruct_path(const char *path, const char **start, const char **end)
{
  const char *ptr = path;
  *start = path;
  while(is_range_char(*ptr)) ptr++;
  if(ptr > path && *ptr == ':') { ptr++; *start = ptr; }
  ptr = *end = path + strlen(path);
  while(ptr > (*start) + 1) {
    ptr--;
    if(*ptr == ':') { *end = ptr; break; }
    else if(!is_range_char(*ptr)) break;
  }
}
static int cmp_filter_into(const void *a, const void *b)
{
  const uint32_t *x = a, *y = b; /* pairs {from, into} */
  return (x[1] > y[1]) - (x[1] < y[1]);
}
static void filter_set_cols(McxCtxFile *f, size_t srcncols, size_t into_offset)
{
  const char *ps, *pe;
  deconstruct_path(f->input, &ps, &pe);
  char *from_fltr = (*pe == ':') ? strdup(pe + 1) : NULL;
  char *into_fltr = NULL;
  if(ps > f->input) { size_t n = (size_t)(ps - 1 - f->input); into_fltr = malloc(n + 1); memcpy(into_fltr, f->input, n); into_fltr[n] = 0; }
  size_t ncols, i;
  if(from_fltr) {
    int s = range_get_num(from_fltr, srcncols - 1);
    if(s < 0) mcx_die("Invalid filter path: %s (from size: %zu)", f->input, srcncols);
    ncols = (size_t)s;
  } else ncols = srcncols;
  if(into_fltr) {
    int s = range_get_num(into_fltr, SIZE_MAX);
    if(s < 0 || (s != 1 && (size_t)s != ncols)) mcx_die("Invalid filter path: %s (s:%i ncols:%zu)", f->input, s, ncols);
  }
  size_t tmp_n = ncols > srcncols ? ncols : srcncols, e;
  if(from_fltr && (e = range_emitted(from_fltr, srcncols - 1)) > tmp_n) tmp_n = e;
  if(into_fltr && (e = range_emitted(into_fltr, SIZE_MAX)) > tmp_n) tmp_n = e;
  if(tmp_n > ((size_t)1 << 28)) mcx_die("Invalid filter path: %s", f->input);
  size_t *tmp = calloc(tmp_n + 1, sizeof(size_t));
  if(!tmp) mcx_die("Out of memory");
  uint32_t *pairs = calloc(ncols, 2 * sizeof(uint32_t));
  if(from_fltr) {
    if(range_parse_array(from_fltr, tmp, srcncols - 1) == -1) mcx_die("Invalid filter path: %s", f->input);
    for(i = 0; i < ncols; i++) pairs[2 * i] = (uint32_t)tmp[i];
  } else for(i = 0; i < ncols; i++) pairs[2 * i] = (uint32_t)i;
  if(into_fltr) {
    int s = range_parse_array_fill(into_fltr, tmp, SIZE_MAX, ncols);
    if(s < 0 || (size_t)s != ncols) mcx_die("Invalid filter path: %s (s:%i ncols:%zu)", f->input, s, ncols);
    for(i = 0; i < ncols; i++) pairs[2 * i + 1] = (uint32_t)tmp[i];
  } else for(i = 0; i < ncols; i++) pairs[2 * i + 1] = (uint32_t)(into_offset + i);
  /* the reference sorts by into with qsort (filters_sort_by_into); ties keep no defined order there either,
   * and nothing below depends on the order of equal `into` entries */
  qsort(pairs, ncols, 2 * sizeof(uint32_t), cmp_filter_into);
  f->nfilter = (uint32_t)ncols;
  f->from_col = malloc(ncols * sizeof(uint32_t)); f->into_col = malloc(ncols * sizeof(uint32_t));
  f->into_ncols = 0;
  for(i = 0; i < ncols; i++) {
    f->from_col[i] = pairs[2 * i]; f->into_col[i] = pairs[2 * i + 1];
    if(f->into_col[i] + 1 > f->into_ncols) f->into_ncols = f->into_col[i] + 1;
  }
  free(pairs); free(tmp); free(from_fltr); free(into_fltr);
}

/* ---- graph_file_reader.c ---------------------------------------------------------------- */
static void gfread(McxCtxFile *f, void *ptr, size_t n, const char *what)
{
  size_t got = fread(ptr, 1, n, f->fh);
  if(got != n) mcx_die("Couldn't read '%s': expected %zu; recieved: %zu; [file: %s]\n", what, n, got, f->path);
}
static char *read_str(McxCtxFile *f, const char *what, size_t idx, size_t *bytes_read)
{
  uint32_t len;
  gfread(f, &len, 4, what);
  if(len > 10000) mcx_die("Very big sample name. Length: %u", len);
  char *s = malloc((size_t)len + 1);
  gfread(f, s, len, what);
  s[len] = '\0';
  if(strlen(s) != len)
    mcx_warn("Sample %zu name has length %u but is only %zu chars long (premature '\\0') [path: %s]\n", idx, len, strlen(s), f->path);
  *bytes_read += 4 + len;
  return s;
}
static size_t read_header(McxCtxFile *f)
{
  size_t i, bytes_read = 0;
  char magic[7]; magic[6] = '\0';
  gfread(f, magic, 6, "Magic word");
  if(strcmp(magic, "CORTEX") != 0) mcx_die("Magic word doesn't match '%s' (start): %s", "CORTEX", f->path);
  bytes_read += 6;
  gfread(f, &f->version, 4, "graph version"); gfread(f, &f->kmer_size, 4, "kmer size");
  gfread(f, &f->num_of_bitfields, 4, "num of bitfields"); gfread(f, &f->num_of_cols, 4, "number of colours");
  bytes_read += 16;
  if(f->version > 7 || f->version < 4)
    mcx_die("Sorry, we only support graph file versions 4, 5, 6 & 7 [version: %u; path: %s]\n", f->version, f->path);
  if(f->kmer_size % 2 == 0) mcx_die("kmer size is not an odd number [kmer_size: %u; path: %s]\n", f->kmer_size, f->path);
  if(f->kmer_size < 3) mcx_die("kmer size is less than three [kmer_size: %u; path: %s]\n", f->kmer_size, f->path);
  if(f->num_of_bitfields * 32 < f->kmer_size)
    mcx_die("Not enough bitfields for kmer size [kmer_size: %u; bitfields: %u; path: %s]\n", f->kmer_size, f->num_of_bitfields, f->path);
  if((f->num_of_bitfields - 1) * 32 >= f->kmer_size) mcx_die("using more than the minimum number of bitfields [path: %s]\n", f->path);
  if(f->num_of_cols == 0) mcx_die("number of colours is zero [path: %s]\n", f->path);
  if(f->num_of_cols > 10000) mcx_die("Very high number of colours: %zu [path: %s]", (size_t)f->num_of_cols, f->path);
  f->ginfo = calloc(f->num_of_cols, sizeof(McxGInfo));
  for(i = 0; i < f->num_of_cols; i++) mcx_ginfo_init(&f->ginfo[i]);
  for(i = 0; i < f->num_of_cols; i++) gfread(f, &f->ginfo[i].mean_read_length, 4, "mean read length for each colour");
  for(i = 0; i < f->num_of_cols; i++) gfread(f, &f->ginfo[i].total_sequence, 8, "total sequance loaded for each colour");
  bytes_read += f->num_of_cols * 12u;
  if(f->version >= 6) {
    for(i = 0; i < f->num_of_cols; i++) {
      free(f->ginfo[i].sample_name);
      f->ginfo[i].sample_name = read_str(f, "sample name", i, &bytes_read);
    }
    f->seq_err_raw = calloc(f->num_of_cols, 16);
    f->clean_flags_raw = calloc(f->num_of_cols, 4);
    for(i = 0; i < f->num_of_cols; i++) {
      gfread(f, f->seq_err_raw[i], 16, "seq error rates");
      memcpy(&f->ginfo[i].seq_err, f->seq_err_raw[i], 10);
    }
    bytes_read += sizeof(long double) * f->num_of_cols;
    for(i = 0; i < f->num_of_cols; i++) {
      McxCleaning *c = &f->ginfo[i].cleaning;
      uint8_t fl[4]; uint32_t thr_unitigs = 0, thr_kmers = 0;
      gfread(f, fl, 4, "cleaning flags");
      memcpy(f->clean_flags_raw[i], fl, 4); /* the header goes back out byte for byte (`sort`), even a flag that is not 0 / 1 */
      c->cleaned_tips = fl[0]; c->cleaned_unitigs = fl[1]; c->cleaned_kmers = fl[2]; c->is_graph_intersection = fl[3];
      gfread(f, &thr_unitigs, 4, "remove low covg unitig threshold");
      gfread(f, &thr_kmers, 4, "remove low covg kmer threshold");
      bytes_read += 12;
      if(f->version <= 6) {
        if(!c->cleaned_unitigs && thr_unitigs == (uint32_t)-1) thr_unitigs = 0;
        if(!c->cleaned_kmers && thr_kmers == (uint32_t)-1) thr_kmers = 0;
      }
      if(!c->cleaned_unitigs && thr_unitigs > 0) {
        mcx_warn("Graph header gives cleaning threshold for unitig when no cleaning was performed [path: %s]", f->path);
        thr_unitigs = 0;
      }
      if(!c->cleaned_kmers && thr_kmers > 0) {
        mcx_warn("Graph header gives cleaning threshold for nodes when no cleaning was performed [path: %s]", f->path);
        thr_kmers = 0;
      }
      c->clean_unitigs_thresh = thr_unitigs; c->clean_kmers_thresh = thr_kmers;
      free(c->intersection_name);
      c->intersection_name = read_str(f, "cleaned against graph name", i, &bytes_read);
    }
  }
  gfread(f, magic, 6, "magic word (end)");
  if(strcmp(magic, "CORTEX") != 0) mcx_die("Magic word doesn't match '%s' (end): '%s' [path: %s]\n", "CORTEX", magic, f->path);
  bytes_read += 6;
  return bytes_read;
}

McxCtxFile *mcx_ctx_open(const char *input, size_t into_offset)
{
  McxCtxFile *f = calloc(1, sizeof(*f));
  const char *ps, *pe;
  f->input = strdup(input);
  deconstruct_path(input, &ps, &pe);
  f->path = malloc((size_t)(pe - ps) + 1);
  memcpy(f->path, ps, (size_t)(pe - ps)); f->path[pe - ps] = '\0';
  f->file_size = -1; f->num_of_kmers = -1;
  struct stat st;
  if(strcmp(f->path, "-") != 0) {
    if(stat(f->path, &st) == 0) f->file_size = st.st_size;
    else mcx_warn("Couldn't get file size: %s", f->path);
  }
  f->fh = strcmp(f->path, "-") == 0 ? stdin : fopen(f->path, "r");
  if(!f->fh) mcx_die("Cannot open file: %s [%s]", f->path, strerror(errno));
  setvbuf(f->fh, NULL, _IOFBF, 1u << 20);
  f->hdr_size = read_header(f);
  filter_set_cols(f, f->num_of_cols, into_offset);
  if(f->file_size != -1) {
    size_t bytes_per_kmer = 8u * MCX_CTX_W(f) + 5u * (size_t)f->num_of_cols;
    size_t remaining = (size_t)f->file_size - f->hdr_size;
    f->num_of_kmers = (int64_t)(remaining / bytes_per_kmer);
    if(remaining % bytes_per_kmer != 0)
      mcx_warn("Truncated graph file: %s [bytes per kmer: %zu remaining: %zu; fsize: %zu; header: %zu; nkmers: %zu]",
               f->path, bytes_per_kmer, remaining, (size_t)f->file_size, f->hdr_size, (size_t)f->num_of_kmers);
  }
  return f;
}

void mcx_ctx_close(McxCtxFile *f)
{
  if(!f) return;
  if(f->fh && f->fh != stdin) fclose(f->fh);
  if(f->ginfo) { for(uint32_t i = 0; i < f->num_of_cols; i++) mcx_ginfo_free(&f->ginfo[i]); free(f->ginfo); }
  free(f->seq_err_raw); free(f->clean_flags_raw);
  free(f->from_col); free(f->into_col); free(f->input); free(f->path); free(f);
}

bool mcx_ctx_filter_is_direct(const McxCtxFile *f)
{
  if(f->nfilter != f->num_of_cols) return false;
  for(uint32_t i = 0; i < f->nfilter; i++) if(f->from_col[i] != i) return false;
  return true;
}

size_t mcx_ctx_write_header_raw(FILE *fh, const McxCtxFile *f)
{
  size_t b = 0; uint32_t i, C = f->num_of_cols;
#define PUT(p, n) do { if(fwrite((p), 1, (n), fh) != (size_t)(n)) mcx_die("Cannot write file"); b += (n); } while(0)
  PUT("CORTEX", 6);
  PUT(&f->version, 4); PUT(&f->kmer_size, 4); PUT(&f->num_of_bitfields, 4); PUT(&f->num_of_cols, 4);
  for(i = 0; i < C; i++) PUT(&f->ginfo[i].mean_read_length, 4);
  for(i = 0; i < C; i++) PUT(&f->ginfo[i].total_sequence, 8);
  if(f->version >= 6) {
    for(i = 0; i < C; i++) { uint32_t len = (uint32_t)strlen(f->ginfo[i].sample_name); PUT(&len, 4); PUT(f->ginfo[i].sample_name, len); }
    for(i = 0; i < C; i++) PUT(f->seq_err_raw[i], 16);
    for(i = 0; i < C; i++) {
      const McxCleaning *c = &f->ginfo[i].cleaning;
      uint32_t tu = c->cleaned_unitigs ? c->clean_unitigs_thresh : 0, tk = c->cleaned_kmers ? c->clean_kmers_thresh : 0;
      uint32_t len = (uint32_t)strlen(c->intersection_name);
      PUT(f->clean_flags_raw[i], 4); PUT(&tu, 4); PUT(&tk, 4); PUT(&len, 4); PUT(c->intersection_name, len);
    }
  }
  PUT("CORTEX", 6);
#undef PUT
  return b;
}

/* ---- graphs_load.c ------------------------------------------------------------------------ */
void mcx_ctx_flatten(McxCtxFile *f, uint32_t intocol)
{
  for(uint32_t i = 0; i < f->nfilter; i++) f->into_col[i] = intocol;
  f->into_ncols = intocol + 1;
}

/* graph_file_read_raw (src/graph/graph_file_reader.c:368-370): a record whose top key word has bits above 2k set */
void mcx_ctx_check_records(const McxCtxFile *f, const unsigned char *recs, size_t n)
{
  const size_t rec_bytes = 8u * MCX_CTX_W(f) + 5u * (size_t)f->num_of_cols;
  const unsigned top_bits = 2u * (f->kmer_size & 31u); /* k is odd: 2..62 */
  /* :360-361: the bytes of the key just read (sizeof(BinaryKmer)) against the header's bitfield count, as an int */
  if(n && (int)(sizeof(uint64_t) * (size_t)f->num_of_bitfields) != (int)(8u * MCX_CTX_W(f))) mcx_die("Unexpected end of file: %s", f->path);
  for(size_t i = 0; i < n; i++) {
    uint64_t w0; memcpy(&w0, recs + i * rec_bytes, 8);
    if(w0 >> top_bits) mcx_die("Oversized kmer in path [kmer: %u]: %s", f->kmer_size, f->path);
  }
}

int mcx_ctx_load(mcx_graph *g, McxCtxFile *f, McxGInfo *ginfo, size_t graph_ncols, uint32_t load_flags,
                 uint64_t *nkmers_read, uint64_t *nkmers_loaded, uint64_t *nkmers_novel)
{
  char a[64], b[64];
  /* file_filter_status + graph_loading_print_status */
  mcx_status("[FileFilter] Reading file %s [%u src colour%s]", f->path, f->num_of_cols, f->num_of_cols == 1 ? "" : "s");
  mcx_ulong_to_str(f->num_of_kmers < 0 ? 0 : (uint64_t)f->num_of_kmers, a); mcx_bytes_to_str(f->file_size < 0 ? 0 : (uint64_t)f->file_size, b);
  mcx_status("[GReader] %s kmers, %s filesize", a, b);
  /* graph_load_ginfo */
  if(f->into_ncols > graph_ncols)
    mcx_die("Program has not assigned enough colours! [colours in graph: %zu vs file: %zu; path: %s]", graph_ncols, (size_t)f->into_ncols, f->path);
  if(ginfo) for(uint32_t i = 0; i < f->nfilter; i++) mcx_ginfo_merge(&ginfo[f->into_col[i]], &f->ginfo[f->from_col[i]]);

  if(f->fh != stdin && fseek(f->fh, (long)f->hdr_size, SEEK_SET) != 0) mcx_die("fseek failed: %s", strerror(errno));
  const size_t rec_bytes = 8u * MCX_CTX_W(f) + 5u * (size_t)f->num_of_cols;
  size_t chunk_recs = LOAD_CHUNK_BYTES / rec_bytes; if(chunk_recs == 0) chunk_recs = 1;
  unsigned char *buf = malloc(chunk_recs * rec_bytes);
  uint64_t nread = 0, nloaded = 0, nnovel = 0;
  int r = MCX_OK;
  for(;;) {
    size_t got = fread(buf, 1, chunk_recs * rec_bytes, f->fh);
    if(got == 0) break;
    if(got % rec_bytes != 0) mcx_die("Unexpected end of file: %s", f->path);
    mcx_ctx_check_records(f, buf, got / rec_bytes);
    uint64_t l = 0, nv = 0;
    r = mcx_graph_load_records(g, buf, got / rec_bytes, f->num_of_cols, MCX_MEM_HOST, f->from_col, f->into_col, f->nfilter,
                               load_flags, &l, &nv);
    if(r) break;
    nread += got / rec_bytes; nloaded += l; nnovel += nv;
  }
  free(buf);
  if(r == MCX_OK && f->num_of_kmers >= 0 && nread != (uint64_t)f->num_of_kmers)
    mcx_warn("%s kmers in the graph file than expected [exp: %zu; act: %zu; path: %s]",
             nread > (uint64_t)f->num_of_kmers ? "More" : "Fewer", (size_t)f->num_of_kmers, (size_t)nread, f->path);
  mcx_ulong_to_str(nloaded, a); mcx_ulong_to_str(nread, b);
  mcx_status("[GReader] Loaded %s / %s kmers", a, b);
  if(nkmers_read) *nkmers_read = nread;
  if(nkmers_loaded) *nkmers_loaded = nloaded;
  if(nkmers_novel) *nkmers_novel = nnovel;
  return r;
}
