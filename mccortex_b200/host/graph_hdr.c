/* graph_hdr.c -- per-colour header metadata and the .ctx v6 header bytes.
 *
 * Bit-exactness notes: mean_read_length goes through the
 * same lossy double round trips as src/basic/graph_info.c:116-170, and seq_err is the raw
 * x87 long double the reference fwrites (src/graph/graph_writer.c:93-94): 10 value bytes +
 * 6 zero padding bytes on x86-64.  The expressions below are written in the reference's
 * order of evaluation on purpose; x86-64 gcc gives the same bits. */
#include "mcx_host.h"
#include <stdlib.h>
#include <string.h>

/* src/basic/graph_info.c:4-10,60-67 */
void mcx_ginfo_init(McxGInfo *g)
{
  g->mean_read_length = 0; g->total_sequence = 0; g->seq_err = 0.01;
  g->sample_name = strdup("undefined");
  memset(&g->cleaning, 0, sizeof(g->cleaning));
  g->cleaning.intersection_name = strdup("undefined");
}
void mcx_ginfo_free(McxGInfo *g)
{
  free(g->sample_name); g->sample_name = NULL;
  free(g->cleaning.intersection_name); g->cleaning.intersection_name = NULL;
}
static char *str_append(char *dst, const char *sep, const char *src)
{
  size_t n = strlen(dst) + strlen(sep) + strlen(src) + 1;
  char *s = malloc(n);
  strcpy(s, dst); strcat(s, sep); strcat(s, src);
  free(dst);
  return s;
}
/* graph_info_append_intersect, src/basic/graph_info.c:88-101 */
static void cleaning_append_intersect(McxCleaning *c, const char *name)
{
  if(!c->is_graph_intersection) { free(c->intersection_name); c->intersection_name = strdup(name); }
  else c->intersection_name = str_append(c->intersection_name, ",", name);
  c->is_graph_intersection = true;
}
/* error_cleaning_merge, src/basic/graph_info.c:34-58 */
static void cleaning_merge(McxCleaning *dst, const McxCleaning *src)
{
  dst->cleaned_tips |= src->cleaned_tips;
  dst->cleaned_unitigs |= src->cleaned_unitigs;
  dst->cleaned_kmers |= src->cleaned_kmers;
  if(src->clean_unitigs_thresh > 0 && (dst->clean_unitigs_thresh == 0 || src->clean_unitigs_thresh < dst->clean_unitigs_thresh))
    dst->clean_unitigs_thresh = src->clean_unitigs_thresh;
  if(src->clean_kmers_thresh > 0 && (dst->clean_kmers_thresh == 0 || src->clean_kmers_thresh < dst->clean_kmers_thresh))
    dst->clean_kmers_thresh = src->clean_kmers_thresh;
  if(src->is_graph_intersection) cleaning_append_intersect(dst, src->intersection_name);
  dst->is_graph_intersection |= src->is_graph_intersection;
}
void mcx_ginfo_set_name(McxGInfo *g, const char *name) { free(g->sample_name); g->sample_name = strdup(name); }

void mcx_ginfo_update_contigs(McxGInfo *g, uint64_t added_seq, uint64_t num_contigs)
{
  if(!added_seq && !num_contigs) return;
  size_t have = 0;
  if(g->total_sequence && g->mean_read_length)
    have = ((double)g->total_sequence / g->mean_read_length) + 0.5;
  if(have + num_contigs > 0)
    g->mean_read_length = (uint32_t)((double)(g->total_sequence + added_seq) / (have + num_contigs));
  g->total_sequence += added_seq;
}

void mcx_ginfo_merge(McxGInfo *dst, const McxGInfo *src)
{
  if(strcmp(src->sample_name, "undefined") != 0) {
    if(strcmp(dst->sample_name, "undefined") == 0) mcx_ginfo_set_name(dst, src->sample_name);
    else dst->sample_name = str_append(dst->sample_name, ",", src->sample_name);
  }
  uint64_t total = dst->total_sequence + src->total_sequence;
  if(total > 0) {
    dst->seq_err = (dst->seq_err * dst->total_sequence + src->seq_err * src->total_sequence) / total;
    size_t src_contigs = 0;
    if(src->total_sequence && src->mean_read_length)
      src_contigs = ((double)src->total_sequence / src->mean_read_length) + 0.5;
    mcx_ginfo_update_contigs(dst, src->total_sequence, src_contigs);
  }
  cleaning_merge(&dst->cleaning, &src->cleaning);
  dst->total_sequence = total;
}

static size_t put(FILE *fh, const void *p, size_t n)
{
  if(fwrite(p, 1, n, fh) != n) mcx_die("Cannot write to file");
  return n;
}

size_t mcx_write_ctx_header(FILE *fh, uint32_t kmer_size, uint32_t ncols, const McxGInfo *ginfo)
{
  uint32_t i;
  McxGInfo *h = calloc(ncols, sizeof(McxGInfo));
  for(i = 0; i < ncols; i++) { mcx_ginfo_init(&h[i]); mcx_ginfo_merge(&h[i], &ginfo[i]); }
  size_t b = mcx_write_ctx_header_as_is(fh, kmer_size, ncols, h, 0);
  for(i = 0; i < ncols; i++) mcx_ginfo_free(&h[i]);
  free(h);
  return b;
}

/* graph_write_header (src/graph/graph_writer.c:62-110) of a header whose colours are h[] as they stand */
size_t mcx_write_ctx_header_as_is(FILE *fh, uint32_t kmer_size, uint32_t ncols, const McxGInfo *h, uint32_t nbitfields)
{
  size_t b = 0; uint32_t i;
  uint32_t version = 6, W = nbitfields ? nbitfields : (kmer_size + 31) / 32;
  b += put(fh, "CORTEX", 6);
  b += put(fh, &version, 4); b += put(fh, &kmer_size, 4); b += put(fh, &W, 4); b += put(fh, &ncols, 4);
  for(i = 0; i < ncols; i++) b += put(fh, &h[i].mean_read_length, 4);
  for(i = 0; i < ncols; i++) b += put(fh, &h[i].total_sequence, 8);
  for(i = 0; i < ncols; i++) {
    uint32_t len = (uint32_t)strlen(h[i].sample_name);
    b += put(fh, &len, 4); b += put(fh, h[i].sample_name, len);
  }
  for(i = 0; i < ncols; i++) {
    unsigned char ld[sizeof(long double)];
    memset(ld, 0, sizeof(ld)); memcpy(ld, &h[i].seq_err, 10);
    b += put(fh, ld, sizeof(ld));
  }
  for(i = 0; i < ncols; i++) {
    /* write_error_cleaning_object, src/graph/graph_writer.c:33-58 */
    const McxCleaning *c = &h[i].cleaning;
    unsigned char flags[4] = {c->cleaned_tips, c->cleaned_unitigs, c->cleaned_kmers, c->is_graph_intersection};
    uint32_t len = (uint32_t)strlen(c->intersection_name);
    uint32_t thr_unitigs = c->cleaned_unitigs ? c->clean_unitigs_thresh : 0, thr_kmers = c->cleaned_kmers ? c->clean_kmers_thresh : 0;
    b += put(fh, flags, 4); b += put(fh, &thr_unitigs, 4); b += put(fh, &thr_kmers, 4);
    b += put(fh, &len, 4); b += put(fh, c->intersection_name, len);
  }
  b += put(fh, "CORTEX", 6);
  return b;
}
