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

void mcx_ginfo_init(McxGInfo *g)
{
  g->mean_read_length = 0; g->total_sequence = 0; g->seq_err = 0.01;
  g->sample_name = strdup("undefined");
}
void mcx_ginfo_free(McxGInfo *g) { free(g->sample_name); g->sample_name = NULL; }
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
    else {
      size_t n = strlen(dst->sample_name) + 1 + strlen(src->sample_name) + 1;
      char *s = malloc(n);
      strcpy(s, dst->sample_name); strcat(s, ","); strcat(s, src->sample_name);
      free(dst->sample_name); dst->sample_name = s;
    }
  }
  uint64_t total = dst->total_sequence + src->total_sequence;
  if(total > 0) {
    dst->seq_err = (dst->seq_err * dst->total_sequence + src->seq_err * src->total_sequence) / total;
    size_t src_contigs = 0;
    if(src->total_sequence && src->mean_read_length)
      src_contigs = ((double)src->total_sequence / src->mean_read_length) + 0.5;
    mcx_ginfo_update_contigs(dst, src->total_sequence, src_contigs);
  }
  dst->total_sequence = total;
}

static size_t put(FILE *fh, const void *p, size_t n)
{
  if(fwrite(p, 1, n, fh) != n) mcx_die("Cannot write to file");
  return n;
}

size_t mcx_write_ctx_header(FILE *fh, uint32_t kmer_size, uint32_t ncols, const McxGInfo *ginfo)
{
  size_t b = 0; uint32_t i;
  uint32_t version = 6, W = (kmer_size + 31) / 32;
  McxGInfo *h = calloc(ncols, sizeof(McxGInfo));
  for(i = 0; i < ncols; i++) { mcx_ginfo_init(&h[i]); mcx_ginfo_merge(&h[i], &ginfo[i]); }
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
    /* ErrorCleaning of a freshly built graph: nothing cleaned, no intersection (graph_info.c:4-10) */
    unsigned char flags[4] = {0, 0, 0, 0}; uint32_t zero = 0, len = 9;
    b += put(fh, flags, 4); b += put(fh, &zero, 4); b += put(fh, &zero, 4);
    b += put(fh, &len, 4); b += put(fh, "undefined", 9);
  }
  b += put(fh, "CORTEX", 6);
  for(i = 0; i < ncols; i++) mcx_ginfo_free(&h[i]);
  free(h);
  return b;
}
