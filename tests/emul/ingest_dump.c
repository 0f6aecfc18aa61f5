/* tests/emul/ingest_dump.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Links the host driver's sequence ingest (mccortex_b200/host/seq_ingest.c) against stubs of the three
 * library calls it makes, so the CPU-only tests can see what `mccortex-b200 build --remove-pcr` would hand
 * to the GPU: the LINES bytes, the parallel quality bytes, one mate byte per read and the threshold.
 * tests/emul/emul_frontend --pcr then runs the device math on that dump.
 *
 * usage: ingest_dump <out-prefix> <mode se|pe|il> <fq_cutoff> <fq_offset> <hp> <matedir 0..3> <file1> [<file2>]
 *   writes <out-prefix>.lines .qual .mate and prints "fq_cutoff=<with offset> nreads=<n> nbatches=<n>"
 * mode `file`: the plain build ingest (mcx_load_seq_file: sequential reader, or the multi-threaded one of
 *   seq_ingest_par.c when the file is eligible); .lines / .qual are what mcx_graph_add_reads would receive, in call order
 */
#include "../../mccortex_b200/host/mcx_host.h"
#include <stdlib.h>
#include <string.h>

static FILE *f_lines, *f_qual, *f_mate;
static unsigned last_cut; static unsigned long long nreads_seen, nbatches;

int mcx_graph_add_reads_pcr(mcx_graph *g, const mcx_read_batch *b, const uint64_t *read_off, const uint8_t *mate, uint64_t nreads)
{
  (void)g;
  if(read_off[0] != 0 || read_off[nreads] != b->nbytes) { fprintf(stderr, "bad offsets\n"); exit(3); }
  for(uint64_t r = 0; r < nreads; r++)
    if(b->seq[read_off[r + 1] - 1] != '\n' || memchr(b->seq + read_off[r], '\n', read_off[r + 1] - 1 - read_off[r])) { fprintf(stderr, "offsets are not line starts\n"); exit(3); }
  fwrite(b->seq, 1, b->nbytes, f_lines);
  if(b->qual) fwrite(b->qual, 1, b->nbytes, f_qual);
  else for(uint64_t i = 0; i < b->nbytes; i++) fputc(0x7F, f_qual);
  fwrite(mate, 1, nreads, f_mate);
  if(b->fq_cutoff) last_cut = b->fq_cutoff;
  nreads_seen += nreads; nbatches++;
  return MCX_OK;
}
static int file_mode;
int mcx_graph_add_reads(mcx_graph *g, const mcx_read_batch *b)
{
  (void)g;
  if(!file_mode) { fprintf(stderr, "unexpected mcx_graph_add_reads\n"); exit(3); }
  if(b->layout != MCX_LAYOUT_LINES || b->mem != MCX_MEM_HOST || !b->nbytes || b->seq[b->nbytes - 1] != '\n') { fprintf(stderr, "bad batch\n"); exit(3); }
  fwrite(b->seq, 1, b->nbytes, f_lines);
  if(b->qual) fwrite(b->qual, 1, b->nbytes, f_qual);
  else for(uint64_t i = 0; i < b->nbytes; i++) fputc(0x7F, f_qual);
  for(uint64_t i = 0; i < b->nbytes; i++) nreads_seen += b->seq[i] == '\n';
  if(b->fq_cutoff) last_cut = b->fq_cutoff;
  nbatches++;
  return MCX_OK;
}
int mcx_graph_sync(mcx_graph *g, mcx_load_stats *st) { (void)g; memset(st, 0, sizeof(*st)); return MCX_OK; }

int main(int argc, char **argv)
{
  if(argc < 8 || (!strcmp(argv[2], "pe") && argc < 9)) { fprintf(stderr, "usage: %s <out-prefix> <se|pe|il> <fq_cutoff> <fq_offset> <hp> <matedir> <file1> [<file2>]\n", argv[0]); return 2; }
  char path[4096];
  snprintf(path, sizeof(path), "%s.lines", argv[1]); f_lines = fopen(path, "wb");
  snprintf(path, sizeof(path), "%s.qual", argv[1]); f_qual = fopen(path, "wb");
  snprintf(path, sizeof(path), "%s.mate", argv[1]); f_mate = fopen(path, "wb");
  McxLoadPrefs prefs; memset(&prefs, 0, sizeof(prefs));
  prefs.fq_cutoff = (uint8_t)atoi(argv[3]); prefs.fq_offset = (uint8_t)atoi(argv[4]); prefs.hp_cutoff = (uint8_t)atoi(argv[5]);
  prefs.matedir = (uint8_t)atoi(argv[6]); prefs.remove_pcr = true;
  mcx_msg_out = NULL;
  McxSeqFile *a = mcx_seq_open(argv[7]), *b = !strcmp(argv[2], "pe") ? mcx_seq_open(argv[8]) : NULL;
  if(!a || (!strcmp(argv[2], "pe") && !b)) { fprintf(stderr, "cannot open input\n"); return 2; }
  mcx_load_stats st; memset(&st, 0, sizeof(st));
  int r;
  if(!strcmp(argv[2], "file")) { file_mode = 1; prefs.remove_pcr = false; r = mcx_load_seq_file(NULL, a, &prefs, &st); }
  else r = mcx_load_seq_pcr(NULL, a, b, !strcmp(argv[2], "il"), &prefs, &st);
  fclose(f_lines); fclose(f_qual); fclose(f_mate);
  printf("fq_cutoff=%u nreads=%llu nbatches=%llu se=%llu pe=%llu\n", last_cut, nreads_seen, nbatches,
         (unsigned long long)st.num_se_reads, (unsigned long long)st.num_pe_reads);
  return r;
}
