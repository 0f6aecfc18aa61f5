/*
 * oracle/hts_stubs.c -- TEST INFRASTRUCTURE, not product code.
 *
 * The reference links htslib for SAM/BAM/CRAM input (libs/seq_file/seq_file.h:
 * 146-242,487-545,613-615).  That input format is out of scope for the build
 * hot path, and building htslib needs its own configure
 * step, so the oracle binary links these stand-ins instead: a read_t still gets
 * its (unused) bam1_t, and any attempt to open a .sam/.bam/.cram file fails
 * the way a missing file would.
 */
#include <stdlib.h>
#include <stdio.h>
#include "htslib/sam.h"

const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";

bam1_t *bam_init1(void) { return (bam1_t*)calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t *b) { if(b) { free(b->data); free(b); } }
void bam_hdr_destroy(bam_hdr_t *h) { (void)h; }
htsFile *hts_open(const char *fn, const char *mode) { (void)fn; (void)mode; return NULL; }
int hts_close(htsFile *fp) { (void)fp; return 0; }
const htsFormat *hts_get_format(htsFile *fp) { (void)fp; return NULL; }
bam_hdr_t *sam_hdr_read(samFile *fp) { (void)fp; return NULL; }
int sam_read1(samFile *fp, bam_hdr_t *h, bam1_t *b)
{
  (void)fp; (void)h; (void)b;
  fprintf(stderr, "oracle: SAM/BAM/CRAM input is not built into the oracle binary\n");
  abort();
}
