/* ctx_sort.c -- `mccortex-b200 sort [options] <in.ctx>`
 *
 * Drop-in for the reference's `mccortexNN sort` (src/commands/ctx_sort.c): same options, same result
 * (the records of the graph file in ascending key order, header untouched; in place unless -o).
 * The reference reads the whole file, qsorts an array of pointers and writes the records back; here
 * the records are sorted on the GPU (mcx_sort_records: radix sort of the keys + gather).
 */
#include "mcx_host.h"
#include <errno.h>
#include <fcntl.h>
#include <getopt.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define CMD "mccortex-b200"

static const char sort_usage[] =
"usage: "CMD" sort [options] <in.ctx>\n"
"\n"
"  Sort a cortex graph file. Loads entire graph into memory then sorts (on the GPU).\n"
"\n"
"  -h, --help              This help message\n"
"  -q, --quiet             Silence status output normally printed to STDERR\n"
"  -f, --force             Overwrite output files\n"
"  -m, --memory <mem>      Memory to use\n"
"  -n, --nkmers <kmers>    Number of hash table entries (e.g. 1G ~ 1 billion)\n"
"  -o, --out <out.ctx>     Output file [default: overwrite input]\n"
"  -D, --device <id>       CUDA device [default: 0]\n"
"\n";

static struct option longopts[] = {
  {"help", no_argument, NULL, 'h'},   {"force", no_argument, NULL, 'f'},
  {"memory", required_argument, NULL, 'm'}, {"nkmers", required_argument, NULL, 'n'},
  {"out", required_argument, NULL, 'o'},    {"device", required_argument, NULL, 'D'},
  {NULL, 0, NULL, 0}};

int mcx_cmd_sort(int argc, char **argv)
{
  const char *out_path = NULL;
  bool force = false, mem_set = false, nkmers_set = false;
  size_t mem_to_use = MCX_DEFAULT_MEM, num_kmers_arg = 0;
  int device = 0, c;

  while((c = getopt_long_only(argc, argv, "hfm:n:o:D:", longopts, NULL)) != -1) {
    switch(c) {
      case 0: break;
      case 'h': mcx_print_usage(sort_usage, NULL); break;
      case 'f': if(force) mcx_print_usage(sort_usage, "-f, --force given twice"); force = true; break;
      case 'm': if(mem_set) mcx_print_usage(sort_usage, "-m, --memory <M> specifed more than once");
                if(!mcx_mem_to_integer(optarg, &mem_to_use)) mcx_print_usage(sort_usage, "-m, --memory <M> requires a size e.g. 1GB: %s", optarg);
                mem_set = true; break;
      case 'n': if(nkmers_set) mcx_print_usage(sort_usage, "-n, --nkmers <N> specifed more than once");
                if(!mcx_mem_to_integer(optarg, &num_kmers_arg)) mcx_print_usage(sort_usage, "-n, --nkmers <M> requires a size e.g. 1G: %s", optarg);
                nkmers_set = true; break;
      case 'o': if(out_path) mcx_print_usage(sort_usage, "-o, --out given twice"); out_path = optarg; break;
      case 'D': device = atoi(optarg); break;
      default: mcx_die("`"CMD" sort -h` for help. Bad option: %s", argv[optind - 1]);
    }
  }
  if(optind + 1 != argc) mcx_print_usage(sort_usage, "Require exactly one input graph file (.ctx)");
  const char *ctx_path = argv[optind];

  McxCtxFile *f = mcx_ctx_open(ctx_path, 0);
  if(!mcx_ctx_filter_is_direct(f)) mcx_die("Cannot open graph file with a filter ('in.ctx:blah' syntax)");
  if(f->kmer_size > 63) mcx_die("Please recompile with correct kmer size (%u)", f->kmer_size);

  size_t num_kmers;
  if(f->num_of_kmers < 0) {
    if(!nkmers_set) mcx_die("If reading from a stream, must give -n <num_kmers>");
    num_kmers = num_kmers_arg;
  } else num_kmers = (size_t)f->num_of_kmers;

  /* futil_fopen_create: refuses to overwrite without -f */
  FILE *fout = NULL;
  if(out_path) {
    int mode = O_CREAT | O_EXCL | O_WRONLY | O_TRUNC;
    if(force) mode &= ~O_EXCL;
    int fd = strcmp(out_path, "-") == 0 ? -1 : open(out_path, mode, 0666);
    if(strcmp(out_path, "-") != 0 && fd < 0) {
      if(errno == EEXIST) mcx_die("File already exists: %s", out_path);
      mcx_die("Cannot write to file: %s [%s]", out_path, strerror(errno));
    }
    fout = fd < 0 ? stdout : fdopen(fd, "w");
  }

  const size_t ncols = f->num_of_cols, kmer_mem = 8u * MCX_CTX_W(f) + 5u * ncols;
  const size_t memory = (sizeof(char *) + kmer_mem) * num_kmers;
  char mem_str[64]; mcx_bytes_to_str(memory, mem_str);
  if(memory > mem_to_use) mcx_die("Require at least %s memory", mem_str);
  mcx_status("[memory] Total: %s", mem_str);

  unsigned char *mem = malloc(kmer_mem * num_kmers + 1);
  if(!mem) mcx_die("Out of memory");
  if(f->fh != stdin && fseek(f->fh, (long)f->hdr_size, SEEK_SET) != 0) mcx_die("fseek failed");
  size_t nkread = fread(mem, 1, num_kmers * kmer_mem, f->fh);
  if(nkread != num_kmers * kmer_mem) mcx_die("Could only read %zu bytes [<%zu]", nkread, num_kmers * kmer_mem);
  char tmpc;
  if(fread(&tmpc, 1, 1, f->fh) != 0) mcx_die("More kmers in file than believed (kmers: %zu ncols: %zu).", num_kmers, ncols);
  mcx_status("Read %zu kmers with %zu colour%s", num_kmers, ncols, ncols == 1 ? "" : "s");

  if(mcx_device_count() == 0) mcx_die("No CUDA device: "CMD" has no CPU fallback");
  int r = mcx_sort_records(device, f->kmer_size, (uint32_t)ncols, mem, num_kmers, mem);
  if(r) mcx_die("mcx_sort_records failed [%i]: %s", r, mcx_last_error());

  if(out_path) mcx_ctx_write_header_raw(fout, f);
  else {
    /* in place: reopen for writing at the first record (the reference opens the file "r+") */
    fout = fopen(f->path, "r+");
    if(!fout) mcx_die("Cannot open file: %s [%s]", f->path, strerror(errno));
    if(fseek(fout, (long)f->hdr_size, SEEK_SET) != 0) mcx_die("fseek failed");
  }
  if(fwrite(mem, kmer_mem, num_kmers, fout) != num_kmers) mcx_die("Cannot write to file");
  if(fout != stdout) fclose(fout); else fflush(fout);
  mcx_ctx_close(f);
  free(mem);
  return EXIT_SUCCESS;
}
