/* ctx_join.c -- `mccortex-b200 join`: merge graph files into one, colours side by side or on top of each other.
 *
 * Command line and results of the reference's `join` (src/commands/ctx_join.c, graph_writer_merge_mkhdr /
 * graph_writer_stream_mkhdr in src/graph/graph_writer.c:398-647) without --intersect:
 *   - files are opened with the colour filter syntax [offset:]in.ctx[:cols]; without an offset a file's colours
 *     follow those of the files before it (graph_file_open2(..., into_offset = colours so far), ctx_join.c:131);
 *   - the output header is the merge of the input headers, colour by colour (graph_file_merge_header,
 *     graph_file_reader.c:61-75), written as it stands;
 *   - one input: the file is filtered as a stream, records stay in input order whether or not --sort was given
 *     (ctx_join.c:185-201, graph_writer_stream, graph_writer.c:398-453) -- host only;
 *   - else every file is merged into the device table (mcx_graph_load_records: coverage adds saturating, edges
 *     OR, a k-mer without coverage in the selected colours is skipped) and dumped, sorted with --sort.  The
 *     reference loads `--ncols` colours at a time and rewrites the file per group when memory is short
 *     (graph_writer.c:551-641); a B200 holds all colours at once, so --ncols is accepted and has no effect.
 */
#include "mcx_host.h"
#include <errno.h>
#include <fcntl.h>
#include <getopt.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define CMD "mccortex-b200"

static const char join_usage[] =
"usage: "CMD" join [options] in1.ctx [[offset:]in2.ctx[:1,2,4-5] ...]\n"
"\n"
"  Merge cortex graphs.\n"
"\n"
"  -h, --help              This help message\n"
"  -q, --quiet             Silence status output normally printed to STDERR\n"
"  -f, --force             Overwrite output files\n"
"  -o, --out <out.ctx>     Output file [required]\n"
"  -m, --memory <mem>      Memory to use\n"
"  -n, --nkmers <kmers>    Number of hash table entries (e.g. 1G ~ 1 billion)\n"
"  -N, --ncols <c>         Accepted for compatibility (all colours are loaded at once)\n"
"  -S, --sort              Output sorted graph file\n"
"  -D, --device <id>       CUDA device [default: 0]\n"
"\n"
"  Files can be specified with specific colours: samples.ctx:2,3\n"
"  Offset specifies where to load the first colour: 3:samples.ctx\n"
"  (-i, --intersect is not supported by this command yet)\n"
"\n";

static struct option longopts[] = {
  {"help", no_argument, NULL, 'h'},         {"out", required_argument, NULL, 'o'},
  {"force", no_argument, NULL, 'f'},        {"memory", required_argument, NULL, 'm'},
  {"nkmers", required_argument, NULL, 'n'}, {"ncols", required_argument, NULL, 'N'},
  {"intersect", required_argument, NULL, 'i'}, {"sort", no_argument, NULL, 'S'},
  {"device", required_argument, NULL, 'D'}, {NULL, 0, NULL, 0}};

static void die_lib(int r, const char *what)
{
  if(r == MCX_ERR_TABLE_FULL) mcx_die("Hash table is full"); /* src/graph/hash_table.c:119-123 */
  if(r == MCX_ERR_NO_DEVICE) mcx_die("No CUDA device: "CMD" has no CPU fallback");
  mcx_die("%s failed [%i]: %s", what, r, mcx_last_error());
}

static inline uint32_t safe_add_covg(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; return s > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)s; }

/* graph_writer_stream: one file through its colour filter, record order kept */
static uint64_t stream_filter(McxCtxFile *f, FILE *out, uint32_t out_ncols)
{
  const size_t W = MCX_CTX_W(f), in_rec = 8 * W + 5 * (size_t)f->num_of_cols, out_rec = 8 * W + 5 * (size_t)out_ncols;
  if(f->fh != stdin && fseek(f->fh, (long)f->hdr_size, SEEK_SET) != 0) mcx_die("fseek failed: %s", strerror(errno));
  const size_t chunk = 1u << 16;
  unsigned char *in = malloc(chunk * in_rec), *ob = malloc(chunk * out_rec);
  uint32_t *cv = malloc(4 * (size_t)out_ncols); unsigned char *ed = malloc(out_ncols);
  if(!in || !ob || !cv || !ed) mcx_die("Out of memory");
  uint64_t dumped = 0;
  for(;;) {
    size_t got = fread(in, 1, chunk * in_rec, f->fh);
    if(got == 0) break;
    if(got % in_rec != 0) mcx_die("Unexpected end of file: %s", f->path);
    size_t n = got / in_rec, w = 0;
    mcx_ctx_check_records(f, in, n);
    for(size_t r = 0; r < n; r++) {
      const unsigned char *rec = in + r * in_rec;
      memset(cv, 0, 4 * (size_t)out_ncols); memset(ed, 0, out_ncols);
      for(uint32_t i = 0; i < f->nfilter; i++) {
        uint32_t c; memcpy(&c, rec + 8 * W + 4 * (size_t)f->from_col[i], 4);
        cv[f->into_col[i]] = safe_add_covg(cv[f->into_col[i]], c);
        ed[f->into_col[i]] |= rec[8 * W + 4 * (size_t)f->num_of_cols + f->from_col[i]];
      }
      uint32_t keep = 0;
      for(uint32_t c = 0; c < out_ncols; c++) keep |= cv[c];
      if(!keep) continue;
      unsigned char *o = ob + w * out_rec;
      memcpy(o, rec, 8 * W); memcpy(o + 8 * W, cv, 4 * (size_t)out_ncols); memcpy(o + 8 * W + 4 * (size_t)out_ncols, ed, out_ncols);
      w++;
    }
    if(w && fwrite(ob, out_rec, w, out) != w) mcx_die("Cannot write to file");
    dumped += w;
  }
  free(in); free(ob); free(cv); free(ed);
  return dumped;
}

int mcx_cmd_join(int argc, char **argv)
{
  const char *out_path = NULL;
  bool force = false, mem_set = false, nkmers_set = false, sort_kmers = false;
  size_t mem_to_use = MCX_DEFAULT_MEM, num_kmers_arg = MCX_DEFAULT_NKMERS, use_ncols = 0;
  int device = 0, c;

  while((c = getopt_long_only(argc, argv, "hfo:m:n:N:i:SD:", longopts, NULL)) != -1) {
    switch(c) {
      case 0: break;
      case 'h': mcx_print_usage(join_usage, NULL); break;
      case 'o': if(out_path) mcx_print_usage(join_usage, "-o, --out <out.ctx> given more than once"); out_path = optarg; break;
      case 'f': if(force) mcx_print_usage(join_usage, "-f, --force given twice"); force = true; break;
      case 'm': if(mem_set) mcx_print_usage(join_usage, "-m, --memory <M> specifed more than once");
                if(!mcx_mem_to_integer(optarg, &mem_to_use)) mcx_print_usage(join_usage, "-m, --memory <M> requires a size e.g. 1GB: %s", optarg);
                mem_set = true; break;
      case 'n': if(nkmers_set) mcx_print_usage(join_usage, "-n, --nkmers <N> specifed more than once");
                if(!mcx_mem_to_integer(optarg, &num_kmers_arg)) mcx_print_usage(join_usage, "-n, --nkmers <M> requires a size e.g. 1G: %s", optarg);
                nkmers_set = true; break;
      case 'N': if(use_ncols) mcx_print_usage(join_usage, "-N, --ncols <c> given twice");
                use_ncols = (size_t)atol(optarg);
                if(use_ncols == 0) mcx_print_usage(join_usage, "-N, --ncols <c> must be > 0: %s", optarg);
                break;
      case 'i': mcx_die("--intersect is not supported by `"CMD" join` (use `"CMD" build --intersect`, or the reference's join)");
      case 'S': if(sort_kmers) mcx_print_usage(join_usage, "-S, --sort given twice"); sort_kmers = true; break;
      case 'D': device = atoi(optarg); break;
      default: mcx_die("`"CMD" join -h` for help. Bad option: %s", argv[optind - 1]);
    }
  }
  if(!out_path) mcx_print_usage(join_usage, "--out <out.ctx> required");
  if(optind >= argc) mcx_print_usage(join_usage, "Please specify at least one input graph file");

  const size_t nfiles = (size_t)(argc - optind);
  McxCtxFile **files = calloc(nfiles, sizeof(*files));
  mcx_status("Probing %zu graph files and 0 intersect files", nfiles);
  size_t ctx_max_cols = 0, i;
  uint64_t ctx_max_kmers = 0, ctx_sum_kmers = 0;
  for(i = 0; i < nfiles; i++) {
    files[i] = mcx_ctx_open(argv[optind + i], ctx_max_cols);
    if(files[0]->kmer_size != files[i]->kmer_size)
      mcx_print_usage(join_usage, "Kmer sizes don't match [%u vs %u]", files[0]->kmer_size, files[i]->kmer_size);
    if(files[i]->into_ncols > ctx_max_cols) ctx_max_cols = files[i]->into_ncols;
    uint64_t nk = files[i]->num_of_kmers < 0 ? 0 : (uint64_t)files[i]->num_of_kmers;
    if(nk > ctx_max_kmers) ctx_max_kmers = nk;
    ctx_sum_kmers += nk;
  }
  const uint32_t kmer_size = files[0]->kmer_size;
  if(kmer_size > 63 || !(kmer_size & 1u)) mcx_die("Unsupported kmer size (%u)", kmer_size);
  const bool to_stdout = strcmp(out_path, "-") == 0;
  if(use_ncols) {
    if(use_ncols < ctx_max_cols && to_stdout) mcx_die("I need %zu colours if outputting to STDOUT (--ncols)", ctx_max_cols);
    if(use_ncols > ctx_max_cols) mcx_warn("I only need %zu colour%s ('--ncols %zu' ignored)", ctx_max_cols, ctx_max_cols == 1 ? "" : "s", use_ncols);
  }

  /* futil_create_output: refuses to overwrite without -f */
  FILE *out = stdout;
  if(!to_stdout) {
    int mode = O_CREAT | O_EXCL | O_WRONLY | O_TRUNC;
    if(force) mode &= ~O_EXCL;
    int fd = open(out_path, mode, 0666);
    if(fd < 0) {
      if(errno == EEXIST) mcx_die("File already exists: %s", out_path);
      mcx_die("Cannot write to file: %s [%s]", out_path, strerror(errno));
    }
    out = fdopen(fd, "w");
  }
  setvbuf(out, NULL, _IOFBF, 4u << 20);
  mcx_status("Output %zu cols; from %zu files; intersecting 0 graphs; ", ctx_max_cols, nfiles);

  /* graph_file_merge_header over the files, in order */
  McxGInfo *ginfo = calloc(ctx_max_cols, sizeof(McxGInfo));
  for(i = 0; i < ctx_max_cols; i++) mcx_ginfo_init(&ginfo[i]);

  uint64_t nrec = 0;
  if(nfiles == 1) { /* ctx_join.c:185-201: one file is always streamed, --sort or not */
    McxCtxFile *f = files[0];
    for(uint32_t j = 0; j < f->nfilter; j++) mcx_ginfo_merge(&ginfo[f->into_col[j]], &f->ginfo[f->from_col[j]]);
    mcx_status("Filtering %s to %s with stream filter", f->path, to_stdout ? "STDOUT" : out_path);
    mcx_write_ctx_header_as_is(out, kmer_size, (uint32_t)ctx_max_cols, ginfo, files[nfiles - 1]->num_of_bitfields);
    nrec = stream_filter(f, out, (uint32_t)ctx_max_cols);
  } else {
    /* ctx_join.c:212-226: the reference sizes its table for ONE colour in memory (then takes as many as fit) */
    const size_t W = (kmer_size + 31) / 32;
    size_t bits_per_kmer = 64 * W + 40 + (sort_kmers ? 64 : 0), graph_mem = 0;
    size_t kmers_in_hash = mcx_get_kmers_in_hash(mem_to_use, mem_set, num_kmers_arg, nkmers_set, bits_per_kmer,
                                                 (int64_t)ctx_max_kmers, (int64_t)ctx_sum_kmers, true, &graph_mem);
    if(graph_mem > mem_to_use) { char m[64]; mcx_bytes_to_str(graph_mem, m); mcx_die("Need to set higher memory limit [ at least -m %s ]", m); }
    if(mcx_device_count() == 0) mcx_die("No CUDA device: "CMD" has no CPU fallback");
    mcx_graph *g = NULL;
    int r = mcx_graph_create(kmer_size, (uint32_t)ctx_max_cols, kmers_in_hash, device, 0, &g);
    if(r) die_lib(r, "mcx_graph_create");
    mcx_status("Loading and saving %zu colours at once", ctx_max_cols);
    for(i = 0; i < nfiles; i++) {
      r = mcx_ctx_load(g, files[i], ginfo, ctx_max_cols, 0, NULL, NULL, NULL);
      if(r) die_lib(r, "loading graph file");
    }
    mcx_load_stats st;
    r = mcx_graph_sync(g, &st);
    if(r) die_lib(r, "loading graph file");
    mcx_write_ctx_header_as_is(out, kmer_size, (uint32_t)ctx_max_cols, ginfo, files[nfiles - 1]->num_of_bitfields);
    uint32_t rec_bytes = 0;
    r = mcx_graph_export_begin(g, sort_kmers ? 1 : 0, &nrec, &rec_bytes);
    if(r) die_lib(r, "mcx_graph_export_begin");
    size_t chunk_recs = (64u << 20) / rec_bytes;
    char *buf = malloc(chunk_recs * rec_bytes);
    if(!buf) mcx_die("Out of memory");
    for(uint64_t at = 0; at < nrec; at += chunk_recs) {
      uint64_t n = nrec - at < chunk_recs ? nrec - at : chunk_recs;
      r = mcx_graph_export_read(g, at, n, buf);
      if(r) die_lib(r, "mcx_graph_export_read");
      if(fwrite(buf, rec_bytes, n, out) != n) mcx_die("Cannot write to file");
    }
    free(buf);
    mcx_graph_export_end(g);
    mcx_graph_destroy(g);
  }
  if(out != stdout) fclose(out); else fflush(out);
  { char a[64]; mcx_ulong_to_str(nrec, a);
    mcx_status("[graphwriter] Dumped %s kmers in %zu colour%s into: %s (format version: 6)", a, ctx_max_cols,
               ctx_max_cols == 1 ? "" : "s", to_stdout ? "STDOUT" : out_path); }
  for(i = 0; i < nfiles; i++) mcx_ctx_close(files[i]);
  for(i = 0; i < ctx_max_cols; i++) mcx_ginfo_free(&ginfo[i]);
  free(ginfo); free(files);
  return EXIT_SUCCESS;
}
