/* ctx_build.c -- `mccortex-b200 build [options] <out.ctx>`
 *
 * Drop-in for the reference's `mccortexNN build` (src/commands/ctx_build.c, dispatched from
 * src/main/mccortex.c:279-332): same options, same ordering rules, same exit status and the
 * same .ctx v6 bytes (with -S), but the graph lives on a B200 behind include/mcx_gpu.h.
 * One binary serves every odd k in 3..63 and writes W = ceil(k/32) in the header, which is
 * what mccortex31 / mccortex63 write for their k ranges.
 *
 * Not (yet) supported, and rejected with an error rather than silently ignored:
 *   SAM/BAM/CRAM input.
 */
#include "mcx_host.h"
#include <ctype.h>
#include <errno.h>
#include <fcntl.h>
#include <getopt.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <time.h>
#include <unistd.h>
#include <zlib.h>

#define CMD "mccortex-b200"
#define MAX_IO_THREADS 10 /* src/global/global.h:41: tasks per build_graph() call */
#define MIN_KMER 3
#define MAX_KMER 63

static const char build_usage[] =
"usage: "CMD" build [options] <out.ctx>\n"
"\n"
"  Build a cortex graph on a B200 GPU.  \n"
"\n"
"  -h, --help               This help message\n"
"  -q, --quiet              Silence status output normally printed to STDERR\n"
"  -f, --force              Overwrite output files\n"
"  -m, --memory <mem>       Memory to use\n"
"  -n, --nkmers <kmers>     Number of hash table entries (e.g. 1G ~ 1 billion)\n"
"  -t, --threads <T>        Accepted for compatibility (the GPU does the work)\n"
"  -k, --kmer <kmer>        Kmer size must be odd (63 >= k >= 3)\n"
"  -s, --sample <name>      Sample name (required before any seq args)\n"
"  -1, --seq <in.fa>        Load sequence data\n"
"  -2, --seq2 <in1:in2>     Load paired end sequence data\n"
"  -i, --seqi <in.fq>       Load paired end sequence from a single file\n"
"  -Q, --fq-cutoff <Q>      Filter quality scores [default: 0 (off)]\n"
"  -O, --fq-offset <N>      FASTQ ASCII offset    [default: 0 (auto-detect)]\n"
"  -H, --cut-hp <bp>        Breaks reads at homopolymers >= <bp> [default: off]\n"
"  -p, --remove-pcr         Remove (or keep) PCR duplicate reads\n"
"  -P, --keep-pcr           Don't do PCR duplicate removal [default]\n"
"  -M, --matepair <orient>  Mate pair orientation: FF,FR,RF,RR [default: FR]\n"
"  -g, --graph <in.ctx>     Load samples from a graph file (.ctx)\n"
"  -I, --intersect <i.ctx>  Only load kmers that appear in i.ctx. Multiple -I will merge\n"
"  -S, --sort               Output a graph file ordered by kmer\n"
"  -D, --device <id[,id..]> CUDA device [default: 0]; several: one replica of the graph per device, every\n"
"                           batch of reads goes to one of them, the replicas are merged before the dump\n"
"      --shard              with several devices: ONE graph, hash-partitioned over the devices (each holds 1/N of\n"
"                           the k-mers; k-mers travel to their owner over NVLink); sequence inputs only\n"
"\n"
"  Note: Argument must come before input file\n"
"  --sample <name> is required before sequence input can be loaded.\n"
"  Consecutive sequence options are loaded into the same colour.\n"
"\n";

static struct option longopts[] = {
  {"help", no_argument, NULL, 'h'},       {"memory", required_argument, NULL, 'm'},
  {"nkmers", required_argument, NULL, 'n'}, {"threads", required_argument, NULL, 't'},
  {"force", no_argument, NULL, 'f'},      {"kmer", required_argument, NULL, 'k'},
  {"sample", required_argument, NULL, 's'}, {"sort", no_argument, NULL, 'S'},
  {"seq", required_argument, NULL, '1'},  {"seq2", required_argument, NULL, '2'},
  {"seqi", required_argument, NULL, 'i'}, {"matepair", required_argument, NULL, 'M'},
  {"fq-cutoff", required_argument, NULL, 'Q'}, {"fq-offset", required_argument, NULL, 'O'},
  {"cut-hp", required_argument, NULL, 'H'}, {"remove-pcr", no_argument, NULL, 'p'},
  {"keep-pcr", no_argument, NULL, 'P'},   {"graph", required_argument, NULL, 'g'},
  {"intersect", required_argument, NULL, 'I'}, {"device", required_argument, NULL, 'D'},
  {"shard", no_argument, NULL, 1000},
  {NULL, 0, NULL, 0}};

typedef struct {
  McxSeqFile *file, *file2;   /* file2: the second file of a --seq2 pair kept together (--remove-pcr only) */
  bool interleaved;           /* --seqi */
  McxLoadPrefs prefs;
  mcx_load_stats stats;
} BuildTask;

static BuildTask *tasks = NULL; static size_t ntasks = 0, tasks_cap = 0;
static char **sample_names = NULL; static size_t *sample_cols = NULL; static size_t nsamples = 0;
static McxCtxFile **gfiles = NULL; static size_t ngfiles = 0;
static McxCtxFile **ifiles = NULL; static size_t nifiles = 0;   /* --intersect graphs */
static size_t nthreads = 0, kmer_size = 0, output_colours = 0;
static bool mem_set = false, nkmers_set = false, force = false, sort_kmers = false;
static size_t mem_to_use = MCX_DEFAULT_MEM, num_kmers = MCX_DEFAULT_NKMERS;
static int device = 0;
#define MAX_DEVICES 16
static int devices[MAX_DEVICES] = {0}, ndevices = 1;
static bool shard_mode = false;    /* --shard: one graph partitioned over the devices (mcx_shardset_*) instead of replicas */
static char *out_path = NULL;

#define usage_err(...) mcx_print_usage(build_usage, __VA_ARGS__)

static void opt_name(int c, char *out, size_t n)
{
  const struct option *o;
  snprintf(out, n, "-%c, --Unknown", c);
  for(o = longopts; o->name; o++) if(o->val == c) { snprintf(out, n, "-%c, --%s", c, o->name); return; }
}

static size_t parse_size(const char *cmd, const char *arg, bool nonzero)
{
  char *end; unsigned long v;
  if(*arg < '0' || *arg > '9') usage_err("%s requires an int x >= 0: %s", cmd, arg);
  v = strtoul(arg, &end, 10);
  if(*end) usage_err("%s requires an int x >= 0: %s", cmd, arg);
  if(nonzero && v == 0) usage_err("%s <N> must be > 0: %s", cmd, arg);
  return v;
}
static uint8_t parse_uint8(const char *cmd, const char *arg)
{
  size_t v = parse_size(cmd, arg, false);
  if(v > 255) usage_err("%s requires an int 0 <= x < 256: %s", cmd, arg);
  return (uint8_t)v;
}

/* src/commands/ctx_build.c:119-131 */
static void check_sample_name(const char *s)
{
  const char *p;
  if(strlen(s) < 1) mcx_die("Sample name is too short: '%s'", s);
  if(!strcmp(s, "undefined")) mcx_die("Bad sample name: '%s'", s);
  if(!strcmp(s, "noname")) mcx_die("Bad sample name: '%s'", s);
  if(s[0] == '.') mcx_die("Sample name should start with a dot: '%s'", s);
  for(p = s; *p; p++) {
    if(isspace((unsigned char)*p)) mcx_die("Sample name should not contain whitespace: '%s'", s);
    if(!isgraph((unsigned char)*p)) mcx_die("Bad character in sample name: '%s'", s);
  }
}

static bool has_ext(const char *p, const char *ext)
{
  size_t n = strlen(p), m = strlen(ext);
  return n >= m && strcasecmp(p + n - m, ext) == 0;
}

static McxSeqFile *open_seq(const char *path, char opt)
{
  if(has_ext(path, ".sam") || has_ext(path, ".bam") || has_ext(path, ".cram"))
    mcx_die("SAM/BAM/CRAM input is not supported by "CMD": %s", path);
  McxSeqFile *sf = mcx_seq_open(path);
  if(!sf) mcx_die("Cannot open -%c file: %s", opt, path);
  return sf;
}

static void push_task(const char *path, char opt, const McxLoadPrefs *prefs)
{
  McxSeqFile *sf = open_seq(path, opt);
  if(ntasks == tasks_cap) { tasks_cap = tasks_cap ? tasks_cap * 2 : 16; tasks = realloc(tasks, tasks_cap * sizeof(*tasks)); }
  memset(&tasks[ntasks], 0, sizeof(tasks[ntasks]));
  tasks[ntasks].file = sf; tasks[ntasks].prefs = *prefs;
  ntasks++;
}

/* src/commands/ctx_build.c:99-117 (add_task) + src/basic/async_read_io.c:27-80: without
 * --remove-pcr a --seq2 pair is two independent single-end tasks */
static void add_seq_arg(char opt, char *arg, const McxLoadPrefs *prefs)
{
  if(prefs->fq_offset >= 128) mcx_die("fq-offset too big: %i", (int)prefs->fq_offset);
  if(prefs->fq_offset + prefs->fq_cutoff >= 128) mcx_die("fq-cutoff too big: %i", prefs->fq_offset + prefs->fq_cutoff);
  if(opt == '2') {
    char *sep = strchr(arg, ':');
    if(!sep) sep = strchr(arg, ',');
    if(!sep || strchr(sep + 1, *sep)) mcx_die("Expected -%c <in1>:<in2>", opt);
    *sep = '\0';
    push_task(arg, opt, prefs);
    if(prefs->remove_pcr) tasks[ntasks - 1].file2 = open_seq(sep + 1, opt); /* submit paired end reads together */
    else push_task(sep + 1, opt, prefs);
  } else {
    push_task(arg, opt, prefs);
    tasks[ntasks - 1].interleaved = (opt == 'i');
  }
}

static void parse_args(int argc, char **argv)
{
  McxLoadPrefs prefs; memset(&prefs, 0, sizeof(prefs));
  prefs.matedir = 1; /* SEQ_LOADING_PREFS_INIT: READPAIR_FR */
  int intocolour = -1, c;
  bool sample_named = false, pref_unused = false, kmer_given = false, threads_given = false;
  char cmd[100];
  const char shortopts[] = "+hm:n:t:fk:s:S1:2:i:M:Q:O:H:pPg:I:D:";

  while((c = getopt_long_only(argc, argv, shortopts + 1, longopts, NULL)) != -1) {
    opt_name(c, cmd, sizeof(cmd));
    switch(c) {
      case 0: break;
      case 'h': usage_err(NULL); break;
      case 't': if(threads_given) usage_err("%s given twice", cmd); threads_given = true;
                nthreads = parse_size(cmd, optarg, true); break;
      case 'm': if(mem_set) usage_err("-m, --memory <M> specifed more than once");
                if(!mcx_mem_to_integer(optarg, &mem_to_use)) usage_err("-m, --memory <M> requires a size e.g. 1GB: %s", optarg);
                if(mem_to_use == 0) usage_err("--memory <M> cannot be zero");
                mem_set = true; break;
      case 'n': if(nkmers_set) usage_err("-n, --nkmers <N> specifed more than once");
                if(!mcx_mem_to_integer(optarg, &num_kmers)) usage_err("-n, --nkmers <M> requires a size e.g. 1G: %s", optarg);
                if(num_kmers == 0) usage_err("--nkmer <N> cannot be zero");
                nkmers_set = true; break;
      case 'f': if(force) usage_err("%s given twice", cmd); force = true; break;
      case 'k': if(kmer_given) usage_err("%s given twice", cmd); kmer_given = true;
                kmer_size = parse_size(cmd, optarg, true);
                if(kmer_size < MIN_KMER || kmer_size > MAX_KMER) mcx_die("Please recompile with correct kmer size (%zu)", kmer_size);
                if(!(kmer_size & 1)) mcx_die("Invalid kmer-size (%zu): requires odd number %i <= k <= %i", kmer_size, MIN_KMER, MAX_KMER);
                break;
      case 's':
        intocolour++;
        check_sample_name(optarg);
        sample_names = realloc(sample_names, (nsamples + 1) * sizeof(char *));
        sample_cols = realloc(sample_cols, (nsamples + 1) * sizeof(size_t));
        sample_cols[nsamples] = (size_t)intocolour;
        sample_names[nsamples++] = optarg;
        sample_named = true;
        break;
      case 'S': if(sort_kmers) usage_err("%s given twice", cmd); sort_kmers = true; break;
      case '1': case '2': case 'i':
        pref_unused = false;
        if(!sample_named) usage_err("Please give sample name first [-s,--sample <name>]");
        prefs.colour = (uint32_t)intocolour;
        add_seq_arg((char)c, optarg, &prefs);
        break;
      case 'M':
        if(strcmp(optarg, "FF") && strcmp(optarg, "FR") && strcmp(optarg, "RF") && strcmp(optarg, "RR"))
          mcx_die("-M,--matepair <orient> must be one of: FF,FR,RF,RR");
        prefs.matedir = (uint8_t)((optarg[0] == 'R' ? 2 : 0) | (optarg[1] == 'R' ? 1 : 0)); /* cortex_types.h:17-25 */
        pref_unused = true; break; /* only matters with --remove-pcr */
      case 'O': prefs.fq_offset = parse_uint8(cmd, optarg); pref_unused = true; break;
      case 'Q': prefs.fq_cutoff = parse_uint8(cmd, optarg); pref_unused = true; break;
      case 'H': prefs.hp_cutoff = parse_uint8(cmd, optarg); pref_unused = true; break;
      case 'p': prefs.remove_pcr = true; pref_unused = true; break;
      case 'P': prefs.remove_pcr = false; pref_unused = true; break;
      case 'g': { /* src/commands/ctx_build.c:189-196 */
        if(intocolour == -1) intocolour = 0;
        McxCtxFile *gf = mcx_ctx_open(optarg, (size_t)intocolour);
        if((int)gf->into_ncols - 1 > intocolour) intocolour = (int)gf->into_ncols - 1;
        gfiles = realloc(gfiles, (ngfiles + 1) * sizeof(*gfiles));
        gfiles[ngfiles++] = gf;
        sample_named = false;
        break;
      }
      case 'I': { /* src/commands/ctx_build.c:197-204 */
        McxCtxFile *gf = mcx_ctx_open(optarg, 0);
        if(gf->into_ncols > 1) mcx_warn("Flattening intersection graph into colour 0: %s", optarg);
        mcx_ctx_flatten(gf, 0);
        ifiles = realloc(ifiles, (nifiles + 1) * sizeof(*ifiles));
        ifiles[nifiles++] = gf;
        break;
      }
      case 'D': {
        /* -D 2 or -D 0,1,2,3 */
        char *list = strdup(optarg), *tok, *save = NULL;
        ndevices = 0;
        for(tok = strtok_r(list, ",", &save); tok; tok = strtok_r(NULL, ",", &save)) {
          if(ndevices == MAX_DEVICES) mcx_die("%s: at most %d devices", cmd, MAX_DEVICES);
          devices[ndevices++] = (int)parse_size(cmd, tok, false);
        }
        if(ndevices == 0) mcx_die("%s <id[,id..]> requires an argument", cmd);
        device = devices[0];
        free(list);
        break;
      }
      case 1000: shard_mode = true; break;
      case ':': case '?':
        mcx_die("`"CMD" build -h` for help. Bad option: %s", argv[optind - 1]);
      default: mcx_die("Bad option: %s", cmd);
    }
  }
  if(!nthreads) nthreads = 2;

  if(optind + 1 > argc) usage_err("Expected exactly one graph file");
  else if(optind + 1 < argc) usage_err("Expected only one graph file. What is this: '%s'", argv[optind]);
  out_path = argv[optind];
  mcx_status("Saving graph to: %s", strcmp(out_path, "-") ? out_path : "STDOUT");

  if(nsamples == 0) usage_err("No inputs given");
  if(pref_unused) usage_err("Arguments not given BEFORE sequence file");
  if(!kmer_size) mcx_die("kmer size not set with -k <K>");
  for(size_t i = 0; i < ngfiles; i++)
    if(gfiles[i]->kmer_size != kmer_size)
      usage_err("Input graph kmer_size doesn't match [%u vs %zu]: %s", gfiles[i]->kmer_size, kmer_size, gfiles[i]->input);
  output_colours = (size_t)(intocolour + (sample_named ? 1 : 0));
}

/* src/basic/file_util.c:164-174 */
static void create_output(const char *path)
{
  if(strcmp(path, "-") == 0) return;
  int mode = O_CREAT | O_EXCL | O_WRONLY | O_APPEND;
  if(force) mode &= ~O_EXCL;
  int fd = open(path, mode, 0666);
  if(fd < 0) {
    if(errno == EEXIST) mcx_die("File already exists: %s", path);
    else mcx_die("Cannot write to file: %s [%s]", path, strerror(errno));
  }
  close(fd);
}

static void die_mcx(int r, const char *what)
{
  if(r == MCX_ERR_TABLE_FULL) mcx_die("Hash table is full"); /* src/graph/hash_table.c:119-123 */
  if(r == MCX_ERR_NOMEM) mcx_die("Out of memory on the GPU (%s)", mcx_last_error());
  if(r == MCX_ERR_NO_DEVICE) mcx_die("No CUDA device: "CMD" has no CPU fallback");
  if(r == MCX_ERR_UNSUPPORTED) mcx_die("%s: not supported (%s)", what, mcx_last_error());
  mcx_die("%s failed [%i]: %s", what, r, mcx_last_error());
}

/* src/tools/build_graph.c:352-386 */
static void print_task_stats(const BuildTask *t)
{
  char a[64], b[64], c[64];
  mcx_status("[task] input: %s colour: %u", mcx_seq_path(t->file), t->prefs.colour);
  mcx_ulong_to_str(t->stats.num_se_reads, a); mcx_ulong_to_str(t->stats.num_pe_reads, b);
  mcx_status("  SE reads: %s  PE reads: %s", a, b);
  if(t->stats.num_good_reads != UINT64_MAX) {
    mcx_ulong_to_str(t->stats.num_good_reads, a); mcx_ulong_to_str(t->stats.num_bad_reads, b);
    mcx_status("  good reads: %s  bad reads: %s", a, b);
  }
  mcx_ulong_to_str(t->stats.num_dup_se_reads, a); mcx_ulong_to_str(t->stats.num_dup_pe_pairs, b);
  mcx_status("  dup SE reads: %s  dup PE pairs: %s", a, b);
  mcx_ulong_to_str(t->stats.total_bases_read, a); mcx_ulong_to_str(t->stats.total_bases_loaded, b);
  mcx_status("  bases read: %s  bases loaded: %s", a, b);
  mcx_ulong_to_str(t->stats.contigs_parsed, a); mcx_ulong_to_str(t->stats.num_kmers_loaded, b);
  mcx_ulong_to_str(t->stats.num_kmers_novel, c);
  mcx_status("  num contigs: %s  num kmers: %s novel kmers: %s", a, b, c);
}

/* one of several files of one colour that are read at the same time */
typedef struct { pthread_t thread; BuildTask *task; mcx_graph *g; int rc; } FileLoad;
static void *file_load_main(void *arg)
{
  FileLoad *fl = arg;
  fl->rc = mcx_load_seq_file(fl->g, fl->task->file, &fl->task->prefs, &fl->task->stats);
  return NULL;
}

/* the device graph, created on its own thread (see ctx_build) */
static struct {
  pthread_t thread; bool joined;
  uint32_t k, ncols, flags; uint64_t capacity; int device; bool host_batches;
  mcx_graph *g; int rc; const char *what;
  mcx_graph *gs[MAX_DEVICES]; /* gs[0] == g; more with -D a,b,..: replicas on the other devices */
  mcx_shardset *ss;           /* --shard: the partitioned graph (then g is only a non-NULL token for the ingest) */
  unsigned next;              /* round robin over the replicas (under mcx_ingest.lock when files load concurrently) */
  int done; /* set (release) by the thread when rc / g are final; polled (acquire) by graph_ready */
} ginit;

static void *graph_init_main(void *arg)
{
  (void)arg;
  if(mcx_device_count() == 0) { ginit.rc = MCX_ERR_NO_DEVICE; ginit.what = "device"; }
  else if(shard_mode) {
    ginit.rc = mcx_shardset_create(ginit.k, ginit.ncols, ginit.capacity, devices, (uint32_t)ndevices, &ginit.ss);
    ginit.what = "mcx_shardset_create";
    ginit.g = (mcx_graph *)ginit.ss;
  }
  else {
    for(int d = 0; d < ndevices && !ginit.rc; d++) {
      ginit.rc = mcx_graph_create(ginit.k, ginit.ncols, ginit.capacity, devices[d], ginit.flags, &ginit.gs[d]);
      ginit.what = "mcx_graph_create";
      if(!ginit.rc && ginit.host_batches) { ginit.rc = mcx_graph_prepare_host(ginit.gs[d]); ginit.what = "mcx_graph_prepare_host"; }
    }
    ginit.g = ginit.gs[0];
  }
  __atomic_store_n(&ginit.done, 1, __ATOMIC_RELEASE);
  return NULL;
}
/* -D a,b,..: which replica takes the next batch; joining all of them */
static mcx_graph *route_replica(mcx_graph *g) { (void)g; return ginit.gs[ginit.next++ % (unsigned)ndevices]; }
static int sync_replicas(mcx_graph *g, mcx_load_stats *st)
{
  (void)g;
  memset(st, 0, sizeof(*st));
  int rc = MCX_OK;
  for(int d = 0; d < ndevices; d++) {
    mcx_load_stats s;
    int r = mcx_graph_sync(ginit.gs[d], &s);
    if(r && !rc) rc = r;
    mcx_add_load_stats(st, &s);
  }
  return rc;
}
/* --shard: the ingest's batches go to the shard set */
static int submit_shard(mcx_graph *g, const mcx_read_batch *b) { (void)g; return mcx_shardset_add_reads(ginit.ss, b); }
static int sync_shards(mcx_graph *g, mcx_load_stats *st) { (void)g; return mcx_shardset_sync(ginit.ss, st); }
/* fold replica d into replica 0: its records (coverage adds saturating, edges OR -- mcx_graph_load_records) */
static void merge_replica(int d, uint32_t ncols)
{
  uint64_t nrec = 0; uint32_t rec_bytes = 0;
  int r = mcx_graph_export_begin(ginit.gs[d], 0, &nrec, &rec_bytes);
  if(r) die_mcx(r, "mcx_graph_export_begin");
  uint32_t *cols = malloc(4 * (size_t)ncols);
  for(uint32_t c = 0; c < ncols; c++) cols[c] = c;
  const size_t chunk_recs = (64u << 20) / rec_bytes;
  char *buf = malloc(chunk_recs * rec_bytes);
  if(!cols || !buf) mcx_die("Out of memory");
  for(uint64_t at = 0; at < nrec; at += chunk_recs) {
    const uint64_t n = nrec - at < chunk_recs ? nrec - at : chunk_recs;
    r = mcx_graph_export_read(ginit.gs[d], at, n, buf);
    if(r) die_mcx(r, "mcx_graph_export_read");
    r = mcx_graph_load_records(ginit.gs[0], buf, n, ncols, MCX_MEM_HOST, cols, cols, ncols, 0, NULL, NULL);
    if(r) die_mcx(r, "merging replicas");
  }
  free(buf); free(cols);
  mcx_graph_export_end(ginit.gs[d]);
  mcx_graph_destroy(ginit.gs[d]); ginit.gs[d] = NULL;
  { char a[64]; mcx_ulong_to_str(nrec, a); mcx_status("[replica] merged %s kmers of the graph on GPU %i into GPU %i", a, devices[d], devices[0]); }
}

static bool graph_ready(void *ctx) { (void)ctx; return __atomic_load_n(&ginit.done, __ATOMIC_ACQUIRE) != 0; }
static mcx_graph *graph_wait(void *ctx)
{
  (void)ctx;
  static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  pthread_mutex_lock(&mu); /* several file threads may ask */
  if(!ginit.joined) {
    pthread_join(ginit.thread, NULL); ginit.joined = true;
    if(ginit.rc == MCX_ERR_NO_DEVICE && !ginit.g && !strcmp(ginit.what, "device")) mcx_die("No CUDA device: "CMD" has no CPU fallback");
    if(ginit.rc) die_mcx(ginit.rc, ginit.what);
    char a[64]; mcx_ulong_to_str(ginit.capacity, a);
    for(int d = 0; d < ndevices; d++) {
      if(shard_mode) mcx_status("[hasht] Allocating shard %i of a device table with %s entries on GPU %i", d, a, devices[d]);
      else mcx_status("[hasht] Allocating device table with %s entries on GPU %i", a, devices[d]);
    }
    mcx_phase("cuda init + table (joined)");
  }
  pthread_mutex_unlock(&mu);
  return ginit.g;
}

/* Large outputs: the records come off the device in 32 MB chunks into pinned buffers (DMA at PCIe speed) and
 * writer threads put each chunk at its place in the file with pwrite -- the reference's dump is one thread calling
 * fwrite per field (src/graph/graph_writer.c:116-127).  Returns false (nothing written) if the file is not seekable. */
#define OUT_NBUF 4
#define OUT_CHUNK (32u << 20)
typedef struct {
  int fd; char *buf[OUT_NBUF]; size_t nbytes[OUT_NBUF]; off_t off[OUT_NBUF]; int state[OUT_NBUF]; /* 0 free, 1 full, 2 being written */
  bool done; int err;
  pthread_mutex_t mu; pthread_cond_t cv;
} OutPipe;
static void *out_writer(void *arg)
{
  OutPipe *op = arg;
  for(;;) {
    int s = -1;
    pthread_mutex_lock(&op->mu);
    for(;;) {
      for(int i = 0; i < OUT_NBUF; i++) if(op->state[i] == 1) { s = i; break; }
      if(s >= 0 || op->done) break;
      pthread_cond_wait(&op->cv, &op->mu);
    }
    if(s < 0) { pthread_mutex_unlock(&op->mu); return NULL; }
    op->state[s] = 2;
    pthread_mutex_unlock(&op->mu);
    size_t at = 0; int err = 0;
    while(at < op->nbytes[s]) {
      ssize_t w = pwrite(op->fd, op->buf[s] + at, op->nbytes[s] - at, op->off[s] + (off_t)at);
      if(w <= 0) { if(w < 0 && errno == EINTR) continue; err = 1; break; }
      at += (size_t)w;
    }
    pthread_mutex_lock(&op->mu);
    if(err) op->err = 1;
    op->state[s] = 0;
    pthread_cond_broadcast(&op->cv);
    pthread_mutex_unlock(&op->mu);
  }
}
static bool write_records_parallel(mcx_graph *g, FILE *fh, uint64_t nrec, uint32_t rec_bytes)
{
  if(fflush(fh) != 0) return false;
  const off_t base = ftello(fh);
  if(base < 0) return false;
  OutPipe op; memset(&op, 0, sizeof(op));
  op.fd = fileno(fh);
  for(int i = 0; i < OUT_NBUF; i++)
    if(mcx_host_alloc((void **)&op.buf[i], OUT_CHUNK) != MCX_OK) { while(i-- > 0) mcx_host_free(op.buf[i]); return false; }
  pthread_mutex_init(&op.mu, NULL); pthread_cond_init(&op.cv, NULL);
  pthread_t th[OUT_NBUF];
  for(int i = 0; i < OUT_NBUF; i++) if(pthread_create(&th[i], NULL, out_writer, &op) != 0) mcx_die("Cannot start a thread");
  uint64_t chunk_recs = OUT_CHUNK / rec_bytes;
  if(getenv("MCX_OUT_CHUNK_RECS") && atol(getenv("MCX_OUT_CHUNK_RECS")) > 0 && (uint64_t)atol(getenv("MCX_OUT_CHUNK_RECS")) < chunk_recs)
    chunk_recs = (uint64_t)atol(getenv("MCX_OUT_CHUNK_RECS")); /* tests: many small chunks */
  for(uint64_t at = 0; at < nrec; at += chunk_recs) {
    const uint64_t n = nrec - at < chunk_recs ? nrec - at : chunk_recs;
    int s = -1;
    pthread_mutex_lock(&op.mu);
    for(;;) {
      for(int i = 0; i < OUT_NBUF; i++) if(op.state[i] == 0) { s = i; break; }
      if(s >= 0) break;
      pthread_cond_wait(&op.cv, &op.mu);
    }
    pthread_mutex_unlock(&op.mu);
    int r = mcx_graph_export_read(g, at, n, op.buf[s]);
    if(r) die_mcx(r, "mcx_graph_export_read");
    pthread_mutex_lock(&op.mu);
    op.nbytes[s] = (size_t)(n * rec_bytes); op.off[s] = base + (off_t)(at * rec_bytes); op.state[s] = 1;
    pthread_cond_broadcast(&op.cv);
    pthread_mutex_unlock(&op.mu);
  }
  pthread_mutex_lock(&op.mu); op.done = true; pthread_cond_broadcast(&op.cv); pthread_mutex_unlock(&op.mu);
  for(int i = 0; i < OUT_NBUF; i++) pthread_join(th[i], NULL);
  for(int i = 0; i < OUT_NBUF; i++) mcx_host_free(op.buf[i]);
  pthread_mutex_destroy(&op.mu); pthread_cond_destroy(&op.cv);
  if(op.err) mcx_die("Cannot write to file");
  if(fseeko(fh, base + (off_t)(nrec * rec_bytes), SEEK_SET) != 0) mcx_die("Cannot write to file");
  return true;
}

static int ctx_build(int argc, char **argv)
{
  size_t i, s, t;
  mcx_phase("start");
  parse_args(argc, argv);

  size_t max_kmers = 0;
  for(i = 0; i < ngfiles; i++) {
    mcx_status("[FileFilter] Reading file %s [%u src colour%s]", gfiles[i]->path, gfiles[i]->num_of_cols, gfiles[i]->num_of_cols == 1 ? "" : "s");
    max_kmers += gfiles[i]->num_of_kmers < 0 ? 0 : (size_t)gfiles[i]->num_of_kmers;
  }
  for(s = t = 0; s < nsamples || t < ntasks;) {
    if(t == ntasks || (s < nsamples && sample_cols[s] <= tasks[t].prefs.colour)) { mcx_status("[sample] %zu: %s", s, sample_names[s]); s++; }
    else {
      char off[32] = "auto-detect", cut[32] = "off", hp[32] = "off";
      if(tasks[t].prefs.fq_offset) sprintf(off, "%u", tasks[t].prefs.fq_offset);
      if(tasks[t].prefs.fq_cutoff) sprintf(cut, "%u", tasks[t].prefs.fq_cutoff);
      if(tasks[t].prefs.hp_cutoff) sprintf(hp, "%u", tasks[t].prefs.hp_cutoff);
      mcx_status("[task] %s%s%s; FASTQ offset: %s, threshold: %s; cut homopolymers: %s; remove PCR duplicates: %s; colour: %u\n",
                 mcx_seq_path(tasks[t].file), tasks[t].file2 ? ", " : "", tasks[t].file2 ? mcx_seq_path(tasks[t].file2) : "",
                 off, cut, hp, tasks[t].prefs.remove_pcr ? "yes" : "no", tasks[t].prefs.colour);
      t++;
    }
  }

  /* src/commands/ctx_build.c:285-289 + src/basic/async_read_io.c:313-334: 5 x file bytes */
  for(t = 0; t < ntasks; t++) {
    int64_t fsize = mcx_seq_file_size(tasks[t].file), fsize2 = tasks[t].file2 ? mcx_seq_file_size(tasks[t].file2) : 0;
    if(fsize < 0 || fsize2 < 0) { max_kmers = SIZE_MAX; break; }
    max_kmers += (size_t)(fsize + fsize2) * 5;
  }

  /* src/commands/ctx_build.c:293-303: intersecting: every read only updates k-mers of the intersection
   * graphs, and those bound the table */
  /* src/commands/ctx_build.c:259-261 */
  bool remove_pcr_used = false;
  for(t = 0; t < ntasks; t++) remove_pcr_used |= tasks[t].prefs.remove_pcr;
  if(nifiles > 0) {
    if(remove_pcr_used) usage_err("Cannot use --remove-pcr and --intersect"); /* ctx_build.c:294-295 */
    for(t = 0; t < ntasks; t++) tasks[t].prefs.must_exist = true;
    max_kmers = 0;
    for(i = 0; i < nifiles; i++) max_kmers += ifiles[i]->num_of_kmers < 0 ? 0 : (size_t)ifiles[i]->num_of_kmers;
  }

  /* src/commands/ctx_build.c:311-322 */
  size_t W = (kmer_size + 31) / 32, graph_mem;
  /* (the reference's arithmetic, so that -m gives the same number of k-mers: its read-start marks are two
   * bits per k-mer; the device keeps two u32 per slot on top of that) */
  size_t bits_per_kmer = W * 64 + (32 + 8) * output_colours + (nifiles > 0 ? 8 : 0) + (remove_pcr_used ? 2 : 0) + (sort_kmers ? 64 : 0);
  size_t kmers_in_hash = mcx_get_kmers_in_hash(mem_to_use, mem_set, num_kmers, nkmers_set, bits_per_kmer, 0,
                                               (int64_t)max_kmers, true, &graph_mem);
  if(graph_mem > mem_to_use) { char m[64]; mcx_bytes_to_str(graph_mem, m); mcx_die("Need to set higher memory limit [ at least -m %s ]", m); }

  create_output(out_path);
  mcx_status("Writing %zu colour graph to %s\n", output_colours, strcmp(out_path, "-") ? out_path : "STDOUT");

  /* CUDA start-up, the context and the table take 0.5-1.5 s: they run on a second thread while this one
   * starts parsing the first sequence file (seq_ingest.c keeps the parsed batches until the graph exists) */
  if(ndevices > 1 && remove_pcr_used) mcx_die("--remove-pcr needs the reads in order on one table: use one device");
  if(shard_mode) {
    if(ndevices < 2) mcx_die("--shard needs several devices: -D 0,1[,..]");
    if(nifiles > 0 || ngfiles > 0) mcx_die("--shard builds from sequence only (no --graph / --intersect)");
  }
  ginit.k = (uint32_t)kmer_size; ginit.ncols = (uint32_t)output_colours; ginit.capacity = kmers_in_hash; ginit.device = device;
  ginit.flags = (nifiles > 0 ? MCX_GRAPH_INTERSECT : 0) | (remove_pcr_used ? MCX_GRAPH_READSTRT : 0);
  ginit.host_batches = ntasks > 0;
  if(pthread_create(&ginit.thread, NULL, graph_init_main, NULL) != 0) mcx_die("Cannot start a thread");
  mcx_graph_source.wait = graph_wait; mcx_graph_source.ready = graph_ready; mcx_graph_source.ctx = NULL;
  if(shard_mode) { mcx_ingest.submit = submit_shard; mcx_ingest.sync = sync_shards; }
  else if(ndevices > 1) { mcx_ingest.route = route_replica; mcx_ingest.sync = sync_replicas; }
  mcx_graph *g = NULL;
  int r = 0;
  if(nifiles > 0 || ngfiles > 0 || remove_pcr_used) g = graph_wait(NULL); /* these need the device right away */

  McxGInfo *ginfo = calloc(output_colours, sizeof(McxGInfo));
  for(i = 0; i < output_colours; i++) mcx_ginfo_init(&ginfo[i]);

  /* src/commands/ctx_build.c:362-377: graph files first (their header metadata is merged into the
   * colours they load into), then the --sample names OVERWRITE the names of the colours they name */
  /* src/commands/ctx_build.c:346-361: intersection graphs first -- k-mers and one flattened edge set, no
   * coverage, no header metadata */
  for(i = 0; i < nifiles; i++) {
    if(ifiles[i]->kmer_size != kmer_size)
      mcx_die("Graph has different kmer size [kmer_size: %u vs %zu; path: %s]", ifiles[i]->kmer_size, kmer_size, ifiles[i]->path);
    for(int d = 0; d < ndevices; d++) { /* every replica looks its reads up in the intersection graph */
      r = mcx_ctx_load(ginit.gs[d], ifiles[i], NULL, output_colours, MCX_LOAD_INTO_ISEC, NULL, NULL, NULL);
      if(r) die_mcx(r, "loading intersection graph");
      mcx_load_stats st;
      r = mcx_graph_sync(ginit.gs[d], &st);
      if(r) die_mcx(r, "loading intersection graph");
    }
    mcx_ctx_close(ifiles[i]);
  }
  for(i = 0; i < ngfiles; i++) {
    uint64_t nread = 0, nloaded = 0, nnovel = 0;
    r = mcx_ctx_load(g, gfiles[i], ginfo, output_colours, nifiles > 0 ? (MCX_LOAD_MUST_EXIST | MCX_LOAD_MASK_ISEC) : 0u,
                     &nread, &nloaded, &nnovel);
    if(r) die_mcx(r, "loading graph file");
    mcx_load_stats st; /* fold the novel k-mers of the file into the table occupancy */
    r = mcx_graph_sync(g, &st);
    if(r) die_mcx(r, "loading graph file");
    uint64_t nk0 = 0, cap0 = 0; mcx_graph_stats(g, &nk0, &cap0);
    { char a[64], b[64]; mcx_ulong_to_str(nk0, a); mcx_ulong_to_str(cap0, b);
      mcx_status("[hasht] table occupancy: %s / %s (%.2f%%)", a, b, cap0 ? 100.0 * nk0 / cap0 : 0.0); }
    mcx_ctx_close(gfiles[i]);
  }
  for(i = 0; i < nsamples; i++) mcx_ginfo_set_name(&ginfo[sample_cols[i]], sample_names[i]);

  /* src/commands/ctx_build.c:389-407: build_graph() on batches of <= 10 tasks.  Quirk Q1
   * (src/tools/build_graph.c:242 vs :276,:285-300): within one call every task's header
   * statistics are credited to the batch's FIRST task, i.e. to that task's colour. */
  /* With --remove-pcr anywhere (ctx_build.c:386-403): one call per run of <= 10 tasks of ONE colour, and the
   * read-start marks are wiped when the colour changes.  The reference reads the files of one call
   * concurrently; so does this driver for consecutive files of one colour (not with --remove-pcr: order matters). */
  size_t start, end; uint32_t prev_colour = 0;
  for(start = 0; start < ntasks; start = end) {
    end = start + MAX_IO_THREADS < ntasks ? start + MAX_IO_THREADS : ntasks;
    if(remove_pcr_used) {
      uint32_t colour = tasks[start].prefs.colour;
      if(colour != prev_colour) { r = mcx_graph_pcr_reset(g); if(r) die_mcx(r, "mcx_graph_pcr_reset"); }
      end = start + 1;
      while(end < ntasks && end - start < MAX_IO_THREADS && tasks[end].prefs.colour == colour) end++;
      prev_colour = colour;
    }
    mcx_load_stats credited; memset(&credited, 0, sizeof(credited));
    for(t = start; t < end; t++) {
      /* consecutive files of one colour are read at the same time, like the reader threads of one build_graph() call
       * (src/basic/async_read_io.c); their counters cannot be told apart, so the totals go to the first of them --
       * where the reference's own bookkeeping puts them anyway (quirk Q1) */
      size_t run_end = t + 1;
      while(!tasks[t].prefs.remove_pcr && run_end < end && !tasks[run_end].prefs.remove_pcr &&
            tasks[run_end].prefs.colour == tasks[t].prefs.colour) run_end++;
      if(run_end - t > 1 && !(getenv("MCX_FILE_THREADS") && atoi(getenv("MCX_FILE_THREADS")) == 1)) {
        FileLoad fl[MAX_IO_THREADS];
        mcx_ingest.concurrent = true; mcx_ingest.nfiles = (int)(run_end - t);
        for(size_t i = t; i < run_end; i++) {
          fl[i - t].task = &tasks[i]; fl[i - t].g = g; fl[i - t].rc = 0;
          if(pthread_create(&fl[i - t].thread, NULL, file_load_main, &fl[i - t]) != 0) mcx_die("Cannot start a thread");
        }
        for(size_t i = t; i < run_end; i++) { pthread_join(fl[i - t].thread, NULL); if(fl[i - t].rc) die_mcx(fl[i - t].rc, "loading sequence"); }
        mcx_ingest.concurrent = false; mcx_ingest.nfiles = 0;
        g = graph_wait(NULL);
        mcx_load_stats st;
        r = mcx_sync_reads(g, &st);
        if(r) die_mcx(r, "loading sequence");
        mcx_add_load_stats(&tasks[t].stats, &st);
        mcx_phase("sequence files loaded");
        credited.total_bases_loaded += tasks[t].stats.total_bases_loaded;
        credited.contigs_parsed += tasks[t].stats.contigs_parsed;
        t = run_end - 1;
        continue;
      }
      if(tasks[t].prefs.remove_pcr)
        r = mcx_load_seq_pcr(g, tasks[t].file, tasks[t].file2, tasks[t].interleaved, &tasks[t].prefs, &tasks[t].stats);
      else
      r = mcx_load_seq_file(g, tasks[t].file, &tasks[t].prefs, &tasks[t].stats); /* g may still be NULL: see mcx_graph_source */
      if(r) die_mcx(r, "loading sequence");
      mcx_phase("sequence file loaded");
      credited.total_bases_loaded += tasks[t].stats.total_bases_loaded;
      credited.contigs_parsed += tasks[t].stats.contigs_parsed;
    }
    mcx_ginfo_update_contigs(&ginfo[tasks[start].prefs.colour], credited.total_bases_loaded, credited.contigs_parsed);
  }

  g = graph_wait(NULL);
  /* several devices: the replicas' tables are folded into the first one (graph files went there alone) */
  for(int d = 1; d < ndevices && !shard_mode; d++) merge_replica(d, (uint32_t)output_colours);
  if(ndevices > 1 && !shard_mode) { mcx_load_stats st; r = mcx_graph_sync(g, &st); if(r) die_mcx(r, "merging replicas"); mcx_phase("replicas merged"); }
  /* src/commands/ctx_build.c:409-413 */
  if(nifiles > 0) {
    r = mcx_graph_finish_intersect(g, NULL);
    if(r) die_mcx(r, "mcx_graph_finish_intersect");
  }

  uint64_t nk = 0, cap = 0;
  if(shard_mode) mcx_shardset_stats(ginit.ss, &nk, &cap); else mcx_graph_stats(g, &nk, &cap);
  { char a[64], b[64]; mcx_ulong_to_str(nk, a); mcx_ulong_to_str(cap, b);
    mcx_status("[hasht] table occupancy: %s / %s (%.2f%%)", a, b, cap ? 100.0 * nk / cap : 0.0); }
  for(t = 0; t < ntasks; t++) { print_task_stats(&tasks[t]); mcx_seq_close(tasks[t].file); mcx_seq_close(tasks[t].file2); }

  mcx_status("Dumping graph...\n");
  FILE *fh = strcmp(out_path, "-") ? fopen(out_path, "w") : stdout;
  if(!fh) mcx_die("Cannot open file: %s [%s]", out_path, strerror(errno));
  setvbuf(fh, NULL, _IOFBF, 4u << 20);
  mcx_write_ctx_header(fh, (uint32_t)kmer_size, (uint32_t)output_colours, ginfo);

  uint64_t nrec = 0; uint32_t rec_bytes = 0;
  if(shard_mode) {
    /* every shard dumps its records (sorted: ascending keys); the file is the merge of the P runs, made while writing
     * (the reference iterates ONE table: HASH_ITERATE_SORTED, src/graph/hash_table.c:362-374) */
    r = mcx_shardset_export_begin(ginit.ss, sort_kmers ? 1 : 0, &nrec, &rec_bytes);
    if(r) die_mcx(r, "mcx_shardset_export_begin");
    mcx_phase("export: compact + sort (per shard)");
    const uint64_t chunk_recs = (64u << 20) / rec_bytes;
    char *buf = malloc(chunk_recs * rec_bytes);
    if(!buf) mcx_die("Out of memory");
    uint64_t written = 0, got = 0;
    do {
      r = mcx_shardset_export_next(ginit.ss, buf, chunk_recs, &got);
      if(r) die_mcx(r, "mcx_shardset_export_next");
      if(got && fwrite(buf, rec_bytes, got, fh) != got) mcx_die("Cannot write to file");
      written += got;
    } while(got);
    free(buf);
    if(written != nrec) mcx_die("sharded dump: %llu of %llu records", (unsigned long long)written, (unsigned long long)nrec);
    mcx_shardset_export_end(ginit.ss);
    if(fh != stdout) fclose(fh); else fflush(fh);
    mcx_phase("export: merge + write");
    { char a[64]; mcx_ulong_to_str(nrec, a);
      mcx_status("[graphwriter] Dumped %s kmers in %zu colour%s into: %s (format version: 6)", a, output_colours,
                 output_colours == 1 ? "" : "s", strcmp(out_path, "-") ? out_path : "STDOUT"); }
    for(i = 0; i < output_colours; i++) mcx_ginfo_free(&ginfo[i]);
    free(ginfo); free(tasks); free(sample_names);
    mcx_shardset_destroy(ginit.ss);
    mcx_phase("destroy");
    return EXIT_SUCCESS;
  }
  r = mcx_graph_export_begin(g, sort_kmers ? 1 : 0, &nrec, &rec_bytes);
  if(r) die_mcx(r, "mcx_graph_export_begin");
  mcx_phase("export: compact + sort");
  uint64_t pipe_min = 256u << 20; /* MCX_OUT_PIPE_MIN=<bytes> overrides (tests) */
  if(getenv("MCX_OUT_PIPE_MIN")) pipe_min = (uint64_t)atoll(getenv("MCX_OUT_PIPE_MIN"));
  const bool big_regular = fh != stdout && (uint64_t)nrec * rec_bytes >= pipe_min;
  if(!big_regular || !write_records_parallel(g, fh, nrec, rec_bytes)) {
    size_t chunk_recs = (64u << 20) / rec_bytes;
    char *buf = malloc(chunk_recs * rec_bytes);
    for(uint64_t at = 0; at < nrec; at += chunk_recs) {
      uint64_t n = nrec - at < chunk_recs ? nrec - at : chunk_recs;
      r = mcx_graph_export_read(g, at, n, buf);
      if(r) die_mcx(r, "mcx_graph_export_read");
      if(fwrite(buf, rec_bytes, n, fh) != n) mcx_die("Cannot write to file");
    }
    free(buf);
  }
  mcx_graph_export_end(g);
  if(fh != stdout) fclose(fh); else fflush(fh);
  mcx_phase("export: D2H + write");
  { char a[64]; mcx_ulong_to_str(nrec, a);
    mcx_status("[graphwriter] Dumped %s kmers in %zu colour%s into: %s (format version: 6)", a, output_colours,
               output_colours == 1 ? "" : "s", strcmp(out_path, "-") ? out_path : "STDOUT"); }

  for(i = 0; i < output_colours; i++) mcx_ginfo_free(&ginfo[i]);
  free(ginfo); free(tasks); free(sample_names);
  mcx_graph_destroy(g);
  mcx_phase("destroy");
  return EXIT_SUCCESS;
}

/* src/main/mccortex.c:255-277: strip -q / --quiet anywhere on the line */
static bool remove_quiet_flags(int *argcp, char **argv)
{
  bool q = false; int i, j, argc = *argcp;
  for(i = j = 1; i < argc; i++) {
    if(!strcmp(argv[i], "--quiet") || !strcmp(argv[i], "-q")) { q = true; continue; }
    if(argv[i][0] == '-' && argv[i][1] != '-') {
      char *p, *w;
      for(p = w = argv[i] + 1; *p; p++) { if(*p == 'q') q = true; else *w++ = *p; }
      *w = '\0';
    }
    argv[j++] = argv[i];
  }
  *argcp = j;
  return q;
}

int main(int argc, char **argv)
{
  time_t t0 = time(NULL);
  mcx_msg_out = stderr;
  bool is_sort = argc >= 2 && strcasecmp(argv[1], "sort") == 0, is_join = argc >= 2 && strcasecmp(argv[1], "join") == 0;
  if(argc < 2 || (strcasecmp(argv[1], "build") != 0 && !is_sort && !is_join)) {
    fprintf(stderr, "\nusage: "CMD" <build|sort|join> [options] <out.ctx>\n"
                    "  build   construct a CORTEX v6 graph file on a B200 for the rest of McCortex\n"
                    "  sort    sort the k-mers of a graph file\n"
                    "  join    merge graph files\n\n");
    return EXIT_FAILURE;
  }
  if(argc == 2 && !is_sort && !is_join) mcx_print_usage(build_usage, NULL);
  /* command line for the log, before -q is stripped */
  size_t len = 0; int i;
  for(i = 0; i < argc; i++) len += strlen(argv[i]) + 1;
  char *line = malloc(len + 1); line[0] = 0;
  for(i = 0; i < argc; i++) { strcat(line, argv[i]); if(i + 1 < argc) strcat(line, " "); }
  if(remove_quiet_flags(&argc, argv)) mcx_msg_out = NULL;
  mcx_status("[cmd] %s", line);
  { char cwd[4096]; if(getcwd(cwd, sizeof(cwd))) mcx_status("[cwd] %s", cwd); }
  mcx_status("[version] "CMD" sm_100a zlib=%s k=%i..%i", ZLIB_VERSION, MIN_KMER, MAX_KMER);
  free(line);

  char *tmp = argv[1]; argv[1] = argv[0]; argv[0] = tmp;
  int ret = is_sort ? mcx_cmd_sort(argc - 1, argv + 1) : (is_join ? mcx_cmd_join(argc - 1, argv + 1) : ctx_build(argc - 1, argv + 1));
  mcx_status(ret == 0 ? "Done." : "Fail.");
  mcx_status("[time] %.2lf seconds\n", difftime(time(NULL), t0));
  /* every output file is closed: leave without the CUDA runtime's exit handlers (tearing the context down in user
   * space takes 0.3-0.7 s; the kernel driver reclaims the device memory either way) */
  fflush(stdout); fflush(stderr);
  _exit(ret);
}
