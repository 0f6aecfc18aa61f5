/*
 * mcx_synth.c -- deterministic synthetic read generator.
 *
 * Counter-based RNG splitmix64(seed + counter), so any read can be generated
 * independently (pthreads over read ranges here; the same formulae can run anywhere).
 *   genome base j      = top 2 bits of splitmix64(S_g + j)
 *   read i             : start = splitmix64(S_r + i) mod (G - L + 1)
 *                        strand = bit 63 of splitmix64(S_r + i + 2^40) (1 = reverse complement)
 *                        base j substituted iff splitmix64(S_e + i*256 + j) < p_err * 2^64,
 *                        new = (old + 1 + ((h >> 62) % 3)) & 3   (applied to the read as emitted)
 * Output layouts: LINES (read + '\n', what libmcxgpu ingests directly) or FASTA
 * (">r\n" + read + "\n", what the reference CLI ingests).
 *
 * Built as libmcxsynth.so (ctypes, used by bench.py / tests) and as the `mcx-synth` CLI.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define S_G 0x5EED0001ull
#define S_R 0x5EED0002ull
#define S_E 0x5EED0003ull

static inline uint64_t splitmix64(uint64_t x)
{
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

/* run fn(lo, hi, arg) over [0, n) split across the host's cores */
typedef void (*range_fn)(uint64_t lo, uint64_t hi, void *arg);
typedef struct { range_fn fn; uint64_t lo, hi; void *arg; } RangeJob;
static void *range_thread(void *p) { RangeJob *j = (RangeJob*)p; j->fn(j->lo, j->hi, j->arg); return NULL; }
static void parallel_ranges(uint64_t n, range_fn fn, void *arg)
{
  long nc = sysconf(_SC_NPROCESSORS_ONLN);
  uint64_t T = nc < 1 ? 1 : (nc > 64 ? 64 : (uint64_t)nc), t;
  if(n < 4096 || T == 1) { fn(0, n, arg); return; }
  pthread_t th[64]; RangeJob jobs[64];
  for(t = 0; t < T; t++) {
    jobs[t].fn = fn; jobs[t].arg = arg; jobs[t].lo = n * t / T; jobs[t].hi = n * (t + 1) / T;
    pthread_create(&th[t], NULL, range_thread, &jobs[t]);
  }
  for(t = 0; t < T; t++) pthread_join(th[t], NULL);
}

/* genome as 2-bit codes, one per byte */
typedef struct { uint8_t *g; uint64_t seed; } GenomeJob;
static void genome_range(uint64_t lo, uint64_t hi, void *arg)
{
  GenomeJob *a = (GenomeJob*)arg; uint64_t j;
  for(j = lo; j < hi; j++) a->g[j] = (uint8_t)(splitmix64(a->seed + j) >> 62);
}
void mcx_synth_genome(uint8_t *g, uint64_t G, uint64_t seed_xor)
{
  GenomeJob a = { g, S_G ^ seed_xor };
  parallel_ranges(G, genome_range, &a);
}

/* bytes one read occupies in the output */
uint64_t mcx_synth_read_bytes(uint32_t L, int fasta) { return (uint64_t)L + 1u + (fasta ? 3u : 0u); }

/* reads [first, first+n) -> out (n * mcx_synth_read_bytes bytes) */
typedef struct { char *out; uint64_t first; uint32_t L; const uint8_t *genome; uint64_t G; double p_err; int fasta; uint64_t sx; } ReadsJob;
static void reads_range(uint64_t lo, uint64_t hi, void *arg)
{
  static const char ACGT[4] = {'A', 'C', 'G', 'T'};
  const ReadsJob *a = (const ReadsJob*)arg;
  const uint32_t L = a->L;
  const uint64_t stride = mcx_synth_read_bytes(L, a->fasta);
  /* p_err * 2^64 without overflow for p_err == 1 */
  const long double thr_ld = (long double)a->p_err * 18446744073709551616.0L;
  const uint64_t thr = thr_ld >= 18446744073709551615.0L ? UINT64_MAX : (uint64_t)thr_ld;
  uint64_t ii, j;
  for(ii = lo; ii < hi; ii++) {
    uint64_t i = a->first + ii;
    char *o = a->out + ii * stride;
    uint64_t start = splitmix64((S_R ^ a->sx) + i) % (a->G - L + 1);
    int rev = (int)(splitmix64((S_R ^ a->sx) + i + (1ull << 40)) >> 63);
    if(a->fasta) { *o++ = '>'; *o++ = 'r'; *o++ = '\n'; }
    for(j = 0; j < L; j++) {
      uint8_t b = rev ? (uint8_t)(3u - a->genome[start + L - 1 - j]) : a->genome[start + j];
      if(a->p_err > 0) {
        uint64_t h = splitmix64((S_E ^ a->sx) + i * 256u + j);
        if(h < thr) b = (uint8_t)((b + 1u + ((h >> 62) % 3u)) & 3u);
      }
      o[j] = ACGT[b];
    }
    o[L] = '\n';
  }
}
void mcx_synth_reads(char *out, uint64_t first, uint64_t n, uint32_t L, const uint8_t *genome, uint64_t G,
                     double p_err, int fasta, uint64_t seed_xor)
{
  ReadsJob a = { out, first, L, genome, G, p_err, fasta, seed_xor };
  parallel_ranges(n, reads_range, &a);
}

#ifdef MCX_SYNTH_MAIN
/* mcx-synth <G> <first_read> <nreads> <L> <p_err> <fasta:0|1> [seed_xor] > out */
int main(int argc, char **argv)
{
  if(argc < 7) { fprintf(stderr, "usage: %s <G> <first> <nreads> <L> <p_err> <fasta 0|1> [seed_xor]\n", argv[0]); return 2; }
  uint64_t G = strtoull(argv[1], 0, 10), first = strtoull(argv[2], 0, 10), n = strtoull(argv[3], 0, 10);
  uint32_t L = (uint32_t)atoi(argv[4]); double p = atof(argv[5]); int fasta = atoi(argv[6]);
  uint64_t sx = argc > 7 ? strtoull(argv[7], 0, 0) : 0;
  uint8_t *g = (uint8_t*)malloc(G);
  mcx_synth_genome(g, G, sx);
  const uint64_t block = 1u << 18, stride = mcx_synth_read_bytes(L, fasta);
  char *buf = (char*)malloc(block * stride);
  for(uint64_t at = 0; at < n; at += block) {
    uint64_t m = n - at < block ? n - at : block;
    mcx_synth_reads(buf, first + at, m, L, g, G, p, fasta, sx);
    if(fwrite(buf, stride, m, stdout) != m) { perror("write"); return 1; }
  }
  free(buf); free(g);
  return 0;
}
#endif
