// mcx_abi.cu -- extern "C" layer of libmcxgpu.so (see include/mcx_gpu.h).
//
// Owns: the device table, counters, the H2D staging ring for host batches, and the
// export buffer.  No CPU implementation of any part of the hot path lives here: if
// there is no CUDA device every entry point fails with MCX_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/mcx_gpu.h"
#include "mcx_build.h"
#include "mcx_chunk.cuh"

#define MCX_NSTAGE 3
#define MCX_STAGE_POS (32ull << 20)                 /* positions per staged piece of a host batch */
#define MCX_STAGE_BYTES (MCX_STAGE_POS + 256ull)    /* + look-back / look-ahead / alignment slack */

static thread_local char g_err[256] = "";
static int fail_cuda(cudaError_t e, const char *what)
{
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return e == cudaErrorMemoryAllocation ? MCX_ERR_NOMEM : MCX_ERR_CUDA;
}
#define CU(x) do { cudaError_t e_ = (x); if(e_ != cudaSuccess) return fail_cuda(e_, #x); } while(0)

struct mcx_graph {
  int device;
  uint32_t k, W, ncols;
  uint64_t capacity;
  McxTable table;
  unsigned long long *d_counters;
  cudaStream_t own_primary;            // used unless the caller installs a stream
  cudaStream_t streams[MCX_NSTAGE];    // staging ring: H2D + kernel of one piece per slot
  cudaEvent_t events[MCX_NSTAGE];
  cudaEvent_t ev_fork;
  uint8_t *d_stage[MCX_NSTAGE];
  uint8_t *h_stage[MCX_NSTAGE];
  bool stage_ready;
  int next;
  cudaStream_t user_stream; bool use_user_stream;
  uint64_t occ_bound;      // upper bound on occurrences ever sent (saturation guard)
  uint64_t pend_positions; // byte positions queued since the last sync
  uint64_t pend_offsets_reads, pend_offsets_bases; // OFFSETS batches: counted on the host
  uint64_t nkmers;         // slots claimed so far (updated at sync)
  McxExport exp; bool exp_valid;
  uint8_t *d_tmp; size_t d_tmp_bytes; // scratch of the batch being queued (OFFSETS -> LINES repack, quality batches, graph records):
  uint8_t *d_tmpv[2]; size_t d_tmp_bytesv[2]; cudaEvent_t ev_tmp[2]; int tmp_cur; // one of two, so that batch b+1 is queued while b runs
  uint8_t *d_isec;         // build --intersect: one edge byte per slot (Edges *isec_edges, ctx_build.c:341-343), else NULL
  uint64_t front_pending;  // positions queued since the front table was last flushed (its counters are 32-bit)
  uint32_t *d_first;       // build --remove-pcr: first read ordinal per (slot, orientation) (mcx_pcr.cuh), else NULL
  uint32_t pcr_ord;        // ordinal of the next read of this colour
  uint8_t *d_pcr; size_t d_pcr_bytes; // --remove-pcr: device copy of the batch being filtered
  bool sharded;  // front table holds records of keys owned by other shards: only mcx_graph_flush_sharded may empty it
  // `make EXPERIMENTS=1` builds only: MCX_KERNEL=warp selects kernel A2 (mcx_build_warp.cu) for inserting builds with
  // k <= 31 and no quality / homopolymer cut-off; MCX_CLASSES=2|4: one front table per key class, one launch per class
  bool warp_kernel;
  uint32_t ncls_log2;
};

// front table geometry: ncls class tables of (4 << bits) slots each; all tags first, then all counters
static size_t front_slots(const mcx_graph *g) { return (size_t)4u << g->table.front_set_bits; }
static size_t front_bytes(const mcx_graph *g) { return (front_slots(g) << g->ncls_log2) * 12u; }
static McxTable class_table(const mcx_graph *g, uint32_t c)
{
  McxTable t = g->table;
  if(t.front) { t.front = g->table.front + c * front_slots(g); t.front_cnt = g->table.front_cnt + c * front_slots(g); }
  return t;
}
// merge every class's front table into the big table (or, sharded, into the owners' bins)
static cudaError_t flush_front(mcx_graph *g, cudaStream_t st, const McxTupleBins *bins = nullptr)
{
  if(!g->table.front_set_bits) return cudaSuccess;
  for(uint32_t c = 0; c < (1u << g->ncls_log2); c++) {
    const McxTable t = class_table(g, c);
    cudaError_t e = bins ? mcx_launch_front_flush_sharded(t, *bins, g->occ_bound >= 0xF0000000ull, g->d_counters, st)
                         : mcx_launch_front_flush(t, g->occ_bound >= 0xF0000000ull, g->d_counters, st);
    if(e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

extern "C" int mcx_device_count(void)
{
  int n = 0;
  if(cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
extern "C" const char *mcx_last_error(void) { return g_err; }

extern "C" int mcx_host_alloc(void **ptr, size_t bytes)
{
  if(!ptr) return MCX_ERR_BAD_ARG;
  if(mcx_device_count() == 0) return MCX_ERR_NO_DEVICE;
  CU(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return MCX_OK;
}
extern "C" int mcx_host_free(void *ptr) { if(ptr) CU(cudaFreeHost(ptr)); return MCX_OK; }

// All work of a graph is ordered on its primary stream (the caller's, or our own).  The
// host-staging path fans out to the ring streams and joins back (event fork/join), so
// events recorded on the primary stream bracket everything a call enqueued.
static cudaStream_t primary(mcx_graph *g) { return g->use_user_stream ? g->user_stream : g->own_primary; }
// MCX_TIMING=1: wall clock of the steps of mcx_graph_create on stderr
#include <time.h>
static void abi_phase(const char *what)
{
  static int on = -1; static struct timespec last;
  struct timespec now;
  if(on < 0) { on = getenv("MCX_TIMING") != NULL; clock_gettime(CLOCK_MONOTONIC, &last); }
  if(!on) return;
  clock_gettime(CLOCK_MONOTONIC, &now);
  fprintf(stderr, "[phase]     %-24s +%.3f s\n", what, (double)(now.tv_sec - last.tv_sec) + 1e-9 * (double)(now.tv_nsec - last.tv_nsec));
  last = now;
}

extern "C" int mcx_graph_create(uint32_t k, uint32_t ncols, uint64_t capacity, int device, uint32_t flags, mcx_graph **out)
{
  if(!out || k < 3 || k > 63 || !(k & 1u) || ncols == 0 || ncols > 4096 || capacity == 0) return MCX_ERR_BAD_ARG;
  abi_phase("create: enter");
  int ndev = mcx_device_count();
  abi_phase("create: device count");
  if(ndev == 0) { snprintf(g_err, sizeof(g_err), "no CUDA device: libmcxgpu has no CPU fallback"); return MCX_ERR_NO_DEVICE; }
  if(device < 0 || device >= ndev) return MCX_ERR_BAD_ARG;
  CU(cudaSetDevice(device));
  CU(cudaFree(0));
  abi_phase("create: context");
  mcx_graph *g = (mcx_graph *)calloc(1, sizeof(*g));
  if(!g) return MCX_ERR_NOMEM;
  g->device = device; g->k = k; g->W = (k + 31u) / 32u; g->ncols = ncols;
  g->capacity = capacity;
  g->table.nslots = (capacity + 1ull) & ~1ull;
  g->table.stride = mcx_slot_words(g->W, ncols);
  g->table.ncols = ncols;
  size_t bytes = (size_t)g->table.nslots * g->table.stride * 4u;
  cudaError_t e = cudaMalloc(&g->table.slots, bytes);
  if(e == cudaErrorMemoryAllocation) { cudaGetLastError(); mcx_pool_trim(device); e = cudaMalloc(&g->table.slots, bytes); } // (export scratch cached by the pool)
  if(e != cudaSuccess) { free(g); return fail_cuda(e, "cudaMalloc(table)"); }
  abi_phase("create: table malloc");
  e = cudaMalloc(&g->d_counters, MCX_NCOUNTERS_ALL * sizeof(unsigned long long));
  if(e != cudaSuccess) { cudaFree(g->table.slots); free(g); return fail_cuda(e, "cudaMalloc(counters)"); }
  cudaStreamCreateWithFlags(&g->own_primary, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&g->ev_fork, cudaEventDisableTiming);
  for(int i = 0; i < MCX_NSTAGE; i++) {
    cudaStreamCreateWithFlags(&g->streams[i], cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&g->events[i], cudaEventDisableTiming);
  }
  cudaMemsetAsync(g->table.slots, 0, bytes, g->own_primary);
  cudaMemsetAsync(g->d_counters, 0, MCX_NCOUNTERS_ALL * sizeof(unsigned long long), g->own_primary);
  e = cudaStreamSynchronize(g->own_primary);
  if(e != cudaSuccess) { int r = fail_cuda(e, "memset(table)"); mcx_graph_destroy(g); return r; }
  abi_phase("create: streams + memset");
  // front table (it counts one colour at a time and is flushed when the colour changes): sized to sit in L2 (64 MB = 2^21 sets of four 8-byte
  // slots); MCX_FRONT_BITS=0 disables it, other values are for experiments
  if(flags & MCX_GRAPH_INTERSECT) {
    e = cudaMalloc(&g->d_isec, (size_t)g->table.nslots + 8);
    if(e != cudaSuccess) { int r = fail_cuda(e, "cudaMalloc(isec_edges)"); mcx_graph_destroy(g); return r; }
    cudaMemset(g->d_isec, 0, (size_t)g->table.nslots + 8);
  }
  if(flags & MCX_GRAPH_READSTRT) {
    e = cudaMalloc(&g->d_first, (size_t)g->table.nslots * 8u);
    if(e != cudaSuccess) { int r = fail_cuda(e, "cudaMalloc(read starts)"); mcx_graph_destroy(g); return r; }
    cudaMemset(g->d_first, 0xFF, (size_t)g->table.nslots * 8u);
  }
  // (an intersected build only looks k-mers up: no front table)
  if(!(flags & MCX_GRAPH_INTERSECT)) {
    // k <= 31: 2^21 sets of four 8-byte ways = 64 MB of tags + 32 MB of counters, L2-resident.  k > 31: two 16-byte ways
    // per set, so the same bytes hold half as many keys -- fewer than the 4.6 M hot k-mers of the bench genome; 2^22 sets
    // (128 MB + 32 MB, no longer L2-resident) measured faster on configs[2]: 27.2 against 22.7 G k-mers/s
    // (profiles/r2al_k63_grid.txt) -- a hot-pass hit that misses L2 is still cheaper than a parked big-table update
    uint32_t bits = g->W == 1u ? 21 : 22;
    if(const char *m = getenv("MCX_FRONT_BITS")) bits = (uint32_t)atoi(m);
    if(bits) {
      if(bits < 16) bits = 16;
      if(bits > 24) bits = 24;
      g->table.front_set_bits = bits;
      g->table.front_words = g->W; // k > 31: the same bytes hold two 16-byte ways per set (and half the counters are unused)
#ifdef MCX_EXPERIMENTS
      if(const char *m = getenv("MCX_CLASSES")) if(g->W == 1u) { int v = atoi(m); g->ncls_log2 = v >= 4 ? 2u : (v >= 2 ? 1u : 0u); }
#endif
      // one allocation: per class (4 << bits) 8-byte tags -- all classes' tags first --, then the 4-byte counters
      e = cudaMalloc(&g->table.front, front_bytes(g));
      if(e != cudaSuccess) { g->table.front_set_bits = 0; int r = fail_cuda(e, "cudaMalloc(front)"); mcx_graph_destroy(g); return r; }
      g->table.front_cnt = reinterpret_cast<unsigned int *>(g->table.front + (front_slots(g) << g->ncls_log2));
      cudaMemset(g->table.front, 0, front_bytes(g));
    }
  }
#ifdef MCX_EXPERIMENTS
  if(const char *m = getenv("MCX_KERNEL")) g->warp_kernel = strcmp(m, "warp") == 0;
#endif
  abi_phase("create: front table");
  *out = g;
  return MCX_OK;
}

extern "C" int mcx_graph_destroy(mcx_graph *g)
{
  if(!g) return MCX_OK;
  cudaSetDevice(g->device);
  cudaDeviceSynchronize();
  mcx_export_free(&g->exp);
  for(int i = 0; i < MCX_NSTAGE; i++) {
    if(g->d_stage[i]) cudaFree(g->d_stage[i]);
    if(g->h_stage[i]) cudaFreeHost(g->h_stage[i]);
    if(g->streams[i]) cudaStreamDestroy(g->streams[i]);
    if(g->events[i]) cudaEventDestroy(g->events[i]);
  }
  if(g->own_primary) cudaStreamDestroy(g->own_primary);
  if(g->ev_fork) cudaEventDestroy(g->ev_fork);
  g->d_tmpv[g->tmp_cur] = g->d_tmp;
  for(int i = 0; i < 2; i++) { if(g->d_tmpv[i]) cudaFree(g->d_tmpv[i]); if(g->ev_tmp[i]) cudaEventDestroy(g->ev_tmp[i]); }
  if(g->d_isec) cudaFree(g->d_isec);
  if(g->d_first) cudaFree(g->d_first);
  if(g->d_pcr) cudaFree(g->d_pcr);
  if(g->table.front) cudaFree(g->table.front);
  if(g->d_counters) cudaFree(g->d_counters);
  if(g->table.slots) cudaFree(g->table.slots);
  free(g);
  return MCX_OK;
}

static int sync_all(mcx_graph *g)
{
  CU(cudaSetDevice(g->device));
  for(int i = 0; i < MCX_NSTAGE; i++) CU(cudaStreamSynchronize(g->streams[i]));
  CU(cudaStreamSynchronize(primary(g)));
  return MCX_OK;
}

extern "C" int mcx_graph_clear(mcx_graph *g)
{
  if(!g) return MCX_ERR_BAD_ARG;
  CU(cudaSetDevice(g->device));
  cudaStream_t st = primary(g);
  CU(cudaMemsetAsync(g->table.slots, 0, (size_t)g->table.nslots * g->table.stride * 4u, st));
  CU(cudaMemsetAsync(g->d_counters, 0, MCX_NCOUNTERS_ALL * sizeof(unsigned long long), st));
  if(g->table.front) CU(cudaMemsetAsync(g->table.front, 0, front_bytes(g), st));
  if(g->d_isec) CU(cudaMemsetAsync(g->d_isec, 0, (size_t)g->table.nslots + 8, st));
  if(g->d_first) CU(cudaMemsetAsync(g->d_first, 0xFF, (size_t)g->table.nslots * 8u, st));
  g->pcr_ord = 0;
  g->front_pending = 0;
  g->sharded = false;
  g->occ_bound = 0; g->pend_positions = 0; g->pend_offsets_reads = g->pend_offsets_bases = 0; g->nkmers = 0;
  return MCX_OK;
}

extern "C" int mcx_graph_set_stream(mcx_graph *g, void *cuda_stream)
{
  if(!g) return MCX_ERR_BAD_ARG;
  int r = sync_all(g); if(r) return r;
  g->use_user_stream = cuda_stream != NULL;
  g->user_stream = (cudaStream_t)cuda_stream;
  return MCX_OK;
}

static int ensure_stage(mcx_graph *g)
{
  if(g->stage_ready) return MCX_OK;
  abi_phase("stage: enter");
  for(int i = 0; i < MCX_NSTAGE; i++) {
    CU(cudaMalloc(&g->d_stage[i], MCX_STAGE_BYTES));
    CU(cudaHostAlloc(&g->h_stage[i], MCX_STAGE_BYTES, cudaHostAllocDefault));
  }
  abi_phase("stage: 3 x 32 MB pinned");
  g->stage_ready = true;
  return MCX_OK;
}

extern "C" int mcx_graph_prepare_host(mcx_graph *g)
{
  if(!g) return MCX_ERR_BAD_ARG;
  CU(cudaSetDevice(g->device));
  return ensure_stage(g);
}

static McxBuildParams make_params(mcx_graph *g, const mcx_read_batch *b, const uint8_t *dseq, uint64_t nbytes,
                                  uint64_t r_begin, uint64_t r_end)
{
  McxBuildParams p;
  p.seq = dseq; p.nbytes = nbytes; p.r_begin = r_begin; p.r_end = r_end;
  p.k = g->k; p.hp_cutoff = b->hp_cutoff; p.colour = b->colour;
  p.may_saturate = g->occ_bound >= 0xF0000000ull;
  p.counters = g->d_counters;
  p.qual = nullptr; p.qcut = 0; p.summary = nullptr;
  p.cls = 0; p.ncls_log2 = 0;
  p.run_tiles = 0; if(const char *m = getenv("MCX_W_RUN")) p.run_tiles = (uint32_t)atoi(m);
  return p;
}

// The front table's counters are 32 bits wide: merge it into the big table before a counter could
// wrap, i.e. before 2^32 - 2^28 positions (an upper bound on the occurrences of any one k-mer)
// have been queued since the last flush.
#define MCX_FRONT_SPAN 0xEF000000ull  /* most positions one launch may cover */
static int front_colour(mcx_graph *g, uint32_t colour)
{
  if(!g->table.front_set_bits || g->table.front_colour == colour) return MCX_OK;
  if(g->sharded) { snprintf(g_err, sizeof(g_err), "mcx_graph_flush_sharded must run before the colour changes"); return MCX_ERR_UNSUPPORTED; }
  if(g->front_pending) {
    CU(flush_front(g, primary(g)));
    g->front_pending = 0;
  }
  // the tags carry the old colour's edge bits: start the new colour with an empty front table
  CU(cudaMemsetAsync(g->table.front, 0, front_bytes(g), primary(g)));
  g->table.front_colour = colour;
  return MCX_OK;
}
static int front_guard(mcx_graph *g, uint64_t positions)
{
  if(!g->table.front_set_bits) return MCX_OK;
  if(positions > MCX_FRONT_SPAN) { snprintf(g_err, sizeof(g_err), "batch too large: split it into pieces of < 3.7e9 bytes"); return MCX_ERR_UNSUPPORTED; }
  if(g->front_pending + positions >= 0xF0000000ull) {
    if(g->sharded) { snprintf(g_err, sizeof(g_err), "mcx_graph_flush_sharded must run at least every 4e9 positions"); return MCX_ERR_UNSUPPORTED; }
    CU(flush_front(g, primary(g)));
    g->front_pending = 0;
  }
  g->front_pending += positions;
  return MCX_OK;
}

// one launch over [r_begin, r_end) of a LINES buffer: the fused insert kernel, or (must_exist) the lookup kernel
static cudaError_t launch_build(mcx_graph *g, const mcx_read_batch *b, const McxBuildParams &p, cudaStream_t st)
{
  if(b->must_exist) return mcx_launch_build_lookup(p, g->table, st);
#ifdef MCX_EXPERIMENTS
  if(g->warp_kernel && mcx_warp_kernel_supports(p)) {
    // one launch per key class over the same reads, each with the class's own front table
    for(uint32_t c = 0; c < (1u << g->ncls_log2); c++) {
      McxBuildParams pc = p; pc.cls = c; pc.ncls_log2 = g->table.front_set_bits ? g->ncls_log2 : 0u;
      cudaError_t e = mcx_launch_build_warp(pc, class_table(g, c), st);
      if(e != cudaSuccess || !pc.ncls_log2) return e;
    }
    return cudaSuccess;
  }
#endif
  return mcx_launch_build_fused(p, g->table, st);
}

// LINES batch resident on the device
static int add_lines_device(mcx_graph *g, const mcx_read_batch *b, const uint8_t *dseq, uint64_t nbytes)
{
  if(((uintptr_t)dseq & 15u) != 0) { snprintf(g_err, sizeof(g_err), "device seq buffer must be 16-byte aligned"); return MCX_ERR_BAD_ARG; }
  g->occ_bound += nbytes;
  // one launch per span of <= MCX_FRONT_SPAN positions (the whole buffer stays visible to every
  // launch, so windows and edges across a cut see their neighbours)
  const uint64_t span = MCX_FRONT_SPAN;
  for(uint64_t lo = 0; lo < nbytes; lo += span) {
    const uint64_t hi = lo + span < nbytes ? lo + span : nbytes;
    int r = front_guard(g, hi - lo); if(r) return r;
    McxBuildParams p = make_params(g, b, dseq, nbytes, lo, hi);
    CU(launch_build(g, b, p, primary(g)));
  }
  g->pend_positions += nbytes;
  return MCX_OK;
}

// LINES batch in host memory: cut into pieces of MCX_STAGE_POS positions; each piece ships
// with 16 bytes of look-back and 80 of look-ahead so windows and edges that straddle a cut
// see their neighbours; H2D and kernels of consecutive pieces overlap on a ring of streams.
static int add_lines_host(mcx_graph *g, const mcx_read_batch *b, const uint8_t *hseq, uint64_t nbytes)
{
  int r = ensure_stage(g); if(r) return r;
  cudaPointerAttributes attr;
  bool pinned = (cudaPointerGetAttributes(&attr, hseq) == cudaSuccess) && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  g->occ_bound += nbytes;
  for(uint64_t span_lo = 0; span_lo < nbytes; span_lo += MCX_FRONT_SPAN - MCX_FRONT_SPAN % MCX_STAGE_POS) {
  const uint64_t span_hi = span_lo + (MCX_FRONT_SPAN - MCX_FRONT_SPAN % MCX_STAGE_POS) < nbytes ? span_lo + (MCX_FRONT_SPAN - MCX_FRONT_SPAN % MCX_STAGE_POS) : nbytes;
  { int rg = front_guard(g, span_hi - span_lo); if(rg) return rg; } // may queue a flush on the primary stream: the previous span has been joined into it
  CU(cudaEventRecord(g->ev_fork, primary(g)));
  bool used[MCX_NSTAGE] = {false, false, false};
  for(uint64_t pos = span_lo; pos < span_hi; pos += MCX_STAGE_POS) {
    uint64_t pend = pos + MCX_STAGE_POS < span_hi ? pos + MCX_STAGE_POS : span_hi;
    uint64_t b0 = pos ? pos - MCX_LB : 0;
    uint64_t b1 = pend + MCX_TAIL < nbytes ? pend + MCX_TAIL : nbytes;
    int s = g->next; g->next = (g->next + 1) % MCX_NSTAGE;
    cudaStream_t st = g->streams[s];
    CU(cudaEventSynchronize(g->events[s])); // previous user of this slot's staging buffers is done
    if(!used[s]) { CU(cudaStreamWaitEvent(st, g->ev_fork, 0)); used[s] = true; }
    const uint8_t *src = hseq + b0;
    if(!pinned) { memcpy(g->h_stage[s], src, b1 - b0); src = g->h_stage[s]; }
    CU(cudaMemcpyAsync(g->d_stage[s], src, b1 - b0, cudaMemcpyHostToDevice, st));
    McxBuildParams p = make_params(g, b, g->d_stage[s], b1 - b0, pos - b0, pend - b0);
    CU(launch_build(g, b, p, st));
    CU(cudaEventRecord(g->events[s], st));
  }
  for(int s = 0; s < MCX_NSTAGE; s++) if(used[s]) CU(cudaStreamWaitEvent(primary(g), g->events[s], 0));
  }
  g->pend_positions += nbytes;
  return MCX_OK;
}

// Side paths (quality batches, OFFSETS batches) stage through device scratch.  Two scratch buffers alternate: the one a
// batch takes is free once the batch queued two calls earlier has run (its event), so the host queues batch b+1 while
// batch b's copies and kernels run -- no cudaStreamSynchronize per batch (round 1 had one: the production pipeline's
// --fq-cutoff builds ran without any overlap).
static int tmp_rotate(mcx_graph *g)
{
  g->d_tmpv[g->tmp_cur] = g->d_tmp; g->d_tmp_bytesv[g->tmp_cur] = g->d_tmp_bytes;
  g->tmp_cur ^= 1;
  if(!g->ev_tmp[g->tmp_cur]) CU(cudaEventCreateWithFlags(&g->ev_tmp[g->tmp_cur], cudaEventDisableTiming));
  else CU(cudaEventSynchronize(g->ev_tmp[g->tmp_cur]));
  g->d_tmp = g->d_tmpv[g->tmp_cur]; g->d_tmp_bytes = g->d_tmp_bytesv[g->tmp_cur];
  return MCX_OK;
}
static int tmp_release(mcx_graph *g, cudaStream_t st)
{
  CU(cudaEventRecord(g->ev_tmp[g->tmp_cur], st));
  return MCX_OK;
}
static int ensure_tmp(mcx_graph *g, size_t bytes)
{
  if(g->d_tmp_bytes >= bytes) return MCX_OK;
  int r = sync_all(g); if(r) return r;
  if(g->d_tmp) cudaFree(g->d_tmp);
  g->d_tmp = NULL; g->d_tmp_bytes = 0;
  CU(cudaMalloc(&g->d_tmp, bytes));
  g->d_tmp_bytes = bytes;
  return MCX_OK;
}

// Quality cut-off batches (reference: qual_cutoff > 0 in seq_contig_start2/end2).  Contig
// membership then needs a carry across chunks, so a launch must start at a read boundary:
// the whole batch is brought to the device (seq + qual, LINES layout) and run as ONE launch
// pair (summary pass + insert pass).  Host batches are copied synchronously (no piece overlap
// on this path yet).
static int add_reads_qual(mcx_graph *g, const mcx_read_batch *b)
{
  if(b->nbytes == 0) return MCX_OK;
  cudaStream_t st = primary(g);
  const bool offsets = b->layout == MCX_LAYOUT_OFFSETS;
  if(offsets && !b->offsets) return MCX_ERR_BAD_ARG;
  if(!offsets && b->layout != MCX_LAYOUT_LINES) return MCX_ERR_BAD_ARG;
  const uint64_t lines_bytes = b->nbytes + (offsets ? b->nreads : 0);
  const size_t A = 256;
  size_t seq_off = 0, qual_off = (lines_bytes + 16 + A - 1) / A * A, sum_off = qual_off + (lines_bytes + 16 + A - 1) / A * A;
  size_t sum_bytes = lines_bytes / MCX_T + 4;
  size_t raw_seq_off = (sum_off + sum_bytes + A - 1) / A * A, raw_qual_off = raw_seq_off + (b->nbytes + A - 1) / A * A;
  size_t off_off = raw_qual_off + (b->nbytes + A - 1) / A * A;
  size_t need = off_off + (offsets ? (size_t)(b->nreads + 1) * 8 : 0) + A;
  bool direct = !offsets && b->mem == MCX_MEM_DEVICE && (((uintptr_t)b->seq | (uintptr_t)b->qual) & 15u) == 0;
  int r = tmp_rotate(g); if(r) return r;
  r = ensure_tmp(g, direct ? sum_off + sum_bytes + A : need); if(r) return r;
  const uint8_t *dseq, *dqual;
  if(direct) { dseq = (const uint8_t *)b->seq; dqual = (const uint8_t *)b->qual; }
  else if(!offsets) {
    cudaMemcpyKind kind = b->mem == MCX_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    CU(cudaMemcpyAsync(g->d_tmp + seq_off, b->seq, b->nbytes, kind, st));
    CU(cudaMemcpyAsync(g->d_tmp + qual_off, b->qual, b->nbytes, kind, st));
    dseq = g->d_tmp + seq_off; dqual = g->d_tmp + qual_off;
  } else {
    cudaMemcpyKind kind = b->mem == MCX_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    const uint8_t *rs = (const uint8_t *)b->seq, *rq = (const uint8_t *)b->qual;
    if(b->mem == MCX_MEM_HOST) {
      CU(cudaMemcpyAsync(g->d_tmp + raw_seq_off, b->seq, b->nbytes, kind, st));
      CU(cudaMemcpyAsync(g->d_tmp + raw_qual_off, b->qual, b->nbytes, kind, st));
      rs = g->d_tmp + raw_seq_off; rq = g->d_tmp + raw_qual_off;
    }
    CU(cudaMemcpyAsync(g->d_tmp + off_off, b->offsets, (size_t)(b->nreads + 1) * 8, kind, st));
    CU(mcx_launch_repack_lines(rs, (const uint64_t *)(g->d_tmp + off_off), b->nreads, g->d_tmp + seq_off, st));
    CU(mcx_launch_repack_lines(rq, (const uint64_t *)(g->d_tmp + off_off), b->nreads, g->d_tmp + qual_off, st));
    dseq = g->d_tmp + seq_off; dqual = g->d_tmp + qual_off;
  }
  g->occ_bound += lines_bytes;
  McxBuildParams p = make_params(g, b, dseq, lines_bytes, 0, lines_bytes);
  p.qual = dqual; p.qcut = b->fq_cutoff; p.summary = g->d_tmp + sum_off;
  if(b->must_exist) { CU(mcx_launch_contig_summary(p, st)); CU(mcx_launch_build_lookup(p, g->table, st)); }
  else CU(mcx_launch_build_fused_qual(p, g->table, st));
  g->pend_positions += lines_bytes;
  return tmp_release(g, st); // (pageable host sources were staged by cudaMemcpyAsync before it returned; pinned ones stay the caller's until sync)
}

extern "C" int mcx_graph_add_reads(mcx_graph *g, const mcx_read_batch *b)
{
  if(!g || !b || b->colour >= g->ncols) return MCX_ERR_BAD_ARG;
  if(b->nbytes && !b->seq) return MCX_ERR_BAD_ARG;
  if(b->hp_cutoff == 1 || b->hp_cutoff > g->k) {
    snprintf(g_err, sizeof(g_err), "hp_cutoff must be 0 or in [2, k]"); return MCX_ERR_UNSUPPORTED;
  }
  if(b->fq_cutoff >= 127) { snprintf(g_err, sizeof(g_err), "fq_cutoff (incl. offset) must be < 127"); return MCX_ERR_UNSUPPORTED; }
  CU(cudaSetDevice(g->device));
  if(b->must_exist && g->table.front_set_bits) {
    // everything counted so far must be in the big table before k-mers are looked up there
    CU(flush_front(g, primary(g)));
    g->front_pending = 0;
  }
  if(g->exp_valid) { mcx_export_free(&g->exp); g->exp_valid = false; }
  { int r = front_colour(g, b->colour); if(r) return r; }
  if((b->fq_cutoff && b->qual) || b->layout != MCX_LAYOUT_LINES) { int r = front_guard(g, b->nbytes + b->nreads); if(r) return r; }
  if(b->fq_cutoff && b->qual) return add_reads_qual(g, b);

  if(b->layout == MCX_LAYOUT_LINES) {
    if(b->nbytes == 0) return MCX_OK;
    if(b->mem == MCX_MEM_DEVICE) return add_lines_device(g, b, (const uint8_t *)b->seq, b->nbytes);
    if(b->mem == MCX_MEM_HOST) return add_lines_host(g, b, (const uint8_t *)b->seq, b->nbytes);
    return MCX_ERR_BAD_ARG;
  }
  if(b->layout != MCX_LAYOUT_OFFSETS || (!b->offsets && b->nreads)) return MCX_ERR_BAD_ARG;
  if(b->nreads == 0) return MCX_OK;

  // OFFSETS -> LINES on the device: read r shifts right by r bytes and gains a terminator
  const uint64_t lines_bytes = b->nbytes + b->nreads;
  const size_t off_bytes = (size_t)(b->nreads + 1) * sizeof(uint64_t);
  const size_t lines_off = 0, offs_off = (lines_bytes + 255) & ~(size_t)255;
  size_t raw_off = (offs_off + off_bytes + 255) & ~(size_t)255;
  size_t need = raw_off + (b->mem == MCX_MEM_HOST ? ((b->nbytes + 255) & ~(size_t)255) : 0);
  int r = tmp_rotate(g); if(r) return r;
  r = ensure_tmp(g, need + 256); if(r) return r;
  cudaStream_t st = primary(g);
  uint64_t *d_off = (uint64_t *)(g->d_tmp + offs_off);
  const uint8_t *d_raw = (const uint8_t *)b->seq;
  if(b->mem == MCX_MEM_HOST) {
    if(b->offsets[0] != 0 || b->offsets[b->nreads] != b->nbytes) return MCX_ERR_BAD_ARG;
    CU(cudaMemcpyAsync(g->d_tmp + raw_off, b->seq, b->nbytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_off, b->offsets, off_bytes, cudaMemcpyHostToDevice, st));
    d_raw = g->d_tmp + raw_off;
  } else if(b->mem == MCX_MEM_DEVICE) {
    CU(cudaMemcpyAsync(d_off, b->offsets, off_bytes, cudaMemcpyDeviceToDevice, st));
  } else return MCX_ERR_BAD_ARG;
  CU(mcx_launch_repack_lines(d_raw, d_off, b->nreads, g->d_tmp + lines_off, st));
  g->occ_bound += lines_bytes;
  McxBuildParams p = make_params(g, b, g->d_tmp + lines_off, lines_bytes, 0, lines_bytes);
  CU(launch_build(g, b, p, st));
  g->pend_positions += lines_bytes;
  return tmp_release(g, st);
}

// ---- build --remove-pcr ---------------------------------------------------------------------
extern "C" int mcx_graph_pcr_reset(mcx_graph *g)
{
  if(!g || !g->d_first) return MCX_ERR_BAD_ARG;
  CU(cudaSetDevice(g->device));
  CU(cudaMemsetAsync(g->d_first, 0xFF, (size_t)g->table.nslots * 8u, primary(g)));
  g->pcr_ord = 0;
  return MCX_OK;
}

extern "C" int mcx_graph_add_reads_pcr(mcx_graph *g, const mcx_read_batch *b, const uint64_t *read_off, const uint8_t *mate,
                                       uint64_t nreads)
{
  if(!g || !b || b->colour >= g->ncols || b->layout != MCX_LAYOUT_LINES) return MCX_ERR_BAD_ARG;
  if(!g->d_first) { snprintf(g_err, sizeof(g_err), "graph was not created with MCX_GRAPH_READSTRT"); return MCX_ERR_BAD_ARG; }
  if(b->must_exist) { snprintf(g_err, sizeof(g_err), "remove-pcr and must-exist exclude each other (build_graph.c:197)"); return MCX_ERR_UNSUPPORTED; }
  if(nreads == 0 || b->nbytes == 0) return MCX_OK;
  if(!b->seq || !read_off || !mate || (b->mem != MCX_MEM_HOST && b->mem != MCX_MEM_DEVICE)) return MCX_ERR_BAD_ARG;
  if(b->hp_cutoff == 1 || b->hp_cutoff > g->k) { snprintf(g_err, sizeof(g_err), "hp_cutoff must be 0 or in [2, k]"); return MCX_ERR_UNSUPPORTED; }
  if(b->fq_cutoff >= 127) { snprintf(g_err, sizeof(g_err), "fq_cutoff (incl. offset) must be < 127"); return MCX_ERR_UNSUPPORTED; }
  if((uint64_t)g->pcr_ord + nreads >= 0xFFFFFFFFull) { snprintf(g_err, sizeof(g_err), "remove-pcr: more than 4e9 reads in one colour"); return MCX_ERR_UNSUPPORTED; }
  CU(cudaSetDevice(g->device));
  cudaStream_t st = primary(g);
  const bool useq = b->fq_cutoff && b->qual;
  const bool host = b->mem == MCX_MEM_HOST;
  const size_t A = 256, sb = ((size_t)b->nbytes + 16 + A - 1) / A * A;
  const size_t seq_off = 0, qual_off = host ? sb : 0, off_off = qual_off + (host && useq ? sb : 0);
  const size_t mate_off = off_off + (host ? ((size_t)(nreads + 1) * 8 + A - 1) / A * A : 0);
  const size_t node_off = mate_off + (host ? ((size_t)nreads + A - 1) / A * A : 0);
  const size_t need = node_off + (size_t)nreads * 8 + A;
  if(g->d_pcr_bytes < need) {
    int r = sync_all(g); if(r) return r;
    if(g->d_pcr) cudaFree(g->d_pcr);
    g->d_pcr = NULL; g->d_pcr_bytes = 0;
    CU(cudaMalloc(&g->d_pcr, need + need / 4));
    g->d_pcr_bytes = need + need / 4;
  }
  uint8_t *dseq = (uint8_t *)b->seq, *dqual = useq ? (uint8_t *)b->qual : nullptr;
  const uint64_t *doff = read_off; const uint8_t *dmate = mate;
  if(host) {
    if(read_off[0] != 0 || read_off[nreads] != b->nbytes) return MCX_ERR_BAD_ARG;
    dseq = g->d_pcr + seq_off;
    CU(cudaMemcpyAsync(dseq, b->seq, b->nbytes, cudaMemcpyHostToDevice, st));
    if(useq) { dqual = g->d_pcr + qual_off; CU(cudaMemcpyAsync(dqual, b->qual, b->nbytes, cudaMemcpyHostToDevice, st)); }
    CU(cudaMemcpyAsync(g->d_pcr + off_off, read_off, (size_t)(nreads + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(g->d_pcr + mate_off, mate, (size_t)nreads, cudaMemcpyHostToDevice, st));
    doff = (const uint64_t *)(g->d_pcr + off_off); dmate = g->d_pcr + mate_off;
  }
  CU(mcx_launch_pcr_filter(dseq, dqual, doff, dmate, nreads, g->k, useq ? b->fq_cutoff : 0u, b->hp_cutoff, g->table, g->d_first,
                           (uint64_t *)(g->d_pcr + node_off), g->pcr_ord, g->d_counters, st));
  g->pcr_ord += (uint32_t)nreads;
  mcx_read_batch db = *b;
  db.mem = MCX_MEM_DEVICE; db.seq = (const char *)dseq; db.qual = (const char *)dqual;
  int r = mcx_graph_add_reads(g, &db);
  if(r) return r;
  CU(cudaStreamSynchronize(st)); // the device copy (and pageable host sources) are reused by the next batch
  return MCX_OK;
}

extern "C" int mcx_graph_add_str(mcx_graph *g, uint32_t colour, const char *seq, size_t len)
{
  if(!g || !seq) return MCX_ERR_BAD_ARG;
  char *buf = (char *)malloc(len + 1);
  if(!buf) return MCX_ERR_NOMEM;
  memcpy(buf, seq, len); buf[len] = '\n';
  mcx_read_batch b; memset(&b, 0, sizeof(b));
  b.seq = buf; b.nbytes = len + 1; b.layout = MCX_LAYOUT_LINES; b.mem = MCX_MEM_HOST; b.colour = colour;
  int r = mcx_graph_add_reads(g, &b);
  if(r == MCX_OK) r = sync_all(g);
  free(buf);
  return r;
}

extern "C" int mcx_graph_sync(mcx_graph *g, mcx_load_stats *stats)
{
  if(!g) return MCX_ERR_BAD_ARG;
  int r = sync_all(g); if(r) return r;
  if(g->sharded) { snprintf(g_err, sizeof(g_err), "mcx_graph_flush_sharded must run before sync"); return MCX_ERR_BAD_ARG; }
  CU(flush_front(g, primary(g)));
  g->front_pending = 0;
  CU(cudaStreamSynchronize(primary(g)));
  unsigned long long c[MCX_NCOUNTERS_ALL];
  CU(cudaMemcpy(c, g->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
  CU(cudaMemset(g->d_counters, 0, sizeof(c)));
  g->nkmers += c[MCX_CNT_NOVEL];
  if(stats) {
    memset(stats, 0, sizeof(*stats));
    stats->num_kmers_loaded = c[MCX_CNT_KMERS];
    stats->num_kmers_novel = c[MCX_CNT_NOVEL];
    stats->contigs_parsed = c[MCX_CNT_CONTIGS];
    // every contig of n windows spans n + k - 1 bases (build_graph.c:173-176)
    // (must-exist builds: windows whose k-mer is not in the graph still belong to their contig, build_graph.c:173-181)
    stats->total_bases_loaded = c[MCX_CNT_KMERS] + c[MCX_CNT_NOTFOUND] + (uint64_t)(g->k - 1u) * c[MCX_CNT_CONTIGS];
    stats->num_se_reads = c[MCX_CNT_READS];
    stats->total_bases_read = g->pend_positions - c[MCX_CNT_READS];
    stats->num_good_reads = UINT64_MAX;
    stats->num_bad_reads = UINT64_MAX;
    stats->num_dup_se_reads = c[MCX_CNT_DUP_SE];
    stats->num_dup_pe_pairs = c[MCX_CNT_DUP_PE];
  }
  g->pend_positions = 0;
  if(c[MCX_CNT_FULL]) { snprintf(g_err, sizeof(g_err), "Hash table is full"); return MCX_ERR_TABLE_FULL; }
  return MCX_OK;
}

extern "C" int mcx_graph_flush(mcx_graph *g)
{
  if(!g) return MCX_ERR_BAD_ARG;
  if(g->sharded) { snprintf(g_err, sizeof(g_err), "mcx_graph_flush_sharded must be used on a sharded graph"); return MCX_ERR_BAD_ARG; }
  CU(cudaSetDevice(g->device));
  CU(flush_front(g, primary(g)));
  g->front_pending = 0;
  return MCX_OK;
}

extern "C" int mcx_graph_stats(mcx_graph *g, uint64_t *nkmers, uint64_t *capacity)
{
  if(!g) return MCX_ERR_BAD_ARG;
  if(nkmers) *nkmers = g->nkmers;
  if(capacity) *capacity = g->capacity;
  return MCX_OK;
}

extern "C" int mcx_graph_export_begin(mcx_graph *g, int sorted, uint64_t *nrecords, uint32_t *record_bytes)
{
  if(!g) return MCX_ERR_BAD_ARG;
  int r = sync_all(g); if(r) return r;
  if(g->sharded) { snprintf(g_err, sizeof(g_err), "mcx_graph_flush_sharded must run before export"); return MCX_ERR_BAD_ARG; }
  CU(flush_front(g, primary(g)));
  g->front_pending = 0;
  if(g->exp_valid) { mcx_export_free(&g->exp); g->exp_valid = false; }
  cudaError_t e = mcx_export_build(g->table, g->k, sorted != 0, &g->exp, primary(g));
  if(e != cudaSuccess) return fail_cuda(e, "export");
  g->exp_valid = true;
  if(nrecords) *nrecords = g->exp.nrec;
  if(record_bytes) *record_bytes = g->exp.rec_bytes;
  return MCX_OK;
}

extern "C" int mcx_graph_export_read(mcx_graph *g, uint64_t first, uint64_t n, void *host_dst)
{
  if(!g || !g->exp_valid || first + n > g->exp.nrec || (!host_dst && n)) return MCX_ERR_BAD_ARG;
  if(n == 0) return MCX_OK;
  CU(cudaSetDevice(g->device));
  CU(cudaMemcpy(host_dst, g->exp.records + first * g->exp.rec_bytes, n * g->exp.rec_bytes, cudaMemcpyDeviceToHost));
  return MCX_OK;
}

extern "C" int mcx_graph_export_end(mcx_graph *g)
{
  if(!g) return MCX_ERR_BAD_ARG;
  mcx_export_free(&g->exp); g->exp_valid = false;
  return MCX_OK;
}

// bins described either by one contiguous allocation (bin d at d * cap) or by one pointer per destination
static int fill_bins(McxTupleBins *bins, uint32_t W, uint32_t nparts, uint32_t my_part, uint64_t cap,
                     uint64_t *keys_out, uint32_t *meta_out, uint64_t *const *keys_dst, uint32_t *const *meta_dst,
                     uint64_t *counts_out)
{
  if(nparts > MCX_MAX_PARTS) { snprintf(g_err, sizeof(g_err), "at most %d shards", MCX_MAX_PARTS); return MCX_ERR_UNSUPPORTED; }
  for(uint32_t d = 0; d < MCX_MAX_PARTS; d++) { bins->keys[d] = nullptr; bins->meta[d] = nullptr; }
  for(uint32_t d = 0; d < nparts; d++) {
    bins->keys[d] = keys_dst ? keys_dst[d] : keys_out + (uint64_t)d * cap * W;
    bins->meta[d] = meta_dst ? meta_dst[d] : meta_out + (uint64_t)d * cap;
    if(d != my_part && (!bins->keys[d] || !bins->meta[d])) return MCX_ERR_BAD_ARG;
  }
  bins->cursor = (unsigned long long *)counts_out; bins->cap = cap; bins->nparts = nparts; bins->my_part = my_part;
  return MCX_OK;
}

extern "C" int mcx_kmer_tuples(mcx_graph *g, const mcx_read_batch *b, uint32_t nparts, uint64_t cap_per_part,
                               uint64_t *keys_out, uint32_t *masks_out, uint64_t *counts_out)
{
  if(!g || !b || !nparts || !cap_per_part || !keys_out || !masks_out || !counts_out) return MCX_ERR_BAD_ARG;
  if(b->layout != MCX_LAYOUT_LINES || b->mem != MCX_MEM_DEVICE || ((uintptr_t)b->seq & 15u)) return MCX_ERR_BAD_ARG;
  if(b->hp_cutoff == 1 || b->hp_cutoff > g->k || (b->fq_cutoff && b->qual) || b->must_exist) return MCX_ERR_UNSUPPORTED;
  CU(cudaSetDevice(g->device));
  cudaStream_t st = primary(g);
  CU(cudaMemsetAsync(counts_out, 0, nparts * sizeof(uint64_t), st));
  McxTupleBins bins;
  { int r = fill_bins(&bins, g->W, nparts, UINT32_MAX, cap_per_part, keys_out, masks_out, NULL, NULL, counts_out); if(r) return r; }
  bins.my_part = 0;
  McxBuildParams p = make_params(g, b, (const uint8_t *)b->seq, b->nbytes, 0, b->nbytes);
  CU(mcx_launch_kmer_tuples(p, bins, st));
  g->pend_positions += b->nbytes;
  return MCX_OK;
}

// sharded build, per batch: local front table + local big table for owned keys + tuples for the rest
// [r_begin, r_end): the positions of the (device) buffer this launch owns; r_end == 0 means the whole buffer
static int add_reads_sharded(mcx_graph *g, const mcx_read_batch *b, uint32_t nparts, uint32_t my_part, uint64_t cap_per_part,
                             uint64_t *keys_out, uint32_t *meta_out, uint64_t *const *keys_dst, uint32_t *const *meta_dst,
                             uint64_t *counts_out, uint64_t r_begin = 0, uint64_t r_end = 0)
{
  if(!g || !b || nparts < 2 || my_part >= nparts || !cap_per_part || !counts_out) return MCX_ERR_BAD_ARG;
  if(!(keys_out && meta_out) && !(keys_dst && meta_dst)) return MCX_ERR_BAD_ARG;
  if(b->layout != MCX_LAYOUT_LINES || b->mem != MCX_MEM_DEVICE || ((uintptr_t)b->seq & 15u)) return MCX_ERR_BAD_ARG;
  if(b->hp_cutoff == 1 || b->hp_cutoff > g->k || b->must_exist) return MCX_ERR_UNSUPPORTED;
  const bool useq = b->fq_cutoff && b->qual;
  if(useq && (b->fq_cutoff >= 127 || ((uintptr_t)b->qual & 15u))) return MCX_ERR_BAD_ARG;
  if(useq && (r_begin != 0 || (r_end != 0 && r_end != b->nbytes))) {
    snprintf(g_err, sizeof(g_err), "a quality cut-off launch must cover whole reads"); return MCX_ERR_BAD_ARG;
  }
  if(b->colour >= g->ncols) return MCX_ERR_BAD_ARG;
  CU(cudaSetDevice(g->device));
  if(g->exp_valid) { mcx_export_free(&g->exp); g->exp_valid = false; }
  cudaStream_t st = primary(g);
  McxTupleBins bins;
  { int r = fill_bins(&bins, g->W, nparts, my_part, cap_per_part, keys_out, meta_out, keys_dst, meta_dst, counts_out); if(r) return r; }
  { int r = front_colour(g, b->colour); if(r) return r; }
  g->sharded = true;
  if(r_end == 0) r_end = b->nbytes;
  if(r_begin > r_end || r_end > b->nbytes) return MCX_ERR_BAD_ARG;
  { int r = front_guard(g, r_end - r_begin); if(r) return r; }
  CU(cudaMemsetAsync(counts_out, 0, nparts * sizeof(uint64_t), st));
  g->occ_bound += r_end - r_begin;
  McxBuildParams p = make_params(g, b, (const uint8_t *)b->seq, b->nbytes, r_begin, r_end);
  if(useq) { // the carry summaries of the quality pass (one byte per chunk) live in the rotating scratch
    int r = tmp_rotate(g); if(r) return r;
    r = ensure_tmp(g, b->nbytes / MCX_T + 260); if(r) return r;
    p.qual = (const uint8_t *)b->qual; p.qcut = b->fq_cutoff; p.summary = g->d_tmp;
  }
#ifdef MCX_EXPERIMENTS
  if(g->warp_kernel && mcx_warp_kernel_supports(p)) {
    for(uint32_t c = 0; c < (1u << g->ncls_log2); c++) {
      McxBuildParams pc = p; pc.cls = c; pc.ncls_log2 = g->table.front_set_bits ? g->ncls_log2 : 0u;
      CU(mcx_launch_build_warp_sharded(pc, class_table(g, c), bins, st));
      if(!pc.ncls_log2) break;
    }
  } else
#endif
  CU(mcx_launch_build_sharded(p, g->table, bins, st));
  if(useq) { int r = tmp_release(g, st); if(r) return r; }
  g->pend_positions += r_end - r_begin;
  g->sharded = true;
  return MCX_OK;
}

extern "C" int mcx_graph_add_reads_sharded(mcx_graph *g, const mcx_read_batch *b, uint32_t nparts, uint32_t my_part,
                                           uint64_t cap_per_part, uint64_t *keys_out, uint32_t *meta_out, uint64_t *counts_out)
{
  return add_reads_sharded(g, b, nparts, my_part, cap_per_part, keys_out, meta_out, NULL, NULL, counts_out);
}

extern "C" int mcx_graph_add_reads_routed(mcx_graph *g, const mcx_read_batch *b, uint32_t nparts, uint32_t my_part,
                                          uint64_t cap_per_part, uint64_t *const *keys_dst, uint32_t *const *meta_dst,
                                          uint64_t *counts_out)
{
  return add_reads_sharded(g, b, nparts, my_part, cap_per_part, NULL, NULL, keys_dst, meta_dst, counts_out);
}

// sharded build, end of a step: empty the front table -- owned records into the local big table,
// the others (aggregated: one tuple per k-mer, not per occurrence) into the bins
static int flush_sharded(mcx_graph *g, uint32_t nparts, uint32_t my_part, uint64_t cap_per_part, uint64_t *keys_out,
                         uint32_t *meta_out, uint64_t *const *keys_dst, uint32_t *const *meta_dst, uint64_t *counts_out)
{
  if(!g || nparts < 2 || my_part >= nparts || !cap_per_part || !counts_out) return MCX_ERR_BAD_ARG;
  if(!(keys_out && meta_out) && !(keys_dst && meta_dst)) return MCX_ERR_BAD_ARG;
  CU(cudaSetDevice(g->device));
  cudaStream_t st = primary(g);
  McxTupleBins bins;
  { int r = fill_bins(&bins, g->W, nparts, my_part, cap_per_part, keys_out, meta_out, keys_dst, meta_dst, counts_out); if(r) return r; }
  CU(cudaMemsetAsync(counts_out, 0, nparts * sizeof(uint64_t), st));
  CU(flush_front(g, st, &bins));
  g->sharded = false; g->front_pending = 0;
  return MCX_OK;
}

extern "C" int mcx_graph_flush_sharded(mcx_graph *g, uint32_t nparts, uint32_t my_part, uint64_t cap_per_part,
                                       uint64_t *keys_out, uint32_t *meta_out, uint64_t *counts_out)
{
  return flush_sharded(g, nparts, my_part, cap_per_part, keys_out, meta_out, NULL, NULL, counts_out);
}

extern "C" int mcx_graph_flush_routed(mcx_graph *g, uint32_t nparts, uint32_t my_part, uint64_t cap_per_part,
                                      uint64_t *const *keys_dst, uint32_t *const *meta_dst, uint64_t *counts_out)
{
  return flush_sharded(g, nparts, my_part, cap_per_part, NULL, NULL, keys_dst, meta_dst, counts_out);
}

extern "C" int mcx_graph_insert_tuples(mcx_graph *g, const uint64_t *keys, const uint32_t *masks, uint64_t n, uint32_t colour)
{
  if(!g || colour >= g->ncols || (n && (!keys || !masks))) return MCX_ERR_BAD_ARG;
  CU(cudaSetDevice(g->device));
  if(g->exp_valid) { mcx_export_free(&g->exp); g->exp_valid = false; }
  g->occ_bound += n;
  CU(mcx_launch_insert_tuples(keys, masks, n, NULL, g->k, g->table, colour, g->occ_bound >= 0xF0000000ull, g->d_counters,
                              primary(g)));
  return MCX_OK;
}

// same, but the number of tuples is a device word (written by the sender, exchanged on the stream):
// nothing on this path makes the host wait for the GPU
extern "C" int mcx_graph_insert_tuples_on(mcx_graph *g, void *cuda_stream, const uint64_t *keys, const uint32_t *masks,
                                          const uint64_t *n_dev, uint64_t n_max, uint32_t colour)
{
  if(!g || colour >= g->ncols || !n_dev || (n_max && (!keys || !masks))) return MCX_ERR_BAD_ARG;
  CU(cudaSetDevice(g->device));
  if(g->exp_valid) { mcx_export_free(&g->exp); g->exp_valid = false; }
  g->occ_bound += n_max;
  CU(mcx_launch_insert_tuples(keys, masks, n_max, n_dev, g->k, g->table, colour, g->occ_bound >= 0xF0000000ull, g->d_counters,
                              cuda_stream ? (cudaStream_t)cuda_stream : primary(g)));
  return MCX_OK;
}

extern "C" int mcx_graph_insert_tuples_n(mcx_graph *g, const uint64_t *keys, const uint32_t *masks, const uint64_t *n_dev,
                                         uint64_t n_max, uint32_t colour)
{
  if(!g || colour >= g->ncols || !n_dev || (n_max && (!keys || !masks))) return MCX_ERR_BAD_ARG;
  CU(cudaSetDevice(g->device));
  if(g->exp_valid) { mcx_export_free(&g->exp); g->exp_valid = false; }
  g->occ_bound += n_max; // every tuple may carry an aggregated count: the saturation guard switches on early, never late
  CU(mcx_launch_insert_tuples(keys, masks, n_max, n_dev, g->k, g->table, colour, g->occ_bound >= 0xF0000000ull, g->d_counters,
                              primary(g)));
  return MCX_OK;
}

// replaces graph_load() for records the caller has read from a .ctx file (any number of file colours;
// the (from, into) pairs are the reference's FileFilter).  Synchronous: returns after the records
// are merged, with the counts the reference keeps in GraphLoadingStats.
extern "C" int mcx_graph_load_records(mcx_graph *g, const void *records, uint64_t nrecords, uint32_t file_ncols, uint32_t mem,
                                      const uint32_t *from_col, const uint32_t *into_col, uint32_t nmap, uint32_t flags,
                                      uint64_t *nkmers_loaded, uint64_t *nkmers_novel)
{
  if(!g || !file_ncols || (nrecords && !records) || (nmap && (!from_col || !into_col))) return MCX_ERR_BAD_ARG;
  for(uint32_t m = 0; m < nmap; m++) if(from_col[m] >= file_ncols || into_col[m] >= g->ncols) return MCX_ERR_BAD_ARG;
  if((flags & (MCX_LOAD_INTO_ISEC | MCX_LOAD_MASK_ISEC)) && !g->d_isec) {
    snprintf(g_err, sizeof(g_err), "graph was not created with MCX_GRAPH_INTERSECT"); return MCX_ERR_BAD_ARG;
  }
  CU(cudaSetDevice(g->device));
  if(g->exp_valid) { mcx_export_free(&g->exp); g->exp_valid = false; }
  if(nkmers_loaded) *nkmers_loaded = 0;
  if(nkmers_novel) *nkmers_novel = 0;
  if(nrecords == 0 || nmap == 0) return MCX_OK;
  cudaStream_t st = primary(g);
  // must-exist loads look keys up in the big table: k-mers that so far live only in the front table would count as absent
  if((flags & MCX_LOAD_MUST_EXIST) && g->table.front_set_bits && g->front_pending && !g->sharded) {
    CU(flush_front(g, st));
    g->front_pending = 0;
  }
  const size_t rec_bytes = 8u * g->W + 5u * (size_t)file_ncols, bytes = rec_bytes * nrecords;
  const size_t map_off = (bytes + 255) & ~(size_t)255;
  const uint8_t *drecs = (const uint8_t *)records;
  int r = ensure_tmp(g, map_off + 8u * (size_t)nmap + 256 + (mem == MCX_MEM_HOST ? 0 : 0)); if(r) return r;
  if(mem == MCX_MEM_HOST) { CU(cudaMemcpyAsync(g->d_tmp, records, bytes, cudaMemcpyHostToDevice, st)); drecs = g->d_tmp; }
  else if(mem != MCX_MEM_DEVICE) return MCX_ERR_BAD_ARG;
  uint32_t *d_from = (uint32_t *)(g->d_tmp + map_off), *d_into = d_from + nmap;
  CU(cudaMemcpyAsync(d_from, from_col, 4u * (size_t)nmap, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_into, into_col, 4u * (size_t)nmap, cudaMemcpyHostToDevice, st));
  // counts of THIS call: read the running counters before and after
  unsigned long long c0[MCX_NCOUNTERS_ALL], c1[MCX_NCOUNTERS_ALL];
  CU(cudaMemcpyAsync(c0, g->d_counters, sizeof(c0), cudaMemcpyDeviceToHost, st));
  g->occ_bound = 0xF0000000ull; // file coverages can be anything: saturation-aware adds from here on
  CU(mcx_launch_load_records(drecs, nrecords, file_ncols, d_from, d_into, nmap, flags, g->k, g->table, g->d_isec, g->d_counters, st));
  CU(cudaMemcpyAsync(c1, g->d_counters, sizeof(c1), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if(nkmers_loaded) *nkmers_loaded = c1[MCX_CNT_RECS_LOADED] - c0[MCX_CNT_RECS_LOADED];
  if(nkmers_novel) *nkmers_novel = c1[MCX_CNT_NOVEL] - c0[MCX_CNT_NOVEL];
  if(c1[MCX_CNT_FULL]) { snprintf(g_err, sizeof(g_err), "Hash table is full"); return MCX_ERR_TABLE_FULL; }
  return MCX_OK;
}

extern "C" int mcx_graph_finish_intersect(mcx_graph *g, uint64_t *nkmers)
{
  if(!g) return MCX_ERR_BAD_ARG;
  if(!g->d_isec) { snprintf(g_err, sizeof(g_err), "graph was not created with MCX_GRAPH_INTERSECT"); return MCX_ERR_BAD_ARG; }
  CU(cudaSetDevice(g->device));
  int r = sync_all(g); if(r) return r;
  if(g->exp_valid) { mcx_export_free(&g->exp); g->exp_valid = false; }
  cudaStream_t st = primary(g);
  unsigned long long *d_n = g->d_counters + MCX_CNT_INSERTED; // scratch counter (unused on this path)
  CU(cudaMemsetAsync(d_n, 0, sizeof(*d_n), st));
  CU(mcx_launch_finish_intersect(g->table, g->W, g->d_isec, d_n, st));
  unsigned long long n = 0;
  CU(cudaMemcpyAsync(&n, d_n, sizeof(n), cudaMemcpyDeviceToHost, st));
  CU(cudaMemsetAsync(d_n, 0, sizeof(*d_n), st));
  CU(cudaStreamSynchronize(st));
  g->nkmers = n;
  if(nkmers) *nkmers = n;
  return MCX_OK;
}

// replaces the body of ctx_sort (src/commands/ctx_sort.c:117-155): nrecords packed .ctx records in host
// memory -> the same records in ascending key order (records_out may equal records_in)
extern "C" int mcx_sort_records(int device, uint32_t kmer_size, uint32_t ncols, const void *records_in, uint64_t nrecords,
                                void *records_out)
{
  if(kmer_size < 3 || kmer_size > 63 || !(kmer_size & 1u) || !ncols || (nrecords && (!records_in || !records_out))) return MCX_ERR_BAD_ARG;
  if(mcx_device_count() == 0) { snprintf(g_err, sizeof(g_err), "no CUDA device: libmcxgpu has no CPU fallback"); return MCX_ERR_NO_DEVICE; }
  if(nrecords == 0) return MCX_OK;
  CU(cudaSetDevice(device));
  const size_t bytes = (size_t)nrecords * (8u * ((kmer_size + 31u) / 32u) + 5u * (size_t)ncols);
  uint8_t *d_in = NULL, *d_out = NULL;
  cudaError_t e = cudaMalloc(&d_in, bytes + 16);
  if(e == cudaSuccess) e = cudaMalloc(&d_out, bytes + 16);
  if(e == cudaSuccess) e = cudaMemcpy(d_in, records_in, bytes, cudaMemcpyHostToDevice);
  if(e == cudaSuccess) e = mcx_sort_records_device(d_in, nrecords, kmer_size, ncols, d_out, 0);
  if(e == cudaSuccess) e = cudaMemcpy(records_out, d_out, bytes, cudaMemcpyDeviceToHost);
  if(d_in) cudaFree(d_in);
  if(d_out) cudaFree(d_out);
  if(e != cudaSuccess) return fail_cuda(e, "mcx_sort_records");
  return MCX_OK;
}

// ---- device buffers that peers can map (one process per GPU: CUDA IPC over NVLink) ----------
extern "C" int mcx_device_alloc(int device, size_t bytes, void **dptr)
{
  if(!dptr) return MCX_ERR_BAD_ARG;
  if(mcx_device_count() == 0) return MCX_ERR_NO_DEVICE;
  CU(cudaSetDevice(device));
  CU(cudaMalloc(dptr, bytes ? bytes : 1));
  CU(cudaMemset(*dptr, 0, bytes ? bytes : 1));
  return MCX_OK;
}
extern "C" int mcx_device_free(int device, void *dptr)
{
  if(!dptr) return MCX_OK;
  CU(cudaSetDevice(device));
  CU(cudaFree(dptr));
  return MCX_OK;
}
extern "C" int mcx_ipc_export(const void *dptr, unsigned char handle[64])
{
  if(!dptr || !handle) return MCX_ERR_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, (void *)dptr));
  memcpy(handle, &h, 64);
  return MCX_OK;
}
extern "C" int mcx_ipc_open(int device, const unsigned char handle[64], void **dptr)
{
  if(!dptr || !handle) return MCX_ERR_BAD_ARG;
  CU(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CU(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
  return MCX_OK;
}
extern "C" int mcx_ipc_close(int device, void *dptr)
{
  if(!dptr) return MCX_OK;
  CU(cudaSetDevice(device));
  CU(cudaIpcCloseMemHandle(dptr));
  return MCX_OK;
}

extern "C" uint32_t mcx_key_owner(const uint64_t *key_words, uint32_t k, uint32_t nparts)
{
  uint32_t hb, hc;
  if(k <= 31) { McxKmer<1> key; key.b[0] = key_words[0]; hc = mcx_lookup3<1>(key, 0u, &hb); }
  else { McxKmer<2> key; key.b[0] = key_words[0]; key.b[1] = key_words[1]; hc = mcx_lookup3<2>(key, 0u, &hb); }
  return mcx_owner(hc, nparts);
}


// =====================================================================================================================
// Shard set: the sharded build driven from ONE process (the C driver's `build -D 0,1,.. --shard`).
// One table shard per device, owner(key) = top bits of Lookup3.  Host batches are cut into pieces of
// MCX_STAGE_POS positions, piece i goes to device i mod P: H2D on the device's copy stream, then the sharded kernel
// (mcx_build_sharded_kernel: local front table, local inserts for owned keys, tuples for the others stored straight
// into the owner's receive ring -- peer memory over NVLink, cudaDeviceEnablePeerAccess), then kernel C on every owner.
// Ordering is events only (the host never waits inside a batch except for a free staging slot):
//   produced[d]     d's kernel, hence its stores into the owners' rings and its counters, has completed
//   inserted[o][d]  owner o has inserted region d of its ring: d may overwrite it (and its counters) with its next piece
// Replaces what ctx_build.c:389-425 does with one shared-memory table: build_graph() + graph_writer_save_mkhdr().
#define MCX_SS_NSLOT 2
struct mcx_shardset {
  uint32_t k, W, ncols, P;
  int dev[MCX_MAX_PARTS];
  mcx_graph *g[MCX_MAX_PARTS];
  uint64_t cap;                                  // tuples per (owner, sender) ring region
  uint64_t *ring_k[MCX_MAX_PARTS]; uint32_t *ring_m[MCX_MAX_PARTS];   // on the OWNER: P regions
  uint64_t *counts[MCX_MAX_PARTS];               // on the SENDER: tuples it produced for each owner
  cudaStream_t copy[MCX_MAX_PARTS];
  uint8_t *d_stage[MCX_MAX_PARTS][MCX_SS_NSLOT], *h_stage[MCX_MAX_PARTS][MCX_SS_NSLOT];
  uint8_t *d_qstage[MCX_MAX_PARTS][MCX_SS_NSLOT], *h_qstage[MCX_MAX_PARTS][MCX_SS_NSLOT];   // quality bytes (allocated on first use)
  cudaEvent_t staged[MCX_MAX_PARTS][MCX_SS_NSLOT], stage_free[MCX_MAX_PARTS][MCX_SS_NSLOT];
  int slot[MCX_MAX_PARTS];
  cudaEvent_t produced[MCX_MAX_PARTS], inserted[MCX_MAX_PARTS][MCX_MAX_PARTS];
  bool region_busy[MCX_MAX_PARTS][MCX_MAX_PARTS];
  uint32_t next, colour; bool dirty;             // round robin; colour of the pieces since the last flush
  uint64_t pending[MCX_MAX_PARTS];               // positions since the last front-table flush, per device
  uint64_t bytes_submitted;
  // export: one sorted run per shard, merged on the host while reading
  bool exp_on, exp_sorted; uint32_t rec_bytes;
  uint64_t exp_n[MCX_MAX_PARTS], exp_at[MCX_MAX_PARTS];   // records of shard d / records already fetched
  uint8_t *exp_buf[MCX_MAX_PARTS]; uint64_t exp_have[MCX_MAX_PARTS], exp_pos[MCX_MAX_PARTS];  // host chunk: records in it / consumed
  uint32_t exp_cur;
};
#define MCX_SS_CHUNK_RECS (1u << 20)

static void ss_peer_pointers(mcx_shardset *s, uint32_t d, uint64_t **kd, uint32_t **md)
{
  for(uint32_t o = 0; o < s->P; o++) { kd[o] = s->ring_k[o] + (uint64_t)d * s->cap * s->W; md[o] = s->ring_m[o] + (uint64_t)d * s->cap; }
}
// after d has produced (kernel or flush): every owner inserts region d of its ring
static int ss_insert_all(mcx_shardset *s, uint32_t d, uint32_t colour)
{
  CU(cudaSetDevice(s->dev[d]));
  CU(cudaEventRecord(s->produced[d], primary(s->g[d])));
  for(uint32_t o = 0; o < s->P; o++) {
    if(o == d) continue;
    mcx_graph *go = s->g[o];
    CU(cudaSetDevice(s->dev[o]));
    CU(cudaStreamWaitEvent(primary(go), s->produced[d], 0));
    go->occ_bound += s->cap;
    CU(mcx_launch_insert_tuples(s->ring_k[o] + (uint64_t)d * s->cap * s->W, s->ring_m[o] + (uint64_t)d * s->cap, s->cap,
                                s->counts[d] + o, s->k, go->table, colour, go->occ_bound >= 0xF0000000ull, go->d_counters, primary(go)));
    CU(cudaEventRecord(s->inserted[o][d], primary(go)));
    s->region_busy[o][d] = true;
  }
  return MCX_OK;
}
// d may write its regions (and counters) again only after the owners have read them
static int ss_wait_regions(mcx_shardset *s, uint32_t d)
{
  CU(cudaSetDevice(s->dev[d]));
  for(uint32_t o = 0; o < s->P; o++)
    if(o != d && s->region_busy[o][d]) { CU(cudaStreamWaitEvent(primary(s->g[d]), s->inserted[o][d], 0)); s->region_busy[o][d] = false; }
  return MCX_OK;
}
static int ss_flush(mcx_shardset *s)
{
  if(!s->dirty) return MCX_OK;
  for(uint32_t d = 0; d < s->P; d++) {
    uint64_t *kd[MCX_MAX_PARTS]; uint32_t *md[MCX_MAX_PARTS];
    ss_peer_pointers(s, d, kd, md);
    { int r = ss_wait_regions(s, d); if(r) return r; }
    { int r = flush_sharded(s->g[d], s->P, d, s->cap, NULL, NULL, kd, md, s->counts[d]); if(r) return r; }
    { int r = ss_insert_all(s, d, s->colour); if(r) return r; }
    s->pending[d] = 0;
  }
  s->dirty = false;
  return MCX_OK;
}

extern "C" int mcx_shardset_destroy(mcx_shardset *s)
{
  if(!s) return MCX_OK;
  for(uint32_t d = 0; d < s->P; d++) { cudaSetDevice(s->dev[d]); cudaDeviceSynchronize(); }
  for(uint32_t d = 0; d < s->P; d++) {
    cudaSetDevice(s->dev[d]);
    if(s->exp_buf[d]) cudaFreeHost(s->exp_buf[d]);
    for(int i = 0; i < MCX_SS_NSLOT; i++) {
      if(s->d_stage[d][i]) cudaFree(s->d_stage[d][i]);
      if(s->h_stage[d][i]) cudaFreeHost(s->h_stage[d][i]);
      if(s->d_qstage[d][i]) cudaFree(s->d_qstage[d][i]);
      if(s->h_qstage[d][i]) cudaFreeHost(s->h_qstage[d][i]);
      if(s->staged[d][i]) cudaEventDestroy(s->staged[d][i]);
      if(s->stage_free[d][i]) cudaEventDestroy(s->stage_free[d][i]);
    }
    if(s->copy[d]) cudaStreamDestroy(s->copy[d]);
    if(s->produced[d]) cudaEventDestroy(s->produced[d]);
    for(uint32_t o = 0; o < s->P; o++) if(s->inserted[d][o]) cudaEventDestroy(s->inserted[d][o]);
    if(s->ring_k[d]) cudaFree(s->ring_k[d]);
    if(s->ring_m[d]) cudaFree(s->ring_m[d]);
    if(s->counts[d]) cudaFree(s->counts[d]);
    if(s->g[d]) mcx_graph_destroy(s->g[d]);
  }
  free(s);
  return MCX_OK;
}

extern "C" int mcx_shardset_create(uint32_t k, uint32_t ncols, uint64_t capacity, const int *devices, uint32_t ndevices, mcx_shardset **out)
{
  if(!out || !devices || ndevices < 2 || ndevices > MCX_MAX_PARTS || capacity == 0) return MCX_ERR_BAD_ARG;
  const int ndev = mcx_device_count();
  if(ndev == 0) { snprintf(g_err, sizeof(g_err), "no CUDA device: libmcxgpu has no CPU fallback"); return MCX_ERR_NO_DEVICE; }
  for(uint32_t d = 0; d < ndevices; d++) {
    if(devices[d] < 0 || devices[d] >= ndev) return MCX_ERR_BAD_ARG;
    for(uint32_t e = 0; e < d; e++) if(devices[e] == devices[d]) { snprintf(g_err, sizeof(g_err), "a device may hold one shard only"); return MCX_ERR_BAD_ARG; }
  }
  mcx_shardset *s = (mcx_shardset *)calloc(1, sizeof(*s));
  if(!s) return MCX_ERR_NOMEM;
  s->k = k; s->W = (k + 31u) / 32u; s->ncols = ncols; s->P = ndevices;
  // a piece is MCX_STAGE_POS positions; in the worst case (nothing absorbed by the front table) its occurrences
  // spread evenly over the owners: 1.5 x that + the front table's flush (at most one tuple per slot, spread likewise)
  s->cap = (MCX_STAGE_POS / ndevices) * 3u / 2u + (1u << 20);
  // a shard holds 1/P of the k-mers (+ 10 % for the spread of a hash partition)
  const uint64_t cap_shard = capacity / ndevices + capacity / ndevices / 10u + 1024u;
  int rc = MCX_OK;
  for(uint32_t d = 0; d < ndevices && rc == MCX_OK; d++) {
    s->dev[d] = devices[d];
    rc = mcx_graph_create(k, ncols, cap_shard, devices[d], 0, &s->g[d]);
  }
#define SS_CU(x) do { if(rc == MCX_OK) { cudaError_t e_ = (x); if(e_ != cudaSuccess) rc = fail_cuda(e_, #x); } } while(0)
  for(uint32_t d = 0; d < ndevices && rc == MCX_OK; d++) {
    SS_CU(cudaSetDevice(devices[d]));
    for(uint32_t o = 0; o < ndevices && rc == MCX_OK; o++) {
      if(o == d) continue;
      int can = 0;
      SS_CU(cudaDeviceCanAccessPeer(&can, devices[d], devices[o]));
      if(rc == MCX_OK && !can) { snprintf(g_err, sizeof(g_err), "GPU %d cannot access GPU %d (no peer access)", devices[d], devices[o]); rc = MCX_ERR_UNSUPPORTED; }
      if(rc == MCX_OK) { cudaError_t e = cudaDeviceEnablePeerAccess(devices[o], 0); if(e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = fail_cuda(e, "cudaDeviceEnablePeerAccess"); cudaGetLastError(); }
    }
    SS_CU(cudaMalloc(&s->ring_k[d], (size_t)ndevices * s->cap * s->W * sizeof(uint64_t)));
    SS_CU(cudaMalloc(&s->ring_m[d], (size_t)ndevices * s->cap * sizeof(uint32_t)));
    SS_CU(cudaMalloc(&s->counts[d], MCX_MAX_PARTS * sizeof(uint64_t)));
    SS_CU(cudaMemset(s->counts[d], 0, MCX_MAX_PARTS * sizeof(uint64_t)));
    SS_CU(cudaStreamCreateWithFlags(&s->copy[d], cudaStreamNonBlocking));
    SS_CU(cudaEventCreateWithFlags(&s->produced[d], cudaEventDisableTiming));
    for(uint32_t o = 0; o < ndevices; o++) SS_CU(cudaEventCreateWithFlags(&s->inserted[d][o], cudaEventDisableTiming));
    for(int i = 0; i < MCX_SS_NSLOT; i++) {
      SS_CU(cudaMalloc(&s->d_stage[d][i], MCX_STAGE_BYTES));
      SS_CU(cudaHostAlloc(&s->h_stage[d][i], MCX_STAGE_BYTES, cudaHostAllocDefault));
      SS_CU(cudaEventCreateWithFlags(&s->staged[d][i], cudaEventDisableTiming));
      SS_CU(cudaEventCreateWithFlags(&s->stage_free[d][i], cudaEventDisableTiming));
    }
  }
#undef SS_CU
  if(rc != MCX_OK) { mcx_shardset_destroy(s); return rc; }
  *out = s;
  return MCX_OK;
}

// replaces build_graph() over one shared table (src/tools/build_graph.c:192-301) for a host batch in LINES layout
extern "C" int mcx_shardset_add_reads(mcx_shardset *s, const mcx_read_batch *b)
{
  if(!s || !b || b->colour >= s->ncols) return MCX_ERR_BAD_ARG;
  if(b->layout != MCX_LAYOUT_LINES || b->mem != MCX_MEM_HOST || (b->nbytes && !b->seq)) return MCX_ERR_BAD_ARG;
  if(b->must_exist) { snprintf(g_err, sizeof(g_err), "sharded builds cannot intersect (must_exist)"); return MCX_ERR_UNSUPPORTED; }
  if(b->hp_cutoff == 1 || b->hp_cutoff > s->k) { snprintf(g_err, sizeof(g_err), "hp_cutoff must be 0 or in [2, k]"); return MCX_ERR_UNSUPPORTED; }
  if(b->fq_cutoff >= 127) { snprintf(g_err, sizeof(g_err), "fq_cutoff (incl. offset) must be < 127"); return MCX_ERR_UNSUPPORTED; }
  if(b->nbytes == 0) return MCX_OK;
  if(s->exp_on) return MCX_ERR_BAD_ARG;
  if(s->dirty && b->colour != s->colour) { int r = ss_flush(s); if(r) return r; }   // the front tables count one colour at a time
  s->colour = b->colour;
  const bool useq = b->fq_cutoff && b->qual;
  const uint8_t *hseq = (const uint8_t *)b->seq, *hqual = (const uint8_t *)b->qual;
  const uint64_t nbytes = b->nbytes;
  cudaPointerAttributes attr;
  const bool pinned = (cudaPointerGetAttributes(&attr, hseq) == cudaSuccess) && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  for(uint64_t pos = 0; pos < nbytes;) {
    uint64_t pend = pos + MCX_STAGE_POS < nbytes ? pos + MCX_STAGE_POS : nbytes;
    uint64_t b0 = pos ? pos - MCX_LB : 0, b1 = pend + MCX_TAIL < nbytes ? pend + MCX_TAIL : nbytes;
    if(useq) {
      // contig membership under a quality cut-off carries across chunks: a launch covers whole reads (cut after a terminator)
      if(pend < nbytes) {
        const void *nl = memrchr(hseq + pos, '\n', pend - pos);
        if(!nl) { snprintf(g_err, sizeof(g_err), "a read longer than %llu bases cannot be quality-filtered in a sharded build", (unsigned long long)MCX_STAGE_POS); return MCX_ERR_UNSUPPORTED; }
        pend = (uint64_t)((const uint8_t *)nl - hseq) + 1u;
      }
      b0 = pos; b1 = pend;   // (pos is 16-byte aligned in the source only by chance: the staging copy realigns it)
    }
    // every device's front table counts in 32 bits: flush them all before any could wrap
    for(uint32_t d = 0; d < s->P; d++) if(s->pending[d] + MCX_STAGE_POS >= 0xE0000000ull) { int r = ss_flush(s); if(r) return r; break; }
    const uint32_t d = s->next++ % s->P;
    mcx_graph *g = s->g[d];
    CU(cudaSetDevice(s->dev[d]));
    const int sl = s->slot[d]; s->slot[d] = (sl + 1) % MCX_SS_NSLOT;
    CU(cudaEventSynchronize(s->stage_free[d][sl]));       // the kernel that last read this staging slot is done
    const uint8_t *src = hseq + b0;
    if(!pinned) { memcpy(s->h_stage[d][sl], src, b1 - b0); src = s->h_stage[d][sl]; }
    CU(cudaMemcpyAsync(s->d_stage[d][sl], src, b1 - b0, cudaMemcpyHostToDevice, s->copy[d]));
    if(useq) {
      if(!s->d_qstage[d][sl]) { CU(cudaMalloc(&s->d_qstage[d][sl], MCX_STAGE_BYTES)); CU(cudaHostAlloc(&s->h_qstage[d][sl], MCX_STAGE_BYTES, cudaHostAllocDefault)); }
      const uint8_t *qsrc = hqual + b0;
      if(!pinned) { memcpy(s->h_qstage[d][sl], qsrc, b1 - b0); qsrc = s->h_qstage[d][sl]; }
      CU(cudaMemcpyAsync(s->d_qstage[d][sl], qsrc, b1 - b0, cudaMemcpyHostToDevice, s->copy[d]));
    }
    CU(cudaEventRecord(s->staged[d][sl], s->copy[d]));
    CU(cudaStreamWaitEvent(primary(g), s->staged[d][sl], 0));
    { int r = ss_wait_regions(s, d); if(r) return r; }
    uint64_t *kd[MCX_MAX_PARTS]; uint32_t *md[MCX_MAX_PARTS];
    ss_peer_pointers(s, d, kd, md);
    mcx_read_batch db = *b;
    db.mem = MCX_MEM_DEVICE; db.seq = (const char *)s->d_stage[d][sl]; db.nbytes = b1 - b0;
    db.qual = useq ? (const char *)s->d_qstage[d][sl] : NULL;
    { int r = add_reads_sharded(g, &db, s->P, d, s->cap, NULL, NULL, kd, md, s->counts[d], useq ? 0 : pos - b0, useq ? 0 : pend - b0); if(r) return r; }
    CU(cudaEventRecord(s->stage_free[d][sl], primary(g)));
    { int r = ss_insert_all(s, d, b->colour); if(r) return r; }
    s->pending[d] += pend - pos;
    s->dirty = true;
    pos = pend;
  }
  s->bytes_submitted += nbytes;
  return MCX_OK;
}

// join everything; stats = the counters of all shards since the previous sync
extern "C" int mcx_shardset_sync(mcx_shardset *s, mcx_load_stats *stats)
{
  if(!s) return MCX_ERR_BAD_ARG;
  { int r = ss_flush(s); if(r) return r; }
  for(uint32_t d = 0; d < s->P; d++) { CU(cudaSetDevice(s->dev[d])); CU(cudaStreamSynchronize(s->copy[d])); CU(cudaStreamSynchronize(primary(s->g[d]))); }
  mcx_load_stats tot; memset(&tot, 0, sizeof(tot));
  int rc = MCX_OK;
  for(uint32_t d = 0; d < s->P; d++) {
    mcx_load_stats st;
    const int r = mcx_graph_sync(s->g[d], &st);
    if(r && rc == MCX_OK) rc = r;
    tot.total_bases_read += st.total_bases_read; tot.total_bases_loaded += st.total_bases_loaded;
    tot.contigs_parsed += st.contigs_parsed; tot.num_kmers_loaded += st.num_kmers_loaded; tot.num_kmers_novel += st.num_kmers_novel;
    tot.num_se_reads += st.num_se_reads;
  }
  tot.num_good_reads = tot.num_bad_reads = UINT64_MAX;
  if(stats) *stats = tot;
  return rc;
}

extern "C" int mcx_shardset_stats(mcx_shardset *s, uint64_t *nkmers, uint64_t *capacity)
{
  if(!s) return MCX_ERR_BAD_ARG;
  uint64_t n = 0, c = 0;
  for(uint32_t d = 0; d < s->P; d++) { n += s->g[d]->nkmers; c += s->g[d]->capacity; }
  if(nkmers) *nkmers = n;
  if(capacity) *capacity = c;
  return MCX_OK;
}

// replaces graph_writer_save_mkhdr's record stream (src/graph/graph_writer.c:182-193): every shard exports its records
// (sorted: ascending keys); ownership is by hash, so the file order is the P-way merge of the shards' runs
extern "C" int mcx_shardset_export_begin(mcx_shardset *s, int sorted, uint64_t *nrecords, uint32_t *record_bytes)
{
  if(!s) return MCX_ERR_BAD_ARG;
  { int r = mcx_shardset_sync(s, NULL); if(r) return r; }
  uint64_t tot = 0;
  for(uint32_t d = 0; d < s->P; d++) {
    uint32_t rb = 0;
    int r = mcx_graph_export_begin(s->g[d], sorted, &s->exp_n[d], &rb);
    if(r) return r;
    s->rec_bytes = rb; s->exp_at[d] = s->exp_have[d] = s->exp_pos[d] = 0;
    tot += s->exp_n[d];
    if(!s->exp_buf[d]) { CU(cudaSetDevice(s->dev[d])); CU(cudaHostAlloc(&s->exp_buf[d], (size_t)MCX_SS_CHUNK_RECS * rb, cudaHostAllocDefault)); }
  }
  s->exp_on = true; s->exp_sorted = sorted != 0; s->exp_cur = 0;
  if(nrecords) *nrecords = tot;
  if(record_bytes) *record_bytes = s->rec_bytes;
  return MCX_OK;
}
static int ss_refill(mcx_shardset *s, uint32_t d)
{
  const uint64_t left = s->exp_n[d] - s->exp_at[d];
  const uint64_t n = left < MCX_SS_CHUNK_RECS ? left : MCX_SS_CHUNK_RECS;
  s->exp_have[d] = n; s->exp_pos[d] = 0;
  if(n == 0) return MCX_OK;
  int r = mcx_graph_export_read(s->g[d], s->exp_at[d], n, s->exp_buf[d]);
  s->exp_at[d] += n;
  return r;
}
// the next (at most max_records) records of the merged stream -> host_dst; *got = 0 at the end
extern "C" int mcx_shardset_export_next(mcx_shardset *s, void *host_dst, uint64_t max_records, uint64_t *got)
{
  if(!s || !s->exp_on || !got || (!host_dst && max_records)) return MCX_ERR_BAD_ARG;
  const uint32_t rb = s->rec_bytes, W = s->W;
  uint8_t *out = (uint8_t *)host_dst;
  uint64_t n = 0;
  while(n < max_records) {
    uint32_t best = UINT32_MAX; uint64_t b0 = 0, b1 = 0;
    for(uint32_t d = s->exp_sorted ? 0 : s->exp_cur; d < s->P; d++) {
      if(s->exp_pos[d] == s->exp_have[d]) {
        if(s->exp_at[d] == s->exp_n[d]) { if(!s->exp_sorted) { s->exp_cur = d + 1; } continue; }
        int r = ss_refill(s, d); if(r) return r;
      }
      const uint8_t *rec = s->exp_buf[d] + s->exp_pos[d] * rb;
      uint64_t k0, k1 = 0;
      memcpy(&k0, rec, 8); if(W == 2) memcpy(&k1, rec + 8, 8);
      if(!s->exp_sorted) { best = d; break; }
      if(best == UINT32_MAX || k0 < b0 || (k0 == b0 && k1 < b1)) { best = d; b0 = k0; b1 = k1; }
    }
    if(best == UINT32_MAX) break;
    if(!s->exp_sorted) { // concatenation: copy what is left of this shard's chunk at once
      uint64_t m = s->exp_have[best] - s->exp_pos[best];
      if(m > max_records - n) m = max_records - n;
      memcpy(out + n * rb, s->exp_buf[best] + s->exp_pos[best] * rb, m * rb);
      s->exp_pos[best] += m; n += m;
      continue;
    }
    memcpy(out + n * rb, s->exp_buf[best] + s->exp_pos[best] * rb, rb);
    s->exp_pos[best]++; n++;
  }
  *got = n;
  return MCX_OK;
}
extern "C" int mcx_shardset_export_end(mcx_shardset *s)
{
  if(!s) return MCX_ERR_BAD_ARG;
  for(uint32_t d = 0; d < s->P; d++) mcx_graph_export_end(s->g[d]);
  s->exp_on = false;
  return MCX_OK;
}
