// mcx_chunk.cuh -- chunk geometry + the two per-chunk phases of the reads->k-mers
// stage, written as host/device functions over plain arrays.  On the GPU the
// arrays are shared memory (mcx_build.cu); in tests/emul they are host arrays,
// so the exact same code is checked against the oracle without a GPU.
//
// A chunk is MCX_T consecutive window start positions of the batch byte buffer
// (LINES layout: every read is followed by one '\n', so any non-ACGT byte also
// separates reads and a window can never span two reads).  The chunk stages
//   raw[0 .. MCX_RAW) = buffer bytes [chunk_start - MCX_LB, chunk_start + MCX_T + MCX_TAIL)
// i.e. 16 bytes of look-back (for the base before a window: the incoming edge)
// and k+1 <= 64 bytes of look-ahead (the window body and the base after it).
#pragma once
#include "mcx_device.cuh"

#ifndef MCX_CTA_THREADS
#define MCX_CTA_THREADS 256u             /* threads per CTA of the chunk kernels (128 was measured: see profiles/r1g_experiments.txt) */
#endif
#define MCX_T     (8u * MCX_CTA_THREADS) /* window starts per chunk: MCX_WPT per thread */
#define MCX_LB    16u                   /* look-back bytes */
#define MCX_TAIL  80u                   /* look-ahead bytes (>= 64 + 1, multiple of 16) */
#define MCX_RAW   (MCX_LB + MCX_T + MCX_TAIL)   /* 2144 bytes, multiple of 16 */
#define MCX_PKW   (MCX_RAW / 16u + 4u)  /* packed words + over-read padding */
#define MCX_MSW   (MCX_RAW / 32u + 4u)  /* mask words + over-read padding */
#define MCX_VW    ((MCX_LB + MCX_T + 1u + 31u) / 32u + 3u) /* in-contig mask words: staged positions 0 .. LB+T, + over-read padding */
#define MCX_SEQ_PAD 256u                /* readable slack required after nbytes of a device buffer */

// Phase 1, one 16-byte item: raw bytes -> packed bases, bad / eq / newline bits.
//   w[4]     : the 16 raw bytes as little-endian u32
//   prev     : the byte before w[0] (anything for item 0)
//   gpos     : buffer offset of byte 0 of this item (may be "negative" = huge for the chunk-0 look-back)
//   nbytes   : bytes of real data in the buffer; positions >= nbytes read as terminators
MCX_HD void mcx_convert16(const uint32_t w[4], uint32_t prev, uint64_t gpos, uint64_t nbytes,
                          uint32_t *pk, uint32_t *bad16, uint32_t *eq16, uint32_t *nl16)
{
  uint32_t p = 0, b = 0, e = 0, n = 0;
#pragma unroll
  for(int i = 0; i < 4; i++) {
    p = (p << 8) | mcx_pack4(w[i]);
    b |= mcx_bad4(w[i]) << (4 * i);
    e |= mcx_eqprev4(w[i], i ? (w[i - 1] >> 24) : prev) << (4 * i);
    n |= mcx_nl4(w[i]) << (4 * i);
  }
  // bytes outside [0, nbytes): not data
  if(gpos >= nbytes || gpos + 16u > nbytes) { // first test also catches wrapped (negative) gpos
    uint32_t oob = 0;
    for(uint32_t i = 0; i < 16u; i++) if(gpos + i >= nbytes) oob |= 1u << i;
    b |= oob; n &= ~oob; e &= ~oob;
  }
  *pk = p; *bad16 = b; *eq16 = e; *nl16 = n;
}

// Quality bits of one 16-byte item (reference: seq_contig_start2 requires qual > cutoff for every
// base of the FIRST k-mer of a contig, seq_contig_end2 extends while qual >= cutoff:
// src/basic/seq_reader.c:84,149; `char` quality vs uint8_t cutoff compare after integer
// promotion, so bytes >= 0x80 are negative and fail).  Bit i of *weak16: byte i < cutoff
// (cannot extend a contig); bit i of *strong16: byte i <= cutoff (cannot start one).
MCX_HD void mcx_qual16(const uint32_t q[4], uint32_t qcut, uint32_t *weak16, uint32_t *strong16)
{
  uint32_t wk = 0, st = 0;
  for(uint32_t i = 0; i < 16u; i++) {
    int v = (int)(signed char)((q[i >> 2] >> (8u * (i & 3u))) & 0xFFu);
    wk |= (uint32_t)(v < (int)qcut) << i;
    st |= (uint32_t)(v <= (int)qcut) << i;
  }
  *weak16 = wk; *strong16 = st;
}

// Contig membership with a quality cut-off is not a local predicate:
//   in_contig(p) = ev(p) && (sv(p) || in_contig(p-1))
// ev = window may EXTEND a contig (all bases ACGT, qual >= cutoff, no homopolymer run),
// sv = window may START one (qual > cutoff instead).  g = ev&sv generates, ev&~sv propagates:
// this is the carry chain of the addition ev + (ev&sv), so one 32-bit add resolves 32 windows.
// ev/sv/x: bit strings of nwords u32; cin = in_contig of the window before bit 0.
// Writes x (may be NULL) and returns in_contig of the last bit.
MCX_HD uint32_t mcx_contig_chain(const uint32_t *ev, const uint32_t *sv, uint32_t nwords, uint32_t cin, uint32_t *x)
{
  for(uint32_t w = 0; w < nwords; w++) {
    uint32_t a = ev[w], b = ev[w] & sv[w];
    uint64_t sum = (uint64_t)a + b + cin;
    uint32_t into = (uint32_t)sum ^ a ^ b;      // carry INTO each bit
    cin = (uint32_t)(sum >> 32) & 1u;           // carry out of bit 31
    if(x) x[w] = (into >> 1) | (cin << 31);     // carry OUT of each bit = in_contig
  }
  return cin;
}

// ---- the same chain, warp-wide: the lane-local pieces (the device wrappers in mcx_build.cu add the shuffles, tests/emul
// walks the 32 lanes on the CPU).  Lane l owns the MCX_CHAIN_WPL consecutive mask words starting at l * MCX_CHAIN_WPL.
// The carry-in of the chunk is planted at bit cb of word 0 (ev = sv = cin there, everything below cleared) and the chain
// itself starts with carry 0 -- exactly what the serial callers of mcx_contig_chain do.
#define MCX_CHAIN_WPL ((MCX_VW + 31u) / 32u)
MCX_HD void mcx_chain_lane_load(const uint32_t *vm, const uint32_t *sv, uint32_t cb, uint32_t cin, uint32_t lane,
                                uint32_t a[MCX_CHAIN_WPL], uint32_t b[MCX_CHAIN_WPL])
{
#pragma unroll
  for(uint32_t i = 0; i < MCX_CHAIN_WPL; i++) {
    const uint32_t w = lane * MCX_CHAIN_WPL + i;
    uint32_t e = w < MCX_VW ? vm[w] : 0u, t = w < MCX_VW ? sv[w] : 0u;
    if(w == 0) {
      const uint32_t keep = (~0u << cb) & ~(1u << cb);
      e = (e & keep) | (cin << cb); t = (t & keep) | (cin << cb);
    }
    a[i] = e; b[i] = e & t;
  }
}
// carry out of the lane's words for carry-in 0 (*f0) and 1 (*f1)
MCX_HD void mcx_chain_lane_carry(const uint32_t a[MCX_CHAIN_WPL], const uint32_t b[MCX_CHAIN_WPL], uint32_t *f0, uint32_t *f1)
{
  uint32_t c0 = 0u, c1 = 1u;
#pragma unroll
  for(uint32_t i = 0; i < MCX_CHAIN_WPL; i++) {
    c0 = (uint32_t)(((uint64_t)a[i] + b[i] + c0) >> 32);
    c1 = (uint32_t)(((uint64_t)a[i] + b[i] + c1) >> 32);
  }
  *f0 = c0; *f1 = c1;
}
// (f0, f1) := this lane's carry function applied after the one of the lanes to its left (l0, l1)
MCX_HD void mcx_chain_compose(uint32_t l0, uint32_t l1, uint32_t *f0, uint32_t *f1)
{
  const uint32_t n0 = l0 ? *f1 : *f0, n1 = l1 ? *f1 : *f0;
  *f0 = n0; *f1 = n1;
}
// the lane's words again, with its real carry-in c: in_contig bits written over vm
MCX_HD void mcx_chain_lane_store(uint32_t *vm, const uint32_t a[MCX_CHAIN_WPL], const uint32_t b[MCX_CHAIN_WPL], uint32_t c, uint32_t lane)
{
#pragma unroll
  for(uint32_t i = 0; i < MCX_CHAIN_WPL; i++) {
    const uint32_t w = lane * MCX_CHAIN_WPL + i;
    const uint64_t sum = (uint64_t)a[i] + b[i] + c;
    const uint32_t into = (uint32_t)sum ^ a[i] ^ b[i];
    c = (uint32_t)(sum >> 32) & 1u;
    if(w < MCX_VW) vm[w] = (into >> 1) | (c << 31);
  }
}
// Chunk summary without walking the chain.  With z = the last window in (cb, f] that cannot extend a contig (ev = 0),
// window f is in a contig iff some window after z can start one (sv = 1) -- or, if there is no such z, iff the carry-in
// was set.  Lane l looks at words l, l + 32, ...: its candidate for z (-1: none) ...
MCX_HD uint32_t mcx_summary_word_mask(uint32_t w, uint32_t cb, uint32_t f)
{
  uint32_t m = ~0u;
  if(w == 0) m &= ~0u << (cb + 1u);
  if(w == (f >> 5) && (f & 31u) != 31u) m &= (1u << ((f & 31u) + 1u)) - 1u;
  return m;
}
MCX_HD uint32_t mcx_clz32(uint32_t x) // x != 0
{
#if defined(__CUDA_ARCH__)
  return (uint32_t)__clz((int)x);
#else
  return (uint32_t)__builtin_clz(x);
#endif
}
MCX_HD int mcx_summary_lane_last_zero(const uint32_t *ev, uint32_t cb, uint32_t f, uint32_t lane)
{
  int z = -1;
  for(uint32_t w = lane; w <= (f >> 5); w += 32u) {
    const uint32_t zeros = ~ev[w] & mcx_summary_word_mask(w, cb, f);
    if(zeros) { const int c = (int)(w * 32u + 31u) - (int)mcx_clz32(zeros); z = c > z ? c : z; }
  }
  return z;
}
// ... and, z being the maximum over the lanes, whether one of its words has a start bit after z
MCX_HD uint32_t mcx_summary_lane_starts(const uint32_t *sv, uint32_t cb, uint32_t f, int z, uint32_t lane)
{
  uint32_t any = 0;
  for(uint32_t w = lane; w <= (f >> 5); w += 32u) {
    uint32_t m = mcx_summary_word_mask(w, cb, f);
    if((int)(w * 32u + 31u) <= z) continue;
    if((int)(w * 32u) <= z) m &= ~0u << (uint32_t)(z - (int)(w * 32u) + 1);
    any |= sv[w] & m;
  }
  return any;
}

// ---------------------------------------------------------------------------
// Phase 2a, word-parallel: 32 windows per call instead of one.
// Masks are indexed by STAGED POSITION q (byte q of raw[]); window q = bases q .. q+k-1.
// ---------------------------------------------------------------------------
struct McxBits128 { uint64_t lo, hi; };
MCX_HD McxBits128 mcx_b128_load(const uint32_t *m, uint32_t j)
{
  McxBits128 x;
  x.lo = ((uint64_t)m[j + 1] << 32) | m[j];
  x.hi = ((uint64_t)m[j + 3] << 32) | m[j + 2];
  return x;
}
MCX_HD McxBits128 mcx_b128_shr(McxBits128 x, uint32_t s) // 1 <= s <= 63
{
  McxBits128 r; r.lo = (x.lo >> s) | (x.hi << (64u - s)); r.hi = x.hi >> s; return r;
}
// out[q] = OR_{d<n} in[q+d]   (n >= 1; bits beyond the 128 loaded read as 0)
MCX_HD McxBits128 mcx_b128_dilate(McxBits128 x, uint32_t n)
{
  uint32_t have = 1;
  while(have < n) {
    uint32_t s = (n - have < have) ? (n - have) : have;
    McxBits128 y = mcx_b128_shr(x, s);
    x.lo |= y.lo; x.hi |= y.hi; have += s;
  }
  return x;
}
// out[q] = AND_{d<n} in[q+d]
MCX_HD McxBits128 mcx_b128_erode(McxBits128 x, uint32_t n)
{
  uint32_t have = 1;
  while(have < n) {
    uint32_t s = (n - have < have) ? (n - have) : have;
    McxBits128 y = mcx_b128_shr(x, s);
    x.lo &= y.lo; x.hi &= y.hi; have += s;
  }
  return x;
}

// Loadable-window bits for positions 32j .. 32j+31 (row B, local form; see mcx_window_ok in
// mcx_device.cuh for the per-window statement of the same predicate and DESIGN.md for why it
// equals seq_contig_start2/seq_contig_end2).  Needs mask words j .. j+3.
MCX_HD uint32_t mcx_valid_word(const uint32_t *bad, const uint32_t *eq, uint32_t j, uint32_t k, uint32_t hp_cutoff)
{
  McxBits128 d = mcx_b128_dilate(mcx_b128_load(bad, j), k);      // any bad base in [q, q+k)
  uint32_t v = ~(uint32_t)d.lo;
  if(hp_cutoff > 1u) {
    // a run of hp equal chars inside the window = hp-1 consecutive eq bits starting in [q+1, q+k-hp+1]
    McxBits128 a = mcx_b128_erode(mcx_b128_load(eq, j), hp_cutoff - 1u);
    McxBits128 h = mcx_b128_dilate(mcx_b128_shr(a, 1u), k - hp_cutoff + 1u);
    v &= ~(uint32_t)h.lo;
  }
  return v;
}

// ---------------------------------------------------------------------------
// Phase 2b: one thread walks MCX_WPT consecutive windows with ROLLING forward and
// reverse-complement k-mers (shift in one base per step) instead of re-extracting and
// re-reversing every window: ~4x fewer instructions per window (ncu: the first kernel was
// issue-bound at 15 warp-instructions per occurrence once the table accesses hit L2).
// ---------------------------------------------------------------------------
#define MCX_WPT 8u /* windows per thread: 256 threads x 8 = MCX_T */


template <int W> MCX_HD void mcx_roll(McxKmer<W> &f, McxKmer<W> &r, uint32_t b, uint32_t k);
template <> MCX_HD void mcx_roll<1>(McxKmer<1> &f, McxKmer<1> &r, uint32_t b, uint32_t k)
{
  f.b[0] = ((f.b[0] << 2) | b) & (~0ull >> (64u - 2u * k));
  r.b[0] = (r.b[0] >> 2) | ((uint64_t)(3u - b) << (2u * k - 2u));
}
template <> MCX_HD void mcx_roll<2>(McxKmer<2> &f, McxKmer<2> &r, uint32_t b, uint32_t k)
{
  const uint32_t top = 2u * (k - 32u); // bits used in b[0], 2..62
  f.b[0] = ((f.b[0] << 2) | (f.b[1] >> 62)) & (~0ull >> (64u - top));
  f.b[1] = (f.b[1] << 2) | b;
  r.b[1] = (r.b[1] >> 2) | (r.b[0] << 62);
  r.b[0] = (r.b[0] >> 2) | ((uint64_t)(3u - b) << (top - 2u));
}
template <int W> MCX_HD uint32_t mcx_first_base(const McxKmer<W> &f, uint32_t k)
{
  return (uint32_t)(f.b[0] >> (W == 1 ? 2u * k - 2u : 2u * (k - 32u) - 2u)) & 3u;
}

// Thread t of the CTA: windows at staged positions MCX_LB + 8t + j, j < 8.  `vmask` is indexed by
// staged position (bit q = window q is in a contig).  The windows are produced MCX_HALF at a time:
// fn(keys, emasks, valid, starts, j0) receives the canonical keys / edge masks of windows
// j0 .. j0+MCX_HALF-1, the bit mask of those that are in a contig and the subset that begin one.
// A group is complete before any table access so that the sink can keep several probe loads in
// flight (the kernel is L2-latency bound otherwise) without holding all eight keys in registers.
#define MCX_HALF 4u
// mcx_thread_occurrences2: `pre(j0)` is called before the keys of each group are computed, `fn` after it even
// when the group has no window in a contig (valid == 0); returns false if the thread has no window in a
// contig at all (then neither hook was called).  The warp-specialised kernel stages EVERY window.
template <int W, class P, class F>
MCX_HD bool mcx_thread_occurrences2(const uint32_t *pk, const uint32_t *vmask, uint32_t t, uint32_t k, P &&pre, F &&fn);

template <int W, class F>
MCX_HD void mcx_thread_occurrences(const uint32_t *pk, const uint32_t *vmask, uint32_t t, uint32_t k, F &&fn)
{
  mcx_thread_occurrences2<W>(pk, vmask, t, k, [](uint32_t) {}, [&](const McxKmer<W> *keys, const uint32_t *emasks, uint32_t valid,
                                                                uint32_t starts, uint32_t j0) { if(valid) fn(keys, emasks, valid, starts, j0); });
}

template <int W, class P, class F>
MCX_HD bool mcx_thread_occurrences2(const uint32_t *pk, const uint32_t *vmask, uint32_t t, uint32_t k, P &&pre, F &&fn)
{
  const uint32_t p0 = MCX_LB + MCX_WPT * t;
  // in_contig bits of windows p0-1 .. p0+8 (bit 0 = the window before ours)
  const uint32_t vb = (uint32_t)(mcx_get64bits(vmask, p0 - 1u)) & 0x3FFu;
  if(!(vb & 0x1FEu)) return false;
  McxKmer<W> f = mcx_kmer_at<W>(pk, p0, k);
  McxKmer<W> r = mcx_kmer_revcomp<W>(f, k);
  const uint64_t nx = mcx_get32bases(pk, p0 + k);      // bases p0+k .. : the ones shifted in
  uint32_t prev = mcx_get_base(pk, p0 - 1u);           // base before the current window
#pragma unroll
  for(uint32_t j0 = 0; j0 < MCX_WPT; j0 += MCX_HALF) {
    pre(j0);
    McxKmer<W> keys[MCX_HALF]; uint32_t emasks[MCX_HALF];
#pragma unroll
    for(uint32_t i = 0; i < MCX_HALF; i++) {
      const uint32_t j = j0 + i;
      const uint32_t next = (uint32_t)(nx >> (62u - 2u * j)) & 3u;
      bool rc_lt;
      if(W == 1) rc_lt = r.b[0] < f.b[0];
      else rc_lt = (r.b[0] < f.b[0]) || (r.b[0] == f.b[0] && r.b[W - 1] < f.b[W - 1]);
      keys[i] = rc_lt ? r : f;
      emasks[i] = mcx_edge_mask(rc_lt ? 1u : 0u, (vb >> j) & 1u, prev, (vb >> (j + 2u)) & 1u, next);
      prev = mcx_first_base<W>(f, k);
      mcx_roll<W>(f, r, next, k);
    }
    const uint32_t valid = (vb >> (j0 + 1u)) & ((1u << MCX_HALF) - 1u);
    const uint32_t starts = valid & ~(vb >> j0);
    fn(keys, emasks, valid, starts, j0);
  }
  return true;
}
