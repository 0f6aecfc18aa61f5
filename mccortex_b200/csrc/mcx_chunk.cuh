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

#define MCX_T     2048u                 /* window starts per chunk */
#define MCX_LB    16u                   /* look-back bytes */
#define MCX_TAIL  80u                   /* look-ahead bytes (>= 64 + 1, multiple of 16) */
#define MCX_RAW   (MCX_LB + MCX_T + MCX_TAIL)   /* 2144 bytes, multiple of 16 */
#define MCX_PKW   (MCX_RAW / 16u + 4u)  /* packed words + over-read padding */
#define MCX_MSW   (MCX_RAW / 32u + 4u)  /* mask words + over-read padding */
#define MCX_VW    ((MCX_T + 2u + 31u) / 32u + 1u) /* valid-mask words for windows -1 .. T */
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

// Phase 2a: is the window with local index i (buffer start = chunk_start - 1 + i) loadable?
MCX_HD bool mcx_chunk_window_ok(const uint32_t *bad, const uint32_t *eq, uint32_t i, uint32_t k, uint32_t hp_cutoff)
{
  return mcx_window_ok(bad, eq, MCX_LB - 1u + i, k, hp_cutoff);
}

// Phase 2b: everything one occurrence contributes, from the staged arrays.
template <int W> struct McxOcc { McxKmer<W> key; uint32_t hc, hb, emask, orient; };

template <int W>
MCX_HD McxOcc<W> mcx_chunk_occurrence(const uint32_t *pk, const uint32_t *vmask, uint32_t i, uint32_t k)
{
  McxOcc<W> o;
  uint32_t p = MCX_LB - 1u + i; // position of the window's first base in the staged arrays
  McxKmer<W> f = mcx_kmer_at<W>(pk, p, k);
  o.key = mcx_kmer_key<W>(f, k, &o.orient);
  o.hc = mcx_lookup3<W>(o.key, 0u, &o.hb);
  bool has_prev = mcx_get_bit(vmask, i - 1u), has_next = mcx_get_bit(vmask, i + 1u);
  o.emask = mcx_edge_mask(o.orient, has_prev, mcx_get_base(pk, p - 1u), has_next, mcx_get_base(pk, p + k));
  return o;
}
