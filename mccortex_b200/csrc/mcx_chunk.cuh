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
