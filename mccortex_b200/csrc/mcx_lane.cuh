// mcx_lane.cuh -- the warp-autonomous front end, lane by lane (k <= 31, no quality / homopolymer cut-off).
//
// The batch byte buffer (LINES layout) is cut into TILES of 512 window-start positions = 32 PIECES of
// 16 bytes.  A warp owns a run of consecutive tiles; lane l of the warp owns piece 32t + l of tile t,
// i.e. the MCX_LW = 16 windows that start in it.  For those it needs
//   * its own piece and the two that follow (window 15 of the lane ends at base 15 + 30 = 45, the
//     base after it is 46 < 48): packed bases pk0..pk2 and 48 "bad base" bits,
//   * the last base of the piece before (the base that precedes window 0: the incoming edge) and
//     whether it is bad.
// On the GPU (mcx_build_warp.cu) the neighbouring pieces come from the neighbouring lanes by shuffle --
// no shared-memory staging of converted data, no CTA barrier; everything here is MCX_HD so that
// tests/emul runs the same lane math on the CPU against the oracle.
//
// What it replaces in the reference is what mcx_chunk.cuh replaces (rows A-D, G of SURVEY 8a):
//   dna_char_to_nuc_arr               src/basic/dna.c:8-25
//   seq_contig_start2 / end2          src/basic/seq_reader.c:61-172   (no quality, no homopolymer rule:
//                                     contigs = maximal runs of windows without a non-ACGT base)
//   binary_kmer_left_shift_add, binary_kmer_reverse_complement, binary_kmer_get_key
//                                     src/graph/binary_kmer.{h,c}
//   db_graph_add_edge_mt (local form) src/graph/db_graph.c:152-166
#pragma once
#include "mcx_chunk.cuh"

#define MCX_LW 16u          /* windows per lane = bytes per piece */
#define MCX_TILE 512u       /* positions per tile = 32 lanes x MCX_LW */

// One piece: 16 raw bytes (little-endian u32 x 4) -> packed bases (base 0 in the top two bits),
// bad bits (bit i = byte i is not ACGTacgt), newline bits.  gpos = buffer offset of byte 0; bytes at
// offsets >= nbytes are not data: bad, and never a read terminator.
MCX_HD void mcx_piece_convert(const uint32_t w[4], uint64_t gpos, uint64_t nbytes, uint32_t *pk, uint32_t *bad16, uint32_t *nl16)
{
  uint32_t p = 0, b = 0, n = 0;
#pragma unroll
  for(int i = 0; i < 4; i++) {
    p = (p << 8) | mcx_pack4(w[i]);
    b |= mcx_bad4(w[i]) << (4 * i);
  }
  if(b) { // terminators are rare (one piece in ten on 150 bp reads): only then look for '\n'
#pragma unroll
    for(int i = 0; i < 4; i++) n |= mcx_nl4(w[i]) << (4 * i);
  }
  if(gpos + 16u > nbytes) {
    const uint32_t oob = gpos >= nbytes ? 0xFFFFu : (0xFFFFu << (uint32_t)(nbytes - gpos)) & 0xFFFFu;
    b |= oob; n &= ~oob;
  }
  *pk = p; *bad16 = b; *nl16 = n;
}

// bit i of the result = AND of x[i .. i+n-1]  (n >= 1; bits shifted in from above are 0)
MCX_HD uint64_t mcx_and_run64(uint64_t x, uint32_t n)
{
  uint32_t have = 1;
  while(have < n) {
    const uint32_t s = (n - have < have) ? (n - have) : have;
    x &= x >> s;
    have += s;
  }
  return x;
}

// in-contig bits of windows -1 .. 16 of a lane (bit j+1 = window j; no quality / homopolymer rule:
// a window is loadable iff none of its k bases is bad).  bad48: bit i = base i of the lane's three
// pieces is bad; prev_bad: the base before base 0 is bad.
MCX_HD uint32_t mcx_lane_valid(uint64_t bad48, uint32_t prev_bad, uint32_t k)
{
  const uint64_t B = (bad48 << 1) | (uint64_t)(prev_bad & 1u) | (~0ull << 49);
  return (uint32_t)mcx_and_run64(~B, k) & 0x3FFFFu;
}

// The lane's 16 windows, MCX_HALF at a time: fn(keys, emasks, valid, starts, j0) is called for
// j0 = 0, 4, 8, 12 -- ALWAYS four times, also when no window of the group is in a contig, so that a
// sink that uses warp collectives stays convergent.  valid / starts are 4-bit masks.
//   pk0..pk2   packed bases of the lane's piece and the two that follow
//   vb         mcx_lane_valid()
//   prev_base  2-bit code of the base before base 0 (anything if it is bad)
template <class F>
MCX_HD void mcx_lane_windows(uint32_t pk0, uint32_t pk1, uint32_t pk2, uint32_t vb, uint32_t prev_base, uint32_t k, F &&fn)
{
  const uint64_t b01 = ((uint64_t)pk0 << 32) | pk1;
  McxKmer<1> f; f.b[0] = b01 >> (64u - 2u * k);
  McxKmer<1> r = mcx_kmer_revcomp<1>(f, k);
  // bases k .. k+15 (the ones shifted in), base k in the top two bits
  const uint64_t s = (k >> 4) ? (((uint64_t)pk1 << 32) | pk2) : b01;
  const uint32_t nx = (uint32_t)((s << ((k & 15u) * 2u)) >> 32);
  uint32_t prev = prev_base;
#if defined(__CUDA_ARCH__) && defined(MCX_LANE_ROLLED)
#pragma unroll 1   /* one copy of the group body: a quarter of the code (instruction cache) */
#else
#pragma unroll
#endif
  for(uint32_t j0 = 0; j0 < MCX_LW; j0 += MCX_HALF) {
    McxKmer<1> keys[MCX_HALF]; uint32_t emasks[MCX_HALF];
#pragma unroll
    for(uint32_t i = 0; i < MCX_HALF; i++) {
      const uint32_t j = j0 + i;
      const uint32_t next = (nx >> (30u - 2u * j)) & 3u;
      const bool rc_lt = r.b[0] < f.b[0];
      keys[i] = rc_lt ? r : f;
      emasks[i] = mcx_edge_mask(rc_lt ? 1u : 0u, (vb >> j) & 1u, prev, (vb >> (j + 2u)) & 1u, next);
      prev = mcx_first_base<1>(f, k);
      mcx_roll<1>(f, r, next, k);
    }
    const uint32_t valid = (vb >> (j0 + 1u)) & ((1u << MCX_HALF) - 1u);
    const uint32_t starts = valid & ~(vb >> j0);
    fn(keys, emasks, valid, starts, j0);
  }
}

// windows of piece `gpos` (a multiple of 16) owned by a launch over [r_begin, r_end): 16-bit mask
MCX_HD uint32_t mcx_piece_own(uint64_t gpos, uint64_t r_begin, uint64_t r_end)
{
  if(gpos >= r_begin && gpos + MCX_LW <= r_end) return 0xFFFFu;
  if(gpos + MCX_LW <= r_begin || gpos >= r_end) return 0u;
  const uint32_t lo = r_begin > gpos ? (uint32_t)(r_begin - gpos) : 0u;
  const uint32_t hi = r_end - gpos < MCX_LW ? (uint32_t)(r_end - gpos) : MCX_LW;
  return ((1u << hi) - 1u) & ~((1u << lo) - 1u);
}
