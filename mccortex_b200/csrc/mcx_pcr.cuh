// mcx_pcr.cuh -- build --remove-pcr: the per-read math, MCX_HD so tests/emul can run it on the CPU.
//
// Replaces (reference, relative to /root/reference):
//   seq_reads_are_novel             src/tools/build_graph.c:35-92
//   seq_reader_orient_mp_FF         src/basic/seq_reader.c:506-510
//   seq_read_reverse_complement     libs/seq_file/seq_file.h:758-777
//   seq_contig_start(r, 0, ...)     src/basic/seq_reader.c:61-117   (first contig only)
//
// The reference keeps two bits per k-mer ("a read started here", forward / reverse) and asks, read by
// read: does every mate that has a k-mer start on a bit that is already set?  Then the read (pair) is a
// duplicate and is not loaded; otherwise its bits are set.  After ANY read has been through that test
// its bits are set (they were before, or they are now), so bit (node, orient) is set when read i arrives
// iff some read j < i starts there.  That removes the order dependence:
//     first[node, orient] = min { i : read (pair) i has a mate starting there }
//     duplicate(i)        = for every mate m of i that has a k-mer: first[start(m)] < i
// which is one atomicMin pass over the reads followed by one compare pass.  It reproduces the reference
// run with one worker thread (reads in file order); with several threads the reference's outcome
// depends on scheduling.
#pragma once
#include "mcx_device.cuh"

#define MCX_PCR_NONE 0xFFFFFFFFFFFFFFFFull   /* node of a read without a k-mer */
#define MCX_PCR_UNSET 0xFFFFFFFFu            /* first[] of a start nobody has used */
/* mcx_graph_add_reads_pcr mate[] bytes */
#define MCX_MATE_SINGLE 0u
#define MCX_MATE_FIRST  1u   /* the next read is its mate */
#define MCX_MATE_SECOND 2u
#define MCX_MATE_KIND   3u
#define MCX_MATE_REVCOMP 4u  /* reverse-complement the read (and reverse its qualities) before anything else */

MCX_HD bool mcx_is_acgt(uint32_t c)
{
  uint32_t u = c & 0xDFu;
  return u == 0x41u || u == 0x43u || u == 0x47u || u == 0x54u;
}

// only ACGTacgt change (seq_file.h:715-723); case is kept
MCX_HD uint8_t mcx_complement_char(uint8_t c)
{
  uint32_t u = c & 0xDFu, lower = c & 0x20u;
  uint32_t r = u == 0x41u ? 0x54u : u == 0x54u ? 0x41u : u == 0x43u ? 0x47u : u == 0x47u ? 0x43u : 0u;
  return r ? (uint8_t)(r | lower) : c;
}

// lanes lane, lane + nlanes, ... of one read: swap ends, complementing the bases; qualities (parallel
// bytes, may be NULL) are only reversed
MCX_HD void mcx_pcr_revcomp_lanes(uint8_t *seq, uint8_t *qual, uint64_t len, uint32_t lane, uint32_t nlanes)
{
  for(uint64_t i = lane; i < len / 2; i += nlanes) {
    const uint64_t j = len - 1 - i;
    const uint8_t a = seq[i], b = seq[j];
    seq[i] = mcx_complement_char(b); seq[j] = mcx_complement_char(a);
    if(qual) { const uint8_t qa = qual[i]; qual[i] = qual[j]; qual[j] = qa; }
  }
  if((len & 1) && lane == 0) seq[len / 2] = mcx_complement_char(seq[len / 2]);
}

// Start of the first contig of a read: the smallest p with [p, p+k) all ACGT, every quality > qcut
// (qcut != 0 and qual != NULL; signed compare like the reference's char) and no hp consecutive equal
// bytes (hp != 0; raw compare).  seq_contig_start2 reaches the same p by jumping past the last offender
// of each window it tries.  Returns len when there is none.
MCX_HD uint64_t mcx_first_contig_start(const uint8_t *seq, const uint8_t *qual, uint64_t len, uint32_t k, uint32_t qcut, uint32_t hp)
{
  uint64_t from = 0; uint32_t run = 1; uint32_t prev = 0x100u;
  for(uint64_t i = 0; i < len; i++) {
    const uint32_t c = seq[i];
    bool ok = mcx_is_acgt(c);
    if(qcut && qual) ok = ok && ((int)(int8_t)qual[i] > (int)qcut);
    if(!ok) from = i + 1;
    run = (c == prev) ? run + 1u : 1u;
    prev = c;
    if(hp && run >= hp && i + 2u - hp > from) from = i + 2u - hp;
    if(i + 1u >= from + k) return from;
  }
  return len;
}

// binary_kmer_from_str (src/basic/binary_kmer.c:156-186) over k bytes known to be ACGTacgt
template <int W> MCX_HD McxKmer<W> mcx_kmer_from_ascii(const uint8_t *s, uint32_t k)
{
  McxKmer<W> f;
#pragma unroll
  for(int w = 0; w < W; w++) f.b[w] = 0;
  for(uint32_t i = 0; i < k; i++) {
    const uint32_t c = s[i], code = ((c >> 1) ^ (c >> 2)) & 3u;
#pragma unroll
    for(int w = 0; w + 1 < W; w++) f.b[w] = (f.b[w] << 2) | (f.b[w + 1] >> 62);
    f.b[W - 1] = (f.b[W - 1] << 2) | code;
  }
  return f;
}

// ordinal of the read (pair) that read r belongs to: the batch index of its first mate
MCX_HD uint64_t mcx_pcr_leader(uint64_t r, uint32_t mate) { return (mate & MCX_MATE_KIND) == MCX_MATE_SECOND ? r - 1 : r; }

// the compare pass for read r: node[] and first[] as written by the mark pass of this batch (and
// first[] by all earlier batches of the colour)
MCX_HD bool mcx_pcr_is_dup(uint64_t r, const uint8_t *mate, const uint64_t *node, const uint32_t *first, uint32_t ord_base)
{
  const uint64_t l = mcx_pcr_leader(r, mate[r]);
  const uint32_t ord = ord_base + (uint32_t)l;
  bool dup = node[l] == MCX_PCR_NONE || first[node[l]] < ord;
  if((mate[l] & MCX_MATE_KIND) == MCX_MATE_FIRST) dup = dup && (node[l + 1] == MCX_PCR_NONE || first[node[l + 1]] < ord);
  return dup;
}
