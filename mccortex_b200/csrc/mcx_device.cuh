// mcx_device.cuh -- per-base / per-window integer math of the build hot path.
//
// Everything here is MCX_HD (__host__ __device__) and free of CUDA-only
// intrinsics so that tests/emul can execute the very same code on the CPU and
// compare it, window by window, with the oracle (this container has no GPU).
// The kernels in mcx_build.cu call these functions on shared-memory arrays.
//
// Encoding conventions (chosen for the GPU, not the reference's):
//   * packed bases: u32 word j holds bases 16j..16j+15, base 16j in the TOP two
//     bits, so a k-mer is a big-endian bit string and "first base = most
//     significant bits" exactly like BinaryKmer (reference
//     src/graph/binary_kmer.h:7,39-49).
//   * bit masks (bad / eq / valid): bit i of the little-endian bit string
//     (word i>>5, bit i&31) belongs to base / window i.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MCX_HD __host__ __device__ __forceinline__
#else
#define MCX_HD inline
#endif

#define MCX_KEY_FLAG (1ULL << 63) /* "slot assigned" bit, same role as BKMER_SET_FLAG hash_table.h:14-15 */

// ---------------------------------------------------------------------------
// Row A (src/basic/dna.c:8-25): ASCII -> 2-bit code, four bases per u32.
// A/a=0 C/c=1 G/g=2 T/t=3 via ((c>>1)^(c>>2))&3; validity = letter in ACGTacgt.
// ---------------------------------------------------------------------------
MCX_HD uint32_t mcx_bytes_eq4(uint32_t a, uint32_t b)
{
  // 0x80 in every byte lane where a == b (classic zero-byte test on a^b)
  uint32_t x = a ^ b;
  uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
  return ~(t | x | 0x7F7F7F7Fu);
}

// 4 ASCII bytes (little-endian: byte 0 = first base) -> 8 packed bits,
// first base in bits 7:6.
MCX_HD uint32_t mcx_pack4(uint32_t w)
{
  uint32_t c = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
  return (c * 0x40100401u) >> 24; // 2^30 + 2^20 + 2^10 + 1 : gathers the four 2-bit fields, no carries
}

// 4 ASCII bytes -> 4 bits, bit i set iff byte i is NOT one of ACGTacgt.
MCX_HD uint32_t mcx_bad4(uint32_t w)
{
  uint32_t u = w & 0xDFDFDFDFu; // fold case
  uint32_t ok = mcx_bytes_eq4(u, 0x41414141u) | mcx_bytes_eq4(u, 0x43434343u) |
                mcx_bytes_eq4(u, 0x47474747u) | mcx_bytes_eq4(u, 0x54545454u);
  uint32_t x = (~ok >> 7) & 0x01010101u;
  return ((x * 0x01020408u) >> 24) & 0xFu; // byte i -> bit i
}

// 4 bytes -> 4 bits, bit i set iff byte i == byte i-1 (prev = byte before byte 0).
// Raw byte compare, case sensitive, exactly like seq[i-1]==seq[i] in
// src/basic/seq_reader.c:97,157.
MCX_HD uint32_t mcx_eqprev4(uint32_t w, uint32_t prev)
{
  uint32_t sh = (w << 8) | (prev & 0xFFu);
  uint32_t x = (mcx_bytes_eq4(w, sh) >> 7) & 0x01010101u;
  return ((x * 0x01020408u) >> 24) & 0xFu;
}

// 4 bytes -> 4 bits, bit i set iff byte i == '\n' (read terminator of the LINES layout)
MCX_HD uint32_t mcx_nl4(uint32_t w)
{
  uint32_t x = (mcx_bytes_eq4(w, 0x0A0A0A0Au) >> 7) & 0x01010101u;
  return ((x * 0x01020408u) >> 24) & 0xFu;
}

// ---------------------------------------------------------------------------
// Bit-string access
// ---------------------------------------------------------------------------
// 32 bases starting at base p, base p in the top two bits. Reads words p>>4 .. (p>>4)+2.
MCX_HD uint64_t mcx_get32bases(const uint32_t *pk, uint32_t p)
{
  uint32_t j = p >> 4, o = (p & 15u) * 2u;
  uint64_t hi = ((uint64_t)pk[j] << 32) | pk[j + 1];
  uint64_t lo = ((uint64_t)pk[j + 2] << o) >> 32;
  return (hi << o) | lo;
}

MCX_HD uint32_t mcx_get_base(const uint32_t *pk, uint32_t p)
{
  return (pk[p >> 4] >> (30u - 2u * (p & 15u))) & 3u;
}

// 64 mask bits starting at bit p (bit p -> result bit 0). Reads words p>>5 .. (p>>5)+2.
MCX_HD uint64_t mcx_get64bits(const uint32_t *m, uint32_t p)
{
  uint32_t j = p >> 5, o = p & 31u;
  uint64_t lo = ((uint64_t)m[j + 1] << 32) | m[j];
  uint64_t hi = (uint64_t)m[j + 2];
  return o ? ((lo >> o) | (hi << (64u - o))) : lo;
}

MCX_HD uint32_t mcx_get_bit(const uint32_t *m, uint32_t p) { return (m[p >> 5] >> (p & 31u)) & 1u; }

// true iff x contains a run of >= n consecutive 1 bits (n >= 1)
MCX_HD bool mcx_has_run(uint64_t x, uint32_t n)
{
  // shift-and doubling: after processing, bit i set iff bits i..i+n-1 were all set
  uint32_t have = 1;
  while(have < n && x) {
    uint32_t s = (n - have < have) ? (n - have) : have;
    x &= x >> s;
    have += s;
  }
  return x != 0;
}

// ---------------------------------------------------------------------------
// Rows C/D: k-mer words, reverse complement, canonical key
// ---------------------------------------------------------------------------
template <int W> struct McxKmer { uint64_t b[W]; }; // b[0] most significant, like BinaryKmer

// reverse the order of the 32 two-bit fields of x and complement them
MCX_HD uint64_t mcx_revcomp64(uint64_t x)
{
  x = ~x;
#if defined(__CUDA_ARCH__)
  x = __brevll(x);
#else
  x = ((x >> 32) | (x << 32));
  x = ((x & 0xFFFF0000FFFF0000ull) >> 16) | ((x & 0x0000FFFF0000FFFFull) << 16);
  x = ((x & 0xFF00FF00FF00FF00ull) >> 8) | ((x & 0x00FF00FF00FF00FFull) << 8);
  x = ((x & 0xF0F0F0F0F0F0F0F0ull) >> 4) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
  x = ((x & 0xCCCCCCCCCCCCCCCCull) >> 2) | ((x & 0x3333333333333333ull) << 2);
  x = ((x & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((x & 0x5555555555555555ull) << 1);
#endif
  // full bit reversal also swapped the two bits inside every field: swap back
  return ((x & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((x & 0x5555555555555555ull) << 1);
}

// k-mer starting at base p of the packed array (reference: binary_kmer_from_str
// binary_kmer.c:156-186 gives the same words)
template <int W> MCX_HD McxKmer<W> mcx_kmer_at(const uint32_t *pk, uint32_t p, uint32_t k);
template <> MCX_HD McxKmer<1> mcx_kmer_at<1>(const uint32_t *pk, uint32_t p, uint32_t k)
{
  McxKmer<1> r; r.b[0] = mcx_get32bases(pk, p) >> (64u - 2u * k); return r;
}
template <> MCX_HD McxKmer<2> mcx_kmer_at<2>(const uint32_t *pk, uint32_t p, uint32_t k)
{
  McxKmer<2> r;
  r.b[0] = mcx_get32bases(pk, p) >> (128u - 2u * k); // first k-32 bases, right aligned (k in 33..63)
  r.b[1] = mcx_get32bases(pk, p + k - 32u);          // last 32 bases
  return r;
}

// reference: binary_kmer_reverse_complement binary_kmer.c:102-133
template <int W> MCX_HD McxKmer<W> mcx_kmer_revcomp(const McxKmer<W> &f, uint32_t k);
template <> MCX_HD McxKmer<1> mcx_kmer_revcomp<1>(const McxKmer<1> &f, uint32_t k)
{
  McxKmer<1> r; r.b[0] = mcx_revcomp64(f.b[0]) >> (64u - 2u * k); return r;
}
template <> MCX_HD McxKmer<2> mcx_kmer_revcomp<2>(const McxKmer<2> &f, uint32_t k)
{
  McxKmer<2> r;
  uint32_t s = 128u - 2u * k; // 2..62
  uint64_t hi = mcx_revcomp64(f.b[1]), lo = mcx_revcomp64(f.b[0]);
  r.b[0] = hi >> s;
  r.b[1] = (hi << (64u - s)) | (lo >> s);
  return r;
}

// reference: binary_kmer_get_key binary_kmer.c:43-57; orient = 0 FORWARD / 1 REVERSE (db_node.h:109-110)
template <int W> MCX_HD McxKmer<W> mcx_kmer_key(const McxKmer<W> &f, uint32_t k, uint32_t *orient)
{
  McxKmer<W> rc = mcx_kmer_revcomp<W>(f, k);
  bool rc_lt;
  if(W == 1) rc_lt = rc.b[0] < f.b[0];
  else rc_lt = (rc.b[0] < f.b[0]) || (rc.b[0] == f.b[0] && rc.b[W - 1] < f.b[W - 1]);
  *orient = rc_lt ? 1u : 0u;
  return rc_lt ? rc : f;
}

// ---------------------------------------------------------------------------
// Row E: Lookup3 hashlittle specialised to 8*W key bytes
// (reference src/kmer/kmer_hash.h:89-97,124-133,162-211).  Returns c; *b2 gets
// the second lane (what lookup3's hashlittle2 calls *pb).
// ---------------------------------------------------------------------------
MCX_HD uint32_t mcx_rot(uint32_t x, uint32_t k) { return (x << k) | (x >> (32u - k)); }

template <int W> MCX_HD uint32_t mcx_lookup3(const McxKmer<W> &key, uint32_t initval, uint32_t *b2)
{
  uint32_t a, b, c;
  a = b = c = 0xdeadbeefu + (uint32_t)(8 * W) + initval;
  a += (uint32_t)key.b[0];
  b += (uint32_t)(key.b[0] >> 32);
  if(W == 2) {
    c += (uint32_t)key.b[W - 1];
    a -= c;  a ^= mcx_rot(c, 4);  c += b;
    b -= a;  b ^= mcx_rot(a, 6);  a += c;
    c -= b;  c ^= mcx_rot(b, 8);  b += a;
    a -= c;  a ^= mcx_rot(c,16);  c += b;
    b -= a;  b ^= mcx_rot(a,19);  a += c;
    c -= b;  c ^= mcx_rot(b, 4);  b += a;
    a += (uint32_t)(key.b[W - 1] >> 32);
  }
  c ^= b; c -= mcx_rot(b,14);
  a ^= c; a -= mcx_rot(c,11);
  b ^= a; b -= mcx_rot(a,25);
  c ^= b; c -= mcx_rot(b,16);
  a ^= c; a -= mcx_rot(c,4);
  b ^= a; b -= mcx_rot(a,14);
  c ^= b; c -= mcx_rot(b,24);
  *b2 = b;
  return c;
}

// ---------------------------------------------------------------------------
// Row B, local form (no quality cut-off): window p is loadable iff its k bases
// are all ACGT and (hp_cutoff>0) it contains no run of >= hp_cutoff equal
// characters.  See DESIGN.md "contig rules as window predicates" for the proof
// that maximal runs of such windows are exactly the contigs that
// seq_contig_start2 / seq_contig_end2 (seq_reader.c:61-172) produce when
// hp_cutoff <= k.
//   bad : bit i = base i is not ACGT (read terminators are non-ACGT, so a window
//         can never span two reads)
//   eq  : bit i = byte i equals byte i-1
// ---------------------------------------------------------------------------
MCX_HD bool mcx_window_ok(const uint32_t *bad, const uint32_t *eq, uint32_t p, uint32_t k, uint32_t hp_cutoff)
{
  uint64_t kmask = (k >= 64u) ? ~0ull : ((1ull << k) - 1ull);
  if(mcx_get64bits(bad, p) & kmask) return false;
  if(hp_cutoff > 1u) {
    // a run of hp equal chars inside [p,p+k) == hp-1 consecutive eq bits inside [p+1,p+k)
    uint64_t e = mcx_get64bits(eq, p + 1u) & (kmask >> 1);
    if(mcx_has_run(e, hp_cutoff - 1u)) return false;
  }
  return true;
}

// ---------------------------------------------------------------------------
// Row G, local form: the edge bits one occurrence ORs into its own node
// (reference db_graph_add_edge_mt db_graph.c:152-166, nuc_orient_to_edge
// db_node.h:180): next base on the strand read, previous base complemented on
// the other strand.
// ---------------------------------------------------------------------------
MCX_HD uint32_t mcx_edge_mask(uint32_t orient, bool has_prev, uint32_t prev_nuc, bool has_next, uint32_t next_nuc)
{
  uint32_t m = 0;
  if(has_next) m |= 1u << (next_nuc + 4u * orient);
  if(has_prev) m |= 1u << ((~prev_nuc & 3u) + 4u * (orient ^ 1u));
  return m;
}

// ---------------------------------------------------------------------------
// Table geometry (device table layout is ours; only the sorted dump is contract)
//   slot = [ key: W x u64 | covg: C x u32 | edges: C x u8 padded to u32 ] padded to 16 B
// ---------------------------------------------------------------------------
MCX_HD uint32_t mcx_slot_words(uint32_t W, uint32_t C)
{
  uint32_t w = 2u * W + C + (C + 3u) / 4u;
  return (w + 3u) & ~3u;
}

// hash -> first slot of the probe sequence.  `b` (top) and `c` are both
// well-mixed lanes of lookup3; multi-GPU ownership uses the top bits of `c`
// (mcx_owner), so slot choice is driven by `b` first to stay independent of it.
MCX_HD uint64_t mcx_mulhi64(uint64_t a, uint64_t b)
{
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
MCX_HD uint64_t mcx_home_slot(uint32_t c, uint32_t b, uint64_t nslots)
{
  return mcx_mulhi64(((uint64_t)b << 32) | c, nslots);
}
MCX_HD uint32_t mcx_owner(uint32_t c, uint32_t nparts) { return (uint32_t)(((uint64_t)c * nparts) >> 32); }

// ---------------------------------------------------------------------------
// Front table (k <= 31, one colour).  TWO regions, because of what the L2 does (measured with
// scripts/probe_bench*.cu on a B200, 64 MB footprint, random sets): 32-byte loads alone run at
// 265 G/s, 32-bit REDs alone at 190 G/s, but a load followed by a RED into the SAME region only
// at 60 G/s -- any mix of reads and writes over one footprint costs 3x.  With the tags in a
// region that is (almost) only read and the counters in a region that is only written, the same
// pair runs at ~100 G/s:
//   tags     : 8-byte slots, four per 32-byte sector (= one set), read by the probe load;
//              written when a key claims a way and when an edge bit appears (a handful of times
//              per k-mer)
//   counters : one u32 per slot, only ever the target of RED.ADD
// A key (<= 62 bits = a 30-bit high word kh and a 32-bit low word kl) is sent through a
// BIJECTION built from three Feistel-style steps that use only 32-bit multiplies,
//     y1 = kh ^ ((kl * C1) >> 2);   x = kl ^ (y1 * C2);   y = y1 ^ ((x * C3) >> 2);
// (each step xors one half with a function of the other, so it is its own inverse given the
// other half).  set = low S bits of y, tag = (x, y >> S): together they identify the key exactly,
// so the key itself is not stored -- twice as many hot k-mers per L2 byte as key-carrying slots.
//   tag.lo = x
//   tag.hi = [ 0 ... | edges : 8 | occupied : 1 | y >> S : 30-S bits ]
// An empty slot is all zero; the occupied bit makes a tag compare against an empty slot fail
// without a separate test.
// ---------------------------------------------------------------------------
#define MCX_FH_C1 0x9E3779B1u   /* odd */
#define MCX_FH_C2 0x85EBCA6Bu   /* odd */
#define MCX_FH_C3 0xC2B2AE35u   /* odd */
struct McxFKey { uint32_t x, y; };
MCX_HD McxFKey mcx_fhash(uint64_t key)
{
  const uint32_t kl = (uint32_t)key, kh = (uint32_t)(key >> 32);
  McxFKey r;
  uint32_t y = kh ^ ((kl * MCX_FH_C1) >> 2);
  r.x = kl ^ (y * MCX_FH_C2);
  r.y = y ^ ((r.x * MCX_FH_C3) >> 2);
  return r;
}
MCX_HD uint64_t mcx_fhash_inv(uint32_t x, uint32_t y)
{
  const uint32_t y1 = y ^ ((x * MCX_FH_C3) >> 2);
  const uint32_t kl = x ^ (y1 * MCX_FH_C2);
  const uint32_t kh = y1 ^ ((kl * MCX_FH_C1) >> 2);
  return ((uint64_t)kh << 32) | kl;
}
// ---- 33 <= k <= 63: the same table with 16-byte tags, two ways per 32-byte set -------------------------------------
// A key is a high word kh (<= 62 bits) and a low word kl (64 bits).  The same three Feistel steps, 64 bits wide, with
// m(v, C) = (v ^ (v >> 31)) * C as the round function:
//     y1 = kh ^ (m(kl, C1) >> 2);   x = kl ^ m(y1, C2);   y = y1 ^ (m(x, C3) >> 2);          (y: 62 bits)
// set = TOP S bits of y (the well-mixed end of the last product), tag = (x, low 62 - S bits of y):
//   tag.lo = x
//   tag.hi = [ displaced : 1 (bit 63) | 0 ... | edges : 8 | occupied : 1 | low 62 - S bits of y ]
// Counters: one u32 per way, in their own region, as for k <= 31.  An empty way is sixteen zero bytes.
#define MCX_FH2_C1 0x9E3779B97F4A7C15ull
#define MCX_FH2_C2 0xC2B2AE3D27D4EB4Full
#define MCX_FH2_C3 0xD6E8FEB86659FD93ull
struct McxFKey2 { uint64_t x, y; };
MCX_HD uint64_t mcx_fmix64(uint64_t v, uint64_t c) { return (v ^ (v >> 31)) * c; }
MCX_HD McxFKey2 mcx_fhash2(uint64_t kh, uint64_t kl)
{
  McxFKey2 r;
  const uint64_t y1 = kh ^ (mcx_fmix64(kl, MCX_FH2_C1) >> 2);
  r.x = kl ^ mcx_fmix64(y1, MCX_FH2_C2);
  r.y = y1 ^ (mcx_fmix64(r.x, MCX_FH2_C3) >> 2);
  return r;
}
MCX_HD void mcx_fhash2_inv(uint64_t x, uint64_t y, uint64_t *kh, uint64_t *kl)
{
  const uint64_t y1 = y ^ (mcx_fmix64(x, MCX_FH2_C3) >> 2);
  *kl = x ^ mcx_fmix64(y1, MCX_FH2_C2);
  *kh = y1 ^ (mcx_fmix64(*kl, MCX_FH2_C1) >> 2);
}
struct McxFrontGeom2 { uint32_t S, tshift, eshift; uint64_t occ, mask; }; // mask: tag bits + occupied + displaced
#define MCX_FRONT2_DISPLACED (1ull << 63)
MCX_HD McxFrontGeom2 mcx_front_geom2_bits(uint32_t S)
{
  McxFrontGeom2 g;
  g.S = S; g.tshift = 62u - S;
  g.occ = 1ull << g.tshift;
  g.mask = ((g.occ << 1) - 1ull) | MCX_FRONT2_DISPLACED;
  g.eshift = g.tshift + 1u;
  return g;
}

// geometry of the tag's hi word for S set bits (16 <= S <= 24)
struct McxFrontGeom { uint32_t S, setmask, occ, tagmask, eshift; };
MCX_HD McxFrontGeom mcx_front_geom_bits(uint32_t S)
{
  McxFrontGeom g;
  g.S = S; g.setmask = (1u << S) - 1u;
  g.occ = 1u << (30u - S);            // occupied flag, just above the tag bits
  g.tagmask = (g.occ << 1) - 1u;      // tag bits + occupied flag
  g.eshift = 31u - S;                 // edges field
  return g;
}
