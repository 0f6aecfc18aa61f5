// mcx_table.cuh -- device-resident dBGraph table: find-or-insert + coverage + edges.
//
// Replaces (reference, relative to /root/reference):
//   hash_table_find_or_insert_mt   src/graph/hash_table.c:250-281
//   db_node_increment_coverage_mt  src/graph/db_node.c:139-144
//   db_node_set_col_edge_mt        src/graph/db_node.h:273-274
// The reference takes a per-bucket bit-spinlock, scans a <=48-slot bucket and
// rehashes up to 20 times; here the key slot itself is the lock: one
// atomicCAS (64-bit for k<=31, 128-bit for k<=63) claims an empty slot, probing
// is linear over 32-byte sectors, and coverage/edges live in the same slot so a
// hit costs one sector read + one RED.
//
// Slot layout (u32 words, stride = mcx_slot_words(W,C), 16-byte aligned):
//   [0 .. 2W)            key words, word 0 of the u64 b[0] carries MCX_KEY_FLAG
//   [2W .. 2W+C)         covg[c]
//   [2W+C .. )           edges bytes, colour c at byte c
#pragma once
#include "mcx_device.cuh"

struct McxTable {
  uint32_t *slots;      // nslots * stride u32
  uint64_t nslots;      // always even
  uint32_t stride;      // u32 words per slot
  uint32_t ncols;
  // L2-resident front table (k <= 31, one colour): a dense write-combining cache in front of
  // the big table.  The big table is >> L2 and a hot k-mer there drags a whole 128-byte L2 line
  // for 16 useful bytes, so on high-coverage input the hot set does not fit L2 (ncu, first
  // kernel: 127 B of DRAM traffic per occurrence, L2 hit rate 30 %).  Front slots are 8 bytes,
  //   [ count : 64-8-T bits | edges : 8 | tag : T = 62 - front_set_bits ]
  // four per 32-byte sector = one set; set index and tag are the two halves of the bijective
  // mcx_phi(key), so the key is implied exactly.  A key claims a way if one is free (first
  // come, never evicted); its occurrences are then counted here at L2 speed.  Counts that fill
  // 1/8 of the count field are moved to the big table on the fly (atomicAnd returns and clears
  // the count field), everything else by mcx_front_flush_kernel at sync.  Sums and ORs
  // commute and every record is merged exactly once, so the final table is identical.
  unsigned long long *front;   // (4 << front_set_bits) slots, or nullptr
  uint32_t front_set_bits;     // log2(number of sets); 0 = no front table
};

#if defined(__CUDACC__)

// 32-byte probe load (SASS: LDG.E.256).  .ca / .cg / volatile / L1::no_allocate flavours and
// two 16-byte loads were measured (profiles/r1_exp_ld_modes.txt): no difference except that two
// 16-byte loads are 45 % slower, so the plain form stays.
__device__ __forceinline__ void mcx_ld256(const void *p, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d)
{
  asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}

__device__ __forceinline__ void mcx_ld128(const void *p, uint64_t &a, uint64_t &b)
{
  asm volatile("ld.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
}

// 128-bit compare-and-swap (sm_90+): returns the previous 16 bytes
__device__ __forceinline__ void mcx_cas128(void *p, uint64_t cmp_lo, uint64_t cmp_hi, uint64_t new_lo, uint64_t new_hi,
                                           uint64_t &old_lo, uint64_t &old_hi)
{
  asm volatile("{\n\t.reg .b128 c, s, r;\n\t"
               "mov.b128 c, {%2, %3};\n\t"
               "mov.b128 s, {%4, %5};\n\t"
               "atom.global.cas.b128 r, [%6], c, s;\n\t"
               "mov.b128 {%0, %1}, r;\n\t}"
               : "=l"(old_lo), "=l"(old_hi)
               : "l"(cmp_lo), "l"(cmp_hi), "l"(new_lo), "l"(new_hi), "l"(p)
               : "memory");
}

// saturating coverage add (reference: CAS loop capped at COVG_MAX, db_node.c:139-144).
// A plain RED is exact unless the counter is within reach of 2^32.  `may_saturate` is a
// launch-uniform flag the host clears while fewer than ~4e9 occurrences were ever sent to the
// graph (then no counter can be near the cap).  When it is set, a snapshot of the counter that
// came with the probe load (possibly stale by the few 1e5 increments in flight) still proves a
// RED safe if it is below 0xF0000000 - n; otherwise fall back to the reference's CAS loop.
__device__ __forceinline__ void mcx_covg_add(uint32_t *cv, uint32_t n, bool may_saturate, bool have_snap = false,
                                             uint32_t snap = 0)
{
  if(n == 0) return;
  if(!may_saturate || (have_snap && snap < 0xF0000000u - n)) { atomicAdd(cv, n); return; }
  uint32_t v = *(volatile uint32_t *)cv;
  if(v < 0xF0000000u - n) { atomicAdd(cv, n); return; }
  while(v != 0xFFFFFFFFu) {
    uint32_t nv = (v + n < v) ? 0xFFFFFFFFu : v + n;
    uint32_t old = atomicCAS(cv, v, nv);
    if(old == v) break;
    v = old;
  }
}

__device__ __forceinline__ void mcx_edges_or(uint32_t *slot, uint32_t W, uint32_t C, uint32_t colour, uint32_t emask,
                                             uint32_t known_word, bool known)
{
  if(!emask) return;
  uint32_t *e = slot + 2u * W + C + (colour >> 2);
  uint32_t bits = emask << ((colour & 3u) * 8u);
  uint32_t cur = known ? known_word : *(volatile uint32_t *)e; // possibly stale: only ever misses bits => extra OR
  if((cur & bits) != bits) atomicOr(e, bits);
}

// find-or-insert in the big table, then covg[colour] += n (saturating) and edges |= emask.
// Returns 0 = found, 1 = novel, 2 = table full.
template <int W>
__device__ __forceinline__ int mcx_table_add(const McxTable &t, const McxKmer<W> &key, uint32_t hc, uint32_t hb,
                                             uint32_t colour, uint32_t emask, uint32_t n, bool may_saturate);

// ---- k <= 31 --------------------------------------------------------------
template <>
__device__ __forceinline__ int mcx_table_add<1>(const McxTable &t, const McxKmer<1> &key, uint32_t hc, uint32_t hb,
                                                uint32_t colour, uint32_t emask, uint32_t n, bool may_saturate)
{
  const uint64_t keyf = key.b[0] | MCX_KEY_FLAG;
  uint64_t idx = mcx_home_slot(hc, hb, t.nslots);
  int novel = 0;
  if(t.stride == 4u) {
    // 16-byte slots {key, covg, edges}: probe one 32-byte sector (= 2 slots) per load
    idx &= ~1ull;
    for(uint64_t probes = 0; probes < t.nslots; probes += 2) {
      uint32_t *s = t.slots + idx * 4u;
      uint64_t k0, m0, k1, m1;
      mcx_ld256(s, k0, m0, k1, m1);
      uint32_t *hit = nullptr; uint64_t meta = 0;
      if(k0 == keyf) { hit = s; meta = m0; }
      else if(k1 == keyf) { hit = s + 4; meta = m1; }
      else if(k0 == 0 || k1 == 0) {
        // first empty slot of the sector (slot 0 before slot 1 keeps the probe order total)
        uint32_t *cand = (k0 == 0) ? s : s + 4;
        uint64_t old = atomicCAS((unsigned long long *)cand, 0ull, (unsigned long long)keyf);
        if(old == 0) { hit = cand; novel = 1; }
        else if(old == keyf) { hit = cand; }
        else if(cand == s) {
          // lost slot 0 to another key: slot 1 of the same sector is next in probe order
          uint64_t o1 = (k1 == 0) ? atomicCAS((unsigned long long *)(s + 4), 0ull, (unsigned long long)keyf) : k1;
          if(o1 == 0) { hit = s + 4; novel = 1; }
          else if(o1 == keyf) { hit = s + 4; }
        }
        meta = 0; // freshly claimed or raced: treat edges as unknown-empty => OR is issued
      }
      if(hit) {
        mcx_covg_add(hit + 2, n, may_saturate, true, (uint32_t)meta);
        mcx_edges_or(hit, 1, 1, 0, emask, (uint32_t)(meta >> 32), true);
        return novel;
      }
      idx += 2; if(idx >= t.nslots) idx = 0;
    }
    return 2;
  }
  // generic stride (C > 1)
  for(uint64_t probes = 0; probes < t.nslots; probes++) {
    uint32_t *s = t.slots + idx * (uint64_t)t.stride;
    uint64_t cur = *(volatile uint64_t *)s;
    if(cur == 0) {
      cur = atomicCAS((unsigned long long *)s, 0ull, (unsigned long long)keyf);
      if(cur == 0) { novel = 1; cur = keyf; }
    }
    if(cur == keyf) {
      mcx_covg_add(s + 2 + colour, n, may_saturate);
      mcx_edges_or(s, 1, t.ncols, colour, emask, 0, false);
      return novel;
    }
    idx++; if(idx >= t.nslots) idx = 0;
  }
  return 2;
}

// ---- 33 <= k <= 63 ---------------------------------------------------------
template <>
__device__ __forceinline__ int mcx_table_add<2>(const McxTable &t, const McxKmer<2> &key, uint32_t hc, uint32_t hb,
                                                uint32_t colour, uint32_t emask, uint32_t n, bool may_saturate)
{
  const uint64_t k0f = key.b[0] | MCX_KEY_FLAG, k1 = key.b[1];
  uint64_t idx = mcx_home_slot(hc, hb, t.nslots);
  int novel = 0;
  for(uint64_t probes = 0; probes < t.nslots; probes++) {
    uint32_t *s = t.slots + idx * (uint64_t)t.stride;
    uint64_t c0, c1, m0 = 0, m1 = 0;
    bool have_meta = (t.stride == 8u);
    if(have_meta) mcx_ld256(s, c0, c1, m0, m1); // 32-byte slot: key + covg + edges in one sector
    else mcx_ld128(s, c0, c1); // one 16-byte transaction: never a torn view of a 128-bit CAS
    if(c0 == 0) {
      mcx_cas128(s, 0ull, 0ull, k0f, k1, c0, c1);
      if(c0 == 0) { novel = 1; c0 = k0f; c1 = k1; }
      have_meta = false;
    }
    if(c0 == k0f && c1 == k1) {
      mcx_covg_add(s + 4 + colour, n, may_saturate);
      // edges word for colour c sits at u32 index 4 + C + (c>>2); with the 256-bit
      // load we hold u32 words 4..7 in (m0, m1)
      uint32_t ew_idx = 4u + t.ncols + (colour >> 2);
      bool known = have_meta && ew_idx < 8u;
      uint32_t ew = 0;
      if(known) { uint64_t m = (ew_idx < 6u) ? m0 : m1; ew = (uint32_t)(m >> (32u * (ew_idx & 1u))); }
      mcx_edges_or(s, 2, t.ncols, colour, emask, ew, known);
      return novel;
    }
    idx++; if(idx >= t.nslots) idx = 0;
  }
  return 2;
}

// ---- front table ------------------------------------------------------------
struct McxFrontGeom { uint32_t T; uint64_t tagmask, one; };
__device__ __forceinline__ McxFrontGeom mcx_front_geom(const McxTable &t)
{
  McxFrontGeom g; g.T = 62u - t.front_set_bits; g.tagmask = (1ull << g.T) - 1ull; g.one = 1ull << (g.T + 8u);
  return g;
}

// Slow side of the front table for one occurrence of `key` (k <= 31): re-reads the set, claims a
// free way if the key is not there, drains a count field that is 1/8 full.  Returns true if the
// Returns 0xFFFFFFFF if the front table did NOT absorb the occurrence; otherwise the count that
// this thread has just taken OUT of the front table and must add to the big table (0 almost always).
static __device__ __noinline__ uint32_t mcx_front_add_slow(McxTable t, uint64_t key, uint32_t emask)
{
  const McxFrontGeom g = mcx_front_geom(t);
  const uint64_t y = mcx_phi(key), tag = y & g.tagmask, ebits = (uint64_t)emask << g.T;
  unsigned long long *set = t.front + ((y >> g.T) << 2);
  uint64_t v0, v1, v2, v3;
  mcx_ld256(set, v0, v1, v2, v3);
  uint32_t drained = 0;
#pragma unroll
  for(int w = 0; w < 4; w++) {
    const uint64_t v = w == 0 ? v0 : (w == 1 ? v1 : (w == 2 ? v2 : v3));
    if(v != 0 && (v & g.tagmask) == tag) {
      atomicAdd(&set[w], (unsigned long long)g.one);
      if((v & ebits) != ebits) atomicOr(&set[w], (unsigned long long)ebits);
      if((v >> (g.T + 8u)) >= (1ull << (53u - g.T))) {
        // the count field is 56-T >= 12 bits wide (15 at the default size): move it to the big
        // table once it is 1/8 full, long before it can wrap (a wrap would need 7/8 of the range,
        // > 28k increments of ONE address at the default size, inside one load->RED latency; the
        // L2 atomic unit retires ~1 per clock per address)
        uint64_t old = atomicAnd(&set[w], (unsigned long long)(g.one - 1ull));
        drained = (uint32_t)(old >> (g.T + 8u));
      }
      return drained;
    }
  }
#pragma unroll
  for(int w = 0; w < 4; w++) {
    const uint64_t v = w == 0 ? v0 : (w == 1 ? v1 : (w == 2 ? v2 : v3));
    if(v == 0) {
      uint64_t old = atomicCAS(&set[w], 0ull, (unsigned long long)(tag | ebits | g.one));
      if(old == 0) return 0u;                              // claimed, first count and edges included
      if((old & g.tagmask) == tag) {                       // lost the race to the same key
        atomicAdd(&set[w], (unsigned long long)g.one);
        if((old & ebits) != ebits) atomicOr(&set[w], (unsigned long long)ebits);
        return 0u;
      }
    }
  }
  return 0xFFFFFFFFu;
}

// Fast side: the set has already been loaded (v0..v3).  Handles the overwhelmingly common case --
// the key sits in the set, its edge bits are already there, its count field is far from full --
// with ONE RED and returns true; anything else returns false and goes to mcx_front_add_slow.
__device__ __forceinline__ bool mcx_front_hit(const McxFrontGeom &g, unsigned long long *set, uint64_t tag, uint64_t ebits,
                                              uint64_t v0, uint64_t v1, uint64_t v2, uint64_t v3)
{
  const uint64_t lim = 1ull << (53u - g.T);
  int w = -1; uint64_t v = 0;
  if((v0 & g.tagmask) == tag && v0 != 0) { w = 0; v = v0; }
  else if((v1 & g.tagmask) == tag && v1 != 0) { w = 1; v = v1; }
  else if((v2 & g.tagmask) == tag && v2 != 0) { w = 2; v = v2; }
  else if((v3 & g.tagmask) == tag && v3 != 0) { w = 3; v = v3; }
  if(w < 0 || (v & ebits) != ebits || (v >> (g.T + 8u)) >= lim) return false;
  // count and edge fields sit entirely in the high 32-bit word (T + 8 >= 32 for every allowed
  // size), so the increment is a 32-bit RED like the big table's covg++ (64-bit REDs measured
  // ~30 % slower here)
  atomicAdd(reinterpret_cast<unsigned int *>(&set[w]) + 1, (unsigned int)(g.one >> 32));
  return true;
}

#endif // __CUDACC__
