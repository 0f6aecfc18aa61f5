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
  // L2-resident front table (16-byte slots {key, count, edges}, k <= 31, one colour only):
  // a dense write-combining cache in front of the big table.  The big table is >> L2 and a
  // hot k-mer there drags a whole 128-byte L2 line for 16 useful bytes, so on high-coverage
  // input the hot set does not fit L2 (ncu: 127 B of DRAM traffic per occurrence, L2 hit
  // rate 30 %).  Keys that win a front slot (first come, 2-way per 32-byte sector, no
  // probing beyond it) are counted here at L2 speed and merged into the big table by
  // mcx_front_flush_kernel; everything else falls through to the big table.  Sums and ORs
  // commute, so the final table is identical.
  uint32_t *front;      // front_nslots * 4 u32, or nullptr
  uint64_t front_nslots;
  uint32_t front_ways;  // 2 = one sector, 4 = a second sector chosen by an independent hash
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

// saturating coverage increment (reference: CAS loop capped at COVG_MAX, db_node.c:139-144).
// A plain RED is exact unless the counter is within reach of 2^32.  `may_saturate` is a
// launch-uniform flag the host clears while fewer than ~4e9 occurrences were ever sent to the
// graph (then no counter can be near the cap).  When it is set, a snapshot of the counter that
// came with the probe load (possibly stale by the few 1e5 increments in flight) still proves a
// RED safe if it is below 0xF0000000; otherwise fall back to the reference's CAS loop.
__device__ __forceinline__ void mcx_covg_inc(uint32_t *cv, bool may_saturate, bool have_snap = false, uint32_t snap = 0)
{
  if(!may_saturate || (have_snap && snap < 0xF0000000u)) { atomicAdd(cv, 1u); return; }
  uint32_t v = *(volatile uint32_t *)cv;
  if(v < 0xF0000000u) { atomicAdd(cv, 1u); return; }
  while(v != 0xFFFFFFFFu) {
    uint32_t old = atomicCAS(cv, v, v + 1u);
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

// find-or-insert + coverage + edges for one occurrence.
// Returns 0 = found, 1 = novel, 2 = table full.
template <int W>
__device__ __forceinline__ int mcx_table_add(const McxTable &t, const McxKmer<W> &key, uint32_t hc, uint32_t hb,
                                             uint32_t colour, uint32_t emask, bool may_saturate);

// ---- k <= 31 --------------------------------------------------------------
template <>
__device__ __forceinline__ int mcx_table_add<1>(const McxTable &t, const McxKmer<1> &key, uint32_t hc, uint32_t hb,
                                                uint32_t colour, uint32_t emask, bool may_saturate)
{
  const uint64_t keyf = key.b[0] | MCX_KEY_FLAG;
  uint64_t idx = mcx_home_slot(hc, hb, t.nslots);
  int novel = 0;
  if(t.stride == 4u) {
    if(t.front_nslots) {
      // front table: two ways per 32-byte sector, claim-if-empty, never evict; with
      // front_ways == 4 a second sector (independent hash) is tried when the first is taken
      uint64_t fi = mcx_mulhi64(((uint64_t)hc << 32) | hb, t.front_nslots) & ~1ull;
#pragma unroll 1
      for(uint32_t way = 0; way < t.front_ways; way += 2) {
        uint32_t *s = t.front + fi * 4u;
        uint64_t k0, m0, k1, m1;
        mcx_ld256(s, k0, m0, k1, m1);
        uint32_t *hit = nullptr; uint64_t meta = 0;
        if(k0 == keyf) { hit = s; meta = m0; }
        else if(k1 == keyf) { hit = s + 4; meta = m1; }
        else if(k0 == 0 || k1 == 0) {
          uint32_t *cand = (k0 == 0) ? s : s + 4;
          uint64_t old = atomicCAS((unsigned long long *)cand, 0ull, (unsigned long long)keyf);
          if(old == 0 || old == keyf) hit = cand;
          else if(cand == s && k1 == 0) {
            old = atomicCAS((unsigned long long *)(s + 4), 0ull, (unsigned long long)keyf);
            if(old == 0 || old == keyf) hit = s + 4;
          }
        }
        if(hit) {
          mcx_covg_inc(hit + 2, true, true, (uint32_t)meta); // front counters always guard against wrap
          mcx_edges_or(hit, 1, 1, 0, emask, (uint32_t)(meta >> 32), true);
          return 0; // novelty is decided when the entry is merged into the big table
        }
        fi = mcx_mulhi64((((uint64_t)hb * 0x9E3779B1u) << 32) ^ (((uint64_t)hc << 32) | hb) ^ hc, t.front_nslots) & ~1ull;
      }
    }
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
        mcx_covg_inc(hit + 2, may_saturate, true, (uint32_t)meta);
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
      mcx_covg_inc(s + 2 + colour, may_saturate);
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
                                                uint32_t colour, uint32_t emask, bool may_saturate)
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
      mcx_covg_inc(s + 4 + colour, may_saturate);
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

#endif // __CUDACC__
