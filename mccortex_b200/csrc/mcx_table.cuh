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
  // L2-resident front table (one colour at a time; k <= 31 here, the 16-byte-tag variant for k > 31 in mcx_device.cuh): a dense write-combining cache in front of
  // the big table.  The big table is >> L2 and a hot k-mer there drags a whole 128-byte L2 line
  // for 16 useful bytes, so on high-coverage input the hot set does not fit L2 (ncu, first
  // kernel: 127 B of DRAM traffic per occurrence, L2 hit rate 30 %).  Layout, the bijective hash
  // that makes (set, tag) identify the key exactly, and why tags and counters live in separate
  // regions are in mcx_device.cuh.  A key claims a way if one is free (first come, never
  // evicted); its occurrences are then counted here at L2 speed (one 32-byte load of the read-
  // mostly tags, one 32-bit RED into the write-only counters).  mcx_front_flush_kernel merges
  // every record into the big table at sync; sums and ORs commute and every record is merged
  // exactly once, so the final table is identical.  The counters are 32 bits wide: the host
  // flushes before 2^32 - 2^28 positions have been queued since the last flush, so none can wrap.
  unsigned long long *front;   // tags: (4 << front_set_bits) slots, or nullptr
  unsigned int *front_cnt;     // counters: one per slot
  uint32_t front_set_bits;     // log2(number of sets); 0 = no front table
  uint32_t front_colour;       // the ONE colour the front table is counting (it is flushed when the colour changes)
  uint32_t front_words;        // key words of the build (1: 8-byte tags, four ways per set; 2: 16-byte tags, two ways)
};

#if defined(__CUDACC__)

// 32-byte probe load (SASS: LDG.E.256).  .ca / .cg / volatile / L1::no_allocate flavours, L2 eviction-policy
// hints and two 16-byte loads were measured (profiles/r1_exp_ld_modes.txt, r1g_experiments.txt 5, 13): no
// difference except that two 16-byte loads are 45 % slower, so the plain form stays.
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
  // (n itself may be anything up to 2^32-1 when it comes from a graph file: 0xF0000000u - n must not wrap)
  if(!may_saturate || (have_snap && n < 0xF0000000u && snap < 0xF0000000u - n)) { atomicAdd(cv, n); return; }
  uint32_t v = *(volatile uint32_t *)cv;
  if(n < 0xF0000000u && v < 0xF0000000u - n) { atomicAdd(cv, n); return; }
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

// Probing stops after MCX_PROBE_CAP slots: the table then counts as full.  Linear probing over a table sized
// by the reference's rule (distinct / 0.75, src/graph/cmd_mem.c) has runs of a few hundred slots; runs of 2^16 need a
// load above 0.98.  Without the cap every missing k-mer of an undersized table would scan the whole table
// (the reference gives up after 20 full buckets: "Hash table is full", src/graph/hash_table.c:119-123,280).
// Insert and find use the same cap, so a key is always found where it was put.
#define MCX_PROBE_CAP 65536ull
__device__ __forceinline__ uint64_t mcx_probe_limit(const McxTable &t) { return t.nslots < MCX_PROBE_CAP ? t.nslots : MCX_PROBE_CAP; }

// find-or-insert in the big table, then covg[colour] += n (saturating) and edges |= emask.
// Returns 0 = found, 1 = novel, 2 = table full.
// `pre` (may be NULL): the 32 bytes at mcx_table_home(), already loaded by the caller so that several
// probes can be in flight per thread (mcx_table_prefetchable() says whether that sector is what the
// first probe reads).
template <int W>
__device__ __forceinline__ int mcx_table_add(const McxTable &t, const McxKmer<W> &key, uint32_t hc, uint32_t hb,
                                             uint32_t colour, uint32_t emask, uint32_t n, bool may_saturate,
                                             const uint64_t *pre = nullptr);
__device__ __forceinline__ bool mcx_table_prefetchable(const McxTable &t) { return t.stride == 4u || t.stride == 8u; }
__device__ __forceinline__ const uint32_t *mcx_table_home(const McxTable &t, uint32_t hc, uint32_t hb)
{
  uint64_t idx = mcx_home_slot(hc, hb, t.nslots);
  if(t.stride == 4u) idx &= ~1ull;
  return t.slots + idx * (uint64_t)t.stride;
}

// ---- k <= 31 --------------------------------------------------------------
template <>
__device__ __forceinline__ int mcx_table_add<1>(const McxTable &t, const McxKmer<1> &key, uint32_t hc, uint32_t hb,
                                                uint32_t colour, uint32_t emask, uint32_t n, bool may_saturate,
                                                const uint64_t *pre)
{
  const uint64_t keyf = key.b[0] | MCX_KEY_FLAG;
  uint64_t idx = mcx_home_slot(hc, hb, t.nslots);
  int novel = 0;
  if(t.stride == 4u) {
    // 16-byte slots {key, covg, edges}: probe one 32-byte sector (= 2 slots) per load
    idx &= ~1ull;
    for(uint64_t probes = 0, lim = mcx_probe_limit(t); probes < lim; probes += 2) {
      uint32_t *s = t.slots + idx * 4u;
      uint64_t k0, m0, k1, m1;
      if(probes == 0 && pre) { k0 = pre[0]; m0 = pre[1]; k1 = pre[2]; m1 = pre[3]; }
      else mcx_ld256(s, k0, m0, k1, m1);
      uint32_t *hit = nullptr; uint64_t meta = 0;
      if(k0 == keyf) { hit = s; meta = m0; }
      else if(k1 == keyf) { hit = s + 4; meta = m1; }
      else if(k0 == 0 || k1 == 0) {
        // first empty slot of the sector (slot 0 before slot 1 keeps the probe order total).  A novel k-mer takes the
        // WHOLE 16-byte slot with one 128-bit compare-and-swap -- key, coverage and edges at once -- instead of a 64-bit
        // CAS on the key followed by a RED on the coverage and a RED on the edges: one L2 atomic instead of three for what
        // is most of the work on cold data (error k-mers; all of configs[4]).  An empty slot is all zero: nothing ever
        // writes coverage or edges before the key.
        const uint64_t val = (uint64_t)n | ((uint64_t)emask << 32);
        uint32_t *cand = (k0 == 0) ? s : s + 4;
        uint64_t olo, ohi;
        mcx_cas128(cand, 0ull, 0ull, keyf, val, olo, ohi);
        if(olo == 0) return 1;
        if(olo == keyf) { hit = cand; meta = ohi; }
        else if(cand == s) {
          // lost slot 0 to another key: slot 1 of the same sector is next in probe order
          if(k1 == 0) {
            mcx_cas128(s + 4, 0ull, 0ull, keyf, val, olo, ohi);
            if(olo == 0) return 1;
            if(olo == keyf) { hit = s + 4; meta = ohi; }
          }
        }
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
  for(uint64_t probes = 0, lim = mcx_probe_limit(t); probes < lim; probes++) {
    uint32_t *s = t.slots + idx * (uint64_t)t.stride;
    uint64_t cur = (probes == 0 && pre) ? pre[0] : *(volatile uint64_t *)s;
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
                                                uint32_t colour, uint32_t emask, uint32_t n, bool may_saturate,
                                                const uint64_t *pre)
{
  const uint64_t k0f = key.b[0] | MCX_KEY_FLAG, k1 = key.b[1];
  uint64_t idx = mcx_home_slot(hc, hb, t.nslots);
  int novel = 0;
  for(uint64_t probes = 0, lim = mcx_probe_limit(t); probes < lim; probes++) {
    uint32_t *s = t.slots + idx * (uint64_t)t.stride;
    uint64_t c0, c1, m0 = 0, m1 = 0;
    bool have_meta = (t.stride == 8u);
    if(probes == 0 && pre) { c0 = pre[0]; c1 = pre[1]; m0 = pre[2]; m1 = pre[3]; }
    else if(have_meta) mcx_ld256(s, c0, c1, m0, m1); // 32-byte slot: key + covg + edges in one sector
    else mcx_ld128(s, c0, c1); // one 16-byte transaction: never a torn view of a 128-bit CAS
    if(c0 == 0) {
      mcx_cas128(s, 0ull, 0ull, k0f, k1, c0, c1);
      // claimed: the rest of the slot is still all zero as loaded (nothing writes coverage or edges before the key), so the
      // edge word is known and mcx_edges_or needs no load.  Lost the race: whoever won may have written them since.
      if(c0 == 0) { novel = 1; c0 = k0f; c1 = k1; }
      else have_meta = false;
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

// generic find / find-or-insert: returns the slot (u32 pointer) or nullptr (not found / table full).
// Probe order is the one mcx_table_add uses, so both see the same slots.
template <int W>
__device__ __forceinline__ uint32_t *mcx_table_slot(const McxTable &t, const McxKmer<W> &key, bool insert, int *novel, int *full)
{
  uint32_t hb, hc = mcx_lookup3<W>(key, 0u, &hb);
  uint64_t idx = mcx_home_slot(hc, hb, t.nslots);
  if(t.stride == 4u) idx &= ~1ull;
  const uint64_t k0f = key.b[0] | MCX_KEY_FLAG;
  for(uint64_t probes = 0, lim = mcx_probe_limit(t); probes < lim; probes++) {
    uint32_t *s = t.slots + idx * (uint64_t)t.stride;
    if(W == 1) {
      uint64_t cur = *(volatile uint64_t *)s;
      if(cur == 0) {
        if(!insert) return nullptr;
        cur = atomicCAS((unsigned long long *)s, 0ull, (unsigned long long)k0f);
        if(cur == 0) { *novel = 1; return s; }
      }
      if(cur == k0f) return s;
    } else {
      uint64_t c0, c1;
      mcx_ld128(s, c0, c1);
      if(c0 == 0) {
        if(!insert) return nullptr;
        mcx_cas128(s, 0ull, 0ull, k0f, key.b[W - 1], c0, c1);
        if(c0 == 0) { *novel = 1; return s; }
      }
      if(c0 == k0f && c1 == key.b[W - 1]) return s;
    }
    idx++; if(idx >= t.nslots) idx = 0;
  }
  *full = 1;
  return nullptr;
}

// ---- front table ------------------------------------------------------------
__device__ __forceinline__ McxFrontGeom mcx_front_geom(const McxTable &t) { return mcx_front_geom_bits(t.front_set_bits); }

// Slow side of the front table for one occurrence of `key` (k <= 31): re-reads the set, counts the
// occurrence if the key is there, claims a free way if it is not, adds a missing edge bit.
// A key that finds its home sector full of other keys (4.8 % of the hot k-mers of the bench
// workload: Poisson(2.2) keys per 4-way set) may live DISPLACED in the neighbouring sector
// (set ^ 1), marked by the top bit of its tag word so that it is never mistaken for a key of that
// set.  The hot pass only looks at the home sector, so a displaced key is always parked and ends
// up here: two L2 loads and a RED instead of a DRAM round trip to the big table -- and, in the
// sharded build, no tuple.  Returns false if the front table cannot absorb the occurrence (both
// sectors are full of other keys): then it belongs to the big table.
#define MCX_FRONT_DISPLACED 0x80000000u
// v[4]: the sector as the caller loaded it (possibly a little stale: a way that has been claimed since makes the CAS
// fail, and the value the CAS returns says whether the same key took it)
__device__ __forceinline__ bool mcx_front_resolve_sector(const McxTable &t, const McxFrontGeom &g, uint64_t s4, uint32_t x,
                                                         uint32_t th, uint32_t eb, const uint64_t v[4])
{
  const uint32_t mask = g.tagmask | MCX_FRONT_DISPLACED;
  unsigned long long *set = t.front + s4;
  int w = -1; uint32_t seen_hi = 0;
#pragma unroll
  for(int i = 3; i >= 0; i--) {
    const uint32_t lo = (uint32_t)v[i], hi = (uint32_t)(v[i] >> 32);
    if(lo == x && ((hi ^ th) & mask) == 0u) { w = i; seen_hi = hi; }
  }
  if(w < 0) {
#pragma unroll
    for(int i = 0; i < 4; i++) {
      if(w < 0 && v[i] == 0) {
        const uint64_t mine = ((uint64_t)(th | eb) << 32) | x;
        const uint64_t old = atomicCAS(&set[i], 0ull, (unsigned long long)mine);
        const uint32_t olo = (uint32_t)old, ohi = (uint32_t)(old >> 32);
        if(old == 0) { w = i; seen_hi = th | eb; }                                  // claimed, edges included
        else if(olo == x && ((ohi ^ th) & mask) == 0u) { w = i; seen_hi = ohi; }    // lost the race to the same key
      }
    }
    if(w < 0) return false;
  }
  atomicAdd(t.front_cnt + s4 + (uint32_t)w, 1u);
  if((seen_hi & eb) != eb) atomicOr(reinterpret_cast<unsigned int *>(&set[w]) + 1, eb);
  return true;
}
static __device__ __noinline__ bool mcx_front_add_slow(McxTable t, uint64_t key, uint32_t emask)
{
  const McxFrontGeom g = mcx_front_geom(t);
  const McxFKey fk = mcx_fhash(key);
  const uint32_t th = (fk.y >> g.S) | g.occ, eb = emask << g.eshift;
  const uint64_t s4 = (uint64_t)(fk.y & g.setmask) << 2;
  // both sectors are loaded at once: a displaced key (half of what is parked) costs one L2 round trip, not two
  uint64_t a[4], b[4];
  mcx_ld256(t.front + s4, a[0], a[1], a[2], a[3]);
  mcx_ld256(t.front + (s4 ^ 4ull), b[0], b[1], b[2], b[3]);
  if(mcx_front_resolve_sector(t, g, s4, fk.x, th, eb, a)) return true;
  return mcx_front_resolve_sector(t, g, s4 ^ 4ull, fk.x, th | MCX_FRONT_DISPLACED, eb, b);
}
// Fast side: the set has already been loaded (v0..v3).  Handles the overwhelmingly common case --
// the key sits in the set and its edge bits are already there -- with ONE 32-bit RED into the
// counter region and returns true; anything else returns false (-> parked, mcx_front_add_slow).
__device__ __forceinline__ bool mcx_front_hit(const McxFrontGeom &g, unsigned int *cnt_set, uint32_t x, uint32_t th, uint32_t eb,
                                              uint64_t v0, uint64_t v1, uint64_t v2, uint64_t v3)
{
  const uint32_t mask = g.tagmask | MCX_FRONT_DISPLACED; // a displaced entry belongs to the neighbouring set: never a match here
  const uint32_t h0 = (uint32_t)(v0 >> 32), h1 = (uint32_t)(v1 >> 32), h2 = (uint32_t)(v2 >> 32), h3 = (uint32_t)(v3 >> 32);
  const bool m0 = ((uint32_t)v0 == x) & (((h0 ^ th) & mask) == 0u);
  const bool m1 = ((uint32_t)v1 == x) & (((h1 ^ th) & mask) == 0u);
  const bool m2 = ((uint32_t)v2 == x) & (((h2 ^ th) & mask) == 0u);
  const bool m3 = ((uint32_t)v3 == x) & (((h3 ^ th) & mask) == 0u);
  const uint32_t hi = m0 ? h0 : (m1 ? h1 : (m2 ? h2 : h3));
  const uint32_t way = m0 ? 0u : (m1 ? 1u : (m2 ? 2u : 3u));
  if(!(m0 | m1 | m2 | m3) || (hi & eb) != eb) return false;
  atomicAdd(cnt_set + way, 1u);
  return true;
}

// ---- front table, 33 <= k <= 63: two 16-byte ways per set (layout: mcx_device.cuh) ------------------------------------
// v[4] = {lo0, hi0, lo1, hi1} of one set as the caller loaded it.  Same rules as k <= 31: first come, never evicted;
// a key whose home set is full may live displaced in set ^ 1.
__device__ __forceinline__ bool mcx_front2_resolve_sector(const McxTable &t, const McxFrontGeom2 &g, uint64_t set, uint64_t x,
                                                          uint64_t th, uint64_t eb, const uint64_t v[4])
{
  unsigned long long *ways = t.front + (set << 2);
  int w = -1; uint64_t seen_hi = 0;
#pragma unroll
  for(int i = 1; i >= 0; i--)
    if(v[2 * i] == x && ((v[2 * i + 1] ^ th) & g.mask) == 0ull) { w = i; seen_hi = v[2 * i + 1]; }
  if(w < 0) {
#pragma unroll
    for(int i = 0; i < 2; i++) {
      if(w < 0 && (v[2 * i] | v[2 * i + 1]) == 0ull) {
        uint64_t olo, ohi;
        mcx_cas128(ways + 2 * i, 0ull, 0ull, x, th | eb, olo, ohi);
        if((olo | ohi) == 0ull) { w = i; seen_hi = th | eb; }                         // claimed, edges included
        else if(olo == x && ((ohi ^ th) & g.mask) == 0ull) { w = i; seen_hi = ohi; }  // lost the race to the same key
      }
    }
    if(w < 0) return false;
  }
  atomicAdd(t.front_cnt + (set << 1) + (uint32_t)w, 1u);
  if((seen_hi & eb) != eb) atomicOr(ways + 2 * w + 1, (unsigned long long)eb);
  return true;
}
static __device__ __noinline__ bool mcx_front2_add_slow(McxTable t, uint64_t kh, uint64_t kl, uint32_t emask)
{
  const McxFrontGeom2 g = mcx_front_geom2_bits(t.front_set_bits);
  const McxFKey2 fk = mcx_fhash2(kh, kl);
  const uint64_t set = fk.y >> g.tshift, th = (fk.y & (g.occ - 1ull)) | g.occ, eb = (uint64_t)emask << g.eshift;
  uint64_t a[4], b[4];
  mcx_ld256(t.front + (set << 2), a[0], a[1], a[2], a[3]);
  mcx_ld256(t.front + ((set ^ 1ull) << 2), b[0], b[1], b[2], b[3]);
  if(mcx_front2_resolve_sector(t, g, set, fk.x, th, eb, a)) return true;
  return mcx_front2_resolve_sector(t, g, set ^ 1ull, fk.x, th | MCX_FRONT2_DISPLACED, eb, b);
}
// fast side: one 32-bit RED if the key sits in its home set with the edge bits already there
__device__ __forceinline__ bool mcx_front2_hit(const McxFrontGeom2 &g, unsigned int *cnt_set, uint64_t x, uint64_t th, uint64_t eb,
                                               uint64_t lo0, uint64_t hi0, uint64_t lo1, uint64_t hi1)
{
  const bool m0 = (lo0 == x) & (((hi0 ^ th) & g.mask) == 0ull);
  const bool m1 = (lo1 == x) & (((hi1 ^ th) & g.mask) == 0ull);
  const uint64_t hi = m0 ? hi0 : hi1;
  if(!(m0 | m1) || (hi & eb) != eb) return false;
  atomicAdd(cnt_set + (m0 ? 0u : 1u), 1u);
  return true;
}


#endif // __CUDACC__
