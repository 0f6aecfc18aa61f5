// tests/emul/emul_frontend.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Executes the host/device (MCX_HD) math of mccortex_b200/csrc/mcx_device.cuh and
// mcx_chunk.cuh on the CPU, walking chunks exactly like mcx_front_end() in
// mcx_build.cu (phase 1 / 2a / 2b over the same staged arrays), and writes the sorted
// .ctx records + counters so tests can compare them with the oracle WITHOUT a GPU.
// It exists because the build container has no GPU; it is never linked into
// libmcxgpu.so and the product never calls it.
//
// usage: emul_frontend <lines-file> <k> <hp_cutoff> <r_piece> [<qual-lines-file> <qcut>]  -> stdout: records
//        r_piece: positions per simulated launch (0 = one launch), to exercise the
//        host-staging cut logic of mcx_abi.cu (look-back 16 / look-ahead 80).
//        qual-lines-file: quality bytes parallel to the lines file (quality cut-off mode:
//        two passes per launch like mcx_launch_build_fused_qual; launches must start at read
//        boundaries, so r_piece must be 0).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>
#include <array>
#include "../../mccortex_b200/csrc/mcx_chunk.cuh"
#include "../../mccortex_b200/csrc/mcx_pcr.cuh"
#include "../../mccortex_b200/csrc/mcx_lane.cuh"

struct Rec { uint32_t covg; uint8_t edges; };
typedef std::map<std::array<uint64_t, 2>, Rec> Table;

struct Counters { uint64_t kmers = 0, novel = 0, contigs = 0, reads = 0; };

template <int W>
static void run_launch(const uint8_t *seq, uint64_t nbytes, uint64_t r_begin, uint64_t r_end, uint32_t k, uint32_t hp,
                       Table &tab, Counters &cnt, const uint8_t *qual = nullptr, uint32_t qcut = 0)
{
  std::vector<uint8_t> raw(MCX_RAW), qraw(MCX_RAW);
  std::vector<uint32_t> pk(MCX_PKW), bad(MCX_MSW), bads(MCX_MSW), eq(MCX_MSW), vmask(MCX_VW), svm(MCX_VW);
  uint64_t c_first = r_begin / MCX_T, c_last = (r_end + MCX_T - 1) / MCX_T;
  std::vector<uint8_t> summary(c_last - c_first + 1);
  for(int pass = qual ? 0 : 1; pass < 2; pass++)   // pass 0 = summary kernel (quality mode only)
  for(uint64_t chunk = c_first; chunk < c_last; chunk++) {
    uint64_t cs = chunk * (uint64_t)MCX_T;
    // staging (issue_chunk_load): bytes outside the copied range hold garbage
    memset(raw.data(), 0x5A, MCX_RAW);
    uint64_t src_off = cs ? cs - MCX_LB : 0; uint32_t dst_off = cs ? 0 : MCX_LB;
    uint64_t avail = (nbytes - src_off + 15ull) & ~15ull; uint32_t want = MCX_RAW - dst_off;
    uint32_t bytes = avail < want ? (uint32_t)avail : want;
    for(uint32_t i = 0; i < bytes; i++) raw[dst_off + i] = (src_off + i < nbytes) ? seq[src_off + i] : 0xEE;
    if(qual) { memset(qraw.data(), 0x11, MCX_RAW); for(uint32_t i = 0; i < bytes; i++) qraw[dst_off + i] = (src_off + i < nbytes) ? qual[src_off + i] : 0x22; }
    for(uint32_t t = 0; t < 4; t++) { pk[MCX_RAW / 16u + t] = 0; bad[MCX_RAW / 32u + t] = 0xFFFFFFFFu; eq[MCX_RAW / 32u + t] = 0; }
    // phase 1
    for(uint32_t tid = 0; tid < MCX_RAW / 16u; tid++) {
      uint32_t w[4]; memcpy(w, &raw[tid * 16u], 16);
      uint32_t prev = tid ? raw[tid * 16u - 1u] : 0u;
      uint64_t gpos = cs - MCX_LB + tid * 16ull;
      uint32_t p, b16, e16, n16;
      mcx_convert16(w, prev, gpos, nbytes, &p, &b16, &e16, &n16);
      pk[tid] = p;
      if(qual) {
        uint32_t q[4]; memcpy(q, &qraw[tid * 16u], 16);
        uint32_t wk16, st16; mcx_qual16(q, qcut, &wk16, &st16);
        ((uint16_t *)bads.data())[tid] = (uint16_t)(b16 | st16);
        b16 |= wk16;
      }
      ((uint16_t *)bad.data())[tid] = (uint16_t)b16;
      ((uint16_t *)eq.data())[tid] = (uint16_t)e16;
      if(pass == 1 && n16 && tid >= MCX_LB / 16u && tid < (MCX_LB + MCX_T) / 16u)
        for(uint32_t i = 0; i < 16u; i++)
          if(((n16 >> i) & 1u) && gpos + i >= r_begin && gpos + i < r_end) cnt.reads++;
    }
    // phase 2a (word-parallel, masks indexed by staged position)
    for(uint32_t w = 0; w < MCX_VW; w++) {
      bool live = w < (MCX_LB + MCX_T + 32u) / 32u;
      vmask[w] = live ? mcx_valid_word(bad.data(), eq.data(), w, k, hp) : 0u;
      svm[w] = (qual && live) ? mcx_valid_word(bads.data(), eq.data(), w, k, hp) : 0u;
    }
    if(qual) {
      const uint32_t cb = MCX_LB - 1u, keep = ~0u << cb;
      const uint32_t ev0 = vmask[0] & keep & ~(1u << cb), sv0 = svm[0] & keep & ~(1u << cb);
      if(pass == 0) {
        uint32_t out = 0;
        for(uint32_t cin = 0; cin < 2u; cin++) {
          vmask[0] = ev0 | (cin << cb); svm[0] = sv0 | (cin << cb);
          std::vector<uint32_t> x(MCX_VW);
          mcx_contig_chain(vmask.data(), svm.data(), MCX_VW, 0u, x.data());
          out |= mcx_get_bit(x.data(), MCX_LB - 1u + MCX_T) << cin;
        }
        summary[chunk - c_first] = (uint8_t)out;
        continue;
      }
      uint32_t cin = 0; uint64_t j = chunk;
      while(j > c_first) { uint32_t sm = summary[--j - c_first]; if(sm == 0u) { cin = 0; break; } if(sm == 3u) { cin = 1; break; } }
      vmask[0] = ev0 | (cin << cb); svm[0] = sv0 | (cin << cb);
      mcx_contig_chain(vmask.data(), svm.data(), MCX_VW, 0u, vmask.data());
    }
    // phase 2b: thread t walks 8 consecutive windows with rolling k-mers, four at a time
    for(uint32_t t = 0; t < MCX_T / MCX_WPT; t++) {
      mcx_thread_occurrences<W>(pk.data(), vmask.data(), t, k,
        [&](const McxKmer<W> *keys, const uint32_t *emasks, uint32_t valid, uint32_t starts, uint32_t j0) {
          for(uint32_t i = 0; i < MCX_HALF; i++) {
            uint64_t g = cs + MCX_WPT * t + j0 + i;
            if(!((valid >> i) & 1u) || g < r_begin || g >= r_end) continue;
            cnt.kmers++;
            cnt.contigs += (starts >> i) & 1u;
            std::array<uint64_t, 2> key = {keys[i].b[0], W == 2 ? keys[i].b[W - 1] : 0};
            auto it = tab.find(key);
            if(it == tab.end()) { tab[key] = Rec{1, (uint8_t)emasks[i]}; cnt.novel++; }
            else { if(it->second.covg != 0xFFFFFFFFu) it->second.covg++; it->second.edges |= (uint8_t)emasks[i]; }
          }
        });
    }
  }
}

static std::vector<uint8_t> slurp(const char *path)
{
  std::vector<uint8_t> data; uint8_t buf[1 << 16]; size_t n;
  FILE *f = fopen(path, "rb"); if(!f) { perror(path); exit(2); }
  while((n = fread(buf, 1, sizeof(buf), f)) > 0) data.insert(data.end(), buf, buf + n);
  fclose(f);
  return data;
}


// --lane: the warp-autonomous front end (mcx_lane.cuh) walked tile by tile, lane by lane, like
// mcx_build_warp_kernel does (pieces of 16 bytes; a lane sees its piece, the two after it and the last
// base of the one before); what the GPU gets from neighbouring lanes by shuffle is read from the buffer here.
static void lane_piece(const uint8_t *seq, uint64_t nbytes, uint64_t g, uint32_t *pk, uint32_t *bad, uint32_t *nl)
{
  const uint64_t npieces = (nbytes + 15) >> 4;
  uint32_t w[4] = {0, 0, 0, 0};
  if(g < npieces) {
    uint8_t raw[16];
    for(uint32_t i = 0; i < 16; i++) raw[i] = (g * 16 + i < nbytes) ? seq[g * 16 + i] : (uint8_t)('A' + (i & 1)); // garbage that looks like data
    memcpy(w, raw, 16);
  }
  mcx_piece_convert(w, g * 16, nbytes, pk, bad, nl);
}
static void run_launch_lane(const uint8_t *seq, uint64_t nbytes, uint64_t r_begin, uint64_t r_end, uint32_t k, Table &tab, Counters &cnt)
{
  const uint64_t T0 = r_begin / MCX_TILE, T1 = (r_end + MCX_TILE - 1) / MCX_TILE;
  for(uint64_t t = T0; t < T1; t++)
    for(uint32_t lane = 0; lane < 32; lane++) {
      const uint64_t g = 32 * t + lane;
      uint32_t pk[3], bad[3], nl[3];
      for(int i = 0; i < 3; i++) lane_piece(seq, nbytes, g + i, &pk[i], &bad[i], &nl[i]);
      uint32_t prev_base = 0, prev_bad = 1;
      if(g > 0) { uint32_t ppk, pbad, pnl; lane_piece(seq, nbytes, g - 1, &ppk, &pbad, &pnl); prev_base = ppk & 3u; prev_bad = (pbad >> 15) & 1u; }
      const uint32_t own = mcx_piece_own(g * 16, r_begin, r_end);
      cnt.reads += __builtin_popcount(nl[0] & own);
      const uint64_t bad48 = (uint64_t)bad[0] | ((uint64_t)bad[1] << 16) | ((uint64_t)bad[2] << 32);
      const uint32_t vb = mcx_lane_valid(bad48, prev_bad, k);
      uint32_t calls = 0;
      mcx_lane_windows(pk[0], pk[1], pk[2], vb, prev_base, k,
        [&](const McxKmer<1> *keys, const uint32_t *emasks, uint32_t valid, uint32_t starts, uint32_t j0) {
          calls++;
          valid &= own >> j0;
          for(uint32_t i = 0; i < MCX_HALF; i++) {
            if(!((valid >> i) & 1u)) continue;
            cnt.kmers++;
            cnt.contigs += (starts >> i) & 1u;
            std::array<uint64_t, 2> key = {keys[i].b[0], 0};
            auto it = tab.find(key);
            if(it == tab.end()) { tab[key] = Rec{1, (uint8_t)emasks[i]}; cnt.novel++; }
            else { if(it->second.covg != 0xFFFFFFFFu) it->second.covg++; it->second.edges |= (uint8_t)emasks[i]; }
          }
        });
      if(calls != MCX_LW / MCX_HALF) { fprintf(stderr, "lane front end must call its sink for every group\n"); exit(3); }
    }
}
static int main_lane(int argc, char **argv)
{
  if(argc < 5) { fprintf(stderr, "usage: %s --lane <lines-file> <k> <r_piece>\n", argv[0]); return 2; }
  std::vector<uint8_t> data = slurp(argv[2]);
  uint32_t k = (uint32_t)atoi(argv[3]);
  uint64_t piece = strtoull(argv[4], NULL, 10), nbytes = data.size();
  if(k > 31) { fprintf(stderr, "--lane: k <= 31\n"); return 2; }
  if(piece == 0) piece = nbytes ? nbytes : 1;
  Table tab; Counters cnt;
  for(uint64_t pos = 0; pos < nbytes; pos += piece) {
    uint64_t pend = pos + piece < nbytes ? pos + piece : nbytes;
    uint64_t b0 = pos ? pos - MCX_LB : 0, b1 = pend + MCX_TAIL < nbytes ? pend + MCX_TAIL : nbytes;
    if(b0 % 16) { fprintf(stderr, "piece must be a multiple of 16\n"); return 2; }
    run_launch_lane(data.data() + b0, b1 - b0, pos - b0, pend - b0, k, tab, cnt);
  }
  for(auto &kv : tab) {
    fwrite(&kv.first[0], 8, 1, stdout);
    fwrite(&kv.second.covg, 4, 1, stdout);
    fwrite(&kv.second.edges, 1, 1, stdout);
  }
  fprintf(stderr, "kmers=%llu novel=%llu contigs=%llu reads=%llu\n", (unsigned long long)cnt.kmers,
          (unsigned long long)cnt.novel, (unsigned long long)cnt.contigs, (unsigned long long)cnt.reads);
  return 0;
}

// --pcr <lines-file> <k> <hp> <qual-lines-file|-> <qcut> <mate-file> <batch_reads>
// the three passes of mcx_pcr.cu (orient / mark / mask) batch by batch with the MCX_HD math of mcx_pcr.cuh,
// a std::map standing in for the table's slot numbers, then the normal front end over the filtered lines
static int main_pcr(int argc, char **argv)
{
  if(argc < 9) { fprintf(stderr, "usage: %s --pcr <lines> <k> <hp> <qual|-> <qcut> <mates> <batch_reads>\n", argv[0]); return 2; }
  std::vector<uint8_t> data = slurp(argv[2]), qdata, mate = slurp(argv[7]);
  uint32_t k = (uint32_t)atoi(argv[3]), hp = (uint32_t)atoi(argv[4]), qcut = (uint32_t)atoi(argv[6]);
  if(strcmp(argv[5], "-") != 0) qdata = slurp(argv[5]);
  uint64_t batch = strtoull(argv[8], NULL, 10);
  uint8_t *qp = qdata.empty() ? nullptr : qdata.data();
  std::vector<uint64_t> off(1, 0);
  for(uint64_t i = 0; i < data.size(); i++) if(data[i] == '\n') off.push_back(i + 1);
  uint64_t nreads = off.size() - 1;
  if(mate.size() != nreads || (qp && qdata.size() != data.size())) { fprintf(stderr, "mate/qual files do not match the lines\n"); return 2; }
  if(batch == 0) batch = nreads ? nreads : 1;
  std::map<std::array<uint64_t, 3>, uint64_t> ids;   // (key words, orient) -> node number
  std::vector<uint32_t> first;
  std::vector<uint64_t> node(nreads);
  uint64_t dup_se = 0, dup_pe = 0; uint32_t ord_base = 0;
  for(uint64_t b0 = 0, n; b0 < nreads; b0 += n) {
    n = nreads - b0 < batch ? nreads - b0 : batch;
    if((mate[b0 + n - 1] & MCX_MATE_KIND) == MCX_MATE_FIRST) n++;   // a pair never straddles two batches
    const uint64_t *boff = off.data() + b0; const uint8_t *bmate = mate.data() + b0; uint64_t *bnode = node.data() + b0;
    for(uint64_t r = 0; r < n; r++) if(bmate[r] & MCX_MATE_REVCOMP)
      for(uint32_t lane = 0; lane < 32; lane++)
        mcx_pcr_revcomp_lanes(data.data() + boff[r], qp ? qp + boff[r] : nullptr, boff[r + 1] - boff[r] - 1, lane, 32);
    for(uint64_t r = 0; r < n; r++) {
      uint64_t lo = boff[r], len = boff[r + 1] - lo - 1;
      uint64_t start = mcx_first_contig_start(data.data() + lo, qp ? qp + lo : nullptr, len, k, qp ? qcut : 0, hp);
      bnode[r] = MCX_PCR_NONE;
      if(start >= len) continue;
      uint32_t orient; std::array<uint64_t, 3> id = {0, 0, 0};
      if(k <= 31) { McxKmer<1> key = mcx_kmer_key<1>(mcx_kmer_from_ascii<1>(data.data() + lo + start, k), k, &orient); id[0] = key.b[0]; }
      else { McxKmer<2> key = mcx_kmer_key<2>(mcx_kmer_from_ascii<2>(data.data() + lo + start, k), k, &orient); id[0] = key.b[0]; id[1] = key.b[1]; }
      id[2] = orient;
      auto it = ids.find(id);
      if(it == ids.end()) { it = ids.emplace(id, (uint64_t)first.size()).first; first.push_back(MCX_PCR_UNSET); }
      bnode[r] = it->second;
      uint32_t ord = ord_base + (uint32_t)mcx_pcr_leader(r, bmate[r]);
      if(ord < first[it->second]) first[it->second] = ord;
    }
    for(uint64_t r = 0; r < n; r++) {
      if(!mcx_pcr_is_dup(r, bmate, bnode, first.data(), ord_base)) continue;
      for(uint64_t i = boff[r]; i + 1 < boff[r + 1]; i++) data[i] = 'N';
      uint32_t kind = bmate[r] & MCX_MATE_KIND;
      if(kind == MCX_MATE_FIRST) dup_pe++; else if(kind == MCX_MATE_SINGLE) dup_se++;
    }
    ord_base += (uint32_t)n;
  }
  Table tab; Counters cnt;
  uint64_t nbytes = data.size();
  if(nbytes) {
    if(k <= 31) run_launch<1>(data.data(), nbytes, 0, nbytes, k, hp, tab, cnt, qp, qcut);
    else run_launch<2>(data.data(), nbytes, 0, nbytes, k, hp, tab, cnt, qp, qcut);
  }
  int W = k <= 31 ? 1 : 2;
  for(auto &kv : tab) {
    fwrite(&kv.first[0], 8, 1, stdout);
    if(W == 2) fwrite(&kv.first[1], 8, 1, stdout);
    fwrite(&kv.second.covg, 4, 1, stdout);
    fwrite(&kv.second.edges, 1, 1, stdout);
  }
  fprintf(stderr, "kmers=%llu novel=%llu contigs=%llu reads=%llu dupse=%llu duppe=%llu\n", (unsigned long long)cnt.kmers,
          (unsigned long long)cnt.novel, (unsigned long long)cnt.contigs, (unsigned long long)cnt.reads,
          (unsigned long long)dup_se, (unsigned long long)dup_pe);
  return 0;
}

// --fhash <n> <seed>: the front tables' key <-> (set, tag) maps are bijections (inverse round trip, both widths) and
// spread k-mer-like keys over the sets (prints the worst set load for 2^S * 2 keys over 2^S sets, S = 16)
static int main_fhash(int argc, char **argv)
{
  uint64_t n = argc > 2 ? strtoull(argv[2], NULL, 10) : 100000, z = argc > 3 ? strtoull(argv[3], NULL, 10) : 1;
  auto rnd = [&z]() { z += 0x9E3779B97F4A7C15ull; uint64_t x = z; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); };
  const uint32_t S = 16;
  std::vector<uint32_t> load1(1u << S, 0), load2(1u << S, 0);
  const McxFrontGeom g1 = mcx_front_geom_bits(S); const McxFrontGeom2 g2 = mcx_front_geom2_bits(S);
  uint64_t bad = 0;
  for(uint64_t i = 0; i < n; i++) {
    // keys the way a graph has them: random, or runs of one base / neighbours that differ in one base
    uint64_t kh = rnd() >> 2, kl = rnd();
    if(i % 7 == 3) { kh = 0; kl = i; } else if(i % 7 == 5) { kh = (0x5555555555555555ull * (i & 3)) >> 2; kl = 0x5555555555555555ull * (i & 3) ^ (i << 2); }
    McxFKey a = mcx_fhash(((kh & 0x3FFFFFFFull) << 32) | (uint32_t)kl);
    if(mcx_fhash_inv(a.x, a.y) != (((kh & 0x3FFFFFFFull) << 32) | (uint32_t)kl) || (a.y >> 30)) bad++;
    McxFKey2 b = mcx_fhash2(kh, kl);
    uint64_t rh, rl; mcx_fhash2_inv(b.x, b.y, &rh, &rl);
    if(rh != kh || rl != kl || (b.y >> 62)) bad++;
    // set + tag bits rebuild y exactly
    const uint64_t set = b.y >> g2.tshift, th = b.y & (g2.occ - 1ull);
    if(((set << g2.tshift) | th) != b.y || set >> S) bad++;
    load1[a.y & g1.setmask]++; load2[set]++;
  }
  uint32_t m1 = 0, m2 = 0;
  for(uint32_t i = 0; i < (1u << S); i++) { if(load1[i] > m1) m1 = load1[i]; if(load2[i] > m2) m2 = load2[i]; }
  printf("bad=%llu max_load1=%u max_load2=%u mean=%.3f\n", (unsigned long long)bad, m1, m2, (double)n / (1u << S));
  return bad ? 1 : 0;
}

// --chain <n> <seed>: the warp-wide contig chain and the closed-form chunk summary of the quality modes, lane by lane
// on the CPU (mcx_chain_lane_* / mcx_summary_lane_*, mcx_chunk.cuh -- the pieces the device wrappers in mcx_build.cu
// glue together with shuffles), against the serial mcx_contig_chain on random masks of every density.
static int main_chain(int argc, char **argv)
{
  uint64_t n = argc > 2 ? strtoull(argv[2], NULL, 10) : 20000, z0 = argc > 3 ? strtoull(argv[3], NULL, 10) : 1;
  auto rnd = [&z0]() { z0 += 0x9E3779B97F4A7C15ull; uint64_t x = z0; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); };
  const uint32_t cb = MCX_LB - 1u, f = MCX_LB - 1u + MCX_T;
  uint64_t bad = 0, in_contig_f = 0, depends = 0;
  for(uint64_t it = 0; it < n; it++) {
    // densities from "almost all set" (long contigs: the interesting carries) to sparse
    const uint32_t de = (uint32_t)(rnd() % 7), ds = (uint32_t)(rnd() % 7);
    auto word = [&](uint32_t d) { uint32_t w = (uint32_t)rnd(); for(uint32_t i = 0; i < d; i++) w |= (uint32_t)rnd(); if(d == 6) w = ~0u; return w; };
    std::vector<uint32_t> ev(MCX_VW), sv(MCX_VW);
    for(uint32_t w = 0; w < MCX_VW; w++) { ev[w] = word(de); sv[w] = (uint32_t)rnd() & (uint32_t)rnd() & (ds < 3 ? (uint32_t)rnd() : ~0u) & ev[w]; if(ds == 6) sv[w] = 0; }
    if(it % 5 == 0) for(uint32_t w = (f >> 5) + 1; w < MCX_VW; w++) ev[w] = sv[w] = 0; // what the kernel has beyond the chunk
    for(uint32_t cin = 0; cin < 2u; cin++) {
      // serial reference, as the first version of the kernel did it
      std::vector<uint32_t> e0 = ev, s0 = sv, want(MCX_VW);
      const uint32_t keep = (~0u << cb) & ~(1u << cb);
      e0[0] = (e0[0] & keep) | (cin << cb); s0[0] = (s0[0] & keep) | (cin << cb);
      mcx_contig_chain(e0.data(), s0.data(), MCX_VW, 0u, want.data());
      // 32 lanes
      std::vector<uint32_t> vm = ev;
      uint32_t a[32][MCX_CHAIN_WPL], b[32][MCX_CHAIN_WPL], f0[32], f1[32];
      for(uint32_t l = 0; l < 32; l++) { mcx_chain_lane_load(vm.data(), sv.data(), cb, cin, l, a[l], b[l]); mcx_chain_lane_carry(a[l], b[l], &f0[l], &f1[l]); }
      for(uint32_t d = 1; d < 32; d <<= 1) {
        uint32_t p0[32], p1[32]; memcpy(p0, f0, sizeof(p0)); memcpy(p1, f1, sizeof(p1)); // __shfl_up reads the values before the step
        for(uint32_t l = d; l < 32; l++) mcx_chain_compose(p0[l - d], p1[l - d], &f0[l], &f1[l]);
      }
      for(uint32_t l = 0; l < 32; l++) mcx_chain_lane_store(vm.data(), a[l], b[l], l ? f0[l - 1] : 0u, l);
      if(vm != want) bad++;
      // summary: closed form against the chain's bit f
      int z = -1;
      for(uint32_t l = 0; l < 32; l++) { int c = mcx_summary_lane_last_zero(ev.data(), cb, f, l); if(c > z) z = c; }
      uint32_t any = 0;
      for(uint32_t l = 0; l < 32; l++) any |= mcx_summary_lane_starts(sv.data(), cb, f, z, l);
      const uint32_t out = (any ? 1u : 0u) | (((any ? 1u : 0u) | (z < 0 ? 1u : 0u)) << 1);
      if(((out >> cin) & 1u) != mcx_get_bit(want.data(), f)) bad++;
      in_contig_f += mcx_get_bit(want.data(), f); if(cin == 1u && (out == 2u)) depends++;
    }
  }
  printf("bad=%llu cases=%llu last_window_in_contig=%llu carry_dependent=%llu\n", (unsigned long long)bad, (unsigned long long)(2 * n),
         (unsigned long long)in_contig_f, (unsigned long long)depends);
  return bad ? 1 : 0;
}

int main(int argc, char **argv)
{
  if(argc > 1 && strcmp(argv[1], "--chain") == 0) return main_chain(argc, argv);
  if(argc > 1 && strcmp(argv[1], "--fhash") == 0) return main_fhash(argc, argv);
  if(argc > 1 && strcmp(argv[1], "--pcr") == 0) return main_pcr(argc, argv);
  if(argc > 1 && strcmp(argv[1], "--lane") == 0) return main_lane(argc, argv);
  if(argc < 5) { fprintf(stderr, "usage: %s <lines-file> <k> <hp> <r_piece>\n", argv[0]); return 2; }
  FILE *f = fopen(argv[1], "rb"); if(!f) { perror(argv[1]); return 2; }
  std::vector<uint8_t> data; uint8_t buf[1 << 16]; size_t n;
  while((n = fread(buf, 1, sizeof(buf), f)) > 0) data.insert(data.end(), buf, buf + n);
  fclose(f);
  uint32_t k = (uint32_t)atoi(argv[2]), hp = (uint32_t)atoi(argv[3]);
  uint64_t piece = strtoull(argv[4], NULL, 10), nbytes = data.size();
  std::vector<uint8_t> qdata; uint32_t qcut = 0;
  if(argc >= 7) {
    FILE *qf = fopen(argv[5], "rb"); if(!qf) { perror(argv[5]); return 2; }
    while((n = fread(buf, 1, sizeof(buf), qf)) > 0) qdata.insert(qdata.end(), buf, buf + n);
    fclose(qf);
    qcut = (uint32_t)atoi(argv[6]);
    if(qdata.size() != data.size() || piece != 0) { fprintf(stderr, "qual file must parallel the lines file; r_piece must be 0\n"); return 2; }
  }
  const uint8_t *qp = qdata.empty() ? nullptr : qdata.data();
  if(piece == 0) piece = nbytes ? nbytes : 1;
  Table tab; Counters cnt;
  for(uint64_t pos = 0; pos < nbytes; pos += piece) {
    uint64_t pend = pos + piece < nbytes ? pos + piece : nbytes;
    uint64_t b0 = pos ? pos - MCX_LB : 0, b1 = pend + MCX_TAIL < nbytes ? pend + MCX_TAIL : nbytes;
    if(b0 % 16) { fprintf(stderr, "piece must be a multiple of 16\n"); return 2; }
    if(k <= 31) run_launch<1>(data.data() + b0, b1 - b0, pos - b0, pend - b0, k, hp, tab, cnt, qp ? qp + b0 : nullptr, qcut);
    else run_launch<2>(data.data() + b0, b1 - b0, pos - b0, pend - b0, k, hp, tab, cnt, qp ? qp + b0 : nullptr, qcut);
  }
  int W = k <= 31 ? 1 : 2;
  for(auto &kv : tab) {
    fwrite(&kv.first[0], 8, 1, stdout);
    if(W == 2) fwrite(&kv.first[1], 8, 1, stdout);
    fwrite(&kv.second.covg, 4, 1, stdout);
    fwrite(&kv.second.edges, 1, 1, stdout);
  }
  fprintf(stderr, "kmers=%llu novel=%llu contigs=%llu reads=%llu\n", (unsigned long long)cnt.kmers,
          (unsigned long long)cnt.novel, (unsigned long long)cnt.contigs, (unsigned long long)cnt.reads);
  return 0;
}
