/* sizing.c -- `-n/-m` -> number of k-mer slots, with the reference's rules so the same
 * command line asks for the same capacity (src/basic/hash_mem.c:5-51, hash_mem.h:12-14,
 * src/graph/cmd_mem.c:38-130).  The device table does not use buckets; only the resulting
 * capacity (and the error conditions) carry over. */
#include "mcx_host.h"

static size_t ht_mem(size_t bktsize, size_t nbkts, size_t nbits) { return (bktsize * nbkts * nbits) / 8 + nbkts * 2; }

size_t mcx_hash_table_cap(uint64_t nkmers, uint64_t *num_bkts, uint8_t *bkt_size)
{
  uint64_t bits = 10, nb, bs;
  while(nkmers / (1UL << bits) > MCX_MAX_BUCKET_SIZE) bits++;
  nb = 1UL << bits;
  bs = (nkmers + nb - 1) / nb;
  if(bs < 1) bs = 1;
  if(num_bkts) *num_bkts = nb;
  if(bkt_size) *bkt_size = (uint8_t)bs;
  return nb * bs;
}

size_t mcx_hash_table_mem(uint64_t nkmers, size_t entrybits, uint64_t *nkmers_out)
{
  uint64_t nb, cap; uint8_t bs;
  cap = mcx_hash_table_cap(nkmers, &nb, &bs);
  if(nkmers_out) *nkmers_out = cap;
  return ht_mem(bs, nb, entrybits);
}

size_t mcx_hash_table_mem_limit(size_t memlimit, size_t entrybits, uint64_t *nkmers_out)
{
  size_t bs, bits = 10, nb = 1UL << bits, nk;
  while(ht_mem(MCX_MAX_BUCKET_SIZE, nb, entrybits) < memlimit) { bits++; nb = 1UL << bits; }
  bs = (memlimit - nb * 2) / ((nb * entrybits) / 8);
  if(bs == 0) {
    bits--; nb = 1UL << bits;
    nk = bs * nb;
    bs = nk / nb; if(bs < 1) bs = 1;
  }
  if(bs > MCX_MAX_BUCKET_SIZE) bs = MCX_MAX_BUCKET_SIZE;
  if(nkmers_out) *nkmers_out = nb * bs;
  return ht_mem(bs, nb, entrybits);
}

size_t mcx_get_kmers_in_hash(size_t mem_to_use, bool mem_to_use_set, size_t num_kmers, bool num_kmers_set,
                             size_t entry_bits, int64_t min_req, int64_t max_req, bool use_mem_limit,
                             size_t *graph_mem_ptr)
{
  uint64_t kmers_in_hash = 0;
  size_t graph_mem = 0, min_kmers_mem;
  char gm[64], mu[64], kh[64], mk[64], mm[64];

  mcx_status("[memory] %zu bits per kmer", entry_bits);

  if(num_kmers_set) graph_mem = mcx_hash_table_mem(num_kmers, entry_bits, &kmers_in_hash);
  else if(use_mem_limit) graph_mem = mcx_hash_table_mem_limit(mem_to_use, entry_bits, &kmers_in_hash);
  else if(min_req > 0) graph_mem = mcx_hash_table_mem((size_t)(min_req / MCX_IDEAL_OCCUPANCY), entry_bits, &kmers_in_hash);

  if(max_req > 0 && !num_kmers_set) {
    size_t gm2; uint64_t kh2;
    gm2 = mcx_hash_table_mem(max_req / MCX_IDEAL_OCCUPANCY, entry_bits, &kh2);
    if(gm2 < graph_mem) { graph_mem = gm2; kmers_in_hash = kh2; }
  }
  if(kmers_in_hash < 1024) graph_mem = mcx_hash_table_mem(1024, entry_bits, &kmers_in_hash);

  uint64_t min_nkmers = min_req < 0 ? 1024 : (uint64_t)min_req;
  min_kmers_mem = mcx_hash_table_mem(min_nkmers, entry_bits, NULL);
  mcx_bytes_to_str(graph_mem, gm); mcx_bytes_to_str(mem_to_use, mu); mcx_bytes_to_str(min_kmers_mem, mm);
  mcx_ulong_to_str(kmers_in_hash, kh); mcx_ulong_to_str(min_nkmers, mk);

  if(min_req >= 0) {
    if(kmers_in_hash < (uint64_t)min_req)
      mcx_die("Not enough kmers in hash: require at least %s kmers (min memory: %s)", mk, mm);
    else if(kmers_in_hash < min_req / MCX_WARN_OCCUPANCY)
      mcx_warn("Expected hash table occupancy %.2f%% [%zu / %zu](you may want to increase -n or -m)",
               (100.0 * min_req) / kmers_in_hash, (size_t)min_req, (size_t)kmers_in_hash);
  }
  if(mem_to_use_set && num_kmers_set) {
    if(num_kmers > kmers_in_hash)
      mcx_die("-n <kmers> requires more memory than given with -m <mem> [%s > %s]", gm, mu);
    else if(use_mem_limit && graph_mem < mem_to_use)
      mcx_status("Note: Using less memory than requested (%s < %s); allows for %s kmers", gm, mu, kh);
  }
  if(graph_mem > mem_to_use)
    mcx_die("Not enough memory for requested graph: require at least %s [>%s]", gm, mu);

  mcx_status("[memory] graph: %s", gm);
  if(graph_mem_ptr) *graph_mem_ptr = graph_mem;
  return kmers_in_hash;
}
