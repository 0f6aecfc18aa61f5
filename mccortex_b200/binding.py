"""ctypes mirror of include/mcx_gpu.h (the C ABI of libmcxgpu.so).

Names and argument meaning follow the reference interface the ABI replaces
(db_graph_alloc / build_graph / build_graph_from_str_mt / graph_writer_save_mkhdr,
see the header for file:line), so tests read like the reference's own tests.
Nothing here computes anything: every method is one C call.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))

MCX_LAYOUT_LINES, MCX_LAYOUT_OFFSETS = 0, 1
MCX_MEM_HOST, MCX_MEM_DEVICE = 0, 1
MCX_GRAPH_INTERSECT = 1
MCX_GRAPH_READSTRT = 2
MCX_MATE_SINGLE, MCX_MATE_FIRST, MCX_MATE_SECOND, MCX_MATE_REVCOMP = 0, 1, 2, 4
MCX_LOAD_MUST_EXIST, MCX_LOAD_INTO_ISEC, MCX_LOAD_MASK_ISEC = 1, 2, 4

_STATUS = {0: "MCX_OK", 1: "MCX_ERR_BAD_ARG", 2: "MCX_ERR_CUDA", 3: "MCX_ERR_TABLE_FULL",
           4: "MCX_ERR_NOMEM", 5: "MCX_ERR_UNSUPPORTED", 6: "MCX_ERR_NO_DEVICE"}


class McxError(RuntimeError):
    def __init__(self, code, what, detail=""):
        self.code = code
        self.status = _STATUS.get(code, str(code))
        super().__init__("%s failed: %s%s" % (what, self.status, (" (" + detail + ")") if detail else ""))


class ReadBatch(C.Structure):
    _fields_ = [("seq", C.c_void_p), ("qual", C.c_void_p), ("offsets", C.c_void_p),
                ("nreads", C.c_uint64), ("nbytes", C.c_uint64),
                ("layout", C.c_uint32), ("mem", C.c_uint32), ("colour", C.c_uint32),
                ("fq_cutoff", C.c_uint8), ("hp_cutoff", C.c_uint8), ("must_exist", C.c_uint8), ("reserved", C.c_uint8)]


class LoadStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "total_bases_read", "total_bases_loaded", "contigs_parsed", "num_kmers_loaded",
        "num_kmers_novel", "num_se_reads", "num_pe_reads", "num_good_reads", "num_bad_reads",
        "num_dup_se_reads", "num_dup_pe_pairs")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def lib_path():
    return os.path.join(HERE, "lib", "libmcxgpu.so")


def driver_path():
    return os.path.join(HERE, "bin", "mccortex-b200")


def build_native(targets=("all",)):
    """Compile libmcxgpu.so (nvcc, sm_100a) and the C host driver in-tree."""
    subprocess.check_call(["make", "-s", "-C", HERE] + list(targets))


_lib = None


def lib():
    """Load libmcxgpu.so.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise ImportError("%s is missing: run `make -C mccortex_b200` (or __graft_entry__.build()); "
                          "mccortex_b200 has no CPU fallback" % p)
    L = C.CDLL(p)
    vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
    L.mcx_device_count.restype = C.c_int
    L.mcx_last_error.restype = C.c_char_p
    L.mcx_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.mcx_host_free.argtypes = [vp]
    L.mcx_graph_create.argtypes = [u32, u32, u64, C.c_int, u32, C.POINTER(vp)]
    L.mcx_graph_destroy.argtypes = [vp]
    L.mcx_graph_clear.argtypes = [vp]
    L.mcx_graph_set_stream.argtypes = [vp, vp]
    L.mcx_graph_add_reads.argtypes = [vp, C.POINTER(ReadBatch)]
    L.mcx_graph_add_reads_pcr.argtypes = [vp, C.POINTER(ReadBatch), vp, vp, C.c_uint64]
    L.mcx_graph_pcr_reset.argtypes = [vp]
    L.mcx_graph_add_str.argtypes = [vp, u32, C.c_char_p, C.c_size_t]
    L.mcx_graph_prepare_host.argtypes = [vp]
    L.mcx_graph_sync.argtypes = [vp, C.POINTER(LoadStats)]
    L.mcx_graph_flush.argtypes = [vp]
    L.mcx_graph_finish_intersect.argtypes = [vp, C.POINTER(u64)]
    L.mcx_graph_load_records.argtypes = [vp, vp, u64, u32, u32, C.POINTER(u32), C.POINTER(u32), u32, u32, C.POINTER(u64), C.POINTER(u64)]
    L.mcx_graph_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.mcx_graph_export_begin.argtypes = [vp, C.c_int, C.POINTER(u64), C.POINTER(u32)]
    L.mcx_graph_export_read.argtypes = [vp, u64, u64, vp]
    L.mcx_graph_export_end.argtypes = [vp]
    L.mcx_kmer_tuples.argtypes = [vp, C.POINTER(ReadBatch), u32, u64, vp, vp, vp]
    L.mcx_graph_insert_tuples.argtypes = [vp, vp, vp, u64, u32]
    L.mcx_graph_add_reads_sharded.argtypes = [vp, C.POINTER(ReadBatch), u32, u32, u64, vp, vp, vp]
    L.mcx_graph_flush_sharded.argtypes = [vp, u32, u32, u64, vp, vp, vp]
    L.mcx_graph_insert_tuples_n.argtypes = [vp, vp, vp, vp, u64, u32]
    L.mcx_graph_insert_tuples_on.argtypes = [vp, vp, vp, vp, vp, u64, u32]
    L.mcx_graph_add_reads_routed.argtypes = [vp, C.POINTER(ReadBatch), u32, u32, u64, C.POINTER(vp), C.POINTER(vp), vp]
    L.mcx_graph_flush_routed.argtypes = [vp, u32, u32, u64, C.POINTER(vp), C.POINTER(vp), vp]
    L.mcx_device_alloc.argtypes = [C.c_int, C.c_size_t, C.POINTER(vp)]
    L.mcx_device_free.argtypes = [C.c_int, vp]
    L.mcx_ipc_export.argtypes = [vp, C.c_char_p]
    L.mcx_ipc_open.argtypes = [C.c_int, C.c_char_p, C.POINTER(vp)]
    L.mcx_ipc_close.argtypes = [C.c_int, vp]
    L.mcx_sort_records.argtypes = [C.c_int, u32, u32, vp, u64, vp]
    L.mcx_key_owner.restype = u32
    L.mcx_key_owner.argtypes = [C.POINTER(u64), u32, u32]
    _lib = L
    return L


def _ck(code, what):
    if code != 0:
        raise McxError(code, what, (lib().mcx_last_error() or b"").decode(errors="replace"))


def device_count():
    return int(lib().mcx_device_count())


def host_alloc(nbytes):
    """Pinned host buffer (address) for read batches: H2D without a staging copy."""
    p = C.c_void_p()
    _ck(lib().mcx_host_alloc(C.byref(p), nbytes), "mcx_host_alloc")
    return p.value


def host_free(addr):
    _ck(lib().mcx_host_free(C.c_void_p(addr)), "mcx_host_free")


def device_alloc(device, nbytes):
    """cudaMalloc'ed, zeroed device buffer (address) that a peer process can map through CUDA IPC"""
    p = C.c_void_p()
    _ck(lib().mcx_device_alloc(device, nbytes, C.byref(p)), "mcx_device_alloc")
    return p.value


def device_free(device, addr):
    _ck(lib().mcx_device_free(device, C.c_void_p(addr)), "mcx_device_free")


def ipc_export(addr):
    h = C.create_string_buffer(64)
    _ck(lib().mcx_ipc_export(C.c_void_p(addr), h), "mcx_ipc_export")
    return h.raw


def ipc_open(device, handle):
    p = C.c_void_p()
    _ck(lib().mcx_ipc_open(device, C.create_string_buffer(bytes(handle), 64), C.byref(p)), "mcx_ipc_open")
    return p.value


def ipc_close(device, addr):
    _ck(lib().mcx_ipc_close(device, C.c_void_p(addr)), "mcx_ipc_close")


def sort_records(kmer_size, ncols, records, device=0):
    """mcx_sort_records: packed .ctx records (bytes) -> the same records in ascending key order (bytes)"""
    rb = 8 * ((kmer_size + 31) // 32) + 5 * ncols
    assert len(records) % rb == 0
    src = C.create_string_buffer(bytes(records), max(len(records), 1))
    dst = C.create_string_buffer(max(len(records), 1))
    _ck(lib().mcx_sort_records(device, kmer_size, ncols, src, len(records) // rb, dst), "mcx_sort_records")
    return dst.raw[:len(records)]


def key_owner(key_words, k, nparts):
    arr = (C.c_uint64 * len(key_words))(*key_words)
    return int(lib().mcx_key_owner(arr, k, nparts))


class Graph:
    """Device-resident coloured de Bruijn graph shard (reference: dBGraph, src/graph/db_graph.h:23-56)."""

    def __init__(self, kmer_size, ncols=1, capacity=1 << 20, device=0, flags=0):
        self.k, self.ncols, self.capacity, self.device = kmer_size, ncols, capacity, device
        self.W = (kmer_size + 31) // 32
        h = C.c_void_p()
        _ck(lib().mcx_graph_create(kmer_size, ncols, capacity, device, flags, C.byref(h)), "mcx_graph_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().mcx_graph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clear(self):
        _ck(lib().mcx_graph_clear(self.h), "mcx_graph_clear")

    def set_stream(self, cuda_stream):
        _ck(lib().mcx_graph_set_stream(self.h, C.c_void_p(cuda_stream or None)), "mcx_graph_set_stream")

    # -- build_graph() ----------------------------------------------------------------
    def _batch(self, seq_addr, nbytes, layout, mem, colour, hp_cutoff, offsets_addr=None, nreads=0,
               qual_addr=None, fq_cutoff=0):
        b = ReadBatch()
        b.seq, b.qual, b.offsets = seq_addr, qual_addr, offsets_addr
        b.nreads, b.nbytes, b.layout, b.mem, b.colour = nreads, nbytes, layout, mem, colour
        b.fq_cutoff, b.hp_cutoff = fq_cutoff, hp_cutoff
        return b

    def add_reads_raw(self, seq_addr, nbytes, layout=MCX_LAYOUT_LINES, mem=MCX_MEM_HOST, colour=0, hp_cutoff=0,
                      offsets_addr=None, nreads=0, qual_addr=None, fq_cutoff=0, must_exist=False):
        b = self._batch(seq_addr, nbytes, layout, mem, colour, hp_cutoff, offsets_addr, nreads, qual_addr, fq_cutoff)
        b.must_exist = 1 if must_exist else 0
        _ck(lib().mcx_graph_add_reads(self.h, C.byref(b)), "mcx_graph_add_reads")

    def add_lines(self, data, colour=0, hp_cutoff=0, qual=None, fq_cutoff=0, must_exist=False):
        """data: bytes in LINES layout (each read followed by one newline), host memory.
        qual: bytes parallel to data (0x7F where a read has no quality); fq_cutoff includes the ASCII offset."""
        buf = C.create_string_buffer(bytes(data), len(data))
        qbuf = C.create_string_buffer(bytes(qual), len(qual)) if qual is not None else None
        assert qbuf is None or len(qual) == len(data)
        self.add_reads_raw(C.addressof(buf), len(data), MCX_LAYOUT_LINES, MCX_MEM_HOST, colour, hp_cutoff,
                           qual_addr=C.addressof(qbuf) if qbuf is not None else None, fq_cutoff=fq_cutoff, must_exist=must_exist)

    def add_reads(self, reads, colour=0, hp_cutoff=0, quals=None, fq_cutoff=0):
        """reads: list of bytes/str; shipped in OFFSETS layout (reads abut + offsets[n+1]).
        quals: optional list of quality strings (padded / cut to the read length here)."""
        reads = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
        blob = b"".join(reads)
        qbuf = None
        if quals is not None:
            qq = []
            for r, q in zip(reads, quals):
                q = (q.encode("latin1") if isinstance(q, str) else bytes(q or b""))[:len(r)]
                qq.append(q + b"\x7f" * (len(r) - len(q)))
            qblob = b"".join(qq)
            qbuf = C.create_string_buffer(qblob, max(len(qblob), 1))
        offs = (C.c_uint64 * (len(reads) + 1))()
        o = 0
        for i, r in enumerate(reads):
            offs[i] = o
            o += len(r)
        offs[len(reads)] = o
        buf = C.create_string_buffer(blob, max(len(blob), 1))
        self.add_reads_raw(C.addressof(buf), len(blob), MCX_LAYOUT_OFFSETS, MCX_MEM_HOST, colour, hp_cutoff,
                           C.addressof(offs), len(reads),
                           qual_addr=C.addressof(qbuf) if qbuf is not None else None, fq_cutoff=fq_cutoff)

    def add_units_pcr(self, units, quals=None, colour=0, fq_cutoff=0, hp_cutoff=0, matedir=1):
        """build --remove-pcr for one batch (mcx_graph_add_reads_pcr): units = [(seq,) | (seq1, seq2), ...] in
        reading order, quals parallel to it (strings, '' = none); matedir 0 FF, 1 FR, 2 RF, 3 RR.
        fq_cutoff includes the ASCII offset.  The graph must have been created with MCX_GRAPH_READSTRT."""
        lines, qlines, mates, offs, o = [], [], bytearray(), [], 0
        for i, u in enumerate(units):
            for m, r in enumerate(u):
                r = r.encode() if isinstance(r, str) else bytes(r)
                q = quals[i][m] if quals is not None else b""
                q = (q.encode("latin1") if isinstance(q, str) else bytes(q or b""))[:len(r)]
                flip = (matedir & 2) if m == 0 else (matedir & 1)
                if q and len(q) < len(r) and flip:
                    q = q + b"." * (len(r) - len(q))   # what seq_read_reverse_complement means to pad with
                lines.append(r + b"\n")
                qlines.append(q + b"\x7f" * (len(r) - len(q) + 1))
                mates.append((0 if len(u) == 1 else 1 + m) | (MCX_MATE_REVCOMP if flip else 0))
                offs.append(o)
                o += len(r) + 1
        offs.append(o)
        blob, qblob = b"".join(lines), b"".join(qlines)
        buf = C.create_string_buffer(blob, max(len(blob), 1))
        qbuf = C.create_string_buffer(qblob, max(len(qblob), 1))
        oarr = (C.c_uint64 * len(offs))(*offs)
        marr = C.create_string_buffer(bytes(mates), max(len(mates), 1))
        b = self._batch(C.addressof(buf), len(blob), MCX_LAYOUT_LINES, MCX_MEM_HOST, colour, hp_cutoff,
                        qual_addr=C.addressof(qbuf) if fq_cutoff else None, fq_cutoff=fq_cutoff)
        _ck(lib().mcx_graph_add_reads_pcr(self.h, C.byref(b), C.addressof(oarr), C.addressof(marr), len(mates)),
            "mcx_graph_add_reads_pcr")

    def pcr_reset(self):
        _ck(lib().mcx_graph_pcr_reset(self.h), "mcx_graph_pcr_reset")

    def prepare_host(self):
        """allocate the pinned staging ring for host batches now (otherwise the first host batch does it)"""
        _ck(lib().mcx_graph_prepare_host(self.h), "mcx_graph_prepare_host")

    def add_str(self, seq, colour=0):
        if isinstance(seq, str):
            seq = seq.encode()
        _ck(lib().mcx_graph_add_str(self.h, colour, seq, len(seq)), "mcx_graph_add_str")

    # -- graph_load() ------------------------------------------------------------------
    def load_records(self, records, file_ncols, from_cols, into_cols, flags=0):
        """records: bytes of packed .ctx records (W x u64 key, file_ncols x u32 covg, file_ncols x u8 edges);
        (from_cols[i], into_cols[i]) = the reference's FileFilter.  Returns (nkmers_loaded, nkmers_novel)."""
        rb = 8 * self.W + 5 * file_ncols
        assert len(records) % rb == 0 and len(from_cols) == len(into_cols)
        buf = C.create_string_buffer(bytes(records), max(len(records), 1))
        fr = (C.c_uint32 * len(from_cols))(*from_cols)
        to = (C.c_uint32 * len(into_cols))(*into_cols)
        nl, nn = C.c_uint64(), C.c_uint64()
        _ck(lib().mcx_graph_load_records(self.h, buf, len(records) // rb, file_ncols, MCX_MEM_HOST, fr, to, len(from_cols),
                                         flags, C.byref(nl), C.byref(nn)), "mcx_graph_load_records")
        return int(nl.value), int(nn.value)

    def finish_intersect(self):
        n = C.c_uint64()
        _ck(lib().mcx_graph_finish_intersect(self.h, C.byref(n)), "mcx_graph_finish_intersect")
        return int(n.value)

    def sync(self):
        st = LoadStats()
        _ck(lib().mcx_graph_sync(self.h, C.byref(st)), "mcx_graph_sync")
        return st

    def flush(self):
        _ck(lib().mcx_graph_flush(self.h), "mcx_graph_flush")

    def stats(self):
        n, cap = C.c_uint64(), C.c_uint64()
        _ck(lib().mcx_graph_stats(self.h, C.byref(n), C.byref(cap)), "mcx_graph_stats")
        return int(n.value), int(cap.value)

    # -- graph_writer: records only (the header belongs to the host driver) --------------
    def export_records(self, sorted=True):
        n, rb = C.c_uint64(), C.c_uint32()
        _ck(lib().mcx_graph_export_begin(self.h, 1 if sorted else 0, C.byref(n), C.byref(rb)), "mcx_graph_export_begin")
        try:
            buf = C.create_string_buffer(max(int(n.value) * int(rb.value), 1))
            _ck(lib().mcx_graph_export_read(self.h, 0, n.value, buf), "mcx_graph_export_read")
            return buf.raw[:int(n.value) * int(rb.value)], int(n.value), int(rb.value)
        finally:
            lib().mcx_graph_export_end(self.h)

    # -- multi-GPU pieces ------------------------------------------------------------------
    def kmer_tuples(self, seq_dev_addr, nbytes, nparts, cap_per_part, keys_addr, masks_addr, counts_addr, hp_cutoff=0):
        b = self._batch(seq_dev_addr, nbytes, MCX_LAYOUT_LINES, MCX_MEM_DEVICE, 0, hp_cutoff)
        _ck(lib().mcx_kmer_tuples(self.h, C.byref(b), nparts, cap_per_part, keys_addr, masks_addr, counts_addr),
            "mcx_kmer_tuples")

    def add_reads_sharded(self, seq_dev_addr, nbytes, nparts, my_part, cap_per_part, keys_addr, meta_addr, counts_addr,
                          hp_cutoff=0, colour=0, qual_dev_addr=None, fq_cutoff=0):
        b = self._batch(seq_dev_addr, nbytes, MCX_LAYOUT_LINES, MCX_MEM_DEVICE, colour, hp_cutoff, qual_addr=qual_dev_addr, fq_cutoff=fq_cutoff)
        _ck(lib().mcx_graph_add_reads_sharded(self.h, C.byref(b), nparts, my_part, cap_per_part, keys_addr, meta_addr,
                                              counts_addr), "mcx_graph_add_reads_sharded")

    def flush_sharded(self, nparts, my_part, cap_per_part, keys_addr, meta_addr, counts_addr):
        _ck(lib().mcx_graph_flush_sharded(self.h, nparts, my_part, cap_per_part, keys_addr, meta_addr, counts_addr),
            "mcx_graph_flush_sharded")

    @staticmethod
    def _ptrs(addrs):
        return (C.c_void_p * len(addrs))(*[C.c_void_p(a or None) for a in addrs])

    def add_reads_routed(self, seq_dev_addr, nbytes, nparts, my_part, cap_per_part, keys_addrs, meta_addrs, counts_addr,
                         hp_cutoff=0, colour=0):
        """keys_addrs / meta_addrs: one device address per destination shard (local or peer-mapped)"""
        b = self._batch(seq_dev_addr, nbytes, MCX_LAYOUT_LINES, MCX_MEM_DEVICE, colour, hp_cutoff)
        _ck(lib().mcx_graph_add_reads_routed(self.h, C.byref(b), nparts, my_part, cap_per_part, self._ptrs(keys_addrs),
                                             self._ptrs(meta_addrs), counts_addr), "mcx_graph_add_reads_routed")

    def flush_routed(self, nparts, my_part, cap_per_part, keys_addrs, meta_addrs, counts_addr):
        _ck(lib().mcx_graph_flush_routed(self.h, nparts, my_part, cap_per_part, self._ptrs(keys_addrs),
                                         self._ptrs(meta_addrs), counts_addr), "mcx_graph_flush_routed")

    def insert_tuples_on(self, cuda_stream, keys_addr, masks_addr, n_dev_addr, n_max, colour=0):
        _ck(lib().mcx_graph_insert_tuples_on(self.h, C.c_void_p(cuda_stream or None), keys_addr, masks_addr, n_dev_addr,
                                             n_max, colour), "mcx_graph_insert_tuples_on")

    def insert_tuples_n(self, keys_addr, masks_addr, n_dev_addr, n_max, colour=0):
        _ck(lib().mcx_graph_insert_tuples_n(self.h, keys_addr, masks_addr, n_dev_addr, n_max, colour),
            "mcx_graph_insert_tuples_n")

    def insert_tuples(self, keys_addr, masks_addr, n, colour=0):
        _ck(lib().mcx_graph_insert_tuples(self.h, keys_addr, masks_addr, n, colour), "mcx_graph_insert_tuples")
