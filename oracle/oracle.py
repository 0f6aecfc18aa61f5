"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes wrapper over oracle/liboracle.so (the plain-C restatement in
mcx_oracle.c) plus helpers that drive the compiled, unmodified reference
binaries in oracle/_ref/ (built by `make -C oracle ref`).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")


def build(ref=True):
    """Compile liboracle.so and (when /root/reference is present) oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle"])
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "total_bases_read", "total_bases_loaded", "contigs_parsed",
        "num_kmers_loaded", "num_kmers_novel",
        "num_se_reads", "num_pe_reads", "num_good_reads", "num_bad_reads")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(ref=False)
        L = C.CDLL(LIB_PATH)
        L.orc_char_to_nuc.restype = C.c_uint8
        L.orc_char_to_nuc.argtypes = [C.c_char]
        L.orc_lookup3.restype = C.c_uint32
        L.orc_lookup3.argtypes = [C.POINTER(C.c_uint64), C.c_int, C.c_uint32, C.POINTER(C.c_uint32)]
        L.orc_hashtest_xor.restype = C.c_uint32
        L.orc_hashtest_xor.argtypes = [C.c_uint64]
        L.orc_graph_new.restype = C.c_void_p
        L.orc_graph_new.argtypes = [C.c_size_t, C.c_size_t, C.c_uint64]
        L.orc_graph_free.argtypes = [C.c_void_p]
        L.orc_graph_nkmers.restype = C.c_uint64
        L.orc_graph_nkmers.argtypes = [C.c_void_p]
        L.orc_graph_is_full.restype = C.c_int
        L.orc_graph_is_full.argtypes = [C.c_void_p]
        L.orc_graph_set_name.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p]
        L.orc_graph_add_read.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                         C.c_size_t, C.c_uint8, C.c_uint8, C.c_uint8, C.POINTER(Stats)]
        L.orc_graph_add_contig.restype = C.c_size_t
        L.orc_graph_add_contig.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.orc_graph_update_ginfo.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(Stats)]
        L.orc_graph_write_header.restype = C.c_size_t
        L.orc_graph_write_header.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_graph_dump_sorted.restype = C.c_size_t
        L.orc_graph_dump_sorted.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_graph_load_file.restype = C.c_long
        L.orc_graph_load_file.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_uint8, C.c_uint8,
                                          C.c_uint8, C.POINTER(Stats)]
        L.orc_read_windows.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_size_t,
                                       C.c_uint8, C.c_uint8, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_hash_table_cap.restype = C.c_uint64
        L.orc_hash_table_cap.argtypes = [C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)]
        L.orc_hash_table_mem_limit.restype = C.c_uint64
        L.orc_hash_table_mem_limit.argtypes = [C.c_size_t, C.c_size_t, C.POINTER(C.c_uint64)]
        L.orc_guess_fq_offset.restype = C.c_int
        L.orc_guess_fq_offset.argtypes = [C.c_char_p, C.c_size_t]
        _lib = L
    return _lib


def lookup3(words, initval=0):
    """bklk3_hashlittle over W u64 words -> (c, b)."""
    arr = (C.c_uint64 * len(words))(*words)
    b = C.c_uint32(0)
    c = lib().orc_lookup3(arr, len(words), initval, C.byref(b))
    return int(c), int(b.value)


class Graph:
    """In-memory oracle graph (rows A-H of SURVEY.md section 8)."""

    def __init__(self, k, ncols=1, capacity=1 << 20):
        self.k, self.ncols = k, ncols
        self.h = lib().orc_graph_new(k, ncols, capacity)

    def close(self):
        if self.h:
            lib().orc_graph_free(self.h)
            self.h = None

    __del__ = close

    def set_name(self, col, name):
        lib().orc_graph_set_name(self.h, col, name.encode())

    def add_read(self, seq, qual=None, colour=0, fq_cutoff=0, fq_offset=0, hp_cutoff=0, stats=None):
        st = stats if stats is not None else Stats()
        if isinstance(seq, str):
            seq = seq.encode()
        if isinstance(qual, str):
            qual = qual.encode()
        lib().orc_graph_add_read(self.h, seq, len(seq), qual, len(qual) if qual else 0,
                                 colour, fq_cutoff, fq_offset, hp_cutoff, C.byref(st))
        return st

    def load_file(self, path, colour=0, fq_cutoff=0, fq_offset=0, hp_cutoff=0, stats=None):
        st = stats if stats is not None else Stats()
        r = lib().orc_graph_load_file(self.h, path.encode(), colour, fq_cutoff, fq_offset, hp_cutoff, C.byref(st))
        if r == -1000000000:
            raise IOError("cannot open " + path)
        return st

    def update_ginfo(self, colour, stats):
        lib().orc_graph_update_ginfo(self.h, colour, C.byref(stats))

    @property
    def nkmers(self):
        return int(lib().orc_graph_nkmers(self.h))

    @property
    def full(self):
        return bool(lib().orc_graph_is_full(self.h))

    def header(self):
        n = lib().orc_graph_write_header(self.h, None)
        buf = C.create_string_buffer(n)
        lib().orc_graph_write_header(self.h, buf)
        return buf.raw[:n]

    def dump_sorted(self):
        n = lib().orc_graph_dump_sorted(self.h, None)
        buf = C.create_string_buffer(n)
        m = lib().orc_graph_dump_sorted(self.h, buf)
        assert m == n
        return buf.raw[:n]


MAX_IO_THREADS = 10  # src/global/global.h:41


def build_ctx(k, samples, capacity=1 << 20):
    """Oracle equivalent of `mccortex build -k K -S [--sample name --seq file ...]`.

    samples: list of (name, [task, ...]); task = path or dict(path=, fq_cutoff=, fq_offset=, hp_cutoff=).
    Reproduces quirk Q1: per build_graph() call (<=10 consecutive tasks,
    ctx_build.c:389-407) all header stats are credited to the first task's colour.
    Returns (ctx_bytes, [per-batch Stats]).
    """
    g = Graph(k, len(samples), capacity)
    tasks = []
    for col, (name, files) in enumerate(samples):
        g.set_name(col, name)
        for t in files:
            if isinstance(t, str):
                t = dict(path=t)
            tasks.append((col, t))
    batches = []
    for start in range(0, len(tasks), MAX_IO_THREADS):
        st = Stats()
        for col, t in tasks[start:start + MAX_IO_THREADS]:
            g.load_file(t["path"], col, t.get("fq_cutoff", 0), t.get("fq_offset", 0), t.get("hp_cutoff", 0), st)
        g.update_ginfo(tasks[start][0], st)
        batches.append(st)
    out = g.dump_sorted()
    full = g.full
    g.close()
    if full:
        raise RuntimeError("Hash table is full")
    return out, batches


# ---------------------------------------------------------------- reference binaries

def ref_binary(k):
    """Path of the compiled reference binary serving kmer size k, or None."""
    name = "mccortex31" if k <= 31 else "mccortex63"
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None


def ref_run(k, args, check=True, capture=True):
    exe = ref_binary(k)
    if exe is None:
        raise FileNotFoundError("oracle/_ref not built (run `make -C oracle ref` where /root/reference exists)")
    return subprocess.run([exe] + list(args), check=check,
                          stdout=subprocess.PIPE if capture else None,
                          stderr=subprocess.PIPE if capture else None)


def ref_build(k, build_args, out_path, threads=2, nkmers="1M", mem="1G", sort=True):
    """`mccortexNN build -q -f -t T -m M -n N -k K [-S] <build_args> out`"""
    args = ["build", "-q", "-f", "-t", str(threads), "-m", mem, "-n", str(nkmers), "-k", str(k)]
    if sort:
        args.append("-S")
    args += list(build_args) + [out_path]
    ref_run(k, args)
    with open(out_path, "rb") as f:
        return f.read()
