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
        "num_se_reads", "num_pe_reads", "num_good_reads", "num_bad_reads",
        "num_dup_se_reads", "num_dup_pe_pairs")]

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
        L.orc_graph_load_records.restype = C.c_uint64
        L.orc_graph_load_records.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint32),
                                             C.POINTER(C.c_uint32), C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]
        L.orc_graph_merge_file_ginfo.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_char_p, C.c_char_p,
                                                 C.c_char_p, C.c_uint32, C.c_uint32, C.c_char_p]
        L.orc_graph_set_intersect.argtypes = [C.c_void_p, C.c_int]
        L.orc_graph_finish_intersect.argtypes = [C.c_void_p]
        L.orc_graph_wipe_readstrt.argtypes = [C.c_void_p]
        L.orc_graph_add_reads_pcr.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                              C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_size_t,
                                              C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint8, C.c_int, C.POINTER(Stats)]
        L.orc_graph_load_pcr.restype = C.c_long
        L.orc_graph_load_pcr.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_size_t, C.c_uint8, C.c_uint8,
                                         C.c_uint8, C.c_int, C.POINTER(Stats)]
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

    def wipe_readstrt(self):
        lib().orc_graph_wipe_readstrt(self.h)

    def add_reads_pcr(self, seq1, qual1=None, seq2=None, qual2=None, colour=0, fq_cutoff=0, fq_offset=0, hp_cutoff=0,
                      matedir=1, stats=None):
        """one read (seq2 None) or one pair through build_graph_from_reads_mt with --remove-pcr"""
        st = stats if stats is not None else Stats()
        enc = lambda x: x.encode("latin1") if isinstance(x, str) else x
        seq1, qual1, seq2, qual2 = enc(seq1), enc(qual1), enc(seq2), enc(qual2)
        lib().orc_graph_add_reads_pcr(self.h, seq1, len(seq1), qual1, len(qual1) if qual1 else 0,
                                      seq2, len(seq2) if seq2 else 0, qual2, len(qual2) if qual2 else 0,
                                      colour, fq_cutoff, fq_offset, fq_offset, hp_cutoff, matedir, C.byref(st))
        return st

    def load_pcr(self, path1, path2=None, interleaved=False, colour=0, fq_cutoff=0, fq_offset=0, hp_cutoff=0, matedir=1,
                 stats=None):
        st = stats if stats is not None else Stats()
        mode = 1 if path2 else (2 if interleaved else 0)
        r = lib().orc_graph_load_pcr(self.h, path1.encode(), path2.encode() if path2 else None, mode, colour, fq_cutoff,
                                     fq_offset, hp_cutoff, matedir, C.byref(st))
        if r == -1000000000:
            raise IOError("cannot open " + path1)
        return st

    def set_intersect(self, must_exist_reads=True):
        lib().orc_graph_set_intersect(self.h, int(must_exist_reads))

    def finish_intersect(self):
        lib().orc_graph_finish_intersect(self.h)

    def load_ctx(self, spec, into_offset=0, must_exist=False, isec=False, mask_isec=False):
        """graph_load() of `[into:]path[:from]` (file_filter.c syntax); returns (ctx, loaded, novel).
        isec: an intersection graph (flattened into colour 0, edges only, no header metadata);
        mask_isec: edges ANDed with the intersection edge set (must_exist_in_edges)."""
        ctx = CtxFile(spec, into_offset)
        if isec:
            ctx.filter = [(f, 0) for f, _ in ctx.filter]
            ctx.into_ncols = 1
        for fr, into in ([] if isec else ctx.filter):
            h = ctx.ginfo[fr]
            lib().orc_graph_merge_file_ginfo(self.h, into, h["mean"], h["total"], h["name"], h["seq_err"], h["flags"],
                                             h["thr_unitigs"], h["thr_kmers"], h["isec_name"])
        n = len(ctx.records) // ctx.rec_bytes
        fr = (C.c_uint32 * len(ctx.filter))(*[f for f, _ in ctx.filter])
        to = (C.c_uint32 * len(ctx.filter))(*[t for _, t in ctx.filter])
        novel = C.c_uint64(0)
        flags = (1 if must_exist else 0) | (2 if isec else 0) | (4 if mask_isec else 0)
        loaded = lib().orc_graph_load_records(self.h, ctx.records, n, ctx.ncols, fr, to, len(ctx.filter), flags,
                                              C.byref(novel))
        return ctx, int(loaded), int(novel.value)

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


def _parse_range(s, range_max, count_only=False):
    """src/basic/range.c: '1,3-5' -> [1,3,4,5]; '' -> 0..range_max.  A descending range a-b does NOT stop at b in the
    reference (range.c:67-68: `for(j = start; j <= start; j--)`): it emits a, a-1, ..., 0 -- while range_get_num
    (:39-55) counts |a-b|+1 entries; count_only returns that count.  [probed against the compiled reference]"""
    out = []
    for part in [p for p in s.split(",") if p != ""]:
        if part == "*":
            out += list(range(0, range_max + 1))
            continue
        a, _, b = part.partition("-")
        a = int(a)
        b = int(b) if b != "" else a
        if a > range_max or b > range_max:
            raise ValueError("Invalid filter path")
        if count_only:
            out += [0] * (abs(a - b) + 1)
        else:
            out += list(range(a, b + 1)) if a <= b else list(range(a, -1, -1))
    return out if out else list(range(0, range_max + 1))


class CtxFile:
    """A .ctx file + colour filter: graph_file_open2 (graph_file_reader.c:272-322), file_filter_set_cols
    (file_filter.c:77-153).  Test infrastructure: reads the whole file."""

    def __init__(self, spec, into_offset=0):
        import re
        import struct
        m = re.match(r"^(?:([0-9,\-]+):)?(.*?)(?::([0-9,\-]*))?$", spec)
        into_f, self.path, from_f = m.group(1), m.group(2), m.group(3)
        data = open(self.path, "rb").read()
        assert data[:6] == b"CORTEX"
        self.version, self.k, self.W, self.ncols = struct.unpack_from("<IIII", data, 6)
        off = 22
        means = struct.unpack_from("<%dI" % self.ncols, data, off); off += 4 * self.ncols
        totals = struct.unpack_from("<%dQ" % self.ncols, data, off); off += 8 * self.ncols
        names = []
        for _ in range(self.ncols):
            (ln,) = struct.unpack_from("<I", data, off); off += 4
            names.append(data[off:off + ln]); off += ln
        errs = [data[off + 16 * i: off + 16 * i + 16] for i in range(self.ncols)]; off += 16 * self.ncols
        self.ginfo = []
        for i in range(self.ncols):
            flags = data[off:off + 4]; tu, tk, ln = struct.unpack_from("<III", data, off + 4); off += 16
            isec = data[off:off + ln]; off += ln
            self.ginfo.append(dict(mean=means[i], total=totals[i], name=names[i].split(b"\0")[0], seq_err=errs[i], flags=flags,
                                   thr_unitigs=tu, thr_kmers=tk, isec_name=isec.split(b"\0")[0]))
        assert data[off:off + 6] == b"CORTEX"
        off += 6
        self.hdr_size = off
        self.rec_bytes = 8 * self.W + 5 * self.ncols
        self.records = data[off:off + (len(data) - off) // self.rec_bytes * self.rec_bytes]
        # file_filter_set_cols (file_filter.c:94-138): the FIRST range_get_num() entries of the parsed `from` list are used
        frm = _parse_range(from_f or "", self.ncols - 1)[:len(_parse_range(from_f or "", self.ncols - 1, count_only=True))]
        if into_f is not None:
            into = _parse_range(into_f, 1 << 30)
            if len(into) == 1:
                into = into * len(frm)
            assert len(into) == len(frm), "Invalid filter path"
        else:
            into = [into_offset + i for i in range(len(frm))]
        self.filter = sorted(zip(frm, into), key=lambda p: p[1])
        self.into_ncols = max(i for _, i in self.filter) + 1


MAX_IO_THREADS = 10  # src/global/global.h:41


def build_ctx_args(k, args, capacity=1 << 20):
    """Oracle equivalent of `mccortex build -k K -S <args>` for the argument forms
    `-g [into:]in.ctx[:from]`, `-s name`, `-1 file` (and -Q/-O/-H before a file), in command-line order:
    colour bookkeeping of ctx_build.c:158-196, graphs loaded before reads (:362-377), --sample names
    overwrite the names of the colours they name (:379-382), stats credited per batch of <= 10 tasks (Q1)."""
    intocolour, sample_named = -1, False
    names, tasks, graphs = [], [], []
    isecs = [args[i + 1] for i, a in enumerate(args) if a in ("-I", "--intersect")]
    fq_cutoff = fq_offset = hp_cutoff = 0
    remove_pcr, matedir = False, 1   # SEQ_LOADING_PREFS_INIT: READPAIR_FR
    it = iter(args)
    for a in it:
        if a in ("-s", "--sample"):
            intocolour += 1
            names.append((intocolour, next(it)))
            sample_named = True
        elif a in ("-1", "--seq", "-2", "--seq2", "-i", "--seqi"):
            t = dict(path=next(it), path2=None, interleaved=a in ("-i", "--seqi"), fq_cutoff=fq_cutoff, fq_offset=fq_offset,
                     hp_cutoff=hp_cutoff, remove_pcr=remove_pcr, matedir=matedir)
            if a in ("-2", "--seq2"):
                t["path"], t["path2"] = t["path"].split(":")
                if not remove_pcr:   # add_task, ctx_build.c:99-118: two single-end tasks (quirk Q5)
                    tasks.append((intocolour, dict(t, path2=None)))
                    t = dict(t, path=t["path2"], path2=None)
            tasks.append((intocolour, t))
        elif a in ("-p", "--remove-pcr"):
            remove_pcr = True
        elif a in ("-P", "--keep-pcr"):
            remove_pcr = False
        elif a in ("-M", "--matepair"):
            matedir = ["FF", "FR", "RF", "RR"].index(next(it))
        elif a in ("-Q", "--fq-cutoff"):
            fq_cutoff = int(next(it))
        elif a in ("-O", "--fq-offset"):
            fq_offset = int(next(it))
        elif a in ("-H", "--cut-hp"):
            hp_cutoff = int(next(it))
        elif a in ("-g", "--graph"):
            if intocolour == -1:
                intocolour = 0
            spec = next(it)
            ctx = CtxFile(spec, intocolour)
            assert ctx.k == k
            intocolour = max(intocolour, ctx.into_ncols - 1)
            graphs.append((spec, ctx.filter[0][1] if False else None, intocolour, ctx))
            sample_named = False
        elif a in ("-I", "--intersect"):
            next(it)
        else:
            raise ValueError("unsupported build argument for the oracle: " + a)
    ncols = intocolour + (1 if sample_named else 0)
    g = Graph(k, ncols, capacity)
    if isecs:   # ctx_build.c:341-361
        g.set_intersect(True)
        for spec in isecs:
            g.load_ctx(spec, 0, isec=True)
    # into_offset of each graph = intocolour at the time its -g was parsed
    intocolour = -1
    it = iter(args)
    for a in it:
        if a in ("-s", "--sample"):
            intocolour += 1; next(it)
        elif a in ("-g", "--graph"):
            if intocolour == -1:
                intocolour = 0
            spec = next(it)
            ctx, _, _ = g.load_ctx(spec, intocolour, must_exist=bool(isecs), mask_isec=bool(isecs))
            intocolour = max(intocolour, ctx.into_ncols - 1)
        elif a in ("-1", "--seq", "-2", "--seq2", "-i", "--seqi", "-Q", "--fq-cutoff", "-O", "--fq-offset", "-H", "--cut-hp",
                   "-I", "--intersect", "-M", "--matepair"):
            next(it)
    for col, name in names:
        g.set_name(col, name)
    # ctx_build.c:384-407: with --remove-pcr anywhere, one build_graph() call per run of <= 10 tasks of one
    # colour and the read-start marks are wiped when the colour changes; the files of one call are read in
    # command-line order here (the reference reads them concurrently)
    remove_pcr_used = any(t["remove_pcr"] for _, t in tasks)
    start, prev_col = 0, 0
    while start < len(tasks):
        col0 = tasks[start][0]
        end = min(start + MAX_IO_THREADS, len(tasks))
        if remove_pcr_used:
            if col0 != prev_col:
                g.wipe_readstrt()
            end = start + 1
            while end < len(tasks) and end - start < MAX_IO_THREADS and tasks[end][0] == col0:
                end += 1
        st = Stats()
        for col, t in tasks[start:end]:
            if t["remove_pcr"]:
                g.load_pcr(t["path"], t["path2"], t["interleaved"], col, t["fq_cutoff"], t["fq_offset"], t["hp_cutoff"],
                           t["matedir"], st)
            else:
                g.load_file(t["path"], col, t["fq_cutoff"], t["fq_offset"], t["hp_cutoff"], st)
        g.update_ginfo(col0, st)
        start, prev_col = end, col0
    if isecs:   # ctx_build.c:409-413
        g.finish_intersect()
    out = g.dump_sorted()
    full = g.full
    g.close()
    if full:
        raise RuntimeError("Hash table is full")
    return out


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
