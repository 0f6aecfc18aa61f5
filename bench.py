#!/usr/bin/env python
"""bench.py -- k-mers/s inserted during `build` (k=31) on B200, per BASELINE.json.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--reads R]

Own arm (default): one "step" = one whole `build` of the workload (BASELINE.json configs[1]:
50M x 150 bp synthetic reads of a 4.6 Mbp genome, 0.1 % substitutions, k=31, 1 colour).
  value : zero the table, insert every k-mer occurrence; inputs resident in HBM (LINES layout),
          CUDA events on the launching stream.
  e2e   : the whole job through the C ABI from a pinned HOST buffer: H2D of the reads, insert,
          sorted export (mcx_graph_export_begin(sorted=1)) and D2H of every .ctx record into host
          memory, all inside the timed region.
  extra : cli (one `mccortex-b200 build -S` process, FASTA on tmpfs -> sorted .ctx), config5 (the cold-table
          regime of configs[4]: a 1/32 slice of 600M reads of a 3 Gbp genome on this GPU, and the insert kernel
          alone on a table >> L2), other_configs (configs[2] k=63, configs[3] four colours; kernel only).
Reference arm (--impl reference): the compiled, unmodified reference (oracle/_ref/mccortex31
build) on the box's host cores, each step a bounded sample (2M reads) of the same workload.

Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 31
READ_LEN = 150
GENOME = int(os.environ.get("MCX_BENCH_GENOME", 4_600_000))   # env overrides are for experiments only
P_ERR = float(os.environ.get("MCX_BENCH_PERR", 0.001))
DEFAULT_READS = 50_000_000
B_ALG = 19.25  # algorithmic bytes per k-mer occurrence, k=31 single GPU (SURVEY 8d / DESIGN.md)
B_ALG_MULTI = 43.25


def synth_lib():
    p = os.path.join(ROOT, "mccortex_b200", "lib", "libmcxsynth.so")
    L = C.CDLL(p)
    L.mcx_synth_genome.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
    L.mcx_synth_reads.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64,
                                  C.c_double, C.c_int, C.c_uint64]
    L.mcx_synth_read_bytes.restype = C.c_uint64
    L.mcx_synth_read_bytes.argtypes = [C.c_uint32, C.c_int]
    return L


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu summary, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def write_fasta_sample(path, first, nreads, genome_buf, SL):
    stride = SL.mcx_synth_read_bytes(READ_LEN, 1)
    block = 1 << 18
    buf = C.create_string_buffer(int(block * stride))
    with open(path, "wb") as f:
        at = 0
        while at < nreads:
            m = min(block, nreads - at)
            SL.mcx_synth_reads(buf, first + at, m, READ_LEN, genome_buf, GENOME, P_ERR, 1, 0)
            f.write(buf.raw[:int(m * stride)])
            at += m


def run_reference_build(fasta, nreads, threads, tmpdir):
    """one `mccortex31 build` of the sample; returns (seconds, k-mer occurrences)"""
    from oracle import oracle as O
    exe = O.ref_binary(K)
    out = os.path.join(tmpdir, "ref.ctx")
    # capacity as the reference would size it: distinct <= genome + 31 * expected errors, / 0.75
    est = int((GENOME + nreads * READ_LEN * P_ERR * K * 1.1) / 0.75)
    args = [exe, "build", "-f", "-t", str(threads), "-m", "60G", "-n", str(est), "-k", str(K),
            "--sample", "s", "--seq", fasta, out]
    t0 = time.perf_counter()
    r = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("reference build failed: " + r.stderr[-2000:])
    nk = nreads * (READ_LEN - K + 1)
    for line in r.stderr.splitlines():
        if "num kmers:" in line:
            nk = int(line.split("num kmers:")[1].split()[0].replace(",", ""))
    try:
        os.remove(out)
    except OSError:
        pass
    return dt, nk


REF_SAMPLE_READS = 2_000_000   # reads per reference step: 240 M k-mer occurrences at 87x coverage, ~10 s on 16 cores


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    if O.ref_binary(K) is None:
        O.build(ref=True)
    SL = synth_lib()
    genome = C.create_string_buffer(GENOME)
    SL.mcx_synth_genome(genome, GENOME, 0)
    nreads = min(REF_SAMPLE_READS, args.reads)
    threads = min(os.cpu_count() or 1, 32)
    tmpdir = tempfile.mkdtemp(prefix="mcxref", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    fasta = os.path.join(tmpdir, "sample.fa")
    write_fasta_sample(fasta, 0, nreads, genome, SL)
    for _ in range(args.warmup):
        run_reference_build(fasta, nreads, threads, tmpdir)
    tot, nk_tot = 0.0, 0
    for _ in range(args.steps):
        dt, nk = run_reference_build(fasta, nreads, threads, tmpdir)
        tot += dt; nk_tot += nk
    os.remove(fasta); os.rmdir(tmpdir)
    val = nk_tot / tot
    sample = "first %d reads (%d k-mer occurrences/step) of the workload; whole `mccortex31 build` process, unsorted dump to tmpfs" % (
        nreads, nk_tot // max(1, args.steps))
    line = {
        "impl": "reference", "metric": "kmers_per_sec_build_k31", "value": val, "unit": "k-mers/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": dict(workload_config(args, 1), reads_per_step=nreads,
                       note="each reference step builds the first %d reads of the workload (a bounded sample); "
                            "the own arm builds all %d" % (nreads, args.reads)),
        "cpu_baseline": {"value": val, "unit": "k-mers/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    return {"workload": "configs[1]: %d x %d bp synthetic reads/GPU, genome %d bp, p_err %g, k=%d, 1 colour" % (
                args.reads, READ_LEN, GENOME, P_ERR, K),
            "reads_per_gpu": args.reads, "read_len": READ_LEN, "kmer": K, "colours": 1,
            "l2": "inputs (%.1f GB reads, table >> 126 MB L2) exceed L2; table re-zeroed every step" % (
                args.reads * (READ_LEN + 1) / 1e9),
            "parallelism": "1 GPU fused kernel" if n == 1 else "%d GPUs: hash-partitioned shards, all-to-all of (key,mask) tuples" % n}


def own_arm(args):
    import torch
    import mccortex_b200 as M

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist = None
        torch.cuda.set_device(0)
    dev = torch.device("cuda", local)
    assert M.device_count() > 0, "bench needs a CUDA device: there is no CPU fallback"

    if world > 1:
        from mccortex_b200.multi import bench_multi
        return bench_multi(args, rank, world, local, dist)

    # ---- synthetic workload: generated on the host into PINNED memory (e2e source), copied once to HBM
    SL = synth_lib()
    R = args.reads
    stride = READ_LEN + 1
    nbytes = R * stride
    genome = C.create_string_buffer(GENOME)
    SL.mcx_synth_genome(genome, GENOME, 0)
    t0 = time.perf_counter()
    host = M.host_alloc(nbytes + 4096)
    SL.mcx_synth_reads(host, 0, R, READ_LEN, genome, GENOME, P_ERR, 0, 0)
    t_gen = time.perf_counter() - t0
    occ_per_step = R * (READ_LEN - K + 1)

    dseq = torch.empty(nbytes + 4096, dtype=torch.uint8, device=dev)
    harr = (C.c_uint8 * nbytes).from_address(host)
    hview = torch.frombuffer(harr, dtype=torch.uint8)
    dseq[:nbytes].copy_(hview)
    torch.cuda.synchronize()

    distinct_est = int(GENOME + R * READ_LEN * P_ERR * K * 1.05)
    capacity = int(distinct_est / 0.75)
    g = M.Graph(K, 1, capacity, device=local)
    stream = torch.cuda.Stream(device=dev)  # a real (non-null) stream: the library orders all its work on it
    torch.cuda.set_stream(stream)
    g.set_stream(stream.cuda_stream)

    def step_device():
        g.clear()
        g.add_reads_raw(dseq.data_ptr(), nbytes, M.MCX_LAYOUT_LINES, M.MCX_MEM_DEVICE)
        g.flush()

    # ---- warm-up
    for _ in range(args.warmup):
        step_device()
    st = g.sync()
    # (clear() also zeroes the counters, so they hold the last step only)
    assert args.warmup == 0 or st.num_kmers_loaded == occ_per_step, (st.as_dict(), occ_per_step)

    sampler = ClockSampler(local)
    sampler.start()
    # ---- value: device-resident inputs, CUDA events on the launching stream
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    ev0.record(stream)
    for i in range(args.steps):
        g.clear()
        kev[i][0].record(stream)
        g.add_reads_raw(dseq.data_ptr(), nbytes, M.MCX_LAYOUT_LINES, M.MCX_MEM_DEVICE)
        kev[i][1].record(stream)
        g.flush()  # front table -> big table: the step ends with the whole graph in the big table
    ev1.record(stream)
    torch.cuda.synchronize()
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    st = g.sync()
    assert st.num_kmers_loaded == occ_per_step, (st.as_dict(), occ_per_step)
    value = occ_per_step * args.steps / (ms_total * 1e-3)

    # ---- e2e: the whole job through the C ABI.  Pinned host reads -> insert -> sorted export -> every .ctx record back
    #      in (pinned) host memory, all inside the timed region; wall clock and CUDA events, the larger one counts.
    def step_e2e(dst, dst_bytes):
        g.clear()
        g.add_reads_raw(host, nbytes, M.MCX_LAYOUT_LINES, M.MCX_MEM_HOST)
        nrec, rb = C.c_uint64(), C.c_uint32()
        M.binding._ck(M.lib().mcx_graph_export_begin(g.h, 1, C.byref(nrec), C.byref(rb)), "export_begin")
        nb_out = int(nrec.value) * int(rb.value)
        assert nb_out <= dst_bytes, (nb_out, dst_bytes)
        M.binding._ck(M.lib().mcx_graph_export_read(g.h, 0, nrec.value, C.c_void_p(dst)), "export_read")
        M.lib().mcx_graph_export_end(g.h)
        return int(nrec.value), int(rb.value)

    distinct = g.stats()[0]
    out_bytes = (distinct + (distinct >> 6) + 1024) * 13
    hout = M.host_alloc(out_bytes)
    e_steps = max(1, min(args.steps, 3))
    nrec_e, rb_e = step_e2e(hout, out_bytes)  # warm the staging ring and the allocator
    torch.cuda.synchronize()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record(stream)
    t0 = time.perf_counter()
    for _ in range(e_steps):
        nrec_e, rb_e = step_e2e(hout, out_bytes)
    ee1.record(stream)
    torch.cuda.synchronize()
    e2e_wall = time.perf_counter() - t0
    e2e_ms = ee0.elapsed_time(ee1)
    e2e_val = occ_per_step * e_steps / (max(e2e_ms * 1e-3, e2e_wall))
    st_h = g.sync()
    assert st_h.num_kmers_loaded == occ_per_step and nrec_e == distinct, (st_h.as_dict(), nrec_e, distinct)
    # the records that came back are a sorted .ctx body: strictly ascending keys (cheap check on a sample), md5 for the record
    import hashlib
    rec_view = (C.c_uint8 * (nrec_e * rb_e)).from_address(hout)
    md5_records = hashlib.md5(rec_view).hexdigest() if not args.no_extras else None
    import struct
    probe = [struct.unpack_from("<Q", rec_view, i * rb_e)[0] for i in range(0, nrec_e, max(1, nrec_e // 4096))]
    assert all(x < y for x, y in zip(probe, probe[1:])), "export is not in ascending key order"
    clocks = sampler.stop()
    M.host_free(hout)

    # ---- CPU baseline: the compiled reference on a bounded sample, host cores of this box
    cpu = None
    if not args.no_cpu_baseline:
        try:
            from oracle import oracle as O
            if O.ref_binary(K) is not None:
                n_s = REF_SAMPLE_READS if R >= REF_SAMPLE_READS else R
                threads = min(os.cpu_count() or 1, 32)
                tmpdir = tempfile.mkdtemp(prefix="mcxref", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
                fasta = os.path.join(tmpdir, "sample.fa")
                write_fasta_sample(fasta, 0, n_s, genome, SL)
                run_reference_build(fasta, n_s, threads, tmpdir)   # untimed: the first run of the binary on a box is ~2x slower
                dt, nk = run_reference_build(fasta, n_s, threads, tmpdir)
                os.remove(fasta); os.rmdir(tmpdir)
                cpu = {"value": nk / dt, "unit": "k-mers/s", "cores": threads, "kind": "reference",
                       "sample": "first %d reads (%d k-mer occurrences) of the workload, `mccortex31 build -t %d`, whole process %.1f s "
                                 "(second of two runs)" % (n_s, nk, threads, dt)}
        except Exception as ex:  # the baseline is informative; never lose the GPU numbers over it
            cpu = {"value": None, "unit": "k-mers/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % ex}

    g.close()
    extra = {"distinct_kmers": distinct, "table_slots": capacity, "host_gen_s": t_gen,
             "export_records": nrec_e, "export_md5": md5_records,
             "e2e_breakdown": "H2D %.2f GB + insert + sorted export + D2H %.2f GB per step" % (nbytes / 1e9, nrec_e * rb_e / 1e9)}
    if not args.no_extras:
        for name, fn in (("cli", lambda: extra_cli(args, SL, genome, occ_per_step)),
                         ("other_configs", lambda: extra_other_configs(M, torch, dseq, nbytes, R, stream)),
                         ("config5", lambda: extra_config5(args, M, torch, SL, dev, stream))):
            try:
                if name == "config5":
                    del dseq
                    torch.cuda.empty_cache()
                extra[name] = fn()
            except Exception as ex:  # extras never cost the headline numbers
                extra[name] = {"failed": "%s: %s" % (type(ex).__name__, ex)}

    peak, peak_src = measured_peaks()
    achieved = occ_per_step * B_ALG / (kernel_ms * 1e-3) / 1e9
    tr = ncu_traffic()
    # dram bytes per launch: the ncu --set full capture is of a shorter launch of the same kernel on
    # the same workload (a 50M-read launch x ~40 replays does not fit a profiling call); its
    # bytes per k-mer occurrence are scaled to the launch timed here
    traffic = None
    if tr:
        traffic = tr.get("dram_bytes_per_launch") or (tr.get("dram_bytes_per_kmer") or 0) * occ_per_step or None
    line = {
        "metric": "kmers_per_sec_build_k31", "value": value, "unit": "k-mers/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, 1),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     # physical DRAM utilisation next to the logical (algorithmic) fraction
                     "dram_util": (traffic / (kernel_ms * 1e-3) / 1e9 / peak) if traffic else None,
                     "kernel": "mcx_build_fused_kernel<1>", "kernel_ms": kernel_ms,
                     "alg_bytes_per_kmer": B_ALG, "kmers_per_launch": occ_per_step,
                     "traffic_note": (tr or {}).get("note")},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_val, "unit": "k-mers/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nrec_e * rb_e + 64 + 72,
                "steps": e_steps, "ms_per_step": 1e3 * max(e2e_ms * 1e-3, e2e_wall) / e_steps},
        # per step in the `value` region: one mcx_build_fused_kernel per span of <= 0xEF000000 positions (the front
        # table's 32-bit counters are merged into the big table between spans) + as many mcx_front_flush_kernel
        "gpu_launches": 2 * args.steps * (-(-nbytes // 0xEF000000)),
        "clocks": clocks,
        "extra": extra,
    }
    print(json.dumps(line), flush=True)
    M.host_free(host)


def _time_build(M, torch, g, stream, batches, iters=3):
    """best of `iters`: clear, add the device-resident batches [(addr, nbytes, colour)], flush; returns (ms, stats)"""
    best = 1e30
    for _ in range(iters):
        g.clear(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for addr, nb, col in batches:
            g.add_reads_raw(addr, nb, M.MCX_LAYOUT_LINES, M.MCX_MEM_DEVICE, colour=col)
        g.flush()
        e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, g.sync()


def extra_other_configs(M, torch, dseq, nbytes, R, stream):
    """configs[2] (k=63) and configs[3] (four colours) on the same reads, kernel only: parity cases (tests/) whose speed
    is worth knowing, not bench lines"""
    peak, _ = measured_peaks()
    stride = READ_LEN + 1
    out = {}
    for label, k, ncols, balg in (("configs[2] k=63, 1 colour", 63, 1, 27.70), ("configs[3] k=31, 4 colours (4 samples x R/4 reads)", 31, 4, 19.25)):
        occ = R * (READ_LEN - k + 1)
        cap = int((GENOME + R * READ_LEN * P_ERR * k * 1.05) / 0.75)
        g = M.Graph(k, ncols, cap); g.set_stream(stream.cuda_stream)
        per = R // ncols * stride
        batches = [(dseq.data_ptr() + c * per, per if c + 1 < ncols else nbytes - c * per, c) for c in range(ncols)]
        ms, st = _time_build(M, torch, g, stream, batches)
        assert st.num_kmers_loaded == occ, (label, st.num_kmers_loaded, occ)
        out[label] = {"value": occ / (ms * 1e-3), "unit": "k-mers/s", "ms_per_step": ms, "reads": R, "distinct_kmers": g.stats()[0],
                      "roofline_frac": occ * balg / (ms * 1e-3) / 1e9 / peak, "alg_bytes_per_kmer": balg}
        g.close()
    # the production pipeline's flags (scripts/make-pipeline.pl:342-346: --fq-cutoff 10): the same reads with a quality string
    # per read (Phred+33: 'F' everywhere, 1 % of the bases at or below the cut-off), quality bytes device-resident next to the
    # bases; two kernels per launch (carry summary + insert).  Rate = positions in contigs actually loaded per second.
    k = K
    qual = torch.full((nbytes + 4096,), 70, dtype=torch.uint8, device=dseq.device)
    low = torch.rand(nbytes, device=dseq.device) < 0.01
    qual[:nbytes][low] = 40
    del low
    g = M.Graph(k, 1, int((GENOME + R * READ_LEN * P_ERR * k * 1.05) / 0.75)); g.set_stream(stream.cuda_stream)
    best = 1e30
    for _ in range(3):
        g.clear(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step = (1 << 30) // (16 * stride) * (16 * stride)   # launches start at read boundaries and at 16-byte aligned addresses
        for lo in range(0, nbytes, step):
            g.add_reads_raw(dseq.data_ptr() + lo, min(step, nbytes - lo), M.MCX_LAYOUT_LINES, M.MCX_MEM_DEVICE,
                            qual_addr=qual.data_ptr() + lo, fq_cutoff=43)
        g.flush()
        e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    st = g.sync()
    out["configs[1] with --fq-cutoff 10 (1 % of the bases at or below the cut-off)"] = {
        "value": st.num_kmers_loaded / (best * 1e-3), "unit": "k-mers/s", "ms_per_step": best, "reads": R, "kmers_loaded": st.num_kmers_loaded,
        "positions_per_s": nbytes / (best * 1e-3), "roofline_frac": st.num_kmers_loaded * 19.25 / (best * 1e-3) / 1e9 / peak}
    g.close()
    return out


def extra_cli(args, SL, genome, occ):
    """the whole `mccortex-b200 build -S` process on the workload: FASTA on tmpfs -> sorted .ctx on tmpfs"""
    import shutil
    import mccortex_b200 as M
    R = args.reads
    need = R * (READ_LEN + 4) + R * 40   # FASTA + output
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > need * 1.2 else None
    tmpdir = tempfile.mkdtemp(prefix="mcxcli", dir=base)
    try:
        fasta, out = os.path.join(tmpdir, "reads.fa"), os.path.join(tmpdir, "out.ctx")
        write_fasta_sample(fasta, 0, R, genome, SL)
        est = int((GENOME + R * READ_LEN * P_ERR * K * 1.05) / 0.75)
        cmd = [M.driver_path(), "build", "-f", "-q", "-S", "-m", "100G", "-n", str(est), "-k", str(K), "--sample", "s", "--seq", fasta, out]
        t0 = time.perf_counter()
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"failed": r.stderr[-500:]}
        return {"wall_s": wall, "kmers_per_s": occ / wall, "fasta_bytes": os.path.getsize(fasta), "ctx_bytes": os.path.getsize(out),
                "command": "mccortex-b200 build -S -k 31 -n %d --sample s --seq reads.fa out.ctx (tmpfs: %s)" % (est, bool(base))}
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)


def extra_config5(args, M, torch, SL, dev, stream):
    """the cold-table regime of configs[4] (600M x 150 bp reads of a 3 Gbp genome over 8 GPUs) on ONE GPU: a 1/32 slice of
    the reads into a local table (every k-mer is seen about once: the front table absorbs nothing, every occurrence is a
    random DRAM sector of a table >> L2), and the insert kernel (kernel C) alone on distinct random keys"""
    peak, _ = measured_peaks()
    G5, R5 = int(os.environ.get("MCX_BENCH_G5", 3_000_000_000)), args.config5_reads
    stride = READ_LEN + 1
    nb5 = R5 * stride
    t0 = time.perf_counter()
    genome5 = C.create_string_buffer(G5)
    SL.mcx_synth_genome(genome5, G5, 5)
    host5 = M.host_alloc(nb5 + 4096)
    SL.mcx_synth_reads(host5, 0, R5, READ_LEN, genome5, G5, P_ERR, 0, 5)
    del genome5
    t_gen = time.perf_counter() - t0
    d5 = torch.empty(nb5 + 4096, dtype=torch.uint8, device=dev)
    d5[:nb5].copy_(torch.frombuffer((C.c_uint8 * nb5).from_address(host5), dtype=torch.uint8))
    torch.cuda.synchronize()
    M.host_free(host5)
    occ = R5 * (READ_LEN - K + 1)
    cap = int(occ * 1.02 / 0.75)           # (nearly) every occurrence is a distinct k-mer at 0.9x coverage
    g = M.Graph(K, 1, cap); g.set_stream(stream.cuda_stream)
    ms, st = _time_build(M, torch, g, stream, [(d5.data_ptr(), nb5, 0)], iters=2)
    assert st.num_kmers_loaded == occ, (st.num_kmers_loaded, occ)
    distinct = g.stats()[0]
    out = {"workload": "1/32 slice of configs[4]: %d x %d bp reads of a %d bp genome, p_err %g, k=%d, local table of %d slots (%.0f GB)" % (
               R5, READ_LEN, G5, P_ERR, K, cap, cap * 16 / 1e9),
           "value": occ / (ms * 1e-3), "unit": "k-mers/s", "ms_per_step": ms, "distinct_kmers": distinct,
           "occurrences_per_distinct": occ / max(1, distinct), "tuples_per_step": 0, "nvlink_bytes": 0,
           "frac": occ * B_ALG / (ms * 1e-3) / 1e9 / peak, "alg_bytes_per_kmer": B_ALG,
           "dram_bytes_per_occurrence": (ncu_traffic() or {}).get("config5_dram_bytes_per_kmer"), "host_gen_s": t_gen}
    g.close(); del d5
    torch.cuda.empty_cache()
    # the insert kernel alone: distinct random keys -> a 2^30-slot (17 GB) table, then the same keys again (all found)
    n = 128_000_000
    g = M.Graph(K, 1, 1 << 30); g.set_stream(stream.cuda_stream)
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    keys = torch.randint(0, 1 << 62, (n,), dtype=torch.int64, device=dev, generator=gen)
    meta = torch.full((n,), (1 << 8) | 0x21, dtype=torch.int32, device=dev)
    alg = 8 + 8 + 2 + 12   # key compare + covg RMW + edge RMW + the 12-byte tuple read
    res = {}
    for label in ("novel", "found"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(stream)
        g.insert_tuples(keys.data_ptr(), meta.data_ptr(), n)
        e1.record(stream); torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        res[label] = {"inserts_per_s": n / (t * 1e-3), "ms": t, "logical_gb_s": n * alg / (t * 1e-3) / 1e9,
                      "frac_logical": n * alg / (t * 1e-3) / 1e9 / peak,
                      "frac_physical_64B": n * 64 / (t * 1e-3) / 1e9 / peak}
    stn = g.sync()
    assert stn.num_kmers_novel >= n * 0.999
    g.close()
    out["insert_kernel_cold"] = dict(res, tuples=n, table_slots=1 << 30, alg_bytes_per_insert=alg,
                                     note="mcx_insert_tuples_kernel<1> alone; physical floor = one 32-byte sector each way per insert")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--reads", type=int, default=DEFAULT_READS, help="reads per GPU (default: configs[1] = 50M)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip extra.cli / extra.other_configs / extra.config5")
    ap.add_argument("--config5-reads", type=int, default=18_750_000, help="reads of the configs[4] slice (default: 600M / 32)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        own_arm(args)


if __name__ == "__main__":
    main()
