"""Writes tests/golden/: small inputs + the .ctx the COMPILED REFERENCE produces for them.

Run in the build container (needs oracle/_ref, i.e. /root/reference):
    python tests/golden/make_golden.py
The inputs are seeded; the .ctx files and cases.json are committed so the GPU box
(which has no /root/reference) checks against the reference's own bytes.
"""
import hashlib
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from conftest import rand_reads, EDGE_READS, write_pcr_files  # noqa: E402


def write_fa(path, reads, wrap=None):
    with open(path, "w") as f:
        for i, r in enumerate(reads):
            f.write(">r%d\n" % i)
            if wrap:
                for j in range(0, len(r), wrap):
                    f.write(r[j:j + wrap] + "\n")
            else:
                f.write(r + "\n")


def write_fq(path, reads, rng, qoff=33, qlo=2, qhi=40):
    with open(path, "w") as f:
        for i, r in enumerate(reads):
            q = "".join(chr(qoff + (rng.randint(qlo, qhi) if rng.random() < 0.15 else rng.randint(25, qhi))) for _ in r)
            f.write("@r%d\n%s\n+\n%s\n" % (i, r, q))


def main():
    rng = random.Random(424242)
    cases = []

    def add(name, k, samples):
        args = []
        for s in samples:
            args += ["-s", s["name"]]
            for t in s["tasks"]:
                if t.get("fq_cutoff"):
                    args += ["-Q", str(t["fq_cutoff"])]
                if t.get("fq_offset"):
                    args += ["-O", str(t["fq_offset"])]
                if t.get("hp_cutoff"):
                    args += ["-H", str(t["hp_cutoff"])]
                args += ["-1", os.path.join(HERE, t["file"])]
                # -Q/-O/-H persist for following inputs in the reference: reset explicitly
                if t.get("fq_cutoff") or t.get("hp_cutoff") or t.get("fq_offset"):
                    pass
        ctx = name + ".ctx"
        data = O.ref_build(k, args, os.path.join(HERE, ctx), threads=3)
        cases.append(dict(name=name, k=k, samples=samples, ctx=ctx, md5=hashlib.md5(data).hexdigest(),
                          ref_args=[a.replace(HERE + "/", "") for a in args]))
        print(name, len(data), cases[-1]["md5"])

    # tiny 2-colour graph from SURVEY 8c
    open(os.path.join(HERE, "g1.fa"), "w").write(
        ">r1\nACGTACGTTAGCNNACGTTAGCATCGATCGGATCGAT\n>r2\nacgtacgttagc\n>r3\nAC\n>r4\nTTTTTTTTTTTTTTT\nGGGGGGGGGCA\n")
    open(os.path.join(HERE, "g2.fa"), "w").write(">s1\nATCGATCCGATCGATGCTAACGT\n")
    add("tiny_k11_c2", 11, [dict(name="one", tasks=[dict(file="g1.fa")]), dict(name="two", tasks=[dict(file="g2.fa")])])

    write_fa(os.path.join(HERE, "a.fa"), rand_reads(rng, 150, 150, 4000) + EDGE_READS)
    add("reads_k31", 31, [dict(name="s", tasks=[dict(file="a.fa")])])
    add("reads_k63", 63, [dict(name="s", tasks=[dict(file="a.fa")])])
    add("reads_k33", 33, [dict(name="s", tasks=[dict(file="a.fa")])])
    write_fa(os.path.join(HERE, "b.fa"), rand_reads(rng, 60, (20, 300), 4000), wrap=60)
    add("two_colours_k21", 21, [dict(name="x", tasks=[dict(file="a.fa")]), dict(name="y", tasks=[dict(file="b.fa")])])
    add("empty_colour_k31", 31, [dict(name="e", tasks=[]), dict(name="l", tasks=[dict(file="b.fa")])])
    write_fa(os.path.join(HERE, "long.fa"), rand_reads(rng, 1, 20000, 30000, pN=0.0005), wrap=80)
    add("long_record_k31", 31, [dict(name="chr", tasks=[dict(file="long.fa")])])
    add("hp4_k21", 21, [dict(name="h", tasks=[dict(file="a.fa", hp_cutoff=4)])])
    write_fq(os.path.join(HERE, "q.fq"), rand_reads(rng, 150, (30, 200), 4000, perr=0.02), rng)
    add("fq10_k21", 21, [dict(name="q", tasks=[dict(file="q.fq", fq_cutoff=10)])])
    add("fq20_hp5_k15", 15, [dict(name="q", tasks=[dict(file="q.fq", fq_cutoff=20, hp_cutoff=5)])])

    # build --graph: graph files (the goldens above) merged with reads.  Arguments are
    # given raw; "@/" stands for this directory.
    graph_cases = []

    def add_graph(name, k, raw, threads=3):
        args = [a.replace("@/", HERE + "/") for a in raw]
        ctx = name + ".ctx"
        data = O.ref_build(k, args, os.path.join(HERE, ctx), threads=threads)
        graph_cases.append(dict(name=name, k=k, ctx=ctx, md5=hashlib.md5(data).hexdigest(), ref_args=raw))
        print(name, len(data), graph_cases[-1]["md5"])

    add_graph("graph_then_reads_k21", 21, ["-g", "@/two_colours_k21.ctx", "-s", "z", "-1", "@/b.fa"])
    add_graph("graph_from_filter_k21", 21, ["-g", "@/two_colours_k21.ctx:1", "-s", "z", "-1", "@/a.fa"])
    add_graph("graph_flatten_k21", 21, ["-g", "0:@/two_colours_k21.ctx:0,1", "-s", "z", "-1", "@/b.fa"])
    add_graph("graphs_same_colour_k31", 31, ["-g", "@/reads_k31.ctx", "-g", "@/long_record_k31.ctx", "-s", "s2", "-1", "@/a.fa"])
    add_graph("graph_k63", 63, ["-g", "@/reads_k63.ctx", "-s", "t", "-1", "@/b.fa"])
    add_graph("sample_graph_sample_k21", 21, ["-s", "first", "-1", "@/b.fa", "-g", "@/two_colours_k21.ctx:1,0", "-s", "last", "-1", "@/g1.fa"])

    # build --intersect: only k-mers of the intersection graph(s); reads and graphs are looked up, never inserted
    add_graph("isec_reads_k21", 21, ["-I", "@/two_colours_k21.ctx", "-s", "z", "-1", "@/b.fa", "-1", "@/a.fa", "-s", "w", "-1", "@/g1.fa"])
    add_graph("isec_filter_graph_k21", 21, ["-I", "@/two_colours_k21.ctx:1", "-g", "@/hp4_k21.ctx", "-s", "z", "-1", "@/b.fa"])
    add_graph("isec_two_graphs_hp_k21", 21, ["-I", "@/hp4_k21.ctx", "-I", "@/fq10_k21.ctx", "-g", "@/two_colours_k21.ctx", "-s", "z",
                                             "-H", "5", "-1", "@/a.fa", "-1", "@/q.fq"])
    add_graph("isec_long_k31", 31, ["-I", "@/long_record_k31.ctx", "-s", "s2", "-1", "@/a.fa", "-1", "@/long.fa"])
    add_graph("isec_fq10_k15", 15, ["-I", "@/fq20_hp5_k15.ctx", "-s", "z", "-Q", "10", "-1", "@/q.fq"])
    add_graph("isec_fq25_hp4_k15", 15, ["-I", "@/fq20_hp5_k15.ctx", "-s", "z", "-Q", "25", "-H", "4", "-1", "@/q.fq", "-1", "@/a.fa"])
    add_graph("isec_k63", 63, ["-I", "@/reads_k63.ctx", "-s", "t", "-1", "@/b.fa", "-1", "@/a.fa"])

    # build --remove-pcr: the reference with ONE worker thread and one input task per colour is
    # deterministic (reads are tested in file order); that run is the golden.  se.fa / p1.fq + p2.fq / il.fq come
    # from their own seed so the files above keep their bytes.
    pdir = os.path.join(HERE, "pcr")
    os.makedirs(pdir, exist_ok=True)
    write_pcr_files(random.Random(31337), pdir, n=250)
    add_graph("pcr_se_k21", 21, ["-p", "-s", "a", "-1", "@/pcr/se.fa"], threads=1)
    add_graph("pcr_se_rr_k31", 31, ["-p", "-M", "RR", "-s", "a", "-1", "@/pcr/se.fa"], threads=1)
    add_graph("pcr_pe_k21", 21, ["-p", "-s", "a", "-2", "@/pcr/p1.fq:@/pcr/p2.fq"], threads=1)
    add_graph("pcr_pe_fq10_hp4_k21", 21, ["-p", "-Q", "10", "-H", "4", "-s", "a", "-2", "@/pcr/p1.fq:@/pcr/p2.fq"], threads=1)
    add_graph("pcr_pe_ff_k33", 33, ["-p", "-M", "FF", "-s", "a", "-2", "@/pcr/p1.fq:@/pcr/p2.fq"], threads=1)
    add_graph("pcr_pe_rf_fq12_k21", 21, ["-p", "-M", "RF", "-Q", "12", "-s", "a", "-2", "@/pcr/p1.fq:@/pcr/p2.fq"], threads=1)
    add_graph("pcr_il_k21", 21, ["-p", "-s", "a", "-i", "@/pcr/il.fq"], threads=1)
    add_graph("pcr_il_fq10_rr_k63", 63, ["-p", "-Q", "10", "-M", "RR", "-s", "a", "-i", "@/pcr/il.fq"], threads=1)
    add_graph("pcr_three_colours_k21", 21, ["-p", "-s", "a", "-1", "@/pcr/se.fa", "-P", "-s", "b", "-2", "@/pcr/p1.fq:@/pcr/p2.fq",
                                            "-p", "-s", "c", "-i", "@/pcr/il.fq"], threads=1)
    add_graph("pcr_same_file_two_colours_k21", 21, ["-p", "-s", "a", "-1", "@/pcr/se.fa", "-s", "b", "-1", "@/pcr/se.fa"], threads=1)

    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(cases, f, indent=1)
    with open(os.path.join(HERE, "graph_cases.json"), "w") as f:
        json.dump(graph_cases, f, indent=1)


if __name__ == "__main__":
    main()
