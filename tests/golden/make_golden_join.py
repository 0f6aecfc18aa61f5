"""Writes tests/golden/join_*.ctx + join_cases.json: what the COMPILED REFERENCE's `join` makes of golden graph files
that are already committed (two_colours_k21.ctx, fq10_k21.ctx, hp4_k21.ctx, graph_k63.ctx, reads_k63.ctx).

Run in the build container (needs oracle/_ref with join, i.e. /root/reference):
    python tests/golden/make_golden_join.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

CASES = [
    # name, k, join arguments ('@/' = tests/golden/), sorted?
    ("join_side_by_side_k21", 21, ["@/two_colours_k21.ctx", "@/fq10_k21.ctx", "@/hp4_k21.ctx"], True),
    ("join_on_top_k21", 21, ["0:@/two_colours_k21.ctx:1", "0:@/fq10_k21.ctx", "1:@/hp4_k21.ctx", "1:@/two_colours_k21.ctx:0"], True),
    ("join_picked_k21", 21, ["@/two_colours_k21.ctx:1,0", "3:@/hp4_k21.ctx"], True),
    ("join_stream_filter_k21", 21, ["@/two_colours_k21.ctx:1"], False),
    ("join_stream_identity_k21", 21, ["@/two_colours_k21.ctx:0,1", ], False),
    ("join_k63", 63, ["@/graph_k63.ctx", "0:@/reads_k63.ctx"], True),
]


def main():
    out = []
    for name, k, args, sort in CASES:
        path = os.path.join(HERE, name + ".ctx")
        real = [a.replace("@/", HERE + "/") for a in args]
        O.ref_run(k, ["join", "-q", "-f", "-m", "1G", "-n", "1M"] + (["-S"] if sort else []) + ["-o", path] + real)
        data = open(path, "rb").read()
        out.append(dict(name=name, k=k, ctx=name + ".ctx", md5=hashlib.md5(data).hexdigest(), args=args, sort=sort))
        print(name, len(data), out[-1]["md5"])
    with open(os.path.join(HERE, "join_cases.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
