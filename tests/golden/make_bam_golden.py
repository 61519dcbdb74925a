"""Generates tests/golden/tiny_rich.bam + tiny_rich.json: a small BAM with what real files carry (read names of varying
length, random bases, optional fields, two contigs, unmapped reads at the end, deflate level 6) and, per refID, the sha256 of
every field AS THE REFERENCE'S OWN samtools-0.1.18 RETURNS IT (bam_read1 through oracle/_ref/libref_harness.so::ref_bam_records).
Run in the development container (needs /root/reference built into oracle/_ref):  python tests/golden/make_bam_golden.py"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from rsicnv_b200 import synth  # noqa: E402
from bind import ref_bam_records  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    contigs = [("1", 40000), ("2", 30000), ("MT", 16569)]
    r0, _ = synth.make_reads(40000, 101, None, coverage=5, n_events=0, tid=0, frac_indel=0.3)
    r1, _ = synth.make_reads(30000, 102, None, coverage=4, n_events=0, tid=1)
    path = os.path.join(HERE, "tiny_rich.bam")
    synth.write_bam(path, contigs, {0: r0, 1: r1}, level=6, rich=17, unmapped_tail=7, block_size=4000)
    out = {"file_sha": hashlib.sha256(open(path, "rb").read()).hexdigest(), "names": [c[0] for c in contigs], "lens": [c[1] for c in contigs], "tids": {}}
    for tid in (0, 1, -1):
        R = ref_bam_records(path, tid)
        out["tids"][str(tid)] = {"n": int(len(R["pos"])), **{k: sha(v) for k, v in R.items()}}
    json.dump(out, open(os.path.join(HERE, "tiny_rich.json"), "w"), indent=1)
    print(os.path.getsize(path), {t: v["n"] for t, v in out["tids"].items()})


if __name__ == "__main__":
    main()
