"""manual profiling aid: rsigpu_bam_feed of one synthetic BAM file image (for ncu -k regex:k_bgzf_inflate)"""
import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from rsicnv_b200 import api, synth
L = int(sys.argv[1]) if len(sys.argv) > 1 else 12_000_000
fa = synth.make_fasta(L, 19)
reads, _ = synth.make_reads(L, 19, fa, coverage=30, n_events=8)
path = "/tmp/prof_inf.bam"
synth.write_bam(path, [("19", L)], {0: reads}, level=1, random_seq=7, threads=16)
data = np.fromfile(path, np.uint8)
h = api.parse_bam_header(data)
dec = api.Context()
for rep in range(2):
    dec.bam_begin(1)
    t0 = time.perf_counter()
    consumed, runs = dec.bam_feed(data[h["coff"]:], skip=h["skip"])
    print("feed %.1f ms" % (1e3 * (time.perf_counter() - t0)), runs)
    dec.bam_end()
