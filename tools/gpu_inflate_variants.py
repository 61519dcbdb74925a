"""manual tuning aid: feed time of one chr19-shaped BAM file image for several builds of the library (argv: library paths)"""
import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from rsicnv_b200 import api, synth
L = synth.CHR19_LEN
fa = synth.make_fasta(L, 19)
reads, _ = synth.make_reads(L, 19, fa, coverage=30, n_events=20)
path = "/tmp/prof.bam"
synth.write_bam(path, [("19", L)], {0: reads}, level=1, random_seq=7, threads=16)
data = np.fromfile(path, np.uint8)
h = api.parse_bam_header(data)
libs = sys.argv[1:] or [api.DEFAULT_LIB]
lib0 = api.load_library(libs[0])
pin = C.c_void_p(); assert lib0.rsigpu_pinned_alloc(C.c_size_t(len(data)), C.byref(pin)) == 0
C.memmove(pin, data.ctypes.data, len(data))
for lp in libs:
    dec = api.Context(lib=lp)
    for rep in range(3):
        if rep == 2:
            dec.set_profile(True)
        dec.bam_begin(1)
        t0 = time.perf_counter()
        consumed, runs = dec.bam_feed(pin.value + h["coff"], len(data) - h["coff"], skip=h["skip"])
        t1 = time.perf_counter()
        dec.bam_end()
    pr = dict((nm, ms) for nm, ms, n in dec.profile())
    print("%-28s feed %.1f ms, inflate %.2f ms, runs %s" % (os.path.basename(lp), 1e3 * (t1 - t0), pr.get("k_bgzf_inflate", -1), runs), flush=True)
    dec.close()
