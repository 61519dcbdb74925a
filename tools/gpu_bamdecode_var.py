"""manual tuning aid: decode the same chr19-shaped BAM file image with several builds of the library, per-kernel times"""
import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from rsicnv_b200 import api, synth
L = synth.CHR19_LEN
libs = sys.argv[1:] or [api.DEFAULT_LIB]
fa = synth.make_fasta(L, 19)
reads, _ = synth.make_reads(L, 19, fa, coverage=30, n_events=20)
path = "/tmp/prof.bam"
synth.write_bam(path, [("19", L)], {0: reads}, level=1, random_seq=7, threads=32)
data = np.fromfile(path, np.uint8)
h = api.parse_bam_header(data)
lib0 = api.load_library(libs[0])
pin = C.c_void_p(); assert lib0.rsigpu_pinned_alloc(C.c_size_t(len(data)), C.byref(pin)) == 0
C.memmove(pin, data.ctypes.data, len(data))
for lib in libs:
    dec = api.Context(lib=lib)
    for rep in range(3):
        dec.set_profile(rep == 2)
        dec.bam_begin(1)
        t0 = time.perf_counter()
        consumed, runs = dec.bam_feed(pin.value + h["coff"], len(data) - h["coff"], skip=h["skip"])
        t1 = time.perf_counter()
        if rep == 2:
            got = dec.bam_run_reads(0)
        dec.bam_end()
    ok = all(np.array_equal(got[k].astype(np.int64), np.asarray(reads[k]).astype(np.int64)) for k in got)
    prof = {nm: ms for nm, ms, n in dec.profile()}
    print("%s: feed %.1f ms, inflate %.2f ms, chain %.2f, fields %.2f, payload %.2f, identical %s" % (os.path.basename(lib), 1e3 * (t1 - t0), prof.get("k_bgzf_inflate", 0), prof.get("k_bam_chain", 0), prof.get("k_bam_fields", 0), prof.get("k_bam_payload", 0), ok), flush=True)
    dec.close()
