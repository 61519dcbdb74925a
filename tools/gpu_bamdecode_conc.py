"""manual tuning aid: K contexts decode + run the same chr19-shaped BAM file image concurrently, for several builds of the library"""
import sys, os, time, ctypes as C
from concurrent.futures import ThreadPoolExecutor
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
fpin = C.c_void_p(); assert lib0.rsigpu_pinned_alloc(C.c_size_t(L), C.byref(fpin)) == 0
C.memmove(fpin, fa.ctypes.data, L)

def one(cx):
    cx.set_reference_ptr(fpin.value, L); cx.pileup_begin(); cx.bam_begin(1)
    off = h["coff"]; first = True
    while off < len(data):
        consumed, runs = cx.bam_feed(pin.value + off, len(data) - off, skip=h["skip"] if first else 0)
        for i, (tid, nr) in enumerate(runs):
            cx.bam_take(i, cx)
        if consumed == 0:
            break
        first = False; off += consumed
    cx.bam_end(); cx.have_reads()
    return len(cx.run())

for lib in libs:
    for K in (1, 4):
        ctxs = [api.Context(lib=lib, minq=0, min_baseQ=10) for _ in range(K)]
        pool = ThreadPoolExecutor(K)
        for _ in range(2):
            n = list(pool.map(one, ctxs))
        t0 = time.perf_counter()
        for _ in range(3):
            n = list(pool.map(one, ctxs))
        dt = (time.perf_counter() - t0) / 3
        print("%s K=%d: %.1f ms per step, %.1f ms per contig, calls %s" % (os.path.basename(lib), K, 1e3 * dt, 1e3 * dt / K, n), flush=True)
        for cx in ctxs:
            cx.close()
        pool.shutdown()
