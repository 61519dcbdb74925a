"""manual: host-clock trace (RSIGPU_TRACE=1) of 4 contexts decoding + running the same BAM image concurrently"""
import sys, os, time, ctypes as C
from concurrent.futures import ThreadPoolExecutor
os.environ["RSIGPU_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from rsicnv_b200 import api, synth
L = synth.CHR19_LEN
fa = synth.make_fasta(L, 19)
reads, _ = synth.make_reads(L, 19, fa, coverage=30, n_events=20)
path = "/tmp/prof.bam"
synth.write_bam(path, [("19", L)], {0: reads}, level=1, random_seq=7, threads=32)
data = np.fromfile(path, np.uint8)
h = api.parse_bam_header(data)
lib0 = api.load_library()
pin = C.c_void_p(); assert lib0.rsigpu_pinned_alloc(C.c_size_t(len(data)), C.byref(pin)) == 0
C.memmove(pin, data.ctypes.data, len(data))
fpin = C.c_void_p(); assert lib0.rsigpu_pinned_alloc(C.c_size_t(L), C.byref(fpin)) == 0
C.memmove(fpin, fa.ctypes.data, L)

def one(cx):
    cx.set_reference_ptr(fpin.value, L); cx.pileup_begin(); cx.bam_begin(1)
    consumed, runs = cx.bam_feed(pin.value + h["coff"], len(data) - h["coff"], skip=h["skip"])
    for i, (tid, nr) in enumerate(runs):
        cx.bam_take(i, cx)
    cx.bam_end(); cx.have_reads()
    return len(cx.run())

K = 4
ctxs = [api.Context(minq=0, min_baseQ=10) for _ in range(K)]
pool = ThreadPoolExecutor(K)
for _ in range(2):
    list(pool.map(one, ctxs))
sys.stderr.write("==== traced steps ====\n"); sys.stderr.flush()
t0 = time.perf_counter()
list(pool.map(lambda cx: [one(cx) for _ in range(2)], ctxs))
sys.stderr.write("==== total %.1f ms for %d contigs ====\n" % (1e3 * (time.perf_counter() - t0), 2 * K))
