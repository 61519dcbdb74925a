"""manual: do K concurrent rsigpu_bam_feed calls (K decoder contexts, K host threads, same pinned BAM image) overlap on the GPU?
Prints the wall time of a round of K feeds and, with per-kernel profiling on, each context's k_bgzf_inflate duration."""
import sys, os, time, ctypes as C
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from rsicnv_b200 import api, synth
L = int(sys.argv[1]) if len(sys.argv) > 1 else synth.CHR19_LEN
KS = [int(k) for k in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8]
fa = synth.make_fasta(L, 19)
reads, _ = synth.make_reads(L, 19, fa, coverage=30, n_events=20)
path = "/tmp/prof.bam"
synth.write_bam(path, [("19", L)], {0: reads}, level=1, random_seq=7, threads=32)
data = np.fromfile(path, np.uint8)
h = api.parse_bam_header(data)
LIB = sys.argv[3] if len(sys.argv) > 3 else None
lib0 = api.load_library(LIB)
print("library", LIB or api.DEFAULT_LIB, flush=True)
pin = C.c_void_p(); assert lib0.rsigpu_pinned_alloc(C.c_size_t(len(data)), C.byref(pin)) == 0
C.memmove(pin, data.ctypes.data, len(data))
print("BAM image %.1f MB" % (len(data) / 1e6), flush=True)

def feed(cx):
    cx.bam_begin(1)
    t0 = time.perf_counter()
    consumed, runs = cx.bam_feed(pin.value + h["coff"], len(data) - h["coff"], skip=h["skip"])
    t1 = time.perf_counter()
    cx.bam_end()
    return t0, t1

ctxs = [api.Context(lib=LIB) for _ in range(max(KS))]
for K in KS:
    pool = ThreadPoolExecutor(K)
    cs = ctxs[:K]
    for _ in range(2):
        list(pool.map(feed, cs))
    t0 = time.perf_counter()
    R = 3
    spans = []
    for _ in range(R):
        spans.append(list(pool.map(feed, cs)))
    wall = (time.perf_counter() - t0) / R
    last = spans[-1]; base = min(a for a, b in last)
    print("K=%d: %.1f ms per round, %.1f ms per feed; feed spans (ms) %s" % (K, 1e3 * wall, 1e3 * wall / K, " ".join("%.0f-%.0f" % (1e3 * (a - base), 1e3 * (b - base)) for a, b in last)), flush=True)
    for cx in cs:
        cx.set_profile(True)
    list(pool.map(feed, cs))
    print("      profiled: inflate ms per context", ["%.1f" % dict((nm, ms) for nm, ms, n in cx.profile()).get("k_bgzf_inflate", -1) for cx in cs], flush=True)
    for cx in cs:
        cx.set_profile(False)
    pool.shutdown()
