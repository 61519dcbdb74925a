"""manual profiling aid: a chr19-shaped BAM FILE through the GPU decoder (rsigpu_bam_feed / take) and the hot path"""
import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from rsicnv_b200 import api, synth
L = int(sys.argv[1]) if len(sys.argv) > 1 else synth.CHR19_LEN
chunk = (int(sys.argv[2]) if len(sys.argv) > 2 else 4096) << 20
fa = synth.make_fasta(L, 19)
reads, _ = synth.make_reads(L, 19, fa, coverage=30, n_events=20)
path = "/tmp/prof.bam"
t = time.time(); synth.write_bam(path, [("19", L)], {0: reads}, level=1, random_seq=7, threads=16); print("write_bam %.1f s, %d bytes" % (time.time() - t, os.path.getsize(path)))
data = np.fromfile(path, np.uint8)
h = api.parse_bam_header(data)
lib = api.load_library()
pin = C.c_void_p()
assert lib.rsigpu_pinned_alloc(C.c_size_t(len(data)), C.byref(pin)) == 0
C.memmove(pin, data.ctypes.data, len(data))
dec = api.Context(); ctx = api.Context(minq=0, min_baseQ=10)
for rep in range(3):
    if rep == 2:
        dec.set_profile(True)
    ctx.set_reference(fa); ctx.pileup_begin()
    dec.bam_begin(1)
    t0 = time.perf_counter(); tf = 0.0; tt = 0.0
    off = h["coff"]; first = True; nfeeds = 0
    while off < len(data):
        n = min(chunk, len(data) - off)
        a = time.perf_counter()
        consumed, runs = dec.bam_feed(pin.value + off, n, skip=h["skip"] if first else 0)
        b = time.perf_counter()
        for i, (tid, nr) in enumerate(runs):
            dec.bam_take(i, ctx)
        c = time.perf_counter()
        tf += b - a; tt += c - b; nfeeds += 1
        if consumed == 0:
            break
        first = False; off += consumed
    dec.bam_end(); ctx.have_reads()
    t1 = time.perf_counter()
    calls = ctx.run()
    t2 = time.perf_counter()
    print("rep %d: decode %.1f ms (feed %.1f, take %.1f, %d feeds), run %.1f ms, %d calls, rewalked %d" % (rep, 1e3 * (t1 - t0), 1e3 * tf, 1e3 * tt, nfeeds, 1e3 * (t2 - t1), len(calls), dec.debug_state()["bam_rewalked"]))
for nm, ms, n in sorted(dec.profile(), key=lambda x: -x[1])[:10]:
    print("%-24s %9.4f ms x%d" % (nm, ms, n))
# parity with the host-pushed reads
ref = api.Context(minq=0, min_baseQ=10); ref.set_reference(fa); ref.pileup_begin(); ref.pileup_push(reads); ref.have_reads()
want = ref.run()
print("calls identical to pileup_push path:", len(want) == len(calls) and all(bytes(x) == bytes(y) for x, y in zip(want, calls)))
