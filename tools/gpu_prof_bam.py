"""manual profiling aid: one chr19-shaped BAM contig (reads as SoA) through the real library (for ncu)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from rsicnv_b200 import api, synth
L = int(sys.argv[1]) if len(sys.argv) > 1 else synth.CHR19_LEN
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
fa = synth.make_fasta(L, 19)
reads, _ = synth.make_reads(L, 19, fa, coverage=30, n_events=20)
ctx = api.Context(minq=0, min_baseQ=10)
ctx.set_reference(fa); ctx.pileup_begin(); ctx.pileup_push(reads); ctx.have_reads()
for _ in range(reps):
    calls = ctx.run()
print(len(calls), len(reads["pos"]), ctx.stage_ms())
ctx.set_profile(True); ctx.run()
for nm, ms, n in sorted(ctx.profile(), key=lambda x: -x[1])[:16]:
    print("%-24s %8.4f ms x%d" % (nm, ms, n))
