"""manual profiling aid: one depth-path contig of a given size through the real library (for ncu)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import make_case
from rsicnv_b200 import api
L = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_011
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
fa, d, _ = make_case(L, 23, stress=False, n_events=20, lens=(2000, 5000, 10000, 30000, 100000))
ctx = api.Context()
ctx.set_reference(fa); ctx.set_depth(d)
for _ in range(reps):
    calls = ctx.run()
print(len(calls), ctx.stage_ms(), ctx.debug_state()["Lmax"])
ctx.set_profile(True); ctx.run()
for nm, ms, n in sorted(ctx.profile(), key=lambda x: -x[1])[:8]:
    print("%-24s %8.4f ms x%d" % (nm, ms, n))
ds = ctx.debug_state()
print({k: (v / 1.965e6 if k not in ("cp_tests", "cp_cnvlen_total", "cp_nref_total") else v) for k, v in ds.items() if k.startswith("cp_")}, "(ms at 1965 MHz)")
