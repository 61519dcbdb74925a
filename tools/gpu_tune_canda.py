import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import make_case
from rsicnv_b200 import api, synth
L = synth.CHR19_LEN
fa, d, _ = make_case(L, 23, n_events=20, lens=(2000, 5000, 10000, 30000, 100000))
ctx = api.Context(); ctx.set_reference(fa); ctx.set_depth(d)
for nt in (32, 64, 128, 256, 512):
    ctx.set_cand_threads(nt)
    ctx.run(); ctx.run()
    ctx.set_profile(True); calls = ctx.run()
    pr = {k: v for k, v, n in ctx.profile()}
    ctx.set_profile(False)
    print(nt, len(calls), "cand_a %.4f cand_b %.4f cand_c %.4f" % (pr["k_cand_a"], pr["k_cand_b"], pr["k_cand_c"]), ctx.stage_ms()["total"])
