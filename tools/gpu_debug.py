"""manual GPU debugging aid: stage-by-stage comparison of a small contig against the oracle"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from bind import Lib
from common import make_case
from rsicnv_b200 import api
lib = sys.argv[1] if len(sys.argv) > 1 else None
L = int(sys.argv[2]) if len(sys.argv) > 2 else 400_000
fa, d, _ = make_case(L, 1)
o = Lib("oracle"); o.set_params()
ro = o.depth_path(d, fa, 3, want_bins=True); o.set_params(); ro1 = o.depth_path(d, fa, 1)
ctx = api.Context(lib=lib)
ctx.set_reference(fa); ctx.set_depth(d)
print("raw depth equal", np.array_equal(ctx.array(api.ARR_RAW_DEPTH), d))
print("noseq", ctx.array(api.ARR_NOSEQ_BEG), ctx.array(api.ARR_NOSEQ_END))
try:
    ctx.load_finish()
except Exception as e:
    print("load_finish:", e)
print(ctx.debug_state())
rdc = ctx.array(api.ARR_DEPTH)
print("depth equal", np.array_equal(rdc, ro1["depth"]), rdc[:10], ro1["depth"][:10], int((rdc != ro1["depth"]).sum()))
print("oracle stats", ro1["stats"])
try:
    calls = ctx.run()
    print(ctx.debug_state())
    nb = ctx.chr_stats().nbins
    print("status equal", np.array_equal(ctx.array(api.ARR_BIN_STATUS), ro["bins"][3][:nb]), len(calls), len(ro["calls"]))
except Exception as e:
    print("run:", e)
