"""small depth + BAM cases through the real library (for compute-sanitizer runs)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import make_case
from rsicnv_b200 import api, synth
fa, d, _ = make_case(500_003, 39, stress=True)
with api.Context() as ctx:
    ctx.set_reference(fa); ctx.set_depth(d); print("depth calls", len(ctx.run()))
L = 10_300_000
fa = synth.make_fasta(L, 3)
reads, _ = synth.make_reads(L, 3, fa, coverage=3, n_events=4, lens=(3000, 8000, 20000))
with api.Context(minq=0, min_baseQ=10) as ctx:
    ctx.set_reference(fa); ctx.pileup_begin(); ctx.pileup_push(reads); ctx.have_reads(); print("bam calls", len(ctx.run()))
