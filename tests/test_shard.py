"""Chromosome sharding: LPT balance on the b37 lengths and the ordered gather over a world_size-2 gloo group."""
import os
import subprocess
import sys

from rsicnv_b200 import shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_balance_b37():
    lens = list(synth.B37_LENS.values())
    assert shard.imbalance(lens, 2) < 1.01 and shard.imbalance(lens, 4) < 1.02 and shard.imbalance(lens, 8) < 1.05   # 1.002 / 1.008 / 1.038
    a = shard.lpt_assign(lens, 8)
    assert sorted(set(a)) == list(range(8))
    assert shard.lpt_assign([5, 5, 5], 1) == [0, 0, 0]


WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch.distributed as dist
from rsicnv_b200 import shard
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["MASTER_PORT"], rank=rank, world_size=world)
lens = [300, 250, 200, 120, 100, 90]
assign = shard.lpt_assign(lens, world)
mine = {t: ["contig%%d row%%d rank%%d" %% (t, k, rank) for k in range(t %% 3)] for t in range(len(lens)) if assign[t] == rank}
rows = shard.gather_rows(mine, len(lens), rank, world)
if rank == 0:
    want = [r for t in range(len(lens)) for r in ["contig%%d row%%d rank%%d" %% (t, k, assign[t]) for k in range(t %% 3)]]
    assert rows == want, (rows, want)
    print("gather ok", len(rows))
else:
    assert rows is None
dist.destroy_process_group()
''' % ROOT


def test_ordered_gather_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 400)
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_PORT=port, MASTER_ADDR="127.0.0.1"),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "gather ok" in outs[0]
