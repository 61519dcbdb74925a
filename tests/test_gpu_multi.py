"""Multi-GPU paths (need >= 2 devices: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`; skipped on a
one-GPU box).  Contigs are independent units (rsi.cpp:2189-2217), so every multi-GPU run must reproduce the one-GPU table
byte for byte:
  * the C ABI: a decoder context on GPU 0 hands its runs to a contig context on GPU 1 (rsigpu_bam_take across devices);
  * the CLI with `-gpus 2`: per-GPU decoders started from the .bai offsets, and the one-pass route without the index whose
    decoder sits on GPU 0 and feeds the contexts of both GPUs;
  * bench.py's whole-genome workload under torchrun with 2 ranks: same table hash as with 1 rank.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from bind import REF_BAMTOOL, REF_BIN, have_ref
from rsicnv_b200 import api, synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "rsicnv_b200", "bin", "rsicnv")


def _ndev():
    return api.load_library().rsigpu_num_devices()


def _table(path):
    return [ln for ln in open(path).read().splitlines() if not ln.startswith("#input")]


@pytest.fixture(scope="module")
def two_gpus():
    if _ndev() < 2:
        pytest.skip("needs two GPUs")


def test_bam_take_across_gpus(two_gpus, gpu_lib, tmp_path):
    L = 10_400_000
    fa = synth.make_fasta(L, 81)
    reads, _ = synth.make_reads(L, 81, fa, coverage=10, n_events=5, lens=(3000, 8000, 20000))
    bam = str(tmp_path / "t.bam")
    synth.write_bam(bam, [("1", L)], {0: reads}, level=1, random_seq=3)
    data = np.fromfile(bam, np.uint8)
    h = api.parse_bam_header(data)
    outs = []
    for dst_dev in (0, 1):
        with api.Context(device=0, lib=gpu_lib) as dec, api.Context(device=dst_dev, lib=gpu_lib, minq=0, min_baseQ=10) as ctx:
            ctx.set_reference(fa); ctx.pileup_begin()
            dec.bam_begin(1)
            consumed, runs = dec.bam_feed(data[h["coff"]:], skip=h["skip"])
            for i, (tid, n) in enumerate(runs):
                if tid == 0:
                    dec.bam_take(i, ctx)
            dec.bam_end(); ctx.pileup_end()
            raw = ctx.array(api.ARR_RAW_DEPTH).copy()
            calls = ctx.run()
            outs.append((raw.tobytes(), b"".join(bytes(c) for c in calls)))
    assert outs[0] == outs[1] and len(outs[0][1]) > 0


def test_cli_two_gpus_equals_one(two_gpus, tmp_path):
    if not os.path.exists(CLI):
        subprocess.run(["make", "-s", "cli"], cwd=ROOT, check=True)
    lens = [10_900_000, 10_300_000, 10_700_000, 10_250_000]
    names = ["1", "2", "3", "X"]
    fas = [synth.make_fasta(L, 90 + i) for i, L in enumerate(lens)]
    reads = {i: synth.make_reads(L, 90 + i, fas[i], coverage=8, n_events=3, lens=(4000, 9000, 20000), tid=i)[0] for i, L in enumerate(lens)}
    bam = str(tmp_path / "t.bam"); fasta = str(tmp_path / "t.fa")
    synth.write_bam(bam, list(zip(names, lens)), reads, level=1, rich=31, unmapped_tail=30)
    synth.write_fasta_multi(fasta, list(zip(names, fas)))
    common = ["rsi", "-b", bam, "-f", fasta, "-q", "0", "-Q", "10", "-np"]
    tables = {}
    if have_ref():
        subprocess.run([REF_BAMTOOL, "index", bam], check=True)
        subprocess.run([REF_BIN] + common + ["-o", str(tmp_path / "ref.txt")], check=True, capture_output=True)
        tables["reference"] = _table(str(tmp_path / "ref.txt"))
    for key, extra, env in (("1gpu", ["-gpus", "1"], {}), ("2gpu_indexed", ["-gpus", "2"], {}), ("2gpu_one_pass", ["-gpus", "2"], {"RSICNV_NO_INDEX": "1"}),
                            ("2gpu_hostdecode", ["-gpus", "2", "-hostdecode"], {})):
        out = subprocess.run([CLI] + common + extra + ["-o", str(tmp_path / (key + ".txt"))], capture_output=True, text=True, env=dict(os.environ, **env))
        assert out.returncode == 0, (key, out.stderr)
        tables[key] = _table(str(tmp_path / (key + ".txt")))
        if "2gpu" in key:
            assert "on GPU 1" in out.stderr, (key, out.stderr)        # the second GPU really took contigs
    first = tables["1gpu"]
    assert len(first) > 4
    for key, t in tables.items():
        assert t == first, key


def test_bench_whole_genome_two_ranks_same_table(two_gpus, tmp_path):
    """bench.py --workload wg at 1 and 2 ranks (small genome scale): identical table hash, i.e. LPT sharding + the ordered gather
    change nothing"""
    env = dict(os.environ)
    args = ["--workload", "wg", "--genome-scale", "0.05", "--steps", "1", "--warmup", "1", "--no-cli", "--profile-steps", "1"]
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=env, cwd=ROOT, timeout=900)
    assert r1.returncode == 0, r1.stderr[-2000:]
    r2 = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29571",
                         os.path.join(ROOT, "bench.py"), "--gpus", "2"] + args, capture_output=True, text=True, env=env, cwd=ROOT, timeout=900)
    assert r2.returncode == 0, r2.stderr[-2000:]
    j1 = json.loads([ln for ln in r1.stdout.splitlines() if ln.startswith("{")][-1])
    j2 = json.loads([ln for ln in r2.stdout.splitlines() if ln.startswith("{")][-1])
    assert j1["table_sha1"] == j2["table_sha1"] and j1["calls"] == j2["calls"] and j1["calls"] > 0
    assert j2["n_gpus"] == 2 and j2["scaling"] == "strong"
