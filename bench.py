#!/usr/bin/env python
"""bench.py -- Gbases/s of the B200-native `rsicnv rsi` depth -> CNV-call path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload chr19_bam|chr19_depth]

One "step" = one pass of the hot path over one synthetic b37-chr19-shaped contig (59,128,983 bp, 30x, 20
planted DEL/DUP) per GPU.  `value` = bases of all contigs processed per second with the inputs already
resident in HBM (device time, CUDA events on the launching stream, max over ranks); `e2e` = the same
through the C ABI with pinned HOST buffers, host<->device copies inside the timed region.  N > 1: one
process per GPU (torchrun), contigs are independent units => no data-path collective, weak scaling.
`--impl reference` times the reference's own CPU implementation (oracle/_ref, built from
/root/reference by oracle/Makefile.ref) on the box's host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rsicnv_b200 import synth  # noqa: E402

CHR19 = synth.CHR19_LEN
METRIC = "Gbases/s depth->RSI CNV calls"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, gpu_index: int):
        self.idx = gpu_index; self.proc = None; self.lines = []

    def start(self):
        q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_inputs(seed: int, L: int):
    fa = synth.make_fasta(L, seed)
    depth, events = synth.make_depth(L, seed, fa, n_events=20)
    return fa, depth, events


def make_bam_inputs(seed: int, L: int, coverage: float = 30.0):
    fa = synth.make_fasta(L, seed)
    reads, events = synth.make_reads(L, seed, fa, coverage=coverage, n_events=20)
    return fa, reads, events


READ_FIELDS = ("pos", "mpos", "isize", "mtid", "flag", "mapq", "cigar_off", "cigar", "qual_off", "qual")


# algorithmic HBM bytes of one launch of each streaming kernel (DESIGN.md "kernels"); L = contig, Lc = after N removal
def kernel_bytes(name: str, L: int, Lc: int, reads_bytes: int = 0, qual_bytes: int = 0):
    return {
        "k_qual_mask": qual_bytes,        # every base quality once (the 1-bit-per-base mask it writes is overhead, not counted)
        "k_gc_table": 5 * L,          # depth 4 + FASTA 1, read once
        "k_gc_adjust": 9 * L,         # depth 4 + FASTA 1 read, adjusted depth 4 written
        "k_bins": 4 * Lc,             # compacted depth read once
        "k_pileup_tile": reads_bytes - qual_bytes + 4 * L,   # every read record's core fields + CIGAR once + depth written once
    }.get(name, 0)


# DRAM traffic per launch on the chr19 workloads, from the ncu --set full captures kept under profiles/ (bytes)
NCU_TRAFFIC = {"k_qual_mask": 1.679254e9 + 203.3e6, "k_pileup_tile": 0.615778e9 + 231.7e6, "k_gc_table": 0.295681e9 + 4.9e6,
               "k_gc_adjust": 0.295735e9 + 181.3e6, "k_bins": 0.223885e9 + 11.3e6}


def ref_worker(args):
    """one reference-CPU process: the reference's own functions on an in-memory contig (oracle/_ref/libref_harness.so)"""
    seed, L, kind = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from bind import Lib
    fa, depth, _ = make_inputs(seed, L)
    r = Lib(kind); r.set_params()
    t0 = time.perf_counter()
    res = r.depth_path(depth, fa, 3)
    return time.perf_counter() - t0, len(res["calls"])


def ref_bam_worker(args):
    """one reference-CPU process on the BAM path: the UNMODIFIED reference CLI on its own sorted + indexed synthetic BAM"""
    seed, L, workdir = args[:3]
    with_ours = len(args) > 3 and args[3]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from bind import REF_BAMTOOL, REF_BIN
    d = os.path.join(workdir, f"s{seed}")
    os.makedirs(d, exist_ok=True)
    fa, reads, _ = make_bam_inputs(seed, L)
    synth.write_fasta(os.path.join(d, "t.fa"), "19", fa)
    synth.write_bam(os.path.join(d, "t.bam"), [("19", L)], {0: reads})
    subprocess.run([REF_BAMTOOL, "index", os.path.join(d, "t.bam")], check=True)
    t0 = time.perf_counter()
    subprocess.run([REF_BIN, "rsi", "-b", os.path.join(d, "t.bam"), "-f", os.path.join(d, "t.fa"), "-q", "0", "-Q", "10", "-np", "-o", os.path.join(d, "out.txt")],
                   check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    dt = time.perf_counter() - t0
    ncalls = sum(1 for ln in open(os.path.join(d, "out.txt")) if not ln.startswith("#"))
    ours = None
    cli = os.path.join(ROOT, "rsicnv_b200", "bin", "rsicnv")
    if with_ours and os.path.exists(cli):
        # the same files through this repo's CLI (host BGZF/BAM decode + the CUDA path); second run = CUDA context and page cache warm
        ts = []; notes = []
        for extra in ([], [], [], ["-hostdecode"]):
            t0 = time.perf_counter()
            r = subprocess.run([cli, "rsi", "-b", os.path.join(d, "t.bam"), "-f", os.path.join(d, "t.fa"), "-q", "0", "-Q", "10", "-np", "-o", os.path.join(d, "ours.txt")] + extra,
                               check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=dict(os.environ, RSICNV_TIMING="1"))
            ts.append(time.perf_counter() - t0)
            notes += [ln for ln in r.stderr.splitlines() if ln.startswith("#timing")]
        ours = {"first_s": ts[0], "second_s": ts[1], "runs_s": ts[:3], "hostdecode_s": ts[3], "timing": notes,
                "identical_table": open(os.path.join(d, "ours.txt"), "rb").read() == open(os.path.join(d, "out.txt"), "rb").read()}
    return dt, ncalls, ours


def ref_bam_port_worker(args):
    """fallback when oracle/_ref is absent: the oracle restatement of the same path on the read SoA"""
    seed, L = args[:2]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from bind import Lib, oracle_bam_path
    fa, reads, _ = make_bam_inputs(seed, L)
    o = Lib("oracle")
    t0 = time.perf_counter()
    res = oracle_bam_path(o, reads, fa, minq=0, min_baseQ=10)
    return time.perf_counter() - t0, len(res["calls"])


def time_reference_bam(sample_len: int, procs: int, seed0: int, with_ours: bool = False):
    import multiprocessing as mp
    import tempfile
    ctx = mp.get_context("spawn")
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "rsicnv"))
    if not have_ref:
        ref_kind()
    with tempfile.TemporaryDirectory() as td, ctx.Pool(procs) as pool:
        out = pool.map(ref_bam_worker if have_ref else ref_bam_port_worker, [(seed0 + i, sample_len, td, with_ours) for i in range(procs)])
    return max(o[0] for o in out), out, ("reference" if have_ref else "port")


def ref_kind():
    """"reference" = the unmodified reference objects built into oracle/_ref; "port" = the oracle restatement"""
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")):
        return "ref", "reference"
    if not os.path.exists(os.path.join(ROOT, "oracle", "librsi_oracle.so")):
        subprocess.run(["make", "-s", "oracle"], cwd=ROOT, check=True)
    return "oracle", "port"


def time_reference(sample_len: int, procs: int, seed0: int):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    lib_kind = ref_kind()[0]
    with ctx.Pool(procs) as pool:
        t0 = time.perf_counter()
        out = pool.map(ref_worker, [(seed0 + i, sample_len, lib_kind) for i in range(procs)])
        wall = time.perf_counter() - t0
    cpu = max(o[0] for o in out)       # slowest worker's compute time (input synthesis excluded)
    return cpu, wall, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="chr19_bam", choices=["chr19_bam", "chr19_depth"])
    ap.add_argument("--len", type=int, default=CHR19, help="contig length (debug; the contract uses the default)")
    ap.add_argument("--cpu-sample", type=int, default=12_000_000, help="contig length of the bounded CPU-reference sample (> 10 Mbp for the BAM path)")
    ap.add_argument("--profile-steps", type=int, default=2)
    ap.add_argument("--contigs-per-step", type=int, default=4,
                    help="independent chr19-shaped contigs in flight per GPU and step (one context + stream + host thread each, as the whole-genome CLI runs them)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    L = a.len
    bam = a.workload == "chr19_bam"
    if bam:
        workload = (f"chr19_bam (BASELINE.json configs[1]): rsicnv rsi -b <simulated sorted chr19-shaped BAM, {L} bp, 30x 2x100bp pairs, 20 planted DEL/DUP> "
                    f"-f <synthetic FASTA> -q 0 -Q 10 -m 101 -np (full pileup + RP/Q0 path; reads enter the C ABI as decoded SoA batches)")
        l2 = "inputs (read SoA ~150 B/read x 17.7 M reads + FASTA) are far larger than the 126 MB L2"
    else:
        workload = f"chr19_depth (BASELINE.json configs[0]): rsicnv rsi -d <synthetic chr19-shaped depth, {L} bp, 30x NB-like, 20 planted DEL/DUP> -c 19 -f <synthetic FASTA> -m 101 -np"
        l2 = "inputs (depth 4 B/base + FASTA 1 B/base = %.0f MB per contig) are larger than the 126 MB L2" % (5 * L / 1e6)
    K = max(1, a.contigs_per_step)
    config = {"workload": workload, "contigs_per_step_per_gpu": K, "contig_bp": L, "m": 101,
              "parallelism": f"contig-sharded x{world} GPUs, {K} contigs in flight per GPU (independent contexts/streams)", "l2": l2}

    if a.impl == "reference":
        if rank != 0:
            return 0
        procs = min(os.cpu_count() or 1, 32)
        sample = min(a.cpu_sample, L)
        kind = ref_kind()[1]
        times = []
        for s in range(a.warmup + a.steps):
            if s < a.warmup and s > 0:
                continue  # one warm-up pass is enough for a CPU job (page cache / import)
            if bam:
                cpu, out, kind = time_reference_bam(sample, procs, 1000 + 97 * s)
            else:
                cpu, wall, out = time_reference(sample, procs, 1000 + 97 * s)
            if s >= a.warmup:
                times.append(cpu)
        ms = 1e3 * float(np.mean(times))
        val = procs * sample / (ms / 1e3) / 1e9
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Gbases/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": val, "unit": "Gbases/s", "cores": procs, "kind": kind,
                                 "sample": (f"{procs} processes x one {sample} bp 30x synthetic BAM each through the unmodified `rsicnv rsi -b` CLI (BGZF/BAM decode included)" if bam else
                                            f"{procs} processes x one {sample} bp synthetic contig each through the reference's own checkgccontent..detectcnv..sd_filters "
                                            f"(in-memory depth array, same boundary as the C ABI; text parsing excluded)")},
                "e2e": {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from rsicnv_b200 import api

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the product has no CPU path"}))
        return 2
    torch.cuda.set_device(local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"   # NCCL's version banner goes to stdout at VERSION/INFO level; stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # pinned host buffers (the e2e leg copies from these every step)
    reads_bytes = 0
    qual_bytes = 0
    if bam:
        fa, reads, events = make_bam_inputs(19 + rank, L)
        pins = {k: torch.from_numpy(reads[k]).pin_memory() for k in READ_FIELDS}
        batch = api.ReadBatch()
        batch.n_reads = len(reads["pos"]); batch.tid = 0
        for k in READ_FIELDS:
            setattr(batch, k, pins[k].data_ptr())
        nreads = len(reads["pos"])
        h2d = L + sum(int(pins[k].numel() * pins[k].element_size()) for k in READ_FIELDS)
        # bytes the pileup kernel has to read once: pos, flag, mapq, CIGAR offsets + ops, quality offsets + qualities
        reads_bytes = sum(int(pins[k].numel() * pins[k].element_size()) for k in ("pos", "flag", "mapq", "cigar_off", "cigar", "qual_off", "qual"))
        qual_bytes = int(pins["qual"].numel())
        config["reads"] = nreads
        ctxs = [api.Context(device=local, minq=0, min_baseQ=10) for _ in range(K)]
        # the same reads as a BAM FILE image (BGZF level 1, random bases so that it compresses like a real one): the
        # end-to-end leg starts from these bytes, as the reference does (samtools bgzf/bam readers, BAM in page cache)
        import tempfile
        with tempfile.TemporaryDirectory() as td:
            t0 = time.perf_counter()
            synth.write_bam(os.path.join(td, "t.bam"), [("19", L)], {0: reads}, level=1, random_seq=7 + rank, threads=max(4, min(32, (os.cpu_count() or 8) // world)))
            bam_np = np.fromfile(os.path.join(td, "t.bam"), np.uint8)
        bam_hdr = api.parse_bam_header(bam_np)
        bam_pin = torch.from_numpy(bam_np).pin_memory()
        config["bam_file_bytes"] = int(bam_np.size); config["bam_write_s"] = round(time.perf_counter() - t0, 1)
    else:
        fa, depth, events = make_inputs(19 + rank, L)
        dp_pin = torch.from_numpy(depth).pin_memory()
        h2d = 5 * L
        ctxs = [api.Context(device=local) for _ in range(K)]
    ctx = ctxs[0]
    fa_pin = torch.from_numpy(fa).pin_memory()
    bufs = [(api.Cnv * 65536)() for _ in range(K)]
    buf = bufs[0]
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(K)

    def stage_one(cx):
        cx.set_reference_ptr(fa_pin.data_ptr(), L)
        if bam:
            cx.pileup_begin()
            cx._ck(cx.lib.rsigpu_pileup_push(cx.h, C.byref(batch)))
            cx.have_reads()
        else:
            cx.set_depth_ptr(dp_pin.data_ptr(), L)

    def stage_inputs():
        for cx in ctxs:
            stage_one(cx)

    def run_all():
        """one step: the K resident contigs go through the hot path concurrently (ctypes releases the GIL)"""
        return list(pool.map(lambda kb: kb[0].run_count(kb[1], 65536), zip(ctxs, bufs)))

    def e2e_all(n=1):
        def one(kb):
            for _ in range(n):
                stage_one(kb[0])
                r = kb[0].run_count(kb[1], 65536)
            return r
        return list(pool.map(one, zip(ctxs, bufs)))

    def file_one(kb):
        """BAM file bytes (pinned host memory) -> BGZF inflate + record decode on the GPU -> pileup -> ... -> calls on the host"""
        cx, out = kb
        cx.set_reference_ptr(fa_pin.data_ptr(), L)
        cx.pileup_begin()
        cx.bam_begin(len(bam_hdr["names"]))
        off = bam_hdr["coff"]; first = True; n = int(bam_np.size)
        while off < n:
            consumed, runs = cx.bam_feed(bam_pin.data_ptr() + off, n - off, skip=bam_hdr["skip"] if first else 0)
            for i, (tid, nr) in enumerate(runs):
                if tid == 0:
                    cx.bam_take(i, cx)
            if consumed == 0:
                break
            first = False; off += consumed
        cx.bam_end(); cx.have_reads()
        return cx.run_count(out, 65536)

    def file_all(n=1):
        """every context runs n contigs back to back in its own host thread (no barrier between the steps of different contexts:
        the copies of one contig overlap the kernels of another, as in a whole-genome run)"""
        return list(pool.map(lambda kb: [file_one(kb) for _ in range(n)][-1], zip(ctxs, bufs)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg: inputs staged once, each step = the whole hot path on the K resident contigs
    stage_inputs()
    for _ in range(a.warmup):
        ncalls = run_all()[0]
    sampler = ClockSampler(local); sampler.start()
    barrier()
    l0 = sum(cx.launch_count() for cx in ctxs)
    t0 = time.perf_counter()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    dev_ms = 0.0; stages = None
    for _ in range(a.steps):
        ts = time.perf_counter()
        ncalls = run_all()[0]
        torch.cuda.synchronize()
        dev_ms += 1e3 * (time.perf_counter() - ts)        # K streams overlap: the step time is the bracketed wall time of the step
        sm = ctx.stage_ms()
        stages = sm if stages is None else {k: stages[k] + sm[k] for k in sm}
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop()
    launches = sum(cx.launch_count() for cx in ctxs) - l0
    st = ctx.chr_stats()
    # ---- end-to-end leg: pinned host buffers -> H2D -> hot path -> calls on the host, every step, K contigs in flight
    for _ in range(max(1, a.warmup // 2)):
        e2e_all()
    barrier()
    t0 = time.perf_counter()
    ne = e2e_all(a.steps)[0]
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    file_ms = None
    if bam:
        nf = file_all()[0]
        assert nf == ne, "decoded-on-GPU path disagrees with the staged-reads path"
        barrier()
        t0 = time.perf_counter()
        nf = file_all(a.steps)[0]
        barrier()
        file_ms = 1e3 * (time.perf_counter() - t0)
    # ---- per-kernel device times (extra profiled steps, CUDA events around every launch on the context's stream)
    ctx.set_profile(True)
    for _ in range(a.profile_steps):
        ctx.run_count(buf, 65536)
    prof = ctx.profile()
    ctx.set_profile(False)
    dprof = []
    if bam:
        ctx.set_profile(True)
        file_one((ctx, buf))
        dprof = [(nm, ms, n) for nm, ms, n in ctx.profile() if nm.startswith("k_bgzf") or nm.startswith("k_bam")]
        ctx.set_profile(False)

    t = torch.tensor([dev_ms, wall_ms, e2e_ms, file_ms or 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, e2e_ms, file_ms = [float(x) for x in t.tolist()]
    if rank == 0:
        peak, peak_src = peaks()
        total_bases = world * K * L * a.steps
        value = total_bases / (dev_ms / 1e3) / 1e9
        kern = []
        for name, ms, n in prof:
            b = kernel_bytes(name, L, st.compact_len, reads_bytes, qual_bytes)
            avg = ms / max(n, 1)
            kern.append({"kernel": name, "launches_per_step": n / a.profile_steps, "avg_ms": avg, "ms_per_step": ms / a.profile_steps,
                         "algorithmic_bytes": b, "gbs": (b / 1e9) / (avg / 1e3) if b and avg > 0 else None})
        kern.sort(key=lambda k: -k["ms_per_step"])
        stream = [k for k in kern if k["algorithmic_bytes"]]
        # the HBM-bound (per-base / per-read streaming) kernel with the largest time; kernels without algorithmic bytes are the
        # bin-level and candidate-list kernels (latency-bound on L2-resident data), listed under "kernels"
        top = stream[0] if stream else (kern[0] if kern else None)
        roof = None
        if top:
            ach = top["gbs"] or 0.0
            roof = {"bound": "hbm", "kernel": top["kernel"], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": NCU_TRAFFIC.get(top["kernel"]) if L == CHR19 else None,
                    "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full (profiles/r1_streaming_full_f.txt)",
                    "top_kernel_by_time": kern[0]["kernel"], "top_kernel_ms": kern[0]["avg_ms"],
                    "peak_source": peak_src, "share_of_step": top["ms_per_step"] / max(sum(k["ms_per_step"] for k in kern), 1e-9),
                    "how": "algorithmic bytes per launch / average launch duration (CUDA events on the context's stream, %d extra profiled steps)" % a.profile_steps,
                    "streaming_kernels": [{"kernel": k["kernel"], "gbs": k["gbs"], "frac": (k["gbs"] or 0) / peak, "ms": k["avg_ms"]} for k in stream]}
        line = {"metric": METRIC, "value": value, "unit": "Gbases/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": dev_ms / a.steps, "wall_ms_per_step": wall_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8 qualities -> int32 depth / f64 statistics", "data": "synthetic", "config": config, "clocks": clocks,
                "e2e": None,
                "gpu_launches": int(launches), "calls_per_contig": int(ncalls), "roofline": roof,
                "single_contig_stage_ms": {k: v / a.steps for k, v in (stages or {}).items()},
                "timing": "e2e legs: K host threads each run `steps` contigs back to back, one barrier + synchronize on both sides of the whole region; "
                          "value: wall time of each step bracketed by torch.cuda.synchronize (K contexts on K streams overlap, per-context CUDA-event "
                          "times are in single_contig_stage_ms); kernels: CUDA events around every launch of one context in extra profiled steps",
                "kernels": kern[:12]}
        soa = {"value": total_bases / (e2e_ms / 1e3) / 1e9, "unit": "Gbases/s", "h2d_bytes_per_step": int(h2d) * K,
               "d2h_bytes_per_step": int(ne * 128 + 53000) * K, "ms_per_step": e2e_ms / a.steps}
        if bam:
            # headline: from the BAM FILE's bytes, like the reference arm (BGZF inflate + BAM record decoding inside the timed region, on the GPU)
            line["e2e"] = {"value": total_bases / (file_ms / 1e3) / 1e9, "unit": "Gbases/s", "h2d_bytes_per_step": (int(bam_np.size) + L) * K,
                           "d2h_bytes_per_step": int(ne * 128 + 53000) * K, "ms_per_step": file_ms / a.steps,
                           "input": "BAM file image (BGZF) + FASTA contig in pinned host memory -> rsigpu_bam_feed/take -> rsigpu_run -> calls on the host"}
            soa["input"] = "already-decoded reads (structure of arrays, rsigpu_pileup_push) + FASTA contig in pinned host memory"
            line["e2e_decoded_reads"] = soa
        else:
            line["e2e"] = soa
        if dprof:
            dec_bytes = int(sum(np.diff(reads["qual_off"].astype(np.int64)) * 3 // 2 + 38 + 4 * np.diff(reads["cigar_off"].astype(np.int64))))
            line["decode_kernels"] = [{"kernel": nm, "launches": n, "ms": ms,
                                       **({"compressed_bytes": int(bam_np.size), "decoded_bytes": dec_bytes, "decoded_gbs": dec_bytes / 1e9 / (ms / 1e3)} if nm == "k_bgzf_inflate" else {})}
                                      for nm, ms, n in sorted(dprof, key=lambda x: -x[1])]
        if world == 1:
            # release this process's device memory and pinned buffers before other processes (the CLIs) are timed
            for cx in ctxs:
                cx.close()
            if bam:
                del pins, bam_pin
            del fa_pin
            torch.cuda.empty_cache()
            sample = min(a.cpu_sample, L)
            if bam:
                cpu, out, kind = time_reference_bam(sample, 1, 19, with_ours=True)
                if len(out[0]) > 2 and out[0][2]:
                    o = out[0][2]
                    line["cli_e2e"] = {"sample_bp": sample, "reference_cli_s": cpu, "this_cli_first_s": o["first_s"], "this_cli_second_s": o["second_s"], "this_cli_hostdecode_s": o["hostdecode_s"], "this_cli_timing": o["timing"],
                                       "this_cli_runs_s": o["runs_s"], "identical_table": o["identical_table"], "speedup_second": cpu / o["second_s"], "speedup_best": cpu / min(o["runs_s"]),
                                       "what": "BAM + FASTA files -> CNV table through each CLI (process start and CUDA context creation included; this CLI decodes the BAM on the GPU, "
                                               "-hostdecode on host threads; its start-up varies by seconds on a box whose GPU is held by the bench process itself)"}
                line["cpu_baseline"] = {"value": sample / cpu / 1e9, "unit": "Gbases/s", "cores": 1, "kind": kind,
                                        "sample": f"one {sample} bp 30x synthetic BAM through the unmodified `rsicnv rsi -b ... -q 0 -Q 10 -np` CLI on one host core "
                                                  f"(the reference is single-threaded; BGZF/BAM decode included, BAM in page cache)"}
            else:
                cpu, _, out = time_reference(sample, 1, 19)
                line["cpu_baseline"] = {"value": sample / cpu / 1e9, "unit": "Gbases/s", "cores": 1, "kind": ref_kind()[1],
                                        "sample": f"one {sample} bp synthetic contig through checkgccontent..detectcnv..sd_filters on one host core "
                                                  f"(the reference is single-threaded; in-memory depth array, text parsing excluded)"}
        print(json.dumps(line))
    for cx in ctxs:
        cx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
