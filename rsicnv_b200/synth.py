"""Seeded synthetic inputs for the rsicnv `rsi` hot path (no network: every input is simulated).

Shapes follow SURVEY.md §8(d): a b37-shaped contig (uppercase ACGT with N blocks), an NB-like
per-base depth track with planted DEL/DUP segments, and 2x100 bp FR read pairs for the BAM path.
Nothing here is on the measured path; it only fabricates inputs for tests and bench.py.
"""
from __future__ import annotations

import numpy as np

CHR19_LEN = 59_128_983
B37_LENS = {
    "1": 249250621, "2": 243199373, "3": 198022430, "4": 191154276, "5": 180915260, "6": 171115067,
    "7": 159138663, "8": 146364022, "9": 141213431, "10": 135534747, "11": 135006516, "12": 133851895,
    "13": 115169878, "14": 107349540, "15": 102531392, "16": 90354753, "17": 81195210, "18": 78077248,
    "19": 59128983, "20": 63025520, "21": 48129895, "22": 51304566, "X": 155270560, "Y": 59373566,
}
EVENT_LENS = (2_000, 5_000, 10_000, 30_000, 100_000)


def n_blocks_for(L: int) -> list[tuple[int, int]]:
    """N blocks shaped like b37 chr19: telomere gap, centromere, tail (scaled for small L)."""
    if L >= 40_000_000:
        return [(0, 60_000), (int(L * 0.41657), int(L * 0.41657) + 3_100_000), (L - 10_000, L)]
    a = max(1000, L // 1000)
    c0 = int(L * 0.41657)
    return [(0, a), (c0, c0 + max(2000, L // 20)), (L - max(500, a // 6), L)]


def make_fasta(L: int, seed: int, n_blocks=None) -> np.ndarray:
    """uint8 ASCII contig; GC fraction drifts regionally (0.30..0.60) so that GC strata are populated."""
    rng = np.random.default_rng(seed)
    if n_blocks is None:
        n_blocks = n_blocks_for(L)
    seg = 5_000
    nseg = (L + seg - 1) // seg
    gcf = 0.45 + 0.15 * np.sin(np.arange(nseg) * 0.37 + seed) * rng.uniform(0.3, 1.0, nseg)
    p = np.repeat(gcf, seg)[:L].astype(np.float32)
    u = rng.random(L, dtype=np.float32)
    v = rng.integers(0, 2, L, dtype=np.uint8)
    is_gc = u < p
    out = np.where(is_gc, np.where(v == 0, ord("G"), ord("C")), np.where(v == 0, ord("A"), ord("T"))).astype(np.uint8)
    for a, b in n_blocks:
        out[a:b] = ord("N")
    return out


def plant_events(L: int, seed: int, n_events: int, n_blocks, lens=EVENT_LENS, margin=None):
    """Alternating DEL (x0.5) / DUP (x1.5) events away from N blocks and contig ends."""
    rng = np.random.default_rng(seed + 7)
    ev = []
    if n_events == 0:
        return ev
    margin = margin if margin is not None else max(3 * max(lens), L // 200)
    slots = np.linspace(margin, L - margin, n_events + 2)[1:-1]
    for k, c in enumerate(slots):
        ln = int(lens[k % len(lens)])
        s = int(c + rng.integers(-ln, ln))
        e = s + ln
        bad = any(s < b + 3 * ln and e > a - 3 * ln for a, b in n_blocks)
        if bad or s < margin or e > L - margin:
            continue
        ev.append((s, e, 0.5 if k % 2 == 0 else 1.5))
    return ev


def make_depth(L: int, seed: int, fasta: np.ndarray | None = None, n_events: int = 20, mean: float = 30.0,
               shape: float = 40.0, lens=EVENT_LENS, gc_bias: float = 0.25, events=None, block: int = 100):
    """NB-like depth: moving-average(100) of Poisson(Gamma(shape, mean) per `block` bases) with a mild GC bias,
    0 inside N blocks, planted copy-number events.  Returns (int32 depth, events)."""
    rng = np.random.default_rng(seed)
    if fasta is None:
        fasta = make_fasta(L, seed)
    n_blocks = _n_runs(fasta)
    if events is None:
        events = plant_events(L, seed, n_events, n_blocks, lens)
    nblk = (L + block - 1) // block
    lam = np.repeat(rng.gamma(shape, mean / shape, nblk).astype(np.float32), block)[:L].copy()
    cn = np.ones(L, np.float32)
    for s, e, f in events:
        cn[s:e] = f
    # regional GC effect (201-bp window GC fraction), multiplicative
    gc = ((fasta == ord("G")) | (fasta == ord("C"))).astype(np.float32)
    cs = np.concatenate(([0.0], np.cumsum(gc, dtype=np.float64)))
    w = 201
    lo = np.clip(np.arange(L) - w // 2, 0, max(L - w, 0))
    frac = ((cs[np.minimum(lo + w, L)] - cs[lo]) / w).astype(np.float32)
    lam *= cn * (1.0 + gc_bias * (frac - 0.45) / 0.15)
    np.maximum(lam, 0, out=lam)
    raw = rng.poisson(lam).astype(np.float64)
    k = 100
    c2 = np.concatenate(([0.0], np.cumsum(raw)))
    hi = np.minimum(np.arange(L) + k // 2, L)
    lo2 = np.maximum(np.arange(L) - k // 2, 0)
    d = np.rint((c2[hi] - c2[lo2]) / (hi - lo2)).astype(np.int32)
    d[fasta == ord("N")] = 0
    return d, events


def _n_runs(fasta: np.ndarray) -> list[tuple[int, int]]:
    isn = (fasta == ord("N")).astype(np.int8)
    d = np.diff(np.concatenate(([0], isn, [0])))
    return list(zip(np.flatnonzero(d == 1).tolist(), np.flatnonzero(d == -1).tolist()))


# --------------------------------------------------------------------------------------------
# files
def write_fasta(path: str, name: str, seq: np.ndarray, width: int = 60) -> None:
    L = len(seq)
    header = f">{name}\n".encode()
    nfull = L // width
    body = np.empty(L + nfull + (1 if L % width else 0), np.uint8)
    if nfull:
        blk = body[: nfull * (width + 1)].reshape(nfull, width + 1)
        blk[:, :width] = seq[: nfull * width].reshape(nfull, width)
        blk[:, width] = 10
    if L % width:
        body[nfull * (width + 1):-1] = seq[nfull * width:]
        body[-1] = 10
    with open(path, "wb") as f:
        f.write(header)
        f.write(body.tobytes())
    with open(path + ".fai", "w") as f:
        f.write(f"{name}\t{L}\t{len(header)}\t{width}\t{width + 1}\n")


def write_fasta_multi(path: str, contigs: list[tuple[str, np.ndarray]], width: int = 60) -> None:
    off = 0
    fai = []
    with open(path, "wb") as f:
        for name, seq in contigs:
            L = len(seq)
            header = f">{name}\n".encode()
            f.write(header)
            off += len(header)
            fai.append(f"{name}\t{L}\t{off}\t{width}\t{width + 1}\n")
            nfull = L // width
            body = np.empty(L + nfull + (1 if L % width else 0), np.uint8)
            if nfull:
                blk = body[: nfull * (width + 1)].reshape(nfull, width + 1)
                blk[:, :width] = seq[: nfull * width].reshape(nfull, width)
                blk[:, width] = 10
            if L % width:
                body[nfull * (width + 1):-1] = seq[nfull * width:]
                body[-1] = 10
            f.write(body.tobytes())
            off += len(body)
    with open(path + ".fai", "w") as f:
        f.writelines(fai)


def write_depth_file(path: str, depth: np.ndarray, chunk: int = 4_000_000) -> None:
    """`pos<TAB>depth` lines, 1-based (what `rsicnv rsi -s` writes: loaddata.cpp:464-470)."""
    with open(path, "wb") as f:
        for a in range(0, len(depth), chunk):
            d = depth[a:a + chunk]
            pos = np.arange(a + 1, a + 1 + len(d))
            f.write("\n".join(f"{p}\t{v}" for p, v in zip(pos.tolist(), d.tolist())).encode())
            f.write(b"\n")


def stress_events(L: int, seed: int, n_blocks) -> list[tuple[int, int, float]]:
    """Adversarial event layout for the candidate stage: same-type neighbours separated by short gaps
    (merge paths), nested copy-number levels (multi-level RSI status -> multisegments), weak events
    that fail the depth test, events hugging contig ends and N blocks."""
    rng = np.random.default_rng(seed + 101)
    ev: list[tuple[int, int, float]] = []
    free = []
    prev = 0
    for a, b in sorted(n_blocks) + [(L, L)]:
        if a - prev > 60_000:
            free.append((prev + 3_000, a - 3_000))
        prev = b
    factors = [0.5, 1.5, 0.0, 2.0, 0.8, 1.2, 0.65, 1.35]
    for lo, hi in free:
        p = lo + int(rng.integers(500, 4_000))
        while p < hi - 40_000:
            kind = int(rng.integers(0, 5))
            f = factors[int(rng.integers(0, len(factors)))]
            ln = int(rng.choice([600, 1200, 2500, 5000, 9000, 20000]))
            if kind == 0:      # isolated
                ev.append((p, p + ln, f)); p += ln
            elif kind == 1:    # two same-type events with a short gap
                g = int(rng.choice([150, 400, 900, 2500]))
                ev.append((p, p + ln, f)); ev.append((p + ln + g, p + 2 * ln + g, f)); p += 2 * ln + g
            elif kind == 2:    # nested: outer mild, inner strong
                inner = ln // 3
                ev.append((p, p + ln, 0.6 if f < 1 else 1.4))
                ev.append((p + inner, p + 2 * inner, 0.1 if f < 1 else 2.2)); p += ln
            elif kind == 3:    # staircase of three levels
                ev.append((p, p + ln, 0.75 if f < 1 else 1.25)); ev.append((p + ln, p + 2 * ln, 0.5 if f < 1 else 1.5))
                ev.append((p + 2 * ln, p + 3 * ln, 0.25 if f < 1 else 1.75)); p += 3 * ln
            else:              # opposite types back to back
                ev.append((p, p + ln, 0.5)); ev.append((p + ln + 300, p + 2 * ln + 300, 1.5)); p += 2 * ln + 300
            p += int(rng.integers(3_000, 60_000))
    return [(s, e, f) for s, e, f in ev if e < L]


def apply_events(cn: np.ndarray, events) -> None:
    for s, e, f in events:
        cn[s:e] = f
