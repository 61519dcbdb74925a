"""Contig sharding across GPUs.  The path partitions by chromosome (every iteration of the reference's loop,
rsi.cpp:2189-2217, uses chromosome-local statistics only), so ranks take whole contigs, longest first onto the
least-loaded rank (LPT), and the only exchange is the ordered gather of the call rows to rank 0 (rows are written
in BAM-header order, rsi.cpp:2116-2127, 1594)."""
from __future__ import annotations


def lpt_assign(lengths: list[int], world: int) -> list[int]:
    """rank of every contig; ties keep header order (same rule as rsicnv_b200/host/main.cpp)"""
    order = sorted(range(len(lengths)), key=lambda i: -lengths[i])
    load = [0] * world
    out = [0] * len(lengths)
    for i in order:
        r = min(range(world), key=lambda k: load[k])
        out[i] = r
        load[r] += lengths[i]
    return out


def imbalance(lengths: list[int], world: int) -> float:
    a = lpt_assign(lengths, world)
    load = [0] * world
    for i, r in enumerate(a):
        load[r] += lengths[i]
    return max(load) / (sum(lengths) / world)


def gather_rows(rows_by_contig: dict[int, list[str]], n_contigs: int, rank: int, world: int, group=None) -> list[str] | None:
    """ordered gather of the per-contig table rows to rank 0 (torch.distributed, any backend)"""
    import torch.distributed as dist
    if world == 1:
        return [r for t in range(n_contigs) for r in rows_by_contig.get(t, [])]
    out = [None] * world if rank == 0 else None
    dist.gather_object(rows_by_contig, out, dst=0, group=group)
    if rank != 0:
        return None
    merged: dict[int, list[str]] = {}
    for d in out:
        merged.update(d)
    return [r for t in range(n_contigs) for r in merged.get(t, [])]
