"""Multi-GPU plumbing: record sharding and the single exchange step of the path.

Step 1 is independent per record, so records are sharded over ranks in CONTIGUOUS ranges of the
pooled record list (reference records, then target records).  Concatenating the per-rank minimizer
lists in rank order therefore reproduces the single-GPU (record, pos) order exactly, and the
result does not depend on the GPU count.

Steps 2-3 need one exchange: uniqueness is per ASSEMBLY, not per GPU (bin/ntjoin_utils.py:182-187),
so the FULL per-rank lists (the multiset, not the locally-unique set) are all-gathered before the
global count.  torch.distributed carries it (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_ranges(offsets_per_asm, world):
    """offsets_per_asm: list of uint64 arrays (n_records+1) in assembly order.
    Returns ranges[rank][asm] = (first_record, end_record), balanced by cumulative bases."""
    lens = np.concatenate([np.diff(np.asarray(o).astype(np.int64)) for o in offsets_per_asm])
    cum = np.concatenate([[0], np.cumsum(lens)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.argmin(np.abs(cum - total * r / world))))
    cuts.append(len(lens))
    cuts = np.maximum.accumulate(cuts)
    starts = np.cumsum([0] + [len(o) - 1 for o in offsets_per_asm])
    out = []
    for r in range(world):
        a0, a1 = cuts[r], cuts[r + 1]
        per = []
        for i in range(len(offsets_per_asm)):
            lo = min(max(a0, starts[i]), starts[i + 1])
            hi = max(lo, min(a1, starts[i + 1]))
            per.append((int(lo - starts[i]), int(hi - starts[i])))
        out.append(per)
    return out


def all_gather_minimizers(hashes, contigs, first_record, group=None):
    """hashes: int64 tensor (bit pattern of the uint64 out_hash), contigs: int32 tensor of LOCAL record
    ids, both in (record, pos) order on this rank; first_record: global id of this rank's first record.
    Returns (all_hashes, all_contigs) with GLOBAL record ids, identical on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = hashes.device
    n = hashes.numel()
    counts = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, torch.tensor([n], dtype=torch.int64, device=dev), group=group)
    counts = counts.cpu().numpy()
    mx = max(1, int(counts.max()))
    hbuf = torch.zeros(mx, dtype=torch.int64, device=dev)
    cbuf = torch.zeros(mx, dtype=torch.int32, device=dev)
    if n:
        hbuf[:n] = hashes
        cbuf[:n] = contigs + int(first_record)
    gh = torch.empty(world * mx, dtype=torch.int64, device=dev)
    gc = torch.empty(world * mx, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(gh, hbuf, group=group)
    dist.all_gather_into_tensor(gc, cbuf, group=group)
    hh = torch.cat([gh[r * mx:r * mx + int(counts[r])] for r in range(world)])
    cc = torch.cat([gc[r * mx:r * mx + int(counts[r])] for r in range(world)])
    return hh, cc


class DeviceArray:
    """Zero-copy view of an engine-owned device array for torch.as_tensor (__cuda_array_interface__)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}
