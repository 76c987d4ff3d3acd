"""Multi-GPU plumbing: record sharding and the exchange steps of the path.

Step 1 is independent per record, so records are sharded over ranks in CONTIGUOUS ranges of the
pooled record list (reference records, then target records).  Concatenating the per-rank minimizer
lists in rank order therefore reproduces the single-GPU (record, pos) order exactly, and the
result does not depend on the GPU count.

Steps 2-3 need exchanges: uniqueness is per ASSEMBLY, not per GPU (bin/ntjoin_utils.py:182-187), and
the intersection is over all assemblies (:155-157).  `distributed_filter_and_edges` runs them with
every rank doing 1/world of the work on globally indexed arrays:

    all-gather(hashes)      FULL per-rank lists (the multiset, not the locally-unique sets)
    stage mark              rank r owns the hash range r: unique / found-in-all / vertex ids
    all-reduce(sum, mk)     N x u32 marks, zero outside the owned entries (+ the per-rank vertex counts)
    stage adjacency         own records: ordered survivors, adjacent pairs -> successor table
    all-reduce(sum, succ)   n_asm x nV x u32, zero outside own sightings
    stage edges             support masks + ownership of the local sightings, first-source table
    all-reduce(min, srcmin) nV x u32
    stage finish            local edge shard with global order keys

The stages are device kernels behind the C ABI (Engine.dist_stages()); the collectives are
torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests).  The orchestration is a
generator that yields one request per collective, so the same code is driven by a real process
group (`TorchComm`) or, in tests, by several simulated ranks in lock-step (`run_lockstep`).
"""
import numpy as np


def shard_ranges(offsets_per_asm, world):
    """offsets_per_asm: list of uint64 arrays (n_records+1) in assembly order.
    Returns ranges[rank][asm] = (first_record, end_record): every rank gets ONE contiguous record range of EVERY
    assembly, so concatenating the ranks' lists per assembly keeps the (record, pos) order.  Cuts fall on record
    boundaries; the assemblies are cut one after the other (coarsest records first) and each later assembly
    compensates the imbalance left by the earlier ones, so the per-rank totals track total/world as closely as the
    finest assembly allows (chromosome-scale references are evened out by the target's contigs)."""
    offs = [np.asarray(o).astype(np.int64) for o in offsets_per_asm]
    n_asm = len(offs)
    totals = [int(o[-1] - o[0]) for o in offs]
    order = sorted(range(n_asm), key=lambda a: -(int(np.diff(offs[a]).max()) if len(offs[a]) > 1 else 0))
    loads = np.zeros(world, dtype=np.int64)
    cuts = [None] * n_asm
    done = 0
    for a in order:
        o = offs[a] - offs[a][0]
        done += totals[a]
        c = [0]
        assigned = 0                                   # bases of this assembly given to lower ranks
        for r in range(world - 1):
            want = done * (r + 1) // world - int(loads[:r + 1].sum()) + 0   # bases this assembly should add to ranks 0..r
            want = min(max(want, assigned), totals[a])
            j = int(np.searchsorted(o, want))
            if j > 0 and (j >= len(o) or want - o[j - 1] <= o[j] - want):
                j -= 1
            j = min(max(j, c[-1]), len(o) - 1)
            c.append(j)
            assigned = int(o[j])
        c.append(len(o) - 1)
        cuts[a] = c
        for r in range(world):
            loads[r] += int(o[c[r + 1]] - o[c[r]])
    return [[(cuts[a][r], cuts[a][r + 1]) for a in range(n_asm)] for r in range(world)]


def gather_and_write_tsv(path, out_hash, pos, contig, forward, first_record, names, k, fasta_path=None,
                         with_pos=True, with_strand=False, with_seq=True, group=None):
    """Seam S2 with several GPUs: every rank passes the minimizers of ITS record range (contig = local record ids,
    first_record = global id of its first record; ranks hold contiguous ranges in rank order) and rank 0 writes the
    assembly's one `<fasta>.k<k>.w<w>.tsv` exactly as a single process would (`names` = all record names; the
    k-mer text of --seq comes from `fasta_path`, read on rank 0 with the engine's host reader)."""
    import ctypes as C
    import torch.distributed as dist
    from ._lib import check, load_library
    from .engine import HostSketch
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = (np.asarray(out_hash, dtype=np.uint64), np.asarray(pos, dtype=np.uint32),
            np.asarray(contig, dtype=np.uint32) + np.uint32(first_record), np.asarray(forward, dtype=np.uint8))
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0, group=group)
    if rank != 0:
        return None
    oh, ps, cg, fw = (np.concatenate([p[i] for p in parts]) for i in range(4))
    offsets = seq = handle = None
    lib = load_library()
    try:
        if with_seq:
            handle = C.c_void_p()
            check(lib, lib.mxe_fasta_read(str(fasta_path).encode(), C.byref(handle)))
            n, offs, text = C.c_uint32(), C.c_void_p(), C.c_void_p()
            check(lib, lib.mxe_fasta_view(handle, C.byref(n), C.byref(offs), C.byref(text)))
            if n.value != len(names):
                raise ValueError(f"{fasta_path} has {n.value} records, {len(names)} names were given")
            offsets = np.ctypeslib.as_array(C.cast(offs, C.POINTER(C.c_uint64)), shape=(n.value + 1,))
            total = int(offsets[-1])
            seq = np.ctypeslib.as_array(C.cast(text, C.POINTER(C.c_uint8)), shape=(max(1, total),))
        hs = HostSketch(oh, ps, cg, names, k, forward=fw, offsets=offsets, seq=seq)
        try:
            hs.write_tsv(path, pos=with_pos, strand=with_strand, seq=with_seq)
        finally:
            hs.close()
    finally:
        if handle is not None:
            lib.mxe_fasta_free(handle)
    return len(oh)


class DeviceArray:
    """Zero-copy view of an engine-owned device array for torch.as_tensor (__cuda_array_interface__)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


# ------------------------------------------------------------------------------------------------
# distributed steps 2-3
# ------------------------------------------------------------------------------------------------
class Layout:
    """Global minimizer index space: assemblies in order, inside an assembly the ranks in order."""

    def __init__(self, counts):
        self.counts = np.asarray(counts, dtype=np.int64)            # [world, n_asm]
        self.world, self.n_asm = self.counts.shape
        self.asm_off = np.concatenate([[0], np.cumsum(self.counts.sum(axis=0))]).astype(np.int64)
        before = np.cumsum(self.counts, axis=0) - self.counts      # minimizers of lower ranks, per assembly
        self.goff = self.asm_off[:-1][None, :] + before             # [world, n_asm] global index of each slice
        self.N = int(self.asm_off[-1])


def _filter_steps(stages, hashes, contigs, weights, rank, world, device):
    """Generator: yields (op, payload) per collective and receives its result.
    hashes[a]: int64 tensor (bit pattern of the uint64 out_hash) of this rank's minimizers of assembly a
    in (record, pos) order; contigs[a]: int32 tensor of their (local) record ids."""
    import torch
    n_asm = len(hashes)
    counts = yield ("counts", [int(h.numel()) for h in hashes])
    lay = Layout(counts)
    keys = yield ("keys", (hashes, lay))                         # int64[N] in the global layout
    mk = torch.empty(lay.N + world, dtype=torch.int32, device=device)   # marks + one slot per rank for its vertex count
    handle, nv_local = stages.mark(keys, lay.asm_off, rank, world, mk)
    try:
        mk[lay.N:].zero_()
        mk[lay.N + rank] = int(nv_local)
        yield ("sum", mk)                                          # the vertex counts ride along with the marks
        nvs = mk[lay.N:].cpu().numpy().astype(np.int64)
        vbase = np.concatenate([[0], np.cumsum(nvs)]).astype(np.int64)
        n_v = int(vbase[-1])
        succ = torch.empty(n_asm * max(1, n_v), dtype=torch.int32, device=device)
        stages.adjacency(handle, mk, vbase, lay.goff[rank], lay.counts[rank], contigs, succ)
        yield ("sum", succ)
        srcmin = torch.empty(max(1, n_v), dtype=torch.int32, device=device)
        n_e_local = stages.edges(handle, succ, srcmin)
        yield ("min", srcmin)
    except BaseException:
        stages.abort(handle)
        raise
    shard = stages.finish(handle, srcmin, weights)
    return DistShard(shard, lay, rank, vbase, int(n_e_local), keep=(keys, mk, succ, srcmin))


def _a2a_steps(stages, hashes, contigs, weights, rank, world, device):
    """All-to-all formulation (traffic and work shrink with 1/world).  Same inputs and result as `_filter_steps`; every item travels once:
    keys to their hash owner, marks back, sightings to the owners of their two vertices."""
    import torch
    n_asm = len(hashes)
    handle, cnt, send_keys = stages.partition(hashes, rank, world)            # cnt[owner][asm]
    try:
        allc = yield ("counts", [int(x) for x in cnt.reshape(-1)])
        c3 = allc.reshape(world, world, n_asm)                                # [source][owner][asm]
        send_splits = c3[rank].sum(axis=1)
        recv_counts = np.ascontiguousarray(c3[:, rank, :])                    # [source][asm] addressed to this owner
        recv_splits = recv_counts.sum(axis=1)
        n_recv = int(recv_splits.sum())
        recv_keys = yield ("a2a", (send_keys, send_splits, recv_splits, 1))
        ret_marks = torch.empty(max(1, n_recv), dtype=torch.int32, device=device)
        nv_local = stages.mark(handle, recv_keys, recv_counts, ret_marks)
        marks = yield ("a2a", (ret_marks[:n_recv], recv_splits, send_splits, 1))   # same order as send_keys
        lay = Layout(c3.sum(axis=1))                                          # local counts of every rank -> global indices
        rec_cnt, send_rec = stages.sightings(handle, marks, contigs, lay.goff[rank], world)
        allr = yield ("counts", [int(x) for x in rec_cnt] + [int(nv_local)])
        rec_send, rec_recv = allr[rank, :world], allr[:, rank]
        vbase = np.concatenate([[0], np.cumsum(allr[:, world])]).astype(np.int64)
        recv_rec = yield ("a2a", (send_rec, rec_send, rec_recv, 3))
    except BaseException:
        stages.abort(handle)
        raise
    n_rec = int(rec_recv.sum())
    shard = stages.finish(handle, recv_rec, n_rec, lay.N, weights)
    return DistShard(shard, lay, rank, vbase, None, keep=(recv_keys, ret_marks, marks, recv_rec))


class DistShard:
    """This rank's part of the result of steps 2-3: flags of its own minimizers, the vertices of its
    hash range (ascending) and its edges with global order keys.  `merge_shards` reassembles the
    single-GPU result from the fetched shards of all ranks."""

    def __init__(self, result, layout, rank, vbase, n_edges_local, keep=None):
        self.result, self.layout, self.rank, self.vbase, self.n_edges_local = result, layout, rank, vbase, n_edges_local
        self._keep = keep

    def counts(self):
        return self.result.counts()

    def fetch(self, copy=True):
        d = dict(self.result.fetch(copy=copy))
        n_e = self.result.counts()[2] if self.n_edges_local is None else self.n_edges_local
        d["edge_key"] = self.result.edge_keys if n_e else np.empty(0, dtype=np.uint64)
        return d

    def close(self):
        self.result.close()
        self._keep = None


def merge_shards(shards):
    """shards: fetched dicts of all ranks, in rank order.  Returns the single-GPU result layout:
    uniq/keep per assembly, vertices ascending, edges in the order of bin/ntjoin_utils.py:115."""
    n_asm = len(shards[0]["uniq"])
    out = {"uniq": [np.concatenate([s["uniq"][a] for s in shards]) for a in range(n_asm)],
           "keep": [np.concatenate([s["keep"][a] for s in shards]) for a in range(n_asm)],
           "vertices": np.concatenate([s["vertices"] for s in shards])}
    key = np.concatenate([s["edge_key"] for s in shards])
    order = np.argsort(key, kind="stable")
    for name in ("edge_u", "edge_v", "support", "weight"):
        out[name] = np.concatenate([s[name] for s in shards])[order]
    return out


class TorchComm:
    """Collectives of `_filter_steps` over a torch.distributed process group."""

    def __init__(self, device, group=None, timing=False):
        import torch.distributed as dist
        self.dist, self.group, self.device = dist, group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.timing, self.spans = timing, []          # (op, start event, end event) on the current stream

    def counts(self, values):
        import torch
        mine = torch.tensor(values, dtype=torch.int64, device=self.device)
        out = torch.empty(self.world * len(values), dtype=torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(out, mine, group=self.group)
        return out.cpu().numpy().reshape(self.world, len(values))

    def keys(self, hashes, lay):
        """one padded all-gather of the concatenated local lists, then slice copies into the global layout"""
        import torch
        local_n = lay.counts.sum(axis=1)
        mx = max(1, int(local_n.max()))
        buf = torch.empty(mx, dtype=torch.int64, device=self.device)
        at = 0
        for h in hashes:
            if h.numel():
                buf[at:at + h.numel()] = h
            at += h.numel()
        if at < mx:
            buf[at:].zero_()
        gathered = torch.empty(self.world * mx, dtype=torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(gathered, buf, group=self.group)
        keys = torch.empty(max(1, lay.N), dtype=torch.int64, device=self.device)
        for r in range(self.world):
            at = r * mx
            for a in range(lay.n_asm):
                n = int(lay.counts[r, a])
                if n:
                    g = int(lay.goff[r, a])
                    keys[g:g + n] = gathered[at:at + n]
                at += n
        return keys

    def a2a(self, send, send_splits, recv_splits, width):
        """all_to_all_single of `width` elements per item; splits in items"""
        import torch
        n_out = int(np.sum(recv_splits)) * width
        out = torch.empty(max(1, n_out), dtype=send.dtype, device=self.device)[:n_out]
        self.dist.all_to_all_single(out, send, output_split_sizes=[int(x) * width for x in recv_splits],
                                    input_split_sizes=[int(x) * width for x in send_splits], group=self.group)
        return out

    def reduce(self, op, tensor):
        self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM if op == "sum" else self.dist.ReduceOp.MIN, group=self.group)

    def execute(self, req):
        op, payload = req
        if self.timing:
            import torch
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        if op == "counts":
            out = self.counts(payload)
        elif op == "keys":
            out = self.keys(*payload)
        elif op == "a2a":
            out = self.a2a(*payload)
        else:
            out = self.reduce(op, payload)
        if self.timing:
            ev1.record()
            self.spans.append((op, ev0, ev1))
        return out

    def timing_ms(self, reset=True):
        """device milliseconds per collective kind since the last reset (needs timing=True)"""
        import torch
        torch.cuda.synchronize()
        out = {}
        for op, a, b in self.spans:
            out[op] = out.get(op, 0.0) + a.elapsed_time(b)
        if reset:
            self.spans = []
        return out


def distributed_filter_and_edges(stages, hashes, contigs, weights, comm):
    """Steps 2-3 across the ranks of `comm` (a TorchComm).  Returns this rank's DistShard.
    `stages` selects the formulation: Engine.dist_stages() (all-reduce; default, faster while a step is latency-bound)
    or Engine.a2a_stages() (all-to-all; per-rank work and traffic shrink with 1/world).  The engine must run on torch's current stream (Engine.set_stream)."""
    steps = _a2a_steps if hasattr(stages, "partition") else _filter_steps
    gen = steps(stages, hashes, contigs, weights, comm.rank, comm.world, comm.device)
    try:
        req = next(gen)
        while True:
            req = gen.send(comm.execute(req))
    except StopIteration as stop:
        return stop.value


def run_lockstep(stage_list, hashes_per_rank, contigs_per_rank, weights, device):
    """Drive `world` simulated ranks of `_filter_steps` in one process (tests: the same orchestration and
    the same device stages as the real multi-process run, with the collectives done in place)."""
    import torch
    world = len(stage_list)
    steps = _a2a_steps if hasattr(stage_list[0], "partition") else _filter_steps
    gens = [steps(stage_list[r], hashes_per_rank[r], contigs_per_rank[r], weights, r, world, device) for r in range(world)]
    reqs = [next(g) for g in gens]
    results = [None] * world
    while any(r is not None for r in reqs):
        op = reqs[0][0]
        assert all(r[0] == op for r in reqs)
        if op == "counts":
            resp = [np.asarray([r[1] for r in reqs], dtype=np.int64)] * world
        elif op == "keys":
            lay = reqs[0][1][1]
            keys = torch.empty(max(1, lay.N), dtype=torch.int64, device=device)
            for r in range(world):
                for a in range(lay.n_asm):
                    n = int(lay.counts[r, a])
                    if n:
                        g = int(lay.goff[r, a])
                        keys[g:g + n] = reqs[r][1][0][a]
            resp = [keys.clone() for _ in range(world)]
        elif op == "a2a":
            resp = []
            for d in range(world):
                parts = []
                for src in range(world):
                    send, send_splits, _recv, width = reqs[src][1]
                    at = int(np.sum(send_splits[:d])) * width
                    parts.append(send[at:at + int(send_splits[d]) * width])
                resp.append(torch.cat(parts) if parts else torch.empty(0, dtype=reqs[0][1][0].dtype, device=device))
        else:
            stack = torch.stack([r[1] for r in reqs])
            red = stack.sum(dim=0, dtype=torch.int32) if op == "sum" else stack.min(dim=0).values
            for r in reqs:
                r[1].copy_(red)
            resp = [None] * world
        nxt = []
        for i, g in enumerate(gens):
            try:
                nxt.append(g.send(resp[i]))
            except StopIteration as stop:
                results[i] = stop.value
                nxt.append(None)
        reqs = nxt
    return results
