"""numpy stand-ins for the four device stages of the multi-GPU steps 2-3 (include/mxe.h: mxe_dist_*).

TEST INFRASTRUCTURE: lets the CPU tests drive ntjoin_b200.dist's orchestration (real gloo collectives or the
lock-step simulator) without a GPU.  Same contracts as Engine.dist_stages(): tensors in, tensors filled in place.
The single-process semantics they distribute are bin/ntjoin_utils.py:167-193, :152-165, :94-115.
"""
import numpy as np


def _u32(t):
    return t.numpy().view(np.uint32)


def owner_of(h, world):
    return ((h >> np.uint64(32)) * np.uint64(world)) >> np.uint64(32)


class _State:
    pass


class _Shard:
    """mimics FilterResult for DistShard (fetch / edge_keys / counts / close)"""

    def __init__(self, d, keys):
        self._d, self.edge_keys = d, keys

    def fetch(self, copy=True):
        return self._d

    def counts(self):
        return sum(len(u) for u in self._d["uniq"]), len(self._d["vertices"]), len(self._d["edge_u"])

    def close(self):
        pass


class NumpyDistStages:
    def mark(self, keys, asm_off, rank, world, mk):
        st = _State()
        st.keys = keys.numpy().view(np.uint64)
        st.asm_off = np.asarray(asm_off, dtype=np.int64)
        st.n_asm, st.world, st.rank = len(asm_off) - 1, world, rank
        N = int(st.asm_off[-1])
        m = _u32(mk)
        m[:] = 0
        h_all = st.keys[:N]
        idx = np.nonzero(owner_of(h_all, world) == rank)[0]
        st.vertices = np.empty(0, dtype=np.uint64)
        if len(idx):
            h = h_all[idx]
            asm = np.searchsorted(st.asm_off, idx, side="right") - 1
            _, inv, cnt = np.unique(np.stack([h, asm.astype(np.uint64)], axis=1), axis=0, return_inverse=True, return_counts=True)
            uniq = cnt[inv.ravel()] == 1
            hu, hinv, hcnt = np.unique(h, return_inverse=True, return_counts=True)
            n_uniq = np.bincount(hinv, weights=uniq.astype(np.float64), minlength=len(hu))
            keep_h = (hcnt == st.n_asm) & (n_uniq == st.n_asm)          # once in EVERY assembly
            vid_h = np.cumsum(keep_h) - 1
            keep = keep_h[hinv]
            m[idx] = (uniq.astype(np.uint32) << np.uint32(31)) | np.where(keep, vid_h[hinv] + 1, 0).astype(np.uint32)
            st.vertices = hu[keep_h]
        return st, len(st.vertices)

    def adjacency(self, st, mk, vbase, loc_off, loc_n, contigs, succ):
        m = _u32(mk)
        s = _u32(succ)
        s[:] = 0
        vbase = np.asarray(vbase, dtype=np.int64)
        st.nV = int(vbase[-1])
        st.luniq, st.lkeep, sight = [], [], []
        cv, cg, ca, cc = [], [], [], []
        for a in range(st.n_asm):
            g = int(loc_off[a]) + np.arange(int(loc_n[a]), dtype=np.int64)
            mm = m[g] if len(g) else np.empty(0, dtype=np.uint32)
            keep = (mm & np.uint32(0x7FFFFFFF)) != 0
            st.luniq.append((mm >> np.uint32(31)).astype(bool))
            st.lkeep.append(keep)
            gk = g[keep]
            vid = vbase[owner_of(st.keys[gk], st.world).astype(np.int64)] + (mm[keep] & np.uint32(0x7FFFFFFF)).astype(np.int64) - 1
            cv.append(vid); cg.append(gk); ca.append(np.full(len(gk), a, dtype=np.int64))
            cc.append(contigs[a].numpy().astype(np.int64)[keep] if len(g) else np.empty(0, dtype=np.int64))
        st.cvid, st.cidx, st.casm, ctg = map(np.concatenate, (cv, cg, ca, cc))
        n = len(st.cvid)
        st.eflag = np.zeros(n, dtype=bool)
        if n > 1:
            st.eflag[:-1] = (st.casm[:-1] == st.casm[1:]) & (ctg[:-1] == ctg[1:])
        j = np.nonzero(st.eflag)[0]
        s[st.casm[j] * st.nV + st.cvid[j]] = (st.cvid[j + 1] + 1).astype(np.uint32)

    def edges(self, st, succ, srcmin):
        s = _u32(succ)
        sm = _u32(srcmin)
        sm[:] = 0x7F7F7F7F
        j = np.nonzero(st.eflag)[0]
        v, x = st.cvid[j], st.cvid[j + 1]
        mask = np.zeros(len(j), dtype=np.uint32)
        for b in range(st.n_asm):
            hit = (s[b * st.nV + v] == (x + 1).astype(np.uint32)) | (s[b * st.nV + x] == (v + 1).astype(np.uint32))
            mask |= hit.astype(np.uint32) << np.uint32(b)
        first = np.array([(int(mm) & -int(mm)).bit_length() - 1 for mm in mask], dtype=np.int64)
        own = first == st.casm[j]
        st.q0, st.emask = j[own], mask[own]
        np.minimum.at(sm, st.cvid[st.q0], st.cidx[st.q0].astype(np.uint32))
        return len(st.q0)

    def finish(self, st, srcmin, weights):
        sm = _u32(srcmin)
        key = (sm[st.cvid[st.q0]].astype(np.uint64) << np.uint64(32)) | st.cidx[st.q0].astype(np.uint64)
        order = np.argsort(key, kind="stable")
        q0, mask, key = st.q0[order], st.emask[order], key[order]
        w = np.zeros(len(q0), dtype=np.float64)
        for a in range(st.n_asm):                       # Python's sum(): assembly order, starting from 0
            w = np.where((mask >> np.uint32(a)) & np.uint32(1), w + float(weights[a]), w)
        d = {"uniq": st.luniq, "keep": st.lkeep, "vertices": st.vertices,
             "edge_u": st.keys[st.cidx[q0]] if len(q0) else np.empty(0, dtype=np.uint64),
             "edge_v": st.keys[st.cidx[q0 + 1]] if len(q0) else np.empty(0, dtype=np.uint64),
             "support": mask.astype(np.uint32), "weight": w}
        return _Shard(d, key)

    def abort(self, st):
        pass


# ------------------------------------------------------------------------------------------------
# all-to-all formulation (include/mxe.h: mxe_a2a_*)
# ------------------------------------------------------------------------------------------------
VBITS = 27


def _i64(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64).copy())


class NumpyA2AStages:
    def partition(self, hashes, rank, world):
        st = _State()
        st.n_asm, st.world, st.rank = len(hashes), world, rank
        hs = [h.numpy().view(np.uint64) for h in hashes]
        st.lofs = np.concatenate([[0], np.cumsum([len(h) for h in hs])]).astype(np.int64)
        st.lhash = np.concatenate(hs) if hs else np.empty(0, dtype=np.uint64)
        st.lasm = np.concatenate([np.full(len(h), a, dtype=np.int64) for a, h in enumerate(hs)]) if hs else np.empty(0, dtype=np.int64)
        own = owner_of(st.lhash, world).astype(np.int64)
        st.perm = np.argsort(own, kind="stable")
        cnt = np.zeros((world, st.n_asm), dtype=np.int64)
        np.add.at(cnt, (own, st.lasm), 1)
        return st, cnt, _i64(st.lhash[st.perm])

    def mark(self, st, recv_keys, recv_counts, ret_marks):
        h = recv_keys.numpy().view(np.uint64)
        rc = np.asarray(recv_counts, dtype=np.int64).reshape(st.world, st.n_asm)
        asm = np.concatenate([np.full(int(rc[r, a]), a, dtype=np.int64) for r in range(st.world) for a in range(st.n_asm)]) \
            if rc.sum() else np.empty(0, dtype=np.int64)
        out = _u32(ret_marks)
        st.vertices = np.empty(0, dtype=np.uint64)
        if len(h):
            _, inv, cnt = np.unique(np.stack([h, asm.astype(np.uint64)], axis=1), axis=0, return_inverse=True, return_counts=True)
            uniq = cnt[inv.ravel()] == 1
            hu, hinv, hcnt = np.unique(h, return_inverse=True, return_counts=True)
            n_uniq = np.bincount(hinv, weights=uniq.astype(np.float64), minlength=len(hu))
            keep_h = (hcnt == st.n_asm) & (n_uniq == st.n_asm)
            vid_h = np.cumsum(keep_h) - 1
            out[:len(h)] = (uniq.astype(np.uint32) << np.uint32(31)) | np.where(keep_h[hinv], vid_h[hinv] + 1, 0).astype(np.uint32)
            st.vertices = hu[keep_h]
        return len(st.vertices)

    def sightings(self, st, marks, contigs, goff, world):
        L = len(st.lhash)
        lmark = np.zeros(L, dtype=np.uint32)
        lmark[st.perm] = marks.numpy().view(np.uint32)[:L]
        keep = (lmark & np.uint32(0x7FFFFFFF)) != 0
        st.luniq = [(lmark[st.lofs[a]:st.lofs[a + 1]] >> np.uint32(31)).astype(bool) for a in range(st.n_asm)]
        st.lkeep = [keep[st.lofs[a]:st.lofs[a + 1]] for a in range(st.n_asm)]
        ctg = np.concatenate([c.numpy().astype(np.int64) for c in contigs]) if L else np.empty(0, dtype=np.int64)
        li = np.nonzero(keep)[0]
        own = owner_of(st.lhash[li], world).astype(np.uint64)
        cid = (own << np.uint64(VBITS)) | ((lmark[li] & np.uint32(0x7FFFFFFF)).astype(np.uint64) - np.uint64(1))
        a = st.lasm[li]
        g = (np.asarray(goff, dtype=np.int64)[a] + (li - st.lofs[a])).astype(np.uint64)
        h = st.lhash[li]
        j = np.nonzero((a[:-1] == a[1:]) & (ctg[li][:-1] == ctg[li][1:]))[0] if len(li) > 1 else np.empty(0, dtype=np.int64)
        vmask = np.uint64((1 << VBITS) - 1)
        succ = np.stack([h[j + 1], ((cid[j] & vmask) << np.uint64(32)) | cid[j + 1], (g[j] << np.uint64(8)) | (a[j].astype(np.uint64) << np.uint64(1))], axis=1)
        pred = np.stack([h[j], ((cid[j + 1] & vmask) << np.uint64(32)) | cid[j], (a[j].astype(np.uint64) << np.uint64(1)) | np.uint64(1)], axis=1)
        rec = np.empty((2 * len(j), 3), dtype=np.uint64)
        rec[0::2], rec[1::2] = succ, pred
        dest = np.empty(2 * len(j), dtype=np.int64)
        dest[0::2], dest[1::2] = (cid[j] >> np.uint64(VBITS)).astype(np.int64), (cid[j + 1] >> np.uint64(VBITS)).astype(np.int64)
        order = np.argsort(dest, kind="stable")
        cnt = np.bincount(dest, minlength=world).astype(np.int64)
        return cnt, _i64(rec[order].reshape(-1))

    def finish(self, st, recv_rec, n_rec, n_global, weights):
        rec = recv_rec.numpy().view(np.uint64)[:3 * n_rec].reshape(n_rec, 3)
        nV = len(st.vertices)
        vloc = (rec[:, 1] >> np.uint64(32)).astype(np.int64)
        other = (rec[:, 1] & np.uint64(0xFFFFFFFF)).astype(np.int64) + 1
        a = ((rec[:, 2] >> np.uint64(1)) & np.uint64(0x7F)).astype(np.int64)
        is_pred = (rec[:, 2] & np.uint64(1)).astype(bool)
        succ = np.zeros(st.n_asm * max(1, nV), dtype=np.int64)
        pred = np.zeros(st.n_asm * max(1, nV), dtype=np.int64)
        succ[a[~is_pred] * nV + vloc[~is_pred]] = other[~is_pred]
        pred[a[is_pred] * nV + vloc[is_pred]] = other[is_pred]
        s = np.nonzero(~is_pred)[0]
        mask = np.zeros(len(s), dtype=np.uint32)
        for b in range(st.n_asm):
            hit = (succ[b * nV + vloc[s]] == other[s]) | (pred[b * nV + vloc[s]] == other[s])
            mask |= hit.astype(np.uint32) << np.uint32(b)
        first = np.array([(int(m) & -int(m)).bit_length() - 1 for m in mask], dtype=np.int64)
        own = first == a[s]
        e, emask = s[own], mask[own]
        g = (rec[e, 2] >> np.uint64(8)).astype(np.int64)
        srcmin = np.full(max(1, nV), np.iinfo(np.int64).max, dtype=np.int64)
        np.minimum.at(srcmin, vloc[e], g)
        key = (srcmin[vloc[e]].astype(np.uint64) << np.uint64(32)) | g.astype(np.uint64)
        order = np.lexsort((a[e], srcmin[vloc[e]]))
        e, emask, key = e[order], emask[order], key[order]
        w = np.zeros(len(e), dtype=np.float64)
        for b in range(st.n_asm):
            w = np.where((emask >> np.uint32(b)) & np.uint32(1), w + float(weights[b]), w)
        d = {"uniq": st.luniq, "keep": st.lkeep, "vertices": st.vertices,
             "edge_u": st.vertices[vloc[e]] if len(e) else np.empty(0, dtype=np.uint64),
             "edge_v": rec[e, 0] if len(e) else np.empty(0, dtype=np.uint64),
             "support": emask.astype(np.uint32), "weight": w}
        return _Shard(d, key)

    def abort(self, st):
        pass
