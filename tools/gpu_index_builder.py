"""GPU index builder for bench/test SET-UP (not part of the search hot path, not shipped in the
C-ABI library).

The reference's own builder (BKT::Index::BuildIndex: balanced k-means tree + TP-tree KNN + refine,
BKTIndex.cpp:739-772) needs hours for 1M x 768 on a few host cores, so the
bench cannot build its index with it inside one GPU lease.  This module builds, with plain torch
ops on the GPU, an index with the SAME data structures the reference searches and persists:

  * a BKT (BKTree.h:25-32, :546-627): every data point is exactly one tree node; internal nodes hold
    the member nearest to a k-means centroid (K = BKTKmeansK = 32), sets <= BKTLeafSize (8) become leaf
    children, the root has centerid = N, a (-1,-1,-1) sentinel ends the array;
  * a fixed-degree relative-neighbourhood graph (RelativeNeighborhoodGraph.h:18-35 prune rule applied
    to the `cand` nearest neighbours of every point), rows -1 padded (NeighborhoodGraph.h:626-649);

and writes them in the reference's on-disk format (indexloader.ini, vectors.bin, tree.bin, graph.bin,
deletes.bin: VectorIndex.cpp:197-222, Dataset.h:146-180, BKTree.h:635-645, NeighborhoodGraph.h:606-615,
Labelset.h:78-83), so the UNMODIFIED reference loads the very same folder (VectorIndex::LoadIndex) for
the CPU baseline.  Parity is defined on identical index files, so any index the reference can load is
a legitimate input; this one is merely fast to make.
"""
import math
import os
import time

import numpy as np
import torch


def _sq_norms(x, chunk=4000000):
    """Row-wise squared norms without materialising x**2 for the whole array."""
    if x.shape[0] <= chunk:
        return (x.float() ** 2).sum(1)
    out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    for s in range(0, x.shape[0], chunk):
        out[s:s + chunk] = (x[s:s + chunk].float() ** 2).sum(1)
    return out


# ---------------------------------------------------------------------------------------------
# BKT
# ---------------------------------------------------------------------------------------------
def _assign_dense(x, xn, pts, slot, cent, centn, K, chunk_cols_budget=1 << 27):
    """For points `pts` (ids) whose group slot is `slot`, return argmin over the K centroids of that
    slot and the distance.  cent: [P*K, dim] (slot-major).  Dense GEMM against all P*K centroids,
    then each point gathers its own K columns."""
    P = cent.shape[0] // K
    n = pts.shape[0]
    best = torch.empty(n, dtype=torch.int64, device=x.device)
    bestd = torch.empty(n, dtype=torch.float32, device=x.device)
    chunk = max(256, min(n, chunk_cols_budget // max(1, P * K)))
    ar = torch.arange(K, device=x.device)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        ids = pts[s:e]
        xs = x[ids]
        if P == 1:
            d = xn[ids][:, None] - 2.0 * (xs @ cent.t()) + centn[None, :]
        else:
            full = xs @ cent.t()  # [c, P*K]
            cols = slot[s:e, None] * K + ar[None, :]
            d = xn[ids][:, None] - 2.0 * torch.gather(full, 1, cols) + centn[cols]
            del full
        dm, am = d.min(dim=1)
        best[s:e] = am
        bestd[s:e] = dm
    return best, bestd


def build_bkt(x, kmeans_k=32, leaf_size=8, iters=2, seed=0, log=None):
    """Returns (nodes int32 [N+2, 3] on CPU, tree_starts int32 [1])."""
    dev = x.device
    N = x.shape[0]
    K = kmeans_k
    final_t = K * (leaf_size + 1)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    xn = _sq_norms(x)
    nodes = torch.full((N + 2, 3), -1, dtype=torch.int32, device=dev)
    nodes[0, 0] = N
    next_free = 1
    # open groups: parent node index per group, members sorted by group, offsets
    grp_parent = torch.zeros(1, dtype=torch.int64, device=dev)
    members = torch.randperm(N, generator=g, device=dev)
    offsets = torch.tensor([0, N], dtype=torch.int64, device=dev)
    level = 0
    while grp_parent.numel() > 0:
        P = grp_parent.numel()
        sizes = offsets[1:] - offsets[:-1]
        gid = torch.repeat_interleave(torch.arange(P, device=dev), sizes)  # group of each member
        pos = torch.arange(members.numel(), device=dev) - offsets[:-1][gid]  # position inside its group
        small = sizes <= leaf_size
        big = sizes > final_t
        mid = ~small & ~big

        # number of children per group
        nchild = torch.zeros(P, dtype=torch.int64, device=dev)
        nchild[small] = sizes[small]
        nchild[mid] = (sizes[mid] + leaf_size) // (leaf_size + 1)

        # ---- big groups: k-means with K centroids, centre = member nearest its centroid ----
        child_of_member = torch.full_like(members, -1)  # local child rank (0..nchild-1) of each member
        is_center = torch.zeros_like(members, dtype=torch.bool)
        if big.any():
            bslot_of_group = torch.full((P,), -1, dtype=torch.int64, device=dev)
            bidx = torch.nonzero(big).flatten()
            Pb = bidx.numel()
            bslot_of_group[bidx] = torch.arange(Pb, device=dev)
            msel = big[gid]
            bm = members[msel]
            bslot = bslot_of_group[gid[msel]]
            bpos = pos[msel]
            # init: the first K members of each (already shuffled) group
            cent = torch.zeros((Pb * K, x.shape[1]), dtype=torch.float32, device=dev)
            init = bpos < K
            cent[bslot[init] * K + bpos[init]] = x[bm[init]].float()
            for it in range(iters + 1):
                centn = _sq_norms(cent)
                a, ad = _assign_dense(x, xn, bm, bslot, cent, centn, K)
                key = bslot * K + a
                if it == iters:
                    break
                cnt = torch.bincount(key, minlength=Pb * K).float()
                newc = torch.zeros_like(cent)
                newc.index_add_(0, key, x[bm].float())
                nz = cnt > 0
                cent[nz] = newc[nz] / cnt[nz, None]
            cnt = torch.bincount(key, minlength=Pb * K)
            # centre of each non-empty cluster = member with the smallest distance to the centroid
            mind = torch.full((Pb * K,), float("inf"), device=dev)
            mind.scatter_reduce_(0, key, ad, reduce="amin")
            cand = torch.nonzero(ad <= mind[key]).flatten()
            first = torch.full((Pb * K,), bm.numel(), dtype=torch.int64, device=dev)
            first.scatter_reduce_(0, key[cand], cand, reduce="amin")
            nonempty = cnt > 0
            # local child rank of each cluster inside its group
            rank = (torch.cumsum(nonempty.view(Pb, K).long(), dim=1) - 1).view(-1)
            nchild[bidx] = nonempty.view(Pb, K).sum(1)
            midx = torch.nonzero(msel).flatten()
            child_of_member[midx] = rank[key]
            cpos = first[nonempty]
            is_center[midx[cpos]] = True

        # ---- mid groups: ceil(s / (leaf+1)) children, first c members are centres, rest round-robin ----
        if mid.any():
            msel = mid[gid]
            c = nchild[gid[msel]]
            p = pos[msel]
            midx = torch.nonzero(msel).flatten()
            child_of_member[midx] = p % c
            is_center[midx[p < c]] = True

        # ---- small groups: every member is a leaf child ----
        if small.any():
            msel = small[gid]
            midx = torch.nonzero(msel).flatten()
            child_of_member[midx] = pos[msel]
            is_center[midx] = True

        # ---- allocate child nodes (contiguous per parent) ----
        cstart = next_free + torch.cumsum(nchild, 0) - nchild
        nodes[grp_parent, 1] = cstart.int()
        nodes[grp_parent, 2] = (cstart + nchild).int()
        total_children = int(nchild.sum().item())
        child_node = cstart[gid] + child_of_member  # node index of the child each member falls under
        cm = torch.nonzero(is_center).flatten()
        nodes[child_node[cm], 0] = members[cm].int()
        next_free += total_children

        # ---- next level: non-centre members grouped by child node ----
        rest = torch.nonzero(~is_center).flatten()
        if rest.numel() == 0:
            break
        ckey = child_node[rest]
        order = torch.argsort(ckey, stable=True)
        ckey = ckey[order]
        members = members[rest][order]
        uniq, counts = torch.unique_consecutive(ckey, return_counts=True)
        grp_parent = uniq
        offsets = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(counts, 0)])
        if log:
            log("bkt level %d: %d groups, %d nodes so far" % (level, P, next_free))
        level += 1
    assert next_free == N + 1, (next_free, N)
    return nodes.cpu().numpy(), np.array([0], np.int32)


# ---------------------------------------------------------------------------------------------
# BKT for very large N: same node layout, but a big group is cut into its <= 32 children by five rounds of
# balanced random-direction bisection (O(N dim) per round) instead of k-means against all centroids of the level
# (O(N * centroids * dim), which stops being practical around 10^7-10^8 points).  The centre of a child is the
# member nearest to the child's mean, exactly as in build_bkt.
# ---------------------------------------------------------------------------------------------
def _segment_stats(seg, nseg):
    cnt = torch.bincount(seg, minlength=nseg)
    off = torch.cumsum(cnt, 0) - cnt
    return cnt, off


def build_bkt_balanced(x, kmeans_k=32, leaf_size=8, seed=0, log=None):
    dev = x.device
    N, dim = x.shape
    K = kmeans_k
    rounds = int(math.log2(K))
    assert (1 << rounds) == K, "balanced builder needs a power-of-two fan-out"
    final_t = K * (leaf_size + 1)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    nodes = torch.full((N + 2, 3), -1, dtype=torch.int32, device=dev)
    nodes[0, 0] = N
    next_free = 1
    grp_parent = torch.zeros(1, dtype=torch.int64, device=dev)
    members = torch.randperm(N, generator=g, device=dev)
    offsets = torch.tensor([0, N], dtype=torch.int64, device=dev)
    chunk = max(1, (1 << 26) // dim)
    level = 0
    while grp_parent.numel() > 0:
        P = grp_parent.numel()
        sizes = offsets[1:] - offsets[:-1]
        gid = torch.repeat_interleave(torch.arange(P, device=dev), sizes)
        pos = torch.arange(members.numel(), device=dev) - offsets[:-1][gid]
        small = sizes <= leaf_size
        big = sizes > final_t
        mid = ~small & ~big
        nchild = torch.zeros(P, dtype=torch.int64, device=dev)
        nchild[small] = sizes[small]
        nchild[mid] = (sizes[mid] + leaf_size) // (leaf_size + 1)
        child_of_member = torch.full_like(members, -1)
        is_center = torch.zeros_like(members, dtype=torch.bool)

        if big.any():
            bidx = torch.nonzero(big).flatten()
            Pb = bidx.numel()
            brank = torch.full((P,), -1, dtype=torch.int64, device=dev)
            brank[bidx] = torch.arange(Pb, device=dev)
            msel = big[gid]
            midx = torch.nonzero(msel).flatten()          # positions in `members`
            bm = members[midx]
            seg = brank[gid[midx]]                          # one segment per big group to start with
            nseg = Pb
            for _ in range(rounds):
                cnt, off = _segment_stats(seg, nseg)
                # members are kept sorted by segment, so segment s occupies [off[s], off[s]+cnt[s])
                ra = off + (torch.rand(nseg, generator=g, device=dev) * cnt).long().clamp_max_(1 << 62)
                rb = off + (torch.rand(nseg, generator=g, device=dev) * cnt).long()
                ra = torch.minimum(ra, off + cnt - 1)
                rb = torch.minimum(rb, off + cnt - 1)
                dirs = x[bm[ra]] - x[bm[rb]]
                proj = torch.empty(bm.numel(), device=dev)
                for s0 in range(0, bm.numel(), chunk):
                    e0 = min(bm.numel(), s0 + chunk)
                    proj[s0:e0] = (x[bm[s0:e0]] * dirs[seg[s0:e0]]).sum(1)
                order = torch.argsort(proj, stable=True)
                order = order[torch.argsort(seg[order], stable=True)]
                bm = bm[order]
                midx = midx[order]
                seg = seg[order]
                p_in = torch.arange(bm.numel(), device=dev) - off[seg]
                seg = seg * 2 + (p_in >= (cnt[seg] + 1) // 2).long()
                nseg *= 2
            # child segments: mean, centre = member nearest to its segment mean
            cnt, off = _segment_stats(seg, nseg)
            mean = torch.zeros((nseg, dim), device=dev)
            for s0 in range(0, bm.numel(), chunk):
                e0 = min(bm.numel(), s0 + chunk)
                mean.index_add_(0, seg[s0:e0], x[bm[s0:e0]])
            mean /= cnt.clamp_min(1)[:, None].float()
            dmean = torch.empty(bm.numel(), device=dev)
            for s0 in range(0, bm.numel(), chunk):
                e0 = min(bm.numel(), s0 + chunk)
                dmean[s0:e0] = ((x[bm[s0:e0]] - mean[seg[s0:e0]]) ** 2).sum(1)
            mind = torch.full((nseg,), float("inf"), device=dev)
            mind.scatter_reduce_(0, seg, dmean, reduce="amin")
            candp = torch.nonzero(dmean <= mind[seg]).flatten()
            first = torch.full((nseg,), bm.numel(), dtype=torch.int64, device=dev)
            first.scatter_reduce_(0, seg[candp], candp, reduce="amin")
            nonempty = cnt > 0
            rank = (torch.cumsum(nonempty.view(Pb, K).long(), dim=1) - 1).view(-1)
            nchild[bidx] = nonempty.view(Pb, K).sum(1)
            child_of_member[midx] = rank[seg]
            is_center[midx[first[nonempty]]] = True
            # (members/offsets order is irrelevant below: everything is expressed per position in `members`)

        if mid.any():
            msel = mid[gid]
            c = nchild[gid[msel]]
            p_ = pos[msel]
            mi = torch.nonzero(msel).flatten()
            child_of_member[mi] = p_ % c
            is_center[mi[p_ < c]] = True
        if small.any():
            msel = small[gid]
            mi = torch.nonzero(msel).flatten()
            child_of_member[mi] = pos[msel]
            is_center[mi] = True

        cstart = next_free + torch.cumsum(nchild, 0) - nchild
        nodes[grp_parent, 1] = cstart.int()
        nodes[grp_parent, 2] = (cstart + nchild).int()
        child_node = cstart[gid] + child_of_member
        cm = torch.nonzero(is_center).flatten()
        nodes[child_node[cm], 0] = members[cm].int()
        next_free += int(nchild.sum().item())

        rest = torch.nonzero(~is_center).flatten()
        if rest.numel() == 0:
            break
        ckey = child_node[rest]
        order = torch.argsort(ckey, stable=True)
        ckey = ckey[order]
        members = members[rest][order]
        uniq, counts = torch.unique_consecutive(ckey, return_counts=True)
        grp_parent = uniq
        offsets = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(counts, 0)])
        if log:
            log("bkt(balanced) level %d: %d groups, %d nodes so far" % (level, P, next_free))
        level += 1
    assert next_free == N + 1, (next_free, N)
    return nodes.cpu().numpy(), np.array([0], np.int32)


# ---------------------------------------------------------------------------------------------
# KD-tree (KDTree.h:22-28 node, :61-446 build): split on the dimension of largest variance at the mean,
# values < split go left (Subdivide, KDTree.h:401-446), an all-equal node is split evenly, single points
# become leaves encoded as -(id)-1.  Level-synchronous on the GPU.
# ---------------------------------------------------------------------------------------------
def build_kdt(x, log=None):
    """Returns (nodes [N, 4] as int32 view {left, right, split_dim, split_value bits}, tree_starts [1])."""
    dev = x.device
    N, dim = x.shape
    left = torch.zeros(N, dtype=torch.int32, device=dev)
    right = torch.zeros(N, dtype=torch.int32, device=dev)
    sdim = torch.zeros(N, dtype=torch.int32, device=dev)
    sval = torch.zeros(N, dtype=torch.float32, device=dev)
    perm = torch.arange(N, device=dev)
    seg_node = torch.zeros(1, dtype=torch.int64, device=dev)        # tree node index of each open segment
    sizes = torch.tensor([N], dtype=torch.int64, device=dev)
    next_free = 1
    level = 0
    while seg_node.numel() > 0:
        S = seg_node.numel()
        seg = torch.repeat_interleave(torch.arange(S, device=dev), sizes)
        offs = torch.cumsum(sizes, 0) - sizes
        pos = torch.arange(perm.numel(), device=dev) - offs[seg]
        xs = x[perm]
        cnt = sizes.float()[:, None]
        mean = torch.zeros((S, dim), device=dev).index_add_(0, seg, xs) / cnt
        var = torch.zeros((S, dim), device=dev).index_add_(0, seg, xs * xs) / cnt - mean * mean
        sd = var.argmax(1)
        sv = mean[torch.arange(S, device=dev), sd]
        val = xs[torch.arange(xs.shape[0], device=dev), sd[seg]]
        go_right = val >= sv[seg]
        nright = torch.zeros(S, dtype=torch.int64, device=dev).index_add_(0, seg, go_right.long())
        degenerate = (nright == 0) | (nright == sizes)
        if degenerate.any():  # all equal along the split dimension: split evenly (KDTree.h:438-443)
            half = (sizes // 2)[seg]
            go_right = torch.where(degenerate[seg], pos >= half, go_right)
            nright = torch.zeros(S, dtype=torch.int64, device=dev).index_add_(0, seg, go_right.long())
        nleft = sizes - nright
        sdim[seg_node] = sd.int()
        sval[seg_node] = sv
        # order members: by segment, left side first
        order = torch.argsort(seg * 2 + go_right.long(), stable=True)
        perm = perm[order]
        child_sizes = torch.stack([nleft, nright], 1).reshape(-1)    # [2S] in (seg, side) order
        child_first = torch.cumsum(child_sizes, 0) - child_sizes
        is_leaf = child_sizes == 1
        internal = ~is_leaf
        n_int = int(internal.sum().item())
        child_index = torch.full((2 * S,), -1, dtype=torch.int64, device=dev)
        child_index[internal] = next_free + torch.arange(n_int, device=dev)
        child_index[is_leaf] = -perm[child_first[is_leaf]] - 1
        ci = child_index.view(S, 2)
        left[seg_node] = ci[:, 0].int()
        right[seg_node] = ci[:, 1].int()
        next_free += n_int
        # next level: members of internal children, in order
        keep_child = internal
        child_of_pos = torch.repeat_interleave(torch.arange(2 * S, device=dev), child_sizes)
        keep = keep_child[child_of_pos]
        perm = perm[keep]
        seg_node = child_index[internal]
        sizes = child_sizes[internal]
        if log and level % 4 == 0:
            log("kdt level %d: %d segments" % (level, S))
        level += 1
    nodes = torch.stack([left, right, sdim, sval.view(torch.int32)], 1).contiguous()
    return nodes.cpu().numpy(), np.array([0], np.int32)


# ---------------------------------------------------------------------------------------------
# relative-neighbourhood graph
# ---------------------------------------------------------------------------------------------
def build_rng_graph(x, degree=32, cand=64, rng_factor=1.0, fill=True, row_chunk=None, log=None):
    """Exact `cand`-nearest neighbours by brute force, then the RNG prune rule of
    RelativeNeighborhoodGraph::RebuildNeighbors; with fill=True the rows are then topped up with the
    nearest rejected candidates (the reference prunes ~1000 candidates per node and ends with rows that
    are ~90% full; pruning only `cand` would leave rows a third full, which is not the workload the
    reference's searches see).  Returns int32 [N, degree] on CPU (-1 padded)."""
    dev = x.device
    N, dim = x.shape
    cand = min(cand, N - 1)
    xn = _sq_norms(x)
    graph = torch.full((N, degree), -1, dtype=torch.int32, device=dev)
    if row_chunk is None:
        row_chunk = max(64, min(8192, (1 << 31) // max(1, N)))  # <= 8 GiB of fp32 scores per chunk
    sub = max(64, min(row_chunk, (1 << 28) // max(1, cand * dim)))  # gather chunk: <= 1 GiB fp32
    t0 = time.time()
    for s in range(0, N, row_chunk):
        e = min(N, s + row_chunk)
        d = xn[s:e, None] - 2.0 * (x[s:e] @ x.t()) + xn[None, :]
        d[torch.arange(e - s, device=dev), torch.arange(s, e, device=dev)] = float("inf")
        dn, idx = torch.topk(d, cand, dim=1, largest=False, sorted=True)
        del d
        for ss in range(0, e - s, sub):
            ee = min(e - s, ss + sub)
            ci = idx[ss:ee]                       # [b, cand]
            cd = dn[ss:ee]
            cv = x[ci.reshape(-1)].view(ee - ss, cand, dim)
            cn = xn[ci]
            pd = cn[:, :, None] - 2.0 * torch.bmm(cv, cv.transpose(1, 2)) + cn[:, None, :]
            del cv
            acc = torch.zeros((ee - ss, cand), dtype=torch.bool, device=dev)
            count = torch.zeros(ee - ss, dtype=torch.int64, device=dev)
            out = torch.full((ee - ss, degree), -1, dtype=torch.int32, device=dev)
            rows = torch.arange(ee - ss, device=dev)
            for j in range(cand):
                bad = ((pd[:, :, j] * rng_factor < cd[:, j, None]) & acc).any(dim=1)
                ok = ~bad & (count < degree)
                acc[:, j] = ok
                r = rows[ok]
                out[r, count[ok]] = ci[ok, j].int()
                count += ok.long()
            if fill:
                for j in range(cand):
                    ok = ~acc[:, j] & (count < degree)
                    r = rows[ok]
                    out[r, count[ok]] = ci[ok, j].int()
                    count += ok.long()
            graph[s + ss:s + ee] = out
        if log and (s // row_chunk) % 16 == 0:
            log("graph rows %d/%d (%.1fs)" % (e, N, time.time() - t0))
    return graph.cpu().numpy()


# ---------------------------------------------------------------------------------------------
# kNN candidates at scale: random-projection partition trees with brute force inside the leaves -- the
# reference's own recipe for the initial graph (NeighborhoodGraph.h:126-360: TPTNumber trees, leaves <= TPTLeafSize,
# all-pairs inside a leaf, keep the best), level-synchronous on the GPU.  O(trees * N * leaf * dim) instead of O(N^2).
# ---------------------------------------------------------------------------------------------
def _tpt_leaves(x, leaf, gen):
    """One balanced random-projection tree. Returns perm (ids ordered by leaf) and the leaf size (last leaf padded)."""
    dev = x.device
    N, dim = x.shape
    perm = torch.randperm(N, generator=gen, device=dev)
    seg_len = N
    nseg = 1
    while seg_len > leaf:
        # every segment splits at the median of the projection on the direction between two of its random members
        offs = (torch.arange(nseg, device=dev) * seg_len)
        # (nseg * seg_len can exceed N by a few elements when N is not a power of two: clamp into the array)
        a = perm[(offs + torch.randint(0, seg_len, (nseg,), generator=gen, device=dev)).clamp_max(N - 1)]
        b = perm[(offs + torch.randint(0, seg_len, (nseg,), generator=gen, device=dev)).clamp_max(N - 1)]
        dirs = (x[a] - x[b])                                            # [nseg, dim]
        n_full = nseg * seg_len
        proj = torch.empty(N, device=dev)
        chunk = max(1, (1 << 26) // dim)
        segid = torch.arange(N, device=dev) // seg_len
        segid.clamp_(max=nseg - 1)
        for s in range(0, N, chunk):
            e = min(N, s + chunk)
            proj[s:e] = (x[perm[s:e]] * dirs[segid[s:e]]).sum(1)
        # sort inside segments: key = segment, then projection
        order = torch.argsort(proj, stable=True)
        order = order[torch.argsort(segid[order], stable=True)]
        perm = perm[order]
        seg_len = (seg_len + 1) // 2
        nseg *= 2
        # after the split the two halves of every old segment are [0, ceil(len/2)) and the rest; with N not a power of
        # two the tail segments are a little shorter, which only means slightly unbalanced leaves
        if nseg * seg_len < N:
            seg_len += 1
    return perm, seg_len


def build_knn_tpt(x, k=48, trees=8, leaf=1024, seed=0, log=None):
    """Approximate k nearest neighbours of every point. Returns (ids [N,k] int64, dists [N,k]) sorted by distance."""
    dev = x.device
    N, dim = x.shape
    g = torch.Generator(device=dev)
    g.manual_seed(seed + 77)
    xn = _sq_norms(x)
    best_d = torch.full((N, k), float("inf"), device=dev)
    best_i = torch.full((N, k), -1, dtype=torch.int64, device=dev)
    t0 = time.time()
    for t in range(trees):
        perm, m = _tpt_leaves(x, leaf, g)
        nleaf = (N + m - 1) // m
        pad = nleaf * m - N
        if pad:
            perm = torch.cat([perm, perm[:pad]])                      # pad the last leaf with repeats (masked below)
        lchunk = max(1, (1 << 28) // (m * m))
        kk = min(k, m - 1)
        for ls in range(0, nleaf, lchunk):
            le = min(nleaf, ls + lchunk)
            ids = perm[ls * m:le * m].view(le - ls, m)                  # [L, m]
            xl = x[ids.reshape(-1)].view(le - ls, m, dim)
            nl = xn[ids]
            d = nl[:, :, None] - 2.0 * torch.bmm(xl, xl.transpose(1, 2)) + nl[:, None, :]
            del xl
            d.diagonal(dim1=1, dim2=2).fill_(float("inf"))
            if pad and le == nleaf:                                   # padded duplicates of the last leaf: never candidates
                d[-1, :, m - pad:] = float("inf")
            cd, ci = torch.topk(d, kk, dim=2, largest=False)            # [L, m, kk]
            del d
            cand_ids = torch.gather(ids[:, None, :].expand(-1, m, -1), 2, ci)
            rows = ids.reshape(-1)
            cd = cd.reshape(-1, kk)
            cand_ids = cand_ids.reshape(-1, kk)
            if pad and le == nleaf:
                keep = torch.ones(rows.numel(), dtype=torch.bool, device=dev)
                keep[-pad:] = False
                rows, cd, cand_ids = rows[keep], cd[keep], cand_ids[keep]
            # merge with the running best of those rows: concatenate, drop duplicate ids, keep the k smallest
            md = torch.cat([best_d[rows], cd], 1)
            mi = torch.cat([best_i[rows], cand_ids], 1)
            o = torch.argsort(mi, dim=1, stable=True)
            mi_s = torch.gather(mi, 1, o)
            md_s = torch.gather(md, 1, o)
            dup = torch.zeros_like(mi_s, dtype=torch.bool)
            dup[:, 1:] = (mi_s[:, 1:] == mi_s[:, :-1]) & (mi_s[:, 1:] >= 0)
            md_s[dup] = float("inf")
            nd, no = torch.topk(md_s, k, dim=1, largest=False)
            best_d[rows] = nd
            best_i[rows] = torch.gather(mi_s, 1, no)
        if log:
            log("tpt tree %d/%d done (%.1fs)" % (t + 1, trees, time.time() - t0))
    best_i[torch.isinf(best_d)] = -1
    return best_i, best_d


def build_rng_graph_from_candidates(x, cand_i, cand_d, degree=32, rng_factor=1.0, fill=True, log=None):
    """RNG prune (RelativeNeighborhoodGraph::RebuildNeighbors) of given sorted candidate lists. int32 [N, degree] on CPU."""
    dev = x.device
    N, dim = x.shape
    cand = cand_i.shape[1]
    xn = _sq_norms(x)
    graph = torch.full((N, degree), -1, dtype=torch.int32, device=dev)
    sub = max(64, min(65536, (1 << 28) // max(1, cand * dim)))
    t0 = time.time()
    for s in range(0, N, sub):
        e = min(N, s + sub)
        ci = cand_i[s:e]
        cd = cand_d[s:e]
        valid = ci >= 0
        cis = ci.clamp_min(0)
        cv = x[cis.reshape(-1)].view(e - s, cand, dim)
        cn = xn[cis]
        pd = cn[:, :, None] - 2.0 * torch.bmm(cv, cv.transpose(1, 2)) + cn[:, None, :]
        del cv
        acc = torch.zeros((e - s, cand), dtype=torch.bool, device=dev)
        count = torch.zeros(e - s, dtype=torch.int64, device=dev)
        out = torch.full((e - s, degree), -1, dtype=torch.int32, device=dev)
        rows = torch.arange(e - s, device=dev)
        for j in range(cand):
            bad = ((pd[:, :, j] * rng_factor < cd[:, j, None]) & acc).any(dim=1)
            ok = ~bad & (count < degree) & valid[:, j]
            acc[:, j] = ok
            r = rows[ok]
            out[r, count[ok]] = ci[ok, j].int()
            count += ok.long()
        if fill:
            for j in range(cand):
                ok = ~acc[:, j] & (count < degree) & valid[:, j]
                r = rows[ok]
                out[r, count[ok]] = ci[ok, j].int()
                count += ok.long()
        graph[s:e] = out
        if log and (s // sub) % 32 == 0:
            log("rng rows %d/%d (%.1fs)" % (e, N, time.time() - t0))
    return graph.cpu().numpy()


# ---------------------------------------------------------------------------------------------
# folder writer (reference on-disk format) and exact ground truth
# ---------------------------------------------------------------------------------------------
INI_TEMPLATE = """[Index]
IndexAlgoType=BKT
ValueType=Float

TreeFilePath=tree.bin
GraphFilePath=graph.bin
VectorFilePath=vectors.bin
DeleteVectorFilePath=deletes.bin
EnableBfs=0
BKTNumber=1
BKTKmeansK={kmeans_k}
BKTLeafSize={leaf_size}
Samples=1000
BKTLambdaFactor=100.000000
TPTNumber=32
TPTLeafSize=2000
NumTopDimensionTpTreeSplit=5
NeighborhoodSize={degree}
GraphNeighborhoodScale=2.000000
GraphCEFScale=2.000000
RefineIterations=2
EnableRebuild=0
CEF=1000
AddCEF=500
MaxCheckForRefineGraph=8192
RNGFactor=1.000000
GPUGraphType=2
GPURefineSteps=0
GPURefineDepth=30
GPULeafSize=500
HeadNumGPUs=1
TPTBalanceFactor=2
NumberOfThreads={threads}
DistCalcMethod={metric}
DeletePercentageForRefine=0.400000
AddCountForRebuild=1000
MaxCheck=8192
ThresholdOfNumberOfContinuousNoBetterPropagation=3
NumberOfInitialDynamicPivots=50
NumberOfOtherDynamicPivots=4
HashTableExponent=2
DataBlockSize=1048576
DataCapacity=2147483647
MetaRecordSize=10
"""


# ---------------------------------------------------------------------------------------------
# synthetic PQ / OPQ quantizer (the reference cannot train OPQ natively, so the
# codebooks are plain per-subspace k-means and the rotation a random orthonormal matrix, written in the
# reference's SaveQuantizer format so the reference loads the same file)
# ---------------------------------------------------------------------------------------------
def train_quantizer_gpu(x, m, ks=256, opq=True, seed=0, iters=8, sample=200000):
    """x: float32 [N, dim] on the device. Returns (codebooks [m, ks, dsub], rotation [dim, dim] or None);
    the rotation is applied as x @ rotation (OPQQuantizer::m_VectorMatrixMultiply with the transposed matrix)."""
    dev = x.device
    N, dim = x.shape
    dsub = dim // m
    assert dsub * m == dim
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    rot = None
    if opq:
        a = torch.randn((dim, dim), generator=g, device=dev, dtype=torch.float32)
        rot, _ = torch.linalg.qr(a)
        rot = rot.contiguous()
    idx = torch.randperm(N, generator=g, device=dev)[:min(N, sample)]
    s = x[idx]
    if rot is not None:
        s = s @ rot
    s = s.view(-1, m, dsub).transpose(0, 1).contiguous()            # [m, S, dsub]
    S = s.shape[1]
    cb = s[:, torch.randperm(S, generator=g, device=dev)[:ks], :].clone()   # [m, ks, dsub]
    if cb.shape[1] < ks:
        cb = torch.cat([cb, cb[:, :ks - cb.shape[1]]], 1)
    for _ in range(iters):
        d = (s * s).sum(2)[:, :, None] - 2.0 * torch.bmm(s, cb.transpose(1, 2)) + (cb * cb).sum(2)[:, None, :]
        a = d.argmin(2)                                              # [m, S]
        onehot = torch.zeros((m, S, ks), device=dev, dtype=torch.float32).scatter_(2, a[:, :, None], 1.0)
        cnt = onehot.sum(1)                                          # [m, ks]
        newc = torch.bmm(onehot.transpose(1, 2), s)                  # [m, ks, dsub]
        nz = cnt > 0
        cb[nz] = newc[nz] / cnt[nz][:, None]
    return cb.contiguous(), rot


def encode_gpu(x, codebooks, rotation, chunk=262144):
    """Nearest codeword per sub-vector (plain fp32; the stored codes are just data for the index)."""
    m, ks, dsub = codebooks.shape
    chunk = max(1024, min(chunk, (1 << 29) // (m * ks)))  # <= 2 GiB for the [m, chunk, ks] distance block
    out = torch.empty((x.shape[0], m), dtype=torch.uint8, device=x.device)
    cn = (codebooks * codebooks).sum(2)
    for s in range(0, x.shape[0], chunk):
        xs = x[s:s + chunk]
        if rotation is not None:
            xs = xs @ rotation
        xs = xs.view(-1, m, dsub).transpose(0, 1)                    # [m, c, dsub]
        d = -2.0 * torch.bmm(xs, codebooks.transpose(1, 2)) + cn[:, None, :]
        out[s:s + chunk] = d.argmin(2).transpose(0, 1).to(torch.uint8)
    return out


def quantizer_blob(codebooks, rotation, rtype):
    """PQQuantizer::SaveQuantizer / OPQQuantizer::SaveQuantizer layout. rtype: 0 int8, 1 uint8, 2 int16, 3 float."""
    m, ks, dsub = codebooks.shape
    qtype = 2 if rotation is not None else 1
    b = np.array([qtype, rtype], np.uint8).tobytes() + np.array([m, ks, dsub], np.int32).tobytes()
    b += np.ascontiguousarray(codebooks.cpu().numpy(), np.float32).tobytes()
    if rotation is not None:
        b += np.ascontiguousarray(rotation.cpu().numpy(), np.float32).tobytes()
    return b


KDT_INI_HEAD = """[Index]
IndexAlgoType=KDT
ValueType=Float

TreeFilePath=tree.bin
GraphFilePath=graph.bin
VectorFilePath=vectors.bin
DeleteVectorFilePath=deletes.bin
KDTNumber=1
NumTopDimensionKDTSplit=5
Samples=100
IsOldVersion=false
"""


def save_index_folder(folder, vectors, graph, nodes, tree_starts, metric, kmeans_k=32, leaf_size=8,
                      quantizer=None, algo="BKT", value_type="Float"):
    """vectors: float32 numpy [N, dim] (already normalised for cosine), or uint8 PQ codes [N, M] together with
    `quantizer` = bytes of the quantizer file."""
    os.makedirs(folder, exist_ok=True)
    N, dim = vectors.shape
    quantized = quantizer is not None
    with open(os.path.join(folder, "vectors.bin"), "wb") as f:
        np.array([N, dim], np.int32).tofile(f)
        vdt = {"Float": np.float32, "Int8": np.int8, "UInt8": np.uint8}[value_type]
        np.ascontiguousarray(vectors, np.uint8 if quantized else vdt).tofile(f)
    if quantized:
        with open(os.path.join(folder, "quantizer.bin"), "wb") as f:
            f.write(quantizer)
    with open(os.path.join(folder, "graph.bin"), "wb") as f:
        np.array([N, graph.shape[1]], np.int32).tofile(f)
        np.ascontiguousarray(graph, np.int32).tofile(f)
    with open(os.path.join(folder, "tree.bin"), "wb") as f:
        np.array([tree_starts.shape[0]], np.int32).tofile(f)
        np.ascontiguousarray(tree_starts, np.int32).tofile(f)
        np.array([nodes.shape[0]], np.int32).tofile(f)
        np.ascontiguousarray(nodes, np.int32).tofile(f)
    with open(os.path.join(folder, "deletes.bin"), "wb") as f:
        np.array([0, N, 1], np.int32).tofile(f)  # Labelset: deleted count, then Dataset<int8>(N x 1)
        np.zeros(N, np.int8).tofile(f)
    with open(os.path.join(folder, "indexloader.ini"), "w") as f:
        f.write(ini_text(metric, graph.shape[1], kmeans_k, leaf_size, quantized, algo, value_type))


def ini_text(metric, degree, kmeans_k=32, leaf_size=8, quantized=False, algo="BKT", value_type="Float"):
    """The indexloader.ini the reference's LoadIndex reads (VectorIndex.cpp:197-222, :617-681 / :745-792)."""
    text = INI_TEMPLATE
    if algo == "KDT":
        # same [Index] body with the KD-tree parameters in place of the BKT ones (KDT/ParameterDefinitionList.h)
        body = INI_TEMPLATE.split("TPTNumber=32", 1)[1]
        text = KDT_INI_HEAD + "TPTNumber=32" + body.replace("NumTopDimensionTpTreeSplit", "NumTopDimensionTPTSplit")
    if quantized:
        text = "[Quantizer]\nQuantizerFilePath=quantizer.bin\n\n" + text.replace("ValueType=Float", "ValueType=UInt8")
    elif value_type != "Float":
        text = text.replace("ValueType=Float", "ValueType=" + value_type)
    return text.format(kmeans_k=kmeans_k, leaf_size=leaf_size, degree=degree, threads=os.cpu_count() or 1, metric=metric)


def exact_topk(x, q, k, metric, chunk=2048):
    """Exact ground truth ids [nq, k] (L2: squared distance; Cosine: 1 - dot on unit vectors)."""
    out = []
    xn = _sq_norms(x)
    chunk = max(32, min(chunk, (1 << 31) // max(1, x.shape[0])))  # <= 8 GiB of fp32 scores at a time
    for s in range(0, q.shape[0], chunk):
        qs = q[s:s + chunk]
        if metric == "L2":
            d = (qs ** 2).sum(1)[:, None] - 2.0 * (qs @ x.t()) + xn[None, :]
        else:
            d = 1.0 - qs @ x.t()
        out.append(torch.topk(d, k, dim=1, largest=False).indices)
    return torch.cat(out).cpu().numpy()


def build_index(x, metric="L2", degree=32, cand=64, kmeans_k=32, leaf_size=8, seed=0, log=None, algo="BKT",
                tpt_above=2500000, tpt_trees=8, balanced_above=12000000):
    """x: float32 tensor [N, dim] on the build device (unit rows for Cosine). Returns numpy arrays."""
    t = time.time()
    if algo == "KDT":
        nodes, starts = build_kdt(x, log=log)
    elif x.shape[0] > balanced_above:
        nodes, starts = build_bkt_balanced(x, kmeans_k=kmeans_k, leaf_size=leaf_size, seed=seed, log=log)
    else:
        nodes, starts = build_bkt(x, kmeans_k=kmeans_k, leaf_size=leaf_size, seed=seed, log=log)
    t_tree = time.time() - t
    t = time.time()
    if x.shape[0] > tpt_above:   # brute force is O(N^2): beyond a few million points use the partition-tree candidates
        kc = min(cand, 48) if x.shape[0] <= 50000000 else 32   # candidate lists are N x k x 12 bytes on the device
        ci, cdist = build_knn_tpt(x, k=kc, trees=tpt_trees, leaf=1024, seed=seed, log=log)
        graph = build_rng_graph_from_candidates(x, ci, cdist, degree=degree, log=log)
        del ci, cdist
    else:
        graph = build_rng_graph(x, degree=degree, cand=cand, log=log)
    t_graph = time.time() - t
    if log:
        log("index built: tree %.1fs, graph %.1fs" % (t_tree, t_graph))
    return nodes, starts, graph
