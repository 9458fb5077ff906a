"""Skeleton extraction from the affinity matrix (drop-in for the reference's utils/dyna_utils.py).

One-off host logic (K = 24 nodes, result cached by HSVRNNBVH), so it stays on the CPU; unlike the reference it
needs no networkx: shortest paths come from a small Dijkstra.  Path lengths are accumulated from the source
outwards, which keeps the float values (hop counts plus 1e-5 tie-break bumps) identical to the reference's.
"""
from __future__ import annotations

import heapq
from collections import namedtuple

import numpy as np
import torch

Priority = namedtuple("Priority", ["values", "indices"])


def _shortest_paths(mask: np.ndarray, weights: np.ndarray, big: float) -> np.ndarray:
    """mask[a, b] != 0 where an (undirected) edge exists, weights[a, b] its length (may be 0 for an edge that
    was added after the weights were frozen) -> (K, K) path lengths, `big` if unreachable."""
    K = mask.shape[0]
    nbrs = [np.nonzero(mask[a])[0].tolist() for a in range(K)]
    out = np.full((K, K), big, dtype=np.float64)
    for src in range(K):
        best = {src: 0.0}
        heap = [(0.0, src)]
        closed = [False] * K
        while heap:
            d, u = heapq.heappop(heap)
            if closed[u]:
                continue
            closed[u] = True
            out[src, u] = d
            for v in nbrs[u]:
                cand = d + float(weights[u, v])
                if v not in best or cand < best[v]:
                    best[v] = cand
                    heapq.heappush(heap, (cand, v))
    return out


def _components(adj: np.ndarray) -> int:
    K = adj.shape[0]
    seen = [False] * K
    count = 0
    for s in range(K):
        if seen[s]:
            continue
        count += 1
        stack = [s]
        seen[s] = True
        while stack:
            u = stack.pop()
            for v in np.nonzero(adj[u])[0]:
                if not seen[v]:
                    seen[v] = True
                    stack.append(int(v))
    return count


def process_affinity_glob(affinity, BIG_NUM=1e4):
    """affinity (nneighbor, K, K, 1) -> (A (K, K) tree adjacency, priority (values, indices): nodes by ascending
    tree distance from the root, parents (K,) int64).  Follows reference dyna_utils.py:6-171 step by step."""
    n, K = affinity.shape[0], affinity.shape[1]
    dev = affinity.device
    with torch.no_grad():
        infl_t = affinity.max(dim=0).values.squeeze(-1)
        infl = infl_t.detach().cpu().numpy()
        picks = infl_t.topk(n, dim=-1).indices.cpu().numpy()
        adj = np.zeros((K, K), dtype=np.float32)
        for k in range(K):
            adj[k, picks[k]] = 1
        adj = np.maximum(adj, adj.T)

        dist = _shortest_paths(adj, adj, BIG_NUM)
        if _components(adj) > 1:
            # join the first unreachable component (lowest total-distance rank) to the provisional root
            total = dist.sum(axis=-1)
            root = int(total.argmin())
            rank = np.empty(K)
            rank[total.argsort()] = np.arange(K)
            unreachable = np.where(dist[root] == BIG_NUM)[0]
            pick = unreachable[0]
            for c in unreachable[1:]:
                if rank[pick] > rank[c]:
                    pick = c
            adj[root, pick] = adj[pick, root] = 1
            dist = _shortest_paths(adj, adj, BIG_NUM)

        # nodes with identical total distance: bump the edge to the less influential one by 1e-5
        total = dist.sum(axis=-1)
        wadj = adj.copy()
        for k in range(K - 1):
            for q in range(k + 1, K):
                if total[k] != total[q]:
                    continue
                q_nb = set(np.nonzero(adj[q])[0].tolist())
                for m in np.nonzero(adj[k])[0]:
                    if int(m) in q_nb:
                        loser = q if infl[m, k] > infl[m, q] else k
                        wadj[m, loser] += 1e-5
                        wadj[loser, m] += 1e-5
        dist = torch.from_numpy(_shortest_paths(adj, wadj, BIG_NUM))

        root = int(dist.sum(dim=-1).topk(K, largest=False).indices[0])
        depth = dist[root]
        first = int(depth.topk(K, largest=False).indices[0])
        parents = []
        for k in range(K):
            if k == root:
                parents.append(k)
                continue
            nb = np.nonzero(adj[k])[0]
            choice, gap = None, -1e3
            for m in nb:
                delta = depth[m] - depth[k]
                if delta < 0 and delta > gap:
                    choice, gap = int(m), delta
                elif delta < 0 and delta == gap:
                    if infl[k, m] > infl[k, choice]:
                        choice, gap = int(m), delta
                elif delta == 0:
                    shared, shared_depth = None, 1e4
                    for mm in np.nonzero(adj[m])[0]:
                        if mm in nb and depth[mm] < depth[m] and shared_depth > depth[mm]:
                            shared, shared_depth = int(mm), depth[mm]
                    if shared is not None and infl[shared, m] > infl[shared, k]:
                        choice, gap = int(m), delta
            if choice is None:
                choice = first
                adj[k, choice] = adj[choice, k] = 1
            parents.append(choice)

        tree = np.zeros((K, K), dtype=np.float64)
        for k, pa in enumerate(parents):
            if k != pa:
                tree[k, pa] = tree[pa, k] = 1
        tdist = torch.from_numpy(_shortest_paths(tree, wadj, BIG_NUM))
        pr = tdist[root].topk(K, dim=-1, largest=False)
        priority = Priority(values=pr.values.to(dev), indices=pr.indices.to(dev))
        return torch.from_numpy(tree).to(dev), priority, torch.tensor(parents, dtype=torch.int64, device=dev)
