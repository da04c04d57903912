"""Seeded synthetic power-law graphs of the BASELINE.json shapes.

Datasets are not available offline, so every workload is a Chung-Lu style
undirected simple graph, symmetrised the way the reference symmetrises OGB edge
lists (reference preprocess.py:65-69) and stored as the *structure* of the
row-normalised adjacency (reference utils.py:56-64, main.py:267-270).  Only
``indptr``/``indices`` are kept: the hot path never reads ``lap_matrix.data``
(the adjacency values are recomputed from full-graph degrees in
``create_coo_tensor``, reference cuda_spmm.cu:800).

Generator parameters are fixed per shape and reported with every number
(SURVEY.md section 8(d)): ``alpha`` is the rank exponent of the expected-degree
sequence ``w_i ~ (i + i0)^-alpha`` and ``max_degree`` caps the hub.  The edge
count after duplicate/self-loop removal is topped up until it hits the target.
"""
from __future__ import annotations

import dataclasses
import hashlib
import os

import numpy as np


@dataclasses.dataclass(frozen=True)
class GraphShape:
    name: str
    num_nodes: int
    num_undirected_edges: int   # directed nnz of the symmetrised adjacency = 2x this
    feat_dim: int
    num_classes: int
    max_degree: int             # cap on the expected degree of the largest hub
    alpha: float = 0.6          # rank exponent of the expected-degree sequence
    self_loops: bool = False    # GCN uses row_normalize(adj + I) (reference main.py:269-270)


# Shapes named by BASELINE.json `configs`.
SHAPES = {
    # Cora: 2,708 nodes / 10,556 directed edges / 1,433-d / 7 classes, GCN => +I
    "cora": GraphShape("cora", 2708, 5278, 1433, 7, max_degree=168, self_loops=True),
    # Reddit: 232,965 nodes / 114.6 M directed edges / 602-d / 41 classes, GraphSAGE => no +I
    "reddit": GraphShape("reddit", 232965, 57307946, 602, 41, max_degree=21657),
    # ogbn-products: 2,449,029 nodes / 61.86 M undirected edges / 100-d / 47 classes, GCN => +I
    "products": GraphShape("products", 2449029, 61859140, 100, 47, max_degree=17481, self_loops=True),
    # ogbn-papers100M scaled by 1/16 in nodes and edges (111 M nodes / 1.6 B undirected edges do not fit the
    # time budget of a bench run; degree statistics are kept: mean directed degree ~29), 128-d, GCN => +I
    "papers16": GraphShape("papers16", 6941000, 100000000, 128, 172, max_degree=15000, self_loops=True),
    # small shapes used by the CPU test-suite and smoke()
    "tiny": GraphShape("tiny", 600, 3000, 37, 5, max_degree=90, self_loops=True),
    "small": GraphShape("small", 20000, 400000, 100, 16, max_degree=2500),
}


@dataclasses.dataclass
class Graph:
    shape: GraphShape
    indptr: np.ndarray        # int64 [N+1]
    indices: np.ndarray       # int32 [nnz]   sorted within each row, no duplicates
    train_nodes: np.ndarray   # int64, random 66 % of nodes (seed 2)
    valid_nodes: np.ndarray
    test_nodes: np.ndarray

    @property
    def num_nodes(self) -> int:
        return self.shape.num_nodes

    @property
    def nnz(self) -> int:
        return int(self.indptr[-1])

    def degrees(self) -> np.ndarray:
        return np.diff(self.indptr)

    def to_scipy(self, dtype=np.float32):
        """Row-normalised adjacency as scipy CSR (what reference main.py calls lap_matrix)."""
        import scipy.sparse as sp
        deg = self.degrees()
        data = np.repeat((1.0 / np.maximum(deg, 1)).astype(dtype), deg)
        return sp.csr_matrix((data, self.indices, self.indptr), shape=(self.num_nodes, self.num_nodes))


def _expected_degree_weights(shape: GraphShape) -> np.ndarray:
    n = shape.num_nodes
    mean_deg = 2.0 * shape.num_undirected_edges / n
    ratio = max(shape.max_degree / mean_deg, 1.0)
    ranks = np.arange(n, dtype=np.float64)
    # find i0 with w_0 / mean(w) == ratio by bisection on log scale (monotone decreasing in i0)
    lo, hi = 1e-3, float(n) * 1e3
    for _ in range(80):
        mid = np.sqrt(lo * hi)
        w = (ranks + mid) ** (-shape.alpha)
        if w[0] / w.mean() > ratio:
            lo = mid
        else:
            hi = mid
    w = (ranks + np.sqrt(lo * hi)) ** (-shape.alpha)
    return w / w.sum()


def _sorted_unique(keys: np.ndarray) -> np.ndarray:
    """np.unique for int64 keys via sort + neighbour compare (np.unique's hash path is ~50x slower here)."""
    if keys.size == 0:
        return keys
    s = np.sort(keys)
    m = np.empty(s.size, dtype=bool)
    m[0] = True
    np.not_equal(s[1:], s[:-1], out=m[1:])
    return s[m]


def _sample_pairs(rng, table, count, n):
    """`count` endpoint pairs drawn from the expected-degree distribution through a guide table
    (inverse CDF quantised to len(table) buckets); self loops dropped; returned as lo*n+hi keys."""
    t = table.size
    u = table[(rng.random(count) * t).astype(np.int64)].astype(np.int64)
    v = table[(rng.random(count) * t).astype(np.int64)].astype(np.int64)
    keep = u != v
    u, v = u[keep], v[keep]
    lo = np.minimum(u, v)
    hi = np.maximum(u, v)
    return lo * n + hi


def generate(shape: GraphShape | str, seed: int = 0) -> Graph:
    """Chung-Lu undirected simple graph with exactly ``num_undirected_edges`` edges."""
    if isinstance(shape, str):
        shape = SHAPES[shape]
    n, target = shape.num_nodes, shape.num_undirected_edges
    rng = np.random.Generator(np.random.PCG64(seed))
    # node ids are shuffled so that hubs are not the low ids
    perm = rng.permutation(n)
    cdf = np.cumsum(_expected_degree_weights(shape))
    cdf /= cdf[-1]
    tsize = 1 << int(np.clip(np.ceil(np.log2(n)) + 4, 16, 26))
    table = np.minimum(np.searchsorted(cdf, (np.arange(tsize) + 0.5) / tsize, side="right"), n - 1).astype(np.int32)

    keys = np.empty(0, dtype=np.int64)
    need = target
    while keys.size < target:
        batch = _sample_pairs(rng, table, int(need * 1.15) + 1024, n)
        keys = _sorted_unique(np.concatenate([keys, batch]))
        need = max(target - keys.size, 0) * 2 + 1024
    if keys.size > target:
        drop = rng.choice(keys.size, keys.size - target, replace=False)
        mask = np.ones(keys.size, dtype=bool)
        mask[drop] = False
        keys = keys[mask]
    lo = perm[keys // n]
    hi = perm[keys % n]
    del keys

    parts = [lo * n + hi, hi * n + lo]
    del lo, hi
    if shape.self_loops:
        eye = np.arange(n, dtype=np.int64)
        parts.append(eye * n + eye)
    directed = np.sort(np.concatenate(parts))
    del parts
    rows = directed // n
    indices = (directed - rows * n).astype(np.int32 if n < 2**31 else np.int64)
    del directed
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=indptr[1:])
    del rows

    split = np.random.Generator(np.random.PCG64(2)).permutation(n)
    n_train = int(0.66 * n)
    n_val = int(0.10 * n)
    return Graph(shape, indptr, indices,
                 np.sort(split[:n_train]), np.sort(split[n_train:n_train + n_val]),
                 np.sort(split[n_train + n_val:]))


def features(shape: GraphShape | str, seed: int = 1, rows: np.ndarray | None = None,
             dtype=np.float32) -> np.ndarray:
    """N(0,1) fp32 features (standardised like reference preprocess.py:29-31).

    With ``rows`` given only those rows are generated; row r is a pure function of
    (seed, r), so shards generated on different ranks agree bit for bit."""
    if isinstance(shape, str):
        shape = SHAPES[shape]
    if rows is None:
        rows = np.arange(shape.num_nodes, dtype=np.int64)
    rows = np.asarray(rows, dtype=np.int64)
    # counter-based: hash (seed,row,col) -> two uint32 -> Box-Muller; pure numpy, vectorised
    f = shape.feat_dim
    out = np.empty((rows.size, f), dtype=dtype)
    cols = np.arange(f, dtype=np.uint64)
    step = max(1, (1 << 22) // max(f, 1))
    for s in range(0, rows.size, step):
        r = rows[s:s + step].astype(np.uint64)[:, None]
        x = (r * np.uint64(0x9E3779B97F4A7C15) + cols * np.uint64(0xC2B2AE3D27D4EB4F)
             + np.uint64((int(seed) * 0x165667B19E3779F9) & 0xFFFFFFFFFFFFFFFF))
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xFF51AFD7ED558CCD)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xC4CEB9FE1A85EC53)
        x ^= x >> np.uint64(33)
        u1 = ((x >> np.uint64(40)).astype(np.float64) + 0.5) / float(1 << 24)
        u2 = ((x & np.uint64(0xFFFFFF)).astype(np.float64) + 0.5) / float(1 << 24)
        out[s:s + step] = (np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)).astype(dtype)
    return out


def labels(shape: GraphShape | str, seed: int = 3) -> np.ndarray:
    """Uniform class ids int64 [N]."""
    if isinstance(shape, str):
        shape = SHAPES[shape]
    return np.random.Generator(np.random.PCG64(seed)).integers(0, shape.num_classes, shape.num_nodes)


def cache_path(shape: GraphShape, seed: int, root: str | None = None) -> str:
    root = root or os.environ.get("GNN_B200_CACHE", "/tmp/gnn_b200_cache")
    tag = hashlib.sha1(repr((dataclasses.astuple(shape), seed)).encode()).hexdigest()[:12]
    # graphs above ~64 MB (Reddit-shaped: 460 MB) go to a sub-directory that .gpurunignore lists, so a cache written
    # by a local run never rides along in the repository snapshot; ranks of one box still share it
    if shape.num_undirected_edges * 8 > (64 << 20):
        root = os.path.join(root, "big")
    return os.path.join(root, f"graph_{shape.name}_{tag}.npz")


def generate_cached(shape: GraphShape | str, seed: int = 0, root: str | None = None) -> Graph:
    if isinstance(shape, str):
        shape = SHAPES[shape]
    path = cache_path(shape, seed, root)
    if os.path.exists(path):
        z = np.load(path)
        return Graph(shape, z["indptr"], z["indices"], z["train"], z["valid"], z["test"])
    g = generate(shape, seed)
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.savez(path + ".tmp.npz", indptr=g.indptr, indices=g.indices, train=g.train_nodes,
                 valid=g.valid_nodes, test=g.test_nodes)
        os.replace(path + ".tmp.npz", path)
    except OSError:
        pass
    return g
