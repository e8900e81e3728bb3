"""Region-graph construction on the host (orchestration; stays off the GPU by design).

Mirrors /root/reference/src/treeStructure.jl (buildTree :4-21, getSplits :23-129, _buildSplit :131-210,
_buildSum :212-243, _buildGP :245-307, build :405-437), the node types of
src/DeepStructuredMixtures.jl:40-71 and getOverlap (src/fit.jl:12-39).  The partitions produced here are
INPUTS of libdsmgp (leaf_obs is consumed verbatim); the library never re-derives them.

Randomness: Julia's RNG stream cannot be reproduced without Julia; the draws are made in the same order
from the same distributions (Beta(2,2), rand(1:2), Categorical, Dirichlet) with a NumPy Generator.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _native as nat
from .kernels import KernelFunction

EPS = 1e-8   # DeepStructuredMixtures.jl:27


@dataclass
class DSMGPConfig:                      # DeepStructuredMixtures.jl:91-101
    meanFun: Optional["ConstMean"]
    kernels: Union[KernelFunction, List[KernelFunction]]
    observationNoise: float
    minData: int
    K: int          # number of splits per GPSplitNode
    V: int          # number of children under a GPSumNode
    depth: int
    bnoise: float
    sumRoot: bool


@dataclass
class ConstMean:                        # means.jl:7-18
    m: float


class Node:
    id: int = -1


@dataclass
class GPNode(Node):                     # DeepStructuredMixtures.jl:61-71 (dist lives in the device handle)
    obs: np.ndarray                     # ascending 1-based global rows (Julia convention)
    lb: np.ndarray
    ub: np.ndarray
    nobs: int
    kernelid: int                       # 1-based like the reference
    mean: float                         # ConstMean value
    kernel: KernelFunction = None
    logNoise: float = 0.0
    leaf_index: int = -1
    id: int = -1


@dataclass
class GPSplitNode(Node):                # DeepStructuredMixtures.jl:52-59
    lowerBound: np.ndarray
    upperBound: np.ndarray
    split: List[Tuple[int, float]]      # (d 1-based, s)
    children: List[Node] = field(default_factory=list)
    id: int = -1


@dataclass
class GPSumNode(Node):                  # DeepStructuredMixtures.jl:40-45
    children: List[Node] = field(default_factory=list)
    logweights: List[float] = field(default_factory=list)
    kernel_mixture: bool = False        # GPSumNode{T,GPNode} built by _buildGP for KernelFunction[...]
    id: int = -1

    def add(self, child: Node, logw: float):
        self.children.append(child)
        self.logweights.append(logw)


def getLeaves(node: Node) -> List[GPNode]:          # fit.jl:9-10
    if isinstance(node, GPNode):
        return [node]
    out: List[GPNode] = []
    for c in node.children:
        out.extend(getLeaves(c))
    return out


# ---------------------------------------------------------------------------------------------
def getSplits(X: np.ndarray, lowerBound, upperBound, minData: int, eps: float, K: int, d: int,
              rng: np.random.Generator, depth: int = 1) -> List[float]:
    """treeStructure.jl:23-129 (d 0-based here)."""
    K_ = depth ** 2
    s: List[float] = []
    col = X[:, d]
    l = max(lowerBound[d], col.min())
    u = min(upperBound[d], col.max())
    v = u - l
    sel = col[(col > l) & (col <= u)]                       # :40  (excludes the minimum point, App. B Q12)
    if sel.size > minData * 2:
        m = float(np.median(sel))                           # :49
        z1 = z2 = 0
        c = 0
        s_new = float(np.mean(col))
        while z1 == 0 or z2 == 0:
            a = rng.beta(2.0, 2.0) * v + l                  # :52
            s_new = float(eps * a + (1 - eps) * m)          # :54
            z1 = int(np.sum(sel <= s_new))
            z2 = int(np.sum(sel > s_new))
            c += 1
            if c > 100:                                     # :61-64
                return s
        zi = int(rng.integers(1, 3))                        # rand(1:2) :67

        def left():
            ub = np.array(upperBound, dtype=float); ub[d] = s_new
            return getSplits(X, lowerBound, ub, minData, eps, K, d, rng, depth + 1)

        def right():
            lb = np.array(upperBound, dtype=float); lb[d] = s_new     # copy(upperBound) sic, App. B Q11
            return getSplits(X, lb, upperBound, minData, eps, K, d, rng, depth + 1)

        if zi == 1:
            if z1 > minData and K_ < K:
                s.extend(left()); K_ += 1
            if z2 > minData and K_ < K:
                s.extend(right())
        else:
            if z2 > minData and K_ < K:
                s.extend(right()); K_ += 1
            if z1 > minData and K_ < K:
                s.extend(left())
        s.append(s_new)
    return s


def _buildGP(X, y, lb, ub, config: DSMGPConfig, observations: np.ndarray, rng) -> Node:
    """treeStructure.jl:245-307."""
    mean = float(np.mean(y)) if config.meanFun is None else float(config.meanFun.m)     # :271,292
    if isinstance(config.kernels, (list, tuple)):
        w = rng.dirichlet(np.ones(len(config.kernels)))                                 # :260
        node = GPSumNode(kernel_mixture=True)
        for v, kern in enumerate(config.kernels):
            node.add(GPNode(obs=observations.copy(), lb=np.array(lb), ub=np.array(ub), nobs=len(observations),
                            kernelid=v + 1, mean=mean, kernel=kern.copy(), logNoise=config.observationNoise),
                     math.log(w[v]))
        return node
    return GPNode(obs=observations.copy(), lb=np.array(lb), ub=np.array(ub), nobs=len(observations), kernelid=1,
                  mean=mean, kernel=config.kernels.copy(), logNoise=config.observationNoise)


def _buildSplit(X, y, lowerBound, upperBound, config: DSMGPConfig, depth: int, observations: np.ndarray,
                rng, d: int = 0) -> Node:
    """treeStructure.jl:131-210 (d 0-based)."""
    s = sorted(getSplits(X, lowerBound, upperBound, config.minData, config.bnoise, config.K, d, rng))
    split = [(d + 1, si) for si in s] + [(d + 1, float(upperBound[d]))]
    node = GPSplitNode(lowerBound=np.array(lowerBound), upperBound=np.array(upperBound), split=split)
    lb = np.array(lowerBound, dtype=float)
    ub = np.array(upperBound, dtype=float)
    col = X[:, d]
    if s:
        for (_, si) in split:
            lb_ = lb.copy(); ub_ = ub.copy(); ub_[d] = si
            idx = np.nonzero((col > lb_[d]) & (col <= ub_[d]))[0]
            if depth < config.depth and idx.size > config.minData:
                if config.sumRoot:
                    child = _buildSum(X[idx], y[idx], lb_, ub_, config, depth, observations[idx], rng)
                else:
                    child = _buildSplit(X[idx], y[idx], lb_, ub_, config, depth, observations[idx], rng)
            else:
                child = _buildGP(X[idx], y[idx], lb_, ub_, config, observations[idx], rng)
            node.children.append(child)
            lb[d] = si
        return node
    idx = np.nonzero((col > lowerBound[d]) & (col <= upperBound[d]))[0]
    return _buildGP(X[idx], y[idx], np.array(lowerBound), np.array(upperBound), config, observations[idx], rng)


def _buildSum(X, y, lowerBound, upperBound, config: DSMGPConfig, depth: int, observations: np.ndarray, rng) -> Node:
    """treeStructure.jl:212-243."""
    V = config.V
    node = GPSumNode()
    phi = X.max(axis=0) - X.min(axis=0)
    phi = phi / phi.sum()
    for _ in range(V):
        d = int(rng.choice(len(phi), p=phi))                # rand(Categorical(phi)) :236
        node.add(_buildSplit(X, y, lowerBound, upperBound, config, depth + 1, observations, rng, d=d), -math.log(V))
    return node


# ---------------------------------------------------------------------------------------------
# The same construction with the data passes on the device (`dsmgp_part_*`): X stays in HBM, every node of
# the recursion is an index list there.  The recursion and EVERY random draw are the code above (same order, same generator),
# so the region graph is bit-identical to the host builder's; what changes is that getSplits works on the node's SORTED column
# (min / max / median / counts become binary searches), `findall` becomes a stable K-way partition kernel and X[idx,:] is never
# copied.
class DevicePartition:
    def __init__(self, X: np.ndarray):
        import ctypes as C
        self._C = C
        self._lib = nat.lib()
        self._p = C.c_void_p()
        self.X = nat.colmajor(X)
        self.N, self.D = self.X.shape
        nat.check(self._lib.dsmgp_part_create(nat.p_d(self.X), self.N, self.D, C.byref(self._p)))

    def close(self):
        if self._p and self._p.value:
            self._lib.dsmgp_part_destroy(self._p)
            self._p = self._C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self, node: int) -> int:
        return int(self._lib.dsmgp_part_size(self._p, node))

    def range(self, node: int):
        mn = np.zeros(self.D); mx = np.zeros(self.D)
        nat.check(self._lib.dsmgp_part_range(self._p, node, nat.p_d(mn), nat.p_d(mx)))
        return mn, mx

    def sorted_column(self, node: int, d: int) -> np.ndarray:
        out = np.zeros(self.size(node))
        nat.check(self._lib.dsmgp_part_sorted_column(self._p, node, d, nat.p_d(out)))
        return out

    def split(self, node: int, d: int, lower, upper):
        lo = nat.f64(lower); up = nat.f64(upper)
        K = lo.size
        ch = np.zeros(K, dtype=np.int64); sz = np.zeros(K, dtype=np.int64)
        nat.check(self._lib.dsmgp_part_split(self._p, node, d, nat.p_d(lo), nat.p_d(up), K, nat.p_i64(ch), nat.p_i64(sz)))
        return ch, sz

    def rows(self, node: int) -> np.ndarray:
        out = np.zeros(self.size(node), dtype=np.int64)
        nat.check(self._lib.dsmgp_part_rows(self._p, node, nat.p_i64(out)))
        return out


def _getSplits_sorted(col: np.ndarray, lowerBound, upperBound, minData: int, eps: float, K: int, d: int,
                      rng: np.random.Generator, depth: int = 1) -> List[float]:
    """getSplits (treeStructure.jl:23-129) on the node's SORTED column: identical draws and results to `getSplits`."""
    K_ = depth ** 2
    s: List[float] = []
    l = max(lowerBound[d], col[0])
    u = min(upperBound[d], col[-1])
    v = u - l
    i0 = int(np.searchsorted(col, l, side="right"))          # (col > l)
    i1 = int(np.searchsorted(col, u, side="right"))          # (col <= u)
    sel = col[i0:i1]
    if sel.size > minData * 2:
        n = sel.size
        m = float(sel[n // 2]) if n % 2 else float((sel[n // 2 - 1] + sel[n // 2]) / 2.0)      # median(sel) :49
        z1 = z2 = 0
        c = 0
        s_new = 0.0
        while z1 == 0 or z2 == 0:
            a = rng.beta(2.0, 2.0) * v + l
            s_new = float(eps * a + (1 - eps) * m)
            z1 = int(np.searchsorted(sel, s_new, side="right"))
            z2 = n - z1
            c += 1
            if c > 100:
                return s
        zi = int(rng.integers(1, 3))

        def left():
            ub = np.array(upperBound, dtype=float); ub[d] = s_new
            return _getSplits_sorted(col, lowerBound, ub, minData, eps, K, d, rng, depth + 1)

        def right():
            lb = np.array(upperBound, dtype=float); lb[d] = s_new
            return _getSplits_sorted(col, lb, upperBound, minData, eps, K, d, rng, depth + 1)

        if zi == 1:
            if z1 > minData and K_ < K:
                s.extend(left()); K_ += 1
            if z2 > minData and K_ < K:
                s.extend(right())
        else:
            if z2 > minData and K_ < K:
                s.extend(right()); K_ += 1
            if z1 > minData and K_ < K:
                s.extend(left())
        s.append(s_new)
    return s


def _buildGP_dev(part: DevicePartition, y, lb, ub, config: DSMGPConfig, node: int, rng) -> Node:
    obs = part.rows(node)
    return _buildGP(None, y[obs - 1], lb, ub, config, obs, rng)


def _buildSplit_dev(part: DevicePartition, y, lowerBound, upperBound, config: DSMGPConfig, depth: int, node: int, rng, d: int = 0) -> Node:
    col = part.sorted_column(node, d)
    s = sorted(_getSplits_sorted(col, lowerBound, upperBound, config.minData, config.bnoise, config.K, d, rng))
    split = [(d + 1, si) for si in s] + [(d + 1, float(upperBound[d]))]
    out = GPSplitNode(lowerBound=np.array(lowerBound), upperBound=np.array(upperBound), split=split)
    lb = np.array(lowerBound, dtype=float)
    ub = np.array(upperBound, dtype=float)
    if s:
        lows, ups, bounds = [], [], []
        for (_, si) in split:
            lb_ = lb.copy(); ub_ = ub.copy(); ub_[d] = si
            lows.append(lb_[d]); ups.append(ub_[d]); bounds.append((lb_, ub_))
            lb[d] = si
        children, sizes = part.split(node, d, lows, ups)
        for (lb_, ub_), ch, n in zip(bounds, children, sizes):
            if depth < config.depth and n > config.minData:
                if config.sumRoot:
                    child = _buildSum_dev(part, y, lb_, ub_, config, depth, int(ch), rng)
                else:
                    child = _buildSplit_dev(part, y, lb_, ub_, config, depth, int(ch), rng)
            else:
                child = _buildGP_dev(part, y, lb_, ub_, config, int(ch), rng)
            out.children.append(child)
        return out
    children, _ = part.split(node, d, [lowerBound[d]], [upperBound[d]])
    return _buildGP_dev(part, y, np.array(lowerBound), np.array(upperBound), config, int(children[0]), rng)


def _buildSum_dev(part: DevicePartition, y, lowerBound, upperBound, config: DSMGPConfig, depth: int, node: int, rng) -> Node:
    V = config.V
    out = GPSumNode()
    mn, mx = part.range(node)
    phi = mx - mn
    phi = phi / phi.sum()
    for _ in range(V):
        d = int(rng.choice(len(phi), p=phi))
        out.add(_buildSplit_dev(part, y, lowerBound, upperBound, config, depth + 1, node, rng, d=d), -math.log(V))
    return out


def buildTree_device(X: np.ndarray, y: np.ndarray, config: DSMGPConfig, rng) -> Node:
    """buildTree (treeStructure.jl:4-21) with the data passes on the device; same graph as `buildTree` for the same generator."""
    N, D = X.shape
    assert N == len(y)
    assert np.all(np.isfinite(X))
    y = np.asarray(y, dtype=np.float64)
    part = DevicePartition(X)
    try:
        lb = np.full(D, -np.inf)
        ub = np.full(D, np.inf)
        if config.sumRoot:
            return _buildSum_dev(part, y, lb, ub, config, 0, 0, rng)
        return _buildSplit_dev(part, y, lb, ub, config, 0, 0, rng)
    finally:
        part.close()


def buildTree(X: np.ndarray, y: np.ndarray, config: DSMGPConfig, rng) -> Node:
    """treeStructure.jl:4-21."""
    N, D = X.shape
    assert N == len(y)
    assert np.all(np.isfinite(X))
    lb = np.full(D, -np.inf)
    ub = np.full(D, np.inf)
    obs = np.arange(1, N + 1, dtype=np.int64)
    if config.sumRoot:
        return _buildSum(X, y, lb, ub, config, 0, obs, rng)
    return _buildSplit(X, y, lb, ub, config, 0, obs, rng)


# ---------------------------------------------------------------------------------------------
def number_nodes(root: Node) -> List[Node]:
    """Post-order numbering: children precede parents, root last (the order dsmgp_tree requires)."""
    order: List[Node] = []

    def rec(n: Node):
        if not isinstance(n, GPNode):
            for c in n.children:
                rec(c)
        n.id = len(order)
        order.append(n)

    rec(root)
    for i, lf in enumerate(getLeaves(root)):
        lf.leaf_index = i
    return order


def flatten(root: Node) -> Tuple[nat.FlatTree, List[GPNode]]:
    order = number_nodes(root)
    nn = len(order)
    node_type = np.zeros(nn, dtype=np.int32)
    child_ptr = np.zeros(nn + 1, dtype=np.int64)
    child_idx: List[int] = []
    leaf_of_node = np.full(nn, -1, dtype=np.int64)
    split_dim = np.full(nn, -1, dtype=np.int32)
    split_ptr = np.zeros(nn + 1, dtype=np.int64)
    split_val: List[float] = []
    for i, n in enumerate(order):
        if isinstance(n, GPNode):
            node_type[i] = nat.NODE_LEAF
            leaf_of_node[i] = n.leaf_index
        else:
            if isinstance(n, GPSplitNode):
                node_type[i] = nat.NODE_SPLIT
                split_dim[i] = n.split[0][0] - 1
                split_val.extend(s for (_, s) in n.split)
            else:
                node_type[i] = nat.NODE_KSUM if n.kernel_mixture else nat.NODE_SUM
            child_idx.extend(c.id for c in n.children)
        child_ptr[i + 1] = len(child_idx)
        split_ptr[i + 1] = len(split_val)
    ft = nat.FlatTree(node_type, child_ptr, np.array(child_idx, dtype=np.int64), leaf_of_node, split_dim, split_ptr,
                      np.array(split_val, dtype=np.float64), root.id)
    return ft, getLeaves(root)


def getOverlap(root: Node, N: int) -> np.ndarray:
    """fit.jl:12-39.  D[n,m] = 1 - |obs_n \\ obs_m| / |obs_n| for leaves below different children of a common
    sum node (count multiplied by kernelid equality, so D = 1 across kernels).  Host-side BitArray work, as in
    the reference (treeStructure.jl:428-431)."""
    leaves = getLeaves(root)
    L = len(leaves)
    D = np.zeros((L, L))
    bits = np.zeros((L, N), dtype=bool)
    for i, lf in enumerate(leaves):
        bits[i, lf.obs - 1] = True
        lf.leaf_index = i
    cnt = bits.sum(axis=1)

    def rec(node: Node) -> List[int]:
        if isinstance(node, GPNode):
            return [node.leaf_index]
        r = [rec(c) for c in node.children]
        if isinstance(node, GPSumNode):
            for i in range(len(r)):
                for j in range(i + 1, len(r)):
                    for n in r[i]:
                        for m in r[j]:
                            same = 1 if leaves[n].kernelid == leaves[m].kernelid else 0
                            inter = int(np.count_nonzero(bits[n] & bits[m]))
                            D[n, m] = 1.0 - ((cnt[n] - inter) * same) / cnt[n]
                            D[m, n] = 1.0 - ((cnt[m] - inter) * same) / cnt[m]
        return [x for sub in r for x in sub]

    rec(root)
    return D
