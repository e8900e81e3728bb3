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
