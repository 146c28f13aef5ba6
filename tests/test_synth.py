"""Synthetic geometries used by bench.py and the idea behind the fused routed objective, on the CPU.

* synth.d8_convergent: one drainage tree whose outlet block polygon sees every subbasin (checked with the dense oracle).
* For a FIXED gauge polygon the routed discharge is max over the polygon's maximal candidate nodes of the sum of QS over
  the node's upstream subbasins (QS >= 0): restated here on the CPU twin of the routing operator and held to the dense
  oracle's AreaD8 + per-polygon maximum (R/Routing_GR2MSemiDistr.R:100-123) for random QS and every polygon."""
import numpy as np
import pytest

import oracle
from oracle import router_twin
from gr2msemidistr_b200 import synth


def gauge_owner(op, gauge, nsub):
    """CPU restatement of wfa_router_gauge_upstream on the twin's forest (children lists, node_src, candidates)."""
    nn = len(op["nodes"])
    parent = np.full(nn, -1)
    for v, ch in enumerate(op["children"]):
        for c in ch:
            parent[c] = v
    cand = sorted(set(op["cand"][gauge]))
    iscand = np.zeros(nn, bool)
    iscand[cand] = True
    owner, k = np.full(nsub, -1), 0
    for c in cand:
        v, maximal = parent[c], True
        while v >= 0:
            if iscand[v]:
                maximal = False
                break
            v = parent[v]
        if not maximal:
            continue
        stack = [c]
        while stack:
            v = stack.pop()
            if op["node_src"][v] >= 0:
                owner[op["node_src"][v]] = k
            stack.extend(op["children"][v])
        k += 1
    return k, owner


def test_convergent_tree_outlet_sees_every_subbasin():
    n, nb = 240, 12
    d8 = synth.d8_convergent(n).numpy()
    src, off, cells = synth.block_polygons(n, n, nb, nb, margin=max(1, n // nb // 4))
    QR = oracle.route(d8, np.ones((len(src), 1)), src, off, cells)
    assert QR[-1, 0] == nb * nb                        # the outlet block accumulates the whole basin
    assert np.isfinite(QR).all() and QR.min() >= 1     # no polygon is edge-contaminated, every one sees its own source
    op = router_twin.build(d8, src, off, cells)
    k, owner = gauge_owner(op, len(src) - 1, len(src))
    assert k == 1 and (owner == 0).all()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_routing_at_a_fixed_gauge_is_a_maximum_of_masked_sums(seed):
    rng = np.random.default_rng(seed)
    nrow, ncol = 72, 96
    d8 = synth.d8_grid(nrow, ncol, seed=seed).numpy()
    src, off, cells = synth.block_polygons(nrow, ncol, 6, 8, margin=1)
    nsub = len(src)
    QS = np.round(rng.uniform(0, 300, (nsub, 5)), 1)
    QS[rng.random(QS.shape) < 0.2] = 0.0
    QR = oracle.route(d8, QS, src, off, cells)
    op = router_twin.build(d8, src, off, cells)
    seen = 0
    for g in range(nsub):
        k, owner = gauge_owner(op, g, nsub)
        if k == 0:
            assert not np.any(QR[g] > 0)               # no source-carrying path crosses the polygon's interior
            continue
        want = np.max([QS[owner == j].sum(axis=0) for j in range(k)], axis=0)
        assert np.allclose(QR[g], want, rtol=1e-13, atol=1e-10)
        seen += k > 1
    assert seen > 0                                    # square blocks on a tilted plane: several parallel streams
