"""CPU restatement of the month-invariant routing operator (csrc/router.cu) -- TEST INFRASTRUCTURE ONLY.

Same algebra as the CUDA operator, written the slow and obvious way, so that the collapse itself
(R/Routing_GR2MSemiDistr.R:77-81,100-123: one weight per subbasin, AreaD8, per-polygon maximum) can be checked
on a CPU box against the dense oracle (oracle.route), with no GPU involved:

  node    = a source cell, or a cell with >= 2 upstream neighbours that have a source above them
  A[node] = QS[source at node] (else 0) + sum over child nodes in AreaD8's k = 1..8 order of the neighbour
            the child's path arrives through
  A[cell] = A[nearest node at or above it] below a source, +0.0 elsewhere
"""
from __future__ import annotations

import numpy as np

from . import d8_plan

DX = (0, 1, 1, 0, -1, -1, -1, 0, 1)
DY = (0, 0, -1, -1, -1, 0, 1, 1, 1)


def build(flowdir, src_cell, poly_off, poly_cells, contcheck=True):
    """Month-invariant part: nodes in evaluation order, their k-ordered children, per-polygon candidates."""
    d = np.ascontiguousarray(flowdir, dtype=np.int16)
    nrow, ncol = d.shape
    plan = d8_plan(d)
    receiver, inmask, level, contam = plan["receiver"], plan["inmask"], plan["level"], plan["contam"]
    nsub = len(src_cell)
    srcmap = {}
    for s in range(nsub):                                   # duplicates: the last assignment wins (R's `[<-`)
        c = int(src_cell[s])
        if level[c] >= 0:
            srcmap[c] = s
    active = set()
    for c in srcmap:                                        # everything at or below a source
        while c >= 0 and level[c] >= 0 and c not in active:
            active.add(c)
            c = int(receiver[c])

    def upstream(c):                                        # (k, cell) of the neighbours draining into c, k ascending
        return [(k, c + DY[k] * ncol + DX[k]) for k in range(1, 9) if inmask[c] >> (k - 1) & 1]

    nodes = [c for c in active if c in srcmap or sum(u in active for _, u in upstream(c)) != 1]
    nodes.sort(key=lambda c: (level[c], c))                 # upstream before downstream
    node_id = {c: i for i, c in enumerate(nodes)}
    seg = {}
    children = [[] for _ in nodes]
    for i, c0 in enumerate(nodes):
        seg[c0] = i
        prev, c = c0, int(receiver[c0])
        while c >= 0 and level[c] >= 0:
            if c in node_id:
                k = next(k for k, u in upstream(c) if u == prev)
                children[node_id[c]].append((k, i))
                break
            seg[c] = i
            prev, c = c, int(receiver[c])
    children = [[i for _, i in sorted(ch)] for ch in children]
    cand, haszero = [], []
    for s in range(nsub):
        seen, zero = [], False
        for c in poly_cells[poly_off[s]:poly_off[s + 1]]:
            c = int(c)
            if level[c] < 0 or (contcheck and contam[c]):
                continue
            if c in active:
                if seg[c] not in seen:
                    seen.append(seg[c])
            else:
                zero = True
        cand.append(seen)
        haszero.append(zero)
    return dict(nodes=nodes, node_src=[srcmap.get(c, -1) for c in nodes], children=children, cand=cand,
                haszero=haszero, nactive=len(active))


def apply(op, QS, f32=False):
    """QS [nsub, ncols] -> QR [nsub, ncols]: what oracle.route gives (a zero may differ in sign)."""
    QS = np.asarray(QS, dtype=np.float64)
    T = np.float32 if f32 else np.float64
    ncols = QS.shape[1]
    val = np.zeros((len(op["nodes"]), ncols), dtype=T)
    for i, s in enumerate(op["node_src"]):
        acc = QS[s].astype(T) if s >= 0 else np.zeros(ncols, dtype=T)
        for j in op["children"][i]:
            acc = (acc + val[j]).astype(T)
        val[i] = acc
    QR = np.full(QS.shape, -np.inf)
    for s, (cand, zero) in enumerate(zip(op["cand"], op["haszero"])):
        m = np.full(ncols, 0.0 if zero else -np.inf)
        for i in cand:
            v = val[i].astype(np.float64)
            m = np.where(v > m, v, m)                      # NaN never wins, like na.rm = TRUE
        QR[s] = m
    return QR
