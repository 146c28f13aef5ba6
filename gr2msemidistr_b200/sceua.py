"""Population-synchronous SCE-UA (Duan, Sorooshian & Gupta 1992/1994) -- the caller side of
R/Optim_GR2MSemiDistr.R:225-229, where the reference hands OFUN to rtop::sceua.

rtop::sceua evaluates one parameter vector at a time; here every complex performs its competitive
complex evolution (CCE) step in lock-step, and several independent searches (`restarts`) advance together,
so each round yields one batch of candidate points for gr2m_objective_batch (the initial populations are
one batch of restarts*ngs*(2n+1) points, each evolution round up to three batches of <= restarts*ngs
points: reflection, contraction, random replacement).  Algorithm constants
follow rtop::sceua's defaults (npg = 2n+1, nps = n+1, nspl = npg, alpha = 1, beta = 0.5, kstop = 5,
pcento = 0.01, peps = 1e-4, initial point included).  The random stream is NumPy's, not R's, so the
search path differs from rtop's; parity is on objective values, not on the trajectory.

`shard = (rank, world, allgather)` makes every rank evaluate its slice of each batch and exchange
the values (allgather(local_values, counts) -> full vector); all ranks draw the same random numbers.
"""
from __future__ import annotations

import numpy as np


def split_counts(n: int, world: int):
    return [n * (r + 1) // world - n * r // world for r in range(world)]


def _evaluate(obj, X, shard):
    X = np.ascontiguousarray(X, np.float64)
    if shard is None or shard[1] == 1:
        return np.asarray(obj(X), np.float64)
    rank, world, allgather = shard
    counts = split_counts(len(X), world)
    lo = sum(counts[:rank])
    local = np.asarray(obj(X[lo:lo + counts[rank]]), np.float64) if counts[rank] else np.empty(0)
    return np.asarray(allgather(local, counts), np.float64)


def _search(x0, lower, upper, maxn, ngs, kstop, pcento, peps, rng, max_generations):
    """One SCE-UA search as a coroutine: yields batches of candidate points X[m, n] and is sent their objective
    values f[m]; returns its result dict.  All `ngs` complexes take their CCE step in lock-step."""
    n = lower.size
    npg, nps = 2 * n + 1, n + 1
    nspl = npg
    npt = ngs * npg
    bound = upper - lower
    alpha, beta = 1.0, 0.5

    def clean(f):
        return np.where(np.isnan(f), np.inf, f)       # NA objective (e.g. sd(sim) = 0) never wins

    X = lower + rng.random((npt, n)) * bound
    if x0 is not None:
        X[0] = np.clip(x0, lower, upper)              # iniflg = 1
    F = clean((yield X))
    icall = npt
    order = np.argsort(F, kind="stable")
    X, F = X[order], F[order]
    history = [float(F[0])]
    # triangular selection probabilities over the npg ranks of a complex
    prob = 2.0 * (npg - np.arange(npg)) / (npg * (npg + 1.0))
    gen = 0
    while icall < maxn and gen < max_generations:
        gen += 1
        cx = X.reshape(npg, ngs, n).transpose(1, 0, 2).copy()     # complex k holds points k, k+ngs, ...
        cf = F.reshape(npg, ngs).T.copy()
        for _ in range(nspl):
            # choose nps distinct members per complex (weighted towards the better ones), keep rank order
            keys = rng.random((ngs, npg)) ** (1.0 / prob)         # Efraimidis-Spirakis weighted sampling
            sel = np.sort(np.argsort(-keys, axis=1)[:, :nps], axis=1)
            rows = np.arange(ngs)[:, None]
            sx, sf = cx[rows, sel], cf[rows, sel]
            worst, fw = sx[:, -1], sf[:, -1]
            ce = sx[:, :-1].mean(axis=1)
            snew = ce + alpha * (ce - worst)
            out = np.any((snew < lower) | (snew > upper), axis=1)
            if out.any():
                snew[out] = lower + rng.random((int(out.sum()), n)) * bound
            fnew = clean((yield snew))
            icall += ngs
            bad = fnew > fw
            if bad.any():
                idx = np.flatnonzero(bad)
                cand = worst[idx] + beta * (ce[idx] - worst[idx])
                fc = clean((yield cand))
                icall += len(idx)
                snew[idx], fnew[idx] = cand, fc
                still = fc > fw[idx]
                if still.any():
                    idx2 = idx[still]
                    cand2 = lower + rng.random((len(idx2), n)) * bound
                    fc2 = clean((yield cand2))
                    icall += len(idx2)
                    snew[idx2], fnew[idx2] = cand2, fc2
            wpos = sel[:, -1]
            cx[np.arange(ngs), wpos] = snew
            cf[np.arange(ngs), wpos] = fnew
            o = np.argsort(cf, axis=1, kind="stable")
            cx, cf = cx[rows, o], cf[rows, o]
            if icall >= maxn:
                break
        X = cx.transpose(1, 0, 2).reshape(npt, n)
        F = cf.T.reshape(npt)
        order = np.argsort(F, kind="stable")
        X, F = X[order], F[order]
        history.append(float(F[0]))
        # convergence: parameter spread (peps) or stalled improvement over kstop shuffles (pcento)
        gnrng = np.exp(np.mean(np.log(np.maximum((X.max(0) - X.min(0)) / bound, 1e-300))))
        if gnrng < peps:
            break
        if len(history) > kstop:
            ref = np.mean(np.abs(history[-kstop - 1:]))
            if ref > 0 and abs(history[-1] - history[-kstop - 1]) / ref * 100.0 < pcento:
                break
    return {"par": X[0].copy(), "value": float(F[0]), "evaluations": int(icall), "generations": gen,
            "history": history}


def sceua_batched(obj, x0, lower, upper, maxn=5000, ngs=5, restarts=1, kstop=5, pcento=0.01, peps=1e-4, seed=0,
                  shard=None, max_generations=10_000):
    """Minimise obj over the box [lower, upper].  obj(X[m, n]) -> f[m].

    `ngs` complexes per search (rtop::sceua's default 5, which is what the reference gets, Optim:225-229) and
    `maxn` evaluations per search, as in the reference.  The batch the GPU sees is widened by `restarts`
    independent searches advancing in lock-step (restart 0 starts from x0 like the reference; the others from
    random populations of their own streams), not by inflating the population at a fixed budget -- a wide
    population with maxn = 5000 gets two shuffles and does not converge.  Returns the best search's
    dict(par, value, generations, history) with `evaluations` summed over all searches and `restart` = its index."""
    x0 = np.asarray(x0, np.float64)
    lower = np.asarray(lower, np.float64)
    upper = np.asarray(upper, np.float64)
    if ngs is None:
        ngs = 5
    if ngs < 1 or restarts < 1:
        raise ValueError("ngs and restarts must be >= 1")
    live, asked, done = {}, {}, {}
    for r in range(restarts):
        rng = np.random.default_rng(seed if restarts == 1 else [seed, r])
        live[r] = _search(x0 if r == 0 else None, lower, upper, maxn, ngs, kstop, pcento, peps, rng, max_generations)
        asked[r] = next(live[r])
    while live:
        ids = sorted(live)
        F = _evaluate(obj, np.concatenate([asked[r] for r in ids]), shard)
        lo = 0
        for r in ids:
            m = len(asked[r])
            try:
                asked[r] = live[r].send(F[lo:lo + m])
            except StopIteration as fin:
                done[r] = fin.value
                del live[r], asked[r]
            lo += m
    best = min(sorted(done), key=lambda r: done[r]["value"])
    res = dict(done[best])
    res["evaluations"] = int(sum(d["evaluations"] for d in done.values()))
    res["restart"] = best
    return res
