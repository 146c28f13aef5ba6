"""The N > 1 host path on CPU: two gloo ranks shard the parameter sets of every objective batch,
all-gather the values and must reproduce the single-process result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    import backend_stub
    from gr2msemidistr_b200 import dist as gdist
    from gr2msemidistr_b200 import synth
    from gr2msemidistr_b200.sceua import _evaluate, sceua_batched
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P, E, area, nd = synth.forcing(5, 48)
        region = np.zeros(5, np.int32)
        qobs = synth.qobs_from(np.full(48, 40.0))
        calls = []

        def obj(X):
            calls.append(len(X))
            with backend_stub.Basin(P, E, area, nd, region, 1) as b:
                return b.objective_batch(X, qobs, 6, 1)["of"]

        shard = gdist.shard_spec()
        assert shard is not None and shard[0] == rank and shard[1] == world
        # every rank ends up with the whole table although each contributed only its rows (5 rows over 2 ranks: 3 + 2)
        Pd, Ed = gdist.replicate_forcing(P, E)
        assert np.array_equal(Pd.numpy(), P) and np.array_equal(Ed.numpy(), E)
        par = synth.param_sets(21)                               # 21 sets over 2 ranks: 10 + 11
        full = _evaluate(obj, par, shard)
        lo, hi = gdist.shard_range(21, rank, world)
        assert calls[-1] == hi - lo
        res = sceua_batched(obj, [1000, 1, 1, 1], [1, 0.01, 0.8, 0.8], [2000, 2, 1.2, 1.2], maxn=300, ngs=3, seed=7, shard=shard)
        q.put((rank, full, res["par"], res["value"], res["evaluations"], sum(calls)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gloo_ranks_match_single_process():
    import torch.multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import backend_stub
    from gr2msemidistr_b200 import synth
    from gr2msemidistr_b200.sceua import sceua_batched, split_counts
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    P, E, area, nd = synth.forcing(5, 48)
    region = np.zeros(5, np.int32)
    qobs = synth.qobs_from(np.full(48, 40.0))
    ncalls = []

    def obj(X):
        ncalls.append(len(X))
        with backend_stub.Basin(P, E, area, nd, region, 1) as b:
            return b.objective_batch(X, qobs, 6, 1)["of"]

    ref_full = obj(synth.param_sets(21))
    ref = sceua_batched(obj, [1000, 1, 1, 1], [1, 0.01, 0.8, 0.8], [2000, 2, 1.2, 1.2], maxn=300, ngs=3, seed=7)
    for rank, full, par, val, nev, nlocal in out:
        assert np.array_equal(full, ref_full, equal_nan=True)
        assert np.array_equal(par, ref["par"]) and val == ref["value"] and nev == ref["evaluations"]
    # the two ranks together evaluated exactly what one process evaluates, split almost evenly
    assert out[0][5] + out[1][5] == sum(ncalls)
    assert abs(out[0][5] - out[1][5]) <= len(ncalls)
    assert split_counts(10_000, 8) == [1250] * 8 and sum(split_counts(10_001, 8)) == 10_001
