"""Device-pointer entry points, caller-provided streams and error behaviour on a real GPU."""
import ctypes as C

import numpy as np
import pytest

from gr2msemidistr_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    import __graft_entry__ as g
    g.build()
    from gr2msemidistr_b200 import capi as m
    assert m.device_count() >= 1
    m.set_device(0)
    return m


def test_objective_batch_dev_on_a_torch_stream(capi):
    import torch
    P, E, area, nd = synth.forcing(300, 120, seed=77)
    region = np.zeros(300, np.int32)
    par = synth.param_sets(50)
    qobs = synth.qobs_from(np.full(120, 80.0))
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(device=dev)
    with capi.Basin(P, E, area, nd, region, 1) as b:
        host = b.objective_batch(par, qobs, 12, capi.OF_KGE, want_qt=True)
        b.set_stream(st.cuda_stream)
        with torch.cuda.stream(st):
            par_d = torch.from_numpy(par).to(dev, non_blocking=True)
            of_d = torch.empty(50, dtype=torch.float64, device=dev)
            crit_d = torch.empty((50, 3), dtype=torch.float64, device=dev)
            qt_d = torch.empty((50, 120), dtype=torch.float64, device=dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            b.objective_batch_dev(par_d.data_ptr(), 50, qobs, 12, capi.OF_KGE, 0, of_d.data_ptr(), crit_d.data_ptr(), qt_d.data_ptr())
            e1.record()
        st.synchronize()
        assert e0.elapsed_time(e1) > 0 and b.last_kernel_ms() > 0
        assert np.array_equal(of_d.cpu().numpy(), host["of"], equal_nan=True)
        assert np.array_equal(crit_d.cpu().numpy(), host["crit"], equal_nan=True)
        assert np.array_equal(qt_d.cpu().numpy(), host["qt"])
        b.set_stream(None)
        again = b.objective_batch(par, qobs, 12, capi.OF_KGE)
        assert np.array_equal(again["of"], host["of"], equal_nan=True)      # deterministic run to run


def test_accumulate_dev_in_place_on_torch_memory(capi, oracle, rng):
    import torch
    d = synth.d8_grid(200, 300, seed=3)
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(device=dev)
    W = rng.uniform(0, 50, (200 * 300, 8))
    with capi.Plan(dev_ptr=d.to(dev).data_ptr(), shape=(200, 300)) as p:
        ref = oracle.d8_plan(d.numpy())
        got = p.export()
        assert np.array_equal(got["order"], ref["order"]) and np.array_equal(got["level"], ref["level"])
        p.set_stream(st.cuda_stream)
        with torch.cuda.stream(st):
            A = torch.from_numpy(W).to(dev)
            p.accumulate_dev(A.data_ptr(), 8, 8)
            p.mask_dev(A.data_ptr(), 8, 8)
        st.synchronize()
        wide, tail = p.last_phase_ms()
        assert wide >= 0 and tail >= 0
        A = A.cpu().numpy()
    for m in (0, 7):
        Ar = oracle.aread8(d.numpy(), W[:, m].reshape(200, 300), contcheck=True).ravel()
        assert np.array_equal(np.nan_to_num(A[:, m], nan=-1), np.nan_to_num(Ar, nan=-1))


def test_errors_are_reported_not_fatal(capi):
    lib = capi.lib()
    P, E, area, nd = synth.forcing(4, 12)
    with capi.Basin(P, E, area, nd, np.zeros(4, np.int32), 1) as b:
        with pytest.raises(ValueError):
            b.run(np.ones(3))                                   # wrong parameter count (binding check)
        with pytest.raises(capi.GR2MError) as ei:
            b.objective_batch(np.ones((2, 4)), np.ones(12), warmup=12)       # warm-up swallows the whole period
        assert ei.value.code == capi.ERR_INVALID and "warmup" in str(ei.value)
        with pytest.raises(capi.GR2MError):
            b.run(np.ones(4), S0=np.ones(4), R0=None)           # IniState needs both stores
        # NaN forcing / parameters propagate as NaN, they do not crash or hang
        out = b.objective_batch(np.array([[np.nan, 1.0, 1.0, 1.0], [300.0, 0.9, 1.0, 1.0]]), np.arange(1.0, 13.0), 0, capi.OF_NSE)
        assert np.isnan(out["of"][0]) and np.isfinite(out["of"][1])
    Pn = P.copy(); Pn[1, 3] = np.nan
    with capi.Basin(Pn, E, area, nd, np.zeros(4, np.int32), 1) as b:
        r = b.run(np.array(synth.SET0))
        assert np.isnan(r["QS"][1, 3:]).all() and np.isfinite(r["QS"][0]).all()
    assert lib.gr2m_set_device(99) == capi.ERR_INVALID
    with pytest.raises(ValueError):
        capi.Plan(np.ones((3, 3), np.int16)).accumulate(np.ones((2, 2)))     # wrong grid shape


def test_plain_c_caller_runs_on_the_gpu(capi, tmp_path):
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "c_abi_demo")
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(root, "include"), "-o", exe,
                           os.path.join(root, "examples", "c_abi_demo.c"), "-L", libdir, "-lgr2m_b200",
                           f"-Wl,-rpath,{libdir}", "-lm"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "set 0: 1 - NSE = 0.000" in out.stdout and "set 1:" in out.stdout


def test_basin_from_device_tables(capi):
    # gr2m_basin_create_dev: the forcing tables already in HBM (what replicate_forcing hands over)
    import torch
    P, E, area, nd = synth.forcing(70, 36)
    region = (np.arange(70) % 3).astype(np.int32)
    par = synth.param_sets(9, nreg=3)
    qobs = synth.qobs_from(np.full(36, 55.0))
    with capi.Basin(P, E, area, nd, region, 3) as b:
        want = b.objective_batch(par, qobs, 6, capi.OF_NSE)
        sim = b.run(par[0])
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        Pd, Ed = torch.from_numpy(P).cuda(non_blocking=True), torch.from_numpy(E).cuda(non_blocking=True)
    with capi.Basin(None, None, area, nd, region, 3, dev_ptrs=(Pd.data_ptr(), Ed.data_ptr()), shape=P.shape,
                    stream=st.cuda_stream) as b:
        got = b.objective_batch(par, qobs, 6, capi.OF_NSE)
        sim2 = b.run(par[0])
    assert np.array_equal(got["of"], want["of"]) and np.array_equal(got["crit"], want["crit"], equal_nan=True)
    for k in ("PR", "AE", "SM", "RU", "QS"):
        assert np.array_equal(sim[k], sim2[k])
