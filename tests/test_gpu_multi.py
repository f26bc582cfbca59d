"""One process, several GPUs: the multi-device handle (ilqgb_create_multi) shards a batch over the devices and gathers into the
caller's single arrays; per-device kernel attributes (the > 48 KB dynamic shared memory opt-in of the backward pass) are set on
every device a handle runs on.  Needs two visible GPUs (gpurun --gpus 2); skipped otherwise."""
import numpy as np
import pytest

import ilqg_b200
import oracle_lib
import parity_util as PU
from ilqg_b200 import workloads as W

pytestmark = pytest.mark.gpu


def _n_dev():
    return ilqg_b200.Library("car", 0).device_count()


def test_multi_device_handle_on_one_device_equals_plain_handle():
    """the sharding layer itself (contiguous shards, chunks per device, gather by global problem index) with a single GPU listed
    twice is not allowed to change anything either: devices=[0] and devices=[0, 0]"""
    B, T = 200, 60
    x0, u0 = W.car_batch(B, T=T, seed=81)
    outs = []
    for devs in (None, [0], [0, 0]):
        s = ilqg_b200.BatchSolver("car", 0, B, T, devices=devs, chunks=2 if devs else 0)
        s.set_options({"max_iter": 8}); s.set_params(W.CAR_PARAMS)
        out = s.solve_host(x0, u0)
        assert s.devices() == (len(devs) if devs else 1)
        out2 = s.solve(x0, u0)
        for k in out:
            assert np.array_equal(out[k], out2[k]), k
        out["l"] = s.get("l"); out["lam"] = s.get("lambda"); out["nbp"] = s.get_int("n_backpass")
        outs.append(out)
        s.close()
    for o in outs[1:]:
        for k in outs[0]:
            assert np.array_equal(outs[0][k], o[k]), k
    ora = oracle_lib.OracleLib(PU.oracle_kinds("car", 0)[0], "car", 0).solve_batch(x0, u0, W.CAR_PARAMS, {"max_iter": 8.0}, 4, want_traj=True)
    assert np.array_equal(outs[0]["cost"], ora["cost"]) and np.array_equal(outs[0]["x"], ora["x"])


@pytest.mark.skipif(_n_dev() < 2, reason="needs two GPUs")
def test_batch_sharded_over_two_gpus_in_one_process():
    B, T = 333, 80
    x0, u0 = W.car_batch(B, T=T, seed=83)
    s = ilqg_b200.BatchSolver("car", 0, B, T, devices=[0, 1])
    s.set_options({"max_iter": 10}); s.set_params(W.CAR_PARAMS)
    a = s.solve_host(x0, u0)
    b = s.solve(x0, u0)
    assert s.devices() == 2
    s.close()
    one = ilqg_b200.BatchSolver("car", 0, B, T, device=1)
    one.set_options({"max_iter": 10}); one.set_params(W.CAR_PARAMS)
    c = one.solve(x0, u0)
    one.close()
    for k in a:
        assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k]), k


@pytest.mark.skipif(_n_dev() < 2, reason="needs two GPUs")
def test_large_shared_memory_kernels_on_every_device():
    """carhx FULL_DDP=1 needs 2 * 55 * 64 * 8 = 56 320 B of dynamic shared memory in the backward pass, above the 48 KB default:
    the opt-in is a per-device function attribute (it was applied on the first device only)."""
    T = 120
    x0, u0 = W.car_batch(3, T=T, seed=85)
    outs = []
    for dev in (0, 1):
        s = ilqg_b200.BatchSolver("carhx", 1, 3, T, device=dev)
        s.set_options({"max_iter": 6}); s.set_params(W.CARHX_PARAMS)
        outs.append(s.solve(x0, u0))
        s.close()
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]), k
    q = ilqg_b200.BatchSolver("quad", 1, 5, 100, devices=[1, 0])
    q.set_options({"max_iter": 4}); q.set_params(W.QUAD_PARAMS)
    xq, uq = W.quad_batch(5, T=100)
    oq = q.solve(xq, uq)
    q.close()
    q1 = ilqg_b200.BatchSolver("quad", 1, 5, 100)
    q1.set_options({"max_iter": 4}); q1.set_params(W.QUAD_PARAMS)
    o1 = q1.solve(xq, uq)
    q1.close()
    for k in oq:
        assert np.array_equal(oq[k], o1[k]), k
