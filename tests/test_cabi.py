"""The C ABI: every symbol declared in include/ilqg_b200.h is exported by each built library, static facts are right,
option validation matches the reference's messages, and the product fails loudly without a GPU (no CPU path)."""
import ctypes as C
import os
import re

import pytest

import ilqg_b200
import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBS = [("car", 0), ("car", 1), ("brachi", 0), ("brachi", 1), ("quad", 0), ("quad", 1), ("carhx", 0), ("brachi_hli", 1), ("pend", 0), ("pend", 1)]


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ilqg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ilqgb_\w+)\s*\(", src)))


@pytest.mark.parametrize("problem,ddp", LIBS)
def test_every_declared_symbol_is_exported(problem, ddp):
    syms = declared_symbols()
    assert len(syms) >= 25
    lib = C.CDLL(ilqg_b200.lib_path(problem, ddp))
    for s in syms:
        assert hasattr(lib, s), f"{s} missing from {problem} ddp{ddp}"


def test_static_facts():
    car = ilqg_b200.Library("car", 0)
    assert (car.problem, car.nx, car.nu, car.full_ddp) == ("Car", 4, 2, 0)
    assert car.param_names == ["cf", "cu", "cx", "d", "h", "limA", "limW", "pf", "px"]
    assert car.param_sizes == [4, 2, 2, 1, 1, 2, 2, 4, 2]
    assert car.deriv_doubles_per_step == 18
    assert ilqg_b200.Library("car", 1).deriv_doubles_per_step == 33
    br = ilqg_b200.Library("brachi", 1)
    assert (br.problem, br.nx, br.nu, br.full_ddp, br.param_names) == ("Brachi", 1, 1, 1, ["dx", "g", "yf"])


def test_option_validation_matches_reference_messages():
    L = ilqg_b200.Library("car", 0)
    s = oracle_lib.OracleLib("port", "car", 0).solver(4)
    ref = oracle_lib.OracleLib("reference", "car", 0).solver(4) if oracle_lib.available("reference", "car", 0) else None
    names = ["alpha", "tolFun", "tolConstraint", "tolGrad", "max_iter", "lambdaInit", "dlambdaInit", "lambdaFactor", "lambdaMax",
             "lambdaMin", "regType", "zMin", "debug_level", "w_pen_init_l", "w_pen_init_f", "w_pen_max_l", "w_pen_max_f",
             "w_pen_fact1", "w_pen_fact2", "w_pen_init", "nonsense"]
    values = [-1.0, 0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 6.0, 7.0, [1.0, 0.5], [0.5, 1.0], [0.5, 0.5], [2.0, 1.0]]
    for n in names:
        for v in values:
            want = s.set_opt_raw(n, v)
            assert L.validate_option(n, v) == want, (n, v)
            if ref is not None:
                assert ref.set_opt_raw(n, v) == want, (n, v)


def test_no_cpu_fallback():
    L = ilqg_b200.Library("car", 0)
    if L.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ilqg_b200.BatchSolver("car", 0, 4, 10)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(ilqg_b200, "LIB_DIR", "/nonexistent")
    ilqg_b200.Library._cache.pop(("quadx", 0, None), None)
    with pytest.raises(RuntimeError, match="not found"):
        ilqg_b200.Library("quadx", 0)
