"""The solver kernel's LM state machine, without a GPU.

loam_livox_b200/csrc/lm_core.cuh (the restatement of ceres' TrustRegionMinimizer + LevenbergMarquardtStrategy + projected Armijo line search that
lm_solve_kernel runs) is plain C++: tests/cpp/lm_host.cpp compiles it for the host together with the kernel's orchestration of one LM step (both
ComputeStep hypotheses evaluated ahead of the accept test, the gradient test evaluated aside, the pending hand-over).  Here it is driven by the ORACLE's
evaluations (cost, J^T J, J^T r at every trial point) and its trajectory is compared with the oracle's own solver: same number of evaluations and LM
iterations, same termination, same solution to 1e-9 -- on unconstrained problems, on problems where the translation bounds bind (line search), with the
2-iteration cap of solve #1.  The speculative path must also equal the plain serial path bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from loam_livox_b200 import synthetic as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lmh(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("lmhost") / "liblmhost.so")
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "cpp", "lm_host.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    L = C.CDLL(so)
    f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L.lmh_init.argtypes = [C.c_void_p, f64p, C.c_double, C.c_int, f64p]
    L.lmh_step.argtypes = [C.c_void_p, f64p, C.c_int, f64p]
    L.lmh_step.restype = C.c_int
    L.lmh_summary.argtypes = [C.c_void_p, f64p]
    L.lmh_sizeof.restype = C.c_int
    return L


def _sums(oracle, blocks, guess, x, bound):
    cost, g, H = oracle.evaluate(blocks, guess.q, guess.t, x, bound=bound)
    return np.ascontiguousarray(np.concatenate([H[np.triu_indices(6)], g, [cost, float(blocks.shape[0])]]))


def _run(lmh, oracle, blocks, guess, x0, bound, max_iter, speculative):
    st = C.create_string_buffer(lmh.lmh_sizeof())
    x = np.zeros(7)
    lmh.lmh_init(st, np.ascontiguousarray(x0, dtype=np.float64), bound, max_iter, x)
    evals = 0
    for _ in range(400):
        sums = _sums(oracle, blocks, guess, x, bound)
        evals += 1
        nxt = np.zeros(7)
        done = lmh.lmh_step(st, sums, int(speculative), nxt)
        x = nxt
        if done:
            break
    assert done
    summ = np.zeros(6)
    lmh.lmh_summary(st, summ)
    return x, dict(initial_cost=summ[0], final_cost=summ[1], iterations=int(summ[2]), termination=int(summ[3]), evaluations=int(summ[4]))


@pytest.mark.parametrize("seed,bound,max_iter", [(0, 0.3, 50), (1, 0.3, 2), (2, 0.02, 50), (3, 0.005, 50), (4, 10.0, 50), (5, 0.05, 2)])
def test_lm_core_follows_the_oracle_solver(lmh, oracle, seed, bound, max_iter):
    from test_oracle import _blocks
    b, guess, _, _ = _blocks(oracle, seed=seed)
    x0 = np.array([0, 0, 0, 1, 0, 0, 0], float)
    xo, so = oracle.solve(b, guess.q, guess.t, x0, max_iter, bound=bound)
    xs, ss = _run(lmh, oracle, b, guess, x0, bound, max_iter, speculative=True)
    xp, sp = _run(lmh, oracle, b, guess, x0, bound, max_iter, speculative=False)
    # the speculative orchestration (what the kernel does) equals the plain serial state machine, bit for bit
    assert np.array_equal(xs, xp) and ss == sp
    # and both follow the oracle's solver
    assert ss["iterations"] == int(so["iterations"])
    assert abs(ss["initial_cost"] - so["initial_cost"]) <= 1e-12 * so["initial_cost"] and abs(ss["final_cost"] - so["final_cost"]) <= 1e-9 * so["final_cost"]
    assert np.allclose(xs, xo, rtol=0, atol=1e-9), (xs, xo)
    if bound <= 0.02:
        assert np.abs(xs[4:]).max() <= bound * (1 + 1e-12)      # the bounds really bind in these cases
