"""Host-side check of the K3 step/adjoint algebra (hy2dl_b200/csrc/conceptual_models.cuh) against
the reference's golden outputs and autograd gradients.  The CUDA kernels compile this very header;
here it is built with g++ into a tiny harness (tests/host/) -- test infrastructure, never shipped.
Tolerance: rtol 1e-4 (the north-star bound) with -ffp-contract=off so that rounding matches the
device build of these files (-fmad=false)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from _util import DIRECT_CASES, ROOT, assert_close, bucket_of, golden

SRC = os.path.join(ROOT, "tests", "host", "conceptual_host.cpp")
LIB = os.path.join(ROOT, "tests", "host", "_build", "libconceptual_host.so")
SHM_NAMES = ["dd", "f_thr", "sumax", "beta", "perc", "kf", "ki", "kb"]
HBV_NAMES = ["BETA", "FC", "K0", "K1", "K2", "LP", "PERC", "UZL", "TT", "CFMAX", "CFR", "CWH", "BETAET"]
SHM_STATES = ["ss", "sf", "su", "si", "sb"]
HBV_STATES = ["SNOWPACK", "MELTWATER", "SM", "SUZ", "SLZ"]
# model id (include/hy2dl_b200.h HY2DL_MODEL_*), parameter order, storage order
MODELS = {"shm": (0, SHM_NAMES, SHM_STATES), "hbv": (1, HBV_NAMES, HBV_STATES),
          "linear_reservoir": (2, ["ki", "aux_ET"], ["si"]),
          "nonsense": (3, ["dd", "sumax", "beta", "ki", "kb"], ["ss", "su", "si", "sb"])}


@pytest.fixture(scope="module")
def host():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, SRC], check=True)
    return C.CDLL(LIB)


def fp(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("name", DIRECT_CASES)
def test_step_and_adjoint_match_reference(host, name):
    g = golden(name)
    B, T, m, ncols, dyn = [int(v) for v in g["meta"]]
    model, names, snames = MODELS[bucket_of(name)]
    NS = len(snames)
    N = B * m
    xc = g["x_conceptual"]
    tm = (xc[:, :, 2] + xc[:, :, 3]) / 2 if ncols == 4 else xc[:, :, 2]
    forc = np.ascontiguousarray(np.stack([xc[:, :, 0].T, xc[:, :, 1].T, tm.T], axis=1), dtype=np.float32)  # [T,3,B]
    par, dpar, ts = [], [], []
    for n in names:
        key = "param." + n
        if key not in g:
            par.append(None); dpar.append(None); ts.append(0)
            continue
        p = g[key]  # [B, T or 1, m]
        plane = np.ascontiguousarray(p.transpose(1, 0, 2).reshape(-1, N), dtype=np.float32)
        par.append(plane if dyn else plane[0].copy())
        dpar.append(np.zeros_like(par[-1]))
        ts.append(N if dyn else 0)
    P = (C.c_void_p * 16)(*[None if a is None else a.ctypes.data for a in par])
    DP = (C.c_void_p * 16)(*[None if a is None else a.ctypes.data for a in dpar])
    TS = (C.c_int64 * 16)(*ts)
    init = np.ascontiguousarray(np.stack([g["init." + s].reshape(N) for s in snames]), dtype=np.float32)
    dstates = np.ascontiguousarray(
        np.stack([g["ws." + s].transpose(1, 0, 2).reshape(T, N) for s in snames], axis=1), dtype=np.float32)
    dq = np.ascontiguousarray(g["wq"][:, :, 0].T, dtype=np.float32)
    states = np.zeros((T, NS, N), dtype=np.float32)
    q = np.zeros((T, B), dtype=np.float32)
    dinit = np.zeros((NS, N), dtype=np.float32)
    host.conceptual_host(model, C.c_int64(B), C.c_int64(T), C.c_int64(m), fp(forc), P, TS, fp(init),
                         int("param.BETAET" in g or model != 1), fp(states), fp(q), fp(dq), fp(dstates), DP, fp(dinit))
    tol = dict(rtol=1e-4, atol=1e-6)
    assert_close(q.T[:, :, None], g["y_hat"], what="y_hat", **tol)
    for i, s in enumerate(snames):
        assert_close(states[:, i].reshape(T, B, m).transpose(1, 0, 2), g["state." + s], what="state " + s, **tol)
        assert_close(dinit[i].reshape(B, m), g["grad.init." + s], what="grad init " + s, **tol)
    for n, d in zip(names, dpar):
        if d is None:
            continue
        got = d.reshape(-1, B, m).transpose(1, 0, 2)
        assert_close(got, g["grad.param." + n], what="grad param " + n, **tol)
