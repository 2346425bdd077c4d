"""Pins of the CPU oracle against the reference's own code (needs /root/reference; skipped on the GPU box).

1. Gauss tables and 1D Lagrange polynomials of oracle/driver.c and of the CUDA emitter == the oomph-lib sources compiled
   into oracle/_ref/libref_shim.so (bit-exact).
2. The generated-format plugin + driver compiled against the reference's jitbridge.h / jitbridge_hang.h gives bit-identical
   element matrices and assembled CSR as the build against the oracle's restated headers.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from problems import make_oracle, make_problem

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.join(os.path.dirname(HERE), "oracle")
HAVE_REF = os.path.isdir("/root/reference/src")
pytestmark = pytest.mark.skipif(not HAVE_REF, reason="reference tree not present")


@pytest.fixture(scope="module")
def shim():
    subprocess.run(["make", "-C", ORACLE], check=True, capture_output=True)
    return ctypes.CDLL(os.path.join(ORACLE, "_ref", "libref_shim.so"))


def test_gauss_tables_bit_exact(shim):
    from oracle import build_plugin
    from pyoomph_b200.cuda_emitter import gauss_rule
    pb = make_problem("poisson", 2)
    drv = ctypes.CDLL(build_plugin(pb["code"], pb["code"].name))
    for dim, npt in ((2, 9), (3, 27)):
        kn, w = gauss_rule(dim)
        for ipt in range(npt):
            k_ref, w_ref = (ctypes.c_double * 3)(), ctypes.c_double()
            k_or, w_or = (ctypes.c_double * 3)(), ctypes.c_double()
            shim.ref_gauss(dim, ipt, k_ref, ctypes.byref(w_ref))
            drv.oracle_gauss(dim, ipt, k_or, ctypes.byref(w_or))
            assert list(k_ref)[:dim] == list(k_or)[:dim] == list(kn[ipt])
            assert w_ref.value == w_or.value == w[ipt]
    # the mistyped literal really is there
    k = (ctypes.c_double * 3)(); w_ = ctypes.c_double()
    shim.ref_gauss(2, 8, k, ctypes.byref(w_))
    assert k[0] == 0.774596662941483 and k[0] != 0.774596669241483


def test_lagrange_polynomials_bit_exact(shim):
    from oracle import build_plugin
    from pyoomph_b200.cuda_emitter import _lag
    pb = make_problem("poisson", 2)
    drv = ctypes.CDLL(build_plugin(pb["code"], pb["code"].name))
    rng = np.random.default_rng(0)
    pts = list(rng.uniform(-1, 1, 20)) + [-0.774596669241483, 0.0, 0.774596662941483, 0.77459666924148]
    for order in (2, 3):
        for s in pts:
            p, d = (ctypes.c_double * 3)(), (ctypes.c_double * 3)()
            shim.ref_lagrange(order, ctypes.c_double(s), p, d)
            P, D = _lag(order, float(s))
            assert list(p)[:order] == P and list(d)[:order] == D
            # oracle: 2D tensor product at (s, s) contains the 1D values as products
            psi, dpsi = (ctypes.c_double * 9)(), (ctypes.c_double * 18)()
            sv = (ctypes.c_double * 2)(s, s)
            drv.oracle_dshape_local(2, order, sv, psi, dpsi)
            # (the driver is built like the reference's JIT code, -O3 -march=native, so gcc may contract 1.0-s*s into an
            #  FMA: the last bit can differ from the uncontracted shim; the reference's own bits depend on its build flags)
            for i in range(order):
                for j in range(order):
                    assert abs(psi[i * order + j] - p[i] * p[j]) <= 4e-16
                    assert abs(dpsi[(i * order + j) * 2 + 0] - p[i] * d[j]) <= 8e-16
                    assert abs(dpsi[(i * order + j) * 2 + 1] - d[i] * p[j]) <= 8e-16


@pytest.mark.parametrize("kind,N", [("ns_unsteady", 3), ("ale", 3), ("heat3d", 2), ("ale_axi_obs", 3)])
def test_reference_headers_give_identical_results(kind, N):
    pb = make_problem(kind, N)
    out = []
    for ref in (False, True):
        op = make_oracle(pb, reference_headers=ref)
        e = op.element(pb["mesh"].n_elem // 2, flag=2)
        r, m = op.assemble(flag=2)
        obs = (np.array(list(op.evaluate_integral_expressions().values())),) if pb["code"].integral_expressions else ()
        out.append(e + (r,) + tuple(x for mat in m for x in mat) + obs)     # EvalIntegralExpression through jitbridge.h:469 too
        op.close()
    for a, b in zip(*out):
        assert np.array_equal(a, b)
