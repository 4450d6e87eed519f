import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "optimized-number-theoretic-transform-implementations_b200"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ntt():
    """The product package (ctypes view of libntt_b200.so); built on demand."""
    try:
        return importlib.import_module(PKG)
    except ImportError:
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, PKG)], check=True)
        return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The reference's own compiled code (oracle/_ref); tests that need it skip when it is absent."""
    from oracle.pyoracle import Reference
    return Reference()


@pytest.fixture(scope="session")
def golden_cases():
    with open(os.path.join(GOLDEN, "cases.json")) as fh:
        return json.load(fh)["cases"]


@pytest.fixture(scope="session")
def golden_case0():
    with open(os.path.join(GOLDEN, "case0_full.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden_synth():
    with open(os.path.join(GOLDEN, "synthetic.json")) as fh:
        return json.load(fh)["sets"]


class CaseTables:
    """Reference-format tables of one parameter set, built by the ORACLE (the checker side)."""

    def __init__(self, oracle, m, q, psi, psi_inv, n_inv):
        self.m, self.N, self.q = m, 1 << m, q
        self.psi, self.psi_inv, self.n_inv = psi, psi_inv, n_inv
        self.w, self.w_con = oracle.tables(self.N, q, psi)
        self.w_inv, self.w_inv_con = oracle.tables(self.N, q, psi_inv)
        self.n_inv_con = oracle.companion(n_inv, q)


@pytest.fixture(scope="session")
def case_tables(oracle, golden_cases):
    cache = {}

    def get(idx):
        if idx not in cache:
            c = golden_cases[idx]
            cache[idx] = CaseTables(oracle, c["m"], c["q"], c["w"], c["w_inv"], c["n_inv"])
        return cache[idx]

    return get


def edge_inputs(N, q):
    z = np.zeros(N, dtype=np.uint64)
    d0 = z.copy(); d0[0] = 1
    dl = z.copy(); dl[N - 1] = 1
    return {"zero": z, "qm1": np.full(N, q - 1, dtype=np.uint64), "delta0": d0, "deltaN": dl}
