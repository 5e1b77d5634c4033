import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not p.endswith("scalars.npz"))


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def assert_results_equal(got, want, skip=None, label=""):
    """Bit-exact comparison of two (offsets, positions, status) CSR results, per query."""
    g_off, g_pos, g_st = got
    w_off, w_pos, w_st = want
    Q = len(w_st)
    assert len(g_st) == Q and len(g_off) == Q + 1, label
    if skip is None or not np.any(skip):
        if np.array_equal(g_off, w_off) and np.array_equal(g_pos, w_pos) and np.array_equal(g_st, w_st):
            return
    for i in range(Q):
        if skip is not None and skip[i]:
            continue
        a = g_pos[int(g_off[i]):int(g_off[i + 1])]
        b = w_pos[int(w_off[i]):int(w_off[i + 1])]
        assert int(g_st[i]) == int(w_st[i]), f"{label}: query {i}: status {g_st[i]} != {w_st[i]}"
        assert np.array_equal(a, b), f"{label}: query {i}: {a[:8]}.. ({a.size}) != {b[:8]}.. ({b.size})"


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import bindings
    bindings.build()
    return bindings


@pytest.fixture(scope="session", autouse=True)
def _guarded_allocations_stay_intact():
    """KMER_B200_GUARD=1 python -m pytest tests -m gpu: the whole GPU suite with canary zones around every device
    allocation of the library (include/kmer_b200.h); a store outside a buffer anywhere in the session fails it here."""
    yield
    if os.environ.get("KMER_B200_GUARD") and "kmer_index_b200" in sys.modules:
        import kmer_index_b200
        assert kmer_index_b200.guard_violations() == 0, "a kernel stored outside one of the library's device buffers"
