"""CPU tests of the drop-in boundary: the shared libraries load, export exactly what include/ntt_b200.h
declares, the reference-named shim exports the reference's symbols, and -- on a box without a GPU -- every
compute entry point fails loudly instead of falling back to a CPU path."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ntt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ntt_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(ntt):
    declared = header_symbols()
    assert len(declared) >= 35
    assert sorted(ntt.EXPORTS) == declared, "python EXPORTS list and the header disagree"
    lib = ctypes.CDLL(ntt.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libntt_b200.so does not export %s" % name
    # nothing else leaks out of the library
    out = subprocess.run(["nm", "-D", "--defined-only", ntt.LIB_PATH], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert sorted(exported) == declared


def test_dropin_library_exports_reference_names(ntt):
    out = subprocess.run(["nm", "-D", "--defined-only", ntt.DROPIN_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    # the three non-inline functions of include/ntt_reference.h (lines 13, 33, 44)
    assert exported == ["fwd_ntt_ref_harvey_lazy", "fwd_ntt_ref_harvey_lazy_dbl", "inv_ntt_ref_harvey"]
    ctypes.CDLL(ntt.DROPIN_PATH)  # resolves libntt_b200.so through its $ORIGIN rpath


def test_version_and_device_count(ntt):
    assert "sm_100a" in ntt.version()
    assert ntt.device_count() >= 0


def test_product_does_not_reference_the_oracle():
    """The shipped path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "optimized-number-theoretic-transform-implementations_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".c", ".h", ".cu", ".cuh", ".py")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), "%s mentions the oracle" % f
    out = subprocess.run(["ldd", os.path.join(pkg, "libntt_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "ntt_ref" not in out


def test_no_cpu_fallback_without_gpu(ntt):
    if ntt.device_count() > 0:
        pytest.skip("a GPU is present; the loud-failure path is for CPU-only boxes")
    with pytest.raises(ntt.NttError, match="no CUDA device"):
        ntt.Plan.from_psi(256, 7681, 62)
    w = ntt.calc_w(62, 256, 7681)
    wc = ntt.calc_w_con(w, 7681)
    a = np.arange(256, dtype=np.uint64)
    with pytest.raises(ntt.NttError):
        ntt.fwd_ntt_ref_harvey(a, 256, 7681, w, wc)
    assert np.array_equal(a, np.arange(256, dtype=np.uint64)), "input must be untouched on failure"


def test_argument_validation_messages(ntt):
    """Shape errors are reported before any device work (same on CPU and GPU boxes)."""
    h = ctypes.c_void_p()
    rc = ntt.lib.ntt_b200_plan_create_psi(ctypes.byref(h), 0, 255, 7681, 62)
    assert rc == -1 and b"power of two" in ntt.lib.ntt_b200_last_error()
    rc = ntt.lib.ntt_b200_plan_create_psi(ctypes.byref(h), 0, 256, 7680, 62)
    assert rc == -1 and b"odd" in ntt.lib.ntt_b200_last_error()
    rc = ntt.lib.ntt_b200_plan_create_psi(ctypes.byref(h), 0, 256, 1 << 62, 62)
    assert rc == -1


def test_shard_range_covers_the_batch_contiguously(ntt):
    """ntt_b200_shard_range (pure host arithmetic): contiguous shards, sizes differ by at most one, every unit once."""
    import ctypes as C
    for batch in (0, 1, 7, 48, 4096, 4099):
        for parts in (1, 2, 3, 8):
            pos, sizes = 0, []
            for i in range(parts):
                f, c = C.c_size_t(), C.c_size_t()
                ntt.lib.ntt_b200_shard_range(batch, parts, i, C.byref(f), C.byref(c))
                assert f.value == pos
                pos += c.value
                sizes.append(c.value)
            assert pos == batch and max(sizes) - min(sizes) <= 1
    f, c = C.c_size_t(5), C.c_size_t(5)
    ntt.lib.ntt_b200_shard_range(10, 4, 9, C.byref(f), C.byref(c))      # index out of range: empty shard
    assert (f.value, c.value) == (0, 0)


def test_multi_gpu_api_fails_loudly_without_a_device(ntt):
    if ntt.device_count() > 0:
        import pytest
        pytest.skip("a GPU is present")
    import pytest
    with pytest.raises(ntt.NttError) as e:
        ntt.MultiPlan(1 << 10, 0x1FFFFFC800001, ntt.min_primitive_root(1 << 10, 0x1FFFFFC800001), [0])
    assert "no CUDA device" in str(e.value) or "CPU fallback" in str(e.value)
