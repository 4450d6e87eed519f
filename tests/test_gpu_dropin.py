"""GPU test: the reference's OWN correctness driver (tests/main.c + tests/test_correctness.c, compiled
unmodified by oracle/Makefile into oracle/_ref/ntt-variants-dropin) linked against libntt_b200_dropin.so.
Its fwd_ntt_ref_harvey / inv_ntt_ref_harvey / fwd_ntt_ref_harvey_dbl calls run on the GPU; every other
variant it cross-checks (SEAL, radix-4, radix-4x4) is the reference's CPU code and must memcmp-equal our
output on all 19 cases (tests/test_correctness.c:256-284)."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
BINARY = os.path.join(ROOT, "oracle", "_ref", "ntt-variants-dropin")


def test_reference_driver_passes_with_gpu_dropin():
    if not os.path.exists(BINARY):
        pytest.skip("oracle/_ref/ntt-variants-dropin not built (needs /root/reference at build time)")
    out = subprocess.run([BINARY], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    text = out.stdout
    # the driver's exit code is always 0 (tests/main.c:41-44); failures only show as "Bad results" lines
    assert "Bad results" not in text, [l for l in text.splitlines() if "Bad" in l][:5]
    assert text.count("Test ") >= 19, text[-500:]  # the driver stops at the first failing case
    assert text.count("Running inv_ntt_ref_harvey") >= 19


BENCH = os.path.join(ROOT, "oracle", "_ref", "ntt-variants-bench-dropin")


def test_reference_bench_driver_runs_with_gpu_dropin():
    """The reference's own bench driver (tests/main.c + tests/bench.c, -DTEST_SPEED) linked against the drop-in, in
    its single-function mode (`ntt-variants-bench 0` = fwd_ntt_ref_harvey on tests[9], tests/main.c:12-17): the
    number it prints is the single-polynomial GPU round trip (H2D + kernel + D2H + sync per call) in ns, measured by
    the reference's MEASURE macro.  Function 1 (fwd_ntt_seal, the reference's CPU code) runs beside it."""
    if not os.path.exists(BENCH):
        pytest.skip("oracle/_ref/ntt-variants-bench-dropin not built (needs /root/reference at build time)")
    ns = {}
    for func in (0, 1):
        out = subprocess.run([BENCH, str(func)], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        assert "cycle=" in out.stdout, out.stdout[-300:]
        ns[func] = int(out.stdout.split("cycle=")[1].split()[0])
    print("single-polynomial N=2^14 latency: GPU drop-in %d ns per call, reference fwd_ntt_seal (CPU) %d ns" % (ns[0], ns[1]))
    assert 0 < ns[0] < 50_000_000 and ns[1] > 0
