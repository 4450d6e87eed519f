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
