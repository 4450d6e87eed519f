"""examples/batch_ntt.c: the C-ABI header is plain C11, the library links from C, and the program behaves on both
kinds of box (exit 2 with a clear message on a CPU-only box -- no fallback; exit 0 after an exact round trip on a GPU)."""
import os
import subprocess

import pytest

from conftest import PKG

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, PKG)


def _build(tmp_path, name="batch_ntt"):
    exe = str(tmp_path / name)
    cmd = ["gcc", "-std=c11", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", name + ".c"), "-L", LIBDIR, "-lntt_b200", "-Wl,-rpath," + LIBDIR, "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


def test_c_example_builds_and_refuses_to_run_without_a_gpu(ntt, tmp_path):
    exe = _build(tmp_path)
    if ntt.device_count() > 0:
        pytest.skip("a GPU is present; see the gpu-marked test")
    r = subprocess.run([exe, "10", "3"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("logn,batch", [(14, 64), (10, 5), (16, 3)])
def test_c_example_round_trip(ntt, tmp_path, logn, batch):
    exe = _build(tmp_path)
    r = subprocess.run([exe, str(logn), str(batch)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "round trip exact" in r.stdout and "tables match calc_w" in r.stdout


def test_multi_gpu_c_example_builds_and_refuses_to_run_without_a_gpu(ntt, tmp_path):
    exe = _build(tmp_path, "multi_gpu_ntt")
    if ntt.device_count() > 0:
        pytest.skip("a GPU is present; see the gpu-marked test")
    r = subprocess.run([exe, "10", "3"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("logn,per_gpu", [(14, 64), (12, 7)])
def test_multi_gpu_c_example(ntt, tmp_path, logn, per_gpu):
    """examples/multi_gpu_ntt.c: the device-list form of the C-ABI (every visible GPU), sharded == single device."""
    exe = _build(tmp_path, "multi_gpu_ntt")
    r = subprocess.run([exe, str(logn), str(per_gpu)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "sharded == single device" in r.stdout and "round trip exact" in r.stdout
