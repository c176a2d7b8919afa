"""tests/c_abi_smoke.c: a plain-C program that links -ltritd (no Python in the loop).  CPU: it must build, link
against every symbol it uses and find that the library refuses without a CUDA device.  GPU: it runs the 7x6x5 case
through tritd_admm_f64 and compares with the outputs of the reference's own .m source."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200", "tritd")


def _build(tmp_path):
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-L" + LIBDIR, "-ltritd",
                           "-Wl,-rpath," + LIBDIR, "-lm", "-o", exe])
    return exe


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_c_program_links_and_refuses_without_gpu(tmp_path):
    exe = _build(tmp_path)
    if _has_gpu():
        pytest.skip("GPU present: covered by the gpu test")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "no CPU fallback" in out.stdout


@pytest.mark.gpu
def test_c_program_runs_the_solver(tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "c_abi_smoke: ok" in out.stdout
