"""The C ABI used from plain C (no Python / PyTorch in the caller)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "c_abi", "c_abi_smoke.c")
LIBDIR = os.path.join(ROOT, "mdp_playground_b200")


def _compile(out):
    cmd = ["gcc", "-O1", SRC, "-I", os.path.join(ROOT, "include"),
           "-I", "/usr/local/cuda/include", "-L", LIBDIR, "-lmdpp_b200",
           "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + LIBDIR,
           "-Wl,-rpath,/usr/local/cuda/lib64", "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr


def test_header_compiles_as_plain_c(tmp_path):
    """include/mdpp_b200.h is valid C and the smoke program links against the
    library's exported symbols (no GPU needed to build)."""
    _compile(str(tmp_path / "smoke"))


@pytest.mark.gpu
def test_c_caller_matches_scalar_restatement(tmp_path):
    exe = str(tmp_path / "smoke")
    _compile(exe)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "mismatches=0" in res.stdout
