"""A non-Python host: tests/c_harness/find_eqb.c is compiled with gcc against include/cloudatlas_b200.h and run
against the shared library with host buffers -- the ccall sequence of the Julia binding (ODEModel constructor,
model.f, model.Df, ensemble f from page-locked memory, hookstepsolve(model, ...)), i.e. the reference's
test/runtests.jl:40-82 and notebooks/find-eqb.ipynb."""
import os
import subprocess

import pytest

import __graft_entry__ as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_harness", "find_eqb.c")
LIBDIR = os.path.join(ROOT, "cloudatlas.jl_b200", "lib")


def build(tmp_path):
    exe = str(tmp_path / "find_eqb")
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-L", LIBDIR,
           "-lcloudatlas_b200", "-lm", f"-Wl,-rpath,{LIBDIR}", "-o", exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_c_harness_compiles_and_links(tmp_path):
    """CPU: the header is valid C and every entry point the harness uses is exported."""
    g.build()
    build(tmp_path)


@pytest.mark.gpu
def test_c_harness_runs(tmp_path):
    exe = build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "harness ok" in out.stdout
