"""Runs the C++ drop-in's own parity binary (arrow_b200/cpp/b200_host_test.cc): identical inputs
through arrow::compute::CallFunction / Acero with the stock CPU registry of the installed reference
binary and with the nested B200 registry; results and error texts must match."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "arrow_b200", "lib", "b200_host_test")


@pytest.mark.gpu
def test_cpp_dropin_parity():
    assert os.path.exists(EXE), "run __graft_entry__.build() first"
    p = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    tail = "\n".join(p.stdout.splitlines()[-15:])
    assert p.returncode == 0, tail + p.stderr[-2000:]
    assert "PASS" in p.stdout and "FAIL" not in p.stdout, tail


def test_cpp_dropin_built_and_links_reference_binaries():
    """CPU: the plugin is built, links the installed reference libraries and exports its entry points."""
    so = os.path.join(ROOT, "arrow_b200", "lib", "libarrow_b200_compute.so")
    assert os.path.exists(so), "run __graft_entry__.build() first"
    ldd = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    assert "libarrow.so.2400" in ldd and "libarrow_compute.so.2400" in ldd and "libarrow_b200.so" in ldd
    assert "not found" not in ldd, ldd
    nm = subprocess.run(["nm", "-D", "--defined-only", "-C", so], capture_output=True, text=True).stdout
    for sym in ("arrow_b200::Runtime::Get", "arrow_b200::RegisterFunctions", "arrow_b200::MakeGrouper",
                "arrow_b200::RegisterAceroNodes", "arrow_b200::ToDevice"):
        assert sym in nm, sym
