"""The C++ host side (include/xworld_b200.hpp, the SimulatorInterface-shaped facade) compiled with g++
against the C-ABI library: config 1 (SimpleGame, CPU plumbing) everywhere, simple_race too on a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    import __graft_entry__ as g
    g.build()
    exe = str(tmp_path / "test_interface")
    libdir = os.path.join(ROOT, "xworld_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O1", "-Wall", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "test_interface.cpp"),
                           "-L" + libdir, "-lxworld_b200", "-Wl,-rpath," + libdir])
    return exe


def test_cpp_facade_simple_game(tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_facade_on_gpu(tmp_path):
    out = subprocess.run([_build(tmp_path), "gpu"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
