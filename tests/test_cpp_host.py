"""The C++ mirror of ICAR's plugin surface (include/icar_b200_mod.hpp) and the reference's two test scenarios
written against it (tests/cpp/test_reference_scenarios.cpp): the compiled-host-language side of the boundary.
Without a GPU the program must build, link against libicar_b200.so and stop loudly; on a GPU its printed
rows are compared with the frozen vectors."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_reference_scenarios.cpp")


@pytest.fixture(scope="module")
def program(tmp_path_factory):
    import icar_b200
    icar_b200.lib()                                   # the library has to exist (fails loudly otherwise)
    exe = str(tmp_path_factory.mktemp("cpp") / "scenarios")
    libdir = os.path.join(ROOT, "icar_b200")
    cmd = ["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC,
           "-L", libdir, "-licar_b200", f"-Wl,-rpath,{libdir}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_cpp_host_builds_and_fails_loudly_without_a_gpu(program):
    if has_gpu():
        pytest.skip("a GPU is present: the scenarios themselves run in the gpu tests")
    r = subprocess.run([program, "plugin"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3
    assert "no CPU fallback" in r.stderr and "error stop" in r.stderr


@pytest.mark.gpu
def test_reference_scenarios_from_cpp(program, golden_dir):
    r = subprocess.run([program], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    # src/tests/test_mpdata.f90: first revolution against the frozen row (bit for bit: the leaf call is exact)
    row = np.array([float(x) for x in [l for l in lines if l.startswith("mpdata loop 1:")][0].split(":")[1].split()], np.float32)
    g = np.load(os.path.join(golden_dir, "mpdata_scenario_f32.npz"))["rows"][0]
    assert np.array_equal(row, g), (row, g)
    # src/tests/test_mp_simple.f90: 100 chained calls, 2e-6 of each column's maximum
    tab = np.array([[float(x) for x in l.split(":")[1].split()] for l in lines if l.startswith("mp_simple call")], np.float32)
    want = np.load(os.path.join(golden_dir, "mp_simple_scenario_f32.npz"))["table"]
    assert tab.shape == want.shape
    assert (np.abs(tab - want).max(axis=0) / np.abs(want).max(axis=0) <= 2e-6).all()
    # the plugin surface: vapour condensed, something reached the ground or stayed as cloud, five scalars advected
    last = [l for l in lines if l.startswith("plugin surface:")][0]
    qv0, qv1, cloud, precipitation, nvars = [float(x) for x in re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?", last)]
    assert qv1 < qv0 and (cloud > 0 or precipitation > 0) and nvars == 5
