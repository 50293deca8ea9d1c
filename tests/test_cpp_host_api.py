"""The C++ host-side mirror of the reference API (include/ruckig_b200.hpp) over the C ABI.

tests/cpp/host_api_test.cpp holds ports of the reference's first integration tests (test_suite/tests/tests.rs:33-303)
written against that header. Without a GPU the program must build, link against libruckig_b200.so and be refused by the
library (there is no CPU path); on a B200 (-m gpu) the ports run.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "rsruckig_b200", "csrc")


@pytest.fixture(scope="module")
def binary(tmp_path_factory):
    if not os.path.exists(os.path.join(LIBDIR, "libruckig_b200.so")):
        import __graft_entry__ as entry
        entry.build()
    exe = str(tmp_path_factory.mktemp("cpp") / "host_api_test")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "cpp", "host_api_test.cpp"),
           "-L" + LIBDIR, "-lruckig_b200", "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_builds_and_the_library_refuses_to_work_without_a_device(binary):
    r = subprocess.run([binary, "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    import torch
    if not torch.cuda.is_available():
        assert "RK_ERR_NO_DEVICE" in r.stdout


@pytest.mark.gpu
def test_reference_tests_through_the_cpp_api(binary):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    r = subprocess.run([binary], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().startswith("OK"), r.stdout + r.stderr
