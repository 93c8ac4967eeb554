"""include/vkjit.hpp — the header-only C++17 mirror of the reference's `vkjit-rust` crate (Var with Clone/Drop
ownership, operators, eval / schedule, the free functions) over the C ABI.  tests/cpp_client.cpp is compiled with
-Wall -Wextra -Werror; its host part needs no GPU, its device part replays src/main.rs and the two front-end tests of
libs/vkjit-rust/src/types.rs:213-242 on the B200.  (Named test_zz_* so that it runs after the parity suites.)"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    exe = str(tmp_path / "cpp_client")
    lib_dir = os.path.join(ROOT, "vkjit_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp_client.cpp"), "-o", exe, "-L", lib_dir, "-lvkjit_b200",
                           "-Wl,-rpath," + lib_dir])
    return exe


def test_cpp_front_end_host_part(tmp_path):
    r = subprocess.run([build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpp client ok" in r.stdout


@pytest.mark.gpu
def test_cpp_front_end_on_the_device(tmp_path):
    r = subprocess.run([build(tmp_path), "--device"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpp client ok (device)" in r.stdout
