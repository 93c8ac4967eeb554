"""include/vkjit_b200.h is a plain C header: a C99 client compiles with -Wall -Werror -pedantic, links against
libvkjit_b200.so and drives trace construction + codegen through the ABI."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c99_client_builds_and_runs(tmp_path):
    exe = str(tmp_path / "c_client")
    lib_dir = os.path.join(ROOT, "vkjit_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_client.c"), "-o", exe, "-L", lib_dir, "-lvkjit_b200",
                           "-Wl,-rpath," + lib_dir])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c client ok" in r.stdout
