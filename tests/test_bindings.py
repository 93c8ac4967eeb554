"""The (uncompiled) Rust FFI crate must declare exactly the entry points of include/vkjit_b200.h."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rust_sys_crate_covers_the_c_abi():
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "vkjit_b200.h")).read(), flags=re.S)
    c_names = set(re.findall(r"\b(vkjit_[a-z0-9_]+)\s*\(", hdr))
    rs = open(os.path.join(ROOT, "bindings", "rust", "vkjit-sys", "src", "lib.rs")).read()
    rs_names = set(re.findall(r"pub fn (vkjit_[a-z0-9_]+)\(", rs))
    assert c_names == rs_names, (sorted(c_names - rs_names), sorted(rs_names - c_names))


def test_rust_sys_crate_is_what_the_generator_emits():
    """lib.rs's extern block is generated (bindings/rust/gen_sys.py): signatures, not only names, follow the header."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bindings", "rust", "gen_sys.py"), "--check"])
    assert r.returncode == 0, "run `python bindings/rust/gen_sys.py` after changing include/vkjit_b200.h"
