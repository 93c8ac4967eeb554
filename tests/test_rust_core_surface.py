"""bindings/rust/vkjit-core-b200 (the blind Rust replacement of `vkjit_core`, VERDICT r01 task 6) must offer every
public item a front-end can reach in the reference crate: each `pub fn` of `impl Ir` (internal.rs:167-542), `VarId`,
`Var::ty`, `VarType`'s methods, the module paths and re-exports vkjit-rust / vkjit-python import, and every C entry point
it calls must exist in the generated `vkjit-sys` crate.  No Rust toolchain exists in this image, so this is a textual check."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CRATE = os.path.join(ROOT, "bindings", "rust", "vkjit-core-b200", "src")
SURFACE = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_rust_surface.json")))
# SPIR-V specific items with no meaning for a CUDA backend (vartype.rs:84-119)
NOT_APPLICABLE = {"VarType": {"to_spirv"}}


def crate_fns(path):
    text = open(os.path.join(CRATE, path)).read()
    names = set(re.findall(r"pub fn ([a-z_0-9]+)", text))
    names |= set(re.findall(r"^\s*[bu]op!\((\w+),", text, flags=re.M))      # bop!(add, 0) / uop!(neg, 0) define `pub fn`
    return names, text


def test_every_public_fn_of_the_reference_ir_has_a_counterpart():
    mine, text = crate_fns("internal.rs")
    missing = [f["name"] for f in SURFACE["Ir"] if f["name"] not in mine]
    assert not missing, missing
    for group in ("VarId", "Var"):
        block = text[text.index(f"impl {group} {{"):]
        for f in SURFACE[group]:
            assert re.search(r"pub fn %s\b" % f["name"], block), (group, f["name"])
    vt, _ = crate_fns("vartype.rs")
    for f in SURFACE["VarType"]:
        if f["name"] not in NOT_APPLICABLE["VarType"]:
            assert f["name"] in vt, f["name"]


def test_module_paths_and_reexports_the_front_ends_import():
    lib = open(os.path.join(CRATE, "lib.rs")).read()
    # vkjit-rust: `use vkjit_core::Ir`, `vkjit_core::vartype::VarType`, `vkjit_core::{AsVarType, VarId, VarType}`
    for mod in ("internal", "vartype"):
        assert re.search(r"pub mod %s;" % mod, lib)
    flat = " ".join(re.findall(r"pub use ([^;]+);", lib))
    for item in ("Ir", "VarId", "AsVarType", "VarType"):
        assert re.search(r"\b%s\b" % item, flat), item
    toml = open(os.path.join(CRATE, "..", "Cargo.toml")).read()
    assert 'name = "vkjit-core"' in toml and 'name = "vkjit_core"' in toml    # same crate name: a path swap is the whole switch


def test_every_c_entry_point_the_crate_calls_is_declared_in_vkjit_sys():
    sys_rs = open(os.path.join(ROOT, "bindings", "rust", "vkjit-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (vkjit_[a-z0-9_]+)\(", sys_rs))
    used = set()
    for dirpath, _, files in os.walk(CRATE):
        for fn in files:
            used |= set(re.findall(r"sys::(vkjit_[a-z0-9_]+)\(", open(os.path.join(dirpath, fn)).read()))
    assert used and used <= declared, sorted(used - declared)
    # argument counts agree with the declarations (catches a drifted signature without a compiler)
    for dirpath, _, files in os.walk(CRATE):
        for fn in files:
            text = open(os.path.join(dirpath, fn)).read()
            for m in re.finditer(r"sys::(vkjit_[a-z0-9_]+)\(", text):
                depth, i, args, cur = 1, m.end(), 0, ""
                while depth:
                    ch = text[i]
                    if ch in "([{":
                        depth += 1
                    elif ch in ")]}":
                        depth -= 1
                    if depth == 1 and ch == ",":
                        args += 1; cur = ""
                    elif depth >= 1:
                        cur += ch
                    i += 1
                n_call = args + (1 if cur.strip() else 0)
                decl = re.search(r"pub fn %s\(([^)]*)\)" % m.group(1), sys_rs).group(1)
                n_decl = len([a for a in decl.split(",") if a.strip()])
                assert n_call == n_decl, (fn, m.group(1), n_call, n_decl)


def test_surface_fixture_is_current_when_the_reference_is_present():
    if not os.path.isdir("/root/reference/libs/vkjit-core/src"):
        return
    import subprocess
    import sys
    before = open(os.path.join(ROOT, "tests", "golden", "reference_rust_surface.json")).read()
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "golden", "make_rust_surface.py")], stdout=subprocess.DEVNULL)
    assert open(os.path.join(ROOT, "tests", "golden", "reference_rust_surface.json")).read() == before
