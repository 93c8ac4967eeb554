#!/usr/bin/env python
"""Lists the public surface of the reference's `vkjit_core` that front-ends can call — every `pub fn` of `impl Ir`
(libs/vkjit-core/src/internal.rs:167-542, the `bop!` expansions included), of `impl VarId`, `impl Var`, `impl VarType`
and the crate's re-exports — into tests/golden/reference_rust_surface.json.  Names and line numbers only, no code.
tests/test_rust_core_surface.py checks the B200 replacement crate (bindings/rust/vkjit-core-b200) against it."""
import json
import os
import re

REF = "/root/reference/libs/vkjit-core/src"
HERE = os.path.dirname(os.path.abspath(__file__))


def pub_fns(text, impl_header):
    """names of `pub fn` inside the first `impl <header> {` block, with line numbers"""
    start = text.index(impl_header)
    depth, i, out = 0, text.index("{", start), []
    body_start = i
    while True:
        if text[i] == "{":
            depth += 1
        elif text[i] == "}":
            depth -= 1
            if depth == 0:
                break
        i += 1
    body = text[body_start:i]
    line0 = text[:body_start].count("\n") + 1
    for m in re.finditer(r"pub fn ([a-z_0-9]+)", body):
        out.append({"name": m.group(1), "line": line0 + body[:m.start()].count("\n")})
    for m in re.finditer(r"^\s*bop!\((\w+)\);", body, flags=re.M):   # bop!(Add) expands to `pub fn add`
        out.append({"name": m.group(1).lower(), "line": line0 + body[:m.start()].count("\n"), "via": "bop! macro (internal.rs:146-166)"})
    return out


def main():
    internal = open(os.path.join(REF, "internal.rs")).read()
    vartype = open(os.path.join(REF, "vartype.rs")).read()
    lib = open(os.path.join(REF, "lib.rs")).read()
    out = {
        "source": "DoeringChristian/vkjit libs/vkjit-core/src (names and line numbers only)",
        "Ir": pub_fns(internal, "impl Ir {"),
        "VarId": pub_fns(internal, "impl VarId {"),
        "Var": pub_fns(internal, "impl Var {"),
        "VarType": [f for f in pub_fns(vartype, "impl VarType {")],
        "reexports": re.findall(r"pub use ([^;]+);", lib),
        "modules": re.findall(r"pub mod (\w+);", lib),
    }
    with open(os.path.join(HERE, "reference_rust_surface.json"), "w") as f:
        json.dump(out, f, indent=1)
    print({k: len(v) for k, v in out.items() if isinstance(v, list)})


if __name__ == "__main__":
    main()
