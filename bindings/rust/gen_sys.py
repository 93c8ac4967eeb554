#!/usr/bin/env python
"""Regenerates the `extern "C"` block of bindings/rust/vkjit-sys/src/lib.rs from include/vkjit_b200.h.

    python bindings/rust/gen_sys.py            # rewrite lib.rs in place
    python bindings/rust/gen_sys.py --check    # exit 1 if lib.rs is out of date (tests/test_bindings.py)

Everything above the `extern "C" {` line of lib.rs (types, constants, the stats struct) is hand-written and kept.
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HDR = os.path.join(ROOT, "include", "vkjit_b200.h")
OUT = os.path.join(ROOT, "bindings", "rust", "vkjit-sys", "src", "lib.rs")

SCALARS = {"int32_t": "i32", "uint32_t": "u32", "uint64_t": "u64", "int64_t": "i64", "size_t": "usize", "float": "f32", "double": "f64",
           "char": "c_char", "void": "c_void", "int": "i32", "uint8_t": "u8",
           "vkjit_status": "vkjit_status", "vkjit_type": "vkjit_type", "vkjit_var": "vkjit_var", "vkjit_ir": "vkjit_ir",
           "vkjit_stats_t": "vkjit_stats_t"}
RUST_KEYWORDS = {"type", "in", "ref", "box", "fn", "mod", "move", "self", "impl", "loop", "match", "use", "where"}


def rust_type(ctype: str) -> str:
    t = ctype.strip()
    m = re.fullmatch(r"void\s*\(\s*\*\s*\)\s*\(\s*void\s*\*\s*\)", t)
    if m:
        return "Option<unsafe extern \"C\" fn(*mut c_void)>"
    stars = t.count("*")
    t = t.replace("*", " ")
    toks = t.split()
    const = "const" in toks
    base = [k for k in toks if k not in ("const", "struct", "unsigned")]
    name = SCALARS[base[0]] if base else "u32"
    if "unsigned" in toks and not base:
        name = "u32"
    for i in range(stars):
        # `const T*` -> *const T; a second level (`T* const*`) stays *const as well when the header says const
        name = ("*const " if const else "*mut ") + name
    return name


def parse(hdr_text: str):
    text = re.sub(r"/\*.*?\*/", "", hdr_text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    fns = []
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(vkjit_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        if ret.startswith("typedef"):
            continue
        params = []
        if args and args != "void":
            depth, cur, parts = 0, "", []
            for ch in args:
                if ch == "(":
                    depth += 1
                if ch == ")":
                    depth -= 1
                if ch == "," and depth == 0:
                    parts.append(cur); cur = ""
                else:
                    cur += ch
            parts.append(cur)
            for i, a in enumerate(parts):
                a = a.strip()
                fp = re.fullmatch(r"void\s*\(\s*\*\s*([A-Za-z_0-9]+)\s*\)\s*\(\s*void\s*\*\s*\)", a)
                if fp:
                    params.append((fp.group(1), "Option<unsafe extern \"C\" fn(*mut c_void)>"))
                    continue
                mm = re.fullmatch(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)", a)
                ctype, pname = mm.group(1), mm.group(2)
                if not ctype.strip():  # unnamed parameter
                    ctype, pname = a, f"a{i}"
                if pname in RUST_KEYWORDS or pname == "out":
                    pname += "_"
                params.append((pname, rust_type(ctype)))
        rret = "" if ret == "void" else " -> " + rust_type(ret)
        fns.append((name, params, rret))
    return fns


def render(fns):
    lines = ['extern "C" {']
    for name, params, rret in fns:
        lines.append("    pub fn %s(%s)%s;" % (name, ", ".join(f"{n}: {t}" for n, t in params), rret))
    lines.append("}")
    return "\n".join(lines) + "\n"


def main():
    fns = parse(open(HDR).read())
    cur = open(OUT).read()
    head = cur[:cur.index('extern "C" {')]
    new = head + render(fns)
    if "--check" in sys.argv:
        sys.exit(0 if new == cur else 1)
    open(OUT, "w").write(new)
    print(f"{len(fns)} entry points written to {os.path.relpath(OUT, ROOT)}")


if __name__ == "__main__":
    main()
