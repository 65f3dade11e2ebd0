"""bindings/rust/src/ffi.rs against include/slzw.h: the Rust crate a salzweg maintainer would add
 cannot be compiled in this image (no cargo / rustc), so its raw declarations are
at least diffed against the header -- every function by name and arity, every constant by value."""
import os
import re

from tests.conftest import ROOT


def _header():
    with open(os.path.join(ROOT, "include", "slzw.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    fns = {}
    for m in re.finditer(r"SLZW_API\s+[^;]*?\b(slzw_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = " ".join(m.group(2).split())
        fns[m.group(1)] = 0 if args == "void" else len(args.split(","))
    consts = {k: int(v) for k, v in re.findall(r"(SLZW_(?:OK|ERR_\w+|RC_\w+))\s*=\s*(-?\d+)", text)}
    consts.update({k: int(v) for k, v in re.findall(r"#define\s+(SLZW_(?:FLAVOUR|PREDICTOR)_\w+)\s+(\d+)", text)})
    return fns, consts


def _rust():
    with open(os.path.join(ROOT, "bindings", "rust", "src", "ffi.rs")) as f:
        text = f.read()
    block = text[text.index('extern "C" {'):]
    fns = {}
    for m in re.finditer(r"pub fn (slzw_\w+)\(([^)]*)\)", block):
        args = m.group(2).strip()
        fns[m.group(1)] = 0 if not args else len(args.split(","))
    consts = {k: int(v) for k, v in re.findall(r"pub const (SLZW_\w+): \w+ = (-?\d+);", text)}
    return fns, consts


def test_rust_ffi_declares_the_whole_header():
    h_fns, h_consts = _header()
    r_fns, r_consts = _rust()
    assert len(h_fns) >= 30
    assert sorted(r_fns) == sorted(h_fns)
    assert r_fns == h_fns  # arities
    assert r_consts == h_consts


def test_rust_facade_uses_only_declared_functions():
    with open(os.path.join(ROOT, "bindings", "rust", "src", "lib.rs")) as f:
        text = f.read()
    used = set(re.findall(r"\b(slzw_\w+)\s*\(", text))
    assert used and used <= set(_rust()[0])
