"""The Rust binding (rust/blu-consensus-sys) cannot be compiled in this image (no cargo / rustc), so its shape is checked by
hand against the C header: the #[repr(C)] structs must list the header's fields in the header's order with matching
widths, and every `extern "C"` function it declares must be a symbol the library exports."""
import os
import re

from blutils_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RS = open(os.path.join(ROOT, "rust", "blu-consensus-sys", "src", "lib.rs")).read()
HDR = open(os.path.join(ROOT, "include", "blu_consensus.h")).read()

C2RUST = {"int32_t": "i32", "uint32_t": "u32", "uint64_t": "u64", "int64_t": "i64", "double": "f64", "uint8_t": "u8", "int8_t": "i8"}


def c_struct_fields(name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), HDR, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        m = re.match(r"(\w+)\s+(\w+)(?:\[(\d+)\])?$", decl)
        assert m, decl
        ty, field, n = m.groups()
        out.append((field, C2RUST[ty] if n is None else "[%s; %s]" % (C2RUST[ty], n)))
    return out


def rust_struct_fields(name):
    body = re.search(r"pub struct %s \{(.*?)\n\}" % name, RS, re.S).group(1)
    out = []
    for line in body.splitlines():
        line = line.split("//")[0].strip().rstrip(",")
        if not line:
            continue
        m = re.match(r"pub (\w+): (.+)$", line)
        assert m, line
        out.append((m.group(1), m.group(2).strip()))
    return out


def test_blu_opts_layout_matches_header():
    assert rust_struct_fields("blu_opts") == c_struct_fields("blu_opts")
    # ... and the ctypes mirror the tests themselves use
    assert [f[0] for f in _ffi.blu_opts._fields_] == [f[0] for f in c_struct_fields("blu_opts")]


def test_extern_functions_exist_in_the_library():
    block = re.search(r'extern "C" \{(.*?)\n\}', RS, re.S).group(1)
    names = re.findall(r"pub fn (\w+)\(", block)
    assert len(names) >= 15
    exported = {n for n, _, _ in _ffi.SYMBOLS}
    assert set(names) <= exported, set(names) - exported
    lib = _ffi.lib()
    for n in names:
        assert hasattr(lib, n)
