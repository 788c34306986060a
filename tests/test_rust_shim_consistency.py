"""bindings/rust: the crate cannot be compiled here (no Rust toolchain), so its `extern "C"` block
is checked against include/diffsol_b200.h textually: the same set of entry points, the same number of arguments each, the
same field order in `dsb_options`, and every symbol present in the built library."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def c_prototypes():
    text = open(os.path.join(ROOT, "include", "diffsol_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"^(?:int|int64_t|void|const char\*)\s+(dsb_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.M | re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return protos, text


def rust_prototypes():
    text = open(os.path.join(ROOT, "bindings", "rust", "src", "ffi.rs")).read()
    protos = {}
    for m in re.finditer(r"pub fn (dsb_\w+)\s*\(([^;]*?)\)\s*(?:->\s*[^;]+)?;", text, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    return protos, text


def test_every_entry_point_is_declared_with_the_same_arity():
    c, _ = c_prototypes()
    r, _ = rust_prototypes()
    assert len(c) >= 50
    assert set(c) == set(r), (sorted(set(c) - set(r)), sorted(set(r) - set(c)))
    assert {k: v for k, v in c.items() if r[k] != v} == {}


def test_options_struct_has_the_same_field_order():
    _, ctext = c_prototypes()
    _, rtext = rust_prototypes()
    cbody = re.search(r"typedef struct dsb_options \{(.*?)\} dsb_options;", ctext, flags=re.S).group(1)
    cfields = re.findall(r"(int32_t|double)\s+(\w+)\s*;", cbody)
    rbody = re.search(r"pub struct dsb_options \{(.*?)\n\}", rtext, flags=re.S).group(1)
    rfields = re.findall(r"pub (\w+): (i32|f64)", rbody)
    assert [(n, {"int32_t": "i32", "double": "f64"}[t]) for t, n in cfields] == rfields and len(rfields) == 22


def test_every_declared_symbol_is_exported_by_the_library():
    from diffsol_b200 import build
    lib = ctypes.CDLL(build.build())
    r, _ = rust_prototypes()
    for name in r:
        assert hasattr(lib, name), name
