"""integration/b200.rs cannot be compiled here (no rustc): check what can be checked textually - every `extern "C"` function
it declares exists in include/vors_b200.h with the same parameter count, and `VorsConfig` mirrors `vors_config` field for field."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RS = open(os.path.join(ROOT, "integration", "b200.rs")).read()
H = open(os.path.join(ROOT, "include", "vors_b200.h")).read()


def _c_functions():
    text = re.sub(r"/\*.*?\*/", "", H, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(vors_\w+)\s*\(([^;{}]*?)\)\s*;", text, re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_extern_block_matches_header():
    cfun = _c_functions()
    decls = re.findall(r"\bfn\s+(vors_\w+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", RS, re.S)
    assert len(decls) >= 10
    for name, args in decls:
        assert name in cfun, f"{name} is not declared in include/vors_b200.h"
        n = 0 if not args.strip() else len([a for a in args.split(",") if a.strip()])
        assert n == cfun[name], f"{name}: {n} parameters in b200.rs, {cfun[name]} in the header"


def test_config_struct_field_order():
    body = re.search(r"typedef struct vors_config \{(.*?)\} vors_config;", H, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = []
    for m in re.finditer(r"\b(uint32_t|int32_t|float)\s+([^;]+);", body):
        for nm in m.group(2).split(","):
            c_fields.append((nm.strip(), {"uint32_t": "u32", "int32_t": "i32", "float": "f32"}[m.group(1)]))
    rs_body = re.search(r"pub struct VorsConfig \{(.*?)\n\}", RS, re.S).group(1)
    rs_body = re.sub(r"//[^\n]*", "", rs_body)
    rs_fields = re.findall(r"pub\s+(\w+)\s*:\s*(\w+)", rs_body)
    assert rs_fields == c_fields


def test_shim_keeps_the_reference_signatures():
    # inverse_compositional.rs:74-80, 170-176, 243 - the three public entry points vors_track.rs calls
    assert "pub fn init(self, depth_ts: f64, depth_map: &DMatrix<u16>, img_ts: f64, img: DMatrix<u8>) -> Tracker" in RS
    assert "pub fn track(&mut self, depth_time: f64, depth_map: &DMatrix<u16>, img_time: f64, img: DMatrix<u8>)" in RS
    assert "pub fn current_frame(&self) -> (f64, Iso3)" in RS
