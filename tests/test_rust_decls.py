"""The Rust `-sys` crate (rust/r4w-b200-sys/src/lib.rs) cannot be compiled in this image (no rustc), so its agreement with
the C header is checked textually: the same entry points on both sides, the same number of arguments, scalar types that
correspond, and `#[repr(C)]` structs whose field lists give the sizes of the C structs (through the ctypes mirrors)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "r4w_b200.h")
RUST = os.path.join(ROOT, "rust", "r4w-b200-sys", "src", "lib.rs")


def _strip_c(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def _split_args(s):
    s = s.strip()
    return [] if s in ("", "void") else [a.strip() for a in s.split(",")]


def c_functions():
    text = _strip_c(open(HEADER).read())
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w \*]*?)\b(r4wb_\w+)\s*\(([^;{}]*?)\)\s*;", text):
        out[m.group(2)] = (m.group(1).strip(), _split_args(m.group(3)))
    return out


def rust_functions():
    text = re.sub(r"//[^\n]*", " ", open(RUST).read())
    body = text[text.index('extern "C" {'):]
    out = {}
    for m in re.finditer(r"pub fn (r4wb_\w+)\s*\(([^)]*)\)\s*(?:->\s*([^;]+))?;", body):
        out[m.group(1)] = ((m.group(3) or "()").strip(), _split_args(m.group(2)))
    return out


C2RUST = {"uint64_t": "u64", "uint32_t": "u32", "uint8_t": "u8", "int8_t": "i8", "double": "f64", "int": "c_int", "size_t": "usize",
          "r4wb_error": "c_int", "r4wb_fmt": "c_int", "r4wb_mem": "c_int", "void": "c_void", "char": "c_char"}


def _rust_type_of_c(decl, is_return=False):
    """'const r4wb_pcps* h' -> '*const r4wb_pcps'"""
    d = decl.strip()
    if not is_return:
        d = re.sub(r"\b[A-Za-z_]\w*$", "", d).strip() if not d.endswith("*") else d      # drop the parameter name
    stars = d.count("*")
    const = "const" in d.split("*")[0].split()
    base = [w for w in d.replace("*", " ").split() if w != "const"][0]
    t = C2RUST.get(base, base)
    for i in range(stars):
        t = ("*const " if (const and i == 0) else "*mut ") + t
    return t


def test_same_entry_points_and_signatures():
    cf, rf = c_functions(), rust_functions()
    assert len(cf) >= 50
    assert set(cf) == set(rf), (sorted(set(cf) - set(rf)), sorted(set(rf) - set(cf)))
    for name, (cret, cargs) in cf.items():
        rret, rargs = rf[name]
        assert len(cargs) == len(rargs), name
        for ca, ra in zip(cargs, rargs):
            want = _rust_type_of_c(ca)
            got = ra.split(":", 1)[1].strip()
            assert got == want, (name, ca, ra, want)
        want_ret = "()" if cret == "void" else _rust_type_of_c(cret, is_return=True)
        assert rret == want_ret, (name, cret, rret)


def test_every_declared_symbol_is_exported():
    from r4w_b200 import _lib
    lib = C.CDLL(_lib.LIB_PATH) if hasattr(_lib, "LIB_PATH") else _lib.lib()
    for name in rust_functions():
        assert hasattr(lib, name), name


RUST_SIZES = {"u8": (1, 1), "u32": (4, 4), "u64": (8, 8), "f64": (8, 8)}


def rust_struct_layout(name, text, seen=None):
    """(size, align) of a #[repr(C)] struct from its field list, C layout rules"""
    m = re.search(r"pub struct %s\s*\{(.*?)\n\}" % re.escape(name), text, flags=re.S)
    assert m, name
    off, align = 0, 1
    for f in re.finditer(r"pub\s+\w+\s*:\s*([^,\n]+),", m.group(1)):
        t = f.group(1).strip()
        arr = re.match(r"\[(\w+);\s*(\d+)\]", t)
        count = int(arr.group(2)) if arr else 1
        base = arr.group(1) if arr else t
        if base.startswith("*"):
            sz, al = 8, 8
        elif base in RUST_SIZES:
            sz, al = RUST_SIZES[base]
        else:
            sz, al = rust_struct_layout(base, text)
        off = (off + al - 1) // al * al + sz * count
        align = max(align, al)
    return (off + align - 1) // align * align, align


def test_struct_sizes_match_the_c_structs():
    from r4w_b200 import config as cfgmod, _lib, acquisition, tracking
    text = re.sub(r"//[^\n]*", " ", open(RUST).read())
    mirrors = {}
    for mod in (cfgmod, _lib, acquisition, tracking):
        for v in vars(mod).values():
            if isinstance(v, type) and issubclass(v, C.Structure) and v is not C.Structure:
                mirrors[C.sizeof(v)] = v
    # sizes of the C structs, computed by the compiler through a tiny probe would need gcc; the ctypes mirrors are already
    # checked against the header by tests/test_cabi.py, so compare with those
    want = {"r4wb_lla": 24, "r4wb_sat_cfg": 96, "r4wb_output_cfg": 48, "r4wb_acq_result": 48, "r4wb_track_cfg": 64, "r4wb_track_state": 64,
            "r4wb_sat_status": 88}
    for name, size in want.items():
        assert rust_struct_layout(name, text)[0] == size, (name, rust_struct_layout(name, text))
        assert size in mirrors, (name, size, sorted(mirrors))
    big = {n: rust_struct_layout(n, text)[0] for n in ("r4wb_receiver_cfg", "r4wb_environment_cfg", "r4wb_scenario_cfg")}
    for n, size in big.items():
        assert size in mirrors, (n, size, sorted(mirrors))


def test_struct_sizes_match_gcc(tmp_path):
    """sizeof of every struct of the header as gcc lays it out == the #[repr(C)] layout of the Rust declaration"""
    import subprocess
    names = ["r4wb_lla", "r4wb_sat_cfg", "r4wb_receiver_cfg", "r4wb_environment_cfg", "r4wb_output_cfg", "r4wb_scenario_cfg",
             "r4wb_sat_status", "r4wb_acq_result", "r4wb_track_cfg", "r4wb_track_state"]
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include "r4w_b200.h"\nint main(void){' +
                   "".join('printf("%s %%zu\\n", sizeof(%s));' % (n, n) for n in names) + "return 0;}\n")
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    text = re.sub(r"//[^\n]*", " ", open(RUST).read())
    for n in names:
        assert rust_struct_layout(n, text)[0] == int(got[n]), (n, got[n])
