"""The C-ABI boundary: libdhqr.so loads without a GPU and exports exactly what include/dhqr.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dhqr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dhqr_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    import dhqr_b200 as D
    assert _declared() == sorted(D._lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    import dhqr_b200 as D
    lib = D._lib.load()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.dhqr_version() == 100


def test_no_oracle_or_cpu_fallback_in_product():
    # the product path must never import / link the oracle
    pkg = os.path.join(ROOT, "distributedhouseholderqr.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                bad = re.findall(r"^\s*(?:import|from)\s+\S*oracle|#\s*include[^\n]*oracle|dlopen\([^\n]*oracle|libdhqr_oracle", text, flags=re.M)
                assert not bad, (os.path.join(dirpath, f), bad)
    import subprocess
    out = subprocess.run(["ldd", os.path.join(pkg, "libdhqr.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_fails_loudly_without_gpu():
    import torch
    import dhqr_b200 as D
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(D._lib.DhqrError) as e:
        D.Handle(0)
    assert "cuda" in str(e.value).lower()


def test_argument_errors_are_lapack_style():
    import dhqr_b200 as D
    lib = D._lib.load()
    # null handle -> -1 on every entry point that takes one
    assert lib.dhqr_qr_f64(None, 4, 2, 0, 2, None, 4, None, 0, None) == -1
    assert lib.dhqr_solve_f64(None, 4, 2, 0, 2, None, 4, None, None, 4, 1, None) == -1
    assert lib.dhqr_set_option(None, b"nb", 64) == -1
    assert b"null handle" in lib.dhqr_last_error()
    assert lib.dhqr_create(None, 0) == -1


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/dhqr.h must be consumable by a C compiler (cgo / ccall / ctypes users never see C++), and every declared
    function must resolve against libdhqr.so at link time."""
    import re, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "dhqr.h")).read()
    names = sorted(set(re.findall(r"^(?:int|const char \*)\s*(dhqr_[a-z0-9_]+)\s*\(", hdr, flags=re.M)))
    assert len(names) >= 20
    src = tmp_path / "use.c"
    body = "\n".join(f"    p[{i}] = (fn_t)&{n};" for i, n in enumerate(names))
    src.write_text('#include "dhqr.h"\n#include <stdio.h>\ntypedef void (*fn_t)(void);\nint main(void) {\n    fn_t p[%d];\n%s\n'
                   '    printf("%%d %%d\\n", dhqr_version(), (int)(sizeof(p) / sizeof(p[0])));\n    return p[0] == 0;\n}\n' % (len(names), body))
    exe = tmp_path / "use"
    libdir = os.path.join(root, "distributedhouseholderqr.jl_b200")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-l:libdhqr.so", f"-Wl,-rpath,{libdir}", "-Wl,--allow-shlib-undefined"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    if out.returncode == 0:                      # loads without a GPU as long as libcudart resolves; no CUDA call is made
        ver, cnt = out.stdout.split()
        assert int(ver) == 100 and int(cnt) == len(names)
