"""CPU: the C-ABI library loads and exports exactly the symbols include/tokamak_b200.h declares; the
Python binding covers all of them; the product path fails loudly without a GPU and never touches oracle/."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tokamak_b200.h")
PKG = os.path.join(ROOT, "tokamak-zk-evm_b200")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tkm_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    p = os.path.join(PKG, "lib", "libtokamak_b200.so")
    if not os.path.exists(p):
        subprocess.check_call(["make", "-s", "-j8", "-C", PKG])
    return p


def test_header_symbols_exported(lib_path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = set(re.findall(r" T (tkm_[a-z0-9_]+)", out))
    syms = header_symbols()
    assert len(syms) > 50
    missing = [s for s in syms if s not in exported]
    assert not missing, f"declared in header but not exported: {missing}"
    extra = [s for s in exported if s not in syms]
    assert not extra, f"exported but not declared in header: {extra}"


def test_python_binding_covers_header(lib_path):
    from tokamak_b200 import ffi

    lib = ffi.load()
    bound = set(ffi.SIGNATURES) | set(ffi.STRING_FUNCS)
    assert bound == set(header_symbols())
    assert b"sm_100a" in lib.tkm_version()


def test_sass_is_sm100a_only(lib_path):
    out = subprocess.check_output(["cuobjdump", "--list-elf", lib_path], text=True)
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_fails_loudly_without_gpu(lib_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import tokamak_b200 as T

    with pytest.raises(T.TkmError) as e:
        T.Context(0)
    assert e.value.status == -5 and "no CPU fallback" in str(e.value)


def test_product_does_not_reference_oracle():
    bad = []
    for dirpath, _, files in os.walk(PKG):
        if os.path.basename(dirpath) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".rs", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"(import\s+(pyref|oracle_ffi|oracle_backend)|from\s+(pyref|oracle_ffi|oracle_backend)|liboracle|oracle/)", txt):
                    # comments that merely state the rule are allowed only in ffi.py's docstring
                    if not (f == "ffi.py" and "nothing here imports oracle/" in txt and txt.count("oracle") == 1):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_rust_sys_crate_declares_every_header_symbol():
    """The -sys crate sources (host/rust/tokamak-b200-sys) are not compiled here (no Rust toolchain), so at least keep
    them in step with the header: one `pub fn` per exported prototype."""
    hdr = open(os.path.join(ROOT, "include", "tokamak_b200.h")).read()
    rs = open(os.path.join(PKG, "host", "rust", "tokamak-b200-sys", "src", "lib.rs")).read()
    syms = set(re.findall(r"\b(tkm_[a-z0-9_]+)\s*\(", hdr))
    have = set(re.findall(r"pub fn (tkm_[a-z0-9_]+)", rs))
    assert not (syms - have), sorted(syms - have)
