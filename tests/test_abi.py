"""CPU-side checks of the C ABI: the shared library loads, exports every symbol include/b200ctr.h declares, and the
ctypes prototypes parsed from the header are sane.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

from b200ctr import _lib


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    assert _lib.LIB_PATH.startswith(_lib.REPO_ROOT)


def test_every_declared_symbol_is_exported():
    decl = _lib.parse_header()
    assert len(decl) >= 40
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\sT\s+(b200ctr_\w+)", out))
    missing = set(decl) - exported
    assert not missing, f"declared in include/b200ctr.h but not exported: {sorted(missing)}"
    extra = exported - set(decl)
    assert not extra, f"exported but undeclared: {sorted(extra)}"


def test_loader_binds_prototypes_and_version():
    lib = _lib.load()
    assert lib.b200ctr_version() == 1
    assert isinstance(lib.b200ctr_last_error(), bytes)
    proto = _lib.parse_header()
    ret, args = proto["b200ctr_gemm"]
    assert ret is ctypes.c_int and len(args) == 15
    assert args[0] is ctypes.c_int and args[2] is ctypes.c_int64 and args[5] is ctypes.c_void_p
    ret, args = proto["b200ctr_cat_fwd"]
    assert _lib.Dropout in args
    assert proto["b200ctr_gemm_workspace_bytes"][0] is ctypes.c_int64


def test_struct_layout_matches_header_field_order():
    src = open(_lib.HEADER).read()
    body = re.search(r"typedef struct \{([^{}]*)\} b200ctr_epilogue_t;", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [re.split(r"[\s\*]+", f.strip())[-1] for f in body.split(";") if f.strip()]
    assert names == [f[0] for f in _lib.Epilogue._fields_]
    body = re.search(r"typedef struct \{([^{}]*)\} b200ctr_dropout_t;", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [re.split(r"[\s\*]+", f.strip())[-1] for f in body.split(";") if f.strip()]
    assert names == [f[0] for f in _lib.Dropout._fields_]


def test_argument_validation_without_a_gpu():
    """Entry points validate arguments before touching the device, so bad calls fail cleanly even here."""
    lib = _lib.load()
    rc = lib.b200ctr_gemm(0, 0, 4, 4, 4, None, 4, None, 4, None, 4, None, 1, None, None)
    assert rc == -1 and b"null" in lib.b200ctr_last_error()
    rc = lib.b200ctr_retrieve_fwd(None, 1, None, None, None, None, 1, 8, 4, 64, 0, 10, None, None, None, None, None)
    assert rc == -1
    assert lib.b200ctr_gemm_workspace_bytes(10, 20, 30, 4) == 4 * 10 * 20 * 4
    assert lib.b200ctr_gemm_workspace_bytes(10, 20, 30, 1) == 0


def test_sass_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
