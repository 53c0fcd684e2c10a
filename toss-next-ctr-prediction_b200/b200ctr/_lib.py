"""ctypes binding of libb200ctr.so (the C ABI declared in include/b200ctr.h).

The prototypes are parsed from the header itself, so the Python side cannot drift from the ABI and
`tests/test_abi.py` can check that every declared symbol is exported.  There is no fallback: if the
library is missing or fails to load, importing a compute op raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(os.path.dirname(_HERE))
HEADER = os.path.join(REPO_ROOT, "include", "b200ctr.h")
LIB_PATH = os.path.join(_HERE, "lib", "libb200ctr.so")
CSRC = os.path.join(os.path.dirname(_HERE), "csrc")


class Dropout(ctypes.Structure):
    _fields_ = [("p", ctypes.c_float), ("seed_lo", ctypes.c_uint32), ("seed_hi", ctypes.c_uint32),
                ("stream_id", ctypes.c_uint32)]


class Epilogue(ctypes.Structure):
    _fields_ = [("alpha", ctypes.c_float), ("bias", ctypes.c_void_p), ("act", ctypes.c_int),
                ("pre_out", ctypes.c_void_p), ("ld_pre", ctypes.c_int64), ("drop", Dropout),
                ("mask_mode", ctypes.c_int), ("mask_src", ctypes.c_void_p), ("ld_mask", ctypes.c_int64),
                ("mask_scale", ctypes.c_float), ("residual", ctypes.c_void_p), ("ld_res", ctypes.c_int64),
                ("accum", ctypes.c_void_p), ("ld_accum", ctypes.c_int64)]


NO_DROP = Dropout(0.0, 0, 0, 0)

_SCALARS = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "float": ctypes.c_float, "double": ctypes.c_double,
            "int32_t": ctypes.c_int32, "uint32_t": ctypes.c_uint32, "uint64_t": ctypes.c_uint64, "b200ctr_dropout_t": Dropout}
_RET = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "const char*": ctypes.c_char_p}


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object]]]:
    """Return {symbol: (restype, [argtypes])} for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)
    out: Dict[str, Tuple[object, List[object]]] = {}
    for m in re.finditer(r"(const char\*|int64_t|int)\s+(b200ctr_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes: List[object] = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.rsplit(" ", 1)[0].replace("const ", "").strip()
                    argtypes.append(_SCALARS[ty])
        out[name] = (_RET[ret], argtypes)
    return out


_lib = None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into b200ctr/lib/libb200ctr.so (nvcc cross-compiles without a GPU)."""
    import subprocess
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, capture_output=not verbose)
    r = subprocess.run(["make", "-C", CSRC, "-j", str(os.cpu_count() or 4)], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libb200ctr.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB_PATH


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(or `make -C toss-next-ctr-prediction_b200/csrc`). There is no CPU/PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (ret, argtypes) in parse_header().items():
        fn = getattr(lib, name)            # AttributeError here == header/library mismatch: fail loudly
        fn.restype = ret
        fn.argtypes = argtypes
    if lib.b200ctr_version() != 1:
        raise RuntimeError("libb200ctr ABI version mismatch")
    _lib = lib
    return lib


def last_error() -> str:
    return load().b200ctr_last_error().decode()


class B200CtrError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        raise B200CtrError(f"{what or 'libb200ctr'} failed with code {rc}: {last_error()}")
