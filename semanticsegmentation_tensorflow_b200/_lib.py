"""ctypes binding of libsegk.so.  Prototypes are generated from `include/segk.h`, so the
Python side can never drift from the C ABI.  Loading fails loudly when the library is
missing: there is no CPU fallback on the product path."""
from __future__ import annotations

import ctypes
import os
import re

from .build import library_path

_HEADER = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "segk.h"))

SEGK_OK, SEGK_EINVAL, SEGK_ECUDA, SEGK_ENOMEM = 0, -1, -2, -3
EPI_RELU, EPI_OUT_F32 = 1, 2
DT_BF16, DT_F32, DT_U8 = 0, 1, 2

_SCALARS = {
    "int": ctypes.c_int, "unsigned": ctypes.c_uint, "float": ctypes.c_float, "double": ctypes.c_double,
    "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64, "size_t": ctypes.c_size_t,
}
_RET = dict(_SCALARS)
_RET["const char*"] = ctypes.c_char_p


class SegkError(RuntimeError):
    def __init__(self, code: int, what: str, msg: str):
        super().__init__(f"{what} failed with status {code}: {msg}")
        self.code = code


def parse_header(path: str = _HEADER):
    """-> {name: (restype_str, [(ctype_str, argname), ...])} for every function declared."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)
    decls = {}
    for m in re.finditer(r"(const char\*|int64_t|size_t|int)\s+(segk_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        parsed = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                parsed.append((mm.group(1).strip().replace(" *", "*"), mm.group(2)))
        decls[name] = (ret, parsed)
    return decls


def _ctype(t: str):
    if t == "const char*":
        return ctypes.c_char_p
    if t.endswith("*"):
        return ctypes.c_void_p
    return _SCALARS[t]


class Library:
    """Loaded libsegk.so with typed entry points; raises SegkError on non-zero status."""

    def __init__(self, path: str | None = None):
        path = path or library_path()
        if not os.path.exists(path):
            raise ImportError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "The segmentation ops have no CPU fallback.")
        self.path = path
        self.cdll = ctypes.CDLL(path)
        self.decls = parse_header()
        for name, (ret, args) in self.decls.items():
            fn = getattr(self.cdll, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = _RET[ret]
            fn.argtypes = [_ctype(t) for t, _ in args]
        if self.cdll.segk_abi_version() != 1:
            raise ImportError("libsegk.so ABI version mismatch")


_LIB = None


def lib() -> Library:
    global _LIB
    if _LIB is None:
        _LIB = Library()
    return _LIB


class Context:
    """One segk_ctx per device (thread-compatible, see include/segk.h)."""

    def __init__(self, device: int = 0):
        self.lib = lib()
        self.c = self.lib.cdll
        h = ctypes.c_void_p()
        rc = self.c.segk_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise SegkError(rc, "segk_create", "no sm_100 device %d (or CUDA unavailable)" % device)
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            self.c.segk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_error(self) -> str:
        return (self.c.segk_last_error(self.h) or b"").decode(errors="replace")

    def call(self, name: str, *args):
        rc = getattr(self.c, name)(self.h, *args)
        if rc != 0:
            raise SegkError(rc, name, self.last_error())

    def set_tuning(self, key: str, value: int):
        self.call("segk_set_tuning", key.encode(), int(value))

    @property
    def launches(self) -> int:
        return int(self.c.segk_launch_count(self.h))

    @property
    def sm_count(self) -> int:
        return int(self.c.segk_sm_count(self.h))
