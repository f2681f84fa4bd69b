"""ctypes binding of libdmayolo.so, generated from include/dmayolo.h.

The header is the single source of truth for the C-ABI: every `typedef struct dmay_*_params`
is parsed into a ctypes.Structure and every `int dmay_*(const X* p, dmay_stream_t stream)`
prototype into a bound function.  There is no fallback: if the shared library is missing
(or fails to load) `lib()` raises, so a GPU path can never silently run on something else.
"""
from __future__ import annotations

import ctypes
import re
from pathlib import Path

PKG = Path(__file__).resolve().parent
HEADER = PKG.parent / "include" / "dmayolo.h"
SO = Path(__import__("os").environ.get("DMAY_SO", PKG / "libdmayolo.so"))  # DMAY_SO: A/B a second build

_CTYPES = {
    "const void*": ctypes.c_void_p,
    "void*": ctypes.c_void_p,
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
}


class DmayError(RuntimeError):
    pass


def parse_header(text: str | None = None):
    """-> (structs: {name: [(field, ctype)]}, funcs: {name: (restype, [argtype names])}, consts)"""
    if text is None:
        text = HEADER.read_text()
    text_nc = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    structs = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} (\w+);", text_nc, flags=re.S):
        name, body = m.group(3), m.group(2)
        fields = []
        for line in body.strip().splitlines():
            line = line.strip().rstrip(";")
            if not line:
                continue
            fm = re.match(r"(const void\*|void\*|long long|int|float|double)\s+(\w+)$", line)
            if not fm:
                raise ValueError(f"unparseable field in {name}: {line!r}")
            fields.append((fm.group(2), _CTYPES[fm.group(1)]))
        structs[name] = fields
    funcs = {}
    for m in re.finditer(r"^(int|long long|const char\*)\s+(dmay_\w+)\((.*?)\);", text_nc, flags=re.M | re.S):
        funcs[m.group(2)] = (m.group(1), [a.strip() for a in m.group(3).split(",")])
    consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define (DMAY_\w+) (-?\d+)", text_nc)}
    return structs, funcs, consts


_STRUCTS, _FUNCS, CONSTS = parse_header()


def _make_struct(name, fields):
    return type(name, (ctypes.Structure,), {"_fields_": fields})


STRUCTS = {n: _make_struct(n, f) for n, f in _STRUCTS.items()}
_FIELD_NAMES = {n: {k for k, _ in f} for n, f in _STRUCTS.items()}
_lib = None


def exported_symbols():
    return sorted(_FUNCS)


def lib():
    """Load libdmayolo.so (once).  Raises DmayError when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not SO.exists():
        raise DmayError(
            f"{SO} is missing: the sm_100a extension is the only compute path (no CPU/eager fallback). "
            "Build it with `python -m dma_yolo_b200.build` (or __graft_entry__.build()).")
    try:
        L = ctypes.CDLL(str(SO))
    except OSError as e:  # pragma: no cover
        raise DmayError(f"cannot load {SO}: {e}") from e
    for fname, (rtype, args) in _FUNCS.items():
        fn = getattr(L, fname)  # AttributeError here == header/library mismatch
        fn.restype = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "const char*": ctypes.c_char_p}[rtype]
        argtypes = []
        for a in args:
            if a in ("void", ""):
                continue
            if a.startswith("const dmay_") and "*" in a:
                sname = a.split()[1].rstrip("*")
                argtypes.append(ctypes.POINTER(STRUCTS[sname]))
            elif a.startswith("dmay_stream_t"):
                argtypes.append(ctypes.c_void_p)
            elif a.startswith("const void*") or a.startswith("void*"):
                argtypes.append(ctypes.c_void_p)
            elif a.startswith("long long"):
                argtypes.append(ctypes.c_longlong)
            elif a.startswith("int"):
                argtypes.append(ctypes.c_int)
            else:
                raise ValueError(f"unparseable argument {a!r} of {fname}")
        fn.argtypes = argtypes
    _lib = L
    return L


def strerror(code: int) -> str:
    return lib().dmay_strerror(int(code)).decode()


def check(code: int, what: str):
    if code != 0:
        raise DmayError(f"{what} failed: {code} ({strerror(code)})")


def call(fname: str, stream: int, **fields):
    """Fill the params struct of `fname` from keyword arguments and launch on `stream`."""
    L = lib()
    sname = _FUNCS[fname][1][0].split()[1].rstrip("*")
    st = STRUCTS[sname]()
    names = _FIELD_NAMES[sname]
    for k, v in fields.items():
        if k not in names:
            raise AttributeError(f"{sname} has no field {k!r}")
        setattr(st, k, v)
    check(getattr(L, fname)(ctypes.byref(st), ctypes.c_void_p(stream)), fname)


def launch_count() -> int:
    return int(lib().dmay_launch_count())
