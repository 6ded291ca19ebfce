"""ctypes binding of libdram_b200.so — signatures are generated from include/dram_b200.h, so the binding cannot
drift from the C ABI.  There is NO fallback: if the library is missing or a call fails, we raise."""
import ctypes
import os
import re

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(PKG_DIR, "..", "include", "dram_b200.h")
LIB_PATH = os.path.join(PKG_DIR, "libdram_b200.so")

_CTYPES = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float, "double": ctypes.c_double,
           "size_t": ctypes.c_size_t}


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)
    protos = {}
    for m in re.finditer(r"\b(const\s+char\s*\*|int|size_t)\s+(dram_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if "char" in ret else _CTYPES[ret.strip()]
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    t = re.sub(r"\bconst\b", "", a).strip()
                    t = re.sub(r"\s+\w+$", "", t).strip()          # drop the parameter name
                    argtypes.append(_CTYPES[t])
        protos[name] = (restype, argtypes)
    return protos


class DramLibraryError(RuntimeError):
    pass


_LIB = None


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise DramLibraryError(
            f"{LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the DRAM hot path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in parse_header().items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _LIB = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().dram_last_error().decode("utf-8", "replace")
        raise DramLibraryError(f"{what} failed (code {rc}): {msg}")
