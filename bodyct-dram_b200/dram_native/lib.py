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
    for m in re.finditer(r"\b(const\s+char\s*\*|long\s+long|int|size_t)\s+(dram_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if "char" in ret else _CTYPES[" ".join(ret.split())]
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


# kernels launched per C-ABI call (for the bench's `gpu_launches` claim); memsets are not counted
KERNELS_PER_CALL = {"dram_peer_mailbox_bytes": 0, "dram_peer_max_doubles": 0, "dram_peer_max_ranks": 0, "dram_peer_alloc": 0,
                    "dram_peer_free": 0, "dram_peer_export": 0, "dram_peer_open": 0, "dram_peer_close": 0,
                    "dram_conv3d_umma_fwd_stat_rows": 0, "dram_conv3d_umma_fwd_kernel": 0, "dram_conv3d_umma_wgrad_kernel": 0, "dram_bn_stats_from_partials_workspace_bytes": 0,
                    "dram_bn_stats_from_partials": 2, "dram_version": 0, "dram_sm_arch": 0, "dram_last_error": 0, "dram_device_check": 0,
                    "dram_pcm_num_offsets": 0, "dram_pcm_qk_floats": 0, "dram_labelled_sum_workspace_bytes": 0, "dram_labelled_sum": 2, "dram_pcm_bwd_ws_floats": 0, "dram_conv3d_umma_wgrad_workspace_bytes": 0,
                    "dram_upsample2x_concat_fwd": 2, "dram_upsample2x_concat_planes": 2, "dram_upsample2x_concat_bwd": 1, "dram_conv3d_umma_wgrad": 2,
                    "dram_pcm_fwd": 2, "dram_pcm_bwd": 3, "dram_pointwise8_planes_supported": 0,
                    "dram_pointwise8_planes_wgrad_workspace_bytes": 0, "dram_pointwise8_planes_wgrad": 2, "dram_label_bboxes": 2}


class Profile:
    """Launch accounting.  `enabled` additionally brackets every call with CUDA events on the launching stream."""

    def __init__(self):
        self.enabled = False
        self.reset()

    def reset(self):
        self.launches = 0
        self.calls = {}
        self.records = []          # (name, note, start_event, end_event)
        self.pending_note = None

    def note(self, **kw):
        """Attach algorithmic flops / bytes to the NEXT call (consumed by it)."""
        self.pending_note = kw

    def summary(self, by_tag=False):
        """-> {name: {"calls", "ms", "flops", "bytes"}} after a device synchronize (by_tag: key = name + layer tag)."""
        out = {}
        for name, note, s, e in self.records:
            key = name + ":" + str(note.get("tag")) if (by_tag and note and note.get("tag")) else name
            d = out.setdefault(key, {"calls": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["calls"] += 1
            d["ms"] += s.elapsed_time(e)
            if note:
                d["flops"] += float(note.get("flops", 0.0))
                d["bytes"] += float(note.get("bytes", 0.0))
        return out


PROFILE = Profile()


def _instrument(name, fn):
    nk = KERNELS_PER_CALL.get(name, 1)
    if nk == 0:
        return fn

    def call(*args):
        P = PROFILE
        P.launches += nk
        P.calls[name] = P.calls.get(name, 0) + 1
        if P.enabled:
            import torch
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            note, P.pending_note = P.pending_note, None
            s.record()
            rc = fn(*args)
            e.record()
            P.records.append((name, note, s, e))
            return rc
        P.pending_note = None
        return fn(*args)

    return call


class _Lib:
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
    cdll = ctypes.CDLL(LIB_PATH)
    lib = _Lib()
    for name, (restype, argtypes) in parse_header().items():
        fn = getattr(cdll, name)         # AttributeError if the .so lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
        setattr(lib, name, _instrument(name, fn))
    lib.cdll = cdll
    _LIB = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().dram_last_error().decode("utf-8", "replace")
        raise DramLibraryError(f"{what} failed (code {rc}): {msg}")
