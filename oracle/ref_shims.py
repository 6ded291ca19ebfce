"""Load the UNMODIFIED reference modules (`/root/reference/dram/{parts,models,metrics,utils}.py`) on CPU.

TEST INFRASTRUCTURE ONLY.  Nothing under `bodyct-dram_b200/` may import this file.  It is used
  * by `tests/golden/make_golden.py` to generate the committed golden vectors, and
  * by the `-m "not gpu"` tests, when `/root/reference` exists, to pin `oracle/dram_oracle.py`
    (our own CPU restatement) against the real reference code.
`/root/reference` does not exist on the GPU box, so no `-m gpu` test, `smoke()` or `bench.py` leg reads it.

What is shimmed:
  * empty stub modules for SimpleITK / skimage (imported at module top by utils.py:7-8, data_transforms.py:5,
    never called on the model/loss path);
  * a ~60-line fake `dgl` implementing the degree-bucketed UDF semantics that `models.PCM` relies on
    (`DGLGraph(nx_graph)`, `.to`, `.ndata`, `update_all(message_func, reduce_func)`,
    `dgl.transform.remove_self_loop`); DGL itself is not vendored and is unpinned in the reference
    (Dockerfile:130-137 builds master; README.md:12 says 0.6.x) -> "parity unpinned" for DGL internals,
    anchored on the reference's own call sites models.py:256-258,330-349;
  * `torch.Tensor.cuda` -> identity while the loss runs (metrics.py:136,173 hard-code `.cuda()`).
The reference modules are registered under private names (`ref_models`, ...) so they never collide with the
same-named drop-in modules of this repo.
"""
import importlib
import os
import sys
import types
from collections import namedtuple

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("DRAM_REFERENCE_ROOT", "/root/reference/dram")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models.py"))


# ------------------------------------------------------------------------------------------------ fake DGL
class _Batch:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _FakeGraph:
    """Edge-list graph with DGL-0.6 style `update_all(msg, reduce)` (in-degree bucketing)."""

    def __init__(self, nx_graph=None, src=None, dst=None, n=None):
        if nx_graph is not None:
            e = np.asarray(list(nx_graph.edges()), dtype=np.int64).reshape(-1, 2)
            src, dst, n = torch.from_numpy(e[:, 0].copy()), torch.from_numpy(e[:, 1].copy()), nx_graph.number_of_nodes()
        self.src, self.dst, self.n = src, dst, n
        self.ndata = {}

    def to(self, device):
        return self

    def number_of_nodes(self):
        return self.n

    def number_of_edges(self):
        return int(self.src.numel())

    def update_all(self, message_func, reduce_func):
        order = torch.argsort(self.dst, stable=True)
        src_s = self.src[order]
        deg = torch.bincount(self.dst, minlength=self.n)
        start = torch.cumsum(deg, 0) - deg
        msgs = message_func(_Batch(src={k: v[src_s] for k, v in self.ndata.items()}))
        results = {}
        for d in torch.unique(deg).tolist():
            if d == 0:
                continue
            nodes = torch.nonzero(deg == d).flatten()
            eidx = start[nodes][:, None] + torch.arange(d)[None, :]
            mailbox = {k: v[eidx.reshape(-1)].view(len(nodes), d, *v.shape[1:]) for k, v in msgs.items()}
            out = reduce_func(_Batch(data={k: v[nodes] for k, v in self.ndata.items()}, mailbox=mailbox))
            for k, v in out.items():
                if k not in results:
                    results[k] = torch.zeros(self.n, *v.shape[1:], dtype=v.dtype)
                results[k] = results[k].index_put((nodes,), v)
        self.ndata.update(results)


def _remove_self_loop(g):
    keep = g.src != g.dst
    return _FakeGraph(src=g.src[keep], dst=g.dst[keep], n=g.n)


def _install_stubs():
    for n in ["SimpleITK", "skimage", "skimage.filters", "skimage.filters.thresholding", "skimage.exposure"]:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules["skimage"].filters = sys.modules["skimage.filters"]
    sys.modules["skimage"].exposure = sys.modules["skimage.exposure"]
    sys.modules["skimage.filters"].thresholding = sys.modules["skimage.filters.thresholding"]
    for a in ["sitkNearestNeighbor", "sitkLinear", "sitkGaussian", "sitkLabelGaussian", "sitkBSpline",
              "sitkHammingWindowedSinc", "sitkCosineWindowedSinc", "sitkWelchWindowedSinc",
              "sitkLanczosWindowedSinc"]:
        setattr(sys.modules["SimpleITK"], a, 0)
    dgl = types.ModuleType("dgl")
    dgl.DGLGraph = _FakeGraph
    dgl.transform = types.ModuleType("dgl.transform")
    dgl.transform.remove_self_loop = _remove_self_loop
    sys.modules["dgl"] = dgl
    sys.modules["dgl.transform"] = dgl.transform


Reference = namedtuple("Reference", "models parts metrics utils Settings root")
_CACHE = {}
_FLAT = ["parts", "utils", "models", "metrics", "data_transforms"]


def load_reference():
    """Import the reference's flat modules under private names and return them."""
    if "ref" in _CACHE:
        return _CACHE["ref"]
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    if not hasattr(np, "bool"):          # utils.py:237 uses np.bool (numpy<1.24); only on degenerate branch
        np.bool = bool
    _install_stubs()
    saved = {k: sys.modules.pop(k) for k in _FLAT if k in sys.modules}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mods = {k: importlib.import_module(k) for k in ["parts", "utils", "models", "metrics"]}
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in _FLAT:
            m = sys.modules.pop(k, None)
            if m is not None:
                sys.modules["ref_" + k] = m
        sys.modules.update(saved)
    # models.py:520-546 dumps debug tiles through OpenCV whenever the loss has set `trace_path` (always, metrics.py:202);
    # the drawing helper is reporting code outside the hot path -> no-op.
    mods["models"].draw_mask_tile_singleview_heatmap = lambda *a, **k: None
    ref = Reference(mods["models"], mods["parts"], mods["metrics"], mods["utils"], mods["utils"].Settings,
                    REFERENCE_ROOT)
    _CACHE["ref"] = ref
    return ref


class cpu_cuda_identity:
    """Context manager: make `.cuda()` a no-op so metrics.py:136,173 run on CPU."""

    def __enter__(self):
        self._orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self_, *a, **k: self_
        return self

    def __exit__(self, *exc):
        torch.Tensor.cuda = self._orig
        return False


def load_settings(name):
    """`Settings` object of one of the reference's exp_settings files (st_dram_ref.py / st_dram_ref_att.py)."""
    ref = load_reference()
    return ref.Settings(os.path.join(ref.root, "exp_settings", name))


def build_reference_model(model_cfg, seed=0):
    """Construct + HeNorm-initialise a reference model exactly as job_runner.py:360-382 does."""
    ref = load_reference()
    cfg = dict(model_cfg)
    method = cfg.pop("method")
    cls = getattr(ref.models, method.split(".")[-1])
    torch.manual_seed(seed)
    m = cls(**cfg)
    m.init(ref.models.HeNorm(mode="fan_in"))
    return m


class LossHost:
    """Stand-in for the job runner `obj` the loss reads (metrics.py:172,198-199)."""

    def __init__(self, freq=None):
        self.ctss_frequency_map = freq or {k: 1.0 / 6 for k in range(6)}
        self.debug_path = "/tmp/dram_oracle_debug"
        self.epoch_n = 0


def trace_metas(batch, size):
    """`metas` dict the attention model's trace branch indexes (models.py:526-527)."""
    return {"uid": [f"synthetic_{i}" for i in range(batch)], "original_size": [tuple(size)] * batch}
