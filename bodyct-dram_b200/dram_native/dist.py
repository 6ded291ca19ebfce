"""Data-parallel plumbing (new functionality: the reference has no distributed code).

One process per GPU, torch.distributed (NCCL over NVLink/NVSwitch on the GPU box, gloo in CPU tests).  Parity target =
the single-process reference run on the GLOBAL batch, hence three kinds of exchange:
  1. BatchNorm batch statistics (sum, sumsq, count forward; sum dz, sum dz*xhat backward) are all-reduced so every rank
     normalises with the global batch statistics (SyncBN semantics) — 2*C doubles per layer, latency-bound;
  2. parameter gradients are SUM-all-reduced in flat buckets, launched asynchronously as soon as a bucket's gradients
     exist (reverse layer order) so NCCL overlaps the rest of the backward pass;
  3. the loss's batch-global scalars (metrics.py:30,37,42,48) — see metrics.py in this package.
"""
import torch
import torch.distributed as td

_STATE = {"group": None, "sync_bn": True, "reducer": None, "peer": None}


def configure(group=None, sync_bn=True):
    _STATE["group"], _STATE["sync_bn"] = group, sync_bn


def world_size():
    return td.get_world_size(_STATE["group"]) if td.is_available() and td.is_initialized() else 1


def rank():
    return td.get_rank(_STATE["group"]) if td.is_available() and td.is_initialized() else 0


def active():
    """data parallel exchanges on?  (`suspend()` turns them off inside an initialised process group: every rank then trains
    on its own, which is how bench.py measures the per-GPU free-running step time next to the synchronised one)"""
    return world_size() > 1 and not _STATE.get("suspended", False)


def suspend(flag=True):
    _STATE["suspended"] = bool(flag)


def shard(items):
    """This rank's share of a (sorted) work list: items[rank::world] — scans / lobe batches shard with no collective."""
    items = list(items)
    return items[rank()::world_size()] if active() else items


def all_gather_object(obj):
    """-> [obj of rank 0, obj of rank 1, ...] (small host objects: validation results)"""
    if not active():
        return [obj]
    out = [None] * world_size()
    td.all_gather_object(out, obj, group=_STATE["group"])
    return out


def check_uniform_batch(shape_key):
    """Every rank must step the same batch shape: `allreduce_stats` multiplies the local voxel count by the world size
    instead of exchanging it, and a shape change re-captures the step's CUDA graph (NCCL calls inside) on every rank.
    One small host-side gather per NEW shape; raises on a mismatch (use drop_last=True loaders under data parallelism)."""
    keys = all_gather_object(tuple(shape_key))
    if any(k != keys[0] for k in keys):
        raise RuntimeError(f"data parallel training needs the same batch shape on every rank, got {keys}")


class PeerGroup:
    """NVLink peer-memory mailboxes for the step's small exchanges (csrc/peer.cu): every rank allocates a mailbox with
    cudaMalloc, exports it through CUDA IPC, and maps every other rank's mailbox.  One kernel per exchange: push to all
    peers, signal, wait, reduce in rank order.  `DRAM_PEER=0` (or a failed IPC mapping) keeps the NCCL path."""

    def __init__(self):
        import ctypes
        from . import lib as _lib
        self._lib, self._ct = _lib, ctypes
        L = _lib.load()
        self.rank, self.world = rank(), world_size()
        if self.world > L.dram_peer_max_ranks():
            raise RuntimeError(f"peer mailboxes are built for <= {L.dram_peer_max_ranks()} ranks of one node")
        mine = ctypes.c_void_p()
        _lib.check(L.dram_peer_alloc(ctypes.byref(mine)), "peer_alloc")
        self.mine = mine
        handle = (ctypes.c_ubyte * 64)()
        _lib.check(L.dram_peer_export(mine, handle), "peer_export")
        handles = [None] * self.world
        td.all_gather_object(handles, bytes(handle), group=_STATE["group"])
        self.opened = []
        ptrs = []
        for q, h in enumerate(handles):
            if q == self.rank:
                ptrs.append(mine.value)
                continue
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
            p = ctypes.c_void_p()
            _lib.check(L.dram_peer_open(buf, ctypes.byref(p)), f"peer_open(rank {q})")
            self.opened.append(p)
            ptrs.append(p.value)
        self.boxes = (ctypes.c_void_p * self.world)(*ptrs)
        self.max_doubles = L.dram_peer_max_doubles()
        td.barrier(group=_STATE["group"])

    def _arr(self, *ptrs):
        return (self._ct.c_void_p * len(ptrs))(*ptrs)

    def all_reduce_(self, t, out=None):
        """SUM of a small contiguous float64 CUDA tensor over the ranks (in place unless `out`)."""
        out = t if out is None else out
        self._lib.check(self._lib.load().dram_peer_allreduce_f64(self.boxes, self._arr(t.data_ptr()), self._arr(out.data_ptr()),
                                                                 t.numel(), self.rank, self.world, 1,
                                                                 torch.cuda.current_stream().cuda_stream), "peer_allreduce_f64")
        return out

    def bn_finalize(self, sums, count, gamma, beta, running_mean, running_var, momentum, eps, n_updates):
        """global BatchNorm statistics + finalize in one kernel -> (mean, rstd, scale, shift, global count)"""
        C = sums.numel() // 2
        gsums = torch.empty(2 * C + 1, device=sums.device, dtype=torch.float64)
        o = torch.empty((4, C), device=sums.device, dtype=torch.float32)
        p = lambda x: 0 if x is None else x.data_ptr()
        cnt = (self._ct.c_double * 1)(float(count))
        self._lib.check(self._lib.load().dram_bn_finalize_peer(
            self.boxes, self._arr(sums.data_ptr()), cnt, self._arr(gsums.data_ptr()), self.rank, self.world, 1,
            self._arr(p(gamma)), self._arr(p(beta)), self._arr(p(running_mean)), self._arr(p(running_var)), float(momentum),
            float(eps), int(n_updates), self._arr(o[0].data_ptr()), self._arr(o[1].data_ptr()), self._arr(o[2].data_ptr()),
            self._arr(o[3].data_ptr()), C, torch.cuda.current_stream().cuda_stream), "bn_finalize_peer")
        return o[0], o[1], o[2], o[3]


def init_peer():
    """Set up the peer mailboxes once per process (NCCL backend, all ranks on one node); returns the group or None."""
    import os
    if _STATE.get("peer") is not None or not active() or os.environ.get("DRAM_PEER", "1") != "1":
        return _STATE.get("peer")
    if td.get_backend(_STATE["group"]) != "nccl" or not torch.cuda.is_available():
        return None
    try:
        _STATE["peer"] = PeerGroup()
    except Exception as e:                                   # noqa: BLE001  (no IPC between the ranks: NCCL does the exchanges)
        import logging
        logging.getLogger("dram.dist").warning("peer mailboxes unavailable (%s): BatchNorm / loss exchanges go through NCCL", e)
        _STATE["peer"] = None
    return _STATE["peer"]


def peer():
    return _STATE.get("peer") if (active() and _STATE["sync_bn"]) else None


def all_reduce_(t):
    if active():
        pg = _STATE.get("peer")
        if pg is not None and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and 0 < t.numel() <= pg.max_doubles:
            return pg.all_reduce_(t)
        td.all_reduce(t, op=td.ReduceOp.SUM, group=_STATE["group"])
    return t


def allreduce_stats(sums, count):
    """In-place SUM of the BatchNorm forward sums across ranks; returns the global element count.
    Every rank holds the same per-GPU batch (weak scaling), so the count needs no exchange (and no host sync)."""
    if not (active() and _STATE["sync_bn"]):
        return count
    all_reduce_(sums)
    return float(count) * world_size()


def allreduce_sums(sums):
    """Global copy of the BatchNorm backward sums (the local ones stay intact: they are this rank's dgamma/dbeta)."""
    if not (active() and _STATE["sync_bn"]):
        return sums
    pg = _STATE.get("peer")
    if pg is not None and sums.is_cuda and sums.is_contiguous() and sums.numel() <= pg.max_doubles:
        return pg.all_reduce_(sums, out=torch.empty_like(sums))      # one kernel, the local sums stay untouched
    g = sums.clone()
    all_reduce_(g)
    return g


class GradReducer:
    """Gradient all-reduce (SUM) over ONE flat fp32 buffer that the parameters' `.grad` tensors are views of.

    Layout: parameters in REVERSE registration order (the order backward produces them), split into contiguous buckets of
    ~bucket_mb.  The tensor-core wgrad kernels write `dw` straight into the buffer (`grad_view`, handed to autograd, which
    adopts the tensor as `.grad` without a copy); any gradient that arrives in other storage is copied in once by the
    post-accumulate-grad hook and `.grad` is re-pointed at the view.  Nothing is packed or unpacked around the all-reduce
    (round 1 moved every gradient through two `copy_` kernels: ~170 launches per step).
    overlap=True: a bucket's async all_reduce is issued as soon as its last gradient exists (NCCL runs under the rest of the
    backward pass); overlap=False (default, DRAM_GRAD_OVERLAP=0): one all_reduce over the whole buffer in `finish()` —
    the persistent one-CTA-per-SM convolution kernels leave NCCL no SM to run on concurrently, so the overlap mostly costs
    them a second wave.
    `scale`: the reference's loss terms are batch SUMS / global means formed before backward, so the correct reduction is
    a plain SUM (scale=1); pass 1/world_size for mean semantics."""

    def __init__(self, params, bucket_mb=16.0, scale=1.0, overlap=None):
        import os
        self.params = [p for p in params if p.requires_grad]
        self.scale = scale
        self.overlap = (os.environ.get("DRAM_GRAD_OVERLAP", "0") == "1") if overlap is None else bool(overlap)
        self.slot, self.buckets, self.owner = {}, [], {}
        off, start, cur = 0, 0, []
        for p in reversed(self.params):
            self.slot[id(p)] = (off, p.numel())
            self.owner[id(p)] = len(self.buckets)
            cur.append(p)
            off += (p.numel() + 3) // 4 * 4                       # 16-byte aligned views (vector stores in the kernels)
            if (off - start) * 4 >= bucket_mb * (1 << 20):
                self.buckets.append((start, off, cur))
                start, cur = off, []
        if cur:
            self.buckets.append((start, off, cur))
        self.total = off
        self.flat = None
        self.by_ptr = {}
        self.work = []
        self.hooks = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
        _STATE["reducer"] = self
        self.reset()

    def _ensure(self):
        dev = self.params[0].device
        if self.flat is None or self.flat.device != dev:
            self.flat = torch.zeros(self.total, device=dev, dtype=torch.float32)
            self.by_ptr = {p.data_ptr(): p for p in self.params}
        return self.flat

    def view(self, p):
        off, n = self.slot[id(p)]
        return self._ensure()[off:off + n].view_as(p)

    def grad_view(self, w):
        """Storage for a freshly computed gradient of parameter `w` (looked up by address: autograd hands the Functions a
        different Python object), or None when `w` is not reduced here or already holds a gradient (accumulation)."""
        if not active() or not self.params:
            return None
        self._ensure()
        p = self.by_ptr.get(w.data_ptr())
        if p is None or p.grad is not None or tuple(p.shape) != tuple(w.shape):
            return None
        return self.view(p)

    def reset(self):
        self.pending = [len(b[2]) for b in self.buckets]
        self.seen = set()
        self.work = []

    def _adopt(self, p):
        v = self.view(p)
        if p.grad is None:
            v.zero_()
        elif p.grad.data_ptr() != v.data_ptr():
            v.copy_(p.grad)
        else:
            return
        p.grad = v

    def _hook(self, p):
        if not active():
            return
        self._adopt(p)
        self.seen.add(id(p))
        bi = self.owner[id(p)]
        self.pending[bi] -= 1
        if self.pending[bi] == 0 and self.overlap:
            self._launch(bi)

    def _launch(self, bi):
        from . import functional as _F
        _F.SIDE.join()                    # weight gradients may be produced on the side stream (functional._SideStream)
        start, end, _ = self.buckets[bi]
        self.pending[bi] = -1
        self.work.append(td.all_reduce(self._ensure()[start:end], op=td.ReduceOp.SUM, group=_STATE["group"], async_op=True))

    def finish(self):
        """All-reduce what has not been launched yet, wait, leave the sums in `.grad`.  Call after loss.backward()."""
        if active():
            for p in self.params:                          # parameters that never received a gradient this step: zeros
                if id(p) not in self.seen:
                    self._adopt(p)
            if self.overlap:
                for bi in range(len(self.buckets)):
                    if self.pending[bi] >= 0:
                        self._launch(bi)
            else:
                from . import functional as _F
                _F.SIDE.join()
                self.work.append(td.all_reduce(self._ensure(), op=td.ReduceOp.SUM, group=_STATE["group"], async_op=True))
            from . import lib as _lib
            prof = _lib.PROFILE.enabled                       # bench.py's per-kernel events: time the exposed part of the all-reduce
            if prof:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            for w in self.work:
                w.wait()
            if prof:
                e1.record()
                _lib.PROFILE.records.append(("nccl_grad_allreduce_exposed", {"bytes": 4.0 * self.total}, e0, e1))
            if self.scale != 1.0:
                self.flat.mul_(self.scale)
        self.reset()

    def remove(self):
        for h in self.hooks:
            h.remove()
        if _STATE.get("reducer") is self:
            _STATE["reducer"] = None


def grad_out(w, shape=None):
    """Tensor a backward kernel should write the gradient of parameter `w` into: its slice of the data-parallel flat
    gradient buffer when there is one (no copy before the all-reduce), else fresh memory."""
    r = _STATE.get("reducer")
    v = r.grad_view(w) if r is not None else None
    if v is not None:
        return v
    return torch.empty(tuple(w.shape) if shape is None else shape, device=w.device, dtype=torch.float32)
