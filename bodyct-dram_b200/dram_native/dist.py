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

_STATE = {"group": None, "sync_bn": True}


def configure(group=None, sync_bn=True):
    _STATE["group"], _STATE["sync_bn"] = group, sync_bn


def world_size():
    return td.get_world_size(_STATE["group"]) if td.is_available() and td.is_initialized() else 1


def rank():
    return td.get_rank(_STATE["group"]) if td.is_available() and td.is_initialized() else 0


def active():
    return world_size() > 1


def all_reduce_(t):
    if active():
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
    g = sums.clone()
    all_reduce_(g)
    return g


class GradReducer:
    """Bucketed asynchronous gradient all-reduce (SUM) driven by post-accumulate-grad hooks.

    Parameters are packed, in REVERSE registration order (the order backward produces them), into flat buckets of
    ~bucket_mb; when the last gradient of a bucket has been accumulated the bucket is copied into its flat buffer and
    an async all_reduce is issued; `finish()` waits and scatters the sums back into `.grad`.
    `scale`: the reference's loss terms are batch SUMS / global means formed before backward, so the correct reduction
    is a plain SUM (scale=1); pass 1/world_size for mean semantics."""

    def __init__(self, params, bucket_mb=16.0, scale=1.0):
        self.params = [p for p in params if p.requires_grad]
        self.scale = scale
        self.buckets = []
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_mb * (1 << 20):
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(cur)
        self.owner = {}
        for bi, b in enumerate(self.buckets):
            for p in b:
                self.owner[p] = bi
        self.flat = [None] * len(self.buckets)
        self.pending = [0] * len(self.buckets)
        self.work = []
        self.hooks = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
        self.reset()

    def reset(self):
        self.pending = [len(b) for b in self.buckets]
        self.work = []

    def _hook(self, p):
        if not active():
            return
        bi = self.owner[p]
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi):
        from . import functional as _F
        _F.SIDE.join()                    # weight gradients are produced on the side stream (functional._SideStream)
        b = self.buckets[bi]
        n = sum(p.numel() for p in b)
        if self.flat[bi] is None or self.flat[bi].numel() != n or self.flat[bi].device != b[0].device:
            self.flat[bi] = torch.empty(n, device=b[0].device, dtype=torch.float32)
        flat, off = self.flat[bi], 0
        for p in b:
            k = p.numel()
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            flat[off:off + k].copy_(g.reshape(-1))
            off += k
        self.work.append((bi, td.all_reduce(flat, op=td.ReduceOp.SUM, group=_STATE["group"], async_op=True)))

    def finish(self):
        """Wait for every bucket and write the reduced gradients back.  Call after loss.backward()."""
        if active():
            for bi in range(len(self.buckets)):          # parameters that never received a gradient this step
                if self.pending[bi] > 0:
                    self._launch(bi)
            for bi, w in self.work:
                w.wait()
                flat, off = self.flat[bi], 0
                for p in self.buckets[bi]:
                    k = p.numel()
                    g = flat[off:off + k].view_as(p)
                    if self.scale != 1.0:
                        g = g * self.scale
                    if p.grad is None:
                        p.grad = g.clone()
                    else:
                        p.grad.copy_(g)
                    off += k
        self.reset()

    def remove(self):
        for h in self.hooks:
            h.remove()
