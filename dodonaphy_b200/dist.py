"""Multi-GPU sharding of the hot path: one process per GPU, ``torch.distributed`` (NCCL over
NVLink on the GPUs; gloo in the CPU tests) for the only exchanges the path has.

The reference has no distributed code; the two independent axes are

* **samples / chains** (``vi.py:328-333``, ``mcmc.py:136-146``): every rank evaluates its own
  contiguous slice of the batch with the full alignment; nothing is exchanged by the kernels.  The
  driver needs every per-sample lnL (IWAE logsumexp / acceptance tests), so ``gather_samples`` does
  one all_gather of B doubles.
* **site-pattern blocks** (long alignments): patterns are conditionally independent given the tree,
  so every rank holds L/world patterns of every tip, runs the same trees on its block and the
  partial lnL (and the partial d lnL/d blens, 2S-2 doubles per tree) are summed with one all_reduce
  -- the "scalar NCCL allreduce" of BASELINE.json.  Reduction order inside a rank is fixed, across
  ranks it is NCCL's: results agree with 1 GPU to ~1e-13 relative.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous near-equal partition of range(n): returns (lo, hi) for this rank."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_samples(x, rank=None, world=None):
    """Slice the leading (sample/chain) axis for this rank."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def shard_sites(tipmask, weights, rank=None, world=None):
    """Slice the site-pattern axis of the (S, L) tip masks and (L,) weights for this rank."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(tipmask.shape[1], rank, world)
    return np.ascontiguousarray(tipmask[:, lo:hi]), np.ascontiguousarray(weights[lo:hi])


def gather_samples(local, total, group=None):
    """all_gather of per-sample results when the batch was split with ``shard_samples``.
    ``local``: (b_local, ...) tensor; returns (total, ...) on every rank."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    sizes = [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]
    if len(set(sizes)) == 1:
        out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def reduce_site_blocks(lnl, grad_blens=None, group=None):
    """Sum the per-block partial lnL (B,) and partial gradients (B, 2S-2) over ranks in ONE all_reduce."""
    if dist.get_world_size(group) == 1:
        return lnl, grad_blens
    if grad_blens is None:
        out = lnl.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out, None
    packed = torch.cat([lnl.reshape(-1, 1), grad_blens.reshape(lnl.numel(), -1)], dim=1).contiguous()
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed[:, 0].reshape(lnl.shape), packed[:, 1:].reshape(grad_blens.shape)


class SiteShardedStep:
    """One evaluation of a site-sharded likelihood as ONE replayable unit: this rank's pruning pass
    writes lnL and d lnL/d blens of its pattern block into a single packed buffer (B, 1 + 2S-2), and the
    all_reduce over ranks runs on that buffer in place -- kernels and collective captured together in a
    CUDA graph (NCCL collectives are capturable), so an evaluation costs one graph launch instead of
    the Python launch sequence plus a pack / reduce / unpack round trip.  ``run(peel, blens)`` copies the
    new tree into the captured inputs and replays; results are views of the packed buffer."""

    def __init__(self, engine, model, peel, blens, rates=None, mix=False, grad=True, group=None, graph=True, reduce=True):
        self.engine, self.model, self.rates, self.mix, self.grad, self.group = engine, model, rates, mix, grad, group
        dev = engine.device
        self.peel = peel.detach().to(dev, dtype=torch.int64).reshape(-1, engine.S - 1, 3).contiguous().clone()
        self.blens = blens.detach().to(dev, dtype=torch.float64).reshape(-1, engine.E).contiguous().clone()
        B = int(self.blens.shape[0])
        if B == 1:
            self.packed = torch.zeros(1 + engine.E if grad else 1, dtype=torch.float64, device=dev)
            self.lnl = self.packed[:1]
            self.g = self.packed[1:].reshape(1, engine.E) if grad else None
        else:  # separate regions of one buffer: [B lnL | B x E gradients]
            self.packed = torch.zeros(B * (1 + engine.E) if grad else B, dtype=torch.float64, device=dev)
            self.lnl = self.packed[:B]
            self.g = self.packed[B:].reshape(B, engine.E) if grad else None
        self.world = dist.get_world_size(group) if (reduce and dist.is_initialized()) else 1
        self.graph = None
        if graph:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._step()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step()
                self.graph = g
            except Exception as exc:  # a collective that cannot be captured here: launch by launch
                self.graph_error = repr(exc)
                torch.cuda.synchronize(dev)

    def _step(self):
        self.engine.loglik(self.peel, self.blens, self.model, rates=self.rates, mix=self.mix, grad=self.grad,
                           out=(self.lnl, self.g))
        if self.world > 1:
            dist.all_reduce(self.packed, op=dist.ReduceOp.SUM, group=self.group)

    def run(self, peel=None, blens=None):
        if peel is not None:
            self.peel.copy_(peel.reshape(self.peel.shape), non_blocking=True)
        if blens is not None:
            self.blens.copy_(blens.reshape(self.blens.shape), non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step()
        return self.lnl, self.g


class SiteShardedLikelihood:
    """Long-alignment likelihood (BASELINE.json configs[4]): this rank's block of site patterns on
    this rank's GPU, one all_reduce per evaluation."""

    def __init__(self, tipmask, weights, device=None, group=None):
        from .engine import PruneEngine

        self.group = group
        tm, w = shard_sites(np.asarray(tipmask), np.asarray(weights)) if dist.is_initialized() else (tipmask, weights)
        self.engine = PruneEngine(tm, w, device=device)

    def loglik(self, peel, blens, model, rates=None, mix=False, grad=False):
        ll, g = self.engine.loglik(peel, blens, model, rates=rates, mix=mix, grad=grad)
        if dist.is_initialized():
            ll, g = reduce_site_blocks(ll, g, group=self.group)
        return ll, g
