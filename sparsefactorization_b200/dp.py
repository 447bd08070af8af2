"""Data-parallel training of PSFNet: one process per GPU, batch sharded by rank, one flat-bucket
gradient all-reduce per step (NCCL over NVLink on the GPU box, gloo in the CPU tests).

The reference has no distributed code at all; this is the DP loop the north star
asks for around the same ``TrainModel`` semantics (SyntheticExperiments/psf_utils.py:48-137).
Design notes:
  * every sample's chain is independent, so ranks exchange nothing on the data path; the only
    collective is ``all_reduce(SUM)`` of the flattened parameter gradients followed by ``/ world``
    (local-mean losses over equal shards average to the global-mean gradient);
  * parameters that receive no gradient (``pos_embedding`` without positional embedding,
    ``embedding`` for Adding) are left out of the bucket (decided once from the first backward pass - a
    property of the model, identical on every rank), so no ``find_unused_parameters`` machinery is needed;
  * no per-parameter pack / unpack launches: one multi-tensor copy into the flat buffer, ``ReduceOp.AVG`` inside
    NCCL, and ``p.grad`` re-pointed at the buffer slices afterwards;
  * ranks start from identical weights (broadcast from rank 0).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized()


def world_size() -> int:
    return dist.get_world_size() if is_dist() else 1


def rank() -> int:
    return dist.get_rank() if is_dist() else 0


def shard_range(n_items: int, rank_: Optional[int] = None, world: Optional[int] = None, keep_tail: bool = False):
    """Contiguous shard of ``range(n_items)`` for a rank.  Training shards are equal (the tail that does not divide is
    dropped, so that every rank takes the same number of steps); ``keep_tail`` gives the remainder to the last rank
    (evaluation covers every sample, as the reference does)."""
    rank_ = rank() if rank_ is None else rank_
    world = world_size() if world is None else world
    per = n_items // world
    hi = (rank_ + 1) * per
    if keep_tail and rank_ == world - 1:
        hi = n_items
    return rank_ * per, hi


class GradBucket:
    """One flat fp32 buffer over the parameters that actually receive a gradient.

    Which ones do is decided once, after the first backward pass (``pos_embedding`` without positional embedding and
    ``embedding`` for Adding are constructed but unused, psf.py:98-107 vs :151-162; the answer is a property of the
    model, hence identical on every rank).  ``pack`` is one multi-tensor copy; after the all-reduce every ``p.grad`` is
    re-pointed at its slice of the buffer, so there is no copy back."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.candidates: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.candidates:
            raise ValueError("no trainable parameters")
        self.params: Optional[List[torch.nn.Parameter]] = None
        self.flat = None
        self.views = None

    def _build(self) -> None:
        self.params = [p for p in self.candidates if p.grad is not None]
        if not self.params:
            raise RuntimeError("backward produced no gradient")
        dev, dtype = self.params[0].device, self.params[0].dtype
        sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(sizes), dtype=dtype, device=dev)
        self.views = [v.view_as(p) for v, p in zip(self.flat.split(sizes), self.params)]

    def pack(self) -> torch.Tensor:
        if self.params is None:
            self._build()
        grads = []
        for p in self.params:
            if p.grad is None:
                raise RuntimeError("a parameter that had a gradient in the first step has none now; "
                                   "rebuild the DataParallel wrapper after changing the model")
            grads.append(p.grad)
        if any(g.data_ptr() != v.data_ptr() for g, v in zip(grads, self.views)):
            torch._foreach_copy_(self.views, grads)          # one multi-tensor launch
        return self.flat

    def unpack(self) -> None:
        for p, v in zip(self.params, self.views):
            p.grad = v                                       # re-point, no copy


class DataParallel:
    """Minimal DP wrapper: ``broadcast_parameters()`` once, ``allreduce_gradients()`` after backward."""

    def __init__(self, net: torch.nn.Module, process_group=None):
        self.net = net
        self.group = process_group
        self.bucket = GradBucket(net.parameters())

    def broadcast_parameters(self, src: int = 0) -> None:
        if world_size() == 1:
            return
        with torch.no_grad():
            for t in list(self.net.parameters()) + list(self.net.buffers()):
                dist.broadcast(t.data, src=src, group=self.group)

    def allreduce_gradients(self) -> None:
        """Mean of the ranks' gradients in place: one all-reduce of the live bucket (0.62 MB at cfg2 - latency bound)."""
        w = world_size()
        if w == 1:
            return
        flat = self.bucket.pack()
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)      # the division happens inside NCCL
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(w)
        self.bucket.unpack()


def train_step(net, dp: Optional[DataParallel], optimizer, loss_fn, X, Y, clip_grad_norm: Optional[float] = None):
    """zero_grad -> forward -> loss -> backward -> (all-reduce) -> step.  Returns the detached local loss."""
    optimizer.zero_grad(set_to_none=True)
    pred = net(X)
    out = loss_fn(pred.squeeze(), Y)
    out.backward()
    if dp is not None:
        dp.allreduce_gradients()
    if clip_grad_norm is not None:
        torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=clip_grad_norm)
    optimizer.step()
    return out.detach()


@torch.no_grad()
def evaluate(net, X, Y, loss_fn, problem: str, batch_size: int):
    """Loss / accuracy over this rank's shard, all-reduced to global values (percent accuracy)."""
    dev = X.device
    stats = torch.zeros(3, dtype=torch.float64, device=dev)   # loss_sum, correct, total
    for i in range(0, X.shape[0], batch_size):              # the last batch may be short: every sample counts
        xb, yb = X[i:i + batch_size], Y[i:i + batch_size]
        pred = net(xb)
        stats[0] += loss_fn(pred.reshape(yb.shape[0], -1).squeeze(-1), yb).double() * yb.shape[0]
        if problem == 'adding':
            stats[1] += (torch.abs(pred.reshape(-1) - yb) < 0.04).sum()
        else:
            stats[1] += pred.max(1)[1].eq(yb).sum()
        stats[2] += yb.shape[0]
    if is_dist():
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    total = max(float(stats[2]), 1.0)
    return float(stats[0]) / total, 100.0 * float(stats[1]) / total


def TrainModelDP(net, train_data, val_data, test_data, n_epochs, test_freq, optimizer, loss, problem,
                 saving_criteria, batch_size, shuffle_seed=0, verbose=True, clip_grad_norm=None):
    """Data-parallel counterpart of ``TrainModel``.

    ``*_data`` are ``(X, Y)`` tensor pairs holding the FULL dataset on every rank (already on the
    training device); each rank trains on its contiguous shard with per-GPU batch ``batch_size``.
    Rank 0 prints the reference's per-epoch lines and saves the state dict under the same rule.
    ``clip_grad_norm``: max gradient norm after the all-reduce (the Genome loop clips at 1.0, Genome_Clf/psf_utils.py:73).
    """
    dp = DataParallel(net)
    dp.broadcast_parameters()
    Xtr, Ytr = train_data
    lo, hi = shard_range(Xtr.shape[0])
    Xs, Ys = Xtr[lo:hi], Ytr[lo:hi]
    history = []
    for epoch in range(n_epochs):
        gen = torch.Generator(device=Xs.device).manual_seed(shuffle_seed + epoch * 7919 + rank())
        perm = torch.randperm(Xs.shape[0], generator=gen, device=Xs.device)
        running, steps = None, 0
        for i in range(0, Xs.shape[0] - batch_size + 1, batch_size):
            idx = perm[i:i + batch_size]
            out = train_step(net, dp, optimizer, loss, Xs[idx], Ys[idx], clip_grad_norm)
            running = out if running is None else running + out
            steps += 1
        mean_loss = torch.zeros((), device=Xs.device) if running is None else running / max(steps, 1)
        if is_dist():
            dist.all_reduce(mean_loss, op=dist.ReduceOp.SUM)
            mean_loss /= world_size()
        rec = {"epoch": epoch, "train_loss": float(mean_loss)}
        if rank() == 0 and verbose:
            print("Epoch {} - Training loss:  {}".format(epoch, rec["train_loss"]))
        if epoch % test_freq == 0:
            net.eval()
            vs = shard_range(val_data[0].shape[0], keep_tail=True)
            ts = shard_range(test_data[0].shape[0], keep_tail=True)
            vl, va = evaluate(net, val_data[0][vs[0]:vs[1]], val_data[1][vs[0]:vs[1]], loss, problem, batch_size)
            tl, ta = evaluate(net, test_data[0][ts[0]:ts[1]], test_data[1][ts[0]:ts[1]], loss, problem, batch_size)
            net.train()
            rec.update(val_loss=vl, val_acc=va, test_loss=tl, test_acc=ta)
            if rank() == 0 and verbose:
                print("Val  loss: {}".format(vl))
                print("Test loss: {}".format(tl))
                print("Val  accuracy: {}".format(va))
                print("Test accuracy: {}".format(ta))
                print('_' * 40)
            if rank() == 0 and ta > saving_criteria:
                torch.save(net.state_dict(), '{}_epoch{}_acc{}.pt'.format(problem, epoch, ta))
        history.append(rec)
    return history
