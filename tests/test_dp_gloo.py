"""Host-side logic of the data-parallel loop on CPU: two gloo ranks (no GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sparsefactorization_b200 import dp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.body = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.GELU(), torch.nn.Linear(16, 1))
        self.unused = torch.nn.Embedding(4, 3)       # a parameter that never receives a gradient

    def forward(self, x):
        return self.body(x)


def _make_net():
    torch.manual_seed(0)
    return _Net()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1)
        X, Y = torch.randn(8, 6, generator=g), torch.randn(8, generator=g)
        net = _make_net()
        if rank == 1:                                  # ranks start different; broadcast must fix that
            with torch.no_grad():
                for p in net.parameters():
                    p.add_(1.0)
        wrap = dp.DataParallel(net)
        wrap.broadcast_parameters()
        lo, hi = dp.shard_range(8)
        opt = torch.optim.SGD(net.parameters(), lr=0.1)
        loss = dp.train_step(net, wrap, opt, torch.nn.MSELoss(), X[lo:hi], Y[lo:hi])
        assert net.unused.weight.grad is None          # never received a gradient: stays out of the bucket
        grads = torch.cat([p.grad.reshape(-1) for p in net.parameters() if p.grad is not None])
        params = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
        # TrainModelDP: two epochs on a tiny regression set, metrics all-reduced
        net2 = _make_net()
        hist = dp.TrainModelDP(net2, (X, Y), (X, Y), (X, Y), n_epochs=2, test_freq=1,
                               optimizer=torch.optim.SGD(net2.parameters(), lr=0.05), loss=torch.nn.MSELoss(),
                               problem="adding", saving_criteria=1e9, batch_size=2, verbose=False)
        q.put((rank, float(loss), grads.tolist(), params.tolist(), hist))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradients_equal_single_process_full_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference on the full batch
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(8, 6, generator=g), torch.randn(8, generator=g)
    net = _make_net()
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    dp.train_step(net, None, opt, torch.nn.MSELoss(), X, Y)
    ref_grads = torch.cat([p.grad.reshape(-1) for p in net.parameters() if p.grad is not None])
    ref_params = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    for rank, loss, grads, params, hist in res:
        torch.testing.assert_close(torch.tensor(grads), ref_grads, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(torch.tensor(params), ref_params, rtol=1e-5, atol=1e-6)
        assert len(hist) == 2 and hist[1]["train_loss"] <= hist[0]["train_loss"] + 1e-6
        assert "test_acc" in hist[0]
    # both ranks report the same all-reduced metrics
    assert res[0][4][1]["test_loss"] == pytest.approx(res[1][4][1]["test_loss"])


def test_shard_range_and_bucket():
    assert dp.shard_range(10, 0, 4) == (0, 2) and dp.shard_range(10, 3, 4) == (6, 8)
    assert dp.shard_range(10, 3, 4, keep_tail=True) == (6, 10) and dp.shard_range(10, 0, 4, keep_tail=True) == (0, 2)
    net = _make_net()
    b = dp.GradBucket(net.parameters())
    net(torch.randn(3, 6)).sum().backward()
    want = torch.cat([p.grad.reshape(-1) for p in net.body.parameters()])
    flat = b.pack().clone()
    # only the parameters that received a gradient are in the bucket (decided at the first pack)
    assert flat.numel() == sum(p.numel() for p in net.body.parameters())
    torch.testing.assert_close(flat, want)
    b.flat.mul_(2)
    b.unpack()
    got = torch.cat([p.grad.reshape(-1) for p in net.body.parameters()])
    torch.testing.assert_close(got, 2 * flat)
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(b.params, b.views))   # re-pointed, not copied
    assert net.unused.weight.grad is None
    # a second step: fresh gradients are copied into the same buffer
    for p in net.parameters():
        p.grad = None
    net(torch.randn(3, 6)).sum().backward()
    want2 = torch.cat([p.grad.reshape(-1) for p in net.body.parameters()])
    torch.testing.assert_close(b.pack(), want2)


def test_evaluate_counts_the_ragged_tail():
    torch.manual_seed(0)
    net = _make_net()
    X, Y = torch.randn(7, 6), torch.randn(7)
    loss, acc = dp.evaluate(net, X, Y, torch.nn.MSELoss(), "adding", batch_size=3)     # batches of 3, 3, 1
    want = float(torch.nn.functional.mse_loss(net(X).squeeze(), Y))
    assert abs(loss - want) < 1e-6


def test_dataset_creator_label_dtypes():
    from sparsefactorization_b200 import psf_utils
    from sparsefactorization_b200.genome import psf_utils as genome_utils
    from sparsefactorization_b200.lra import psf_utils as lra_utils
    data = torch.zeros(4, 3)
    assert psf_utils.DatasetCreator(data, torch.arange(4, dtype=torch.int32))[1][1].dtype == torch.long   # ListOps targets
    assert psf_utils.DatasetCreator(data, torch.rand(4))[1][1].dtype == torch.float32                      # Adding targets
    assert lra_utils.DatasetCreator(data, torch.arange(4, dtype=torch.int32))[0][1].dtype == torch.long
    assert genome_utils.DatasetCreator(data, torch.arange(4, dtype=torch.int16))[0][1].dtype == torch.long
    assert len(lra_utils.DatasetCreator(data, torch.arange(4))) == 4


def _clip_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        X, Y = 50.0 * torch.randn(8, 6, generator=g), 50.0 * torch.randn(8, generator=g)      # large gradients
        net = _make_net()
        wrap = dp.DataParallel(net)
        wrap.broadcast_parameters()
        lo, hi = dp.shard_range(8)
        opt = torch.optim.SGD(net.parameters(), lr=0.0)
        dp.train_step(net, wrap, opt, torch.nn.MSELoss(), X[lo:hi], Y[lo:hi], clip_grad_norm=1.0)
        grads = torch.cat([p.grad.reshape(-1) for p in net.parameters() if p.grad is not None])
        hist = dp.TrainModelDP(_make_net(), (X, Y), (X, Y), (X, Y), n_epochs=1, test_freq=1,
                               optimizer=torch.optim.SGD(net.parameters(), lr=0.0), loss=torch.nn.MSELoss(), problem="adding",
                               saving_criteria=1e9, batch_size=2, verbose=False, clip_grad_norm=1.0)
        q.put((rank, float(grads.norm()), grads.tolist(), len(hist)))
    finally:
        dist.destroy_process_group()


def test_clip_grad_norm_applies_after_the_all_reduce():
    """Genome loop (Genome_Clf/psf_utils.py:73): the clip sees the averaged gradient, identically on every rank."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_clip_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    (_, n0, g0, h0), (_, n1, g1, h1) = res
    assert abs(n0 - 1.0) < 1e-4 and abs(n1 - 1.0) < 1e-4          # clipped to the max norm
    assert g0 == g1 and h0 == h1 == 1
