"""CPU, world_size 2, gloo: the data-parallel plumbing (bucketed gradient all-reduce with DDP-mean semantics,
replica broadcast, batch-axis sharding).  The CUDA kernels are not involved: the model here is a small
torch module, so that the N>1 host logic is covered without a GPU."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_model():
    torch.manual_seed(3)
    enc = torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.Tanh(), torch.nn.Linear(16, 8))
    model = torch.nn.Module()
    model.res2plus1d = torch.nn.Module()
    model.res2plus1d.conv1 = enc[0]
    model.res2plus1d.conv2 = enc[2]
    for name in ("conv3", "conv4", "conv5"):
        setattr(model.res2plus1d, name, torch.nn.Linear(8, 8))
    model.linear = torch.nn.Linear(8, 2)

    def fwd(x):
        h = torch.tanh(model.res2plus1d.conv1(x))
        h = model.res2plus1d.conv2(h)
        for name in ("conv3", "conv4", "conv5"):
            h = torch.tanh(getattr(model.res2plus1d, name)(h))
        return model.linear(h)

    model.forward = fwd
    return model


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import dp_b200  # noqa: F401
    from dp_b200 import distributed as dpd

    r, lr, w = dpd.init_distributed("gloo")
    assert (r, w) == (rank, world)
    model = _make_model()
    if rank == 1:   # replicas must be made identical by the trainer's broadcast
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    g = torch.Generator().manual_seed(11)
    X = torch.randn(8, 12, generator=g)
    Y = torch.randint(0, 2, (8,), generator=g)
    shard = dpd.shard_range(8, rank, world)
    xs, ys = X[shard.start:shard.stop], Y[shard.start:shard.stop]
    loss_fn = lambda o, t: torch.nn.functional.cross_entropy(o, t, reduction="sum")  # noqa: E731  (Focal/CE are sums)
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    trainer = dpd.DataParallelTrainer(model, loss_fn, opt, max_norm_grad=None)
    loss, out = trainer.step(xs, ys)
    grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    torch.save({"grads": grads, "loss": loss.detach(), "params": {n: p.detach().clone() for n, p in model.named_parameters()}},
               os.path.join(out_dir, f"rank{rank}.pt"))
    # second step: the hooks re-arm
    loss2, _ = trainer.step(xs, ys)
    assert torch.allclose(loss2, loss)
    for n, p in model.named_parameters():
        assert torch.allclose(p.grad, grads[n], atol=1e-6), n
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_bucketed_allreduce_is_ddp_mean(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]
    # replicas identical after broadcast, gradients identical after the all-reduce
    for n in res[0]["params"]:
        assert torch.equal(res[0]["params"][n], res[1]["params"][n]), n
        assert torch.allclose(res[0]["grads"][n], res[1]["grads"][n], atol=1e-7), n
    # reference: mean over ranks of the per-shard gradients of the same (rank-0) weights
    model = _make_model()
    g = torch.Generator().manual_seed(11)
    X = torch.randn(8, 12, generator=g)
    Y = torch.randint(0, 2, (8,), generator=g)
    acc = {n: torch.zeros_like(p) for n, p in model.named_parameters()}
    for r in range(world):
        model.zero_grad()
        sl = slice(r * 4, (r + 1) * 4)
        torch.nn.functional.cross_entropy(model(X[sl]), Y[sl], reduction="sum").backward()
        for n, p in model.named_parameters():
            acc[n] += p.grad / world
    for n in acc:
        assert torch.allclose(res[0]["grads"][n], acc[n], atol=1e-6), n
