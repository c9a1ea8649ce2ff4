"""world_size-2 gloo test of the gradient all-reduce plumbing (istnet_b200/parallel.py) on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


class Toy(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(8, 16)
        self.b = nn.Linear(16, 4)
        self.unused = nn.Linear(4, 4)  # never receives a gradient (like feats.fc)
        self.frozen = nn.Linear(4, 4)
        for p in self.frozen.parameters():
            p.requires_grad_(False)

    def forward(self, x):
        return self.b(torch.relu(self.a(x)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from istnet_b200.parallel import GradAllReducer, broadcast_module

    torch.manual_seed(100 + rank)  # different init per rank: broadcast must fix it
    m = Toy()
    broadcast_module(m)
    red = GradAllReducer(m, bucket_mb=0.0002)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(3, 8, 8, generator=g)  # 3 steps, global batch 8
    out = []
    for step in range(3):
        red.zero_grad()
        xs = x[step, rank * 4 : (rank + 1) * 4]
        m(xs).pow(2).mean().backward()
        red.finish()
        out.append([p.grad.clone() if p.grad is not None else None for p in m.parameters()])
    if rank == 0:
        ret["grads"] = out
        ret["state"] = {k: v.clone() for k, v in m.state_dict().items()}
        ret["nb"] = red.num_buckets()
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_allreduce_matches_full_batch_gradient():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    ref = Toy()
    ref.load_state_dict(ret["state"])
    g = torch.Generator().manual_seed(7)
    x = torch.randn(3, 8, 8, generator=g)
    assert ret["nb"] >= 2
    for step in range(3):
        ref.zero_grad(set_to_none=True)
        ref(x[step]).pow(2).mean().backward()
        for p, got in zip(ref.parameters(), ret["grads"][step]):
            if p.grad is None:
                assert got is None
            else:
                assert torch.allclose(p.grad, got, atol=1e-6), step


def test_gather_and_bind_grads_pack_fresh_gradients_into_the_buckets():
    """CUDA-graph path: gradients produced from grad=None are packed into the flat buckets by one multi-tensor copy."""
    from istnet_b200.parallel import GradAllReducer

    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    unused = torch.nn.Parameter(torch.zeros(4))  # never receives a gradient: must stay out of the buckets
    m.register_parameter("unused", unused)
    red = GradAllReducer(m, bucket_mb=1e-4)
    x = torch.randn(6, 5)
    m(x).square().sum().backward()
    red.finish()  # first step: discovers the used parameters and builds the buckets
    assert red.num_buckets() >= 2 and unused.grad is None
    for p in m.parameters():
        p.grad = None
    m(2 * x).square().sum().backward()  # fresh tensors, as autograd installs them inside the captured step
    want = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    red.gather_grads()
    red.bind_grads()
    flat = torch.cat([b["flat"] for b in red.buckets])
    assert flat.numel() == sum((v.numel() + 63) // 64 * 64 for v in want.values())  # every slot starts 256-byte aligned
    for n, p in m.named_parameters():
        if n in want:
            assert torch.equal(p.grad, want[n])
            assert any(p.grad.data_ptr() >= b["flat"].data_ptr() and p.grad.data_ptr() < b["flat"].data_ptr() + 4 * b["flat"].numel() for b in red.buckets)


def _worker_graph_path(rank, world, port, ret):
    """The N > 1 CUDA-graph flow of bench.py without a graph: first step eager (discovers the buckets), hooks removed, then every
    step starts from grad=None, ends with gather_grads() (the captured multi-tensor copy) + reduce_all(), bind_grads() once."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from istnet_b200.parallel import GradAllReducer, broadcast_module

    torch.manual_seed(200 + rank)
    m = Toy()
    broadcast_module(m)
    red = GradAllReducer(m, bucket_mb=0.0002)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(3, 8, 8, generator=g)
    red.zero_grad()
    m(x[0, rank * 4 : (rank + 1) * 4]).pow(2).mean().backward()
    red.finish()
    red.remove_hooks()
    out = []
    for step in (1, 2):
        for p in m.parameters():
            p.grad = None
        m(x[step, rank * 4 : (rank + 1) * 4]).pow(2).mean().backward()
        red.gather_grads()
        red.reduce_all()
        red.bind_grads()
        out.append([p.grad.clone() if p.grad is not None else None for p in m.parameters()])
    if rank == 0:
        ret["grads"] = out
        ret["state"] = {k: v.clone() for k, v in m.state_dict().items()}
    dist.barrier()
    dist.destroy_process_group()


def test_graph_path_gather_reduce_bind_matches_full_batch_gradient():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_graph_path, args=(2, _free_port(), ret), nprocs=2, join=True)
    ref = Toy()
    ref.load_state_dict(ret["state"])
    g = torch.Generator().manual_seed(9)
    x = torch.randn(3, 8, 8, generator=g)
    for i, step in enumerate((1, 2)):
        ref.zero_grad(set_to_none=True)
        ref(x[step]).pow(2).mean().backward()
        for (n, p), got in zip(ref.named_parameters(), ret["grads"][i]):
            if p.grad is None:
                assert got is None, n
            else:
                assert torch.allclose(p.grad, got, atol=1e-6), (step, n)


def _worker_overlap_path(rank, world, port, ret):
    """The per-bucket tail of the captured step (parallel.DataParallelStep) driven by the post-accumulate hooks, and the eager path
    after an `optimizer.zero_grad(set_to_none=True)` (reference solver, utils/solver.py:94) detached the gradients from the buckets."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from istnet_b200.parallel import GradAllReducer, broadcast_module

    torch.manual_seed(300 + rank)
    m = Toy()
    broadcast_module(m)
    red = GradAllReducer(m, bucket_mb=0.0002, early=lambda n: n.startswith("b."))
    g = torch.Generator().manual_seed(11)
    x = torch.randn(4, 8, 8, generator=g)
    red.zero_grad()
    m(x[0, rank * 4 : (rank + 1) * 4]).pow(2).mean().backward()
    red.finish()
    assert red.buckets[0]["early"] and not red.buckets[-1]["early"]
    out = []
    # (a) hook-driven tail, step starts from grad=None; buckets end up holding the SUM over ranks
    for p in m.parameters():
        p.grad = None
    red.begin_capture_overlap([])
    m(x[1, rank * 4 : (rank + 1) * 4]).pow(2).mean().backward()
    red.end_capture_overlap()
    red.bind_grads()
    assert abs(red.bucket_scale - 1.0 / world) < 1e-12
    out.append([p.grad.clone() * red.bucket_scale if p.grad is not None else None for p in m.parameters()])
    # (b) eager path where the caller set the gradients to None (torch's zero_grad default) instead of using red.zero_grad()
    for step in (2, 3):
        m.zero_grad(set_to_none=True)
        m(x[step, rank * 4 : (rank + 1) * 4]).pow(2).mean().backward()
        red.finish()
        for b in red.buckets:
            for p, v in zip(b["params"], b["views"]):
                assert p.grad.data_ptr() == v.data_ptr()
        out.append([p.grad.clone() if p.grad is not None else None for p in m.parameters()])
    if rank == 0:
        ret["grads"] = out
        ret["state"] = {k: v.clone() for k, v in m.state_dict().items()}
    dist.barrier()
    dist.destroy_process_group()


def test_hook_driven_bucket_tail_and_set_to_none_eager_path():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_overlap_path, args=(2, _free_port(), ret), nprocs=2, join=True)
    ref = Toy()
    ref.load_state_dict(ret["state"])
    g = torch.Generator().manual_seed(11)
    x = torch.randn(4, 8, 8, generator=g)
    for i, step in enumerate((1, 2, 3)):
        ref.zero_grad(set_to_none=True)
        ref(x[step]).pow(2).mean().backward()
        for (n, p), got in zip(ref.named_parameters(), ret["grads"][i]):
            if p.grad is None:
                assert got is None, n
            else:
                assert torch.allclose(p.grad, got, atol=1e-6), (step, n)
