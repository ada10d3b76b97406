"""N > 1 host logic on CPU: 2 ranks over gloo (the GPU path uses the same code with NCCL).

Checks the loop harness (mmgl_b200/train.py): rank-disjoint data, DDP gradient averaging equals the single-process
gradient over the union of the ranks' micro-batches, ``no_sync`` under accumulation gives the same update as
all-reducing every micro-step, and max-over-ranks timing.  A small fp64 regression module stands in for the model
(the model's kernels are CUDA-only by design)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmgl_b200 import train


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = torch.nn.Linear(6, 5).double()
        self.b = torch.nn.Linear(5, 1).double()
        self.frozen = torch.nn.Linear(6, 6).double()
        for p in self.frozen.parameters():
            p.requires_grad = False

    def forward(self, x, y):
        return ((self.b(torch.tanh(self.a(self.frozen(x)))) - y) ** 2).mean()


def _batch(rank, step, n=4):
    g = torch.Generator().manual_seed(train.rank_seed(1234, rank, step))
    return dict(x=torch.randn(n, 6, generator=g, dtype=torch.float64), y=torch.randn(n, 1, generator=g, dtype=torch.float64))


def _worker(rank, world, port, accum, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    net = torch.nn.parallel.DistributedDataParallel(Toy())
    opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr=0.1)
    loss = train.optimizer_step(net, opt, [_batch(rank, s) for s in range(accum)], lambda m, b: m(**b),
                                accum_steps=accum, grad_clip=10.0)
    mx = train.max_over_ranks(float(rank + 1), "cpu")
    if rank == 0:
        torch.save({"state": {k: v.clone() for k, v in net.module.state_dict().items()}, "loss": loss, "max": mx}, out)
    dist.barrier()
    dist.destroy_process_group()


class Toy32(Toy):
    """fp32 twin: its two weight matrices take FlatGradSync's flat-buffer path, its biases the flattened-small path"""

    def __init__(self):
        super().__init__()
        self.float()


def _worker_flat(rank, world, port, accum, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    net = Toy32()
    if rank == 1:                       # FlatGradSync broadcasts rank 0's parameters: start rank 1 somewhere else
        with torch.no_grad():
            for p in net.parameters():
                p.add_(1.0)
    sync = train.FlatGradSync(net)
    opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr=0.1)
    f32 = lambda b: {k: v.float() for k, v in b.items()}
    loss = train.optimizer_step(net, opt, [f32(_batch(rank, s)) for s in range(accum)], lambda m, b: m(**b),
                                accum_steps=accum, grad_clip=10.0, grad_sync=sync)
    if rank == 0:
        torch.save({"state": {k: v.clone() for k, v in net.state_dict().items()}, "loss": loss,
                    "views": all(p.grad is None or p.grad.data_ptr() == sync.slot(p).data_ptr() for p in sync.big)}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("accum", [1, 2])
def test_two_rank_flat_grad_sync_matches_single_process(tmp_path, accum):
    """FlatGradSync (one all-reduce of a flat gradient buffer after backward, mmgl_b200/train.py) gives the same update as
    one process running the union of both ranks' micro-batches, from rank 0's initial parameters."""
    world, port = 2, _free_port()
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker_flat, args=(world, port, accum, out), nprocs=world, join=True)
    got = torch.load(out)
    ref = Toy32()
    opt = torch.optim.SGD([p for p in ref.parameters() if p.requires_grad], lr=0.1)
    total = 0.0
    for s in range(accum):
        for r in range(world):
            b = {k: v.float() for k, v in _batch(r, s).items()}
            loss = ref(**b) / (accum * world)
            loss.backward()
            total += float(loss) if r == 0 else 0.0
    torch.nn.utils.clip_grad_norm_([p for p in ref.parameters() if p.requires_grad], 10.0)
    opt.step()
    for k, v in ref.state_dict().items():
        assert torch.allclose(got["state"][k], v, rtol=1e-5, atol=1e-6), k


@pytest.mark.parametrize("accum", [1, 3])
def test_two_rank_ddp_matches_single_process(tmp_path, accum):
    world, port = 2, _free_port()
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, port, accum, out), nprocs=world, join=True)
    got = torch.load(out, weights_only=False)
    assert got["max"] == 2.0
    # single process over the union of both ranks' micro-batches
    ref = Toy()
    opt = torch.optim.SGD([p for p in ref.parameters() if p.requires_grad], lr=0.1)
    for s in range(accum):
        for r in range(world):
            (ref(**_batch(r, s)) / (accum * world)).backward()
    torch.nn.utils.clip_grad_norm_([p for p in ref.parameters() if p.requires_grad], 10.0)
    opt.step()
    for k, v in ref.state_dict().items():
        torch.testing.assert_close(got["state"][k], v, rtol=1e-12, atol=1e-12)
    # ranks drew different data
    assert not torch.equal(_batch(0, 0)["x"], _batch(1, 0)["x"])


def test_synthetic_batches_are_rank_disjoint_and_well_formed():
    from mmgl_b200 import synth
    spec = synth.BatchSpec(batch=3, max_input_length=32, max_output_length=16, text_neighbors=4, image_neighbors=2,
                           image_size=8, with_lpe=True, with_graph=True)
    b0, b1 = synth.make_batch(spec, 1234), synth.make_batch(spec, 1334)
    assert not torch.equal(b0["input_ids"], b1["input_ids"])
    assert torch.equal(synth.make_batch(spec, 1234)["input_ids"], b0["input_ids"]), "seeded generation must repeat"
    n = spec.text_neighbors + spec.image_neighbors
    for b in (b0, b1):
        assert b["input_ids"].shape == (3, 48) and b["labels"].shape == (3, 48)
        locs = torch.cat((b["text_locations"], b["image_locations"]), 1).sort(1).values
        assert torch.equal(locs, torch.arange(n).expand(3, n)), "locations must be a permutation of the bank slots"
        # valid neighbors occupy the first slots, padding the last (wikiweb2m/data.py:349-454)
        valid = torch.zeros(3, n, dtype=torch.bool)
        valid.scatter_(1, b["text_locations"], b["neighbor_pos_ids"] > 0)
        valid.scatter_(1, b["image_locations"], b["neighbor_images_pos_ids"] > 0)
        cnt = valid.sum(1)
        assert all(bool(valid[r, :cnt[r]].all()) and not bool(valid[r, cnt[r]:].any()) for r in range(3))
        assert b["lpe"].shape == (3, n + 1, n - 4) and b["graph"].shape == (3, n + 1, n + 1)
        rows = b["graph"].sum(-1)
        assert bool(((rows - 1).abs() < 1e-5).logical_or(rows == 0).all()), "graph rows are normalised or empty"
        assert (b["attention_mask"].sum(1) >= 2).all()
