"""Host-side multi-GPU logic on CPU: world_size 2 over gloo (one process per rank, rendezvous on 127.0.0.1)."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pu3 = importlib.import_module("3pu_pytorch_b200")
        D = pu3.dist
        out = {}
        # 1. sharded eval with an uneven split and gather (the net is a stand-in: scaling + 2x replication)
        net = lambda x, ratio=None: torch.cat([x * 2, x * 3], dim=2)
        patches = torch.arange(5 * 3 * 4, dtype=torch.float32).view(5, 3, 4)
        lo, hi, part = D.upsample_sharded(net, patches, ratio=2, gather=False)
        out["range"] = (lo, hi, tuple(part.shape))
        _, _, full = D.upsample_sharded(net, patches, ratio=2, gather=True)
        out["gather_ok"] = bool(torch.equal(full, net(patches)))
        # 2. gradient all-reduce on the flat buffer == gradient of the whole batch
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Conv1d(3, 8, 1), torch.nn.ReLU(), torch.nn.Conv1d(8, 3, 1))
        ref = torch.nn.Sequential(torch.nn.Conv1d(3, 8, 1), torch.nn.ReLU(), torch.nn.Conv1d(8, 3, 1))
        ref.load_state_dict(model.state_dict())
        x = torch.randn(6, 3, 10, generator=torch.Generator().manual_seed(1))
        opt = D.FlatAdam(model)
        opt.zero_grad()
        l, h = D.shard_range(6, rank, world)
        model(x[l:h]).square().mean().backward()
        scale = opt.reduce_gradients()
        ref(x).square().mean().backward()
        # segments of the flat buffer are padded to 128 bytes: compare through the per-parameter views
        out["grad_err"] = max(float((p.grad * scale - r.grad).abs().max())
                              for p, r in zip(model.parameters(), ref.parameters()))
        out["views"] = bool(all(p.grad.data_ptr() >= opt.flat_grad.data_ptr() for p in model.parameters()))
        # a rank-local optimizer inside an initialised job never enters a collective (bench.py side legs)
        local = D.FlatAdam(torch.nn.Conv1d(3, 4, 1), data_parallel=False)
        out["local_scale"] = local.reduce_gradients() if rank == 0 else 1.0     # only ONE rank calls it: must not hang
        try:
            opt.step()
            out["cpu_step"] = "ran"
        except RuntimeError as e:
            out["cpu_step"] = "raised" if "CUDA" in str(e) else str(e)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0]["range"] == (0, 3, (3, 3, 8)) and results[1]["range"] == (3, 5, (2, 3, 8))
    for r in range(world):
        assert results[r]["gather_ok"] and results[r]["views"]
        assert results[r]["grad_err"] < 1e-6
        assert results[r]["local_scale"] == 1.0
        assert results[r]["cpu_step"] == "raised"     # the optimizer kernel has no CPU fallback


def test_shard_range_covers_everything():
    pu3 = importlib.import_module("3pu_pytorch_b200")
    for total in (0, 1, 7, 32, 48, 256):
        for world in (1, 2, 3, 8):
            spans = [pu3.dist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1
