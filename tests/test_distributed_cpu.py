"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row sharding, the flat-bucket
gradient all-reduce and the sharded training step.  The model on each rank is the CPU oracle
(allowed in tests only); the logic under test is odevio_b200.distributed, which the GPU path uses
unchanged with the nccl backend."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _small_opt(**over):
    from oracle.pose_odernn import default_opt
    return default_opt(v_f_len=24, i_f_len=8, ode_hidden_dim=16, ode_fn_num_layers=2, ode_solver="rk4", **over)


def _data(B, S=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    fv, fi = torch.randn(B, S, 24, generator=g), torch.randn(B, S, 8, generator=g)
    ts = torch.arange(S + 1, dtype=torch.float32).repeat(B, 1) * 0.1
    gts = 0.1 * torch.randn(B, S, 6, generator=g)
    return fv, fi, ts, gts


def _worker(rank, world, port, B, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.pose_odernn import OraclePoseODERNN
        from odevio_b200 import distributed as D
        torch.manual_seed(0)
        torch.set_num_threads(1)
        model = OraclePoseODERNN(_small_opt())
        fv, fi, ts, gts = _data(B)
        a, b = D.shard_rows(B, rank, world)
        opt = D.make_optimizer(model, lr=1e-2)
        # (1) gradient exchange: weighted flat-bucket all-reduce == global-batch gradient
        poses, _ = model(fv[a:b], fi[a:b], ts[a:b])
        D.pose_loss(poses, gts[a:b]).backward()
        n = D.allreduce_pose_net_grads(model, world, shard_rows_count=b - a, global_rows=B)
        grads = {k: p.grad.clone() for k, p in model.named_parameters()}
        # (2) one full sharded step (equal shards path when B % world == 0)
        if B % world == 0:
            loss = D.train_step(model, opt, fv[a:b], fi[a:b], ts[a:b], gts[a:b], world_size=world)
        torch.save({"grads": grads, "n": n, "state": model.state_dict(), "rows": (a, b)},
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])
def test_two_rank_gradient_allreduce_equals_big_batch(tmp_path, B):
    from oracle.pose_odernn import OraclePoseODERNN
    from odevio_b200 import distributed as D
    port = _free_port()
    mp.spawn(_worker, args=(2, port, B, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    assert r0["rows"][0] == 0 and r0["rows"][1] == r1["rows"][0] and r1["rows"][1] == B
    # single-process reference on the full batch
    torch.manual_seed(0)
    torch.set_num_threads(1)
    model = OraclePoseODERNN(_small_opt())
    fv, fi, ts, gts = _data(B)
    opt = D.make_optimizer(model, lr=1e-2)
    poses, _ = model(fv, fi, ts)
    D.pose_loss(poses, gts).backward()
    n_params = sum(p.numel() for p in D.pose_net_params(model))
    assert r0["n"] == r1["n"] == n_params
    for k, p in model.named_parameters():
        assert torch.allclose(r0["grads"][k], p.grad, rtol=1e-4, atol=1e-7), k
        assert torch.equal(r0["grads"][k], r1["grads"][k]), k          # ranks agree bit-for-bit
    if B % 2 == 0:
        D.train_step(model, opt, fv, fi, ts, gts, world_size=1)
        for k, v in model.state_dict().items():
            assert torch.allclose(r0["state"][k], v, rtol=1e-4, atol=1e-6), k
            assert torch.equal(r0["state"][k], r1["state"][k]), k


def test_shard_rows_partition():
    from odevio_b200.distributed import shard_rows
    for B in (0, 1, 7, 8, 1024, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_rows(8, 2, 2)
