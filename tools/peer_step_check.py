"""Run under torchrun (N >= 2 GPUs): gradient all-reduce + clip + Adam as ONE kernel over NVLink peer memory
(PeerFusedPoseNetAdam.step_allreduce -> odevio_allreduce_adam_peer) against NCCL all-reduce + the three-launch step
(FusedPoseNetAdam) from the same initial state with different per-rank gradients; prints the largest parameter deviation,
the time per call of both, and PEER_STEP_OK."""
import copy
import os

import torch
import torch.distributed as dist

import odevio_b200
from odevio_b200.distributed import pose_net_params
from odevio_b200.training import FusedPoseNetAdam, PeerFusedPoseNetAdam
from oracle.pose_odernn import default_opt

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
a = odevio_b200.PoseODERNN(default_opt()).to(dev)
b = copy.deepcopy(a)
peer = PeerFusedPoseNetAdam(a, lr=1e-3)
base = FusedPoseNetAdam(b, lr=1e-3)
g = torch.Generator().manual_seed(100 + rank)          # every rank its own gradients
worst = 0.0
for step in range(5):
    scale = 50.0 if step == 2 else 0.01
    for pa, pb in zip(pose_net_params(a), pose_net_params(b)):
        gr = (scale * torch.randn(pa.shape, generator=g) / pa.numel() ** 0.5).to(dev)
        pa.grad, pb.grad = gr.clone(), gr.clone()
    peer.gather_grads(); peer.step_allreduce()
    flat = base.gather_grads(1.0 / world)
    dist.all_reduce(flat); base.step()
    torch.cuda.synchronize()
    assert abs(peer.norm_coef[0].item() - base.norm_coef[0].item()) <= 1e-5 * base.norm_coef[0].item(), step
    for pa, pb in zip(pose_net_params(a), pose_net_params(b)):
        worst = max(worst, ((pa - pb).abs().max() / pb.abs().max().clamp_min(1e-12)).item())
# every rank ends with the same parameters
chk = peer.flat.clone()
dist.broadcast(chk, 0)
assert torch.equal(chk, peer.flat), "ranks diverged"
assert worst <= 1e-5, worst
# timing (gradients already gathered)
def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t_peer = timed(peer.step_allreduce)
def nccl_step():
    dist.all_reduce(base.grads); base.step()
t_nccl = timed(nccl_step)
if rank == 0:
    print(f"world {world}: worst relative parameter deviation {worst:.2e}; one-kernel peer step {t_peer * 1e3:.1f} us, "
          f"NCCL all-reduce + 3 launches {t_nccl * 1e3:.1f} us (bucket {peer.flat.numel() * 4 / 1e6:.1f} MB)")
    print("PEER_STEP_OK")
dist.destroy_process_group()
