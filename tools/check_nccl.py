"""Multi-GPU parity: points sharded over WORLD_SIZE ranks + one fused NCCL allreduce per
evaluation must reproduce the single-GPU gradient/loss and keep replicas bit-identical.
launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_nccl.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pinn_based_online_pde_calculator_b200 import PinnEngine, shard_range
from pinn_based_online_pde_calculator_b200.workloads import init_params, make_points, make_workload

rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr_}"))
wl = make_workload("C2", n_col=400_003)
x_col, x_bd, u_bd = make_points(wl)  # identical on every rank (rank=0 seed)
eng = PinnEngine(wl.net, wl.eq, n_bc=4, device=lr_)
idt = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{lr_}")
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(PinnEngine.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
eng.init_nccl(bytes(idt.cpu().numpy().tobytes()), rank, world)
eng.set_params(init_params(wl.net))
b, e = shard_range(wl.n_col, rank, world)
sp = [shard_range(len(a), rank, world) for a in x_bd]
eng.set_points(x_col[b:e], [a[s:t] for a, (s, t) in zip(x_bd, sp)], [a[s:t] for a, (s, t) in zip(u_bd, sp)])
eng.set_global_counts(wl.n_col, [len(a) for a in x_bd])
eng.set_loss(1.0, 2.0)
g, info = eng.loss_grad()
eng.adam_init()
rows = eng.adam_steps(5, 1e-3)
# L-BFGS on the sharded problem: gradient and loss sums allreduced per evaluation, the two-loop recursion and the line
# search run replicated on identical numbers (no further collective)
res_lb, rows_lb = eng.lbfgs(15, 1e-12)
p = torch.as_tensor(eng.get_params()).cuda()
# replicas stay bit-identical
pl = [torch.empty_like(p) for _ in range(world)]
dist.all_gather(pl, p)
same = all(torch.equal(pl[0], q) for q in pl)
if rank == 0:
    ref = PinnEngine(wl.net, wl.eq, n_bc=4, device=lr_)
    ref.set_params(init_params(wl.net))
    ref.set_points(x_col, x_bd, u_bd)
    ref.set_loss(1.0, 2.0)
    g1, info1 = ref.loss_grad()
    ref.adam_init()
    rows1 = ref.adam_steps(5, 1e-3)
    res1, rows_lb1 = ref.lbfgs(15, 1e-12)
    lb_rel = abs(res_lb["final_loss"] / res1["final_loss"] - 1)
    print(f"world={world} L-BFGS: {res_lb['iterations']} iterations / {res_lb['evaluations']} evaluations, final loss {res_lb['final_loss']:.6e} "
          f"(single GPU: {res1['iterations']} / {res1['evaluations']}, {res1['final_loss']:.6e}, rel {lb_rel:.2e}), mode syncs {eng.lbfgs_host_syncs()}")
    assert lb_rel < 5e-3 and res_lb["final_loss"] < rows[-1, 0]
    rel = float((g - g1).norm() / g1.norm())
    print(f"world={world} grad rel err vs single GPU: {rel:.3e}; loss_info rel: {np.abs(info / info1 - 1).max():.3e}; "
          f"adam loss rows rel: {np.abs(rows[:, 0] / rows1[:, 0] - 1).max():.3e}; replicas identical: {same}")
    assert rel < 1e-5 and same
dist.barrier()
dist.destroy_process_group()
