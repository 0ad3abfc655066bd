#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): the slab-decomposed application with the NCCL halo
exchange must reproduce, bit for bit, the single-GPU application of the same global problem computed on every rank."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import deo_b200 as D
from deo_b200 import _lib
from deo_b200.dist import SlabContext, SlabPlan, extended_slab

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
_lib.check(D.load_library().deo_init(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = SlabContext.from_torch_distributed(device=local)
ROBIN = ((1.0, 0.5, 0.25), (1.0, -0.5, 0.75))
fails = 0
for name, shape, a, dtype, reps in [("fused-launch", (256, 192, 160 * world), 4, np.float64, 3), ("three-launch (thin slabs)", (128, 96, 40 * world), 4, np.float64, 2),
                                    ("fused f32 a=6", (256, 128, 144 * world), 6, np.float32, 2), ("table (upwind + nonuniform)", (128, 96, 150 * world), 4, np.float64, 2),
                                    ("shifted tiles (2^k + 1 extents)", (130, 65, 160 * world), 4, np.float64, 2),
                                    ("odd row length (cp.async loader)", (129, 66, 120 * world), 4, np.float64, 2)]:
    h = tuple(1.0 / (s + 1) for s in shape)
    u = np.asfortranarray(np.random.default_rng(5).uniform(-1, 1, shape).astype(dtype))
    if name.startswith("table"):
        dxs = [(hh * (1 + 0.3 * np.sin(2 * np.pi * np.arange(1, s + 2) / (s + 1)))).astype(dtype) for s, hh in zip(shape, h)]
        c = np.sin(6 * np.pi * np.arange(1, shape[2] + 1) / shape[2]).astype(dtype)
        A = D.CenteredDifference[1](2, a, dxs[0], shape[0]) + D.CenteredDifference[2](2, a, dxs[1], shape[1]) + \
            D.CenteredDifference[3](2, a, dxs[2], shape[2]) + D.UpwindDifference[3](1, 2, dxs[2], shape[2], c)
        Q = D.compose(*D.RobinBC(*ROBIN, dxs, 1, shape, dtype=dtype))
    else:
        A = D.CenteredDifference[1](2, a, h[0], shape[0], dtype=dtype) + D.CenteredDifference[2](2, a, h[1], shape[1], dtype=dtype) + \
            D.CenteredDifference[3](2, a, h[2], shape[2], dtype=dtype)
        Q = D.compose(*D.RobinBC(*ROBIN, h, 1, shape, dtype=dtype))
    G = A * Q
    want = ((G) * D.DeviceArray.from_host(u)).to_host()                      # whole problem on this GPU
    plan = SlabPlan(G, shape, dtype, ctx=ctx)
    ext = extended_slab(u, rank, world, plan.halo, fill=np.nan)               # halos poisoned: they must come from the exchange
    ud = D.DeviceArray.from_host(ext)
    dud = D.DeviceArray(plan.local_out_shape, dtype)
    for _ in range(reps):                                                     # repeated applications reuse the flag / events
        plan.apply(dud, ud)
    D.sync()
    got = dud.to_host()
    ok = np.array_equal(got, want[..., plan.start:plan.start + plan.count])
    t = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print(f"{name:32s} shape={shape} {np.dtype(dtype).name} kernel={plan.info[0]} ranks={world}: {'OK (bitwise)' if t.item() == 0 else 'MISMATCH'}", flush=True)
    fails += int(t.item())
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if fails else 0)
