"""N>1 host logic on CPU: two processes over gloo.  Covers what the host runtime does around the library
for slab-decomposed plans -- NCCL unique-id hand-over, slab bookkeeping, scatter, the halo-exchange
protocol (same plane offsets as csrc/dist.cu) and gather -- and checks the reassembled result of applying
the oracle slab by slab equals the oracle on the global array."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import deo_b200  # noqa: F401
        from deo_b200 import dist as DD
        from oracle import oracle as O

        # 1. the 128-byte NCCL id reaches every rank unchanged
        uid = DD.broadcast_unique_id()
        assert len(uid) == 128

        # 2. scatter + halo exchange reproduce the [halo | own | halo] blocks
        shape, halo = (12, 10, 23), 2
        u = np.asfortranarray(np.random.default_rng(7).uniform(-1, 1, shape)) if rank == 0 else None
        ext = DD.scatter_slabs(u, shape, np.float64, halo)
        s, c = DD.slab_bounds(shape[-1], world, rank)
        assert ext.shape == shape[:-1] + (c + 2 * halo,)
        ug = np.asfortranarray(np.random.default_rng(7).uniform(-1, 1, shape))   # same seed: every rank can check
        assert np.array_equal(ext[..., halo:halo + c], ug[..., s:s + c])
        wiped = ext.copy(order="F")
        wiped[..., :halo] = np.nan
        wiped[..., halo + c:] = np.nan
        DD.exchange_halos_host(wiped, halo)
        lo, hi = max(s - halo, 0), min(s + c + halo, shape[-1])
        assert np.array_equal(wiped[..., lo - (s - halo):hi - (s - halo)], ug[..., lo:hi])

        # 3. slab-by-slab application (oracle as the stand-in for the kernel) reassembles to the global result:
        #    interior slabs need exactly `halo` planes from their neighbours and nothing else
        h = tuple(1.0 / (n + 1) for n in shape)
        ops = [O.CenteredDifference(2, 4, h[a], shape[a], axis=a + 1) for a in range(3)]
        bcs = {a + 1: O.RobinBC((1.0, 0.5, 0.25), (1.0, -0.5, 0.75), h[a], 1) for a in range(3)}
        want = O.apply_sum(ops, ug, bcs)
        local = want[..., s:s + c].copy(order="F")       # what this rank's kernel writes
        got = DD.gather_slabs(local, shape)
        if rank == 0:
            assert np.array_equal(got, want)
        # locality of the stencil: du on [s, s+c) depends only on planes [s-halo, s+c+halo)
        pert = ug.copy(order="F")
        if s - halo - 1 >= 0:
            pert[..., : s - halo] += 1.0
        if s + c + halo < shape[-1]:
            pert[..., s + c + halo:] -= 1.0
        assert np.array_equal(O.apply_sum(ops, pert, bcs)[..., s:s + c], local)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + "".join(traceback.format_exception(e))))


@pytest.mark.timeout(300)
def test_world_size_2_gloo_slab_host_logic():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"
