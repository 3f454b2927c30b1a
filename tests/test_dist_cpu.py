"""Host-side logic of the multi-GPU path on CPU: two processes over gloo (no GPU needed).
Covers the slab partition (particle_comm_initial between ranks, src/particle.h:19-20), the
bootstrap of the library's NCCL communicator id, and the per-rank parameters."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import conf_path, ROOT


def _worker(rank, world, port, conf, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cpic_b200
        from cpic_b200 import load_conf, init_particles
        from cpic_b200.dist import env_rank, partition, slab_rank, broadcast_id
        assert env_rank() == (rank, world, rank)
        params, run = load_conf(conf, rank=rank, nranks=world, device=rank)
        assert (params.rank, params.nranks) == (rank, world) and params.ny % world == 0
        parts = init_particles(conf)                 # every rank generates the same global population
        mine = partition(parts, params, rank)
        # 1. every particle is owned by exactly one rank, and lies in that rank's slab
        counts = torch.tensor([len(p["id"]) for p in mine], dtype=torch.int64)
        total = counts.clone()
        dist.all_reduce(total)
        assert total.tolist() == run.nparticles
        Ls = params.Ly / world
        for p in mine:
            assert ((p["y"] >= rank * Ls) & (p["y"] < (rank + 1) * Ls + 1e-12)).all()
            assert (slab_rank(params, p["y"]) == rank).all()
        ids = [torch.zeros(int(total[0]), dtype=torch.bool) for _ in range(1)]
        ids[0][torch.from_numpy(mine[0]["id"])] = True
        seen = ids[0].to(torch.int32)
        dist.all_reduce(seen)
        assert (seen == 1).all()
        # 2. the NCCL id made on rank 0 reaches every rank unchanged
        import ctypes as C
        ident = bytes(128)
        if rank == 0:
            buf = (C.c_char * 128)()
            rc = cpic_b200.lib().cpic_b200_comm_id(buf)
            ident = bytes(buf) if rc == 0 else bytes(range(128))
        got = broadcast_id(ident, dist)
        gathered = [None] * world
        dist.all_gather_object(gathered, got)
        assert all(g == gathered[0] for g in gathered) and any(gathered[0])
        q.put((rank, "ok", counts.tolist()))
    except Exception as e:                           # noqa: BLE001
        q.put((rank, f"{type(e).__name__}: {e}", None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("conf", ["uniform-small.conf", "2d-2species-small.conf"])
def test_two_ranks_gloo(conf):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 100)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, conf_path(conf), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
    assert all(min(r[2]) > 0 for r in res)          # both slabs hold particles
