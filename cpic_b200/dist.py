"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed only carries the
128-byte NCCL id to the ranks; the data path is the library's own NCCL traffic
(cpic_b200/csrc/comm.cu). Mirrors the reference's rank layout: Y slabs of ny/P rows
(src/sim.c:177-187, src/field.c:145-155)."""
import os

import numpy as np


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def slab_rank(params, y):
    """Rank owning a Y coordinate: row = floor(y/dy) clamped, rank = row // (ny/P)."""
    dy = params.Ly / params.ny
    row = np.floor(np.asarray(y) * (1.0 / dy)).astype(np.int64)
    row = np.clip(row, 0, params.ny - 1)
    return row // (params.ny // params.nranks)


def partition(parts, params, rank):
    """particle_comm_initial between ranks (src/particle.h:19-20): the particles of every
    species that fall in this rank's slab, input order kept."""
    out = []
    for p in parts:
        sel = slab_rank(params, p["y"]) == rank
        out.append({k: np.ascontiguousarray(v[sel]) for k, v in p.items()})
    return out


def broadcast_id(id_bytes, dist, device=None):
    """Rank 0's id to everybody over an initialised torch.distributed group (nccl or gloo)."""
    import torch
    t = torch.zeros(128, dtype=torch.uint8, device=device)
    if dist.get_rank() == 0:
        t.copy_(torch.frombuffer(bytearray(id_bytes), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())


def bootstrap(sim, dist, device=None):
    """ncclCommInitRank on every rank with rank 0's unique id."""
    ident = sim.comm_id() if dist.get_rank() == 0 else bytes(128)
    sim.comm_init(broadcast_id(ident, dist, device))


def set_particles_collective(sim, parts, dist, device=None):
    """set_particles on every rank with one block capacity for all of them (the particle
    exchange buffers are sized from it, so the ranks must agree)."""
    import torch
    for i, p in enumerate(parts):
        sim.set_particles(i, p["id"], p["x"], p["y"], p["ux"], p["uy"], p.get("uz"))
        cap = torch.tensor([sim.capacity(i)], dtype=torch.int64, device=device)
        dist.all_reduce(cap, op=dist.ReduceOp.MAX)
        if int(cap.item()) != sim.capacity(i):
            sim.reserve(i, int(cap.item()))
            sim.set_particles(i, p["id"], p["x"], p["y"], p["ux"], p["uy"], p.get("uz"))
