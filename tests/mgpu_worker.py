"""One rank of the multi-GPU parity test (launched by torchrun from test_gpu_multi.py):
runs `steps` sim_steps of a conf on WORLD_SIZE GPUs and checks this rank's slab against the
single-rank oracle, which every rank computes for itself."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import torch
import torch.distributed as dist

from cpic_b200 import Sim, load_conf, init_particles
from cpic_b200.dist import env_rank, partition, bootstrap, set_particles_collective
from _parity import oracle_from, relerr, TOL


def main():
    conf, steps = sys.argv[1], int(sys.argv[2])
    fused = len(sys.argv) < 4 or sys.argv[3] == "fused"
    rank, world, local = env_rank()
    # MGPU_DEVICE=cpu: the ranks are CPU processes running the kernels under tests/simt (gloo carries
    # the id, tests/simt/fake_nccl.c the library's traffic); default: one GPU per rank over NCCL
    dev = os.environ.get("MGPU_DEVICE", "cuda")
    if dev == "cuda":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    if conf.startswith("random:"):
        # a random configuration of tests/test_gpu_parity.py::_random_case, the same on every rank
        from test_gpu_zz_robustness import _random_case
        params, parts, _, _ = _random_case(int(conf.split(":")[1]))
        params.rank, params.nranks, params.device = rank, world, local
        if params.ny % world or (params.ny // world) % max(params.block_cells, 1):
            if rank == 0:
                print(f"MGPU-OK world={world} conf={conf} skipped (slab height)")
            dist.destroy_process_group()
            return
    else:
        params, run = load_conf(conf, rank=rank, nranks=world, device=local)
        parts = init_particles(conf)
    if os.environ.get("MGPU_TIGHT"):
        params.capacity_factor = 1.01        # forces the collective capacity growth (check_capacity)
    o = oracle_from(params, parts)
    if os.environ.get("MGPU_STREAMED"):
        # sim_init with the streamed initialisation: every rank draws the whole population in batches,
        # tallies all slabs (one capacity everywhere, no communication) and keeps its own
        g = Sim.from_conf(conf, rank=rank, nranks=world, device=local, stream_batch=int(os.environ["MGPU_STREAMED"]))
    else:
        g = Sim(params)
        set_particles_collective(g, partition(parts, params, rank), dist, device=dev)
    bootstrap(g, dist, device=dev)
    o.pre_step()
    g.pre_step()
    nyl = params.ny // world
    r0 = rank * nyl
    worst = {}

    def check(tag):
        g.sync()
        e = {}

        def slab_err(mine, ref_slab, ref_all):
            # max|a - b| over this rank's rows against max|b| over the WHOLE field: a slab
            # whose own values happen to be small must not inflate the measure
            scale = np.abs(ref_all).max() if ref_all.size else 0.0
            d = np.abs(np.asarray(mine) - np.asarray(ref_slab)).max() if ref_slab.size else 0.0
            return d / scale if scale > 0 else d

        e["rho"] = slab_err(g.field("rho"), o.field("rho")[r0:r0 + nyl], o.field("rho"))
        og = o.field("phi_ghost")          # rows: -1, 0 .. ny-1, ny, ny+1 (periodic images)
        rows = [(r0 - 1 + k) % params.ny for k in range(nyl + 3)]
        e["phi"] = slab_err(g.field("phi_ghost"), og[1:params.ny + 1][rows], og)
        e["Ex"] = slab_err(g.field("Ex"), o.field("Ex")[r0:r0 + nyl + 1], o.field("Ex"))
        e["Ey"] = slab_err(g.field("Ey"), o.field("Ey")[r0:r0 + nyl + 1], o.field("Ey"))
        from cpic_b200.dist import slab_rank
        for i in range(len(params.q)):
            a, b = g.particles(i), o.particles(i)
            sel = slab_rank(params, b["y"]) == rank
            ids = b["id"][sel]
            assert len(a["id"]) == len(ids) and (a["id"] == ids).all(), \
                f"{tag}: rank {rank} species {i} holds {len(a['id'])} particles, the oracle puts {len(ids)} in its slab"
            umax = max(np.abs(b["ux"]).max(initial=0), np.abs(b["uy"]).max(initial=0), 1e-300)
            e[f"x{i}"] = np.abs(a["x"] - b["x"][sel]).max(initial=0) / params.Lx
            e[f"y{i}"] = np.abs(a["y"] - b["y"][sel]).max(initial=0) / params.Ly
            e[f"ux{i}"] = np.abs(a["ux"] - b["ux"][sel]).max(initial=0) / umax
            e[f"uy{i}"] = np.abs(a["uy"] - b["uy"][sel]).max(initial=0) / umax
        bad = {k: v for k, v in e.items() if not v <= TOL}
        assert not bad, f"{tag}: rank {rank}: beyond {TOL}: {bad}"
        for k, v in e.items():
            worst[k] = max(worst.get(k, 0), v)

    check("after sim_init")
    if os.environ.get("MGPU_RUN"):
        # the bench's way: cpic_b200_run / run_timed over all the steps (capacity check agreed over the
        # ranks every 32 steps inside the library), compared once at the end
        half = steps // 2
        g.run(half)
        g.run_timed(steps - half)
        for it in range(steps):
            o.step()
        check(f"after {steps} steps of cpic_b200_run")
        steps = 0
    for it in range(steps):
        if fused:
            g.step()
        else:
            g.step_staged()
        o.step()
        check(f"iteration {it}")
    t = torch.tensor([max(worst.values())], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    caps = [g.capacity(i) for i in range(len(params.q))]
    ct = torch.tensor(caps, device=dev, dtype=torch.int64)
    cmax, cmin = ct.clone(), ct.clone()
    dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(cmin, op=dist.ReduceOp.MIN)
    assert torch.equal(cmax, cmin), f"block capacities differ between ranks: {caps}"
    if rank == 0:
        print(f"MGPU-CAPS {caps}")
        print(f"MGPU-OK world={world} conf={os.path.basename(conf)} steps={steps} worst={t.item():.2e}")
    g.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
