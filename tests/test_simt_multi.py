"""Several ranks on the CPU: the multi-GPU path of the library (Y slabs, particle faces with
far movers, rho/phi ghost rows, distributed FFT, capacity agreement; cpic_b200/csrc/comm.cu)
with the kernels run by the SIMT interpreter of tests/simt and NCCL replaced by
tests/simt/fake_nccl.c (files between processes). Same worker and same assertions against the
single-rank oracle as tests/test_gpu_multi.py runs on real GPUs; torch.distributed (gloo)
only carries the 128-byte id. Test infrastructure only -- see tests/test_simt_check.py."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import conf_path, ROOT
from test_simt_check import simt_build, SIMT  # noqa: F401  (fixture)


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_ranks_cpu(lib, n, conf, steps, mode="fused", env=None, timeout=600):
    base = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()), WORLD_SIZE=str(n),
                MGPU_DEVICE="cpu", CPIC_B200_LIB=lib, CPIC_B200_SIMT_CHECK="1",
                CPIC_B200_NCCL=os.path.join(SIMT, "_build", "libfake_nccl.so"), **(env or {}))
    procs = []
    for r in range(n):
        e = dict(base, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mgpu_worker.py"),
                                       conf_path(conf), str(steps), mode], env=e, cwd=ROOT,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=timeout)[0])
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    bad = [i for i, p in enumerate(procs) if p.returncode != 0]
    assert not bad and "MGPU-OK" in outs[0], "\n".join(f"--- rank {i}\n{o[-2500:]}" for i, o in enumerate(outs))
    return outs[0]


@pytest.mark.parametrize("conf,mode", [("uniform-small.conf", "fused"), ("2d-2species-small.conf", "staged"),
                                       ("two-streams.conf", "fused"), ("far-beam.conf", "fused")])
def test_two_ranks_on_cpu(simt_build, conf, mode):
    run_ranks_cpu(simt_build, 2, conf, 10, mode)


def test_four_ranks_on_cpu(simt_build):
    run_ranks_cpu(simt_build, 4, "uniform-small.conf", 10)


def test_capacity_growth_is_agreed_between_ranks_on_cpu(simt_build):
    out = run_ranks_cpu(simt_build, 2, "uniform-small.conf", 40, env={"MGPU_TIGHT": "1"})
    assert "MGPU-CAPS" in out


def test_streamed_initialisation_on_two_ranks_on_cpu(simt_build):
    run_ranks_cpu(simt_build, 2, "2d-2species-small.conf", 6, env={"MGPU_STREAMED": "1000"})


def test_run_and_run_timed_on_two_ranks_on_cpu(simt_build):
    """cpic_b200_run / cpic_b200_run_timed as bench.py calls them with several ranks: 70 steps, so that
    the collective capacity check inside the library (every 32 steps) runs twice, with tight
    capacities so that it has something to agree on."""
    out = run_ranks_cpu(simt_build, 2, "uniform-small.conf", 70, env={"MGPU_RUN": "1", "MGPU_TIGHT": "1"}, timeout=900)
    assert "MGPU-CAPS" in out



def test_bench_script_on_two_ranks_on_cpu(simt_build):
    """bench.py with two ranks at the tiny size: the multi_rank_check of the line (N ranks against one rank,
    product only) and the weak-scaling workload over the multi-rank path."""
    import json
    base = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()), WORLD_SIZE="2", BENCH_TINY="1",
                CPIC_B200_LIB=simt_build, CPIC_B200_SIMT_CHECK="1",
                CPIC_B200_NCCL=os.path.join(SIMT, "_build", "libfake_nccl.so"))
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "3", "--warmup", "3",
                               "--no-e2e"], env=dict(base, RANK=str(r), LOCAL_RANK=str(r)), cwd=ROOT,
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=900) for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[1][-2000:] for o in outs)
    d = json.loads(outs[0][0].strip().splitlines()[-1])
    assert d["n_gpus"] == 2 and d["multi_rank_check"]["ok"] and d["multi_rank_check"]["worst"] <= 1e-12
    assert d["multi_rank_check"]["count_conserved"] and d["multi_rank_check"]["particles_on_another_rank_than_at_start"] > 0
    assert outs[1][0].strip() == ""          # rank 0 alone prints
