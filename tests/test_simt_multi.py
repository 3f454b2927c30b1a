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


def run_ranks_cpu(lib, n, conf, steps, mode="fused", env=None, timeout=600, raw=False):
    base = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()), WORLD_SIZE=str(n),
                MGPU_DEVICE="cpu", CPIC_B200_LIB=lib, CPIC_B200_SIMT_CHECK="1",
                CPIC_B200_NCCL=os.path.join(SIMT, "_build", "libfake_nccl.so"), **(env or {}))
    procs = []
    for r in range(n):
        e = dict(base, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mgpu_worker.py"),
                                       conf if raw else conf_path(conf), str(steps), mode], env=e, cwd=ROOT,
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


def _run_driver(lib, conf, rank=None, nranks=1, extra_env=None):
    code = ("import ctypes as C, sys; L = C.CDLL(sys.argv[1]); "
            "argv = (C.c_char_p * 3)(b'cpic', b'-q', sys.argv[2].encode()); sys.exit(L.cpic_b200_main(3, argv))")
    env = dict(os.environ, CPIC_B200_SIMT_CHECK="1", CPIC_B200_NCCL=os.path.join(SIMT, "_build", "libfake_nccl.so"), **(extra_env or {}))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    if rank is not None:
        env.update(CPIC_B200_RANK=str(rank), CPIC_B200_NRANKS=str(nranks), CPIC_B200_DEVICE="0")
    return subprocess.Popen([sys.executable, "-c", code, lib, conf], cwd=ROOT, env=env, stdout=subprocess.PIPE,
                            stderr=subprocess.PIPE, text=True)


def test_own_driver_on_two_ranks_on_cpu(simt_build, tmp_path):
    """`mpirun -n 2 cpic <conf>` as two processes of cpic_b200_main (one per GPU; here the interpreter): the
    reference's initial conditions drawn per slab, the communicator id through a file, and at the end the
    energies of the two slabs add up to those of the single-rank run."""
    import re
    text = open(conf_path("2d-2species-small.conf")).read()
    text = re.sub(r"cycles\s*=\s*\d+", "cycles = 6", text)
    conf = tmp_path / "run.conf"
    conf.write_text(text)
    one = _run_driver(simt_build, str(conf))
    out1, err1 = one.communicate(timeout=600)
    assert one.returncode == 0, err1[-2000:]
    idf = str(tmp_path / "id")
    procs = [_run_driver(simt_build, str(conf), r, 2, {"CPIC_B200_ID_FILE": idf}) for r in range(2)]
    outs = [p.communicate(timeout=900) for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[1][-1500:] for o in outs)
    assert "Simulation ends" in outs[0][0] and "Simulation ends" not in outs[1][0]
    energy = lambda text: [(float(a), float(b)) for a, b in re.findall(r"kinetic (\S+) potential (\S+)", text)]
    ke = sum(energy(o[0])[0][0] for o in outs)
    pe = sum(energy(o[0])[0][1] for o in outs)
    # the single-rank driver prints no energies: take them from the library
    ref = subprocess.run([sys.executable, "-c",
                          "import sys; sys.path.insert(0, %r); from cpic_b200 import Sim; s = Sim.from_conf(sys.argv[1]); s.run(6); print(*s.energy())" % ROOT,
                          str(conf)], cwd=ROOT, capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, CPIC_B200_LIB=simt_build, CPIC_B200_SIMT_CHECK="1"))
    assert ref.returncode == 0, ref.stderr[-1500:]
    ke1, pe1 = (float(v) for v in ref.stdout.split()[-2:])
    assert abs(ke - ke1) <= 1e-11 * abs(ke1) and abs(pe - pe1) <= 1e-9 * max(abs(pe1), 1e-300), (ke, ke1, pe, pe1)


@pytest.mark.parametrize("ranks,seed", [(2, 1), (2, 3), (2, 4), (4, 7)])
def test_random_configurations_on_several_ranks_on_cpu(simt_build, ranks, seed):
    """Random configurations (grid, block size, species, fields, velocities up to a block per step: tests/
    test_gpu_zz_robustness.py::_random_case) on 2 and 4 ranks against the single-rank oracle."""
    run_ranks_cpu(simt_build, ranks, f"random:{seed}", 6, raw=True)
