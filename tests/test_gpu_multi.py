"""Multi-GPU parity (needs >= 2 GPUs): P Y-slabs over NCCL against the single-rank oracle."""
import os
import subprocess
import sys

import pytest

from conftest import conf_path, ROOT

pytestmark = pytest.mark.gpu


def ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def run_ranks(n, conf, steps, mode="fused", port=29611, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mgpu_worker.py"), conf if conf.startswith("random:") else conf_path(conf), str(steps), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0 and "MGPU-OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("conf,mode", [("uniform-small.conf", "fused"), ("2d-2species-small.conf", "staged"),
                                       ("two-streams.conf", "fused"), ("far-beam.conf", "fused")])
def test_two_ranks(conf, mode):
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    run_ranks(2, conf, 10, mode)


@pytest.mark.parametrize("conf,mode,env", [("uniform-small.conf", "fused", None), ("far-beam-tall.conf", "fused", None),
                                          ("2d-2species-small.conf", "staged", None),
                                          ("uniform-small.conf", "fused", {"MGPU_RUN": "1"})])
def test_four_ranks(conf, mode, env):
    if ngpus() < 4:
        pytest.skip("needs 4 GPUs")
    run_ranks(4, conf, 12, mode, port=29613, env=env)


@pytest.mark.parametrize("conf,mode,env", [("uniform-tall.conf", "fused", None), ("far-beam-tall.conf", "fused", None),
                                          ("uniform-tall.conf", "staged", None),
                                          ("uniform-tall.conf", "fused", {"MGPU_RUN": "1", "MGPU_TIGHT": "1"})])
def test_eight_ranks(conf, mode, env):
    if ngpus() < 8:
        pytest.skip("needs 8 GPUs")
    run_ranks(8, conf, 12 if not env else 40, mode, port=29617, env=env)


def test_two_ranks_over_nccl():
    """The same exchange without peer memory (CPIC_B200_P2P=0): the face regions, halos and FFT transposes
    travel by NCCL send/recv."""
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    run_ranks(2, "far-beam.conf", 10, "fused", port=29619, env={"CPIC_B200_P2P": "0"})


def test_two_ranks_capacity_growth():
    """Tight block capacities: the ranks must agree on a larger capacity on the fly
    (check_capacity over NCCL) and stay identical to the oracle."""
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = run_ranks(2, "uniform-small.conf", 40, "fused", port=29615, env={"MGPU_TIGHT": "1"})
    assert "MGPU-CAPS" in out


def test_own_driver_on_two_ranks(tmp_path):
    """`mpirun -n 2 cpic <conf>` as two processes of cpic_b200_cli, one per GPU (rank and communicator id through
    the environment and a file): the energies of the two slabs add up to those of the single-rank run."""
    import re
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    cli = os.path.join(ROOT, "cpic_b200", "cpic_b200_cli")
    text = open(conf_path("2d-2species-small.conf")).read()
    text = re.sub(r"cycles\s*=\s*\d+", "cycles = 8", text)
    conf = tmp_path / "run.conf"
    conf.write_text(text)
    base = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    procs = [subprocess.Popen([cli, "-q", str(conf)], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                              env=dict(base, CPIC_B200_RANK=str(r), CPIC_B200_NRANKS="2", CPIC_B200_DEVICE=str(r),
                                       CPIC_B200_ID_FILE=str(tmp_path / "id"))) for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[1][-1500:] for o in outs)
    assert "Simulation ends" in outs[0][0]
    energy = lambda t: [(float(a), float(b)) for a, b in re.findall(r"kinetic (\S+) potential (\S+)", t)][0]
    ke = sum(energy(o[0])[0] for o in outs)
    pe = sum(energy(o[0])[1] for o in outs)
    from cpic_b200 import Sim
    s = Sim.from_conf(str(conf))
    s.run(8)
    ke1, pe1 = s.energy()
    s.close()
    assert abs(ke - ke1) <= 1e-11 * abs(ke1) and abs(pe - pe1) <= 1e-9 * max(abs(pe1), 1e-300), (ke, ke1, pe, pe1)


@pytest.mark.parametrize("seed", [1, 3, 4, 7])
def test_two_ranks_random_configurations(seed):
    """Random configurations (tests/test_gpu_zz_robustness.py::_random_case) on two ranks against the oracle."""
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    run_ranks(2, f"random:{seed}", 8, "fused", port=29621 + seed)
