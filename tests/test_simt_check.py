"""The device code on the CPU: cpic_b200/csrc/{kernels.cuh,sim.cu,comm.cu} compiled with g++
against the lockstep SIMT interpreter of tests/simt (fibers per CUDA thread, warp collectives
that wait for their participants, asynchronous copies that land at their wait, poisoned fresh
memory), then the gpu-marked parity tests run against that build in a subprocess.

This is test infrastructure: the package never builds or loads tests/simt, and a parity claim
is only ever made by the `-m gpu` run on the B200. What it buys on a box without a GPU is a
check of the kernels' logic -- indices, compaction order, capacities, barrier protocols,
races between lanes that hardware lockstep hides (it found two in round 1) -- against the
same oracle and the same assertions as the GPU run."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

SIMT = os.path.join(ROOT, "tests", "simt")
LIB = os.path.join(SIMT, "_build", "libcpic_b200_simt.so")


@pytest.fixture(scope="module")
def simt_build():
    r = subprocess.run(["make", "-C", SIMT], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert os.path.exists(LIB)
    return LIB


@pytest.fixture(scope="module")
def simt_build_alt():
    """The compile-time alternatives of the kernels: 16-byte cp.async instead of TMA bulk copies for the
    segment batches (OWN_BULK=0), one deposit launch per species (DEP_FUSED=0), a two-stage ring."""
    r = subprocess.run(["make", "-C", SIMT, "B=_build_alt", "EXTRA=-DOWN_BULK=0 -DDEP_FUSED=0 -DPIPE_STAGES=2"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return os.path.join(SIMT, "_build_alt", "libcpic_b200_simt.so")


@pytest.fixture(scope="module")
def simt_build_batch_major():
    """Batch-major segments (SEG_AOSOA=1): the 32 slots of a batch hold their six arrays in 1536
    contiguous bytes -- prepared for measurement, off in the shipped build."""
    r = subprocess.run(["make", "-C", SIMT, "B=_build_aosoa", "EXTRA=-DSEG_AOSOA=1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return os.path.join(SIMT, "_build_aosoa", "libcpic_b200_simt.so")


def run_under_interpreter(lib, args, timeout=900):
    env = dict(os.environ, CPIC_B200_LIB=lib, CPIC_B200_SIMT_CHECK="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", "-m", "gpu"] + args,
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout and "skipped" not in r.stdout, tail
    return r.stdout


def test_interpreter_selftest(simt_build):
    """Collectives, CTA barriers, late-landing cp.async / TMA copies, zero fill of tensor tiles,
    the driver's tensor-map checks and the deadlock report, each against a known answer."""
    r = subprocess.run([os.path.join(SIMT, "_build", "selftest")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "selftest: ok" in r.stdout


def test_product_library_is_not_the_interpreter():
    """The shipped library must not contain the interpreter, and the package must not look for it."""
    import cpic_b200._lib as L
    path = L.lib_path() if "CPIC_B200_LIB" not in os.environ else os.path.join(ROOT, "cpic_b200", "libcpic_b200.so")
    assert "simt" not in os.path.basename(path)
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
    assert "simt_check_reset" not in out and "simt_switch" not in out
    for name in os.listdir(os.path.join(ROOT, "cpic_b200")):
        if name.endswith(".py"):
            assert "simt" not in open(os.path.join(ROOT, "cpic_b200", name)).read(), name


def test_the_python_view_refuses_the_interpreter_build_outside_the_test_suite(simt_build):
    env = dict(os.environ, CPIC_B200_LIB=simt_build)
    env.pop("CPIC_B200_SIMT_CHECK", None)
    r = subprocess.run([sys.executable, "-c", "import cpic_b200._lib as L; L.lib()"], cwd=ROOT, env=env,
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU path" in r.stderr, r.stderr[-500:]


def test_kernel_parity_suite_under_the_interpreter(simt_build):
    """tests/test_gpu_parity.py, every case: solver, stage_field_E, deposit, the first 10 steps of
    seven configurations staged and fused, bitwise fused == staged, determinism, hot beam, velocity
    limit, ragged species, far movers, image round trip -- same oracle, same 1e-12."""
    out = run_under_interpreter(simt_build, ["tests/test_gpu_parity.py"])
    assert int(out.strip().splitlines()[-1].split()[0]) >= 36, out[-500:]


def test_robustness_suite_under_the_interpreter(simt_build):
    """tests/test_gpu_zz_robustness.py: streamed initialisation, 120 random configurations, exchange
    regions that overflow into the far-mover list."""
    run_under_interpreter(simt_build, ["tests/test_gpu_zz_robustness.py", "-k", "not two_ranks"])


def test_alternative_kernel_paths_under_the_interpreter(simt_build_alt):
    """The switches that are off in the shipped build keep passing the same parity suite."""
    run_under_interpreter(simt_build_alt, ["tests/test_gpu_parity.py", "-k",
                                           "first_10_steps or bitwise or far_movers or hot_beam or deposit"])


def test_batch_major_segments_under_the_interpreter(simt_build_batch_major):
    run_under_interpreter(simt_build_batch_major, ["tests/test_gpu_parity.py", "-k",
                                                   "first_10_steps or bitwise or far_movers or image_round_trip or ragged"])


def test_reference_drivers_through_the_dropin_under_the_interpreter(simt_build):
    """The reference's own test/cyclotron.c and its `cpic -q <conf>` main with output enabled, linked to
    dropin/cpic_b200_stages.c, with the C ABI served by the interpreted kernels (the 50000-cycle
    constant-speed and the 1200-step harmonic drivers are left to the GPU run)."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "dropin_cpic")):
        pytest.skip("drop-in drivers not built (needs /root/reference at build time)")
    run_under_interpreter(simt_build, ["tests/test_gpu_dropin.py", "-k", "cyclotron or cli_and_output"])


def test_bench_script_under_the_interpreter(simt_build):
    """bench.py itself, every leg of its default run (main workload, per-stage pass, staged stages, C-ABI e2e,
    the other workloads, the multi-process CPU reference), at a tiny size (BENCH_TINY) against the
    interpreted kernels: the one JSON line must parse and carry the contract's keys."""
    import json
    env = dict(os.environ, CPIC_B200_LIB=simt_build, CPIC_B200_SIMT_CHECK="1", BENCH_TINY="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["config"]["workload"].startswith("D:") and d["gpu_launches"] > 0
    assert set(d["other_workloads"]) == {"A", "2s"} and all("value" in v for v in d["other_workloads"].values())
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    if d["cpu_baseline"]["kind"] == "reference":
        assert d["cpu_baseline"]["cores"] >= 1


def test_bench_reference_arm(simt_build):
    """`bench.py --impl reference` (the unmodified reference over forked ranks, oracle/_ref/cpic_ref_mp) at
    the tiny size: one JSON line with impl, cpu_baseline and e2e; no product library is imported."""
    import json
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "cpic_ref_mp")):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    env = dict(os.environ, BENCH_TINY="1")
    env.pop("CPIC_B200_LIB", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "4", "--warmup", "3"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference" and d["value"] > 0
    assert d["warmup"] == 3 and d["steps"] == 4
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["config"]["workload"].startswith("D:")


def test_physics_under_the_interpreter(simt_build):
    """Two-stream growth rate and energy history, constant speed and cyclotron orbits (the 1200-step
    harmonic golden trajectory is left to the GPU run: minutes of interpretation)."""
    run_under_interpreter(simt_build, ["tests/test_gpu_physics.py", "-k", "two_stream or constant_speed or cyclotron"])


def test_own_driver_and_asynchronous_output_under_the_interpreter(simt_build, tmp_path):
    """cpic_b200_main (the stand-alone driver: cpic's command line and run loop) with output enabled, served by
    the interpreted kernels: grids copied on the copy stream into pinned staging and written by the
    background thread in aligned slices -- same file sizes, values (1e-12) and XDMF text as the CPU
    reference's own output.c."""
    import numpy as np
    ref = os.path.join(ROOT, "oracle", "_ref", "cpic_ref")
    if not os.path.exists(ref):
        pytest.skip("cpic_ref not built (needs /root/reference at build time)")
    text = open(os.path.join(ROOT, "conf", "two-streams.conf")).read().replace("cycles = 800", "cycles = 4")
    outs = {}
    for name in ("gpu", "cpu"):
        d = tmp_path / name
        conf = tmp_path / f"{name}.conf"
        conf.write_text(text + f'\noutput = {{ path = "{d}" slices = 4 alignment = 4096 }}\n')
        if name == "cpu":
            r = subprocess.run([ref, "-q", str(conf)], cwd=ROOT, capture_output=True, text=True, timeout=300)
        else:
            code = ("import ctypes as C, sys; L = C.CDLL(sys.argv[1]); "
                    "argv = (C.c_char_p * 3)(b'cpic', b'-q', sys.argv[2].encode()); sys.exit(L.cpic_b200_main(3, argv))")
            r = subprocess.run([sys.executable, "-c", code, simt_build, str(conf)], cwd=ROOT, capture_output=True, text=True,
                               timeout=600, env=dict(os.environ, CPIC_B200_SIMT_CHECK="1"))
        assert r.returncode == 0, r.stderr[-2000:] + r.stdout[-500:]
        outs[name] = d
    nx = 64
    for it in range(4):
        for f in ("rho", "phi", "E_X", "E_Y"):
            pa, pb = outs["gpu"] / "bin" / str(it) / f"{f}.bin", outs["cpu"] / "bin" / str(it) / f"{f}.bin"
            assert os.path.getsize(pa) == os.path.getsize(pb) and os.path.getsize(pa) % 4096 == 0
            a, b = np.fromfile(pa), np.fromfile(pb)
            ok = np.isfinite(b) & (np.abs(b) < 1e300)
            if f in ("rho", "phi"):
                ok &= (np.arange(a.size) % (nx + 2)) < nx
            n_live = {"rho": 65, "phi": 67, "E_X": 65, "E_Y": 65}[f] * (nx + 2 if f in ("rho", "phi") else nx)
            ok &= np.arange(a.size) < n_live
            assert np.abs(a[ok] - b[ok]).max() <= 1e-12 * np.abs(b[ok]).max(), (it, f)
        assert (outs["gpu"] / "xdmf" / f"fields-iter{it}.xdmf").read_text() == (outs["cpu"] / "xdmf" / f"fields-iter{it}.xdmf").read_text()
