"""CPU tests of the host side: the `.conf` reader, sim_read_config/sim_prepare semantics, the
host particle initialisers (bit-identical to the reference), and the C ABI surface."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import conf_path, ROOT
import cpic_b200
from cpic_b200 import load_conf, init_particles, Cpic_b200Error, Params, Sim
from _refbind import RefSim, ref_available

needs_ref = pytest.mark.skipif(not ref_available("ref"), reason="oracle/_ref not built (needs /root/reference)")


def test_library_exports_every_declared_symbol():
    """Every function include/cpic_b200.h declares is exported by libcpic_b200.so."""
    hdr = open(os.path.join(ROOT, "include", "cpic_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cpic_b200_[A-Za-z0-9_]+)\s*\(", hdr))
    declared -= {"cpic_b200_sim_t", "cpic_b200_conf_t"}
    L = C.CDLL(cpic_b200.lib_path())
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    assert declared == set(cpic_b200.EXPORTS), declared ^ set(cpic_b200.EXPORTS)
    assert b"sm_100a" in cpic_b200.lib().cpic_b200_version()


def test_no_cpu_fallback():
    """Without a GPU the product refuses to run (there is no CPU path to fall back to)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(Cpic_b200Error) as e:
        Sim(Params(16, 16, 1.0, 1.0, 0.1, 1.0))
    assert e.value.code == 2 and "no CPU path" in str(e.value)


def test_conf_params_match_reference_prepare():
    p, run = load_conf(conf_path("2d-2species.conf"))
    assert (p.nx, p.ny, p.Lx, p.Ly, p.dt) == (1024, 1024, 4.0, 4.0, 5.0e-3)
    assert p.B == (0.0, 0.0, -0.2) and p.q == (-1.0, 1.0) and p.m == (1.0e-2, 2.0)
    assert run.nparticles == [5000000, 5000000] and run.seed == 138 and run.solver == "MFT"
    assert run.cycles == 100 and not run.output_enabled and run.output_alignment == 512


def test_conf_grammar(tmp_path):
    inc = tmp_path / "constants.conf"
    inc.write_text('constants = { light_speed = 2.99792458e+8; vacuum_permittivity : 1.5 }\n')
    main = tmp_path / "t.conf"
    main.write_text('''
# hash comment
// slash comment
/* block
   comment */
@include "constants.conf"
species = ( { name = "a" "b"; particles = 12L, charge = -1.0 mass = 2.5e-1
              drift_velocity = [ 1.0, -2.0 ] init_method = "position delta"
              position_delta = [0.5, 0.25]; position_init = [0.0, 0.0] } )
field = { magnetic = [0.0, 0.0, 1.0e-1] }
grid : { points = [ 0x10, 8 ] }
simulation = { dimensions = 2 cycles = 5 time_step = 1.0e-2 random_seed = 7 solver = "MFT"
  enable_fftw_threads = 0 plasma_chunks = 2 pblock_nmax = 64 space_length = [2.0, 1.0]
  sampling_period = { energy = 0 field = 0 particle = 0 } stop_SEM = 0.0 realtime_plot = 0 }
output = { path = "/tmp/x" slices = 4 alignment = 4096 }
''')
    p, run = load_conf(main)
    assert (p.nx, p.ny, p.e0, p.plasma_chunks) == (16, 8, 1.5, 2)
    assert p.q == (-1.0,) and p.m == (0.25,) and run.nparticles == [12]
    assert run.output_enabled and run.output_path == "/tmp/x" and run.output_slices == 4
    parts = init_particles(main)
    assert np.allclose(parts[0]["x"], np.fmod(0.5 * np.arange(12), 2.0))
    assert (parts[0]["ux"] == 1.0).all() and (parts[0]["uy"] == -2.0).all()


@pytest.mark.parametrize("text,msg", [
    ("simulation = { dimensions = 2 }", 'Failed to read parameter "simulation.cycles"'),
    ("simulation = { dimensions = 2", "unexpected end of file"),
    ("x = [1, 2.0]", "mismatched element type"),
])
def test_conf_errors(tmp_path, text, msg):
    f = tmp_path / "bad.conf"
    f.write_text(text)
    with pytest.raises(Cpic_b200Error) as e:
        load_conf(f)
    assert msg in str(e.value)


def test_reference_1d_era_confs_are_rejected_like_the_reference():
    """files without the current schema's keys fail with the reference's message."""
    src = "/root/reference/conf/two-streams.conf"
    if not os.path.exists(src):
        pytest.skip("reference tree not mounted")
    with pytest.raises(Cpic_b200Error) as e:
        load_conf(src)
    assert "Failed to read parameter" in str(e.value)


@needs_ref
@pytest.mark.parametrize("conf", ["uniform-small.conf", "2d-2species-small.conf", "2d-2species-delta.conf",
                                  "two-streams.conf", "cyclotron.conf", "constant-speed.conf", "harmonic-64.conf"])
def test_host_init_is_bit_identical_to_reference(conf):
    """plasma_init + particles_init (src/plasma.c:62-128, src/particle.c:17-168) including the
    glibc rand() stream order over chunks and species."""
    parts = init_particles(conf_path(conf))
    r = RefSim(conf_path(conf), "ref")
    p, _ = load_conf(conf_path(conf))
    assert (r.nx, r.ny, r.dt, r.e0) == (p.nx, p.ny, p.dt, p.e0)
    for i, mine in enumerate(parts):
        ref = r.particles(i)
        for k in ("id", "x", "y", "ux", "uy"):
            assert np.array_equal(ref[k], mine[k]), (conf, i, k)


def test_ref_nprocs_striping():
    """With ref_nprocs = 2 the ids are striped over 2*plasma_chunks chunks and each process
    seeds rand() with seed + rank (src/sim.c:153, src/plasma.c:62-63)."""
    a = init_particles(conf_path("uniform-small.conf"), ref_nprocs=1)[0]
    b = init_particles(conf_path("uniform-small.conf"), ref_nprocs=2)[0]
    assert np.array_equal(a["id"], b["id"])
    assert not np.array_equal(a["x"], b["x"])
    # process 0 of 2 draws the same stream as the single process, for different ids:
    # chunk 0 of 2 -> ids 0, 2, 4 .. (1 process)  vs  chunk 0, proc 0 of 2 -> ids 0, 4, 8 ..
    n = len(a["id"])
    first_a = a["x"][0:n:2][: n // 4]
    first_b = b["x"][0:n:4][: n // 4]
    assert np.array_equal(first_a, first_b)


def test_cli_usage_and_no_gpu_message():
    """cpic_b200_cli keeps the reference's command line (src/cpic.c:30-38, :114-127)."""
    import subprocess
    cli = os.path.join(ROOT, "cpic_b200", "cpic_b200_cli")
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage:" in r.stderr and "<config file>" in r.stderr
    r = subprocess.run([cli, "-q", "/nonexistent.conf"], capture_output=True, text=True)
    assert r.returncode == 1 and "Configuration read failed" in r.stderr
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([cli, "-q", conf_path("cyclotron.conf")], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU path" in r.stderr


def test_glibc_rand_recurrence_and_jump_ahead(tmp_path):
    """csrc/host/glibc_rand.h restates glibc's rand() (TYPE_3: r[i] = r[i-31] + r[i-3]) so that the device
    initialiser can draw the reference's initial conditions from any point of the stream: the seeding,
    5000 values and two jump-aheads (matrix powers) against the C library itself, six seeds."""
    import subprocess
    exe = str(tmp_path / "glibc_rand_check")
    src = os.path.join(ROOT, "tests", "glibc_rand_check.cpp")
    r = subprocess.run(["g++", "-O2", "-I", os.path.join(ROOT, "cpic_b200", "csrc", "host"), "-o", exe, src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "bad=0" in r.stdout, r.stdout[-500:]


def test_reference_arm_leaves_other_ranks_silent():
    """`bench.py --impl reference` under torchrun: rank 0 alone runs and prints; the other ranks exit 0 without work."""
    import subprocess
    import sys
    env = dict(os.environ, RANK="3", WORLD_SIZE="8", LOCAL_RANK="3", BENCH_TINY="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "8"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
