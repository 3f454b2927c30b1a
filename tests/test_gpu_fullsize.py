"""BASELINE.json configs[1] at its full size (conf/2d-2species.conf: 1024^2 grid, 1e7 particles, the
workload bench.py measures): parity with the oracle for the first steps, then the
size-independent properties of the path over a longer run."""
import numpy as np
import pytest

from conftest import conf_path
from _parity import oracle_from, gpu_from, field_errors, particle_errors, assert_close
from cpic_b200 import load_conf, init_particles

pytestmark = pytest.mark.gpu


def test_config_A_full_size():
    conf = conf_path("2d-2species.conf")
    params, run = load_conf(conf)
    parts = init_particles(conf)
    n = [len(p["id"]) for p in parts]
    assert n == [5_000_000, 5_000_000]
    o = oracle_from(params, parts)
    g = gpu_from(params, parts)
    o.pre_step()
    g.pre_step()
    assert_close(field_errors(g, o), what="config A after sim_init")
    for it in range(10):      # BASELINE.json north_star: the first 10 steps
        g.step()
        o.step()
        g.sync()
        assert_close(field_errors(g, o), what=f"config A fields, iteration {it}")
        assert_close(particle_errors(g, o, params), what=f"config A particles, iteration {it}")

    # properties that hold at any size
    ke0, _ = g.energy()
    g.run(60)
    # 1. comm_plasma loses nobody (src/comm_plasma.c: every collected particle is injected somewhere)
    assert [g.num_particles(i) for i in range(2)] == n
    ids = g.particles(0)["id"]
    assert np.array_equal(ids, np.arange(n[0]))
    # 2. every particle deposits exactly -q/e0 (CIC weights sum to one): total charge is conserved
    rho = g.field("rho")
    expect = sum(-q / params.e0 * k for q, k in zip(params.q, n))
    each = sum(abs(q) / params.e0 * k for q, k in zip(params.q, n))
    assert abs(rho.sum() - expect) <= 1e-10 * each
    # 3. the solve leaves phi with zero mean (G(0,0) = 0, src/solver.c:268-269)
    phi = g.field("phi")
    assert abs(phi.mean()) <= 1e-12 * np.abs(phi).max()
    # 4. periodic wrap: every particle inside [0, L] (src/comm_plasma.c:738-747)
    for i in range(2):
        p = g.particles(i, sort=False)
        assert p["x"].min() >= 0.0 and p["x"].max() <= params.Lx
        assert p["y"].min() >= 0.0 and p["y"].max() <= params.Ly
        assert np.all(p["uz"] == 0.0)
    # 5. B only rotates and the field energy is tiny in this weakly coupled plasma: kinetic energy drifts little
    ke1, _ = g.energy()
    assert abs(ke1 - ke0) / ke0 < 1e-4


def test_config_B_cyclotron_2048_first_steps():
    """BASELINE.json configs[2]: conf/cyclotron-2048.conf (uniform B, 2048^2 grid, 1e8 particles) from the
    reference's own initial conditions: sim_init and three iterations against the oracle."""
    conf = conf_path("cyclotron-2048.conf")
    params, run = load_conf(conf)
    parts = init_particles(conf)
    assert sum(len(p["id"]) for p in parts) == 100_000_000
    o = oracle_from(params, parts)
    g = gpu_from(params, parts)
    o.pre_step()
    g.pre_step()
    assert_close(field_errors(g, o), what="config B after sim_init")
    for it in range(3):
        g.step()
        o.step()
        g.sync()
        assert_close(field_errors(g, o), what=f"config B fields, iteration {it}")
        assert_close(particle_errors(g, o, params), what=f"config B particles, iteration {it}")


def test_config_D_size_independent_properties():
    """BASELINE.json configs[4] at its full single-GPU size (2048^2 cells, 2.5e8 particles: bench.py's default
    workload, device initialiser): what must hold at any size -- comm_plasma loses nobody, every particle
    deposits exactly -q/e0 (CIC weights sum to one), the solve leaves phi with zero mean, the energies stay
    finite and the kinetic energy of this weakly coupled beam drifts little."""
    import bench
    from cpic_b200 import Sim
    w = bench.WORKLOADS["D"]
    conf = bench.scaled_conf(conf_path(w["conf"]), w, w["nps"])
    params, run = load_conf(conf)
    params.outbox_fraction = 0.19
    g = Sim(params)
    n = w["nps"]
    for i in range(2):
        g.init_beam(i, n, id0=0, drift=w["drift"][i], spread=w["spread"][i], seed=138 + i)
    g.pre_step()
    ke0, _ = g.energy()
    g.run(20)
    assert [g.num_particles(i) for i in range(2)] == [n, n]
    rho = g.field("rho")
    each = sum(abs(q) / params.e0 * n for q in params.q)
    expect = sum(-q / params.e0 * n for q in params.q)
    assert abs(rho.sum() - expect) <= 1e-10 * each
    phi = g.field("phi")
    assert abs(phi.mean()) <= 1e-12 * np.abs(phi).max()
    ke1, pe1 = g.energy()
    assert np.isfinite(ke1) and np.isfinite(pe1) and abs(ke1 - ke0) / ke0 < 1e-3
    g.close()
