"""GPU robustness tests, kept in a file that sorts after the established suites: the streamed
initialisation, random configurations against the oracle, exchange regions that overflow. Same
oracle and tolerance as tests/test_gpu_parity.py; they also run under the SIMT interpreter in the CPU
suite (tests/test_simt_check.py)."""
import numpy as np
import pytest

from conftest import conf_path
from _parity import (TOL, pair_from_conf, oracle_from, gpu_from, field_errors, particle_errors, assert_close)
from cpic_b200 import Sim, Params

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("conf,batch", [("2d-2species-small.conf", 777), ("two-streams.conf", 64), ("far-beam.conf", 1 << 20)])
def test_streamed_initialisation(conf, batch):
    """cpic_b200_sim_from_conf_streamed: the reference's initial conditions generated and uploaded in
    batches (count, reserve, append). Same particles, same capacity as the one-shot path; inside a block
    they are ordered by batch, so sums differ in the last bits only."""
    g, o, params, _ = pair_from_conf(conf_path(conf))
    s = Sim.from_conf(conf_path(conf), stream_batch=batch)
    for i in range(len(params.q)):
        assert s.capacity(i) == g.capacity(i)
        assert s.num_particles(i) == g.num_particles(i)
    for it in range(5):
        s.sync()
        assert_close(field_errors(s, o), what=f"{conf} streamed, iteration {it}")
        assert_close(particle_errors(s, o, params), what=f"{conf} streamed, iteration {it}")
        s.step()
        o.step()


def test_streamed_initialisation_rejects_misuse():
    p = Params(64, 64, 4.0, 4.0, 0.01, 1.0)
    s = Sim(p)
    x = np.array([1.0, 5.0])
    y = np.array([1.0, 1.0])
    L = s.L
    ptr = lambda a: a.ctypes.data
    assert L.cpic_b200_reserve_counted(s.h, 0) != 0                      # nothing counted yet
    assert L.cpic_b200_count_particles(s.h, 0, 2, ptr(x), ptr(y)) != 0   # x = 5 is outside [0, 4]
    ids = np.arange(1, dtype=np.int64)
    one = np.array([1.0])
    assert L.cpic_b200_add_particles(s.h, 0, 1, ptr(ids), ptr(one), ptr(one), ptr(one), ptr(one), None) != 0  # no storage
    assert L.cpic_b200_count_particles(s.h, 0, 1, ptr(one), ptr(one)) == 0
    assert L.cpic_b200_reserve_counted(s.h, 0) == 0
    assert L.cpic_b200_add_particles(s.h, 0, 1, ptr(ids), ptr(one), ptr(one), ptr(one), ptr(one), None) == 0
    assert s.num_particles(0) == 1


def _random_case(seed):
    """A random but reference-legal configuration: grid 8..128 per side, particle blocks of 4/8/16
    cells, 1-3 species with mixed charges and masses, B along z, uniform or clustered populations of
    0..3000 particles, speeds from 0.1 to 20 cells per step (far movers) under the reference's umax."""
    rng = np.random.default_rng(seed)
    nx, ny = int(2 ** rng.integers(3, 8)), int(2 ** rng.integers(3, 8))
    bc = min(int(rng.choice([4, 8, 16])), nx, ny)
    dx = float(rng.choice([0.25, 0.5, 1.0, 0.03125]))
    nsp = int(rng.integers(1, 4))
    q = tuple(float(rng.choice([-1.0, 1.0, -2.0])) for _ in range(nsp))
    m = tuple(float(rng.choice([1.0, 0.5, 1836.0, 0.01])) for _ in range(nsp))
    B = (0.0, 0.0, float(rng.choice([0.0, -1.0, 0.5, 3.0])))
    dt = float(rng.choice([0.01, 0.05]))
    e0 = float(rng.choice([100.0, 1e4]))
    chunks = min(int(rng.choice([1, 2, 4])), nx // 2)
    p = Params(nx, ny, nx * dx, ny * dx, dt, e0, B, q, m, chunks, block_cells=bc,
               keep_particle_E=bool(rng.integers(0, 2)))
    umax = (nx // chunks) * dx / dt
    parts, idbase = [], 0
    for _ in range(nsp):
        n = int(rng.choice([0, 1, 7, 200, 3000]))
        v = min(float(rng.choice([0.1, 1.0, 6.0, 20.0])) * dx / dt, 0.45 * umax, 0.45 * ny * dx / dt)
        if rng.integers(0, 2):
            x = (rng.normal(0.5, 0.05, n) % 1.0) * p.Lx
            y = (rng.normal(0.5, 0.05, n) % 1.0) * p.Ly
        else:
            x, y = rng.random(n) * p.Lx, rng.random(n) * p.Ly
        parts.append({"id": np.arange(idbase, idbase + n, dtype=np.int64), "x": x, "y": y,
                      "ux": (rng.random(n) * 2 - 1) * v, "uy": (rng.random(n) * 2 - 1) * v, "uz": np.zeros(n)})
        idbase += n
    return p, parts, bool(rng.integers(0, 2)), int(rng.integers(3, 7))


@pytest.mark.parametrize("first", range(0, 120, 20))
def test_random_configurations(first):
    """Twenty random configurations per case against the oracle. Loud refusals are legitimate outcomes
    (the reference's velocity limit; an exchange region that a clustered beam overflows in one step);
    a wrong number is not."""
    from cpic_b200._lib import Cpic_b200Error
    ran = 0
    for seed in range(first, first + 20):
        p, parts, staged, steps = _random_case(seed)
        o = oracle_from(p, parts)
        g = gpu_from(p, parts)
        try:
            o.pre_step()
            g.pre_step()
            for it in range(steps):
                (g.step_staged if staged else g.step)()
                o.step()
                g.sync()
                errs = {**field_errors(g, o), **particle_errors(g, o, p)}
                # rounding differences grow with the plasma's own instabilities: 1e-12 at the start
                # (the parity protocol), two digits of slack by the last of these steps
                assert_close(errs, tol=TOL * 10 ** min(it, 2), what=f"seed {seed}, iteration {it}")
            ran += 1
        except Cpic_b200Error as e:
            assert "exchange region" in str(e) or "velocity" in str(e), (seed, str(e))
        except RuntimeError as e:
            assert "velocity limit" in str(e), (seed, str(e))
        finally:
            g.close()
    assert ran >= 10


def test_full_exchange_regions_spill_into_the_far_list():
    """A beam that sends ~200 particles per block and step across one face into regions of 64 slots:
    the leavers that do not fit are listed like far movers and placed by k_far_insert -- same
    particles, same fields as the oracle, no refusal."""
    rng = np.random.default_rng(5)
    nx = ny = 64
    dx, dt = 0.5, 0.05
    p = Params(nx, ny, nx * dx, ny * dx, dt, 1e4, (0.0, 0.0, 0.0), (-1.0,), (1.0,), 1, block_cells=8, outbox_fraction=0.05)
    n = 30000
    parts = [{"id": np.arange(n, dtype=np.int64), "x": rng.random(n) * p.Lx, "y": rng.random(n) * p.Ly,
              "ux": np.full(n, 5 * dx / dt), "uy": np.full(n, -3 * dx / dt), "uz": np.zeros(n)}]
    o = oracle_from(p, parts)
    g = gpu_from(p, parts)
    o.pre_step()
    g.pre_step()
    assert g.occupancy(0)["side_cap"] == 64
    for it in range(3):
        g.step()
        o.step()
        g.sync()
        assert_close({**field_errors(g, o), **particle_errors(g, o, p)}, what=f"iteration {it}")


def test_host_getters_grow_a_segment_that_cannot_take_its_arrivals():
    """A uniform species collapsing onto a cluster of opposite charge: a block ends up with arrivals
    from all eight neighbours. Without a cpic_b200_sync in between nothing has grown the capacity;
    reading the particles must do it rather than drop the arrivals."""
    rng = np.random.default_rng(1743)
    nx, ny, dx, dt, n = 64, 32, 0.5, 0.05, 3000
    p = Params(nx, ny, nx * dx, ny * dx, dt, 1.0, (0.0, 0.0, 0.5), (1.0, -2.0), (0.5, 1.0), 1, block_cells=4)
    v = dx / dt
    parts = [{"id": np.arange(n, dtype=np.int64), "x": rng.random(n) * p.Lx, "y": rng.random(n) * p.Ly,
              "ux": (rng.random(n) * 2 - 1) * v, "uy": (rng.random(n) * 2 - 1) * v, "uz": np.zeros(n)},
             {"id": np.arange(n, 2 * n, dtype=np.int64), "x": (rng.normal(0.5, 0.05, n) % 1.0) * p.Lx,
              "y": (rng.normal(0.5, 0.05, n) % 1.0) * p.Ly,
              "ux": (rng.random(n) * 2 - 1) * v, "uy": (rng.random(n) * 2 - 1) * v, "uz": np.zeros(n)}]
    o = oracle_from(p, parts)
    g = gpu_from(p, parts)
    o.pre_step()
    g.pre_step()
    cap0 = g.capacity(0)
    for it in range(8):
        g.step()
        o.step()
        occ = g.occupancy(0)               # read-only: own particles + pending arrivals of the fullest block
        if occ["block"] > occ["block_cap"]:
            break
    else:
        pytest.fail("the population never outgrew a segment: the case no longer exercises the path")
    assert sum(len(g.particles(i)["id"]) for i in range(len(p.q))) == sum(len(q["id"]) for q in parts)
    assert g.capacity(0) > cap0
    g.sync()
    assert_close(particle_errors(g, o, p), tol=1e-9, what="after the growth")
    for it in range(3):
        g.step()
        o.step()
    g.sync()
    assert_close(particle_errors(g, o, p), tol=1e-8, what="three steps later")


def test_two_ranks_streamed_initialisation():
    from test_gpu_multi import ngpus, run_ranks
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    run_ranks(2, "2d-2species-small.conf", 6, port=29617, env={"MGPU_STREAMED": "1000"})

