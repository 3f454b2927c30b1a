"""GPU parity tests: the CUDA path through the C ABI against the oracle on identical
initial conditions. Tolerance: 1e-12 relative (BASELINE.json north_star) in the norms of
SURVEY 8c: max|a-b|/max|b| per field, |dx|/L and |du|/max|u| per particle."""
import numpy as np
import pytest

from conftest import conf_path
from _parity import (TOL, relerr, pair_from_conf, oracle_from, gpu_from, field_errors, particle_errors,
                     assert_close)
from cpic_b200 import Sim, Params, load_conf, init_particles

pytestmark = pytest.mark.gpu

CONFS = ["uniform-small.conf", "2d-2species-small.conf", "2d-2species-delta.conf", "two-streams.conf",
         "harmonic-64.conf", "cyclotron.conf", "far-beam.conf"]


@pytest.mark.parametrize("shape", [(64, 64), (128, 32), (4, 4), (8, 16), (256, 128)])
def test_solver_parity(shape):
    """MFT_solve (src/solver.c:465-509): cuFFT + k_green against the oracle's FFT."""
    nx, ny = shape
    p = Params(nx, ny, 4.0, 4.0 * ny / nx, 0.01, 1.0)
    g = Sim(p)
    o = oracle_from(p, [])
    rho = np.random.default_rng(nx * 1000 + ny).standard_normal((ny + 1, nx))
    raw = np.zeros(g.field_shape("rho"))
    raw[:, :nx] = rho
    g.set_raw_field("rho", raw)
    o.set_rho(rho)
    g.solve()
    g.sync()
    o.solve()
    assert relerr(g.field("phi"), o.field("phi")) <= TOL


@pytest.mark.parametrize("shape", [(64, 64), (4, 4), (16, 8)])
def test_stage_field_E_parity(shape):
    """stage_field_E (src/field.c:450-501): solve + phi ghosts + centred difference."""
    nx, ny = shape
    p = Params(nx, ny, 2.0, 2.0 * ny / nx, 0.01, 1.0)
    g = Sim(p)
    o = oracle_from(p, [])
    rho = np.random.default_rng(7).standard_normal((ny + 1, nx))
    raw = np.zeros(g.field_shape("rho"))
    raw[:, :nx] = rho
    g.set_raw_field("rho", raw)
    o.set_rho(rho)
    g.stage_field_E()
    g.sync()
    o.stage_field_E()
    assert_close(field_errors(g, o, ("phi_ghost", "Ex", "Ey")), what=f"stage_field_E {shape}")


@pytest.mark.parametrize("conf", CONFS)
def test_deposit_parity(conf):
    """stage_field_rho (src/field.c:268-356) right after the upload."""
    g, o, params, _ = pair_from_conf(conf_path(conf))
    g.sync()
    assert_close(field_errors(g, o, ("rho_ghost",)), what=conf)
    # total deposited charge: every particle contributes exactly -q/e0
    tot = sum(-q / params.e0 * len(o.particles(i)["id"]) for i, q in enumerate(params.q))
    got = g.field("rho").sum()
    assert abs(got - tot) <= 1e-9 * max(1.0, abs(tot))


@pytest.mark.parametrize("conf", CONFS)
def test_first_10_steps_staged(conf):
    """The four stage calls, separately, for iterations 0..9 (0 is the rewind
    step), compared after every stage-complete step. Includes the per-particle E."""
    g, o, params, _ = pair_from_conf(conf_path(conf))
    assert_close(field_errors(g, o), what=f"{conf} after sim_init")
    for it in range(10):
        g.step_staged()
        o.step()
        g.sync()
        assert_close(field_errors(g, o), what=f"{conf} fields, iteration {it}")
        assert_close(particle_errors(g, o, params, with_E=True), what=f"{conf} particles, iteration {it}")


@pytest.mark.parametrize("conf", CONFS)
def test_first_10_steps_fused(conf):
    """cpic_b200_step (gather+push fused, E never stored) for iterations 0..9."""
    g, o, params, _ = pair_from_conf(conf_path(conf))
    for it in range(10):
        g.step()
        o.step()
        g.sync()
        assert_close(field_errors(g, o), what=f"{conf} fields, iteration {it}")
        assert_close(particle_errors(g, o, params), what=f"{conf} particles, iteration {it}")
    assert g.iter == 10


def test_fused_equals_staged_bitwise():
    """Fusing gather into the push must not change a single bit."""
    a, _, params, _ = pair_from_conf(conf_path("2d-2species-small.conf"))
    b, _, _, _ = pair_from_conf(conf_path("2d-2species-small.conf"))
    for _ in range(6):
        a.step()
        b.step_staged()
    for i in range(len(params.q)):
        pa, pb = a.particles(i, sort=False), b.particles(i, sort=False)
        for k in ("id", "x", "y", "ux", "uy", "uz"):
            assert np.array_equal(pa[k], pb[k]), (i, k, np.abs(pa[k] - pb[k]).max())
    for name in ("rho", "phi", "Ex", "Ey"):
        assert np.array_equal(a.raw_field(name), b.raw_field(name)), name


def test_run_to_run_determinism():
    """Sums are ordered: two runs give identical bits, particle order included."""
    runs = []
    for _ in range(2):
        g, _, params, _ = pair_from_conf(conf_path("uniform-small.conf"))
        for _ in range(12):
            g.step()
        g.sync()
        runs.append(([g.raw_field(n) for n in ("rho", "phi", "Ex", "Ey")],
                     [g.particles(i, sort=False) for i in range(len(params.q))]))
        g.close()
    for fa, fb in zip(runs[0][0], runs[1][0]):
        assert np.array_equal(fa, fb)
    for pa, pb in zip(runs[0][1], runs[1][1]):
        for k in pa:
            assert np.array_equal(pa[k], pb[k]), k


def test_hot_beam_migration():
    """comm_plasma (src/comm_plasma.c:1122-1142): a beam that crosses about one cell per step
    in both directions, so that a large share of every particle block migrates each step."""
    nx = ny = 64
    n = 30000
    rng = np.random.default_rng(5)
    L = 8.0
    dx = L / nx
    dt = 0.05
    p = Params(nx, ny, L, L, dt, 1.0e6, (0.0, 0.0, 0.2), (-1.0,), (1.0,), plasma_chunks=1)
    v = 0.9 * dx / dt
    parts = [{"id": np.arange(n), "x": rng.uniform(0, L, n), "y": rng.uniform(0, L, n),
              "ux": rng.choice([-v, v], n), "uy": rng.choice([-v, v], n)}]
    o = oracle_from(p, parts)
    g = gpu_from(p, parts)
    o.pre_step()
    g.pre_step()
    for it in range(20):
        g.step()
        o.step()
        g.sync()
        assert g.num_particles(0) == n
        assert_close(field_errors(g, o), what=f"hot beam fields, iteration {it}")
        assert_close(particle_errors(g, o, p), what=f"hot beam particles, iteration {it}")


def test_velocity_limit_is_reported():
    """check_velocity (src/mover.c:97-137) aborts in the reference; here it is an error code."""
    from cpic_b200 import Cpic_b200Error
    p = Params(16, 16, 1.0, 1.0, 1.0, 1.0, (0, 0, 0), (-1.0,), (1.0,), plasma_chunks=16)
    parts = [{"id": np.arange(4), "x": np.full(4, 0.5), "y": np.full(4, 0.5),
              "ux": np.full(4, 100.0), "uy": np.zeros(4)}]
    g = gpu_from(p, parts)
    g.pre_step()
    g.step()
    with pytest.raises(Cpic_b200Error) as e:
        g.sync()
    assert e.value.code == 3


def test_empty_and_ragged_species():
    """A species with zero particles and one with a single particle (tail-pack cases of
    src/interpolate.c:329-344 have no analogue here, but empty blocks must work)."""
    p = Params(32, 16, 2.0, 1.0, 0.01, 1.0, (0, 0, 0.1), (-1.0, 1.0), (1.0, 4.0))
    parts = [{"id": np.zeros(0, np.int64), "x": np.zeros(0), "y": np.zeros(0), "ux": np.zeros(0), "uy": np.zeros(0)},
             {"id": np.array([7]), "x": np.array([1.99]), "y": np.array([0.999]), "ux": np.array([0.5]), "uy": np.array([0.7])}]
    o = oracle_from(p, parts)
    g = gpu_from(p, parts)
    o.pre_step()
    g.pre_step()
    for it in range(8):
        g.step_staged()
        o.step()
        g.sync()
        assert_close(field_errors(g, o), what=f"ragged, iteration {it}")
        assert_close(particle_errors(g, o, p, with_E=True), what=f"ragged particles, iteration {it}")


def test_far_movers():
    """Particles that jump several particle blocks in one step (allowed by the reference up to
    one chunk, src/sim.c:198-200) take the far-mover path and must still match the oracle."""
    nx = ny = 64
    n = 4000
    rng = np.random.default_rng(11)
    L = 8.0
    dx = L / nx
    dt = 0.05
    p = Params(nx, ny, L, L, dt, 1.0e6, (0.0, 0.0, 0.1), (-1.0,), (1.0,), plasma_chunks=1)
    v = 20.0 * dx / dt          # 20 cells per step: 2.5 blocks
    parts = [{"id": np.arange(n), "x": rng.uniform(0, L, n), "y": rng.uniform(0, L, n),
              "ux": np.where(np.arange(n) % 50 == 0, v, 0.1 * v / 20), "uy": np.where(np.arange(n) % 75 == 0, -v, 0.0)}]
    o = oracle_from(p, parts)
    g = gpu_from(p, parts)
    o.pre_step()
    g.pre_step()
    for it in range(8):
        g.step()
        o.step()
        g.sync()
        assert g.num_particles(0) == n
        assert_close(field_errors(g, o), what=f"far movers fields, iteration {it}")
        assert_close(particle_errors(g, o, p), what=f"far movers particles, iteration {it}")


def test_image_round_trip_is_exact():
    """cpic_b200_image_download / _upload (the e2e path of bench.py): the device state after an
    upload continues bit-identically."""
    import ctypes as C
    a, _, params, _ = pair_from_conf(conf_path("2d-2species-small.conf"))
    b, _, _, _ = pair_from_conf(conf_path("2d-2species-small.conf"))
    for _ in range(4):
        a.step()
        b.step()
    L = b.L
    n = L.cpic_b200_image_bytes(b.h)
    host = L.cpic_b200_host_alloc(n)
    assert host
    for _ in range(5):
        assert L.cpic_b200_image_download(b.h, host, n) == 0
        assert L.cpic_b200_image_upload(b.h, host, n) == 0
        a.step()
        b.step()
    a.sync()
    b.sync()
    L.cpic_b200_host_free(host)
    for name in ("rho", "phi", "Ex", "Ey"):
        assert relerr(b.raw_field(name), a.raw_field(name)) <= TOL
    for i in range(len(params.q)):
        pa, pb = a.particles(i), b.particles(i)
        for k in ("id", "x", "y", "ux", "uy"):
            assert np.array_equal(pa[k], pb[k]), (i, k)


@pytest.mark.parametrize("conf,nprocs", [("2d-2species-small.conf", 1), ("uniform-small.conf", 1), ("two-streams.conf", 1),
                                         ("far-beam.conf", 1), ("2d-2species-small.conf", 4)])
def test_device_initialiser_draws_the_reference_initial_conditions(conf, nprocs):
    """cpic_b200_sim_from_conf_device: glibc's rand() stream advanced by matrix powers on the host and
    drawn on the device in the reference's order (src/particle.c:17-21, 69-73; src/plasma.c:62-128) --
    bit for bit the particles of the host initialiser, and the same fields after sim_init."""
    path = conf_path(conf)
    a = Sim.from_conf(path, ref_nprocs=nprocs)
    b = Sim.from_conf(path, ref_nprocs=nprocs, on_device=True, stream_batch=4096)
    for i in range(a.nspecies):
        pa, pb = a.particles(i), b.particles(i)
        assert np.array_equal(pa["id"], pb["id"])
        for k in ("x", "y", "ux", "uy", "uz"):
            assert np.array_equal(pa[k], pb[k]), (conf, i, k)
    for k in ("rho", "phi", "Ex", "Ey"):
        assert relerr(b.field(k), a.field(k)) <= 1e-13, (conf, k)
    a.close()
    b.close()


def test_step_with_the_state_in_host_memory():
    """cpic_b200_step_host (the e2e path of bench.py): upload, sim_step and download of the particle image
    species by species on three streams give the state of a resident cpic_b200_step (the deposit groups
    absorbed arrivals differently from pending ones: rounding-level differences in rho, hence 1e-12)."""
    import ctypes as C
    path = conf_path("2d-2species-small.conf")
    a = Sim.from_conf(path)
    b = Sim.from_conf(path)
    L = b.L
    nbytes = L.cpic_b200_image_bytes(b.h)
    host = L.cpic_b200_host_alloc(nbytes)
    assert host
    try:
        assert L.cpic_b200_image_download(b.h, host, nbytes) == 0
        for it in range(6):
            a.step()
            rc = L.cpic_b200_step_host(b.h, host, nbytes)
            assert rc == 0, L.cpic_b200_last_error()
        a.sync()
        assert b.iter == a.iter
        for k in ("rho", "phi", "Ex", "Ey"):
            assert relerr(b.raw_field(k)[:, :a.params.nx], a.raw_field(k)[:, :a.params.nx]) <= TOL, k
        for i in range(a.nspecies):
            pa, pb = a.particles(i), b.particles(i)
            assert np.array_equal(pa["id"], pb["id"])
            for k, scale in (("x", a.params.Lx), ("y", a.params.Ly), ("ux", np.abs(pa["ux"]).max()), ("uy", np.abs(pa["uy"]).max())):
                assert np.abs(pa[k] - pb[k]).max() / max(scale, 1e-300) <= TOL, (i, k)
        # and the image itself holds the same particles
        img = (C.c_char * nbytes).from_address(host)
        n0 = np.frombuffer(img, np.int64, 1, 0)[0]
        assert n0 == a.num_particles(0)
    finally:
        L.cpic_b200_host_free(host)
        a.close()
        b.close()


@pytest.mark.parametrize("conf,bands", [("2d-2species-small.conf", 4), ("far-beam.conf", 3), ("uniform-small.conf", 8), ("two-streams.conf", 1)])
def test_banded_step_with_the_state_in_host_memory(conf, bands):
    """cpic_b200_step_host_banded: the image in bands of block rows, every band uploaded, pushed, absorbed,
    packed and downloaded as soon as its neighbours allow (far movers: the species is packed again) -- the
    state of a resident cpic_b200_step to 1e-12, and the image holds exactly the device's particles."""
    import ctypes as C
    path = conf_path(conf)
    a = Sim.from_conf(path)
    b = Sim.from_conf(path)
    L = b.L
    nbytes = L.cpic_b200_banded_image_bytes(b.h, bands)
    assert nbytes > 0
    host = L.cpic_b200_host_alloc(nbytes)
    assert host
    try:
        assert L.cpic_b200_banded_image_download(b.h, host, nbytes, bands) == 0, L.cpic_b200_last_error()
        for it in range(8):
            a.step()
            rc = L.cpic_b200_step_host_banded(b.h, host, nbytes)
            assert rc == 0, L.cpic_b200_last_error()
        a.sync()
        assert b.iter == a.iter
        for k in ("rho", "phi", "Ex", "Ey"):
            assert relerr(b.raw_field(k)[:, :a.params.nx], a.raw_field(k)[:, :a.params.nx]) <= TOL, k
        for i in range(a.nspecies):
            pa, pb = a.particles(i), b.particles(i)
            assert np.array_equal(pa["id"], pb["id"])
            for k, scale in (("x", a.params.Lx), ("y", a.params.Ly), ("ux", np.abs(pa["ux"]).max()), ("uy", np.abs(pa["uy"]).max())):
                assert np.abs(pa[k] - pb[k]).max() / max(scale, 1e-300) <= TOL, (i, k)
        # the image: header, then per species and band n, cap, counts, six arrays of cap entries
        raw = np.frombuffer((C.c_char * nbytes).from_address(host), np.uint8)
        hdr = raw[:32].view(np.int64)
        assert hdr[1] == min(bands, hdr[3]) or hdr[1] <= bands
        off, nb_total = 32, int(hdr[3])
        nbands = int(hdr[1])
        for i in range(a.nspecies):
            ids = []
            for j in range(nbands):
                n, cap = (int(v) for v in raw[off:off + 16].view(np.int64))
                off += 16
                r0, r1 = b_rows(a, nbands, j), b_rows(a, nbands, j + 1)
                nblk = (r1 - r0) * nbx_of(a)
                cnt = raw[off:off + 4 * nblk].view(np.int32)
                assert cnt.sum() == n
                off += (4 * nblk + 7) // 8 * 8
                arr = raw[off:off + 48 * cap].view(np.float64).reshape(6, cap)
                ids.append(arr[5, :n].view(np.int64).copy())
                off += 48 * cap
            ids = np.sort(np.concatenate(ids))
            assert np.array_equal(ids, np.sort(b.particles(i)["id"]))
    finally:
        L.cpic_b200_host_free(host)
        a.close()
        b.close()


def nbx_of(sim):
    return sim.params.nx // 8 if sim.params.nx % 8 == 0 else sim.params.nx // _blk(sim.params.nx)


def _blk(n):
    d = 8
    while d > 1 and n % d:
        d >>= 1
    return d


def b_rows(sim, bands, j):
    nby = sim.params.ny // _blk(sim.params.ny)
    return nby * j // bands


def test_particles_in_host_order():
    """cpic_b200_set_host_order / cpic_b200_get_particles_ordered (the drop-in's list refresh): entry k of every
    array belongs to the k-th id of the host's order, whatever the particle blocks did to the order on the
    device; an id that the order does not know is an error, as is a wrong count."""
    import ctypes as C
    path = conf_path("2d-2species-small.conf")
    s = Sim.from_conf(path)
    s.run(5)
    L = s.L
    rng = np.random.default_rng(7)
    for i in range(s.nspecies):
        ref = s.particles(i)                      # sorted by id
        n = len(ref["id"])
        order = rng.permutation(ref["id"]).astype(np.int64)
        assert L.cpic_b200_set_host_order(s.h, i, n, order.ctypes.data_as(C.c_void_p)) == 0
        out = {k: np.empty(n) for k in ("x", "y", "ux", "uy", "uz", "Ex", "Ey")}
        rc = L.cpic_b200_get_particles_ordered(s.h, i, n, *[out[k].ctypes.data_as(C.c_void_p) for k in ("x", "y", "ux", "uy", "uz", "Ex", "Ey")])
        assert rc == 0, L.cpic_b200_last_error()
        pos = np.searchsorted(ref["id"], order)
        for k in ("x", "y", "ux", "uy", "uz"):
            assert np.array_equal(out[k], ref[k][pos]), (i, k)
        # wrong count
        assert L.cpic_b200_get_particles_ordered(s.h, i, n - 1, *[None] * 7) != 0
        # an order that misses a particle
        short = order[:-1].copy()
        assert L.cpic_b200_set_host_order(s.h, i, n - 1, short.ctypes.data_as(C.c_void_p)) == 0
        assert L.cpic_b200_get_particles_ordered(s.h, i, n - 1, *[out[k].ctypes.data_as(C.c_void_p) for k in ("x", "y", "ux", "uy", "uz", "Ex", "Ey")]) != 0
    s.close()
