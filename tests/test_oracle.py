"""CPU tests that pin the oracle (oracle/cpic_oracle.c) before anything is trusted to it:
against the reference's own golden vectors and known answers, and -- when the compiled
reference is present (oracle/_ref, built from /root/reference by oracle/Makefile) -- against
the unmodified reference step by step."""
import os

import numpy as np
import pytest

from conftest import conf_path, ROOT
from _oraclebind import OracleSim, rfft2, irfft2
from _parity import oracle_from, relerr, TOL
from _refbind import RefSim, ref_available
from cpic_b200 import load_conf, init_particles

GOLDEN = os.path.join(ROOT, "tests", "golden")
needs_ref = pytest.mark.skipif(not ref_available("ref_acc"), reason="oracle/_ref not built (needs /root/reference)")


def oracle_for(conf):
    p, run = load_conf(conf_path(conf))
    o = oracle_from(p, init_particles(conf_path(conf)))
    o.pre_step()
    return o, p, run


@pytest.mark.parametrize("shape", [(8, 8), (64, 32), (16, 128), (12, 20), (1, 4)])
def test_fft_against_numpy(shape):
    """The stand-in for FFTW (third party, not vendored: fftw 3.3.6, src/build.mk:47-48)."""
    ny, nx = shape
    a = np.random.default_rng(ny * 131 + nx).standard_normal((ny, nx))
    g = np.fft.rfft2(a)
    assert np.abs(rfft2(a) - g).max() <= 1e-13 * max(1.0, np.abs(g).max())
    assert np.abs(irfft2(g, nx) / a.size - a).max() <= 1e-13


def test_cic_known_answer():
    """test/interpolate.c.disabled:9-27: x=(3.25,3.75), x0=(2,1), dx=1 gives
    w00=3/16, w10=1/16, w01=9/16, w11=3/16 and i0=(1,2)."""
    o = OracleSim(8, 8, 8.0, 8.0, 0.1, 1.0, (0, 0, 0), [(-1.0, 1.0)])
    w, ix, iy = o.weights(3.25 - 2.0, 3.75 - 1.0)
    assert (ix, iy) == (1, 2)
    assert w == [3 / 16, 9 / 16, 1 / 16, 3 / 16]      # order: w00, w01, w10, w11


def test_weights_use_dx_x_for_y():
    """SURVEY F2 (src/interpolate.c:87-88): with dx != dy the Y offset is formed with dx[X]."""
    o = OracleSim(8, 8, 8.0, 4.0, 0.1, 1.0, (0, 0, 0), [(-1.0, 1.0)])      # dx = 1, dy = 0.5
    w, ix, iy = o.weights(0.25, 1.3)
    rely = (1.3 - 2 * 1.0) * 2.0         # floor(1.3/0.5) = 2 cells, times dx[X] = 1.0 (not 0.5)
    assert iy == 2 and abs((w[1] + w[3]) - rely) < 1e-15


def test_harmonic_line_probe():
    """test/harmonic/E.csv: ParaView probe along y=4 of the 64x64 harmonic case at start:
    E_X, phi and rho, 6 significant digits."""
    o, p, _ = oracle_for("harmonic-64.conf")
    E = np.genfromtxt(os.path.join(GOLDEN, "E.csv"), delimiter=",", names=True)
    ix = np.round(E["Points_0"] / (p.Lx / p.nx)).astype(int)
    iy = int(round(4.0 / (p.Ly / p.ny)))
    sel = ix < p.nx
    assert np.abs(o.field("phi")[iy][ix[sel]] - E["phi"][sel]).max() < 2e-6
    assert np.abs(o.field("Ex")[iy][ix[sel]] - E["E_X"][sel]).max() < 5e-6
    assert np.abs(o.field("rho")[iy][ix[sel]] - E["rho"][sel]).max() < 1e-6
    assert np.abs(o.field("Ey")[iy][ix[sel]] - E["E_Y"][sel]).max() < 5e-6


def test_harmonic_trajectory_golden():
    """test/harmonic/harm.r0x, harm.E0x: x and E_x of particle 0 per iteration of
    conf/harmonic.conf (1024^2), as printed by test/harmonic.c:112-120 (7 digits)."""
    r0 = np.loadtxt(os.path.join(GOLDEN, "harm.r0x"))
    E0 = np.loadtxt(os.path.join(GOLDEN, "harm.E0x"))
    o, _, _ = oracle_for("harmonic.conf")
    n = 40
    xs, Es = [], []
    for _ in range(n):
        o.step()
        p = o.particles(0)
        xs.append(p["x"][0])
        Es.append(p["Ex"][0])
    assert np.abs(np.array(xs) - r0[:n]).max() < 1e-6
    assert np.abs(np.array(Es) - E0[:n]).max() < 1e-8


def test_cyclotron_analytic():
    """test/cyclotron.c:108-131,199-215: the radius error stays below v*dt^2."""
    o, p, _ = oracle_for("cyclotron.conf")
    q = o.particles(0)
    u = np.array([q["ux"][0], q["uy"][0], 0.0])
    v = np.linalg.norm(u)
    radius = v / (abs(p.q[0]) * p.B[2] / p.m[0])
    tmp = np.cross(u, np.array(p.B))
    tmp *= radius / np.linalg.norm(tmp)
    center = np.array([q["x"][0], q["y"][0], 0.0]) + tmp
    worst = 0.0
    for _ in range(1500):
        o.step()
        q = o.particles(0)
        worst = max(worst, abs(np.hypot(q["x"][0] - center[0], q["y"][0] - center[1]) - abs(radius)))
    assert worst < v * p.dt ** 2


def test_constant_speed():
    """test/constant-speed.c:13,:78-91: no self force, velocity constant to 1e-10."""
    o, _, _ = oracle_for("constant-speed.conf")
    u0 = o.particles(0)
    for _ in range(200):
        o.step()
    u1 = o.particles(0)
    assert abs(u1["ux"][0] - u0["ux"][0]) < 1e-10 and abs(u1["uy"][0] - u0["uy"][0]) < 1e-10


@needs_ref
@pytest.mark.parametrize("conf", ["uniform-small.conf", "2d-2species-small.conf", "two-streams.conf",
                                  "2d-2species-delta.conf", "harmonic-64.conf"])
def test_oracle_against_reference(conf):
    """Ten sim_steps of the reference's own objects (accumulate-correct build, see F1) and of
    the restatement from the same post-sim_init particles: fields and particles to 1e-12."""
    r = RefSim(conf_path(conf), "ref_acc")
    o = OracleSim(r.nx, r.ny, r.Lx, r.Ly, r.dt, r.e0, (r.Bx, r.By, r.Bz), [(q, m) for q, m, n in r.species],
                  r.lib.ref_nchunks(r.h))
    for i in range(r.nspecies):
        p = r.particles(i)
        o.set_particles(i, p["id"], p["x"], p["y"], p["ux"], p["uy"], p["uz"])
    o.pre_step()
    for it in range(11):
        for k in ("rho_ghost", "phi", "Ex", "Ey"):
            assert relerr(o.field(k), r.field(k)) <= TOL, (conf, it, k)
        for i in range(r.nspecies):
            a, b = o.particles(i), r.particles(i)
            assert (a["id"] == b["id"]).all()
            assert np.abs(a["x"] - b["x"]).max() / r.Lx <= TOL and np.abs(a["y"] - b["y"]).max() / r.Ly <= TOL
            umax = max(np.abs(b["ux"]).max(), np.abs(b["uy"]).max(), 1e-300)
            for k in ("ux", "uy"):
                assert np.abs(a[k] - b[k]).max() / umax <= TOL, (conf, it, k)
        r.step()
        o.step()


@needs_ref
@pytest.mark.parametrize("conf", ["2d-2species-delta.conf", "uniform-small.conf"])
def test_f1_lost_deposits(conf):
    """the unmodified reference drops deposits when two lanes of a pack share a cell
    (src/simd_avx2.h:226-249). The census counts them, the lossy emulation reproduces the
    reference's rho bit for bit, and the accumulate-correct oracle differs exactly there."""
    r = RefSim(conf_path(conf), "ref")
    nch = r.lib.ref_nchunks(r.h)
    o = OracleSim(r.nx, r.ny, r.Lx, r.Ly, r.dt, r.e0, (r.Bx, r.By, r.Bz), [(q, m) for q, m, n in r.species], nch)
    packs, lost = r.collision_census()
    assert packs > 0 and lost >= packs
    for _ in range(3):
        o.rho_raw[:] = 0
        for ic in range(nch):
            for i in range(r.nspecies):
                p = r.particles(i, sort=False)
                sel = p["chunk"] == ic
                o.deposit_lossy(r.species[i][0], p["x"][sel], p["y"][sel], ic * r.Lx / nch, 0.0)
        o.rho_raw[0, :r.nx] += o.rho_raw[r.ny, :r.nx]
        ref = r.field("rho_ghost")
        assert relerr(o.field("rho_ghost"), ref) < 1e-15
        r.step()


@needs_ref
def test_reference_reproduces_its_golden_vectors():
    """The compiled reference (behind the MPI/libconfig/FFTW shims) against its own
    harm.r0x / harm.E0x: this is what pins oracle/_ref itself."""
    r0 = np.loadtxt(os.path.join(GOLDEN, "harm.r0x"))
    E0 = np.loadtxt(os.path.join(GOLDEN, "harm.E0x"))
    r = RefSim(conf_path("harmonic.conf"), "ref")
    n = 25
    xs, Es = [], []
    for _ in range(n):
        r.step()
        p = r.particles(0)
        xs.append(p["x"][0])
        Es.append(p["Ex"][0])
    assert np.abs(np.array(xs) - r0[:n]).max() < 1e-6
    assert np.abs(np.array(Es) - E0[:n]).max() < 1e-8


def test_two_stream_growth_rate():
    """BASELINE configs[0]: the k=1 mode of two cold beams (v0 = +-1, total plasma frequency
    sqrt(2), i.e. beam frequency 1) grows at gamma = 0.486 in cold-fluid theory,
    w^2 = k^2 v0^2 + wp^2 - wp sqrt(4 k^2 v0^2 + wp^2). The oracle's field energy must grow at
    2*gamma within 15% over the linear phase."""
    o, p, _ = oracle_for("two-streams.conf")
    t, ex2 = [], []
    for it in range(360):
        o.step()
        t.append((it + 1) * p.dt)
        ex2.append((o.field("Ex")[:p.ny] ** 2).sum())
    t, ex2 = np.array(t), np.array(ex2)
    sel = (t > 6.0) & (t < 14.0)
    slope = np.polyfit(t[sel], np.log(ex2[sel]), 1)[0]
    gamma = np.sqrt(np.sqrt(5.0) - 2.0)
    assert abs(slope / 2.0 - gamma) / gamma < 0.15, (slope / 2.0, gamma)


# ---- the multi-process shims behind the CPU baseline (oracle/shim/shim_mpi_mp.c, shim_fftw.c -DSHIM_MP)

def _mp_run(conf, nprocs, steps, tmp):
    """oracle/_ref/ref_mp_check: the unmodified reference (accumulate-correct deposit) over `nprocs`
    forked ranks; returns the assembled slabs and the particles sorted by id."""
    import subprocess
    d = os.path.join(tmp, f"p{nprocs}")
    os.makedirs(d, exist_ok=True)
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_mp_check")
    r = subprocess.run([exe, conf, str(steps), d], env=dict(os.environ, CPIC_SHIM_NPROCS=str(nprocs)),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    fields = {k: [] for k in ("rho", "phi", "Ex", "Ey")}
    parts = None
    for rank in range(nprocs):
        b = open(os.path.join(d, f"rank{rank}.bin"), "rb").read()
        nx, ny, ns = (int(v) for v in np.frombuffer(b, np.int64, 3, 0))
        off = 24
        for k in fields:
            rows = int(np.frombuffer(b, np.int64, 1, off)[0])
            off += 8
            fields[k].append(np.frombuffer(b, np.float64, rows * nx, off).reshape(rows, nx)[:ny])
            off += 8 * rows * nx
        if parts is None:
            parts = [{k: [] for k in ("id", "x", "y", "ux", "uy")} for _ in range(ns)]
        for s in range(ns):
            n = int(np.frombuffer(b, np.int64, 1, off)[0])
            off += 8
            parts[s]["id"].append(np.frombuffer(b, np.int64, n, off))
            off += 8 * n
            for k in ("x", "y", "ux", "uy"):
                parts[s][k].append(np.frombuffer(b, np.float64, n, off))
                off += 8 * n
    out = []
    for s in parts:
        q = {k: np.concatenate(v) for k, v in s.items()}
        o = np.argsort(q["id"], kind="stable")
        out.append({k: v[o] for k, v in q.items()})
    return {k: np.vstack(v) for k, v in fields.items()}, out


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_mp_check")),
                    reason="oracle/_ref/ref_mp_check not built (needs /root/reference)")
@pytest.mark.parametrize("conf,ranks", [("two-streams.conf", (2, 4)), ("2d-2species-delta.conf", (2,))])
def test_reference_over_forked_ranks_equals_single_rank(conf, ranks, tmp_path):
    """The CPU baseline runs the unmodified reference with one forked rank per core over socket and
    shared-memory stand-ins for MPI and FFTW-MPI: 10 steps on P ranks must reproduce the single-rank
    run (position-delta initial conditions do not depend on the number of ranks)."""
    F1, P1 = _mp_run(conf_path(conf), 1, 10, str(tmp_path))
    for nprocs in ranks:
        F, P = _mp_run(conf_path(conf), nprocs, 10, str(tmp_path))
        for k in F1:
            assert relerr(F[k], F1[k]) <= TOL, (conf, nprocs, k)
        for a, b in zip(P, P1):
            assert np.array_equal(a["id"], b["id"])
            for k, scale in (("x", 1.0), ("y", 1.0), ("ux", np.abs(b["ux"]).max()), ("uy", max(np.abs(b["uy"]).max(), 1e-300))):
                assert np.abs(a[k] - b[k]).max() / scale <= 1e-11, (conf, nprocs, k)
