"""The CUDA path against the compiled, UNMODIFIED reference itself on the B200 (oracle/_ref travels
to the GPU box): the chain GPU = restatement = reference closed in one test, plus the part of
SURVEY F1 that can only be stated against the unmodified build."""
import os

import numpy as np
import pytest

from conftest import conf_path
from _parity import relerr, TOL, gpu_from
from _refbind import RefSim, ref_available
from cpic_b200 import load_conf, init_particles

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_available("ref_acc"), reason="oracle/_ref not built (needs /root/reference at build time)")]


def _compare(g, r, params, tag):
    g.sync()
    for k in ("rho_ghost", "phi_ghost", "Ex", "Ey"):
        assert relerr(g.field(k), r.field(k)) <= TOL, f"{tag}: {k} {relerr(g.field(k), r.field(k))}"
    for i in range(len(params.q)):
        a, b = g.particles(i), r.particles(i)
        assert np.array_equal(a["id"], b["id"]), tag
        umax = max(np.abs(b["ux"]).max(), np.abs(b["uy"]).max(), 1e-300)
        for k, scale in (("x", params.Lx), ("y", params.Ly), ("ux", umax), ("uy", umax)):
            assert np.abs(a[k] - b[k]).max() / scale <= TOL, f"{tag}: species {i} {k}"


@pytest.mark.parametrize("conf", ["2d-2species-small.conf", "two-streams.conf", "uniform-small.conf"])
def test_first_10_steps_against_the_compiled_reference(conf):
    """sim_init + iterations 0..9 of the reference's own objects (accumulate-correct deposit variant,
    SURVEY F1) against cpic_b200 on identical initial conditions: rho, phi, E and particles to 1e-12."""
    path = conf_path(conf)
    params, _ = load_conf(path)
    r = RefSim(path, "ref_acc")
    g = gpu_from(params, init_particles(path))
    g.pre_step()
    _compare(g, r, params, f"{conf} after sim_init")
    for it in range(10):
        g.step()
        r.step()
        _compare(g, r, params, f"{conf} iteration {it}")
    g.close()


def test_unmodified_reference_differs_only_by_its_lost_deposits():
    """against the UNMODIFIED build rho differs exactly where the census says deposits were
    dropped (src/simd_avx2.h:226-249), and nowhere else."""
    path = conf_path("2d-2species-small.conf")
    params, _ = load_conf(path)
    r = RefSim(path, "ref")
    g = gpu_from(params, init_particles(path))
    g.pre_step()
    g.sync()
    packs, lost = r.collision_census()
    d = np.abs(g.field("rho_ghost") - r.field("rho_ghost"))
    scale = np.abs(r.field("rho_ghost")).max()
    nodes = int((d > 1e-12 * scale).sum())
    if lost == 0:
        assert nodes == 0
    else:
        # every lost deposit touches at most four nodes
        assert 0 < nodes <= 4 * lost
    g.close()
