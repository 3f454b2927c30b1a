"""Shared helpers of the parity tests: build the oracle and the GPU simulation from the
same `.conf`, and compare them with the norms SURVEY section 8c/8d defines."""
import numpy as np

import cpic_b200
from cpic_b200 import Sim, load_conf, init_particles
from _oraclebind import OracleSim

TOL = 1e-12   # BASELINE.json north_star: rho, phi, E and particle x/u within 1e-12 relative


def relerr(a, b):
    """max|a-b| / max|b| (the max-norm-relative measure of SURVEY 8c)."""
    a = np.asarray(a)
    b = np.asarray(b)
    scale = np.abs(b).max() if b.size else 0.0
    d = np.abs(a - b).max() if b.size else 0.0
    return d / scale if scale > 0 else d


def oracle_from(params, parts):
    o = OracleSim(params.nx, params.ny, params.Lx, params.Ly, params.dt, params.e0, params.B,
                  list(zip(params.q, params.m)), params.plasma_chunks)
    for i, p in enumerate(parts):
        o.set_particles(i, p["id"], p["x"], p["y"], p["ux"], p["uy"], p.get("uz"))
    return o


def gpu_from(params, parts, **kw):
    for k, v in kw.items():
        setattr(params, k, v)
    s = Sim(params)
    for i, p in enumerate(parts):
        s.set_particles(i, p["id"], p["x"], p["y"], p["ux"], p["uy"], p.get("uz"))
    return s


def pair_from_conf(path, **kw):
    params, run = load_conf(path)
    parts = init_particles(path)
    o = oracle_from(params, parts)
    g = gpu_from(params, parts, **kw)
    o.pre_step()
    g.pre_step()
    return g, o, params, run


def field_errors(g, o, names=("rho_ghost", "phi_ghost", "Ex", "Ey")):
    return {k: relerr(g.field(k), o.field(k)) for k in names}


def particle_errors(g, o, params, with_E=False):
    """Per-particle |dx|/L and |du|/max|u|, particles keyed by id."""
    worst = {}
    for i in range(len(params.q)):
        a, b = g.particles(i), o.particles(i)
        assert len(a["id"]) == len(b["id"]), f"species {i}: {len(a['id'])} particles on the GPU, {len(b['id'])} in the oracle"
        assert (a["id"] == b["id"]).all()
        e = {"x": np.abs(a["x"] - b["x"]).max(initial=0.0) / params.Lx,
             "y": np.abs(a["y"] - b["y"]).max(initial=0.0) / params.Ly}
        umax = max(np.abs(b["ux"]).max(initial=0.0), np.abs(b["uy"]).max(initial=0.0), 1e-300)
        e["ux"] = np.abs(a["ux"] - b["ux"]).max(initial=0.0) / umax
        e["uy"] = np.abs(a["uy"] - b["uy"]).max(initial=0.0) / umax
        e["uz"] = np.abs(a["uz"] - b["uz"]).max(initial=0.0)
        if with_E:
            # the scale of E is that of the grid field (a lone particle gathers ~0 from its own field)
            emax = max(np.abs(o.field("Ex")).max(), np.abs(o.field("Ey")).max(), 1e-300)
            e["Ex"] = np.abs(a["Ex"] - b["Ex"]).max(initial=0.0) / emax
            e["Ey"] = np.abs(a["Ey"] - b["Ey"]).max(initial=0.0) / emax
        for k, v in e.items():
            worst[k] = max(worst.get(k, 0.0), v)
    return worst


def assert_close(errs, tol=TOL, what=""):
    bad = {k: v for k, v in errs.items() if not (v <= tol)}
    assert not bad, f"{what}: beyond {tol:g}: {bad} (all: {errs})"
