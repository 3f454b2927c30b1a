"""Python view of the C ABI: mirrors the reference's sim_t life cycle
(sim_init / sim_step / the four stage functions, reference src/sim.c)."""
import ctypes as C
from dataclasses import dataclass, field as dfield
from typing import List, Sequence

import numpy as np

from ._lib import lib, check, ParamsC, RunC, MAX_SPECIES

FIELDS = {"rho": 0, "phi": 1, "Ex": 2, "Ey": 3}


@dataclass
class Params:
    """cpic_b200_params_t (what sim_read_config + sim_prepare produce, src/sim.c:38-206)."""
    nx: int
    ny: int
    Lx: float
    Ly: float
    dt: float
    e0: float
    B: Sequence[float] = (0.0, 0.0, 0.0)
    q: Sequence[float] = (-1.0,)
    m: Sequence[float] = (1.0,)
    plasma_chunks: int = 1
    rank: int = 0
    nranks: int = 1
    device: int = -1
    capacity_factor: float = 0.0
    keep_particle_E: bool = False
    outbox_fraction: float = 0.0
    block_cells: int = 0

    def to_c(self):
        p = ParamsC()
        p.nx, p.ny, p.Lx, p.Ly, p.dt, p.e0 = self.nx, self.ny, self.Lx, self.Ly, self.dt, self.e0
        for i in range(3):
            p.B[i] = self.B[i]
        p.plasma_chunks = self.plasma_chunks
        p.nspecies = len(self.q)
        for i, (q, m) in enumerate(zip(self.q, self.m)):
            p.q[i], p.m[i] = q, m
        p.rank, p.nranks, p.device = self.rank, self.nranks, self.device
        p.capacity_factor = self.capacity_factor
        p.keep_particle_E = int(self.keep_particle_E)
        p.outbox_fraction = self.outbox_fraction
        p.block_cells = self.block_cells
        return p

    @staticmethod
    def from_c(p):
        n = p.nspecies
        return Params(p.nx, p.ny, p.Lx, p.Ly, p.dt, p.e0, tuple(p.B), tuple(p.q[:n]), tuple(p.m[:n]),
                      p.plasma_chunks, p.rank, p.nranks, p.device, p.capacity_factor, bool(p.keep_particle_E),
                      p.outbox_fraction, p.block_cells)


@dataclass
class Run:
    cycles: int = 0
    seed: int = 0
    stop_SEM: float = 0.0
    solver: str = "MFT"
    output_enabled: bool = False
    output_path: str = ""
    output_slices: int = 1
    output_alignment: int = 512
    nparticles: List[int] = dfield(default_factory=list)

    @staticmethod
    def from_c(r, nspecies):
        return Run(r.cycles, r.seed, r.stop_SEM, r.solver.decode(), bool(r.output_enabled),
                   r.output_path.decode(), r.output_slices, r.output_alignment, list(r.nparticles[:nspecies]))


def load_conf(path, rank=0, nranks=1, device=-1):
    """sim_read_config + sim_prepare on a cpic `.conf` (host only, no GPU needed)."""
    L = lib()
    h = C.c_void_p()
    check(L.cpic_b200_conf_load(str(path).encode(), C.byref(h)))
    try:
        p, r = ParamsC(), RunC()
        check(L.cpic_b200_conf_params(h, rank, nranks, device, C.byref(p), C.byref(r)))
        return Params.from_c(p), Run.from_c(r, p.nspecies)
    finally:
        L.cpic_b200_conf_free(h)


def init_particles(path, ref_nprocs=1):
    """plasma_init on the host (src/plasma.c, src/particle.c): list of dicts, one per species,
    arrays in id order, bit-identical to the reference started with `ref_nprocs` processes."""
    L = lib()
    h = C.c_void_p()
    check(L.cpic_b200_conf_load(str(path).encode(), C.byref(h)))
    try:
        p, r = ParamsC(), RunC()
        check(L.cpic_b200_conf_params(h, 0, 1, -1, C.byref(p), C.byref(r)))
        ns = p.nspecies
        out = []
        for i in range(ns):
            n = r.nparticles[i]
            out.append({"id": np.zeros(n, np.int64), "x": np.zeros(n), "y": np.zeros(n),
                        "ux": np.zeros(n), "uy": np.zeros(n), "uz": np.zeros(n)})
        arr = lambda k: (C.c_void_p * ns)(*[o[k].ctypes.data for o in out])
        check(L.cpic_b200_conf_init_particles(h, ref_nprocs, arr("id"), arr("x"), arr("y"), arr("ux"), arr("uy")))
        return out
    finally:
        L.cpic_b200_conf_free(h)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Sim:
    """One rank's simulation on one B200. Methods carry the reference's names."""

    def __init__(self, params: Params = None, _handle=None, _run=None):
        self.L = lib()
        self.h = C.c_void_p()
        self.run_info = _run
        if _handle is not None:
            self.h = _handle
            self.params = params
        else:
            self.params = params
            pc = params.to_c()
            check(self.L.cpic_b200_create(C.byref(pc), C.byref(self.h)))
        self.nspecies = len(self.params.q)
        self.ny_local = self.params.ny // self.params.nranks

    @classmethod
    def from_conf(cls, path, rank=0, nranks=1, device=-1, ref_nprocs=1, stream_batch=0, on_device=False):
        """sim_init (src/sim.c:238-320). stream_batch > 0: the population is generated and uploaded in
        batches of that many particles (host memory stays small whatever the population); on_device: the
        reference's initial conditions are drawn on the GPU (glibc rand() stream by jump-ahead)."""
        L = lib()
        h = C.c_void_p()
        r = RunC()
        if on_device:
            check(L.cpic_b200_sim_from_conf_device(str(path).encode(), rank, nranks, device, ref_nprocs,
                                                   stream_batch, C.byref(h), C.byref(r)))
        elif stream_batch > 0:
            check(L.cpic_b200_sim_from_conf_streamed(str(path).encode(), rank, nranks, device, ref_nprocs,
                                                     stream_batch, C.byref(h), C.byref(r)))
        else:
            check(L.cpic_b200_sim_from_conf(str(path).encode(), rank, nranks, device, ref_nprocs, C.byref(h), C.byref(r)))
        params, run = load_conf(path, rank, nranks, device)
        return cls(params, _handle=h, _run=run)

    def close(self):
        if self.h:
            self.L.cpic_b200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- multi GPU
    def comm_id(self):
        buf = (C.c_char * 128)()
        check(self.L.cpic_b200_comm_id(buf))
        return bytes(buf)

    def comm_init(self, id128: bytes):
        buf = (C.c_char * 128).from_buffer_copy(id128)
        check(self.L.cpic_b200_comm_init(self.h, buf))

    # ---- particles
    def set_particles(self, species, id, x, y, ux, uy, uz=None):
        a = [np.ascontiguousarray(id, np.int64)] + [np.ascontiguousarray(v, np.float64) for v in (x, y, ux, uy)]
        uz = None if uz is None else np.ascontiguousarray(uz, np.float64)
        check(self.L.cpic_b200_set_particles(self.h, species, len(a[0]), *[_ptr(v) for v in a], _ptr(uz)))

    def capacity(self, species):
        return self.L.cpic_b200_capacity(self.h, species)

    def occupancy(self, species):
        o = (C.c_int64 * 6)()
        check(self.L.cpic_b200_occupancy(self.h, species, C.byref(o)))
        return dict(zip(("block", "block_cap", "side", "side_cap", "corner", "corner_cap"), o))

    def reserve(self, species, capacity):
        check(self.L.cpic_b200_reserve(self.h, species, capacity))

    def init_uniform(self, species, n, id0=0, vx=0.0, vy=0.0, seed=138):
        check(self.L.cpic_b200_init_uniform(self.h, species, n, id0, vx, vy, seed))

    def init_beam(self, species, n, id0=0, drift=(0.0, 0.0), spread=(0.0, 0.0), seed=138):
        """Device initialiser: uniform positions, u = drift + U(-spread, spread) per axis."""
        check(self.L.cpic_b200_init_beam(self.h, species, n, id0, drift[0], drift[1], spread[0], spread[1], seed))

    def num_particles(self, species):
        return self.L.cpic_b200_num_particles(self.h, species)

    def particles(self, species, sort=True):
        n = self.num_particles(species)
        out = {"id": np.empty(n, np.int64)}
        for k in ("x", "y", "ux", "uy", "uz", "Ex", "Ey"):
            out[k] = np.empty(n, np.float64)
        got = self.L.cpic_b200_get_particles(self.h, species, n, *[_ptr(out[k]) for k in
                                             ("id", "x", "y", "ux", "uy", "uz", "Ex", "Ey")])
        if got != n:
            check(2 if got < 0 else 0)
            raise RuntimeError(f"particle count changed during download ({got} != {n})")
        if sort:
            o = np.argsort(out["id"], kind="stable")
            out = {k: v[o] for k, v in out.items()}
        return out

    # ---- stages (src/sim.c:503,517,525,536)
    def stage_field_E(self):
        check(self.L.cpic_b200_stage_field_E(self.h))

    def stage_plasma_E(self):
        check(self.L.cpic_b200_stage_plasma_E(self.h))

    def stage_plasma_r(self):
        check(self.L.cpic_b200_stage_plasma_r(self.h))

    def stage_field_rho(self):
        check(self.L.cpic_b200_stage_field_rho(self.h))

    def pre_step(self):
        check(self.L.cpic_b200_pre_step(self.h))

    def step(self):
        check(self.L.cpic_b200_step(self.h))

    def step_staged(self):
        """sim_step through the four separate stage calls, as the reference driver makes them."""
        self.stage_field_E()
        self.stage_plasma_E()
        self.stage_plasma_r()
        self.stage_field_rho()
        self.iter = self.iter + 1

    def run(self, steps):
        check(self.L.cpic_b200_run(self.h, steps))

    def run_timed(self, steps):
        """`steps` sim_steps; returns the device time in ms (CUDA events on the sim's stream)."""
        ms = C.c_double()
        check(self.L.cpic_b200_run_timed(self.h, steps, C.byref(ms)))
        return ms.value

    def sync(self):
        check(self.L.cpic_b200_sync(self.h))

    def solve(self):
        check(self.L.cpic_b200_solve(self.h))

    @property
    def iter(self):
        return self.L.cpic_b200_iter(self.h)

    @iter.setter
    def iter(self, v):
        check(self.L.cpic_b200_set_iter(self.h, v))

    # ---- fields
    def field_shape(self, name):
        rows, stride = C.c_int64(), C.c_int64()
        check(self.L.cpic_b200_field_shape(self.h, FIELDS[name], C.byref(rows), C.byref(stride)))
        return rows.value, stride.value

    def raw_field(self, name):
        """The array in the reference's padded layout (rows x stride)."""
        rows, stride = self.field_shape(name)
        out = np.empty((rows, stride), np.float64)
        check(self.L.cpic_b200_get_field(self.h, FIELDS[name], _ptr(out)))
        return out

    def set_raw_field(self, name, a):
        rows, stride = self.field_shape(name)
        a = np.ascontiguousarray(a, np.float64)
        assert a.shape == (rows, stride), (a.shape, rows, stride)
        check(self.L.cpic_b200_set_field(self.h, FIELDS[name], _ptr(a)))

    def field(self, name):
        """Views comparable with the oracle: rho/phi slab (ny x nx); rho_ghost (ny+1 x nx);
        phi_ghost (ny+3 x nx); Ex/Ey (ny+1 x nx)."""
        nx, ny = self.params.nx, self.ny_local
        if name == "rho":
            return self.raw_field("rho")[:ny, :nx]
        if name == "rho_ghost":
            return self.raw_field("rho")[:ny + 1, :nx]
        if name == "phi":
            return self.raw_field("phi")[1:ny + 1, :nx]
        if name == "phi_ghost":
            return self.raw_field("phi")[:, :nx]
        if name in ("Ex", "Ey"):
            return self.raw_field(name)
        raise KeyError(name)

    def energy(self):
        ke, pe = C.c_double(), C.c_double()
        check(self.L.cpic_b200_energy(self.h, C.byref(ke), C.byref(pe)))
        return ke.value, pe.value

    # ---- measurement
    def timing(self, enable=True):
        check(self.L.cpic_b200_timing(self.h, int(enable)))

    def get_timing(self):
        ms = (C.c_double * 6)()
        n = C.c_int64()
        check(self.L.cpic_b200_get_timing(self.h, C.byref(ms), C.byref(n)))
        return dict(zip(("field_E", "gather_push", "exchange", "field_rho", "solver", "gather"), ms)), n.value
