"""ctypes binding for oracle/libcpic_oracle.so (the plain-C restatement). Test
infrastructure only: used by tests/, __graft_entry__.smoke() and bench.py's CPU baseline."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libcpic_oracle.so")


class _Species(C.Structure):
    _fields_ = [("q", C.c_double), ("m", C.c_double), ("n", C.c_longlong), ("id", C.POINTER(C.c_longlong))] + \
        [(k, C.POINTER(C.c_double)) for k in ("x", "y", "z", "ux", "uy", "uz", "Ex", "Ey")]


class _Sim(C.Structure):
    _fields_ = [("nx", C.c_longlong), ("ny", C.c_longlong), ("S", C.c_longlong),
                ("L", C.c_double * 2), ("dx", C.c_double * 2), ("dt", C.c_double), ("e0", C.c_double),
                ("B", C.c_double * 3), ("umax", C.c_double * 3), ("iter", C.c_longlong),
                ("nspecies", C.c_int), ("sp", C.POINTER(_Species)),
                ("rho", C.POINTER(C.c_double)), ("phi", C.POINTER(C.c_double)),
                ("Ex", C.POINTER(C.c_double)), ("Ey", C.POINTER(C.c_double)),
                ("G", C.POINTER(C.c_double)), ("gre", C.POINTER(C.c_double)), ("gim", C.POINTER(C.c_double)),
                ("aborted", C.c_int)]


def build():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(ORACLE_DIR, "cpic_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        dp = C.POINTER(C.c_double)
        L.oracle_create.restype = C.POINTER(_Sim)
        L.oracle_create.argtypes = [C.c_longlong, C.c_longlong, C.c_double, C.c_double, C.c_double, C.c_double,
                                    dp, C.c_longlong, C.c_int, dp, dp]
        L.oracle_destroy.argtypes = [C.POINTER(_Sim)]
        L.oracle_srand.argtypes = [C.c_uint]
        L.oracle_alloc_species.argtypes = [C.POINTER(_Sim), C.c_int, C.c_longlong]
        L.oracle_init_randpos_chunk.argtypes = [C.POINTER(_Sim), C.c_int, C.c_longlong, C.c_longlong, dp]
        L.oracle_init_delta.argtypes = [C.POINTER(_Sim), C.c_int, dp, dp, dp]
        L.oracle_set_particles.argtypes = [C.POINTER(_Sim), C.c_int, C.c_longlong] + [C.c_void_p] * 6
        for f in ("oracle_stage_field_rho", "oracle_stage_field_E", "oracle_stage_plasma_E", "oracle_pre_step",
                  "oracle_solve", "oracle_phi_ghosts", "oracle_field_E"):
            getattr(L, f).argtypes = [C.POINTER(_Sim)]
            getattr(L, f).restype = None
        for f in ("oracle_stage_plasma_r", "oracle_step"):
            getattr(L, f).argtypes = [C.POINTER(_Sim)]
            getattr(L, f).restype = C.c_int
        L.oracle_weights.argtypes = [C.POINTER(_Sim), C.c_double, C.c_double, dp,
                                     C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        L.oracle_rfft2.argtypes = [C.c_longlong, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]
        L.oracle_irfft2.argtypes = [C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong]
        L.oracle_deposit_lossy.argtypes = [C.POINTER(_Sim), C.c_double, C.c_longlong, C.c_void_p, C.c_void_p,
                                           C.c_double, C.c_double]
        L.oracle_kinetic_energy.argtypes = [C.POINTER(_Sim)]
        L.oracle_kinetic_energy.restype = C.c_double
        L.oracle_field_energy.argtypes = [C.POINTER(_Sim)]
        L.oracle_field_energy.restype = C.c_double
        _lib = L
    return _lib


def _darr(v):
    return (C.c_double * len(v))(*[float(x) for x in v])


class OracleSim:
    def __init__(self, nx, ny, Lx, Ly, dt, e0, B, species_qm, plasma_chunks=1):
        """species_qm: list of (q, m)."""
        self.L = lib()
        q = _darr([s[0] for s in species_qm])
        m = _darr([s[1] for s in species_qm])
        self.p = self.L.oracle_create(nx, ny, Lx, Ly, dt, e0, _darr(B), plasma_chunks, len(species_qm), q, m)
        self.s = self.p.contents
        self.nx, self.ny, self.S = nx, ny, self.s.S
        self.nspecies = len(species_qm)

    def __del__(self):
        try:
            self.L.oracle_destroy(self.p)
        except Exception:
            pass

    # --- particles
    def set_particles(self, i, id, x, y, ux, uy, uz=None):
        a = [np.ascontiguousarray(id, np.int64)] + [np.ascontiguousarray(v, np.float64) for v in (x, y, ux, uy)]
        a.append(np.ascontiguousarray(uz if uz is not None else np.zeros(len(a[0])), np.float64))
        self.L.oracle_set_particles(self.p, i, len(a[0]), *[v.ctypes.data_as(C.c_void_p) for v in a])

    def srand(self, seed):
        self.L.oracle_srand(seed)

    def alloc(self, i, n):
        self.L.oracle_alloc_species(self.p, i, n)

    def init_randpos_chunk(self, i, ic, nchunks, v):
        self.L.oracle_init_randpos_chunk(self.p, i, ic, nchunks, _darr(v))

    def init_delta(self, i, r0, dr, v):
        self.L.oracle_init_delta(self.p, i, _darr(r0), _darr(dr), _darr(v))

    def particles(self, i):
        sp = self.s.sp[i]
        n = sp.n
        out = {"id": np.ctypeslib.as_array(sp.id, (n,)).copy() if n else np.empty(0, np.int64)}
        for k in ("x", "y", "z", "ux", "uy", "uz", "Ex", "Ey"):
            out[k] = np.ctypeslib.as_array(getattr(sp, k), (n,)).copy() if n else np.empty(0)
        return out

    # --- fields (views into the oracle's memory)
    def _view(self, ptr, rows, ld):
        return np.ctypeslib.as_array(ptr, (rows, ld))

    @property
    def rho_raw(self):
        return self._view(self.s.rho, self.ny + 1, self.S)

    @property
    def phi_raw(self):
        return self._view(self.s.phi, self.ny + 3, self.S)

    def field(self, name):
        nx, ny = self.nx, self.ny
        if name == "rho":
            return self.rho_raw[:ny, :nx].copy()
        if name == "rho_ghost":
            return self.rho_raw[:ny + 1, :nx].copy()
        if name == "phi":
            return self.phi_raw[1:ny + 1, :nx].copy()
        if name == "phi_ghost":
            return self.phi_raw[:, :nx].copy()
        if name == "Ex":
            return self._view(self.s.Ex, ny + 1, nx).copy()
        if name == "Ey":
            return self._view(self.s.Ey, ny + 1, nx).copy()
        raise KeyError(name)

    def set_rho(self, a):
        self.rho_raw[:a.shape[0], :self.nx] = a

    def set_E(self, ex, ey):
        self._view(self.s.Ex, self.ny + 1, self.nx)[:] = ex
        self._view(self.s.Ey, self.ny + 1, self.nx)[:] = ey

    @property
    def iter(self):
        return self.s.iter

    @iter.setter
    def iter(self, v):
        self.s.iter = v

    def pre_step(self):
        self.L.oracle_pre_step(self.p)

    def step(self):
        if self.L.oracle_step(self.p):
            raise RuntimeError("oracle: velocity limit exceeded (check_velocity)")

    def stage_field_rho(self):
        self.L.oracle_stage_field_rho(self.p)

    def stage_field_E(self):
        self.L.oracle_stage_field_E(self.p)

    def stage_plasma_E(self):
        self.L.oracle_stage_plasma_E(self.p)

    def stage_plasma_r(self):
        if self.L.oracle_stage_plasma_r(self.p):
            raise RuntimeError("oracle: velocity limit exceeded (check_velocity)")

    def solve(self):
        self.L.oracle_solve(self.p)

    def weights(self, x, y):
        w = (C.c_double * 4)()
        ix, iy = C.c_longlong(), C.c_longlong()
        self.L.oracle_weights(self.p, x, y, w, C.byref(ix), C.byref(iy))
        return list(w), ix.value, iy.value

    def deposit_lossy(self, q, x, y, gx, gy):
        x = np.ascontiguousarray(x, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        self.L.oracle_deposit_lossy(self.p, q, len(x), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p),
                                    gx, gy)

    def kinetic_energy(self):
        return self.L.oracle_kinetic_energy(self.p)

    def field_energy(self):
        return self.L.oracle_field_energy(self.p)


def rfft2(a):
    ny, nx = a.shape
    a = np.ascontiguousarray(a, np.float64)
    re = np.empty((ny, nx // 2 + 1))
    im = np.empty((ny, nx // 2 + 1))
    lib().oracle_rfft2(ny, nx, a.ctypes.data_as(C.c_void_p), nx, re.ctypes.data_as(C.c_void_p),
                       im.ctypes.data_as(C.c_void_p))
    return re + 1j * im


def irfft2(g, nx):
    ny = g.shape[0]
    re = np.ascontiguousarray(g.real)
    im = np.ascontiguousarray(g.imag)
    out = np.empty((ny, nx))
    lib().oracle_irfft2(ny, nx, re.ctypes.data_as(C.c_void_p), im.ctypes.data_as(C.c_void_p),
                        out.ctypes.data_as(C.c_void_p), nx)
    return out
