"""ctypes binding for oracle/_ref/libcpic_ref*.so (the unmodified reference built by
oracle/Makefile). Test infrastructure only."""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def ref_available(variant="ref"):
    return os.path.exists(os.path.join(REF_DIR, f"libcpic_{variant}.so"))


class RefSim:
    """One reference simulation (sim_init + sim_step), variant 'ref' (unmodified) or
    'ref_acc' (accumulate-correct deposit)."""

    def __init__(self, conf_path, variant="ref"):
        path = os.path.join(REF_DIR, f"libcpic_{variant}.so")
        # RTLD_LOCAL + a private copy per variant keeps the two variants' globals apart
        self.lib = L = C.CDLL(path, mode=os.RTLD_LOCAL)
        L.ref_open.restype = C.c_void_p
        L.ref_open.argtypes = [C.c_char_p]
        for f in ("ref_step", "ref_stage_field_E"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("ref_stage_plasma_E", "ref_stage_plasma_r", "ref_stage_field_rho", "ref_advance_iter"):
            getattr(L, f).restype = None
            getattr(L, f).argtypes = [C.c_void_p]
        L.ref_iter.restype = C.c_longlong
        L.ref_iter.argtypes = [C.c_void_p]
        L.ref_cycles.restype = C.c_longlong
        L.ref_cycles.argtypes = [C.c_void_p]
        L.ref_nspecies.restype = C.c_int
        L.ref_nspecies.argtypes = [C.c_void_p]
        L.ref_nchunks.restype = C.c_longlong
        L.ref_nchunks.argtypes = [C.c_void_p]
        L.ref_scalar.restype = C.c_double
        L.ref_scalar.argtypes = [C.c_void_p, C.c_int]
        L.ref_grid.restype = None
        L.ref_grid.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        L.ref_specie.restype = None
        L.ref_specie.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
        L.ref_get_particles.restype = C.c_longlong
        L.ref_get_particles.argtypes = [C.c_void_p, C.c_int, C.c_longlong] + [C.c_void_p] * 10
        L.ref_collision_census.restype = C.c_longlong
        L.ref_collision_census.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        L.ref_get_field.restype = C.c_longlong
        L.ref_get_field.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_set_field.restype = C.c_longlong
        L.ref_set_field.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ref_timer.restype = C.c_double
        L.ref_timer.argtypes = [C.c_void_p, C.c_int]
        self.h = L.ref_open(os.fsencode(conf_path))
        if not self.h:
            raise RuntimeError(f"reference sim_init failed for {conf_path}")
        nx, ny = C.c_longlong(), C.c_longlong()
        L.ref_grid(self.h, C.byref(nx), C.byref(ny))
        self.nx, self.ny = nx.value, ny.value
        self.nspecies = L.ref_nspecies(self.h)
        names = ["dt", "e0", "Lx", "Ly", "dx", "dy", "Bx", "By", "Bz", "umax_x", "umax_y", "umax_z"]
        for i, n in enumerate(names):
            setattr(self, n, L.ref_scalar(self.h, i))
        self.species = []
        for i in range(self.nspecies):
            q, m, n = C.c_double(), C.c_double(), C.c_longlong()
            L.ref_specie(self.h, i, C.byref(q), C.byref(m), C.byref(n))
            self.species.append((q.value, m.value, n.value))

    @property
    def iter(self):
        return self.lib.ref_iter(self.h)

    def step(self):
        rc = self.lib.ref_step(self.h)
        if rc:
            raise RuntimeError("reference sim_step failed")

    def stage_field_E(self):
        self.lib.ref_stage_field_E(self.h)

    def stage_plasma_E(self):
        self.lib.ref_stage_plasma_E(self.h)

    def stage_plasma_r(self):
        self.lib.ref_stage_plasma_r(self.h)

    def stage_field_rho(self):
        self.lib.ref_stage_field_rho(self.h)

    def advance_iter(self):
        self.lib.ref_advance_iter(self.h)

    def particles(self, species, sort=True):
        """dict of arrays for one species; sorted by particle id when sort=True,
        otherwise in the reference's list order."""
        n = self.lib.ref_get_particles(self.h, species, 0, *([None] * 10))
        out = {"id": np.empty(n, np.int64), "chunk": np.empty(n, np.int64)}
        for k in ("x", "y", "z", "ux", "uy", "uz", "Ex", "Ey"):
            out[k] = np.empty(n, np.float64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self.lib.ref_get_particles(self.h, species, n, p(out["id"]), p(out["x"]), p(out["y"]), p(out["z"]),
                                   p(out["ux"]), p(out["uy"]), p(out["uz"]), p(out["Ex"]), p(out["Ey"]),
                                   p(out["chunk"]))
        if sort:
            o = np.argsort(out["id"], kind="stable")
            out = {k: v[o] for k, v in out.items()}
        return out

    def collision_census(self):
        lost = C.c_longlong()
        packs = self.lib.ref_collision_census(self.h, C.byref(lost))
        return packs, lost.value

    _FIELDS = {"rho": 0, "phi": 1, "Ex": 2, "Ey": 3, "rho_ghost": 4, "phi_ghost": 5}

    def field(self, name):
        which = self._FIELDS[name]
        rows = self.lib.ref_get_field(self.h, which, None)
        out = np.empty((rows, self.nx), np.float64)
        self.lib.ref_get_field(self.h, which, out.ctypes.data_as(C.c_void_p))
        return out

    def set_field(self, name, a):
        a = np.ascontiguousarray(a, np.float64)
        self.lib.ref_set_field(self.h, self._FIELDS[name], a)

    def timer(self, which):
        return self.lib.ref_timer(self.h, which)
