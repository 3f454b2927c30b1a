"""Loader and prototypes for libcpic_b200.so (include/cpic_b200.h)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
MAX_SPECIES = 8


class Cpic_b200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cpic_b200 error {code}: {msg}")
        self.code = code


def lib_path():
    # CPIC_B200_LIB: an alternative build of the same library (kernel tuning experiments)
    return os.environ.get("CPIC_B200_LIB") or os.path.join(_HERE, "libcpic_b200.so")


def build(force=False):
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-s", "-C", os.path.join(_HERE, "csrc")]
    if force:
        subprocess.check_call(cmd + ["clean"])
    subprocess.check_call(cmd)
    return lib_path()


class ParamsC(C.Structure):
    _fields_ = [("nx", C.c_int64), ("ny", C.c_int64), ("Lx", C.c_double), ("Ly", C.c_double),
                ("dt", C.c_double), ("e0", C.c_double), ("B", C.c_double * 3),
                ("plasma_chunks", C.c_int64), ("nspecies", C.c_int32),
                ("q", C.c_double * MAX_SPECIES), ("m", C.c_double * MAX_SPECIES),
                ("rank", C.c_int32), ("nranks", C.c_int32), ("device", C.c_int32),
                ("capacity_factor", C.c_double), ("keep_particle_E", C.c_int32),
                ("block_cells", C.c_int32), ("outbox_fraction", C.c_double)]


class RunC(C.Structure):
    _fields_ = [("cycles", C.c_int64), ("seed", C.c_uint32), ("stop_SEM", C.c_double),
                ("period_energy", C.c_int64), ("period_field", C.c_int64), ("period_particle", C.c_int64),
                ("solver", C.c_char * 16), ("output_enabled", C.c_int32), ("output_path", C.c_char * 4096),
                ("output_slices", C.c_int64), ("output_alignment", C.c_int64),
                ("nparticles", C.c_int64 * MAX_SPECIES)]


_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes): every symbol include/cpic_b200.h declares
EXPORTS = {
    "cpic_b200_version": (C.c_char_p, []),
    "cpic_b200_last_error": (C.c_char_p, []),
    "cpic_b200_create": (_i, [C.POINTER(ParamsC), _pp]),
    "cpic_b200_destroy": (None, [_vp]),
    "cpic_b200_comm_id": (_i, [_vp]),
    "cpic_b200_comm_init": (_i, [_vp, _vp]),
    "cpic_b200_set_particles": (_i, [_vp, _i, _i64] + [_vp] * 6),
    "cpic_b200_capacity": (_i64, [_vp, _i]),
    "cpic_b200_reserve": (_i, [_vp, _i, _i64]),
    "cpic_b200_count_particles": (_i, [_vp, _i, _i64, _vp, _vp]),
    "cpic_b200_reserve_counted": (_i, [_vp, _i]),
    "cpic_b200_add_particles": (_i, [_vp, _i, _i64] + [_vp] * 6),
    "cpic_b200_occupancy": (_i, [_vp, _i, C.POINTER(_i64 * 6)]),
    "cpic_b200_num_particles": (_i64, [_vp, _i]),
    "cpic_b200_get_particles": (_i64, [_vp, _i, _i64] + [_vp] * 8),
    "cpic_b200_set_host_order": (_i, [_vp, _i, _i64, _vp]),
    "cpic_b200_get_particles_ordered": (_i, [_vp, _i, _i64] + [_vp] * 7),
    "cpic_b200_init_uniform": (_i, [_vp, _i, _i64, _i64, _d, _d, C.c_uint64]),
    "cpic_b200_init_beam": (_i, [_vp, _i, _i64, _i64, _d, _d, _d, _d, C.c_uint64]),
    "cpic_b200_stage_field_E": (_i, [_vp]),
    "cpic_b200_stage_plasma_E": (_i, [_vp]),
    "cpic_b200_stage_plasma_r": (_i, [_vp]),
    "cpic_b200_stage_field_rho": (_i, [_vp]),
    "cpic_b200_pre_step": (_i, [_vp]),
    "cpic_b200_step": (_i, [_vp]),
    "cpic_b200_run": (_i, [_vp, _i64]),
    "cpic_b200_run_timed": (_i, [_vp, _i64, C.POINTER(_d)]),
    "cpic_b200_iter": (_i64, [_vp]),
    "cpic_b200_set_iter": (_i, [_vp, _i64]),
    "cpic_b200_sync": (_i, [_vp]),
    "cpic_b200_field_shape": (_i, [_vp, _i, C.POINTER(_i64), C.POINTER(_i64)]),
    "cpic_b200_get_field": (_i, [_vp, _i, _vp]),
    "cpic_b200_set_field": (_i, [_vp, _i, _vp]),
    "cpic_b200_solve": (_i, [_vp]),
    "cpic_b200_energy": (_i, [_vp, C.POINTER(_d), C.POINTER(_d)]),
    "cpic_b200_timing": (_i, [_vp, _i]),
    "cpic_b200_get_timing": (_i, [_vp, C.POINTER(_d * 6), C.POINTER(_i64)]),
    "cpic_b200_image_bytes": (_i64, [_vp]),
    "cpic_b200_image_download": (_i, [_vp, _vp, _i64]),
    "cpic_b200_image_upload": (_i, [_vp, _vp, _i64]),
    "cpic_b200_step_host": (_i, [_vp, _vp, _i64]),
    "cpic_b200_banded_image_bytes": (_i64, [_vp, _i]),
    "cpic_b200_banded_image_download": (_i, [_vp, _vp, _i64, _i]),
    "cpic_b200_step_host_banded": (_i, [_vp, _vp, _i64]),
    "cpic_b200_host_alloc": (_vp, [C.c_size_t]),
    "cpic_b200_host_free": (None, [_vp]),
    "cpic_b200_conf_load": (_i, [C.c_char_p, _pp]),
    "cpic_b200_conf_free": (None, [_vp]),
    "cpic_b200_conf_params": (_i, [_vp, _i, _i, _i, C.POINTER(ParamsC), C.POINTER(RunC)]),
    "cpic_b200_conf_init_particles": (_i, [_vp, _i] + [_pp] * 5),
    "cpic_b200_sim_from_conf": (_i, [C.c_char_p, _i, _i, _i, _i, _pp, C.POINTER(RunC)]),
    "cpic_b200_sim_from_conf_streamed": (_i, [C.c_char_p, _i, _i, _i, _i, _i64, _pp, C.POINTER(RunC)]),
    "cpic_b200_sim_from_conf_device": (_i, [C.c_char_p, _i, _i, _i, _i, _i64, _pp, C.POINTER(RunC)]),
    "cpic_b200_init_reference": (_i, [_vp, _i, _vp, _i64]),
    "cpic_b200_conf_stream_particles": (_i, [_vp, _i, _i64, _vp, _vp]),
    "cpic_b200_write_fields": (_i, [_vp, C.c_char_p, _i64, _i64, _i64, _i64, _d, _d]),
    "cpic_b200_write_fields_async": (_i, [_vp, C.c_char_p, _i64, _i64, _i64, _i64, _i64, _d, _d]),
    "cpic_b200_output_wait": (_i, [_vp]),
    "cpic_b200_get_fields_begin": (_i, [_vp, _vp]),
    "cpic_b200_get_fields_end": (_i, [_vp]),
    "cpic_b200_main": (_i, [_i, C.POINTER(C.c_char_p)]),
}

_lib = None


def lib():
    """The loaded library. There is no fallback: a missing extension is an error."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: build it with cpic_b200.build() "
                              "(nvcc -gencode arch=compute_100a,code=sm_100a); cpic_b200 has no CPU path")
        L = C.CDLL(path, mode=os.RTLD_GLOBAL)
        for name, (res, args) in EXPORTS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        version = L.cpic_b200_version().decode()
        if "sm_100a" not in version and os.environ.get("CPIC_B200_SIMT_CHECK") != "1":
            # CPIC_B200_LIB may name another build of the CUDA library, never a CPU one: the build
            # that the test suite interprets on the CPU is only accepted inside that suite
            raise ImportError(f"{path} is not a CUDA build of cpic_b200 ({version}); cpic_b200 has no CPU path")
        _lib = L
    return _lib


def check(rc):
    if rc:
        raise Cpic_b200Error(rc, lib().cpic_b200_last_error().decode(errors="replace"))
