"""ctypes binding of libtracefly_b200.so — the C-ABI declared in include/tracefly_b200.h.

There is NO fallback: if the shared library is missing this module raises at import, and
every entry point raises if no sm_100 device is present (TF_ERR_NODEVICE).  ctypes releases
the GIL around each foreign call, so the poller threads of UAV_Navigation/main.py (MAIN:82-86)
keep running while kernels are enqueued / synchronised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TRACEFLY_B200_LIB: load another build of the SAME library (A/B measurements of kernel variants, build.py)
LIB_PATH = os.environ.get("TRACEFLY_B200_LIB") or os.path.join(_HERE, "_build", "libtracefly_b200.so")

TF_OK = 0
TF_ERR_INVALID, TF_ERR_SHAPE, TF_ERR_UNSUPPORTED, TF_ERR_CUDA = -1, -2, -3, -4
TF_ERR_NOMEM, TF_ERR_NODEVICE, TF_ERR_WORKSPACE = -5, -6, -7
SEM_NONE, SEM_HR, SEM_PF = 0, 1, 2
KIND_SPHERE, KIND_CUBE, KIND_CYLINDER, KIND_AABB = 0, 1, 2, 3
SEL_TRUNCATION, SEL_TOURNAMENT, SEL_TOP2 = 0, 1, 2
CX_SINGLE_POINT, CX_UNIFORM = 0, 1
MUT_LATTICE, MUT_LATTICE_PF, MUT_GAUSSIAN = 0, 1, 2
W_MAX, TOPK_MAX = 256, 8192


class Params(C.Structure):  # tf_params
    _fields_ = [("w_safety", C.c_float), ("w_energy", C.c_float), ("w_wind", C.c_float), ("w_time", C.c_float),
                ("w_turn", C.c_float), ("dist_zero_guard", C.c_int32), ("hazard_weight", C.c_float),
                ("samples_per_segment", C.c_int32)]


class GaParams(C.Structure):  # tf_ga_params
    _fields_ = [("n_elite", C.c_int32), ("selection", C.c_int32), ("n_top", C.c_int32), ("tournament_k", C.c_int32),
                ("crossover", C.c_int32), ("mutation", C.c_int32), ("mutation_rate", C.c_float), ("sigma", C.c_float),
                ("stream_id", C.c_uint32), ("reserve_sms", C.c_int32)]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a).  tracefly_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_vp, _i, _i64, _u64, _u32, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_uint32, C.c_float, C.c_double
_SIGS = {
    "tf_env_create": (_i, [_i, _vp, _i, _f, C.POINTER(_vp)]),
    "tf_env_create_f64": (_i, [_i, _vp, _i, _d, C.POINTER(_vp)]),
    "tf_env_set_dynamic": (_i, [_vp, _vp, _i, _f]),
    "tf_env_build_static_index": (_i, [_vp, C.POINTER(_f), C.POINTER(_f), _f, _i, _f, _i, _vp]),
    "tf_env_static_index_stats": (_i, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_d), C.POINTER(_i), C.POINTER(_d)]),
    "tf_env_build_dynamic_index": (_i, [_vp, _i, _i, C.POINTER(_f), C.POINTER(_f), _f, _i, _vp]),
    "tf_env_dynamic_index_stats": (_i, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_d), C.POINTER(_i)]),
    "tf_env_set_voxels": (_i, [_vp, _vp, _i, _i, _i]),
    "tf_env_set_hazard": (_i, [_vp, _vp, _i, _i, _i, _f, _f, _f, _f, _f]),
    "tf_env_destroy": (_i, [_vp]),
    "tf_voxel_grid_bytes": (_i64, [_i, _i, _i]),
    "tf_voxelize": (_i, [_i, _vp, _i, _f, _vp, _i, _i, _i, _vp]),
    "tf_voxelize_points": (_i, [_vp, _i64, C.POINTER(_d), _d, _vp, _i, _i, _i, _vp]),
    "tf_find_path_gae": (_i, [_vp, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "tf_py_shuffle6_tape": (_i, [_vp, _vp, _i, _vp]),
    "tf_encode_waypoints": (_i, [_vp, _i, _i64, _i, _i, _d, _vp, _vp]),
    "tf_roulette_cumweights": (_i, [_vp, _i, _vp, _vp]),
    "tf_select_roulette": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "tf_voxel_pack_linear": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "tf_params_default": (None, [C.POINTER(Params)]),
    "tf_ga_params_default": (None, [C.POINTER(GaParams)]),
    "tf_fitness": (_i, [_vp, _vp, _vp, _i, _i, C.POINTER(Params), _vp, _vp, _vp]),
    "tf_fitness_f64": (_i, [_vp, _vp, _vp, _i, _i, C.POINTER(Params), _vp, _vp, _vp]),
    "tf_point_query_f64": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "tf_topk_workspace_bytes": (_i64, [_i, _i, _i]),
    "tf_rank_topk": (_i, [_vp, _i, _i, _i, _vp, _vp, _i64, _vp]),
    "tf_rank_topk_f64": (_i, [_vp, _i, _i, _i, _vp, _vp, _i64, _vp]),
    "tf_select_set_workspace_bytes": (_i64, [_i]),
    "tf_select_set": (_i, [_vp, _i, _i, _i, _vp, _vp, _i64, _vp]),
    "tf_ga_step": (_i, [_vp, _vp, _vp, _vp, _vp, _u64, _u32, C.POINTER(GaParams), _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "tf_ga_step_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _u64, _u32, C.POINTER(GaParams), _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "tf_np_hr_tape": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _vp]),
    "tf_generation": (_i, [_vp, _vp, _vp, _vp, _vp, _u64, _u32, C.POINTER(GaParams), C.POINTER(Params), _i, _i,
                           _vp, _vp, _vp, _vp, _vp, _vp]),
    "tf_generation_ranked": (_i, [_vp, _vp, _vp, _vp, _vp, _u64, _u32, C.POINTER(GaParams), C.POINTER(Params), _i, _i,
                                  _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i64, _vp]),
    "tf_track_best": (_i, [_vp, _vp, _i, _i, _u32, _vp, _vp, _vp, _vp]),
    "tf_track_best_f64": (_i, [_vp, _vp, _i, _i, _u32, _vp, _vp, _vp, _vp]),
    "tf_init_diagonal": (_i, [C.POINTER(_d), C.POINTER(_d), _i, _i, _vp, C.POINTER(_i), _vp]),
    "tf_init_diagonal_f64": (_i, [C.POINTER(_d), C.POINTER(_d), _i, _i, _vp, C.POINTER(_i), _vp]),
    "tf_init_walks": (_i, [_u64, _vp, C.POINTER(_f), C.POINTER(_f), C.POINTER(_f), _i, _i, _vp, _vp]),
    "tf_pack_elites": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "tf_replace_worst": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp]),
    "tf_fitness_gae": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "tf_unpack_i16": (_i, [_vp, _vp, _i64, _vp]),
    "tf_pack_i16": (_i, [_vp, _vp, _i64, _vp]),
    "tf_version": (_i, []),
    "tf_strerror": (C.c_char_p, [_i]),
    "tf_last_cuda_error": (_i, []),
    "tf_device_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i64)]),
    "tf_philox4x32_10": (None, [_u32, _u32, _u32, _u32, _u32, _u32, C.POINTER(_u32)]),
    "tf_unpack_xyz": (_i, [_vp, _vp, _i64, _vp]),
    "tf_pack_xyz": (_i, [_vp, _vp, _i64, _vp]),
    "tf_find_path_pf": (_i, [_vp, C.POINTER(_d), C.POINTER(_d), _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "tf_selftest_sqrt": (_i, [_u32, _u64, _vp, _vp]),
    "tf_selftest_rcp": (_i, [_u32, _u64, _vp, _vp]),
    "tf_set_option": (_i, [_i, _i]),
    "tf_launch_count": (_i64, []),
}
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)  # AttributeError here = header / library mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args

EXPORTED = tuple(_SIGS)


class TraceFlyError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        msg = lib.tf_strerror(status).decode()
        if status == TF_ERR_CUDA:
            msg += f" [cudaError {lib.tf_last_cuda_error()}]"
        super().__init__(f"{where}: {msg} ({status})")


def check(status, where):
    if status == TF_OK:
        return
    if status in (TF_ERR_INVALID, TF_ERR_SHAPE):
        raise ValueError(str(TraceFlyError(status, where)))
    raise TraceFlyError(status, where)


def default_params(**kw):
    p = Params()
    lib.tf_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def default_ga_params(**kw):
    g = GaParams()
    lib.tf_ga_params_default(C.byref(g))
    for k, v in kw.items():
        setattr(g, k, v)
    return g


def philox(key0, key1, c0, c1, c2, c3):
    out = (C.c_uint32 * 4)()
    lib.tf_philox4x32_10(key0, key1, c0, c1, c2, c3, out)
    return [int(v) for v in out]


CODEC_QUANT_F32, CODEC_DELTA_ROUND = 0, 1
DODGE = {"up": 0, "down": 1, "left": 2, "right": 3, "forward": 4, "backward": 5}
OPT_DISABLE_STREAM = 1
OPT_STREAM_SCHEDULE = 2
OPT_RESERVE_SMS = 3


def set_option(option, value):
    check(lib.tf_set_option(option, int(value)), "tf_set_option")


def launch_count():
    return int(lib.tf_launch_count())


if os.environ.get("TRACEFLY_STREAM_SCHEDULE"):      # A/B knob for measurements (see TF_OPT_STREAM_SCHEDULE)
    set_option(OPT_STREAM_SCHEDULE, int(os.environ["TRACEFLY_STREAM_SCHEDULE"]))
