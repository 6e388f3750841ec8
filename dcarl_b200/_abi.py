"""ctypes binding of libdcarl_b200.so (include/dcarl_b200.h).  No torch types cross the
boundary: tensors are passed as ``data_ptr()`` integers, the stream as a raw handle.

The library is the only compute path of this package.  If it is missing, or no CUDA
device is present, every compute call raises — there is no CPU fallback.
"""
import ctypes as C
import os
import subprocess

_DIR = os.path.dirname(os.path.abspath(__file__))
# DCARL_B200_LIB selects another build of the same library (kernel A/B runs, tools/gpu_ab.sh)
LIB_PATH = os.environ.get("DCARL_B200_LIB") or os.path.join(_DIR, "lib", "libdcarl_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_DIR), "include", "dcarl_b200.h")
_lib = None
ABI_VERSION = 2


class DcarlError(RuntimeError):
    """A C-ABI call returned non-zero (code < 0: bad argument index, > 0: cudaError_t)."""

    def __init__(self, code, msg):
        super().__init__(f"dcarl_b200: {msg} (code {code})")
        self.code = code


class Params(C.Structure):
    """dcarl_params — constants of test_DCARL.py:10-12, 43-52."""
    _fields_ = [("alpha", C.c_double), ("scale", C.c_double), ("cap", C.c_double),
                ("prior_rule", C.c_double), ("prior_other", C.c_double),
                ("activation_bonus", C.c_double), ("n_thres", C.c_int32), ("rule_act", C.c_int32)]

    def __init__(self, alpha=0.05, scale=150.0, cap=100.0, prior_rule=100.0, prior_other=-50.0,
                 activation_bonus=0.9, n_thres=10, rule_act=0):
        super().__init__(alpha, scale, cap, prior_rule, prior_other, activation_bonus, n_thres, rule_act)


class RlsParams(C.Structure):
    """dcarl_rls_params — constants of RLS.act_test (RLS.py:120-157)."""
    _fields_ = [("rule_mean_gate", C.c_double), ("confidence_thres", C.c_double),
                ("visited_times_thres", C.c_int32), ("min_candidate_visits", C.c_int32)]

    def __init__(self, rule_mean_gate=-0.1, confidence_thres=0.5, visited_times_thres=10,
                 min_candidate_visits=5):
        super().__init__(rule_mean_gate, confidence_thres, visited_times_thres, min_candidate_visits)


MAX_PEERS = 16


class PeerCtx(C.Structure):
    """dcarl_peer_ctx — peer-mapped all-reduce workspaces of the ranks of one NVLink box."""
    _fields_ = [("slots", C.c_void_p * MAX_PEERS), ("flags", C.c_void_p * MAX_PEERS), ("epoch", C.c_void_p),
                ("rank", C.c_int32), ("world", C.c_int32), ("n_max", C.c_int32), ("pad_", C.c_int32)]


class RolloutParams(C.Structure):
    """dcarl_rollout_params — constants of the candidate-trajectory rollout (see the header)."""
    _fields_ = [("dt", C.c_double), ("robot_radius", C.c_double), ("move_gap", C.c_double),
                ("ticks_per_point", C.c_double), ("collision_reward", C.c_double), ("pass_y", C.c_double),
                ("stuck_speed", C.c_double), ("stay_thres", C.c_double), ("n_pred", C.c_int32),
                ("obstacles_considered", C.c_int32)]


FRENET_MAX_D, FRENET_MAX_T, FRENET_MAX_V = 16, 4, 8


class FrenetParams(C.Structure):
    """dcarl_frenet_params — the sample grids and limits of JunctionTrajectoryPlanner.py:13-39."""
    _fields_ = [("d_samples", C.c_double * FRENET_MAX_D), ("T_samples", C.c_double * FRENET_MAX_T),
                ("tv_samples", C.c_double * FRENET_MAX_V), ("n_t", C.c_int32 * FRENET_MAX_T),
                ("n_d", C.c_int32), ("n_T", C.c_int32), ("n_v", C.c_int32), ("n_pred", C.c_int32),
                ("obstacles_considered", C.c_int32), ("pad_", C.c_int32),
                ("dt", C.c_double), ("target_speed", C.c_double),
                ("max_speed", C.c_double), ("max_accel", C.c_double), ("max_curvature", C.c_double),
                ("robot_radius", C.c_double), ("move_gap", C.c_double),
                ("kj", C.c_double), ("kt", C.c_double), ("kd", C.c_double), ("klat", C.c_double), ("klon", C.c_double)]


_vp, _i32, _i64, _sz, _d, _f, _u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t, C.c_double, C.c_float, C.c_uint64
_int = C.c_int

# name -> (restype, argtypes); must list every symbol include/dcarl_b200.h declares
SIGNATURES = {
    "dcarl_last_error": (C.c_char_p, []),
    "dcarl_abi_version": (_int, []),
    "dcarl_device_info": (_int, [C.POINTER(_int)] * 3),
    "dcarl_default_params": (None, [C.POINTER(Params)]),
    "dcarl_default_rls_params": (None, [C.POINTER(RlsParams)]),
    "dcarl_bounds": (_int, [_vp, _int, _vp, _i64, C.POINTER(Params), _vp, _vp]),
    "dcarl_stream_bounds_ws": (_sz, [_i64, _i64, _i32, _i32]),
    "dcarl_stream_rows_fit": (_int, [_i32, _i32]),
    "dcarl_stream_bounds": (_int, [_vp, _vp, _vp, _int, _i64, _i64, _i32, _i32, C.POINTER(Params),
                                   _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _int, _vp]),
    "dcarl_stream_bounds_u8": (_int, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, C.POINTER(Params),
                                      _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dcarl_cell_values": (_int, [_vp, _vp, _vp, _vp, _i64, C.POINTER(Params), _vp, _vp]),
    "dcarl_stream_stats": (_int, [_vp, _int, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp]),
    "dcarl_stream_stats_reduce": (_int, [_vp, _int, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _i32, _int,
                                         C.POINTER(PeerCtx), _vp]),
    "dcarl_peer_ws_bytes": (_sz, [_i32, _i32]),
    "dcarl_peer_ws_layout": (None, [_i32, _i32, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)]),
    "dcarl_peer_allreduce": (_int, [_vp, _i32, C.POINTER(PeerCtx), _vp]),
    "dcarl_allreduce_stats": (_int, [_vp, _i32, _vp, _vp]),
    "dcarl_sample_returns": (_int, [_vp, _int, _vp, _vp, _int, _vp, _i64, _i32, _i32, _d, _d, _d, _d,
                                    _vp, _vp, _int, _vp]),
    "dcarl_sample_returns_philox": (_int, [_vp, _int, _u64, _i64, _i32, _i32, _d, _d, _d, _d,
                                           _vp, _vp, _vp, _int, _vp]),
    "dcarl_box_moments_ws": (_sz, [_i32, _i32, _i32]),
    "dcarl_box_moments": (_int, [_vp, _vp, _int, _i64, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _sz, _vp]),
    "dcarl_box_moments_ranges": (_int, [_vp, _vp, _int, _i64, _i32, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp]),
    "dcarl_box_query_ws": (_sz, [_i32, _i32, _i32]),
    "dcarl_box_query": (_int, [_vp, _vp, _int, _i64, _i32, _vp, _vp, _i32, _i32, C.POINTER(RlsParams), _vp, _vp, _vp,
                               _vp, _sz, C.POINTER(PeerCtx), _vp]),
    "dcarl_box_select": (_int, [_vp, _i32, _i32, C.POINTER(RlsParams), _vp, _vp, _vp]),
    "dcarl_add_data_counts": (_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "dcarl_add_data_rows": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _d, _vp, _vp, _vp]),
    "dcarl_rollout_cv": (_int, [_vp, _vp, _i64, _i32, _f, _f, _f, _vp, _i32, _vp, _vp, _vp, _vp]),
    "dcarl_default_rollout_params": (None, [C.POINTER(RolloutParams)]),
    "dcarl_rollout_candidates": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, C.POINTER(RolloutParams),
                                        _vp, _vp, _vp, _vp]),
    "dcarl_default_frenet_params": (None, [C.POINTER(FrenetParams)]),
    "dcarl_spline_build": (_int, [_vp, _vp, _i32, _vp, _vp]),
    "dcarl_frenet_start": (_int, [_vp, _vp, _i32, _vp, _vp, _i64, _vp, _vp]),
    "dcarl_frenet_plan": (_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i64, C.POINTER(FrenetParams),
                                 _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into lib/libdcarl_b200.so (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(_DIR, "csrc"), "-j8"], stdout=out)
    return LIB_PATH


def lib():
    """Load the shared library (once).  Raises loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"dcarl_b200: CUDA library not built ({LIB_PATH} missing). Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C dcarl_b200/csrc`. "
                "There is no CPU fallback.")
        try:
            import torch  # noqa: F401  (loads the CUDA runtime the library shares)
        except Exception:
            pass
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.dcarl_abi_version() != ABI_VERSION:
            raise RuntimeError("dcarl_b200: ABI version mismatch between _abi.py and the built library")
        _lib = L
    return _lib


def check(code):
    if code != 0:
        raise DcarlError(code, lib().dcarl_last_error().decode())


def device_info():
    sm, mj, mn = _int(), _int(), _int()
    check(lib().dcarl_device_info(C.byref(sm), C.byref(mj), C.byref(mn)))
    return sm.value, mj.value, mn.value
