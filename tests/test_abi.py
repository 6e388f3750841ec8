"""CPU-side checks of the drop-in boundary: the shared library loads, exports exactly what
include/dcarl_b200.h declares, and validates arguments before touching the device."""
import ctypes as C
import os
import re
import subprocess

import pytest

from dcarl_b200 import _abi


def _header_functions():
    src = open(_abi.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dcarl_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_in_tree():
    assert os.path.exists(_abi.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.realpath(_abi.LIB_PATH).startswith(os.path.realpath(os.path.dirname(os.path.dirname(_abi.__file__))))


def test_header_bindings_and_exports_agree():
    declared = _header_functions()
    assert declared and sorted(_abi.SIGNATURES) == declared
    out = subprocess.check_output(["nm", "-D", "--defined-only", _abi.LIB_PATH], text=True)
    exported = sorted(set(re.findall(r"\sT\s+(dcarl_[a-z0-9_]+)$", out, flags=re.M)))
    assert exported == declared


def test_library_targets_sm_100a_only():
    out = subprocess.check_output(["cuobjdump", "-lelf", _abi.LIB_PATH], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_loads_and_reports_version_and_defaults():
    L = _abi.lib()
    assert L.dcarl_abi_version() == _abi.ABI_VERSION == 2
    p = _abi.Params(alpha=0, scale=0, cap=0, prior_rule=0, prior_other=0, activation_bonus=0, n_thres=0, rule_act=5)
    L.dcarl_default_params(C.byref(p))
    assert (p.alpha, p.scale, p.cap, p.prior_rule, p.prior_other, p.activation_bonus, p.n_thres, p.rule_act) == \
           (0.05, 150.0, 100.0, 100.0, -50.0, 0.9, 10, 0)
    r = _abi.RlsParams(0, 0, 0, 0)
    L.dcarl_default_rls_params(C.byref(r))
    assert (r.rule_mean_gate, r.confidence_thres, r.visited_times_thres, r.min_candidate_visits) == (-0.1, 0.5, 10, 5)


def test_argument_validation_needs_no_device():
    L = _abi.lib()
    p = _abi.Params()
    # S <= 0 is argument 7 of dcarl_stream_bounds
    rc = L.dcarl_stream_bounds(None, None, None, 0, 1, 1, 0, 11, p, None, 0, None, None, None, None, None, None, 0, 0, None)
    assert rc == -7 and b"S <= 0" in L.dcarl_last_error()
    rc = L.dcarl_stream_bounds(None, None, None, 0, 1, 1, 20, 11, _abi.Params(rule_act=11), None, 0, None, None, None,
                               None, None, None, 0, 0, None)
    assert rc == -9 and b"rule_act" in L.dcarl_last_error()
    rc = L.dcarl_box_moments(None, None, 0, -1, 20, None, None, 1, 8, None, None, 0, None)
    assert rc == -4
    rc = L.dcarl_box_query(None, None, 1, 5, 0, None, None, 1, 8, _abi.RlsParams(), None, None, None, None, 0, None, None)
    assert rc == -5 and b"D must be" in L.dcarl_last_error()
    bad = _abi.PeerCtx()
    bad.world, bad.rank, bad.n_max = 99, 0, 8
    assert L.dcarl_peer_allreduce(C.c_void_p(16), 4, C.byref(bad), None) == -3 and b"DCARL_MAX_PEERS" in L.dcarl_last_error()
    assert L.dcarl_peer_ws_bytes(8, 768) >= 2 * 8 * 768 * 8 + 2 * 8 * 4 + 4
    rp = _abi.RolloutParams()
    L.dcarl_default_rollout_params(C.byref(rp))
    assert (rp.dt, rp.robot_radius, rp.ticks_per_point, rp.collision_reward, rp.pass_y, rp.n_pred) == (0.3, 1.0, 6.0, -100.0, 73.7, 15)
    assert L.dcarl_rollout_candidates(None, None, None, None, None, None, 0, 5, 40, 14, rp, None, None, None, None) == -9
    assert L.dcarl_stream_bounds_u8(None, None, None, 1, 1, 300, 4, p, None, None, None, None, None, None, None) == -13
    rc = L.dcarl_bounds(None, 1, None, 3, _abi.Params(alpha=0.0), None, None)
    assert rc == -5 and b"alpha" in L.dcarl_last_error()
    with pytest.raises(_abi.DcarlError):
        _abi.check(rc)
    # empty problems are fine without a device
    assert L.dcarl_stream_bounds(None, None, None, 0, 0, 10, 20, 11, p, None, 0, None, None, None, None, None, None, 0, 0, None) == 0
    assert L.dcarl_box_select(None, 0, 8, _abi.RlsParams(), None, None, None) == 0
    assert L.dcarl_stream_bounds_ws(4096, 20000, 20, 11) >= 4096 * 220 * 28
    assert L.dcarl_stream_rows_fit(20, 11) == 1 and L.dcarl_stream_rows_fit(40, 256) == 0
    # the planner entry points validate their arguments on the host, before any launch
    fp = _abi.FrenetParams()
    L.dcarl_default_frenet_params(fp)
    assert (fp.n_d, fp.n_T, fp.n_v, fp.n_t[0], fp.n_pred) == (5, 1, 2, 14, 15) and abs(fp.target_speed - 30 / 3.6) < 1e-15
    assert L.dcarl_frenet_plan(None, None, 0, None, None, None, None, None, 0, 0, fp, None, None, None, None, None, None, None, None) == 0
    fp.n_d = 17
    assert L.dcarl_frenet_plan(None, None, 0, None, None, None, None, None, 0, 5, fp, None, None, None, None, None, None, None, None) == -11
    assert b"n_d" in L.dcarl_last_error()
    fp.n_d, fp.n_v = 16, 8
    assert L.dcarl_frenet_plan(None, None, 0, None, None, None, None, None, 0, 5, fp, None, None, None, None, None, None, None, None) == -11
    assert b"32 candidates" in L.dcarl_last_error()
    assert L.dcarl_spline_build(None, None, 0, None, None) == 0 and L.dcarl_frenet_start(None, None, 0, None, None, 0, None, None) == 0


def test_product_has_no_cpu_fallback_and_never_imports_the_oracle():
    import dcarl_b200
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            dcarl_b200.upper_bound([1.0, 2.0, 3.0])
    pkg = os.path.dirname(dcarl_b200.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "oracle" not in text.lower().replace("no cpu fallback", ""), f
