"""dcarl_b200 — B200 (sm_100a) implementation of DCARL's data-parallel hot path.

Same Python entry points as the reference (zhcao92/DCARL), executed by hand-written CUDA kernels
behind a C ABI (include/dcarl_b200.h, dcarl_b200/csrc/*.cu, bound in _abi.py with ctypes):

    upper_bound, lower_bound, CI_lower_bound, mean_value      test_DCARL.py:10-28
    stream_bounds / simulate / StreamPlan                     test_DCARL.py main loops
    Data_Generation, add_an_act_data, random_state_norm       data_sampling.py
    RLS, BoxBuffer                                            stable_baselines/deepq/RLS.py
    rollout_cv                                                constant-velocity surrogate
    rollout_candidates, collect_and_select                    dqn_value_collect.py:86-145 + TestScenario_Town03.py:403-422
    JunctionTrajectoryPlanner, frenet_plan, Routes            Agent/zzz/JunctionTrajectoryPlanner.py (+ predict.py)

There is no CPU fallback: a compute call without the built library or without a CUDA device raises.
"""
from ._abi import DcarlError, Params, RlsParams, build, LIB_PATH  # noqa: F401
from .bounds import CI_lower_bound, bounds_batch, cell_values, lower_bound, mean_value, upper_bound  # noqa: F401
from .stream import StreamPlan, StreamResult, merge_stats, simulate, stream_bounds  # noqa: F401
from .sampling import (Data_Generation, add_an_act_data, generate, random_state_manual,  # noqa: F401
                       random_state_norm, sample_returns, sample_returns_philox)
from .rls import RLS, BoxBuffer, VISITED_STATE_DIST  # noqa: F401
from .rollout import collect_and_select, rollout_candidates, rollout_cv, rollout_params  # noqa: F401
from .frenet import (FrenetParams, FrenetResult, JunctionTrajectoryPlanner, Routes, TrajectoryAction,  # noqa: F401
                     dense_polyline2d, frenet_params, frenet_plan, frenet_start)
from .io import load_rls_buffer, load_sim_data, save_rls_buffer  # noqa: F401
from . import dist  # noqa: F401

__version__ = "0.1.0"
