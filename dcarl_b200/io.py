"""Ingest of the reference's on-disk formats into device buffers (SURVEY "next" row f1).

* ``data.npy`` / ``data_carla.npy``: float64 [N, 4] rows ``[idx, state, act, R]`` as written by
  data_sampling.py:65 and read by test_DCARL.py:32-33.
* ``visited_state.txt`` / ``visited_value.txt``: the RLS buffer, ``%f`` text with 21 and 2 columns
  (RLS.py:56-59), reloaded at construction (RLS.py:50-53).
"""
import os

import numpy as np
import torch

from .rls import BoxBuffer, VISITED_STATE_DIST


def load_sim_data(path, n_rows=None, r_dtype=np.float64, device=None):
    """-> (idx int32, act int32, R) for ``stream_bounds``; CUDA tensors when ``device`` is given."""
    data = np.load(path)
    if data.ndim != 2 or data.shape[1] != 4:
        raise ValueError(f"{path}: expected a float [N, 4] array of [idx, state, act, R] rows")
    data = data[:n_rows] if n_rows is not None else data
    idx, act = data[:, 0].astype(np.int32), data[:, 2].astype(np.int32)
    if not (np.array_equal(idx, data[:, 0]) and np.array_equal(act, data[:, 2])):
        raise ValueError(f"{path}: idx / act columns are not integral")
    R = np.ascontiguousarray(data[:, 3].astype(r_dtype))
    if device is None:
        return idx, act, R
    return tuple(torch.from_numpy(x).to(device) for x in (idx, act, R))


def load_rls_buffer(directory=".", obs_dimension=20, dist=VISITED_STATE_DIST, build_index=True):
    """visited_state.txt + visited_value.txt -> BoxBuffer on the current device (column 1 of the
    value file is the return, RLS.py:174), optionally with the sorted index built."""
    states = np.loadtxt(os.path.join(directory, "visited_state.txt")).reshape(-1, obs_dimension + 1)
    values = np.loadtxt(os.path.join(directory, "visited_value.txt")).reshape(-1, 2)
    if len(states) != len(values):
        raise ValueError("visited_state.txt and visited_value.txt disagree on the row count")
    buf = BoxBuffer(obs_dimension, dist, capacity=max(len(states), 16))
    if len(states):
        buf.append(states, values[:, 1])
        if build_index:
            buf.build_index()
    return buf


def save_rls_buffer(buf, directory=".", actions=None):
    """Write the device buffer back in the reference's text format (checkpoint of the agent's
    experience).  ``actions`` defaults to the stored action column (the first column of
    visited_value.txt is the action, RLS.py:194)."""
    st = buf.state.cpu().numpy().astype(np.float64)
    rt = buf.ret.cpu().numpy().astype(np.float64)
    a = st[:, buf.D] if actions is None else np.asarray(actions, np.float64)
    np.savetxt(os.path.join(directory, "visited_state.txt"), st, fmt="%f")
    np.savetxt(os.path.join(directory, "visited_value.txt"), np.stack([a, rt], 1), fmt="%f")
