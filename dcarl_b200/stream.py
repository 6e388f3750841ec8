"""Streaming confidence-bound estimator + argmax selector on the GPU.

Drop-in for the ``__main__`` loops of Simulation_testing/Simulation_1/test_DCARL.py:73-99 and
Simulation_testing/Simulation_2/test_DCARL.py:72-105, batched over independent scenario
instances.  Array in, array out: numpy in -> numpy out (host<->device copies inside),
CUDA tensors in -> CUDA tensors out (no copies, asynchronous on the current stream).
"""
from dataclasses import dataclass
from typing import Optional

import ctypes as C

import numpy as np
import torch

from . import _abi, _dev

STATS_HEADER = 5   # DCARL_STATS_HEADER


@dataclass
class StreamResult:
    value: object            # [n_inst, n_steps] max_a value[idx][a] per row        (step_TSRL_value)
    act: object              # [n_inst, n_steps] int32 argmax per row              (step_TSRL_act)
    activation_step: object  # [n_inst, S] int64, -1 = never                        (activation_step)
    overall: Optional[object]  # [n_inst, n_steps] or None                          (overall_value)
    act_hist: Optional[object] = None   # [A] uint64 histogram of act (int64 tensor)
    stats: Optional[object] = None      # [5 + A] float64 summary vector (see dcarl_stream_stats)


class StreamPlan:
    """Pre-allocated device (and pinned host) buffers for repeated batches of one shape."""

    def __init__(self, n_inst, n_steps, S, A, params=None, *, r_dtype=torch.float32,
                 out_dtype=torch.float32, want_overall=True, want_stats=True, variant=0, device=None,
                 key_dtype=torch.int32):
        """key_dtype=torch.uint8 selects the compact wire format (dcarl_stream_bounds_u8): idx / act in
        and the argmax out are one byte each — 6 + 9 bytes per scenario-step instead of 12 + 12 on the
        host <-> device link.  It needs S, A <= 256, float32 rows and outputs, and a table that fits
        shared memory; results are identical to the int32 format."""
        _dev.require_cuda()
        self.n_inst, self.n_steps, self.S, self.A = int(n_inst), int(n_steps), int(S), int(A)
        self.params = params or _abi.Params()
        self.r_dtype, self.out_dtype = r_dtype, out_dtype
        self.want_overall, self.want_stats, self.variant = want_overall, want_stats, int(variant)
        self.key_dtype = key_dtype
        self.want_stats_kernel = True      # dist.MixedStreamBatch chains the summary kernels of several plans itself
        if key_dtype not in (torch.int32, torch.uint8):
            raise ValueError("key_dtype must be torch.int32 or torch.uint8")
        if key_dtype == torch.uint8:
            if r_dtype != torch.float32 or out_dtype != torch.float32 or self.variant == 1:
                raise ValueError("the compact wire format is float32 in / float32 out on the row-parallel kernel")
            if self.S > 256 or self.A > 256 or not _abi.lib().dcarl_stream_rows_fit(self.S, self.A):
                raise ValueError("the compact wire format needs S, A <= 256 and an S*A table that fits shared memory")
        self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
        shp = (self.n_inst, self.n_steps)
        e = lambda s, d: torch.empty(s, dtype=d, device=self.device)
        self.value = e(shp, out_dtype)
        self.act = e(shp, key_dtype)
        self.overall = e(shp, out_dtype) if want_overall else None
        self.activation_step = e((self.n_inst, self.S), torch.int64)
        self.act_hist = torch.zeros(self.A, dtype=torch.int64, device=self.device) if want_stats else None
        self.bad_rows = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.stats = e((STATS_HEADER + self.A,), torch.float64) if want_stats else None
        L = _abi.lib()
        walker = self.variant == 1 or (self.variant == 0 and not L.dcarl_stream_rows_fit(self.S, self.A))
        need = L.dcarl_stream_bounds_ws(self.n_inst, self.n_steps, self.S, self.A) if walker else 0
        self.ws = e((max(need, 16),), torch.uint8)
        self._ws_bytes = need
        self._pinned = None
        self._streams = None

    # -- device path -------------------------------------------------------------------------
    def _launch(self, pi, pa, pr, n_inst, value, act, astep, overall, ws, stream):
        """One dcarl_stream_bounds[_u8] call on raw pointers (c_void_p) for `n_inst` instances."""
        L = _abi.lib()
        sh = _dev.stream_handle(stream)
        if self.key_dtype == torch.uint8:
            _abi.check(L.dcarl_stream_bounds_u8(
                pi, pa, pr, n_inst, self.n_steps, self.S, self.A, self.params, value, act, astep, overall,
                _dev.ptr(self.act_hist), _dev.ptr(self.bad_rows), sh))
        else:
            _abi.check(L.dcarl_stream_bounds(
                pi, pa, pr, int(self.r_dtype == torch.float64), n_inst, self.n_steps, self.S, self.A, self.params,
                value, int(self.out_dtype == torch.float64), act, astep, overall,
                _dev.ptr(self.act_hist), _dev.ptr(self.bad_rows), ws, self._ws_bytes, self.variant, sh))

    def _stats(self, stream):
        _abi.check(_abi.lib().dcarl_stream_stats(
            _dev.ptr(self.value), int(self.out_dtype == torch.float64), _dev.ptr(self.overall),
            _dev.ptr(self.activation_step), _dev.ptr(self.act_hist), self.n_inst, self.n_steps,
            self.S, self.A, _dev.ptr(self.stats), _dev.stream_handle(stream)))

    def run_device(self, idx, act, R, stream=None, check=False):
        """idx/act key_dtype, R r_dtype: CUDA tensors [n_inst, n_steps] on the plan's device.
        Asynchronous on `stream` (default: the current stream of the plan's device)."""
        for t, d in ((idx, self.key_dtype), (act, self.key_dtype), (R, self.r_dtype)):
            if not (t.is_cuda and t.device == self.device and t.dtype == d and t.is_contiguous()
                    and tuple(t.shape) == (self.n_inst, self.n_steps)):
                raise ValueError(f"run_device: expected contiguous CUDA tensors [n_inst, n_steps] on {self.device} "
                                 f"of dtype {self.key_dtype} (idx, act) and {self.r_dtype} (R)")
        # the C ABI launches on the CURRENT device: make the plan's device current, and issue the
        # zeroing of the accumulators on the same stream as the kernels that add into them
        with torch.cuda.device(self.device):
            st = stream if stream is not None else torch.cuda.current_stream(self.device)
            with torch.cuda.stream(st):
                if self.want_stats:
                    self.act_hist.zero_()
                self.bad_rows.zero_()
                self._launch(_dev.ptr(idx), _dev.ptr(act), _dev.ptr(R), self.n_inst, _dev.ptr(self.value),
                             _dev.ptr(self.act), _dev.ptr(self.activation_step), _dev.ptr(self.overall),
                             _dev.ptr(self.ws), st)
                if self.want_stats and self.want_stats_kernel:
                    self._stats(st)
        if check:
            self.raise_on_bad_rows()
        return StreamResult(self.value, self.act, self.activation_step, self.overall, self.act_hist, self.stats)

    def raise_on_bad_rows(self):
        n = int(self.bad_rows.item())
        if n:
            raise IndexError(f"{n} rows have idx outside [0, {self.S}) or act outside [0, {self.A}) "
                             "(the reference raises IndexError on such a row)")

    # -- host path: pinned staging, copies inside --------------------------------------------
    def _pin(self):
        if self._pinned is None:
            shp = (self.n_inst, self.n_steps)
            pe = lambda d: torch.empty(shp, dtype=d).pin_memory()
            self._pinned = dict(idx=pe(self.key_dtype), act=pe(self.key_dtype), R=pe(self.r_dtype),
                                value=pe(self.out_dtype), oact=pe(self.key_dtype),
                                overall=pe(self.out_dtype) if self.want_overall else None,
                                astep=torch.empty((self.n_inst, self.S), dtype=torch.int64).pin_memory())
            e = lambda d: torch.empty(shp, dtype=d, device=self.device)
            self._din = (e(self.key_dtype), e(self.key_dtype), e(self.r_dtype))
        return self._pinned

    def _key_bytes(self):
        return 1 if self.key_dtype == torch.uint8 else 4

    def h2d_bytes(self):
        return self.n_inst * self.n_steps * (2 * self._key_bytes() + (8 if self.r_dtype == torch.float64 else 4))

    def d2h_bytes(self):
        w = 8 if self.out_dtype == torch.float64 else 4
        return (self.n_inst * self.n_steps * (self._key_bytes() + w * (2 if self.want_overall else 1))
                + self.n_inst * self.S * 8)

    def _launch_range(self, lo, hi, stream):
        """Kernel launch for instances [lo, hi) of the plan's own input/output buffers."""
        ns = self.n_steps
        kb = self._key_bytes()
        rb = 8 if self.r_dtype == torch.float64 else 4
        ob = 8 if self.out_dtype == torch.float64 else 4
        di, da, dr = self._din
        off = lambda t, w: C.c_void_p(t.data_ptr() + lo * ns * w) if t is not None else None
        self._launch(off(di, kb), off(da, kb), off(dr, rb), hi - lo, off(self.value, ob), off(self.act, kb),
                     C.c_void_p(self.activation_step.data_ptr() + lo * self.S * 8), off(self.overall, ob),
                     C.c_void_p(self.ws.data_ptr() + (lo * (self._ws_bytes // max(self.n_inst, 1)))), stream)

    def run_pinned(self, chunks=8):
        """Inputs already in the plan's pinned host buffers (``plan.pinned()['idx'|'act'|'R']``).
        Instances are cut into `chunks` pieces that flow through three streams — host->device copy,
        kernel, device->host copy — so the two PCIe directions and the kernel overlap.  Returns the
        pinned buffers after everything has drained."""
        pin = self._pin()
        with torch.cuda.device(self.device):
            return self._run_pinned(pin, chunks)

    def _run_pinned(self, pin, chunks):
        if self._streams is None:
            self._streams = tuple(torch.cuda.Stream(device=self.device) for _ in range(3))
        s_in, s_k, s_out = self._streams
        cur = torch.cuda.current_stream(self.device)
        for st in self._streams:
            st.wait_stream(cur)
        chunks = max(1, min(int(chunks), self.n_inst))
        edges = [self.n_inst * c // chunks for c in range(chunks + 1)]
        with torch.cuda.stream(s_k):
            if self.want_stats:
                self.act_hist.zero_()
            self.bad_rows.zero_()
        outs = [(pin["value"], self.value), (pin["oact"], self.act), (pin["astep"], self.activation_step)]
        if self.want_overall:
            outs.append((pin["overall"], self.overall))
        for lo, hi in zip(edges[:-1], edges[1:]):
            if hi == lo:
                continue
            with torch.cuda.stream(s_in):
                for d, h in zip(self._din, (pin["idx"], pin["act"], pin["R"])):
                    d[lo:hi].copy_(h[lo:hi], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            s_k.wait_event(ev_in)
            self._launch_range(lo, hi, s_k)
            ev_k = torch.cuda.Event()
            ev_k.record(s_k)
            s_out.wait_event(ev_k)
            with torch.cuda.stream(s_out):
                for h, d in outs:
                    h[lo:hi].copy_(d[lo:hi], non_blocking=True)
        if self.want_stats:
            self._stats(s_k)
        for st in self._streams:
            cur.wait_stream(st)
        cur.synchronize()
        return pin

    def pinned(self):
        return self._pin()

    def run_host(self, idx, act, R):
        """numpy [n_inst, n_steps] in, numpy out."""
        pin = self._pin()
        pin["idx"].numpy()[...] = idx
        pin["act"].numpy()[...] = act
        pin["R"].numpy()[...] = R
        self.run_pinned()
        self.raise_on_bad_rows()
        return StreamResult(pin["value"].numpy().copy(), pin["oact"].numpy().copy(), pin["astep"].numpy().copy(),
                            pin["overall"].numpy().copy() if self.want_overall else None,
                            self.act_hist.cpu().numpy() if self.want_stats else None,
                            self.stats.cpu().numpy() if self.want_stats else None)


def stream_bounds(idx, act, R, S, A, params=None, *, want_overall=True, out_dtype=None, variant=0,
                  want_stats=False, key_dtype=torch.int32):
    """One-shot batched call.  idx/act/R: [n_steps] or [n_inst, n_steps]; numpy (host) or CUDA tensors.

    R float32 computes from float32 returns, anything else from float64; out_dtype defaults to
    float64 unless R is float32.  key_dtype=torch.uint8 runs the compact wire format (one-byte keys
    in, one-byte argmax out; float32 only).  Raises IndexError when a row's idx/act is out of range."""
    _dev.require_cuda()
    on_device = _dev.is_device(R)
    r_dt = _dev.float_dtype_of(R)
    o_dt = out_dtype or (torch.float32 if r_dt == torch.float32 else torch.float64)
    if isinstance(o_dt, np.dtype) or o_dt in (np.float32, np.float64):
        o_dt = torch.float32 if np.dtype(o_dt) == np.float32 else torch.float64
    if key_dtype == torch.uint8 and not on_device:
        for name, k, hi in (("idx", idx, S), ("act", act, A)):     # a one-byte cast must not hide a bad key
            k = np.asarray(k)
            if k.size and (k.min() < 0 or k.max() > 255):
                raise IndexError(f"{name} outside [0, {hi}) (the reference raises IndexError on such a row)")
    di, da, dr = _dev.as_device(idx, key_dtype), _dev.as_device(act, key_dtype), _dev.as_device(R, r_dt)
    single = di.dim() == 1
    if single:
        di, da, dr = di[None], da[None], dr[None]
    n_inst, n_steps = di.shape
    plan = StreamPlan(n_inst, n_steps, S, A, params, r_dtype=r_dt, out_dtype=o_dt, want_overall=want_overall,
                      want_stats=want_stats, variant=variant, device=di.device, key_dtype=key_dtype)
    res = plan.run_device(di.contiguous(), da.contiguous(), dr.contiguous(), check=True)
    sq = (lambda t: t[0]) if single else (lambda t: t)
    conv = (lambda t: t) if on_device else (lambda t: t.cpu().numpy())
    f = lambda t: None if t is None else conv(sq(t))
    return StreamResult(f(res.value), f(res.act), f(res.activation_step), f(res.overall),
                        None if res.act_hist is None else conv(res.act_hist),
                        None if res.stats is None else conv(res.stats))


def simulate(data, true_action_values, state_num, action_num, rule_act=0, n_thres=10, n_rows=20000):
    """The body of test_DCARL.py as a function (ST1:33-99 / ST2:32-105).

    data: float64 [N, 4] rows [idx, state, act, R] (``np.load('.../data.npy')``); the first n_rows
    are streamed like ``data[0:20000]``.  Returns the script's globals: per-state lists
    step_TSRL_value / step_TSRL_act / true_step_TSRL_value, activation_step, overall_value."""
    data = np.asarray(data, np.float64)[:n_rows]
    idx = data[:, 0].astype(np.int32)
    act = data[:, 2].astype(np.int32)
    res = stream_bounds(idx, act, data[:, 3], state_num, action_num,
                        _abi.Params(rule_act=rule_act, n_thres=n_thres), want_overall=True)
    tav = np.asarray(true_action_values)
    by_state = [np.nonzero(idx == s)[0] for s in range(state_num)]
    return dict(
        step_TSRL_value=[res.value[r].tolist() for r in by_state],
        step_TSRL_act=[res.act[r].astype(np.int64).tolist() for r in by_state],
        true_step_TSRL_value=[tav[s][res.act[r]].tolist() for s, r in enumerate(by_state)],
        activation_step=res.activation_step.copy(),
        overall_value=res.overall.tolist(),
        step_value_rows=res.value, step_act_rows=res.act)


def merge_stats(stats_list, A_max=None):
    """Sum the summary vectors of several groups of instances whose action counts differ (BASELINE
    config 5 mixes Simulation_1-shaped S=1/A=30 and Simulation_2-shaped S=20/A=11 scenarios): the
    action histogram tail is zero-padded to the widest group.  Tensors in -> tensor out."""
    A_max = A_max or max(int(st.numel()) - STATS_HEADER for st in stats_list)
    out = torch.zeros(STATS_HEADER + A_max, dtype=torch.float64, device=stats_list[0].device)
    for st in stats_list:
        out[:st.numel()] += st
    return out
