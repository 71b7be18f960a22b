"""Torch-tensor front end of the C ABI (include/mbd_b200.h).

PyTorch is plumbing here: it owns device memory and the CUDA stream; every compute call goes
through ctypes into libmbd_b200.so with raw device pointers.  There is no CPU/eager fallback:
all functions raise `MbdError` without a CUDA device.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import MbdError, check, key_ptr

RUN = 64  # samples per sequential run in mbd_weighted_sum (kRun in csrc/mbd_b200.cu)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(t: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    if not t.is_cuda:
        raise MbdError("expected a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype or not t.is_contiguous():
        raise MbdError(f"expected a contiguous {dtype} tensor, got {t.dtype} contiguous={t.is_contiguous()}")
    return t


def set_kernel_variant(v: int):
    """0 = auto, 1 = lane-per-link kernel, 2/3/4 = warp-per-link kernel with CTA / named / mbarrier sync,
    5 = two links per warp / 16 samples per CTA,
    6 = two interleaved sample groups per CTA (all bit-identical)."""
    check(_lib.lib().mbd_set_kernel_variant(int(v)), "mbd_set_kernel_variant")


class Model:
    """Device-resident compiled model (mbd_model_create)."""

    def __init__(self, blob: np.ndarray, device: Optional[torch.device] = None):
        _lib.require_gpu()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        blob = np.ascontiguousarray(blob, dtype=np.uint32)
        self.blob = blob
        hdr = blob.view(np.int32)
        self.L, self.nu, self.n_frames, self.reward, self.ntrack = (int(hdr[i]) for i in range(1, 6))
        with torch.cuda.device(self.device):
            self._h = _lib.lib().mbd_model_create(blob.ctypes.data_as(_lib.c_u32p), blob.size)
        if not self._h:
            raise MbdError("mbd_model_create failed: " + _lib.lib().mbd_last_error().decode())

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().mbd_model_destroy(ctypes.c_void_p(h))
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass

    @property
    def handle(self):
        return ctypes.c_void_p(self._h)


def sample(key, n_total: int, n_begin: int, n_local: int, HNu: int, sigma: float, Ybar: torch.Tensor,
           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _lib.require_gpu()
    Ybar = _dev(Ybar)
    out = torch.empty((n_local, HNu), device=Ybar.device, dtype=torch.float32) if out is None else _dev(out)
    k, kp = key_ptr(key)
    check(_lib.lib().mbd_sample(kp, n_total, n_begin, n_local, HNu, ctypes.c_float(sigma), _p(Ybar), _p(out), _stream()), "mbd_sample")
    return out


def rollout(model: Model, state_init: torch.Tensor, Y0s: torch.Tensor, xref: Optional[torch.Tensor] = None,
            want_rewss=False, want_final=False, want_track=False, nsub_override: int = 0):
    """vmap(rollout_us)(state_init, Y0s): Y0s [n,H,nu] -> dict(rews [n], rewss, logpd, final, track)."""
    state_init, Y0s = _dev(state_init), _dev(Y0s)
    n, H, nu = Y0s.shape
    if nu != model.nu or state_init.numel() != model.L * 13:
        raise MbdError("shape mismatch between model, state_init and Y0s")
    dev = Y0s.device
    rews = torch.empty(n, device=dev)
    rewss = torch.empty((n, H), device=dev) if want_rewss else None
    final = torch.empty((n, model.L, 13), device=dev) if want_final else None
    track = torch.empty((n, H, model.ntrack, 3), device=dev) if want_track else None
    logpd, href = None, 0
    if xref is not None:
        xref = _dev(xref)
        href = xref.shape[1]
        logpd = torch.empty(n, device=dev)
    check(_lib.lib().mbd_rollout(model.handle, _p(state_init), _p(Y0s), n, H, _p(rewss), _p(rews), _p(xref), href, _p(logpd),
                                 _p(final), _p(track), nsub_override, _stream()), "mbd_rollout")
    return dict(rews=rews, rewss=rewss, logpd=logpd, final=final, track=track)


def sample_rollout(model: Model, state_init, key, n_total, n_begin, n_local, H, sigma, Ybar, Y0s_out, rews_out,
                   xref=None, logpd_out=None):
    """The fused hot-path kernel: writes Y0s_out [n_local,H*nu], rews_out [n_local] (+logpd_out)."""
    k, kp = key_ptr(key)
    href = 0 if xref is None else xref.shape[1]
    check(_lib.lib().mbd_sample_rollout(model.handle, _p(_dev(state_init)), kp, n_total, n_begin, n_local, H, ctypes.c_float(sigma),
                                        _p(_dev(Ybar)), _p(_dev(Y0s_out)), _p(_dev(rews_out)), _p(xref), href, _p(logpd_out),
                                        _stream()), "mbd_sample_rollout")


EUNSUPPORTED = -4


def reverse_step(model: Model, state_init, key, n, H, sigma, Ybar_i, temp, coef, Y0s_out, rews_out, weights_out, scalars_out,
                 runs_scratch, out) -> bool:
    """reverse_once as ONE cooperative kernel.  Returns False when the configuration is not covered by the
    fused kernel (the caller then issues the separate launches)."""
    k, kp = key_ptr(key)
    c = (ctypes.c_float * 5)(*[float(v) for v in coef])
    rc = _lib.lib().mbd_reverse_step(model.handle, _p(_dev(state_init)), kp, n, H, ctypes.c_float(sigma), _p(_dev(Ybar_i)),
                                     ctypes.c_float(temp), c, _p(_dev(Y0s_out)), _p(_dev(rews_out)), _p(_dev(weights_out)),
                                     _p(_dev(scalars_out)), _p(_dev(runs_scratch)), _p(_dev(out)), _stream())
    if rc == EUNSUPPORTED:
        return False
    check(rc, "mbd_reverse_step")
    return True


def car2d_rollout(params, x0, Y0s, xref=None, want_rewss=False, want_traj=False, key=None, n_total=0, n_begin=0,
                  sigma=0.0, Ybar=None, rews_out=None, logpd_out=None):
    """Car2d rollouts; with `key` the noise is drawn in-kernel and written to Y0s [n,H,2]."""
    params, x0, Y0s = _dev(params), _dev(x0), _dev(Y0s)
    n, H, _ = Y0s.shape
    dev = Y0s.device
    rews = torch.empty(n, device=dev) if rews_out is None else rews_out
    rewss = torch.empty((n, H), device=dev) if want_rewss else None
    traj = torch.empty((n, H, 3), device=dev) if want_traj else None
    logpd, href = None, 0
    if xref is not None:
        href = xref.shape[0]
        logpd = torch.empty(n, device=dev) if logpd_out is None else logpd_out
    kp = None
    if key is not None:
        k, kp = key_ptr(key)
    check(_lib.lib().mbd_car2d_rollout(_p(params), _p(x0), kp, n_total, n_begin, n, H, ctypes.c_float(sigma), _p(Ybar), _p(Y0s),
                                       _p(rewss), _p(rews), _p(xref), href, _p(logpd), _p(traj), _stream()), "mbd_car2d_rollout")
    return dict(rews=rews, rewss=rewss, logpd=logpd, traj=traj)


def pusht_rollout(params, x0, Y0s, want_rewss=False, want_final=False, want_traj=False, key=None, n_total=0, n_begin=0,
                  sigma=0.0, Ybar=None, rews_out=None):
    """pushT rollouts (csrc/pusht.cuh); with `key` the noise is drawn in-kernel and written to Y0s [n,H,2]."""
    params, x0, Y0s = _dev(params), _dev(x0), _dev(Y0s)
    n, H, _ = Y0s.shape
    dev = Y0s.device
    rews = torch.empty(n, device=dev) if rews_out is None else rews_out
    rewss = torch.empty((n, H), device=dev) if want_rewss else None
    final = torch.empty((n, 16), device=dev) if want_final else None
    traj = torch.empty((n, H, 16), device=dev) if want_traj else None
    kp = None
    if key is not None:
        k, kp = key_ptr(key)
    check(_lib.lib().mbd_pusht_rollout(_p(params), _p(x0), kp, n_total, n_begin, n, H, ctypes.c_float(sigma), _p(Ybar), _p(Y0s),
                                       _p(rewss), _p(rews), _p(final), _p(traj), _stream()), "mbd_pusht_rollout")
    return dict(rews=rews, rewss=rewss, final=final, traj=traj, logpd=None)


def softmax_weights(rews_all, logpd_all, n_begin, n_local, temp, rew_xref, weights_out, scalars_out, scratch):
    n_total = rews_all.numel()
    check(_lib.lib().mbd_softmax_weights(_p(_dev(rews_all)), _p(logpd_all), n_total, n_begin, n_local, ctypes.c_float(temp),
                                         ctypes.c_float(rew_xref), _p(_dev(weights_out)), _p(_dev(scalars_out)), _p(_dev(scratch)),
                                         _stream()), "mbd_softmax_weights")


def weighted_sum(weights, Y0s, HNu, scratch, partial_out):
    n_local = weights.numel()
    check(_lib.lib().mbd_weighted_sum(_p(_dev(weights)), _p(_dev(Y0s)), n_local, HNu, _p(_dev(scratch)), _p(_dev(partial_out)),
                                      _stream()), "mbd_weighted_sum")


def weighted_sum_runs(weights, Y0s, HNu, runs_out) -> int:
    """first stage only; returns the number of 64-sample runs written to runs_out [nruns, HNu]"""
    rc = _lib.lib().mbd_weighted_sum_runs(_p(_dev(weights)), _p(_dev(Y0s)), weights.numel(), HNu, _p(_dev(runs_out)), _stream())
    if rc <= 0:
        check(rc if rc < 0 else -1, "mbd_weighted_sum_runs")
    return rc


def weighted_sqerr_sum(weights, Y0s, mu, HNu, scratch, partial_out):
    n_local = weights.numel()
    check(_lib.lib().mbd_weighted_sqerr_sum(_p(_dev(weights)), _p(_dev(Y0s)), _p(_dev(mu)), n_local, HNu, _p(_dev(scratch)),
                                            _p(_dev(partial_out)), _stream()), "mbd_weighted_sqerr_sum")


def peer_gather(peer_ptrs, P, rank, src_off_words, count, flag_off_words, epoch, dst, err):
    arr = (ctypes.c_uint64 * P)(*[int(p) for p in peer_ptrs])
    check(_lib.lib().mbd_peer_gather(arr, P, rank, src_off_words, count, flag_off_words, ctypes.c_uint32(epoch), _p(_dev(dst)),
                                     ctypes.c_void_p(err.data_ptr()), _stream()), "mbd_peer_gather")


def update(partials, P, HNu, Ybar_i, coef, out):
    c = (ctypes.c_float * 5)(*[float(v) for v in coef])
    check(_lib.lib().mbd_update(_p(_dev(partials)), P, HNu, _p(_dev(Ybar_i)), c, _p(_dev(out)), _stream()), "mbd_update")


def step_launch(plan: "_lib.StepPlan"):
    """one diffusion step (three launches, parameters in device memory): mbd_step_launch"""
    check(_lib.lib().mbd_step_launch(ctypes.byref(plan), _stream()), "mbd_step_launch")


def ffma_peak(device: Optional[torch.device] = None, iters: int = 4096) -> float:
    """measured fp32 FFMA throughput of the device in TFLOP/s (the fp32 roofline denominator)"""
    _lib.require_gpu()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(dev):
        scratch = torch.empty(1 << 20, device=dev)
        out = ctypes.c_float(0.0)
        check(_lib.lib().mbd_ffma_peak(_p(scratch), int(iters), ctypes.byref(out), _stream()), "mbd_ffma_peak")
    return float(out.value)


class Event:
    """raw CUDA event owned by the library (mbd_event_*): recordable from inside the C launch sequence"""

    def __init__(self):
        self.h = _lib.lib().mbd_event_create()
        if not self.h:
            raise MbdError("mbd_event_create failed")

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                _lib.lib().mbd_event_destroy(ctypes.c_void_p(h))
            except Exception:  # noqa: BLE001
                pass

    def record(self):
        check(_lib.lib().mbd_event_record(ctypes.c_void_p(self.h), _stream()), "mbd_event_record")

    def synchronize(self):
        check(_lib.lib().mbd_event_sync(ctypes.c_void_p(self.h)), "mbd_event_sync")

    def elapsed_ms(self, later: "Event") -> float:
        return float(_lib.lib().mbd_event_elapsed_ms(ctypes.c_void_p(self.h), ctypes.c_void_p(later.h)))


def step_launch_timed(plan: "_lib.StepPlan", before: Event, mid: Event, mid2: Event, after: Event):
    """events: before the rollout kernel | after it | after the statistics kernel | after the update kernel"""
    check(_lib.lib().mbd_step_launch_ev(ctypes.byref(plan), ctypes.c_void_p(before.h), ctypes.c_void_p(mid.h), ctypes.c_void_p(mid2.h),
                                         ctypes.c_void_p(after.h), _stream()), "mbd_step_launch_ev")
