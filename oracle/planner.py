"""CPU ORACLE (test infrastructure) — numpy restatement of the planner arithmetic of
/root/reference/mbd/planners/mbd_planner.py:84-148 on top of the C rollouts in mbd_oracle.c.

The statistics mirror the reference's jnp calls one to one with numpy float32 (mean, population
std, guard, softmax with max subtraction, einsum); the GPU path uses its own deterministic
reduction order, so comparisons at this level are tolerance based (rtol 1e-4 as north_star
states), while everything below (noise, per-sample returns) is compared bit for bit.
Pinned: the schedule KATs of SURVEY section 8(d) (sigmas[-1] = 0.6305 / 0.7981 / 0.8839).
"""
from __future__ import annotations

import numpy as np

from . import oracle as orc

f32 = np.float32


def make_schedule(beta0, betaT, Ndiffuse):
    """mbd_planner.py:84-87."""
    # jnp.linspace in float32 [jax-recalled]: start*(1 - k/(N-1)) + stop*(k/(N-1)), endpoint appended (see mbd_b200/planners/engine.py)
    import os
    if os.environ.get("MBD_LINSPACE", "jax") == "numpy" or Ndiffuse < 2:
        betas = np.linspace(beta0, betaT, Ndiffuse, dtype=f32)
    else:
        step = (np.arange(Ndiffuse - 1, dtype=f32) / f32(Ndiffuse - 1)).astype(f32)
        betas = np.concatenate([((f32(beta0) * (f32(1.0) - step)).astype(f32) + (f32(betaT) * step).astype(f32)).astype(f32),
                                np.array([betaT], dtype=f32)])
    alphas = (f32(1.0) - betas).astype(f32)
    alphas_bar = np.cumprod(alphas, dtype=f32)
    sigmas = np.sqrt(f32(1.0) - alphas_bar).astype(f32)
    return betas, alphas, alphas_bar, sigmas


def softmax(x):
    """jax.nn.softmax: exp(x - max) / sum."""
    e = np.exp((x - x.max()).astype(f32)).astype(f32)
    return (e / e.sum(dtype=f32)).astype(f32)


def reverse_once_stats(rews, Y0s, temp, logpd=None, rew_xref=0.0):
    """mbd_planner.py:110-128: returns (Ybar [HNu], rews.mean(), weights)."""
    rews = rews.astype(f32)
    rew_std = rews.std(dtype=f32)
    rew_std = f32(1.0) if rew_std < 1e-4 else rew_std
    rew_mean = rews.mean(dtype=f32)
    logp0 = ((rews - rew_mean) / rew_std / f32(temp)).astype(f32)
    if logpd is not None:
        xl = (logpd - logpd.max()).astype(f32)
        logpdemo = ((xl + f32(rew_xref) - rew_mean) / rew_std / f32(temp)).astype(f32)
        mask = logpdemo > logp0
        logp0 = np.where(mask, logpdemo, logp0).astype(f32)
        logp0 = ((logp0 - logp0.mean(dtype=f32)) / logp0.std(dtype=f32) / f32(temp)).astype(f32)
    w = softmax(logp0)
    Ybar = np.einsum("n,nj->j", w.astype(np.float64), Y0s.astype(np.float64)).astype(f32)  # exact-ish fp32 dot
    return Ybar, rew_mean, w


def update(Ybar_i, Ybar, alphas, alphas_bar, i):
    """mbd_planner.py:100,130-133 literally, float32."""
    ab = f32(alphas_bar[i])
    Yi = (Ybar_i * np.sqrt(ab)).astype(f32)
    score = (f32(1.0) / (f32(1.0) - ab) * (-Yi + np.sqrt(ab) * Ybar)).astype(f32)
    Yim1 = (f32(1.0) / np.sqrt(f32(alphas[i])) * (Yi + (f32(1.0) - ab) * score)).astype(f32)
    return (Yim1 / np.sqrt(f32(alphas_bar[i - 1]))).astype(f32)


class OracleEnv:
    """Rollout backend for the oracle planner: kind 'xpbd' (blob, state), 'car2d' (params, x0) or 'pusht' (params, x0)."""

    def __init__(self, kind, Nu, **kw):
        self.kind, self.Nu, self.kw = kind, Nu, kw

    def rollout(self, Y0s, H, xref=None, nthreads=0):
        n = Y0s.shape[0]
        Y = Y0s.reshape(n, H, self.Nu)
        if self.kind == "xpbd" and self.kw.get("simd") and xref is None:
            out = orc.simd_rollout(self.kw["blob"], self.kw["state"], Y, nthreads=nthreads)   # SIMD across samples, same bits
            if out is not None:
                return out
        if self.kind == "xpbd":
            return orc.xpbd_rollout(self.kw["blob"], self.kw["state"], Y, xref=xref, nthreads=nthreads)
        if self.kind == "pusht":
            return orc.pusht_rollout(self.kw["params"], self.kw["x0"], Y, nthreads=nthreads)
        return orc.car2d_rollout(self.kw["params"], self.kw["x0"], Y, xref=xref, nthreads=nthreads)


def reverse_once(env: OracleEnv, key, Nsample, H, sigma, Ybar_i, temp, alphas, alphas_bar, i, xref=None, rew_xref=0.0,
                 nthreads=0):
    """One diffusion step; `key` is Y0s_rng (already split off).  Returns dict."""
    HNu = H * env.Nu
    Y0s = orc.sample_Y0s(key, Nsample, HNu, sigma, Ybar_i, nthreads=nthreads)
    out = env.rollout(Y0s, H, xref=xref, nthreads=nthreads)
    Ybar, rew_mean, w = reverse_once_stats(out["rews"], Y0s, temp, logpd=out["logpd"] if xref is not None else None,
                                           rew_xref=rew_xref)
    return dict(Y0s=Y0s, rews=out["rews"], logpd=out["logpd"], Ybar=Ybar, weights=w, rew_mean=rew_mean,
                Ybar_im1=update(Ybar_i, Ybar, alphas, alphas_bar, i))


def run_diffusion(env: OracleEnv, seed, Nsample, H, Ndiffuse, temp, beta0=1e-4, betaT=1e-2, xref=None, rew_xref=0.0,
                  nthreads=0, key_after_reset=None):
    """mbd_planner.py:38-182 (without rendering).  Returns (rew_final, Yi [Ndiffuse-1, HNu], rews)."""
    rng = orc.prng_key(seed)
    rng, _rng_reset = orc.split(rng)
    betas, alphas, alphas_bar, sigmas = make_schedule(beta0, betaT, Ndiffuse)
    rng_exp, rng = orc.split(rng)
    r = rng_exp
    Yb = np.zeros(H * env.Nu, dtype=f32)
    Ybars, rews = [], []
    for i in range(Ndiffuse - 1, 0, -1):
        r, k = orc.split(r)
        o = reverse_once(env, k, Nsample, H, float(sigmas[i]), Yb, temp, alphas, alphas_bar, i, xref=xref, rew_xref=rew_xref,
                         nthreads=nthreads)
        Yb = o["Ybar_im1"]
        Ybars.append(Yb)
        rews.append(o["rew_mean"])
    Yi = np.stack(Ybars)
    fin = env.rollout(Yi[-1][None], H, nthreads=1)
    return float(fin["rews"][0]), Yi, np.array(rews, dtype=f32)


def update_once(env: OracleEnv, key, Nsample, H, sigma, mu_0t, temp, method, nthreads=0):
    """path_integral.py:111-127 + the three update rules (:33-52).  Returns dict(mu, sigma, rew_mean, idx)."""
    HNu = H * env.Nu
    Y0s = orc.sample_Y0s(key, Nsample, HNu, sigma, mu_0t, nthreads=nthreads)
    rews = env.rollout(Y0s, H, nthreads=nthreads)["rews"]
    std = rews.std(dtype=f32)
    std = f32(1.0) if std < 1e-4 else std   # shared guard (see mbd_b200/planners/path_integral.py docstring)
    logp0 = ((rews - rews.mean(dtype=f32)) / std / f32(temp)).astype(f32)
    w = softmax(logp0)
    idx = None
    if method == "cem":
        idx = np.argsort(w, kind="stable")[::-1][:10]
        mu = Y0s[idx].mean(axis=0, dtype=f32)
    else:
        mu = np.einsum("n,nj->j", w.astype(np.float64), Y0s.astype(np.float64)).astype(f32)
        if method == "cma-es":
            err = (Y0s - mu_0t[None]).astype(f32)
            var = np.einsum("n,nj->j", w.astype(np.float64), (err * err).astype(np.float64)).astype(f32)
            sigma = float(np.sqrt(var).mean(dtype=f32)) * sigma
            sigma = max(sigma, 1e-3)
    return dict(Y0s=Y0s, rews=rews, weights=w, mu=mu, sigma=sigma, rew_mean=rews.mean(dtype=f32), idx=idx)
