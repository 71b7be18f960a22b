"""Host-side `pipeline_init`: (q, qd) -> per-link COM-frame state [L,13] float32.

Restates brax.positional.pipeline.init = kinematics.forward + com.from_world (call site
/root/reference/mbd/envs/humanoidrun.py:29).  Runs once per solve on the host in float64
and is rounded once to float32; it is the INPUT of the hot path (state_init), not part of it.
The derived quantities Brax also stores (j, jd, a_p, a_c) are re-derived from (x_i, xd_i) at
the start of every physics step by the kernel and the oracle alike.
State row: x_i.pos(3) x_i.rot(4, wxyz) xd_i.ang(3) xd_i.vel(3).
"""
from __future__ import annotations

import numpy as np

from .mjcf import System, quat_mul, rotate


def _quat_rot_axis(axis, angle):
    s = np.sin(angle / 2)
    return np.array([np.cos(angle / 2), axis[0] * s, axis[1] * s, axis[2] * s])


def forward(sys: System, q: np.ndarray, qd: np.ndarray):
    """kinematics.forward: world link transforms x (pos, rot) and motions xd (ang, vel)."""
    L = sys.num_links()
    q = np.asarray(q, dtype=np.float64)
    qd = np.asarray(qd, dtype=np.float64)
    xpos, xrot = np.zeros((L, 3)), np.zeros((L, 4))
    xang, xvel = np.zeros((L, 3)), np.zeros((L, 3))
    for l in range(L):
        qs, ds = int(sys.link_q_start[l]), int(sys.link_dof_start[l])
        par = sys.link_parents[l]
        if sys.link_types[l] == "f":
            xpos[l] = q[qs:qs + 3]
            r = q[qs + 3:qs + 7]
            xrot[l] = r / np.linalg.norm(r)  # MuJoCo semantics: free-joint quaternions are unit
            xvel[l] = qd[ds:ds + 3]
            xang[l] = qd[ds + 3:ds + 6]
            continue
        nd = int(sys.link_types[l])
        if np.any(sys.dof_is_slide[ds:ds + nd]):
            raise NotImplementedError("slide joints")
        # stacked hinges: j.rot = r0*r1*r2; jd.ang = w0 + R(r0) w1 + R(r0 r1) w2 (link-transform frame)
        jrot = np.array([1.0, 0, 0, 0])
        jang = np.zeros(3)
        for k in range(nd):
            axis = sys.dof_axis[ds + k]
            jang = jang + rotate(axis * qd[ds + k], jrot)
            jrot = quat_mul(jrot, _quat_rot_axis(axis, q[qs + k]))
        jp = sys.joint_pos[l]
        jpos = jp - rotate(jp, jrot)  # rotation about the joint anchor
        ppos = xpos[par] if par >= 0 else np.zeros(3)
        prot = xrot[par] if par >= 0 else np.array([1.0, 0, 0, 0])
        pang = xang[par] if par >= 0 else np.zeros(3)
        pvel = xvel[par] if par >= 0 else np.zeros(3)
        tpos = ppos + rotate(sys.link_pos[l], prot)
        trot = quat_mul(prot, sys.link_rot[l])
        xpos[l] = tpos + rotate(jpos, trot)
        xrot[l] = quat_mul(trot, jrot)
        xrot[l] /= np.linalg.norm(xrot[l])
        w_rel = rotate(jang, trot)
        xang[l] = pang + w_rel
        anchor = xpos[l] + rotate(jp, xrot[l])
        xvel[l] = pvel + np.cross(pang, xpos[l] - ppos) + np.cross(w_rel, xpos[l] - anchor)
    return xpos, xrot, xang, xvel


def pipeline_init(sys: System, q, qd, links=None) -> np.ndarray:
    xpos, xrot, xang, xvel = forward(sys, q, qd)
    L = sys.num_links()
    st = np.zeros((L, 13), dtype=np.float64)
    for l in range(L):
        rc = rotate(sys.com[l], xrot[l])
        st[l, 0:3] = xpos[l] + rc                       # com.from_world: x_i = x.do(inertia.transform)
        st[l, 3:7] = xrot[l]
        st[l, 7:10] = xang[l]
        st[l, 10:13] = xvel[l] + np.cross(xang[l], rc)  # xd_i.vel = xd.vel + ang x (x_i.pos - x.pos)
    if links is not None:
        st = st[list(links)]
    return st.astype(np.float32)


def to_world(sys: System, state: np.ndarray, links=None):
    """com.to_world on a [L,13] state: returns x.pos [L,3], x.rot, xd.ang, xd.vel (float64 host view)."""
    links = list(range(sys.num_links())) if links is None else list(links)
    st = np.asarray(state, dtype=np.float64)
    pos, vel = np.zeros((len(links), 3)), np.zeros((len(links), 3))
    for i, l in enumerate(links):
        rc = rotate(sys.com[l], st[i, 3:7])
        pos[i] = st[i, 0:3] - rc
        vel[i] = st[i, 10:13] + np.cross(rc, st[i, 7:10])
    return pos, st[:, 3:7].copy(), st[:, 7:10].copy(), vel
