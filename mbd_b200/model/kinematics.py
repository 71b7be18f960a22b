"""Host-side `pipeline_init`: (q, qd) -> per-link COM-frame state [L,13] float32.

Restates brax.positional.pipeline.init = kinematics.forward + com.from_world (call site
/root/reference/mbd/envs/humanoidrun.py:29).  Runs once per solve on the host in float64
and is rounded once to float32; it is the INPUT of the hot path (state_init), not part of it.
The derived quantities Brax also stores (j, jd, a_p, a_c) are re-derived from (x_i, xd_i) at
the start of every physics step by the kernel and the oracle alike.
State row: x_i.pos(3) x_i.rot(4, wxyz) xd_i.ang(3) xd_i.vel(3).
"""
from __future__ import annotations

import os

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
            # MuJoCo semantics: free-joint quaternions are unit.  Brax's positional kinematics.forward uses q[3:7] as it comes
            # [brax-recalled]: after the +-0.01 reset noise that is a ~1e-2 relative difference in the reset pose.
            # MBD_FREE_QUAT_NORMALIZE=0 selects the un-normalised reading (compatibility switch; unpinned either way).
            xrot[l] = r / np.linalg.norm(r) if os.environ.get("MBD_FREE_QUAT_NORMALIZE", "1") != "0" else r
            xvel[l] = qd[ds:ds + 3]
            xang[l] = qd[ds + 3:ds + 6]
            continue
        nd = int(sys.link_types[l])
        # stacked hinges: j.rot = r0*r1*r2; jd.ang = w0 + R(r0) w1 + R(r0 r1) w2 (link-transform frame)
        jrot = np.array([1.0, 0, 0, 0])
        jang = np.zeros(3)
        slide_pos, slide_vel = np.zeros(3), np.zeros(3)
        for k in range(nd):
            axis = sys.dof_axis[ds + k]
            if sys.dof_is_slide[ds + k]:  # prismatic dof: translation along the axis (link-transform frame)
                slide_pos = slide_pos + rotate(axis * (q[qs + k] - sys.ref(ds + k)), jrot)
                slide_vel = slide_vel + rotate(axis * qd[ds + k], jrot)
                continue
            jang = jang + rotate(axis * qd[ds + k], jrot)
            jrot = quat_mul(jrot, _quat_rot_axis(axis, q[qs + k] - sys.ref(ds + k)))
        jp = sys.joint_pos[l]
        jpos = jp - rotate(jp, jrot) + slide_pos  # rotation about the joint anchor (+ slide offset)
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
        xvel[l] = pvel + np.cross(pang, xpos[l] - ppos) + np.cross(w_rel, xpos[l] - anchor) + rotate(slide_vel, trot)
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


def inverse(sys: System, xpos, xrot, xang, xvel, links=None):
    """kinematics.inverse (restated): joint coordinates (q, qd) from world link transforms.
    Hinge angles are the intrinsic x-y'-z'' Euler angles of the joint-frame relative rotation,
    rates are the projections of the relative angular velocity on the instantaneous axes —
    the same quantities the physics step uses (axis_angle_ang)."""
    links = list(range(sys.num_links())) if links is None else list(links)
    q = np.array(sys.init_q, dtype=np.float64).copy()
    qd = np.zeros(sys.qd_size())
    for l in links:
        qs, ds = int(sys.link_q_start[l]), int(sys.link_dof_start[l])
        if sys.link_types[l] == "f":
            q[qs:qs + 3], q[qs + 3:qs + 7] = xpos[l], xrot[l]
            qd[ds:ds + 3], qd[ds + 3:ds + 6] = xvel[l], xang[l]
            continue
        nd = int(sys.link_types[l])
        par = sys.link_parents[l]
        prot = xrot[par] if par >= 0 else np.array([1.0, 0, 0, 0])
        pang = xang[par] if par >= 0 else np.zeros(3)
        a_p = quat_mul(quat_mul(prot, sys.link_rot[l]), sys.joint_rot[l])
        a_c = quat_mul(xrot[l], sys.joint_rot[l])
        j = quat_mul(a_p * np.array([1.0, -1, -1, -1]), a_c)
        w, x, y, z = j
        r00, r01, r02 = 1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)
        r12, r22 = 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)
        par_sign = sys.joint_parity[l]
        ang = [np.arctan2(-r12, r22), np.arctan2(r02, np.hypot(r00, r01)), par_sign * np.arctan2(-r01, r00)]
        lon = np.array([0.0, r22, -r12])
        lon = lon / (np.linalg.norm(lon) + 1e-30)
        axes = [np.array([1.0, 0, 0]), lon, par_sign * np.array([r02, r12, r22])]
        jd = rotate(xang[l] - pang, a_p * np.array([1.0, -1, -1, -1]))
        # slide dofs (world-parented links only, blob.pack enforces it): coordinate along the fixed parent-side axis
        trot = quat_mul(prot, sys.link_rot[l])
        ppos = xpos[par] if par >= 0 else np.zeros(3)
        anchor_p = ppos + rotate(sys.link_pos[l] + rotate(sys.joint_pos[l], sys.link_rot[l]), prot)
        rcw = rotate(sys.joint_pos[l], xrot[l])
        d = xpos[l] + rcw - anchor_p
        v_anchor = xvel[l] + np.cross(xang[l], rcw)
        for k in range(nd):
            if sys.dof_is_slide[ds + k]:
                axis_w = rotate(sys.dof_axis[ds + k], trot)
                q[qs + k] = sys.ref(ds + k) + float(np.dot(d, axis_w))
                qd[ds + k] = float(np.dot(v_anchor, axis_w))
                continue
            q[qs + k] = ang[k] + sys.ref(ds + k)
            qd[ds + k] = float(np.dot(axes[k], jd))
    return q.astype(np.float32), qd.astype(np.float32)
