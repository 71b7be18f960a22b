"""MJCF -> `System` compiler for the subset of MuJoCo XML the MBD envs use.

Replaces `brax.io.mjcf.load` at the call sites `/root/reference/mbd/envs/humanoidrun.py:15`,
`humanoidtrack.py:16`, `hopper.py:14` (Brax itself is NOT vendored in the reference; this
is a from-scratch restatement of its documented behaviour, see DESIGN.md "parity unpinned"):

  * XML-level fusing of joint-less bodies into their parent (Brax `_fuse_bodies`):
    the humanoid feet become geoms of the shins.
  * MuJoCo `inertiafromgeom` rules for sphere / capsule (density 1000 by default):
    mass, centre of mass and inertia tensor per body via the parallel-axis theorem.
  * `compiler angle="degree"`, `<default>` classes for joint / geom / motor, body quats
    are normalised, joint axes are normalised.
  * Brax `<custom><numeric>` parameters (constraint_*, ang_damping, joint_scale_*, ...).
  * links = MuJoCo bodies 1.. in depth-first order; `link_types` is 'f' for a free joint or
    the number of stacked 1-dof joints; dofs/actuators are mapped by joint name.

Only what the positional (XPBD) pipeline reads is produced.  Everything is computed in
float64 and rounded once to float32 when the device blob is packed (see blob.py).
"""
from __future__ import annotations

import dataclasses
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional

import numpy as np

# ----------------------------------------------------------------------------------
# small float64 helpers (host side only)
# ----------------------------------------------------------------------------------


def _vec(s: Optional[str], default) -> np.ndarray:
    if s is None:
        return np.array(default, dtype=np.float64)
    return np.array([float(v) for v in s.split()], dtype=np.float64)


def quat_mul(u, v):
    return np.array([
        u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3],
        u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2],
        u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1],
        u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0],
    ])


def quat_inv(q):
    return q * np.array([1.0, -1.0, -1.0, -1.0])


def rotate(v, q):
    s, u = q[0], q[1:]
    return 2 * np.dot(u, v) * u + (s * s - np.dot(u, u)) * v + 2 * s * np.cross(u, v)


def quat_from_3x3(m):
    """Rotation matrix (columns = frame axes) -> unit quaternion [w,x,y,z]."""
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = np.array([0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s])
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = np.array([(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s])
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = np.array([(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s])
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = np.array([(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s])
    q = q / np.linalg.norm(q)
    return q if q[0] >= 0 else -q


def quat_to_3x3(q):
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def orthogonals(a):
    """Brax `math.orthogonals`: two unit vectors orthogonal to unit vector a."""
    y, z = np.array([0.0, 1.0, 0.0]), np.array([0.0, 0.0, 1.0])
    b = y if (-0.5 < a[1] < 0.5) else z
    b = b - a * np.dot(a, b)
    b = b / np.linalg.norm(b)
    return b, np.cross(a, b)


# ----------------------------------------------------------------------------------
# parsed model
# ----------------------------------------------------------------------------------


@dataclasses.dataclass
class Geom:
    name: str
    type: str  # sphere | capsule | plane
    pos: np.ndarray  # centre in body frame
    quat: np.ndarray  # orientation in body frame (capsule: local z = axis)
    size: np.ndarray  # sphere: [r]; capsule: [r, half_length]; plane: size
    density: float
    contype: int
    conaffinity: int
    friction: float
    body: int = -1  # link index (-1 = world)


@dataclasses.dataclass
class Joint:
    name: str
    type: str  # free | hinge | slide
    pos: np.ndarray
    axis: np.ndarray
    range: np.ndarray  # radians (hinge) / metres (slide); +-inf when unlimited
    stiffness: float
    damping: float
    armature: float
    ref: float = 0.0  # MuJoCo `ref`: the joint coordinate of the pose written in the XML (qpos0)


@dataclasses.dataclass
class Body:
    name: str
    pos: np.ndarray
    quat: np.ndarray
    parent: int
    joints: List[Joint]
    geoms: List[Geom]


@dataclasses.dataclass
class System:
    """The fields of Brax's `System` that the positional pipeline and the MBD envs touch."""

    link_names: List[str]
    link_parents: List[int]
    link_types: str
    # link.transform (relative to parent link frame; identity for free links)
    link_pos: np.ndarray  # [L,3]
    link_rot: np.ndarray  # [L,4]
    # link.joint (anchor + joint frame, in the child link frame)
    joint_pos: np.ndarray  # [L,3]
    joint_rot: np.ndarray  # [L,4]
    joint_parity: np.ndarray  # [L] (+-1; -1 when the 3 hinge axes are left-handed)
    # link.inertia
    mass: np.ndarray  # [L]
    com: np.ndarray  # [L,3] inertia.transform.pos
    inertia: np.ndarray  # [L,3,3] about COM in link frame (unused when spring_inertia_scale == 1)
    # dofs (qd layout: free = 6 [vel, ang], then one per 1-dof joint)
    dof_link: np.ndarray  # [nv]
    dof_axis: np.ndarray  # [nv,3] in link frame (zeros for free dofs)
    dof_is_slide: np.ndarray  # [nv] bool
    dof_stiffness: np.ndarray
    dof_damping: np.ndarray
    dof_armature: np.ndarray
    dof_limit: np.ndarray  # [nv,2]
    link_dof_start: np.ndarray  # [L] first qd index of each link
    link_q_start: np.ndarray  # [L] first q index of each link
    # actuators
    act_names: List[str]
    act_qd_id: np.ndarray  # [nu]
    act_q_id: np.ndarray
    act_gear: np.ndarray
    act_ctrl_range: np.ndarray  # [nu,2]
    act_gain: np.ndarray
    act_bias_q: np.ndarray
    act_bias_qd: np.ndarray
    # geoms / static contact list
    geoms: List[Geom]
    contacts: List[dict]  # each: {link, pos[3] (link frame), radius, friction, geom}
    # options and Brax custom numerics
    dt: float
    gravity: np.ndarray
    init_q: np.ndarray
    custom: Dict[str, float]
    dof_ref: Optional[np.ndarray] = None  # [nv] MuJoCo `ref` per 1-dof joint (zeros when absent): displacement = q - ref

    def ref(self, d: int) -> float:
        return 0.0 if self.dof_ref is None else float(self.dof_ref[d])

    # Brax-compatible accessors used by the reference envs ---------------------------
    def q_size(self) -> int:
        return int(self.init_q.shape[0])

    def qd_size(self) -> int:
        return int(self.dof_link.shape[0])

    def act_size(self) -> int:
        return int(self.act_qd_id.shape[0])

    def num_links(self) -> int:
        return len(self.link_names)

    @property
    def ang_damping(self):
        return self.custom["ang_damping"]

    @property
    def vel_damping(self):
        return self.custom["vel_damping"]


_CUSTOM_DEFAULTS = {
    # Brax defaults (brax/io/mjcf.py) for the positional / spring pipelines
    "vel_damping": 0.0,
    "ang_damping": 0.0,
    "baumgarte_erp": 0.1,
    "spring_mass_scale": 0.0,
    "spring_inertia_scale": 0.0,
    "joint_scale_pos": 0.5,
    "joint_scale_ang": 0.2,
    "collide_scale": 1.0,
    "elasticity": 0.0,
    "constraint_stiffness": 2000.0,
    "constraint_limit_stiffness": 1000.0,
    "constraint_vel_damping": 100.0,
    "constraint_ang_damping": 0.0,
    "matrix_inv_iterations": 10,
    "solver_maxls": 20,
}


class _Defaults:
    """<default> handling: top-level defaults + named classes (childclass / class attr)."""

    def __init__(self, root: ET.Element):
        self.classes: Dict[str, Dict[str, Dict[str, str]]] = {}
        top = root.find("default")
        self._walk(top, "main", {})

    def _walk(self, node, name, inherited):
        cur = {k: dict(v) for k, v in inherited.items()}
        if node is not None:
            for child in node:
                if child.tag == "default":
                    continue
                cur.setdefault(child.tag, {}).update(child.attrib)
        self.classes[name] = cur
        if node is not None:
            for child in node.findall("default"):
                self._walk(child, child.get("class"), cur)

    def resolve(self, elem: ET.Element, tag: str, active_class: str) -> Dict[str, str]:
        cls = elem.get("class", active_class)
        out = dict(self.classes.get(cls, self.classes["main"]).get(tag, {}))
        out.update(elem.attrib)
        return out


def _capsule_frame_from_fromto(fromto: np.ndarray):
    a, b = fromto[:3], fromto[3:]
    centre = 0.5 * (a + b)
    d = b - a
    half = 0.5 * np.linalg.norm(d)
    z = d / (np.linalg.norm(d) + 1e-300)
    x, y = orthogonals(z)
    quat = quat_from_3x3(np.stack([x, y, z], axis=1))
    return centre, quat, half


def _geom_mass_inertia(g: Geom):
    """MuJoCo inertiafromgeom for one geom: mass, inertia tensor about its centre (geom frame)."""
    rho = g.density
    if g.type == "sphere":
        r = g.size[0]
        m = rho * 4.0 / 3.0 * np.pi * r ** 3
        i = np.eye(3) * (0.4 * m * r * r)
    elif g.type == "capsule":
        r, h = g.size[0], g.size[1]
        mc = rho * np.pi * r * r * 2 * h
        ms = rho * 4.0 / 3.0 * np.pi * r ** 3
        m = mc + ms
        izz = mc * r * r / 2 + 0.4 * ms * r * r
        ixx = mc * (r * r / 4 + h * h / 3) + ms * (0.4 * r * r + h * h + 0.75 * r * h)
        i = np.diag([ixx, ixx, izz])
    else:
        return 0.0, np.zeros((3, 3))
    return m, i


def _parse_geom(elem, defaults, active_class, degree: bool = True) -> Geom:
    a = defaults.resolve(elem, "geom", active_class)
    typ = a.get("type", "sphere")
    size = _vec(a.get("size"), [0.0])
    pos = _vec(a.get("pos"), [0, 0, 0])
    quat = _vec(a.get("quat"), [1, 0, 0, 0])
    quat = quat / np.linalg.norm(quat)
    if "axisangle" in a:  # MuJoCo axisangle="x y z angle" (angle in the compiler's unit)
        aa = _vec(a["axisangle"], None)
        ang = np.deg2rad(aa[3]) if degree else aa[3]
        ax = aa[:3] / np.linalg.norm(aa[:3])
        quat = np.concatenate([[np.cos(ang / 2)], ax * np.sin(ang / 2)])
    if typ == "capsule":
        if "fromto" in a:
            pos, quat, half = _capsule_frame_from_fromto(_vec(a["fromto"], None))
            size = np.array([size[0], half])
        else:
            size = np.array([size[0], size[1]])
    elif typ not in ("sphere", "plane"):
        raise NotImplementedError(f"geom type {typ!r} is outside the positional-env subset")
    fr = _vec(a.get("friction"), [1.0, 0.005, 0.0001])
    return Geom(
        name=a.get("name", ""), type=typ, pos=pos, quat=quat, size=size,
        density=float(a.get("density", 1000.0)),
        contype=int(a.get("contype", 1)), conaffinity=int(a.get("conaffinity", 1)),
        friction=float(fr[0]),
    )


def _parse_joint(elem, defaults, active_class, degree: bool) -> Joint:
    a = defaults.resolve(elem, "joint", active_class)
    typ = a.get("type", "hinge")
    if typ not in ("free", "hinge", "slide"):
        raise NotImplementedError(f"joint type {typ!r}")
    axis = _vec(a.get("axis"), [0, 0, 1])
    axis = axis / np.linalg.norm(axis)
    limited = a.get("limited", "auto")
    has_range = "range" in a
    rng = _vec(a.get("range"), [0, 0])
    if typ == "hinge" and degree:
        rng = np.deg2rad(rng)
    is_limited = (limited == "true") or (limited == "auto" and has_range)
    if typ == "free" or not is_limited:
        rng = np.array([-np.inf, np.inf])
    return Joint(
        name=a.get("name", ""), type=typ, pos=_vec(a.get("pos"), [0, 0, 0]), axis=axis, range=rng,
        stiffness=float(a.get("stiffness", 0.0)), damping=float(a.get("damping", 0.0)),
        armature=float(a.get("armature", 0.0)),
        ref=(np.deg2rad(float(a.get("ref", 0.0))) if (typ == "hinge" and degree) else float(a.get("ref", 0.0))),
    )


def _collect_bodies(elem, parent, defaults, active_class, degree, out: List[Body], world_geoms: List[Geom]):
    """Depth-first walk.  Joint-less bodies are fused into their parent (Brax `_fuse_bodies`)."""
    for child in elem:
        if child.tag == "geom" and parent == -1:
            g = _parse_geom(child, defaults, active_class, degree)
            g.body = -1
            world_geoms.append(g)
    for b in elem.findall("body"):
        cls = b.get("childclass", active_class)
        pos = _vec(b.get("pos"), [0, 0, 0])
        quat = _vec(b.get("quat"), [1, 0, 0, 0])
        quat = quat / np.linalg.norm(quat)
        joints = [_parse_joint(j, defaults, cls, degree) for j in b.findall("joint")]
        if b.find("freejoint") is not None:
            joints = [Joint("root", "free", np.zeros(3), np.array([0, 0, 1.0]), np.array([-np.inf, np.inf]), 0, 0, 0)]
        geoms = [_parse_geom(g, defaults, cls, degree) for g in b.findall("geom")]
        if not joints and parent >= 0:
            # fuse into parent: re-express geoms (and the subtree) in the parent's frame
            for g in geoms:
                g.pos = pos + rotate(g.pos, quat)
                g.quat = quat_mul(quat, g.quat)
                out[parent].geoms.append(g)
            sub = ET.Element("fused")
            for gb in b.findall("body"):
                gb2 = ET.fromstring(ET.tostring(gb))
                p2 = pos + rotate(_vec(gb2.get("pos"), [0, 0, 0]), quat)
                q2 = quat_mul(quat, _vec(gb2.get("quat"), [1, 0, 0, 0]))
                gb2.set("pos", " ".join(repr(float(v)) for v in p2))
                gb2.set("quat", " ".join(repr(float(v)) for v in q2))
                sub.append(gb2)
            _collect_bodies(sub, parent, defaults, cls, degree, out, world_geoms)
            continue
        idx = len(out)
        out.append(Body(b.get("name", f"body{idx}"), pos, quat, parent, joints, geoms))
        _collect_bodies(b, idx, defaults, cls, degree, out, world_geoms)


def load(path: str) -> System:
    root = ET.parse(path).getroot()
    compiler = root.find("compiler")
    degree = True
    if compiler is not None and compiler.get("angle", "degree") == "radian":
        degree = False
    defaults = _Defaults(root)
    option = root.find("option")
    dt = float(option.get("timestep", 0.002)) if option is not None else 0.002
    gravity = _vec(option.get("gravity") if option is not None else None, [0, 0, -9.81])

    custom = dict(_CUSTOM_DEFAULTS)
    cust = root.find("custom")
    if cust is not None:
        for n in cust.findall("numeric"):
            vals = [float(v) for v in n.get("data").split()]
            custom[n.get("name")] = vals[0] if len(vals) == 1 else np.array(vals)

    bodies: List[Body] = []
    world_geoms: List[Geom] = []
    _collect_bodies(root.find("worldbody"), -1, defaults, "main", degree, bodies, world_geoms)
    L = len(bodies)

    link_types = ""
    for b in bodies:
        if b.joints and b.joints[0].type == "free":
            link_types += "f"
        else:
            if not 1 <= len(b.joints) <= 3:
                raise NotImplementedError(f"link {b.name}: {len(b.joints)} joints")
            link_types += str(len(b.joints))

    link_pos = np.zeros((L, 3)); link_rot = np.tile([1.0, 0, 0, 0], (L, 1))
    joint_pos = np.zeros((L, 3)); joint_rot = np.tile([1.0, 0, 0, 0], (L, 1)); parity = np.ones(L)
    mass = np.zeros(L); com = np.zeros((L, 3)); inertia = np.zeros((L, 3, 3))
    dof_link, dof_axis, dof_slide, dof_k, dof_d, dof_arm, dof_lim, dof_ref = [], [], [], [], [], [], [], []
    link_dof_start = np.zeros(L, dtype=np.int64); link_q_start = np.zeros(L, dtype=np.int64)
    init_q: List[float] = []
    joint_dof: Dict[str, int] = {}; joint_q: Dict[str, int] = {}
    all_geoms: List[Geom] = list(world_geoms)

    for i, b in enumerate(bodies):
        link_dof_start[i] = len(dof_link); link_q_start[i] = len(init_q)
        if link_types[i] == "f":
            # "mujoco stores free q in world frame, so clear link transform for free links"
            init_q += list(b.pos) + list(b.quat)
            joint_dof[b.joints[0].name] = len(dof_link); joint_q[b.joints[0].name] = link_q_start[i]
            for _ in range(6):
                dof_link.append(i); dof_axis.append(np.zeros(3)); dof_slide.append(False)
                dof_k.append(0.0); dof_d.append(0.0); dof_arm.append(0.0); dof_lim.append([-np.inf, np.inf]); dof_ref.append(0.0)
        else:
            link_pos[i], link_rot[i] = b.pos, b.quat
            axes = [j.axis for j in b.joints]
            joint_pos[i] = b.joints[0].pos
            for j in b.joints:
                joint_dof[j.name] = len(dof_link); joint_q[j.name] = len(init_q)
                init_q.append(j.ref)   # MuJoCo qpos0 = ref
                dof_ref.append(j.ref)
                dof_link.append(i); dof_axis.append(j.axis); dof_slide.append(j.type == "slide")
                dof_k.append(j.stiffness); dof_d.append(j.damping); dof_arm.append(j.armature)
                dof_lim.append(list(j.range))
            # joint frame (Brax kinematics.link_to_joint_frame): x = dof0 axis, y = dof1, z = dof2
            if len(axes) == 1:
                o1, o2 = orthogonals(axes[0])
                frame = np.stack([axes[0], o1, o2], axis=1)
            elif len(axes) == 2:
                frame = np.stack([axes[0], axes[1], np.cross(axes[0], axes[1])], axis=1)
            else:
                par = float(np.sign(np.dot(np.cross(axes[0], axes[1]), axes[2])))
                parity[i] = par
                frame = np.stack([axes[0], axes[1], axes[2] * par], axis=1)
            if np.abs(frame.T @ frame - np.eye(3)).max() > 1e-6:
                raise NotImplementedError(f"link {b.name}: stacked joint axes are not orthonormal")
            joint_rot[i] = quat_from_3x3(frame)
        # inertia from geoms
        m_tot, mc = 0.0, np.zeros(3)
        parts = []
        for g in b.geoms:
            g.body = i
            all_geoms.append(g)
            m, ig = _geom_mass_inertia(g)
            parts.append((m, ig, g))
            m_tot += m; mc += m * g.pos
        if m_tot <= 0:
            raise ValueError(f"body {b.name} has no mass")
        c = mc / m_tot
        it = np.zeros((3, 3))
        for m, ig, g in parts:
            r = quat_to_3x3(g.quat)
            d = g.pos - c
            it += r @ ig @ r.T + m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
        mass[i], com[i], inertia[i] = m_tot, c, it

    if compiler is not None and compiler.get("settotalmass") is not None:
        # MuJoCo compiler settotalmass: scale every body mass (and inertia) so that the masses add up to the given value
        k = float(compiler.get("settotalmass")) / mass.sum()
        mass, inertia = mass * k, inertia * k

    # actuators (motor / general with gear; position actuators carry bias)
    act_names, act_qd, act_q, act_gear, act_ctrl, act_gain, act_bq, act_bqd = [], [], [], [], [], [], [], []
    actuator = root.find("actuator")
    if actuator is not None:
        for a_el in actuator:
            a = defaults.resolve(a_el, a_el.tag, "main")
            if a_el.tag not in ("motor", "general", "position"):
                raise NotImplementedError(f"actuator {a_el.tag}")
            jn = a["joint"]
            act_names.append(a.get("name", jn))
            act_qd.append(joint_dof[jn]); act_q.append(joint_q[jn])
            act_gear.append(_vec(a.get("gear"), [1.0])[0])
            if a.get("ctrllimited", "false") == "true" or ("ctrlrange" in a and a.get("ctrllimited", "auto") == "auto"):
                act_ctrl.append(list(_vec(a.get("ctrlrange"), [0, 0])))
            else:
                act_ctrl.append([-np.inf, np.inf])
            if a_el.tag == "position":
                kp = float(a.get("kp", 1.0)); act_gain.append(kp); act_bq.append(-kp); act_bqd.append(0.0)
            else:
                act_gain.append(1.0); act_bq.append(0.0); act_bqd.append(0.0)

    # static contact list: sphere (on a link) vs plane (on the world), filtered by
    # (contype_a & conaffinity_b) | (contype_b & conaffinity_a)   [MuJoCo broadphase rule]
    contacts = []
    planes = [g for g in all_geoms if g.type == "plane" and g.body == -1]
    for g in all_geoms:
        if g.body < 0 or g.type == "plane":
            continue
        for pl in planes:
            if (g.contype & pl.conaffinity) | (pl.contype & g.conaffinity):
                if g.type == "sphere":
                    contacts.append(dict(link=g.body, pos=g.pos.copy(), radius=float(g.size[0]),
                                         friction=max(g.friction, pl.friction), geom=g.name,
                                         plane_pos=pl.pos.copy(), plane_normal=rotate(np.array([0, 0, 1.0]), pl.quat)))
                elif g.type == "capsule":
                    axis = rotate(np.array([0, 0, 1.0]), g.quat)
                    for sgn in (-1.0, 1.0):  # MJX plane_capsule: one contact per end cap
                        contacts.append(dict(link=g.body, pos=g.pos + sgn * g.size[1] * axis, radius=float(g.size[0]),
                                             friction=max(g.friction, pl.friction), geom=g.name,
                                             plane_pos=pl.pos.copy(), plane_normal=rotate(np.array([0, 0, 1.0]), pl.quat)))

    return System(
        link_names=[b.name for b in bodies], link_parents=[b.parent for b in bodies], link_types=link_types,
        link_pos=link_pos, link_rot=link_rot, joint_pos=joint_pos, joint_rot=joint_rot, joint_parity=parity,
        mass=mass, com=com, inertia=inertia,
        dof_link=np.array(dof_link, dtype=np.int64), dof_axis=np.array(dof_axis), dof_is_slide=np.array(dof_slide),
        dof_stiffness=np.array(dof_k), dof_damping=np.array(dof_d), dof_armature=np.array(dof_arm),
        dof_limit=np.array(dof_lim), link_dof_start=link_dof_start, link_q_start=link_q_start,
        act_names=act_names, act_qd_id=np.array(act_qd, dtype=np.int64), act_q_id=np.array(act_q, dtype=np.int64),
        act_gear=np.array(act_gear), act_ctrl_range=np.array(act_ctrl).reshape(-1, 2), act_gain=np.array(act_gain),
        act_bias_q=np.array(act_bq), act_bias_qd=np.array(act_bqd),
        geoms=all_geoms, contacts=contacts, dt=dt, gravity=gravity,
        # Brax: a <custom><numeric name="init_qpos"> overrides the MuJoCo qpos0
        init_q=(np.asarray(custom["init_qpos"], dtype=np.float64) if isinstance(custom.get("init_qpos"), np.ndarray)
                and len(custom["init_qpos"]) == len(init_q) else np.array(init_q)),
        custom={k: v for k, v in custom.items() if not isinstance(v, np.ndarray)},
        dof_ref=np.array(dof_ref),
    )
