"""Deterministic random MJCF models inside the positional-env subset (test infrastructure, NOT reference assets).

Used by the differential tests: every model a seed produces goes through the same compiler (`mbd_b200.model`), the CPU oracle
and the CUDA kernels, and the two must agree bit for bit — topology handling (children order, up to 4 children, several
world-parented trees, 1..16 links), stacked hinges with arbitrary orthonormal axes, planar / sliding / free roots and 0..6 plane
contacts per link are all drawn at random instead of being the handful of shapes the fixed envs happen to have.
"""
from __future__ import annotations

import numpy as np

MAXL, MAXCHILD, MAXDOF, MAXCON = 16, 4, 3, 6


def _fmt(v):
    return " ".join(f"{float(x):.9g}" for x in np.atleast_1d(v))


def _frame(rng):
    """a random right-handed orthonormal frame (columns)"""
    a = rng.normal(size=(3, 3))
    q, _ = np.linalg.qr(a)
    if np.linalg.det(q) < 0:
        q[:, 2] = -q[:, 2]
    return q


def random_model(seed: int, max_links: int = MAXL, links: int = 0, topology: str = ""):
    """Returns (xml_text, facts) for seed; facts = dict(L, nu, roots, ncon, max_children).
    topology: "" random tree | "chain" (every link has one child: the deepest tree, L - 1 links with children) |
    "lonely" (the last link is a second, single-link free body that collides: a leaf WITHOUT a joint to a parent)."""
    rng = np.random.default_rng(seed)
    L = int(rng.integers(1, max_links + 1))
    if links:
        L = int(links)      # fixed link count (11 = the shape the specialised humanoid kernels are instantiated for)
    # ---- topology: parent[i] < i, at most MAXCHILD children, one or two world-parented trees
    parent = [-1]
    nchild = [0] * L
    second_root = L >= 4 and rng.random() < 0.25
    if topology:
        second_root = False
    for i in range(1, L):
        if topology == "chain":
            parent.append(i - 1); nchild[i - 1] += 1
            continue
        if topology == "lonely" and i == L - 1:
            parent.append(-1)
            continue
        if second_root and i == L - 1 - int(rng.integers(0, max(1, L // 3))):
            parent.append(-1)
            second_root = False
            continue
        cand = [p for p in range(i) if nchild[p] < MAXCHILD]
        # prefer recent bodies (chains) but allow bushy trees
        w = np.array([1.0 + 2.0 * (p == i - 1) + 1.0 * (nchild[p] > 0) for p in cand])
        p = int(rng.choice(cand, p=w / w.sum()))
        parent.append(p)
        nchild[p] += 1
    depth = [0] * L
    for i in range(1, L):
        depth[i] = 0 if parent[i] < 0 else depth[parent[i]] + 1
    # ---- bodies
    bodies = []
    act = []
    ncon_total = 0
    for i in range(L):
        b = {"name": f"b{i}", "joints": [], "geoms": []}
        if parent[i] < 0:
            kind = rng.choice(["free", "planar", "slide", "hinge_only"], p=[0.45, 0.3, 0.15, 0.1])
            if topology == "lonely" and i == L - 1:
                kind = "free"
            b["pos"] = np.array([0.6 * i * (i > 0), 0.0, 0.45 + 0.22 * max(depth) + 0.3 * rng.random()])
            if kind == "free":
                b["joints"].append(dict(type="free", name=f"j{i}_free"))
            elif kind == "planar":   # hopper / walker2d / halfcheetah style root
                b["joints"].append(dict(type="slide", name=f"j{i}_x", axis="1 0 0", damping=_fmt(rng.choice([0.0, 0.5]))))
                b["joints"].append(dict(type="slide", name=f"j{i}_z", axis="0 0 1", ref=_fmt(b["pos"][2]) if rng.random() < 0.5 else None))
                b["joints"].append(dict(type="hinge", name=f"j{i}_y", axis="0 1 0"))
            elif kind == "slide":    # cartpole style cart, optionally limited and actuated
                lim = rng.random() < 0.6
                sax = str(rng.choice(["1 0 0", "0 1 0"]))
                b["joints"].append(dict(type="slide", name=f"j{i}_s", axis=sax,
                                        limited="true" if lim else "false", range="-0.8 0.8" if lim else None,
                                        stiffness=_fmt(rng.choice([0.0, 3.0])), damping=_fmt(rng.choice([0.0, 1.0]))))
                if rng.random() < 0.5:
                    b["joints"].append(dict(type="hinge", name=f"j{i}_h", axis="0 1 0" if sax == "1 0 0" else "1 0 0"))
                if rng.random() < 0.8:
                    act.append((f"j{i}_s", float(rng.uniform(5, 40))))
            else:
                F = _frame(rng)
                for k in range(int(rng.integers(1, MAXDOF + 1))):
                    b["joints"].append(dict(type="hinge", name=f"j{i}_{k}", axis=_fmt(F[:, k]), limited="true",
                                            range=_fmt([-rng.uniform(10, 80), rng.uniform(10, 80)]),
                                            stiffness=_fmt(rng.choice([0.0, 4.0])), damping=_fmt(rng.uniform(0, 2))))
        else:
            d = rng.normal(size=3)
            d[2] = -abs(d[2]) * 0.7
            d = d / np.linalg.norm(d) * rng.uniform(0.12, 0.22)
            b["pos"] = d
            F = _frame(rng) if rng.random() < 0.6 else np.eye(3)[:, rng.permutation(3)]
            for k in range(int(rng.choice([1, 2, 3], p=[0.5, 0.3, 0.2]))):
                lo, hi = -rng.uniform(5, 70), rng.uniform(5, 70)
                if rng.random() < 0.15:       # a range that does not contain zero (the joint starts outside its limits)
                    lo, hi = 10.0, 10.0 + rng.uniform(10, 50)
                b["joints"].append(dict(type="hinge", name=f"j{i}_{k}", axis=_fmt(F[:, k]), limited="true" if rng.random() < 0.9 else "false",
                                        range=_fmt([lo, hi]), stiffness=_fmt(rng.choice([0.0, 2.0, 8.0])),
                                        damping=_fmt(rng.uniform(0, 2)), armature=_fmt(rng.choice([0.0, 0.01]))))
                if rng.random() < 0.75:
                    act.append((f"j{i}_{k}", float(rng.uniform(10, 80))))
        # geoms: mass always comes from geoms; contacts from the ones with contype 1
        ng = int(rng.choice([1, 2, 3], p=[0.6, 0.3, 0.1]))
        con_here = 0
        for g in range(ng):
            collide = rng.random() < 0.45 or (topology == "lonely" and i == L - 1)
            if rng.random() < 0.55:
                e = rng.normal(size=3)
                e = e / np.linalg.norm(e) * rng.uniform(0.08, 0.2)
                need = 2
                geom = dict(type="capsule", fromto=_fmt(np.concatenate([np.zeros(3), e])), size=_fmt(rng.uniform(0.03, 0.07)))
            else:
                need = 1
                geom = dict(type="sphere", pos=_fmt(rng.normal(size=3) * 0.05), size=_fmt(rng.uniform(0.04, 0.1)))
            if collide and con_here + need > MAXCON:
                collide = False
            con_here += need if collide else 0
            geom.update(name=f"g{i}_{g}", contype="1" if collide else "0", density=_fmt(rng.uniform(200, 1200)),
                        friction=_fmt([rng.choice([0.5, 1.0, 1.5]), 0.1, 0.1]))
            b["geoms"].append(geom)
        ncon_total += con_here
        bodies.append(b)
    if not act:   # at least one control input: put a motor on the first hinge / slide that exists
        for b in bodies:
            js = [j for j in b["joints"] if j["type"] != "free"]
            if js:
                act.append((js[0]["name"], 20.0))
                break
    # ---- XML
    def attrs(d):
        return " ".join(f'{k}="{v}"' for k, v in d.items() if v is not None)

    def emit(i, ind):
        b = bodies[i]
        pad = "  " * ind
        out = [f'{pad}<body name="{b["name"]}" pos="{_fmt(b["pos"])}">']
        for j in b["joints"]:
            out.append(f"{pad}  <joint {attrs(j)}/>")
        for g in b["geoms"]:
            out.append(f"{pad}  <geom {attrs(g)}/>")
        for c in range(L):
            if parent[c] == i:
                out.extend(emit(c, ind + 1))
        out.append(f"{pad}</body>")
        return out

    body_xml = []
    for i in range(L):
        if parent[i] < 0:
            body_xml.extend(emit(i, 2))
    cad = rng.choice([0.0, 20.0, 50.0])
    xml = "\n".join([
        f'<!-- generated by tests/modelgen.py, seed {seed}: test infrastructure, not a reference asset -->',
        f'<mujoco model="fuzz_{seed}">',
        '  <compiler angle="degree" inertiafromgeom="true"/>',
        f'  <option timestep="{_fmt(rng.choice([0.003, 0.005, 0.008]))}"/>',
        "  <custom>",
        f'    <numeric data="{_fmt(cad)}" name="constraint_ang_damping"/>',
        f'    <numeric data="{_fmt(rng.choice([0.0, -0.05]))}" name="ang_damping"/>',
        f'    <numeric data="{_fmt(rng.choice([0.5, 0.3]))}" name="joint_scale_pos"/>',
        f'    <numeric data="{_fmt(rng.choice([0.2, 0.1]))}" name="joint_scale_ang"/>',
        '    <numeric data="0" name="spring_mass_scale"/>',
        '    <numeric data="1" name="spring_inertia_scale"/>',
        "  </custom>",
        "  <default>",
        '    <joint armature="0" damping="0"/>',
        '    <geom conaffinity="0" contype="0"/>',
        '    <motor ctrllimited="true" ctrlrange="-1 1"/>',
        "  </default>",
        "  <worldbody>",
        '    <geom conaffinity="1" contype="0" name="floor" pos="0 0 0" size="40 40 1" type="plane" friction="1 0.5 0.5"/>',
        *body_xml,
        "  </worldbody>",
        "  <actuator>",
        *[f'    <motor gear="{_fmt(g)}" joint="{j}"/>' for j, g in act],
        "  </actuator>",
        "</mujoco>",
        "",
    ])
    facts = dict(L=L, nu=len(act), roots=sum(1 for p in parent if p < 0), ncon=ncon_total, max_children=max(nchild) if L else 0,
                 parent=parent)
    return xml, facts
