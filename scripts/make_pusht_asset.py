"""Derives mbd_b200/assets/pusht.json (model DATA: masses, sizes, joint and actuator parameters) from the reference's
pushT.xml (/root/reference/mbd/assets/pushT.xml, loaded by /root/reference/mbd/envs/pushT.py:18).  Run in a checkout that has
the reference tree; the JSON travels with the repo.

    python scripts/make_pusht_asset.py [path/to/pushT.xml]

Only what the planar generalized pipeline of include/mbd_pusht.h needs is extracted; MuJoCo defaults are filled in where the XML
is silent (solref 0.02 1, solimp 0.9 0.95 0.001 0.5 2, opt.iterations 100, geom friction 1 0.005 0.0001, contype/conaffinity 1)."""
import json
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def vec(s, n=None):
    v = [float(x) for x in s.split()]
    assert n is None or len(v) == n, (s, n)
    return v


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/mbd/assets/pushT.xml"
    root = ET.parse(src).getroot()
    opt = root.find("option")
    out = {"source": "derived from the reference's mbd/assets/pushT.xml by scripts/make_pusht_asset.py",
           "timestep": float(opt.get("timestep")), "gravity": vec(opt.get("gravity"), 3), "iterations": int(opt.get("iterations", 100)),
           "solref": [0.02, 1.0], "solimp": [0.9, 0.95, 0.001, 0.5, 2.0], "bodies": {}}
    assert all(g == 0.0 for g in out["gravity"]), "the planar pipeline assumes no gravity (pushT.xml: gravity 0 0 0)"
    for body in root.find("worldbody").findall("body"):
        name = body.get("name")
        assert vec(body.get("pos"), 3) == [0.0, 0.0, 0.0]
        joints = []
        for j in body.findall("joint"):
            joints.append({"name": j.get("name"), "type": j.get("type"), "axis": vec(j.get("axis"), 3),
                           "limited": j.get("limited", "false") == "true",
                           "range": vec(j.get("range"), 2) if j.get("range") else None,
                           "damping": float(j.get("damping", 0.0)), "stiffness": float(j.get("stiffness", 0.0)),
                           "armature": float(j.get("armature", 0.0))})
        geoms = []
        for g in body.findall("geom"):
            typ = g.get("type")
            size = vec(g.get("size"))
            pos = vec(g.get("pos", "0 0 0"), 3)
            if typ == "sphere":
                vol = 4.0 / 3.0 * np.pi * size[0] ** 3
            elif typ == "box":
                vol = 8.0 * size[0] * size[1] * size[2]
            else:
                raise NotImplementedError(typ)
            mass = float(g.get("mass")) if g.get("mass") else float(g.get("density", 1000.0)) * vol
            geoms.append({"name": g.get("name"), "type": typ, "size": size, "pos": pos, "mass": mass,
                          "friction": vec(g.get("friction", "1 0.005 0.0001"), 3)[0],
                          "collides": int(g.get("contype", 1)) != 0 or int(g.get("conaffinity", 1)) != 0,
                          "rgba": vec(g.get("rgba", "0.5 0.5 0.5 1"), 4)})
        # planar inertia about z through the compound COM
        m = sum(g["mass"] for g in geoms)
        com = [sum(g["mass"] * g["pos"][k] for g in geoms) / m for k in range(3)]
        izz = 0.0
        for g in geoms:
            if g["type"] == "sphere":
                i_own = 0.4 * g["mass"] * g["size"][0] ** 2
            else:
                i_own = g["mass"] / 12.0 * ((2 * g["size"][0]) ** 2 + (2 * g["size"][1]) ** 2)
            izz += i_own + g["mass"] * ((g["pos"][0] - com[0]) ** 2 + (g["pos"][1] - com[1]) ** 2)
        out["bodies"][name] = {"joints": joints, "geoms": geoms, "mass": m, "com": com, "izz": izz}
    table = root.find("worldbody").find("geom")
    out["table"] = {"pos": vec(table.get("pos"), 3), "size": vec(table.get("size"), 3), "rgba": vec(table.get("rgba"), 4)}
    out["actuators"] = [{"joint": a.get("joint"), "gear": vec(a.get("gear"))[0], "ctrlrange": vec(a.get("ctrlrange"), 2)}
                        for a in root.find("actuator").findall("motor")]
    dst = os.path.join(ROOT, "mbd_b200", "assets", "pusht.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", dst)
    for k, b in out["bodies"].items():
        print(k, "mass", b["mass"], "com", b["com"], "izz", b["izz"])


if __name__ == "__main__":
    main()
