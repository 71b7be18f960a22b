"""Compiles the reference's model/demo assets into this package's derived-data files.

Run in the builder container (where /root/reference exists):
    python scripts/compile_assets.py [--ref /root/reference]
Outputs (committed):
    mbd_b200/assets/{humanoidrun,humanoidtrack,humanoidstandup,cartpole}.json   compiled System (mjcf.py)
    mbd_b200/assets/demos.npz   car2d_xref (50,2) f32; jog_xref (5,50,3) f32 built exactly as
                                /root/reference/mbd/envs/humanoidtrack.py:33-44 does
The pickles hold jax.Array objects; they are read with a JAX-free unpickler shim.
"""
import argparse
import io
import os
import pickle
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mbd_b200.model import mjcf, system_io  # noqa: E402


class _Shim(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("jax") and name == "_reconstruct_array":
            def rec(fun, args, arr_state, aval_state):
                a = fun(*args)
                a.__setstate__(arr_state)
                return np.asarray(a)
            return rec
        if module.startswith("numpy.core"):
            module = module.replace("numpy.core", "numpy._core")
        return super().find_class(module, name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    src = os.path.join(a.ref, "mbd", "assets")
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mbd_b200", "assets")
    os.makedirs(dst, exist_ok=True)
    for name in ("humanoidrun", "humanoidtrack", "humanoidstandup", "cartpole"):
        s = mjcf.load(os.path.join(src, name + ".xml"))
        system_io.save(s, os.path.join(dst, name + ".json"))
        print(name, s.link_types, s.num_links(), "links")
    car = np.load(os.path.join(src, "car2d_xref.npy")).astype(np.float32)
    with open(os.path.join(src, "jog_xref.pkl"), "rb") as f:
        d = _Shim(io.BytesIO(f.read())).load()
    H = 50
    names = ["torso", "left_thigh", "right_thigh", "left_shin", "right_shin"]
    xref = []
    for n in names:
        x = np.asarray(d[n], dtype=np.float32)
        if len(x) < H:
            x = np.concatenate([x, np.tile(x[-1:], (H - len(x), 1))], axis=0)
        else:
            x = x[70:H + 70]
        xref.append(x)
    jog = np.stack(xref, axis=0).astype(np.float32)
    np.savez(os.path.join(dst, "demos.npz"), car2d_xref=car, jog_xref=jog)
    print("car2d_xref", car.shape, "jog_xref", jog.shape)


if __name__ == "__main__":
    main()
