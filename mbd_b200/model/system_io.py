"""(De)serialise a compiled `System` as JSON so the package runs where the reference tree
(and its MJCF files) is absent, e.g. on the GPU box.  The JSON files under mbd_b200/assets/
are DERIVED DATA: produced by scripts/compile_assets.py from the reference's
mbd/assets/*.xml with this repo's own compiler (mbd_b200/model/mjcf.py)."""
from __future__ import annotations

import dataclasses
import json

import numpy as np

from .mjcf import Geom, System


def _enc(v):
    if isinstance(v, np.ndarray):
        return {"__nd__": v.dtype.str, "shape": list(v.shape), "data": np.where(np.isfinite(v), v, 0).ravel().tolist()
                if v.dtype.kind != "f" else [float(x) if np.isfinite(x) else ("inf" if x > 0 else "-inf") for x in v.ravel()]}
    if isinstance(v, Geom):
        return {"__geom__": {k: _enc(getattr(v, k)) for k in [f.name for f in dataclasses.fields(Geom)]}}
    if isinstance(v, (list, tuple)):
        return [_enc(x) for x in v]
    if isinstance(v, dict):
        return {k: _enc(x) for k, x in v.items()}
    if isinstance(v, (np.floating, np.integer, np.bool_)):
        return v.item()
    return v


def _dec(v):
    if isinstance(v, dict):
        if "__nd__" in v:
            data = [float(x) if isinstance(x, str) else x for x in v["data"]]
            return np.array(data, dtype=np.dtype(v["__nd__"])).reshape(v["shape"])
        if "__geom__" in v:
            return Geom(**{k: _dec(x) for k, x in v["__geom__"].items()})
        return {k: _dec(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_dec(x) for x in v]
    return v


def dumps(sys: System) -> str:
    return json.dumps({f.name: _enc(getattr(sys, f.name)) for f in dataclasses.fields(System)}, indent=0)


def loads(s: str) -> System:
    d = json.loads(s)
    return System(**{k: _dec(v) for k, v in d.items()})


def save(sys: System, path: str):
    with open(path, "w") as f:
        f.write(dumps(sys))


def load(path: str) -> System:
    with open(path) as f:
        return loads(f.read())
