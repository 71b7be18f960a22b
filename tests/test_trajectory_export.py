"""The rollout export in the structure the reference's tools consume: `render_us` (mbd/utils.py:23-34,
mbd_planner.py:168-178) returns the Brax-visualizer page, `brax_json.dumps` the JSON document `brax.io.json.dumps` /
`scripts/vis_diffusion.py:27-112` build, `BraxLikeSystem` the geom arrays vis_diffusion walks.  The CPU test builds the
states with the host kinematics; the GPU test steps the real env."""
import json
import os

import numpy as np
import pytest

import mbd_b200
from mbd_b200 import prng
from mbd_b200.io import brax_json
from mbd_b200.model import kinematics
from mbd_b200.utils import render_us, rollout_states, trajectory_arrays


def _check_document(env, doc, T):
    L = env.sys.num_links()
    assert doc["opt"]["timestep"] == pytest.approx(env.dt) and doc["link_names"] == list(env.sys.link_names)
    pos, rot = np.array(doc["states"]["x"]["pos"]), np.array(doc["states"]["x"]["rot"])
    assert pos.shape == (T, L, 3) and rot.shape == (T, L, 4) and np.allclose(np.linalg.norm(rot, axis=-1), 1.0, atol=1e-4)
    # geoms are keyed by link name (+ "world" for the floor), each with the fields the viewer reads
    assert set(doc["geoms"]) <= set(env.sys.link_names) | {"world"} and "world" in doc["geoms"]
    assert doc["geoms"]["world"][0]["name"] == "Plane" and doc["geoms"]["world"][0]["link_idx"] == -1
    n = 0
    for name, geoms in doc["geoms"].items():
        for g in geoms:
            n += 1
            assert set(g) == {"name", "link_idx", "pos", "rot", "rgba", "size"} and g["name"] in ("Plane", "Sphere", "Capsule")
            assert g["link_idx"] == (-1 if name == "world" else env.sys.link_names.index(name))
            assert len(g["pos"]) == 3 and len(g["rot"]) == 4 and len(g["rgba"]) == 4 and len(g["size"]) == 3
    assert n == len(env.sys.geoms)


def test_brax_json_document_from_host_kinematics():
    env = mbd_b200.envs.get_env("humanoidrun")
    rng = np.random.default_rng(0)
    states = []
    for _ in range(4):
        q = env.sys.init_q.copy(); q[7:] += rng.uniform(-0.2, 0.2, q.size - 7)
        states.append(env.pipeline_init(q, np.zeros(env.sys.qd_size())))
    doc = json.loads(brax_json.dumps(env.sys, states, env.dt))
    _check_document(env, doc, 4)
    # the torso capsule of the vendored XML: fromto 0 -.07 0 0 .07 0, size .07 -> centre 0, half length .07
    torso = doc["geoms"]["torso"][0]
    assert torso["name"] == "Capsule" and np.allclose(torso["size"][:2], [0.07, 0.07], atol=1e-6)
    bs = brax_json.BraxLikeSystem(env.sys, env.dt)   # what vis_diffusion.py:37-47 walks
    assert bs.ngeom == len(env.sys.geoms) and bs.geom_bodyid.min() == 0 and bs.geom_pos.shape == (bs.ngeom, 3)
    page = brax_json.render(env.sys, states, env.dt)
    assert page.startswith("<html>") and "var system = {" in page and "viewer.js" in page


@pytest.mark.gpu
def test_render_us_on_the_gpu_stepped_env(tmp_path):
    """the real env surface (GPU kernel behind env.step): render_us returns the viewer page; its states are the pipeline
    states BEFORE each step, initial state first, and joint coordinates agree with the world poses at reset"""
    env = mbd_b200.envs.get_env("humanoidrun")
    state = env.reset(prng.split(prng.PRNGKey(0))[1])
    us = np.clip(np.random.default_rng(0).normal(size=(6, env.action_size)) * 0.5, -1, 1).astype(np.float32)
    page = render_us(env.step, env.sys, state, us)
    doc = json.loads(page[page.index("var system = ") + len("var system = "):page.index(";</script>")])
    _check_document(env, doc, 6)
    assert np.allclose(doc["states"]["x"]["pos"][0], np.asarray(state.pipeline_state.x.pos), atol=1e-5)
    assert not np.allclose(doc["states"]["x"]["pos"][5][0], doc["states"]["x"]["pos"][0][0])
    tr = trajectory_arrays(env, rollout_states(env.step, state, us))
    pos = kinematics.forward(env.sys, tr["q"][0].astype(np.float64), tr["qd"][0].astype(np.float64))[0]
    assert np.allclose(pos, tr["pos"][0], atol=1e-5)


@pytest.mark.gpu
def test_cli_writes_the_reference_artefacts(tmp_path, monkeypatch):
    """mbd_planner.py:152-178: mu_0ts.npy (Ndiffuse-1, H, Nu) + rollout.html (+ rollout.json / rollout_states.npz)"""
    from mbd_b200.planners.mbd_planner import Args, run_diffusion
    run_diffusion(Args(env_name="hopper", Nsample=256, Hsample=10, Ndiffuse=5))
    path = os.path.join(os.path.dirname(mbd_b200.__path__[0]), "results", "hopper")
    assert np.load(os.path.join(path, "mu_0ts.npy")).shape == (4, 10, 3)
    doc = json.load(open(os.path.join(path, "rollout.json")))
    assert np.array(doc["states"]["x"]["pos"]).shape == (10, 4, 3) and open(os.path.join(path, "rollout.html")).read().startswith("<html>")
