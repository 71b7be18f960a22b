"""Args / recommended-parameter override logic (mbd_planner.py:18-35,45-69) and the env registry."""
import dataclasses
import os

import pytest

import mbd_b200
from mbd_b200.planners.mbd_planner import Args, apply_recommended_params


def test_args_fields_match_reference():
    names = [f.name for f in dataclasses.fields(Args)]
    assert names == ["seed", "disable_recommended_params", "not_render", "env_name", "Nsample", "Hsample", "Ndiffuse",
                     "temp_sample", "beta0", "betaT", "enable_demo"]
    a = Args()
    assert (a.seed, a.env_name, a.Nsample, a.Hsample, a.Ndiffuse, a.temp_sample, a.beta0, a.betaT, a.enable_demo) == \
        (0, "ant", 2048, 50, 100, 0.1, 1e-4, 1e-2, False)


def test_recommended_overrides(capsys):
    a = apply_recommended_params(Args(env_name="humanoidrun", Nsample=16, Ndiffuse=5, temp_sample=0.7))
    assert (a.Nsample, a.Ndiffuse, a.temp_sample, a.Hsample) == (8192, 300, 0.1, 50)
    assert "override temp_sample to 0.1" in capsys.readouterr().out
    a = apply_recommended_params(Args(env_name="pushT"))
    assert (a.Ndiffuse, a.Hsample, a.temp_sample) == (200, 40, 0.2)
    a = apply_recommended_params(Args(env_name="halfcheetah"))
    assert a.temp_sample == 0.4
    a = apply_recommended_params(Args(env_name="car2d", temp_sample=0.3, Nsample=64))
    assert (a.temp_sample, a.Nsample, a.Ndiffuse) == (0.3, 64, 100)  # car2d has no recommendation
    a = apply_recommended_params(Args(env_name="humanoidrun", Nsample=16, disable_recommended_params=True))
    assert a.Nsample == 16 and a.Ndiffuse == 100


def test_tyro_cli_parses_reference_flags():
    import tyro
    a = tyro.cli(Args, args=["--env_name", "car2d", "--Nsample", "64", "--Hsample", "40", "--enable_demo", "--not_render"])
    assert a.env_name == "car2d" and a.Nsample == 64 and a.Hsample == 40 and a.enable_demo and a.not_render


def test_env_registry():
    assert mbd_b200.envs.get_env("car2d").action_size == 2
    e = mbd_b200.envs.get_env("humanoidrun")
    assert (e.action_size, e.observation_size, round(e.dt, 6)) == (17, 47, 0.042)
    t = mbd_b200.envs.get_env("humanoidtrack")
    assert (t.action_size, round(t.dt, 6), t.xref.shape, t.rew_xref) == (17, 0.03, (5, 50, 3), 1.0)
    assert t.track_body_idx.tolist() == [0, 5, 3, 6, 4] and t.ref_body_idx.tolist() == [11, 12, 13, 14, 15]
    hs = mbd_b200.envs.get_env("humanoidstandup")
    assert (hs.action_size, hs.observation_size, round(hs.dt, 6)) == (17, 47, 0.042)
    with pytest.raises(ValueError, match="Unknown environment"):
        mbd_b200.envs.get_env("nope")
    pt = mbd_b200.envs.get_env("pushT")   # the one `generalized`-backend env (tests/test_pusht.py)
    assert (pt.action_size, pt.observation_size, round(pt.dt, 6), pt.backend) == (2, 16, 0.05, "generalized")


def test_reset_is_the_reference_chain():
    """humanoidrun.py:19-32: split(rng,3) -> uniform(+-0.01) on q (24) and qd (23)."""
    import numpy as np
    from mbd_b200 import prng
    e = mbd_b200.envs.get_env("humanoidrun")
    rng, rng_reset = prng.split(prng.PRNGKey(0))
    st = e.reset(rng_reset)
    _, r1, r2 = prng.split(rng_reset, 3)
    q = e.sys.init_q.astype(np.float32) + prng.uniform(r1, (24,), -0.01, 0.01)
    assert np.allclose(st.pipeline_state.q[7:], q[7:], atol=1e-6)       # hinge angles survive FK -> IK
    assert np.allclose(st.pipeline_state.q[:3], q[:3], atol=1e-6)
    assert st.obs.shape == (47,) and st.reward == 0 and st.done == 0
    st2 = mbd_b200.envs.get_env("humanoidtrack").reset(None)
    assert np.allclose(st2.pipeline_state.q[7:24], 0, atol=1e-7) and np.allclose(st2.pipeline_state.qd, 0)


def test_schedule_upload_helpers():
    """host pieces of the graph-captured solve: the Y0s_rng chain of a whole solve equals the per-step `rng, Y0s_rng = split(rng)`
    of mbd_planner.py:103; the float32 linspace keeps the endpoints exactly and the schedule KATs of SURVEY 8(d)"""
    import numpy as np
    from mbd_b200 import prng
    from mbd_b200.planners import engine as eng
    rng_exp = prng.split(prng.split(prng.PRNGKey(0))[0])[0]
    keys = eng.key_chain(rng_exp, 12)
    r = rng_exp
    for i in range(11, 0, -1):
        r, k = prng.split(r)
        assert np.array_equal(keys[i], k)
    assert not keys[0].any()
    for N, sig in ((100, 0.6305), (200, 0.7981), (300, 0.8839)):
        betas, alphas, alphas_bar, sigmas = eng.make_schedule(1e-4, 1e-2, N)
        assert betas.dtype == np.float32 and betas[0] == np.float32(1e-4) and betas[-1] == np.float32(1e-2) and abs(float(sigmas[-1]) - sig) < 5e-5
    c = eng.update_coef(alphas, alphas_bar, 5)
    assert all(np.asarray(v).dtype == np.float32 for v in c) and len(c) == 5


def test_brax_asset_lookup_order(tmp_path, monkeypatch):
    """envs/base.py::brax_asset: $MBD_BRAX_ASSETS first (a user's own Brax files), then the repo's restated models"""
    from mbd_b200.envs.base import ASSET_DIR, brax_asset
    monkeypatch.delenv("MBD_BRAX_ASSETS", raising=False)
    assert os.path.samefile(brax_asset("hopper.xml"), os.path.join(ASSET_DIR, "hopper.xml"))
    (tmp_path / "hopper.xml").write_text(open(os.path.join(ASSET_DIR, "hopper.xml")).read())
    monkeypatch.setenv("MBD_BRAX_ASSETS", str(tmp_path))
    assert os.path.samefile(brax_asset("hopper.xml"), tmp_path / "hopper.xml")
    with pytest.raises(FileNotFoundError):
        brax_asset("nope.xml")


def test_bench_reference_arm_contract(tmp_path):
    """`bench.py --impl reference` (the driver's CPU arm, no GPU needed): one JSON line with the contract's keys, the same
    `config.workload` string as the GPU arm, threads taken from the cgroup / affinity (not from torchrun's OMP_NUM_THREADS=1)"""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3", "--cpu-samples", "128"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "env-steps/s" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["config"]["workload"] == "humanoidrun Nsample=8192 Hsample=50 n_frames=7 Ndiffuse=300"
    assert line["cpu_baseline"]["cores"] >= 1 and line["e2e"]["h2d_bytes_per_step"] == 0
    # other ranks of a torchrun launch exit silently
    r2 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1"], capture_output=True, text=True,
                        timeout=120, env=dict(env, RANK="1", WORLD_SIZE="2"))
    assert r2.returncode == 0 and r2.stdout.strip() == ""
