#!/usr/bin/env python
"""bench.py — env-steps/s of the MBD reverse-diffusion step (BASELINE.json metric).

A "step" is ONE diffusion step `reverse_once` (/root/reference/mbd/planners/mbd_planner.py:97-135)
on humanoidrun: Nsample x Hsample env steps (x n_frames=7 XPBD substeps) + reward statistics +
softmax weighted mean + update.  value = Nsample*Hsample / t_step, whole job over all ranks.

  python bench.py [--gpus N --steps K --warmup W]        # under torchrun for N > 1
  python bench.py --impl reference ...                   # the CPU restatement (oracle) arm

Timing: per-step CUDA events on the launching (current) stream, >= 3 warm-ups, L2 flushed
(256 MiB memset, untimed) between steps, barrier + synchronize on both sides, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENV_NAME, NSAMPLE, HSAMPLE, NDIFFUSE, TEMP = "humanoidrun", 8192, 50, 300, 0.1
NU, NFRAMES = 17, 7
BYTES_PER_ENV_STEP = 4 * NU + 4  # action row read once + reward written once = 72 B
METRIC = "env-steps/sec (Nsample x Hsample per diffusion step) on humanoidrun"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_summary():
    """numbers of the rollout kernel from the committed ncu --set full capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "rollout_kernel_summary.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:  # noqa: BLE001
            return {}
    return {}


def _ncu_traffic():
    return _ncu_summary().get("dram_bytes_per_launch")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 6 and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def _oracle_setup():
    import mbd_b200
    from mbd_b200 import prng
    from oracle import oracle as orc
    from oracle import planner as opl
    env = mbd_b200.envs.get_env(ENV_NAME)
    rng, rng_reset = prng.split(prng.PRNGKey(0))
    st = env.reset(rng_reset).pipeline_state.raw
    rng_exp, _ = prng.split(rng)
    return env, st, rng_exp, orc, opl


def time_cpu_oracle(n_samples: int, steps: int, warmup: int):
    """The CPU restatement of the same diffusion step (sampling + rollouts + statistics) on all
    host threads, on a bounded sample of n_samples of the Nsample rollouts per step."""
    env, st, key, orc, opl = _oracle_setup()
    from mbd_b200 import prng
    _, alphas, alphas_bar, sigmas = opl.make_schedule(1e-4, 1e-2, NDIFFUSE)
    oenv = opl.OracleEnv("xpbd", NU, blob=env.blob, state=st)
    threads = orc.num_threads()
    Yb = np.zeros(HSAMPLE * NU, np.float32)
    times = []
    i = NDIFFUSE - 1
    for it in range(warmup + steps):
        key, k = prng.split(key)
        t0 = time.perf_counter()
        o = opl.reverse_once(oenv, k, n_samples, HSAMPLE, float(sigmas[i]), Yb, TEMP, alphas, alphas_bar, i)
        dt = time.perf_counter() - t0
        Yb = o["Ybar_im1"]
        i -= 1
        if it >= warmup:
            times.append(dt)
    t = float(np.mean(times))
    return n_samples * HSAMPLE / t, t, threads


def run_reference(args):
    """--impl reference: the reference's CPU path.  JAX/Brax are not installable here (no network, no
    wheels), so this is the CPU restatement (oracle port, OpenMP over samples), labelled as such."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_samples
    val, t, threads = time_cpu_oracle(n, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{ENV_NAME} Nsample={NSAMPLE} Hsample={HSAMPLE} n_frames={NFRAMES} Ndiffuse={NDIFFUSE}",
                   "note": "CPU restatement (JAX/Brax unavailable): C oracle, OpenMP over samples"},
        "cpu_baseline": {"value": val, "unit": "env-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{n} of {NSAMPLE} rollouts x {HSAMPLE} env steps per step, {args.steps} steps"},
        "e2e": {"value": val, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import mbd_b200
    from mbd_b200 import prng
    from mbd_b200.planners import engine as eng

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_total = NSAMPLE * world if args.scaling == "weak" else NSAMPLE
    env = mbd_b200.envs.get_env(ENV_NAME)
    rng, rng_reset = prng.split(prng.PRNGKey(0))
    state_init = env.reset(rng_reset)
    _, alphas, alphas_bar, sigmas = eng.make_schedule(1e-4, 1e-2, NDIFFUSE)
    e = eng.DiffusionEngine(env, n_total, HSAMPLE, TEMP, False, state_init)
    HNu = HSAMPLE * NU
    rng_exp, _ = prng.split(rng)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def chain(nsteps, timed, host_io):
        """runs nsteps diffusion steps of the real chain starting at i = Ndiffuse-1; returns summed ms"""
        key = rng_exp
        Ybar = torch.zeros(HNu, device=dev)
        out = torch.empty(HNu, device=dev)
        h_in = torch.zeros(HNu, dtype=torch.float32).pin_memory()
        h_out = torch.zeros(HNu + 1, dtype=torch.float32).pin_memory()
        tot, kern = 0.0, 0.0
        i = NDIFFUSE - 1
        for _ in range(nsteps):
            key, k = prng.split(key)
            coef = eng.update_coef(alphas, alphas_bar, i)
            flush.fill_(1)
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            if host_io:  # the reference-facing call with HOST buffers: H2D of the iterate, D2H of result + reward
                Ybar.copy_(h_in, non_blocking=True)
            if e.single_kernel:      # one cooperative kernel does the whole step
                e.reverse_once(k, float(sigmas[i]), Ybar, coef, out=out)
                e2.record()
            else:
                e.rollout_phase(k, float(sigmas[i]), Ybar)
                e2.record()
                e.gather_phase()
                e.reduce_phase(Ybar, coef, out)
            if host_io:
                h_out[:HNu].copy_(out, non_blocking=True)
                h_out[HNu:].copy_(e.scalars[:1], non_blocking=True)
            e1.record()
            e1.synchronize()
            if host_io:
                h_in.copy_(h_out[:HNu])
            Ybar, out = out, Ybar
            if timed:
                tot += e0.elapsed_time(e1)
                kern += e0.elapsed_time(e2)
            i -= 1
        return tot, kern

    chain(args.warmup, False, False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    tot_ms, kern_ms = chain(args.steps, True, False)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    chain(1, False, True)
    barrier()
    e2e_ms, _ = chain(args.steps, True, True)
    barrier()
    t = torch.tensor([tot_ms, kern_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot_ms, kern_ms, e2e_ms = (float(v) for v in t.tolist())
    if rank == 0:
        ms_step = tot_ms / args.steps
        value = n_total * HSAMPLE / (ms_step * 1e-3)
        e2e_val = n_total * HSAMPLE / (e2e_ms / args.steps * 1e-3)
        kern_s = kern_ms / args.steps * 1e-3
        alg_bytes = e.n_local * HSAMPLE * BYTES_PER_ENV_STEP  # per launch of the rollout kernel (per rank)
        peak, peak_src = _peaks()
        achieved = alg_bytes / kern_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{ENV_NAME} Nsample={n_total} Hsample={HSAMPLE} n_frames={NFRAMES} Ndiffuse={NDIFFUSE} "
                                   f"(steps i={NDIFFUSE - 1}..{NDIFFUSE - args.steps} of the real chain, seed 0)",
                       "global_samples": n_total, "samples_per_gpu": e.n_local, "parallelism": f"sample-shard x{world}", "exchange": e.exchange,
                       "l2": "flushed between steps (256 MiB memset, untimed)",
                       "substeps_per_s": value * NFRAMES},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "env-steps/s", "h2d_bytes_per_step": HNu * 4 + 8, "d2h_bytes_per_step": HNu * 4 + 4},
            "gpu_launches": e.launches_last_step * args.steps,
            "roofline": {"bound": "hbm", "kernel": ("k_reverse_step_wpl (sampling + rollouts + statistics + weighted mean + update, one launch)"
                                                   if e.single_kernel else f"{_ncu_summary().get('kernel', 'k_rollout_wpl<true,...>')} (fused sampling + rollouts)"), "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": _ncu_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kern_s * 1e3,
                         "note": "path is fp32-issue bound, not HBM bound (SURVEY F7): see DESIGN.md roofline section"},
        }
        # the roofline that actually binds: warp-instruction issue slots.
        # instructions per launch come from the committed ncu capture of the same kernel/workload.
        wi = _ncu_summary().get("warp_instructions")
        if wi and e.n_local == NSAMPLE:
            clk = (clocks or {}).get("sm_mhz") or 1965.0
            peak_issue = 148 * 4 * clk * 1e6  # 1 warp-instruction per SM sub-partition per cycle
            line["roofline_issue"] = {"bound": "fp32 issue slots", "achieved": wi / kern_s, "peak": peak_issue, "unit": "warp-inst/s",
                                      "frac": wi / kern_s / peak_issue, "warp_instructions_per_launch": wi,
                                      "source": "profiles/rollout_kernel_summary.json (ncu smsp__inst_executed.sum)"}
        # fp32 roofline from the ALGORITHMIC flop count (9336 per sample and XPBD substep on humanoidrun, counted by running
        # the physics with an operation-counting scalar type: tests/test_pk_host.py::test_algorithmic_operation_count)
        try:
            clk_mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
            flop = float(e.n_local) * HSAMPLE * NFRAMES * 9336.0
            peak_tf = 148 * 128 * 2 * clk_mhz * 1e6 / 1e12   # 128 fp32 lanes per SM, FMA = 2 flop
            line["roofline_fp32"] = {"bound": "fp32 pipe", "achieved": flop / kern_s / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                                     "frac": flop / kern_s / 1e12 / peak_tf, "flop_per_sample_substep": 9336,
                                     "note": "algorithmic flops (mul, add 1; fma 2; div, rcp, sqrt 1); the device executes ~15 % more"}
        except Exception:  # noqa: BLE001 - never lose the bench line over an explanatory field
            pass
        if world == 1 and not args.no_cpu_baseline:
            val, tcpu, threads = time_cpu_oracle(args.cpu_samples, 3, 1)
            line["cpu_baseline"] = {"value": val, "unit": "env-steps/s", "cores": threads, "kind": "port",
                                    "sample": f"{args.cpu_samples} of {NSAMPLE} rollouts x {HSAMPLE} env steps per step, 3 steps "
                                              f"(CPU restatement; JAX/Brax unavailable)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-samples", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.steps + args.warmup >= NDIFFUSE:
        raise SystemExit("steps + warmup must be < Ndiffuse")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
