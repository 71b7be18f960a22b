"""Turns an ncu --set full report of the rollout kernel into the committed profile summary.
    python scripts/ncu_summarize.py gpurun_out/prof.ncu-rep profiles/r01_rollout_v2_ncu.md "v2 kernel ..." [--json]
"""
import csv, io, json, subprocess, sys

rep, out_md, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.avg.per_cycle_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__icc_request_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg.per_second']
stall = [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h and 'pcsamp' not in h]
with open(out_md, "w") as f:
    f.write(f"# ncu --set full — {title}\n\nCommand: `ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 3 -c 1 "
            f"python scripts/gpu_ncu_target.py <env> <nsample> <variant>` (3 warm-up launches skipped) on 1x B200; workload as in the title (humanoidrun Nsample=8192 Hsample=50 unless stated) "
            f"(2,867,200 XPBD substeps per launch).\n\n| metric | value | unit |\n|---|---|---|\n")
    for k in keys:
        if k in d:
            f.write(f"| {k} | {d[k][0]} | {d[k][1]} |\n")
    f.write("\nWarp stall reasons (warps per issue-active cycle):\n\n| reason | value |\n|---|---|\n")
    for h in sorted(stall, key=lambda h: -float(d[h][0] or 0)):
        f.write(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {d[h][0]} |\n")


def to_bytes(v, u):
    v = float(v)
    return int(v * {"byte": 1, "Kbyte": 1024, "Mbyte": 1024 ** 2, "Gbyte": 1024 ** 3}.get(u, 1))


if "--json" in sys.argv:
    js = {"kernel": title.split(" ")[0], "title": title,
          "dram_bytes_per_launch": to_bytes(*d['dram__bytes_read.sum']) + to_bytes(*d['dram__bytes_write.sum']),
          "duration_ms": float(d['gpu__time_duration.sum'][0]), "warp_instructions": float(d['smsp__inst_executed.sum'][0]),
          "issue_active_pct": float(d['smsp__issue_active.avg.pct_of_peak_sustained_active'][0]),
          "registers_per_thread": int(float(d['launch__registers_per_thread'][0]))}
    json.dump(js, open("profiles/rollout_kernel_summary.json", "w"), indent=1)
    print(js)
