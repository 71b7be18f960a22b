"""Follow-up of scripts/gpu_round2_experiments.py: shared-scheduler group map (B) with a finer start-offset sweep."""
import os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
for mp in ("A", "B"):
    for st in (0, 2000, 3500, 4000, 4500, 5000, 5500, 6000, 7000, 9000):
        try:
            subprocess.run([sys.executable, os.path.join(HERE, "gpu_round2_experiments.py"), "6", mp, str(st)], timeout=60, check=False)
        except subprocess.TimeoutExpired:
            print(f"map={mp} stagger={st}: TIMEOUT", flush=True)
