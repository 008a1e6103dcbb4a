"""Prints the per-kernel CUDA-event times of a short bench run (tuning helper): python scripts/bench_kernels.py"""
import json, subprocess, sys, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu-baseline"],
                     capture_output=True, text=True)
line = [l for l in out.stdout.splitlines() if l.startswith("{")]
if not line:
    print(out.stdout[-2000:], out.stderr[-2000:]); sys.exit(1)
d = json.loads(line[-1])
print(os.environ.get("PSSGP_B200_LIB", "default"), "ms/step", round(d["ms_per_step"], 4), "frac", round(d["step_roofline"]["frac_of_hbm_peak"], 4),
      " ".join(f"{k}={v['avg_us']:.1f}" for k, v in d["kernels"].items()))
