for v in 0 1; do
  PSSGP_FUSED_REVERSE=$v python bench.py --no-extra --no-cpu-baseline 2>/dev/null > gpurun_out/fr_$v.json
done
python - <<'PY'
import json
for v in (0, 1):
    d = json.loads([l for l in open(f"gpurun_out/fr_{v}.json") if l.startswith("{")][-1])
    print("fused_reverse", v, round(d["ms_per_step"], 4), {k: round(x["avg_us"], 1) for k, x in d["kernels"].items()})
PY
