python -m pytest tests -m gpu -x -q -k "msm or prove" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --cpu-steps 0 2>/dev/null > gpurun_out/bench.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("c2", d["ms_per_step"], d["e2e"]["ms_per_step"], "g1 launch_ms", d["roofline"]["launch_ms"], "frac", d["roofline"]["int_pipe"]["frac"])
for k in ("msm_accum_g1","msm_accum_g2","msm_reduce","ntt_pass","msm_sort"):
    print(" ", k, d["phases"][k]["ms_per_step"])
PY
