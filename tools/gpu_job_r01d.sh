mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/test_all.log 2>&1; tail -4 gpurun_out/test_all.log; grep -E "^(FAILED|E  )" gpurun_out/test_all.log | cut -c1-250 | head -20
for mode in tf32x3 tf32 simt; do
  TTDG_CONV=$mode timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full_${mode}.json 2> gpurun_out/bench_err_${mode}.log
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_full_${mode}.json")); print("${mode}", d["value"], "img/s  ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "loss", d["e2e"]["last_loss"], "launches", d["gpu_launches"], "cpu", d["cpu_baseline"]["value"])
except Exception as e:
    print("${mode} FAILED", e); print(open("gpurun_out/bench_err_${mode}.log").read()[-1500:])
PY
done
