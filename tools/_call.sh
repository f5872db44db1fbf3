echo "=== bench pairs"; B200ASR_CONV_WGRAD=bf16x3 timeout 400 python bench.py --no-cpu --no-ref-gpu --steps 10 2>&1 | tail -1 > gpurun_out/r2_bench_pairs.json
echo "=== bench default"; timeout 400 python bench.py --no-cpu --no-ref-gpu --steps 10 2>&1 | tail -1 > gpurun_out/r2_bench_default.json
python - <<'PY'
import json
for n in ("pairs","default"):
    d=json.loads(open(f"gpurun_out/r2_bench_{n}.json").read())
    print(n, d["ms_per_step"], d["config"]["precision"])
    for k in d["kernels"]: print("   %-28s %5.0f launches %7.3f ms %6.1f TF/s" % (k["kernel"], k["launches_per_step"], k["ms_per_step"], k.get("tflops") or 0))
PY
