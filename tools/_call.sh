timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py -m gpu -q -k "pool or vgg or conv3x3" 2>&1 | grep -v Warning | tail -4
timeout 400 python bench.py --no-cpu --no-ref-gpu --steps 10 2>&1 | tail -1 > gpurun_out/r2_bench_pool.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_pool.json").read())
print("ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "loss", d["final_loss"], "launches", d["gpu_launches"])
for k in d["kernels"][:14]: print("   %-28s %5.0f launches %7.3f ms" % (k["kernel"], k["launches_per_step"], k["ms_per_step"]))
PY
