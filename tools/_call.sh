timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | grep -v Warning | tail -3
timeout 400 python bench.py --no-cpu --no-ref-gpu --steps 10 2>&1 | tail -1 > gpurun_out/r2_bench_v4.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_v4.json").read())
print("ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "loss", d["final_loss"], d["roofline"])
for k in d["kernels"][:12]: print("   %-28s %5.0f launches %7.3f ms" % (k["kernel"], k["launches_per_step"], k["ms_per_step"]))
PY
