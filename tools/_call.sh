timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v Warning | tail -4
echo "== bench PDL on"; timeout 400 python bench.py --no-cpu --no-ref-gpu --steps 10 2>&1 | tail -1 > gpurun_out/r2_bench_pdl_on.json
echo "== bench PDL off"; B200ASR_PDL=0 timeout 400 python bench.py --no-cpu --no-ref-gpu --steps 10 2>&1 | tail -1 > gpurun_out/r2_bench_pdl_off.json
python - <<'PY'
import json
for n in ("pdl_on","pdl_off"):
    d=json.loads(open(f"gpurun_out/r2_bench_{n}.json").read())
    print(n, "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "loss", d["final_loss"])
    for k in d["kernels"][:8]: print("   %-28s %5.0f launches %7.3f ms" % (k["kernel"], k["launches_per_step"], k["ms_per_step"]))
PY
for c in cfg3 cfg4 cfg5 cfg1; do timeout 300 python tools/run_configs.py $c 2>&1 | tail -1 | cut -c1-330; done
