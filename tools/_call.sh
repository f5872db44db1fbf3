timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/dp_parity.py 2>&1 | tail -2 > gpurun_out/r2_dp_parity_8gpu.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu --no-ref-gpu 2>&1 | tail -1 > gpurun_out/r2_bench_n8_final.json
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu --no-ref-gpu 2>&1 | tail -1 > gpurun_out/r2_bench_n1_same_box_as_n8.json
cat gpurun_out/r2_dp_parity_8gpu.log
python - <<'PY'
import json
for f in ("r2_bench_n8_final","r2_bench_n1_same_box_as_n8"):
    d=json.loads(open(f"gpurun_out/{f}.json").read())
    print(f, {k:d[k] for k in ("value","n_gpus","ms_per_step","gpu_launches") if k in d}, d["e2e"]["value"], d.get("allreduce"))
PY
