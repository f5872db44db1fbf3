for c in cfg1 cfg2 cfg3 cfg4 cfg5; do timeout 300 python tools/run_configs.py $c 2>&1 | tail -1; done > gpurun_out/r2_configs_default.jsonl
timeout 300 python tools/run_configs.py cfg5 --mode bf16 2>&1 | tail -1 > gpurun_out/r2_configs_cfg5_bf16.jsonl
timeout 300 python tools/run_configs.py cfg2 --mode tf32x3 2>&1 | tail -1 > gpurun_out/r2_configs_cfg2_tf32x3.jsonl
cat gpurun_out/r2_configs_default.jsonl gpurun_out/r2_configs_cfg5_bf16.jsonl gpurun_out/r2_configs_cfg2_tf32x3.jsonl | cut -c1-260
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -4
