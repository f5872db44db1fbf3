NCU="ncu --clock-control none"
B200ASR_ATTN=bf16x3 timeout 500 $NCU --set full --import-source on -k regex:sdpa_fused -s 36 -c 6 -f -o gpurun_out/prof_attn_fused_r2 python tools/one_step.py 3 > gpurun_out/ncu_attn_fused.log 2>&1
tail -3 gpurun_out/ncu_attn_fused.log
B200ASR_ATTN=bf16x3 timeout 300 python bench.py --no-cpu --no-ref-gpu --steps 10 2>&1 | tail -1 > gpurun_out/r2_bench_fused_attn.json
python tools/show_bench.py gpurun_out/r2_bench_fused_attn.json 2>/dev/null | sed -n 1,14p
