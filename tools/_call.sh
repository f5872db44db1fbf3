set -x
timeout 600 python -m pytest tests -m gpu -q -k "emb or bn or batchnorm or cfg3 or reference or implicit" 2>&1 | grep -v Warning | tail -6
timeout 300 python tools/run_configs.py cfg3 2>&1 | tail -1
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/launches_r2.csv python tools/one_step.py 3 > gpurun_out/one_step_ncu.log 2>&1
timeout 400 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/launches_cfg3_r2.csv python tools/one_step.py 3 cfg3 > gpurun_out/one_step_cfg3_ncu.log 2>&1
timeout 400 $NCU --set full --import-source on --kernel-name-base demangled -k regex:WgradPolicy -s 6 -c 3 -f -o gpurun_out/prof_wgrad_r2 python tools/one_step.py 3 > gpurun_out/ncu_wgrad.log 2>&1
timeout 400 $NCU --set full --import-source on --kernel-name-base demangled -k regex:"GemmPolicy<" -s 290 -c 12 -f -o gpurun_out/prof_gemm_r2 python tools/one_step.py 3 > gpurun_out/ncu_gemm.log 2>&1
timeout 400 $NCU --set full --import-source on --kernel-name-base demangled -k regex:"Emb|bn_" -s 14 -c 14 -f -o gpurun_out/prof_emb_r2 python tools/one_step.py 2 cfg3 16 > gpurun_out/ncu_emb.log 2>&1
ls -la gpurun_out/*r2.ncu-rep
