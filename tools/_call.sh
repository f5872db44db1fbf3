set -x
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/launches_r2.csv python tools/one_step.py 3 > gpurun_out/one_step_ncu.log 2>&1
timeout 400 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/launches_cfg3_r2.csv python tools/one_step.py 3 cfg3 > gpurun_out/one_step_cfg3_ncu.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:halo -s 12 -c 6 -f -o gpurun_out/prof_halo_r2 python tools/one_step.py 3 > gpurun_out/ncu_halo.log 2>&1
timeout 400 $NCU --set full --import-source on --kernel-name-base demangled -k regex:"Emb" -s 5 -c 5 -f -o gpurun_out/prof_emb_r2 python tools/one_step.py 2 cfg3 16 > gpurun_out/ncu_emb.log 2>&1
ls -la gpurun_out/*r2.ncu-rep
