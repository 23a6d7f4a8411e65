set -x
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:oz_syrk -s 3 -c 1 -o gpurun_out/prof_oz_syrk_r2 -f python scripts/one_eval.py 16384 1 > gpurun_out/ncu_oz_r2.log 2>&1; tail -2 gpurun_out/ncu_oz_r2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:oz_slice -s 2 -c 1 -o gpurun_out/prof_oz_slice_r2 -f python scripts/one_eval.py 16384 1 > gpurun_out/ncu_ozs_r2.log 2>&1; tail -1 gpurun_out/ncu_ozs_r2.log
