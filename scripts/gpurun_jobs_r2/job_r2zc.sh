mkdir -p gpurun_out
GPK_PROFILE_DUMP=1 timeout 40 python scripts/step_profile.py > gpurun_out/step_profile_r2zc.log 2>&1; tail -3 gpurun_out/step_profile_r2zc.log
