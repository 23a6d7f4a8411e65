set -x
mkdir -p gpurun_out
timeout 200 python scripts/dbg_mauna.py > gpurun_out/dbg_mauna.log 2>&1; cat gpurun_out/dbg_mauna.log
timeout 300 python -m pytest tests/test_gpu_programs.py -q --timeout 200 -k "noise or comp3 or mauna" 2>&1 | grep -v " err " | tail -15
for fx in 0 1; do GPK_OZ_FIXED=$fx timeout 200 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu --no-der 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('FIXED=$fx', d['value'], d['stage_ms_per_eval'], d['parity']['rel_err'], d['clocks']['sm_mhz'], d['gpu_launches'])"; done
GPK_PROFILE_DUMP=1 timeout 100 python scripts/step_profile.py > gpurun_out/step_profile_r2d.log 2>&1; tail -24 gpurun_out/step_profile_r2d.log
