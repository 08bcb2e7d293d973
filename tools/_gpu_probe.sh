set -x
for env in "RV_G4_VARIANT=0 RV_WALK_OCC=4" "RV_G4_VARIANT=3 RV_WALK_OCC=8" "RV_G4_VARIANT=1 RV_WALK_OCC=6"; do
  env $env python bench.py --steps 6 --warmup 3 --e2e-steps 0 --skip-cpu 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.load(sys.stdin); print('$env', d['ms_per_step'], d['roofline']['split_ms'])"
done
ncu --set full --clock-control none --import-source on -k regex:"rv_walk|rv_gather4|rv_apply" -s 9 -c 3 -o gpurun_out/prof_r2_a python bench.py --steps 1 --warmup 3 --e2e-steps 0 --skip-cpu > gpurun_out/ncu_a.log 2>&1
tail -3 gpurun_out/ncu_a.log
