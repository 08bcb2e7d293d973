python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for occ in 4 6; do
  RV_WALK_OCC=$occ python bench.py --steps 6 --warmup 3 --e2e-steps 0 --skip-cpu 2>gpurun_out/w$occ.err | python -c "
import json,sys
l=json.loads(sys.stdin.read()); r=l['roofline']; print('occ $occ', l['ms_per_step'], r['split_ms'])"
done
