python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for env in "RV_G4_VARIANT=0" "RV_G4_VARIANT=4" "RV_G4_VARIANT=3 RV_WALK_OCC=5"; do
  env $env python bench.py --steps 6 --warmup 3 --e2e-steps 0 --parity none --skip-cpu 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.load(sys.stdin); print('$env', d['ms_per_step'], d['roofline']['split_ms'])"
done
