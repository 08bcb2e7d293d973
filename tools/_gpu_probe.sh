RV_G4_VARIANT=5 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or reference_binary_run_here or deterministic or tiled" 2>&1 | tail -4
for v in 4 5 6; do
  RV_G4_VARIANT=$v timeout 300 python bench.py --steps 6 --warmup 3 --e2e-steps 0 --parity none --skip-cpu --no-scaling-base 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.load(sys.stdin); print('variant $v', d['ms_per_step'], d['roofline']['split_ms'])"
done
RV_G4_VARIANT=5 timeout 300 python bench.py --workload 3 --steps 6 --warmup 3 --e2e-steps 0 --parity none --skip-cpu --no-scaling-base 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.load(sys.stdin); print('config 3 variant 5', d['ms_per_step'], d['roofline']['split_ms'])"
