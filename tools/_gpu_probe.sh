timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for v in 4 3; do for k in 2 4; do
  RV_G4_VARIANT=$v python bench.py --workload $k --steps 10 --warmup 3 --e2e-steps 0 --parity none --skip-cpu --no-scaling-base 2> gpurun_out/tmp.err | tail -1 | python -c "
import json,sys; d=json.load(sys.stdin); print('variant $v config $k', d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline'].get('split_ms'), d['roofline']['score_kernel']['kernel_ms'])"
  grep "pileup:" gpurun_out/tmp.err
done; done
python tools/parity_configs.py --configs 1,2,3,5 2>&1 | grep -v "^    " | tail -1
