timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for k in 2 5; do
  python bench.py --workload $k --steps 10 --warmup 3 --e2e-steps 0 --parity none --skip-cpu --no-scaling-base 2> gpurun_out/tmp.err | tail -1 | python -c "
import json,sys; d=json.load(sys.stdin); print('config $k', d['ms_per_step'], d['roofline'].get('split_ms'))"
done
