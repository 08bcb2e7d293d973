export RV_PIPE_TRACE=gpurun_out/pipe_trace.csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for cfg in "16 5" "16 6" "16 8" "15 5"; do
  set -- $cfg
  python bench.py --steps 2 --warmup 3 --e2e-steps 3 --skip-cpu --workers $1 --chunk $2 2>gpurun_out/e.err | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('workers $1 chunk $2 e2e', l['e2e']['sec_per_step'], l['e2e']['value'])"
  grep "rvh_pipeline\|e2e T" gpurun_out/e.err | tail -2 | cut -c1-220
done
