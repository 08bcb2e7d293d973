python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:rv_walk_kernel -c 1 -o gpurun_out/r01_v8walk -f python bench.py --steps 1 --warmup 1 --e2e-steps 0 --skip-cpu > gpurun_out/ncu8w.log 2>&1
RV_WALK_OCC=4 python bench.py --steps 6 --warmup 3 --e2e-steps 0 --skip-cpu 2>gpurun_out/w4.err | python -c "
import json,sys
l=json.loads(sys.stdin.read()); r=l['roofline']; print('occ 4', l['ms_per_step'], r['split_ms'])"
