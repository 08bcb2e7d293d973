python bench.py --steps 4 --warmup 3 --e2e-steps 0 --parity none --skip-cpu 2> gpurun_out/bench_f.err | tail -1 | python -c "
import json,sys; d=json.load(sys.stdin); print(d['ms_per_step'], d['roofline']['split_ms'])"
grep pileup: gpurun_out/bench_f.err
ncu --set full --clock-control none --import-source on -k regex:"rv_walk|rv_apply|rv_pileup_kernel" -s 9 -c 3 -o gpurun_out/prof_r2_b python bench.py --steps 1 --warmup 3 --e2e-steps 0 --parity none --skip-cpu > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
