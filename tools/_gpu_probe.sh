python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err
tail -c 3000 gpurun_out/bench_final_n1.json
tail -5 gpurun_out/bench_final_n1.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
tail -c 1500 gpurun_out/bench_final_ref.json
