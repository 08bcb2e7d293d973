python bench.py > gpurun_out/b14.json 2> gpurun_out/b14.err; cat gpurun_out/b14.json
