timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for k in 2 5; do
  python bench.py --workload $k --steps 6 --warmup 3 --e2e-steps 0 --parity none --skip-cpu --no-scaling-base 2> gpurun_out/tmp.err | tail -1 | python -c "
import json,sys; d=json.load(sys.stdin); print('config $k', d['ms_per_step'], d['roofline']['split_ms'])"
  grep pileup: gpurun_out/tmp.err
done
python tools/parity_configs.py --configs 5,2 2>&1 | grep -v "^    " | tail -3
cd _work/parity_c2_5002600_l1
for pin in 0 1; do RV_PIN_JOBS=$pin ../../build/rabbitvar_b200 -G ref.fa -b "T.bam|N.bam" -N "T|N" -i tiles.bed -c 1 -S 2 -E 3 -g 4 --fisher --th 16 --out /tmp/x.tsv | grep -E "timeline"; done
