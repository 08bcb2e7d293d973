set -x
python tools/parity_configs.py --configs 2,5 2>&1 | grep -v "^    " | tail -3
cd _work/parity_c2_5002600_l1
for w in 2 4 6; do ../../build/rabbitvar_b200 -G ref.fa -b "T.bam|N.bam" -N "T|N" -i tiles.bed -c 1 -S 2 -E 3 -g 4 --fisher --th 16 --workers $w --out /tmp/x.tsv | grep -E "timeline"; done
cd ../..
# launch list of one bench step (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --e2e-steps 0 --parity none --skip-cpu --no-scaling-base > /dev/null 2>&1
# full counters of the four kernels of the pileup stage + the scoring kernels, one launch each
ncu --set full --clock-control none --import-source on -k regex:"rv_pileup_kernel|rv_walk|rv_gather4|rv_apply|rv_score" -s 14 -c 7 -o gpurun_out/prof_r2_c python bench.py --steps 1 --warmup 3 --e2e-steps 0 --parity none --skip-cpu --no-scaling-base > gpurun_out/ncu_c.log 2>&1
tail -2 gpurun_out/ncu_c.log
