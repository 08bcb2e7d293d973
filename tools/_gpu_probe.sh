python tools/parity_configs.py --configs 2,1 --scale 1.0 2>&1 | grep -v "^    " | tail -1
ls _work
cd _work/parity_c2_5002600_l1
run() { ../../build/rabbitvar_b200 -G ref.fa -b "T.bam|N.bam" -N "T|N" -i tiles.bed -c 1 -S 2 -E 3 -g 4 --fisher --out /tmp/x.tsv "$@" | grep -E "timeline" | sed -e 's/\[info\] //'; }
head -2 tiles.bed > /tmp/two.bed
echo "== two tiles, th 1 (CUDA start-up without decode threads beside it)"
for i in 1 2 3 4 5 6; do ../../build/rabbitvar_b200 -G ref.fa -b "T.bam|N.bam" -N "T|N" -i /tmp/two.bed -c 1 -S 2 -E 3 -g 4 --fisher --out /tmp/y.tsv --th 1 | grep timeline; done
echo "== full, th 16"
for i in 1 2 3 4 5 6; do run --th 16; done
echo "== full, th 8"
for i in 1 2 3; do run --th 8; done
