python tools/parity_configs.py --configs 1,2 2>&1 | grep -v "^    " | tail -3
cd _work/parity_c1_1002600_l1
for i in 1 2; do ( time ../../build/rabbitvar_b200 -G ref.fa -b S.bam -N S -i tiles.bed -c 1 -S 2 -E 3 -g 4 --th 16 --out /tmp/x.tsv ) 2>&1 | grep -E "timeline|real|total"; done
cd ../parity_c2_5002600_l1
for w in 2 3; do ( time ../../build/rabbitvar_b200 -G ref.fa -b "T.bam|N.bam" -N "T|N" -i tiles.bed -c 1 -S 2 -E 3 -g 4 --fisher --th 16 --workers $w --out /tmp/x.tsv ) 2>&1 | grep -E "timeline|real|total|jobs"; done
python - <<'PY'
import torch, subprocess, time
x = torch.zeros(1<<28, device="cuda"); torch.cuda.synchronize()
t=time.time(); r=subprocess.run(["../../build/rabbitvar_b200","-G","ref.fa","-b","T.bam|N.bam","-N","T|N","-i","tiles.bed","-c","1","-S","2","-E","3","-g","4","--fisher","--th","16","--out","/tmp/x.tsv"],capture_output=True,text=True); print("with a torch parent holding the GPU:", time.time()-t); print(r.stdout[-400:])
PY
