#!/bin/bash
# ncu captures of round 2 (run on the GPU box through gpurun; reports land in gpurun_out/)
set -x
NCU="ncu --set full --clock-control none"
$NCU -k regex:k_cond_bwd -c 1 -o gpurun_out/r02_T_bwd python tools/kbench.py --reps 1 > /dev/null 2>&1
$NCU -k regex:k_cond_fwd -c 1 -o gpurun_out/r02_T_fwd python tools/kbench.py --reps 1 > /dev/null 2>&1
$NCU -k regex:k_gram -c 1 -o gpurun_out/r02_T_gram python tools/kbench.py --reps 1 > /dev/null 2>&1
$NCU -k regex:k_cond_bwd -c 1 -o gpurun_out/r02_T_bwd_R1 python tools/kbench.py --reps 1 --R 1 > /dev/null 2>&1
$NCU -k regex:k_cond_bwd -c 1 -o gpurun_out/r02_C2_bwd python tools/kbench.py --reps 1 --M 128 > /dev/null 2>&1
$NCU -k regex:k_cond_bwd -c 1 -o gpurun_out/r02_C3_bwd python tools/kbench.py --reps 1 --M 500 --D 16 --R 16 --N 160000 --kind matern52 > /dev/null 2>&1
$NCU -k regex:k_cond_fwd -c 1 -o gpurun_out/r02_C3_fwd python tools/kbench.py --reps 1 --M 500 --D 16 --R 16 --N 160000 --kind matern52 > /dev/null 2>&1
$NCU -k regex:k_cond_fwd -c 1 -o gpurun_out/r02_C5_fwd python tools/kbench.py --reps 1 --M 1024 --R 8 --N 204800 > /dev/null 2>&1
$NCU -k regex:k_cond_bwd -c 1 -o gpurun_out/r02_M1024_bwd python tools/kbench.py --reps 1 --M 1024 --R 8 --N 51200 > /dev/null 2>&1
$NCU -k regex:k_cond_bwd -c 1 -o gpurun_out/r02_C4_bwd python tools/kbench.py --reps 1 --R 1 --N 10000 > /dev/null 2>&1
# the reports stay on the box (64 MiB limit on what travels back): only their raw-page CSVs and the traffic table come home
for r in gpurun_out/r02_*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}_raw.csv 2>/dev/null; done
python tools/ncu_summary.py traffic gpurun_out/r02_T_bwd.ncu-rep "k_cond_bwd|M256,D8,R8,N100000" gpurun_out/r02_traffic.json
python tools/ncu_summary.py traffic gpurun_out/r02_T_fwd.ncu-rep "k_cond_fwd|M256,D8,R8,N100000" gpurun_out/r02_traffic.json
python tools/ncu_summary.py traffic gpurun_out/r02_T_gram.ncu-rep "k_gram|M256,D8,R8,N100000" gpurun_out/r02_traffic.json
python tools/ncu_summary.py traffic gpurun_out/r02_T_bwd_R1.ncu-rep "k_cond_bwd|M256,D8,R1,N100000" gpurun_out/r02_traffic.json
python tools/ncu_summary.py traffic gpurun_out/r02_C2_bwd.ncu-rep "k_cond_bwd|M128,D8,R8,N100000" gpurun_out/r02_traffic.json
python tools/ncu_summary.py traffic gpurun_out/r02_C3_bwd.ncu-rep "k_cond_bwd|M500,D16,R16,N160000" gpurun_out/r02_traffic.json
python tools/ncu_summary.py traffic gpurun_out/r02_C3_fwd.ncu-rep "k_cond_fwd|M500,D16,R16,N160000" gpurun_out/r02_traffic.json
python tools/ncu_summary.py traffic gpurun_out/r02_C5_fwd.ncu-rep "k_cond_fwd|M1024,D8,R8,N204800" gpurun_out/r02_traffic.json
python tools/ncu_summary.py traffic gpurun_out/r02_M1024_bwd.ncu-rep "k_cond_bwd|M1024,D8,R8,N51200" gpurun_out/r02_traffic.json
python tools/ncu_summary.py traffic gpurun_out/r02_C4_bwd.ncu-rep "k_cond_bwd|M256,D8,R1,N10000" gpurun_out/r02_traffic.json
rm -f gpurun_out/r02_*.ncu-rep
python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02_plain.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --no-graph > gpurun_out/r02_ncu_bench.log 2>&1
./tools/fp64_peak > gpurun_out/fp64_peak_r02.txt 2>&1
for s in "--M 256" "--M 256 --R 1" "--M 128" "--M 500 --D 16 --R 16 --N 160000 --kind matern52" "--M 1024 --R 8 --N 51200" "--M 1024 --R 8 --N 204800"; do python tools/kbench.py --reps 4 $s; done > gpurun_out/r02_kbench.txt 2>&1
du -sh gpurun_out
