#!/bin/bash
# Run on the GPU box (gpurun): collects the round's ncu evidence into gpurun_out/ as text.
#   tools/make_profiles.sh <tag>
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
./tools/fp64_peak > $OUT/${TAG}_fp64_peak.json 2>&1
# 1. launch list of the bench command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_c2.csv \
    python bench.py --steps 2 --warmup 3 > $OUT/${TAG}_launches_c2.log 2>&1
# 2. full captures of the pipeline kernels: one Gamma iteration of c2 and of a 128-column stack
for WL in c2 c3; do
  if [ $WL = c2 ]; then ARGS="1 2 c2"; N=9; else ARGS="128 2 c3"; N=5; fi
  ncu --set full --clock-control none --import-source on -k regex:"continuum_kernel|ray_kernel|gamma_kernel|fs_kernel" -s $N -c $N \
      -o $OUT/${TAG}_full_$WL -f python tools/prof_c3.py $ARGS > $OUT/${TAG}_full_$WL.log 2>&1
  python tools/ncu_summary.py $OUT/${TAG}_full_$WL.ncu-rep > $OUT/${TAG}_summary_$WL.txt 2>&1
  ncu -i $OUT/${TAG}_full_$WL.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_traffic.py > $OUT/${TAG}_traffic_$WL.json
done
NCU_KERNEL_ID=::regex:ray_kernel:1 python tools/ncu_lines.py $OUT/${TAG}_full_c3.ncu-rep ray_kernelILi3ELi2ELi1 60 > $OUT/${TAG}_lines_c3_ray_NL1.txt 2>&1
NCU_KERNEL_ID=::regex:gamma_kernel:1 python tools/ncu_lines.py $OUT/${TAG}_full_c3.ncu-rep gamma_kernel 40 > $OUT/${TAG}_lines_c3_gamma.txt 2>&1
rm -f $OUT/${TAG}_full_c2.ncu-rep
ls -la $OUT
