#!/bin/bash
# Run on the GPU box (gpurun): collects the round's ncu evidence into gpurun_out/ as text / json.
#   tools/make_profiles.sh <tag>
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
./tools/fp64_peak > $OUT/${TAG}_fp64_peak.json 2>&1
KREGEX='regex:continuum|ray_|gamma|fs_kernel|stokes_kernel'
# 1. launch list of the bench command itself (cold-cache, serialised: compare SHARES, not absolute times)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 > $OUT/${TAG}_launches_bench.log 2>&1
# 2. full captures of the pipeline kernels: one Gamma iteration of a 128-column stack of config 3, of config 2,
#    and of the 500-depth benchmark shape of the reference (multi-warp ray kernel)
for WL in c3 c2 deep c5; do
  case $WL in
    c3) ARGS="128 2 c3"; N=5;;
    c2) ARGS="1 2 c2"; N=9;;
    deep) ARGS="1 2 deep"; N=5;;
    c5) ARGS="128 2 c5"; N=6;;
  esac
  ncu --set full --clock-control none --import-source on -k "$KREGEX" -s $N -c $N \
      -o $OUT/${TAG}_full_$WL -f python tools/prof_c3.py $ARGS > $OUT/${TAG}_full_$WL.log 2>&1
  python tools/ncu_summary.py $OUT/${TAG}_full_$WL.ncu-rep > $OUT/${TAG}_summary_$WL.txt 2>&1
  ncu -i $OUT/${TAG}_full_$WL.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_traffic.py > $OUT/${TAG}_traffic_$WL.json
done
NCU_KERNEL_ID=::regex:ray_smem:1 python tools/ncu_lines.py $OUT/${TAG}_full_c3.ncu-rep ray_smem_kernelILi3ELi1 60 > $OUT/${TAG}_lines_c3_ray_NL1.txt 2>&1
NCU_KERNEL_ID=::regex:gamma_tile:1 python tools/ncu_lines.py $OUT/${TAG}_full_c3.ncu-rep gamma_tile_kernelILi1 40 > $OUT/${TAG}_lines_c3_gamma.txt 2>&1
rm -f $OUT/${TAG}_full_c2.ncu-rep $OUT/${TAG}_full_deep.ncu-rep $OUT/${TAG}_full_c5.ncu-rep
ls -la $OUT | tail -20
