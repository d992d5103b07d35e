#!/bin/bash
# development aid (under gpurun): per-kernel duration / instructions / pipe utilisation of ONE Gamma iteration
#   tools/kernel_table.sh <tag> <ncol> <workload>     -> gpurun_out/<tag>_ktable.txt
TAG=$1; NCOL=${2:-128}; WL=${3:-c3}
M=gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sass__inst_executed_local_loads,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -k regex:"continuum|ray_|gamma" -s 5 -c 5 --csv --log-file gpurun_out/${TAG}_ktable.csv python tools/prof_c3.py $NCOL 3 $WL > gpurun_out/${TAG}_ktable.log 2>&1
python - <<PY
import csv,collections
rows=list(csv.DictReader(l for l in open('gpurun_out/${TAG}_ktable.csv') if l.startswith('"')))
k=collections.OrderedDict()
for r in rows:
    key=(r['ID'], r['Kernel Name'][:60])
    k.setdefault(key,{})[r['Metric Name']]=r['Metric Value']
out=open('gpurun_out/${TAG}_ktable.txt','w')
tot=0
for (i,name),m in k.items():
    f=lambda x: (lambda v: float(v) if v.replace(".","").isdigit() else float("nan"))(m.get(x,"0").replace(",",""))
    d=f('gpu__time_duration.sum'); tot+=d
    print('%-62s %9.1f us inst %11.0f issue %5.1f%% fp64 %5.1f%% warps %5.1f%% regs %3.0f lld %9.0f dramR %7.1fMB W %7.1fMB'%(name,d/1e3,f('smsp__inst_executed.sum'),f('smsp__issue_active.avg.pct_of_peak_sustained_active'),f('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'),f('sm__warps_active.avg.pct_of_peak_sustained_active'),f('launch__registers_per_thread'),f('sass__inst_executed_local_loads'),f('dram__bytes_read.sum')/1e6,f('dram__bytes_write.sum')/1e6),file=out)
print('total %.1f us'%(tot/1e3),file=out)
PY
cat gpurun_out/${TAG}_ktable.txt
