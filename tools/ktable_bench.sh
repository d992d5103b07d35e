#!/bin/bash
# development aid (under gpurun): per-kernel totals of a short bench.py run of any workload
#   tools/ktable_bench.sh <tag> <bench args...>    -> gpurun_out/<tag>_kbench.txt
TAG=$1; shift
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sass__inst_executed_local_loads,sass__inst_executed_local_stores,l1tex__t_sector_hit_rate.pct
ncu --metrics $M --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_kbench.csv python bench.py "$@" --steps 1 --warmup 3 > gpurun_out/${TAG}_kbench.log 2>&1
python - <<PY
import csv,collections
rows=list(csv.DictReader(l for l in open('gpurun_out/${TAG}_kbench.csv') if l.startswith('"')))
per=collections.OrderedDict()
for r in rows:
    per.setdefault(r['ID'],{'name':r['Kernel Name'][:70]})[r['Metric Name']]=r['Metric Value']
def num(v):
    try: return float(v.replace(',',''))
    except Exception: return 0.0
agg=collections.OrderedDict()
for i,m in per.items():
    f=lambda x: num(m.get(x,"0"))
    a=agg.setdefault(m['name'],dict(n=0,t=0.0,inst=0.0,issue=0.0,fp64=0.0,warps=0.0,regs=0,lld=0.0,lst=0.0,l1=0.0))
    d=f('gpu__time_duration.sum'); a['n']+=1; a['t']+=d; a['inst']+=f('smsp__inst_executed.sum')
    for k,mm in (('issue','smsp__issue_active.avg.pct_of_peak_sustained_active'),('fp64','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'),('warps','sm__warps_active.avg.pct_of_peak_sustained_active'),('l1','l1tex__t_sector_hit_rate.pct')):
        a[k]+=f(mm)*d
    a['regs']=f('launch__registers_per_thread'); a['lld']+=f('sass__inst_executed_local_loads'); a['lst']+=f('sass__inst_executed_local_stores')
out=open('gpurun_out/${TAG}_kbench.txt','w')
tot=sum(a['t'] for a in agg.values())
for name,a in sorted(agg.items(), key=lambda kv:-kv[1]['t']):
    t=a['t'] or 1
    print('%-72s n %4d %10.1f us %5.1f%% inst %12.0f issue %5.1f%% fp64 %5.1f%% warps %5.1f%% L1hit %5.1f%% regs %3.0f lld %10.0f lst %10.0f'%(name,a['n'],a['t']/1e3,100*a['t']/tot,a['inst'],a['issue']/t,a['fp64']/t,a['warps']/t,a['l1']/t,a['regs'],a['lld'],a['lst']),file=out)
print('total %.1f us'%(tot/1e3),file=out)
PY
cat gpurun_out/${TAG}_kbench.txt
