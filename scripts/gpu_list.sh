#!/bin/bash
# ncu launch list (time + DRAM bytes per launch) of scripts/prof_target.py (one cfg2 build + the 10k-query hit batch); tag $1
mkdir -p gpurun_out
R=${1:-k}
timeout -k 10 300 python scripts/prof_target.py 20000 1 > gpurun_out/plain_$R.log 2>&1 &&
timeout -k 10 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv python scripts/prof_target.py 20000 1 > gpurun_out/ncu_list_$R.log 2>&1
echo "ncu list rc=$?"
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_$R.csv')) if len(r)>14 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0].split('::')[-1]; 
    a=agg.setdefault((r[0],name),{})
    a[r[12]]=float(r[14].replace(',',''))
    a['unit_'+r[12]]=r[13]
tot=collections.OrderedDict()
for (i,name),a in agg.items():
    t=a.get('gpu__time_duration.sum',0); u=a.get('unit_gpu__time_duration.sum','ns')
    t*= {'ns':1e-6,'us':1e-3,'ms':1,'s':1e3}.get(u,1e-6)
    def mb(k):
        v=a.get(k,0); uu=a.get('unit_'+k,'byte'); return v*{'byte':1e-6,'Kbyte':1e-3,'Mbyte':1,'Gbyte':1e3}.get(uu,1e-6)
    x=tot.setdefault(name,[0,0.0,0.0,0.0]); x[0]+=1; x[1]+=t; x[2]+=mb('dram__bytes_read.sum'); x[3]+=mb('dram__bytes_write.sum')
for n,(c,t,r,w) in tot.items(): print(f"{n:40s} n={c:3d} ms={t:8.4f} rdMB={r:9.1f} wrMB={w:9.1f}")
print('total ms', sum(x[1] for x in tot.values()))
PY
