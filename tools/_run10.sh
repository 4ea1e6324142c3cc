set -x
python tools/pool_probe.py > gpurun_out/plain_pool.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o /tmp/pool python tools/pool_probe.py > gpurun_out/r02_prof_pool1.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/pool.ncu-rep gpurun_out/r02_pool_ncu_full.csv gemm_tc
ncu -i /tmp/pool.ncu-rep --page source --csv > /tmp/pool_source.csv 2>/dev/null; wc -l /tmp/pool_source.csv
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/pool_source.csv')))
# find header
hi=next(i for i,r in enumerate(rows) if any('Source' in c for c in r) and len(r)>5)
hdr=rows[hi]; print(hdr[:40])
def col(name):
    for i,h in enumerate(hdr):
        if h.strip()==name: return i
    return None
cs=col('Source'); 
samp=[i for i,h in enumerate(hdr) if 'Samples' in h]
print('sample cols', [(i,hdr[i]) for i in samp])
si=samp[0] if samp else None
data=[]
for r in rows[hi+1:]:
    try: v=float(r[si])
    except: continue
    data.append((v,r))
data.sort(key=lambda t:-t[0])
tot=sum(v for v,_ in data)
with open('gpurun_out/r02_pool_source_top.txt','w') as f:
    f.write('total samples %d\n'%tot)
    f.write(' | '.join(hdr)+'\n')
    for v,r in data[:60]:
        f.write(' | '.join(c[:90] for c in r)+'\n')
PY
head -30 gpurun_out/r02_pool_source_top.txt | cut -c1-400
