#!/bin/bash
# round-1 GPU pass p (1 GPU): class-tile gather+scatter — A/B timing at 64^3 + dram bytes under ncu
mkdir -p gpurun_out
OGS_SLICES=65536,4096,512 timeout 900 python tools/quick_ogs.py 64 7 none > gpurun_out/quick_p.log 2>&1; grep -v "^$" gpurun_out/quick_p.log | tail -18
OGS_SLICES=65536 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum --clock-control none -k regex:"ogs_ctile|ogs_tile_pf" -c 60 --csv --log-file gpurun_out/ncu_p.csv python tools/quick_ogs.py 32 7 none > gpurun_out/ncu_p.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/ncu_p.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:60]),{})[r[mi]]=r[vi]
seen=collections.Counter()
for (i,k),m in d.items():
    seen[k]+=1
    if seen[k] in (8,): print(k,m)
PY
