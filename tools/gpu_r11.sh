#!/bin/bash
# 1-GPU call r11: the rewritten large-sample normalisation path (c4: 6.3 MB samples): parity, c4 bench, launch list.
tag=${1:-r11}
out=gpurun_out/$tag
mkdir -p $out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $out/pytest_gpu.log
( timeout 300 python bench.py --workload c4 --steps 300 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_c4.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/ncu_launches_bench_c4.csv \
    python bench.py --workload c4 --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $out/ncu_launch_bench_c4.log 2>&1
tail -4 $out/pytest_gpu.log; cut -c1-250 $out/bench_c4.json; grep -c l2_ $out/ncu_launches_bench_c4.csv
python - <<'PY'
import csv, re
rows=list(csv.reader(open('gpurun_out/'+"r11"+'/ncu_launches_bench_c4.csv')))
hdr=None; seq=[]
for r in rows:
    if r and r[0]=="ID": hdr=r; continue
    if hdr and len(r)==len(hdr) and 'dct::' in r[4]:
        m=re.search(r'dct::(\w+)(<[^(]*)?', r[4]); seq.append((m.group(1)+((m.group(2) or '')[:30]), float(r[-1]), r[8]))
st=[i for i,(k,_,_) in enumerate(seq) if 'JsdOp' in k]
tot=0
for k,v,g in seq[st[2]:st[3]]:
    print(f"{v/1e3:9.1f} us grid {g:14s} {k}"); tot+=v
print("sum", tot/1e3)
PY
