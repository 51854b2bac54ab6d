mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_sort_launches.csv python tools/prof_round2.py --what find64 --queries 67108864 > gpurun_out/r2_sort.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2_sort_launches.csv')))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
for r in rows[hdr+1:]:
    if len(r)>iv: print(f"{float(r[iv].replace(',',''))/1e3:10.1f} us  {r[ik][:80]}")
PY
