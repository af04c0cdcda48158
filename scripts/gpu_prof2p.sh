ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k2p -s 470 -c 4 --csv --log-file gpurun_out/launches_2p.csv python scripts/bench_two_phase.py > gpurun_out/ncu_2p.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_2p.csv')) if len(r)>10]
hdr=rows[0]
for r in rows[1:]:
    d=dict(zip(hdr,r)); print(d['Kernel Name'][:40], d['Metric Name'], d['Metric Value'], d['Metric Unit'])
PY
