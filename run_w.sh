mkdir -p gpurun_out
timeout 60 python tools/tune.py C2 10000 200 wk wk > gpurun_out/w6.log 2>&1
cut -c1-60,100-170 gpurun_out/w6.log
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gsm_ --launch-skip 6 -c 6 --csv --log-file gpurun_out/w6.csv python tools/tune.py C2 10000 200 wk > /dev/null 2>&1
python - <<'PY'
import csv
for r in csv.reader(open('gpurun_out/w6.csv')):
    if len(r)>10 and r[0].isdigit(): print(r[4][:40], r[-3], r[-1])
PY
