mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2c_pytest.log
cat gpurun_out/r2c_pytest.log
timeout 900 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2c_bench.json'))
print(d['value'], d['roofline']['frac'], json.dumps(d['roofline']['fp64']))
for e in d['extra_workloads']: print(json.dumps(e)[:400])
print(d['e2e'])
"
tail -3 gpurun_out/r2c_bench.err
