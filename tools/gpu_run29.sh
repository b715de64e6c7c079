#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/r2H_bench_n1.json 2> gpurun_out/r2H_bench_n1.err
tail -3 gpurun_out/r2H_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2H_bench_n1.json'))
print("headline %.5g e2e %.5g frac %.4f" % (d['value'], d['e2e']['value'], d['roofline']['frac']))
for w in d['extra_workloads']: print(w['config'], "%.4g" % w['value'], "e2e %.4g" % w['e2e']['value'], "frac %.5f" % w['roofline']['frac'])
PY
