#!/bin/bash
# DRAM traffic of one step (all waves) of the config-3 / config-4 workloads; launch list of the default bench
mkdir -p gpurun_out
for w in config3 config4; do
  CMD="python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
  timeout 600 $CMD > gpurun_out/r2u_$w.json 2> gpurun_out/r2u_$w.err
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sweep_batch -c 40 --csv --log-file gpurun_out/r2u_${w}_dram_traffic.csv $CMD > /dev/null 2>&1
  grep -c sweep_batch gpurun_out/r2u_${w}_dram_traffic.csv
done
CMD="python bench.py --no-extras --no-cpu-baseline --steps 2 --warmup 1"
$CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2u_launches_static.csv $CMD > /dev/null 2>&1
python - <<'PY'
import csv
for w in ("config3","config4"):
    rows=[r for r in csv.reader(open("gpurun_out/r2u_%s_dram_traffic.csv"%w)) if len(r)>14 and "sweep_batch" in r[4]]
    per={}
    for r in rows: per.setdefault(int(r[0]),{})[r[12]]=float(r[14])
    ids=sorted(per)
    print(w, "launches", len(ids), [ (round(per[i]['gpu__time_duration.sum']/1e6,1)) for i in ids])
    print(w, "read", [round(per[i]['dram__bytes_read.sum']/1e9,1) for i in ids], "write", [round(per[i]['dram__bytes_write.sum']/1e9,1) for i in ids])
PY
cut -c1-300 gpurun_out/r2u_config3.json
