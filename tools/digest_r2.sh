#!/bin/bash
# Turns the round-2 ncu output of tools/profile_r2.sh (gpurun_out/r2_*) into the digests under profiles/.
cd "$(dirname "$0")/.."
o=gpurun_out p=profiles
python tools/ncu_launches.py $o/r2_launches.csv > $p/r2_launches.txt
python tools/ncu_launches.py $o/r2_traffic.csv --traffic-json $p/traffic.json "k_wf_trace<0, 0>" > $p/r2_traffic.txt
python tools/ncu_summary.py $o/r2_trace.ncu-rep > $p/r2_trace_ncu_summary.txt
python tools/ncu_lines.py $o/r2_trace.ncu-rep "k_wf_trace<(bool)0" 40 > $p/r2_trace_ncu_lines.txt
python tools/ncu_phases.py $o/r2_trace.ncu-rep "k_wf_trace<(bool)0" kf_trace.cuh setup:1-107 push:108-116 popGroup:117-139 refill:140-184 inst:185-255 node:256-274 tri:275-313 pop:314-316 finish:317-327 >> $p/r2_trace_ncu_lines.txt
python tools/ncu_summary.py $o/r2_shade.ncu-rep > $p/r2_shade_ncu_summary.txt
python tools/ncu_lines.py $o/r2_shade.ncu-rep "k_wf_shade" 25 >> $p/r2_shade_ncu_summary.txt
{
  for s in unique unique10m; do
    echo "## $s (8-spp batch, 1920x1080)"; cat $o/r2_${s}_counters.txt; python tools/ncu_launches.py $o/r2_${s}_traffic.csv; echo
  done
} > $p/r2_unique_traffic.txt
python - <<'PY'
import csv, json, subprocess
txt = subprocess.run(['ncu', '-i', 'gpurun_out/r2_trace.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h = rows[0]
for v in rows[2:]:
    d = dict(zip(h, v))
    if 'k_wf_trace<0' in d.get('Kernel Name', ''):
        out = {"kernel": d['Kernel Name'],
               "issue_active_pct": float(d['smsp__issue_active.avg.pct_of_peak_sustained_active'].replace(',', '')),
               "lanes_per_inst": float(d['smsp__thread_inst_executed_per_inst_executed.ratio'].replace(',', '')),
               "source": "ncu --set full, depth-1 closest-hit launch of an 8-spp batch of the bench frame (gpurun_out/r2_trace.ncu-rep, tools/profile_r2.sh)"}
        json.dump(out, open('profiles/issue.json', 'w'), indent=1)
        print(out)
        break
PY
