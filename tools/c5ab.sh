#!/bin/bash
# A/B of UAM_GRID_HALF_CAP on config C5: tools/c5ab.sh 0 4 6 8
for cap in "$@"; do
UAM_GRID_HALF_CAP=$cap timeout 600 python bench_configs.py --only c5 --no-cpu > gpurun_out/c5_cap$cap.json 2> gpurun_out/c5_err.log
python - <<PY
import json
for l in open("gpurun_out/c5_cap$cap.json"):
    l=l.strip()
    if not l.startswith("{"): continue
    d=json.loads(l)
    for c in (d.get("C5") or [d]):
        if "bands" in c: print("cap=$cap bands", c["bands"], "routes/s", round(c["start_goal_queries"]["queries_per_s"],2), "sweeps/s", round(c["full_sweeps"]["queries_per_s"],2), "act", c["full_sweeps"].get("activations_rank0"), "dsweeps", c["full_sweeps"].get("sweeps_rank0"), "rounds", c["full_sweeps"].get("rounds_rank0"), c["start_goal_queries"]["goal_distances_equal_full_sweep"])
PY
done
