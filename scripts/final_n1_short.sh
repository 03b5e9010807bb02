#!/bin/bash
# Single-GPU bench lines after a kernel change late in the round: c1 and c2 in full (cpu_baseline + parity), the other
# configurations without the CPU leg.
set -u
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/v2_c2.json 2> gpurun_out/v2_c2.err; echo "c2 rc=$?"
python bench.py --config c1 --steps 20 --warmup 5 > gpurun_out/v2_c1.json 2> gpurun_out/v2_c1.err; echo "c1 rc=$?"
python bench.py --storage fp32 --steps 20 --warmup 5 --no-pageable --no-cpu > gpurun_out/v2_c2_fp32.json 2>/dev/null; echo "fp32 rc=$?"
for c in c3 c4 c5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu > gpurun_out/v2_$c.json 2> gpurun_out/v2_$c.err; echo "$c rc=$?"
done
python - <<'EOF'
import json
for f in ("c2", "c1", "c2_fp32", "c3", "c4", "c5"):
    try:
        j = json.loads(open("gpurun_out/v2_%s.json" % f).read().strip().splitlines()[-1])
        z = (j.get("roofline") or {}).get("zstep")
        print(f, round(j["value"], 1), round(j["e2e"]["value"], 1), (j.get("parity") or {}).get("ok"), z and round(z["ms"], 4), z and round(z["frac_of_hbm_peak"], 4))
    except Exception as e:
        print(f, "ERR", e)
EOF
