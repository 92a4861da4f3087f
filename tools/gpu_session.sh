#!/bin/bash
# One GPU-box session: variant probe, GPU tests, bench A/B over the CTA order, ncu of the dominant kernel.
# usage (from the repo root, through gpurun): bash tools/gpu_session.sh <tag>
tag=${1:-run}
out=gpurun_out
python tools/variant_probe.py > $out/${tag}_probe.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -40 > $out/${tag}_tests.log
for g in 1 4 8 16; do
  BTCSFLOW_PAIR_GROUP=$g python bench.py --steps 3 --warmup 2 --no-exact --no-parity > $out/${tag}_bench_g$g.log 2>&1
done
ncu --set full --clock-control none --import-source on --kernel-name regex:k_blur_solve_box --launch-skip 20 --launch-count 4 \
    -o $out/${tag}_box python bench.py --steps 1 --warmup 1 --no-exact > $out/${tag}_ncu.log 2>&1
echo "probe OK: $(grep -c '^OK' $out/${tag}_probe.log)"; grep FAIL $out/${tag}_probe.log | head
tail -8 $out/${tag}_tests.log
for g in 1 4 8 16; do python - <<PY
import json
try:
    l=[x for x in open("$out/${tag}_bench_g$g.log") if x.startswith("{")][-1]; d=json.loads(l)
    r=d["roofline"]; print("group $g", round(d["value"]), round(d["e2e"]["value"]), "frac %.3f last %.3f ms upd %.3f ms stage %.3f pipe %.3f" % (r["frac"], r["last_iteration"]["avg_launch_ms"], r["avg_launch_ms"], r["stage_frac"], r["pipeline_frac"]), {k: round(v,2) for k,v in r["stages_ms_per_step"].items()})
except Exception as e: print("group $g failed", e)
PY
done
