#!/bin/bash
# One GPU-box session: GPU tests, bench A/B over environment switches, ncu of the dominant kernel.
# usage (from the repo root, through gpurun): bash tools/gpu_session.sh <tag> ["NAME:ENV=VAL,ENV=VAL" ...]
tag=${1:-run}; shift
out=gpurun_out
variants=("$@")
[ ${#variants[@]} -eq 0 ] && variants=("base:")
if [ -z "$BF_SESSION_NOTESTS" ]; then python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > $out/${tag}_tests.log; tail -5 $out/${tag}_tests.log; fi
for v in "${variants[@]}"; do
  name=${v%%:*}; envs=${v#*:}
  env $(echo $envs | tr ',' ' ') python bench.py --steps 3 --warmup 2 --no-exact --no-parity > $out/${tag}_bench_$name.log 2>&1
  python - <<PY
import json
try:
    l=[x for x in open("$out/${tag}_bench_$name.log") if x.startswith("{")][-1]; d=json.loads(l)
    r=d["roofline"]; print("$name", round(d["value"]), round(d["e2e"]["value"]), "frac %.3f last %.3f ms upd %.3f ms stage %.3f pipe %.3f" % (r["frac"], r["last_iteration"]["avg_launch_ms"], r["avg_launch_ms"], r["stage_frac"], r["pipeline_frac"]), {k: round(v,2) for k,v in r["stages_ms_timed_region"].items()})
except Exception as e: print("$name failed", e)
PY
done
if [ -n "$BF_SESSION_NCU" ]; then
ncu --set full --clock-control none --import-source on --kernel-name regex:k_blur_solve_box --launch-skip 20 --launch-count 4 \
    -o $out/${tag}_box python bench.py --steps 1 --warmup 1 --no-exact --no-parity > $out/${tag}_ncu.log 2>&1
fi
