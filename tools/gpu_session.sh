#!/bin/bash
# One GPU-box session: GPU tests, bench A/B over environment switches, ncu of the dominant kernel.
# usage (from the repo root, through gpurun): bash tools/gpu_session.sh <tag> ["NAME:ENV=VAL,ENV=VAL" ...]
tag=${1:-run}; shift
out=gpurun_out
variants=("$@")
[ ${#variants[@]} -eq 0 ] && variants=("base:")
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > $out/${tag}_tests.log
tail -5 $out/${tag}_tests.log
for v in "${variants[@]}"; do
  name=${v%%:*}; envs=${v#*:}
  env $(echo $envs | tr ',' ' ') python bench.py --steps 3 --warmup 2 --no-exact --no-parity > $out/${tag}_bench_$name.log 2>&1
  python - <<PY
import json
try:
    l=[x for x in open("$out/${tag}_bench_$name.log") if x.startswith("{")][-1]; d=json.loads(l)
    r=d["roofline"]; print("$name", round(d["value"]), round(d["e2e"]["value"]), "frac %.3f last %.3f ms upd %.3f ms stage %.3f pipe %.3f" % (r["frac"], r["last_iteration"]["avg_launch_ms"], r["avg_launch_ms"], r["stage_frac"], r["pipeline_frac"]), {k: round(v,2) for k,v in r["stages_ms_per_step"].items()})
except Exception as e: print("$name failed", e)
PY
done
ncu --set full --clock-control none --import-source on --kernel-name regex:k_blur_solve_box --launch-skip 20 --launch-count 4 \
    -o $out/${tag}_box python bench.py --steps 1 --warmup 1 --no-exact --no-parity > $out/${tag}_ncu.log 2>&1
BTCSFLOW_PAIR_GROUP=8 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --kernel-name regex:"k_blur_solve_box|k_update" --launch-skip 24 --launch-count 8 --csv --log-file $out/${tag}_g8_dram.csv \
    python bench.py --steps 1 --warmup 1 --no-exact --no-parity > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --kernel-name regex:"k_blur_solve_box|k_update" --launch-skip 24 --launch-count 8 --csv --log-file $out/${tag}_g1_dram.csv \
    python bench.py --steps 1 --warmup 1 --no-exact --no-parity > /dev/null 2>&1
