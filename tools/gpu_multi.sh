#!/bin/bash
# Multi-GPU measurements on one box: bash tools/gpu_multi.sh <tag> <N>
tag=$1; n=$2; out=gpurun_out
if [ "$n" = "2" ]; then python -m pytest tests/test_gpu_distributed.py -m gpu -q -rA > $out/${tag}_nccl_test.log 2>&1; grep -E "PASSED|FAILED|SKIPPED|passed|failed|skipped" $out/${tag}_nccl_test.log | tail -4; fi
python bench.py --gpus $n --steps 10 --warmup 3 --no-exact --no-parity > $out/${tag}_weak_n$n.json 2> $out/${tag}_weak_n$n.err; tail -c 300 $out/${tag}_weak_n$n.err
python bench.py --gpus $n --scaling strong --clips 64 --frames 300 --steps 2 --warmup 1 > $out/${tag}_strong_n$n.json 2> $out/${tag}_strong_n$n.err; tail -c 300 $out/${tag}_strong_n$n.err
python - <<PY
import json
for k in ("weak", "strong"):
    try:
        d = json.loads([l for l in open("$out/${tag}_%s_n$n.json" % k) if l.startswith("{")][-1])
        print(k, "N=$n", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 2), d["scaling"])
    except Exception as e:
        print(k, "failed", e)
PY
