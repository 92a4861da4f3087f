#!/bin/bash
# Final measurements of a round on one GPU: GPU tests, default bench line (C2, with parity and exact plan), the other configs,
# strong scaling at N = 1 and the launch list.  The ncu --set full captures are tools/gpu_final_ncu.sh (a separate call: gpurun
# brings back at most 64 MiB).
tag=${1:-final}
out=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $out/${tag}_tests.log; tail -3 $out/${tag}_tests.log
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_ref.json 2> $out/${tag}_ref.err
python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 600 $out/${tag}_bench.err
for c in C1 C4 C5; do python bench.py --config $c --steps 5 --warmup 3 > $out/${tag}_bench_$c.json 2> $out/${tag}_bench_$c.err; tail -c 300 $out/${tag}_bench_$c.err; done
python bench.py --scaling strong --clips 64 --frames 300 --steps 2 --warmup 1 > $out/${tag}_strong1.json 2> $out/${tag}_strong1.err; tail -c 300 $out/${tag}_strong1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-exact --no-parity > /dev/null 2>&1
python - <<PY
import json
for n in ("bench", "bench_C1", "bench_C4", "bench_C5", "strong1", "ref"):
    try:
        d = json.loads([l for l in open("$out/${tag}_%s.json" % n) if l.startswith("{")][-1])
        r = d.get("roofline") or {}
        print(n, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "pipe", r.get("pipeline_frac"), "frac", r.get("frac"), "stage", r.get("stage_frac"))
        if d.get("parity"): print("   parity", json.dumps({k: d["parity"][k] for k in ("compact", "exact")}))
        if d.get("exact_f32_storage"): print("   exact", round(d["exact_f32_storage"]["value"], 1))
        if d.get("cpu_baseline"): print("   cpu", round(d["cpu_baseline"]["value"], 2), d["cpu_baseline"].get("default_threading"))
    except Exception as e:
        print(n, "failed", e)
PY
