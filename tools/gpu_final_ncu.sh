#!/bin/bash
# ncu --set full captures of a round (one GPU): the dominant kernel at the bench's batch, the other kernels of the step, and
# the Gaussian-window kernel at the finest scale of config C4.  Reports stay below gpurun's 64 MiB return limit.
tag=${1:-final}
out=gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name regex:k_blur_solve_box --launch-skip 20 --launch-count 4 \
    -o $out/${tag}_box python bench.py --steps 1 --warmup 1 --no-exact --no-parity > $out/${tag}_ncu.log 2>&1
ncu --set full --clock-control none --kernel-name regex:"k_update|k_polyexp|k_pyr|k_level0" --launch-skip 27 --launch-count 9 \
    -o $out/${tag}_others python bench.py --steps 1 --warmup 1 --no-exact --no-parity > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:k_blur_solve_gauss --launch-skip 33 --launch-count 3 \
    -o $out/${tag}_gauss python bench.py --config C4 --steps 1 --warmup 1 --no-exact --no-parity > /dev/null 2>&1
ls -la $out/*.ncu-rep
