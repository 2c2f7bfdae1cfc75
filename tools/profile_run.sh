#!/bin/bash
# One profiling pass on the GPU box (run under gpurun): smoke, bench lines, ncu launch list and full captures -> gpurun_out/
set -u
TAG=${1:-r2}
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
python tools/raster_bench.py > gpurun_out/${TAG}_raster_bench.json 2>> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-producers > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ssr_|glossy|ssao|deferred|reconstruct" -c 15 -f -o gpurun_out/prof_${TAG} python tools/stage_bench.py --iters 1 --views 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
tail -c 300 gpurun_out/${TAG}_bench.json
