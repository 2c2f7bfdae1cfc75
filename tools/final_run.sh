#!/bin/bash
# Round-end check on the GPU box: every GPU test, smoke, the two bench arms, the producers' timings, the frame from geometry.
TAG=${1:-r1g}
python -m pytest tests -m gpu -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
python tools/raster_bench.py > gpurun_out/${TAG}_raster_bench.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --workload mesh4k --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_mesh4k.json 2>> gpurun_out/${TAG}_bench.err
python - <<P
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], {k:round(v['ms_per_frame'],3) for k,v in d['stages'].items()}, d['producers'])
print(open('gpurun_out/${TAG}_raster_bench.json').read().strip())
m=json.load(open('gpurun_out/${TAG}_bench_mesh4k.json')); print('mesh4k', m['value'], m['ms_per_step'])
P
