python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s6d_bench_n2.json 2> gpurun_out/s6d_bench_n2.err; tail -c 300 gpurun_out/s6d_bench_n2.err; python -c "
import json;d=json.loads(open('gpurun_out/s6d_bench_n2.json').read().strip().splitlines()[-1]);print('N2 views4k', d['value'], d['e2e']['value'], d['e2e']['numa_node'], d['cpu_baseline'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | tail -c 400
nvidia-smi topo -m | head -12; lscpu | grep -i "numa\|socket\|^CPU(s)"
