mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-pageable > gpurun_out/t15_bench_n2.json 2> gpurun_out/t15_bench_n2.err; tail -2 gpurun_out/t15_bench_n2.err; python -c "
import json; d=json.loads(open('gpurun_out/t15_bench_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'], d['gather'])"
