python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 2 --warmup 3 --no-e2e > gpurun_out/t17_n8.json 2> gpurun_out/t17_n8.err; tail -3 gpurun_out/t17_n8.err; python -c "
import json; d=json.loads(open('gpurun_out/t17_n8.json').read().strip().splitlines()[-1]); print(d['value'], d['value_with_gather'], d['gather'])"
