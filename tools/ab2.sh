python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e > gpurun_out/t16_n2.json 2> gpurun_out/t16_n2.err; tail -3 gpurun_out/t16_n2.err; python -c "
import json; d=json.loads(open('gpurun_out/t16_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['value_with_gather'], d['gather'])"
