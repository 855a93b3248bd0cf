mkdir -p gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 --no-pageable > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; tail -2 gpurun_out/r2_bench_n$N.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1]); print('N', d['n_gpus'], 'value', d['value'], 'with gather', d['value_with_gather'], 'e2e', d['e2e']['value'], 'copy ceiling', d['e2e']['copy_only']['value'], d['e2e']['cpu_affinity']); print(d['gather'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 2 --warmup 3 --total-gib 16 --no-e2e > gpurun_out/r2_bench_n${N}_total16gib.json 2> gpurun_out/r2_bench_n${N}_total16gib.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n${N}_total16gib.json').read().strip().splitlines()[-1]); print('16 GiB total: value', d['value'], 'with gather', d['value_with_gather'], d['gather']['ms'], d['gather']['bus_gbs_per_rank_in'])"
