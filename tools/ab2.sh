for m in multicast peers nccl; do echo "== $m"
LZS_B200_GATHER=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e > gpurun_out/t18_n2_$m.json 2> gpurun_out/t18_n2_$m.err; tail -4 gpurun_out/t18_n2_$m.err | cut -c1-300; python -c "
import json; d=json.loads(open('gpurun_out/t18_n2_$m.json').read().strip().splitlines()[-1]); print(d['value'], d['value_with_gather'], d['gather']['ms'], d['gather']['mode'], d['gather']['bus_gbs_per_rank_in'])"
done
