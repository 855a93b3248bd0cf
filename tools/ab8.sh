for m in peers; do echo "== $m"
LZS_B200_GATHER=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 2 --warmup 3 --no-e2e > gpurun_out/t20_n8_$m.json 2> gpurun_out/t20_n8_$m.err; grep -iE "Error" gpurun_out/t20_n8_$m.err | head -3; python -c "
import json; d=json.loads(open('gpurun_out/t20_n8_$m.json').read().strip().splitlines()[-1]); print(d['value'], d['value_with_gather'], d['gather']['ms'], d['gather']['mode'], d['gather']['bus_gbs_per_rank_in'])"
done
