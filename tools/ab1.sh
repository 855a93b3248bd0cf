timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'per kernel',d['roofline']['per_kernel_ms'],'e2e',d['e2e']['value'],'parity',d['parity']['mismatches'])"
