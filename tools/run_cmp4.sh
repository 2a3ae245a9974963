show() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'GFLOPS', round(d['ms_per_step'],3),'ms', d['stages_ms'], d['bins_ms'], (d['roofline_step'] or {}).get('frac'))"; }
for wl in "$@"; do for mode in off on; do echo "== $wl direct=$mode"; BHB200_DIRECT=$mode timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --steps 5 2>&1 | tail -1 | show; done; done
