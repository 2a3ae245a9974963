show() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'GFLOPS', round(d['ms_per_step'],3),'ms', d['stages_ms'], d['bins_ms'], (d['roofline_step'] or {}).get('frac'))"; }
for wl in poisson27 poisson27thin; do echo "== $wl range=off"; BHB200_RANGE=off timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --steps 5 2>&1 | tail -1 | show; done
