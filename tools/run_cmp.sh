# quick comparison of the configs (one JSON line each, trimmed)
show() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],1),'GFLOPS', round(d['ms_per_step'],3),'ms', d['stages_ms'], d['bins_ms'], (d['roofline_step'] or {}).get('frac'))"; }
for mode in off; do echo "== poisson27 range=$mode"; BHB200_RANGE=$mode timeout 300 python bench.py --workload poisson27 --no-e2e --no-cpu-baseline --steps 5 2>&1 | tail -1 | show; done
for wl in poisson27thin; do for mode in off small; do echo "== $wl range=$mode"; BHB200_RANGE=$mode timeout 300 python bench.py --workload $wl --no-e2e --no-cpu-baseline --steps 5 2>&1 | tail -1 | show; done; done
for wl in poisson5 rect; do echo "== $wl"; timeout 600 python bench.py --workload $wl --no-e2e --no-cpu-baseline --steps 5 2>&1 | tail -1 | show; done
