B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-inference --no-extras"
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d.get("phases"))'
for o in "" $EXTRA_OPTS; do echo "== opts: $o"; HG_OPTIONS=$o $B 2>/dev/null | python -c "$P"; done
