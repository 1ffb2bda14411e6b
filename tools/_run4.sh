B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-inference --no-extras"
echo "== base"; $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('phases'))"
echo "== persist_3x3=1"; HG_OPTIONS=persist_3x3=1 $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('phases'))"
echo "== persist_3x3=1 min_units 256"; HG_OPTIONS=persist_3x3=1,persist_min_units=256 $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('phases'))"
