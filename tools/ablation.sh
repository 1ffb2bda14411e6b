# step-level A/B runs: tools/ablation.sh name ENV=... [name ENV=...]...   (one bench.py training run per pair)
B="python bench.py --no-cpu-baseline --no-inference --no-extras --steps ${STEPS:-10} --warmup 3"
while [ $# -ge 2 ]; do
  name=$1; envs=$2; shift 2
  env $envs $B > gpurun_out/abl_$name.json 2> gpurun_out/abl_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/abl_$name.json").read().strip().splitlines()[-1])
    print("$name", d["value"], d["ms_per_step"], d.get("phases"))
except Exception as e:
    print("$name failed", e)
PY
done
