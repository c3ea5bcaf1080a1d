# usage: bash tools/sweep_env.sh VAR kernel_prefix v1 v2 ...   -- bench.py (2 steps) per value, prints the kernel's ms / GB/s from the breakdown
VAR=$1; K=$2; shift 2
for v in "$@"; do
  env $VAR=$v python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-eager > /tmp/sw.json 2> /tmp/sw.err || { echo "$VAR=$v FAILED: $(tail -1 /tmp/sw.err | cut -c1-200)"; continue; }
  python - "$VAR=$v" "$K" <<'PY'
import json, sys
d = json.loads(open('/tmp/sw.json').read().strip().splitlines()[-1])
rows = [b for b in d['breakdown'] if b['kernel'].startswith(sys.argv[2])]
print(sys.argv[1], 'value', round(d['value'], 1), ' '.join(f"{b['kernel']}: {b['ms_per_step']} ms {b['gbs']} GB/s {b['tflops']} TF" for b in rows))
PY
done
