for cfg in "444 256" "296 256" "296 64" "296 32" "148 32" "444 32"; do
  set -- $cfg
  LICHEE_LEAF_GRID=$1 LICHEE_LEAF_RUN=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /tmp/b.json 2>/dev/null
  echo "grid $1 run $2: $(python - <<'PY'
import json
for l in open('/tmp/b.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()})
PY
)"
done
