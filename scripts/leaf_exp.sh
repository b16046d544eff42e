run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /tmp/b.json 2>/dev/null; python - <<'PY'
import json
for l in open('/tmp/b.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d['kernel_launches_per_step'], d['top_list_sha256'][:10])
PY
}
run
