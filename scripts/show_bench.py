"""Prints the headline fields of a bench.py output file (skips non-JSON lines such as NCCL's banner)."""
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print("value %.4g %s | ms/step %.3f | e2e ms/step %.3f | launches %s" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"].get("ms_per_step", 0), d.get("gpu_launches")))
        print("kernel ms/step", {k: round(v, 3) for k, v in d.get("kernel_ms_per_step", {}).items()})
        print("launches/step ", d.get("kernel_launches_per_step"))
        print("digest", d.get("top_list_sha256"), "| roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac")})
        for k in ("strong_scaling", "no_prune", "same_output_baseline"):
            if k in d:
                print(k, json.dumps(d[k])[:900])
