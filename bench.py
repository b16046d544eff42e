#!/usr/bin/env python
"""bench.py -- candidate lineage trees checked+scored per second on the headline workload.

Workload (BASELINE.json configs[4], the largest single-GPU configuration and the one the
north-star speed-up is quoted on): synthetic 64-sample, 120-cluster constraint network built with
the reference's own edge rule under -c (complete edges), -minClusterSize 1, -e 0.1, -s 1000.
One "step" = one whole tree search of that network: enumerate spanning trees in canonical order,
sum-rule check every candidate extension, score every complete valid tree, keep the top 1000.
Parameters.MAX_NUM_TREES is raised from the reference's 100000 to --max-trees so that a step is
long enough to measure a B200 (with the reference's default a step is ~1 ms of launch latency).
The step runs in CANONICAL order (ref_order_budget = -1): on this network the reference's own search spends its
full 10^8 grow() calls and finds no tree (literal port, tests/golden/reference_caps.json; the library's
reference-order replay reproduces that when it is given the budget, tests/test_gpu_bench_config.py, opt-in) -- the
default mode gives up after its replay budget and falls back to exactly this search.  The reference-order mode is measured on the networks where the
reference itself finishes: `bundled_datasets_build_search_ms` (identical output, checked in the run).

  top_list_sha256     : digest of the ranked top-s list; every timed step must produce the same one, and the
                        120x64 / 40x20 digests at the default cap are pinned in tests/golden/canonical_digests.json
  same_output_baseline: the sequential canonical specification (CPU, one core) on the same network with the same
                        caps, output bit-identical (digests compared in the run), next to the GPU end-to-end time

  value : candidate extensions checked per second, whole job, network already resident in HBM
  e2e   : same, through the public API with HOST buffers: create (H2D) + run + ranked trees D2H
  roofline / cpu_baseline / clocks / gpu_launches : see DESIGN.md "Measurement"

N > 1 (torchrun, one rank per GPU): `value` stays the headline workload with the frontier cut at the first
tree levels into N contiguous slices of the canonical order, each rank searching its slice under its own cap
(per-GPU work fixed: weak scaling) and the N ranked lists merged with one NCCL all_gather.  `strong_scaling`
is ONE exhaustive search shared by the N GPUs with dynamic work stealing (lichee_b200/dist.py), timed next to the
same search on one GPU in the same run; the merged digest must equal the single-GPU digest.

--impl reference: the literal CPU port of the reference search (oracle/, the JVM is absent from
this image) on the same network, each step a bounded sample of grow() calls.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "synthetic 64-sample 120-cluster complete network (-c, -minClusterSize 1, -e 0.1, -s 1000)"


def make_workload(args):
    from lichee_b200.synth import make_instance
    net, prm = make_instance(args.clusters, args.samples, seed=args.seed)
    return net.adj, np.ascontiguousarray(net.aaf, dtype=np.float64), prm.vaf_error_margin, net.num_edges


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def ncu_note():
    """Issue-slot and occupancy figures of the hot kernels from the committed ncu summaries (profiles/r02_*_ncu_full.csv,
    `ncu --set full` of the same binary): what binds the kernels when it is not HBM."""
    import csv
    def pick(name, metric, col=-2):
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            return None
        for r in csv.reader(open(path)):
            if r and r[0] == metric:
                return r[1:-1]
        return None
    parts = []
    for label, f in (("leaf_kernel (a burst / a second burst / a short ordinary chunk)", "r02_leaf_kernel_ncu_full.csv"),
                     ("check2_kernel", "r02_check2_kernel_ncu_full.csv"), ("replay_kernel (one warp)", "r02_replay_kernel_ncu_full.csv")):
        iss = pick(f, "smsp__issue_active.avg.pct_of_peak_sustained_active")
        wrp = pick(f, "sm__warps_active.avg.pct_of_peak_sustained_active")
        drm = pick(f, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
        if iss and wrp and drm:
            parts.append(f"{label}: issue slots active {'/'.join(f'{float(x):.0f}' for x in iss)} %, warps active "
                         f"{'/'.join(f'{float(x):.0f}' for x in wrp)} %, DRAM {'/'.join(f'{float(x):.1f}' for x in drm)} % of peak")
    return ("integer/byte frontier work: the kernels are bound by instruction issue and dependent-instruction latency, not by HBM "
            "(ncu --set full, profiles/r02_*_ncu_full.csv): " + "; ".join(parts) + ". DRAM bytes of a launch = its algorithmic bytes; "
            "`traffic` is the DRAM byte count of ONE captured burst launch (profiles/traffic.json).")


def bundled_search_times(lichee_b200, with_cpu):
    """-build search wall time on the bundled ccRCC / HGSC networks (BASELINE configs[0..2], README
    settings, committed fixtures): end to end through the public API from host arrays, median of 5, in the
    library's DEFAULT mode (reference order: replay + reference-order scoring) and in canonical mode, next to
    the CPU port of the reference search; `identical` = ranked list (scores bit for bit, parents, treeNodes
    order) equal to the literal port's, checked here.  These searches are a handful of trees: they measure
    launch latency, not throughput."""
    path = os.path.join(ROOT, "tests", "golden", "networks_bundled.json")
    if not os.path.exists(path):
        return None
    out = {}
    for e in json.load(open(path)):
        if e["clustering"] != "single" or e["complete_edges"]:
            continue
        adj, vaf = e["adj"], np.array(e["aaf"])
        save = 100 if "hgsc" in e["dataset"] else 1
        rec = {"nodes": len(adj)}
        for mode, budget in (("gpu_e2e_ms", 0), ("gpu_canonical_e2e_ms", -1)):
            ts = []
            for _ in range(6):
                t0 = time.perf_counter()
                s = lichee_b200.LineageSearch(adj, vaf, e["vaf_error_margin"], num_save=save, ref_order_budget=budget)
                st = s.run(); trees = s.trees(); s.close()
                ts.append(1e3 * (time.perf_counter() - t0))
            rec[mode] = round(statistics.median(ts[1:]), 3)
            if budget == 0:
                rec["valid_trees"] = int(st.num_trees)
                rec["grow_calls"] = int(st.num_grow_calls)
                rec["reference_order"] = not (st.status & lichee_b200.STATUS_CANONICAL_ORDER)
                ref_trees = trees
        if with_cpu:
            import oracle
            t0 = time.perf_counter()
            r = oracle.search(adj, vaf, e["vaf_error_margin"], top_s=save)
            rec["cpu_port_ms"] = round(1e3 * (time.perf_counter() - t0), 3)
            rec["identical"] = bool(len(ref_trees) == len(r["scores"]) and
                                    all(t.errorScore == r["scores"][k] and t.parent == list(r["parents"][k]) and
                                        t.treeNodes == list(r["orders"][k]) for k, t in enumerate(ref_trees)))
        out[e["dataset"]] = rec
    return out


def same_output_baseline(lichee_b200, args):
    """The sequential canonical specification (oracle/, one core) against the GPU on the SAME network with the
    SAME caps (the reference's defaults: MAX_NUM_TREES 100000), output bit-identical: the clean GPU-vs-CPU ratio
    for identical work.  The literal reference port is beside it for scale: it is a different traversal
    (tests/golden/reference_caps.json: 40x20 reaches the cap after 3.5e7 grow() calls / 57 s, 120x64 finds no tree
    in 10^8 calls / 976 s)."""
    import oracle
    from lichee_b200.synth import make_instance
    out = {}
    for (c, smp, save) in ((120, 64, 1000), (40, 20, 1000)):
        net, prm = make_instance(c, smp, seed=args.seed)
        vaf = np.ascontiguousarray(net.aaf, dtype=np.float64)
        t0 = time.perf_counter()
        r = oracle.canonical_search(net.adj, vaf, prm.vaf_error_margin, top_s=save, max_trees=100000)
        cpu_s = time.perf_counter() - t0
        import hashlib
        hh = hashlib.sha256()
        hh.update(np.ascontiguousarray(r["scores"], dtype=np.float64).tobytes())
        hh.update(np.ascontiguousarray(r["seq"], dtype=np.int64).tobytes())
        hh.update(np.ascontiguousarray(r["parents"], dtype=np.int32).tobytes())
        ts, digest = [], None
        for _ in range(6):
            t0 = time.perf_counter()
            s = lichee_b200.LineageSearch(net.adj, vaf, prm.vaf_error_margin, num_save=save, max_num_trees=100000, ref_order_budget=-1)
            st = s.run(); trees = s.trees(); s.close()
            ts.append(time.perf_counter() - t0)
            digest = lichee_b200.top_list_sha256(trees)
        gpu_s = statistics.median(ts[1:])
        out[f"{c}x{smp} default caps (-s {save})"] = {
            "valid_trees": int(st.num_trees), "cpu_canonical_spec_ms": round(1e3 * cpu_s, 3), "cpu_cores": 1,
            "gpu_e2e_ms": round(1e3 * gpu_s, 3), "ratio": round(cpu_s / gpu_s, 1), "top_list_sha256": digest,
            "identical": digest == hh.hexdigest()}
    return out


def named_shape_times(lichee_b200, adj, vaf, margin, args):
    """Search wall time, end to end from host arrays (median of 5), of (a) the headline network with the
    REFERENCE'S OWN caps (MAX_NUM_TREES 100000, MAX_NUM_GROW_CALLS 1e8, -s 1000): what `-build` does
    with this library in place of getLineageTrees; (b) BASELINE configs[3], 40 clusters x 20 samples,
    -c, with the default caps and with the tree cap raised to 2^28."""
    from lichee_b200.synth import make_instance
    def timed(adj, vaf, margin, save, max_trees, max_grow=100000000):
        ts, st = [], None
        for _ in range(6):
            t0 = time.perf_counter()
            s = lichee_b200.LineageSearch(adj, vaf, margin, num_save=save, max_num_trees=max_trees, max_num_grow_calls=max_grow,
                                          ref_order_budget=-1)
            st = s.run(); s.trees(); s.close()
            ts.append(1e3 * (time.perf_counter() - t0))
        return {"e2e_ms": round(statistics.median(ts[1:]), 3), "valid_trees": int(st.num_trees), "checks": int(st.num_checks),
                "status": int(st.status), "launches": int(st.num_launches), "device_ms": round(float(st.kernel_ms_total), 3)}
    out = {"120x64 default caps (-s 1000)": timed(adj, vaf, margin, args.save, 100000)}
    net, prm = make_instance(40, 20, seed=args.seed)
    v40 = np.ascontiguousarray(net.aaf, dtype=np.float64)
    out["40x20 default caps (-s 100)"] = timed(net.adj, v40, prm.vaf_error_margin, 100, 100000)
    out["40x20 max_trees 2^28, no grow-call cap (-s 100)"] = timed(net.adj, v40, prm.vaf_error_margin, 100, 1 << 28, 1 << 62)
    return out


def strong_scaling_leg(lichee_b200, torch, dist, rank, world, local, args):
    """ONE exhaustive search (no cap binds) shared by all N GPUs with dynamic work stealing and one all_gather
    (lichee_b200/dist.py), next to the same search on one GPU, timed in the same run: the merged top-s digest must
    equal the single-GPU digest.  19 clusters x 9 samples, -c: 3.33e11 valid trees."""
    from lichee_b200.dist import run_work_stealing
    from lichee_b200.synth import make_instance
    net, prm = make_instance(args.strong_clusters, args.strong_samples, seed=args.strong_seed)
    vaf = np.ascontiguousarray(net.aaf, dtype=np.float64)
    store = dist.distributed_c10d._get_default_store()
    s = lichee_b200.LineageSearch(net.adj, vaf, prm.vaf_error_margin, num_save=args.save, max_num_trees=1 << 60,
                                  max_num_grow_calls=1 << 62, device=local, ref_order_budget=-1)
    n_slices = args.slices_per_gpu * world
    dev = f"cuda:{local}"

    def shared():
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        info = run_work_stealing(s, store, n_slices, key="bench/strong", device=dev)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt.item()), info

    shared()                                            # warm-up: level buffers come from the pool afterwards
    t_n, info = shared()
    digest_n = lichee_b200.top_list_sha256(s.trees(), with_seq=False)    # (sequence numbers are per slice)
    mine = {"rank": rank, "slices": len(info["slices_claimed"]), "search_ms": round(1e3 * info["search_s"], 2),
            "merge_ms": round(1e3 * info["merge_s"], 2), "redone": info["slices_redone"]}
    per_rank = [None] * world
    dist.all_gather_object(per_rank, mine)
    t_1, digest_1, trees_1 = 0.0, None, 0
    if rank == 0:                                       # the same search on one GPU (the others wait at the barrier)
        s.run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = s.run()
        torch.cuda.synchronize()
        t_1 = time.perf_counter() - t0
        digest_1 = lichee_b200.top_list_sha256(s.trees(), with_seq=False)
        trees_1 = int(st.num_trees)
    dist.barrier()
    s.close()
    if rank != 0:
        return None
    if digest_1 != digest_n or trees_1 != info["num_trees"]:
        raise SystemExit(f"bench.py: the {world}-GPU search ({info['num_trees']} trees, {digest_n}) differs from the "
                         f"single-GPU search ({trees_1} trees, {digest_1})")
    return {"workload": f"synthetic {args.strong_samples}-sample {args.strong_clusters}-cluster complete network, exhaustive "
                        f"(no cap binds), -s {args.save}: ONE search, {n_slices} slices claimed dynamically, one all_gather",
            "scaling": "strong", "valid_trees": trees_1, "seconds_1gpu": t_1, "seconds_ngpu": t_n, "speedup": t_1 / t_n,
            "efficiency": t_1 / t_n / world, "top_list_sha256": digest_n, "identical_to_1gpu": True, "per_rank": per_rank}


def run_reference(args):
    """Reference arm: literal CPU port of PHYNetwork.grow on the box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    adj, vaf, margin, n_edges = make_workload(args)
    sample_grow = args.ref_grow_calls
    def step():
        t0 = time.perf_counter()
        r = oracle.search(adj, vaf, margin, top_s=args.save, max_trees=100000, max_grow=sample_grow, mode=0)
        return time.perf_counter() - t0, r
    for _ in range(args.warmup):
        step()
    tot_t = tot_c = 0
    for _ in range(args.steps):
        dt, r = step()
        tot_t += dt; tot_c += r["n_checks"]
    val = tot_c / tot_t
    line = {
        "impl": "reference", "metric": "candidate lineage trees checked+scored per second", "value": val,
        "unit": "candidate trees/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "clusters": args.clusters, "samples": args.samples, "edges": n_edges,
                   "seed": args.seed, "num_save": args.save, "max_num_trees_per_gpu": args.max_trees,
                   "parallelism": "single host thread (the reference search is single-threaded)",
                   "order": "the reference's own Gabow-Myers traversal; a bounded sample of it per step: the full search "
                            "of this network runs into MAX_NUM_GROW_CALLS = 10^8 without a valid tree (976 s, "
                            "tests/golden/reference_caps.json)",
                   "valid_trees_per_step": int(r["n_trees"])},
        "cpu_baseline": {"value": val, "unit": "candidate trees/s", "cores": 1, "kind": "port",
                         "sample": f"first {sample_grow} grow() calls of the reference's Gabow-Myers search per step "
                                   f"(literal C port of PHYNetwork.grow; the search is single-threaded in the reference; no JVM in this image)"},
        "e2e": {"value": val, "unit": "candidate trees/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clusters", type=int, default=120)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--save", type=int, default=1000)
    ap.add_argument("--max-trees", type=int, default=1 << 30)
    ap.add_argument("--level-capacity", type=int, default=0)
    ap.add_argument("--ref-grow-calls", type=int, default=1500000)
    ap.add_argument("--cpu-baseline-grow-calls", type=int, default=5000000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strong-clusters", type=int, default=19)
    ap.add_argument("--strong-samples", type=int, default=9)
    ap.add_argument("--strong-seed", type=int, default=1)
    ap.add_argument("--slices-per-gpu", type=int, default=16)
    ap.add_argument("--no-strong", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import lichee_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lichee_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus or world == 1, (world, args.gpus)

    adj, vaf, margin, n_edges = make_workload(args)
    kw = dict(num_save=args.save, max_num_trees=args.max_trees, max_num_grow_calls=1 << 62, device=local,
              part_rank=rank, part_count=world, level_capacity=args.level_capacity, ref_order_budget=-1)

    # pinned host staging of the step's inputs (the arrays the Java host would hand over)
    off = np.zeros(len(adj) + 1, dtype=np.int32)
    for i, a in enumerate(adj):
        off[i + 1] = off[i] + len(a)
    idx = np.array([c for a in adj for c in a], dtype=np.int32)
    h2d_bytes = vaf.nbytes + off.nbytes + idx.nbytes
    # the e2e path copies the network from PINNED host memory every step
    vaf_pinned = torch.from_numpy(vaf).pin_memory()
    vaf = vaf_pinned.numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    resident = lichee_b200.LineageSearch(adj, vaf, margin, **kw)
    rb = resident.record_bytes()
    gather_in = torch.empty(rb * args.save, dtype=torch.uint8, device="cuda")
    gather_out = torch.empty(world * rb * args.save, dtype=torch.uint8, device="cuda") if world > 1 else None

    def merge(search):
        if world > 1:
            search.export_ranked(gather_in.data_ptr())
            dist.all_gather_into_tensor(gather_out, gather_in)
            torch.cuda.synchronize()
            search.merge_ranked(gather_out.data_ptr(), world, device_ptr=True)

    phase_ms = {"search": 0.0, "merge": 0.0}

    def step_resident():
        t0 = time.perf_counter()
        st = resident.run()
        t1 = time.perf_counter()
        merge(resident)
        phase_ms["search"] += 1e3 * (t1 - t0)
        phase_ms["merge"] += 1e3 * (time.perf_counter() - t1)
        return st

    def step_e2e():
        s = lichee_b200.LineageSearch(adj, vaf, margin, **kw)          # H2D of the network
        st = s.run()
        merge(s)
        trees = s.trees()                                               # D2H of the ranked list
        s.close()
        return st, trees

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    phase_ms["search"] = phase_ms["merge"] = 0.0
    t0 = time.perf_counter()
    stats = [step_resident() for _ in range(args.steps)]
    barrier()
    t_dev = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    digests = {lichee_b200.top_list_sha256(resident.trees())}

    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_out = [step_e2e() for _ in range(args.steps)]
    barrier()
    t_e2e = time.perf_counter() - t0
    digests |= {lichee_b200.top_list_sha256(trees) for _, trees in e2e_out}
    if len(digests) != 1:
        raise SystemExit(f"bench.py: the timed steps disagree on the ranked list: {sorted(digests)}")
    top_digest = digests.pop()

    checks = sum(s.num_checks for s in stats)
    scored = sum(s.num_scored for s in stats)
    evaluated = sum(s.num_evaluated for s in stats)
    leaf_items = sum(s.num_leaf_items for s in stats)
    bounded_items = sum(s.num_bounded_items for s in stats)
    launches = sum(s.num_launches for s in stats) + (args.steps * 3 if world > 1 else 0)
    agg = torch.tensor([float(checks), float(scored), float(launches), float(sum(s.num_checks for s, _ in e2e_out)),
                        float(evaluated), float(leaf_items), float(bounded_items)],
                       dtype=torch.float64, device="cuda")
    tmax = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(agg)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    checks_all, scored_all, launches_all, e2e_checks_all, eval_all, leaf_all, bounded_all = [float(x) for x in agg.tolist()]
    t_dev, t_e2e = [float(x) for x in tmax.tolist()]

    per_rank = None
    if world > 1:
        # where each rank's step went: its own slice search (prologue included) and the export / all_gather / merge
        mine = {"rank": rank, "search_ms_per_step": round(phase_ms["search"] / args.steps, 3),
                "merge_ms_per_step": round(phase_ms["merge"] / args.steps, 3),
                "device_ms_per_step": round(sum(s.kernel_ms_total for s in stats) / args.steps, 3),
                "valid_trees_per_step": int(stats[-1].num_trees), "launches_per_step": int(stats[-1].num_launches)}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    strong = None
    if world > 1 and not args.no_strong:
        strong = strong_scaling_leg(lichee_b200, torch, dist, rank, world, local, args)

    if rank == 0:
        K = lichee_b200.SearchStats.KERNELS
        ms = [sum(s.kernel_ms[i] for s in stats) for i in range(5)]
        nl = [sum(s.kernel_launches[i] for s in stats) for i in range(5)]
        by = [sum(s.kernel_bytes[i] for s in stats) for i in range(5)]
        # dominant kernel: leaf_kernel (largest share of the committed ncu launch list, profiles/r02_launches_summary.txt).  The
        # check column cannot be used to pick it: its check2_kernel launches run ahead on a low-priority second stream, so
        # their event times contain the waiting for SMs the main stream holds.
        dom = K.index("leaf") if "leaf" in K else max(range(5), key=lambda i: ms[i])
        peak, peak_src = measured_peak()
        achieved = by[dom] / (ms[dom] * 1e-3) / 1e9 if ms[dom] > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(K[dom])
            except Exception:
                traffic = None
        st0 = stats[-1]
        d2h_bytes = len(e2e_out[-1][1]) * (8 + 8 + 2 * 4 * (args.clusters + 1)) + 256
        line = {
            "metric": "candidate lineage trees checked+scored per second", "value": checks_all / t_dev,
            "unit": "candidate trees/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clusters": args.clusters, "samples": args.samples, "edges": n_edges,
                       "seed": args.seed, "num_save": args.save, "max_num_trees_per_gpu": args.max_trees,
                       "parallelism": f"frontier slices x{world}" if world > 1 else "single GPU",
                       "l2": "frontier chunks are produced and consumed once per step; every level buffer is rewritten "
                             "between timed steps (working set of a step > 126 MB L2)",
                       "order": "canonical (ref_order_budget=-1): followed to its own end the reference's search spends its whole "
                                "MAX_NUM_GROW_CALLS = 10^8 on this network and finds no valid tree (literal port 976 s, "
                                "tests/golden/reference_caps.json; the library's reference-order replay 328 s with the same "
                                "result, profiles/r02_config4_reference_order.log)",
                       "valid_trees_per_step": st0.num_trees * world if world > 1 else st0.num_trees,
                       "complete_trees_scored_per_s": scored_all / t_dev},
            "top_list_sha256": top_digest,
            "what_value_counts": "sum-rule DECISIONS per second (one per candidate extension, what the reference decides with one "
                                 "checkConstraint call); most are carried over from the lexicographic neighbour, read from a "
                                 "precomputed bit, or belong to items dismissed by the lower bound -- see the three rates below",
            "valid_trees_ranked_per_s": scored_all / t_dev,
            "evaluated_trees_per_s": eval_all / t_dev,
            "bounded_fraction": {"last_level_items": bounded_all / max(leaf_all, 1.0),
                                 "valid_trees": 1.0 - eval_all / max(scored_all, 1.0)},
            "e2e": {"value": e2e_checks_all / t_e2e, "unit": "candidate trees/s", "ms_per_step": 1e3 * t_e2e / args.steps,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": K[dom] + "_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "launches": nl[dom], "avg_launch_ms": ms[dom] / max(nl[dom], 1),
                         "algorithmic_bytes_per_launch": by[dom] / max(nl[dom], 1),
                         "note": ncu_note()},
            "kernel_ms_per_step": {K[i]: ms[i] / args.steps for i in range(5)},
            "kernel_launches_per_step": {K[i]: nl[i] / args.steps for i in range(5)},
            "device_ms_per_step": sum(s.kernel_ms_total for s in stats) / args.steps,
            "timing": "value/ms_per_step: host clock between (barrier + cuda synchronize) pairs around the K steps, max over ranks "
                      "(it contains everything, host round trips and the merge included); device_ms_per_step and kernel_ms_per_step: "
                      "CUDA events on the library's own streams (rank 0); the check column includes the check2_kernel launches that "
                      "run ahead on the second stream BESIDE the other columns, so the columns add up to more than the step",
        }
        if per_rank is not None:
            line["per_rank"] = per_rank
        if strong is not None:
            line["strong_scaling"] = strong
        if world == 1:
            # the same search with the lower bound switched off: every valid tree is scored
            os.environ["LICHEE_NO_PRUNE"] = "1"
            nop = lichee_b200.LineageSearch(adj, vaf, margin, **dict(kw, max_num_trees=min(args.max_trees, 1 << 28)))
            del os.environ["LICHEE_NO_PRUNE"]
            nop.run()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nst = nop.run()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            nop.close()
            line["no_prune"] = {"max_num_trees": min(args.max_trees, 1 << 28), "ms_per_step": 1e3 * dt,
                                "evaluated_trees_per_s": nst.num_evaluated / dt, "checks_per_s": nst.num_checks / dt,
                                "note": "LICHEE_NO_PRUNE=1: same search, lower bound off, every valid tree scored"}
            line["bundled_datasets_build_search_ms"] = bundled_search_times(lichee_b200, not args.no_cpu_baseline)
            line["named_shapes"] = named_shape_times(lichee_b200, adj, vaf, margin, args)
            if not args.no_cpu_baseline:
                line["same_output_baseline"] = same_output_baseline(lichee_b200, args)
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            t0 = time.perf_counter()
            r = oracle.search(adj, vaf, margin, top_s=args.save, max_trees=100000, max_grow=args.cpu_baseline_grow_calls, mode=0)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {
                "value": r["n_checks"] / dt, "unit": "candidate trees/s", "cores": 1, "kind": "port",
                "seconds": dt, "valid_trees_found": r["n_trees"],
                "sample": f"first {args.cpu_baseline_grow_calls} grow() calls of the reference's Gabow-Myers search on the same "
                          f"network (literal C port of PHYNetwork.grow/PHYTree.checkConstraint; single-threaded like the reference)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
