#!/usr/bin/env python
"""Randomized fuzz of the CUDA path: python scripts/fuzz_gpu.py <seed> <seconds>.
  canonical mode        against the sequential canonical specification (oracle.canonical_search): bit-exact
                        tree counts, ranked trees, scores, seq numbers and check / grow counts
  reference-order mode  (every third instance, smaller) against the literal port of the reference search
                        (oracle.search): bit-exact counts, ranked trees, treeNodes orders, scores, discovery indices
Instances that need more than 3e6 accepted partial trees are skipped (too slow on the CPU side)."""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, lichee_b200, oracle
from test_gpu_parity import rand_dag
from lichee_b200.synth import make_instance
seed0 = int(sys.argv[1]); budget = float(sys.argv[2])
rng = random.Random(seed0); t0 = time.time(); n_inst = 0; n_trees = 0; bad = 0; n_ref = 0
while time.time() - t0 < budget:
    kind = rng.random()
    if kind < 0.6:
        n = rng.randint(2, 30); S = rng.choice([1, 2, 3, 4, 7, 8, 15, 16, 31, 32, 33, 40, 63, 64])
        adj, vaf = rand_dag(rng, n, S, rng.choice([0.1, 0.2, 0.4, 0.7, 1.0]), rng.choice([0.02, 0.08, 0.2, 0.4, 0.6]))
        margin = rng.choice([0.0, 0.001, 0.05, 0.1, 0.3])
    else:
        net, prm = make_instance(rng.randint(3, 60), rng.choice([2, 5, 12, 20, 33, 64]), rng.randint(0, 10**6),
                                 noise=rng.choice([0.0, 0.005, 0.02]), complete=rng.random() < 0.7)
        adj, vaf, margin = net.adj, np.array(net.aaf), rng.choice([0.02, 0.1])
    cap = rng.choice([1, 2, 50, 3000, 60000, 400000])
    save = rng.choice([1, 3, 64, 1000, 1024])
    kw = {}
    if rng.random() < 0.5: kw['level_capacity'] = rng.choice([1024, 1500, 4096, 50000])
    if n_inst % 3 == 2 and len(adj) <= 14:
        # reference-order mode against the literal port (small: the literal port's search is the slow side)
        cap2, grow2 = rng.choice([1, 50, 3000, 100000]), rng.choice([100, 20000, 2000000])
        s = lichee_b200.LineageSearch(adj, vaf, margin, num_save=save, max_num_trees=cap2, max_num_grow_calls=grow2, ref_order_budget=1 << 22)
        st = s.run(); tr = s.trees(); s.close()
        ref = oracle.search(adj, vaf, margin, top_s=save, max_trees=cap2, max_grow=grow2)
        ok = not (st.status & 4) and st.num_trees == ref['n_trees'] and len(tr) == len(ref['scores'])
        if ok:
            for k, t in enumerate(tr):
                if (t.errorScore != ref['scores'][k] or t.parent != list(ref['parents'][k]) or t.treeNodes != list(ref['orders'][k])
                        or t.seq != ref['dfs_index'][k]): ok = False; break
        if ok and not ref['stopped']: ok = st.num_checks == ref['n_checks'] and st.num_grow_calls == ref['n_grow']
        n_inst += 1; n_ref += 1; n_trees += st.num_trees
        if not ok:
            bad += 1
            print('MISMATCH (reference order)', n_inst, len(adj), vaf.shape, margin, cap2, grow2, save, st.num_trees, ref['n_trees'], flush=True)
            if bad > 3: break
        continue
    s = lichee_b200.LineageSearch(adj, vaf, margin, num_save=save, max_num_trees=cap, max_num_grow_calls=3000000, ref_order_budget=-1, **kw)
    st = s.run(); tr = s.trees(); s.close()
    if st.status & 2: continue                        # too heavy for the sequential CPU specification
    ref = oracle.canonical_search(adj, vaf, margin, top_s=save, max_trees=cap)
    ok = st.num_trees == ref['n_trees'] and len(tr) == len(ref['scores'])
    if ok:
        for k, t in enumerate(tr):
            if t.errorScore != ref['scores'][k] or t.parent != list(ref['parents'][k]) or t.seq != ref['seq'][k]: ok = False; break
    if ok and not (st.status & 3): ok = st.num_checks == ref['n_checks'] and st.num_grow_calls == ref['n_grow']
    n_inst += 1; n_trees += st.num_trees
    if n_inst % 25 == 0: print('seed', seed0, 'progress', n_inst, 'instances', n_trees, 'trees', bad, 'mismatches', round(time.time() - t0), 's', flush=True)
    if not ok:
        bad += 1
        print('MISMATCH', n_inst, len(adj), vaf.shape, margin, cap, save, kw, st.num_trees, ref['n_trees'], flush=True)
        os.makedirs('gpurun_out', exist_ok=True)
        np.save(f'gpurun_out/fuzz_bad_{seed0}_{n_inst}_vaf.npy', vaf); open(f'gpurun_out/fuzz_bad_{seed0}_{n_inst}_adj.txt','w').write(repr((adj, margin, cap, save, kw)))
        if bad > 3: break
print('seed', seed0, 'instances', n_inst, '(reference-order:', n_ref, ') trees', n_trees, 'mismatches', bad, 'in', round(time.time() - t0), 's')
