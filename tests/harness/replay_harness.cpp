// replay_harness.cpp -- TEST INFRASTRUCTURE.  Compiles the library's reference-order state machine
// (lichee_b200/csrc/lichee_replay.cuh, the source the CUDA kernel runs) for the HOST, so that the
// algorithm can be fuzzed against the literal restatement of PHYNetwork.grow in oracle/ on machines
// without a GPU.  The product library never contains or calls this.
#include <stdint.h>
#include <stdlib.h>
#include <vector>

#include "../../lichee_b200/csrc/lichee_replay.cuh"

namespace {

struct HostChecker {             // PHYTree.checkConstraint, PHYTree.java:156-174
    const double *vaf; int S; double margin;
    bool check(replay::State &s, int u, int v, int k, int prev) {
        bool ok = true;
        for (int i = 0; i < S; i++) {
            const double before = prev == replay::kNoChild ? 0.0 : s.psum[(size_t)prev * s.SP + i];
            const double sum = before + vaf[(size_t)v * S + i];
            s.psum[(size_t)k * s.SP + i] = sum;
            if (sum > vaf[(size_t)u * S + i] + margin) ok = false;
        }
        return ok;
    }
};

struct HostRecorder {
    long long cap, count; int n;
    int *orders, *parents;
    bool room() { return count < cap; }
    void write(replay::State &s, long long) {
        for (int i = 0; i < n; i++) orders[count * n + i] = s.tn[i];
        parents[count * n] = -1;
        for (int i = 1; i < n; i++) parents[count * n + s.tn[i]] = s.par[s.tn[i]];
        count++;
    }
};

}  // namespace

// adj in node space; duplicates, self loops and edges into node 0 are dropped here like the library does.
// stats: [0] trees, [1] numGrowCalls, [2] checks, [3] status (1 done, 2 tree cap, 3 grow cap), [4] compactions, [5] checks at the last level.
// Records up to `cap` trees (orders / parents: cap x n) in discovery order; `pause_every` > 0 makes the
// recorder report "full" every that many trees so that pause/resume is exercised.
extern "C" int replay_host_search(int n, int S, const int *adj_off, const int *adj_idx, const double *vaf,
                                  double margin, long long max_trees, long long max_grow, long long cap,
                                  int pause_every, int *orders, int *parents, long long *stats) {
    std::vector<int> off(n + 1, 0), idx;
    for (int u = 0; u < n; u++) {
        const int start = (int)idx.size();
        for (int k = adj_off[u]; k < adj_off[u + 1]; k++) {
            const int v = adj_idx[k];
            if (v == u || v == 0) continue;
            bool dup = false;
            for (size_t j = start; j < idx.size(); j++) if (idx[j] == v) dup = true;
            if (!dup) idx.push_back(v);
        }
        off[u + 1] = (int)idx.size();
    }
    replay::Layout L = replay::make_layout(n, S, (int)idx.size());
    std::vector<unsigned char> blob(L.total);
    replay::init_blob(blob.data(), L, off.data(), idx.data(), max_grow, max_trees);
    replay::State s = replay::bind(blob.data(), L);
    long long leaf_checks = 0;
    HostChecker chk{vaf, S, margin};
    HostRecorder rec{pause_every > 0 ? pause_every : cap, 0, n, orders, parents};
    for (;;) {
        for (;;) {
            const replay::Result r = replay::step(s, rec.room());
            if (r == replay::R_STOP) break;
            if (r == replay::R_CHECK && s.h->tcount == n) leaf_checks++;
            if (r == replay::R_CHECK) s.h->ok = chk.check(s, s.h->cu, s.h->cv, s.h->ck, s.h->cprev) ? 1 : 0;
            if (r == replay::R_RECORD) rec.write(s, s.h->num_trees);
        }
        if (s.h->status != replay::S_RUNNING) break;
        // paused: the record buffer is "full"
        if (rec.count >= cap) break;
        rec.cap = rec.count + pause_every < cap ? rec.count + pause_every : cap;
    }
    stats[0] = s.h->num_trees; stats[1] = s.h->num_grow; stats[2] = s.h->num_checks; stats[3] = s.h->status;
    stats[4] = s.h->compactions;
    stats[5] = leaf_checks;            // checkConstraint calls that add the LAST missing node
    return 0;
}
