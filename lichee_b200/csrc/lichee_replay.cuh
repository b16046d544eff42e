// lichee_replay.cuh -- reference-order enumeration: PHYNetwork.grow (PHYNetwork.java:443-541) as a
// resumable state machine over flat arrays.
//
// Why it exists.  The reference ranks trees with a STABLE sort (Collections.sort, PHYNetwork.java:727),
// keeps the FIRST Parameters.MAX_NUM_TREES trees it discovers (:497) and sums children in the order
// they joined the tree, so its saved top-s list depends on the order in which grow() discovers trees.
// That order is a function of the whole search history: grow() re-appends the edges it popped to the
// END of the adjacency lists (:534-538, addEdge :277-286) and restores frontier edges on TOP of the
// stack f (:502-505).  No order-free parallel search reproduces it, so the library replays the
// reference's own traversal -- one warp, the complete state in shared memory -- and hands the trees
// it discovers, in discovery order, to the parallel scoring and selection kernels.
//
// What is restated, and how each reference structure is held (all node ids are ROW indices: 0 = root):
//   t.treeNodes                 tn[0..tcount)            stack; a node always leaves in LIFO order
//   t.treeEdges.get(p)          the nodes of tn whose par[] is p, in tn order (children join and leave
//                               LIFO, so the insertion order of a child list is the tn order)
//   checkConstraint(p)          psum[k][s] = running sum of p's children up to and including tn[k]:
//                               psum[k] = psum[lastchild[p]] + VAF(tn[k])  (0.0 + VAF for a first child),
//                               the same left-to-right double additions as PHYTree.java:162-165
//   this.edges (adjacency)      adj[u][0..adjlen[u]) in list order; removeEdge shifts, addEdge appends
//   f (ArrayList used as stack) fs[0..top) with tombstones; pos[v*n+u] = index of edge (u,v) in fs or
//                               kNone.  Order of the live entries == order of the ArrayList.
//   edgesRemoved / ff           slices of two stacks (an edge sits in at most one frame's list)
//   edgesAdded                  not stored: when grow() returns, the edges of f that leave v are exactly
//                               the ones that were added (removeAll works by equals())
//   bridge test (:515-530)      L aliases the LIVE tree t (L = t, :447), and when the test runs v has just
//                               been taken out of t and has no children left, so isDescendent(v, w) is
//                               false for every w: the loop ends exactly when no edge into v is left in
//                               the adjacency lists, i.e. indeg[v] == 0 (also while L == null)
//   the two early returns       :490 (MAX_NUM_GROW_CALLS) and :497 (MAX_NUM_TREES): after either one no
//                               further tree can be recorded (a tree is only recorded at the top of a
//                               grow() call and no further call is made), so the replay simply stops
//
// Two statements of the same machine over the same state blob (Layout / Header / init_blob below):
//   * step()            the sequential statement, one thread.  tests/harness/replay_harness.cpp compiles it for
//                       the HOST and tests/test_replay_cpu.py fuzzes it against the literal restatement of
//                       PHYNetwork.grow in oracle/ (thousands of digraphs, both caps, pause/resume).  It is test
//                       infrastructure and documentation: the library has no host path.
//   * replay_kernel     (lichee_search.cu) the warp statement the device runs: the 32 lanes execute the same
//                       transitions with the scalars in registers and split every list scan -- the adjacency list
//                       of v, the edges into v, the restored / re-appended edge lists, the samples of the sum rule,
//                       the words of a recorded tree -- with ballots.  tests/test_gpu_reference_order.py checks it
//                       against the same literal restatement on the GPU.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RP_HD __host__ __device__ __forceinline__
#else
#define RP_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define RP_CTZ(x) (__ffs((int)(x)) - 1)
#else
#define RP_CTZ(x) __builtin_ctz(x)
#endif

namespace replay {

constexpr uint16_t kNone = 0xffffu;
constexpr uint8_t kNoChild = 0xffu;

enum Phase : int { P_ENTER = 0, P_LOOP = 1, P_AFTER_RETURN = 2, P_AFTER = 3, P_FINISH = 4, P_CHECKED = 5, P_RECORDED = 6 };
enum Status : int { S_RUNNING = 0, S_DONE = 1, S_TREE_CAP = 2, S_GROW_CAP = 3 };

struct Header {                 // first bytes of the state blob
    long long num_grow;         // numGrowCalls
    long long num_checks;       // checkConstraint calls
    long long num_trees;        // spanningTrees.size()
    long long max_grow, max_trees;
    long long compactions;
    int top, f_live, ff_top, rem_top;
    int depth, tcount, phase, status;
    unsigned inT[4];
    int cu, cv, ck, cprev;      // the pending checkConstraint call (R_CHECK)
    int ok, pad[3];
};

struct Layout {                 // byte offsets into the blob
    int n, SP, E, fs_cap;
    unsigned o_psum, o_adj, o_adjlen, o_indeg, o_inmask, o_pos, o_fs, o_ff, o_rem;
    unsigned o_ffbase, o_rembase, o_remcnt, o_fu, o_fv, o_tn, o_par, o_last, o_prev, total;
};

inline unsigned rp_align(unsigned x, unsigned a) { return (x + a - 1) / a * a; }

inline Layout make_layout(int n, int SP, int E) {
    Layout L;
    memset(&L, 0, sizeof(L));
    L.n = n; L.SP = SP; L.E = E;
    L.fs_cap = 2 * E + 2 * n + 64;
    if (L.fs_cap < 256) L.fs_cap = 256;
    unsigned o = rp_align((unsigned)sizeof(Header), 16);
    L.o_psum = o;    o = rp_align(o + (unsigned)n * SP * 8u, 16);
    L.o_adj = o;     o = rp_align(o + (unsigned)n * n + 32u, 16);      // (+32: a warp may load 32 entries from the start of any row)
    L.o_adjlen = o;  o = rp_align(o + (unsigned)n, 16);
    L.o_indeg = o;   o = rp_align(o + (unsigned)n, 16);
    L.o_inmask = o;  o = rp_align(o + (unsigned)n * 16u, 16);
    L.o_pos = o;     o = rp_align(o + ((unsigned)n * n + 32u) * 2u, 16);
    L.o_fs = o;      o = rp_align(o + (unsigned)L.fs_cap * 2u, 16);
    L.o_ff = o;      o = rp_align(o + (unsigned)(E + 1) * 2u, 16);
    L.o_rem = o;     o = rp_align(o + (unsigned)(E + 1) * 2u, 16);
    L.o_ffbase = o;  o = rp_align(o + (unsigned)n * 2u, 16);
    L.o_rembase = o; o = rp_align(o + (unsigned)n * 2u, 16);
    L.o_remcnt = o;  o = rp_align(o + (unsigned)n * 2u, 16);
    L.o_fu = o;      o = rp_align(o + (unsigned)n, 16);
    L.o_fv = o;      o = rp_align(o + (unsigned)n, 16);
    L.o_tn = o;      o = rp_align(o + (unsigned)n, 16);
    L.o_par = o;     o = rp_align(o + (unsigned)n, 16);
    L.o_last = o;    o = rp_align(o + (unsigned)n, 16);
    L.o_prev = o;    o = rp_align(o + (unsigned)n, 16);
    L.total = o;
    return L;
}

struct State {                  // typed views of one blob (shared or global memory on the device)
    Header *h;
    double *psum;
    uint8_t *adj, *adjlen, *indeg;
    unsigned *inmask;
    uint16_t *pos, *fs, *ff, *rem, *ffbase, *rembase, *remcnt;
    uint8_t *fu, *fv, *tn, *par, *last, *prev;
    int n, SP, fs_cap;
};

RP_HD State bind(unsigned char *blob, const Layout &L) {
    State s;
    s.h = reinterpret_cast<Header *>(blob);
    s.psum = reinterpret_cast<double *>(blob + L.o_psum);
    s.adj = blob + L.o_adj; s.adjlen = blob + L.o_adjlen; s.indeg = blob + L.o_indeg;
    s.inmask = reinterpret_cast<unsigned *>(blob + L.o_inmask);
    s.pos = reinterpret_cast<uint16_t *>(blob + L.o_pos);
    s.fs = reinterpret_cast<uint16_t *>(blob + L.o_fs);
    s.ff = reinterpret_cast<uint16_t *>(blob + L.o_ff);
    s.rem = reinterpret_cast<uint16_t *>(blob + L.o_rem);
    s.ffbase = reinterpret_cast<uint16_t *>(blob + L.o_ffbase);
    s.rembase = reinterpret_cast<uint16_t *>(blob + L.o_rembase);
    s.remcnt = reinterpret_cast<uint16_t *>(blob + L.o_remcnt);
    s.fu = blob + L.o_fu; s.fv = blob + L.o_fv; s.tn = blob + L.o_tn; s.par = blob + L.o_par;
    s.last = blob + L.o_last; s.prev = blob + L.o_prev;
    s.n = L.n; s.SP = L.SP; s.fs_cap = L.fs_cap;
    return s;
}

// pos is indexed by (target, source): the scan for the edges INTO a node reads consecutive entries
RP_HD int pos_index(int n, int u, int v) { return v * n + u; }

// Initial state of getLineageTrees() (PHYNetwork.java:547-565): t = {root}, f = the root's edges in
// adjacency order.  adj_off/adj_idx: adjacency lists in ROW space, list order kept; the caller has dropped
// duplicates (addEdge refuses them), self loops and edges into the root (the root is in t from the start,
// so such an edge is never pushed and never looked at).  Host only.
inline void init_blob(unsigned char *blob, const Layout &L, const int *adj_off, const int *adj_idx,
                      long long max_grow, long long max_trees) {
    memset(blob, 0, L.total);
    State s = bind(blob, L);
    const int n = L.n;
    memset(s.pos, 0xff, (size_t)n * n * 2);
    memset(s.last, 0xff, n);
    memset(s.prev, 0xff, n);
    for (int u = 0; u < n; u++) {
        int len = 0;
        for (int k = adj_off[u]; k < adj_off[u + 1]; k++) {
            const int v = adj_idx[k];
            s.adj[u * n + len++] = (uint8_t)v;
            s.indeg[v]++;
            s.inmask[v * 4 + (u >> 5)] |= 1u << (u & 31);
        }
        s.adjlen[u] = (uint8_t)len;
    }
    Header *h = s.h;
    h->max_grow = max_grow; h->max_trees = max_trees;
    h->tcount = 1; s.tn[0] = 0; h->inT[0] = 1u;
    for (int i = 0; i < s.adjlen[0]; i++) {
        const int w = s.adj[i];
        s.fs[h->top] = (uint16_t)(w);                    // code of (0, w) = 0 << 7 | w
        s.pos[pos_index(n, 0, w)] = (uint16_t)h->top;
        h->top++; h->f_live++;
    }
    h->phase = P_ENTER; h->depth = 0;
    // `if(nbrs == null || nbrs.size() == 0) return spanningTrees;` -- no grow() call at all
    h->status = s.adjlen[0] == 0 ? S_DONE : S_RUNNING;
}

RP_HD bool in_tree(const Header *h, int v) { return (h->inT[v >> 5] >> (v & 31)) & 1u; }

RP_HD void compact(State &s) {                           // squeeze the tombstones out of fs
    Header *h = s.h;
    int k = 0;
    for (int i = 0; i < h->top; i++) {
        const uint16_t c = s.fs[i];
        if (c != kNone) {
            s.fs[k] = c;
            s.pos[pos_index(s.n, c >> 7, c & 127)] = (uint16_t)k;
            k++;
        }
    }
    h->top = k;
    h->compactions++;
}

RP_HD void push_edge(State &s, uint16_t code) {          // f.add(e)
    Header *h = s.h;
    if (h->top == s.fs_cap) compact(s);
    s.fs[h->top] = code;
    s.pos[pos_index(s.n, code >> 7, code & 127)] = (uint16_t)h->top;
    h->top++;
    h->f_live++;
}

RP_HD void kill_edge(State &s, int u, int v) {           // remove edge (u,v) from f if it is there
    const uint16_t p = s.pos[pos_index(s.n, u, v)];
    if (p != kNone) {
        s.fs[p] = kNone;
        s.pos[pos_index(s.n, u, v)] = kNone;
        s.h->f_live--;
    }
}

// One transition of the state machine, executed by ONE thread (lane 0 of the replay warp on the device).
// The two pieces of per-sample / per-word work are done by the caller between two transitions:
//   R_CHECK   the caller evaluates t.checkConstraint(h->cu) with node h->cv just appended at tree position
//             h->ck (previous child of that parent at position h->cprev, kNoChild = none), fills psum[ck]
//             and stores the outcome in h->ok
//   R_RECORD  the caller stores the complete tree (tn, par) as discovery number h->num_trees
//   R_STOP    the machine has stopped (status != S_RUNNING) or must pause (`room` is false: the record
//             buffer is full; nothing has been counted, calling again with room resumes)
enum Result : int { R_CONTINUE = 0, R_CHECK = 1, R_RECORD = 2, R_STOP = 3 };

RP_HD Result step(State &s, bool room) {
    Header *h = s.h;
    const int n = s.n;
    if (h->status != S_RUNNING) return R_STOP;
    const int d = h->depth;
    switch (h->phase) {
    case P_ENTER: {                                      // grow(t), PHYNetwork.java:443-449
        if (h->tcount == n) {
            if (!room) return R_STOP;                    // pause BEFORE counting the call
            h->num_grow++;
            h->phase = P_RECORDED;
            return R_RECORD;
        }
        h->num_grow++;
        s.ffbase[d] = (uint16_t)h->ff_top;
        h->phase = P_LOOP;
        return R_CONTINUE;
    }
    case P_RECORDED: {
        h->num_trees++;
        if (d == 0) { h->status = S_DONE; return R_STOP; }   // only when the network is the root alone
        h->depth = d - 1;
        h->phase = P_AFTER_RETURN;
        return R_CONTINUE;
    }
    case P_LOOP: {                                       // while(!b && f.size() > 0), :454-460
        if (h->f_live == 0) { h->phase = P_FINISH; return R_CONTINUE; }
        int top = h->top;
        while (s.fs[top - 1] == kNone) top--;
        const uint16_t code = s.fs[--top];
        h->top = top;
        h->f_live--;
        const int u = code >> 7, v = code & 127;
        s.pos[pos_index(n, u, v)] = kNone;
        s.fu[d] = (uint8_t)u; s.fv[d] = (uint8_t)v;
        const int k = h->tcount;                         // t.addNode(v); t.addEdge(u, v)
        const uint8_t prev = s.last[u];
        s.tn[k] = (uint8_t)v; s.par[v] = (uint8_t)u; s.prev[k] = prev;
        s.last[u] = (uint8_t)k;
        h->tcount = k + 1;
        h->inT[v >> 5] |= 1u << (v & 31);
        h->num_checks++;
        h->cu = u; h->cv = v; h->ck = k; h->cprev = prev;
        h->phase = P_CHECKED;
        return R_CHECK;
    }
    case P_CHECKED: {
        if (!h->ok) { h->phase = P_AFTER; return R_CONTINUE; }
        const int v = s.fv[d];
        // edgesAdded: (v, w) for w in edges.get(v), w not in t, :461-470
        const int len = s.adjlen[v];
        for (int i = 0; i < len; i++) {
            const int w = s.adj[v * n + i];
            if (!in_tree(h, w)) push_edge(s, (uint16_t)((v << 7) | w));
        }
        // edgesRemoved: every (w, v) in f, in f order, :473-480
        const int base = h->rem_top;
        int cnt = 0;
        for (int g = 0; g < 4; g++) {
            unsigned m = s.inmask[v * 4 + g] & h->inT[g];
            while (m) {
                const int b = RP_CTZ(m);
                m &= m - 1;
                const int w = g * 32 + b;
                const uint16_t p = s.pos[pos_index(n, w, v)];
                if (p == kNone) continue;
                int j = cnt;                             // insertion sort by position in f
                while (j > 0 && s.pos[pos_index(n, s.rem[base + j - 1] >> 7, v)] > p) { s.rem[base + j] = s.rem[base + j - 1]; j--; }
                s.rem[base + j] = (uint16_t)((w << 7) | v);
                cnt++;
            }
        }
        for (int j = 0; j < cnt; j++) {
            const int w = s.rem[base + j] >> 7;
            s.fs[s.pos[pos_index(n, w, v)]] = kNone;
            s.pos[pos_index(n, w, v)] = kNone;
        }
        h->f_live -= cnt;
        h->rem_top = base + cnt;
        s.rembase[d] = (uint16_t)base; s.remcnt[d] = (uint16_t)cnt;
        if (h->num_grow >= h->max_grow) { h->status = S_GROW_CAP; return R_STOP; }   // :490
        h->depth = d + 1;                                // grow(t), :495
        h->phase = P_ENTER;
        return R_CONTINUE;
    }
    case P_AFTER_RETURN: {                               // :497-505
        if (h->num_trees == h->max_trees) { h->status = S_TREE_CAP; return R_STOP; }
        const int v = s.fv[d];
        const int len = s.adjlen[v];                     // f.removeAll(edgesAdded)
        for (int i = 0; i < len; i++) kill_edge(s, v, s.adj[v * n + i]);
        const int base = s.rembase[d], cnt = s.remcnt[d];
        for (int j = 0; j < cnt; j++) push_edge(s, s.rem[base + j]);   // f.addAll(edgesRemoved)
        h->rem_top = base;
        h->phase = P_AFTER;
        return R_CONTINUE;
    }
    case P_AFTER: {                                      // :508-530
        const int u = s.fu[d], v = s.fv[d];
        const int k = h->tcount - 1;                     // t.removeEdge(u, v): v is the newest node
        s.last[u] = s.prev[k];
        h->tcount = k;
        h->inT[v >> 5] &= ~(1u << (v & 31));
        const int len = s.adjlen[u];                     // this.removeEdge(u, v)
        int i = 0;
        while (i < len && s.adj[u * n + i] != v) i++;
        for (; i + 1 < len; i++) s.adj[u * n + i] = s.adj[u * n + i + 1];
        s.adjlen[u] = (uint8_t)(len - 1);
        s.indeg[v]--;
        s.ff[h->ff_top++] = (uint16_t)((u << 7) | v);    // ff.add(e)
        h->phase = s.indeg[v] == 0 ? P_FINISH : P_LOOP;  // bridge test (see the header comment)
        return R_CONTINUE;
    }
    case P_FINISH: {                                     // :534-540
        const int base = s.ffbase[d];
        for (int i = h->ff_top - 1; i >= base; i--) {
            const uint16_t code = s.ff[i];
            const int u = code >> 7, v = code & 127;
            push_edge(s, code);
            s.adj[u * n + s.adjlen[u]] = (uint8_t)v;
            s.adjlen[u]++;
            s.indeg[v]++;
        }
        h->ff_top = base;
        if (d == 0) { h->status = S_DONE; return R_STOP; }
        h->depth = d - 1;
        h->phase = P_AFTER_RETURN;
        return R_CONTINUE;
    }
    }
    return R_STOP;
}

}  // namespace replay
