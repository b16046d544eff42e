/*
 * lichee_oracle.c -- CPU restatement of LICHeE's lineage-tree search.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in lichee_b200/ may include, link or call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use
 * it, and only as the checker / CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" against a running reference.  The reference is Java; this
 * image has no JVM, release/lichee.jar and lib/weka.jar are missing blobs, and the reference
 * repository holds no unit tests or golden vectors for this path.  The restatement is pinned only
 * (a) statement by statement against the Java sources cited below, (b) against an independent
 * pure-Python restatement (oracle/gm_model.py) on thousands of random instances, and (c) against
 * the two top trees the reference publishes as images (img_demo/RK26.txt.dot.png,
 * img_demo/RMH008.txt.dot.png; see tests/test_published_trees.py).
 *
 * What is restated (read-only reference at /root/reference/LICHeE/src/lineage):
 *   PHYNetwork.getLineageTrees        PHYNetwork.java:547-565
 *   PHYNetwork.grow                   PHYNetwork.java:443-541   (Gabow & Myers '78 with the
 *                                     reference's ArrayList side effects: the frontier stack f and
 *                                     the adjacency lists are re-appended, not restored in place)
 *   PHYNetwork.addEdge / removeEdge   PHYNetwork.java:277-299
 *   PHYNetwork.evaluateLineageTrees   PHYNetwork.java:726-733   (Collections.sort: stable)
 *   PHYTree.addNode/addEdge/removeEdge PHYTree.java:56-96
 *   PHYTree.isDescendent              PHYTree.java:138-154
 *   PHYTree.checkConstraint           PHYTree.java:156-174
 *   PHYTree.computeErrorScore         PHYTree.java:214-233      (parents in descending node id,
 *                                     PHYNode.compareTo PHYNode.java:262-269; Math.pow(x,2)==x*x)
 *   Parameters.MAX_NUM_TREES / MAX_NUM_GROW_CALLS  Parameters.java (100000 / 100000000)
 *
 * mode 0 = the reference, literally.
 * mode 1 = "path-local" variant used to SPECIFY the device's tie-break order: nested grow() calls
 *          leave f and the adjacency lists exactly as they found them.  The set of valid trees is
 *          the same as mode 0; only the enumeration order (hence the child insertion order and the
 *          order among equal scores) can differ.  See DESIGN.md "enumeration order".
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int from, to; } edge_t;

typedef struct {
    int n, S;
    const double *vaf;          /* n x S, row 0 = root (VAF_MAX) */
    double margin;              /* Parameters.VAF_ERROR_MARGIN   */
    long max_trees, max_grow;
    int mode;
    /* PHYNetwork.edges: adjacency lists, order matters */
    int **adj; int *adj_len;
    /* frontier stack f */
    edge_t *f; int f_len, f_cap;
    /* PHYTree t (single live object; L aliases it once a tree has been found) */
    int *t_nodes; int t_nn;     /* treeNodes                         */
    int **t_ch; int *t_nch;     /* treeEdges.get(node)               */
    char *t_has;                /* treeEdges.containsKey(node)       */
    int have_L;
    /* results: clones of complete trees */
    long num_trees, cap_trees;
    int *res_nodes;             /* num_trees x n : treeNodes order   */
    int *res_parent;            /* num_trees x n                     */
    char *res_has;              /* num_trees x n : keys of treeEdges */
    long num_grow, num_checks;
    int stop;                   /* set by the two early returns      */
    int *queue;                 /* scratch for isDescendent          */
} search_t;

static int t_contains(search_t *s, int v) {
    for (int i = 0; i < s->t_nn; i++) if (s->t_nodes[i] == v) return 1;
    return 0;
}
static void t_add_node(search_t *s, int v) {           /* PHYTree.java:56-60 */
    if (!t_contains(s, v)) s->t_nodes[s->t_nn++] = v;
}
static void t_add_edge(search_t *s, int a, int b) {    /* PHYTree.java:62-70 */
    s->t_has[a] = 1;
    for (int i = 0; i < s->t_nch[a]; i++) if (s->t_ch[a][i] == b) return;
    s->t_ch[a][s->t_nch[a]++] = b;
}
static void t_remove_edge(search_t *s, int a, int b) { /* PHYTree.java:72-96 */
    if (s->t_has[a]) {
        for (int i = 0; i < s->t_nch[a]; i++) if (s->t_ch[a][i] == b) {
            memmove(&s->t_ch[a][i], &s->t_ch[a][i + 1], sizeof(int) * (s->t_nch[a] - i - 1));
            s->t_nch[a]--;
            break;
        }
    }
    int connected = 0;
    for (int p = 0; p < s->n && !connected; p++)
        if (s->t_has[p]) for (int i = 0; i < s->t_nch[p]; i++) if (s->t_ch[p][i] == b) { connected = 1; break; }
    if (!connected) {
        for (int i = 0; i < s->t_nn; i++) if (s->t_nodes[i] == b) {
            memmove(&s->t_nodes[i], &s->t_nodes[i + 1], sizeof(int) * (s->t_nn - i - 1));
            s->t_nn--;
            break;
        }
    }
}
static int t_is_descendent(search_t *s, int v, int w) { /* PHYTree.java:138-154 */
    if (!s->t_has[v]) return 0;
    int head = 0, tail = 0;
    for (int i = 0; i < s->t_nch[v]; i++) s->queue[tail++] = s->t_ch[v][i];
    while (head < tail && tail < 4 * s->n) {
        int x = s->queue[head++];
        if (x == w) return 1;
        if (s->t_has[x]) for (int i = 0; i < s->t_nch[x] && tail < 4 * s->n; i++) s->queue[tail++] = s->t_ch[x][i];
    }
    return 0;
}
static int t_check(search_t *s, int p) {               /* PHYTree.java:156-174 */
    s->num_checks++;
    if (!s->t_has[p]) return 1;
    for (int i = 0; i < s->S; i++) {
        double sum = 0;
        for (int k = 0; k < s->t_nch[p]; k++) sum += s->vaf[(size_t)s->t_ch[p][k] * s->S + i];
        if (sum > s->vaf[(size_t)p * s->S + i] + s->margin) return 0;
    }
    return 1;
}
static void g_remove_edge(search_t *s, int a, int b) { /* PHYNetwork.java:289-299 */
    for (int i = 0; i < s->adj_len[a]; i++) if (s->adj[a][i] == b) {
        memmove(&s->adj[a][i], &s->adj[a][i + 1], sizeof(int) * (s->adj_len[a] - i - 1));
        s->adj_len[a]--;
        return;
    }
}
static void g_add_edge(search_t *s, int a, int b) {    /* PHYNetwork.java:277-286 */
    for (int i = 0; i < s->adj_len[a]; i++) if (s->adj[a][i] == b) return;
    s->adj[a][s->adj_len[a]++] = b;
}
static void f_push(search_t *s, edge_t e) {
    if (s->f_len == s->f_cap) { s->f_cap *= 2; s->f = (edge_t *)realloc(s->f, sizeof(edge_t) * s->f_cap); }
    s->f[s->f_len++] = e;
}
static int in_list(const edge_t *l, int n, edge_t e) {
    for (int i = 0; i < n; i++) if (l[i].from == e.from && l[i].to == e.to) return 1;
    return 0;
}
static void f_remove_all(search_t *s, const edge_t *l, int n) { /* ArrayList.removeAll */
    if (n == 0) return;
    int k = 0;
    for (int i = 0; i < s->f_len; i++) if (!in_list(l, n, s->f[i])) s->f[k++] = s->f[i];
    s->f_len = k;
}
static void record_tree(search_t *s) {
    if (s->num_trees == s->cap_trees) {
        s->cap_trees = s->cap_trees ? s->cap_trees * 2 : 1024;
        s->res_nodes = (int *)realloc(s->res_nodes, sizeof(int) * s->cap_trees * s->n);
        s->res_parent = (int *)realloc(s->res_parent, sizeof(int) * s->cap_trees * s->n);
        s->res_has = (char *)realloc(s->res_has, (size_t)s->cap_trees * s->n);
    }
    int *nodes = s->res_nodes + s->num_trees * s->n, *par = s->res_parent + s->num_trees * s->n;
    memcpy(nodes, s->t_nodes, sizeof(int) * s->n);
    for (int i = 0; i < s->n; i++) par[i] = -1;
    for (int p = 0; p < s->n; p++) for (int k = 0; k < s->t_nch[p]; k++) par[s->t_ch[p][k]] = p;
    memcpy(s->res_has + s->num_trees * s->n, s->t_has, s->n);
    s->num_trees++;
}

static void grow(search_t *s) {                        /* PHYNetwork.java:443-541 */
    s->num_grow++;
    if (s->t_nn == s->n) { s->have_L = 1; record_tree(s); return; }
    int ff_len = 0, ff_cap = 16;
    edge_t *ff = (edge_t *)malloc(sizeof(edge_t) * ff_cap);
    int *ff_pos = (int *)malloc(sizeof(int) * ff_cap);     /* mode 1 only */
    edge_t *f_saved = NULL; int f_saved_len = 0;
    int b = 0;
    while (!b && s->f_len > 0) {
        edge_t e = s->f[--s->f_len];
        int v = e.to;
        t_add_node(s, v);
        t_add_edge(s, e.from, v);
        if (t_check(s, e.from)) {
            int na = 0, nr = 0;
            edge_t *added = (edge_t *)malloc(sizeof(edge_t) * (s->adj_len[v] + 1));
            for (int i = 0; i < s->adj_len[v]; i++) {
                int w = s->adj[v][i];
                if (!t_contains(s, w)) { edge_t vw = { v, w }; f_push(s, vw); added[na++] = vw; }
            }
            edge_t *removed = (edge_t *)malloc(sizeof(edge_t) * (s->f_len + 1));
            for (int i = 0; i < s->f_len; i++)
                if (t_contains(s, s->f[i].from) && s->f[i].to == v) removed[nr++] = s->f[i];
            f_remove_all(s, removed, nr);
            if (s->num_grow >= s->max_grow) { s->stop = 1; free(added); free(removed); free(ff); free(ff_pos); return; }
            if (s->mode == 1) {
                f_saved_len = s->f_len;
                f_saved = (edge_t *)malloc(sizeof(edge_t) * (s->f_len + 1));
                memcpy(f_saved, s->f, sizeof(edge_t) * s->f_len);
            }
            grow(s);
            if (s->num_trees == s->max_trees) { s->stop = 1; free(added); free(removed); free(ff); free(ff_pos); free(f_saved); return; }
            if (s->mode == 1) {
                if (!s->stop) { memcpy(s->f, f_saved, sizeof(edge_t) * f_saved_len); s->f_len = f_saved_len; }
                free(f_saved); f_saved = NULL;
            }
            f_remove_all(s, added, na);
            for (int i = 0; i < nr; i++) f_push(s, removed[i]);
            free(added); free(removed);
        }
        t_remove_edge(s, e.from, e.to);
        if (ff_len == ff_cap) { ff_cap *= 2; ff = (edge_t *)realloc(ff, sizeof(edge_t) * ff_cap); ff_pos = (int *)realloc(ff_pos, sizeof(int) * ff_cap); }
        if (s->mode == 1) { int p = 0; while (p < s->adj_len[e.from] && s->adj[e.from][p] != e.to) p++; ff_pos[ff_len] = p; }
        g_remove_edge(s, e.from, e.to);
        ff[ff_len++] = e;
        /* bridge test, PHYNetwork.java:515-530 */
        b = 1;
        for (int w = 0; w < s->n && b; w++)
            for (int i = 0; i < s->adj_len[w]; i++)
                if (s->adj[w][i] == v) {
                    if (!s->have_L || !t_is_descendent(s, v, w)) { b = 0; break; }
                }
    }
    for (int i = ff_len - 1; i >= 0; i--) {
        edge_t e = ff[i];
        f_push(s, e);
        if (s->mode == 1) {
            int a = e.from, p = ff_pos[i];
            memmove(&s->adj[a][p + 1], &s->adj[a][p], sizeof(int) * (s->adj_len[a] - p));
            s->adj[a][p] = e.to; s->adj_len[a]++;
        } else g_add_edge(s, e.from, e.to);
    }
    free(ff); free(ff_pos);
}

/* PHYTree.computeErrorScore on a recorded clone: child lists in insertion order are the
 * relative order of the children inside the clone's treeNodes. */
static double score_tree(const search_t *s, const int *nodes, const int *par) {
    double err = 0;
    for (int p = s->n - 1; p >= 0; p--) {              /* descending node id */
        for (int i = 0; i < s->S; i++) {
            double sum = 0; int any = 0;
            for (int k = 0; k < s->n; k++) { int c = nodes[k]; if (par[c] == p) { sum += s->vaf[(size_t)c * s->S + i]; any = 1; } }
            if (!any) break;
            double a = s->vaf[(size_t)p * s->S + i];
            if (sum > a) err += (sum - a) * (sum - a);
        }
    }
    return sqrt(err);
}

typedef struct { double score; long idx; } ranked_t;
static void merge_sort(ranked_t *a, ranked_t *tmp, long n) { /* stable, like Collections.sort */
    if (n < 2) return;
    long h = n / 2;
    merge_sort(a, tmp, h); merge_sort(a + h, tmp, n - h);
    long i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (a[j].score < a[i].score) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, sizeof(ranked_t) * n);
}

/*
 * adj_off/adj_idx: CSR of PHYNetwork.edges in the reference's adjacency-list order.
 * stats[0]=trees found, [1]=numGrowCalls, [2]=checkConstraint calls, [3]=1 if a cap stopped the search.
 * Outputs hold the first min(top_s, trees) trees of the ranked list: score, parent[n],
 * order[n] = treeNodes (root first, then nodes in the order they joined the tree),
 * dfs_index = position in spanningTrees before sorting.
 * all_scores / all_parents / all_orders (may be NULL): score, parents and treeNodes of every tree in
 * enumeration (discovery) order, up to all_cap trees.
 */
int lichee_oracle_search(int n, int S, const int *adj_off, const int *adj_idx, const double *vaf,
                         double margin, long max_trees, long max_grow, int mode, int top_s,
                         double *out_score, int *out_parent, int *out_order, long *out_dfs_index,
                         long *stats, double *all_scores, int *all_parents, long all_cap, int *all_orders) {
    search_t s; memset(&s, 0, sizeof(s));
    s.n = n; s.S = S; s.vaf = vaf; s.margin = margin; s.max_trees = max_trees; s.max_grow = max_grow; s.mode = mode;
    s.adj = (int **)malloc(sizeof(int *) * n); s.adj_len = (int *)calloc(n, sizeof(int));
    s.t_ch = (int **)malloc(sizeof(int *) * n); s.t_nch = (int *)calloc(n, sizeof(int));
    for (int i = 0; i < n; i++) {
        s.adj[i] = (int *)malloc(sizeof(int) * (n + 1));
        s.t_ch[i] = (int *)malloc(sizeof(int) * (n + 1));
        for (int k = adj_off[i]; k < adj_off[i + 1]; k++) g_add_edge(&s, i, adj_idx[k]);
    }
    s.t_nodes = (int *)malloc(sizeof(int) * (n + 1)); s.t_has = (char *)calloc(n, 1);
    s.queue = (int *)malloc(sizeof(int) * (4 * n + 4));
    s.f_cap = 64; s.f = (edge_t *)malloc(sizeof(edge_t) * s.f_cap);
    t_add_node(&s, 0);
    if (s.adj_len[0] > 0) {
        for (int i = 0; i < s.adj_len[0]; i++) { edge_t e = { 0, s.adj[0][i] }; f_push(&s, e); }
        grow(&s);
    }
    ranked_t *r = (ranked_t *)malloc(sizeof(ranked_t) * (s.num_trees + 1)), *tmp = (ranked_t *)malloc(sizeof(ranked_t) * (s.num_trees + 1));
    for (long i = 0; i < s.num_trees; i++) {
        r[i].score = score_tree(&s, s.res_nodes + i * n, s.res_parent + i * n); r[i].idx = i;
        if (all_scores && i < all_cap) all_scores[i] = r[i].score;
        if (all_parents && i < all_cap) memcpy(all_parents + i * n, s.res_parent + i * n, sizeof(int) * n);
        if (all_orders && i < all_cap) memcpy(all_orders + i * n, s.res_nodes + i * n, sizeof(int) * n);
    }
    merge_sort(r, tmp, s.num_trees);
    long m = s.num_trees < top_s ? s.num_trees : top_s;
    for (long k = 0; k < m; k++) {
        out_score[k] = r[k].score; out_dfs_index[k] = r[k].idx;
        memcpy(out_parent + k * n, s.res_parent + r[k].idx * n, sizeof(int) * n);
        memcpy(out_order + k * n, s.res_nodes + r[k].idx * n, sizeof(int) * n);
    }
    stats[0] = s.num_trees; stats[1] = s.num_grow; stats[2] = s.num_checks; stats[3] = s.stop;
    for (int i = 0; i < n; i++) { free(s.adj[i]); free(s.t_ch[i]); }
    free(s.adj); free(s.adj_len); free(s.t_ch); free(s.t_nch); free(s.t_nodes); free(s.t_has); free(s.queue); free(s.f);
    free(s.res_nodes); free(s.res_parent); free(s.res_has); free(r); free(tmp);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Canonical-order specification (TEST INFRASTRUCTURE): a sequential statement of the order in
 * which lichee_b200 enumerates and sums, used to check the CUDA path bit for bit.  The SET of
 * valid trees and every tree's score are the reference's (up to the floating-point effect of the
 * child summation order); only the enumeration order differs -- DESIGN.md "enumeration order"
 * explains why the reference's own order cannot be reproduced by any parallel search.
 *
 *   sigma   : non-root nodes by descending sum of VAF over samples, ties by ascending id
 *   tree    : one parent per node, parents of a node taken in ascending id
 *   order   : lexicographic in (parent of sigma[0], parent of sigma[1], ...)
 *   sums    : children of a parent are added in sigma order (checkConstraint, PHYTree.java:156-174,
 *             and the per-sample sums of computeErrorScore, PHYTree.java:214-233)
 *   score   : the positive terms (sum - VAF(parent))^2 are added with the device's fixed-shape
 *             reduction (per parent a 32-lane butterfly over samples, then a butterfly over 128
 *             parent slots, slot 0 = root, slot t+1 = sigma[t]) instead of the reference's one-after-another order: a few ulp apart,
 *             far inside the 1e-9 relative tolerance BASELINE.json states
 *   caps    : the first max_trees valid trees in this order
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int n, S, L; const double *vaf; double margin; long max_trees;
    int *sigma; int **cand; int *ncand;
    int *par;                   /* parent of sigma[t] */
    int **ch; int *nch;         /* children per parent, in sigma order */
    long n_trees, n_checks, n_grow;
    int top_s; long kept; double *k_score; long *k_seq; int *k_par;  /* ranked list, insertion sort */
} canon_t;

/* fixed-shape reductions of the device (lichee_search.cu, kernel 4): xor-butterfly over 32 lanes */
static double butterfly32(double *v) {
    for (int d = 16; d > 0; d >>= 1) {
        double t[32];
        for (int l = 0; l < 32; l++) t[l] = v[l] + v[l ^ d];
        memcpy(v, t, sizeof(t));
    }
    return v[0];
}
static double canon_score(canon_t *c) {
    double E[128];
    int slot[128];                       /* reduction slot of a node: 0 = root, t+1 = sigma[t] */
    memset(E, 0, sizeof(E));
    slot[0] = 0;
    for (int t = 0; t < c->L; t++) slot[c->sigma[t]] = t + 1;
    for (int p = 0; p < c->n; p++) {
        if (c->nch[p] == 0) continue;
        double t[64], lane[32];
        memset(t, 0, sizeof(t));
        for (int i = 0; i < c->S; i++) {
            double sum = 0;
            for (int k = 0; k < c->nch[p]; k++) sum += c->vaf[(size_t)c->ch[p][k] * c->S + i];
            double a = c->vaf[(size_t)p * c->S + i];
            if (sum > a) t[i] = (sum - a) * (sum - a);
        }
        for (int l = 0; l < 32; l++) lane[l] = c->S > 32 ? t[l] + t[l + 32] : t[l];
        E[slot[p]] = butterfly32(lane);
    }
    double lane[32];
    for (int l = 0; l < 32; l++) lane[l] = (E[l] + E[l + 32]) + (E[l + 64] + E[l + 96]);
    return sqrt(butterfly32(lane));
}
static void canon_keep(canon_t *c, double score, long seq) {
    if (c->kept == c->top_s && !(score < c->k_score[c->kept - 1])) return;
    long pos = c->kept < c->top_s ? c->kept : c->kept - 1;
    while (pos > 0 && score < c->k_score[pos - 1]) {
        c->k_score[pos] = c->k_score[pos - 1]; c->k_seq[pos] = c->k_seq[pos - 1];
        memcpy(c->k_par + pos * c->n, c->k_par + (pos - 1) * c->n, sizeof(int) * c->n);
        pos--;
    }
    c->k_score[pos] = score; c->k_seq[pos] = seq;
    c->k_par[pos * c->n] = -1;
    for (int t = 0; t < c->L; t++) c->k_par[pos * c->n + c->sigma[t]] = c->par[t];
    if (c->kept < c->top_s) c->kept++;
}
static void canon_rec(canon_t *c, int t) {
    c->n_grow++;
    if (t == c->L) { if (c->n_trees < c->max_trees) { canon_keep(c, canon_score(c), c->n_trees); } c->n_trees++; return; }
    int v = c->sigma[t];
    for (int i = 0; i < c->ncand[t] && c->n_trees < c->max_trees; i++) {
        int p = c->cand[t][i], ok = 1;
        c->n_checks++;
        c->ch[p][c->nch[p]++] = v;
        for (int s = 0; s < c->S && ok; s++) {
            double sum = 0;
            for (int k = 0; k < c->nch[p]; k++) sum += c->vaf[(size_t)c->ch[p][k] * c->S + s];
            if (sum > c->vaf[(size_t)p * c->S + s] + c->margin) ok = 0;
        }
        if (ok) { c->par[t] = p; canon_rec(c, t + 1); }
        c->nch[p]--;
    }
}
/* stats: [0] valid trees (capped), [1] grow calls, [2] checks.  out_sigma[n-1]. */
int lichee_canonical_search(int n, int S, const int *adj_off, const int *adj_idx, const double *vaf,
                            double margin, long max_trees, int top_s, double *out_score, int *out_parent,
                            long *out_seq, long *stats, int *out_sigma) {
    canon_t c; memset(&c, 0, sizeof(c));
    c.n = n; c.S = S; c.L = n - 1; c.vaf = vaf; c.margin = margin; c.max_trees = max_trees; c.top_s = top_s;
    c.sigma = (int *)malloc(sizeof(int) * (n + 1)); c.par = (int *)malloc(sizeof(int) * (n + 1));
    c.cand = (int **)malloc(sizeof(int *) * (n + 1)); c.ncand = (int *)calloc(n + 1, sizeof(int));
    c.ch = (int **)malloc(sizeof(int *) * n); c.nch = (int *)calloc(n, sizeof(int));
    double *mass = (double *)calloc(n, sizeof(double));
    for (int v = 0; v < n; v++) { c.ch[v] = (int *)malloc(sizeof(int) * (n + 1)); for (int s = 0; s < S; s++) mass[v] += vaf[(size_t)v * S + s]; }
    for (int t = 0; t < c.L; t++) c.sigma[t] = t + 1;
    for (int i = 1; i < c.L; i++) {                     /* stable insertion sort, descending mass */
        int x = c.sigma[i], j = i;
        while (j > 0 && mass[c.sigma[j - 1]] < mass[x]) { c.sigma[j] = c.sigma[j - 1]; j--; }
        c.sigma[j] = x;
    }
    int feasible = n > 1 && adj_off[1] > adj_off[0];
    for (int k = 0; k < adj_off[n]; k++) if (adj_idx[k] == 0) feasible = 0;
    for (int t = 0; t < c.L; t++) {
        int v = c.sigma[t];
        c.cand[t] = (int *)malloc(sizeof(int) * (n + 1));
        for (int u = 0; u < n; u++) {                   /* ascending id, duplicates collapse */
            int has = 0;
            for (int k = adj_off[u]; k < adj_off[u + 1]; k++) if (adj_idx[k] == v) has = 1;
            if (has) c.cand[t][c.ncand[t]++] = u;
        }
        if (c.ncand[t] == 0) feasible = 0;
        if (out_sigma) out_sigma[t] = v;
    }
    c.k_score = (double *)malloc(sizeof(double) * (top_s + 1)); c.k_seq = (long *)malloc(sizeof(long) * (top_s + 1));
    c.k_par = (int *)malloc(sizeof(int) * (size_t)(top_s + 1) * n);
    if (feasible) canon_rec(&c, 0); else c.n_grow = 1;
    for (long k = 0; k < c.kept; k++) { out_score[k] = c.k_score[k]; out_seq[k] = c.k_seq[k]; memcpy(out_parent + k * n, c.k_par + k * n, sizeof(int) * n); }
    stats[0] = c.n_trees < max_trees ? c.n_trees : max_trees; stats[1] = c.n_grow; stats[2] = c.n_checks;
    for (int t = 0; t < c.L; t++) free(c.cand[t]);
    for (int v = 0; v < n; v++) free(c.ch[v]);
    free(c.sigma); free(c.par); free(c.cand); free(c.ncand); free(c.ch); free(c.nch); free(mass);
    free(c.k_score); free(c.k_seq); free(c.k_par);
    return 0;
}

/* Evaluates ONE tree given as a parent vector (TEST INFRASTRUCTURE): returns 1 if every parent
 * satisfies the sum rule (PHYTree.checkConstraint with the children in canonical sigma order) and
 * stores the canonical error score; used to re-verify trees returned by searches that are too large
 * to enumerate on the CPU. */
int lichee_canonical_eval_tree(int n, int S, const double *vaf, double margin, const int *parent, double *score_out) {
    canon_t c; memset(&c, 0, sizeof(c));
    c.n = n; c.S = S; c.L = n - 1; c.vaf = vaf; c.margin = margin;
    c.sigma = (int *)malloc(sizeof(int) * (n + 1));
    c.ch = (int **)malloc(sizeof(int *) * n); c.nch = (int *)calloc(n, sizeof(int));
    double *mass = (double *)calloc(n, sizeof(double));
    for (int v = 0; v < n; v++) { c.ch[v] = (int *)malloc(sizeof(int) * (n + 1)); for (int s = 0; s < S; s++) mass[v] += vaf[(size_t)v * S + s]; }
    for (int t = 0; t < c.L; t++) c.sigma[t] = t + 1;
    for (int i = 1; i < c.L; i++) {
        int x = c.sigma[i], j = i;
        while (j > 0 && mass[c.sigma[j - 1]] < mass[x]) { c.sigma[j] = c.sigma[j - 1]; j--; }
        c.sigma[j] = x;
    }
    int ok = 1;
    for (int t = 0; t < c.L; t++) {
        int v = c.sigma[t], p = parent[v];
        if (p < 0 || p >= n || p == v) { ok = 0; break; }
        c.ch[p][c.nch[p]++] = v;
    }
    for (int p = 0; p < n && ok; p++) {
        if (c.nch[p] == 0) continue;
        for (int s = 0; s < S && ok; s++) {
            double sum = 0;
            for (int k = 0; k < c.nch[p]; k++) sum += vaf[(size_t)c.ch[p][k] * S + s];
            if (sum > vaf[(size_t)p * S + s] + margin) ok = 0;
        }
    }
    if (ok && score_out) *score_out = canon_score(&c);
    for (int v = 0; v < n; v++) free(c.ch[v]);
    free(c.sigma); free(c.ch); free(c.nch); free(mass);
    return ok;
}
