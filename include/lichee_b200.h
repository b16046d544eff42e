/*
 * lichee_b200.h -- C ABI of the B200 lineage-tree search (libLICHeE_b200.so).
 *
 * This is the drop-in boundary for the hot path of LICHeE's `-build` command: the enumeration of
 * the spanning trees of the evolutionary constraint network, the per-sample VAF sum-rule check,
 * and the error scoring + ranking of the valid trees.  Everything either side of it (SSNV
 * loading, clustering, network construction, GUI, DOT export) stays in the Java host.
 *
 * Reference interfaces replaced (LICHeE/src/lineage):
 *   PHYNetwork.getLineageTrees()        PHYNetwork.java:547-565  -> lichee_search_run
 *   PHYNetwork.grow(PHYTree)            PHYNetwork.java:443-541  -> device frontier expansion
 *   PHYTree.checkConstraint(PHYNode)    PHYTree.java:156-174     -> device sum-rule check
 *   PHYTree.computeErrorScore()         PHYTree.java:214-233     -> device scoring
 *   PHYNetwork.evaluateLineageTrees()   PHYNetwork.java:726-733  -> device top-s selection
 *   LineageEngine.buildLineage step 5/6 LineageEngine.java:106-134 (caller)
 *
 * Plain pointers and sizes only; no CUDA, torch or C++ types cross this boundary.
 * All functions return LICHEE_OK (0) or a negative error code; lichee_last_error() gives text.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * LICHEE_ERR_NO_DEVICE.
 */
#ifndef LICHEE_B200_H
#define LICHEE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LICHEE_OK                 0
#define LICHEE_ERR_INVALID       -1   /* bad argument (null pointer, size out of range)          */
#define LICHEE_ERR_NO_DEVICE     -2   /* no CUDA device / CUDA runtime error                     */
#define LICHEE_ERR_CYCLIC        -3   /* constraint network is not a DAG                         */
#define LICHEE_ERR_TOO_LARGE     -4   /* more than LICHEE_MAX_NODES nodes / LICHEE_MAX_SAMPLES   */
#define LICHEE_ERR_CUDA          -5   /* a CUDA call failed; see lichee_last_error()             */
#define LICHEE_ERR_STATE         -6   /* call order violated (e.g. result before run)            */

#define LICHEE_MAX_NODES    128       /* network nodes including the germline root (node 0)      */
#define LICHEE_MAX_SAMPLES   64
#define LICHEE_MAX_SAVE    1024       /* upper bound on -s                                       */
#define LICHEE_DEFAULT_REF_ORDER_BUDGET (1ll << 22)   /* grow() calls; every bundled ccRCC / HGSC network needs < 2^18 */

/* status bits returned in lichee_search_stats.status */
#define LICHEE_STATUS_TREE_CAP   1u   /* stopped at max_num_trees   (Parameters.MAX_NUM_TREES)      */
#define LICHEE_STATUS_GROW_CAP   2u   /* stopped at max_num_grow_calls (Parameters.MAX_NUM_GROW_CALLS) */
#define LICHEE_STATUS_CANONICAL_ORDER 4u /* the ranked list follows the CANONICAL enumeration order, not the
                                         reference's discovery order: ties in score, the trees kept under a
                                         binding MAX_NUM_TREES and the order of a node's children (hence the
                                         last bits of a score) may differ from the Java implementation.  Set when
                                         the reference-order replay is switched off (ref_order_budget < 0), when
                                         the search is sharded (part_count > 1, lichee_search_run_slice) or when
                                         the replay did not finish within ref_order_budget grow() calls.  Clear:
                                         counts, ranked trees, treeNodes order and scores are those of
                                         PHYNetwork.getLineageTrees + evaluateLineageTrees, bit for bit.        */

#define LICHEE_K_CHECK   0        /* sum-rule check of every candidate extension (check_kernel,
                                     check2_kernel, and the single-block scheduler descend_kernel) */
#define LICHEE_K_SCAN    1        /* ordered exclusive scan of the pass counts                */
#define LICHEE_K_EMIT    2        /* write surviving children to the next frontier level      */
#define LICHEE_K_LEAF    3        /* error score of complete trees + threshold filter         */
#define LICHEE_K_SELECT  4        /* sequence numbers, radix pre-selection, rank-merge top-s  */
#define LICHEE_NUM_KERNELS 5

typedef struct lichee_search lichee_search;   /* opaque handle: one constraint network on one GPU */

/* Search parameters.  Field <-> reference parameter:
 *   vaf_error_margin    -e                     Parameters.VAF_ERROR_MARGIN      (default 0.1)
 *   num_save            -s                     LineageEngine.Args.numSave       (default 1)
 *   max_num_trees       Parameters.MAX_NUM_TREES        (default 100000)
 *   max_num_grow_calls  Parameters.MAX_NUM_GROW_CALLS   (default 100000000).  Reference-order mode: the reference's own
 *                       counter and early return (PHYNetwork.java:490), exactly.  Canonical mode: a WORK BUDGET -- accepted
 *                       partial trees of a different traversal, checked at chunk boundaries; it cannot mean what :490 means.
 * -minRobustNodeSupport acts before this boundary (it decides which clusters are robust, hence
 * which nodes PHYNetwork.fixNetwork removes between two searches); the host re-submits the
 * adjusted network, see INTEGRATION.md.
 * part_rank/part_count shard one search over several GPUs: the frontier at the first tree levels
 * is cut into part_count contiguous slices of the canonical enumeration order and this handle
 * searches slice part_rank (0/1 = the whole search).
 */
typedef struct lichee_search_params {
    double  vaf_error_margin;
    int32_t num_save;
    int32_t device;              /* CUDA device ordinal                                        */
    int64_t max_num_trees;
    int64_t max_num_grow_calls;
    int32_t part_rank;
    int32_t part_count;
    int64_t level_capacity;      /* frontier items per tree level kept in HBM; 0 = default      */
    int64_t ref_order_budget;    /* grow() calls the reference-order replay may spend before the search
                                    falls back to the canonical order (LICHEE_STATUS_CANONICAL_ORDER):
                                    0 = default (LICHEE_DEFAULT_REF_ORDER_BUDGET), < 0 = canonical order only */
    int64_t reserved[3];
} lichee_search_params;

typedef struct lichee_search_stats {
    int64_t num_trees;           /* valid spanning trees found ("Found N valid tree(s)")        */
    int64_t num_grow_calls;      /* reference-order mode: the reference's numGrowCalls, exactly; canonical mode:
                                    accepted partial trees + 1 (what numGrowCalls would be for that traversal) */
    int64_t num_checks;          /* reference-order mode: checkConstraint calls, exactly; canonical mode: sum-rule
                                    DECISIONS (one per candidate extension; most are carried over, static or
                                    belong to dismissed items -- see num_evaluated)                 */
    int64_t num_scored;          /* complete valid trees ranked: scored, or excluded from the
                                    ranked list by the lower bound of their partial tree        */
    int64_t num_launches;        /* kernels launched by lichee_search_run                       */
    int64_t frontier_bytes;      /* algorithmic HBM bytes moved through the frontier buffers    */
    double  kernel_ms_total;     /* CUDA-event time of the whole device search                  */
    /* per kernel class (LICHEE_K_*): summed CUDA-event time on the launching stream, launches,
     * and ALGORITHMIC bytes (frontier items + masks each launch must read and write once)       */
    double  kernel_ms[LICHEE_NUM_KERNELS];   /* CUDA events per kernel class; the check class includes launches that run ahead on the
                                               library's second stream, BESIDE the other classes: the classes can add up to more
                                               than kernel_ms_total                                                              */
    int64_t kernel_launches[LICHEE_NUM_KERNELS];
    int64_t kernel_bytes[LICHEE_NUM_KERNELS];
    uint32_t status;
    uint32_t num_saved;          /* trees held in the ranked list, <= num_save                  */
    int64_t num_evaluated;       /* complete valid trees whose error score was actually computed (canonical
                                    search: num_scored minus the trees dismissed by the lower bound of their
                                    base; reference order: every recorded tree)                   */
    int64_t num_bounded_items;   /* last-level frontier items skipped by the lower bound         */
    int64_t num_leaf_items;      /* last-level frontier items in total                           */
} lichee_search_stats;

void lichee_default_params(lichee_search_params *p);

/* Ship one constraint network to the device.
 *   n_nodes, n_samples : PHYNetwork.numNodes / numSamples; node 0 is the germline root.
 *   adj_off[n_nodes+1], adj_idx[adj_off[n_nodes]] : PHYNetwork.edges as CSR, children of node i in
 *       adjacency-list order (PHYNetwork.java:75).
 *   vaf[n_nodes * n_samples] : PHYNode.getAAF(sample) per node, row-major; row 0 = VAF_MAX.
 * The arrays are HOST memory and are copied; they may be freed after the call returns. */
int lichee_search_create(lichee_search **out, int32_t n_nodes, int32_t n_samples,
                         const int32_t *adj_off, const int32_t *adj_idx, const double *vaf,
                         const lichee_search_params *params);

/* getLineageTrees() + evaluateLineageTrees(): enumerate, check, score, rank.  Blocking. */
int lichee_search_run(lichee_search *h);

/* Dynamic work stealing: search slice `slice` of `n_slices` contiguous slices of the canonical order
 * (overrides part_rank/part_count).  With keep_ranked != 0 the ranked list, its threshold and the
 * statistics of the earlier slices on this handle are kept, so a GPU can claim slice after slice
 * from a counter shared by all ranks (lichee_b200/dist.py does it with the torch.distributed store).
 * Ties rank by (slice, seq), i.e. the same canonical order as one whole search.  Slices kept on one
 * handle must be searched in ASCENDING order (LICHEE_ERR_STATE otherwise): a tree that ties with the worst
 * held one is admitted only when it precedes it, which the kernels can assume of nothing that comes later.
 * max_num_trees applies to the slice; ONE cap over the whole search (PHYNetwork.java:497, one spanningTrees
 * list) is the caller's job: lichee_b200/dist.py gathers the per-slice counts, re-runs the slice the cap
 * falls into with lichee_search_set_max_num_trees and drops the slices beyond it. */
int lichee_search_run_slice(lichee_search *h, int32_t slice, int32_t n_slices, int32_t keep_ranked);
/* Empty ranked list and zero statistics (a rank that claimed no slice still takes part in the merge). */
int lichee_search_clear(lichee_search *h);

/* Parameters.MAX_NUM_TREES for the runs that follow on this handle. */
int lichee_search_set_max_num_trees(lichee_search *h, int64_t max_num_trees);

int lichee_search_get_stats(const lichee_search *h, lichee_search_stats *out);

/* Ranked tree k (0 = lowest error score), k < stats.num_saved.
 *   score      : PHYTree.getErrorScore()
 *   parent[n]  : parent node of every node (-1 for the root)        -> PHYTree.treeEdges
 *   order[n]   : nodes in the order they join the tree, root first  -> PHYTree.treeNodes; the
 *                children of a node, taken in this order, are its treeEdges list
 *   seq        : position of the tree in the enumeration order (ties in score rank by it)      */
int lichee_search_get_tree(const lichee_search *h, int32_t k, double *score, int32_t *parent,
                           int32_t *order, int64_t *seq);

/* Ranked trees first .. first+count-1 in one call: scores[count], parents[count*n_nodes],
 * orders[count*n_nodes], seqs[count] (any output may be NULL). */
int lichee_search_get_trees(const lichee_search *h, int32_t first, int32_t count, double *scores,
                            int32_t *parents, int32_t *orders, int64_t *seqs);

/* Multi-GPU merge: raw ranked list of this handle (for an allgather) and its inverse.
 * A record is lichee_record_bytes(h) bytes: {double score; int64 seq; int32 part; int32 pad;
 * uint8 parent_of_position[stride]}.  Records stay in DEVICE memory when `device_ptr` is
 * non-zero (so NCCL can move them), else they are copied to the host buffer. */
int64_t lichee_record_bytes(const lichee_search *h);
int lichee_search_export_ranked(const lichee_search *h, void *dst, int device_ptr);
int lichee_search_merge_ranked(lichee_search *h, const void *records, int32_t n_lists, int device_ptr);

/* Canonical enumeration order of this network: sigma[i] = node assigned at tree level i+1
 * (n_nodes-1 entries).  Used by tests and by hosts that want to reproduce `seq`. */
int lichee_search_get_order(const lichee_search *h, int32_t *sigma);

int lichee_search_destroy(lichee_search *h);

/* One-shot convenience used by the JNI stub: create + run + copy the ranked list out.
 * scores[num_save], parents[num_save*n_nodes], orders[num_save*n_nodes]. */
int lichee_get_lineage_trees(int32_t n_nodes, int32_t n_samples, const int32_t *adj_off,
                             const int32_t *adj_idx, const double *vaf,
                             const lichee_search_params *params, lichee_search_stats *stats,
                             double *scores, int32_t *parents, int32_t *orders);

/* Frees the device buffers, streams and events that destroyed handles returned to the process-wide
 * pool (live handles are not touched).  Optional: the pool is otherwise kept until process exit. */
int lichee_release_cached_memory(void);

const char *lichee_last_error(void);
const char *lichee_version(void);
int lichee_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* LICHEE_B200_H */
