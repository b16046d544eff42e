// lichee_search.cu -- B200 (sm_100a) lineage-tree search behind the C ABI in include/lichee_b200.h.
//
// What it replaces in the reference (LICHeE/src/lineage):
//   PHYNetwork.getLineageTrees / grow      PHYNetwork.java:443-565  spanning-tree enumeration
//   PHYTree.checkConstraint                PHYTree.java:156-174     per-sample VAF sum rule
//   PHYTree.computeErrorScore / compareTo  PHYTree.java:214-237     error score, ranking
//   PHYNetwork.evaluateLineageTrees        PHYNetwork.java:726-733  sort, keep the top -s
//
// Design (DESIGN.md has the long version):
//   * The constraint network is a DAG with the germline root as its only source, so a spanning
//     tree is exactly one parent choice per non-root node.  Nodes are placed in a fixed order
//     sigma (descending VAF mass); a frontier item is the byte string of parents chosen so far.
//   * One WARP extends one partial tree: lanes own samples (1 or 2 per lane), the centroid
//     matrix lives in shared memory, the children of a candidate parent are found by comparing
//     the item's parent bytes against the candidate and voting with __ballot_sync, and the
//     per-sample sums are compared against parent VAF + margin with a warp vote.  A warp walks a
//     contiguous run of items and re-decides only what changed between lexicographic neighbours.
//   * The frontier is kept in HBM level by level and walked depth-first in chunks, so memory is
//     bounded by (levels x level_capacity) and trees come out in a deterministic canonical order.
//     check -> exclusive scan -> emit keeps that order without atomics.
//   * Narrow chunks (the first descent, small networks) are scheduled on the device: one block walks
//     the levels on its own (descend_kernel) instead of three launches and a host round trip per level.
//   * The level above the last one (check2_kernel / emit2_kernel) decides the last node's pass masks
//     for all its children with bit operations and totals each item's excess: a lower bound of the
//     score of every tree below it.  Items whose bound already exceeds the s-th best kept score are
//     counted and numbered, never read or scored again.
//   * The last level (leaf_kernel) scores the remaining trees in double precision with a fixed-shape
//     reduction, all trees of an item at once, filtered against the current s-th best score without
//     a square root; survivors go through an exact radix selection and a rank-merge into the
//     device-resident ranked list.
//   No tensor cores: nothing here is a dense contraction.  No CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lichee_b200.h"
#include "lichee_replay.cuh"

namespace {

constexpr int kWarpsPerBlock = 16;                 // 512 threads share one copy of the matrix
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kMaxN = LICHEE_MAX_NODES;
constexpr int kMaxIncremental = 12;                // more changed parent rows than this: recompute the item
constexpr int kCandBatch = 256;                    // candidate slots a warp reserves at a time
constexpr int kSelThreads = 1024;
constexpr int kSelTile = 1024;
constexpr unsigned kFull = 0xffffffffu;

thread_local std::string g_err;

void set_err(const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            set_err("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return LICHEE_ERR_CUDA;                                                      \
        }                                                                                \
    } while (0)

// ------------------------------------------------------------------------------------------
// device-side description of one network
// ------------------------------------------------------------------------------------------
struct DevNet {
    const double *vaf;       // [n][SP]   centroid matrix, row q: q = 0 germline root, q = t+1 node sigma[t]
    const uint8_t *cand;     // [L][128]  candidate parents of sigma[t] as row indices q, in ascending NODE id
    const uint8_t *ncand;    // [L]
    const uint4 *static_ok;  // [L]       bit i: candidate i passes the sum rule when it has no other child
    const double *leaf_static; // [128]   last level: excess E' of candidate i when the last node is its only child
    int n, S, SP, L, stride; // stride = bytes per frontier item (multiple of 16, <= 128)
    int prune;               // every VAF >= 0: a partial tree's excess is a lower bound of every completion's
    double margin;
    uint4 static_both;       // level L-2, bit i: childless candidate i of node v also takes the last node u
    const uint8_t *desc_rows; // [n]     rows in DESCENDING node id: the parent order of computeErrorScore (PHYNode.compareTo)
};

// Frontier item layout: an item is W = stride/4 words; the parent (row index q) of tree position t
// lives in byte t/W of word t%W, so lane l of the warp that loads the item holds positions l, l+W,
// l+2W, l+3W in bytes 0..3 of its word and the children of a parent come out of <= 4 ballots in
// position order.
__host__ __device__ __forceinline__ int item_byte(int t, int W) { return 4 * (t % W) + t / W; }

struct Cand {                // one scored complete tree that beat the current threshold
    double score;
    long long seq;           // position in the canonical enumeration order (of this part)
    int part;                // multi-GPU slice the tree came from
    unsigned origin;         // index of the leaf-parent item (new) or slot | 0x80000000 (kept)
    unsigned pstar;          // parent chosen for the last node (new entries)
    unsigned pad;
};

struct TopMeta {             // device-resident state of the ranked list
    unsigned count;          // trees currently held
    unsigned pad;
    double tau;              // score of the worst held tree once the list is full, else +inf
    unsigned long long n_cand;  // candidates appended by the current leaf launch
    unsigned long long n_valid; // valid complete trees counted by the current leaf launch
    unsigned long long next_run; // work queue of the current leaf launch: next unclaimed run of items
    unsigned long long n_bounded; // items of the current leaf launch whose trees were excluded by the lower bound
    unsigned long long n_eval;   // trees of the current leaf launch whose total was computed
    unsigned long long kmin, kmax; // smallest / largest score bits among the candidates of the current leaf launch
};

struct ScanOut {             // written by the scan, read back by the host once per chunk
    unsigned long long fit;  // (items that fit) << 32 | children they produce
    unsigned long long total;
    unsigned long long live; // emit2_kernel: last-level items written that are not bounded (zeroed by the host)
    unsigned long long trees; // emit2_kernel: valid complete trees below the items written
};

__device__ __forceinline__ bool cand_less(const Cand &a, const Cand &b) {
    if (a.score != b.score) return a.score < b.score;
    if (a.part != b.part) return a.part < b.part;
    return a.seq < b.seq;
}

// Shared-memory staging of the centroid matrix and the candidate table of one level.  The matrix (up
// to 64 KB) comes in with ONE TMA bulk copy (cp.async.bulk -> SASS UBLKCP) that completes on an
// mbarrier, instead of every thread looping over it; the 128-byte candidate table is loaded by the
// threads while the copy is in flight.
__device__ __forceinline__ void stage_network(const DevNet &net, int level, double *vaf_s, uint8_t *cand_s) {
    __shared__ __align__(8) unsigned long long bar;
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar);
    const unsigned dst_a = (unsigned)__cvta_generic_to_shared(vaf_s);
    const unsigned bytes = (unsigned)(net.n * net.SP * sizeof(double));   // multiple of 256
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst_a), "l"(net.vaf), "r"(bytes), "r"(bar_a) : "memory");
    }
    for (int i = threadIdx.x; i < kMaxN; i += blockDim.x) cand_s[i] = net.cand[(size_t)level * kMaxN + i];
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar_a) : "memory");
    }
    __syncthreads();
}

// Row q of the centroid matrix for this lane: samples `lane` and `lane+32`.  With two samples per lane
// the halves are interleaved in shared memory ({x[l], x[l+32]} adjacent) so one 128-bit load gets both.
template <int SPL>
__device__ __forceinline__ void load_row(const double *vaf_s, int q, int SP, int lane, double &x0, double &x1) {
    if (SPL == 2) {
        const double2 v = reinterpret_cast<const double2 *>(vaf_s)[q * 32 + lane];
        x0 = v.x;
        x1 = v.y;
    } else {
        x0 = vaf_s[q * SP + lane];
        x1 = 0.0;
    }
}

// Sum of the centroid rows of the children of parent q inside item word-vector w, in ascending tree
// position: the canonical accumulation order.  rmax = highest byte lane that can hold an
// assigned position.  Returns the sums for samples lane and lane+32.
template <int SPL>
__device__ __forceinline__ void child_sums(unsigned w, unsigned q, int rmax, int W, const double *vaf_s, int SP,
                                           int lane, double &s0, double &s1) {
    const unsigned m = __vcmpeq4(w, q * 0x01010101u);
    s0 = 0.0;
    s1 = 0.0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        if (r > rmax) break;
        unsigned kids = __ballot_sync(kFull, (m >> (8 * r)) & 1u);
        while (kids) {
            const int row = r * W + __ffs(kids);             // position + 1 = row of the child
            kids &= kids - 1;
            double x0, x1;
            load_row<SPL>(vaf_s, row, SP, lane, x0, x1);
            s0 = __dadd_rn(s0, x0);
            if (SPL == 2) s1 = __dadd_rn(s1, x1);
        }
    }
}

// 128-bit set of the rows that are the parent of at least one assigned position of the item.
__device__ __forceinline__ void parent_presence(unsigned w, int rmax, unsigned pm[4]) {
    unsigned m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    for (int r = 0; r <= rmax; r++) {
        const unsigned b = (w >> (8 * r)) & 0xffu;
        const unsigned bit = b < 128u ? 1u << (b & 31) : 0u;
        const unsigned k = b >> 5;
        m0 |= k == 0 ? bit : 0u;
        m1 |= k == 1 ? bit : 0u;
        m2 |= k == 2 ? bit : 0u;
        m3 |= k == 3 ? bit : 0u;
    }
    pm[0] = __reduce_or_sync(kFull, m0);
    pm[1] = __reduce_or_sync(kFull, m1);
    pm[2] = __reduce_or_sync(kFull, m2);
    pm[3] = __reduce_or_sync(kFull, m3);
}

__device__ __forceinline__ bool get_bit(const uint4 &m, unsigned i) {
    const unsigned w = (i >> 5) == 0 ? m.x : (i >> 5) == 1 ? m.y : (i >> 5) == 2 ? m.z : m.w;
    return ((w >> (i & 31)) & 1u) != 0;
}

// Sum rule for candidate row q that already has children in item w: children + the new node (row
// xv) against VAF(q) + margin in every sample.  One copy of the code (the kernels call it from
// several unrolled places; inlined, they outgrow the instruction cache).
template <int SPL>
__device__ __noinline__ bool row_passes(unsigned w, unsigned q, int rmax, int W, const double *vaf_s, int SP, int lane,
                                        double xv0, double xv1, double margin, bool valid0, bool valid1) {
    double s0, s1, a0, a1;
    child_sums<SPL>(w, q, rmax, W, vaf_s, SP, lane, s0, s1);
    load_row<SPL>(vaf_s, q, SP, lane, a0, a1);
    s0 = __dadd_rn(s0, xv0);
    bool bad = valid0 && s0 > __dadd_rn(a0, margin);
    if (SPL == 2) {
        s1 = __dadd_rn(s1, xv1);
        bad |= valid1 && s1 > __dadd_rn(a1, margin);
    }
    return !__any_sync(kFull, bad);
}

// ------------------------------------------------------------------------------------------
// kernel 1: sum-rule check.  One warp per frontier item; for every candidate parent of the node
// placed at this level, test  sum(children VAF) + VAF(node) <= VAF(parent) + margin  per sample.
// Output: 128-bit pass mask + pass count per item.
// ------------------------------------------------------------------------------------------
template <int SPL>
__global__ void __launch_bounds__(kThreads, 3)
check_kernel(DevNet net, int level, const unsigned *__restrict__ items, long long n_items,
             uint4 *__restrict__ mask_out, unsigned *__restrict__ count_out, int run_len) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *vaf_s = reinterpret_cast<double *>(smem_raw);
    uint8_t *cand_s = reinterpret_cast<uint8_t *>(vaf_s + net.n * net.SP);
    __shared__ uint8_t cidx_s[kMaxN];          // by row: candidate index (valid where candq has the bit)
    __shared__ unsigned candq_s[4], okq_s[4];  // rows that are candidates / that pass with no other child
    const int k = net.ncand[level];
    if (threadIdx.x < 4) { candq_s[threadIdx.x] = 0; okq_s[threadIdx.x] = 0; }
    stage_network(net, level, vaf_s, cand_s);
    const uint4 sok4 = net.static_ok[level];
    const unsigned sok[4] = {sok4.x, sok4.y, sok4.z, sok4.w};
    if (threadIdx.x < k) {
        const unsigned q = cand_s[threadIdx.x];
        cidx_s[q] = (uint8_t)threadIdx.x;
        atomicOr(&candq_s[q >> 5], 1u << (q & 31));
        const unsigned sw = (threadIdx.x >> 5) == 0 ? sok4.x : (threadIdx.x >> 5) == 1 ? sok4.y : (threadIdx.x >> 5) == 2 ? sok4.z : sok4.w;
        if ((sw >> (threadIdx.x & 31)) & 1u) atomicOr(&okq_s[q >> 5], 1u << (q & 31));
    }
    __syncthreads();
    const unsigned candq[4] = {candq_s[0], candq_s[1], candq_s[2], candq_s[3]};
    const unsigned okq[4] = {okq_s[0], okq_s[1], okq_s[2], okq_s[3]};

    const int lane = threadIdx.x & 31;
    const int SP = net.SP;
    const int wstride = net.stride >> 2;
    const int rmax = level > 0 ? (level - 1) / wstride : -1;
    double xv0, xv1;
    load_row<SPL>(vaf_s, level + 1, SP, lane, xv0, xv1);
    const bool valid0 = lane < net.S;
    const bool valid1 = SPL == 2 && lane + 32 < net.S;
    const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;

    auto passes = [&](unsigned w, unsigned q) {
        return row_passes<SPL>(w, q, rmax, wstride, vaf_s, SP, lane, xv0, xv1, net.margin, valid0, valid1);
    };

    // Like the last-level kernel, a warp walks a contiguous run of items and, after the first one,
    // re-decides only the candidates whose set of children changed.
    unsigned bits[4] = {0, 0, 0, 0}, um[4] = {0, 0, 0, 0};
    unsigned w_prev = 0;
    for (long long run0 = warp0 * run_len; run0 < n_items; run0 += nwarps * run_len) {
        const long long run1 = run0 + run_len < n_items ? run0 + run_len : n_items;
        for (long long it = run0; it < run1; it++) {
            const unsigned w = lane < wstride ? __ldg(items + it * wstride + lane) : 0xffffffffu;
            bool fresh = it == run0;
            unsigned todo[4] = {0, 0, 0, 0};
            if (!fresh) {
                unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                const unsigned diff = w ^ w_prev;
                for (int r = 0; r <= rmax; r++) {
                    if ((diff >> (8 * r)) & 0xffu) {
                        const unsigned bo = (w_prev >> (8 * r)) & 0xffu, bn = (w >> (8 * r)) & 0xffu;
                        const unsigned bito = bo < 128u ? 1u << (bo & 31) : 0u, bitn = bn < 128u ? 1u << (bn & 31) : 0u;
                        c0 |= ((bo >> 5) == 0 ? bito : 0u) | ((bn >> 5) == 0 ? bitn : 0u);
                        c1 |= ((bo >> 5) == 1 ? bito : 0u) | ((bn >> 5) == 1 ? bitn : 0u);
                        c2 |= ((bo >> 5) == 2 ? bito : 0u) | ((bn >> 5) == 2 ? bitn : 0u);
                        c3 |= ((bo >> 5) == 3 ? bito : 0u) | ((bn >> 5) == 3 ? bitn : 0u);
                    }
                }
                todo[0] = __reduce_or_sync(kFull, c0); todo[1] = __reduce_or_sync(kFull, c1);
                todo[2] = __reduce_or_sync(kFull, c2); todo[3] = __reduce_or_sync(kFull, c3);
                if (__popc(todo[0]) + __popc(todo[1]) + __popc(todo[2]) + __popc(todo[3]) > kMaxIncremental) fresh = true;
            }
            w_prev = w;
            if (fresh) {
                parent_presence(w, rmax, um);
#pragma unroll
                for (int g = 0; g < 4; g++) {               // static indices keep the masks in registers
                    unsigned bg = 0;
                    const int hi = k - g * 32 < 32 ? k - g * 32 : 32;
                    for (int j = 0; j < hi; j++) {
                        const unsigned q = cand_s[g * 32 + j];
                        const unsigned pw = (q >> 5) == 0 ? um[0] : (q >> 5) == 1 ? um[1] : (q >> 5) == 2 ? um[2] : um[3];
                        const bool pass = ((pw >> (q & 31)) & 1u) ? passes(w, q) : ((sok[g] >> j) & 1u) != 0;
                        if (pass) bg |= 1u << j;
                    }
                    bits[g] = bg;
                }
            } else {
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    unsigned rows = todo[g];
                    while (rows) {
                        const int bq = __ffs(rows) - 1;
                        rows &= rows - 1;
                        const unsigned q = g * 32 + bq;
                        const bool here = __ballot_sync(kFull, __vcmpeq4(w, q * 0x01010101u) != 0) != 0;
                        um[g] = here ? um[g] | (1u << bq) : um[g] & ~(1u << bq);
                        if ((candq[g] >> bq) & 1u) {
                            const bool pass = here ? passes(w, q) : ((okq[g] >> bq) & 1u) != 0;
                            const unsigned i = cidx_s[q], bit = 1u << (i & 31);
                            bits[0] = (i >> 5) == 0 ? (pass ? bits[0] | bit : bits[0] & ~bit) : bits[0];
                            bits[1] = (i >> 5) == 1 ? (pass ? bits[1] | bit : bits[1] & ~bit) : bits[1];
                            bits[2] = (i >> 5) == 2 ? (pass ? bits[2] | bit : bits[2] & ~bit) : bits[2];
                            bits[3] = (i >> 5) == 3 ? (pass ? bits[3] | bit : bits[3] & ~bit) : bits[3];
                        }
                    }
                }
            }
            if (lane == 0) {
                mask_out[it] = make_uint4(bits[0], bits[1], bits[2], bits[3]);
                count_out[it] = __popc(bits[0]) + __popc(bits[1]) + __popc(bits[2]) + __popc(bits[3]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// kernels 2a-2c: exclusive scan of the pass counts (keeps the canonical order without atomics)
// and the largest prefix of items whose children fit in the next level's buffer.
// ------------------------------------------------------------------------------------------
constexpr unsigned kCountMask = 0x3fffffffu;   // last-level counts carry a flag bit
constexpr unsigned kBounded = 0x40000000u;     // no tree below this item can enter the ranked list
// what a scan adds up: the plain count, or (last level, live mode) the trees of the items that are not bounded, i.e.
// an upper bound on the candidates a chunk can append
// Live mode with the stored bounds (last level): an item whose base's bound exceeds what the CURRENT threshold
// admits yields no candidate either -- the same test leaf_kernel applies before it touches the item -- so the chunk
// extends over everything a burst earlier in the fill has made hopeless.
__device__ __forceinline__ double admit_bound(const TopMeta *meta, int num_save, bool &full, double &tau);
__device__ __forceinline__ double live_tmax(int live_only, const double *lb, const TopMeta *meta, int num_save) {
    if (!live_only || !lb) return INFINITY;
    bool full;
    double tau;
    return admit_bound(meta, num_save, full, tau);
}
__device__ __forceinline__ unsigned scan_value(unsigned x, int live_only, const double *lb, long long idx, double tmax) {
    if (!live_only) return x & kCountMask;
    if (x & kBounded) return 0u;
    if (lb && __ldg(lb + idx) > tmax) return 0u;
    return x & kCountMask;
}
constexpr int kScanThreads = 256;
constexpr int kScanPer = 8;
constexpr int kScanTile = kScanThreads * kScanPer;

__global__ void __launch_bounds__(kScanThreads)
scan_tiles(const unsigned *__restrict__ in, unsigned *__restrict__ out, unsigned long long *tile_sum,
           long long n, int live_only, const double *__restrict__ lb, const TopMeta *meta, int num_save) {
    __shared__ unsigned warp_tot[kScanThreads / 32];
    const double tmax = live_tmax(live_only, lb, meta, num_save);
    const long long base = (long long)blockIdx.x * kScanTile + threadIdx.x * kScanPer;
    unsigned v[kScanPer], run = 0;
#pragma unroll
    for (int i = 0; i < kScanPer; i++) {
        v[i] = base + i < n ? scan_value(in[base + i], live_only, lb, base + i, tmax) : 0;
        run += v[i];
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned inc = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned t = __shfl_up_sync(kFull, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    unsigned woff = 0;
    for (int i = 0; i < wid; i++) woff += warp_tot[i];
    unsigned ex = woff + inc - run;
#pragma unroll
    for (int i = 0; i < kScanPer; i++) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
    if (threadIdx.x == kScanThreads - 1) tile_sum[blockIdx.x] = woff + inc;
}

__global__ void scan_tile_sums(unsigned long long *tile_sum, int n_tiles, ScanOut *out) {
    // one warp: exclusive scan of the tile totals, 32 at a time
    const int lane = threadIdx.x & 31;
    if (threadIdx.x >= 32) return;
    unsigned long long run = 0;
    for (int base = 0; base < n_tiles; base += 32) {
        const int i = base + lane;
        const unsigned long long t = i < n_tiles ? tile_sum[i] : 0ull;
        unsigned long long inc = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long u = __shfl_up_sync(kFull, inc, d);
            if (lane >= d) inc += u;
        }
        if (i < n_tiles) tile_sum[i] = run + inc - t;
        run += __shfl_sync(kFull, inc, 31);
    }
    if (lane == 0) {
        out->total = run;
        out->fit = 0;
    }
}

__global__ void __launch_bounds__(kScanThreads)
scan_finish(const unsigned *__restrict__ cnt, unsigned *__restrict__ off,
            const unsigned long long *__restrict__ tile_sum, long long n, unsigned long long limit,
            ScanOut *out, int live_only, const double *__restrict__ lb, const TopMeta *meta, int num_save) {
    const double tmax = live_tmax(live_only, lb, meta, num_save);
    const long long base = (long long)blockIdx.x * kScanTile + threadIdx.x * kScanPer;
    const unsigned long long add = tile_sum[blockIdx.x];
    unsigned long long best = 0;
#pragma unroll
    for (int i = 0; i < kScanPer; i++) {
        const long long idx = base + i;
        if (idx < n) {
            const unsigned long long ex = add + off[idx];
            off[idx] = (unsigned)ex;
            const unsigned long long inc = ex + scan_value(cnt[idx], live_only, lb, idx, tmax);
            if (inc <= limit) best = ((unsigned long long)(idx + 1) << 32) | inc;
        }
    }
    // one atomic per block (every thread hitting the same word made this the slowest part of the scan)
    __shared__ unsigned long long wbest[kScanThreads / 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(kFull, best, d);
        best = o > best ? o : best;
    }
    if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < kScanThreads / 32; i++) best = wbest[i] > best ? wbest[i] : best;
        if (best) atomicMax(&out->fit, best);
    }
}

// Single-launch variant for narrow chunks (the first descent, small networks): one block, each
// thread owns a contiguous run.  Wider chunks use the three-kernel tiled scan, which is faster there.
constexpr int kScanSingleThreads = 1024;
constexpr long long kScanSingleMax = 4096;

__global__ void __launch_bounds__(kScanSingleThreads)
scan_single(const unsigned *__restrict__ cnt, unsigned *__restrict__ off, long long n, unsigned long long limit,
            ScanOut *out, int live_only, const double *__restrict__ lb, const TopMeta *meta, int num_save) {
    const double tmax = live_tmax(live_only, lb, meta, num_save);
    __shared__ unsigned warp_tot[kScanSingleThreads / 32];
    __shared__ unsigned long long best_s;
    if (threadIdx.x == 0) best_s = 0;
    const long long per = (n + kScanSingleThreads - 1) / kScanSingleThreads;
    const long long lo = (long long)threadIdx.x * per;
    const long long hi = lo + per < n ? lo + per : n;
    unsigned run = 0;
    for (long long i = lo; i < hi; i++) run += scan_value(cnt[i], live_only, lb, i, tmax);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned inc = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(kFull, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    unsigned woff = 0, total = 0;
    for (int i = 0; i < kScanSingleThreads / 32; i++) {
        if (i < wid) woff += warp_tot[i];
        total += warp_tot[i];
    }
    unsigned long long ex = woff + inc - run, best = 0;
    for (long long i = lo; i < hi; i++) {
        off[i] = (unsigned)ex;
        ex += scan_value(cnt[i], live_only, lb, i, tmax);
        if (ex <= limit) best = ((unsigned long long)(i + 1) << 32) | ex;
    }
    if (best) atomicMax(&best_s, best);
    __syncthreads();
    if (threadIdx.x == 0) { out->fit = best_s; out->total = total; }
}

// ------------------------------------------------------------------------------------------
// kernel 3: emit.  One warp per parent item writes its surviving children, 128 coalesced bytes
// per child, at the offsets the scan assigned (so children appear in canonical order).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
emit_kernel(DevNet net, int level, const unsigned *__restrict__ items, long long n_items,
            const uint4 *__restrict__ mask, const unsigned *__restrict__ off,
            unsigned *__restrict__ out) {
    __shared__ uint8_t cand_s[kMaxN];
    for (int i = threadIdx.x; i < kMaxN; i += blockDim.x) cand_s[i] = net.cand[(size_t)level * kMaxN + i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wstride = net.stride >> 2;
    const int wl = item_byte(level, wstride) >> 2, sh = 8 * (item_byte(level, wstride) & 3);
    const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
    for (long long it = warp0; it < n_items; it += nwarps) {
        const uint4 m4 = mask[it];
        if ((m4.x | m4.y | m4.z | m4.w) == 0) continue;
        const unsigned w = lane < wstride ? __ldg(items + it * wstride + lane) : 0xffffffffu;
        long long dst = (long long)off[it];
        const unsigned mw[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
        for (int g = 0; g < 4; g++) {
            unsigned b = mw[g];
            while (b) {
                const int i = g * 32 + __ffs(b) - 1;
                b &= b - 1;
                unsigned cw = w;
                if (lane == wl) cw = (w & ~(0xffu << sh)) | ((unsigned)cand_s[i] << sh);
                if (lane < wstride) out[dst * wstride + lane] = cw;
                dst++;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// kernel 3b: device-side depth-first scheduling of NARROW chunks.  While the frontier is a handful of
// items per level (the first descent of every search, the whole search of a real-data network) a
// host-driven check -> scan -> emit costs three launches and one host round trip per tree level, ~60 us
// for a few microseconds of work.  Here ONE block stages the centroid matrix once and then runs the
// driver loop itself: pick the deepest non-empty level, size the chunk with the same budget rule as the
// host, check (one warp per (item, slice of the candidate list)), scan in one warp, emit, repeat.  It
// returns to the host when the deepest level is the last one (the fused check+score kernel takes
// over), when a chunk would be wide enough for the grid-wide kernels, or when the search is over.
// Children are written in canonical order like the wide path, so which path ran a chunk is invisible
// in the results.
// ------------------------------------------------------------------------------------------
constexpr int kNarrowThreads = 1024;
constexpr int kNarrowWarps = kNarrowThreads / 32;
constexpr int kNarrowMax = 128;        // items per narrow chunk
constexpr int kNarrowCap = 2048;       // children a narrow chunk may emit (every level buffer holds at least this many)
constexpr int kNarrowIters = 8192;     // chunks per launch (keeps the host's view of the counters fresh)
constexpr double kNarrowSlack = 2.0;   // items a narrow chunk takes beyond what the tree budget asks for (the host's wide chunks: 8;
                                       // measured: the first descent of 118 levels costs 1.0 ms with 8, 0.6 ms with 2)

enum { kStopDone = 0, kStopLeaf = 1, kStopWide = 2, kStopGrow = 3, kStopIters = 4 };

struct Sched {                         // scheduler state, uploaded before and read back after a launch
    long long head[kMaxN], tail[kMaxN];
    unsigned *ptr[kMaxN];              // level buffers
    double fan[kMaxN];                 // observed mean number of surviving children per item of a level
    int seen[kMaxN];
    long long trees_left;              // max_num_trees - valid trees found so far
    long long grow_calls, max_grow_calls;
    long long chunk_max, cap;
    long long bfs_width;               // > 0: breadth-first prologue, whole levels until one is this wide
    long long checks, frontier_bytes, check_bytes, emit_bytes, chunks;
    int stop, stop_level;
};

template <int SPL>
__global__ void __launch_bounds__(kNarrowThreads, 1)
descend_kernel(DevNet net, Sched *sc) {
    // dynamic shared memory: centroid matrix | 128 bytes (stage_network's table slot) | candidate tables of
    // every level
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *vaf_s = reinterpret_cast<double *>(smem_raw);
    uint8_t *cand0_s = reinterpret_cast<uint8_t *>(vaf_s + net.n * net.SP);
    uint8_t *cand_all = cand0_s + kMaxN;                                  // [L][128]
    __shared__ long long head_s[kMaxN], tail_s[kMaxN];
    __shared__ double fan_s[kMaxN];
    __shared__ int seen_s[kMaxN];
    __shared__ uint8_t ncand_s[kMaxN];
    __shared__ signed char rmax_s[kMaxN];                                 // highest byte lane in use at a level
    __shared__ uint8_t wl_s[kMaxN], sh_s[kMaxN];                          // word / bit shift of a level's own byte
    __shared__ unsigned *ptr_s[kMaxN];
    __shared__ unsigned item_s[kNarrowMax * 32];                          // the chunk's items
    __shared__ unsigned mask_s[kNarrowMax * 4], cnt_s[kNarrowMax], off_s[kNarrowMax];
    __shared__ int lv_s, B_s, stop_s, nfit_s;
    __shared__ unsigned total_s;

    const int L = net.L;
    const int W = net.stride >> 2;
    stage_network(net, 0, vaf_s, cand0_s);
    for (int i = threadIdx.x; i < L * (kMaxN / 4); i += blockDim.x)
        reinterpret_cast<unsigned *>(cand_all)[i] = reinterpret_cast<const unsigned *>(net.cand)[i];
    for (int i = threadIdx.x; i < kMaxN; i += blockDim.x) {
        head_s[i] = sc->head[i]; tail_s[i] = sc->tail[i]; fan_s[i] = sc->fan[i]; seen_s[i] = sc->seen[i];
        ncand_s[i] = i < L ? net.ncand[i] : 0;
        ptr_s[i] = sc->ptr[i];
        rmax_s[i] = (signed char)(i > 0 ? (i - 1) / W : -1);
        wl_s[i] = (uint8_t)(item_byte(i, W) >> 2);
        sh_s[i] = (uint8_t)(8 * (item_byte(i, W) & 3));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int SP = net.SP;
    const bool valid0 = lane < net.S;
    const bool valid1 = SPL == 2 && lane + 32 < net.S;
    const long long cap = sc->cap, chunk_max = sc->chunk_max, trees_left = sc->trees_left, max_grow = sc->max_grow_calls;
    const long long bfs_width = sc->bfs_width;
    const long long limit = cap < (long long)kNarrowCap ? cap : (long long)kNarrowCap;
    // scheduler registers (meaningful in warp 0 only)
    long long grow = sc->grow_calls, checks = 0, fbytes = 0, cbytes = 0, ebytes = 0, chunks = 0;
    int hint = L - 1;                  // every level deeper than this is empty

    // Warp 0 picks the next chunk: deepest non-empty level, sized by the host's rule (run_search): what is
    // pending, cut to what the remaining tree budget can use and to what the next level is expected to hold.
    auto schedule = [&](int iter) {
        int lv = -1;
        for (int l = hint; l >= 0; l--) if (tail_s[l] > head_s[l]) { lv = l; break; }
        int stop = -1;
        long long B = 0;
        if (lv < 0) stop = kStopDone;
        else if (lv >= L - 2) stop = kStopLeaf;        // the last two levels belong to check2 / emit2 / leaf
        else if (grow >= max_grow) stop = kStopGrow;
        else if (iter >= kNarrowIters) stop = kStopIters;
        else if (bfs_width > 0) {
            // prologue of a sliced search: the whole level or nothing, until a level is wide enough to be cut
            // (the host applies the same test before it launches, so the first chunk always goes through)
            const long long pending = tail_s[lv] - head_s[lv];
            if (pending >= bfs_width || pending > kNarrowMax || pending * (long long)ncand_s[lv] > limit) stop = kStopWide;
            else B = pending;
        } else {
            // expected complete trees below one item of this level: product of the fan-outs from here down
            double prod = 1.0;
            for (int l = lv + lane; l < L; l += 32) prod *= fmax(seen_s[l] ? fan_s[l] : (double)ncand_s[l], 1e-3);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) prod *= __shfl_xor_sync(kFull, prod, d);
            const double yield = fmin(1e30, prod);
            const long long pending = tail_s[lv] - head_s[lv];
            B = pending < chunk_max ? pending : chunk_max;
            const double want = 1.25 * (double)trees_left / yield + kNarrowSlack;
            if ((double)B > want) B = (long long)want;
            const long long guess = (long long)(0.8 * (double)cap / fmax(1.0, fan_s[lv]));
            const long long gmax = guess > 1024 ? guess : 1024;
            if (B > gmax) B = gmax;
            if (B < 1) B = 1;
            if (B > kNarrowMax) {
                if (iter == 0) B = kNarrowMax;      // the host asked for this chunk: always make progress
                else stop = kStopWide;
            }
        }
        if (lane == 0) { lv_s = lv; B_s = (int)B; stop_s = stop; }
    };
    if (wid == 0) schedule(0);
    __syncthreads();

    for (int iter = 1; stop_s < 0; iter++) {
        const int lv = lv_s, B = B_s;
        const int k = ncand_s[lv];
        const uint8_t *cand_s = cand_all + (size_t)lv * kMaxN;
        const int rmax = rmax_s[lv];
        {
            // stage the chunk's items (plain loads through L2: they may have been written by this block)
            const unsigned *items = ptr_s[lv] + (size_t)head_s[lv] * W;
            for (int i = threadIdx.x; i < B * 32; i += blockDim.x)
                if ((i & 31) < W) item_s[i] = __ldcg(items + (size_t)(i >> 5) * W + (i & 31));
            for (int i = threadIdx.x; i < B * 4; i += blockDim.x) mask_s[i] = 0;
        }
        __syncthreads();

        // check: a task is (item, slice of the candidate list); few items -> more slices per item (a single item is
        // split over all 32 warps).
        int lp = 0;
        while (lp < 5 && (B << (lp + 1)) <= 2 * kNarrowWarps && (2 << lp) <= k) lp++;
        const uint4 sok4 = net.static_ok[lv];
        for (int task = wid; task < (B << lp); task += kNarrowWarps) {
            const int it = task >> lp, part = task & ((1 << lp) - 1);
            const int c0 = (k * part) >> lp, c1 = (k * (part + 1)) >> lp;
            const unsigned w = lane < W ? item_s[it * 32 + lane] : 0xffffffffu;
            double xv0, xv1;
            load_row<SPL>(vaf_s, lv + 1, SP, lane, xv0, xv1);
            unsigned pm[4];
            parent_presence(w, rmax, pm);
            unsigned long long blo = 0, bhi = 0;
            for (int j = c0; j < c1; j++) {
                // a candidate without a child in the item: the precomputed answer (the same doubles, 0 + VAF(node)
                // against VAF(parent) + margin)
                const unsigned q = cand_s[j];
                const unsigned pw = (q >> 5) == 0 ? pm[0] : (q >> 5) == 1 ? pm[1] : (q >> 5) == 2 ? pm[2] : pm[3];
                const bool pass = ((pw >> (q & 31)) & 1u)
                                      ? row_passes<SPL>(w, q, rmax, W, vaf_s, SP, lane, xv0, xv1, net.margin, valid0, valid1)
                                      : get_bit(sok4, (unsigned)j);
                const unsigned long long bit = (unsigned long long)(pass ? 1u : 0u) << (j & 63);
                if (j < 64) blo |= bit; else bhi |= bit;
            }
            if (lane == 0) {
                if ((unsigned)blo) atomicOr(&mask_s[it * 4 + 0], (unsigned)blo);
                if ((unsigned)(blo >> 32)) atomicOr(&mask_s[it * 4 + 1], (unsigned)(blo >> 32));
                if ((unsigned)bhi) atomicOr(&mask_s[it * 4 + 2], (unsigned)bhi);
                if ((unsigned)(bhi >> 32)) atomicOr(&mask_s[it * 4 + 3], (unsigned)(bhi >> 32));
            }
        }
        __syncthreads();

        // scan (one warp, four items per lane): child offsets in canonical order and the longest prefix
        // of items whose children fit the next level
        if (wid == 0) {
            unsigned v[4], run = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int idx = lane * 4 + i;
                v[i] = idx < B ? __popc(mask_s[idx * 4]) + __popc(mask_s[idx * 4 + 1]) + __popc(mask_s[idx * 4 + 2]) + __popc(mask_s[idx * 4 + 3]) : 0u;
                run += v[i];
            }
            unsigned inc = run;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned t = __shfl_up_sync(kFull, inc, d);
                if (lane >= d) inc += t;
            }
            unsigned ex = inc - run;
            int fit = 0;
            unsigned fit_total = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int idx = lane * 4 + i;
                if (idx < B) {
                    cnt_s[idx] = v[i];
                    off_s[idx] = ex;
                    ex += v[i];
                    if ((long long)ex <= limit) { fit = idx + 1; fit_total = ex; }
                }
            }
            // inclusive sums are non-decreasing, so the fitting prefix is the largest fitting index
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const int f2 = __shfl_xor_sync(kFull, fit, d);
                const unsigned t2 = __shfl_xor_sync(kFull, fit_total, d);
                if (f2 > fit) { fit = f2; fit_total = t2; }
            }
            if (lane == 0) { nfit_s = fit; total_s = fit_total; }
        }
        __syncthreads();

        const int n_fit = nfit_s;
        const unsigned total_fit = total_s;
        unsigned *out = ptr_s[lv + 1];
        if (wid == 0) {
            // book the chunk and choose the next one while the other warps emit
            if (lane == 0) {
                head_s[lv] += n_fit;
                if (head_s[lv] == tail_s[lv]) head_s[lv] = tail_s[lv] = 0;
                head_s[lv + 1] = 0;
                tail_s[lv + 1] = total_fit;
                const double f = (double)total_fit / (double)n_fit;
                fan_s[lv] = seen_s[lv] ? 0.5 * fan_s[lv] + 0.5 * f : f;
                seen_s[lv] = 1;
            }
            checks += (long long)n_fit * k;
            grow += total_fit;
            fbytes += ((long long)n_fit + total_fit) * net.stride;
            cbytes += (long long)n_fit * (net.stride + 20);
            ebytes += (long long)n_fit * (net.stride + 20) + (long long)total_fit * net.stride;
            chunks += 1;
            hint = lv + 1;
            __syncwarp();
            schedule(iter);            // writes lv_s / B_s / stop_s: read by the block after the barrier below
        }
        // emit: the surviving children of the fitting items, one coalesced line each
        {
            const int wl = wl_s[lv], sh = sh_s[lv];
            for (int it = wid; it < n_fit; it += kNarrowWarps) {
                if (cnt_s[it] == 0) continue;
                const unsigned w = lane < W ? item_s[it * 32 + lane] : 0xffffffffu;
                size_t dst = off_s[it];
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    unsigned b = mask_s[it * 4 + g];
                    while (b) {
                        const int i = g * 32 + __ffs(b) - 1;
                        b &= b - 1;
                        unsigned cw = w;
                        if (lane == wl) cw = (w & ~(0xffu << sh)) | ((unsigned)cand_s[i] << sh);
                        if (lane < W) out[dst * W + lane] = cw;
                        dst++;
                    }
                }
            }
        }
        __syncthreads();     // the emitted items are visible to the whole block from here on
    }
    for (int i = threadIdx.x; i < kMaxN; i += blockDim.x) {
        sc->head[i] = head_s[i]; sc->tail[i] = tail_s[i]; sc->fan[i] = fan_s[i]; sc->seen[i] = seen_s[i];
    }
    if (threadIdx.x == 0) {
        sc->grow_calls = grow; sc->checks = checks; sc->frontier_bytes = fbytes;
        sc->check_bytes = cbytes; sc->emit_bytes = ebytes; sc->chunks = chunks;
        sc->stop = stop_s; sc->stop_level = lv_s;
    }
}

// ------------------------------------------------------------------------------------------
// kernel 4: last level, scoring.  A warp walks runs of last-level items (every node placed but the last);
// every passing choice p* for the last node is a complete valid tree.
//   score = sqrt( sum over parents p, samples s of max(0, sum(children of p)[s] - VAF(p)[s])^2 )
// (PHYTree.computeErrorScore, PHYTree.java:214-233).  The reference adds the positive terms one
// after another; with noisy centroids a tree has thousands of them, i.e. thousands of dependent
// double additions per tree.  The device uses a FIXED-SHAPE reduction instead (documented in
// DESIGN.md "accumulation order", restated on the CPU by lichee_canonical_search):
//   E[p]  = butterfly over the 32 lanes of ( t[p][lane] + t[p][lane+32] )      t = positive term or 0
//   total = butterfly over the 32 lanes of ( (E[lane]+E[lane+32]) + (E[lane+64]+E[lane+96]) )
// The shape does not depend on the tree, so trees with the same terms get bit-equal scores, the
// result is within a few ulp of the reference's sequential sum (tolerance 1e-9), and all trees of
// one item share every E[p] except E[p*]: O(log n) work per tree.
// Trees that beat the current s-th best score are appended to the candidate list.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double butterfly_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = __dadd_rn(v, __shfl_xor_sync(kFull, v, d));
    return v;
}

template <int SPL>
__device__ __forceinline__ double parent_excess(double s0, double s1, double a0, double a1, bool valid0,
                                                bool valid1) {
    const double d0 = __dsub_rn(s0, a0);
    double t = (valid0 && s0 > a0) ? __dmul_rn(d0, d0) : 0.0;
    if (SPL == 2) {
        const double d1 = __dsub_rn(s1, a1);
        const double t1 = (valid1 && s1 > a1) ? __dmul_rn(d1, d1) : 0.0;
        t = __dadd_rn(t, t1);
    }
    return butterfly_sum(t);
}

// One parent row of a leaf-parent item: excess E without the last node and, for a candidate parent,
// the sum-rule check and the excess F with the last node added.  Returns the pass bit.
template <int SPL>
__device__ __noinline__ bool leaf_row(unsigned w, unsigned q, int rmax, int W, const double *vaf_s, int SP, int lane,
                                      double xv0, double xv1, double margin, bool valid0, bool valid1, bool is_cand,
                                      bool want_f, double &Eq, double &Fq) {
    double s0, s1, a0, a1;
    child_sums<SPL>(w, q, rmax, W, vaf_s, SP, lane, s0, s1);
    load_row<SPL>(vaf_s, q, SP, lane, a0, a1);
    Eq = parent_excess<SPL>(s0, s1, a0, a1, valid0, valid1);
    Fq = 0.0;
    if (!is_cand) return false;
    s0 = __dadd_rn(s0, xv0);
    bool bad = valid0 && s0 > __dadd_rn(a0, margin);
    if (SPL == 2) {
        s1 = __dadd_rn(s1, xv1);
        bad |= valid1 && s1 > __dadd_rn(a1, margin);
    }
    if (__any_sync(kFull, bad)) return false;
    if (want_f) Fq = parent_excess<SPL>(s0, s1, a0, a1, valid0, valid1);
    return true;
}

// Largest total (sum of squared excesses) a tree may have and still enter the ranked list: a tree enters
// a full list only with sqrt_rn(total) < tau.  sqrt_rn is monotone, so that is total <= tmax with tmax =
// the largest double whose rounded square root is below tau: found by stepping a few ulps around
// tau*tau.  No square root for rejected trees.  +inf while the list is not full.
__device__ __forceinline__ double admit_bound(const TopMeta *meta, int num_save, bool &full, double &tau) {
    tau = meta->tau;
    full = meta->count >= (unsigned)num_save;
    double tmax = INFINITY;
    if (full) {
        double t = __dmul_rn(tau, tau);
        for (int i = 0; i < 8 && t > 0.0 && __dsqrt_rn(t) >= tau; i++) t = __longlong_as_double(__double_as_longlong(t) - 1);
        for (int i = 0; i < 8 && __dsqrt_rn(__longlong_as_double(__double_as_longlong(t) + 1)) < tau; i++)
            t = __longlong_as_double(__double_as_longlong(t) + 1);
        tmax = __dsqrt_rn(t) < tau ? t : -1.0;             // tau == 0: nothing can beat it
    }
    return tmax;
}

__device__ __forceinline__ void put_bit(unsigned m[4], unsigned i, bool v) {
    const unsigned bit = 1u << (i & 31), g = i >> 5;
    m[0] = g == 0 ? (v ? m[0] | bit : m[0] & ~bit) : m[0];
    m[1] = g == 1 ? (v ? m[1] | bit : m[1] & ~bit) : m[1];
    m[2] = g == 2 ? (v ? m[2] | bit : m[2] & ~bit) : m[2];
    m[3] = g == 3 ? (v ? m[3] | bit : m[3] & ~bit) : m[3];
}

// ------------------------------------------------------------------------------------------
// kernels 1b / 3c: the level above the last one.  Its items are the BASES of the last level: the
// children of one item differ only in where its node v hangs, and the last node u is then tried under
// every candidate.  One pass over a parent row's children decides three things at once:
//   bit 0  v fits under the row                    (the ordinary check of this level)
//   bit 1  v and then u both fit under the row      (the only last-level decision that involves v)
//   bit 2  u fits under the row without v           (the last-level decision for every other child)
// so the pass mask of every last-level item is assembled by emit2_kernel with bit operations, and the
// last-level kernel does arithmetic only for items that can still reach the ranked list: this kernel
// also totals the excess of the base (fixed-shape reduction, same shape as the scoring) and flags the
// base when that lower bound already exceeds what the full list admits.
// ------------------------------------------------------------------------------------------
template <int SPL>
__device__ __noinline__ unsigned dual_row(unsigned w, unsigned q, int rmax, int W, const double *vaf_s, int SP, int lane,
                                          double xv0, double xv1, double xu0, double xu1, double margin, bool valid0,
                                          bool valid1, bool is_cv, bool is_cu, bool want_e, double &Eq) {
    double s0, s1, a0, a1;
    child_sums<SPL>(w, q, rmax, W, vaf_s, SP, lane, s0, s1);
    load_row<SPL>(vaf_s, q, SP, lane, a0, a1);
    Eq = want_e ? parent_excess<SPL>(s0, s1, a0, a1, valid0, valid1) : 0.0;
    const double c0 = __dadd_rn(a0, margin), c1 = __dadd_rn(a1, margin);
    unsigned r = 0;
    if (is_cv) {
        const double v0 = __dadd_rn(s0, xv0), v1 = __dadd_rn(s1, xv1);
        bool bad = valid0 && v0 > c0;
        if (SPL == 2) bad |= valid1 && v1 > c1;
        if (!__any_sync(kFull, bad)) {
            r |= 1u;
            if (is_cu) {
                bool bad2 = valid0 && __dadd_rn(v0, xu0) > c0;
                if (SPL == 2) bad2 |= valid1 && __dadd_rn(v1, xu1) > c1;
                if (!__any_sync(kFull, bad2)) r |= 2u;
            }
        }
    }
    if (is_cu) {
        bool bad = valid0 && __dadd_rn(s0, xu0) > c0;
        if (SPL == 2) bad |= valid1 && __dadd_rn(s1, xu1) > c1;
        if (!__any_sync(kFull, bad)) r |= 4u;
    }
    return r;
}

template <int SPL>
__global__ void __launch_bounds__(kThreads, 2)
check2_kernel(DevNet net, const unsigned *__restrict__ items, long long n_items, uint4 *__restrict__ mask_out,
              unsigned *__restrict__ count_out, uint4 *__restrict__ both_out, uint4 *__restrict__ masku_out,
              unsigned *__restrict__ flag_out, double *__restrict__ bound_out, const TopMeta *meta, int num_save, int run_len) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *vaf_s = reinterpret_cast<double *>(smem_raw);
    uint8_t *cand_s = reinterpret_cast<uint8_t *>(vaf_s + net.n * net.SP);
    __shared__ uint8_t cidx_s[kMaxN], cidx2_s[kMaxN];     // by row: candidate index for v / for u
    __shared__ unsigned candq_s[4], candq2_s[4];          // rows that are candidates of v / of u
    const int level = net.L - 2;
    const int k = net.ncand[level], k2 = net.ncand[level + 1];
    if (threadIdx.x < 4) { candq_s[threadIdx.x] = 0; candq2_s[threadIdx.x] = 0; }
    stage_network(net, level, vaf_s, cand_s);
    if (threadIdx.x < k) {
        const unsigned q = cand_s[threadIdx.x];
        cidx_s[q] = (uint8_t)threadIdx.x;
        atomicOr(&candq_s[q >> 5], 1u << (q & 31));
    }
    if (threadIdx.x < k2) {
        const unsigned q = net.cand[(size_t)(level + 1) * kMaxN + threadIdx.x];
        cidx2_s[q] = (uint8_t)threadIdx.x;
        atomicOr(&candq2_s[q >> 5], 1u << (q & 31));
    }
    __syncthreads();
    const unsigned candq[4] = {candq_s[0], candq_s[1], candq_s[2], candq_s[3]};
    const unsigned candq2[4] = {candq2_s[0], candq2_s[1], candq2_s[2], candq2_s[3]};
    const uint4 sokv = net.static_ok[level], soku = net.static_ok[level + 1], sboth = net.static_both;

    const int lane = threadIdx.x & 31;
    const int SP = net.SP;
    const int wstride = net.stride >> 2;
    const int rmax = level > 0 ? (level - 1) / wstride : -1;
    double xv0, xv1, xu0, xu1;
    load_row<SPL>(vaf_s, level + 1, SP, lane, xv0, xv1);
    load_row<SPL>(vaf_s, level + 2, SP, lane, xu0, xu1);
    const bool valid0 = lane < net.S;
    const bool valid1 = SPL == 2 && lane + 32 < net.S;
    bool full;
    double tau;
    const double tmax = admit_bound(meta, num_save, full, tau);
    // (the bound is computed whenever the network allows it, also while the list is still filling up: this launch may
    //  run ahead of the selection that fills it, and emit2_kernel compares the bound with the threshold of ITS moment)
    const bool prune = net.prune != 0;
    const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;

    // Neighbouring items are the children of one item of the level above: they differ only in where its node x (tree
    // position L-3, the highest assigned one) hangs.  The state is therefore kept for the item WITHOUT x -- per row:
    // bv0 = v fits (by v's candidate index), bb0 = v and u both fit (same index), bu0 = u fits without v (by u's
    // candidate index), um0 = rows present, E0 = excess of the rows (lane l: rows l, l+32, ..) -- and updated only
    // when a byte other than x's changes (a new sibling group); on top of it ONE row per item is decided with x in
    // place: the parent p that x hangs under (x is the last child in p's sums, the same additions as a full
    // evaluation of the item).
    unsigned bv0[4] = {0, 0, 0, 0}, bb0[4] = {0, 0, 0, 0}, bu0[4] = {0, 0, 0, 0}, um0[4] = {0, 0, 0, 0};
    double E0[4] = {0.0, 0.0, 0.0, 0.0};
    unsigned wb_prev = 0;
    const bool has_x = level >= 1;
    const int wl3 = has_x ? item_byte(level - 1, wstride) >> 2 : 0, sh3 = has_x ? 8 * (item_byte(level - 1, wstride) & 3) : 0;
    // Reference state of the block.  A run used to start with a full evaluation of its first item (every present
    // row: a third of the kernel's instructions at ~20 items per run).  The items of a block are neighbours in the
    // canonical order and share all but their last few bytes, so the block evaluates its FIRST item together -- every
    // warp a 16th of the rows -- and each run starts from that state, redoing only the rows named by the bytes in
    // which its first item differs (or everything, when those are many).  Same arithmetic per row.
    bool have_state = false;
    {
        const long long ref_it = (long long)blockIdx.x * kWarpsPerBlock * run_len;
        if (has_x && ref_it < n_items) {
            __shared__ double eref_s[kMaxN];
            __shared__ unsigned bvref_s[4], bbref_s[4], buref_s[4];
            const int wid = threadIdx.x >> 5;
            if (threadIdx.x < 4) {
                const int t = threadIdx.x;
                bvref_s[t] = t == 0 ? sokv.x : t == 1 ? sokv.y : t == 2 ? sokv.z : sokv.w;
                bbref_s[t] = t == 0 ? sboth.x : t == 1 ? sboth.y : t == 2 ? sboth.z : sboth.w;
                buref_s[t] = t == 0 ? soku.x : t == 1 ? soku.y : t == 2 ? soku.z : soku.w;
            }
            unsigned wr = lane < wstride ? __ldg(items + ref_it * wstride + lane) : 0xffffffffu;
            if (lane == wl3) wr |= 0xffu << sh3;
            parent_presence(wr, rmax, um0);
            __syncthreads();
            int rank = 0;
#pragma unroll
            for (int g = 0; g < 4; g++) {
                unsigned rows = prune ? um0[g] : um0[g] & (candq[g] | candq2[g]);
                while (rows) {
                    const int bq = __ffs(rows) - 1;
                    rows &= rows - 1;
                    if ((rank++ & (kWarpsPerBlock - 1)) != wid) continue;
                    const unsigned q = g * 32 + bq;
                    const bool is_cv = ((candq[g] >> bq) & 1u) != 0, is_cu = ((candq2[g] >> bq) & 1u) != 0;
                    double Eq = 0.0;
                    const unsigned r = dual_row<SPL>(wr, q, rmax, wstride, vaf_s, SP, lane, xv0, xv1, xu0, xu1, net.margin, valid0,
                                                     valid1, is_cv, is_cu, prune, Eq);
                    if (lane == 0) {
                        eref_s[q] = Eq;
                        if (is_cv) {
                            const unsigned iv = cidx_s[q], bit = 1u << (iv & 31);
                            if (r & 1u) atomicOr(&bvref_s[iv >> 5], bit); else atomicAnd(&bvref_s[iv >> 5], ~bit);
                            if (r & 2u) atomicOr(&bbref_s[iv >> 5], bit); else atomicAnd(&bbref_s[iv >> 5], ~bit);
                        }
                        if (is_cu) {
                            const unsigned iu = cidx2_s[q], bit = 1u << (iu & 31);
                            if (r & 4u) atomicOr(&buref_s[iu >> 5], bit); else atomicAnd(&buref_s[iu >> 5], ~bit);
                        }
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int g = 0; g < 4; g++) {
                bv0[g] = bvref_s[g]; bb0[g] = bbref_s[g]; bu0[g] = buref_s[g];
                E0[g] = (prune && ((um0[g] >> lane) & 1u)) ? eref_s[g * 32 + lane] : 0.0;
            }
            wb_prev = wr;
            have_state = true;
        }
    }
    for (long long run0 = warp0 * run_len; run0 < n_items; run0 += nwarps * run_len) {
        const long long run1 = run0 + run_len < n_items ? run0 + run_len : n_items;
        unsigned w_next = lane < wstride ? __ldg(items + run0 * wstride + lane) : 0xffffffffu;
        for (long long it = run0; it < run1; it++) {
            const unsigned w = w_next;
            if (it + 1 < run1) w_next = lane < wstride ? __ldg(items + (it + 1) * wstride + lane) : 0xffffffffu;
            const unsigned wb = (has_x && lane == wl3) ? (w | (0xffu << sh3)) : w;     // the item with x taken out
            bool fresh = !have_state;
            have_state = true;
            unsigned todo[4] = {0, 0, 0, 0};
            if (!fresh) {
                const unsigned diff = wb ^ wb_prev;
                if (__any_sync(kFull, diff != 0)) {
                    unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                    for (int r = 0; r <= rmax; r++) {
                        if ((diff >> (8 * r)) & 0xffu) {
                            const unsigned bo = (wb_prev >> (8 * r)) & 0xffu, bn = (wb >> (8 * r)) & 0xffu;
                            const unsigned bito = bo < 128u ? 1u << (bo & 31) : 0u, bitn = bn < 128u ? 1u << (bn & 31) : 0u;
                            c0 |= ((bo >> 5) == 0 ? bito : 0u) | ((bn >> 5) == 0 ? bitn : 0u);
                            c1 |= ((bo >> 5) == 1 ? bito : 0u) | ((bn >> 5) == 1 ? bitn : 0u);
                            c2 |= ((bo >> 5) == 2 ? bito : 0u) | ((bn >> 5) == 2 ? bitn : 0u);
                            c3 |= ((bo >> 5) == 3 ? bito : 0u) | ((bn >> 5) == 3 ? bitn : 0u);
                        }
                    }
                    todo[0] = __reduce_or_sync(kFull, c0); todo[1] = __reduce_or_sync(kFull, c1);
                    todo[2] = __reduce_or_sync(kFull, c2); todo[3] = __reduce_or_sync(kFull, c3);
                    if (__popc(todo[0]) + __popc(todo[1]) + __popc(todo[2]) + __popc(todo[3]) > kMaxIncremental) fresh = true;
                }
            }
            wb_prev = wb;
            if (fresh) {
                parent_presence(wb, rmax, um0);
                // every candidate starts childless (static answers); present rows overwrite theirs below
                bv0[0] = sokv.x; bv0[1] = sokv.y; bv0[2] = sokv.z; bv0[3] = sokv.w;
                bb0[0] = sboth.x; bb0[1] = sboth.y; bb0[2] = sboth.z; bb0[3] = sboth.w;
                bu0[0] = soku.x; bu0[1] = soku.y; bu0[2] = soku.z; bu0[3] = soku.w;
            }
            if (fresh || (todo[0] | todo[1] | todo[2] | todo[3])) {
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    unsigned rows;
                    if (fresh) {
                        E0[g] = 0.0;
                        rows = prune ? um0[g] : um0[g] & (candq[g] | candq2[g]);
                    } else {
                        rows = todo[g];
                    }
                    while (rows) {
                        const int bq = __ffs(rows) - 1;
                        rows &= rows - 1;
                        const unsigned q = g * 32 + bq;
                        if (!fresh) {
                            const bool here = __ballot_sync(kFull, __vcmpeq4(wb, q * 0x01010101u) != 0) != 0;
                            um0[g] = here ? um0[g] | (1u << bq) : um0[g] & ~(1u << bq);
                        }
                        const bool is_cv = ((candq[g] >> bq) & 1u) != 0, is_cu = ((candq2[g] >> bq) & 1u) != 0;
                        const unsigned iv = cidx_s[q], iu = cidx2_s[q];
                        double Eq = 0.0;
                        unsigned r;
                        if ((um0[g] >> bq) & 1u) {
                            r = (prune || is_cv || is_cu)
                                    ? dual_row<SPL>(wb, q, rmax, wstride, vaf_s, SP, lane, xv0, xv1, xu0, xu1, net.margin, valid0,
                                                    valid1, is_cv, is_cu, prune, Eq)
                                    : 0u;
                        } else {
                            r = (is_cv && get_bit(sokv, iv) ? 1u : 0u) | (is_cv && get_bit(sboth, iv) ? 2u : 0u) |
                                (is_cu && get_bit(soku, iu) ? 4u : 0u);
                        }
                        if (lane == bq) E0[g] = Eq;
                        if (is_cv) { put_bit(bv0, iv, (r & 1u) != 0); put_bit(bb0, iv, (r & 2u) != 0); }
                        if (is_cu) put_bit(bu0, iu, (r & 4u) != 0);
                    }
                }
            }
            // the one row that sees x: its parent p in this item
            unsigned bv[4] = {bv0[0], bv0[1], bv0[2], bv0[3]}, bb[4] = {bb0[0], bb0[1], bb0[2], bb0[3]};
            unsigned bu[4] = {bu0[0], bu0[1], bu0[2], bu0[3]};
            double E[4] = {E0[0], E0[1], E0[2], E0[3]};
            if (has_x) {
                const unsigned p = (__shfl_sync(kFull, w, wl3) >> sh3) & 0xffu;
                const unsigned gp = p >> 5, bp = p & 31;
                const unsigned cq = gp == 0 ? candq[0] : gp == 1 ? candq[1] : gp == 2 ? candq[2] : candq[3];
                const unsigned cq2 = gp == 0 ? candq2[0] : gp == 1 ? candq2[1] : gp == 2 ? candq2[2] : candq2[3];
                const bool is_cv = ((cq >> bp) & 1u) != 0, is_cu = ((cq2 >> bp) & 1u) != 0;
                if (prune || is_cv || is_cu) {
                    double Eq = 0.0;
                    const unsigned r = dual_row<SPL>(w, p, rmax, wstride, vaf_s, SP, lane, xv0, xv1, xu0, xu1, net.margin, valid0,
                                                     valid1, is_cv, is_cu, prune, Eq);
                    if (is_cv) { const unsigned iv = cidx_s[p]; put_bit(bv, iv, (r & 1u) != 0); put_bit(bb, iv, (r & 2u) != 0); }
                    if (is_cu) put_bit(bu, cidx2_s[p], (r & 4u) != 0);
                    if (lane == (int)bp) {
                        E[0] = gp == 0 ? Eq : E[0]; E[1] = gp == 1 ? Eq : E[1];
                        E[2] = gp == 2 ? Eq : E[2]; E[3] = gp == 3 ? Eq : E[3];
                    }
                }
            }
            unsigned flag = 0;
            double tb = 0.0;
            if (prune) {
                tb = butterfly_sum(__dadd_rn(__dadd_rn(E[0], E[1]), __dadd_rn(E[2], E[3])));
                flag = (full && tb > tmax) ? 1u : 0u;
            }
            if (lane == 0) {
                bound_out[it] = tb;                      // kept with the children: a later, tighter threshold re-tests it for free
                mask_out[it] = make_uint4(bv[0], bv[1], bv[2], bv[3]);
                count_out[it] = __popc(bv[0]) + __popc(bv[1]) + __popc(bv[2]) + __popc(bv[3]);
                both_out[it] = make_uint4(bb[0], bb[1], bb[2], bb[3]);
                masku_out[it] = make_uint4(bu[0], bu[1], bu[2], bu[3]);
                flag_out[it] = flag;
            }
        }
    }
}

// emit for the level above the last one: besides the child items it writes, for every child, the pass
// mask of the last node u (the parent's "u without v" mask with the bit of v's parent replaced by the
// "both" decision) and its count.  Children of a BOUNDED parent get the count and the kBounded flag only:
// nothing reads their bytes or masks, so they are not written (their slots in the level buffer stay
// reserved, which keeps item indices and sequence numbers where they belong).
__global__ void __launch_bounds__(kThreads)
emit2_kernel(DevNet net, const unsigned *__restrict__ items, long long n_items, const uint4 *__restrict__ mask,
             const unsigned *__restrict__ off, const uint4 *__restrict__ both, const uint4 *__restrict__ masku,
             const unsigned *__restrict__ flag, const double *__restrict__ bound, unsigned *__restrict__ out,
             uint4 *__restrict__ lmask_out, unsigned *__restrict__ lcnt_out, double *__restrict__ lbound_out, ScanOut *so,
             const TopMeta *meta, int num_save) {
    __shared__ uint8_t cand_s[kMaxN], cidx2_s[kMaxN];
    __shared__ unsigned candq2_s[4];
    const int level = net.L - 2;
    const int k2 = net.ncand[level + 1];
    if (threadIdx.x < 4) candq2_s[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < kMaxN; i += blockDim.x) cand_s[i] = net.cand[(size_t)level * kMaxN + i];
    __syncthreads();
    if (threadIdx.x < k2) {
        const unsigned q = net.cand[(size_t)(level + 1) * kMaxN + threadIdx.x];
        cidx2_s[q] = (uint8_t)threadIdx.x;
        atomicOr(&candq2_s[q >> 5], 1u << (q & 31));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wstride = net.stride >> 2;
    const int wl = item_byte(level, wstride) >> 2, sh = 8 * (item_byte(level, wstride) & 3);
    const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
    unsigned long long my_trees = 0, my_live = 0;       // per lane, reduced at the end
    // The flag was set by check2_kernel against the threshold of ITS launch -- which may have run ahead, beside the
    // selection of the previous fill (run_search, prefetch); the bound came with it, so it is compared once more with
    // what the list admits now: a base the newer threshold dismisses costs count words only and no last-level launch.
    bool full;
    double tau;
    const double tmax = admit_bound(meta, num_save, full, tau);
    const bool prune = net.prune != 0 && full;
    // the next base's words are requested while this one is written (the loop is bound by their latency)
    uint4 m4n = make_uint4(0, 0, 0, 0), mbn = m4n, mun = m4n;
    unsigned fln = 0, offn = 0;
    double tbn = 0.0;
    if (warp0 < n_items) { m4n = mask[warp0]; mbn = both[warp0]; mun = masku[warp0]; fln = flag[warp0]; offn = off[warp0]; tbn = bound[warp0]; }
    for (long long it = warp0; it < n_items; it += nwarps) {
        const uint4 m4 = m4n, mb = mbn, mu = mun;
        const double tb = tbn;
        const unsigned fl = (fln || (prune && tb > tmax)) ? kBounded : 0u;
        long long dst = (long long)offn;
        if (it + nwarps < n_items) {
            const long long nx = it + nwarps;
            m4n = mask[nx]; mbn = both[nx]; mun = masku[nx]; fln = flag[nx]; offn = off[nx]; tbn = bound[nx];
        }
        if ((m4.x | m4.y | m4.z | m4.w) == 0) continue;
        const unsigned mw[4] = {m4.x, m4.y, m4.z, m4.w};
        if (fl) {
            // bounded base: nobody will read its children, only their tree counts are needed (for the number
            // of valid trees and the sequence numbers of later candidates).  Lane j owns candidate g*32+j: where its
            // bit is set it writes the count of that child at the child's rank among the set bits.  No item bytes.
            const unsigned n_u = __popc(mu.x) + __popc(mu.y) + __popc(mu.z) + __popc(mu.w);
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const unsigned b = mw[g];
                if ((b >> lane) & 1u) {
                    const unsigned i = g * 32 + lane, q = cand_s[i];
                    unsigned cnt = n_u;
                    if ((candq2_s[q >> 5] >> (q & 31)) & 1u)
                        cnt = cnt - (get_bit(mu, cidx2_s[q]) ? 1u : 0u) + (get_bit(mb, i) ? 1u : 0u);
                    lcnt_out[dst + __popc(b & ((1u << lane) - 1u))] = cnt | fl;
                    my_trees += cnt;
                }
                dst += __popc(b);
            }
            continue;
        }
        const unsigned w = lane < wstride ? __ldg(items + it * wstride + lane) : 0xffffffffu;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            unsigned b = mw[g];
            while (b) {
                const int i = g * 32 + __ffs(b) - 1;
                b &= b - 1;
                const unsigned q = cand_s[i];
                unsigned cw = w;
                if (lane == wl) cw = (w & ~(0xffu << sh)) | (q << sh);
                if (lane < wstride) out[dst * wstride + lane] = cw;
                if (lane == 0) {
                    unsigned lm[4] = {mu.x, mu.y, mu.z, mu.w};
                    if ((candq2_s[q >> 5] >> (q & 31)) & 1u) put_bit(lm, cidx2_s[q], get_bit(mb, (unsigned)i));
                    const unsigned cnt = __popc(lm[0]) + __popc(lm[1]) + __popc(lm[2]) + __popc(lm[3]);
                    lmask_out[dst] = make_uint4(lm[0], lm[1], lm[2], lm[3]);
                    lcnt_out[dst] = cnt;
                    lbound_out[dst] = tb;
                    my_trees += cnt;
                    my_live += 1;
                }
                dst++;
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        my_trees += __shfl_xor_sync(kFull, my_trees, d);
        my_live += __shfl_xor_sync(kFull, my_live, d);
    }
    if (lane == 0) {
        if (my_trees) atomicAdd(&so->trees, my_trees);
        if (my_live) atomicAdd(&so->live, my_live);
    }
}

// Last level: the pass mask + count of every item were written by emit2_kernel (for the scan that numbers
// the valid trees); this kernel totals the trees of the items that are not bounded and appends those that
// beat the threshold as 16-byte candidates; seq_fixup_kernel takes the square roots and gives them their canonical
// sequence numbers once the scan has run.
template <int SPL>
__global__ void __launch_bounds__(kThreads, 2)
leaf_kernel(DevNet net, const unsigned *__restrict__ items, const ScanOut *__restrict__ fit, long long item_base,
            const unsigned *__restrict__ lcnt, const double *__restrict__ lbound, TopMeta *meta,
            int num_save, uint4 *__restrict__ cands, long long cand_cap, int run_len_arg, const unsigned *__restrict__ ref_item) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *vaf_s = reinterpret_cast<double *>(smem_raw);
    uint8_t *cand_s = reinterpret_cast<uint8_t *>(vaf_s + net.n * net.SP);
    const bool adjacent_only = run_len_arg < 0;            // LICHEE_NO_PREFILTER=1: the round-1 rules (tests: nothing may move)
    const int run_len = run_len_arg < 0 ? -run_len_arg : run_len_arg;
    __shared__ double fstat_s[kMaxN];          // by row: E' of a candidate whose only child is the last node
    __shared__ uint8_t cidx_s[kMaxN];          // by row: candidate index (valid where candq has the bit)
    __shared__ unsigned candq_s[4], okq_s[4];  // rows that are candidates / that pass with no other child
    const int level = net.L - 1;
    const int k = net.ncand[level];
    if (threadIdx.x < 4) { candq_s[threadIdx.x] = 0; okq_s[threadIdx.x] = 0; }
    stage_network(net, level, vaf_s, cand_s);
    if (threadIdx.x < k) {
        const unsigned q = cand_s[threadIdx.x];
        cidx_s[q] = (uint8_t)threadIdx.x;
        fstat_s[q] = net.leaf_static[threadIdx.x];
        atomicOr(&candq_s[q >> 5], 1u << (q & 31));
        const uint4 so = net.static_ok[level];
        const unsigned sw = (threadIdx.x >> 5) == 0 ? so.x : (threadIdx.x >> 5) == 1 ? so.y : (threadIdx.x >> 5) == 2 ? so.z : so.w;
        if ((sw >> (threadIdx.x & 31)) & 1u) atomicOr(&okq_s[q >> 5], 1u << (q & 31));
    }
    __syncthreads();
    const unsigned candq[4] = {candq_s[0], candq_s[1], candq_s[2], candq_s[3]};
    const unsigned okq[4] = {okq_s[0], okq_s[1], okq_s[2], okq_s[3]};

    const int lane = threadIdx.x & 31;
    const int SP = net.SP;
    const bool valid0 = lane < net.S;
    const bool valid1 = SPL == 2 && lane + 32 < net.S;
    const int wstride = net.stride >> 2;
    const int rmax = level > 0 ? (level - 1) / wstride : -1;
    double xv0, xv1;
    load_row<SPL>(vaf_s, level + 1, SP, lane, xv0, xv1);
    bool full;
    double tau;
    const double tmax = admit_bound(meta, num_save, full, tau);
    unsigned long long my_valid = 0, my_eval = 0;
    unsigned my_bounded = 0;
    // candidate slots are reserved kCandBatch at a time per warp (one atomic on the shared counter per batch,
    // not per tree: in a good region millions of trees beat a stale threshold within one launch); slots a
    // warp does not use are written as empty entries (pad = 1), which every consumer skips
    unsigned long long res_next = 0, res_end = 0;
    // A candidate leaves this kernel as 16 bytes: {total (the sum of squared excesses, 8 bytes), item, row of the last
    // node's parent | its candidate index << 8}; item = ~0 marks an empty slot.  seq_fixup_kernel takes the square
    // root, numbers the tree and writes the 32-byte record and the selection keys (in a good region a third of this
    // kernel's instructions went into building and storing the full records of millions of trees).
    auto fill_empty = [&](unsigned long long from, unsigned long long to) {
        const uint4 e = make_uint4(0u, 0x7ff00000u, 0xffffffffu, 0u);
        for (unsigned long long sl = from + lane; sl < to; sl += 32)
            if ((long long)sl < cand_cap) cands[sl] = e;
    };
    // the chunk: the longest prefix of the pending items whose unbounded trees fit the candidate buffer
    // (scanned on the device just before this launch; the host reads the same number afterwards)
    const long long n_items = (long long)(fit->fit >> 32);

    // A warp walks a contiguous run of items.  Neighbours in the canonical order differ in very few
    // parent bytes, and most of the time only in the LAST assigned one: the children of one item of the
    // level above are adjacent and differ in where its node v (tree position L-2) hangs.  The state is
    // therefore kept for the BASE item (the item with v taken out): per parent row q the excess E[q]
    // without the last node and F[q] with it (passing candidates), presence and pass masks.  They carry
    // over in registers and only the rows named by a changed base byte are recomputed.  On top of the base,
    // ONE row per item is evaluated with v in place: the parent p that v hangs under (v is the highest
    // position, hence the last child in p's sum -- same additions as a full evaluation of the item).
    // Lane l keeps rows l, l+32, l+64, l+96.  passq = passing candidates by row.
    double E[4] = {0.0, 0.0, 0.0, 0.0}, F[4] = {0.0, 0.0, 0.0, 0.0};
    unsigned passq[4] = {0, 0, 0, 0}, um[4] = {0, 0, 0, 0};
    unsigned wb_prev = 0;
    // Lower bound: VAFs are non-negative, so hanging further nodes under a parent never lowers its sum and
    // (every operation of the fixed-shape reduction being monotone) never lowers a total.  The total of the
    // BASE item's E rows therefore bounds the total of every tree below it; when it already exceeds what
    // the full ranked list admits, the whole sibling group is only counted (sum-rule pass masks), not
    // scored.  F rows are then filled in lazily: fstale marks rows whose F was skipped.
    unsigned fstale[4] = {0, 0, 0, 0};
    bool hopeless = false;
    const bool prune = net.prune != 0 && full;
    const bool prune_pre = prune && !adjacent_only;
    const bool has_v = level >= 1;                          // a tree position L-2 exists
    const int wl2 = has_v ? item_byte(level - 1, wstride) >> 2 : 0, sh2 = has_v ? 8 * (item_byte(level - 1, wstride) & 3) : 0;
    // The pass mask and the number of valid trees of every item were written by emit2_kernel together with
    // the item (lcnt: count | kBounded), so items whose base was already out of reach when the
    // level above was checked are skipped without being read.  State carries over between consecutively
    // EVALUATED items only.
    // runs are handed out from a counter, so a warp that drew cheap runs simply draws more of them
    long long last_it = -2;
    bool have_state = false;
    // Reference state of the block.  A warp's first evaluated item used to cost a full evaluation of its base (every
    // present parent row: ~6000 instructions, 40 % of a launch inside a good region).  The bases of one fill share all
    // but their last few bytes, so the block evaluates ONE base of the fill together -- ref_item: the first item of
    // the chunk of the level above that this fill came from; every warp takes a 16th of its rows -- and every warp
    // starts from that state: its first item then redoes only the rows named by the bytes that differ (or falls
    // back to the full evaluation when those are many).  Same arithmetic per row, so nothing moves in the results.
    if (ref_item != nullptr && !adjacent_only && has_v) {
        __shared__ double eref_s[kMaxN], fref_s[kMaxN];
        __shared__ unsigned passref_s[4];
        const int wid = threadIdx.x >> 5;
        unsigned wr = lane < wstride ? __ldg(ref_item + lane) : 0xffffffffu;
        if (lane == wl2) wr |= 0xffu << sh2;
        parent_presence(wr, rmax, um);
        if (threadIdx.x < 4) passref_s[threadIdx.x] = candq[threadIdx.x] & ~(threadIdx.x == 0 ? um[0] : threadIdx.x == 1 ? um[1] : threadIdx.x == 2 ? um[2] : um[3]) & okq[threadIdx.x];
        __syncthreads();
        int rank = 0;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            unsigned rows = um[g];
            while (rows) {
                const int bq = __ffs(rows) - 1;
                rows &= rows - 1;
                if ((rank++ & (kWarpsPerBlock - 1)) != wid) continue;
                double Eq, Fq;
                const bool pass = leaf_row<SPL>(wr, g * 32 + bq, rmax, wstride, vaf_s, SP, lane, xv0, xv1, net.margin, valid0, valid1,
                                                ((candq[g] >> bq) & 1u) != 0, true, Eq, Fq);
                if (lane == 0) {
                    eref_s[g * 32 + bq] = Eq;
                    fref_s[g * 32 + bq] = Fq;
                    if (pass) atomicOr(&passref_s[g], 1u << bq);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const bool present = ((um[g] >> lane) & 1u) != 0;
            const bool absent_ok = ((candq[g] & ~um[g] & okq[g]) >> lane) & 1u;
            passq[g] = passref_s[g];
            E[g] = present ? eref_s[g * 32 + lane] : 0.0;
            F[g] = present ? fref_s[g * 32 + lane] : absent_ok ? fstat_s[g * 32 + lane] : 0.0;
        }
        wb_prev = wr;
        have_state = true;
        if (prune) hopeless = butterfly_sum(__dadd_rn(__dadd_rn(E[0], E[1]), __dadd_rn(E[2], E[3]))) > tmax;
    }
    for (;;) {
    long long run0 = 0;
    if (lane == 0) run0 = (long long)atomicAdd(&meta->next_run, 1ull) * run_len;
    run0 = __shfl_sync(kFull, run0, 0);
    if (run0 >= n_items) break;
    const long long run1 = run0 + run_len < n_items ? run0 + run_len : n_items;
    for (long long blk = run0; blk < run1; blk += 32) {
    const long long idx = blk + lane;
    const unsigned cw = idx < run1 ? __ldg(lcnt + idx) : kBounded;
    if (idx < run1) my_valid += cw & kCountMask;
    // not flagged when the level above was checked -- but the threshold has tightened since (a burst earlier in
    // this fill): the base's bound travelled with the item, so the re-test costs one load instead of an evaluation
    bool is_live = (cw & kBounded) == 0;
    if (is_live && prune_pre && __ldg(lbound + idx) > tmax) is_live = false;
    unsigned live = __ballot_sync(kFull, is_live);
    {
        const int nblk = run1 - blk < 32 ? (int)(run1 - blk) : 32;
        my_bounded += nblk - __popc(live);
    }
    unsigned w_next = 0;
    long long next_it = -1;
    while (live) {
        const long long it = blk + __ffs(live) - 1;
        live &= live - 1;
        unsigned w;
        if (next_it == it) w = w_next;
        else w = lane < wstride ? __ldg(items + it * wstride + lane) : 0xffffffffu;
        // the next item's bytes are requested now so that their latency hides behind this item's work
        if (live) {
            next_it = blk + __ffs(live) - 1;
            w_next = lane < wstride ? __ldg(items + next_it * wstride + lane) : 0xffffffffu;
        }
        // base item: v unassigned
        const unsigned wb = (has_v && lane == wl2) ? (w | (0xffu << sh2)) : w;
        unsigned todo[4] = {0, 0, 0, 0};                     // base rows to (re)compute for this item
        // The state belongs to the base of the item evaluated LAST by this warp, wherever that item was: the rows to
        // redo come from the bytes in which the two bases differ, so bounded items in between (97 % of a level on
        // the headline network) do not force a full evaluation of the next live one.
        bool fresh = !have_state || (adjacent_only && it != last_it + 1);
        last_it = it;
        have_state = true;
        if (!fresh) {
            const unsigned diff = wb ^ wb_prev;
            if (__any_sync(kFull, diff != 0)) {
                // rows named by a base byte that changed, in the old or the new item
                unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                for (int r = 0; r <= rmax; r++) {
                    if ((diff >> (8 * r)) & 0xffu) {
                        const unsigned bo = (wb_prev >> (8 * r)) & 0xffu, bn = (wb >> (8 * r)) & 0xffu;
                        const unsigned bito = bo < 128u ? 1u << (bo & 31) : 0u, bitn = bn < 128u ? 1u << (bn & 31) : 0u;
                        c0 |= ((bo >> 5) == 0 ? bito : 0u) | ((bn >> 5) == 0 ? bitn : 0u);
                        c1 |= ((bo >> 5) == 1 ? bito : 0u) | ((bn >> 5) == 1 ? bitn : 0u);
                        c2 |= ((bo >> 5) == 2 ? bito : 0u) | ((bn >> 5) == 2 ? bitn : 0u);
                        c3 |= ((bo >> 5) == 3 ? bito : 0u) | ((bn >> 5) == 3 ? bitn : 0u);
                    }
                }
                todo[0] = __reduce_or_sync(kFull, c0); todo[1] = __reduce_or_sync(kFull, c1);
                todo[2] = __reduce_or_sync(kFull, c2); todo[3] = __reduce_or_sync(kFull, c3);
                if (__popc(todo[0]) + __popc(todo[1]) + __popc(todo[2]) + __popc(todo[3]) > kMaxIncremental) fresh = true;
            }
        }
        wb_prev = wb;
        if (fresh) {
            parent_presence(wb, rmax, um);
            fstale[0] = fstale[1] = fstale[2] = fstale[3] = 0;
        }
        if (fresh || (todo[0] | todo[1] | todo[2] | todo[3])) {
#pragma unroll
        for (int g = 0; g < 4; g++) {
            unsigned rows;
            if (fresh) {
                // candidates with no child in this item: everything about them is static
                const unsigned absent = candq[g] & ~um[g] & okq[g];
                passq[g] = absent;
                E[g] = 0.0;
                F[g] = ((absent >> lane) & 1u) ? fstat_s[g * 32 + lane] : 0.0;
                rows = um[g];
            } else {
                rows = todo[g];
            }
            while (rows) {
                const int bq = __ffs(rows) - 1;
                rows &= rows - 1;
                const unsigned q = g * 32 + bq;
                double Eq = 0.0, Fq = 0.0;
                bool pass = false;
                if (!fresh) {                                // presence of a changed row: any byte equal to q
                    const bool here = __ballot_sync(kFull, __vcmpeq4(wb, q * 0x01010101u) != 0) != 0;
                    um[g] = here ? um[g] | (1u << bq) : um[g] & ~(1u << bq);
                }
                fstale[g] &= ~(1u << bq);
                if ((um[g] >> bq) & 1u) {                   // parent present in the base item
                    pass = leaf_row<SPL>(wb, q, rmax, wstride, vaf_s, SP, lane, xv0, xv1, net.margin, valid0, valid1,
                                         ((candq[g] >> bq) & 1u) != 0, !prune, Eq, Fq);
                    if (prune && pass) fstale[g] |= 1u << bq;
                } else if (((candq[g] & okq[g]) >> bq) & 1u) {          // childless candidate: static
                    Fq = fstat_s[q];
                    pass = true;
                }
                if (lane == bq) { E[g] = Eq; F[g] = Fq; }
                passq[g] = pass ? passq[g] | (1u << bq) : passq[g] & ~(1u << bq);
            }
        }
        }
        if (prune && (fresh || (todo[0] | todo[1] | todo[2] | todo[3]))) {
            // new base: its total bounds every tree of the sibling group from below
            const double tb = butterfly_sum(__dadd_rn(__dadd_rn(E[0], E[1]), __dadd_rn(E[2], E[3])));
            hopeless = tb > tmax;
        }
        if (hopeless) {            // bounded by the current threshold (tighter than the one the level above saw)
            my_bounded++;
            continue;
        }
        if (fstale[0] | fstale[1] | fstale[2] | fstale[3]) {
            // a group that may reach the list: fill in the F rows that were skipped
#pragma unroll
            for (int g = 0; g < 4; g++) {
                unsigned rows = fstale[g];
                fstale[g] = 0;
                while (rows) {
                    const int bq = __ffs(rows) - 1;
                    rows &= rows - 1;
                    double Eq, Fq;
                    leaf_row<SPL>(wb, g * 32 + bq, rmax, wstride, vaf_s, SP, lane, xv0, xv1, net.margin, valid0, valid1, true, true, Eq, Fq);
                    if (lane == bq) F[g] = Fq;
                }
            }
        }
        // the one row that sees v: its parent p in this item
        unsigned pq0 = passq[0], pq1 = passq[1], pq2 = passq[2], pq3 = passq[3];
        double E0 = E[0], E1 = E[1], E2 = E[2], E3 = E[3], F0 = F[0], F1 = F[1], F2 = F[2], F3 = F[3];
        if (has_v) {
            const unsigned p = (__shfl_sync(kFull, w, wl2) >> sh2) & 0xffu;
            const unsigned gp = p >> 5, bp = p & 31;
            const unsigned cq = gp == 0 ? candq[0] : gp == 1 ? candq[1] : gp == 2 ? candq[2] : candq[3];
            const bool is_cand = ((cq >> bp) & 1u) != 0;
            double Ep, Fp;
            const bool pass = leaf_row<SPL>(w, p, rmax, wstride, vaf_s, SP, lane, xv0, xv1, net.margin, valid0, valid1, is_cand, true, Ep, Fp);
            const bool mine = lane == (int)bp;
            const unsigned bitq = 1u << bp;
            if (gp == 0) { if (mine) { E0 = Ep; F0 = Fp; } pq0 = pass ? pq0 | bitq : pq0 & ~bitq; }
            else if (gp == 1) { if (mine) { E1 = Ep; F1 = Fp; } pq1 = pass ? pq1 | bitq : pq1 & ~bitq; }
            else if (gp == 2) { if (mine) { E2 = Ep; F2 = Fp; } pq2 = pass ? pq2 | bitq : pq2 & ~bitq; }
            else { if (mine) { E3 = Ep; F3 = Fp; } pq3 = pass ? pq3 | bitq : pq3 & ~bitq; }
        }
        // Base reduction over the 128 slots, remembering what each lane received at every stage: a
        // tree differs from the base in one slot of one lane, so its total is that lane's partial
        // pushed through the same five additions -- bit-identical to a full butterfly.
        my_eval += __popc(pq0) + __popc(pq1) + __popc(pq2) + __popc(pq3);
        const double P = __dadd_rn(__dadd_rn(E0, E1), __dadd_rn(E2, E3));
        double rcv[5];
        {
            double v = P;
#pragma unroll
            for (int st = 0; st < 5; st++) {
                rcv[st] = __shfl_xor_sync(kFull, v, 16 >> st);
                v = __dadd_rn(v, rcv[st]);
            }
        }
#pragma unroll
        for (int g = 0; g < 4; g++) {
            // this lane's partial with slot g swapped for F
            const double A = g == 0 ? __dadd_rn(__dadd_rn(F0, E1), __dadd_rn(E2, E3))
                           : g == 1 ? __dadd_rn(__dadd_rn(E0, F1), __dadd_rn(E2, E3))
                           : g == 2 ? __dadd_rn(__dadd_rn(E0, E1), __dadd_rn(F2, E3))
                                    : __dadd_rn(__dadd_rn(E0, E1), __dadd_rn(E2, F3));
            const unsigned pqg = g == 0 ? pq0 : g == 1 ? pq1 : g == 2 ? pq2 : pq3;
            double x = A;
#pragma unroll
            for (int st = 0; st < 5; st++) x = __dadd_rn(x, rcv[st]);
            // lane l now holds the total of the tree that hangs the last node under row g*32+l: every
            // tree of the item is evaluated at once, and only the few that may beat the threshold
            // go through the square root and the append (one slot reservation per group, every lane
            // writes its own tree)
            const bool cand = ((pqg >> lane) & 1u) && x <= tmax;
            const unsigned b = __ballot_sync(kFull, cand);
            if (b) {
                const unsigned c_n = __popc(b);
                if (res_end - res_next < c_n) {
                    fill_empty(res_next, res_end);
                    unsigned long long r0 = 0;
                    if (lane == 0) r0 = atomicAdd(&meta->n_cand, (unsigned long long)kCandBatch);
                    res_next = __shfl_sync(kFull, r0, 0);
                    res_end = res_next + kCandBatch;
                }
                if (cand) {
                    // (x <= tmax already says that the rounded square root is below the threshold: admit_bound)
                    const unsigned long long slot = res_next + __popc(b & ((1u << lane) - 1u));
                    if ((long long)slot < cand_cap) {
                        const unsigned q = g * 32 + lane;
                        const long long xb = __double_as_longlong(x);
                        cands[slot] = make_uint4((unsigned)xb, (unsigned)(xb >> 32), (unsigned)(item_base + it), q | ((unsigned)cidx_s[q] << 8));
                    }
                }
                res_next += c_n;
            }
        }
    }
    }
    }
    fill_empty(res_next, res_end);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) my_valid += __shfl_xor_sync(kFull, my_valid, d);
    if (lane == 0 && my_valid) atomicAdd(&meta->n_valid, my_valid);
    if (lane == 0 && my_bounded) atomicAdd(&meta->n_bounded, (unsigned long long)my_bounded);
    if (lane == 0 && my_eval) atomicAdd(&meta->n_eval, my_eval);
}

// A survivor's record from its 16-byte candidate and its keys (score bits, sequence number).
__device__ __forceinline__ Cand cand_record(const uint4 &r, unsigned long long k0, unsigned long long k1, int part) {
    Cand c;
    c.score = __longlong_as_double((long long)k0); c.seq = (long long)k1; c.part = part; c.origin = r.z; c.pstar = r.w & 0xffu; c.pad = 0;
    return c;
}

// Turns the 16-byte candidates of leaf_kernel into ranked-list records: score = sqrt_rn(total), the canonical sequence
// number (rank among the valid trees) now that the scan has numbered the items; retires the ones beyond
// Parameters.MAX_NUM_TREES.  Also writes the selection keys (score bits, seq; ~0 for an empty slot) as two plain
// arrays -- the radix passes then read 8 or 16 bytes per candidate instead of the 32-byte record -- and records the
// smallest and largest score bits of the launch (where the radix passes start).
__global__ void __launch_bounds__(256)
seq_fixup_kernel(const uint4 *__restrict__ c16, Cand *__restrict__ cands, long long m, long long item_base,
                 const uint4 *__restrict__ mask, const unsigned *__restrict__ off, long long seq_base, long long max_trees,
                 unsigned long long *__restrict__ key0, unsigned long long *__restrict__ key1, TopMeta *meta, int num_save, int part,
                 int write_records) {
    __shared__ unsigned long long mn_s[8], mx_s[8];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long kb = ~0ull, seqk = ~0ull;
    bool live = false;
    Cand c;
    c.score = INFINITY; c.seq = 0x7fffffffffffffffll; c.part = 0x7fffffff; c.origin = 0xffffffffu; c.pstar = 0; c.pad = 1;
    if (t < m) {
        const uint4 r = c16[t];
        if (r.z != 0xffffffffu) {
            const double x = __longlong_as_double((long long)(((unsigned long long)r.y << 32) | r.x));
            const double score = __dsqrt_rn(x);
            const long long it = (long long)r.z - item_base;
            const unsigned i = r.w >> 8;
            const uint4 m4 = mask[it];
            const unsigned mw[4] = {m4.x, m4.y, m4.z, m4.w};
            unsigned rank = 0;
            for (unsigned g = 0; g < (i >> 5); g++) rank += __popc(mw[g]);
            rank += __popc(mw[i >> 5] & ((1u << (i & 31)) - 1u));
            const long long seq = seq_base + (long long)off[it] + rank;
            // (a full list admits only scores below its threshold: leaf_kernel's bound on the total says the same)
            const bool admitted = meta->count < (unsigned)num_save || score < meta->tau;
            if (seq < max_trees && admitted) {
                live = true;
                kb = (unsigned long long)__double_as_longlong(score);
                seqk = (unsigned long long)seq;
                c.score = score; c.seq = seq; c.part = part; c.origin = r.z; c.pstar = r.w & 0xffu; c.pad = 0;
            }
        }
        key0[t] = kb; key1[t] = seqk;
        if (write_records) cands[t] = c;       // (a long list: only the survivors' records are built, by radix12_compact)
    }
    // key range of the launch: one pair of atomics per block
    unsigned long long mn = live ? kb : ~0ull, mx = live ? kb : 0ull;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long a = __shfl_xor_sync(kFull, mn, d), b = __shfl_xor_sync(kFull, mx, d);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0) { mn_s[threadIdx.x >> 5] = mn; mx_s[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) { mn = mn_s[w] < mn ? mn_s[w] : mn; mx = mx_s[w] > mx ? mx_s[w] : mx; }
        if (mn <= mx) { atomicMin(&meta->kmin, mn); atomicMax(&meta->kmax, mx); }
    }
}

// ------------------------------------------------------------------------------------------
// kernel 5: block-level top-s selection.  One block keeps the ranked list (<= 1024 trees) in
// shared memory.  Up to 1024 new candidates (the usual case, after the radix pre-selection) are
// sorted and rank-merged with binary searches; longer lists are streamed through in tiles with a
// bitonic sort of 2048 entries.  Keys are (score, part, seq): ties in score go to the earlier tree of
// the canonical order.  The survivors' parent bytes are gathered into the other half of the
// double-buffered tree store and the new threshold is published.
// ------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void bitonic_sort(Cand *e) {
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < N; t += kSelThreads) {
                const int x = t ^ j;
                if (x > t) {
                    const bool up = (t & k) == 0;
                    Cand a = e[t], b = e[x];
                    if (cand_less(b, a) == up) {
                        e[t] = b;
                        e[x] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// Tournament pre-selection for long candidate lists: every block streams kSelPerBlock candidates
// through the same 2048-entry bitonic sort and writes its best 1024 (a superset of the global
// best num_save <= 1024 restricted to its slice) to `out`.  Repeated by the host until the list is
// short enough for the final single-block select_kernel.
constexpr int kSelPerBlock = 8 * kSelTile;

__global__ void __launch_bounds__(kSelThreads)
select_reduce_kernel(const Cand *__restrict__ in, long long m, Cand *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char sel_raw[];
    Cand *e = reinterpret_cast<Cand *>(sel_raw);
    Cand sentinel;
    sentinel.score = INFINITY;
    sentinel.seq = 0x7fffffffffffffffll;
    sentinel.part = 0x7fffffff;
    sentinel.origin = 0xffffffffu;
    sentinel.pstar = 0;
    sentinel.pad = 1;
    const long long lo = (long long)blockIdx.x * kSelPerBlock;
    long long hi = lo + kSelPerBlock;
    if (hi > m) hi = m;
    for (int t = threadIdx.x; t < kSelTile; t += kSelThreads) e[t] = sentinel;
    __syncthreads();
    for (long long base = lo; base < hi; base += kSelTile) {
        for (int t = threadIdx.x; t < kSelTile; t += kSelThreads)
            e[kSelTile + t] = base + t < hi ? in[base + t] : sentinel;
        __syncthreads();
        bitonic_sort<2 * kSelTile>(e);
    }
    for (int t = threadIdx.x; t < kSelTile; t += kSelThreads) out[(long long)blockIdx.x * kSelTile + t] = e[t];
}

__global__ void __launch_bounds__(kSelThreads)
select_kernel(TopMeta *meta, int num_save, const Cand *__restrict__ cands, long long cand_cap,
              Cand *top_cur, Cand *top_next, const unsigned *__restrict__ tree_cur,
              unsigned *__restrict__ tree_next, const unsigned *__restrict__ leaf_items, int wstride,
              int last_level, int external_records, const unsigned char *__restrict__ records,
              long long record_bytes) {
    extern __shared__ __align__(16) unsigned char sel_raw[];
    Cand *e = reinterpret_cast<Cand *>(sel_raw);
    const unsigned count = meta->count;
    unsigned long long m = meta->n_cand;
    if ((long long)m > cand_cap) m = (unsigned long long)cand_cap;
    Cand sentinel;
    sentinel.score = INFINITY;
    sentinel.seq = 0x7fffffffffffffffll;
    sentinel.part = 0x7fffffff;
    sentinel.origin = 0xffffffffu;
    sentinel.pstar = 0;
    sentinel.pad = 1;   // pad == 1 marks a sentinel
    for (int t = threadIdx.x; t < kSelTile; t += kSelThreads) {
        if ((unsigned)t < count) {
            Cand c = top_cur[t];
            c.origin = 0x80000000u | (unsigned)t;
            e[t] = c;
        } else {
            e[t] = sentinel;
        }
    }
    __shared__ unsigned real_n;
    if (threadIdx.x == 0) real_n = 0;
    __syncthreads();
    Cand *res = e;
    if (m <= (unsigned long long)kSelTile) {
        // Common case, a handful of new candidates: rank-merge.  Keys (score, part, seq) are unique,
        // so the final position of an entry is the number of entries smaller than it.
        Cand *fresh = e + kSelTile, *out = e + 2 * kSelTile;
        const int mm = (int)m;
        const int t = threadIdx.x;
        fresh[t] = t < mm ? cands[t] : sentinel;
        out[t] = sentinel;
        __syncthreads();
        // sort the fresh candidates, then every entry's final position is its own index plus the
        // number of entries of the OTHER list that are smaller (binary search; keys are unique)
        bitonic_sort<kSelTile>(fresh);
        const Cand mine_old = e[t], mine_new = fresh[t];
        const bool has_old = (unsigned)t < count, has_new = mine_new.pad == 0;
        if (has_new) {
            int lo = 0, hi = (int)count;            // old entries smaller than my new entry
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (cand_less(e[mid], mine_new)) lo = mid + 1; else hi = mid;
            }
            const unsigned pos = (unsigned)t + (unsigned)lo;
            if (pos < (unsigned)kSelTile) out[pos] = mine_new;
            atomicAdd(&real_n, 1u);
        }
        if (has_old) {
            int lo = 0, hi = kSelTile;              // fresh entries smaller than my old entry (sentinels sort last)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (cand_less(fresh[mid], mine_old)) lo = mid + 1; else hi = mid;
            }
            const unsigned pos = (unsigned)t + (unsigned)lo;
            if (pos < (unsigned)kSelTile) out[pos] = mine_old;
            atomicAdd(&real_n, 1u);
        }
        __syncthreads();
        res = out;
    } else if (external_records == 2) {
        // Merge of ranked lists (one per GPU): every list arrives SORTED, `num_save` entries each.  The held tile is
        // ascending; a list loaded in reverse behind it makes the 2048 entries bitonic, so the eleven compare-exchange
        // steps of a bitonic MERGE order them -- not the 66 steps of a full sort per list.
        for (unsigned long long base = 0; base < m; base += (unsigned)num_save) {
            for (int t = threadIdx.x; t < kSelTile; t += kSelThreads)
                e[2 * kSelTile - 1 - t] = (t < num_save && base + t < m) ? cands[base + t] : sentinel;
            __syncthreads();
            for (int j = kSelTile; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < 2 * kSelTile; t += kSelThreads) {
                    const int x = t ^ j;
                    if (x > t) {
                        const Cand a = e[t], b = e[x];
                        if (cand_less(b, a)) { e[t] = b; e[x] = a; }
                    }
                }
                __syncthreads();
            }
        }
        unsigned mine = 0;
        for (int t = threadIdx.x; t < kSelTile; t += kSelThreads) mine += e[t].pad == 0;
        if (mine) atomicAdd(&real_n, mine);
        __syncthreads();
    } else {
        for (unsigned long long base = 0; base < m; base += kSelTile) {
            for (int t = threadIdx.x; t < kSelTile; t += kSelThreads)
                e[kSelTile + t] = base + t < m ? cands[base + t] : sentinel;
            __syncthreads();
            bitonic_sort<2 * kSelTile>(e);
        }
        // candidates may contain sentinel padding written by the tournament: count the real entries
        unsigned mine = 0;
        for (int t = threadIdx.x; t < kSelTile; t += kSelThreads) mine += e[t].pad == 0;
        if (mine) atomicAdd(&real_n, mine);
        __syncthreads();
    }
    const unsigned newc = real_n < (unsigned)num_save ? real_n : (unsigned)num_save;
    // gather parent bytes of the survivors
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int wl = item_byte(last_level, wstride) >> 2, sh = 8 * (item_byte(last_level, wstride) & 3);
    // (four trees per warp and iteration, their words requested together: the loop is bound by the latency of the loads)
    for (unsigned t0 = wid * 4; t0 < newc; t0 += kSelThreads / 32 * 4) {
        for (int j = lane; j < wstride; j += 32) {      // (reference-order records are 2 x stride bytes: up to 64 words)
            unsigned w[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const unsigned t = t0 + u;
                w[u] = 0;
                if (t >= newc) continue;
                const unsigned origin = res[t].origin, pstar = res[t].pstar;
                if (origin & 0x80000000u) {
                    w[u] = tree_cur[(size_t)(origin & 0x7fffffffu) * wstride + j];
                } else if (external_records) {
                    const unsigned *src = reinterpret_cast<const unsigned *>(records + (size_t)origin * record_bytes + 24);
                    w[u] = src[j];
                } else {
                    w[u] = leaf_items[(size_t)origin * wstride + j];
                    if (j == wl) w[u] = (w[u] & ~(0xffu << sh)) | (pstar << sh);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (t0 + u < newc) tree_next[(size_t)(t0 + u) * wstride + j] = w[u];
        }
    }
    for (int t = threadIdx.x; t < (int)newc; t += kSelThreads) top_next[t] = res[t];
    __syncthreads();
    if (threadIdx.x == 0) {
        meta->count = newc;
        meta->tau = newc >= (unsigned)num_save ? res[newc - 1].score : INFINITY;
        meta->n_cand = 0;
        meta->n_valid = 0;
        meta->next_run = 0; meta->n_bounded = 0; meta->n_eval = 0; meta->kmin = ~0ull; meta->kmax = 0ull;
    }
}

// ------------------------------------------------------------------------------------------
// Radix pre-selection for long candidate lists (bursts: a chunk that enters a better region of the
// search space beats a stale threshold with 10^5-10^6 trees).  Scores are non-negative doubles, so
// their bit patterns order like unsigned integers: sixteen 8-bit passes over (score bits, seq) find
// the num_save-th smallest key exactly, then the candidates at or below it are compacted for the
// rank-merge.  O(m), where
// the tournament of bitonic sorts is O(m log^2).
// ------------------------------------------------------------------------------------------
struct RadixState {
    unsigned long long prefix[2];   // decided high bits of the k-th smallest key: [0] score bits, [1] seq
    unsigned long long k;           // rank still to find inside the current prefix bucket
    unsigned hist[256];
    unsigned long long kept;        // candidates compacted
    unsigned blocks_done;           // last-block-done counter of the current pass
    unsigned ties_fit;              // set after the score passes when every tie at the threshold can be kept
    unsigned long long below;       // candidates with a score below the threshold
};

__global__ void radix_init(RadixState *st, unsigned long long k) {
    st->prefix[0] = st->prefix[1] = 0;
    st->k = k;
    st->kept = 0;
    st->blocks_done = 0;
    st->ties_fit = 0;
    st->below = 0;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) st->hist[i] = 0;
}

// One pass = one launch: every block histograms its share of the candidates for the current byte; the
// last block to finish picks the bucket that holds the k-th smallest key and clears the histogram.
// pass 0..7: bytes of the score (high to low); pass 8..15: bytes of seq among candidates whose score
// equals the selected one (skipped when all ties at the threshold fit in the rank-merge tile anyway).
// (part is constant inside one handle's candidate list.)
__global__ void __launch_bounds__(256)
radix_pass(const Cand *__restrict__ cands, long long m, RadixState *st, int pass, unsigned tile) {
    __shared__ unsigned h[256];
    __shared__ bool last;
    const int word = pass >> 3, shift = 56 - 8 * (pass & 7);
    if (word == 1 && st->ties_fit) return;                    // uniform for the whole grid
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long p0 = st->prefix[0], p1 = st->prefix[1];
    const unsigned long long himask = (pass & 7) == 0 ? 0ull : ~0ull << (shift + 8);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        const Cand c = cands[i];
        if (c.pad) continue;
        const unsigned long long k0 = (unsigned long long)__double_as_longlong(c.score), k1 = (unsigned long long)c.seq;
        if (word == 0) {
            if ((k0 & himask) == (p0 & himask)) atomicAdd(&h[(k0 >> shift) & 0xff], 1u);
        } else if (k0 == p0 && (k1 & himask) == (p1 & himask)) {
            atomicAdd(&h[(k1 >> shift) & 0xff], 1u);
        }
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&st->blocks_done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    h[threadIdx.x] = *(volatile unsigned *)&st->hist[threadIdx.x];     // whole histogram, one load per thread
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long k = st->k, run = 0;
        int b = 0;
        for (; b < 256; b++) {
            if (run + h[b] >= k) break;
            run += h[b];
        }
        if (b == 256) {                               // fewer than k candidates: keep everything
            st->prefix[word] |= 0xffull << shift;
            st->k = 1ull << 62;
        } else {
            st->prefix[word] |= (unsigned long long)b << shift;
            st->k = k - run;
            if (word == 0) st->below += run;
            if (pass == 7) {
                // every candidate with exactly the threshold score: if they all fit next to the ones
                // below it, keep them all and let the rank-merge order them -- no seq passes needed
                if (st->below + h[b] <= (unsigned long long)tile) { st->ties_fit = 1; st->prefix[1] = ~0ull; }
            }
        }
        st->blocks_done = 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) st->hist[i] = 0;
}

__global__ void __launch_bounds__(256)
radix_compact(const Cand *__restrict__ cands, long long m, RadixState *st, Cand *__restrict__ out, long long out_cap) {
    const unsigned long long t0 = st->prefix[0], t1 = st->prefix[1];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        const Cand c = cands[i];
        if (c.pad) continue;
        const unsigned long long k0 = (unsigned long long)__double_as_longlong(c.score), k1 = (unsigned long long)c.seq;
        if (k0 < t0 || (k0 == t0 && k1 <= t1)) {
            const unsigned long long slot = atomicAdd(&st->kept, 1ull);
            if ((long long)slot < out_cap) out[slot] = c;
        }
    }
}

// Radix pre-selection, second form (the one the last-level chunks use): the candidates of a burst agree in their
// leading score bits (scores of neighbouring trees agree to five digits), so byte-wise passes from the top spend
// most launches on bytes that cannot separate anything.  leaf_kernel keeps the smallest and largest candidate key;
// the passes start at the first bit where the two differ, take 12 bits at a time (4096-bin histogram in shared
// memory) and stop as soon as the candidates at or below the bucket of the k-th smallest key fit the rank-merge
// tile.  More trees with exactly the threshold score than the tile holds (common: whole families of trees tie)
// are separated by their sequence numbers the same way, starting at the first bit in which the sequence numbers
// of the chunk differ: typically five or six passes over the candidates instead of sixteen.
struct Radix12State {
    unsigned long long prefix[2];   // decided high bits of the threshold key: [0] score bits, [1] seq
    unsigned long long k;           // rank still to find inside the current bucket
    unsigned long long below;       // candidates below the current bucket
    unsigned long long kept;        // candidates compacted
    unsigned long long thresh[2];   // done: keep every key (score bits, seq) <= thresh
    int word;                       // 0: digits of the score, 1: digits of seq among the candidates with exactly the threshold score
    int shift, width;               // current digit: bits [shift, shift + width) of the current word; shift < 0: none left
    int seq_rest;                   // low bits in which the sequence numbers of this chunk differ
    unsigned done, blocks_done;
    unsigned hist[4096];
};

__global__ void radix12_init(Radix12State *st, const TopMeta *meta, unsigned long long k, unsigned long long seq_lo,
                             unsigned long long seq_hi) {
    if (threadIdx.x == 0) {
        const unsigned long long kmin = meta->kmin, kmax = meta->kmax;
        st->k = k; st->below = 0; st->kept = 0; st->blocks_done = 0; st->done = 0; st->thresh[0] = st->thresh[1] = 0;
        st->word = 0;
        const unsigned long long sx = seq_lo ^ seq_hi;
        const int slead = sx ? __clzll((long long)sx) : 64;
        st->seq_rest = 64 - slead;
        st->prefix[1] = slead == 0 ? 0ull : slead == 64 ? seq_lo : seq_lo & ~(~0ull >> slead);
        if (kmin > kmax) {                              // no candidate at all
            st->done = 1; st->thresh[0] = st->thresh[1] = ~0ull; st->prefix[0] = 0; st->shift = -1; st->width = 0;
        } else {
            const unsigned long long x = kmin ^ kmax;
            const int lead = x ? __clzll((long long)x) : 64;
            const int rest = 64 - lead;                 // low bits in which the candidate scores differ
            st->prefix[0] = lead == 0 ? 0ull : lead == 64 ? kmin : kmin & ~(~0ull >> lead);
            if (rest == 0) {                            // one score: straight to the sequence numbers
                st->word = 1;
                st->width = st->seq_rest < 12 ? st->seq_rest : 12;
                st->shift = st->seq_rest == 0 ? -1 : st->seq_rest - st->width;
            } else {
                st->width = rest < 12 ? rest : 12;
                st->shift = rest - st->width;
            }
        }
    }
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) st->hist[i] = 0;
}

__global__ void __launch_bounds__(256)
radix12_pass(const unsigned long long *__restrict__ key0, const unsigned long long *__restrict__ key1, long long m,
             Radix12State *st, unsigned tile) {
    __shared__ unsigned h[4096];
    __shared__ unsigned part[256];
    __shared__ bool last;
    if (st->done || st->shift < 0) return;              // uniform for the whole grid
    const int word = st->word, shift = st->shift, width = st->width, hi = shift + width;
    const unsigned long long p0 = st->prefix[0], p1 = st->prefix[1];
    const unsigned dmask = (1u << width) - 1u;
    for (int i = threadIdx.x; i < 4096; i += 256) h[i] = 0;
    __syncthreads();
    // four keys per thread and iteration, loaded together (the loop is bound by the latency of its loads)
    for (long long base = (long long)blockIdx.x * 1024; base < m; base += (long long)gridDim.x * 1024) {
        unsigned long long k0v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const long long i = base + j * 256 + threadIdx.x;
            k0v[j] = i < m ? key0[i] : ~0ull;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const unsigned long long k0 = k0v[j];
            if (k0 == ~0ull) continue;                   // empty slot
            if (word == 0) {
                if (hi >= 64 || (k0 >> hi) == (p0 >> hi)) atomicAdd(&h[(unsigned)(k0 >> shift) & dmask], 1u);
            } else if (k0 == p0) {
                const unsigned long long k1 = key1[base + j * 256 + threadIdx.x];
                if (hi >= 64 || (k1 >> hi) == (p1 >> hi)) atomicAdd(&h[(unsigned)(k1 >> shift) & dmask], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4096; i += 256) if (h[i]) atomicAdd(&st->hist[i], h[i]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&st->blocks_done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    unsigned mine = 0;                                   // 16 bins per thread, then a scan of the 256 partial sums
    for (int j = 0; j < 16; j++) { const unsigned v = *(volatile unsigned *)&st->hist[threadIdx.x * 16 + j]; h[threadIdx.x * 16 + j] = v; mine += v; }
    part[threadIdx.x] = mine;
    __syncthreads();
    __shared__ int g_s;
    __shared__ unsigned long long run_s;
    if (threadIdx.x < 32) {
        // warp 0: the group of 16 bins in which the running count reaches k (8 partial sums per lane, then a shuffle scan)
        const unsigned long long k = st->k;
        unsigned long long v[8], tot = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { v[j] = part[threadIdx.x * 8 + j]; tot += v[j]; }
        unsigned long long inc = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(kFull, inc, d);
            if ((int)threadIdx.x >= d) inc += t;
        }
        const unsigned reach = __ballot_sync(kFull, inc >= k);
        if (reach == 0) { if (threadIdx.x == 0) { g_s = 256; run_s = 0; } }
        else if ((int)threadIdx.x == __ffs(reach) - 1) {
            unsigned long long run = inc - tot;
            int j = 0;
            for (; j < 8; j++) { if (run + v[j] >= k) break; run += v[j]; }
            g_s = threadIdx.x * 8 + j; run_s = run;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long k = st->k, run = run_s;
        const int g = g_s;
        if (g == 256) {                                 // fewer than k candidates in the bucket: keep everything
            st->done = 1; st->thresh[0] = st->thresh[1] = ~0ull;
        } else {
            int b = g * 16;
            for (;; b++) { if (run + h[b] >= k) break; run += h[b]; }
            const unsigned long long pre = (word == 0 ? p0 : p1) | ((unsigned long long)b << shift);
            st->prefix[word] = pre; st->k = k - run; st->below += run;
            const unsigned long long low = shift > 0 ? (1ull << shift) - 1ull : 0ull;
            if (st->below + h[b] <= (unsigned long long)tile) {
                st->done = 1;
                if (word == 0) { st->thresh[0] = pre | low; st->thresh[1] = ~0ull; }
                else { st->thresh[0] = p0; st->thresh[1] = pre | low; }
            } else if (shift > 0) {
                const int w = shift < 12 ? shift : 12;
                st->width = w; st->shift = shift - w;
            } else if (word == 0) {
                // more trees with exactly the threshold score than the tile holds: the sequence numbers decide
                st->word = 1;
                st->width = st->seq_rest < 12 ? st->seq_rest : 12;
                st->shift = st->seq_rest == 0 ? -1 : st->seq_rest - st->width;
            } else {
                st->shift = -1;                         // (keys are unique: not reached)
            }
        }
        st->blocks_done = 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4096; i += 256) st->hist[i] = 0;
}

__global__ void __launch_bounds__(256)
radix12_compact(const uint4 *__restrict__ c16, int part, const unsigned long long *__restrict__ key0,
                const unsigned long long *__restrict__ key1, long long m, Radix12State *st, Cand *__restrict__ out, long long out_cap) {
    if (!st->done) return;
    const unsigned long long t0 = st->thresh[0], t1 = st->thresh[1];
    for (long long base = (long long)blockIdx.x * 1024; base < m; base += (long long)gridDim.x * 1024) {
        unsigned long long k0v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const long long i = base + j * 256 + threadIdx.x;
            k0v[j] = i < m ? key0[i] : ~0ull;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const long long i = base + j * 256 + threadIdx.x;
            const unsigned long long k0 = k0v[j];
            if (k0 == ~0ull) continue;
            if (k0 < t0 || (k0 == t0 && key1[i] <= t1)) {
                const unsigned long long k1 = key1[i];
                const unsigned long long slot = atomicAdd(&st->kept, 1ull);
                if ((long long)slot < out_cap) out[slot] = cand_record(c16[i], k0, k1, part);   // only the survivors' records are built
            }
        }
    }
}

__global__ void reset_n_cand(TopMeta *meta) { meta->n_cand = 0; meta->n_valid = 0; meta->next_run = 0; meta->n_bounded = 0; meta->n_eval = 0; meta->kmin = ~0ull; meta->kmax = 0ull; }

__global__ void init_meta(TopMeta *meta) {
    meta->count = 0;
    meta->pad = 0;
    meta->tau = INFINITY;
    meta->n_cand = 0;
    meta->n_valid = 0;
    meta->next_run = 0; meta->n_bounded = 0; meta->n_eval = 0; meta->kmin = ~0ull; meta->kmax = 0ull;
}

// unpack exported records {score, seq, part, pad, bytes} into Cand entries for a merge
__global__ void records_to_cands(const unsigned char *records, long long record_bytes, long long n,
                                 Cand *cands, TopMeta *meta) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const unsigned char *r = records + i * record_bytes;
        Cand c;
        c.score = *reinterpret_cast<const double *>(r);
        c.seq = *reinterpret_cast<const long long *>(r + 8);
        c.part = *reinterpret_cast<const int *>(r + 16);
        c.origin = (unsigned)i;
        c.pstar = 0;
        c.pad = 0;
        const int valid = *reinterpret_cast<const int *>(r + 20);
        if (!valid) { c.score = INFINITY; c.seq = 0x7fffffffffffffffll; c.part = 0x7fffffff; c.pad = 1; }
        cands[i] = c;
    }
    if (i == 0) { meta->count = 0; meta->tau = INFINITY; meta->n_cand = (unsigned long long)n; }
}

__global__ void export_records(const TopMeta *meta, const Cand *top, const unsigned *tree, int wstride,
                               int num_save, unsigned char *records, long long record_bytes) {
    const int t = blockIdx.x;
    if (t >= num_save) return;
    unsigned char *r = records + (size_t)t * record_bytes;
    const bool valid = (unsigned)t < meta->count;
    if (threadIdx.x == 0) {
        *reinterpret_cast<double *>(r) = valid ? top[t].score : INFINITY;
        *reinterpret_cast<long long *>(r + 8) = valid ? top[t].seq : 0;
        *reinterpret_cast<int *>(r + 16) = valid ? top[t].part : 0;
        *reinterpret_cast<int *>(r + 20) = valid ? 1 : 0;
    }
    if ((int)threadIdx.x < wstride)
        reinterpret_cast<unsigned *>(r + 24)[threadIdx.x] = valid ? tree[(size_t)t * wstride + threadIdx.x] : 0xffffffffu;
}

// ------------------------------------------------------------------------------------------
// Reference-order mode (lichee_replay.cuh): the reference's own traversal, replayed by one warp.
//
// replay_kernel   one block; the state blob lives in shared memory when it fits next to the centroid
//                 matrix (every real data set), else it is used in place in global memory.  Lane 0 runs
//                 the transitions; at R_CHECK the warp evaluates the sum rule (lanes own samples, the
//                 running child sums psum[] are lane-private), at R_RECORD it writes the tree as one
//                 record: {score, seq, part, valid, A[stride], B[stride]} with A = the nodes in the order
//                 they joined the tree (treeNodes without the root) and B = their parents, both as ROW
//                 indices in the frontier-item byte layout (item_byte).  Resumable: returns when the
//                 record buffer is full, the step quota is used up or the grow() budget is reached.
// score_ref_kernel one warp per recorded tree: PHYTree.computeErrorScore (PHYTree.java:214-233) in the
//                 reference's accumulation order -- parents in descending node id, per parent the
//                 samples 0..S-1, children added in insertion order, err += (sum - VAF)^2 one term after
//                 the other, then Math.sqrt.  No FMA contraction (explicit _rn intrinsics).
// ------------------------------------------------------------------------------------------
// PHYTree.computeErrorScore (PHYTree.java:214-233) of one recorded tree, by one warp, in the reference's
// accumulation order: parents in descending node id (desc: rows in that order), per parent the samples 0..S-1,
// children added in insertion (treeNodes) order, err += (sum - VAF)^2 one term after the other, Math.sqrt.
// wA / wB: this lane's word of the record's A (treeNodes) and B (parents) byte strings.  Explicit _rn
// intrinsics: no FMA contraction.
template <int SPL>
__device__ __forceinline__ double score_record(unsigned wA, unsigned wB, const double *vaf_s, const uint8_t *desc, int n, int S,
                                               int SP, int L, int W, int lane) {
    const int rmax = (L - 1) / W;
    const bool valid0 = lane < S, valid1 = SPL == 2 && lane + 32 < S;
    unsigned pm[4];
    parent_presence(wB, rmax, pm);
    double err = 0.0;
    for (int kk = 0; kk < n; kk++) {
        const unsigned p = desc[kk];
        if (!((pm[p >> 5] >> (p & 31)) & 1u)) continue;
        const unsigned m = __vcmpeq4(wB, p * 0x01010101u);
        double s0 = 0.0, s1 = 0.0;
        for (int r = 0; r <= rmax; r++) {
            unsigned kids = __ballot_sync(kFull, (m >> (8 * r)) & 1u);
            while (kids) {                               // children in insertion (treeNodes) order
                const int j = __ffs(kids) - 1;
                kids &= kids - 1;
                const unsigned c = (__shfl_sync(kFull, wA, j) >> (8 * r)) & 0xffu;
                double x0, x1;
                load_row<SPL>(vaf_s, (int)c, SP, lane, x0, x1);
                s0 = __dadd_rn(s0, x0);
                if (SPL == 2) s1 = __dadd_rn(s1, x1);
            }
        }
        double a0, a1;
        load_row<SPL>(vaf_s, (int)p, SP, lane, a0, a1);
        double t0 = 0.0, t1 = 0.0;
        if (valid0 && s0 > a0) { const double dd = __dsub_rn(s0, a0); t0 = __dmul_rn(dd, dd); }
        if (SPL == 2 && valid1 && s1 > a1) { const double dd = __dsub_rn(s1, a1); t1 = __dmul_rn(dd, dd); }
        unsigned pos = __ballot_sync(kFull, t0 > 0.0);
        while (pos) {                                    // samples 0..31, one after the other
            const int j = __ffs(pos) - 1;
            pos &= pos - 1;
            err = __dadd_rn(err, __shfl_sync(kFull, t0, j));
        }
        if (SPL == 2) {
            pos = __ballot_sync(kFull, t1 > 0.0);
            while (pos) {                                // samples 32..63
                const int j = __ffs(pos) - 1;
                pos &= pos - 1;
                err = __dadd_rn(err, __shfl_sync(kFull, t1, j));
            }
        }
    }
    return __dsqrt_rn(err);
}

struct ReplayOut {
    long long num_grow, num_checks, num_trees, steps;
    int status, rec_count;
    int fused;                  // 1: the kernel also scored and ranked the trees; the ranked list is in the result block
    int count;                  // fused: trees in the ranked list
};

constexpr int kReplayThreads = 128;
constexpr int kReplaySmemMax = 208 * 1024;  // dynamic shared memory the replay may take (227 KB per block on sm_100a, 17 KB of it static)
constexpr int kReplayRecCap = 1 << 15;       // trees one replay launch may record before the batch is scored and merged
constexpr size_t kPinnedBlob = 256 * 1024;   // pinned: initial replay state (largest layout: 128 nodes, 64 samples, 16256 edges)
constexpr size_t kPinnedResult = (size_t)LICHEE_MAX_SAVE * (32 + 2 * 128);   // pinned: Cand | tree bytes of a fused result
constexpr size_t kPinnedVaf = (size_t)LICHEE_MAX_NODES * 64 * 8;              // pinned: centroid matrix in row space (what the replay stages)
constexpr size_t kPinnedDesc = 256;                                           // pinned: rows in descending node id
constexpr int kFuseMax = 1024;               // a search with at most this many trees is scored and ranked by the replay launch itself

struct FuseKey { double score; unsigned idx, pad; };

struct ReplayArgs {             // (one struct: the argument list had outgrown a readable signature)
    replay::Layout lay;
    const unsigned char *blob_src;   // state to resume from: the pinned initial state (first launch) or blob_dst
    unsigned char *blob_dst;         // device copy the state is saved to when the launch pauses
    int blob_in_smem;
    const double *vaf;               // centroid matrix, row space
    int S, L, stride;
    double margin;
    unsigned char *records;          // record buffer of this launch
    long long record_bytes;
    int rec_cap;
    long long step_quota, grow_limit;
    ReplayOut *out;                  // pinned, device-visible
    // first launch only: initialise the ranked list; fuse: score + rank here when the whole search fits
    int first, fuse, num_save;
    TopMeta *meta;
    const uint8_t *desc_rows;
    unsigned char *result;           // pinned, device-visible: Cand[num_save] | tree bytes (2 x stride each)
};

// SPL: samples per lane (1: <= 32 samples, 2: <= 64).  NG: 32-node groups (1: <= 32 nodes, the latency-oriented loop; 4).
// BS: the state blob is worked on in shared memory -- a template parameter so that every access to it compiles to
// LDS/STS with 32-bit addresses; with a run-time choice between shared and global memory all of them were generic
// loads with 64-bit address arithmetic (a third more instructions, measured).
template <int SPL, int NG, bool BS>
__global__ void __launch_bounds__(kReplayThreads, 1)
replay_kernel(const __grid_constant__ ReplayArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int fuse_count_s;
    const replay::Layout &lay = a.lay;
    double *vaf_s = reinterpret_cast<double *>(smem_raw);
    const int SP = lay.SP, S = a.S, L = a.L, stride = a.stride;
    const double margin = a.margin;
    unsigned char *records = a.records;
    const long long record_bytes = a.record_bytes;
    const int rec_cap = a.rec_cap;
    ReplayOut *out = a.out;
    const unsigned vaf_bytes = (unsigned)lay.n * SP * 8u;
    // the state is worked on in shared memory when it fits, else in place in the device copy
    unsigned char *blob = BS ? smem_raw + vaf_bytes : a.blob_dst;
    {
        // The first launch reads the matrix and the initial state straight from pinned host memory: eight loads per
        // thread are requested before the first one is stored, so a small search pays one or two bus round trips here,
        // not one per 2 KB.
        auto stage = [&](const uint4 *src, uint4 *dst, unsigned n16) {
            for (unsigned base = 0; base < n16; base += blockDim.x * 8) {
                uint4 r[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const unsigned i = base + j * blockDim.x + threadIdx.x;
                    if (i < n16) r[j] = src[i];
                }
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const unsigned i = base + j * blockDim.x + threadIdx.x;
                    if (i < n16) dst[i] = r[j];
                }
            }
        };
        stage(reinterpret_cast<const uint4 *>(a.vaf), reinterpret_cast<uint4 *>(vaf_s), vaf_bytes / 16);
        if (BS || a.blob_src != a.blob_dst)
            stage(reinterpret_cast<const uint4 *>(a.blob_src), reinterpret_cast<uint4 *>(blob), lay.total / 16);
        if (a.first && threadIdx.x == 0) {
            TopMeta *meta = a.meta;
            meta->count = 0; meta->pad = 0; meta->tau = INFINITY; meta->n_cand = 0; meta->n_valid = 0;
            meta->next_run = 0; meta->n_bounded = 0; meta->n_eval = 0; meta->kmin = ~0ull; meta->kmax = 0ull;
        }
        if (threadIdx.x == 0) fuse_count_s = -1;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        // The warp statement of the machine in lichee_replay.cuh (step() there is the sequential statement).
        // All 32 lanes run the same transitions; the scalars of the Header live in registers, identical in
        // every lane; arrays are written by lane 0 (single elements) or by the lanes that own an element of a
        // list scan; a __syncwarp() separates every write from a read by another lane.
        using namespace replay;
        const int lane = threadIdx.x;
        const unsigned lt = (1u << lane) - 1u;
        State s = bind(blob, lay);
        const int n = lay.n;
        const int W = stride >> 2;
        const bool valid0 = lane < S, valid1 = SPL == 2 && lane + 32 < S;
        long long num_grow = s.h->num_grow, num_checks = s.h->num_checks, num_trees = s.h->num_trees;
        long long compactions = s.h->compactions;
        const long long max_grow = s.h->max_grow, max_trees = s.h->max_trees;
        int top = s.h->top, f_live = s.h->f_live, ff_top = s.h->ff_top, rem_top = s.h->rem_top;
        int depth = s.h->depth, tcount = s.h->tcount, phase = s.h->phase, status = s.h->status;
        unsigned t0 = s.h->inT[0], t1 = s.h->inT[1], t2 = s.h->inT[2], t3 = s.h->inT[3];
        const long long grow_start = num_grow;
        const long long step_quota = a.step_quota, grow_limit = a.grow_limit;
        int rec_count = 0;
        // NG = 1: at most 32 nodes (every real data set), one membership word and one pass per list scan
        auto tword = [&](int g) { return NG == 1 ? t0 : g == 0 ? t0 : g == 1 ? t1 : g == 2 ? t2 : t3; };
        auto in_t = [&](int v) { return NG == 1 ? (t0 >> v) & 1u : (tword(v >> 5) >> (v & 31)) & 1u; };
        auto compact = [&]() {                           // squeeze the tombstones out of fs
            __syncwarp();
            int k = 0;
            for (int base = 0; base < top; base += 32) {
                const int i = base + lane;
                const unsigned c = i < top ? s.fs[i] : kNone;
                const unsigned live = __ballot_sync(kFull, c != kNone);
                __syncwarp();
                if (c != kNone) {
                    const int d = k + __popc(live & lt);
                    s.fs[d] = (uint16_t)c;
                    s.pos[pos_index(n, c >> 7, c & 127)] = (uint16_t)d;
                }
                k += __popc(live);
                __syncwarp();
            }
            top = k;
            compactions++;
        };
        auto push_lanes = [&](bool take, unsigned code) {   // f.add(e) for every taking lane, in lane order
            if (top + 32 > lay.fs_cap) compact();
            const unsigned m = __ballot_sync(kFull, take);
            if (take) {
                const int d = top + __popc(m & lt);
                s.fs[d] = (uint16_t)code;
                s.pos[pos_index(n, code >> 7, code & 127)] = (uint16_t)d;
            }
            const int c = __popc(m);
            top += c;
            f_live += c;
        };
        const long long stop_at = grow_limit < grow_start + step_quota ? grow_limit : grow_start + step_quota;
        auto record_tree = [&]() {                       // spanningTrees.add(L.clone()): one record, words split over the lanes
            unsigned char *rp = records + (size_t)rec_count * record_bytes;
            if (lane == 0) {
                *reinterpret_cast<double *>(rp) = 0.0;
                *reinterpret_cast<long long *>(rp + 8) = num_trees;
                *reinterpret_cast<int *>(rp + 16) = 0;
                *reinterpret_cast<int *>(rp + 20) = 1;
            }
            if (lane < W) {
                unsigned wa = 0, wb = 0;
                for (int q = 3; q >= 0; q--) {
                    const int i = q * W + lane;          // tree position i <-> treeNodes[i + 1]
                    unsigned ab = 0xffu, bb = 0xffu;
                    if (i < L) { ab = s.tn[i + 1]; bb = s.par[ab]; }
                    wa = (wa << 8) | ab;
                    wb = (wb << 8) | bb;
                }
                reinterpret_cast<unsigned *>(rp + 24)[lane] = wa;
                reinterpret_cast<unsigned *>(rp + 24 + stride)[lane] = wb;
            }
            rec_count++;
            num_trees++;
        };
        while (status == S_RUNNING) {
            __syncwarp();
            if (phase == P_ENTER) {                      // grow(t), PHYNetwork.java:443-449
                if (num_grow >= stop_at) break;          // pause before counting the call
                if (tcount == n) {
                    if (rec_count >= rec_cap) break;     // record buffer full: pause
                    num_grow++;
                    record_tree();
                    if (depth == 0) { status = S_DONE; break; }      // only when the network is the root alone
                    depth--;
                    phase = P_AFTER_RETURN;
                    continue;
                }
                num_grow++;
                if (lane == 0) s.ffbase[depth] = (uint16_t)ff_top;
                phase = P_LOOP;
            }
            if (phase == P_LOOP) {                       // while(!b && f.size() > 0), :454-495
                if (f_live == 0) { phase = P_FINISH; continue; }
                unsigned code;
                for (;;) {                               // pop: the topmost live entry
                    const int i = top - 1 - lane;
                    const unsigned c = i >= 0 ? s.fs[i] : kNone;
                    const unsigned m = __ballot_sync(kFull, c != kNone);
                    if (m) { const int j = __ffs(m) - 1; code = __shfl_sync(kFull, c, j); top = top - 1 - j; break; }
                    top -= 32;
                }
                f_live--;
                const int u = code >> 7, v = code & 127;
                const int k = tcount;                    // t.addNode(v); t.addEdge(u, v)
                const int prev = s.last[u];
                if (lane == 0) {
                    s.pos[pos_index(n, u, v)] = kNone;
                    s.fu[depth] = (uint8_t)u; s.fv[depth] = (uint8_t)v;
                    s.tn[k] = (uint8_t)v; s.par[v] = (uint8_t)u; s.prev[k] = (uint8_t)prev; s.last[u] = (uint8_t)k;
                }
                tcount = k + 1;
                if (NG == 1) t0 |= 1u << v;
                else { const unsigned bit = 1u << (v & 31); const int g = v >> 5; t0 |= g == 0 ? bit : 0u; t1 |= g == 1 ? bit : 0u; t2 |= g == 2 ? bit : 0u; t3 |= g == 3 ? bit : 0u; }
                num_checks++;
                bool ok;
                {   // t.checkConstraint(u), PHYTree.java:156-174: running sum of u's children, lanes own samples
                    double x0, x1, a0, a1, p0 = 0.0, p1 = 0.0;
                    load_row<SPL>(vaf_s, v, SP, lane, x0, x1);
                    load_row<SPL>(vaf_s, u, SP, lane, a0, a1);
                    if (prev != kNoChild) load_row<SPL>(s.psum, prev, SP, lane, p0, p1);
                    const double s0 = __dadd_rn(p0, x0), s1 = __dadd_rn(p1, x1);
                    if (SPL == 2) reinterpret_cast<double2 *>(s.psum)[k * 32 + lane] = make_double2(s0, s1);
                    else s.psum[k * SP + lane] = s0;
                    bool bad = valid0 && s0 > __dadd_rn(a0, margin);
                    if (SPL == 2) bad |= valid1 && s1 > __dadd_rn(a1, margin);
                    ok = !__any_sync(kFull, bad);
                }
                if (!ok) { phase = P_AFTER; continue; }
                __syncwarp();
                {   // edgesAdded: (v, w) for w in edges.get(v), w not in t, :461-470
                    const int len = s.adjlen[v];
                    for (int base = 0; base < (NG == 1 ? 1 : len); base += 32) {
                        const int i = base + lane;
                        const unsigned w = i < len ? s.adj[v * n + i] : 0u;
                        push_lanes(i < len && !in_t((int)w), ((unsigned)v << 7) | w);
                    }
                }
                {   // edgesRemoved: every (w, v) in f, in f order, :473-480 (every edge of f leaves a tree node)
                    unsigned p[NG], has[NG], rank[NG];
                    int cnt = 0;
#pragma unroll
                    for (int g = 0; g < NG; g++) {
                        const int w = g * 32 + lane;
                        const bool cand = g * 32 < n && ((s.inmask[v * 4 + g] & tword(g)) >> lane) & 1u;
                        p[g] = cand ? s.pos[pos_index(n, w, v)] : kNone;
                        has[g] = __ballot_sync(kFull, p[g] != kNone);
                        cnt += __popc(has[g]);
                        rank[g] = 0;
                    }
                    const int base = rem_top;
                    if (cnt > 0) {
                        if (cnt > 1) {
#pragma unroll
                            for (int g = 0; g < NG; g++) {
                                unsigned m = has[g];
                                while (m) {
                                    const int j = __ffs(m) - 1;
                                    m &= m - 1;
                                    const unsigned pj = __shfl_sync(kFull, p[g], j);
#pragma unroll
                                    for (int g2 = 0; g2 < NG; g2++) rank[g2] += pj < p[g2];
                                }
                            }
                        }
#pragma unroll
                        for (int g = 0; g < NG; g++) {
                            if (p[g] != kNone) {
                                const int w = g * 32 + lane;
                                s.rem[base + rank[g]] = (uint16_t)((w << 7) | v);
                                s.fs[p[g]] = kNone;
                                s.pos[pos_index(n, w, v)] = kNone;
                            }
                        }
                        f_live -= cnt;
                        rem_top = base + cnt;
                    }
                    if (lane == 0) { s.rembase[depth] = (uint16_t)base; s.remcnt[depth] = (uint16_t)cnt; }
                }
                if (num_grow >= max_grow) { status = S_GROW_CAP; break; }   // :490
                depth++;                                 // grow(t), :495
                phase = P_ENTER;
                continue;
            }
            if (phase == P_AFTER_RETURN) {               // :497-505
                if (num_trees == max_trees) { status = S_TREE_CAP; break; }
                const int v = s.fv[depth];
                {   // f.removeAll(edgesAdded): the edges of f that leave v
                    const int len = s.adjlen[v];
                    int killed = 0;
                    for (int base = 0; base < (NG == 1 ? 1 : len); base += 32) {
                        const int i = base + lane;
                        unsigned pp = kNone;
                        int w = 0;
                        if (i < len) { w = s.adj[v * n + i]; pp = s.pos[pos_index(n, v, w)]; }
                        if (pp != kNone) { s.fs[pp] = kNone; s.pos[pos_index(n, v, w)] = kNone; }
                        killed += __popc(__ballot_sync(kFull, pp != kNone));
                    }
                    f_live -= killed;
                }
                {   // f.addAll(edgesRemoved)
                    const int base = s.rembase[depth], cnt = s.remcnt[depth];
                    for (int j0 = 0; j0 < (NG == 1 ? (cnt > 0) : cnt); j0 += 32) {
                        const int j = j0 + lane;
                        push_lanes(j < cnt, j < cnt ? s.rem[base + j] : 0u);
                    }
                    rem_top = base;
                }
                phase = P_AFTER;
                __syncwarp();
            }
            if (phase == P_AFTER) {                      // :508-530
                const int u = s.fu[depth], v = s.fv[depth];
                const int k = tcount - 1;                // t.removeEdge(u, v): v is the newest node
                const int pv = s.prev[k];
                tcount = k;
                if (NG == 1) t0 &= ~(1u << v);
                else { const unsigned bit = ~(1u << (v & 31)); const int g = v >> 5; t0 &= g == 0 ? bit : ~0u; t1 &= g == 1 ? bit : ~0u; t2 &= g == 2 ? bit : ~0u; t3 &= g == 3 ? bit : ~0u; }
                const int len = s.adjlen[u];             // this.removeEdge(u, v): shift the tail of the list left
                const int dv = (int)s.indeg[v] - 1;
                int found = -1;
                for (int base = 0; base < (NG == 1 ? 1 : len); base += 32) {
                    const int i = base + lane;
                    const unsigned a = i < len ? s.adj[u * n + i] : 0xffu;
                    const unsigned m = __ballot_sync(kFull, i < len && a == (unsigned)v);
                    if (found < 0 && m) found = base + __ffs(m) - 1;
                    __syncwarp();
                    if (found >= 0 && i < len && i > found) s.adj[u * n + i - 1] = (uint8_t)a;
                }
                if (lane == 0) {
                    s.last[u] = (uint8_t)pv;
                    s.adjlen[u] = (uint8_t)(len - 1);
                    s.indeg[v] = (uint8_t)dv;
                    s.ff[ff_top] = (uint16_t)((u << 7) | v);         // ff.add(e)
                }
                ff_top++;
                phase = dv == 0 ? P_FINISH : P_LOOP;     // bridge test (lichee_replay.cuh, header comment)
                if (phase == P_LOOP) continue;
                __syncwarp();
            }
            if (phase == P_FINISH) {                     // :534-540: re-push ff in reverse, re-append to the lists
                const int base = s.ffbase[depth], cnt = ff_top - base;
                for (int j0 = 0; j0 < cnt; j0 += 32) {
                    const int j = j0 + lane;
                    const bool take = j < cnt;
                    const unsigned code = take ? s.ff[ff_top - 1 - j] : 0u;
                    push_lanes(take, code);
                    const int u = code >> 7, v = code & 127;
                    // several of the edges may leave the same u / enter the same v: group them, keep lane order
                    const unsigned gu = __match_any_sync(kFull, take ? u : 256 + lane);
                    const unsigned gv = __match_any_sync(kFull, take ? v : 256 + lane);
                    const int al = take ? s.adjlen[u] : 0, di = take ? s.indeg[v] : 0;
                    __syncwarp();
                    if (take) {
                        s.adj[u * n + al + __popc(gu & lt)] = (uint8_t)v;
                        if ((gu >> lane) == 1u) s.adjlen[u] = (uint8_t)(al + __popc(gu));     // highest lane of the group
                        if ((gv >> lane) == 1u) s.indeg[v] = (uint8_t)(di + __popc(gv));
                    }
                    __syncwarp();
                }
                ff_top = base;
                if (depth == 0) { status = S_DONE; break; }
                depth--;
                phase = P_AFTER_RETURN;
            }
        }
        __syncwarp();
        if (lane == 0) {
            Header *h = s.h;
            h->num_grow = num_grow; h->num_checks = num_checks; h->num_trees = num_trees; h->compactions = compactions;
            h->top = top; h->f_live = f_live; h->ff_top = ff_top; h->rem_top = rem_top;
            h->depth = depth; h->tcount = tcount; h->phase = phase; h->status = status;
            h->inT[0] = t0; h->inT[1] = t1; h->inT[2] = t2; h->inT[3] = t3;
            out->num_grow = num_grow; out->num_checks = num_checks; out->num_trees = num_trees;
            out->steps = num_grow - grow_start; out->status = status; out->rec_count = rec_count;
            // the whole search is in this launch's record buffer and small: score and rank it right here
            const bool fuse = a.fuse && status != S_RUNNING && num_trees == (long long)rec_count && rec_count <= kFuseMax;
            out->fused = fuse ? 1 : 0;
            out->count = 0;
            if (fuse) fuse_count_s = rec_count;
        }
    }
    __syncthreads();
    const int fuse_count = fuse_count_s;
    if (fuse_count >= 0) {
        // Fused finish (a search of a real data set is a handful of trees: three more launches and two more
        // host round trips would cost several times the search).  Scores: one warp per tree, the reference's
        // accumulation order (score_record).  Ranking: Collections.sort is stable, so the order is (score,
        // discovery index); bitonic sort of the keys in shared memory.  The ranked list goes straight to the
        // pinned result block: Cand[num_save] | tree bytes.
        __shared__ FuseKey key_s[kFuseMax];
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, W = stride >> 2;
        int N = 1;
        while (N < fuse_count) N <<= 1;
        for (int t = wid; t < fuse_count; t += kReplayThreads / 32) {
            unsigned char *rp = records + (size_t)t * record_bytes;
            const unsigned wA = lane < W ? reinterpret_cast<const unsigned *>(rp + 24)[lane] : 0xffffffffu;
            const unsigned wB = lane < W ? reinterpret_cast<const unsigned *>(rp + 24 + stride)[lane] : 0xffffffffu;
            const double score = score_record<SPL>(wA, wB, vaf_s, a.desc_rows, lay.n, S, SP, L, W, lane);
            if (lane == 0) { key_s[t].score = score; key_s[t].idx = (unsigned)t; }
        }
        for (int t = fuse_count + threadIdx.x; t < N; t += blockDim.x) { key_s[t].score = INFINITY; key_s[t].idx = 0xffffffffu; }
        __syncthreads();
        for (int k = 2; k <= N; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < N; t += blockDim.x) {
                    const int x = t ^ j;
                    if (x > t) {
                        const FuseKey ka = key_s[t], kb = key_s[x];
                        const bool b_less = kb.score < ka.score || (kb.score == ka.score && kb.idx < ka.idx);
                        if (b_less == ((t & k) == 0)) { key_s[t] = kb; key_s[x] = ka; }
                    }
                }
                __syncthreads();
            }
        }
        const int newc = fuse_count < a.num_save ? fuse_count : a.num_save;
        Cand *top = reinterpret_cast<Cand *>(a.result);
        unsigned *tree = reinterpret_cast<unsigned *>(a.result + sizeof(Cand) * (size_t)a.num_save);
        for (int t = threadIdx.x; t < newc; t += blockDim.x) {
            Cand c;
            c.score = key_s[t].score; c.seq = (long long)key_s[t].idx; c.part = 0;
            c.origin = 0x80000000u | (unsigned)t; c.pstar = 0; c.pad = 0;
            top[t] = c;
        }
        for (int t = wid; t < newc; t += kReplayThreads / 32) {
            const unsigned *src = reinterpret_cast<const unsigned *>(records + (size_t)key_s[t].idx * record_bytes + 24);
            for (int j = lane; j < 2 * W; j += 32) tree[(size_t)t * 2 * W + j] = src[j];
        }
        if (threadIdx.x == 0) {
            out->count = newc;
            a.meta->count = 0;                          // the device-resident list is not used by a fused result
        }
        return;                                          // the search is over: no state to save
    }
    if (BS) {
        uint4 *bd = reinterpret_cast<uint4 *>(a.blob_dst);
        const uint4 *bs = reinterpret_cast<const uint4 *>(blob);
        for (unsigned i = threadIdx.x; i < lay.total / 16; i += blockDim.x) bd[i] = bs[i];
    }
}

template <int SPL>
__global__ void __launch_bounds__(kThreads, 2)
score_ref_kernel(DevNet net, unsigned char *records, long long record_bytes, int count, Cand *__restrict__ cands,
                 TopMeta *meta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *vaf_s = reinterpret_cast<double *>(smem_raw);
    uint8_t *cand_s = reinterpret_cast<uint8_t *>(vaf_s + net.n * net.SP);
    __shared__ uint8_t desc_s[kMaxN];
    stage_network(net, 0, vaf_s, cand_s);
    for (int i = threadIdx.x; i < net.n; i += blockDim.x) desc_s[i] = net.desc_rows[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) meta->n_cand = (unsigned long long)count;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int SP = net.SP, W = net.stride >> 2, L = net.L;
    const int warp0 = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5), nwarps = gridDim.x * kWarpsPerBlock;
    for (int t = warp0; t < count; t += nwarps) {
        unsigned char *rp = records + (size_t)t * record_bytes;
        const unsigned wA = lane < W ? reinterpret_cast<const unsigned *>(rp + 24)[lane] : 0xffffffffu;
        const unsigned wB = lane < W ? reinterpret_cast<const unsigned *>(rp + 24 + net.stride)[lane] : 0xffffffffu;
        const double score = score_record<SPL>(wA, wB, vaf_s, desc_s, net.n, net.S, SP, L, W, lane);
        if (lane == 0) {
            *reinterpret_cast<double *>(rp) = score;
            Cand c;
            c.score = score;
            c.seq = *reinterpret_cast<const long long *>(rp + 8);
            c.part = 0; c.origin = (unsigned)t; c.pstar = 0; c.pad = 0;
            cands[t] = c;
        }
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side of the handle
// ------------------------------------------------------------------------------------------
namespace {

// Process-wide pool of device buffers: a search allocates its frontier levels on demand and
// returns them here on destroy, so a host that runs many searches (the fixNetwork loop, several
// datasets) pays cudaMalloc once.  Sizes are rounded up to powers of two.
struct DevicePool {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void *> free_list;
    size_t round_up(size_t n) { size_t r = 1 << 16; while (r < n) r <<= 1; return r; }
    cudaError_t get(int dev, size_t bytes, void **out, size_t *got) {
        const size_t r = round_up(bytes);
        {
            std::lock_guard<std::mutex> g(mu);
            auto it = free_list.find({dev, r});
            if (it != free_list.end()) { *out = it->second; *got = r; free_list.erase(it); return cudaSuccess; }
        }
        *got = r;
        cudaError_t e = cudaMalloc(out, r);
        if (e != cudaSuccess) {            // out of memory: drop the cache and retry once
            cudaGetLastError();
            release(dev);
            e = cudaMalloc(out, r);
        }
        return e;
    }
    void put(int dev, void *p, size_t bytes) {
        if (!p) return;
        std::lock_guard<std::mutex> g(mu);
        free_list.insert({{dev, bytes}, p});
    }
    void release(int dev) {
        std::lock_guard<std::mutex> g(mu);
        for (auto it = free_list.begin(); it != free_list.end();) {
            if (dev < 0 || it->first.first == dev) { cudaSetDevice(it->first.first); cudaFree(it->second); it = free_list.erase(it); } else ++it;
        }
    }
};
DevicePool g_pool;

// Streams, events and the pinned read-back block are pooled as well: creating them costs tens of
// milliseconds, more than a whole search of a real-data network.
struct DevCtx {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk[2][6] = {};
    cudaStream_t stream2 = nullptr;      // check2_kernel of the NEXT chunk, beside the last level of the current fill
    cudaEvent_t evp[3] = {};             // its start / end (timing) and completion (what the main stream waits for)
    void *pinned = nullptr;
    void *pinned_big = nullptr;          // initial replay state | fused result block (device-visible)
};
struct CtxPool {
    std::mutex mu;
    std::multimap<int, DevCtx> free_list;
    bool attrs_set[64] = {};
    bool attrs2_set[64] = {};
    bool get(int dev, DevCtx &c) {
        std::lock_guard<std::mutex> g(mu);
        auto it = free_list.find(dev);
        if (it == free_list.end()) return false;
        c = it->second;
        free_list.erase(it);
        return true;
    }
    void put(int dev, const DevCtx &c) { std::lock_guard<std::mutex> g(mu); free_list.insert({dev, c}); }
};
CtxPool g_ctx;

struct DevBuf {              // one growable device buffer owned by a handle
    void *p = nullptr;
    size_t bytes = 0;
    bool view = false;       // points into the handle's level slab: not returned to the pool
};

}  // namespace

struct lichee_search {
    int n = 0, S = 0, SP = 32, L = 0, stride = 16, device = 0;
    lichee_search_params prm{};
    std::vector<int> sigma;                  // node at position t
    std::vector<uint8_t> ibyte;              // item_byte(t, stride / 4) of every position (get_tree reads a thousand trees: no divisions there)
    std::vector<std::vector<int>> cands;     // candidate parents per position
    bool no_trees = false;                   // some node has no parent at all
    bool ran = false;
    lichee_search_stats stats{};
    // device
    cudaStream_t stream = nullptr;
    DevBuf b_net;                            // centroid matrix + sigma + candidate tables
    std::vector<DevBuf> b_level;             // frontier buffers, grown on demand
    std::vector<long long> head, tail;
    long long cap = 0;                       // upper bound on items per level buffer
    long long cand_cap = 0;                  // upper bound on candidates one last-level launch may append
    DevBuf b_mask, b_cnt, b_off, b_tile, b_small, b_radix, b_radix12, b_key0, b_key1, b_cands[2], b_top[2], b_tree[2];
    DevBuf b_mask2, b_mask3, b_flag, b_bound; // level L-2 chunks: "v and u" mask, "u without v" mask, bounded flag, bound value
    DevBuf b_pf[6];                          // the same six arrays (mask, count, mask2, mask3, flag, bound), second set: consecutive chunks alternate
    cudaStream_t stream2 = nullptr;          // check2_kernel of the next level L-2 chunk runs here, beside the last level of the current fill
    cudaEvent_t evp[3] = {};
    bool no_prefetch = false;                // LICHEE_NO_PREFETCH=1: every kernel on the one stream, in program order
    DevBuf b_lmask, b_lcnt, b_lbound;        // last level, parallel to its item buffer: pass mask, count | flags, bound of the base
    DevBuf b_c16;                            // 16-byte candidates of the running last-level launch (leaf_kernel -> seq_fixup_kernel)
    uint4 sok_last{};                        // static pass mask of the last node (n == 2: the only item's mask)
    DevBuf b_slab;                           // kNarrowCap items for every level (what the narrow scheduler emits into)
    DevBuf b_sched;                          // Sched, device copy
    ScanOut *h_scan = nullptr;               // pinned
    TopMeta *h_meta = nullptr;               // pinned
    Sched *h_sched = nullptr;                // pinned
    bool narrow = true;                      // device-side scheduling of narrow chunks (LICHEE_NO_NARROW=1 turns it off)
    bool no_prefilter = false;               // LICHEE_NO_PREFILTER=1: stored bounds not consulted by the last level and its scan
    int flip = 0;
    long long chunk_max = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t evk[2][6] = {};              // per-chunk kernel boundaries, double buffered
    int evn[2] = {0, 0};
    bool ev_leaf[2] = {false, false};
    // results on host
    std::vector<Cand> h_top;
    std::vector<uint8_t> h_tree;
    DevNet net{};
    size_t smem = 0, smem_leaf = 0;
    // reference-order mode (lichee_replay.cuh)
    int last_slice = -1;                     // highest slice searched into the list kept on this handle (keep_ranked)
    bool cyclic = false;                     // the network has a directed cycle: only the replay can search it
    bool ref_result = false;                 // the ranked list on the host came from the replay (records of 2 x stride bytes)
    replay::Layout rlay{};
    unsigned char *h_big = nullptr;          // pinned: [initial state of getLineageTrees() | fused result block]
    unsigned char *rblob0 = nullptr;         // = h_big: the initial state, read by the first replay launch of every run
    unsigned char *h_result = nullptr;       // = h_big + kPinnedBlob
    double *p_vaf = nullptr;                 // pinned copy of the centroid matrix (row space): the replay reads it from here
    uint8_t *p_desc = nullptr;               // pinned copy of desc_rows
    std::vector<unsigned char> tables;       // [vaf | leaf_static | static_ok | cand | ncand | desc]: uploaded by the first search that needs them
    bool tables_on_device = false;
    bool rblob_in_smem = false;
    DevBuf b_rblob, b_rrec;                  // state blob (device copy), record buffer
    ReplayOut *h_rout = nullptr;             // pinned

    unsigned *level(int lv) { return static_cast<unsigned *>(b_level[lv].p); }
    uint4 *d_mask() { return static_cast<uint4 *>(b_mask.p); }
    unsigned *d_cnt() { return static_cast<unsigned *>(b_cnt.p); }
    unsigned *d_off() { return static_cast<unsigned *>(b_off.p); }
    unsigned long long *d_tile() { return static_cast<unsigned long long *>(b_tile.p); }
    ScanOut *d_scan() { return static_cast<ScanOut *>(b_small.p); }
    TopMeta *d_meta() { return reinterpret_cast<TopMeta *>(static_cast<char *>(b_small.p) + 64); }
    Cand *d_cands(int i) { return static_cast<Cand *>(b_cands[i].p); }
    Cand *d_top(int i) { return static_cast<Cand *>(b_top[i].p); }
    unsigned *d_tree(int i) { return static_cast<unsigned *>(b_tree[i].p); }
};

namespace {

// make `b` at least `bytes` large; contents are NOT preserved (callers only grow empty buffers)
int ensure(lichee_search *h, DevBuf &b, size_t bytes) {
    if (b.bytes >= bytes && b.p) return LICHEE_OK;
    if (b.p && !b.view) g_pool.put(h->device, b.p, b.bytes);
    b.p = nullptr; b.bytes = 0; b.view = false;
    size_t got = 0;
    CK(g_pool.get(h->device, bytes, &b.p, &got));
    b.bytes = got;
    return LICHEE_OK;
}

void give_back(lichee_search *h, DevBuf &b) {
    if (!b.view) g_pool.put(h->device, b.p, b.bytes);
    b.p = nullptr; b.bytes = 0; b.view = false;
}

int grid_for(long long items) {
    long long blocks = (items + kWarpsPerBlock * 4 - 1) / (kWarpsPerBlock * 4);
    const long long maxb = 148 * 3;
    if (blocks > maxb) blocks = maxb;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <int SPL>
int launch_check(lichee_search *h, int lv, const unsigned *items, long long n_items) {
    const int grid = grid_for(n_items);
    const int run_len = (int)std::min<long long>(64, std::max<long long>(1, n_items / ((long long)grid * kWarpsPerBlock)));
    check_kernel<SPL><<<grid, kThreads, h->smem, h->stream>>>(h->net, lv, items, n_items, h->d_mask(), h->d_cnt(), run_len);
    CK(cudaGetLastError());
    return LICHEE_OK;
}

int launch_scan(lichee_search *h, const unsigned *cnt, long long n_items, unsigned long long limit, int live_only = 0,
                const double *lb = nullptr) {
    if (n_items <= kScanSingleMax) {
        scan_single<<<1, kScanSingleThreads, 0, h->stream>>>(cnt, h->d_off(), n_items, limit, h->d_scan(), live_only, lb, h->d_meta(),
                                                           h->prm.num_save);
        h->stats.num_launches += 1;
        h->stats.kernel_launches[LICHEE_K_SCAN] += 1;
    } else {
        const int tiles = (int)((n_items + kScanTile - 1) / kScanTile);
        scan_tiles<<<tiles, kScanThreads, 0, h->stream>>>(cnt, h->d_off(), h->d_tile(), n_items, live_only, lb, h->d_meta(), h->prm.num_save);
        scan_tile_sums<<<1, 32, 0, h->stream>>>(h->d_tile(), tiles, h->d_scan());
        scan_finish<<<tiles, kScanThreads, 0, h->stream>>>(cnt, h->d_off(), h->d_tile(), n_items, limit, h->d_scan(), live_only, lb,
                                                          h->d_meta(), h->prm.num_save);
        h->stats.num_launches += 3;
        h->stats.kernel_launches[LICHEE_K_SCAN] += 3;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->h_scan, h->d_scan(), sizeof(ScanOut), cudaMemcpyDeviceToHost, h->stream));
    return LICHEE_OK;
}

// Device-side set-up of everything beyond the replay's first launch: the network tables, the level slab, the
// scheduler block and the double-buffered ranked list.  Stream-ordered; called by the first search that needs it.
int upload_tables(lichee_search *h) {
    if (h->tables_on_device) return LICHEE_OK;
    const int Lp = std::max(1, h->L);
    CK(cudaMemcpyAsync(h->b_net.p, h->tables.data(), h->tables.size(), cudaMemcpyHostToDevice, h->stream));
    {
        // every level starts as a view of kNarrowCap items into one slab; a level that needs more gets
        // its own buffer from the pool (ensure)
        const size_t per = (size_t)kNarrowCap * h->stride;
        if (ensure(h, h->b_slab, per * Lp) != LICHEE_OK) return LICHEE_ERR_CUDA;
        for (int lv = 0; lv < Lp; lv++) {
            if (!h->b_level[lv].p || h->b_level[lv].view) {
                h->b_level[lv].p = static_cast<char *>(h->b_slab.p) + per * lv;
                h->b_level[lv].bytes = per;
                h->b_level[lv].view = true;
            }
        }
    }
    if (ensure(h, h->b_sched, sizeof(Sched)) != LICHEE_OK) return LICHEE_ERR_CUDA;
    for (int i = 0; i < 2; i++) {
        if (ensure(h, h->b_top[i], sizeof(Cand) * LICHEE_MAX_SAVE) != LICHEE_OK ||
            ensure(h, h->b_tree[i], (size_t)LICHEE_MAX_SAVE * 2 * h->stride) != LICHEE_OK)
            return LICHEE_ERR_CUDA;
    }
    h->tables_on_device = true;
    return LICHEE_OK;
}

// Final selection: merge `m_hint` candidates (already bounded by the caller) into the ranked list.
int run_select(lichee_search *h, const Cand *cands, long long cand_cap, const unsigned *leaf_items, int external,
               const unsigned char *records, long long record_bytes, int words = 0) {
    const size_t sm = sizeof(Cand) * 3 * kSelTile;
    if (words == 0) words = h->stride >> 2;
    select_kernel<<<1, kSelThreads, sm, h->stream>>>(h->d_meta(), h->prm.num_save, cands, cand_cap,
                                                     h->d_top(h->flip), h->d_top(h->flip ^ 1),
                                                     h->d_tree(h->flip), h->d_tree(h->flip ^ 1), leaf_items,
                                                     words, h->L - 1, external, records, record_bytes);
    CK(cudaGetLastError());
    h->flip ^= 1;
    h->stats.num_launches++;
    h->stats.kernel_launches[LICHEE_K_SELECT]++;
    return LICHEE_OK;
}

// Radix pre-selection: on return the (at most kSelTile) survivors sit in b_cands[1] and m is their
// count; returns LICHEE_OK with used=false when ties at the threshold overflow the rank-merge tile
// (then the caller falls back to the tournament on the untouched list in b_cands[0]).
int run_radix_select(lichee_search *h, long long &m, bool &used) {
    used = false;
    if (ensure(h, h->b_radix, sizeof(RadixState)) != LICHEE_OK ||
        ensure(h, h->b_cands[1], sizeof(Cand) * (size_t)kSelTile) != LICHEE_OK)
        return LICHEE_ERR_CUDA;
    RadixState *st = static_cast<RadixState *>(h->b_radix.p);
    const int grid = (int)std::min<long long>((m + 255) / 256, 148 * 8);
    radix_init<<<1, 256, 0, h->stream>>>(st, (unsigned long long)h->prm.num_save);
    for (int pass = 0; pass < 16; pass++)
        radix_pass<<<grid, 256, 0, h->stream>>>(h->d_cands(0), m, st, pass, (unsigned)kSelTile);
    radix_compact<<<grid, 256, 0, h->stream>>>(h->d_cands(0), m, st, h->d_cands(1), (long long)kSelTile);
    CK(cudaGetLastError());
    h->stats.num_launches += 18;
    h->stats.kernel_launches[LICHEE_K_SELECT] += 18;
    unsigned long long kept = 0;
    CK(cudaMemcpyAsync(&kept, &st->kept, sizeof(kept), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (kept > (unsigned long long)kSelTile) return LICHEE_OK;       // massive ties: not usable
    m = (long long)kept;
    used = true;
    return LICHEE_OK;
}

// The 12-bit form for last-level chunks (leaf_kernel recorded the key range in the meta block).  Falls back to
// run_radix_select when equal scores overflow the tile.
// (the candidates' 32-byte records do not exist yet: radix12_compact builds the survivors'; `materialize` writes them
//  all, for the byte-wise fallback)
template <class F>
int run_radix12_select(lichee_search *h, long long &m, bool &used, long long seq_lo, long long seq_hi, int part, F materialize) {
    used = false;
    if (ensure(h, h->b_radix12, sizeof(Radix12State)) != LICHEE_OK ||
        ensure(h, h->b_cands[1], sizeof(Cand) * (size_t)kSelTile) != LICHEE_OK)
        return LICHEE_ERR_CUDA;
    Radix12State *st = static_cast<Radix12State *>(h->b_radix12.p);
    const int grid = (int)std::min<long long>((m + 1023) / 1024, 148 * 8);
    radix12_init<<<1, 256, 0, h->stream>>>(st, h->d_meta(), (unsigned long long)h->prm.num_save, (unsigned long long)seq_lo,
                                           (unsigned long long)seq_hi);
    h->stats.num_launches++;
    h->stats.kernel_launches[LICHEE_K_SELECT]++;
    struct { unsigned long long kept; int shift; unsigned done; } rd{};
    for (int round = 0; round < 2; round++) {           // 6 + 6 digits cover the 128 key bits; finished passes return at once
        const unsigned long long *k0 = static_cast<const unsigned long long *>(h->b_key0.p), *k1 = static_cast<const unsigned long long *>(h->b_key1.p);
        for (int p = 0; p < 6; p++) radix12_pass<<<grid, 256, 0, h->stream>>>(k0, k1, m, st, (unsigned)kSelTile);
        radix12_compact<<<grid, 256, 0, h->stream>>>(static_cast<const uint4 *>(h->b_c16.p), part, k0, k1, m, st, h->d_cands(1), (long long)kSelTile);
        CK(cudaGetLastError());
        h->stats.num_launches += 7;
        h->stats.kernel_launches[LICHEE_K_SELECT] += 7;
        CK(cudaMemcpyAsync(&rd.kept, &st->kept, sizeof(rd.kept), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(&rd.shift, &st->shift, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(&rd.done, &st->done, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (rd.done || rd.shift < 0) break;
    }
    if (getenv("LICHEE_DEBUG")) {
        Radix12State hs;
        cudaMemcpy(&hs, st, offsetof(Radix12State, hist), cudaMemcpyDeviceToHost);
        fprintf(stderr, "radix12: m=%lld done=%u word=%d shift=%d kept=%llu below=%llu k=%llu thresh=%016llx/%llu\n", m, hs.done, hs.word,
                hs.shift, hs.kept, hs.below, hs.k, hs.thresh[0], hs.thresh[1]);
    }
    if (!rd.done || rd.kept > (unsigned long long)kSelTile) {
        if (materialize() != LICHEE_OK) return LICHEE_ERR_CUDA;
        return run_radix_select(h, m, used);
    }
    m = (long long)rd.kept;
    used = true;
    return LICHEE_OK;
}

// Tournament: shrink a long candidate list (m entries in b_cands[0]) to <= kSelPerBlock entries.
// Returns the buffer index holding the survivors and their count.
int run_tournament(lichee_search *h, long long &m, int &src) {
    src = 0;
    const size_t sm = sizeof(Cand) * 2 * kSelTile;
    while (m > kSelTile) {
        const long long blocks = (m + kSelPerBlock - 1) / kSelPerBlock;
        if (ensure(h, h->b_cands[src ^ 1], sizeof(Cand) * (size_t)blocks * kSelTile) != LICHEE_OK) return LICHEE_ERR_CUDA;
        select_reduce_kernel<<<(int)blocks, kSelThreads, sm, h->stream>>>(h->d_cands(src), m, h->d_cands(src ^ 1));
        CK(cudaGetLastError());
        h->stats.num_launches++;
        h->stats.kernel_launches[LICHEE_K_SELECT]++;
        m = blocks * kSelTile;
        src ^= 1;
    }
    return LICHEE_OK;
}

int fetch_top(lichee_search *h) {
    CK(cudaMemcpyAsync(h->h_meta, h->d_meta(), sizeof(TopMeta), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const unsigned count = h->h_meta->count;
    h->h_top.resize(count);
    const size_t tb = h->ref_result ? 2 * (size_t)h->stride : (size_t)h->stride;
    h->h_tree.assign((size_t)count * tb, 0xff);
    if (count) {
        CK(cudaMemcpyAsync(h->h_top.data(), h->d_top(h->flip), sizeof(Cand) * count, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(h->h_tree.data(), h->d_tree(h->flip), (size_t)count * tb, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    h->stats.num_saved = count;
    return LICHEE_OK;
}

// Adds the kernel times of one finished chunk (event set `set`) to the statistics.
// Boundaries: 0 start | 1 after check (or fused leaf) | 2 after scan | 3 after emit (leaf: after the
// host read of the candidate count) | 4 after seq fix-up + select.
int harvest_events(lichee_search *h, int set) {
    const int nb = h->evn[set];
    if (nb < 3) return LICHEE_OK;
    CK(cudaEventSynchronize(h->evk[set][nb - 1]));
    float ms = 0;
    // a last-level chunk runs the fused check+score kernel where the others run the check kernel
    CK(cudaEventElapsedTime(&ms, h->evk[set][0], h->evk[set][1]));
    h->stats.kernel_ms[h->ev_leaf[set] ? LICHEE_K_LEAF : LICHEE_K_CHECK] += ms;
    CK(cudaEventElapsedTime(&ms, h->evk[set][1], h->evk[set][2])); h->stats.kernel_ms[LICHEE_K_SCAN] += ms;
    if (nb > 3 && !h->ev_leaf[set]) {
        CK(cudaEventElapsedTime(&ms, h->evk[set][2], h->evk[set][3]));
        h->stats.kernel_ms[LICHEE_K_EMIT] += ms;
    }
    if (nb > 4) { CK(cudaEventElapsedTime(&ms, h->evk[set][3], h->evk[set][4])); h->stats.kernel_ms[LICHEE_K_SELECT] += ms; }
    h->evn[set] = 0;
    return LICHEE_OK;
}

// Searches slice R of P of the canonical order.  keep = continue an earlier run_search on this handle:
// the ranked list, its threshold and the statistics carry over (dynamic work stealing: a rank claims
// slice after slice from a shared counter).
template <int SPL>
int run_search(lichee_search *h, int R, int P, bool keep) {
    const int L = h->L;
    if (upload_tables(h) != LICHEE_OK) return LICHEE_ERR_CUDA;
    lichee_search_stats &st = h->stats;
    const lichee_search_stats before = st;
    st = lichee_search_stats{};
    st.num_grow_calls = 1;
    if (!keep) {
        init_meta<<<1, 1, 0, h->stream>>>(h->d_meta());
        h->flip = 0;
    }
    for (int lv = 0; lv < L; lv++) h->head[lv] = h->tail[lv] = 0;
    auto finish = [&]() {
        if (keep) {                       // statistics accumulate over the slices of one search
            st.num_trees += before.num_trees; st.num_grow_calls += before.num_grow_calls - 1;
            st.num_checks += before.num_checks; st.num_scored += before.num_scored;
            st.num_evaluated += before.num_evaluated; st.num_bounded_items += before.num_bounded_items; st.num_leaf_items += before.num_leaf_items;
            st.num_launches += before.num_launches; st.frontier_bytes += before.frontier_bytes;
            st.kernel_ms_total += before.kernel_ms_total; st.status |= before.status;
            for (int i = 0; i < LICHEE_NUM_KERNELS; i++) {
                st.kernel_ms[i] += before.kernel_ms[i]; st.kernel_launches[i] += before.kernel_launches[i];
                st.kernel_bytes[i] += before.kernel_bytes[i];
            }
        }
        return fetch_top(h);
    };
    if (h->no_trees || L == 0) return finish();

    if (ensure(h, h->b_level[0], (size_t)h->stride) != LICHEE_OK) return LICHEE_ERR_CUDA;
    CK(cudaMemsetAsync(h->level(0), 0xff, h->stride, h->stream));
    h->tail[0] = 1;
    if (L == 1) {
        // two nodes: the only item is the empty one, its pass mask is the static mask of the last node
        if (ensure(h, h->b_lmask, sizeof(uint4)) != LICHEE_OK || ensure(h, h->b_lcnt, 4) != LICHEE_OK || ensure(h, h->b_lbound, 8) != LICHEE_OK) return LICHEE_ERR_CUDA;
        CK(cudaMemsetAsync(h->b_lbound.p, 0, 8, h->stream));
        const uint4 m = h->sok_last;
        const unsigned c = (unsigned)(__builtin_popcount(m.x) + __builtin_popcount(m.y) + __builtin_popcount(m.z) + __builtin_popcount(m.w));
        CK(cudaMemcpyAsync(h->b_lmask.p, &m, sizeof(m), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->b_lcnt.p, &c, sizeof(c), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    CK(cudaEventRecord(h->ev0, h->stream));

    // fan[lv]: observed mean number of surviving children per item placed at level lv; yield[lv]
    // (recomputed from fan) estimates the complete trees below one item, used to size chunks so
    // that a binding tree cap does not pay for subtrees it will never reach.
    std::vector<double> fan(L, 0.0);
    std::vector<bool> seen(L, false);
    for (int lv = 0; lv < L; lv++) fan[lv] = (double)h->cands[lv].size();
    long long seq_base = 0;
    P = std::max(1, P);
    bool sliced = P <= 1;
    float ms = 0;
    int es = 0;
    h->evn[0] = h->evn[1] = 0;

    // A handful of items: the single-block scheduler walks the levels on its own until the last two levels are
    // reached, a chunk gets wide, or the search ends -- one round trip for all of it.  bfs_width > 0: the
    // breadth-first prologue of a sliced search (whole levels until one is that wide).
    auto run_narrow = [&](long long bfs_width) -> int {
        Sched *hs = h->h_sched;
        for (int l = 0; l < L; l++) {
            hs->head[l] = h->head[l]; hs->tail[l] = h->tail[l]; hs->ptr[l] = h->level(l);
            hs->fan[l] = fan[l]; hs->seen[l] = seen[l] ? 1 : 0;
        }
        hs->trees_left = h->prm.max_num_trees - seq_base;
        hs->grow_calls = st.num_grow_calls; hs->max_grow_calls = h->prm.max_num_grow_calls;
        hs->chunk_max = h->chunk_max; hs->cap = h->cap; hs->bfs_width = bfs_width;
        Sched *ds = static_cast<Sched *>(h->b_sched.p);
        es ^= 1;
        if (harvest_events(h, es) != LICHEE_OK) return LICHEE_ERR_CUDA;
        h->ev_leaf[es] = false;
        CK(cudaEventRecord(h->evk[es][0], h->stream));
        CK(cudaMemcpyAsync(ds, hs, sizeof(Sched), cudaMemcpyHostToDevice, h->stream));
        descend_kernel<SPL><<<1, kNarrowThreads, h->smem + (size_t)L * kMaxN, h->stream>>>(h->net, ds);
        CK(cudaGetLastError());
        CK(cudaEventRecord(h->evk[es][1], h->stream));
        CK(cudaEventRecord(h->evk[es][2], h->stream));
        h->evn[es] = 3;
        CK(cudaMemcpyAsync(hs, ds, sizeof(Sched), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int l = 0; l < L; l++) {
            h->head[l] = hs->head[l]; h->tail[l] = hs->tail[l]; fan[l] = hs->fan[l]; seen[l] = hs->seen[l] != 0;
        }
        st.num_checks += hs->checks;
        st.num_grow_calls = hs->grow_calls;
        st.frontier_bytes += hs->frontier_bytes;
        st.kernel_bytes[LICHEE_K_CHECK] += hs->check_bytes;
        st.kernel_bytes[LICHEE_K_EMIT] += hs->emit_bytes;
        st.num_launches++;
        st.kernel_launches[LICHEE_K_CHECK]++;
        if (getenv("LICHEE_DEBUG")) fprintf(stderr, "narrow: %lld chunks, stop %d at level %d\n", hs->chunks, hs->stop, hs->stop_level);
        return LICHEE_OK;
    };

    bool list_full = false;                  // the ranked list held num_save trees at the last read-back
    // what emit2_kernel reported about the current fill of the last level: items that are not bounded, valid
    // trees below all of them (fill_items = 0: unknown, e.g. after the level was cut into slices)
    long long fill_items = 0, fill_live = 0, fill_trees = 0, fill_parents = 0;
    bool fill_pending = false;               // emit2's counters are still on their way to the host
    const unsigned *fill_ref = nullptr;      // an item of the level above from the chunk the current fill came from (leaf_kernel's reference base)
    // check2_kernel of the NEXT level L-2 chunk is launched on the second stream as soon as the children of the current
    // one have been emitted: it runs beside the last level of the current fill, whose selection kernels are single
    // blocks and whose read-backs leave the GPU idle.  The two sets of result arrays alternate between consecutive
    // chunks; its flags are set against the threshold of the moment and refreshed by emit2_kernel (the bound comes along).
    struct { bool valid = false; long long head = 0, B = 0; int set = 0; } pf;
    int c2set = 0;                           // result arrays of the level L-2 chunk in hand
    auto c2buf = [&](int set, int which) -> DevBuf & {
        if (set) return h->b_pf[which];
        return which == 0 ? h->b_mask : which == 1 ? h->b_cnt : which == 2 ? h->b_mask2 : which == 3 ? h->b_mask3 : which == 4 ? h->b_flag : h->b_bound;
    };
    auto launch_check2 = [&](cudaStream_t strm, int set, const unsigned *c2items, long long nb) -> int {
        static const size_t unit[6] = {sizeof(uint4), 4, sizeof(uint4), sizeof(uint4), 4, 8};
        for (int w = 0; w < 6; w++)
            if (ensure(h, c2buf(set, w), unit[w] * (size_t)nb) != LICHEE_OK) return LICHEE_ERR_CUDA;
        // one run per warp, sized so that the grid is one full wave of two blocks per SM (a run starts from the
        // block's reference state: runs shorter than 8 items do not pay)
        const long long slots = 148ll * 2 * kWarpsPerBlock;
        const int run_len = (int)std::min<long long>(64, std::max<long long>(8, (nb + slots - 1) / slots));
        const int grid = (int)std::min<long long>(148 * 2, (nb + (long long)run_len * kWarpsPerBlock - 1) / ((long long)run_len * kWarpsPerBlock));
        check2_kernel<SPL><<<grid, kThreads, h->smem, strm>>>(
            h->net, c2items, nb, static_cast<uint4 *>(c2buf(set, 0).p), static_cast<unsigned *>(c2buf(set, 1).p),
            static_cast<uint4 *>(c2buf(set, 2).p), static_cast<uint4 *>(c2buf(set, 3).p), static_cast<unsigned *>(c2buf(set, 4).p),
            static_cast<double *>(c2buf(set, 5).p), h->d_meta(), h->prm.num_save, run_len);
        CK(cudaGetLastError());
        st.num_launches++;
        st.kernel_launches[LICHEE_K_CHECK]++;
        st.kernel_bytes[LICHEE_K_CHECK] += nb * (long long)(h->stride + 64);
        return LICHEE_OK;
    };
    // chunk of a level below the last one: what is pending, cut to what the remaining tree budget can use and to
    // what the next level is expected to hold
    auto chunk_rule = [&](int lv, long long pending, bool leaf) -> long long {
        long long B = std::min(pending, h->chunk_max);
        if (sliced) {
            // trees still wanted / trees expected below one item of this level
            double yield = 1.0;
            for (int l = lv; l < L; l++)   // unseen levels: assume every candidate passes, so the first descent stays narrow
                yield = std::min(1e30, yield * std::max(seen[l] ? fan[l] : (double)h->cands[l].size(), 1e-3));
            const double want = 1.25 * (double)(h->prm.max_num_trees - seq_base) / yield + 8.0;
            if ((double)B > want) B = (long long)want;
        }
        if (!leaf) {
            const long long guess = (long long)(0.8 * (double)h->cap / std::max(1.0, fan[lv]));
            B = std::min(B, std::max(1024ll, guess));
        }
        return B < 1 ? 1 : B;
    };
    for (;;) {
        int lv = -1;
        if (!sliced) {
            // breadth-first prologue shared by every part: take whole levels until one is wide
            // enough to be cut into P contiguous slices of the canonical order.
            for (int l = 0; l < L; l++) if (h->tail[l] > h->head[l]) { lv = l; break; }
            if (lv < 0) break;
            const long long width = h->tail[lv] - h->head[lv];
            const bool wide = width >= 64ll * P;
            const bool last = lv == L - 1;
            const bool risky = lv + 1 < L && (double)width * (double)h->cands[lv].size() > 0.5 * (double)h->cap;
            if (wide || last || risky) {
                const long long a = h->head[lv] + width * R / P, b = h->head[lv] + width * (R + 1) / P;
                h->head[lv] = a;
                h->tail[lv] = b;
                sliced = true;
                if (a == b) break;
            } else if (h->narrow && lv < L - 2 && width <= kNarrowMax &&
                       width * (long long)h->cands[lv].size() <= std::min<long long>(h->cap, kNarrowCap)) {
                // narrow levels of the prologue run on the device as well (every rank walks the same levels)
                if (run_narrow(64ll * P) != LICHEE_OK) return LICHEE_ERR_CUDA;
                if (st.num_grow_calls >= h->prm.max_num_grow_calls) { st.status |= LICHEE_STATUS_GROW_CAP; break; }
                continue;
            }
        } else {
            for (int l = L - 1; l >= 0; l--) if (h->tail[l] > h->head[l]) { lv = l; break; }
            if (lv < 0) break;
        }
        const long long pending = h->tail[lv] - h->head[lv];
        const bool leaf = lv == L - 1;
        const long long kcand = (long long)h->cands[lv].size();
        long long B = chunk_rule(lv, pending, leaf);
        if (leaf && fill_pending) {
            CK(cudaStreamSynchronize(h->stream));
            fill_pending = false;
            fill_live = (long long)h->h_scan->live;
            fill_trees = (long long)h->h_scan->trees;
            // emit2's algorithmic bytes: masks, flag and offset of every parent, a count word per child, and
            // item + mask for the children that are not bounded
            st.kernel_bytes[LICHEE_K_EMIT] += fill_parents * 64ll + fill_items * 4ll + fill_live * (long long)(h->stride + 24);
        }
        if (leaf && fill_items > 0 && fill_live == 0 && B == pending && pending == fill_items) {
            // no item of this fill can reach the ranked list: its trees are counted, nothing is launched
            fill_items = 0;
            st.num_checks += B * kcand;
            st.num_bounded_items += B;
            st.num_leaf_items += B;
            st.num_grow_calls += fill_trees;
            const long long before = seq_base;
            seq_base += fill_trees;
            st.num_scored += std::min(seq_base, (long long)h->prm.max_num_trees) - std::min(before, (long long)h->prm.max_num_trees);
            fan[lv] = seen[lv] ? 0.5 * fan[lv] + 0.5 * (double)fill_trees / (double)B : (double)fill_trees / (double)B;
            seen[lv] = true;
            h->head[lv] = h->tail[lv] = 0;
            fill_ref = nullptr;
            if (getenv("LICHEE_DEBUG")) fprintf(stderr, "leaf fill B=%lld counted only: %lld trees\n", B, fill_trees);
            if (seq_base >= h->prm.max_num_trees) { st.status |= LICHEE_STATUS_TREE_CAP; break; }
            if (st.num_grow_calls >= h->prm.max_num_grow_calls) { st.status |= LICHEE_STATUS_GROW_CAP; break; }
            continue;
        }
        if (leaf) fill_items = 0;                 // a chunk of the fill goes through the kernels: totals no longer apply
        if (sliced && lv < L - 2 && h->narrow && B <= kNarrowMax) {
            if (run_narrow(0) != LICHEE_OK) return LICHEE_ERR_CUDA;
            if (st.num_grow_calls >= h->prm.max_num_grow_calls) { st.status |= LICHEE_STATUS_GROW_CAP; break; }
            continue;
        }
        const unsigned *items = h->level(lv) + (size_t)h->head[lv] * (h->stride >> 2);
        const bool dual = lv == L - 2;                       // the level whose children are the last-level items
        // the chunk may have been checked already, beside the last level of the fill before it
        bool prefetched = false;
        if (pf.valid && !leaf) {
            if (dual && pf.head == h->head[lv]) { prefetched = true; B = pf.B; c2set = pf.set; }
            else CK(cudaStreamSynchronize(h->stream2));      // (not reached: the chunk after a fill is the next one of its level)
            pf.valid = false;
        } else if (dual) {
            c2set = 0;
        }
        if ((!dual && !leaf && (ensure(h, h->b_mask, sizeof(uint4) * (size_t)B) != LICHEE_OK || ensure(h, h->b_cnt, 4 * (size_t)B) != LICHEE_OK)) ||
            ensure(h, h->b_off, 4 * (size_t)B) != LICHEE_OK ||
            ensure(h, h->b_tile, 8 * (size_t)(B / kScanTile + 2)) != LICHEE_OK)
            return LICHEE_ERR_CUDA;
        es ^= 1;
        if (harvest_events(h, es) != LICHEE_OK) return LICHEE_ERR_CUDA;   // the chunk before last
        h->ev_leaf[es] = leaf;
        CK(cudaEventRecord(h->evk[es][0], h->stream));
        const long long cand_cap = h->cand_cap;
        const unsigned *lcnt = static_cast<const unsigned *>(h->b_lcnt.p) + (leaf ? h->head[lv] : 0);
        const uint4 *lmask = static_cast<const uint4 *>(h->b_lmask.p) + (leaf ? h->head[lv] : 0);
        if (dual) {
            if (prefetched) {
                // its time on the second stream goes to the check column; the main stream only waits for its end
                CK(cudaStreamWaitEvent(h->stream, h->evp[2], 0));
                CK(cudaEventSynchronize(h->evp[2]));
                float pms = 0;
                CK(cudaEventElapsedTime(&pms, h->evp[0], h->evp[1]));
                st.kernel_ms[LICHEE_K_CHECK] += pms;
                st.num_launches--;                           // (counted again below with the chunk)
            } else if (launch_check2(h->stream, c2set, items, B) != LICHEE_OK) {
                return LICHEE_ERR_CUDA;
            } else {
                st.num_launches--;
            }
        } else if (!leaf) {
            if (launch_check<SPL>(h, lv, items, B) != LICHEE_OK) return LICHEE_ERR_CUDA;
            st.kernel_launches[LICHEE_K_CHECK]++;
            st.kernel_bytes[LICHEE_K_CHECK] += B * (long long)(h->stride + 20);
        } else {
            // last level: scoring of the items that can still reach the ranked list.  Every unbounded tree of
            // the chunk may become a candidate, so the chunk is the longest prefix of the pending items whose
            // unbounded trees fit the candidate buffer: scanned on the device, read by the kernel from
            // device memory and by the host afterwards (no round trip in between, no overflow possible).
            if (ensure(h, h->b_c16, sizeof(uint4) * (size_t)cand_cap) != LICHEE_OK) return LICHEE_ERR_CUDA;
            // (slots are reserved in batches: up to 31 of every 256 and one batch per warp may stay empty)
            // (while the ranked list is still filling up every tree is a candidate, so the first chunk stays small,
            //  yet large enough for a whole search under the reference's MAX_NUM_TREES = 100000)
            long long cand_limit = std::max<long long>(128, (cand_cap - (long long)kCandBatch * 148 * 3 * kWarpsPerBlock) / 5 * 4);
            if (!list_full) cand_limit = std::min<long long>(cand_limit, 1 << 17);
            const double *lb_chunk = h->net.prune && !h->no_prefilter ? static_cast<const double *>(h->b_lbound.p) + h->head[lv] : nullptr;
            if (launch_scan(h, lcnt, B, (unsigned long long)cand_limit, 1, lb_chunk) != LICHEE_OK) return LICHEE_ERR_CUDA;
            st.kernel_bytes[LICHEE_K_SCAN] += B * 12ll;
            const int grid = grid_for(B);
            // runs of items are claimed dynamically by the warps, in 32-item blocks (a bounded item costs a
            // few instructions, an evaluated one hundreds)
            int run_len = 32;    // (measured: runs of 256 items left a tail of ~1 ms per step; a run of bounded items costs a few instructions)
            if (h->no_prefilter) run_len = -run_len;
            leaf_kernel<SPL><<<grid, kThreads, h->smem_leaf, h->stream>>>(
                h->net, items, h->d_scan(), h->head[lv], lcnt, static_cast<const double *>(h->b_lbound.p) + h->head[lv], h->d_meta(), h->prm.num_save,
                static_cast<uint4 *>(h->b_c16.p), cand_cap, run_len, fill_ref);
            CK(cudaGetLastError());
            st.kernel_launches[LICHEE_K_LEAF]++;
        }
        CK(cudaEventRecord(h->evk[es][1], h->stream));
        if (!leaf) {
            const unsigned *cnt_in = dual ? static_cast<const unsigned *>(c2buf(c2set, 1).p) : h->d_cnt();
            if (launch_scan(h, cnt_in, B, (unsigned long long)h->cap) != LICHEE_OK) return LICHEE_ERR_CUDA;
            st.kernel_bytes[LICHEE_K_SCAN] += B * 12ll;
        }
        CK(cudaEventRecord(h->evk[es][2], h->stream));
        h->evn[es] = 3;
        st.num_launches += 1;
        long long leaf_total = 0;
        if (leaf) {
            // the kernel counted its valid trees; the scan that numbers them is only needed when
            // some of them beat the threshold (most chunks: none)
            CK(cudaMemcpyAsync(h->h_meta, h->d_meta(), sizeof(TopMeta), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            B = (long long)(h->h_scan->fit >> 32);           // the prefix the kernel took
            if (B < 1 || h->h_meta->n_cand > (unsigned long long)cand_cap) {
                set_err("last-level chunk accounting broken: %lld items, %llu candidates for %lld slots", B,
                        (unsigned long long)h->h_meta->n_cand, cand_cap);
                return LICHEE_ERR_STATE;
            }
            long long m = (long long)h->h_meta->n_cand;
            leaf_total = (long long)h->h_meta->n_valid;
            list_full = h->h_meta->count >= (unsigned)h->prm.num_save || m >= (long long)h->prm.num_save;
            if (getenv("LICHEE_DEBUG")) fprintf(stderr, "leaf chunk B=%lld cand=%lld total=%lld bounded=%llu tau=%.17g count=%u\n", B, m, leaf_total, h->h_meta->n_bounded, h->h_meta->tau, h->h_meta->count);
            CK(cudaEventRecord(h->evk[es][3], h->stream));
            // algorithmic bytes of the launch: the count word of every item, the item itself where it was
            // evaluated, the candidate records written
            st.kernel_bytes[LICHEE_K_LEAF] += B * 4ll + (B - (long long)h->h_meta->n_bounded) * (long long)h->stride + m * 16ll;
            st.num_evaluated += (long long)h->h_meta->n_eval;
            st.num_bounded_items += (long long)h->h_meta->n_bounded;
            st.num_leaf_items += B;
            if (m > 0) {
                if (launch_scan(h, lcnt, B, ~0ull >> 1) != LICHEE_OK) return LICHEE_ERR_CUDA;
                st.kernel_bytes[LICHEE_K_SCAN] += B * 12ll;
            }
            if (m > 0) {
                if (ensure(h, h->b_key0, 8 * (size_t)m) != LICHEE_OK || ensure(h, h->b_key1, 8 * (size_t)m) != LICHEE_OK ||
                    ensure(h, h->b_cands[0], sizeof(Cand) * (size_t)m) != LICHEE_OK)
                    return LICHEE_ERR_CUDA;
                // a long list gets keys only; the 32-byte records are built for the survivors of the radix selection
                const long long m_all = m;
                const long long item0 = h->head[lv];
                auto fixup = [&](int write_records) -> int {
                    seq_fixup_kernel<<<(int)((m_all + 255) / 256), 256, 0, h->stream>>>(static_cast<const uint4 *>(h->b_c16.p), h->d_cands(0), m_all, item0, lmask,
                                                                                 h->d_off(), seq_base, h->prm.max_num_trees,
                                                                                 static_cast<unsigned long long *>(h->b_key0.p),
                                                                                 static_cast<unsigned long long *>(h->b_key1.p), h->d_meta(),
                                                                                 h->prm.num_save, R, write_records);
                    CK(cudaGetLastError());
                    st.num_launches++;
                    return LICHEE_OK;
                };
                if (fixup(m <= kSelTile ? 1 : 0) != LICHEE_OK) return LICHEE_ERR_CUDA;
                int src = 0;
                bool radix = false;
                if (m > kSelTile && run_radix12_select(h, m, radix, seq_base, seq_base + std::max<long long>(leaf_total, 1) - 1, R, [&]() { return fixup(1); }) != LICHEE_OK)
                    return LICHEE_ERR_CUDA;
                if (radix) src = 1;
                else if (run_tournament(h, m, src) != LICHEE_OK) return LICHEE_ERR_CUDA;
                if (run_select(h, h->d_cands(src), m, h->level(lv), 0, nullptr, 0) != LICHEE_OK) return LICHEE_ERR_CUDA;
                st.kernel_bytes[LICHEE_K_SELECT] += (long long)sizeof(Cand) * (long long)h->h_meta->n_cand;
            } else {
                reset_n_cand<<<1, 1, 0, h->stream>>>(h->d_meta());
            }
            CK(cudaEventRecord(h->evk[es][4], h->stream));
            h->evn[es] = 5;
            const long long total = leaf_total;
            st.num_checks += B * kcand;
            st.num_grow_calls += total;
            st.frontier_bytes += B * (long long)h->stride;
            const long long before = seq_base;
            seq_base += total;
            st.num_scored += std::min(seq_base, (long long)h->prm.max_num_trees) - std::min(before, (long long)h->prm.max_num_trees);
            h->head[lv] += B;
            fan[lv] = seen[lv] ? 0.5 * fan[lv] + 0.5 * (double)total / (double)B : (double)total / (double)B;
            seen[lv] = true;
        } else {
            CK(cudaStreamSynchronize(h->stream));
            const long long n_fit = (long long)(h->h_scan->fit >> 32);
            const long long total_fit = (long long)(h->h_scan->fit & 0xffffffffull);
            if (getenv("LICHEE_DEBUG") && lv >= L - 3) fprintf(stderr, "level %d chunk B=%lld fit=%lld children=%lld prefetched=%d fan=%.2f pending=%lld\n", lv, B, n_fit, total_fit, prefetched ? 1 : 0, fan[lv], pending);
            if (n_fit == 0) {
                set_err("level_capacity %lld too small for one item's children", h->cap);
                return LICHEE_ERR_INVALID;
            }
            if (total_fit > 0) {
                if (ensure(h, h->b_level[lv + 1], (size_t)total_fit * h->stride) != LICHEE_OK) return LICHEE_ERR_CUDA;
                if (dual) {
                    if (ensure(h, h->b_lmask, sizeof(uint4) * (size_t)total_fit) != LICHEE_OK || ensure(h, h->b_lcnt, 4 * (size_t)total_fit) != LICHEE_OK ||
                        ensure(h, h->b_lbound, 8 * (size_t)total_fit) != LICHEE_OK)
                        return LICHEE_ERR_CUDA;
                    CK(cudaMemsetAsync(&h->d_scan()->live, 0, 2 * sizeof(unsigned long long), h->stream));
                    emit2_kernel<<<grid_for(n_fit), kThreads, 0, h->stream>>>(
                        h->net, items, n_fit, static_cast<const uint4 *>(c2buf(c2set, 0).p), h->d_off(), static_cast<const uint4 *>(c2buf(c2set, 2).p),
                        static_cast<const uint4 *>(c2buf(c2set, 3).p), static_cast<const unsigned *>(c2buf(c2set, 4).p),
                        static_cast<const double *>(c2buf(c2set, 5).p), h->level(lv + 1),
                        static_cast<uint4 *>(h->b_lmask.p), static_cast<unsigned *>(h->b_lcnt.p), static_cast<double *>(h->b_lbound.p), h->d_scan(),
                        h->d_meta(), h->prm.num_save);
                    CK(cudaMemcpyAsync(h->h_scan, h->d_scan(), sizeof(ScanOut), cudaMemcpyDeviceToHost, h->stream));
                    fill_pending = true;                      // read at the next last-level chunk
                    fill_ref = items;
                    fill_items = total_fit;
                    fill_parents = n_fit;
                } else {
                    emit_kernel<<<grid_for(n_fit), kThreads, 0, h->stream>>>(h->net, lv, items, n_fit, h->d_mask(),
                                                                            h->d_off(), h->level(lv + 1));
                }
                CK(cudaGetLastError());
                CK(cudaEventRecord(h->evk[es][3], h->stream));
                h->evn[es] = 4;
                st.num_launches++;
                st.kernel_launches[LICHEE_K_EMIT]++;
            }
            h->head[lv + 1] = 0;
            h->tail[lv + 1] = total_fit;
            h->head[lv] += n_fit;
            st.num_checks += n_fit * kcand;
            st.num_grow_calls += total_fit;
            st.frontier_bytes += (n_fit + total_fit) * (long long)h->stride;
            if (!dual) st.kernel_bytes[LICHEE_K_EMIT] += n_fit * (long long)(h->stride + 20) + total_fit * (long long)h->stride;
            const double f = (double)total_fit / (double)n_fit;
            fan[lv] = seen[lv] ? 0.5 * fan[lv] + 0.5 * f : f;
            seen[lv] = true;
            if (dual && sliced && !h->no_prefetch && total_fit > 0 && h->tail[lv] > h->head[lv]) {
                // the next chunk of this level, checked beside the fill that was just emitted
                const long long nb = chunk_rule(lv, h->tail[lv] - h->head[lv], false);
                pf.valid = true; pf.head = h->head[lv]; pf.B = nb; pf.set = c2set ^ 1;
                CK(cudaEventRecord(h->evp[0], h->stream2));
                if (launch_check2(h->stream2, pf.set, h->level(lv) + (size_t)pf.head * (h->stride >> 2), nb) != LICHEE_OK) return LICHEE_ERR_CUDA;
                CK(cudaEventRecord(h->evp[1], h->stream2));
                CK(cudaEventRecord(h->evp[2], h->stream2));
            }
        }
        if (h->head[lv] == h->tail[lv]) {
            h->head[lv] = h->tail[lv] = 0;
            if (leaf) fill_ref = nullptr;
        }
        if (seq_base >= h->prm.max_num_trees) { st.status |= LICHEE_STATUS_TREE_CAP; break; }
        if (st.num_grow_calls >= h->prm.max_num_grow_calls) { st.status |= LICHEE_STATUS_GROW_CAP; break; }
    }
    if (pf.valid) CK(cudaStreamSynchronize(h->stream2));     // a chunk checked ahead that a cap made unnecessary
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (harvest_events(h, 0) != LICHEE_OK || harvest_events(h, 1) != LICHEE_OK) return LICHEE_ERR_CUDA;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    st.kernel_ms_total = ms;
    st.num_trees = std::min(seq_base, (long long)h->prm.max_num_trees);
    return finish();
}

// Reference-order search: replay PHYNetwork.grow on the device, score the discovered trees in the
// reference's accumulation order, rank by (score, discovery index) -- Collections.sort is stable.
// `completed` is false when the replay ran out of its budget of grow() calls before the reference's own
// search would have ended; the caller then falls back to the canonical search.
template <int SPL>
int run_reference_order(lichee_search *h, bool &completed) {
    completed = false;
    lichee_search_stats &st = h->stats;
    st = lichee_search_stats{};
    const long long budget = h->prm.ref_order_budget > 0 ? h->prm.ref_order_budget : LICHEE_DEFAULT_REF_ORDER_BUDGET;
    const replay::Header *h0 = reinterpret_cast<const replay::Header *>(h->rblob0);
    const long long rb = 24 + 2ll * h->stride;
    h->flip = 0;
    h->ref_result = true;
    CK(cudaEventRecord(h->ev0, h->stream));
    int status = h0->status;
    long long num_grow = 0, num_checks = 0, num_trees = 0;
    bool fused = false;
    if (status != replay::S_RUNNING) {                       // the root has no edge: no grow() call at all
        init_meta<<<1, 1, 0, h->stream>>>(h->d_meta());
        CK(cudaGetLastError());
    } else {
        ReplayArgs a;
        a.lay = h->rlay;
        a.blob_src = h->rblob0;                                // pinned initial state, read by the kernel itself
        a.blob_dst = static_cast<unsigned char *>(h->b_rblob.p);
        a.blob_in_smem = h->rblob_in_smem ? 1 : 0;
        a.vaf = h->p_vaf; a.S = h->S; a.L = h->L; a.stride = h->stride; a.margin = h->prm.vaf_error_margin;   // (pinned: staged by the kernel)
        a.record_bytes = rb;
        a.step_quota = 1ll << 22; a.grow_limit = budget;
        a.out = h->h_rout;                                     // pinned: read after the synchronize, no copy
        a.first = 1; a.fuse = 1; a.num_save = h->prm.num_save;
        a.meta = h->d_meta(); a.desc_rows = h->p_desc; a.result = h->h_result;
        const size_t smem = (size_t)h->n * h->SP * sizeof(double) + (h->rblob_in_smem ? h->rlay.total : 0);
        const bool small = h->n <= 32;
        // record buffer: small for the first launch (most real searches end inside it), full size afterwards
        int rec_cap = 4096;
        for (;;) {
            if (ensure(h, h->b_rrec, (size_t)rb * rec_cap) != LICHEE_OK) return LICHEE_ERR_CUDA;
            unsigned char *rec = static_cast<unsigned char *>(h->b_rrec.p);
            a.records = rec; a.rec_cap = rec_cap;
            CK(cudaEventRecord(h->evk[0][0], h->stream));
            if (small && h->rblob_in_smem) replay_kernel<SPL, 1, true><<<1, kReplayThreads, smem, h->stream>>>(a);
            else if (small) replay_kernel<SPL, 1, false><<<1, kReplayThreads, smem, h->stream>>>(a);
            else if (h->rblob_in_smem) replay_kernel<SPL, 4, true><<<1, kReplayThreads, smem, h->stream>>>(a);
            else replay_kernel<SPL, 4, false><<<1, kReplayThreads, smem, h->stream>>>(a);
            CK(cudaGetLastError());
            CK(cudaEventRecord(h->evk[0][1], h->stream));
            CK(cudaStreamSynchronize(h->stream));
            st.num_launches++;
            st.kernel_launches[LICHEE_K_CHECK]++;
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, h->evk[0][0], h->evk[0][1]));
            st.kernel_ms[LICHEE_K_CHECK] += ms;
            status = h->h_rout->status;
            num_grow = h->h_rout->num_grow; num_checks = h->h_rout->num_checks; num_trees = h->h_rout->num_trees;
            const int m = h->h_rout->rec_count;
            st.kernel_bytes[LICHEE_K_CHECK] += (long long)m * rb;
            if (h->h_rout->fused) { fused = true; break; }
            a.first = 0; a.fuse = 0; a.blob_src = a.blob_dst;
            if (upload_tables(h) != LICHEE_OK) return LICHEE_ERR_CUDA;     // (scoring and merging use the device-resident tables and lists)
            if (m > 0) {
                // score the batch (one warp per tree) and merge it into the ranked list
                if (ensure(h, h->b_cands[0], sizeof(Cand) * (size_t)std::max(m, kSelTile)) != LICHEE_OK) return LICHEE_ERR_CUDA;
                const int grid = (int)std::min<long long>(148 * 2, (m + kWarpsPerBlock - 1) / kWarpsPerBlock);
                CK(cudaEventRecord(h->evk[0][2], h->stream));
                score_ref_kernel<SPL><<<grid, kThreads, h->smem, h->stream>>>(h->net, rec, rb, m, h->d_cands(0), h->d_meta());
                CK(cudaGetLastError());
                CK(cudaEventRecord(h->evk[0][3], h->stream));
                st.num_launches++;
                st.kernel_launches[LICHEE_K_LEAF]++;
                st.kernel_bytes[LICHEE_K_LEAF] += (long long)m * (rb + (long long)sizeof(Cand));
                long long mm = m;
                int src = 0;
                bool radix = false;
                if (mm > kSelTile && run_radix_select(h, mm, radix) != LICHEE_OK) return LICHEE_ERR_CUDA;
                if (radix) src = 1;
                else if (run_tournament(h, mm, src) != LICHEE_OK) return LICHEE_ERR_CUDA;
                if (run_select(h, h->d_cands(src), mm, nullptr, 1, rec, rb, h->stride >> 1) != LICHEE_OK) return LICHEE_ERR_CUDA;
                CK(cudaEventRecord(h->evk[0][4], h->stream));
                CK(cudaStreamSynchronize(h->stream));
                CK(cudaEventElapsedTime(&ms, h->evk[0][2], h->evk[0][3])); st.kernel_ms[LICHEE_K_LEAF] += ms;
                CK(cudaEventElapsedTime(&ms, h->evk[0][3], h->evk[0][4])); st.kernel_ms[LICHEE_K_SELECT] += ms;
                st.kernel_bytes[LICHEE_K_SELECT] += (long long)m * (long long)sizeof(Cand);
            }
            if (status != replay::S_RUNNING) break;
            if (num_grow >= budget) break;
            rec_cap = kReplayRecCap;
        }
    }
    if (fused) {
        // one launch, already synchronized: its own time is the device time (no second event, no second round trip)
        st.kernel_ms_total = st.kernel_ms[LICHEE_K_CHECK];
    } else {
        CK(cudaEventRecord(h->ev1, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        st.kernel_ms_total = ms;
    }
    st.num_trees = num_trees;
    st.num_grow_calls = num_grow;
    st.num_checks = num_checks;
    st.num_scored = num_trees;
    st.num_evaluated = num_trees;
    if (status == replay::S_RUNNING) return LICHEE_OK;          // budget used up: not the reference's answer yet
    completed = true;
    if (status == replay::S_TREE_CAP) st.status |= LICHEE_STATUS_TREE_CAP;
    if (status == replay::S_GROW_CAP) st.status |= LICHEE_STATUS_GROW_CAP;
    if (fused) {
        // the ranked list is already in pinned host memory
        const unsigned count = (unsigned)h->h_rout->count;
        const Cand *top = reinterpret_cast<const Cand *>(h->h_result);
        const unsigned char *tree = h->h_result + sizeof(Cand) * (size_t)h->prm.num_save;
        h->h_top.assign(top, top + count);
        h->h_tree.assign(tree, tree + (size_t)count * 2 * h->stride);
        st.num_saved = count;
        return LICHEE_OK;
    }
    return fetch_top(h);
}

// getLineageTrees + evaluateLineageTrees.  Reference order first (exact, sequential, budgeted); the
// canonical parallel search when the replay is off, the search is sharded, or the budget ran out.
int run_whole(lichee_search *h) {
    const bool want_ref = h->prm.ref_order_budget >= 0 && h->prm.part_count <= 1;
    lichee_search_stats ref_stats{};
    bool tried = false;
    if (want_ref) {
        bool completed = false;
        const int rc = h->SP == 32 ? run_reference_order<1>(h, completed) : run_reference_order<2>(h, completed);
        if (rc != LICHEE_OK) return rc;
        if (completed) return LICHEE_OK;
        ref_stats = h->stats;
        tried = true;
    }
    h->ref_result = false;
    if (h->cyclic) {
        set_err(tried ? "constraint network has a directed cycle and the reference-order replay did not finish within "
                        "ref_order_budget grow() calls; the canonical parallel search needs a DAG"
                      : "constraint network has a directed cycle: only the reference-order replay (ref_order_budget >= 0, "
                        "part_count <= 1) can search it");
        return LICHEE_ERR_CYCLIC;
    }
    const int rc = h->SP == 32 ? run_search<1>(h, h->prm.part_rank, h->prm.part_count, false)
                               : run_search<2>(h, h->prm.part_rank, h->prm.part_count, false);
    if (rc != LICHEE_OK) return rc;
    h->stats.status |= LICHEE_STATUS_CANONICAL_ORDER;
    if (tried) {                                            // the time the abandoned replay took is part of the search
        h->stats.kernel_ms_total += ref_stats.kernel_ms_total;
        h->stats.kernel_ms[LICHEE_K_CHECK] += ref_stats.kernel_ms[LICHEE_K_CHECK];
        h->stats.kernel_ms[LICHEE_K_LEAF] += ref_stats.kernel_ms[LICHEE_K_LEAF];
        h->stats.kernel_ms[LICHEE_K_SELECT] += ref_stats.kernel_ms[LICHEE_K_SELECT];
        h->stats.num_launches += ref_stats.num_launches;
        for (int i = 0; i < LICHEE_NUM_KERNELS; i++) h->stats.kernel_launches[i] += ref_stats.kernel_launches[i];
    }
    return LICHEE_OK;
}

}  // namespace

extern "C" {

void lichee_default_params(lichee_search_params *p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->vaf_error_margin = 0.1;          // Parameters.VAF_ERROR_MARGIN
    p->num_save = 1;                    // LineageEngine.Args.numSave
    p->max_num_trees = 100000;          // Parameters.MAX_NUM_TREES
    p->max_num_grow_calls = 100000000;  // Parameters.MAX_NUM_GROW_CALLS
    p->part_rank = 0;
    p->part_count = 1;
}

int lichee_release_cached_memory(void) {
    g_pool.release(-1);
    std::lock_guard<std::mutex> g(g_ctx.mu);
    for (auto &kv : g_ctx.free_list) {
        DevCtx &c = kv.second;
        cudaSetDevice(kv.first);
        if (c.pinned) cudaFreeHost(c.pinned);
        if (c.pinned_big) cudaFreeHost(c.pinned_big);
        if (c.ev0) cudaEventDestroy(c.ev0);
        if (c.ev1) cudaEventDestroy(c.ev1);
        for (int a = 0; a < 2; a++) for (int b = 0; b < 6; b++) if (c.evk[a][b]) cudaEventDestroy(c.evk[a][b]);
        for (int a = 0; a < 3; a++) if (c.evp[a]) cudaEventDestroy(c.evp[a]);
        if (c.stream) cudaStreamDestroy(c.stream);
        if (c.stream2) cudaStreamDestroy(c.stream2);
    }
    g_ctx.free_list.clear();
    return LICHEE_OK;
}

const char *lichee_last_error(void) { return g_err.c_str(); }
const char *lichee_version(void) { return "lichee_b200 0.1 (sm_100a)"; }

int lichee_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int lichee_search_destroy(lichee_search *h) {
    if (!h) return LICHEE_OK;
    if (h->stream) {
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stream);
        if (h->stream2) cudaStreamSynchronize(h->stream2);
    }
    give_back(h, h->b_net);
    for (DevBuf &b : h->b_level) give_back(h, b);
    give_back(h, h->b_mask); give_back(h, h->b_cnt); give_back(h, h->b_off); give_back(h, h->b_tile);
    give_back(h, h->b_small);
    give_back(h, h->b_radix);
    give_back(h, h->b_radix12); give_back(h, h->b_key0); give_back(h, h->b_key1);
    give_back(h, h->b_slab);
    give_back(h, h->b_sched);
    give_back(h, h->b_mask2); give_back(h, h->b_mask3); give_back(h, h->b_flag); give_back(h, h->b_bound);
    for (DevBuf &b : h->b_pf) give_back(h, b);
    give_back(h, h->b_lmask); give_back(h, h->b_lcnt); give_back(h, h->b_lbound); give_back(h, h->b_c16);
    give_back(h, h->b_rblob); give_back(h, h->b_rrec);
    for (int i = 0; i < 2; i++) { give_back(h, h->b_cands[i]); give_back(h, h->b_top[i]); give_back(h, h->b_tree[i]); }
    {
        DevCtx c;
        c.stream = h->stream; c.ev0 = h->ev0; c.ev1 = h->ev1; c.pinned = h->h_scan; c.pinned_big = h->h_big;
        c.stream2 = h->stream2;
        bool whole = c.stream && c.stream2 && c.ev0 && c.ev1 && c.pinned && c.pinned_big;
        for (int x = 0; x < 2; x++) for (int y = 0; y < 6; y++) { c.evk[x][y] = h->evk[x][y]; whole = whole && c.evk[x][y]; }
        for (int x = 0; x < 3; x++) { c.evp[x] = h->evp[x]; whole = whole && c.evp[x]; }
        if (whole) {
            g_ctx.put(h->device, c);
        } else {
            // a create that failed half way: nothing of it goes back to the pool
            if (c.pinned) cudaFreeHost(c.pinned);
            if (c.pinned_big) cudaFreeHost(c.pinned_big);
            if (c.ev0) cudaEventDestroy(c.ev0);
            if (c.ev1) cudaEventDestroy(c.ev1);
            for (int x = 0; x < 2; x++) for (int y = 0; y < 6; y++) if (c.evk[x][y]) cudaEventDestroy(c.evk[x][y]);
            for (int x = 0; x < 3; x++) if (c.evp[x]) cudaEventDestroy(c.evp[x]);
            if (c.stream) cudaStreamDestroy(c.stream);
            if (c.stream2) cudaStreamDestroy(c.stream2);
        }
    }
    delete h;
    return LICHEE_OK;
}

int lichee_search_create(lichee_search **out, int32_t n, int32_t S, const int32_t *adj_off,
                         const int32_t *adj_idx, const double *vaf, const lichee_search_params *params) {
    if (!out || !adj_off || !vaf || !params || n < 1 || S < 1 || (!adj_idx && adj_off[n] > 0)) {
        set_err("lichee_search_create: invalid argument");
        return LICHEE_ERR_INVALID;
    }
    if (n > LICHEE_MAX_NODES || S > LICHEE_MAX_SAMPLES) {
        set_err("network too large: %d nodes (max %d), %d samples (max %d)", n, LICHEE_MAX_NODES, S, LICHEE_MAX_SAMPLES);
        return LICHEE_ERR_TOO_LARGE;
    }
    if (params->num_save < 1 || params->num_save > LICHEE_MAX_SAVE || params->max_num_trees < 1 ||
        params->max_num_grow_calls < 1 || params->part_count < 0 || params->part_rank < 0 ||
        (params->part_count > 0 && params->part_rank >= params->part_count)) {
        set_err("lichee_search_create: parameter out of range (num_save 1..%d)", LICHEE_MAX_SAVE);
        return LICHEE_ERR_INVALID;
    }
    // parents of every node, ascending id.  Dropped: duplicates (PHYNetwork.addEdge refuses them), self loops
    // and edges into the root (the root is in the tree from the start, so neither the reference nor this
    // library ever follows one).  radj: the adjacency lists that remain, in list order.
    if (adj_off[0] != 0) { set_err("adj_off[0] must be 0"); return LICHEE_ERR_INVALID; }
    if (!(params->vaf_error_margin == params->vaf_error_margin) || std::isinf(params->vaf_error_margin)) {
        set_err("vaf_error_margin is not finite"); return LICHEE_ERR_INVALID;
    }
    for (size_t i = 0; i < (size_t)n * S; i++)
        if (!std::isfinite(vaf[i])) { set_err("vaf[%zu] is not finite", i); return LICHEE_ERR_INVALID; }
    std::vector<std::vector<int>> par(n), radj(n);
    std::vector<int> indeg(n, 0);
    for (int u = 0; u < n; u++) {
        if (adj_off[u + 1] < adj_off[u]) { set_err("adj_off not monotone"); return LICHEE_ERR_INVALID; }
        for (int k = adj_off[u]; k < adj_off[u + 1]; k++) {
            const int v = adj_idx[k];
            if (v < 0 || v >= n) { set_err("edge target %d out of range", v); return LICHEE_ERR_INVALID; }
            if (v == u || v == 0) continue;
            if (std::find(radj[u].begin(), radj[u].end(), v) == radj[u].end()) { radj[u].push_back(v); par[v].push_back(u); }
        }
    }
    for (int v = 0; v < n; v++) { std::sort(par[v].begin(), par[v].end()); indeg[v] = (int)par[v].size(); }
    bool cyclic = false;
    {
        std::vector<int> deg = indeg, q;               // Kahn
        for (int v = 0; v < n; v++) if (deg[v] == 0) q.push_back(v);
        size_t qi = 0;
        while (qi < q.size()) {
            const int u = q[qi++];
            for (int v : radj[u]) if (--deg[v] == 0) q.push_back(v);
        }
        // A directed cycle: the parent-choice search needs a DAG, the reference's Gabow-Myers search does not
        // (PHYNetwork.java:114-118,185-232 can produce one inside a group); only the replay can run it.
        cyclic = (int)q.size() != n;
    }
    if (lichee_device_count() < 1) {
        set_err("no CUDA device available: lichee_b200 has no CPU fallback");
        return LICHEE_ERR_NO_DEVICE;
    }
    lichee_search *h = new lichee_search();
    h->n = n; h->S = S; h->prm = *params;
    if (h->prm.part_count < 1) { h->prm.part_count = 1; h->prm.part_rank = 0; }
    h->device = params->device;
    h->L = n - 1;
    h->SP = S <= 32 ? 32 : 64;
    h->stride = std::max(16, ((h->L + 15) / 16) * 16);
    h->ibyte.resize((size_t)std::max(1, h->L));
    for (int t = 0; t < h->L; t++) h->ibyte[t] = (uint8_t)item_byte(t, h->stride >> 2);
    // a tree exists only if the root has children (PHYNetwork.java:557) and every other node a parent
    h->no_trees = radj[0].empty() && n > 1;
    h->cyclic = cyclic;
    for (int v = 1; v < n; v++) if (par[v].empty()) h->no_trees = true;
    if (n == 1) h->no_trees = true;          // getLineageTrees returns an empty list
    // sigma: non-root nodes by descending VAF mass, ties by ascending id
    std::vector<double> mass(n, 0.0);
    for (int v = 0; v < n; v++) { double m = 0; for (int s = 0; s < S; s++) m += vaf[(size_t)v * S + s]; mass[v] = m; }
    h->sigma.resize(h->L);
    for (int t = 0; t < h->L; t++) h->sigma[t] = t + 1;
    std::stable_sort(h->sigma.begin(), h->sigma.end(), [&](int a, int b) { return mass[a] > mass[b]; });
    h->cands.resize(h->L);
    for (int t = 0; t < h->L; t++) h->cands[t] = par[h->sigma[t]];

#define CKD(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            set_err("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_));  \
            lichee_search_destroy(h);                                                              \
            return LICHEE_ERR_CUDA;                                                                \
        }                                                                                          \
    } while (0)
    CKD(cudaSetDevice(h->device));
    {
        DevCtx c;
        if (g_ctx.get(h->device, c)) {
            h->stream = c.stream; h->ev0 = c.ev0; h->ev1 = c.ev1; h->h_scan = static_cast<ScanOut *>(c.pinned);
            h->h_big = static_cast<unsigned char *>(c.pinned_big);
            for (int x = 0; x < 2; x++) for (int y = 0; y < 6; y++) h->evk[x][y] = c.evk[x][y];
            h->stream2 = c.stream2;
            for (int x = 0; x < 3; x++) h->evp[x] = c.evp[x];
        } else {
            // the second stream only fills what the main one leaves idle: lowest priority against highest
            int prio_lo = 0, prio_hi = 0;
            CKD(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            CKD(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_hi));
            CKD(cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, prio_lo));
            for (int a = 0; a < 3; a++) CKD(cudaEventCreate(&h->evp[a]));
            CKD(cudaEventCreate(&h->ev0)); CKD(cudaEventCreate(&h->ev1));
            for (int a = 0; a < 2; a++) for (int b = 0; b < 6; b++) CKD(cudaEventCreate(&h->evk[a][b]));
            CKD(cudaMallocHost(&h->h_scan, sizeof(ScanOut) + sizeof(TopMeta) + sizeof(Sched) + sizeof(ReplayOut)));
            CKD(cudaMallocHost(&h->h_big, kPinnedBlob + kPinnedResult + kPinnedVaf + kPinnedDesc));
        }
    }
    const int Lp = std::max(1, h->L);
    // row index q of every node: 0 = root, t+1 = sigma[t]
    std::vector<int> qof(n, 0);
    for (int t = 0; t < h->L; t++) qof[h->sigma[t]] = t + 1;
    std::vector<double> vp((size_t)n * h->SP, 0.0);
    // SP == 64: the two halves of a row are interleaved ({x[l], x[l+32]} adjacent) for 128-bit loads
    for (int v = 0; v < n; v++)
        for (int s = 0; s < S; s++)
            vp[(size_t)qof[v] * h->SP + (h->SP == 64 ? 2 * (s & 31) + (s >> 5) : s)] = vaf[(size_t)v * S + s];
    std::vector<uint8_t> cd((size_t)Lp * kMaxN, 0), nc(kMaxN, 0);
    std::vector<uint4> sok(Lp, make_uint4(0, 0, 0, 0));
    for (int t = 0; t < h->L; t++) {
        nc[t] = (uint8_t)h->cands[t].size();
        unsigned bits[4] = {0, 0, 0, 0};
        for (size_t i = 0; i < h->cands[t].size(); i++) {
            const int p = h->cands[t][i], v = h->sigma[t];
            cd[(size_t)t * kMaxN + i] = (uint8_t)qof[p];
            // the sum rule for p with v as its only child: 0.0 + VAF(v) > VAF(p) + margin  (same doubles as the kernel)
            bool ok = true;
            for (int s = 0; s < S; s++) {
                volatile double sum = 0.0 + vaf[(size_t)v * S + s];
                volatile double capv = vaf[(size_t)p * S + s] + params->vaf_error_margin;
                if (sum > capv) ok = false;
            }
            if (ok) bits[i >> 5] |= 1u << (i & 31);
        }
        sok[t] = make_uint4(bits[0], bits[1], bits[2], bits[3]);
    }
    // level L-2, candidate i of node v without other children: do v and then the last node u both fit?
    // (0.0 + VAF(v)) + VAF(u) against VAF(p) + margin, the same doubles as dual_row
    uint4 sboth = make_uint4(0, 0, 0, 0);
    if (h->L >= 2) {
        const int t = h->L - 2, v = h->sigma[t], u = h->sigma[t + 1];
        unsigned bits[4] = {0, 0, 0, 0};
        for (size_t i = 0; i < h->cands[t].size(); i++) {
            const int p = h->cands[t][i];
            bool ok = true;
            for (int s = 0; s < S; s++) {
                volatile double sv = 0.0 + vaf[(size_t)v * S + s];
                volatile double svu = sv + vaf[(size_t)u * S + s];
                volatile double capv = vaf[(size_t)p * S + s] + params->vaf_error_margin;
                if (svu > capv) ok = false;
            }
            if (ok) bits[i >> 5] |= 1u << (i & 31);
        }
        sboth = make_uint4(bits[0], bits[1], bits[2], bits[3]);
    }
    if (h->L >= 1) h->sok_last = sok[h->L - 1];
    // last level, candidate whose only child is the last node: its excess E' through the device's
    // fixed-shape reduction (lane value t[l] + t[l+32], xor-butterfly over 32 lanes), same IEEE operations
    std::vector<double> lstat(kMaxN, 0.0);
    if (h->L > 0) {
        const int t = h->L - 1, v = h->sigma[t];
        for (size_t i = 0; i < h->cands[t].size(); i++) {
            const int p = h->cands[t][i];
            volatile double term[64];
            for (int s = 0; s < 64; s++) term[s] = 0.0;
            for (int s = 0; s < S; s++) {
                volatile double sum = 0.0 + vaf[(size_t)v * S + s];
                volatile double a = vaf[(size_t)p * S + s];
                if (sum > a) { volatile double d = sum - a; term[s] = d * d; }
            }
            volatile double lanev[32];
            for (int l = 0; l < 32; l++) lanev[l] = term[l] + term[l + 32];
            for (int d = 16; d > 0; d >>= 1) {
                volatile double nx[32];
                for (int l = 0; l < 32; l++) nx[l] = lanev[l] + lanev[l ^ d];
                for (int l = 0; l < 32; l++) lanev[l] = nx[l];
            }
            lstat[i] = lanev[0];
        }
    }
#define CKE(call) do { if ((call) != LICHEE_OK) { lichee_search_destroy(h); return LICHEE_ERR_CUDA; } } while (0)
    // one buffer for the network: [vaf | leaf_static | static_ok | cand | ncand]
    const size_t o_ls = vp.size() * sizeof(double), o_sok = o_ls + lstat.size() * sizeof(double);
    const size_t o_cand = o_sok + sok.size() * sizeof(uint4), o_ncand = o_cand + cd.size();
    const size_t o_desc = o_ncand + nc.size();
    std::vector<uint8_t> desc(kMaxN, 0);
    for (int k = 0; k < n; k++) desc[k] = (uint8_t)qof[n - 1 - k];
    CKE(ensure(h, h->b_net, o_desc + desc.size()));
    char *nb = static_cast<char *>(h->b_net.p);
    // The tables go to the device with the first search that reads them (upload_tables): a search that ends inside
    // the reference-order replay -- every search of a real data set -- only needs the matrix and desc_rows, which
    // the replay kernel stages straight from pinned host memory.
    h->tables.resize(o_desc + desc.size());
    memcpy(h->tables.data(), vp.data(), vp.size() * sizeof(double));
    memcpy(h->tables.data() + o_ls, lstat.data(), lstat.size() * sizeof(double));
    memcpy(h->tables.data() + o_sok, sok.data(), sok.size() * sizeof(uint4));
    memcpy(h->tables.data() + o_cand, cd.data(), cd.size());
    memcpy(h->tables.data() + o_ncand, nc.data(), nc.size());
    memcpy(h->tables.data() + o_desc, desc.data(), desc.size());
    h->p_vaf = reinterpret_cast<double *>(h->h_big + kPinnedBlob + kPinnedResult);
    h->p_desc = h->h_big + kPinnedBlob + kPinnedResult + kPinnedVaf;
    memcpy(h->p_vaf, vp.data(), vp.size() * sizeof(double));
    memcpy(h->p_desc, desc.data(), desc.size());
    h->cap = params->level_capacity > 0 ? params->level_capacity : (1ll << 21);
    if (h->cap < 1024) h->cap = 1024;
    h->chunk_max = std::min<long long>(h->cap, 1ll << 21);
    h->cand_cap = std::max<long long>(h->cap, 1ll << 23);
    h->b_level.assign(Lp, DevBuf{});
    h->head.assign(Lp, 0); h->tail.assign(Lp, 0);
    CKE(ensure(h, h->b_small, 256));
    h->h_meta = reinterpret_cast<TopMeta *>(h->h_scan + 1);
    h->h_sched = reinterpret_cast<Sched *>(h->h_meta + 1);
    h->h_rout = reinterpret_cast<ReplayOut *>(h->h_sched + 1);
    h->narrow = getenv("LICHEE_NO_NARROW") == nullptr;
    h->no_prefilter = getenv("LICHEE_NO_PREFILTER") != nullptr;
    h->no_prefetch = getenv("LICHEE_NO_PREFETCH") != nullptr;
    h->net.vaf = reinterpret_cast<const double *>(nb);
    h->net.static_ok = reinterpret_cast<const uint4 *>(nb + o_sok);
    h->net.leaf_static = reinterpret_cast<const double *>(nb + o_ls);
    h->net.cand = reinterpret_cast<const uint8_t *>(nb + o_cand);
    h->net.ncand = reinterpret_cast<const uint8_t *>(nb + o_ncand);
    h->net.desc_rows = reinterpret_cast<const uint8_t *>(nb + o_desc);
    h->net.n = n; h->net.S = S; h->net.SP = h->SP; h->net.L = h->L; h->net.stride = h->stride;
    h->net.margin = params->vaf_error_margin;
    h->net.static_both = sboth;
    {
        bool nonneg = params->vaf_error_margin == params->vaf_error_margin;
        for (size_t i = 0; i < (size_t)n * S; i++) if (!(vaf[i] >= 0.0) || !(vaf[i] <= 1e300)) nonneg = false;
        h->net.prune = (nonneg && getenv("LICHEE_NO_PRUNE") == nullptr) ? 1 : 0;
    }
    h->smem = (size_t)n * h->SP * sizeof(double) + kMaxN;
    h->smem_leaf = h->smem;
    if (h->device >= 0 && h->device < 64) {
      std::lock_guard<std::mutex> g(g_ctx.mu);
      if (!g_ctx.attrs_set[h->device]) {
        // opt in to the largest dynamic shared memory any network needs (128 rows x 64 samples)
        const int mx = kMaxN * 64 * (int)sizeof(double) + kMaxN;
        CKD(cudaFuncSetAttribute(check_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        CKD(cudaFuncSetAttribute(check_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        CKD(cudaFuncSetAttribute(leaf_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        CKD(cudaFuncSetAttribute(leaf_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        CKD(cudaFuncSetAttribute(check2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        CKD(cudaFuncSetAttribute(check2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        CKD(cudaFuncSetAttribute(descend_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx + kMaxN * kMaxN));
        CKD(cudaFuncSetAttribute(descend_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx + kMaxN * kMaxN));
        CKD(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(Cand) * 3 * kSelTile)));
        CKD(cudaFuncSetAttribute(select_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(Cand) * 2 * kSelTile)));
        g_ctx.attrs_set[h->device] = true;
      }
    }
    {
        // reference-order replay: adjacency lists in ROW space, list order kept
        std::vector<int> roff(n + 1, 0), ridx;
        std::vector<int> node_of_row(n, 0);
        for (int t = 0; t < h->L; t++) node_of_row[t + 1] = h->sigma[t];
        for (int q = 0; q < n; q++) {
            for (int v : radj[node_of_row[q]]) ridx.push_back(qof[v]);
            roff[q + 1] = (int)ridx.size();
        }
        h->rlay = replay::make_layout(n, h->SP, (int)ridx.size());
        if (h->rlay.total > kPinnedBlob) { set_err("replay state of %u bytes exceeds the pinned block", h->rlay.total); lichee_search_destroy(h); return LICHEE_ERR_STATE; }
        h->rblob0 = h->h_big;
        h->h_result = h->h_big + kPinnedBlob;
        replay::init_blob(h->rblob0, h->rlay, roff.data(), ridx.data(), params->max_num_grow_calls, params->max_num_trees);
        h->rblob_in_smem = (size_t)n * h->SP * sizeof(double) + h->rlay.total <= (size_t)kReplaySmemMax;
        CKE(ensure(h, h->b_rblob, h->rlay.total));
    }
    if (h->device >= 0 && h->device < 64) {
        std::lock_guard<std::mutex> g(g_ctx.mu);
        if (!g_ctx.attrs2_set[h->device]) {
            CKD(cudaFuncSetAttribute(replay_kernel<1, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kReplaySmemMax));
            CKD(cudaFuncSetAttribute(replay_kernel<2, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kReplaySmemMax));
            CKD(cudaFuncSetAttribute(replay_kernel<1, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kReplaySmemMax));
            CKD(cudaFuncSetAttribute(replay_kernel<2, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kReplaySmemMax));
            CKD(cudaFuncSetAttribute(replay_kernel<1, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kReplaySmemMax));
            CKD(cudaFuncSetAttribute(replay_kernel<2, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kReplaySmemMax));
            CKD(cudaFuncSetAttribute(replay_kernel<1, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kReplaySmemMax));
            CKD(cudaFuncSetAttribute(replay_kernel<2, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kReplaySmemMax));
            const int mx = kMaxN * 64 * (int)sizeof(double) + kMaxN;
            CKD(cudaFuncSetAttribute(score_ref_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
            CKD(cudaFuncSetAttribute(score_ref_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
            g_ctx.attrs2_set[h->device] = true;
        }
    }
#undef CKE
#undef CKD
    *out = h;
    return LICHEE_OK;
}

int lichee_search_run(lichee_search *h) {
    if (!h) { set_err("null handle"); return LICHEE_ERR_INVALID; }
    CK(cudaSetDevice(h->device));
    h->last_slice = -1;
    const int rc = run_whole(h);
    if (rc == LICHEE_OK) h->ran = true;
    return rc;
}

int lichee_search_run_slice(lichee_search *h, int32_t slice, int32_t n_slices, int32_t keep_ranked) {
    if (!h) { set_err("null handle"); return LICHEE_ERR_INVALID; }
    if (n_slices < 1 || slice < 0 || slice >= n_slices) { set_err("slice %d of %d out of range", slice, n_slices); return LICHEE_ERR_INVALID; }
    if (keep_ranked && !h->ran) { set_err("keep_ranked needs an earlier run on this handle"); return LICHEE_ERR_STATE; }
    if (h->cyclic) { set_err("constraint network has a directed cycle: a sliced (canonical) search needs a DAG"); return LICHEE_ERR_CYCLIC; }
    if (keep_ranked && h->ref_result) { set_err("keep_ranked: the list on this handle is a reference-order result"); return LICHEE_ERR_STATE; }
    // A tree whose score EQUALS the worst held one is admitted only if it precedes it in the canonical order; the
    // kernels decide that with `score < tau`, which is right exactly when every tree still to come follows the held
    // ones, i.e. when slices are taken in ascending order (a shared counter hands them out that way).
    if (keep_ranked && slice <= h->last_slice) {
        set_err("keep_ranked: slice %d after slice %d -- slices kept on one handle must be searched in ascending order", slice, h->last_slice);
        return LICHEE_ERR_STATE;
    }
    CK(cudaSetDevice(h->device));
    h->ref_result = false;
    const int rc = h->SP == 32 ? run_search<1>(h, slice, n_slices, keep_ranked != 0) : run_search<2>(h, slice, n_slices, keep_ranked != 0);
    if (rc == LICHEE_OK) { h->ran = true; h->stats.status |= LICHEE_STATUS_CANONICAL_ORDER; h->last_slice = slice; }
    return rc;
}

int lichee_search_clear(lichee_search *h) {
    if (!h) { set_err("null handle"); return LICHEE_ERR_INVALID; }
    CK(cudaSetDevice(h->device));
    init_meta<<<1, 1, 0, h->stream>>>(h->d_meta());
    CK(cudaGetLastError());
    h->flip = 0;
    h->ref_result = false;
    h->last_slice = -1;
    h->stats = lichee_search_stats{};
    h->stats.num_grow_calls = 0;
    h->stats.status = LICHEE_STATUS_CANONICAL_ORDER;
    const int rc = fetch_top(h);
    if (rc == LICHEE_OK) h->ran = true;
    return rc;
}

int lichee_search_set_max_num_trees(lichee_search *h, int64_t max_num_trees) {
    if (!h || max_num_trees < 1) { set_err("invalid argument"); return LICHEE_ERR_INVALID; }
    h->prm.max_num_trees = max_num_trees;
    reinterpret_cast<replay::Header *>(h->rblob0)->max_trees = max_num_trees;
    return LICHEE_OK;
}

int lichee_search_get_stats(const lichee_search *h, lichee_search_stats *out) {
    if (!h || !out) { set_err("null argument"); return LICHEE_ERR_INVALID; }
    if (!h->ran) { set_err("lichee_search_run has not completed"); return LICHEE_ERR_STATE; }
    *out = h->stats;
    return LICHEE_OK;
}

int lichee_search_get_order(const lichee_search *h, int32_t *sigma) {
    if (!h || !sigma) { set_err("null argument"); return LICHEE_ERR_INVALID; }
    for (int t = 0; t < h->L; t++) sigma[t] = h->sigma[t];
    return LICHEE_OK;
}

int lichee_search_get_tree(const lichee_search *h, int32_t k, double *score, int32_t *parent,
                           int32_t *order, int64_t *seq) {
    if (!h) { set_err("null handle"); return LICHEE_ERR_INVALID; }
    if (!h->ran) { set_err("lichee_search_run has not completed"); return LICHEE_ERR_STATE; }
    if (k < 0 || (size_t)k >= h->h_top.size()) { set_err("tree index %d out of range", k); return LICHEE_ERR_INVALID; }
    if (score) *score = h->h_top[k].score;
    if (seq) *seq = h->h_top[k].seq;
    if (h->ref_result) {
        // A = treeNodes (without the root) in the order the nodes joined the tree, B = their parents, as rows
        const uint8_t *a = h->h_tree.data() + (size_t)k * 2 * h->stride, *bp = a + h->stride;
        if (parent) parent[0] = -1;
        if (order) order[0] = 0;
        for (int i = 0; i < h->L; i++) {
            const int qa = a[h->ibyte[i]], qb = bp[h->ibyte[i]];
            if (qa < 1 || qa > h->L || qb > h->L) { set_err("corrupt tree record: rows %d <- %d at position %d", qa, qb, i); return LICHEE_ERR_STATE; }
            const int node = h->sigma[qa - 1];
            if (parent) parent[node] = qb == 0 ? 0 : h->sigma[qb - 1];
            if (order) order[i + 1] = node;
        }
        return LICHEE_OK;
    }
    const uint8_t *b = h->h_tree.data() + (size_t)k * h->stride;
    // byte item_byte(t) holds the row index q of the parent of sigma[t]; q = 0 root, q = j+1 node sigma[j]
    int par_of[kMaxN];
    for (int t = 0; t < h->L; t++) {
        const int q = b[h->ibyte[t]];
        if (q > h->L) { set_err("corrupt tree record: parent row %d at position %d", q, t); return LICHEE_ERR_STATE; }
        par_of[t] = q == 0 ? 0 : h->sigma[q - 1];
    }
    if (parent) {
        parent[0] = -1;
        for (int t = 0; t < h->L; t++) parent[h->sigma[t]] = par_of[t];
    }
    if (order) {
        // treeNodes: root first, then a preorder in which a node always follows its parent and the
        // children of one parent keep the canonical (sigma) order used for the sums.  Children lists as
        // CSR (counting sort by parent, stable in t), no allocation.
        int first[kMaxN + 1], fill[kMaxN], kids[kMaxN], stack[kMaxN];
        for (int u = 0; u <= h->n; u++) first[u] = 0;
        for (int t = 0; t < h->L; t++) first[par_of[t] + 1]++;
        for (int u = 0; u < h->n; u++) { first[u + 1] += first[u]; fill[u] = first[u]; }
        for (int t = 0; t < h->L; t++) kids[fill[par_of[t]]++] = h->sigma[t];
        int sp = 0, w = 0;
        stack[sp++] = 0;
        while (sp > 0 && w < h->n) {
            const int u = stack[--sp];
            order[w++] = u;
            for (int i = first[u + 1] - 1; i >= first[u] && sp < kMaxN; i--) stack[sp++] = kids[i];
        }
        for (; w < h->n; w++) order[w] = -1;
    }
    return LICHEE_OK;
}

int lichee_search_get_trees(const lichee_search *h, int32_t first, int32_t count, double *scores,
                            int32_t *parents, int32_t *orders, int64_t *seqs) {
    if (!h) { set_err("null handle"); return LICHEE_ERR_INVALID; }
    for (int32_t k = 0; k < count; k++) {
        const int rc = lichee_search_get_tree(h, first + k, scores ? scores + k : nullptr,
                                              parents ? parents + (size_t)k * h->n : nullptr,
                                              orders ? orders + (size_t)k * h->n : nullptr, seqs ? seqs + k : nullptr);
        if (rc != LICHEE_OK) return rc;
    }
    return LICHEE_OK;
}

int64_t lichee_record_bytes(const lichee_search *h) { return h ? 24 + h->stride : 0; }

int lichee_search_export_ranked(const lichee_search *h, void *dst, int device_ptr) {
    if (!h || !dst) { set_err("null argument"); return LICHEE_ERR_INVALID; }
    if (!h->ran) { set_err("lichee_search_run has not completed"); return LICHEE_ERR_STATE; }
    if (h->ref_result) { set_err("export_ranked: the list is a reference-order result (one sequential search); shard with part_count > 1 or ref_order_budget < 0"); return LICHEE_ERR_STATE; }
    CK(cudaSetDevice(h->device));
    if (upload_tables(const_cast<lichee_search *>(h)) != LICHEE_OK) return LICHEE_ERR_CUDA;
    const long long rb = 24 + h->stride;
    unsigned char *d = nullptr;
    lichee_search *hm = const_cast<lichee_search *>(h);
    if (device_ptr) d = static_cast<unsigned char *>(dst);
    else CK(cudaMalloc(&d, (size_t)rb * h->prm.num_save));
    export_records<<<h->prm.num_save, 32, 0, h->stream>>>(hm->d_meta(), hm->d_top(h->flip), hm->d_tree(h->flip),
                                                          h->stride >> 2, h->prm.num_save, d, rb);
    CK(cudaGetLastError());
    if (!device_ptr) {
        cudaError_t e1 = cudaMemcpyAsync(dst, d, (size_t)rb * h->prm.num_save, cudaMemcpyDeviceToHost, h->stream);
        if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(h->stream);
        cudaFree(d);
        CK(e1);
    } else {
        CK(cudaStreamSynchronize(h->stream));
    }
    return LICHEE_OK;
}

int lichee_search_merge_ranked(lichee_search *h, const void *records, int32_t n_lists, int device_ptr) {
    if (!h || !records || n_lists < 1) { set_err("invalid argument"); return LICHEE_ERR_INVALID; }
    if (!h->ran) { set_err("lichee_search_run has not completed"); return LICHEE_ERR_STATE; }
    if (h->ref_result) { set_err("merge_ranked: the list on this handle is a reference-order result"); return LICHEE_ERR_STATE; }
    CK(cudaSetDevice(h->device));
    if (upload_tables(h) != LICHEE_OK) return LICHEE_ERR_CUDA;
    const long long rb = 24 + h->stride;
    const long long n = (long long)n_lists * h->prm.num_save;
    if (ensure(h, h->b_cands[0], sizeof(Cand) * (size_t)n) != LICHEE_OK) return LICHEE_ERR_CUDA;
    const unsigned char *d = static_cast<const unsigned char *>(records);
    unsigned char *tmp = nullptr;
    if (!device_ptr) {
        CK(cudaMalloc(&tmp, (size_t)rb * n));
        CK(cudaMemcpyAsync(tmp, records, (size_t)rb * n, cudaMemcpyHostToDevice, h->stream));
        d = tmp;
    }
    records_to_cands<<<(int)((n + 255) / 256), 256, 0, h->stream>>>(d, rb, n, h->d_cands(0), h->d_meta());
    CK(cudaGetLastError());
    // (every list is sorted: the merge form of the selection, unless the lists fit one rank-merge tile anyway)
    if (run_select(h, h->d_cands(0), n, nullptr, n > kSelTile ? 2 : 1, d, rb) != LICHEE_OK) { if (tmp) cudaFree(tmp); return LICHEE_ERR_CUDA; }
    const int rc = fetch_top(h);
    if (tmp) cudaFree(tmp);
    return rc;
}

int lichee_get_lineage_trees(int32_t n, int32_t S, const int32_t *adj_off, const int32_t *adj_idx,
                             const double *vaf, const lichee_search_params *params,
                             lichee_search_stats *stats, double *scores, int32_t *parents, int32_t *orders) {
    lichee_search *h = nullptr;
    int rc = lichee_search_create(&h, n, S, adj_off, adj_idx, vaf, params);
    if (rc != LICHEE_OK) return rc;
    rc = lichee_search_run(h);
    if (rc == LICHEE_OK) {
        if (stats) *stats = h->stats;
        for (uint32_t k = 0; k < h->stats.num_saved && rc == LICHEE_OK; k++)
            rc = lichee_search_get_tree(h, (int32_t)k, scores ? scores + k : nullptr,
                                        parents ? parents + (size_t)k * n : nullptr,
                                        orders ? orders + (size_t)k * n : nullptr, nullptr);
    }
    lichee_search_destroy(h);
    return rc;
}

}  // extern "C"
