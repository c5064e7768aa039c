/*
 * kdbuild.cu -- drop-in for kd_tree_build (kd-tree.c:421-477), the serial
 * bottleneck in front of the render path (10-13 s per
 * frame for BASELINE config 2 on one host core).
 *
 * The reference builder is exhaustive: at every node it scores every
 * candidate plane (each dimension x each item x {lower-2EPS, upper+2EPS}) by
 * counting the node's items left / right / straddling (kdtree_split_score,
 * kd-tree.c:294-313) and keeps the first best (strict >, kd-tree.c:323-345).
 * That is 2*D*n^2 comparisons per node -- 1.4e9 at the root of config 2 -- and
 * embarrassingly parallel.  Here one GPU thread owns one candidate and streams
 * the node's item intervals through shared memory; a block reduction keeps
 * (best score, lowest candidate index), the host picks the winner among the
 * per-block results, partitions the item list exactly like
 * kd_tree_split_node (kd-tree.c:381-403: order preserved, straddlers to both
 * sides, infinite objects dropped) and recurses.  Nodes with few items are
 * scored on the host with the same comparisons.
 *
 * The result is the reference's own `kd_tree_t` in host memory (calloc'ed
 * nodes, leaf id/pointer arrays, root AABB, infinite-object list), bit for bit
 * what kd_tree_build produces, so everything downstream -- including the
 * reference's CPU renderer and kd_tree_free -- keeps working.  Tie-breaking of
 * the render path depends on the tree shape (SURVEY note 5); the test compares
 * the flattened trees byte for byte.
 */
#include <cuda_runtime.h>
#include <float.h>
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#include <pthread.h>
#include "ndt_abi.h"
#include "ndt_internal.h"

#define EPS NDT_EPS
#define KD_GPU_MIN_ITEMS 48      /* below this the host scores the node itself */
#define KD_BLOCK 256
#define NDT_KD_MAX_DEVICES 64

/* kd-tree.h:22-26, 40-43 */
typedef struct {
    ndtabi_vec lower, upper;
    int id;
    void *obj_ptr;
} kdb_item;
typedef struct {
    kdb_item **items;
    int n, cap;
} kdb_item_list;

struct Best {
    int score;      /* integer valued in the reference too (kd-tree.c:311) */
    int cand;       /* (dim*n + i)*2 + side; lower wins ties = first best */
};

__device__ __forceinline__ bool better(const Best &a, const Best &b)
{
    return a.score > b.score || (a.score == b.score && a.cand < b.cand);
}

/* one thread per candidate plane of one node */
__global__ void __launch_bounds__(KD_BLOCK)
k_kd_score(const double *__restrict__ lo, const double *__restrict__ hi, int n_total,
           const int *__restrict__ list, int n, int dims, Best *block_best)
{
    __shared__ double s_lo[KD_BLOCK], s_hi[KD_BLOCK];
    __shared__ Best s_best[KD_BLOCK / 32];
    const int per_dim = 2 * n;
    const int blocks_per_dim = (per_dim + KD_BLOCK - 1) / KD_BLOCK;
    const int d = blockIdx.x / blocks_per_dim;
    const int c = (blockIdx.x % blocks_per_dim) * KD_BLOCK + threadIdx.x;   /* candidate within the dimension */
    const bool live = c < per_dim;
    const double *dlo = lo + (size_t)d * n_total, *dhi = hi + (size_t)d * n_total;
    double pos = 0.0;
    if (live) {
        const int it = list[c >> 1];
        pos = (c & 1) ? dhi[it] + 2 * EPS : dlo[it] - 2 * EPS;              /* kd-tree.c:328,336 */
    }
    const double pl = pos - EPS, pr = pos + EPS;
    int left = 0, right = 0;
    for (int base = 0; base < n; base += KD_BLOCK) {
        const int j = base + threadIdx.x;
        if (j < n) { const int it = list[j]; s_lo[threadIdx.x] = dlo[it]; s_hi[threadIdx.x] = dhi[it]; }
        __syncthreads();
        const int m = min(KD_BLOCK, n - base);
        if (live) {
            for (int k = 0; k < m; ++k) {                                   /* kd-tree.c:304-309 */
                const double il = s_lo[k], iu = s_hi[k];
                if (iu < pl) ++left;
                else if (il > pr) ++right;
            }
        }
        __syncthreads();
    }
    Best b;
    b.score = INT_MIN; b.cand = INT_MAX;
    if (live && left > 0 && right > 0) {
        const int unsplit = n - left - right;
        b.score = n - (abs(left - right) + 2 * unsplit);
        b.cand = (d * n + (c >> 1)) * 2 + (c & 1);
    }
    for (int o = 16; o > 0; o >>= 1) {
        Best t;
        t.score = __shfl_down_sync(0xffffffffu, b.score, o);
        t.cand = __shfl_down_sync(0xffffffffu, b.cand, o);
        if (better(t, b)) b = t;
    }
    if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < KD_BLOCK / 32; ++w) if (better(s_best[w], b)) b = s_best[w];
        block_best[blockIdx.x] = b;
    }
}

struct Builder {
    int dims, n_total;
    kdb_item **items;            /* all items, index = id */
    double *h_lo, *h_hi;         /* [dims][n_total] */
    double *d_lo, *d_hi;
    int *d_list; Best *d_best; Best *h_best;
    size_t list_cap, best_cap;
    cudaStream_t st;
    int gpu_nodes, host_nodes, total_nodes;
    const char *err;
    /* ndt_b200_kd_tree_build_bounded: 0 / 0 / 0 = the reference's unbounded recursion */
    int max_depth, leaf_size;
    double max_growth;       /* a split whose two sides hold more than max_growth * n references is not made */
};

/* the reference's search on the host, same comparisons (small nodes) */
static bool host_search(const Builder &B, const int *list, int n, int *out_dim, double *out_pos)
{
    bool found = false;
    int best = INT_MIN;
    for (int d = 0; d < B.dims; ++d) {
        const double *lo = B.h_lo + (size_t)d * B.n_total, *hi = B.h_hi + (size_t)d * B.n_total;
        for (int i = 0; i < n; ++i) {
            for (int side = 0; side < 2; ++side) {
                const double pos = side ? hi[list[i]] + 2 * EPS : lo[list[i]] - 2 * EPS;
                int left = 0, right = 0;
                for (int j = 0; j < n; ++j) {
                    const double il = lo[list[j]], iu = hi[list[j]];
                    if (iu < pos - EPS) ++left;
                    else if (il > pos + EPS) ++right;
                }
                if (left > 0 && right > 0) {
                    const int score = n - (abs(left - right) + 2 * (n - left - right));
                    if (score > best) { best = score; *out_dim = d; *out_pos = pos; found = true; }
                }
            }
        }
    }
    return found;
}

static bool gpu_search(Builder &B, const int *list, int n, int *out_dim, double *out_pos)
{
    const int per_dim = 2 * n;
    const int blocks_per_dim = (per_dim + KD_BLOCK - 1) / KD_BLOCK;
    const int blocks = blocks_per_dim * B.dims;
    if ((size_t)n > B.list_cap || (size_t)blocks > B.best_cap) { B.err = "kd build: scratch too small"; return false; }
    cudaError_t e = cudaMemcpyAsync(B.d_list, list, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, B.st);
    if (e == cudaSuccess) {
        k_kd_score<<<blocks, KD_BLOCK, 0, B.st>>>(B.d_lo, B.d_hi, B.n_total, B.d_list, n, B.dims, B.d_best);
        e = cudaGetLastError();         /* a launch that never started leaves h_best stale: say so instead of decoding it */
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(B.h_best, B.d_best, (size_t)blocks * sizeof(Best), cudaMemcpyDeviceToHost, B.st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(B.st);
    if (e != cudaSuccess) { B.err = cudaGetErrorString(e); return false; }
    Best b; b.score = INT_MIN; b.cand = INT_MAX;
    for (int k = 0; k < blocks; ++k) {
        const Best &t = B.h_best[k];
        if (t.score > b.score || (t.score == b.score && t.cand < b.cand)) b = t;
    }
    if (b.score == INT_MIN) return false;
    const int side = b.cand & 1, rest = b.cand >> 1, d = rest / n, i = rest % n;
    if (b.cand < 0 || d >= B.dims) { B.err = "kd build: the device returned a candidate outside the node"; return false; }
    *out_dim = d;
    *out_pos = side ? B.h_hi[(size_t)d * B.n_total + list[i]] + 2 * EPS
                    : B.h_lo[(size_t)d * B.n_total + list[i]] - 2 * EPS;
    return true;
}

/* kd_tree_split_node, kd-tree.c:315-419 */
static int split_node(Builder &B, ndtabi_kd_node *node, const int *list, int n, int depth = 0)
{
    ++B.total_nodes;
    int split_dim = node->dim;
    double split_pos = 0.0;
    bool found = false;
    const bool stop = (B.max_depth > 0 && depth >= B.max_depth) || (B.leaf_size > 0 && n <= B.leaf_size);
    if (!stop) {
        if (n >= KD_GPU_MIN_ITEMS) { found = gpu_search(B, list, n, &split_dim, &split_pos); ++B.gpu_nodes; }
        else { found = host_search(B, list, n, &split_dim, &split_pos); ++B.host_nodes; }
    }
    if (B.err) return -1;
    if (found && B.max_growth > 0.0) {
        /* bounded build: overlapping objects straddle every plane; do not split when the two sides
         * together would hold more than max_growth * n references (the reference's builder does, and
         * never finishes on scenes like random.c with a few hundred objects: SURVEY note 8) */
        const double *lo = B.h_lo + (size_t)split_dim * B.n_total, *hi = B.h_hi + (size_t)split_dim * B.n_total;
        long both = 0;
        for (int i = 0; i < n; ++i) {
            const double il = lo[list[i]], iu = hi[list[i]];
            if (!(iu < split_pos - EPS) && !(il > split_pos + EPS)) ++both;
        }
        if ((double)(n + both) > B.max_growth * (double)n) found = false;
    }
    if (!found) {   /* leaf, kd-tree.c:360-375 */
        node->num = n;
        node->dim = -1;
        node->boundary = 0.0;
        node->obj_ids = (int *)calloc((size_t)n, sizeof(int *));
        node->objs = (void **)calloc((size_t)n, sizeof(void *));
        for (int i = 0; i < n; ++i) {
            node->obj_ids[i] = B.items[list[i]]->id;
            node->objs[i] = B.items[list[i]]->obj_ptr;
        }
        node->left = node->right = NULL;
        return 1;
    }
    node->dim = split_dim;
    node->boundary = split_pos;
    node->left = (ndtabi_kd_node *)calloc(1, sizeof(ndtabi_kd_node));
    node->right = (ndtabi_kd_node *)calloc(1, sizeof(ndtabi_kd_node));
    node->left->dim = node->right->dim = -1;                    /* kd_node_init */
    int *l = (int *)malloc((size_t)(n ? n : 1) * sizeof(int)), *r = (int *)malloc((size_t)(n ? n : 1) * sizeof(int));
    int nl = 0, nr = 0;
    const double *lo = B.h_lo + (size_t)split_dim * B.n_total, *hi = B.h_hi + (size_t)split_dim * B.n_total;
    for (int i = 0; i < n; ++i) {
        const int it = list[i];
        const double radius = ((ndtabi_object *)B.items[it]->obj_ptr)->bounds.radius;
        if (radius < 0.0) continue;                               /* kd-tree.c:385-389 */
        const double il = lo[it], iu = hi[it];
        if (iu < split_pos - EPS) l[nl++] = it;
        else if (il > split_pos + EPS) r[nr++] = it;
        else { l[nl++] = it; r[nr++] = it; }
    }
    node->left->dim = (node->dim + 1) % B.dims;
    node->right->dim = (node->dim + 1) % B.dims;
    int rc = 0;
    if (nl > 0 && nr > 0) {
        if (split_node(B, node->left, l, nl, depth + 1) < 0) rc = -1;
        if (rc == 0 && split_node(B, node->right, r, nr, depth + 1) < 0) rc = -1;
    }
    free(l); free(r);
    return rc;
}

static int kd_build(void *tree_v, void *items_v, int max_depth, int leaf_size, double max_growth);

extern "C" int ndt_b200_kd_tree_build(void *tree_v, void *items_v)
{
    return kd_build(tree_v, items_v, 0, 0, 0.0);
}

/* The same search (every candidate plane scored on the GPU, kd-tree.c:294-345) with a bounded
 * recursion, for scenes the reference's builder cannot finish (note 8): stops at
 * max_depth, at leaf_size items, and wherever a split would grow the number of references by more than
 * max_growth.  The tree is a valid kd_tree_t for kd_tree_intersect / ndt_b200_flatten; it is NOT the
 * tree the reference would build (it cannot build one), so results are checked against the reference's
 * tree-less trace() instead. */
extern "C" int ndt_b200_kd_tree_build_bounded(void *tree_v, void *items_v, int max_depth, int leaf_size, double max_growth)
{
    if (max_depth < 1 || leaf_size < 1 || !(max_growth > 1.0))
        return ndt_set_error(NDT_B200_E_ARG, "ndt_b200_kd_tree_build_bounded: max_depth >= 1, leaf_size >= 1, max_growth > 1");
    return kd_build(tree_v, items_v, max_depth, leaf_size, max_growth);
}

static int kd_build(void *tree_v, void *items_v, int max_depth, int leaf_size, double max_growth)
{
    ndtabi_kd_tree *tree = (ndtabi_kd_tree *)tree_v;
    kdb_item_list *items = (kdb_item_list *)items_v;
    if (!tree || !items) return ndt_set_error(NDT_B200_E_ARG, "ndt_b200_kd_tree_build: NULL argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return ndt_set_error(NDT_B200_E_CUDA, "no CUDA device; libndt_b200 has no CPU path");
    const int dims = tree->bb_lower.n, n = items->n;
    if (dims < 1 || dims > NDT_MAX_DIM) return ndt_set_error(NDT_B200_E_ARG, "kd tree not initialised (kd_tree_init)");
    if (tree->root == NULL) tree->root = (ndtabi_kd_node *)calloc(1, sizeof(ndtabi_kd_node));

    /* kd-tree.c:428-459 */
    int num_fin = 0, num_inf = 0;
    for (int i = 0; i < n; ++i) {
        if (((ndtabi_object *)items->items[i]->obj_ptr)->bounds.radius >= 0.0) ++num_fin; else ++num_inf;
    }
    tree->root->objs = (void **)calloc((size_t)num_fin, sizeof(void *));
    tree->inf_obj_ptrs = (void **)calloc((size_t)num_inf, sizeof(void *));
    tree->obj_num = 0;
    tree->inf_obj_num = 0;
    Builder B;
    memset(&B, 0, sizeof B);
    B.dims = dims; B.n_total = n; B.items = items->items;
    B.max_depth = max_depth; B.leaf_size = leaf_size; B.max_growth = max_growth;
    B.h_lo = (double *)malloc((size_t)dims * (n ? n : 1) * sizeof(double));
    B.h_hi = (double *)malloc((size_t)dims * (n ? n : 1) * sizeof(double));
    int *root_list = (int *)malloc((size_t)(n ? n : 1) * sizeof(int));
    int n_root = 0;
    for (int i = 0; i < n; ++i) {
        kdb_item *it = items->items[i];
        it->id = i;
        for (int d = 0; d < dims; ++d) { B.h_lo[(size_t)d * n + i] = it->lower.v[d]; B.h_hi[(size_t)d * n + i] = it->upper.v[d]; }
        if (((ndtabi_object *)it->obj_ptr)->bounds.radius >= 0.0) {
            tree->root->objs[tree->obj_num++] = it->obj_ptr;
            root_list[n_root++] = i;
            for (int d = 0; d < dims; ++d) {                     /* aabb_add, kd-tree.c:43-61 */
                if (it->lower.v[d] < tree->bb_lower.v[d]) tree->bb_lower.v[d] = it->lower.v[d];
                if (it->upper.v[d] > tree->bb_upper.v[d]) tree->bb_upper.v[d] = it->upper.v[d];
            }
        } else {
            tree->inf_obj_ptrs[tree->inf_obj_num++] = it->obj_ptr;
        }
    }
    tree->root->dim = 0;
    tree->obj_num = n;                                            /* kd-tree.c:470 */

    /* device scratch is kept for the life of the process (cudaMalloc / cudaFree /
     * stream creation cost 10-300 ms, the build itself 17 ms for 6561 items): one set PER DEVICE --
     * the build runs on whichever device is current for the calling thread -- and one build at a time */
    static struct Scratch { cudaStream_t st; double *d_lo, *d_hi; int *d_list; Best *d_best, *h_best;
                            size_t vb, list_cap, best_cap; } Ks[NDT_KD_MAX_DEVICES];
    static pthread_mutex_t kd_lock = PTHREAD_MUTEX_INITIALIZER;
    int cur_dev = 0;
    if (cudaGetDevice(&cur_dev) != cudaSuccess || cur_dev < 0 || cur_dev >= NDT_KD_MAX_DEVICES)
        return ndt_set_error(NDT_B200_E_CUDA, "ndt_b200_kd_tree_build: no usable current device");
    struct Unlock { pthread_mutex_t *m; ~Unlock() { pthread_mutex_unlock(m); } } unlock_at_exit = { &kd_lock };
    pthread_mutex_lock(&kd_lock);
    Scratch &K = Ks[cur_dev];
    int rc = 0;
    struct timespec t0, t1, t2;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    cudaError_t e = cudaSuccess;
    const size_t vb = (size_t)dims * (n ? n : 1) * sizeof(double);
    const size_t list_cap = (size_t)(n ? n : 1);
    const size_t best_cap = (size_t)dims * ((2 * list_cap + KD_BLOCK - 1) / KD_BLOCK) + 8;
    if (!K.st) e = cudaStreamCreateWithFlags(&K.st, cudaStreamNonBlocking);
    if (e == cudaSuccess && K.vb < vb) {
        cudaFree(K.d_lo); cudaFree(K.d_hi); K.d_lo = K.d_hi = NULL; K.vb = 0;
        e = cudaMalloc(&K.d_lo, vb + vb / 2);
        if (e == cudaSuccess) e = cudaMalloc(&K.d_hi, vb + vb / 2);
        if (e == cudaSuccess) K.vb = vb + vb / 2;
    }
    if (e == cudaSuccess && K.list_cap < list_cap) {
        cudaFree(K.d_list); K.d_list = NULL; K.list_cap = 0;
        e = cudaMalloc(&K.d_list, 2 * list_cap * sizeof(int));
        if (e == cudaSuccess) K.list_cap = 2 * list_cap;
    }
    if (e == cudaSuccess && K.best_cap < best_cap) {
        cudaFree(K.d_best); if (K.h_best) cudaFreeHost(K.h_best);
        K.d_best = K.h_best = NULL; K.best_cap = 0;
        e = cudaMalloc(&K.d_best, 2 * best_cap * sizeof(Best));
        if (e == cudaSuccess) e = cudaMallocHost(&K.h_best, 2 * best_cap * sizeof(Best));
        if (e == cudaSuccess) K.best_cap = 2 * best_cap;
    }
    B.st = K.st; B.d_lo = K.d_lo; B.d_hi = K.d_hi; B.d_list = K.d_list; B.d_best = K.d_best; B.h_best = K.h_best;
    B.list_cap = K.list_cap; B.best_cap = K.best_cap;
    if (e == cudaSuccess) e = cudaMemcpyAsync(B.d_lo, B.h_lo, vb, cudaMemcpyHostToDevice, B.st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(B.d_hi, B.h_hi, vb, cudaMemcpyHostToDevice, B.st);
    if (e != cudaSuccess) {
        rc = ndt_set_error(NDT_B200_E_CUDA, "ndt_b200_kd_tree_build: %s", cudaGetErrorString(e));
    } else {
        cudaStreamSynchronize(B.st);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        rc = split_node(B, tree->root, root_list, n_root);
        clock_gettime(CLOCK_MONOTONIC, &t2);
        if (getenv("NDT_B200_KD_TIMING"))
            fprintf(stderr, "ndt_b200_kd_tree_build: %d items, setup %.2f ms, build %.2f ms (%d nodes: %d searched on the GPU, %d on the host)\n",
                    n, ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec)) * 1e3,
                    ((t2.tv_sec - t1.tv_sec) + 1e-9 * (t2.tv_nsec - t1.tv_nsec)) * 1e3,
                    B.total_nodes, B.gpu_nodes, B.host_nodes);
        if (rc < 0) rc = ndt_set_error(NDT_B200_E_CUDA, "ndt_b200_kd_tree_build: %s", B.err ? B.err : "failed");
    }
    free(B.h_lo); free(B.h_hi); free(root_list);
    return rc;
}
