/*
 * oracle/port.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's CPU algorithms for the hot path.  It is the checker the
 * CUDA path is compared with; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  Nothing under pcrcg_b200/ may.
 *
 * Pinned against the UNMODIFIED reference C++ core (oracle/_ref, built by oracle/Makefile) by
 * tests/test_oracle_pinning.py and against the committed vectors under tests/golden/.
 *
 * Reference files followed (paths relative to /root/reference; "zip!" = cpp_wrappers.zip!cpp_wrappers/):
 *   zip!cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106   grid_subsampling
 *   zip!cpp_subsampling/grid_subsampling/grid_subsampling.cpp:109-211 batch_grid_subsampling
 *   zip!cpp_subsampling/grid_subsampling/grid_subsampling.h:74-79     SampledData::update_points
 *   zip!cpp_subsampling/grid_subsampling/grid_subsampling.h:42-73     SampledData::update_all / _features / _classes
 *   zip!cpp_utils/cloud/cloud.cpp:27-66                               min_point / max_point
 *   cpp_wrappers/cpp_neighbors/neighbors/neighbors.cpp:211-332        batch_nanoflann_neighbors
 *   zip!cpp_utils/nanoflann/nanoflann.hpp:432-440 (L2_Simple_Adaptor), :250 (strict d2 < r2),
 *                                         :208-214,1287 (sort by distance)
 * plus the published behaviour of libstdc++'s std::unordered_map<size_t,...> (GCC 13,
 * bits/hashtable.h _M_insert_bucket_begin / _M_rehash_aux(unique), bits/hashtable_policy.h
 * _Prime_rehash_policy), because the reference's OUTPUT ORDER is that container's iteration order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------- */
/* libstdc++ unordered_map<size_t,T> order model                                               */
/* ------------------------------------------------------------------------------------------- */

/* Bucket-count schedule of GCC 13's _Prime_rehash_policy with max_load_factor 1 and growth 2,
 * starting from the single bucket; checked against the real container by the pinning test
 * (ref_bucket_schedule in oracle/ref_shim.cpp).  The rehash happens BEFORE inserting the element
 * that would make size > bucket_count. */
static const uint64_t k_bucket_schedule[] = {
    13ull, 29ull, 59ull, 127ull, 257ull, 541ull, 1109ull, 2357ull, 5087ull, 10273ull, 20753ull,
    42043ull, 85229ull, 172933ull, 351061ull, 712697ull, 1447153ull, 2938679ull, 5967347ull,
    12117689ull, 24607243ull, 49969847ull, 101473717ull, 206062531ull, 418450807ull,
    849747061ull, 1725587117ull };
#define N_SCHEDULE ((int)(sizeof(k_bucket_schedule) / sizeof(k_bucket_schedule[0])))

int oracle_bucket_schedule(uint64_t* out, int cap)
{
    int n = N_SCHEDULE < cap ? N_SCHEDULE : cap;
    memcpy(out, k_bucket_schedule, sizeof(uint64_t) * (size_t)n);
    return n;
}

/* Singly linked node list + bucket array holding the node BEFORE the first node of the bucket,
 * exactly the libstdc++ layout; (-1) = none, (-2) = the before-begin sentinel. */
typedef struct {
    int64_t* next;      /* next[node]                                   */
    uint64_t* key;      /* key[node]                                    */
    int64_t* bucket;    /* bucket[b] = node before the bucket's first   */
    uint64_t nbkt;
    int64_t head;       /* before_begin.next                            */
    int64_t size;
    uint64_t next_resize;   /* _Prime_rehash_policy::_M_next_resize (0 before the first insert) */
    int sched;
} umap_t;

#define UM_NONE (-1)
#define UM_BB   (-2)

static int64_t um_next(const umap_t* m, int64_t prev) { return prev == UM_BB ? m->head : m->next[prev]; }
static void um_set_next(umap_t* m, int64_t prev, int64_t v) { if (prev == UM_BB) m->head = v; else m->next[prev] = v; }

/* bits/hashtable.h: _M_insert_bucket_begin */
static void um_insert_bucket_begin(umap_t* m, uint64_t b, int64_t node)
{
    if (m->bucket[b] != UM_NONE) {
        m->next[node] = um_next(m, m->bucket[b]);
        um_set_next(m, m->bucket[b], node);
    } else {
        m->next[node] = m->head;
        m->head = node;
        if (m->next[node] != UM_NONE)
            m->bucket[m->key[m->next[node]] % m->nbkt] = node;
        m->bucket[b] = UM_BB;
    }
}

/* bits/hashtable.h: _M_rehash_aux(n, true_type) */
static void um_rehash(umap_t* m, uint64_t nb)
{
    free(m->bucket);
    m->bucket = (int64_t*)malloc(sizeof(int64_t) * nb);
    for (uint64_t i = 0; i < nb; i++) m->bucket[i] = UM_NONE;
    int64_t p = m->head;
    m->head = UM_NONE;
    uint64_t bbegin_bkt = 0;
    while (p != UM_NONE) {
        int64_t nx = m->next[p];
        uint64_t b = m->key[p] % nb;
        if (m->bucket[b] == UM_NONE) {
            m->next[p] = m->head;
            m->head = p;
            m->bucket[b] = UM_BB;
            if (m->next[p] != UM_NONE) m->bucket[bbegin_bkt] = p;
            bbegin_bkt = b;
        } else {
            m->next[p] = um_next(m, m->bucket[b]);
            um_set_next(m, m->bucket[b], p);
        }
        p = nx;
    }
    m->nbkt = nb;
}

static int64_t um_find(const umap_t* m, uint64_t k)
{
    if (m->size == 0) return UM_NONE;
    uint64_t b = k % m->nbkt;
    int64_t prev = m->bucket[b];
    if (prev == UM_NONE) return UM_NONE;
    for (int64_t p = um_next(m, prev); p != UM_NONE; p = m->next[p]) {
        if (m->key[p] == k) return p;
        if (m->key[p] % m->nbkt != b) break;
    }
    return UM_NONE;
}

/* ------------------------------------------------------------------------------------------- */
/* grid subsampling                                                                            */
/* ------------------------------------------------------------------------------------------- */

/* (size_t)floor(x): x86-64 converts through a signed 64-bit truncation for |x| < 2^63, so a
 * (never expected) negative cell index wraps modulo 2^64.  Kept for bit-parity of the key. */
static uint64_t to_size_t(float f) { return (uint64_t)(int64_t)f; }

/* Votes of one voxel and one label column: (label, count) in order of first occurrence -- the unordered_map<int,int> of
 * grid_subsampling.h:22 as far as its CONTENT goes; its iteration order is rebuilt at output time (label_vote_pick). */
typedef struct { int32_t* lab; int32_t* cnt; int32_t D, cap; } votes_t;

static void votes_add(votes_t* v, int32_t label)            /* grid_subsampling.h:56-61  labels[i][*it] += 1 */
{
    for (int32_t e = 0; e < v->D; e++)
        if (v->lab[e] == label) { v->cnt[e]++; return; }
    if (v->D == v->cap) {
        v->cap = v->cap ? 2 * v->cap : 4;
        v->lab = (int32_t*)realloc(v->lab, sizeof(int32_t) * (size_t)v->cap);
        v->cnt = (int32_t*)realloc(v->cnt, sizeof(int32_t) * (size_t)v->cap);
    }
    v->lab[v->D] = label;
    v->cnt[v->D] = 1;
    v->D++;
}

/* grid_subsampling.cpp:100-101: max_element over the map = the FIRST maximal count in iteration order.  The order is obtained
 * by inserting the distinct labels, in order of first occurrence, into the literal container model above
 * (std::hash<int> is the identity: key = (size_t)label). */
static int32_t label_vote_pick(const votes_t* v)
{
    umap_t m;
    int64_t D = v->D;
    m.next = (int64_t*)malloc(sizeof(int64_t) * (size_t)D);
    m.key = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)D);
    m.bucket = (int64_t*)malloc(sizeof(int64_t));
    m.bucket[0] = UM_NONE;
    m.nbkt = 1; m.head = UM_NONE; m.size = 0; m.sched = 0; m.next_resize = 0;
    for (int64_t e = 0; e < D; e++) {
        if ((uint64_t)m.size + 1 > m.next_resize && m.sched < N_SCHEDULE) {
            um_rehash(&m, k_bucket_schedule[m.sched++]);
            m.next_resize = m.nbkt;
        }
        int64_t node = m.size++;
        m.key[node] = (uint64_t)(int64_t)v->lab[e];
        um_insert_bucket_begin(&m, m.key[node] % m.nbkt, node);
    }
    int64_t best = m.head;
    for (int64_t q = m.head; q != UM_NONE; q = m.next[q])
        if (v->cnt[best] < v->cnt[q]) best = q;
    int32_t r = v->lab[best];
    free(m.next); free(m.key); free(m.bucket);
    return r;
}

/* One cloud: grid_subsampling.cpp:5-106.  out must hold 3*n floats.  Optional per-point features f [n, fdim] -> of [M, fdim]
 * (sums in point order divided by (float)count, :88-96) and classes c [n, ldim] -> oc [M, ldim] (:97-102).  Returns M. */
static int64_t subsample_one(const float* p, int64_t n, float dl, float* out, const float* f, int32_t fdim, float* of,
                             const int32_t* c, int32_t ldim, int32_t* oc)
{
    if (n <= 0) return 0;
    /* cloud.cpp:27-66 */
    float mn[3] = { p[0], p[1], p[2] }, mx[3] = { p[0], p[1], p[2] };
    for (int64_t i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) {
            float v = p[3 * i + d];
            if (v < mn[d]) mn[d] = v;
            if (v > mx[d]) mx[d] = v;
        }
    /* grid_subsampling.cpp:27  origin = floor(min * (1/dl)) * dl   (all fp32) */
    float inv = 1 / dl;
    float org[3];
    for (int d = 0; d < 3; d++) org[d] = floorf(mn[d] * inv) * dl;
    /* :30-31 */
    uint64_t NX = to_size_t(floorf((mx[0] - org[0]) / dl)) + 1;
    uint64_t NY = to_size_t(floorf((mx[1] - org[1]) / dl)) + 1;

    umap_t m;
    m.next = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    m.key = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n);
    m.bucket = (int64_t*)malloc(sizeof(int64_t));
    m.bucket[0] = UM_NONE;
    m.nbkt = 1; m.head = UM_NONE; m.size = 0; m.sched = 0; m.next_resize = 0;
    float* sum = (float*)calloc((size_t)n * 3, sizeof(float));
    int* cnt = (int*)calloc((size_t)n, sizeof(int));
    float* fsum = f ? (float*)calloc((size_t)n * (size_t)fdim, sizeof(float)) : NULL;      /* vector<float>(fdim): zeros */
    votes_t* votes = c ? (votes_t*)calloc((size_t)n * (size_t)ldim, sizeof(votes_t)) : NULL;

    for (int64_t i = 0; i < n; i++) {
        /* :53-56  true fp32 divisions */
        uint64_t iX = to_size_t(floorf((p[3 * i + 0] - org[0]) / dl));
        uint64_t iY = to_size_t(floorf((p[3 * i + 1] - org[1]) / dl));
        uint64_t iZ = to_size_t(floorf((p[3 * i + 2] - org[2]) / dl));
        uint64_t k = iX + NX * iY + NX * NY * iZ;
        int64_t node = um_find(&m, k);
        if (node == UM_NONE) {
            /* _M_insert_unique_node: rehash check first (_M_need_rehash(bkt, size, 1)) */
            if ((uint64_t)m.size + 1 > m.next_resize && m.sched < N_SCHEDULE) {
                um_rehash(&m, k_bucket_schedule[m.sched++]);
                m.next_resize = m.nbkt;     /* floor(bucket_count * max_load_factor(1.0)) */
            }
            node = m.size++;
            m.key[node] = k;
            um_insert_bucket_begin(&m, k % m.nbkt, node);
        }
        /* grid_subsampling.h:74-79  count += 1; point += p  (fp32, original order) */
        cnt[node] += 1;
        sum[3 * node + 0] += p[3 * i + 0];
        sum[3 * node + 1] += p[3 * i + 1];
        sum[3 * node + 2] += p[3 * i + 2];
        /* grid_subsampling.h:50,67  transform(features, f_begin, plus<float>) */
        if (f) for (int32_t d = 0; d < fdim; d++) fsum[(size_t)node * fdim + d] += f[(size_t)i * fdim + d];
        if (c) for (int32_t d = 0; d < ldim; d++) votes_add(&votes[(size_t)node * ldim + d], c[(size_t)i * ldim + d]);
    }
    /* :85-87  iterate the container; point * (float)(1.0 / count) */
    int64_t o = 0;
    for (int64_t q = m.head; q != UM_NONE; q = m.next[q], o++) {
        float a = (float)(1.0 / (double)cnt[q]);
        out[3 * o + 0] = sum[3 * q + 0] * a;
        out[3 * o + 1] = sum[3 * q + 1] * a;
        out[3 * o + 2] = sum[3 * q + 2] * a;
        if (f) {                                   /* :90-95  f / (float)count */
            float fc = (float)cnt[q];
            for (int32_t d = 0; d < fdim; d++) of[(size_t)o * fdim + d] = fsum[(size_t)q * fdim + d] / fc;
        }
        if (c) for (int32_t d = 0; d < ldim; d++) oc[(size_t)o * ldim + d] = label_vote_pick(&votes[(size_t)q * ldim + d]);
    }
    if (votes) {
        for (int64_t i = 0; i < n * ldim; i++) { free(votes[i].lab); free(votes[i].cnt); }
        free(votes);
    }
    free(fsum);
    free(m.next); free(m.key); free(m.bucket); free(sum); free(cnt);
    return o;
}

/* grid_subsampling.cpp:109-211.  out_pts must hold 3*n floats.  Returns total M. */
int64_t oracle_subsample_batch(const float* pts, int64_t n, const int32_t* lens, int32_t nb,
                               float dl, int32_t max_p, float* out_pts, int32_t* out_lens)
{
    int64_t maxp = max_p < 1 ? n : max_p;   /* :134 */
    int64_t start = 0, o = 0;
    float* tmp = (float*)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
    for (int32_t b = 0; b < nb; b++) {
        int64_t m = subsample_one(pts + 3 * start, lens[b], dl, tmp, NULL, 0, NULL, NULL, 0, NULL);
        if (m > maxp) m = maxp;             /* :181-204 keep the head */
        memcpy(out_pts + 3 * o, tmp, sizeof(float) * 3 * (size_t)m);
        out_lens[b] = (int32_t)m;
        o += m;
        start += lens[b];
    }
    free(tmp);
    return o;
}

/* grid_subsampling.cpp:109-211 with features [n, fdim] and / or classes [n, ldim] (NULL = absent).  out_feats holds n*fdim
 * floats, out_classes n*ldim ints.  Returns total M, or -2 for ldim > 1 with more than one cloud: the reference slices the
 * classes of the later clouds with a wrong end offset (:157-158) and reads out of bounds, so there is nothing to restate. */
int64_t oracle_subsample_batch_ex(const float* pts, int64_t n, const int32_t* lens, int32_t nb, float dl, int32_t max_p,
                                  const float* feats, int32_t fdim, const int32_t* classes, int32_t ldim, float* out_pts,
                                  int32_t* out_lens, float* out_feats, int32_t* out_classes)
{
    if (classes && ldim > 1 && nb > 1) return -2;
    int64_t maxp = max_p < 1 ? n : max_p;
    int64_t start = 0, o = 0;
    size_t n1 = (size_t)(n > 0 ? n : 1);
    float* tmp = (float*)malloc(sizeof(float) * 3 * n1);
    float* tf = feats ? (float*)malloc(sizeof(float) * n1 * (size_t)fdim) : NULL;
    int32_t* tc = classes ? (int32_t*)malloc(sizeof(int32_t) * n1 * (size_t)ldim) : NULL;
    for (int32_t b = 0; b < nb; b++) {
        int64_t m = subsample_one(pts + 3 * start, lens[b], dl, tmp, feats ? feats + (size_t)start * fdim : NULL, fdim, tf,
                                  classes ? classes + (size_t)start * ldim : NULL, ldim, tc);
        if (m > maxp) m = maxp;             /* :181-204 keep the head of points, features and classes alike */
        memcpy(out_pts + 3 * o, tmp, sizeof(float) * 3 * (size_t)m);
        if (feats) memcpy(out_feats + (size_t)o * fdim, tf, sizeof(float) * (size_t)m * fdim);
        if (classes) memcpy(out_classes + (size_t)o * ldim, tc, sizeof(int32_t) * (size_t)m * ldim);
        out_lens[b] = (int32_t)m;
        o += m;
        start += lens[b];
    }
    free(tmp); free(tf); free(tc);
    return o;
}

/* The vote of one voxel on its own (unit tests of the CUDA vote routine). */
int32_t oracle_label_vote(const int32_t* labels, int64_t n)
{
    votes_t v = { NULL, NULL, 0, 0 };
    for (int64_t i = 0; i < n; i++) votes_add(&v, labels[i]);
    int32_t r = label_vote_pick(&v);
    free(v.lab); free(v.cnt);
    return r;
}

/* Voxel key + origin of one cloud, exposed for unit tests of the CUDA key kernel. */
void oracle_voxel_keys(const float* p, int64_t n, float dl, uint64_t* keys, float* org_out, uint64_t* nxny)
{
    float mn[3] = { p[0], p[1], p[2] }, mx[3] = { p[0], p[1], p[2] };
    for (int64_t i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) {
            float v = p[3 * i + d];
            if (v < mn[d]) mn[d] = v;
            if (v > mx[d]) mx[d] = v;
        }
    float inv = 1 / dl, org[3];
    for (int d = 0; d < 3; d++) org_out[d] = org[d] = floorf(mn[d] * inv) * dl;
    uint64_t NX = to_size_t(floorf((mx[0] - org[0]) / dl)) + 1;
    uint64_t NY = to_size_t(floorf((mx[1] - org[1]) / dl)) + 1;
    nxny[0] = NX; nxny[1] = NY;
    for (int64_t i = 0; i < n; i++) {
        uint64_t iX = to_size_t(floorf((p[3 * i + 0] - org[0]) / dl));
        uint64_t iY = to_size_t(floorf((p[3 * i + 1] - org[1]) / dl));
        uint64_t iZ = to_size_t(floorf((p[3 * i + 2] - org[2]) / dl));
        keys[i] = iX + NX * iY + NX * NY * iZ;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* radius search (canonical order)                                                             */
/* ------------------------------------------------------------------------------------------- */

typedef struct { float d2; int32_t j; } cand_t;

static int cand_cmp(const void* a, const void* b)
{
    const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
    if (x->d2 < y->d2) return -1;
    if (x->d2 > y->d2) return 1;
    return (x->j > y->j) - (x->j < y->j);      /* canonical tie rule: ascending index */
}

/* nanoflann.hpp:432-440: result = 0; for each dim: result += diff*diff  (fp32, no FMA) */
static float d2_ref(const float* a, const float* b)
{
    float r = 0.0f;
    float d0 = a[0] - b[0]; r += d0 * d0;
    float d1 = a[1] - b[1]; r += d1 * d1;
    float d2 = a[2] - b[2]; r += d2 * d2;
    return r;
}

/* Pass 1: counts[i] = |{j in cloud(i): d2 < r*r}| ; returns max count.
 * Brute force: this is the DEFINITION the kd-tree of the reference answers
 * (neighbors.cpp:268-301, nanoflann.hpp:250,1359-1362). */
int32_t oracle_radius_count(const float* q, int64_t nq, const float* s, int64_t ns,
                            const int32_t* ql, const int32_t* sl, int32_t nb, float radius,
                            int32_t* counts)
{
    (void)nq; (void)ns;
    float r2 = radius * radius;                 /* neighbors.cpp:226 */
    int64_t q0 = 0, s0 = 0;
    int32_t mx = 0;
    for (int32_t b = 0; b < nb; b++) {
        for (int64_t i = q0; i < q0 + ql[b]; i++) {
            int32_t c = 0;
            for (int64_t j = s0; j < s0 + sl[b]; j++)
                if (d2_ref(q + 3 * i, s + 3 * j) < r2) c++;
            counts[i] = c;
            if (c > mx) mx = c;
        }
        q0 += ql[b]; s0 += sl[b];
    }
    return mx;
}

/* Pass 2: rows of `width` ints: neighbours ascending (d2, index), global support indices, padded
 * with the shadow index ns (neighbors.cpp:322-324); rows longer than width are truncated (this is
 * the python-side `neighbors[:, :limit]` of datasets/dataloader.py:66-69 applied after sorting). */
void oracle_radius_fill(const float* q, int64_t nq, const float* s, int64_t ns,
                        const int32_t* ql, const int32_t* sl, int32_t nb, float radius,
                        int32_t width, int32_t* out)
{
    (void)nq;
    float r2 = radius * radius;
    int64_t q0 = 0, s0 = 0;
    int32_t maxs = 0;
    for (int32_t b = 0; b < nb; b++) if (sl[b] > maxs) maxs = sl[b];
    cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (size_t)(maxs > 0 ? maxs : 1));
    for (int32_t b = 0; b < nb; b++) {
        for (int64_t i = q0; i < q0 + ql[b]; i++) {
            int32_t k = 0;
            for (int64_t j = s0; j < s0 + sl[b]; j++) {
                float d = d2_ref(q + 3 * i, s + 3 * j);
                if (d < r2) { c[k].d2 = d; c[k].j = (int32_t)j; k++; }
            }
            qsort(c, (size_t)k, sizeof(cand_t), cand_cmp);
            for (int32_t t = 0; t < width; t++) out[i * width + t] = t < k ? c[t].j : (int32_t)ns;
        }
        q0 += ql[b]; s0 += sl[b];
    }
    free(c);
}

/* Tie canonicaliser for rows produced by the REAL reference (oracle/_ref): stable re-order of runs
 * of equal fp32 d2 by ascending index, in place, BEFORE truncation.
 * Returns the number of rows whose order changed. */
int64_t oracle_canonicalise_rows(const float* q, int64_t nq, const float* s, int64_t ns,
                                 int32_t width, int32_t* rows)
{
    int64_t changed = 0;
    cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (size_t)(width > 0 ? width : 1));
    for (int64_t i = 0; i < nq; i++) {
        int32_t k = 0;
        while (k < width && rows[i * width + k] < ns) {
            int32_t j = rows[i * width + k];
            c[k].j = j; c[k].d2 = d2_ref(q + 3 * i, s + 3 * (int64_t)j); k++;
        }
        qsort(c, (size_t)k, sizeof(cand_t), cand_cmp);
        int ch = 0;
        for (int32_t t = 0; t < k; t++) {
            if (rows[i * width + t] != c[t].j) ch = 1;
            rows[i * width + t] = c[t].j;
        }
        changed += ch;
    }
    free(c);
    return changed;
}

/* ------------------------------------------------------------------------------------------- */
/* projection.py:31-61  3D -> 2D projection with depth-consistency test                        */
/* ------------------------------------------------------------------------------------------- */

/* Row r of a 4x4 times (x,y,z,1).  torch.mm on the build host (MKL sgemm, K=4) evaluates this as
 * the fused chain  fma(m3,1, fma(m2,z, fma(m1,y, m0*x)))  -- verified bit-for-bit on 6e5 values by
 * tests/golden/make_golden.py; fmaf() is the correctly rounded single fused multiply-add. */
static float row_dot(const float* m, float x, float y, float z)
{
    float acc = m[0] * x;
    acc = fmaf(m[1], y, acc);
    acc = fmaf(m[2], z, acc);
    acc = fmaf(m[3], 1.0f, acc);
    return acc;
}

/* Returns M.  inds2d: [M,2] (x,y) int64; inds3d: [M] int64 ascending.  No z>0 test, truncation
 * toward zero (".long()"), strict |z - depth| < thresh -- all as the reference. */
int64_t oracle_projection(const float* pts, int64_t n, const float* depth, int32_t H, int32_t W,
                          const float* w2c, const float* K, float thresh,
                          int64_t* inds2d, int64_t* inds3d)
{
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++) {
        float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        float cx = row_dot(w2c + 0, x, y, z), cy = row_dot(w2c + 4, x, y, z), cz = row_dot(w2c + 8, x, y, z);
        float ix = row_dot(K + 0, cx, cy, cz), iy = row_dot(K + 4, cx, cy, cz), iz = row_dot(K + 8, cx, cy, cz);
        float fx = ix / iz, fy = iy / iz;                 /* projection.py:49 */
        /* .long(): C truncation; NaN/inf/out-of-range -> INT64_MIN like x86 cvttss2si */
        int64_t px = (fx == fx && fabsf(fx) < 9.2e18f) ? (int64_t)fx : INT64_MIN;
        int64_t py = (fy == fy && fabsf(fy) < 9.2e18f) ? (int64_t)fy : INT64_MIN;
        if (px < 0 || px >= W || py < 0 || py >= H) continue;          /* :51-53 */
        float d = depth[py * W + px];
        if (!(fabsf(iz - d) < thresh)) continue;                        /* :56 */
        inds2d[2 * m] = px; inds2d[2 * m + 1] = py; inds3d[m] = i; m++;
    }
    return m;
}
