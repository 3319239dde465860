// TEST INFRASTRUCTURE ONLY (see oracle/README.md): extern "C" entry points onto the UNMODIFIED
// reference C++ core, compiled from the sources where they lie (cpp_wrappers.zip, unpacked to a
// temp dir by oracle/Makefile).  Nothing under pcrcg_b200/ may load the library built from this.
//
// Replaces, for ctypes, the CPython glue of the reference:
//   cpp_wrappers.zip!cpp_wrappers/cpp_subsampling/wrapper.cpp:62-333  (subsample_batch)
//   cpp_wrappers/cpp_neighbors/wrapper.cpp:58-238                     (batch_query)
// which no longer compiles against numpy 2 (NPY_IN_ARRAY) / py3.12 (numpy.distutils).
#include "cpp_subsampling/grid_subsampling/grid_subsampling.h"
#include "cpp_neighbors/neighbors/neighbors.h"
#include <algorithm>
#include <cstring>
#include <unordered_map>
#include <cstdlib>

extern "C" {

// Returns number of subsampled points M (or -1 if out capacity too small). out_pts must hold cap*3 floats.
long ref_subsample_batch(const float* pts, long n, const int* lens, int nb, float dl, int max_p,
                         float* out_pts, long cap, int* out_lens)
{
    std::vector<PointXYZ> original_points((const PointXYZ*)pts, (const PointXYZ*)pts + n);
    std::vector<int> original_batches(lens, lens + nb);
    std::vector<PointXYZ> subsampled_points;
    std::vector<float> of, sf;
    std::vector<int> oc, sc, subsampled_batches;
    batch_grid_subsampling(original_points, subsampled_points, of, sf, oc, sc,
                           original_batches, subsampled_batches, dl, max_p);
    long m = (long)subsampled_points.size();
    if (m > cap) return -1;
    std::memcpy(out_pts, subsampled_points.data(), sizeof(float) * 3 * m);
    std::memcpy(out_lens, subsampled_batches.data(), sizeof(int) * nb);
    return m;
}

// Same with per-point features [n, fdim] and / or integer classes [n, ldim] (nullptr = absent): grid_subsampling.cpp:34-102.
// out_feats holds cap*fdim floats, out_classes cap*ldim ints.  ldim > 1 with more than one cloud is refused (-2): the reference
// slices the classes of later clouds with a wrong end offset (grid_subsampling.cpp:157-158), which reads out of bounds.
long ref_subsample_batch_ex(const float* pts, long n, const int* lens, int nb, float dl, int max_p, const float* feats, int fdim,
                            const int* classes, int ldim, float* out_pts, long cap, int* out_lens, float* out_feats, int* out_classes)
{
    if (classes && ldim > 1 && nb > 1) return -2;
    std::vector<PointXYZ> original_points((const PointXYZ*)pts, (const PointXYZ*)pts + n);
    std::vector<int> original_batches(lens, lens + nb);
    std::vector<PointXYZ> subsampled_points;
    std::vector<float> of, sf;
    std::vector<int> oc, sc, subsampled_batches;
    if (feats) of.assign(feats, feats + (size_t)n * fdim);
    if (classes) oc.assign(classes, classes + (size_t)n * ldim);
    batch_grid_subsampling(original_points, subsampled_points, of, sf, oc, sc, original_batches, subsampled_batches, dl, max_p);
    long m = (long)subsampled_points.size();
    if (m > cap) return -1;
    std::memcpy(out_pts, subsampled_points.data(), sizeof(float) * 3 * m);
    std::memcpy(out_lens, subsampled_batches.data(), sizeof(int) * nb);
    if (feats) std::memcpy(out_feats, sf.data(), sizeof(float) * sf.size());
    if (classes) std::memcpy(out_classes, sc.data(), sizeof(int) * sc.size());
    return m;
}

// The vote of one voxel exactly as grid_subsampling.h:56-61 + grid_subsampling.cpp:100-101 do it with THIS libstdc++:
// labels[*it] += 1 in point order, then the first maximal element in the container's iteration order.
int ref_label_vote(const int* labels, long n)
{
    std::unordered_map<int, int> m;
    for (long i = 0; i < n; i++) m[labels[i]] += 1;
    return std::max_element(m.begin(), m.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.second < b.second; })->first;
}

// Two-call protocol: the result is kept in a static vector between the calls.
static std::vector<int> g_neighbors;

// Runs the reference search; returns max_count (row width); total ints = nq*max_count.
long ref_batch_query(const float* q, long nq, const float* s, long ns, const int* ql, const int* sl,
                     int nb, float radius)
{
    std::vector<PointXYZ> queries((const PointXYZ*)q, (const PointXYZ*)q + nq);
    std::vector<PointXYZ> supports((const PointXYZ*)s, (const PointXYZ*)s + ns);
    std::vector<int> q_batches(ql, ql + nb), s_batches(sl, sl + nb);
    g_neighbors.clear();
    batch_nanoflann_neighbors(queries, supports, q_batches, s_batches, g_neighbors, radius);
    if (nq == 0) return 0;
    return (long)(g_neighbors.size() / (size_t)nq);
}

void ref_batch_query_fetch(int* out)
{
    std::memcpy(out, g_neighbors.data(), sizeof(int) * g_neighbors.size());
    std::vector<int>().swap(g_neighbors);
}

// Growth schedule of this libstdc++'s unordered_map<size_t,...> (used to pin the prime table of
// the restatement in oracle/port.c): writes up to cap bucket counts seen while inserting n keys.
int ref_bucket_schedule(long n, long* out, int cap)
{
    std::unordered_map<size_t, int> m;
    int k = 0;
    size_t last = m.bucket_count();
    if (k < cap) out[k++] = (long)last;
    for (long i = 0; i < n; i++) {
        m.emplace((size_t)i, 0);
        if (m.bucket_count() != last) { last = m.bucket_count(); if (k < cap) out[k++] = (long)last; }
    }
    return k;
}

}
