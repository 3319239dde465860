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
#include <cstring>
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
