// C ABI of libpcrcg_b200.so (see include/pcrcg_b200.h).
#include "../../include/pcrcg_b200.h"
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <vector>

namespace pcrcg {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
// ---- profiling ---------------------------------------------------------------------------------
struct ProfRec { int cls; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_event_pool;
static int g_prof_on = 0;
static unsigned long long g_launches = 0;
static std::mutex g_prof_mu;
void count_launches(int n) { g_launches += (unsigned long long)n; }
static cudaEvent_t get_event()
{
    if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void prof_begin(int cls, cudaStream_t st, int* slot)
{
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r{ cls, get_event(), get_event() };
    cudaEventRecord(r.a, st);
    g_prof.push_back(r);
    *slot = (int)g_prof.size() - 1;
}
void prof_end(int slot, cudaStream_t st)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (slot < (int)g_prof.size()) cudaEventRecord(g_prof[slot].b, st);
}
size_t subsample_ws_bytes(int64_t n, int32_t nb);
int subsample_batch_dev(const float*, int64_t, const int32_t*, int32_t, float, int32_t, float*, int32_t*, void*, size_t, cudaStream_t);
size_t subsample_ex_ws_bytes(int64_t n, int32_t nb, int32_t fdim, int32_t ldim);
int subsample_batch_ex_dev(const float*, int64_t, const int32_t*, int32_t, float, int32_t, const float*, int32_t, const int32_t*, int32_t,
                           float*, int32_t*, float*, int32_t*, int32_t*, void*, size_t, cudaStream_t);
int group_starts_dev(const int32_t*, int32_t, int32_t, int32_t*, int32_t*, cudaStream_t);
size_t radius_ws_bytes(int64_t nq, int64_t ns, int32_t nb);
int radius_build_dev(const float*, int64_t, const int32_t*, int32_t, float, void*, size_t, cudaStream_t);
size_t radius_query_ws_bytes(int64_t nq, int32_t nb);
bool radius_cells_preferred(int32_t width);
int radius_query_cells_dev(const float*, int64_t, const int32_t*, int64_t, int32_t, float, int32_t, int32_t, int32_t*, int32_t*, int32_t*,
                           void*, size_t, void*, size_t, int32_t, cudaStream_t);
int radius_query_dev(const float*, int64_t, const int32_t*, int64_t, int32_t, float, int32_t, int32_t, int32_t*, int32_t*, int32_t*,
                     void*, size_t, cudaStream_t);

size_t kpconv_ws_bytes(int64_t nq, int64_t ns, int32_t cin, int32_t K);
int kpconv_forward_dev(const float*, int64_t, const float*, int64_t, const void*, int, int32_t, int32_t, const float*, int32_t, const float*,
                       int32_t, float, const float*, int32_t, float*, void*, size_t, cudaStream_t, const void*, const void*, int32_t, const uint8_t*,
                       const int32_t*, int32_t, double*);
int gemm_tc_core_stats_dev(const void*, const void*, const void*, const void*, int, float*, int, int, int, int, const float*, cudaStream_t,
                           const int32_t*, int, double*, int64_t);
int colstats_final_dev(const double*, const int32_t*, int32_t, int32_t, float, float*, float*, cudaStream_t);
int gemm_dev(const float*, int, const float*, int, int, float*, int, int, int, int, const float*, cudaStream_t);
void gemm_set_force_simt(int);
void gemm_set_stats_dbg(int);
void dense_set_norm_v4(int);
void dense_set_norm_variant(int);
void kpconv_set_agg_simt(int);
void kpconv_set_agg_pipelined(int);
void kpconv_set_small_fused(int);
void kpconv_set_chunk_mb(int);
void kpconv_set_fused(int);
void gemm_set_prof_class(int);
int gemm_tc_core_dev(const void*, const void*, const void*, const void*, int, float*, int, int, int, int, const float*, cudaStream_t);
int split_bf16_dev(const float*, int, int64_t, int, void*, void*, int, cudaStream_t);
int colstats_dev(const float*, int64_t, int32_t, const int32_t*, int32_t, float, float*, float*, cudaStream_t);
int norm_act_dev(const float*, int64_t, int32_t, const int32_t*, int32_t, const float*, const float*, const float*, const float*,
                 const float*, float, float*, void*, void*, int32_t, uint8_t*, cudaStream_t);
int max_pool_dev(const float*, int64_t, int32_t, const void*, int, int64_t, int32_t, int32_t, float*, cudaStream_t);
int max_pool_planes_dev(const void*, const void*, int64_t, int32_t, int32_t, const void*, int, int64_t, int32_t, int32_t, void*, void*, int32_t,
                        cudaStream_t);
int norm_act_planes_dev(const float*, int64_t, int32_t, const int32_t*, int32_t, const float*, const float*, const float*, const float*,
                        const float*, float, float*, void*, void*, int32_t, uint8_t*, const void*, const void*, int32_t, cudaStream_t);
int descriptor_head_dev(const float*, int64_t, int32_t, float*, float*, float*, cudaStream_t);
int closest_pool_dev(const float*, int64_t, int32_t, const void*, int, int64_t, int32_t, float*, cudaStream_t);

int knn_dev(const float*, int64_t, const int32_t*, int32_t, int32_t, int32_t*, cudaStream_t);
int edge_max_stats_dev(const float*, int32_t, const float*, int32_t, const int32_t*, int64_t, int32_t, int32_t, const int32_t*, int32_t, float*,
                       double*, cudaStream_t);
int bias_act_dev(const float*, int64_t, int32_t, const float*, float, float*, cudaStream_t);
int point2node_dev(const float*, int64_t, const int32_t*, const float*, const int32_t*, int32_t, int32_t*, cudaStream_t);
int node_counts_dev(const int32_t*, const uint8_t*, int64_t, const int32_t*, const int32_t*, int32_t, int32_t*, int32_t*, cudaStream_t);
int softmax_rows_dev(float*, int64_t, int32_t, int32_t, float, cudaStream_t);
int l2norm_rows_dev(const float*, int64_t, int32_t, float, float*, cudaStream_t);

int best_match_dev(const float*, int64_t, const float*, int64_t, int32_t, int32_t*, float*, cudaStream_t);
int mutual_dev(const int32_t*, const int32_t*, int64_t, uint8_t*, cudaStream_t);

size_t projection_ws_bytes(int64_t n);
int projection_dev(const float*, int64_t, const float*, int32_t, int32_t, const float*, const float*, float, long long*, long long*, int32_t*,
                   void*, size_t, cudaStream_t);
int project_scatter_batch_dev(const float*, int64_t, const int32_t*, int32_t, const int32_t*, const void*, int32_t, int32_t, int32_t, float, const float*,
                              float*, cudaStream_t);
int project_scatter_dev(const float*, int64_t, int32_t, const float* const*, const float* const*, const float* const*, const float*,
                        const float*, const int32_t*, const int32_t*, int32_t, int32_t, int32_t, float, const float*, float*, cudaStream_t);

// RAII device buffer for the host entry points
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { PCRCG_CUDA(cudaMalloc(&p, bytes > 0 ? bytes : 1)); return PCRCG_OK; }
    template <class T> T* as() { return (T*)p; }
};
}  // namespace pcrcg

using namespace pcrcg;

extern "C" {

const char* pcrcg_last_error(void) { return g_err; }
int pcrcg_version(void) { return 100; }
void pcrcg_free(void* p) { free(p); }

void pcrcg_profile_enable(int32_t on) { g_prof_on = on; }
uint64_t pcrcg_launch_count(void) { return g_launches; }
int32_t pcrcg_profile_classes(void) { return PC_COUNT; }
const char* pcrcg_profile_class_name(int32_t c)
{
    static const char* names[PC_COUNT] = { "subsample", "radius_build", "radius_query", "kpconv_aggregate", "gemm", "norm_act", "pool", "projection", "kpconv_fused", "linear" };
    return (c >= 0 && c < PC_COUNT) ? names[c] : "?";
}
// Synchronises the device, sums elapsed ms and scope counts per class, clears the records.
int pcrcg_profile_report(double* ms, int64_t* counts)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int c = 0; c < PC_COUNT; c++) { ms[c] = 0.0; counts[c] = 0; }
    PCRCG_CUDA(cudaDeviceSynchronize());
    for (auto& r : g_prof) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms[r.cls] += t; counts[r.cls] += 1; }
        g_event_pool.push_back(r.a);
        g_event_pool.push_back(r.b);
    }
    g_prof.clear();
    return PCRCG_OK;
}

size_t pcrcg_subsample_ws_bytes(int64_t n, int32_t nb) { return subsample_ws_bytes(n, nb); }

int pcrcg_subsample_batch_dev(const float* points, int64_t n, const int32_t* lens, int32_t nb, float dl, int32_t max_p,
                              float* out_points, int32_t* out_lens, void* ws, size_t ws_bytes, pcrcg_stream_t stream)
{
    return subsample_batch_dev(points, n, lens, nb, dl, max_p, out_points, out_lens, ws, ws_bytes, (cudaStream_t)stream);
}

int pcrcg_subsample_batch_host(const float* points, int64_t n, const int32_t* lens, int32_t nb, float dl, int32_t max_p,
                               float** out_points, int64_t* out_m, int32_t* out_lens)
{
    PCRCG_REQUIRE(points && lens && out_points && out_m && out_lens, "subsample_batch: null argument");
    PCRCG_REQUIRE(n >= 1 && nb >= 1, "Error");   // the reference raises RuntimeError("Error") on an empty result
    DevBuf dp, dl_, dout, dol, dws;
    size_t wsb = subsample_ws_bytes(n, nb);
    PCRCG_TRY(dp.alloc(sizeof(float) * 3 * n));
    PCRCG_TRY(dl_.alloc(sizeof(int32_t) * nb));
    PCRCG_TRY(dout.alloc(sizeof(float) * 3 * n));
    PCRCG_TRY(dol.alloc(sizeof(int32_t) * nb));
    PCRCG_TRY(dws.alloc(wsb));
    PCRCG_CUDA(cudaMemcpy(dp.p, points, sizeof(float) * 3 * n, cudaMemcpyHostToDevice));
    PCRCG_CUDA(cudaMemcpy(dl_.p, lens, sizeof(int32_t) * nb, cudaMemcpyHostToDevice));
    PCRCG_TRY(subsample_batch_dev(dp.as<float>(), n, dl_.as<int32_t>(), nb, dl, max_p, dout.as<float>(), dol.as<int32_t>(),
                                  dws.p, wsb, 0));
    PCRCG_CUDA(cudaMemcpy(out_lens, dol.p, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost));
    int64_t m = 0;
    for (int b = 0; b < nb; b++) m += out_lens[b];
    PCRCG_REQUIRE(m >= 1, "Error");
    float* o = (float*)malloc(sizeof(float) * 3 * m);
    PCRCG_REQUIRE(o != nullptr, "subsample_batch: out of host memory");
    cudaError_t e = cudaMemcpy(o, dout.p, sizeof(float) * 3 * m, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { free(o); set_error("subsample_batch: D2H failed: %s", cudaGetErrorString(e)); return PCRCG_ERR; }
    *out_points = o;
    *out_m = m;
    return PCRCG_OK;
}

size_t pcrcg_subsample_ex_ws_bytes(int64_t n, int32_t nb, int32_t fdim, int32_t ldim) { return subsample_ex_ws_bytes(n, nb, fdim, ldim); }

int pcrcg_subsample_batch_ex_dev(const float* points, int64_t n, const int32_t* lens, int32_t nb, float dl, int32_t max_p,
                                 const float* features, int32_t fdim, const int32_t* classes, int32_t ldim, float* out_points,
                                 int32_t* out_lens, float* out_features, int32_t* out_classes, int32_t* status, void* ws,
                                 size_t ws_bytes, pcrcg_stream_t stream)
{
    return subsample_batch_ex_dev(points, n, lens, nb, dl, max_p, features, fdim, classes, ldim, out_points, out_lens, out_features,
                                  out_classes, status, ws, ws_bytes, (cudaStream_t)stream);
}

static int d2h_new(void** out, const void* dev, size_t bytes)
{
    void* o = malloc(bytes > 0 ? bytes : 1);
    PCRCG_REQUIRE(o != nullptr, "subsample_batch: out of host memory");
    cudaError_t e = cudaMemcpy(o, dev, bytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { free(o); set_error("subsample_batch: D2H failed: %s", cudaGetErrorString(e)); return PCRCG_ERR; }
    *out = o;
    return PCRCG_OK;
}

int pcrcg_subsample_batch_ex_host(const float* points, int64_t n, const int32_t* lens, int32_t nb, float dl, int32_t max_p,
                                  const float* features, int32_t fdim, const int32_t* classes, int32_t ldim, float** out_points,
                                  int64_t* out_m, int32_t* out_lens, float** out_features, int32_t** out_classes)
{
    PCRCG_REQUIRE(points && lens && out_points && out_m && out_lens, "subsample_batch: null argument");
    PCRCG_REQUIRE((features == nullptr || out_features != nullptr) && (classes == nullptr || out_classes != nullptr), "subsample_batch: null argument");
    PCRCG_REQUIRE(n >= 1 && nb >= 1, "Error");   // the reference raises RuntimeError("Error") on an empty result
    if (features == nullptr) fdim = 0;
    if (classes == nullptr) ldim = 0;
    DevBuf dp, dl_, dout, dol, dws, df, dc, dof, doc, dst;
    size_t wsb = subsample_ex_ws_bytes(n, nb, fdim, ldim);
    PCRCG_TRY(dp.alloc(sizeof(float) * 3 * n));
    PCRCG_TRY(dl_.alloc(sizeof(int32_t) * nb));
    PCRCG_TRY(dout.alloc(sizeof(float) * 3 * n));
    PCRCG_TRY(dol.alloc(sizeof(int32_t) * nb));
    PCRCG_TRY(dws.alloc(wsb));
    PCRCG_TRY(dst.alloc(sizeof(int32_t)));
    PCRCG_CUDA(cudaMemcpy(dp.p, points, sizeof(float) * 3 * n, cudaMemcpyHostToDevice));
    PCRCG_CUDA(cudaMemcpy(dl_.p, lens, sizeof(int32_t) * nb, cudaMemcpyHostToDevice));
    if (fdim) {
        PCRCG_TRY(df.alloc(sizeof(float) * (size_t)n * fdim));
        PCRCG_TRY(dof.alloc(sizeof(float) * (size_t)n * fdim));
        PCRCG_CUDA(cudaMemcpy(df.p, features, sizeof(float) * (size_t)n * fdim, cudaMemcpyHostToDevice));
    }
    if (ldim) {
        PCRCG_TRY(dc.alloc(sizeof(int32_t) * (size_t)n * ldim));
        PCRCG_TRY(doc.alloc(sizeof(int32_t) * (size_t)n * ldim));
        PCRCG_CUDA(cudaMemcpy(dc.p, classes, sizeof(int32_t) * (size_t)n * ldim, cudaMemcpyHostToDevice));
    }
    PCRCG_TRY(subsample_batch_ex_dev(dp.as<float>(), n, dl_.as<int32_t>(), nb, dl, max_p, fdim ? df.as<float>() : nullptr, fdim,
                                     ldim ? dc.as<int32_t>() : nullptr, ldim, dout.as<float>(), dol.as<int32_t>(), dof.as<float>(),
                                     doc.as<int32_t>(), dst.as<int32_t>(), dws.p, wsb, 0));
    PCRCG_CUDA(cudaMemcpy(out_lens, dol.p, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost));
    if (ldim) {
        int32_t status = 0;
        PCRCG_CUDA(cudaMemcpy(&status, dst.p, sizeof(int32_t), cudaMemcpyDeviceToHost));
        PCRCG_REQUIRE(status == 0, "subsample_batch: a voxel holds more than 64 distinct labels in one class column (the tie order of the "
                      "reference's unordered_map is modelled up to 64)");
    }
    int64_t m = 0;
    for (int b = 0; b < nb; b++) m += out_lens[b];
    PCRCG_REQUIRE(m >= 1, "Error");
    void *o = nullptr, *of = nullptr, *oc = nullptr;
    int rc = d2h_new(&o, dout.p, sizeof(float) * 3 * m);
    if (rc == PCRCG_OK && fdim) rc = d2h_new(&of, dof.p, sizeof(float) * (size_t)m * fdim);
    if (rc == PCRCG_OK && ldim) rc = d2h_new(&oc, doc.p, sizeof(int32_t) * (size_t)m * ldim);
    if (rc != PCRCG_OK) { free(o); free(of); free(oc); return rc; }
    *out_points = (float*)o;
    *out_m = m;
    if (fdim) *out_features = (float*)of;
    if (ldim) *out_classes = (int32_t*)oc;
    return PCRCG_OK;
}

int pcrcg_group_starts_dev(const int32_t* lens, int32_t nb, int32_t group, int32_t* out, int32_t* total, pcrcg_stream_t stream)
{
    return group_starts_dev(lens, nb, group, out, total, (cudaStream_t)stream);
}

size_t pcrcg_radius_ws_bytes(int64_t nq, int64_t ns, int32_t nb) { return radius_ws_bytes(nq, ns, nb); }

int pcrcg_radius_build_dev(const float* supports, int64_t ns, const int32_t* s_lens, int32_t nb, float radius, void* ws,
                           size_t ws_bytes, pcrcg_stream_t stream)
{
    return radius_build_dev(supports, ns, s_lens, nb, radius, ws, ws_bytes, (cudaStream_t)stream);
}

int pcrcg_radius_query_dev(const float* queries, int64_t nq, const int32_t* q_lens, int64_t ns, int32_t nb, float radius,
                           int32_t width, int32_t row_stride, int32_t* rows, int32_t* counts, int32_t* max_count, void* ws,
                           size_t ws_bytes, pcrcg_stream_t stream)
{
    return radius_query_dev(queries, nq, q_lens, ns, nb, radius, width, row_stride, rows, counts, max_count, ws, ws_bytes,
                            (cudaStream_t)stream);
}

size_t pcrcg_radius_query_ws_bytes(int64_t nq, int32_t nb) { return radius_query_ws_bytes(nq, nb); }

int pcrcg_radius_query_cells_dev(const float* queries, int64_t nq, const int32_t* q_lens, int64_t ns, int32_t nb, float radius,
                                 int32_t width, int32_t row_stride, int32_t* rows, int32_t* counts, int32_t* max_count, void* ws,
                                 size_t ws_bytes, void* qws, size_t qws_bytes, int32_t queries_are_supports, pcrcg_stream_t stream)
{
    if (!radius_cells_preferred(width))          // wide lists / count-only passes: one warp per query (see radius.cu)
        return radius_query_dev(queries, nq, q_lens, ns, nb, radius, width, row_stride, rows, counts, max_count, ws, ws_bytes,
                                (cudaStream_t)stream);
    return radius_query_cells_dev(queries, nq, q_lens, ns, nb, radius, width, row_stride, rows, counts, max_count, ws, ws_bytes, qws,
                                  qws_bytes, queries_are_supports, (cudaStream_t)stream);
}

int pcrcg_batch_query_host(const float* queries, int64_t nq, const float* supports, int64_t ns, const int32_t* q_lens,
                           const int32_t* s_lens, int32_t nb, float radius, int32_t limit, int32_t** out_rows,
                           int32_t* out_width)
{
    PCRCG_REQUIRE(queries && supports && q_lens && s_lens && out_rows && out_width, "batch_query: null argument");
    PCRCG_REQUIRE(nq >= 1 && ns >= 1 && nb >= 1, "Error");
    DevBuf dq, ds, dql, dsl, dws, dqws, dmax, drows;
    size_t wsb = radius_ws_bytes(nq, ns, nb), qwsb = radius_query_ws_bytes(nq, nb);
    PCRCG_TRY(dqws.alloc(qwsb));
    PCRCG_TRY(dq.alloc(sizeof(float) * 3 * nq));
    PCRCG_TRY(ds.alloc(sizeof(float) * 3 * ns));
    PCRCG_TRY(dql.alloc(sizeof(int32_t) * nb));
    PCRCG_TRY(dsl.alloc(sizeof(int32_t) * nb));
    PCRCG_TRY(dws.alloc(wsb));
    PCRCG_TRY(dmax.alloc(sizeof(int32_t)));
    PCRCG_CUDA(cudaMemcpy(dq.p, queries, sizeof(float) * 3 * nq, cudaMemcpyHostToDevice));
    PCRCG_CUDA(cudaMemcpy(ds.p, supports, sizeof(float) * 3 * ns, cudaMemcpyHostToDevice));
    PCRCG_CUDA(cudaMemcpy(dql.p, q_lens, sizeof(int32_t) * nb, cudaMemcpyHostToDevice));
    PCRCG_CUDA(cudaMemcpy(dsl.p, s_lens, sizeof(int32_t) * nb, cudaMemcpyHostToDevice));
    PCRCG_TRY(radius_build_dev(ds.as<float>(), ns, dsl.as<int32_t>(), nb, radius, dws.p, wsb, 0));
    PCRCG_TRY(radius_query_dev(dq.as<float>(), nq, dql.as<int32_t>(), ns, nb, radius, 0, 0, nullptr, nullptr, dmax.as<int32_t>(),
                               dws.p, wsb, 0));
    int32_t mx = 0;
    PCRCG_CUDA(cudaMemcpy(&mx, dmax.p, sizeof(int32_t), cudaMemcpyDeviceToHost));
    PCRCG_REQUIRE(mx >= 1, "Error");      // cpp_neighbors/wrapper.cpp:201-205
    int32_t width = (limit > 0 && limit < mx) ? limit : mx;
    PCRCG_TRY(drows.alloc(sizeof(int32_t) * (size_t)nq * width));
    PCRCG_TRY(pcrcg_radius_query_cells_dev(dq.as<float>(), nq, dql.as<int32_t>(), ns, nb, radius, width, width, drows.as<int32_t>(), nullptr,
                                           nullptr, dws.p, wsb, dqws.p, qwsb, 0, 0));
    int32_t* o = (int32_t*)malloc(sizeof(int32_t) * (size_t)nq * width);
    PCRCG_REQUIRE(o != nullptr, "batch_query: out of host memory");
    cudaError_t e = cudaMemcpy(o, drows.p, sizeof(int32_t) * (size_t)nq * width, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { free(o); set_error("batch_query: D2H failed: %s", cudaGetErrorString(e)); return PCRCG_ERR; }
    *out_rows = o;
    *out_width = width;
    return PCRCG_OK;
}

size_t pcrcg_kpconv_ws_bytes(int64_t nq, int64_t ns, int32_t cin, int32_t K) { return kpconv_ws_bytes(nq, ns, cin, K); }

int pcrcg_kpconv_forward_dev(const float* q_pts, int64_t nq, const float* s_pts, int64_t ns, const void* neighb_inds, int32_t idx_is_i64,
                             int32_t H, int32_t idx_stride, const float* x, int32_t cin, const float* kernel_points, int32_t K,
                             float KP_extent, const float* weights, int32_t cout, float* out, void* ws, size_t ws_bytes,
                             pcrcg_stream_t stream)
{
    return kpconv_forward_dev(q_pts, nq, s_pts, ns, neighb_inds, idx_is_i64, H, idx_stride, x, cin, kernel_points, K, KP_extent, weights,
                              cout, out, ws, ws_bytes, (cudaStream_t)stream, nullptr, nullptr, 0, nullptr, nullptr, 0, nullptr);
}

int pcrcg_kpconv_forward_split_dev(const float* q_pts, int64_t nq, const float* s_pts, int64_t ns, const void* neighb_inds,
                                   int32_t idx_is_i64, int32_t H, int32_t idx_stride, const float* x, const void* x_hi, const void* x_lo,
                                   int32_t ldxs, const uint8_t* row_positive, int32_t cin, const float* kernel_points, int32_t K,
                                   float KP_extent, const float* weights, int32_t cout, float* out, void* ws, size_t ws_bytes,
                                   pcrcg_stream_t stream)
{
    return kpconv_forward_dev(q_pts, nq, s_pts, ns, neighb_inds, idx_is_i64, H, idx_stride, x, cin, kernel_points, K, KP_extent, weights,
                              cout, out, ws, ws_bytes, (cudaStream_t)stream, x_hi, x_lo, ldxs, row_positive, nullptr, 0, nullptr);
}

int pcrcg_kpconv_forward_stats_dev(const float* q_pts, int64_t nq, const float* s_pts, int64_t ns, const void* neighb_inds,
                                   int32_t idx_is_i64, int32_t H, int32_t idx_stride, const float* x, const void* x_hi, const void* x_lo,
                                   int32_t ldxs, const uint8_t* row_positive, int32_t cin, const float* kernel_points, int32_t K,
                                   float KP_extent, const float* weights, int32_t cout, float* out, void* ws, size_t ws_bytes,
                                   const int32_t* seg_starts, int32_t nseg, double* stats_acc, pcrcg_stream_t stream)
{
    if (stats_acc != nullptr && nseg >= 1) PCRCG_CUDA(cudaMemsetAsync(stats_acc, 0, sizeof(double) * 2 * (size_t)nseg * cout, (cudaStream_t)stream));
    return kpconv_forward_dev(q_pts, nq, s_pts, ns, neighb_inds, idx_is_i64, H, idx_stride, x, cin, kernel_points, K, KP_extent, weights,
                              cout, out, ws, ws_bytes, (cudaStream_t)stream, x_hi, x_lo, ldxs, row_positive, seg_starts, nseg, stats_acc);
}

int pcrcg_gemm_dev(const float* A, int32_t lda, const float* B, int32_t ldb, int32_t b_is_nk, float* C, int32_t ldc, int32_t M, int32_t N,
                   int32_t K, const float* row_scale, pcrcg_stream_t stream)
{
    gemm_set_prof_class(PC_LINEAR);
    const int rc = gemm_dev(A, lda, B, ldb, b_is_nk, C, ldc, M, N, K, row_scale, (cudaStream_t)stream);
    gemm_set_prof_class(PC_GEMM);
    return rc;
}

void pcrcg_gemm_force_simt(int32_t on) { gemm_set_force_simt(on); }

int pcrcg_set_option(const char* name, int32_t value)
{
    if (!strcmp(name, "contraction_simt")) gemm_set_force_simt(value);
    else if (!strcmp(name, "aggregate_simt")) kpconv_set_agg_simt(value);
    else if (!strcmp(name, "aggregate_pipelined")) kpconv_set_agg_pipelined(value);
    else if (!strcmp(name, "first_layer_fused")) kpconv_set_small_fused(value);
    else if (!strcmp(name, "kpconv_chunk_mb")) kpconv_set_chunk_mb(value);
    else if (!strcmp(name, "kpconv_fused")) kpconv_set_fused(value);
    else if (!strcmp(name, "stats_debug")) gemm_set_stats_dbg(value);
    else if (!strcmp(name, "norm_variant")) dense_set_norm_variant(value);
    else if (!strcmp(name, "norm_vectorised")) dense_set_norm_v4(value);
    else { set_error("pcrcg_set_option: unknown option '%s'", name); return PCRCG_ERR; }
    return PCRCG_OK;
}

int pcrcg_split_bf16_dev(const float* x, int32_t ldx, int64_t rows, int32_t cols, void* hi, void* lo, int32_t ldo, pcrcg_stream_t stream)
{
    return split_bf16_dev(x, ldx, rows, cols, hi, lo, ldo, (cudaStream_t)stream);
}

int pcrcg_gemm_bf16x3_dev(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, int32_t ldk, float* C, int32_t ldc,
                          int32_t M, int32_t N, int32_t K, const float* row_scale, pcrcg_stream_t stream)
{
    ProfScope prof(PC_LINEAR, (cudaStream_t)stream, 0);
    return gemm_tc_core_dev(a_hi, a_lo, b_hi, b_lo, ldk, C, ldc, M, N, K, row_scale, (cudaStream_t)stream);
}

int pcrcg_gemm_bf16x3_stats_dev(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, int32_t ldk, float* C, int32_t ldc,
                                int32_t M, int32_t N, int32_t K, const float* row_scale, const int32_t* seg_starts, int32_t nseg,
                                double* stats_acc, pcrcg_stream_t stream)
{
    ProfScope prof(PC_LINEAR, (cudaStream_t)stream, 0);
    if (stats_acc != nullptr && nseg >= 1) PCRCG_CUDA(cudaMemsetAsync(stats_acc, 0, sizeof(double) * 2 * (size_t)nseg * N, (cudaStream_t)stream));
    return gemm_tc_core_stats_dev(a_hi, a_lo, b_hi, b_lo, ldk, C, ldc, M, N, K, row_scale, (cudaStream_t)stream, seg_starts, nseg, stats_acc, 0);
}

int pcrcg_colstats_final_dev(const double* stats_acc, const int32_t* seg_starts, int32_t nseg, int32_t C, float eps, float* mean, float* rstd,
                             pcrcg_stream_t stream)
{
    return colstats_final_dev(stats_acc, seg_starts, nseg, C, eps, mean, rstd, (cudaStream_t)stream);
}

int pcrcg_colstats_dev(const float* x, int64_t n, int32_t C, const int32_t* seg_starts, int32_t nseg, float eps, float* mean, float* rstd,
                       pcrcg_stream_t stream)
{
    return colstats_dev(x, n, C, seg_starts, nseg, eps, mean, rstd, (cudaStream_t)stream);
}

int pcrcg_norm_act_dev(const float* x, int64_t n, int32_t C, const int32_t* seg_starts, int32_t nseg, const float* mean, const float* rstd,
                       const float* sc, const float* sc_mean, const float* sc_rstd, float slope, float* out, void* split_hi, void* split_lo,
                       int32_t split_ld, uint8_t* row_positive, pcrcg_stream_t stream)
{
    return norm_act_dev(x, n, C, seg_starts, nseg, mean, rstd, sc, sc_mean, sc_rstd, slope, out, split_hi, split_lo, split_ld, row_positive,
                        (cudaStream_t)stream);
}

int pcrcg_norm_act_planes_dev(const float* x, int64_t n, int32_t C, const int32_t* seg_starts, int32_t nseg, const float* mean,
                              const float* rstd, const void* sc_hi, const void* sc_lo, int32_t sc_ld, const float* sc_mean, const float* sc_rstd,
                              float slope, float* out, void* split_hi, void* split_lo, int32_t split_ld, uint8_t* row_positive,
                              pcrcg_stream_t stream)
{
    return norm_act_planes_dev(x, n, C, seg_starts, nseg, mean, rstd, nullptr, sc_mean, sc_rstd, slope, out, split_hi, split_lo, split_ld,
                               row_positive, sc_hi, sc_lo, sc_ld, (cudaStream_t)stream);
}

int pcrcg_max_pool_planes_dev(const void* x_hi, const void* x_lo, int64_t ns, int32_t C, int32_t ldx, const void* inds, int32_t idx_is_i64,
                              int64_t nq, int32_t H, int32_t idx_stride, void* out_hi, void* out_lo, int32_t ldo, pcrcg_stream_t stream)
{
    return max_pool_planes_dev(x_hi, x_lo, ns, C, ldx, inds, idx_is_i64, nq, H, idx_stride, out_hi, out_lo, ldo, (cudaStream_t)stream);
}

int pcrcg_descriptor_head_dev(const float* x, int64_t n, int32_t F, float* feats, float* overlap, float* saliency, pcrcg_stream_t stream)
{
    return descriptor_head_dev(x, n, F, feats, overlap, saliency, (cudaStream_t)stream);
}

int pcrcg_max_pool_dev(const float* x, int64_t ns, int32_t C, const void* inds, int32_t idx_is_i64, int64_t nq, int32_t H, int32_t idx_stride,
                       float* out, pcrcg_stream_t stream)
{
    return max_pool_dev(x, ns, C, inds, idx_is_i64, nq, H, idx_stride, out, (cudaStream_t)stream);
}

int pcrcg_closest_pool_dev(const float* x, int64_t ns, int32_t C, const void* inds, int32_t idx_is_i64, int64_t nq, int32_t idx_stride,
                           float* out, pcrcg_stream_t stream)
{
    return closest_pool_dev(x, ns, C, inds, idx_is_i64, nq, idx_stride, out, (cudaStream_t)stream);
}

int pcrcg_knn_dev(const float* points, int64_t n, const int32_t* cloud_starts, int32_t nb, int32_t k, int32_t* out, pcrcg_stream_t stream)
{
    return knn_dev(points, n, cloud_starts, nb, k, out, (cudaStream_t)stream);
}

int pcrcg_edge_max_stats_dev(const float* u, int32_t ldu, const float* v, int32_t ldv, const int32_t* knn, int64_t n, int32_t C, int32_t k,
                             const int32_t* cloud_starts, int32_t nb, float* out, double* stats_acc, pcrcg_stream_t stream)
{
    if (stats_acc != nullptr && nb >= 1) PCRCG_CUDA(cudaMemsetAsync(stats_acc, 0, sizeof(double) * 2 * (size_t)nb * C, (cudaStream_t)stream));
    return edge_max_stats_dev(u, ldu, v, ldv, knn, n, C, k, cloud_starts, nb, out, stats_acc, (cudaStream_t)stream);
}

int pcrcg_point2node_dev(const float* points, int64_t n, const int32_t* point_starts, const float* nodes, const int32_t* node_starts, int32_t nb,
                         int32_t* out, pcrcg_stream_t stream)
{
    return point2node_dev(points, n, point_starts, nodes, node_starts, nb, out, (cudaStream_t)stream);
}

int pcrcg_node_counts_dev(const int32_t* point2node, const uint8_t* visible, int64_t n, const int32_t* point_starts,
                          const int32_t* node_starts, int32_t nb, int32_t* total, int32_t* visible_count, pcrcg_stream_t stream)
{
    return node_counts_dev(point2node, visible, n, point_starts, node_starts, nb, total, visible_count, (cudaStream_t)stream);
}

int pcrcg_bias_act_dev(const float* x, int64_t n, int32_t C, const float* bias, float slope, float* out, pcrcg_stream_t stream)
{
    return bias_act_dev(x, n, C, bias, slope, out, (cudaStream_t)stream);
}

int pcrcg_softmax_rows_dev(float* x, int64_t n, int32_t m, int32_t ld, float scale, pcrcg_stream_t stream)
{
    return softmax_rows_dev(x, n, m, ld, scale, (cudaStream_t)stream);
}

int pcrcg_l2norm_rows_dev(const float* x, int64_t n, int32_t C, float eps, float* out, pcrcg_stream_t stream)
{
    return l2norm_rows_dev(x, n, C, eps, out, (cudaStream_t)stream);
}

int pcrcg_best_match_dev(const float* a, int64_t n, const float* b, int64_t m, int32_t D, int32_t* best_idx, float* best_val,
                         pcrcg_stream_t stream)
{
    return best_match_dev(a, n, b, m, D, best_idx, best_val, (cudaStream_t)stream);
}

int pcrcg_mutual_dev(const int32_t* row_best, const int32_t* col_best, int64_t n, uint8_t* mutual, pcrcg_stream_t stream)
{
    return mutual_dev(row_best, col_best, n, mutual, (cudaStream_t)stream);
}

size_t pcrcg_projection_ws_bytes(int64_t n) { return projection_ws_bytes(n); }

int pcrcg_projection_dev(const float* points, int64_t n, const float* depth, int32_t H, int32_t W, const float* world2camera,
                         const float* intrinsics, float thresh, int64_t* inds2d, int64_t* inds3d, int32_t* count, void* ws, size_t ws_bytes,
                         pcrcg_stream_t stream)
{
    return projection_dev(points, n, depth, H, W, world2camera, intrinsics, thresh, (long long*)inds2d, (long long*)inds3d, count, ws,
                          ws_bytes, (cudaStream_t)stream);
}

int pcrcg_project_scatter_dev(const float* points, int64_t n, int32_t nviews, const float* const* depth, const float* const* feat,
                              const float* const* valid, const float* w2c, const float* k4, const int32_t* row_lo, const int32_t* row_hi,
                              int32_t H, int32_t W, int32_t C, float thresh, const float* base, float* out, pcrcg_stream_t stream)
{
    return project_scatter_dev(points, n, nviews, depth, feat, valid, w2c, k4, row_lo, row_hi, H, W, C, thresh, base, out,
                               (cudaStream_t)stream);
}

int pcrcg_project_scatter_batch_dev(const float* points, int64_t n, const int32_t* cloud_starts, int32_t nb, const int32_t* view_starts,
                                    const void* views, int32_t H, int32_t W, int32_t C, float thresh, const float* base, float* out,
                                    pcrcg_stream_t stream)
{
    return project_scatter_batch_dev(points, n, cloud_starts, nb, view_starts, views, H, W, C, thresh, base, out, (cudaStream_t)stream);
}

}  // extern "C"
