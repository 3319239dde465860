// 3D -> 2D projection with depth-consistency test, and the fused un-projection of 2D image
// features onto the 3D points -- sm_100a.
//
// Replaces projection.py:31-61 (Projection.projection) and the gather/scatter of
// models/architectures.py:273-307,360-370.  Arithmetic contract (oracle/port.c):
//   cam = W2C*[p;1], img = K4*[cam;1], each row evaluated as fma(m3,1,fma(m2,z,fma(m1,y,m0*x)))
//   (what torch.mm/MKL computes), IEEE divide, truncation toward zero, 0<=x<W, 0<=y<H,
//   |img.z - depth[y,x]| < thresh (strict), no z>0 test.  inds3d ascending.
#include "common.cuh"

namespace pcrcg {

struct Mat34 { float m[12]; };      // rows 0..2 of a 4x4

__device__ __forceinline__ float row_dot(const float* m, float x, float y, float z)
{
    float acc = __fmul_rn(m[0], x);
    acc = __fmaf_rn(m[1], y, acc);
    acc = __fmaf_rn(m[2], z, acc);
    acc = __fmaf_rn(m[3], 1.0f, acc);
    return acc;
}

// -> pixel (px,py) and hit flag
__device__ __forceinline__ bool project_point(const Mat34& w2c, const Mat34& k4, float x, float y, float z, const float* __restrict__ depth,
                                              int H, int W, float thresh, int& px, int& py)
{
    float cx = row_dot(w2c.m + 0, x, y, z), cy = row_dot(w2c.m + 4, x, y, z), cz = row_dot(w2c.m + 8, x, y, z);
    float ix = row_dot(k4.m + 0, cx, cy, cz), iy = row_dot(k4.m + 4, cx, cy, cz), iz = row_dot(k4.m + 8, cx, cy, cz);
    float fx = __fdiv_rn(ix, iz), fy = __fdiv_rn(iy, iz);
    if (!(fx == fx) || !(fy == fy) || !(fabsf(fx) < 9.2e18f) || !(fabsf(fy) < 9.2e18f)) return false;
    long long lx = (long long)fx, ly = (long long)fy;          // .long(): truncation toward zero
    if (lx < 0 || lx >= W || ly < 0 || ly >= H) return false;
    float d = depth[(int)ly * W + (int)lx];
    if (!(fabsf(__fsub_rn(iz, d)) < thresh)) return false;
    px = (int)lx; py = (int)ly;
    return true;
}

__global__ void __launch_bounds__(256) k_project_flags(const float* __restrict__ pts, int n, const float* __restrict__ depth, int H, int W,
                                                       Mat34 w2c, Mat34 k4, float thresh, uint32_t* __restrict__ flag,
                                                       uint32_t* __restrict__ pix)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int px = 0, py = 0;
    bool hit = project_point(w2c, k4, pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2], depth, H, W, thresh, px, py);
    flag[i] = hit ? 1u : 0u;
    pix[i] = ((uint32_t)py << 16) | (uint32_t)px;
}

__global__ void __launch_bounds__(256) k_project_compact(const uint32_t* __restrict__ pos, const uint32_t* __restrict__ pix, int n,
                                                         long long* __restrict__ inds2d, long long* __restrict__ inds3d,
                                                         int32_t* __restrict__ count)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *count = (int32_t)pos[n];
    if (i >= n) return;
    if (pos[i + 1] != pos[i]) {
        uint32_t o = pos[i], p = pix[i];
        inds2d[2 * (size_t)o] = (long long)(p & 0xffffu);
        inds2d[2 * (size_t)o + 1] = (long long)(p >> 16);
        inds3d[o] = i;
    }
}

constexpr int MAX_VIEWS = 8;
struct ViewSet {
    const float* depth[MAX_VIEWS];
    const float* feat[MAX_VIEWS];     // [C,H,W]
    const float* valid[MAX_VIEWS];    // [H,W] or nullptr
    Mat34 w2c[MAX_VIEWS];
    Mat34 k4[MAX_VIEWS];
    int32_t row_lo[MAX_VIEWS], row_hi[MAX_VIEWS];   // the view applies to points [row_lo,row_hi) (its cloud)
    int nviews;
};

// One warp per point: the LAST view (in write order) that sees the point wins (architectures.py:367-370),
// its C features * valid + a trailing 1 form the row; unseen points keep base[i] (default 1) everywhere.
__global__ void __launch_bounds__(256) k_project_scatter(const float* __restrict__ pts, int n, ViewSet vs, int H, int W, int C, float thresh,
                                                         const float* __restrict__ base, float* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const float x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
    int win = -1, px = 0, py = 0;
    for (int v = vs.nviews - 1; v >= 0; v--) {
        if (i < vs.row_lo[v] || i >= vs.row_hi[v]) continue;
        if (project_point(vs.w2c[v], vs.k4[v], x, y, z, vs.depth[v], H, W, thresh, px, py)) { win = v; break; }
    }
    float* o = out + (size_t)i * (C + 1);
    if (win < 0) {
        const float b = base ? base[i] : 1.0f;
        for (int c = lane; c <= C; c += 32) o[c] = b;
        return;
    }
    const float* f = vs.feat[win] + (size_t)py * W + px;
    const float vm = vs.valid[win] ? vs.valid[win][(size_t)py * W + px] : 1.0f;
    const size_t plane = (size_t)H * W;
    for (int c = lane; c < C; c += 32) o[c] = vs.valid[win] ? __fmul_rn(f[c * plane], vm) : f[c * plane];
    if (lane == 0) o[C] = 1.0f;
}

// Any number of clouds and views (a stacked batch of fragment pairs): the views live in DEVICE memory, grouped by cloud in
// the reference's write order (views [view_starts[c], view_starts[c+1]) belong to cloud c = rows [cloud_starts[c],
// cloud_starts[c+1])); as above the last view of its cloud that sees a point wins.
struct ViewDev {
    const float* depth;   // [H,W]
    const float* feat;    // [C,H,W]
    const float* valid;   // [H,W] or nullptr
    Mat34 w2c, k4;
};

__global__ void __launch_bounds__(256) k_project_scatter_batch(const float* __restrict__ pts, int n, const int32_t* __restrict__ cloud_starts, int nb,
                                                               const int32_t* __restrict__ view_starts, const ViewDev* __restrict__ views, int H,
                                                               int W, int C, float thresh, const float* __restrict__ base, float* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const float x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
    const int c = cloud_of(cloud_starts, nb, i);
    int win = -1, px = 0, py = 0;
    for (int v = view_starts[c + 1] - 1; v >= view_starts[c]; v--) {
        if (project_point(views[v].w2c, views[v].k4, x, y, z, views[v].depth, H, W, thresh, px, py)) { win = v; break; }
    }
    float* o = out + (size_t)i * (C + 1);
    if (win < 0) {
        const float b = base ? base[i] : 1.0f;
        for (int ch = lane; ch <= C; ch += 32) o[ch] = b;
        return;
    }
    const float* f = views[win].feat + (size_t)py * W + px;
    const float* vmp = views[win].valid;
    const float vm = vmp ? vmp[(size_t)py * W + px] : 1.0f;
    const size_t plane = (size_t)H * W;
    for (int ch = lane; ch < C; ch += 32) o[ch] = vmp ? __fmul_rn(f[ch * plane], vm) : f[ch * plane];
    if (lane == 0) o[C] = 1.0f;
}

int project_scatter_batch_dev(const float* pts, int64_t n, const int32_t* cloud_starts, int32_t nb, const int32_t* view_starts,
                              const void* views, int32_t H, int32_t W, int32_t C, float thresh, const float* base, float* out, cudaStream_t st)
{
    PCRCG_REQUIRE(n >= 0 && n < (1ll << 30) && nb >= 1 && H >= 1 && W >= 1 && C >= 1, "project_scatter: bad dimensions");
    PCRCG_REQUIRE(cloud_starts != nullptr && view_starts != nullptr && views != nullptr, "project_scatter: null argument");
    if (n == 0) return PCRCG_OK;
    ProfScope prof(PC_PROJECT, st, 1);
    k_project_scatter_batch<<<(unsigned)cdiv64(n, 8), 256, 0, st>>>(pts, (int)n, cloud_starts, nb, view_starts, (const ViewDev*)views, H, W, C, thresh,
                                                                    base, out);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

size_t projection_ws_bytes(int64_t n) { return align_up((size_t)(n + 1) * 4, 256) * 2 + scan_ws_bytes(n) + 1024; }

static void to34(const float* m16, Mat34& o) { for (int k = 0; k < 12; k++) o.m[k] = m16[k]; }

// w2c / k4: HOST pointers to 16 floats (row-major 4x4)
int projection_dev(const float* pts, int64_t n, const float* depth, int32_t H, int32_t W, const float* w2c, const float* k4, float thresh,
                   long long* inds2d, long long* inds3d, int32_t* count, void* ws, size_t ws_bytes, cudaStream_t st)
{
    PCRCG_REQUIRE(n >= 0 && n < (1ll << 30) && H >= 1 && W >= 1 && H < 65536 && W < 65536, "projection: bad dimensions");
    Workspace Wk(ws, ws_bytes);
    uint32_t* flag = Wk.take<uint32_t>((size_t)n + 1);
    uint32_t* pix = Wk.take<uint32_t>((size_t)n + 1);
    Wk.off = align_up(Wk.off, 256);
    PCRCG_REQUIRE(ws != nullptr && Wk.off + scan_ws_bytes(n) <= ws_bytes + 256, "projection: workspace too small");
    ProfScope prof(PC_PROJECT, st, 5);
    Mat34 a, b;
    to34(w2c, a);
    to34(k4, b);
    const unsigned g = (unsigned)cdiv64(n > 0 ? n : 1, 256);
    k_project_flags<<<g, 256, 0, st>>>(pts, (int)n, depth, H, W, a, b, thresh, flag, pix);
    PCRCG_CUDA(cudaGetLastError());
    PCRCG_TRY(exclusive_scan_u32(flag, flag, n, Wk.base + Wk.off, ws_bytes - Wk.off, st));
    k_project_compact<<<g, 256, 0, st>>>(flag, pix, (int)n, inds2d, inds3d, count);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

// Arrays of nviews entries; matrices are HOST pointers (nviews x 16 floats); row ranges select the cloud of each view.
int project_scatter_dev(const float* pts, int64_t n, int32_t nviews, const float* const* depth, const float* const* feat,
                        const float* const* valid, const float* w2c, const float* k4, const int32_t* row_lo, const int32_t* row_hi,
                        int32_t H, int32_t W, int32_t C, float thresh, const float* base, float* out, cudaStream_t st)
{
    PCRCG_REQUIRE(nviews >= 0 && nviews <= MAX_VIEWS, "project_scatter: at most %d views", MAX_VIEWS);
    PCRCG_REQUIRE(n >= 0 && n < (1ll << 30) && H >= 1 && W >= 1 && C >= 1, "project_scatter: bad dimensions");
    if (n == 0) return PCRCG_OK;
    ViewSet vs;
    vs.nviews = nviews;
    for (int v = 0; v < nviews; v++) {
        vs.depth[v] = depth[v]; vs.feat[v] = feat[v]; vs.valid[v] = valid ? valid[v] : nullptr;
        to34(w2c + 16 * v, vs.w2c[v]);
        to34(k4 + 16 * v, vs.k4[v]);
        vs.row_lo[v] = row_lo[v]; vs.row_hi[v] = row_hi[v];
    }
    ProfScope prof(PC_PROJECT, st, 1);
    k_project_scatter<<<(unsigned)cdiv64(n, 8), 256, 0, st>>>(pts, (int)n, vs, H, W, C, thresh, base, out);
    PCRCG_CUDA(cudaGetLastError());
    return PCRCG_OK;
}

}  // namespace pcrcg
