// Thread bodies of the three kernels behind subsample_batch(features=, classes=) (grid_subsampling.cpp:34-102): per-voxel feature
// means, per-voxel class votes and the gather of both into the reference's output order.  They work on what the plain pipeline
// of subsample.cu leaves in its workspace:
//   sslot / sidx  the points sorted (stably) by voxel slot: a voxel is a run of equal sslot, its points in ORIGINAL order
//   rank[i]       number of voxel first-occurrences among the points before i (n + 1 entries): rank[first point of a voxel] is
//                 the voxel's first-occurrence rank u, rank[starts[c]] the number of voxels of the clouds before c
//   seqA / seqB   k_order's two list buffers: it starts with L = seqA and swaps the two after every rehash epoch, so the final
//                 list of a cloud (voxel ranks within the cloud, in output order) lies in seqB after an odd number of epochs
// Plain C++ (PCRCG_HD): nvcc compiles them into k_bary_feat / k_label_vote / k_gather_extra, and tests/test_subsample_extras_host.py
// compiles the very same bodies with g++ and runs them "thread" by "thread" against the CPU oracle (test infrastructure: the
// product only ever runs them on the GPU).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "label_vote.h"

#ifdef __CUDA_ARCH__
#define PCRCG_FADD(a, b) __fadd_rn((a), (b))  /* device pass: immune to FMA contraction */
#define PCRCG_FDIV(a, b) __fdiv_rn((a), (b))
#else
#define PCRCG_FADD(a, b) ((a) + (b))          /* host pass (only the g++ test build ever runs it: -ffp-contract=off) */
#define PCRCG_FDIV(a, b) ((a) / (b))
#endif

namespace pcrcg {

// grid_subsampling.h:50,67 (features += f, point order) + .cpp:88-96 (f / (float)count).  Thread j = position in the sorted
// order (works only at a run head), feature columns d0, d0 + dstep, ...
PCRCG_HD void bary_feat_thread(int j, int d0, int dstep, const float* feat, int fdim, const uint32_t* sslot, const uint32_t* sidx,
                               int n, const uint32_t* rank, float* featU)
{
    if (j >= n) return;
    const uint32_t s = sslot[j];
    if (j != 0 && sslot[j - 1] == s) return;
    const size_t u = rank[sidx[j]];
    for (int d = d0; d < fdim; d += dstep) {
        float sum = 0.f;
        int cnt = 0;
        for (int t = j; t < n && sslot[t] == s; t++) {
            sum = PCRCG_FADD(sum, feat[(size_t)sidx[t] * fdim + d]);
            cnt++;
        }
        featU[u * fdim + d] = PCRCG_FDIV(sum, (float)cnt);
    }
}

// grid_subsampling.h:56-61 + .cpp:97-102 (label_vote.h).  Returns true when a voxel held more distinct labels than LV_CAP.
PCRCG_HD bool label_vote_thread(int j, int d0, int dstep, const int32_t* cls, int ldim, const uint32_t* sslot, const uint32_t* sidx,
                                int n, const uint32_t* rank, int32_t* clsU)
{
    if (j >= n) return false;
    const uint32_t s = sslot[j];
    if (j != 0 && sslot[j - 1] == s) return false;
    const size_t u = rank[sidx[j]];
    bool overflow = false;
    for (int d = d0; d < ldim; d += dstep) {
        LabelVote v;
        v.reset();
        for (int t = j; t < n && sslot[t] == s; t++) v.add(cls[(size_t)sidx[t] * ldim + d]);
        overflow = overflow || v.overflow;
        clsU[u * ldim + d] = v.pick();
    }
    return overflow;
}

// Output row e of cloud c is voxel L[e] of the cloud's final list (grid_subsampling.cpp:85-102 walks the container once for
// points, features and classes alike).  Thread t of nthreads working on cloud c; sched = the bucket-count schedule (c_sched).
PCRCG_HD void gather_extra_thread(int c, int t, int nthreads, const uint32_t* rank, const int32_t* starts, const int32_t* out_lens,
                                  const int32_t* out_base, const uint32_t* seqA, const uint32_t* seqB, const uint32_t* sched,
                                  const float* featU, int fdim, float* out_feat, const int32_t* clsU, int ldim, int32_t* out_cls)
{
    const int s0 = starts[c];
    const uint32_t Ub = rank[s0];
    const int M = (int)(rank[starts[c + 1]] - Ub);
    int epochs = 0;
    for (int done = 0; done < M; epochs++) done = (uint32_t)M < sched[epochs] ? M : (int)sched[epochs];
    const uint32_t* L = ((epochs & 1) ? seqB : seqA) + Ub;
    const int m_out = out_lens[c];
    const size_t ob = (size_t)out_base[c];
    if (out_feat != nullptr)
        for (long long k = t; k < (long long)m_out * fdim; k += nthreads) {
            const long long e = k / fdim, d = k - e * fdim;
            out_feat[(ob + e) * fdim + d] = featU[((size_t)Ub + L[e]) * fdim + d];
        }
    if (out_cls != nullptr)
        for (long long k = t; k < (long long)m_out * ldim; k += nthreads) {
            const long long e = k / ldim, d = k - e * ldim;
            out_cls[(ob + e) * ldim + d] = clsU[((size_t)Ub + L[e]) * ldim + d];
        }
}

}  // namespace pcrcg
