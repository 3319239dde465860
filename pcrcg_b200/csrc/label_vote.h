// The class vote of one voxel of the grid subsampling -- grid_subsampling.h:56-61 (labels[i][*it] += 1 in point order) and
// grid_subsampling.cpp:98-102 (max_element over the unordered_map<int, int>: the FIRST maximal element in the container's
// iteration order).  With tied votes the answer is decided by libstdc++'s iteration order, which is reproduced here in closed
// form per rehash epoch, exactly as k_order (subsample.cu) does for the voxels themselves:
//
//   elements  = the distinct labels in order of first occurrence, hash = (size_t)label (std::hash<int> is the identity)
//   epoch j   = rehash of the current list to P_j buckets (13, 29, 59, 127: _Prime_rehash_policy, max_load_factor 1), then
//               insertion of the next distinct labels until P_j elements are held
//   one epoch over the sequence S (current list, then the new labels): buckets in REVERSE order of first appearance in S,
//               the elements of a bucket in REVERSE order of appearance
//
// Plain C++ without dependencies: compiled by nvcc into k_label_vote (subsample.cu) and, by the host test
// tests/test_label_vote_host.py, by g++ next to the live std::unordered_map of oracle/_ref -- the logic is pinned on the CPU
// although it only ever runs on the GPU in the product.
#pragma once

#ifdef __CUDACC__
#define PCRCG_HD __host__ __device__ __forceinline__
#else
#define PCRCG_HD inline
#endif

namespace pcrcg {

constexpr int LV_CAP = 64;                                    // distinct labels per voxel and label column (more: reported)

struct LabelVote {
    int key[LV_CAP];                                          // distinct labels, first-occurrence order
    int cnt[LV_CAP];
    int D;
    bool overflow;

    PCRCG_HD void reset() { D = 0; overflow = false; }

    // labels[i][*it] += 1
    PCRCG_HD void add(int label)
    {
        for (int e = 0; e < D; e++)
            if (key[e] == label) { cnt[e]++; return; }
        if (D == LV_CAP) { overflow = true; return; }
        key[D] = label;
        cnt[D] = 1;
        D++;
    }

    // first maximal element in iteration order
    PCRCG_HD int pick() const
    {
        if (D == 0) return 0;
        int best = 0, ties = 0;
        for (int e = 1; e < D; e++) best = cnt[e] > cnt[best] ? e : best;
        for (int e = 0; e < D; e++) ties += cnt[e] == cnt[best];
        if (ties == 1) return key[best];                     // the order cannot matter
        const unsigned long long sched[4] = { 13ull, 29ull, 59ull, 127ull };
        int A[LV_CAP], B[LV_CAP], bk[LV_CAP], fi[LV_CAP], gs[LV_CAP];
        int* L = A;
        int* Lo = B;
        int done = 0;
        for (int j = 0; done < D; j++) {
            const unsigned long long P = sched[j];
            const int hi = (unsigned long long)D < P ? D : (int)P;
            for (int p = done; p < hi; p++) L[p] = p;
            for (int p = 0; p < hi; p++) {
                bk[p] = (int)((unsigned long long)(long long)key[L[p]] % P);
                gs[p] = 0;
            }
            for (int p = 0; p < hi; p++) {                    // fi[p] = first appearance of p's bucket, gs[f] = size of the bucket opened at f
                int f = p;
                for (int q = 0; q < p; q++)
                    if (bk[q] == bk[p]) { f = q; break; }
                fi[p] = f;
                gs[f]++;
            }
            int later = 0;                                    // gs[f] := number of elements in buckets that open after f
            for (int f = hi - 1; f >= 0; f--) { const int g = gs[f]; gs[f] = later; later += g; }
            for (int p = 0; p < hi; p++) {
                int r = 0;                                    // elements of p's bucket that appear after p
                for (int q = p + 1; q < hi; q++) r += bk[q] == bk[p];
                Lo[gs[fi[p]] + r] = L[p];
            }
            int* t = L; L = Lo; Lo = t;
            done = hi;
        }
        int b = L[0];
        for (int e = 1; e < D; e++)
            if (cnt[b] < cnt[L[e]]) b = L[e];                 // max_element: replaced only by a strictly larger count
        return key[b];
    }
};

}  // namespace pcrcg
