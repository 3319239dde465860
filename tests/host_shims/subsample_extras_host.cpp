// TEST INFRASTRUCTURE: the thread bodies of k_bary_feat / k_label_vote / k_gather_extra (pcrcg_b200/csrc/subsample_extras.h)
// compiled for the host and run "thread" by "thread" in the launch geometry subsample_batch_ex_dev uses, on workspace arrays the
// test builds with NumPy + the CPU oracle (tests/test_subsample_extras_host.py).  Build with -ffp-contract=off.
#include "../../pcrcg_b200/csrc/subsample_extras.h"

extern "C" int host_subsample_extras(int n, int nb, const float* feat, int fdim, const int32_t* cls, int ldim, const uint32_t* sslot,
                                     const uint32_t* sidx, const uint32_t* rank, const int32_t* starts, const int32_t* out_lens,
                                     const int32_t* out_base, const uint32_t* seqA, const uint32_t* seqB, const uint32_t* sched,
                                     float* featU, int32_t* clsU, float* out_feat, int32_t* out_cls)
{
    int overflow = 0;
    const int N = n > 0 ? n : 1;
    if (fdim) {
        const int gx = (N + 255) / 256, gy = fdim < 64 ? fdim : 64;           // dim3 g(cdiv(N, 256), min(fdim, 64)), 256 threads
        for (int by = 0; by < gy; by++)
            for (int bx = 0; bx < gx; bx++)
                for (int tx = 0; tx < 256; tx++) pcrcg::bary_feat_thread(bx * 256 + tx, by, gy, feat, fdim, sslot, sidx, n, rank, featU);
    }
    if (ldim) {
        const int gx = (N + 127) / 128, gy = ldim < 64 ? ldim : 64;           // dim3 g(cdiv(N, 128), min(ldim, 64)), 128 threads
        for (int by = 0; by < gy; by++)
            for (int bx = 0; bx < gx; bx++)
                for (int tx = 0; tx < 128; tx++)
                    if (pcrcg::label_vote_thread(bx * 128 + tx, by, gy, cls, ldim, sslot, sidx, n, rank, clsU)) overflow = 1;
    }
    for (int c = 0; c < nb; c++)                                              // <<<nb, 256>>>
        for (int t = 0; t < 256; t++)
            pcrcg::gather_extra_thread(c, t, 256, rank, starts, out_lens, out_base, seqA, seqB, sched, featU, fdim, fdim ? out_feat : nullptr,
                                       clsU, ldim, ldim ? out_cls : nullptr);
    return overflow;
}
