/* A C caller of libpcrcg_b200.so: grid subsampling of one random cloud through the host-buffer entry point, the call the
 * reference's CPython glue makes (zip!cpp_subsampling/wrapper.cpp:62-333 -> grid_subsampling.cpp:109-211).
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -Lpcrcg_b200 -lpcrcg_b200 -Wl,-rpath,$PWD/pcrcg_b200 -o /tmp/c_abi_demo && /tmp/c_abi_demo
 *
 * Exit status 0 = subsampled on the GPU; 2 = the library reported an error (printed), e.g. no CUDA device: there is no CPU path. */
#include <stdio.h>
#include <stdlib.h>

#include "pcrcg_b200.h"

int main(void)
{
    enum { N = 5000 };
    float* pts = (float*)malloc(sizeof(float) * 3 * N);
    unsigned s = 12345u;
    for (int i = 0; i < 3 * N; i++) {
        s = s * 1664525u + 1013904223u;
        pts[i] = (float)(s >> 8) / 16777216.0f;              /* [0, 1) */
    }
    int32_t lens[1] = { N }, out_lens[1] = { 0 };
    float* out = NULL;
    int64_t m = 0;
    printf("libpcrcg_b200 version %d\n", pcrcg_version());
    int rc = pcrcg_subsample_batch_host(pts, N, lens, 1, 0.1f, 0, &out, &m, out_lens);
    if (rc != 0) {
        printf("pcrcg_subsample_batch_host failed: %s\n", pcrcg_last_error());
        free(pts);
        return 2;
    }
    printf("%d points -> %lld voxels (first barycentre %.6f %.6f %.6f)\n", N, (long long)m, out[0], out[1], out[2]);
    pcrcg_free(out);
    free(pts);
    return 0;
}
