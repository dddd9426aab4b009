// ref_shim.cu -- TEST INFRASTRUCTURE ONLY.
// C-linkage trampolines onto the reference's StructuralLosses launchers
// (evaluation/pytorch_structural_losses/src/nndistance.cuh:1-2 declares them with C++ linkage),
// so tests can drive the recompiled reference kernels through ctypes.  No reference code is copied:
// the objects are compiled by oracle/Makefile straight from /root/reference.
#include <cuda_runtime.h>

void nndistance(int b, int n, const float *xyz, int m, const float *xyz2, float *result, int *result_i,
                float *result2, int *result2_i, cudaStream_t stream);
void nndistancegrad(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_dist1,
                    const int *idx1, const float *grad_dist2, const int *idx2, float *grad_xyz1,
                    float *grad_xyz2, cudaStream_t stream);

void approxmatch(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, float *temp, cudaStream_t stream);
void matchcost(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, float *out, cudaStream_t stream);
void matchcostgrad(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match, float *grad1, float *grad2,
                   cudaStream_t stream);

extern "C" {
// ApproxMatch + MatchCost as MatchCostFunction.forward chains them (match_cost.py:21-23).
// match: [b, m, n] scratch, temp: [b, (n+m)*2] scratch, out: [b].
int ref_match_cost(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, float *temp, float *out, void *stream) {
    try {
        approxmatch(b, n, m, xyz1, xyz2, match, temp, (cudaStream_t)stream);
        matchcost(b, n, m, xyz1, xyz2, match, out, (cudaStream_t)stream);
    } catch (...) {
        return -1;
    }
    return (int)cudaGetLastError();
}
// The three StructuralLossesBackend entry points one by one (pybind/bind.cpp:10-16), for oracle/ref_tree.py.
int ref_approxmatch(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, float *temp, void *stream) {
    try { approxmatch(b, n, m, xyz1, xyz2, match, temp, (cudaStream_t)stream); } catch (...) { return -1; }
    return (int)cudaGetLastError();
}
int ref_matchcost(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, float *out, void *stream) {
    try { matchcost(b, n, m, xyz1, xyz2, match, out, (cudaStream_t)stream); } catch (...) { return -1; }
    return (int)cudaGetLastError();
}
int ref_matchcostgrad(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match, float *g1, float *g2, void *stream) {
    try { matchcostgrad(b, n, m, xyz1, xyz2, match, g1, g2, (cudaStream_t)stream); } catch (...) { return -1; }
    return (int)cudaGetLastError();
}
int ref_nndistance(int b, int n, const float *xyz, int m, const float *xyz2, float *d1, int *i1, float *d2, int *i2,
                   void *stream) {
    nndistance(b, n, xyz, m, xyz2, d1, i1, d2, i2, (cudaStream_t)stream);
    return (int)cudaGetLastError();
}
int ref_nndistancegrad(int b, int n, const float *xyz1, int m, const float *xyz2, const float *g1, const int *i1,
                       const float *g2, const int *i2, float *gx1, float *gx2, void *stream) {
    nndistancegrad(b, n, xyz1, m, xyz2, g1, i1, g2, i2, gx1, gx2, (cudaStream_t)stream);
    return (int)cudaGetLastError();
}
int ref_sync(void) { return (int)cudaDeviceSynchronize(); }
}
