// api.cu -- version / error-string entry points of libpdgn_b200.
#include "common.cuh"

extern "C" int pdgn_abi_version(void) {
    PDGN_RANGE("pdgn_abi_version"); return PDGN_ABI_VERSION; }

extern "C" const char* pdgn_error_string(int code) {
    switch (code) {
        case PDGN_OK: return "success";
        case PDGN_ERR_BAD_ARG: return "pdgn_b200: bad argument (null pointer, negative size or inconsistent range)";
        case PDGN_ERR_UNSUPPORTED: return "pdgn_b200: size outside the supported range of this entry point";
        case PDGN_ERR_WORKSPACE: return "pdgn_b200: workspace missing, misaligned or too small";
        case PDGN_ERR_INDEX: return "pdgn_b200: index out of range (PDGN_B200_VERIFY=1)";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "pdgn_b200: unknown error code";
}
