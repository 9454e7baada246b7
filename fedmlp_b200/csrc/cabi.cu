// cabi.cu — ABI version / status strings / device queries of libfedmlp_b200.
#include "common.cuh"

namespace fmlp {
unsigned long long g_launch_count = 0;
int g_tuning[FMLP_TUNE_COUNT] = {-1, -1, -1, -1};     // -1 = not set: the environment variable / built-in default applies
}

extern "C" unsigned long long fmlp_launch_count(void) { return __atomic_load_n(&fmlp::g_launch_count, __ATOMIC_RELAXED); }

extern "C" int fmlp_abi_version(void) { return FMLP_ABI_VERSION; }

extern "C" const char* fmlp_status_string(int code) {
    switch (code) {
        case FMLP_OK: return "ok";
        case FMLP_ERR_BAD_ARG: return "fedmlp_b200: bad argument (null pointer, negative size, or K/S/C out of range)";
        case FMLP_ERR_UNSUPPORTED: return "fedmlp_b200: unsupported shape or alignment";
        case FMLP_ERR_WORKSPACE: return "fedmlp_b200: workspace too small";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "fedmlp_b200: unknown status";
}

extern "C" int fmlp_sm_count(void) { return fmlp::sm_count(); }

extern "C" int fmlp_set_tuning(int knob, int value) {
    if (knob < 0 || knob >= FMLP_TUNE_COUNT || value < -1) return FMLP_ERR_BAD_ARG;
    __atomic_store_n(&fmlp::g_tuning[knob], value, __ATOMIC_RELAXED);
    return FMLP_OK;
}
extern "C" int fmlp_get_tuning(int knob) {
    if (knob < 0 || knob >= FMLP_TUNE_COUNT) return -1;
    return __atomic_load_n(&fmlp::g_tuning[knob], __ATOMIC_RELAXED);
}
