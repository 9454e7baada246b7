// cabi.cu — ABI version / status strings / device queries of libfedmlp_b200.
#include <cstring>
#include <thread>
#include <vector>

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

// Host utility: n independent memcpy's (dsts[i] <- srcs[i], nbytes[i] bytes), split over `n_threads` host threads by
// bytes.  Used by FedAvg's CPU-state_dict path to pack 727 pageable tensors per client into one pinned buffer at
// memory bandwidth (a per-tensor torch copy costs ~7 us of dispatch each: 62 ms for 8 DenseNet121 clients).
extern "C" int fmlp_host_copy_many(const void* const* srcs, void* const* dsts, const int64_t* nbytes, int64_t n, int n_threads) {
    if (n < 0 || (n > 0 && (!srcs || !dsts || !nbytes))) return FMLP_ERR_BAD_ARG;
    int64_t total = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (nbytes[i] < 0 || (nbytes[i] > 0 && (!srcs[i] || !dsts[i]))) return FMLP_ERR_BAD_ARG;
        total += nbytes[i];
    }
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    if (total < (int64_t)1 << 20 || n_threads == 1) {
        for (int64_t i = 0; i < n; ++i) if (nbytes[i]) memcpy(dsts[i], srcs[i], (size_t)nbytes[i]);
        return FMLP_OK;
    }
    // byte ranges [t*share, (t+1)*share) of the concatenated copy list; a tensor may be split between two threads
    const int64_t share = (total + n_threads - 1) / n_threads;
    auto work = [&](int t) {
        const int64_t lo = (int64_t)t * share, hi = lo + share < total ? lo + share : total;
        int64_t pos = 0;
        for (int64_t i = 0; i < n && pos < hi; ++i) {
            const int64_t b = nbytes[i], a0 = pos > lo ? pos : lo, a1 = pos + b < hi ? pos + b : hi;
            if (a1 > a0) memcpy((char*)dsts[i] + (a0 - pos), (const char*)srcs[i] + (a0 - pos), (size_t)(a1 - a0));
            pos += b;
        }
    };
    std::vector<std::thread> th;
    th.reserve(n_threads - 1);
    for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    return FMLP_OK;
}
