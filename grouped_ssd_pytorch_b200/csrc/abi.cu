// abi.cu — library-level entry points of libgssd_b200.so (include/gssd.h).
#include <atomic>

#include "common.cuh"

namespace gssd {
static std::atomic<unsigned long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace gssd

extern "C" int gssd_abi_version(void) { return GSSD_ABI_VERSION; }

extern "C" uint64_t gssd_launch_count(void) { return gssd::g_launches.load(std::memory_order_relaxed); }

extern "C" const char *gssd_error_string(int code) {
    switch (code) {
        case GSSD_OK: return "ok";
        case GSSD_ERR_ARG: return "invalid argument (null pointer, non-positive size or bad enum)";
        case GSSD_ERR_LIMIT: return "size beyond the kernels' limits (GSSD_MAX_* in gssd.h)";
        case GSSD_ERR_WS: return "workspace too small (see gssd_workspace_bytes)";
        case GSSD_ERR_VALUE: return "value error (variance <= 0 or nms_thresh <= 0)";
        case GSSD_ERR_EMPTY: return "an image has no ground-truth box";
        case GSSD_ERR_UNSUPPORTED: return "shape not supported by the one-launch path (use the two stages)";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

extern "C" size_t gssd_workspace_bytes(int kind, int B, int P, int C, int sum_G, int top_k) {
    (void)C; (void)sum_G; (void)top_k;
    switch (kind) {
        case GSSD_WS_LSE: return 16;
        case GSSD_WS_MATCH: return 16;
        case GSSD_WS_LOSS: return (size_t)(B > 0 ? B : 1) * 8 /* max cluster */ * 2 * sizeof(double) + 16;
        case GSSD_WS_NMS: return 16;
        default: return 0;
    }
}
