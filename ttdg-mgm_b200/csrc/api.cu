// api.cu - version / limits of libttdg_sm100.so (host only).
#include "common.cuh"
#include <string.h>
#include <atomic>

namespace ttdg {
static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace ttdg

extern "C" long long ttdg_launch_count(void) { return ttdg::g_launches.load(); }

extern "C" int ttdg_version(void) { return 100; }   // 0.1.0

extern "C" const char *ttdg_build_info(void) {
    return "libttdg_sm100 0.1.0 | sm_100a | nvcc " TTDG_STR(__CUDACC_VER_MAJOR__) "." TTDG_STR(__CUDACC_VER_MINOR__)
           " | built " __DATE__;
}

extern "C" int ttdg_limit(const char *name) {
    if (!name) return -1;
    if (!strcmp(name, "small_max_dim")) return 96;
    if (!strcmp(name, "lap_max_dim")) return 128;
    if (!strcmp(name, "gagm_max_graphs")) return 64;
    if (!strcmp(name, "gagm_max_nodes")) return 96;
    if (!strcmp(name, "univ")) return 32;
    if (!strcmp(name, "feat_dim")) return 256;
    return -1;
}
