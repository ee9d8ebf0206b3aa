// Error reporting + ABI version for libsparsebev_b200.so.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

namespace sbev {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Implementation selectors (A/B testing of kernel variants).  Defaults come from the environment
// (SBEV_GEMM_IMPL, SBEV_MIX_IMPL, SBEV_SASA_IMPL, SBEV_GATHER_VARIANT), sbev_set_option overrides.
static const char* kOptNames[OPT_COUNT] = {"gemm_impl", "mix_impl", "sasa_impl", "gather_variant", "dense_impl", "dense_cluster", "pdl", "dense_nsplit", "dense_vec4", "dense_fuse_points", "mix_order", "legacy_rotation", "dense_pack", "sasa_kq", "dense_ws", "dense_ws_groups", "gemm_l2_hints", "gather_l2_hint"};
static const char* kOptEnv[OPT_COUNT] = {"SBEV_GEMM_IMPL", "SBEV_MIX_IMPL", "SBEV_SASA_IMPL", "SBEV_GATHER_VARIANT", "SBEV_DENSE_IMPL", "SBEV_DENSE_CLUSTER", "SBEV_PDL", "SBEV_DENSE_NSPLIT", "SBEV_DENSE_VEC4", "SBEV_DENSE_FUSE_POINTS", "SBEV_MIX_ORDER", "SBEV_LEGACY_ROTATION", "SBEV_DENSE_PACK", "SBEV_SASA_KQ", "SBEV_DENSE_WS", "SBEV_DENSE_WS_GROUPS", "SBEV_GEMM_L2_HINTS", "SBEV_GATHER_L2_HINT"};
static const int kOptDefault[OPT_COUNT] = {0, 0, 0, 6, 0, 0, 1, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0};
static int g_opt[OPT_COUNT] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};

int get_option(int id) {
    if (g_opt[id] < 0) {
        const char* e = getenv(kOptEnv[id]);
        g_opt[id] = e ? atoi(e) : kOptDefault[id];
    }
    return g_opt[id];
}
int set_option(const char* name, int value) {
    for (int i = 0; i < OPT_COUNT; ++i)
        if (strcmp(name, kOptNames[i]) == 0) { g_opt[i] = value; return SBEV_OK; }
    set_error("unknown option '%s'", name);
    return SBEV_ERR_INVALID;
}
}  // namespace sbev

extern "C" int sbev_abi_version(void) { return 1; }
extern "C" const char* sbev_last_error(void) { return sbev::g_err; }
extern "C" int sbev_get_option(const char* name) {
    if (!name) return -1;
    for (int i = 0; i < sbev::OPT_COUNT; ++i)
        if (strcmp(name, sbev::kOptNames[i]) == 0) return sbev::get_option(i);
    return -1;
}
extern "C" int sbev_set_option(const char* name, int value) {
    if (!name || value < 0) { sbev::set_error("sbev_set_option: bad arguments"); return SBEV_ERR_INVALID; }
    return sbev::set_option(name, value);
}
