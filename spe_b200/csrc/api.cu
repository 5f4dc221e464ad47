#include "common.cuh"

thread_local char g_spe_err[512] = {0};
std::atomic<int64_t> g_spe_launches{0};

extern "C" __attribute__((visibility("default"))) const char* spe_last_error(void) { return g_spe_err; }
extern "C" __attribute__((visibility("default"))) int spe_version(void) { return 100; }
extern "C" __attribute__((visibility("default"))) int64_t spe_launch_count(void) { return g_spe_launches.load(); }
// Device-wide L1 / shared-memory split preference.  The tcgen05 kernels run with ~200 KB of shared memory per CTA, the small row-wise
// kernels with none: every switch between the two carve-outs makes the SMs drain and reconfigure.  mode 1 = prefer shared memory for
// every kernel that states no preference of its own (cudaFuncCachePreferShared), 0 = driver default.
extern "C" __attribute__((visibility("default"))) int spe_set_cache_config(int mode) {
    return cudaDeviceSetCacheConfig(mode ? cudaFuncCachePreferShared : cudaFuncCachePreferNone) == cudaSuccess ? 0 : -1;
}

// ------------------------------------------------------------------------------------------------
// per-family event profiler (off by default; bench.py turns it on for the timed region)
// ------------------------------------------------------------------------------------------------
#include <vector>
#include <string.h>
#include <stdlib.h>
#include <mutex>
bool g_spe_prof_on = false;
namespace {
struct ProfRec { cudaEvent_t a, b; int fam; double work; char tag[64]; };
std::vector<ProfRec> g_recs;
std::vector<cudaEvent_t> g_pool;
std::mutex g_prof_mu;
cudaEvent_t get_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
}  // namespace
int spe_prof_begin_(int fam, double work, cudaStream_t st, const char* tag) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r; r.a = get_event(); r.b = get_event(); r.fam = fam; r.work = work;
    r.tag[0] = 0;
    if (tag) { strncpy(r.tag, tag, sizeof(r.tag) - 1); r.tag[sizeof(r.tag) - 1] = 0; }
    cudaEventRecord(r.a, st);
    g_recs.push_back(r);
    return (int)g_recs.size() - 1;
}
void spe_prof_end_(int idx, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (idx >= 0 && idx < (int)g_recs.size()) cudaEventRecord(g_recs[idx].b, st);
}
extern "C" __attribute__((visibility("default"))) int spe_prof_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_spe_prof_on = on != 0;
    return 0;
}
// sums per family since the last collect: ms[SPE_FAM_COUNT], work[...], launches[...]; synchronises the device
extern "C" __attribute__((visibility("default"))) int spe_prof_collect(double* ms, double* work, int64_t* launches) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int f = 0; f < SPE_FAM_COUNT; ++f) { ms[f] = 0; work[f] = 0; launches[f] = 0; }
    const char* csv = getenv("SPE_PROF_CSV");      // optional per-launch dump: family,tag,work,ms
    FILE* fp = csv ? fopen(csv, "a") : nullptr;
    for (auto& r : g_recs) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms[r.fam] += t; work[r.fam] += r.work; launches[r.fam] += 1; }
        if (fp) fprintf(fp, "%d,%s,%.6g,%.6f\n", r.fam, r.tag, r.work, t);
        g_pool.push_back(r.a); g_pool.push_back(r.b);
    }
    g_recs.clear();
    if (fp) fclose(fp);
    return 0;
}
extern "C" __attribute__((visibility("default"))) int spe_prof_family_count(void) { return SPE_FAM_COUNT; }
