#include "common.cuh"

thread_local char g_spe_err[512] = {0};
std::atomic<int64_t> g_spe_launches{0};

extern "C" __attribute__((visibility("default"))) const char* spe_last_error(void) { return g_spe_err; }
extern "C" __attribute__((visibility("default"))) int spe_version(void) { return 100; }
extern "C" __attribute__((visibility("default"))) int64_t spe_launch_count(void) { return g_spe_launches.load(); }
