#include "common.cuh"

namespace mvsb200 {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace mvsb200

extern "C" int mvsb200_abi_version(void) { return MVSB200_ABI_VERSION; }
extern "C" const char* mvsb200_last_error(void) { return mvsb200::g_err; }
extern "C" uint64_t mvsb200_launch_count(void) { return mvsb200::g_launches.load(); }
