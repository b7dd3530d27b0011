// Library-wide state of the C ABI: error string, launch counter, ABI version.
#include "common.cuh"

namespace sdet {
thread_local char g_last_error[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace sdet

#ifndef SDET_BUILD_DIGEST
#define SDET_BUILD_DIGEST "unknown"
#endif
extern "C" int sdet_abi_version(void) { return 4; }
// digest of the sources, headers and flags this binary was compiled from (simpledet_b200/build.py source_digest())
extern "C" const char* sdet_build_digest(void) { return SDET_BUILD_DIGEST; }
extern "C" const char* sdet_last_error(void) { return sdet::g_last_error; }
extern "C" uint64_t sdet_launch_count(void) { return sdet::g_launches.load(); }
