// Error plumbing of the C ABI (include/pmd_sm100.h).
#include "common.cuh"

namespace pmd {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
}  // namespace pmd

extern "C" const char* pmd_last_error(void) { return pmd::g_last_error.c_str(); }
extern "C" int pmd_abi_version(void) { return 1; }
