// libtfasr_b200: library-level C ABI (version, status strings, launch accounting).
#include "common.cuh"

namespace tfasr {
std::atomic<int> g_launch_count{0};
}

extern "C" int tfasr_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char* tfasr_status_string(int st) {
  switch (st) {
    case TFASR_OK: return "ok";
    case TFASR_INVALID_VALUE: return "invalid value (bad shape, unsupported configuration, or blank != 0)";
    case TFASR_EXECUTION_FAILED: return "CUDA launch or execution failure";
    case TFASR_WORKSPACE_TOO_SMALL: return "workspace too small";
    default: return "unknown status";
  }
}

extern "C" int tfasr_launch_count(void) { return tfasr::g_launch_count.load(); }
extern "C" void tfasr_launch_count_reset(void) { tfasr::g_launch_count.store(0); }
