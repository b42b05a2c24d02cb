// Library-level entry points of libide3d_b200.so: ABI version, thread-local error text, launch counter.
#include <atomic>

#include "common.cuh"

namespace ide3d {

char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

}  // namespace ide3d

extern "C" int ide3d_abi_version(void) { return IDE3D_ABI_VERSION; }
extern "C" const char* ide3d_last_error(void) { return ide3d::error_buffer(); }
extern "C" uint64_t ide3d_launch_count(void) { return ide3d::g_launches.load(std::memory_order_relaxed); }
