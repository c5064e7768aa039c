/* tools/workload_stats.cpp -- TEST/ANALYSIS TOOL: the CPU emulation of the device core
 * (tests/emu/emu_driver.cpp) built with -DNDT_STATS so core.cuh's NDT_STAT hooks count
 * what a frame's rays actually do (leaf visits, bounding-sphere tests, primitive tests).
 * Used to size the kernels; never part of the product library. */
#define NDT_STATS 1
#include "../tests/emu/emu_driver.cpp"
NdtStats ndt_stats;
bool ndt_stats_box_miss = false;
extern "C" void emu_stats_reset() { memset(&ndt_stats, 0, sizeof ndt_stats); }
extern "C" void emu_stats_get(unsigned long long *out) { memcpy(out, &ndt_stats, sizeof ndt_stats); }
extern "C" int emu_stats_words() { return (int)(sizeof ndt_stats / sizeof(unsigned long long)); }
