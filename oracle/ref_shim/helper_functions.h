/* Stand-in for the CUDA-samples header the reference includes (SpGEMM_cuda/common.h:21).
 * TEST INFRASTRUCTURE (oracle/_ref build only): the minimum the reference uses from it --
 * the StopWatchInterface host timer (bhsparse.h:37-40, bhsparse_cuda.h sdk*Timer calls).
 * Written for this repo; nothing here comes from the CUDA samples. */
#ifndef BHB200_REF_SHIM_HELPER_FUNCTIONS_H
#define BHB200_REF_SHIM_HELPER_FUNCTIONS_H
#include <chrono>

struct StopWatchInterface {
    std::chrono::steady_clock::time_point t0;
    double total_ms = 0.0;
    bool running = false;
};
inline bool sdkCreateTimer(StopWatchInterface **t) { *t = new StopWatchInterface(); return true; }
inline bool sdkDeleteTimer(StopWatchInterface **t) { delete *t; *t = nullptr; return true; }
inline bool sdkStartTimer(StopWatchInterface **t) { (*t)->t0 = std::chrono::steady_clock::now(); (*t)->running = true; return true; }
inline bool sdkStopTimer(StopWatchInterface **t)
{
    if ((*t)->running) {
        (*t)->total_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - (*t)->t0).count();
        (*t)->running = false;
    }
    return true;
}
inline bool sdkResetTimer(StopWatchInterface **t) { (*t)->total_ms = 0.0; (*t)->running = false; return true; }
inline float sdkGetTimerValue(StopWatchInterface **t)
{
    double ms = (*t)->total_ms;
    if ((*t)->running) ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - (*t)->t0).count();
    return (float)ms;
}
#endif
