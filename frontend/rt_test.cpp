// rt_test.cpp — console client of the asynchronous API, the Linux twin of the reference's rt_test_dll/rt_test_dll.cpp:12-44
// (StartRT -> poll status -> WaitRT; optional StopRT after a delay).
//   rt_test <scene.dae> [size] [spp] [depth] [stop_after_seconds]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "YulioRT.h"

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s <scene.dae> [size] [spp] [depth] [stop_after_seconds]\n", argv[0]); return 2; }
    Yulio::ParamsRT params;
    params.size = argc > 2 ? atoi(argv[2]) : 512;
    params.spp = argc > 3 ? atoi(argv[3]) : 16;                 // rt_test_dll.cpp:17 uses 16
    params.depth = argc > 4 ? atoi(argv[4]) : 10;
    const double stopAfter = argc > 5 ? atof(argv[5]) : -1.0;
    const auto t0 = std::chrono::steady_clock::now();
    if (!Yulio::StartRT(argv[1], &params)) { fprintf(stderr, "StartRT failed, error %d\n", (int)Yulio::GetLastErrorRT()); return 1; }
    if (stopAfter >= 0) {
        std::this_thread::sleep_for(std::chrono::duration<double>(stopAfter));
        Yulio::StopRT(true);
    } else {
        Yulio::StatusRT st;
        do {
            std::this_thread::sleep_for(std::chrono::milliseconds(100));
            Yulio::GetCurrentStatusRT(&st);
            printf("state %d progress %.3f\n", (int)st.state, st.progress);
        } while (st.state != Yulio::Done && st.state != Yulio::Stopped && st.lastError == Yulio::NoError);
        Yulio::WaitRT();
    }
    Yulio::StatusRT st; Yulio::GetCurrentStatusRT(&st);
    printf("finished: state %d progress %.3f lastError %d, %.3f s\n", (int)st.state, st.progress, (int)st.lastError,
           std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    return st.lastError == Yulio::NoError ? 0 : 1;
}
