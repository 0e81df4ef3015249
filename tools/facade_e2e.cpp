// facade_e2e.cpp -- the headline workload driven exactly like a caller of the reference drives it: through the C++ class API
// (CVoxelyze::setVoxel, external()->set*, doTimeStep(dt), voxel->position()), here on the facade (libvoxelyze_facade.so over the
// C-ABI over the CUDA kernels).  bench.py runs it at N = 1 and reports the result as `e2e.facade`: per-step cost of the facade's
// change tracking and lazy state mirror on top of vx_step.  usage: facade_e2e <edge> <warmup> <steps> [--build-only]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "Voxelyze.h"

int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 64, warmup = argc > 2 ? atoi(argv[2]) : 5, steps = argc > 3 ? atoi(argv[3]) : 20;
    const bool build_only = argc > 4 && !strcmp(argv[4], "--build-only");
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const auto t0 = now();
    CVoxelyze Vx(0.005);
    CVX_Material* m = Vx.addMaterial(1e6f, 1e3f);
    for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) Vx.setVoxel(m, i, j, k);
    const float load = -1.0f / ((float)n * (float)n);
    for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) {
        Vx.voxel(0, j, k)->external()->setFixedAll();
        Vx.voxel(n - 1, j, k)->external()->setForce(0, 0, load);
    }
    const auto t1 = now();
    if (build_only) { printf("{\"edge\": %d, \"build_s\": %.2f}\n", n, secs(t0, t1)); return 0; }
    const float dt = Vx.recommendedTimeStep();
    CVX_Voxel* probe = Vx.voxel(n - 1, n - 1, n - 1);
    double acc = 0;
    for (int s = 0; s < warmup; s++) { if (!Vx.doTimeStep(dt)) return 2; acc += probe->position().z; }
    const auto t2 = now();
    for (int s = 0; s < steps; s++) { if (!Vx.doTimeStep(dt)) return 2; acc += probe->position().z; }
    const auto t3 = now();
    const double units = (double)Vx.voxelCount() + (double)Vx.linkCount();
    printf("{\"edge\": %d, \"voxels\": %d, \"links\": %d, \"dt\": %.9e, \"build_s\": %.2f, \"first_steps_s\": %.3f, \"steps\": %d, \"ms_per_step\": %.4f, "
           "\"updates_per_s\": %.6e, \"probe_sum\": %.9e}\n", n, Vx.voxelCount(), Vx.linkCount(), dt, secs(t0, t1), secs(t1, t2), steps,
           1e3 * secs(t2, t3) / steps, units * steps / secs(t2, t3), acc);
    return 0;
}
