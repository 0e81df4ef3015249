// voxelyze_solver.cpp -- CVX_LinearSolver of the facade (reference: src/VX_LinearSolver.cpp:49-113).  All arithmetic happens in
// vx_linear_solve on the device.
#include "VX_LinearSolver.h"

bool CVX_LinearSolver::solve()
{
    updateProgress(0, "Forming matrices...");
    errorMsg.clear();
    if (cancelFlag) { cancelFlag = false; errorMsg = "Cancelled\n"; return false; }
    if (vx->voxelCount() == 0) return false;                            // :60
    updateProgress(0.05f, "Solving...");
    const bool ok = vx->staticSolve(relTolerance, maxIterations, &iterations, &residual, &errorMsg);
    if (!ok) { if (errorMsg.empty()) errorMsg = "Solver error\n"; return false; }
    updateProgress(0.9f, "Processing results...");
    return true;
}
