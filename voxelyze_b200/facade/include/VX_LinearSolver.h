// VX_LinearSolver.h -- drop-in CVX_LinearSolver of the voxelyze_b200 facade (reference: include/VX_LinearSolver.h:42-84).
//
// Same public interface: construct on a CVoxelyze, solve() formulates and solves the linearised system and writes the
// resulting positions and angles back into the linked object; progressTick / progressMaxTick / progressMsg / errorMsg /
// cancelFlag as in the reference.  The reference needs the closed-source PARDISO library (solve() returns false without
// it); here the system is solved on the device by vx_linear_solve (include/voxelyze_b200.h: matrix-free preconditioned
// conjugate gradients in FP64), so the extra knobs below exist and cancelFlag is only honoured before the solve starts.
#ifndef VXB200_LINEARSOLVER_H
#define VXB200_LINEARSOLVER_H

#include <string>
#include "Voxelyze.h"

class CVX_LinearSolver {
public:
    CVX_LinearSolver(CVoxelyze* voxelyze) : vx(voxelyze) {}
    bool solve();

    int progressTick = 0;
    int progressMaxTick = 100;
    std::string progressMsg;
    std::string errorMsg;
    bool cancelFlag = false;

    // facade extras (additive)
    double relTolerance = 0.0;      // <= 0: the library default (1e-10 relative residual)
    int maxIterations = 0;          // <= 0: the library default
    int iterations = 0;             // of the last solve
    double residual = 0.0;          // relative residual reached

private:
    CVoxelyze* vx;
    void updateProgress(float percent, const std::string& message) { progressTick = (int)(percent * 100); progressMsg = message; }
};

#endif // VXB200_LINEARSOLVER_H
