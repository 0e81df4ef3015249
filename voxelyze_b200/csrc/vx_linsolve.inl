// vx_linsolve.inl -- C-ABI entry of the static solve (part of vx_capi.cu's extern "C" block; kernels in vx_linsolve.cuh)

namespace {
struct LinWork {
    DevBuf<int> nbr, ijk, ext_vox; DevBuf<uint16_t> mat; DevBuf<unsigned char> fixed; DevBuf<DevExt> ext;
    DevBuf<double2> vec; DevBuf<double> part; DevBuf<LinScalars> sc;
    LinScalars* host = nullptr;
    ~LinWork()
    {
        nbr.release(); ijk.release(); ext_vox.release(); mat.release(); fixed.release(); ext.release(); vec.release(); part.release(); sc.release();
        if (host) cudaFreeHost(host);
    }
};
}

int vx_linear_solve(vx_sim* s, double rel_tol, int max_iter, int* iterations, double* rel_residual)
{
    if (iterations) *iterations = 0;
    if (rel_residual) *rel_residual = 0.0;
    if (!s) return VX_ERR_ARG;
    if (s->call_active) return fail(s, VX_ERR_ARG, "vx_linear_solve inside vx_step_begin .. vx_step_end");
    if (!s->state_ready || s->N_user == 0) return fail(s, VX_ERR_ARG, "vx_linear_solve: no voxels");     // CVX_LinearSolver::solve returns false when dof == 0 (:60)
    for (int i = 0; i < s->N_user && i < (int)s->vflags.size(); i++)
        if (s->vflags[i] & VX_VF_GHOST) return fail(s, VX_ERR_UNSUPPORTED, "vx_linear_solve on a z-slab with ghost layers");
    if (!(rel_tol > 0.0)) rel_tol = 1e-10;
    if (max_iter <= 0) max_iter = 200000;
    NvtxRange nvtx("vx_linear_solve");
    { int rc = flush_ambient(s); if (rc != VX_OK) return rc; }
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));

    const int n = s->N_user;
    const size_t n6 = (size_t)6 * n;
    std::vector<int> nbr(n6, -1);
    for (int l = 0; l < s->L; l++) {
        const int vn = s->lk_vn[l], vp = s->lk_vp[l], ax = s->lk_axis[l];
        if (vn >= n || vp >= n) continue;                                     // links of inert fill cells (none are created, but stay safe)
        nbr[(size_t)(2 * ax) * n + vn] = vp;                                   // the +axis slot of the negative end
        nbr[(size_t)(2 * ax + 1) * n + vp] = vn;
    }
    LinWork w;
    const int grid = std::max(1, std::min(blocks_for(n, VX_LIN_TPB), VX_LIN_MAX_GRID));
    const size_t n3 = (size_t)3 * n;
    CK(w.nbr.alloc(n6)); CK(w.ijk.alloc((size_t)3 * n)); CK(w.mat.alloc(n)); CK(w.fixed.alloc(n));
    CK(w.vec.alloc(7 * n3)); CK(w.part.alloc((size_t)5 * grid)); CK(w.sc.alloc(1));
    CK(cudaHostAlloc((void**)&w.host, sizeof(LinScalars), cudaHostAllocDefault));
    CK(cudaMemcpyAsync(w.nbr.p, nbr.data(), n6 * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemcpyAsync(w.ijk.p, s->ijk.data(), (size_t)3 * n * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemcpyAsync(w.mat.p, s->vmat_id.data(), (size_t)n * sizeof(uint16_t), cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemsetAsync(w.fixed.p, 0, n, s->stream));
    CK(cudaMemsetAsync(w.vec.p, 0, 7 * n3 * sizeof(double2), s->stream));
    CK(cudaMemsetAsync(w.part.p, 0, (size_t)5 * grid * sizeof(double), s->stream));

    LinFrame f{};
    f.n = n; f.grid = grid; f.nbr = w.nbr.p; f.mat = w.mat.p; f.pair_lmat = s->pair_lmat.p; f.n_mat = (int)s->mats.size(); f.lmat = s->lmat_dev.p;
    f.fixed = w.fixed.p; f.ijk = w.ijk.p; f.e2i = s->vox_e2i_dev.p; f.voxel_size = s->vox_size;
    f.x = w.vec.p; f.r = f.x + n3; f.z = f.r + n3; f.y = f.z + n3; f.minv = f.y + n3; f.p[0] = f.minv + n3; f.p[1] = f.p[0] + n3;
    f.part_pap = w.part.p; f.part_rz[0] = f.part_pap + grid; f.part_rz[1] = f.part_rz[0] + grid; f.part_rr[0] = f.part_rz[1] + grid; f.part_rr[1] = f.part_rr[0] + grid;
    f.sc = w.sc.p;
    double2* load = f.p[1];                                                   // not read before iteration 1 writes it

    const int n_ext = (int)s->ext_vox.size();
    if (n_ext) {
        CK(w.ext_vox.alloc(n_ext)); CK(w.ext.alloc(n_ext));
        CK(cudaMemcpyAsync(w.ext_vox.p, s->ext_vox.data(), (size_t)n_ext * sizeof(int), cudaMemcpyHostToDevice, s->stream));
        CK(cudaMemcpyAsync(w.ext.p, s->ext_rows.data(), (size_t)n_ext * sizeof(DevExt), cudaMemcpyHostToDevice, s->stream));
        k_lin_externals<<<blocks_for(n_ext), TPB, 0, s->stream>>>(f, n_ext, w.ext_vox.p, w.ext.p, load);
        s->launches++;
    }
    const Frame fr = s->frame();
    k_lin_start<<<grid, VX_LIN_TPB, 0, s->stream>>>(f, fr.pose0, fr.pose1);
    k_lin_residual0<<<grid, VX_LIN_TPB, 0, s->stream>>>(f, load);
    k_lin_begin<<<1, VX_LIN_TPB, 0, s->stream>>>(f, rel_tol);
    s->launches += 3;
    CK(cudaGetLastError());

    // iterations in batches; the device decides convergence (LinScalars::done), the host looks after every batch
    const int batch = 64;
    int k = 0;
    for (;;) {
        const int upto = std::min(max_iter, k + batch);
        for (; k < upto; k++) {
            k_lin_step_a<<<grid, VX_LIN_TPB, 0, s->stream>>>(f, k);
            k_lin_step_b<<<grid, VX_LIN_TPB, 0, s->stream>>>(f, k);
        }
        k_lin_status<<<1, VX_LIN_TPB, 0, s->stream>>>(f, k);
        s->launches += 2 * batch + 1;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(w.host, w.sc.p, sizeof(LinScalars), cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        if (w.host->done || k >= max_iter) break;
    }
    if (iterations) *iterations = w.host->iters;
    if (rel_residual) *rel_residual = w.host->bb > 0.0 ? sqrt(w.host->rr / w.host->bb) : 0.0;
    if (w.host->done == 2) return fail(s, VX_ERR_SOLVER, "vx_linear_solve: the stiffness matrix is singular (a part of the model is not held) or not positive definite");
    if (!w.host->done) return fail(s, VX_ERR_SOLVER, "vx_linear_solve: no convergence within max_iter iterations");

    k_lin_post<<<grid, VX_LIN_TPB, 0, s->stream>>>(f, fr.pose0, fr.pose1, fr.mom0, fr.mom1);
    s->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s->stream));
    return VX_OK;
}
