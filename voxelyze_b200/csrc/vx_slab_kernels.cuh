// vx_slab_kernels.cuh -- kernels of the peer-memory halo (see vx_slab.inl).
#pragma once

// ------------------------------------------------------------------------------------------------
// peer-memory halo: after the boundary part of a step each slab stores its fresh boundary poses straight
// into the ghost planes of its neighbours (CUDA IPC mappings, NVLink) and then bumps the neighbour's
// arrival counter; the neighbour's next boundary part spins on that counter first.
__global__ void k_halo_push(const double4* __restrict__ src0, const double4* __restrict__ src1, double4* dst0, double4* dst1, int count)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const double4 a = src0[k], b = src1[k];
    dst0[k] = a;
    double* d = reinterpret_cast<double*>(dst1 + k);
    d[0] = b.x; d[1] = b.y; d[2] = b.z;
    // the receiver owns the upper half of w (its flag word, see k_lattice_warp); the lower half is the temperature
    reinterpret_cast<uint32_t*>(d + 3)[0] = (uint32_t)(unsigned long long)__double_as_longlong(b.w);
}
__global__ void k_peer_signal(int* flag, int seq)
{
    __threadfence_system();
    *(volatile int*)flag = seq;
    __threadfence_system();
}
__global__ void k_peer_wait(const int* flags, int need_lo, int need_hi, int* timed_out, long long limit)
{
    const long long t0 = clock64();                        // limit in SM clocks: a lost peer must not hang the GPU
    while (*(volatile const int*)(flags + 0) < need_lo || *(volatile const int*)(flags + 1) < need_hi) {
        if (clock64() - t0 > limit) { *timed_out = 1; break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

