// Batched single points: get_xtb_egrad (reference src/tblite.f90:65-175) for nsys geometries of one composition.
#pragma once
#include "qx_cid.cuh"

namespace qx {

static __global__ void __launch_bounds__(QX_NT, QX_MINB) k_egrad_batch(DevModel m, ScratchLayout L, double *scratch, const double *xyz, double kt, int nsys,
                                                       int *queue, double *energy, double *grad, double *qat, int *stat, int *niter, double *spec) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_next;
    Sm s;
    double *my = scratch + (size_t)blockIdx.x * L.total;
    carve(m, smem, s, my + L.matA);
    const int nat = m.nat;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_next = atomicAdd(queue, 1);
        __syncthreads();
        const int t = s_next;
        if (t >= nsys) break;
        for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) s.xyz[i] = xyz[(size_t)t * 3 * nat + i];
        __syncthreads();
        EgradOut o;
        egrad_cta(m, s, my, L, kt, o, nullptr, spec ? spec + (size_t)t * (2 * m.nao + m.nao * nat + 1) : nullptr);
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) grad[(size_t)t * 3 * nat + i] = s.grad[i];
        for (int i = threadIdx.x; i < nat; i += QX_NT) qat[(size_t)t * nat + i] = s.qat[i];
        if (threadIdx.x == 0) {
            energy[t] = o.energy;
            stat[t] = o.stat == 0 ? 0 : -1;
            if (niter) niter[t] = o.niter;
        }
    }
}

}  // namespace qx
