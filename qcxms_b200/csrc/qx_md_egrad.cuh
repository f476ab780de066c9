// One energy/gradient evaluation + the reference's sanity gate, shared by the MD and the CID kernels.
#pragma once
#include "qx_cid.cuh"

namespace qx {

// one egrad + sanity gate for trajectory t; returns Epot (0 on failure, like the reference's checkqc)
__device__ inline double md_egrad(const DevModel &m, Sm &s, double *my, const ScratchLayout &L, const MdConfig &cfg, double etemp,
                                  double *grad_out, double *achrg_out, int *niter_out, double *qstart = nullptr, double *eigseed = nullptr) {
    const int nat = m.nat;
    EgradOut o;
    egrad_cta(m, s, my, L, etemp * QC_KTOAU, o, qstart, nullptr, eigseed);
    __syncthreads();
    if (threadIdx.x == 0) {
        bool ok = o.stat != -2 && md_checkqc(m, o.energy, s.grad, s.qat, cfg.mchrg);
        s.red[48] = ok ? o.energy : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) grad_out[i] = s.grad[i];
    for (int i = threadIdx.x; i < nat; i += QX_NT) achrg_out[i] = s.qat[i];
    *niter_out = o.niter;
    return s.red[48];
}

}  // namespace qx
