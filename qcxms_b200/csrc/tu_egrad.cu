// Translation unit of the batched single-point kernel (see qx_kernels.h).
#include "qx_kern_egrad.cuh"
#include "qx_kernels.h"

namespace qx {

cudaError_t QX_CAT(tu_egrad_prepare_, QX_VARIANT)(const cudaDeviceProp &prop) { return allow_max_dynamic_smem(k_egrad_batch, prop); }

cudaError_t QX_CAT(tu_egrad_launch_, QX_VARIANT)(int grid, size_t smem, cudaStream_t st, DevModel m, ScratchLayout L, double *scratch, const double *xyz,
                                                 double kt, int nsys, int *queue, double *energy, double *grad, double *qat, int *stat, int *niter,
                                                 double *spec) {
    k_egrad_batch<<<grid, QX_NT, smem, st>>>(m, L, scratch, xyz, kt, nsys, queue, energy, grad, qat, stat, niter, spec);
    return cudaGetLastError();
}

QX_DEFINE_PHASE_READER(QX_CAT(tu_egrad_cycles_, QX_VARIANT))

}  // namespace qx
