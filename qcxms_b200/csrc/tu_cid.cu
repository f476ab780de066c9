// Translation unit of cid(): collision set-up and collision loop (see qx_kernels.h).
#include "qx_kern_cid.cuh"
#include "qx_kernels.h"

namespace qx {

cudaError_t QX_CAT(tu_cid_prepare_, QX_VARIANT)(const cudaDeviceProp &prop) {
    cudaError_t e = allow_max_dynamic_smem(k_cid_init, prop);
    return e == cudaSuccess ? allow_max_dynamic_smem(k_cid_chunk, prop) : e;
}

cudaError_t QX_CAT(tu_cid_init_, QX_VARIANT)(int grid, size_t smem, cudaStream_t st, DevModel m, ScratchLayout L, double *scratch, MdConfig cfg, CidConfig cc,
                                             CidState s, int ntraj, int nuc, int icoll, int *queue) {
    k_cid_init<<<grid, QX_NT, smem, st>>>(m, L, scratch, cfg, cc, s, ntraj, nuc, icoll, queue);
    return cudaGetLastError();
}

cudaError_t QX_CAT(tu_cid_chunk_, QX_VARIANT)(int grid, size_t smem, cudaStream_t st, DevModel m, ScratchLayout L, double *scratch, MdConfig cfg, CidConfig cc,
                                              CidState s, int ntraj, int nuc, int chunk, int *queue) {
    k_cid_chunk<<<grid, QX_NT, smem, st>>>(m, L, scratch, cfg, cc, s, ntraj, nuc, chunk, queue);
    return cudaGetLastError();
}

}  // namespace qx
