// Translation unit of md() in EI mode: set-up kernel (also used by the mean-free-path mode) and the MD loop (see qx_kernels.h).
#define QX_TU_MD_INIT
#include "qx_kern_md.cuh"
#include "qx_kernels.h"

namespace qx {

cudaError_t QX_CAT(tu_md_prepare_, QX_VARIANT)(const cudaDeviceProp &prop) {
    cudaError_t e = allow_max_dynamic_smem(k_md_init, prop);
    return e == cudaSuccess ? allow_max_dynamic_smem(k_md_chunk<false>, prop) : e;
}

cudaError_t QX_CAT(tu_md_occupancy_, QX_VARIANT)(int *per_sm, size_t smem) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, k_md_chunk<false>, QX_NT, smem);
}

cudaError_t QX_CAT(tu_md_init_, QX_VARIANT)(int grid, size_t smem, cudaStream_t st, DevModel m, ScratchLayout L, double *scratch, MdConfig cfg, MdState s,
                                            int ntraj, int *queue) {
    k_md_init<<<grid, QX_NT, smem, st>>>(m, L, scratch, cfg, s, ntraj, queue);
    return cudaGetLastError();
}

cudaError_t QX_CAT(tu_md_chunk_, QX_VARIANT)(int grid, size_t smem, cudaStream_t st, DevModel m, ScratchLayout L, double *scratch, MdConfig cfg, MdState s,
                                             int ntraj, int chunk, int nsub, int step_limit, int *queue, int *progress, unsigned long long *steps_done) {
    k_md_chunk<false><<<grid, QX_NT, smem, st>>>(m, L, scratch, cfg, s, ntraj, chunk, nsub, step_limit, queue, progress, steps_done);
    return cudaGetLastError();
}

QX_DEFINE_PHASE_READER(QX_CAT(tu_md_cycles_, QX_VARIANT))

}  // namespace qx
