// Translation unit of md() in the mean-free-path mode of a CID run (see qx_kernels.h).
#include "qx_kern_md.cuh"
#include "qx_kernels.h"

namespace qx {

cudaError_t QX_CAT(tu_mfp_prepare_, QX_VARIANT)(const cudaDeviceProp &prop) { return allow_max_dynamic_smem(k_md_chunk<true>, prop); }

cudaError_t QX_CAT(tu_mfp_chunk_, QX_VARIANT)(int grid, size_t smem, cudaStream_t st, DevModel m, ScratchLayout L, double *scratch, MdConfig cfg, MdState s,
                                              int ntraj, int chunk, int nsub, int step_limit, int *queue, int *progress, unsigned long long *steps_done) {
    k_md_chunk<true><<<grid, QX_NT, smem, st>>>(m, L, scratch, cfg, s, ntraj, chunk, nsub, step_limit, queue, progress, steps_done);
    return cudaGetLastError();
}

}  // namespace qx
