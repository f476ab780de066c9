// Host-side view of the ensemble kernels.  Every heavy kernel lives in its own translation unit (tu_*.cu) so that the
// library builds in parallel, and every translation unit can be compiled for more than one CTA width (QX_NT): the host
// picks a KernelSet per composition (see context_init in cabi.cu).
#pragma once
#include <cuda_runtime.h>

#include "qx_cid.cuh"

namespace qx {

struct KernelSet {
    int nt;   // threads per CTA the kernels of this set were compiled for
    // single points (tu_egrad.cu)
    cudaError_t (*prepare_egrad)(const cudaDeviceProp &);
    cudaError_t (*egrad_batch)(int grid, size_t smem, cudaStream_t st, DevModel m, ScratchLayout L, double *scratch, const double *xyz, double kt,
                               int nsys, int *queue, double *energy, double *grad, double *qat, int *stat, int *niter, double *spec);
    cudaError_t (*egrad_cycles)(unsigned long long *ph, unsigned long long *sub, unsigned long long *hist);
    // md(), EI mode (tu_md.cu) and mean-free-path mode (tu_mfp.cu)
    cudaError_t (*prepare_md)(const cudaDeviceProp &);
    cudaError_t (*md_occupancy)(int *per_sm, size_t smem);
    cudaError_t (*md_init)(int grid, size_t smem, cudaStream_t st, DevModel m, ScratchLayout L, double *scratch, MdConfig cfg, MdState s, int ntraj,
                           int *queue);
    cudaError_t (*md_chunk)(int grid, size_t smem, cudaStream_t st, DevModel m, ScratchLayout L, double *scratch, MdConfig cfg, MdState s, int ntraj,
                            int chunk, int nsub, int step_limit, int *queue, int *progress, unsigned long long *steps_done);
    cudaError_t (*md_cycles)(unsigned long long *ph, unsigned long long *sub, unsigned long long *hist);
    cudaError_t (*prepare_mfp)(const cudaDeviceProp &);
    cudaError_t (*mfp_chunk)(int grid, size_t smem, cudaStream_t st, DevModel m, ScratchLayout L, double *scratch, MdConfig cfg, MdState s, int ntraj,
                             int chunk, int nsub, int step_limit, int *queue, int *progress, unsigned long long *steps_done);
    // cid() (tu_cid.cu)
    cudaError_t (*prepare_cid)(const cudaDeviceProp &);
    cudaError_t (*cid_init)(int grid, size_t smem, cudaStream_t st, DevModel m, ScratchLayout L, double *scratch, MdConfig cfg, CidConfig cc,
                            CidState s, int ntraj, int nuc, int icoll, int *queue);
    cudaError_t (*cid_chunk)(int grid, size_t smem, cudaStream_t st, DevModel m, ScratchLayout L, double *scratch, MdConfig cfg, CidConfig cc,
                             CidState s, int ntraj, int nuc, int chunk, int *queue);
};

#define QX_CAT2(a, b) a##b
#define QX_CAT(a, b) QX_CAT2(a, b)

// dynamic shared memory limit of a kernel = what the device allows next to the kernel's static shared memory
template <class K>
static inline cudaError_t allow_max_dynamic_smem(K kernel, const cudaDeviceProp &prop) {
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, kernel);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin - (int)a.sharedSizeBytes);
}

// profiling builds (-DQX_PROFILE_PHASES): add the translation unit's counters to ph[16] / sub[16] / hist[64] and reset them
#ifdef QX_PROFILE_PHASES
#define QX_DEFINE_PHASE_READER(name)                                                                \
    cudaError_t name(unsigned long long *ph, unsigned long long *sub, unsigned long long *hist) {   \
        unsigned long long a[16], b[16], c[64], z[64] = {0};                                        \
        cudaError_t e = cudaMemcpyFromSymbol(a, g_phase_cycles, sizeof(a));                         \
        if (e == cudaSuccess) e = cudaMemcpyFromSymbol(b, g_sub_cycles, sizeof(b));                 \
        if (e == cudaSuccess) e = cudaMemcpyFromSymbol(c, g_sweep_hist, sizeof(c));                 \
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(a));                 \
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_sub_cycles, z, sizeof(b));                   \
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_sweep_hist, z, sizeof(c));                   \
        for (int i = 0; i < 16; ++i) { ph[i] += a[i]; sub[i] += b[i]; }                             \
        for (int i = 0; i < 64; ++i) hist[i] += c[i];                                               \
        return e;                                                                                   \
    }
#else
#define QX_DEFINE_PHASE_READER(name) \
    cudaError_t name(unsigned long long *, unsigned long long *, unsigned long long *) { return cudaSuccess; }
#endif

// entry points of the translation units; V = QX_VARIANT of the build (nt320: the default two-CTAs-per-SM kernels)
#define QX_DECLARE_TU_ENTRIES(V)                                                                                                              \
    cudaError_t QX_CAT(tu_egrad_prepare_, V)(const cudaDeviceProp &);                                                                         \
    cudaError_t QX_CAT(tu_egrad_launch_, V)(int, size_t, cudaStream_t, DevModel, ScratchLayout, double *, const double *, double, int, int *, \
                                            double *, double *, double *, int *, int *, double *);                                           \
    cudaError_t QX_CAT(tu_egrad_cycles_, V)(unsigned long long *, unsigned long long *, unsigned long long *);                                \
    cudaError_t QX_CAT(tu_md_prepare_, V)(const cudaDeviceProp &);                                                                            \
    cudaError_t QX_CAT(tu_md_occupancy_, V)(int *, size_t);                                                                                   \
    cudaError_t QX_CAT(tu_md_init_, V)(int, size_t, cudaStream_t, DevModel, ScratchLayout, double *, MdConfig, MdState, int, int *);          \
    cudaError_t QX_CAT(tu_md_chunk_, V)(int, size_t, cudaStream_t, DevModel, ScratchLayout, double *, MdConfig, MdState, int, int, int, int,  \
                                        int *, int *, unsigned long long *);                                                                  \
    cudaError_t QX_CAT(tu_md_cycles_, V)(unsigned long long *, unsigned long long *, unsigned long long *);                                   \
    cudaError_t QX_CAT(tu_mfp_prepare_, V)(const cudaDeviceProp &);                                                                           \
    cudaError_t QX_CAT(tu_mfp_chunk_, V)(int, size_t, cudaStream_t, DevModel, ScratchLayout, double *, MdConfig, MdState, int, int, int, int, \
                                         int *, int *, unsigned long long *);                                                                 \
    cudaError_t QX_CAT(tu_cid_prepare_, V)(const cudaDeviceProp &);                                                                           \
    cudaError_t QX_CAT(tu_cid_init_, V)(int, size_t, cudaStream_t, DevModel, ScratchLayout, double *, MdConfig, CidConfig, CidState, int,     \
                                        int, int, int *);                                                                                     \
    cudaError_t QX_CAT(tu_cid_chunk_, V)(int, size_t, cudaStream_t, DevModel, ScratchLayout, double *, MdConfig, CidConfig, CidState, int,    \
                                         int, int, int *);

#define QX_KERNEL_SET(V, NT)                                                                                                                    \
    {NT, QX_CAT(tu_egrad_prepare_, V), QX_CAT(tu_egrad_launch_, V), QX_CAT(tu_egrad_cycles_, V), QX_CAT(tu_md_prepare_, V),                      \
     QX_CAT(tu_md_occupancy_, V), QX_CAT(tu_md_init_, V), QX_CAT(tu_md_chunk_, V), QX_CAT(tu_md_cycles_, V), QX_CAT(tu_mfp_prepare_, V),        \
     QX_CAT(tu_mfp_chunk_, V), QX_CAT(tu_cid_prepare_, V), QX_CAT(tu_cid_init_, V), QX_CAT(tu_cid_chunk_, V)}

QX_DECLARE_TU_ENTRIES(nt320)
QX_DECLARE_TU_ENTRIES(nt576)
QX_DECLARE_TU_ENTRIES(nt512)

}  // namespace qx
