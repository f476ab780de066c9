// C ABI of qcxms_b200 (include/qcxms_b200.h) and the ensemble kernels.
//
// Kernels are persistent: a grid of (resident CTAs per SM x SM count) CTAs pulls trajectories from an
// atomic work queue; one CTA owns one trajectory at a time and keeps the SCC matrices in shared memory
// and the integral slab in its private, L2-resident global scratch.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/qcxms_b200.h"
#include "qx_host_model.h"
#include "qx_kernels.h"

using namespace qx;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CUDA_OK(expr)                                                                                      \
    do {                                                                                                   \
        cudaError_t e__ = (expr);                                                                          \
        if (e__ != cudaSuccess) return fail(QCXMS_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// ------------------------------------------------------------------------------------ light kernels (the heavy ones: tu_*.cu)
__global__ void __launch_bounds__(QX_NT) k_fragments(DevModel m, const double *xyz, double rcut, int nsys, int *frag, unsigned char *conn, int *stack) {
    const int nat = m.nat;
    for (int t = blockIdx.x; t < nsys; t += gridDim.x)
        md_fragments(m, xyz + (size_t)t * 3 * nat, rcut, conn + (size_t)blockIdx.x * nat * nat, frag + (size_t)t * nat, stack + (size_t)blockIdx.x * nat);
}

// intenergy (reference src/md.f90:715-741): one thread per trajectory, atoms in the reference's order
__global__ void k_intenergy(DevModel m, MdState st, int ntraj, double *fragT, double *e_int) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntraj) return;
    const int nat = m.nat;
    const int *list = st.list + (size_t)t * nat;
    const double *velo = st.velo + (size_t)t * 3 * nat;
    double e[10];
    int n[10];
    for (int i = 0; i < 10; ++i) { e[i] = 0.0; n[i] = 0; }
    for (int i = 0; i < nat; ++i) {
        const int j = list[i] - 1;
        if (j < 0 || j >= 10) continue;
        const double v2 = __dadd_rn(__dadd_rn(__dmul_rn(velo[3 * i], velo[3 * i]), __dmul_rn(velo[3 * i + 1], velo[3 * i + 1])), __dmul_rn(velo[3 * i + 2], velo[3 * i + 2]));
        e[j] = __dadd_rn(e[j], __dmul_rn(__dmul_rn(0.5, m.mass[i]), v2));
        n[j] += 1;
    }
    for (int i = 0; i < 10; ++i) {
        e_int[(size_t)t * 10 + i] = e[i];
        fragT[(size_t)t * 10 + i] = n[i] > 0 ? e[i] / (0.5 * 3 * n[i] * QC_KB) : 0.0;
    }
}

// mass number of the most abundant isotope, H .. Ar
__constant__ int c_nominal_mass[19] = {0, 1, 4, 7, 9, 11, 12, 14, 16, 19, 20, 23, 24, 27, 28, 31, 32, 35, 40};

__global__ void k_histogram(DevModel m, MdState st, int ntraj, int nbins, double *bins) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntraj || st.status[t] != TRJ_FINISHED || !st.mdok[t]) return;
    const int *list = st.list + (size_t)t * m.nat;
    for (int f = 1; f <= 10; ++f) {
        int nominal = 0, natoms = 0;
        for (int i = 0; i < m.nat; ++i) {
            if (list[i] != f) continue;
            const double amu = m.mass[i] * QC_AUTOAMU;
            const int z = m.num[i], std_mass = z <= 18 ? c_nominal_mass[z] : (int)llrint(amu);
            // an isotope label (reference imass) shows as a mass far from the element's natural average
            nominal += fabs(amu - std_mass) > 1.6 ? (int)llrint(amu) : std_mass;
            natoms += 1;
        }
        if (natoms > 0 && nominal < nbins) atomicAdd(&bins[nominal], 1.0);
    }
}

// ------------------------------------------------------------------------------------ host side
struct Context {
    HostModel hm;
    ScratchLayout L{};
    int ncta = 0, device = 0;
    size_t smem = 0;
    double *d_scratch = nullptr;
    int *d_queue = nullptr;
    std::vector<int32_t> key;
    const KernelSet *ks = nullptr;   // kernels compiled for the CTA width this composition runs with
};

static const KernelSet KS_NT320 = QX_KERNEL_SET(nt320, 320);   // two CTAs per SM: the default
static const KernelSet KS_NT576 = QX_KERNEL_SET(nt576, 576);   // one wide CTA per SM
static const KernelSet KS_NT512 = QX_KERNEL_SET(nt512, 512);   // one CTA per SM with 128 registers per thread: large bases

static int context_init(Context &c, int nat, const int32_t *num, const double *mass, int charge, int multiplicity, int device, int nwork, int method = QCXMS_B200_GFN2) {
    CUDA_OK(cudaSetDevice(device));
    std::string why = build_host_model(c.hm, nat, num, mass, charge, multiplicity, method);
    if (!why.empty()) return fail(QCXMS_B200_ERR_UNSUPPORTED, why);
    CUDA_OK(upload_model(c.hm));
    c.device = device;
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    c.smem = (smem_doubles(c.hm.nat, c.hm.nsh, c.hm.nao, c.hm.ld, c.hm.rows8, 0, c.hm.ntype) + 11 * (size_t)c.hm.nat + 16) * sizeof(double) + 64;
    if (c.smem + 2048 > (size_t)prop.sharedMemPerBlockOptin) {   // (2 KB: the kernels' static shared memory)
        // large basis (nao >~ 110): keep the two SCC matrices in the per-CTA global slab -- functional, slower
        c.hm.dev.mat_in_global = 1;
        c.smem = (smem_doubles(c.hm.nat, c.hm.nsh, c.hm.nao, c.hm.ld, c.hm.rows8, 1, c.hm.ntype) + 11 * (size_t)c.hm.nat + 16) * sizeof(double) + 64;
        if (c.smem + 2048 > (size_t)prop.sharedMemPerBlockOptin)
            return fail(QCXMS_B200_ERR_UNSUPPORTED, "system too large for the per-CTA working set (nat = " + std::to_string(c.hm.nat) + ")");
        // The rest of the shared memory -- and, underneath it, the phase-local scratch vectors bsol/pop/d4u, dead while the
        // eigensolver or a GEMM runs (carve(), qx_device.cuh) -- is the block buffer of the blocked Jacobi: 2 * jblock rows of the SCC
        // matrix.  The MD/CID kernels' per-trajectory vectors follow the buffer (DevModel::extras_off).
        const size_t vec = smem_doubles(c.hm.nat, c.hm.nsh, c.hm.nao, c.hm.ld, c.hm.rows8, 1, c.hm.ntype);
        const size_t extras = 11 * (size_t)c.hm.nat + 16;
        const size_t jblk_off = (vec - smem_tail_doubles(c.hm.nat, c.hm.nao, c.hm.ntype) + 1) & ~(size_t)1;
        const size_t limit = ((size_t)prop.sharedMemPerBlockOptin - 2048 - 64) / sizeof(double);   // 2 KB: the kernels' static shared memory
        const size_t cap = limit - jblk_off - extras - 4;
        const char *force_cta = getenv("QCXMS_B200_CTA");
        const int wide = force_cta ? atoi(force_cta) : 512;                        // kernel set of this path (chosen below)
        const int slots = (wide == 288 || wide == 320 ? 320 : wide == 576 ? 576 : 512) / 16;          // sixteen-lane groups of the CTA
        int jb = (int)(cap / (2 * (size_t)c.hm.ld));
        // the pairs of a round of two blocks (jb of them, jb - 1/2 in the first block round) should fill whole passes of `slots` groups:
        // a half-empty second pass every round costs more than the extra block copies of smaller blocks
        if (slots == 20) jb = jb >= 40 ? 40 : (jb > 20 ? 20 : jb);
        else if (jb > slots) jb = slots;
        if (const char *force = getenv("QCXMS_B200_JBLOCK")) { const int f = atoi(force); if (f >= 8 && f <= jb) jb = f; }   // measurement hook
        c.hm.dev.extras_off = (int)(vec + 8);
        if (jb >= 8 && c.hm.nao <= 316) {   // jacobi_rows_blocked: ld <= 320
            const int nb = (c.hm.nao + jb - 1) / jb;
            jb = (c.hm.nao + nb - 1) / nb;          // balanced blocks
            c.hm.dev.jblock = jb;
            size_t end = jblk_off + 2 * (size_t)jb * c.hm.ld;
            if (end < vec) end = vec;
            c.hm.dev.extras_off = (int)((end + 3) & ~(size_t)1);
            c.smem = ((size_t)c.hm.dev.extras_off + extras) * sizeof(double) + 64;
        }
    } else
        c.hm.dev.extras_off = (int)(smem_doubles(c.hm.nat, c.hm.nsh, c.hm.nao, c.hm.ld, c.hm.rows8, 0, c.hm.ntype) + 8);
    {   // the last Jacobi sweep as DMMA products (jacobi_polish / jacobi_polish_gen); QCXMS_B200_POLISH=0: classical sweeps only
        const char *pol = getenv("QCXMS_B200_POLISH");
        c.hm.dev.polish = c.hm.nao >= 16 && !(pol && atoi(pol) == 0);
    }
    c.L = make_layout(c.hm);
    // the device maximum, not this composition's size: host threads set up different compositions concurrently
    // Kernel set: two 320-thread CTAs per SM fill the device best when there are more trajectories than CTA slots.  An ensemble
    // with at most one trajectory per SM (BASELINE config 2 dealt over 8 GPUs: 125 per device) is bound by the latency of a
    // single trajectory's step instead: it runs on 576-thread CTAs, one per SM (shared-memory-resident bases only).
    c.ks = &KS_NT320;
    {
        const char *force = getenv("QCXMS_B200_CTA");   // test / measurement hook: 320 (or 288: the narrow set), 512, 576
        int want = force ? atoi(force) : 0;
        if (want == 288) want = 320;
        if (!c.hm.dev.mat_in_global) {
            // medium bases (72 < nao <~ 110): the working set allows one CTA per SM whatever its width, so it is a wide one with 128
            // registers per thread (the 43 - 52 row pairs of a Jacobi round in one pass of eight-lane groups instead of two, more warps
            // for everything else).  C14H30, 86 AOs: 12.5 k -> 19.1 k, C17H36, 104 AOs: 8.5 k -> 12.9 k single points/s; 576 threads
            // with their 96 registers: 17.7 / 10.6 k
            const bool medium = 2 * (c.smem + 1024) > (size_t)prop.sharedMemPerMultiprocessor;
            if (want == 576) c.ks = &KS_NT576;
            else if (want == 512) c.ks = &KS_NT512;
            else if (want == 0 && medium) c.ks = &KS_NT512;
            else if (want == 0 && nwork > 0 && nwork <= prop.multiProcessorCount) c.ks = &KS_NT576;
        }
        // large bases run one CTA per SM anyway (L2 residency of the SCC matrices, below): the 512-thread CTA rotates 32 row pairs per
        // pass of the blocked Jacobi instead of 18, holds them in its 128 registers per thread, and gives the staged GEMMs 16 warps
        if (c.hm.dev.mat_in_global && c.hm.dev.jblock > 0 && want != 320) c.ks = want == 576 ? &KS_NT576 : &KS_NT512;
    }
    // The wide CTAs have the SM to themselves: three more shared-memory matrices fit, and with them the GEMM-based eigenpair
    // refinement (qx_oa.cuh) that takes the one-sided Jacobi's latency chain out of the SCC (QCXMS_B200_OA=0 keeps the Jacobi).
    if (c.ks == &KS_NT576 && oa_supported(c.hm.nat, c.hm.nao, c.hm.ntype, 576)) {
        const char *off = getenv("QCXMS_B200_OA");
        const size_t smem_oa = (smem_doubles(c.hm.nat, c.hm.nsh, c.hm.nao, c.hm.ld, c.hm.rows8, 0, c.hm.ntype, 1) + 11 * (size_t)c.hm.nat + 16) * sizeof(double) + 64;
        if (!(off && atoi(off) == 0) && smem_oa + 2048 <= (size_t)prop.sharedMemPerBlockOptin) {
            c.hm.dev.oa = 1;
            const char *stop = getenv("QCXMS_B200_OA_STOP");   // measurement hook
            c.hm.dev.oa_stop = stop ? atof(stop) : QX_OA_STOP;
            const char *kap = getenv("QCXMS_B200_OA_KAPPA");
            c.hm.dev.oa_kappa = kap ? atof(kap) : QX_OA_KAPPA;
            c.smem = smem_oa;
            c.hm.dev.extras_off = (int)(smem_doubles(c.hm.nat, c.hm.nsh, c.hm.nao, c.hm.ld, c.hm.rows8, 0, c.hm.ntype, 1) + 8);
        }
    }
    CUDA_OK(c.ks->prepare_egrad(prop));
    CUDA_OK(c.ks->prepare_md(prop));
    CUDA_OK(c.ks->prepare_mfp(prop));
    int per_sm = 0;
    CUDA_OK(c.ks->md_occupancy(&per_sm, c.smem));
    if (per_sm < 1) per_sm = 1;
    // global-slab mode: the Jacobi matrix of every resident CTA has to stay in L2 (two resident CTAs per SM thrashed it for C32H66:
    // 608 instead of 778 single points/s and 0.45 GB of DRAM write-back per single point)
    if (c.hm.dev.mat_in_global && per_sm > 1 &&
        (size_t)per_sm * prop.multiProcessorCount * c.hm.rows8 * c.hm.ld * sizeof(double) > (size_t)prop.l2CacheSize / 2)
        per_sm = 1;
    c.ncta = per_sm * prop.multiProcessorCount;
    if (nwork > 0 && c.ncta > nwork) c.ncta = nwork;
    CUDA_OK(cudaMalloc(&c.d_scratch, c.L.total * sizeof(double) * (size_t)c.ncta));
    CUDA_OK(cudaMalloc(&c.d_queue, sizeof(int)));
    return 0;
}

static void context_free(Context &c) {
    if (c.d_scratch) cudaFree(c.d_scratch);
    if (c.d_queue) cudaFree(c.d_queue);
    if (c.hm.d_blob) cudaFree(c.hm.d_blob);
    c.d_scratch = nullptr; c.d_queue = nullptr; c.hm.d_blob = nullptr;
}

// a cached calculator of the level-1 entry points with its device I/O block and stream
#define QX_EGRAD_SLOTS 6
struct EgradSlot {
    Context ctx;
    void *d_io = nullptr;
    size_t cap = 0;
    cudaStream_t stream = nullptr;
    unsigned long long used = 0;
};

static int egrad_batch_impl(int nsys, int nat, const int32_t *num, const double *xyz, int charge, int multiplicity, int method_id,
                            double etemp, double *qat, double *energy, double *gradient, int32_t *stat, int32_t *niter,
                            std::vector<double> *spec, int *nao_out) {
    if (nsys < 1 || nat < 1 || !num || !xyz || !qat || !energy || !gradient || !stat) return fail(QCXMS_B200_ERR_ARG, "null or empty argument");
    if (method_id != QCXMS_B200_GFN2 && method_id != QCXMS_B200_GFN1) {
        // reference: unknown method -> stat = 5, outputs untouched (src/tblite.f90:114-120).  IPEA1 (id 11) shares the GFN1 model with
        // another element table that is not reconstructed: it reports "unknown method" as well.
        for (int i = 0; i < nsys; ++i) stat[i] = QCXMS_B200_STAT_UNKNOWN_METHOD;
        return 0;
    }
    // The reference calls this entry point once per MD step and, in between, for the fragments of analyse() (src/iniqm.f90:393):
    // the calculators of the last few compositions stay alive (least recently used one is replaced), each with its own device
    // I/O buffers (grown on demand, never freed per call) and stream.
    static std::mutex mtx;
    static EgradSlot slots[QX_EGRAD_SLOTS];
    static unsigned long long tick = 0;
    std::lock_guard<std::mutex> lock(mtx);
    std::vector<int32_t> key(num, num + nat);
    key.push_back(charge); key.push_back(multiplicity); key.push_back(nsys > 1 ? 1 << 20 : 1); key.push_back(method_id);
    for (const char *hook : {"QCXMS_B200_CTA", "QCXMS_B200_OA", "QCXMS_B200_POLISH", "QCXMS_B200_JBLOCK"}) {   // measurement hooks read by context_init
        const char *v = getenv(hook);
        key.push_back(v ? atoi(v) + 1 : 0);
    }
    int dev = 0;
    cudaGetDevice(&dev);
    EgradSlot *sl = nullptr;
    for (auto &c : slots)
        if (c.ctx.key == key && c.ctx.device == dev) sl = &c;
    if (!sl) {
        sl = &slots[0];
        for (auto &c : slots)
            if (c.used < sl->used) sl = &c;
        context_free(sl->ctx);
        sl->ctx.key.clear();
        int rc = context_init(sl->ctx, nat, num, nullptr, charge, multiplicity, dev, nsys > 1 ? 0 : 1, method_id);
        if (rc) { context_free(sl->ctx); return rc; }
        sl->ctx.key = key;
        if (!sl->stream) CUDA_OK(cudaStreamCreateWithFlags(&sl->stream, cudaStreamNonBlocking));
    }
    sl->used = ++tick;
    Context &ctx = sl->ctx;
    const size_t n3 = (size_t)nsys * nat * 3, n1 = (size_t)nsys * nat;
    const size_t nspec = spec ? (size_t)nsys * (2 * ctx.hm.nao + (size_t)ctx.hm.nao * nat + 1) : 0;
    // one device block: xyz | grad | qat | energy | spec | stat | niter
    const size_t ndbl = 2 * n3 + n1 + nsys + nspec, bytes = ndbl * sizeof(double) + 2 * (size_t)nsys * sizeof(int);
    if (sl->cap < bytes) {
        if (sl->d_io) cudaFree(sl->d_io);
        sl->d_io = nullptr; sl->cap = 0;
        CUDA_OK(cudaMalloc(&sl->d_io, bytes));
        sl->cap = bytes;
    }
    double *d_xyz = reinterpret_cast<double *>(sl->d_io), *d_g = d_xyz + n3, *d_q = d_g + n3, *d_e = d_q + n1, *d_spec = spec ? d_e + nsys : nullptr;
    int *d_stat = reinterpret_cast<int *>(d_xyz + ndbl), *d_nit = d_stat + nsys;
    cudaStream_t st = sl->stream;
    CUDA_OK(cudaMemcpyAsync(d_xyz, xyz, n3 * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemsetAsync(ctx.d_queue, 0, sizeof(int), st));
    const int grid = ctx.ncta < nsys ? ctx.ncta : nsys;
    if (nao_out) *nao_out = ctx.hm.nao;
    CUDA_OK(ctx.ks->egrad_batch(grid, ctx.smem, st, ctx.hm.dev, ctx.L, ctx.d_scratch, d_xyz, etemp * QC_KTOAU, nsys, ctx.d_queue, d_e, d_g, d_q, d_stat, d_nit, d_spec));
    if (spec) {
        spec->resize(nspec);
        CUDA_OK(cudaMemcpyAsync(spec->data(), d_spec, nspec * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    CUDA_OK(cudaMemcpyAsync(energy, d_e, nsys * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(gradient, d_g, n3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(qat, d_q, n1 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(stat, d_stat, nsys * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (niter) CUDA_OK(cudaMemcpyAsync(niter, d_nit, nsys * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int qcxms_b200_egrad_batch(int nsys, int nat, const int32_t *num, const double *xyz, int charge, int multiplicity, int method_id,
                                      double etemp, double *qat, double *energy, double *gradient, int32_t *stat, int32_t *niter) {
    return egrad_batch_impl(nsys, nat, num, xyz, charge, multiplicity, method_id, etemp, qat, energy, gradient, stat, niter, nullptr, nullptr);
}

extern "C" int qcxms_b200_egrad(int nat, const int32_t *num, const double *xyz, int charge, int multiplicity, int method_id, double etemp,
                                double *qat, double *energy, double *gradient, int32_t *stat) {
    return qcxms_b200_egrad_batch(1, nat, num, xyz, charge, multiplicity, method_id, etemp, qat, energy, gradient, stat, nullptr);
}

extern "C" int qcxms_b200_basis_size(int nat, const int32_t *num, int method_id, int32_t *nao) {
    if (nat < 1 || !num || !nao) return fail(QCXMS_B200_ERR_ARG, "null or empty argument");
    if (method_id != QCXMS_B200_GFN2 && method_id != QCXMS_B200_GFN1) return fail(QCXMS_B200_ERR_UNSUPPORTED, "GFN2-xTB (method id 2) and GFN1-xTB (1) are implemented");
    if (method_id == QCXMS_B200_GFN1) gfn1_ensure_loaded();
    int n = 0;
    for (int i = 0; i < nat; ++i) {
        if (num[i] < 1 || num[i] > GFN2_MAXZ) return fail(QCXMS_B200_ERR_UNSUPPORTED, "parameters are available for H-Ar");
        if (method_id == QCXMS_B200_GFN1 && !GFN1_EXTRA[num[i]].supported) return fail(QCXMS_B200_ERR_UNSUPPORTED, "no GFN1-xTB parameters for this element");
        const gfn2_elem_t &e = method_id == QCXMS_B200_GFN1 ? GFN1_ELEM[num[i]] : GFN2_ELEM[num[i]];
        for (int k = 0; k < e.nshell; ++k) n += 2 * e.ang[k] + 1;
    }
    *nao = n;
    return 0;
}

// get_xtb_egrad with spec_calc = .true. (reference src/tblite.f90:152-164): the arrays write_qmo puts into tmp.mspec / qcxms.Mspec.tbxtb
extern "C" int qcxms_b200_egrad_spec(int nat, const int32_t *num, const double *xyz, int charge, int multiplicity, int method_id, double etemp,
                                     double *qat, double *energy, double *gradient, int32_t *stat, int32_t *nao_out, int32_t *ihomo,
                                     double *emo, double *focc, double *qmo) {
    if (!nao_out || !ihomo || !emo || !focc || !qmo) return fail(QCXMS_B200_ERR_ARG, "null spec output");
    std::vector<double> spec;
    int nao = 0;
    int rc = egrad_batch_impl(1, nat, num, xyz, charge, multiplicity, method_id, etemp, qat, energy, gradient, stat, nullptr, &spec, &nao);
    if (rc || *stat == QCXMS_B200_STAT_UNKNOWN_METHOD) return rc;
    // orbitals in ascending energy like the reference's LAPACK solver (stable for degenerate levels)
    std::vector<int> ord(nao);
    for (int k = 0; k < nao; ++k) ord[k] = k;
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return spec[a] < spec[b]; });
    for (int k = 0; k < nao; ++k) {
        const int o = ord[k];
        emo[k] = spec[o];
        focc[k] = spec[nao + o];
        // src/mo_energ.f90:41-54: + 1e-10, then normalised over the atoms
        double summa = 0.0;
        for (int j = 0; j < nat; ++j) { qmo[(size_t)k * nat + j] = spec[2 * nao + (size_t)o * nat + j] + 1.e-10; summa = summa + qmo[(size_t)k * nat + j]; }
        for (int j = 0; j < nat; ++j) qmo[(size_t)k * nat + j] = qmo[(size_t)k * nat + j] / summa;
    }
    *nao_out = nao;
    *ihomo = (int)spec[2 * nao + (size_t)nao * nat];
    return 0;
}

extern "C" int qcxms_b200_fragment_structure(int nsys, int nat, const int32_t *num, const double *xyz, double rcut, int32_t *frag) {
    if (nsys < 1 || nat < 1 || !num || !xyz || !frag) return fail(QCXMS_B200_ERR_ARG, "null or empty argument");
    // only the radii table is needed: build a minimal model
    std::vector<double> qcrad(nat), mass(nat);
    for (int i = 0; i < nat; ++i) {
        if (num[i] < 1 || num[i] > ELEM_MAXZ) return fail(QCXMS_B200_ERR_ARG, "bad atomic number");
        qcrad[i] = QC_AATOAU * QCXMS_RAD_AA[num[i]];
        mass[i] = ATOMIC_MASS_AMU[num[i]] * QC_AMUTOAU;
    }
    DevModel m{};
    m.nat = nat;
    const int grid = nsys < 1024 ? nsys : 1024;
    // one device block (freed on every path): radii | xyz | frag | stack | conn
    const size_t b_rad = nat * sizeof(double), b_xyz = (size_t)nsys * nat * 3 * sizeof(double), b_frag = (size_t)nsys * nat * sizeof(int),
                 b_stack = (size_t)grid * nat * sizeof(int), b_conn = (size_t)grid * nat * nat;
    struct Block {
        char *p = nullptr;
        ~Block() { if (p) cudaFree(p); }
    } blk;
    CUDA_OK(cudaMalloc(&blk.p, b_rad + b_xyz + b_frag + b_stack + b_conn));
    double *d_rad = reinterpret_cast<double *>(blk.p), *d_xyz = reinterpret_cast<double *>(blk.p + b_rad);
    int *d_frag = reinterpret_cast<int *>(blk.p + b_rad + b_xyz), *d_stack = reinterpret_cast<int *>(blk.p + b_rad + b_xyz + b_frag);
    unsigned char *d_conn = reinterpret_cast<unsigned char *>(blk.p + b_rad + b_xyz + b_frag + b_stack);
    CUDA_OK(cudaMemcpy(d_rad, qcrad.data(), b_rad, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(d_xyz, xyz, b_xyz, cudaMemcpyHostToDevice));
    m.at_qcrad = d_rad;
    k_fragments<<<grid, QX_NT>>>(m, d_xyz, rcut, nsys, d_frag, d_conn, d_stack);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpy(frag, d_frag, b_frag, cudaMemcpyDeviceToHost));
    return 0;
}

// ------------------------------------------------------------------------------------ ensemble
struct qcxms_b200_ensemble {
    Context ctx;
    MdConfig cfg{};
    MdState st{};
    int ntraj = 0;
    std::vector<void *> allocs;
    unsigned long long *d_steps = nullptr;
    int *d_progress = nullptr;
    double *d_bins = nullptr;
    double *qwarm_buf = nullptr;
    double *gs_buf = nullptr;
    int nbins = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool initialised = false;
    double last_ms = 0.0;
    int64_t launches = 0, scc_iters = 0;
};

template <class T>
static cudaError_t ens_alloc(qcxms_b200_ensemble *h, T **p, size_t n) {
    cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
    if (e == cudaSuccess) {
        h->allocs.push_back(*p);
        e = cudaMemset(*p, 0, n * sizeof(T));
    }
    return e;
}

extern "C" int qcxms_b200_ensemble_create(const qcxms_b200_md_config_t *cfg, int ntraj, int nat, const int32_t *num, const double *mass, int device,
                                          qcxms_b200_ensemble_t **out) {
    if (!cfg || !num || !out || ntraj < 1 || nat < 1) return fail(QCXMS_B200_ERR_ARG, "null or empty argument");
    if (cfg->method_id != QCXMS_B200_GFN2 && cfg->method_id != QCXMS_B200_GFN1) return fail(QCXMS_B200_ERR_UNSUPPORTED, "GFN2-xTB (method id 2) and GFN1-xTB (1) are implemented");
    auto *h = new qcxms_b200_ensemble();
    // multiplicity as the reference's getspin() would pass it (src/utility.f90:449-464); tblite discards it anyway
    int zsum = 0;
    for (int i = 0; i < nat; ++i) zsum += num[i];
    int j = zsum - std::abs(cfg->mchrg);
    int mult = j < 1 ? -1 : 1 + j % 2;
    int rc = context_init(h->ctx, nat, num, mass, cfg->mchrg, mult, device, ntraj, cfg->method_id);
    if (rc) { context_free(h->ctx); delete h; return rc; }
    h->ntraj = ntraj;
    h->cfg.mchrg = cfg->mchrg; h->cfg.nfragexit = cfg->nfragexit; h->cfg.exit_rules = cfg->exit_rules; h->cfg.nmax = cfg->nmax; h->cfg.isec = cfg->isec;
    h->cfg.tstep = cfg->tstep; h->cfg.etemp_in = cfg->etemp_in; h->cfg.ieetemp = cfg->ieetemp; h->cfg.ax = cfg->ax;
    h->cfg.it_mode = 1; h->cfg.tsoll = 0.0; h->cfg.method3 = 0; h->cfg.starting_md = 0;
    const size_t n1 = (size_t)ntraj * nat, n3 = n1 * 3, nt = ntraj;
    MdState &s = h->st;
    cudaError_t e = cudaSuccess;
#define EA(field, n) if (e == cudaSuccess) e = ens_alloc(h, &s.field, n)
    EA(xyz, n3); EA(velo, n3); EA(grad, n3); EA(achrg, n1); EA(velof, n1); EA(avchrg, n1); EA(avxyz, n3);
    EA(eimp, nt); EA(tadd, nt); EA(epot, nt); EA(ekin, nt); EA(ekinstart, nt); EA(etemp, nt); EA(Tav, nt); EA(Epav, nt); EA(Ekav, nt); EA(Edum, nt);
    EA(aTlast, nt); EA(dtime, nt); EA(ttime, nt); EA(fadd, nt);
    EA(nstep, nt); EA(kdump, nt); EA(fconst, nt); EA(morestep, nt); EA(nfrag, nt); EA(status, nt); EA(fragstate, nt); EA(mdok, nt); EA(nadd, nt);
    EA(list, n1); EA(scc_total, nt);
#undef EA
    if (e == cudaSuccess && h->ctx.hm.dev.oa) {   // eigenvector seeds of the eigenpair refinement, per trajectory
        const size_t per = (size_t)QX_OA_NSTORE * h->ctx.hm.nao * h->ctx.hm.nao + 1;
        e = ens_alloc(h, &s.eigseed, (size_t)ntraj * per);
    }
    if (e == cudaSuccess) e = ens_alloc(h, &h->d_steps, 1);
    if (e == cudaSuccess) e = ens_alloc(h, &h->d_progress, nt);
    if (e == cudaSuccess) e = cudaStreamCreate(&h->stream);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
    if (e != cudaSuccess) {
        std::string msg = cudaGetErrorString(e);
        qcxms_b200_ensemble_destroy(h);
        return fail(QCXMS_B200_ERR_CUDA, "ensemble allocation: " + msg);
    }
    *out = h;
    return 0;
}

extern "C" int qcxms_b200_ensemble_destroy(qcxms_b200_ensemble_t *h) {
    if (!h) return 0;
    cudaSetDevice(h->ctx.device);
    for (void *p : h->allocs) cudaFree(p);
    if (h->d_bins) cudaFree(h->d_bins);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    context_free(h->ctx);
    delete h;
    return 0;
}

extern "C" int qcxms_b200_ensemble_set_trajectory(qcxms_b200_ensemble_t *h, int itrj, const double *xyz, const double *velo, const double *velof,
                                                  double eimp, double tadd) {
    if (!h || itrj < 0 || itrj >= h->ntraj || !xyz || !velo || !velof) return fail(QCXMS_B200_ERR_ARG, "bad trajectory index or null array");
    CUDA_OK(cudaSetDevice(h->ctx.device));
    const size_t nat = h->ctx.hm.nat;
    CUDA_OK(cudaMemcpy(h->st.xyz + itrj * nat * 3, xyz, nat * 3 * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(h->st.velo + itrj * nat * 3, velo, nat * 3 * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(h->st.velof + itrj * nat, velof, nat * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(h->st.eimp + itrj, &eimp, sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(h->st.tadd + itrj, &tadd, sizeof(double), cudaMemcpyHostToDevice));
    h->initialised = false;
    return 0;
}

extern "C" int qcxms_b200_ensemble_set_all(qcxms_b200_ensemble_t *h, const double *xyz, const double *velo, const double *velof, const double *eimp,
                                           const double *tadd) {
    if (!h || !xyz || !velo || !velof || !eimp || !tadd) return fail(QCXMS_B200_ERR_ARG, "null array");
    CUDA_OK(cudaSetDevice(h->ctx.device));
    const size_t n1 = (size_t)h->ntraj * h->ctx.hm.nat;
    CUDA_OK(cudaMemcpyAsync(h->st.xyz, xyz, n1 * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaMemcpyAsync(h->st.velo, velo, n1 * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaMemcpyAsync(h->st.velof, velof, n1 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaMemcpyAsync(h->st.eimp, eimp, h->ntraj * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaMemcpyAsync(h->st.tadd, tadd, h->ntraj * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->initialised = false;
    return 0;
}

extern "C" int qcxms_b200_ensemble_set_warm_start(qcxms_b200_ensemble_t *h, int on) {
    if (!h) return fail(QCXMS_B200_ERR_ARG, "null handle");
    CUDA_OK(cudaSetDevice(h->ctx.device));
    if (on && !h->st.qwarm) {
        if (!h->qwarm_buf) {   // allocated once per handle; switching the mode off and on again reuses it
            cudaError_t e = ens_alloc(h, &h->qwarm_buf, (size_t)h->ntraj * (2 * h->ctx.hm.ndim + 1));
            if (e != cudaSuccess) return fail(QCXMS_B200_ERR_CUDA, std::string("warm-start buffer: ") + cudaGetErrorString(e));
        }
        h->st.qwarm = h->qwarm_buf;
        h->initialised = false;   // the populations are seeded by the initial single point of md()
    } else if (!on)
        h->st.qwarm = nullptr;    // (buffer stays owned by the handle)
    return 0;
}

// md() with it = -1 (equilibration) or it = 0 (sampling, every step recorded) -- the two ground-state runs the reference does before
// the production runs (src/main.F90:545-567)
extern "C" int qcxms_b200_ensemble_set_gs_mode(qcxms_b200_ensemble_t *h, int it, double tsoll) {
    if (!h || (it != 0 && it != -1 && it != 1)) return fail(QCXMS_B200_ERR_ARG, "it must be -1 (equilibration), 0 (sampling) or 1 (production)");
    if (it < 0 && !(tsoll > 0.0)) return fail(QCXMS_B200_ERR_ARG, "the equilibration needs a target temperature");
    if (it <= 0 && h->cfg.etemp_in < 0.0) return fail(QCXMS_B200_ERR_ARG, "ground-state runs need an explicit electronic temperature (etemp_in)");
    CUDA_OK(cudaSetDevice(h->ctx.device));
    h->cfg.it_mode = it;
    h->cfg.tsoll = tsoll;
    h->st.gsdump = nullptr;
    if (it == 0) {
        const size_t n = (size_t)h->ntraj * h->cfg.nmax * 6 * h->ctx.hm.nat;
        if (n * sizeof(double) > ((size_t)16 << 30)) return fail(QCXMS_B200_ERR_ARG, "ground-state record buffer would exceed 16 GB: fewer steps or trajectories");
        if (!h->gs_buf) {
            cudaError_t e = ens_alloc(h, &h->gs_buf, n);
            if (e != cudaSuccess) return fail(QCXMS_B200_ERR_CUDA, std::string("ground-state record buffer: ") + cudaGetErrorString(e));
        }
        h->st.gsdump = h->gs_buf;
    }
    h->initialised = false;
    return 0;
}

// records first .. first + count - 1 (0-based steps) of trajectory itrj: [count][nat][6] = xyz, velo of every atom (the lines of qcxms.gs)
extern "C" int qcxms_b200_ensemble_get_gs(qcxms_b200_ensemble_t *h, int itrj, int first, int count, double *xyzvelo) {
    if (!h || !xyzvelo || itrj < 0 || itrj >= h->ntraj || first < 0 || count < 1 || first + count > h->cfg.nmax) return fail(QCXMS_B200_ERR_ARG, "bad argument");
    if (!h->gs_buf || h->cfg.it_mode != 0) return fail(QCXMS_B200_ERR_ARG, "the ensemble is not in ground-state sampling mode");
    CUDA_OK(cudaSetDevice(h->ctx.device));
    const size_t rec = (size_t)6 * h->ctx.hm.nat;
    CUDA_OK(cudaMemcpy(xyzvelo, h->gs_buf + ((size_t)itrj * h->cfg.nmax + first) * rec, (size_t)count * rec * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int qcxms_b200_ensemble_set_mfp(qcxms_b200_ensemble_t *h, int icoll, const double *new_velo) {
    if (!h || icoll < 1 || !new_velo) return fail(QCXMS_B200_ERR_ARG, "icoll >= 1 and new_velo [ntraj] required");
    CUDA_OK(cudaSetDevice(h->ctx.device));
    const size_t nt = h->ntraj, n3 = nt * h->ctx.hm.nat * 3;
    if (!h->st.mfp_d) {
        cudaError_t e = ens_alloc(h, &h->st.mfp_d, nt * 8);
        if (e == cudaSuccess) e = ens_alloc(h, &h->st.mfp_i, nt * 16);
        if (e == cudaSuccess) e = ens_alloc(h, &h->st.avxyz2, n3);
        if (e == cudaSuccess) e = ens_alloc(h, &h->st.store, n3);
        if (e != cudaSuccess) return fail(QCXMS_B200_ERR_CUDA, std::string("mean-free-path buffers: ") + cudaGetErrorString(e));
    }
    std::vector<double> d(nt * 8, 0.0);
    for (size_t t = 0; t < nt; ++t) d[t * 8 + 3] = new_velo[t];   // picked up by the initial single point of md()
    CUDA_OK(cudaMemcpy(h->st.mfp_d, d.data(), d.size() * sizeof(double), cudaMemcpyHostToDevice));
    h->cfg.icoll = icoll;
    h->cfg.method3 = 1;
    h->cfg.starting_md = 0;
    h->initialised = false;
    return 0;
}

// md() as the heating MD before the first collision of an ESI/CID run: reference global method == 3, icoll = 0, starting_md = .true.,
// Tsoll = tscale (src/main.F90:1330-1362; md.f90:428-434).  eimp (E_Scale) and tadd (pretadd) come in through set_all / set_trajectory.
extern "C" int qcxms_b200_ensemble_set_esi(qcxms_b200_ensemble_t *h, double tscale) {
    if (!h || !(tscale > 0.0)) return fail(QCXMS_B200_ERR_ARG, "tscale (K) > 0 required");
    std::vector<double> zero(h->ntraj, 0.0);
    int rc = qcxms_b200_ensemble_set_mfp(h, 1, zero.data());   // allocates and clears the buffers of the method-3 branch of md()
    if (rc) return rc;
    h->cfg.icoll = 0;
    h->cfg.starting_md = 1;
    h->cfg.tsoll = tscale;
    return 0;
}

extern "C" int qcxms_b200_ensemble_get_new_velo(qcxms_b200_ensemble_t *h, double *new_velo) {
    if (!h || !new_velo) return fail(QCXMS_B200_ERR_ARG, "null argument");
    if (!h->st.mfp_d) return fail(QCXMS_B200_ERR_ARG, "ensemble is not in mean-free-path mode");
    CUDA_OK(cudaSetDevice(h->ctx.device));
    std::vector<double> d((size_t)h->ntraj * 8);
    CUDA_OK(cudaMemcpy(d.data(), h->st.mfp_d, d.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int t = 0; t < h->ntraj; ++t) new_velo[t] = d[(size_t)t * 8 + 3];
    return 0;
}

// mean-free-path mode: axyz is the averaged fragment structure once a fragmentation was counted (reference src/md.f90:694-699)
static cudaError_t mfp_fix_axyz(qcxms_b200_ensemble_t *h, size_t t0, size_t nt, double *axyz) {
    if (!h->cfg.method3 || !axyz) return cudaSuccess;
    const size_t n3 = (size_t)h->ctx.hm.nat * 3;
    std::vector<int> mi(nt * 16);
    std::vector<double> st(nt * n3);
    cudaError_t e = cudaMemcpy(mi.data(), h->st.mfp_i + t0 * 16, mi.size() * sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(st.data(), h->st.store + t0 * n3, st.size() * sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return e;
    for (size_t t = 0; t < nt; ++t)
        if (mi[t * 16 + 2] > 1)
            for (size_t i = 0; i < n3; ++i) axyz[t * n3 + i] = st[t * n3 + i];
    return cudaSuccess;
}

extern "C" int qcxms_b200_ensemble_run_md(qcxms_b200_ensemble_t *h, int max_steps, int64_t *steps_done) {
    if (!h) return fail(QCXMS_B200_ERR_ARG, "null handle");
    CUDA_OK(cudaSetDevice(h->ctx.device));
    Context &c = h->ctx;
    const int grid = c.ncta < h->ntraj ? c.ncta : h->ntraj;
    h->launches = 0;
    CUDA_OK(cudaMemsetAsync(h->d_steps, 0, sizeof(unsigned long long), h->stream));
    CUDA_OK(cudaEventRecord(h->ev0, h->stream));
    int base_step = 0;
    if (!h->initialised) {
        CUDA_OK(cudaMemsetAsync(c.d_queue, 0, sizeof(int), h->stream));
        CUDA_OK(c.ks->md_init(grid, c.smem, h->stream, c.hm.dev, c.L, c.d_scratch, h->cfg, h->st, h->ntraj, c.d_queue));
        h->launches += 1;
        h->initialised = true;
    } else {
        // continue: the step limit is relative to the steps already taken (all trajectories advance in lockstep chunks)
        std::vector<int> ns(h->ntraj);
        CUDA_OK(cudaMemcpyAsync(ns.data(), h->st.nstep, h->ntraj * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CUDA_OK(cudaStreamSynchronize(h->stream));
        for (int v : ns) base_step = v > base_step ? v : base_step;
    }
    const int limit = max_steps > 0 ? base_step + max_steps : 0;
    // mean-free-path mode: every counted fragmentation moves the end to nstep + add_steps (src/md.f90:507); the loop ends when no
    // trajectory is running any more, the bound is a safety net only
    const int total = max_steps > 0 ? max_steps : (h->cfg.method3 ? h->cfg.nmax + 64 * (h->ctx.hm.nat / 10 + 1) * 1000 : h->cfg.nmax);
    const int chunk = 64, sub_steps = 8;
    std::vector<int> status(h->ntraj);
    for (int done = 0; done < total; done += chunk) {
        const int this_chunk = total - done < chunk ? total - done : chunk;
        const int nsub = (this_chunk + sub_steps - 1) / sub_steps;
        CUDA_OK(cudaMemsetAsync(c.d_queue, 0, sizeof(int), h->stream));
        CUDA_OK(cudaMemsetAsync(h->d_progress, 0, h->ntraj * sizeof(int), h->stream));
        // the last sub-chunk may be shorter: the kernel bounds every sub-chunk by the launch's step limit as well
        CUDA_OK((h->cfg.method3 ? c.ks->mfp_chunk : c.ks->md_chunk)(grid, c.smem, h->stream, c.hm.dev, c.L, c.d_scratch, h->cfg, h->st, h->ntraj, sub_steps,
                                                                      nsub, max_steps > 0 ? limit : 0, c.d_queue, h->d_progress, h->d_steps));
        h->launches += 1;
        // poll for completion every few chunks (cheap: ntraj ints)
        if ((done / chunk) % 4 == 3 || done + chunk >= total) {
            CUDA_OK(cudaMemcpyAsync(status.data(), h->st.status, h->ntraj * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CUDA_OK(cudaStreamSynchronize(h->stream));
            bool any = false;
            for (int v : status) any = any || v == TRJ_RUNNING;
            if (!any) break;
        }
    }
    CUDA_OK(cudaEventRecord(h->ev1, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    unsigned long long steps = 0;
    CUDA_OK(cudaMemcpy(&steps, h->d_steps, sizeof(steps), cudaMemcpyDeviceToHost));
    if (steps_done) *steps_done = (int64_t)steps;
    std::vector<int> scc(h->ntraj);
    CUDA_OK(cudaMemcpy(scc.data(), h->st.scc_total, h->ntraj * sizeof(int), cudaMemcpyDeviceToHost));
    h->scc_iters = 0;
    for (int v : scc) h->scc_iters += v;
    return 0;
}

extern "C" int qcxms_b200_ensemble_get_result(qcxms_b200_ensemble_t *h, int itrj, double *xyz, double *velo, double *grad, int32_t *list, double *achrg,
                                              double *axyz, qcxms_b200_md_result_t *res) {
    if (!h || itrj < 0 || itrj >= h->ntraj) return fail(QCXMS_B200_ERR_ARG, "bad trajectory index");
    CUDA_OK(cudaSetDevice(h->ctx.device));
    const size_t nat = h->ctx.hm.nat, t = itrj;
    const MdState &s = h->st;
    int nstep, kdump;
    CUDA_OK(cudaMemcpy(&nstep, s.nstep + t, sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(&kdump, s.kdump + t, sizeof(int), cudaMemcpyDeviceToHost));
    if (xyz) CUDA_OK(cudaMemcpy(xyz, s.xyz + t * nat * 3, nat * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (velo) CUDA_OK(cudaMemcpy(velo, s.velo + t * nat * 3, nat * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (grad) CUDA_OK(cudaMemcpy(grad, s.grad + t * nat * 3, nat * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (list) CUDA_OK(cudaMemcpy(list, s.list + t * nat, nat * sizeof(int), cudaMemcpyDeviceToHost));
    // averages over the last kdump steps (reference src/md.f90:688-700)
    if (achrg) {
        CUDA_OK(cudaMemcpy(achrg, s.avchrg + t * nat, nat * sizeof(double), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < nat; ++i) achrg[i] /= kdump;
    }
    if (axyz) {
        CUDA_OK(cudaMemcpy(axyz, s.avxyz + t * nat * 3, nat * 3 * sizeof(double), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < nat * 3; ++i) axyz[i] /= kdump;
        CUDA_OK(mfp_fix_axyz(h, t, 1, axyz));
    }
    if (res) {
        double Tav, Epav, Ekav, aTlast;
        CUDA_OK(cudaMemcpy(&res->mdok, s.mdok + t, sizeof(int), cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(&res->fragstate, s.fragstate + t, sizeof(int), cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(&res->nfrag, s.nfrag + t, sizeof(int), cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(&res->status, s.status + t, sizeof(int), cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(&res->scc_iter_total, s.scc_total + t, sizeof(int), cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(&Tav, s.Tav + t, sizeof(double), cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(&Epav, s.Epav + t, sizeof(double), cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(&Ekav, s.Ekav + t, sizeof(double), cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(&aTlast, s.aTlast + t, sizeof(double), cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(&res->dtime, s.dtime + t, sizeof(double), cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(&res->ttime, s.ttime + t, sizeof(double), cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(&res->Epot, s.epot + t, sizeof(double), cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(&res->Ekin, s.ekin + t, sizeof(double), cudaMemcpyDeviceToHost));
        res->nstep = nstep;
        const int div = nstep > 0 ? nstep : 1;
        res->Tav = Tav / div; res->Epav = Epav / div; res->Ekav = Ekav / div;
        res->aTlast = aTlast / (kdump > 0 ? kdump : 1);
    }
    return 0;
}

extern "C" int qcxms_b200_ensemble_get_all(qcxms_b200_ensemble_t *h, double *xyz, double *velo, double *grad, int32_t *list, double *achrg, double *axyz,
                                           qcxms_b200_md_result_t *res) {
    if (!h) return fail(QCXMS_B200_ERR_ARG, "null handle");
    CUDA_OK(cudaSetDevice(h->ctx.device));
    const size_t nat = h->ctx.hm.nat, nt = h->ntraj, n1 = nt * nat;
    const MdState &s = h->st;
    std::vector<int> nstep(nt), kdump(nt), iv(nt);
    std::vector<double> dv(nt);
    CUDA_OK(cudaMemcpy(nstep.data(), s.nstep, nt * sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(kdump.data(), s.kdump, nt * sizeof(int), cudaMemcpyDeviceToHost));
    if (xyz) CUDA_OK(cudaMemcpy(xyz, s.xyz, n1 * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (velo) CUDA_OK(cudaMemcpy(velo, s.velo, n1 * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (grad) CUDA_OK(cudaMemcpy(grad, s.grad, n1 * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (list) CUDA_OK(cudaMemcpy(list, s.list, n1 * sizeof(int), cudaMemcpyDeviceToHost));
    if (achrg) {
        CUDA_OK(cudaMemcpy(achrg, s.avchrg, n1 * sizeof(double), cudaMemcpyDeviceToHost));
        for (size_t t = 0; t < nt; ++t)
            for (size_t i = 0; i < nat; ++i) achrg[t * nat + i] /= kdump[t];
    }
    if (axyz) {
        CUDA_OK(cudaMemcpy(axyz, s.avxyz, n1 * 3 * sizeof(double), cudaMemcpyDeviceToHost));
        for (size_t t = 0; t < nt; ++t)
            for (size_t i = 0; i < 3 * nat; ++i) axyz[t * 3 * nat + i] /= kdump[t];
        CUDA_OK(mfp_fix_axyz(h, 0, nt, axyz));
    }
    if (res) {
        auto geti = [&](const int *src, int32_t qcxms_b200_md_result_t::*f) -> cudaError_t {
            cudaError_t e = cudaMemcpy(iv.data(), src, nt * sizeof(int), cudaMemcpyDeviceToHost);
            for (size_t t = 0; t < nt; ++t) res[t].*f = iv[t];
            return e;
        };
        auto getd = [&](const double *src, double qcxms_b200_md_result_t::*f) -> cudaError_t {
            cudaError_t e = cudaMemcpy(dv.data(), src, nt * sizeof(double), cudaMemcpyDeviceToHost);
            for (size_t t = 0; t < nt; ++t) res[t].*f = dv[t];
            return e;
        };
        CUDA_OK(geti(s.mdok, &qcxms_b200_md_result_t::mdok));
        CUDA_OK(geti(s.fragstate, &qcxms_b200_md_result_t::fragstate));
        CUDA_OK(geti(s.nfrag, &qcxms_b200_md_result_t::nfrag));
        CUDA_OK(geti(s.status, &qcxms_b200_md_result_t::status));
        CUDA_OK(geti(s.scc_total, &qcxms_b200_md_result_t::scc_iter_total));
        CUDA_OK(getd(s.Tav, &qcxms_b200_md_result_t::Tav));
        CUDA_OK(getd(s.Epav, &qcxms_b200_md_result_t::Epav));
        CUDA_OK(getd(s.Ekav, &qcxms_b200_md_result_t::Ekav));
        CUDA_OK(getd(s.aTlast, &qcxms_b200_md_result_t::aTlast));
        CUDA_OK(getd(s.dtime, &qcxms_b200_md_result_t::dtime));
        CUDA_OK(getd(s.ttime, &qcxms_b200_md_result_t::ttime));
        CUDA_OK(getd(s.epot, &qcxms_b200_md_result_t::Epot));
        CUDA_OK(getd(s.ekin, &qcxms_b200_md_result_t::Ekin));
        for (size_t t = 0; t < nt; ++t) {
            const int div = nstep[t] > 0 ? nstep[t] : 1;
            res[t].nstep = nstep[t];
            res[t].Tav /= div; res[t].Epav /= div; res[t].Ekav /= div;
            res[t].aTlast /= (kdump[t] > 0 ? kdump[t] : 1);
        }
    }
    return 0;
}

extern "C" int qcxms_b200_ensemble_last_timing(qcxms_b200_ensemble_t *h, double *kernel_ms, int64_t *launches, int64_t *scc_iterations) {
    if (!h) return fail(QCXMS_B200_ERR_ARG, "null handle");
    if (kernel_ms) *kernel_ms = h->last_ms;
    if (launches) *launches = h->launches;
    if (scc_iterations) *scc_iterations = h->scc_iters;
    return 0;
}

extern "C" int qcxms_b200_ensemble_intenergy(qcxms_b200_ensemble_t *h, double *fragT, double *e_int) {
    if (!h || !fragT || !e_int) return fail(QCXMS_B200_ERR_ARG, "null argument");
    CUDA_OK(cudaSetDevice(h->ctx.device));
    const size_t n = (size_t)h->ntraj * 10;
    double *d = nullptr;
    CUDA_OK(cudaMalloc(&d, 2 * n * sizeof(double)));
    k_intenergy<<<(h->ntraj + 127) / 128, 128, 0, h->stream>>>(h->ctx.hm.dev, h->st, h->ntraj, d, d + n);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaMemcpy(fragT, d, n * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(e_int, d + n, n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(QCXMS_B200_ERR_CUDA, std::string("intenergy: ") + cudaGetErrorString(e));
    return 0;
}

extern "C" int qcxms_b200_ensemble_histogram(qcxms_b200_ensemble_t *h, int nbins, double *bins_host, void **bins_device) {
    if (!h || nbins < 1) return fail(QCXMS_B200_ERR_ARG, "bad argument");
    CUDA_OK(cudaSetDevice(h->ctx.device));
    if (h->nbins != nbins) {
        if (h->d_bins) cudaFree(h->d_bins);
        CUDA_OK(cudaMalloc(&h->d_bins, nbins * sizeof(double)));
        h->nbins = nbins;
    }
    CUDA_OK(cudaMemsetAsync(h->d_bins, 0, nbins * sizeof(double), h->stream));
    k_histogram<<<(h->ntraj + 127) / 128, 128, 0, h->stream>>>(h->ctx.hm.dev, h->st, h->ntraj, nbins, h->d_bins);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(h->stream));
    if (bins_host) CUDA_OK(cudaMemcpy(bins_host, h->d_bins, nbins * sizeof(double), cudaMemcpyDeviceToHost));
    if (bins_device) *bins_device = h->d_bins;
    return 0;
}

// ------------------------------------------------------------------------------------ the collective (NCCL, loaded on first use)
namespace {
struct NcclUniqueId { char internal[QCXMS_B200_UNIQUE_ID_BYTES]; };
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclUniqueId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string why;
};
NcclApi &nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = getenv("QCXMS_B200_NCCL_LIB");
        const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
            api.why = dlerror();
        }
        if (!api.handle) return;
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle, "ncclCommInitRank"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.handle, "ncclAllReduce"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
        if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy || !api.GetErrorString) {
            api.why = "NCCL entry points missing";
            api.handle = nullptr;
        }
    });
    return api;
}
constexpr int kNcclDouble = 8, kNcclSum = 0;   // ncclFloat64, ncclSum (nccl.h)
}  // namespace

struct qcxms_b200_comm {
    void *comm = nullptr;
    int nranks = 0, rank = 0, device = 0;
    cudaStream_t stream = nullptr;
    double *d_buf = nullptr;
    size_t cap = 0;
};

#define NCCL_OK(expr)                                                                                              \
    do {                                                                                                           \
        int r__ = (expr);                                                                                          \
        if (r__ != 0) return fail(QCXMS_B200_ERR_CUDA, std::string(#expr) + ": " + nccl_api().GetErrorString(r__)); \
    } while (0)

extern "C" int qcxms_b200_comm_unique_id(void *id128) {
    if (!id128) return fail(QCXMS_B200_ERR_ARG, "null argument");
    NcclApi &api = nccl_api();
    if (!api.handle) return fail(QCXMS_B200_ERR_UNSUPPORTED, "NCCL is not available: " + api.why);
    NcclUniqueId id;
    NCCL_OK(api.GetUniqueId(&id));
    memcpy(id128, id.internal, sizeof(id.internal));
    return 0;
}

extern "C" int qcxms_b200_comm_create(const void *id128, int nranks, int rank, int device, qcxms_b200_comm_t **out) {
    if (!id128 || !out || nranks < 1 || rank < 0 || rank >= nranks) return fail(QCXMS_B200_ERR_ARG, "bad argument");
    NcclApi &api = nccl_api();
    if (!api.handle) return fail(QCXMS_B200_ERR_UNSUPPORTED, "NCCL is not available: " + api.why);
    CUDA_OK(cudaSetDevice(device));
    qcxms_b200_comm *c = new qcxms_b200_comm;
    c->nranks = nranks; c->rank = rank; c->device = device;
    NcclUniqueId id;
    memcpy(id.internal, id128, sizeof(id.internal));
    int r = api.CommInitRank(&c->comm, nranks, id, rank);
    if (r != 0) { delete c; return fail(QCXMS_B200_ERR_CUDA, std::string("ncclCommInitRank: ") + api.GetErrorString(r)); }
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { api.CommDestroy(c->comm); delete c; return fail(QCXMS_B200_ERR_CUDA, cudaGetErrorString(e)); }
    *out = c;
    // NCCL connects its channels lazily, at the first collective (22 ms on one B200, more over NVLink): pay that here, in the set-up
    // call every rank makes together, not in the first spectrum reduction of the run
    double one = 1.0;
    const int rc = qcxms_b200_comm_allreduce_sum(c, &one, 1);
    if (rc != 0) { qcxms_b200_comm_destroy(c); *out = nullptr; }
    return rc;
}

extern "C" int qcxms_b200_comm_destroy(qcxms_b200_comm_t *c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->comm) nccl_api().CommDestroy(c->comm);
    if (c->d_buf) cudaFree(c->d_buf);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

extern "C" int qcxms_b200_comm_allreduce_sum(qcxms_b200_comm_t *c, double *inout, int n) {
    if (!c || !inout || n < 1) return fail(QCXMS_B200_ERR_ARG, "bad argument");
    CUDA_OK(cudaSetDevice(c->device));
    if (c->cap < (size_t)n) {
        if (c->d_buf) cudaFree(c->d_buf);
        c->d_buf = nullptr; c->cap = 0;
        CUDA_OK(cudaMalloc(&c->d_buf, (size_t)n * sizeof(double)));
        c->cap = n;
    }
    CUDA_OK(cudaMemcpyAsync(c->d_buf, inout, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NCCL_OK(nccl_api().AllReduce(c->d_buf, c->d_buf, (size_t)n, kNcclDouble, kNcclSum, c->comm, c->stream));
    CUDA_OK(cudaMemcpyAsync(inout, c->d_buf, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int qcxms_b200_ensemble_allreduce_histogram(qcxms_b200_ensemble_t *h, qcxms_b200_comm_t *c, int nbins, double *bins_host) {
    if (!h || !c || !bins_host) return fail(QCXMS_B200_ERR_ARG, "bad argument");
    if (c->device != h->ctx.device) return fail(QCXMS_B200_ERR_ARG, "communicator and ensemble live on different devices");
    void *d_bins = nullptr;
    int rc = qcxms_b200_ensemble_histogram(h, nbins, nullptr, &d_bins);   // synchronises the ensemble's stream
    if (rc) return rc;
    NCCL_OK(nccl_api().AllReduce(d_bins, d_bins, (size_t)nbins, kNcclDouble, kNcclSum, c->comm, c->stream));
    CUDA_OK(cudaMemcpyAsync(bins_host, d_bins, (size_t)nbins * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------------------------ CID host entry
extern "C" int qcxms_b200_cid_batch(const qcxms_b200_cid_config_t *cfg, int ntraj, int nuc, const int32_t *num, const double *mass, int icoll,
                                    double *xyz, double *velo, const double *rnd, const double *velo_cm, double *direc, int32_t *collided,
                                    double *grad, double *achrg, double *axyz, int32_t *list, qcxms_b200_cid_result_t *res, int device) {
    if (!cfg || ntraj < 1 || nuc < 1 || !num || !mass || !xyz || !velo || !rnd || !direc || !collided || !grad || !achrg || !axyz || !list || !res)
        return fail(QCXMS_B200_ERR_ARG, "null or empty argument");
    if (icoll < 1 || (icoll > 1 && !velo_cm)) return fail(QCXMS_B200_ERR_ARG, "icoll >= 1; later collisions need velo_cm");
    if (cfg->method_id != QCXMS_B200_GFN2 && cfg->method_id != QCXMS_B200_GFN1) return fail(QCXMS_B200_ERR_UNSUPPORTED, "GFN2-xTB (method id 2) and GFN1-xTB (1) are implemented");
    if (cfg->gas_z != 2 && cfg->gas_z != 10 && cfg->gas_z != 18 && cfg->gas_z != 7)
        return fail(QCXMS_B200_ERR_UNSUPPORTED, "collision gas must be He, Ne, Ar or N2");
    const int ngas = cfg->gas_z == 7 ? 2 : 1;   // N2: two atoms of mass gas_mass each (reference src/cid.f90:179-180)
    const int nuc0 = nuc + ngas;
    std::vector<int32_t> num0(num, num + nuc);
    std::vector<double> mass0(mass, mass + nuc);
    for (int g = 0; g < ngas; ++g) { num0.push_back(cfg->gas_z); mass0.push_back(cfg->gas_mass); }
    int zsum = 0;
    for (int v : num0) zsum += v;
    const int j = zsum - std::abs(cfg->mchrg);
    const int mult = j < 1 ? -1 : 1 + j % 2;
    Context ctx;
    int rc = context_init(ctx, nuc0, num0.data(), mass0.data(), cfg->mchrg, mult, device, ntraj, cfg->method_id);
    if (rc) { context_free(ctx); return rc; }
    {   // the device maximum, not this composition's size: host threads set up different compositions concurrently
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, device));
        CUDA_OK(ctx.ks->prepare_cid(prop));
    }
    // own (blocking) stream: collisions of different compositions are driven from different host threads and overlap on the GPU
    cudaStream_t strm = nullptr;
    if (cudaStreamCreate(&strm) != cudaSuccess) { context_free(ctx); return fail(QCXMS_B200_ERR_CUDA, "cid: stream creation failed"); }
    MdConfig mc{};
    mc.mchrg = cfg->mchrg; mc.tstep = cfg->tstep;
    CidConfig cc{};
    cc.mchrg = cfg->mchrg; cc.gas_z = cfg->gas_z; cc.eexact = cfg->eexact; cc.manual_dist = cfg->manual_dist;
    cc.ntot = cfg->ntot > 0 ? cfg->ntot : 15000;
    cc.gas_mass = cfg->gas_mass; cc.tstep = cfg->tstep; cc.etemp = cfg->etemp <= 0.0 ? 5000.0 : cfg->etemp; cc.elab = cfg->elab; cc.ecom = cfg->ecom;
    const size_t n3 = (size_t)ntraj * nuc * 3, n30 = (size_t)ntraj * nuc0 * 3;
    std::vector<void *> allocs;
    auto dalloc = [&](size_t bytes) -> void * {
        void *p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
        cudaMemset(p, 0, bytes);
        allocs.push_back(p);
        return p;
    };
    auto cleanup = [&]() { cudaStreamSynchronize(strm); cudaStreamDestroy(strm); for (void *p : allocs) cudaFree(p); context_free(ctx); };
    auto cpy = [&](void *dst, const void *src, size_t bytes, cudaMemcpyKind kind) -> cudaError_t {
        cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, strm);
        return e == cudaSuccess ? cudaStreamSynchronize(strm) : e;
    };
    CidState st{};
    double *d_rnd, *d_vcm = nullptr;
    st.xyz = (double *)dalloc(n3 * 8); st.velo = (double *)dalloc(n3 * 8); st.direc = (double *)dalloc((size_t)ntraj * 3 * 8);
    d_rnd = (double *)dalloc((size_t)ntraj * 9 * 8); d_vcm = (double *)dalloc((size_t)ntraj * 8);
    st.xyz0 = (double *)dalloc(n30 * 8); st.velo0 = (double *)dalloc(n30 * 8); st.grad0 = (double *)dalloc(n30 * 8); st.achrg0 = (double *)dalloc((size_t)ntraj * nuc0 * 8);
    st.avxyz = (double *)dalloc(n3 * 8); st.avxyz2 = (double *)dalloc(n3 * 8); st.store = (double *)dalloc(n3 * 8);
    st.list = (int *)dalloc((size_t)ntraj * nuc * 4);
    st.sc = (CidScalars *)dalloc((size_t)ntraj * sizeof(CidScalars));
    if (!st.xyz || !st.velo || !st.direc || !d_rnd || !d_vcm || !st.xyz0 || !st.velo0 || !st.grad0 || !st.achrg0 || !st.avxyz || !st.avxyz2 || !st.store ||
        !st.list || !st.sc) { cleanup(); return fail(QCXMS_B200_ERR_CUDA, "cid: device allocation failed"); }
    st.rnd = d_rnd; st.velo_cm_in = d_vcm;
    std::vector<CidScalars> hsc(ntraj);
    for (int t = 0; t < ntraj; ++t) { hsc[t] = CidScalars{}; hsc[t].collided = collided[t] ? 1 : 0; }
#define CID_OK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); return fail(QCXMS_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)
    CID_OK(cpy(st.xyz, xyz, n3 * 8, cudaMemcpyHostToDevice));
    CID_OK(cpy(st.velo, velo, n3 * 8, cudaMemcpyHostToDevice));
    CID_OK(cpy(st.direc, direc, (size_t)ntraj * 3 * 8, cudaMemcpyHostToDevice));
    CID_OK(cpy(d_rnd, rnd, (size_t)ntraj * 9 * 8, cudaMemcpyHostToDevice));
    if (velo_cm) CID_OK(cpy(d_vcm, velo_cm, (size_t)ntraj * 8, cudaMemcpyHostToDevice));
    CID_OK(cpy(st.sc, hsc.data(), (size_t)ntraj * sizeof(CidScalars), cudaMemcpyHostToDevice));
    const int grid = ctx.ncta < ntraj ? ctx.ncta : ntraj;
    CID_OK(cudaMemsetAsync(ctx.d_queue, 0, sizeof(int), strm));
    CID_OK(ctx.ks->cid_init(grid, ctx.smem, strm, ctx.hm.dev, ctx.L, ctx.d_scratch, mc, cc, st, ntraj, nuc, icoll, ctx.d_queue));
    const int chunk = 32;
    for (int done = 0; done < cc.ntot + chunk; done += chunk) {
        CID_OK(cudaMemsetAsync(ctx.d_queue, 0, sizeof(int), strm));
        CID_OK(ctx.ks->cid_chunk(grid, ctx.smem, strm, ctx.hm.dev, ctx.L, ctx.d_scratch, mc, cc, st, ntraj, nuc, chunk, ctx.d_queue));
        if ((done / chunk) % 4 == 3 || done + chunk >= cc.ntot) {
            CID_OK(cpy(hsc.data(), st.sc, (size_t)ntraj * sizeof(CidScalars), cudaMemcpyDeviceToHost));
            bool any = false;
            for (const CidScalars &v : hsc) any = any || v.status == TRJ_RUNNING;
            if (!any) break;
        }
    }
    CID_OK(cpy(hsc.data(), st.sc, (size_t)ntraj * sizeof(CidScalars), cudaMemcpyDeviceToHost));
    // hand-back (reference src/cid.f90:1058-1105): the ion part of the collision system; set-up changes to xyz / velo stay
    // visible even when the first single point failed, as in the reference
    std::vector<double> hx(n30), hv(n30), hg(n30), hq((size_t)ntraj * nuc0), hav(n3), hst(n3);
    CID_OK(cpy(hx.data(), st.xyz0, n30 * 8, cudaMemcpyDeviceToHost));
    CID_OK(cpy(hv.data(), st.velo0, n30 * 8, cudaMemcpyDeviceToHost));
    CID_OK(cpy(hg.data(), st.grad0, n30 * 8, cudaMemcpyDeviceToHost));
    CID_OK(cpy(hq.data(), st.achrg0, (size_t)ntraj * nuc0 * 8, cudaMemcpyDeviceToHost));
    CID_OK(cpy(hav.data(), st.avxyz, n3 * 8, cudaMemcpyDeviceToHost));
    CID_OK(cpy(hst.data(), st.store, n3 * 8, cudaMemcpyDeviceToHost));
    CID_OK(cpy(list, st.list, (size_t)ntraj * nuc * 4, cudaMemcpyDeviceToHost));
    CID_OK(cpy(direc, st.direc, (size_t)ntraj * 3 * 8, cudaMemcpyDeviceToHost));
    std::vector<double> sx(n3), sv(n3);
    CID_OK(cpy(sx.data(), st.xyz, n3 * 8, cudaMemcpyDeviceToHost));
    CID_OK(cpy(sv.data(), st.velo, n3 * 8, cudaMemcpyDeviceToHost));
#undef CID_OK
    for (int t = 0; t < ntraj; ++t) {
        const CidScalars &c = hsc[t];
        qcxms_b200_cid_result_t &r = res[t];
        r = qcxms_b200_cid_result_t{};
        r.collided = c.collided; r.scc_iter_total = c.scc_total; r.stopcid = c.stopcid;
        for (int k = 0; k < 3; ++k) r.direc[k] = direc[3 * t + k];
        collided[t] = c.collided;
        const size_t o = (size_t)t * nuc * 3, o0 = (size_t)t * nuc0 * 3;
        if (c.status == TRJ_FAILED) {   // first single point failed: nothing but the set-up happened
            r.status = 2;
            for (int i = 0; i < 3 * nuc; ++i) { xyz[o + i] = sx[o + i]; velo[o + i] = sv[o + i]; }
            continue;
        }
        for (int i = 0; i < 3 * nuc; ++i) {
            xyz[o + i] = hx[o0 + i]; velo[o + i] = hv[o0 + i]; grad[o + i] = hg[o0 + i];
            axyz[o + i] = c.check_fragmented > 1 ? hst[o + i] : hav[o + i] / c.xyzavg_dump;
        }
        for (int i = 0; i < nuc; ++i) achrg[(size_t)t * nuc + i] = hq[(size_t)t * nuc0 + i];
        r.nstep = c.nstep; r.nfrag = c.nfrag; r.velo_cm = c.new_velo; r.aTlast = c.aTlast; r.ttime = c.ttime; r.epot = c.epot;
        r.status = 1;
    }
    cleanup();
    return 0;
}

extern "C" const char *qcxms_b200_last_error(void) { return g_err.c_str(); }
extern "C" const char *qcxms_b200_version(void) { return "qcxms_b200 0.1 (sm_100a)"; }

// profiling builds only (-DQX_PROFILE_PHASES): read and reset the per-phase cycle counters; returns 0 counters otherwise
extern "C" int qcxms_b200_debug_phase_cycles(double *out16) {
    unsigned long long ph[16] = {0}, sub[16] = {0}, sh[64] = {0};
    CUDA_OK(KS_NT320.egrad_cycles(ph, sub, sh));
    CUDA_OK(KS_NT320.md_cycles(ph, sub, sh));
    CUDA_OK(KS_NT576.egrad_cycles(ph, sub, sh));
    CUDA_OK(KS_NT576.md_cycles(ph, sub, sh));
    CUDA_OK(KS_NT512.egrad_cycles(ph, sub, sh));
    CUDA_OK(KS_NT512.md_cycles(ph, sub, sh));
    for (int i = 0; i < 16; ++i) out16[i] = (double)ph[i];
#ifdef QX_PROFILE_PHASES
    fprintf(stderr, "sub-phase cycles (thread 0):");
    for (int i = 0; i < 16; ++i) if (sub[i]) fprintf(stderr, " %d:%.0f", i, (double)sub[i]);
    fprintf(stderr, "\nsweeps per SCC cycle:");
    for (int i = 0; i < 32; ++i) if (sh[32 + i]) fprintf(stderr, " %d:%.2f", i + 1, (double)sh[i] / (double)sh[32 + i]);
    fprintf(stderr, "\n");
#endif
    return 0;
}
