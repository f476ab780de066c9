// C ABI of qcxms_b200 (include/qcxms_b200.h) and the ensemble kernels.
//
// Kernels are persistent: a grid of (resident CTAs per SM x SM count) CTAs pulls trajectories from an
// atomic work queue; one CTA owns one trajectory at a time and keeps the SCC matrices in shared memory
// and the integral slab in its private, L2-resident global scratch.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/qcxms_b200.h"
#include "qx_host_model.h"
#include "qx_cid.cuh"

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

// ------------------------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(QX_NT, 2) k_egrad_batch(DevModel m, ScratchLayout L, double *scratch, const double *xyz, double kt, int nsys,
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

__global__ void __launch_bounds__(QX_NT) k_fragments(DevModel m, const double *xyz, double rcut, int nsys, int *frag, unsigned char *conn, int *stack) {
    const int nat = m.nat;
    for (int t = blockIdx.x; t < nsys; t += gridDim.x)
        md_fragments(m, xyz + (size_t)t * 3 * nat, rcut, conn + (size_t)blockIdx.x * nat * nat, frag + (size_t)t * nat, stack + (size_t)blockIdx.x * nat);
}

// one egrad + sanity gate for trajectory t; returns Epot (0 on failure, like the reference's checkqc)
__device__ inline double md_egrad(const DevModel &m, Sm &s, double *my, const ScratchLayout &L, const MdConfig &cfg, double etemp,
                                  double *grad_out, double *achrg_out, int *niter_out, double *qstart = nullptr) {
    const int nat = m.nat;
    EgradOut o;
    egrad_cta(m, s, my, L, etemp * QC_KTOAU, o, qstart);
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

// steps added after a fragmentation in the mean-free-path MD (reference src/md.f90:233-235)
__device__ inline int mfp_add_steps(int nuc) { return nuc >= 40 ? (nuc / 10) * 1000 : (nuc > 10 ? (nuc / 10) * 500 : 0); }

// scalar state of the mean-free-path mode of md() (reference src/md.f90:91-116, 209-255), one per trajectory in MdState::mfp_d / mfp_i
struct MfpScalars {
    double old_cm[3], new_velo, new_temp, summass, ekin, pad;
    int cnt, count_average, check_fragmented, max_steps, save_natf[10], ops, pad2;
};
static_assert(sizeof(MfpScalars) == 8 * sizeof(double) + 16 * sizeof(int), "MfpScalars layout");
enum { MFP_ZERO_BEFORE = 1, MFP_ACCUM = 2, MFP_ZERO_AFTER = 4, MFP_FINAL = 8 };

__device__ inline void mfp_load(const MdState &st, int t, MfpScalars &q) {
    double *d = (double *)&q;
    int *i = (int *)(d + 8);
    for (int k = 0; k < 8; ++k) d[k] = __ldcg(st.mfp_d + (size_t)t * 8 + k);
    for (int k = 0; k < 16; ++k) i[k] = __ldcg(st.mfp_i + (size_t)t * 16 + k);
}
__device__ inline void mfp_store(const MdState &st, int t, const MfpScalars &q) {
    const double *d = (const double *)&q;
    const int *i = (const int *)(d + 8);
    for (int k = 0; k < 8; ++k) st.mfp_d[(size_t)t * 8 + k] = d[k];
    for (int k = 0; k < 16; ++k) st.mfp_i[(size_t)t * 16 + k] = i[k];
}

// md(): everything before the loop (reference src/md.f90:155-283)
__global__ void __launch_bounds__(QX_NT, 2) k_md_init(DevModel m, ScratchLayout L, double *scratch, MdConfig cfg, MdState st, int ntraj, int *queue) {
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
        if (t >= ntraj) break;
        for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) s.xyz[i] = st.xyz[(size_t)t * 3 * nat + i];
        __syncthreads();
        const double eimp = st.eimp[t];
        const double etemp = cfg.etemp_in < 0.0 ? md_setetemp(cfg, 1, eimp) : cfg.etemp_in;
        int nit = 0;
        double *qw = st.qwarm ? st.qwarm + (size_t)t * (2 * m.ndim + 1) : nullptr;
        if (qw) {   // the first single point of a trajectory has nothing to start from: zero populations == the reference's cold start
            for (int i = threadIdx.x; i < 2 * m.ndim + 1; i += QX_NT) qw[i] = 0.0;
            __syncthreads();
        }
        const double epot = md_egrad(m, s, my, L, cfg, etemp, st.grad + (size_t)t * 3 * nat, st.achrg + (size_t)t * nat, &nit, qw);
        if (threadIdx.x == 0) {
            st.scc_total[t] = nit;
            const double ekin = md_ekinet_seq(nat, st.velo + (size_t)t * 3 * nat, m.mass, 0.0, nullptr);
            const double tadd = st.tadd[t];
            st.ekin[t] = ekin; st.ekinstart[t] = ekin; st.epot[t] = epot; st.etemp[t] = etemp;
            if (cfg.icoll > 0) {   // mean-free-path mode: kinetic energy without the motion of the centre of mass (src/md.f90:246-255, 283)
                MfpScalars q{};
                q.new_velo = st.mfp_d[(size_t)t * 8 + 3];
                cid_center_of_mass(nat, m.mass, st.xyz + (size_t)t * 3 * nat, q.old_cm);
                for (int i = 0; i < nat; ++i) q.summass = q.summass + m.mass[i];
                const double E_kin = 0.5 * q.summass * ((q.new_velo * QC_MSTOAU) * (q.new_velo * QC_MSTOAU));
                const double E_kin_diff = ekin - E_kin;
                q.new_temp = (2 * E_kin_diff) / (3 * QC_KB * nat);
                st.ekin[t] = E_kin_diff;
                q.check_fragmented = 1; q.max_steps = cfg.nmax;
                mfp_store(st, t, q);
            }
            st.Tav[t] = 0; st.Epav[t] = 0; st.Ekav[t] = 0; st.Edum[t] = 0; st.aTlast[t] = 0; st.dtime[t] = 0; st.ttime[t] = 0;
            st.nstep[t] = 0; st.kdump[t] = 50; st.fconst[t] = 0; st.morestep[t] = 0; st.nfrag[t] = 1;
            st.fragstate[t] = 0; st.mdok[t] = 0;
            st.nadd[t] = (int)((tadd + cfg.tstep) / cfg.tstep - 1.0);
            st.fadd[t] = cfg.tstep / (tadd + cfg.tstep);
            st.status[t] = epot == 0.0 ? TRJ_FAILED : TRJ_RUNNING;
        }
        for (int i = threadIdx.x; i < nat; i += QX_NT) { st.avchrg[(size_t)t * nat + i] = 0.0; st.list[(size_t)t * nat + i] = 1; }
        for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) st.avxyz[(size_t)t * 3 * nat + i] = 0.0;
        if (cfg.icoll > 0)
            for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) { st.avxyz2[(size_t)t * 3 * nat + i] = 0.0; st.store[(size_t)t * 3 * nat + i] = 0.0; }
    }
}

// up to `chunk` MD steps (reference src/md.f90:285-682) for every running trajectory
// Work items are (sub-chunk r, trajectory t), r-major, so that the last partial wave of CTAs costs a few steps and
// not a whole chunk.  progress[t] counts the finished sub-chunks of trajectory t in this launch: item (r, t) waits
// until (r-1, t) -- possibly still running on another resident CTA -- is done.
// MFP = true: the mean-free-path md() of a CID run (cfg.icoll >= 1; reference global method == 3): no IEE heating, kinetic energy
// without the centre-of-mass motion, averaged fragment structures, tmax as the only regular exit (src/md.f90:246-255, 466-621, 672).
template <bool MFP>
__global__ void __launch_bounds__(QX_NT, 2) k_md_chunk(DevModel m, ScratchLayout L, double *scratch, MdConfig cfg, MdState st, int ntraj, int chunk,
                                                    int nsub, int step_limit, int *queue, int *progress, unsigned long long *steps_done) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_next, s_flag;
    __shared__ MfpScalars s_q;
    Sm s;
    double *my = scratch + (size_t)blockIdx.x * L.total;
    carve(m, smem, s, my + L.matA);
    const int nat = m.nat;
    const double fstoau = QC_FSTOAU, kB = QC_KB;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_next = atomicAdd(queue, 1);
        __syncthreads();
        const int item = s_next;
        if (item >= ntraj * nsub) break;
        const int sub = item / ntraj, t = item - sub * ntraj;
        if (threadIdx.x == 0) {
            while (atomicAdd(&progress[t], 0) < sub) __nanosleep(200);
            __threadfence();
        }
        __syncthreads();
        if (__ldcg(st.status + t) != TRJ_RUNNING) {
            if (threadIdx.x == 0) { __threadfence(); atomicAdd(&progress[t], 1); }
            continue;
        }
        // per-trajectory arrays live in shared memory for the duration of the work item (read with ld.cg: the previous
        // sub-chunk of this trajectory may have run on another SM)
        double *velo = smem + smem_doubles(m.nat, m.nsh, m.nao, m.ld, m.rows8, m.mat_in_global, m.ntype) + 8, *grad = velo + 3 * nat, *avxyz = grad + 3 * nat, *achrg = avxyz + 3 * nat,
               *avchrg = achrg + nat;
        double *gxyz = st.xyz + (size_t)t * 3 * nat, *gvelo = st.velo + (size_t)t * 3 * nat, *ggrad = st.grad + (size_t)t * 3 * nat;
        double *gachrg = st.achrg + (size_t)t * nat, *gavchrg = st.avchrg + (size_t)t * nat, *gavxyz = st.avxyz + (size_t)t * 3 * nat;
        const double *velof = st.velof + (size_t)t * nat;
        int *list = st.list + (size_t)t * nat;
        for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) { s.xyz[i] = __ldcg(gxyz + i); velo[i] = __ldcg(gvelo + i); grad[i] = __ldcg(ggrad + i); avxyz[i] = __ldcg(gavxyz + i); }
        for (int i = threadIdx.x; i < nat; i += QX_NT) { achrg[i] = __ldcg(gachrg + i); avchrg[i] = __ldcg(gavchrg + i); }
        int scc_add = 0;
        // scalar state, kept redundantly in every thread
        int nstep = __ldcg(st.nstep + t), kdump = __ldcg(st.kdump + t), fconst = __ldcg(st.fconst + t), morestep = __ldcg(st.morestep + t), nfrag = __ldcg(st.nfrag + t);
        int fragstate = 0, mdok = 0, status = TRJ_RUNNING;
        const int nadd = __ldcg(st.nadd + t);
        const double fadd = __ldcg(st.fadd + t), eimp = __ldcg(st.eimp + t), ekinstart = __ldcg(st.ekinstart + t);
        double epot = __ldcg(st.epot + t), ekin = __ldcg(st.ekin + t), etemp = __ldcg(st.etemp + t), Tav = __ldcg(st.Tav + t), Epav = __ldcg(st.Epav + t),
               Ekav = __ldcg(st.Ekav + t), Edum = __ldcg(st.Edum + t);
        double aTlast = __ldcg(st.aTlast + t), dtime = __ldcg(st.dtime + t), ttime = __ldcg(st.ttime + t);
        double *gavxyz2 = nullptr, *gstore = nullptr;
        if (MFP) {
            gavxyz2 = st.avxyz2 + (size_t)t * 3 * nat; gstore = st.store + (size_t)t * 3 * nat;
            if (threadIdx.x == 0) mfp_load(st, t, s_q);
        }
        __syncthreads();
        int done = 0;
        for (int it = 0; it < chunk && status == TRJ_RUNNING; ++it) {
            if (step_limit > 0 && nstep >= step_limit) break;  // pause here: the host asked for a bounded number of steps
            nstep += 1;
            const double T = ekin / (0.5 * 3 * nat * kB);
            Tav += T; Epav += epot; Ekav += ekin;
            double Eav;
            if (nstep > nadd) { Edum += epot + ekin; Eav = Edum / (double)(float)(nstep - nadd); }
            else Eav = epot + ekin;
            const double Eerror = Eav - epot - ekin;
            const bool err1 = epot == 0.0, err2 = fabs(Eerror) > (MFP ? (double)0.2f : (double)0.1f);
            if (err1 || (err2 && cfg.exit_rules)) {
                mdok = ((nfrag > 1 && nfrag <= 4) || cfg.isec > 1) ? 1 : 0;
                status = TRJ_FINISHED;
                break;
            }
            if (kdump > 50 - 1) {
                kdump = 0;
                aTlast = 0.0;
                for (int i = threadIdx.x; i < nat; i += QX_NT) avchrg[i] = 0.0;
                for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) avxyz[i] = 0.0;
            }
            for (int i = threadIdx.x; i < nat; i += QX_NT) avchrg[i] += achrg[i];
            for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) avxyz[i] += s.xyz[i];
            aTlast += MFP ? s_q.new_temp : T;
            // leapfrog (reference md.f90:749-773); kinetic-energy terms summed in the reference order
            for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) {
                const double mass = m.mass[i / 3];
                const double vold = velo[i];
                const double vnew = __dsub_rn(vold, __ddiv_rn(__dmul_rn(cfg.tstep, grad[i]), mass));
                const double vavg = __dmul_rn(0.5, __dadd_rn(vold, vnew));
                const double x = __dadd_rn(s.xyz[i], __dmul_rn(cfg.tstep, vnew));
                velo[i] = vnew;
                s.xyz[i] = x;
                s.vdp[i] = __dmul_rn(0.5, __dmul_rn(__dmul_rn(mass, vavg), vavg));
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                double ke = 0.0;
                for (int i = 0; i < 3 * nat; ++i) ke = __dadd_rn(ke, s.vdp[i]);
                s.red[49] = ke;
            }
            __syncthreads();
            ekin = s.red[49];
            ttime += cfg.tstep / fstoau;
            {
                int nit = 0;
                epot = md_egrad(m, s, my, L, cfg, etemp, grad, achrg, &nit, st.qwarm ? st.qwarm + (size_t)t * (2 * m.ndim + 1) : nullptr);
                scc_add += nit;
            }
            done += 1;
            kdump += 1;
            if (nfrag == 1) morestep = 0;
            if (nfrag > 1 && dtime < 1e-6) dtime = ttime / 1000.0;
            if (!MFP) {
                // IEE heating while the ion is intact
                if (nstep <= nadd && nfrag == 1) {
                    if (!md_impactscale(m, velo, velof, eimp, fadd * nstep, ekinstart, &s_flag)) { status = TRJ_FAILED; break; }
                }
                if (cfg.etemp_in < 0.0) {
                    const double dum = eimp - eimp * (double)(float)nstep / (double)(float)nadd;
                    etemp = md_setetemp(cfg, nfrag, dum);
                }
            }
            md_fragments(m, s.xyz, 3.0, (unsigned char *)(my + L.taskout), list, (int *)(my + L.taskout) + (nat * nat + 3) / 4 + 4);
            if (threadIdx.x == 0) s_flag = md_nfrag(m, list);
            __syncthreads();
            nfrag = s_flag;
            if (MFP) {
                if (nfrag > 6) { status = TRJ_FINISHED; break; }
                if (threadIdx.x == 0) {
                    MfpScalars &q = s_q;
                    // kinetic energy without the centre-of-mass motion (src/md.f90:466-493)
                    double cm[3];
                    cid_center_of_mass(nat, m.mass, s.xyz, cm);
                    const double d0 = cm[0] - q.old_cm[0], d1 = cm[1] - q.old_cm[1], d2 = cm[2] - q.old_cm[2];
                    const double cm_out = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
                    q.new_velo = (cm_out / cfg.tstep) / QC_MSTOAU;
                    q.old_cm[0] = cm[0]; q.old_cm[1] = cm[1]; q.old_cm[2] = cm[2];
                    const double E_kin = 0.5 * q.summass * ((q.new_velo * QC_MSTOAU) * (q.new_velo * QC_MSTOAU));
                    const double E_kin_diff = ekin - E_kin;
                    q.new_temp = (2.0 * E_kin_diff) / (3.0 * QC_KB * nat);
                    q.ekin = E_kin_diff;
                    // averaged fragment structures (src/md.f90:496-621)
                    int ops = 0;
                    if (nfrag > q.check_fragmented) { q.count_average = 1; q.check_fragmented = nfrag; q.max_steps = nstep + mfp_add_steps(nat); }
                    if (nfrag < q.check_fragmented && q.count_average) { q.cnt = 0; ops |= MFP_ZERO_BEFORE; q.count_average = 0; q.check_fragmented = 1; }
                    q.pad2 = 0;
                    if (q.count_average) {
                        q.cnt += 1;
                        ops |= MFP_ACCUM;
                        q.pad2 = q.cnt;   // divisor of this step's store_avxyz
                        int natf[10];
                        for (int i = 0; i < 10; ++i) natf[i] = 0;
                        for (int i = 0; i < nat; ++i) if (list[i] >= 1 && list[i] <= nfrag && list[i] <= 10) natf[list[i] - 1] += 1;
                        for (int i = 0; i < nfrag && i < 10; ++i) {
                            if (q.cnt == 1) q.save_natf[i] = natf[i];
                            if (natf[i] != q.save_natf[i]) { q.cnt = 0; ops |= MFP_ZERO_AFTER; break; }
                        }
                        if (q.cnt == 50) { ops |= MFP_FINAL; q.cnt = 0; q.count_average = 0; }
                    }
                    q.ops = ops;
                }
                __syncthreads();
                ekin = s_q.ekin;
                const int ops = s_q.ops;
                if (ops) {
                    const double cnt = (double)s_q.pad2;
                    for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) {
                        double a2 = __ldcg(gavxyz2 + i), sv = __ldcg(gstore + i);
                        if (ops & MFP_ZERO_BEFORE) { a2 = 0.0; sv = 0.0; }
                        if (ops & MFP_ACCUM) { a2 = a2 + s.xyz[i]; sv = a2 / cnt; }
                        if (ops & MFP_ZERO_AFTER) { a2 = 0.0; sv = 0.0; }
                        if (ops & MFP_FINAL) a2 = 0.0;
                        gavxyz2[i] = a2; gstore[i] = sv;
                    }
                }
                const int max_steps = s_q.max_steps;
                __syncthreads();   // thread 0 rewrites s_q in the next step
                if (nstep >= max_steps) { fragstate = 1; mdok = 1; status = TRJ_FINISHED; break; }
                continue;
            }
            if (cfg.exit_rules) {
                if (nfrag > 6) { status = TRJ_FINISHED; break; }
                if (nfrag > cfg.nfragexit) { fragstate = 1; mdok = 1; status = TRJ_FINISHED; break; }
                fconst = nfrag >= 2 ? fconst + 1 : 0;
                if (fconst > 1000) { fragstate = 2; mdok = 1; status = TRJ_FINISHED; break; }
                if (nfrag >= cfg.nfragexit) {
                    morestep += 1;
                    if (morestep > 250) { fragstate = 1; mdok = 1; status = TRJ_FINISHED; break; }
                }
            }
            if (nstep >= cfg.nmax) { fragstate = 1; mdok = 1; status = TRJ_FINISHED; break; }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) { gxyz[i] = s.xyz[i]; gvelo[i] = velo[i]; ggrad[i] = grad[i]; gavxyz[i] = avxyz[i]; }
        for (int i = threadIdx.x; i < nat; i += QX_NT) { gachrg[i] = achrg[i]; gavchrg[i] = avchrg[i]; }
        if (threadIdx.x == 0) {
            st.scc_total[t] = __ldcg(st.scc_total + t) + scc_add;
            st.nstep[t] = nstep; st.kdump[t] = kdump; st.fconst[t] = fconst; st.morestep[t] = morestep; st.nfrag[t] = nfrag;
            st.epot[t] = epot; st.ekin[t] = ekin; st.etemp[t] = etemp; st.Tav[t] = Tav; st.Epav[t] = Epav; st.Ekav[t] = Ekav; st.Edum[t] = Edum;
            st.aTlast[t] = aTlast; st.dtime[t] = dtime; st.ttime[t] = ttime;
            if (status != TRJ_RUNNING) { st.status[t] = status; st.fragstate[t] = fragstate; st.mdok[t] = mdok; }
            if (MFP) mfp_store(st, t, s_q);
            atomicAdd(steps_done, (unsigned long long)done);
        }
        __syncthreads();   // every thread's global writes of this sub-chunk are issued ...
        if (threadIdx.x == 0) { __threadfence(); atomicAdd(&progress[t], 1); }   // ... and published before the hand-over
    }
}

// ------------------------------------------------------------------------------------ CID (reference src/cid.f90)
// set-up of one collision + the two single points before the loop (iniqm's is only checked, so one evaluation serves both)
__global__ void __launch_bounds__(QX_NT, 2) k_cid_init(DevModel m, ScratchLayout L, double *scratch, MdConfig cfg, CidConfig cc, CidState st, int ntraj,
                                                    int nuc, int icoll, int *queue) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_next;
    Sm s;
    double *my = scratch + (size_t)blockIdx.x * L.total;
    carve(m, smem, s, my + L.matA);
    const int nuc0 = m.nat;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_next = atomicAdd(queue, 1);
        __syncthreads();
        const int t = s_next;
        if (t >= ntraj) break;
        CidScalars *sc = st.sc + t;
        double *xyz0 = st.xyz0 + (size_t)t * 3 * nuc0, *velo0 = st.velo0 + (size_t)t * 3 * nuc0;
        if (threadIdx.x == 0) {
            double tinit, summass, old_cm[3];
            cid_setup_thread0(m, cc, nuc, icoll, st.xyz + (size_t)t * 3 * nuc, st.velo + (size_t)t * 3 * nuc, st.rnd + (size_t)t * 9,
                              st.velo_cm_in ? st.velo_cm_in[t] : 0.0, st.direc + (size_t)t * 3, xyz0, velo0, old_cm, &tinit, &summass);
            CidScalars z{};
            z.total_steps = cc.ntot; z.check_fragmented = 1; z.nfrag = 1; z.collided = sc->collided;
            z.Tinit = tinit; z.summass = summass;
            for (int k = 0; k < 3; ++k) z.old_cm[k] = old_cm[k];
            *sc = z;
            __threadfence_block();
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * nuc0; i += QX_NT) s.xyz[i] = xyz0[i];
        for (int i = threadIdx.x; i < 3 * nuc; i += QX_NT) { st.avxyz[(size_t)t * 3 * nuc + i] = 0.0; st.avxyz2[(size_t)t * 3 * nuc + i] = 0.0; st.store[(size_t)t * 3 * nuc + i] = 0.0; }
        for (int i = threadIdx.x; i < nuc; i += QX_NT) st.list[(size_t)t * nuc + i] = 1;
        __syncthreads();
        int nit = 0;
        const double epot = md_egrad(m, s, my, L, cfg, cc.etemp, st.grad0 + (size_t)t * 3 * nuc0, st.achrg0 + (size_t)t * nuc0, &nit);
        if (threadIdx.x == 0) {
            sc->scc_total = nit; sc->epot = epot;
            if (epot == 0.0) { sc->stopcid = 1; sc->status = TRJ_FAILED; }
            else {
                sc->status = TRJ_RUNNING;
                // distance gas atom -- centre of mass of the ion as it was handed in (reference src/cid.f90:733-737)
                double cm[3];
                cid_center_of_mass(nuc, m.mass, st.xyz + (size_t)t * 3 * nuc, cm);
                const int ig = nuc0 - 1;
                const double d0 = xyz0[3 * ig] - cm[0], d1 = xyz0[3 * ig + 1] - cm[1], d2 = xyz0[3 * ig + 2] - cm[2];
                sc->lowestCOM = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
            }
        }
    }
}

// up to `chunk` steps of the collision loop (reference src/cid.f90:739-1052) for every running trajectory
__global__ void __launch_bounds__(QX_NT, 2) k_cid_chunk(DevModel m, ScratchLayout L, double *scratch, MdConfig cfg, CidConfig cc, CidState st, int ntraj,
                                                     int nuc, int chunk, int *queue) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_next, s_ops, s_cnt, s_stop;
    __shared__ CidScalars sc;
    Sm s;
    double *my = scratch + (size_t)blockIdx.x * L.total;
    carve(m, smem, s, my + L.matA);
    const int nuc0 = m.nat;
    const double autofs = 1.0 / QC_FSTOAU;
    int add_steps = 0;
    if (nuc > 10) add_steps = (nuc / 10) * 500;
    if (nuc >= 40) add_steps = (nuc / 10) * 1000;
    enum { OP_RESET_AV = 1, OP_ZERO_BEFORE = 2, OP_ACCUM = 4, OP_ZERO_AFTER = 8, OP_FINAL = 16 };
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_next = atomicAdd(queue, 1);
        __syncthreads();
        const int t = s_next;
        if (t >= ntraj) break;
        if (st.sc[t].status != TRJ_RUNNING) continue;
        double *velo0 = smem + smem_doubles(m.nat, m.nsh, m.nao, m.ld, m.rows8, m.mat_in_global, m.ntype) + 8, *grad0 = velo0 + 3 * nuc0, *achrg0 = grad0 + 3 * nuc0;
        double *gxyz0 = st.xyz0 + (size_t)t * 3 * nuc0, *gvelo0 = st.velo0 + (size_t)t * 3 * nuc0, *ggrad0 = st.grad0 + (size_t)t * 3 * nuc0,
               *gachrg0 = st.achrg0 + (size_t)t * nuc0;
        double *avxyz = st.avxyz + (size_t)t * 3 * nuc, *avxyz2 = st.avxyz2 + (size_t)t * 3 * nuc, *store = st.store + (size_t)t * 3 * nuc;
        int *list = st.list + (size_t)t * nuc;
        for (int i = threadIdx.x; i < 3 * nuc0; i += QX_NT) { s.xyz[i] = gxyz0[i]; velo0[i] = gvelo0[i]; grad0[i] = ggrad0[i]; }
        for (int i = threadIdx.x; i < nuc0; i += QX_NT) achrg0[i] = gachrg0[i];
        if (threadIdx.x == 0) { sc = st.sc[t]; s_stop = 0; }
        __syncthreads();
        for (int it = 0; it < chunk; ++it) {
            if (threadIdx.x == 0) {
                sc.nstep += 1;
                s_ops = 0;
                if (sc.xyzavg_dump == 50) { sc.xyzavg_dump = 0; s_ops |= OP_RESET_AV; }
                sc.ttime = sc.ttime + cc.tstep * autofs;
                sc.distance_dump += 1; sc.xyzavg_dump += 1;
            }
            __syncthreads();
            if (s_ops & OP_RESET_AV) for (int i = threadIdx.x; i < 3 * nuc; i += QX_NT) avxyz[i] = 0.0;
            for (int i = threadIdx.x; i < 3 * nuc0; i += QX_NT) {   // leapfrog on ion + gas atom
                const double mass = m.mass[i / 3];
                const double vnew = __dsub_rn(velo0[i], __ddiv_rn(__dmul_rn(cc.tstep, grad0[i]), mass));
                velo0[i] = vnew;
                s.xyz[i] = __dadd_rn(s.xyz[i], __dmul_rn(cc.tstep, vnew));
            }
            __syncthreads();
            int nit = 0;
            const double epot = md_egrad(m, s, my, L, cfg, cc.etemp, grad0, achrg0, &nit);
            if (threadIdx.x == 0) { sc.scc_total += nit; sc.epot = epot; }
            if (epot == 0.0) {
                if (threadIdx.x == 0) { sc.stopcid = 1; sc.status = TRJ_FINISHED; }
                break;
            }
            md_fragments(m, s.xyz, 3.0, (unsigned char *)(my + L.taskout), list, (int *)(my + L.taskout) + (nuc * nuc + 3) / 4 + 4, nuc);
            if (threadIdx.x == 0) {
                double cm[3], T;
                cid_center_of_mass(nuc, m.mass, s.xyz, cm);
                const double dc0 = cm[0] - sc.old_cm[0], dc1 = cm[1] - sc.old_cm[1], dc2 = cm[2] - sc.old_cm[2];
                const double cm_out = sqrt(dc0 * dc0 + dc1 * dc1 + dc2 * dc2);
                sc.old_cm[0] = cm[0]; sc.old_cm[1] = cm[1]; sc.old_cm[2] = cm[2];
                sc.new_velo = sc.nstep != 1 ? (cm_out / cc.tstep) / QC_MSTOAU : 0.0;
                const double Ekin = cid_ekinet(nuc, velo0, m.mass, &T);
                const double E_velo = 0.5 * sc.summass * ((sc.new_velo * QC_MSTOAU) * (sc.new_velo * QC_MSTOAU));
                double new_temp = (2 * (Ekin - E_velo)) / (3 * QC_KB * nuc);
                if (sc.nstep == 1) new_temp = sc.Tinit;
                sc.Tav = sc.Tav + new_temp; sc.m = sc.m + 1;
                const double avgT = sc.Tav / sc.m;
                const int nfrag = md_nfrag(m, list, nuc);
                sc.nfrag = nfrag;
                if (nfrag > sc.check_fragmented) { sc.count_average = 1; sc.check_fragmented = nfrag; }
                if (nfrag < sc.check_fragmented && sc.count_average) { sc.cnt = 0; s_ops |= OP_ZERO_BEFORE; sc.count_average = 0; sc.check_fragmented = 1; }
                if (sc.count_average) {
                    sc.cnt += 1;
                    s_ops |= OP_ACCUM;
                    s_cnt = sc.cnt;
                    int natf[10];
                    for (int i = 0; i < 10; ++i) natf[i] = 0;
                    for (int i = 0; i < nuc; ++i) if (list[i] >= 1 && list[i] <= nfrag && list[i] <= 10) natf[list[i] - 1] += 1;
                    for (int i = 0; i < nfrag && i < 10; ++i) {
                        if (sc.cnt == 1) sc.save_natf[i] = natf[i];
                        if (natf[i] != sc.save_natf[i]) { sc.cnt = 0; s_ops |= OP_ZERO_AFTER; break; }
                    }
                    if (sc.cnt == 50) { s_ops |= OP_FINAL; sc.cnt = 0; sc.count_average = 0; }
                }
                sc.aTlast = avgT;
                if (sc.distance_dump == 10) {
                    sc.distance_dump = 0;
                    const int ig = nuc0 - 1;
                    const double d0 = s.xyz[3 * ig] - cm[0], d1 = s.xyz[3 * ig + 1] - cm[1], d2 = s.xyz[3 * ig + 2] - cm[2];
                    const double new_dist = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
                    if (new_dist < sc.lowestCOM) sc.lowestCOM = new_dist;
                    if (sc.lowestCOM < new_dist) sc.step_counter += 1; else sc.step_counter = 0;
                    if (sc.step_counter == 5) {
                        sc.total_steps = sc.nstep + (int)llround(800.0 * (2 * cc.tstep * autofs));
                        sc.collided = 1; sc.Tav = 0; sc.m = 0;
                    }
                }
                if (nfrag > 1 && sc.collided && !sc.fragmented) { sc.total_steps = sc.nstep + add_steps; sc.fragmented = 1; }
                if (sc.nstep >= sc.total_steps) { sc.stopcid = 0; sc.status = TRJ_FINISHED; s_stop = 1; }
            }
            __syncthreads();
            const int ops = s_ops, stop = s_stop;
            const double cnt = (double)s_cnt;
            for (int i = threadIdx.x; i < 3 * nuc; i += QX_NT) {
                avxyz[i] += s.xyz[i];
                if (ops & OP_ZERO_BEFORE) { avxyz2[i] = 0.0; store[i] = 0.0; }
                if (ops & OP_ACCUM) { const double v = avxyz2[i] + s.xyz[i]; avxyz2[i] = v; store[i] = v / cnt; }
                if (ops & OP_ZERO_AFTER) { avxyz2[i] = 0.0; store[i] = 0.0; }
                if (ops & OP_FINAL) avxyz2[i] = 0.0;
            }
            __syncthreads();   // thread 0 rewrites the flags at the top of the next step
            if (stop) break;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * nuc0; i += QX_NT) { gxyz0[i] = s.xyz[i]; gvelo0[i] = velo0[i]; ggrad0[i] = grad0[i]; }
        for (int i = threadIdx.x; i < nuc0; i += QX_NT) gachrg0[i] = achrg0[i];
        if (threadIdx.x == 0) st.sc[t] = sc;
    }
}

__global__ void k_histogram(DevModel m, MdState st, int ntraj, int nbins, double *bins) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntraj || st.status[t] != TRJ_FINISHED || !st.mdok[t]) return;
    const int *list = st.list + (size_t)t * m.nat;
    for (int f = 1; f <= 10; ++f) {
        double mass = 0.0;
        for (int i = 0; i < m.nat; ++i)
            if (list[i] == f) mass += m.mass[i];
        if (mass > 0.0) {
            int bin = (int)llrint(mass * QC_AUTOAMU);
            if (bin >= 0 && bin < nbins) atomicAdd(&bins[bin], 1.0);
        }
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
};

// dynamic shared memory limit of a kernel = what the device allows next to the kernel's static shared memory
template <class K>
static cudaError_t allow_max_dynamic_smem(K kernel, const cudaDeviceProp &prop) {
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, kernel);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin - (int)a.sharedSizeBytes);
}

static int context_init(Context &c, int nat, const int32_t *num, const double *mass, int charge, int multiplicity, int device, int nwork) {
    CUDA_OK(cudaSetDevice(device));
    std::string why = build_host_model(c.hm, nat, num, mass, charge, multiplicity);
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
        // the rest of the shared memory is the block buffer of the blocked Jacobi (2 * jblock rows of the SCC matrix)
        const size_t row = (size_t)c.hm.ld * sizeof(double);
        const size_t avail = (size_t)prop.sharedMemPerBlockOptin - c.smem - 16 - 2048;   // 2 KB: the kernels' static shared memory
        int jb = (int)(avail / (2 * row));
        if (jb > 36) jb = 36;   // 2 x 18 sixteen-lane groups: the pairs of a round of two blocks fill two passes
        if (jb >= 8 && c.hm.nao <= 256) {
            const int nb = (c.hm.nao + jb - 1) / jb;
            jb = (c.hm.nao + nb - 1) / nb;          // balanced blocks
            c.hm.dev.jblock = jb;
            c.smem += 16 + 2 * (size_t)jb * row;
        }
    }
    c.L = make_layout(c.hm);
    // the device maximum, not this composition's size: host threads set up different compositions concurrently
    CUDA_OK(allow_max_dynamic_smem(k_egrad_batch, prop));
    CUDA_OK(allow_max_dynamic_smem(k_md_init, prop));
    CUDA_OK(allow_max_dynamic_smem(k_md_chunk<false>, prop));
    CUDA_OK(allow_max_dynamic_smem(k_md_chunk<true>, prop));
    int per_sm = 0;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_md_chunk<false>, QX_NT, c.smem));
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

static int egrad_batch_impl(int nsys, int nat, const int32_t *num, const double *xyz, int charge, int multiplicity, int method_id,
                            double etemp, double *qat, double *energy, double *gradient, int32_t *stat, int32_t *niter,
                            std::vector<double> *spec, int *nao_out) {
    if (nsys < 1 || nat < 1 || !num || !xyz || !qat || !energy || !gradient || !stat) return fail(QCXMS_B200_ERR_ARG, "null or empty argument");
    if (method_id != QCXMS_B200_GFN2) {
        // reference: unknown method -> stat = 5, outputs untouched (src/tblite.f90:114-120).  GFN1/IPEA1 are not built yet.
        for (int i = 0; i < nsys; ++i) stat[i] = QCXMS_B200_STAT_UNKNOWN_METHOD;
        return 0;
    }
    // cache the model of the last composition: the reference calls this entry point once per MD step
    static std::mutex mtx;
    static Context ctx;
    std::lock_guard<std::mutex> lock(mtx);
    std::vector<int32_t> key(num, num + nat);
    key.push_back(charge); key.push_back(multiplicity); key.push_back(nsys > 1 ? 1 << 20 : 1);
    int dev = 0;
    cudaGetDevice(&dev);
    if (key != ctx.key || ctx.device != dev) {
        context_free(ctx);
        ctx.key.clear();
        int rc = context_init(ctx, nat, num, nullptr, charge, multiplicity, dev, nsys > 1 ? 0 : 1);
        if (rc) { context_free(ctx); return rc; }
        ctx.key = key;
    }
    const size_t n3 = (size_t)nsys * nat * 3;
    double *d_xyz = nullptr, *d_e = nullptr, *d_g = nullptr, *d_q = nullptr;
    int *d_stat = nullptr, *d_nit = nullptr;
    CUDA_OK(cudaMalloc(&d_xyz, n3 * sizeof(double)));
    CUDA_OK(cudaMalloc(&d_g, n3 * sizeof(double)));
    CUDA_OK(cudaMalloc(&d_q, (size_t)nsys * nat * sizeof(double)));
    CUDA_OK(cudaMalloc(&d_e, nsys * sizeof(double)));
    CUDA_OK(cudaMalloc(&d_stat, nsys * sizeof(int)));
    CUDA_OK(cudaMalloc(&d_nit, nsys * sizeof(int)));
    CUDA_OK(cudaMemcpy(d_xyz, xyz, n3 * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemset(ctx.d_queue, 0, sizeof(int)));
    int grid = ctx.ncta < nsys ? ctx.ncta : nsys;
    double *d_spec = nullptr;
    const size_t nspec = (size_t)nsys * (2 * ctx.hm.nao + (size_t)ctx.hm.nao * nat + 1);
    if (spec) CUDA_OK(cudaMalloc(&d_spec, nspec * sizeof(double)));
    if (nao_out) *nao_out = ctx.hm.nao;
    k_egrad_batch<<<grid, QX_NT, ctx.smem>>>(ctx.hm.dev, ctx.L, ctx.d_scratch, d_xyz, etemp * QC_KTOAU, nsys, ctx.d_queue, d_e, d_g, d_q, d_stat, d_nit, d_spec);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaDeviceSynchronize());
    if (spec) {
        spec->resize(nspec);
        CUDA_OK(cudaMemcpy(spec->data(), d_spec, nspec * sizeof(double), cudaMemcpyDeviceToHost));
        cudaFree(d_spec);
    }
    CUDA_OK(cudaMemcpy(energy, d_e, nsys * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(gradient, d_g, n3 * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(qat, d_q, (size_t)nsys * nat * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(stat, d_stat, nsys * sizeof(int), cudaMemcpyDeviceToHost));
    if (niter) CUDA_OK(cudaMemcpy(niter, d_nit, nsys * sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d_xyz); cudaFree(d_g); cudaFree(d_q); cudaFree(d_e); cudaFree(d_stat); cudaFree(d_nit);
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
    if (method_id != QCXMS_B200_GFN2) return fail(QCXMS_B200_ERR_UNSUPPORTED, "only GFN2-xTB (method id 2) is implemented");
    int n = 0;
    for (int i = 0; i < nat; ++i) {
        if (num[i] < 1 || num[i] > GFN2_MAXZ) return fail(QCXMS_B200_ERR_UNSUPPORTED, "GFN2-xTB parameters are available for H-Ar");
        const gfn2_elem_t &e = GFN2_ELEM[num[i]];
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
    double *d_rad, *d_xyz;
    int *d_frag, *d_stack;
    unsigned char *d_conn;
    int grid = nsys < 1024 ? nsys : 1024;
    CUDA_OK(cudaMalloc(&d_rad, nat * sizeof(double)));
    CUDA_OK(cudaMalloc(&d_xyz, (size_t)nsys * nat * 3 * sizeof(double)));
    CUDA_OK(cudaMalloc(&d_frag, (size_t)nsys * nat * sizeof(int)));
    CUDA_OK(cudaMalloc(&d_stack, (size_t)grid * nat * sizeof(int)));
    CUDA_OK(cudaMalloc(&d_conn, (size_t)grid * nat * nat));
    CUDA_OK(cudaMemcpy(d_rad, qcrad.data(), nat * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(d_xyz, xyz, (size_t)nsys * nat * 3 * sizeof(double), cudaMemcpyHostToDevice));
    m.at_qcrad = d_rad;
    k_fragments<<<grid, QX_NT>>>(m, d_xyz, rcut, nsys, d_frag, d_conn, d_stack);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpy(frag, d_frag, (size_t)nsys * nat * sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d_rad); cudaFree(d_xyz); cudaFree(d_frag); cudaFree(d_stack); cudaFree(d_conn);
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
    if (cfg->method_id != QCXMS_B200_GFN2) return fail(QCXMS_B200_ERR_UNSUPPORTED, "only GFN2-xTB (method id 2) is implemented");
    auto *h = new qcxms_b200_ensemble();
    // multiplicity as the reference's getspin() would pass it (src/utility.f90:449-464); tblite discards it anyway
    int zsum = 0;
    for (int i = 0; i < nat; ++i) zsum += num[i];
    int j = zsum - std::abs(cfg->mchrg);
    int mult = j < 1 ? -1 : 1 + j % 2;
    int rc = context_init(h->ctx, nat, num, mass, cfg->mchrg, mult, device, ntraj);
    if (rc) { context_free(h->ctx); delete h; return rc; }
    h->ntraj = ntraj;
    h->cfg.mchrg = cfg->mchrg; h->cfg.nfragexit = cfg->nfragexit; h->cfg.exit_rules = cfg->exit_rules; h->cfg.nmax = cfg->nmax; h->cfg.isec = cfg->isec;
    h->cfg.tstep = cfg->tstep; h->cfg.etemp_in = cfg->etemp_in; h->cfg.ieetemp = cfg->ieetemp; h->cfg.ax = cfg->ax;
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
        cudaError_t e = ens_alloc(h, &h->st.qwarm, (size_t)h->ntraj * (2 * h->ctx.hm.ndim + 1));
        if (e != cudaSuccess) return fail(QCXMS_B200_ERR_CUDA, std::string("warm-start buffer: ") + cudaGetErrorString(e));
        h->initialised = false;   // the populations are seeded by the initial single point of md()
    } else if (!on)
        h->st.qwarm = nullptr;    // (buffer stays owned by the handle)
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
    h->initialised = false;
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
    if (h->cfg.icoll <= 0 || !axyz) return cudaSuccess;
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
        k_md_init<<<grid, QX_NT, c.smem, h->stream>>>(c.hm.dev, c.L, c.d_scratch, h->cfg, h->st, h->ntraj, c.d_queue);
        CUDA_OK(cudaGetLastError());
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
    const int total = max_steps > 0 ? max_steps : (h->cfg.icoll > 0 ? h->cfg.nmax + 64 * (h->ctx.hm.nat / 10 + 1) * 1000 : h->cfg.nmax);
    const int chunk = 64, sub_steps = 8;
    std::vector<int> status(h->ntraj);
    for (int done = 0; done < total; done += chunk) {
        const int this_chunk = total - done < chunk ? total - done : chunk;
        const int nsub = (this_chunk + sub_steps - 1) / sub_steps;
        CUDA_OK(cudaMemsetAsync(c.d_queue, 0, sizeof(int), h->stream));
        CUDA_OK(cudaMemsetAsync(h->d_progress, 0, h->ntraj * sizeof(int), h->stream));
        // the last sub-chunk may be shorter: the kernel bounds every sub-chunk by the launch's step limit as well
        if (h->cfg.icoll > 0)
            k_md_chunk<true><<<grid, QX_NT, c.smem, h->stream>>>(c.hm.dev, c.L, c.d_scratch, h->cfg, h->st, h->ntraj, sub_steps, nsub,
                                                                 max_steps > 0 ? limit : 0, c.d_queue, h->d_progress, h->d_steps);
        else
            k_md_chunk<false><<<grid, QX_NT, c.smem, h->stream>>>(c.hm.dev, c.L, c.d_scratch, h->cfg, h->st, h->ntraj, sub_steps, nsub,
                                                                  max_steps > 0 ? limit : 0, c.d_queue, h->d_progress, h->d_steps);
        CUDA_OK(cudaGetLastError());
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

// ------------------------------------------------------------------------------------ CID host entry
extern "C" int qcxms_b200_cid_batch(const qcxms_b200_cid_config_t *cfg, int ntraj, int nuc, const int32_t *num, const double *mass, int icoll,
                                    double *xyz, double *velo, const double *rnd, const double *velo_cm, double *direc, int32_t *collided,
                                    double *grad, double *achrg, double *axyz, int32_t *list, qcxms_b200_cid_result_t *res, int device) {
    if (!cfg || ntraj < 1 || nuc < 1 || !num || !mass || !xyz || !velo || !rnd || !direc || !collided || !grad || !achrg || !axyz || !list || !res)
        return fail(QCXMS_B200_ERR_ARG, "null or empty argument");
    if (icoll < 1 || (icoll > 1 && !velo_cm)) return fail(QCXMS_B200_ERR_ARG, "icoll >= 1; later collisions need velo_cm");
    if (cfg->method_id != QCXMS_B200_GFN2) return fail(QCXMS_B200_ERR_UNSUPPORTED, "only GFN2-xTB (method id 2) is implemented");
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
    int rc = context_init(ctx, nuc0, num0.data(), mass0.data(), cfg->mchrg, mult, device, ntraj);
    if (rc) { context_free(ctx); return rc; }
    {   // the device maximum, not this composition's size: host threads set up different compositions concurrently
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, device));
        CUDA_OK(allow_max_dynamic_smem(k_cid_init, prop));
        CUDA_OK(allow_max_dynamic_smem(k_cid_chunk, prop));
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
    k_cid_init<<<grid, QX_NT, ctx.smem, strm>>>(ctx.hm.dev, ctx.L, ctx.d_scratch, mc, cc, st, ntraj, nuc, icoll, ctx.d_queue);
    CID_OK(cudaGetLastError());
    const int chunk = 32;
    for (int done = 0; done < cc.ntot + chunk; done += chunk) {
        CID_OK(cudaMemsetAsync(ctx.d_queue, 0, sizeof(int), strm));
        k_cid_chunk<<<grid, QX_NT, ctx.smem, strm>>>(ctx.hm.dev, ctx.L, ctx.d_scratch, mc, cc, st, ntraj, nuc, chunk, ctx.d_queue);
        CID_OK(cudaGetLastError());
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
#ifdef QX_PROFILE_PHASES
    unsigned long long h[16], z[16] = {0};
    CUDA_OK(cudaMemcpyFromSymbol(h, g_phase_cycles, sizeof(h)));
    CUDA_OK(cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)));
    for (int i = 0; i < 16; ++i) out16[i] = (double)h[i];
    {
        unsigned long long sub[16], sz[16] = {0};
        CUDA_OK(cudaMemcpyFromSymbol(sub, g_sub_cycles, sizeof(sub)));
        CUDA_OK(cudaMemcpyToSymbol(g_sub_cycles, sz, sizeof(sz)));
        fprintf(stderr, "sub-phase cycles (thread 0):");
        for (int i = 0; i < 16; ++i) if (sub[i]) fprintf(stderr, " %d:%.0f", i, (double)sub[i]);
        fprintf(stderr, "\n");
    }
    unsigned long long sh[64], sz[64] = {0};
    CUDA_OK(cudaMemcpyFromSymbol(sh, g_sweep_hist, sizeof(sh)));
    CUDA_OK(cudaMemcpyToSymbol(g_sweep_hist, sz, sizeof(sz)));
    fprintf(stderr, "sweeps per SCC cycle:");
    for (int i = 0; i < 32; ++i) if (sh[32 + i]) fprintf(stderr, " %d:%.2f", i + 1, (double)sh[i] / (double)sh[32 + i]);
    fprintf(stderr, "\n");
#else
    for (int i = 0; i < 16; ++i) out16[i] = 0.0;
#endif
    return 0;
}
