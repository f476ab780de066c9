// Ensemble kernels of md() (reference src/md.f90:34-39): set-up (k_md_init) and the MD loop in chunks (k_md_chunk).
// QX_TU_MD_INIT: the translation unit that instantiates k_md_init (it serves both the EI and the mean-free-path mode).
#pragma once
#include "qx_md_egrad.cuh"

namespace qx {

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

#ifdef QX_TU_MD_INIT
// md(): everything before the loop (reference src/md.f90:155-283)
static __global__ void __launch_bounds__(QX_NT, QX_MINB) k_md_init(DevModel m, ScratchLayout L, double *scratch, MdConfig cfg, MdState st, int ntraj, int *queue) {
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
        const double etemp = (cfg.etemp_in < 0.0 && cfg.it_mode > 0) ? md_setetemp(cfg, 1, eimp) : cfg.etemp_in;
        int nit = 0;
        double *qw = st.qwarm ? st.qwarm + (size_t)t * (2 * m.ndim + 1) : nullptr;
        if (qw) {   // the first single point of a trajectory has nothing to start from: zero populations == the reference's cold start
            for (int i = threadIdx.x; i < 2 * m.ndim + 1; i += QX_NT) qw[i] = 0.0;
            __syncthreads();
        }
        double *es = st.eigseed ? st.eigseed + (size_t)t * ((size_t)QX_OA_NSTORE * m.nao * m.nao + 1) : nullptr;
        if (es && threadIdx.x == 0) es[(size_t)QX_OA_NSTORE * m.nao * m.nao] = 0.0;   // no seeds before the first single point
        __syncthreads();
        const double epot = md_egrad(m, s, my, L, cfg, etemp, st.grad + (size_t)t * 3 * nat, st.achrg + (size_t)t * nat, &nit, qw, es);
        if (threadIdx.x == 0) {
            st.scc_total[t] = nit;
            const double ekin = md_ekinet_seq(nat, st.velo + (size_t)t * 3 * nat, m.mass, 0.0, nullptr);
            const double tadd = st.tadd[t];
            st.ekin[t] = ekin; st.ekinstart[t] = ekin; st.epot[t] = epot; st.etemp[t] = etemp;
            if (cfg.method3) {   // mean-free-path mode: kinetic energy without the motion of the centre of mass (src/md.f90:246-255, 283)
                MfpScalars q{};
                q.new_velo = st.mfp_d[(size_t)t * 8 + 3];
                cid_center_of_mass(nat, m.mass, st.xyz + (size_t)t * 3 * nat, q.old_cm);
                for (int i = 0; i < nat; ++i) q.summass = q.summass + m.mass[i];
                const double E_kin = 0.5 * q.summass * ((q.new_velo * QC_MSTOAU) * (q.new_velo * QC_MSTOAU));
                const double E_kin_diff = ekin - E_kin;
                q.new_temp = (2 * E_kin_diff) / (3 * QC_KB * nat);
                if (cfg.icoll > 0) st.ekin[t] = E_kin_diff;   // (the heating MD before the first collision keeps the full kinetic energy, :254)
                q.check_fragmented = 1; q.max_steps = cfg.nmax;
                mfp_store(st, t, q);
            }
            st.Tav[t] = 0; st.Epav[t] = 0; st.Ekav[t] = 0; st.Edum[t] = 0; st.aTlast[t] = 0; st.dtime[t] = 0; st.ttime[t] = 0;
            st.nstep[t] = 0; st.kdump[t] = 50; st.fconst[t] = 0; st.morestep[t] = 0; st.nfrag[t] = 1;
            st.fragstate[t] = 0; st.mdok[t] = 0;
            st.nadd[t] = cfg.it_mode > 0 ? (int)((tadd + cfg.tstep) / cfg.tstep - 1.0) : 0;   // ground-state runs: nadd = 0 (src/md.f90:270-274)
            st.fadd[t] = cfg.tstep / (tadd + cfg.tstep);
            st.status[t] = epot == 0.0 ? TRJ_FAILED : TRJ_RUNNING;
        }
        for (int i = threadIdx.x; i < nat; i += QX_NT) { st.avchrg[(size_t)t * nat + i] = 0.0; st.list[(size_t)t * nat + i] = 1; }
        for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) st.avxyz[(size_t)t * 3 * nat + i] = 0.0;
        if (cfg.method3)
            for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) { st.avxyz2[(size_t)t * 3 * nat + i] = 0.0; st.store[(size_t)t * 3 * nat + i] = 0.0; }
    }
}
#endif

// up to `chunk` MD steps (reference src/md.f90:285-682) for every running trajectory
// Work items are (sub-chunk r, trajectory t), r-major, so that the last partial wave of CTAs costs a few steps and
// not a whole chunk.  progress[t] counts the finished sub-chunks of trajectory t in this launch: item (r, t) waits
// until (r-1, t) -- possibly still running on another resident CTA -- is done.
// MFP = true: the mean-free-path md() of a CID run (cfg.icoll >= 1; reference global method == 3): no IEE heating, kinetic energy
// without the centre-of-mass motion, averaged fragment structures, tmax as the only regular exit (src/md.f90:246-255, 466-621, 672).
template <bool MFP>
static __global__ void __launch_bounds__(QX_NT, QX_MINB) k_md_chunk(DevModel m, ScratchLayout L, double *scratch, MdConfig cfg, MdState st, int ntraj, int chunk,
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
        double *velo = smem + m.extras_off, *grad = velo + 3 * nat, *avxyz = grad + 3 * nat, *achrg = avxyz + 3 * nat,
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
            if (!MFP && cfg.it_mode <= 0) etemp = cfg.etemp_in;   // src/md.f90:290-297
            Tav += T; Epav += epot; Ekav += ekin;
            double Eav;
            if (nstep > nadd) { Edum += epot + ekin; Eav = Edum / (double)(float)(nstep - nadd); }
            else Eav = epot + ekin;
            const double Eerror = (!MFP && cfg.it_mode < 0) ? 0.0 : Eav - epot - ekin;   // no energy-conservation test while equilibrating (:318)
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
            aTlast += (MFP && cfg.icoll > 0) ? s_q.new_temp : T;
            if (!MFP && cfg.it_mode == 0 && st.gsdump) {   // the record of this step in qcxms.gs (src/md.f90:380-385)
                double *rec = st.gsdump + ((size_t)t * cfg.nmax + (nstep - 1)) * 6 * nat;
                for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) { const int a = i / 3, c = i - 3 * a; rec[6 * a + c] = s.xyz[i]; rec[6 * a + 3 + c] = velo[i]; }
            }
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
                epot = md_egrad(m, s, my, L, cfg, etemp, grad, achrg, &nit, st.qwarm ? st.qwarm + (size_t)t * (2 * m.ndim + 1) : nullptr,
                                st.eigseed ? st.eigseed + (size_t)t * ((size_t)QX_OA_NSTORE * m.nao * m.nao + 1) : nullptr);
                scc_add += nit;
            }
            done += 1;
            kdump += 1;
            if (nfrag == 1) morestep = 0;
            if (nfrag > 1 && dtime < 1e-6) dtime = ttime / 1000.0;
            if (!MFP && cfg.it_mode <= 0) {
                // ground-state runs: rescale towards Tsoll while equilibrating (src/md.f90:402-410); no IEE, no fragment check, tmax is the only exit
                if (cfg.it_mode < 0) {
                    const double dum = 100.0 * fabs(Tav / nstep - cfg.tsoll) / cfg.tsoll;
                    if (dum > 5.0 && nstep > 50) {
                        const double f = sqrt(Tav / nstep / cfg.tsoll);
                        __syncthreads();
                        for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) velo[i] = velo[i] / f;
                        __syncthreads();
                    }
                }
                if (nstep >= cfg.nmax) { fragstate = 1; mdok = 1; status = TRJ_FINISHED; break; }
                continue;
            }
            if (!MFP) {
                // IEE heating while the ion is intact
                if (nstep <= nadd && nfrag == 1) {
                    if (!md_impactscale(m, velo, velof, eimp, fadd * nstep, ekinstart, &s_flag)) { status = TRJ_FAILED; break; }
                }
                {   // unconditional in the reference (src/md.f90:443-445): a user ETEMP only serves the first single point
                    const double dum = eimp - eimp * (double)(float)nstep / (double)(float)nadd;
                    etemp = md_setetemp(cfg, nfrag, dum);
                }
            }
            if (MFP && cfg.icoll == 0 && cfg.starting_md && nfrag == 1 && nstep <= nadd) {   // Berendsen thermostat of the heating MD (src/md.f90:428-434)
                const double sca = sqrt(1.0 + ((cfg.tstep / fstoau) / 150) * (cfg.tsoll / T - 1.0));
                for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) velo[i] = sca * velo[i];
                __syncthreads();
            }
            md_fragments(m, s.xyz, 3.0, (unsigned char *)(my + L.taskout), list, (int *)(my + L.taskout) + (nat * nat + 3) / 4 + 4);
            if (threadIdx.x == 0) s_flag = md_nfrag(m, list);
            __syncthreads();
            nfrag = s_flag;
            if (MFP) {
                if (nfrag > 6) { status = TRJ_FINISHED; break; }
                if (threadIdx.x == 0) {
                    MfpScalars &q = s_q;
                    q.ekin = ekin;
                    if (cfg.icoll > 0) {
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
                    }
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

}  // namespace qx
