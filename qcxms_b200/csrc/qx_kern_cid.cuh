// Ensemble kernels of cid() (reference src/cid.f90:24-28): collision set-up and the collision loop in chunks.
#pragma once
#include "qx_md_egrad.cuh"

namespace qx {

// ------------------------------------------------------------------------------------ CID (reference src/cid.f90)
// set-up of one collision + the two single points before the loop (iniqm's is only checked, so one evaluation serves both)
static __global__ void __launch_bounds__(QX_NT, QX_MINB) k_cid_init(DevModel m, ScratchLayout L, double *scratch, MdConfig cfg, CidConfig cc, CidState st, int ntraj,
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
static __global__ void __launch_bounds__(QX_NT, QX_MINB) k_cid_chunk(DevModel m, ScratchLayout L, double *scratch, MdConfig cfg, CidConfig cc, CidState st, int ntraj,
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
        double *velo0 = smem + m.extras_off, *grad0 = velo0 + 3 * nuc0, *achrg0 = grad0 + 3 * nuc0;
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

}  // namespace qx
