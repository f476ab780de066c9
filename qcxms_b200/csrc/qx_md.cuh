// MD side of the ensemble kernel: leapfrog, kinetic energy, IEE heating, fragment connectivity,
// the sanity gate around egrad and the EI exit rules -- one CTA per trajectory.
//   leapfrog            reference src/md.f90:749-773
//   ekinet              reference src/mdinit.f90:57-76
//   impactscale         reference src/impact.f90:12-55   (single-precision literals 0.0002 / 0.001 kept)
//   fragment_structure  reference src/fragments.f90:93-182 (bit-exact integer result)
//   fragmass (nfrag)    reference src/fragments.f90:10-84
//   checkqc / gnorm     reference src/iniqm.f90:684-732  (incl. the g_y-counted-twice quirk)
//   setetemp            reference src/utility.f90:69-86
//   md() state machine  reference src/md.f90:285-700
// Arithmetic that decides integers (distance test, grid search) uses explicit round-to-nearest
// intrinsics so that no FMA contraction can change a comparison against the reference.
#pragma once
#include "params/constants.h"
#include "qx_scc.cuh"

namespace qx {

struct MdConfig {
    int mchrg, nfragexit, exit_rules, nmax, isec;
    double tstep, etemp_in, ieetemp, ax;
    int it_mode; // reference argument `it` of md(): > 0 production run (default 1); 0: ground-state sampling (NVE, every step dumped, qcxms.gs);
                 // -1: ground-state equilibration (velocities rescaled towards tsoll).  src/md.f90:128-131, 290-297, 380-385, 402-410
    double tsoll; // target temperature of the equilibration (reference Tsoll = Tinit)
    int method3;     // 1: md() inside a CID run (reference global method == 3): mean-free-path MD (icoll >= 1) or the heating MD before the
                     // first collision (icoll == 0 with starting_md: Berendsen scaling towards tsoll, src/md.f90:428-434, main.F90:1357-1362)
    int starting_md;
    int icoll;   // 0: EI md(); >= 1: mean-free-path md() of a CID run (reference global method == 3, src/main.F90:1860-1866)
};

// per-trajectory state, SoA over trajectories
struct MdState {
    double *xyz, *velo, *grad, *achrg, *velof;        // [ntraj][3nat] / [ntraj][nat]
    double *avchrg, *avxyz;                           // running sums
    double *eimp, *tadd, *epot, *ekin, *ekinstart, *etemp, *Tav, *Epav, *Ekav, *Edum, *aTlast, *dtime, *ttime, *fadd;
    int *nstep, *kdump, *fconst, *morestep, *nfrag, *status, *fragstate, *mdok, *nadd, *list, *scc_total;
    // mean-free-path mode only (null otherwise): [ntraj][8] old_cm(3), new_velo, new_temp; [ntraj][16] cnt, count_average,
    // check_fragmented, max_steps, save_natf(10); [ntraj][3nat] avxyz2 and store_avxyz of src/md.f90:496-621
    double *mfp_d, *avxyz2, *store;
    int *mfp_i;
    double *qwarm;    // [ntraj][2 ndim + 1] converged populations of the last two steps + count; null unless the opt-in warm start is on
    double *gsdump;   // it_mode == 0: [ntraj][nmax][nat][6] positions and velocities of every step (the records of qcxms.gs), null otherwise
    double *eigseed;  // [ntraj][QX_OA_NSTORE nao^2 + 1] eigenvector seeds of the eigenpair refinement (DevModel::oa), null otherwise
};

enum { TRJ_RUNNING = 0, TRJ_FINISHED = 1, TRJ_FAILED = 2 };

__device__ inline double md_setetemp(const MdConfig &c, int nfrag, double eimp) {
    double etemp = 5000.0 + 20000.0 * c.ax;
    if (eimp > 0.0 && nfrag <= 1) etemp += fmax(eimp, 0.0) * c.ieetemp;
    return etemp;
}

// fragment_structure(nat, oz, xyz, rcut, 1, 0, frag): thread 0 of the CTA propagates labels through the
// connectivity computed by all threads.  conn: nat*nat bytes of scratch.
__device__ inline void md_fragments(const DevModel &m, const double *xyz, double rcut, unsigned char *conn, int *frag, int *stack, int natoms = -1) {
    const int nat = natoms > 0 ? natoms : m.nat;
    for (int t = threadIdx.x; t < nat * nat; t += QX_NT) {
        int i = t / nat, j = t - i * nat;
        unsigned char c = 0;
        if (i != j) {
            // reference evaluates the pair with i < j
            int lo = i < j ? i : j, hi = i < j ? j : i;
            double dx = __dsub_rn(xyz[3 * lo], xyz[3 * hi]), dy = __dsub_rn(xyz[3 * lo + 1], xyz[3 * hi + 1]), dz = __dsub_rn(xyz[3 * lo + 2], xyz[3 * hi + 2]);
            double r = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
            double rcov = __dmul_rn(__dmul_rn(rcut, 0.5), __dadd_rn(m.at_qcrad[lo], m.at_qcrad[hi]));
            c = r < rcov;
        }
        conn[t] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 0; i < nat; ++i) frag[i] = 0;
        int current = 0;
        for (int seed = 0; seed < nat; ++seed) {
            if (frag[seed] != 0) continue;
            ++current;
            int top = 0;
            stack[top++] = seed;
            frag[seed] = current;
            while (top > 0) {
                int i = stack[--top];
                for (int j = 0; j < nat; ++j)
                    if (conn[i * nat + j] && frag[j] == 0) { frag[j] = current; stack[top++] = j; }
            }
        }
    }
    __syncthreads();
}

// number of fragments as fragmass counts them (at most 10 slots, mass > 0)
__device__ inline int md_nfrag(const DevModel &m, const int *frag, int natoms = -1) {
    const int nat = natoms > 0 ? natoms : m.nat;
    int nf = 0;
    for (int f = 1; f <= 10; ++f) {
        double mass = 0.0;
        for (int i = 0; i < nat; ++i)
            if (frag[i] == f) mass += m.mass[i];
        if (mass > 0.0) ++nf;
    }
    return nf;
}

__device__ inline double md_ekinet_seq(int nat, const double *velo, const double *mass, double scalef, const double *velof) {
    double e = 0.0;
    for (int i = 0; i < nat; ++i) {
        double f = velof ? __dadd_rn(1.0, __dmul_rn(velof[i], scalef)) : 1.0;
        double vx = __dmul_rn(velo[3 * i], f), vy = __dmul_rn(velo[3 * i + 1], f), vz = __dmul_rn(velo[3 * i + 2], f);
        double v2 = __dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz));
        e = __dadd_rn(e, __dmul_rn(mass[i], v2));
    }
    return __dmul_rn(e, 0.5);
}

// impactscale: first grid point whose kinetic energy reaches Esoll - 0.001f, applied one grid step further, exactly like the
// reference loop.  The grid is the reference's running sum scal = scal + 0.0002 (single-precision literal, double accumulator,
// src/impact.f90:28-37), NOT k * 0.0002: m.scal_table[k] holds the sum after k additions (host-built, 20001 entries), so the
// candidates are bit-identical to the reference's.  Returns false on the reference's 'error in impactscale' (k >= 20000).
// Warp 0 evaluates 32 candidates per pass.
__device__ inline bool md_impactscale(const DevModel &m, double *velo, const double *velof, double eimp, double ff, double e0, int *flag) {
    const double tol = (double)0.001f;
    const double esoll = __dadd_rn(__dmul_rn(eimp, ff), e0);
    if (threadIdx.x < 32) {
        int found = -1;
        for (int base = 0; base < 20000 && found < 0; base += 32) {
            int c = base + threadIdx.x;  // k = c + 1
            double scal = m.scal_table[c < 20000 ? c : 20000];
            double e = md_ekinet_seq(m.nat, velo, m.mass, scal, velof);
            bool cont = (__dsub_rn(esoll, e) > tol) && (c + 1 < 20000);
            unsigned ball = __ballot_sync(0xffffffffu, !cont);
            if (ball) found = base + __ffs(ball) - 1;
        }
        if (threadIdx.x == 0) *flag = found;
    }
    __syncthreads();
    int found = *flag;
    if (found < 0 || found + 1 >= 20000) return false;
    double scal = m.scal_table[found + 1];
    for (int t = threadIdx.x; t < 3 * m.nat; t += QX_NT) velo[t] = __dmul_rn(velo[t], __dadd_rn(1.0, __dmul_rn(velof[t / 3], scal)));
    __syncthreads();
    return true;
}

// checkqc: returns ok; on failure the caller zeroes the energy (reference: E = 0)
__device__ inline bool md_checkqc(const DevModel &m, double e, const double *grad, const double *qat, int mchrg) {
    if (fabs(e) < 1e-8) return false;
    double gn = 0.0;
    for (int i = 0; i < m.nat; ++i) gn += grad[3 * i] * grad[3 * i] + grad[3 * i + 1] * grad[3 * i + 1] + grad[3 * i + 1] * grad[3 * i + 1];
    gn = sqrt(gn);
    if (gn < 1e-8 || gn > 20.0) return false;
    if (mchrg > 0) {
        double mx = qat[0];
        for (int i = 1; i < m.nat; ++i) mx = fmax(mx, qat[i]);
        if (fabs(mx) < 1e-5) return false;
    }
    return true;
}

}  // namespace qx
