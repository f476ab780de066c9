/* TEST INFRASTRUCTURE -- CPU restatement of the MD side of the QCxMS production-trajectory path.
 * Compiled with -ffp-contract=off so that every expression rounds like the reference's Fortran. */
#include "md_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../qcxms_b200/csrc/params/constants.h"
#include "../qcxms_b200/csrc/params/elem_tables.h"
#include "xtb_oracle.h"

/* reference src/md.f90:749-773 */
void md_oracle_leapfrog(int nat, const double *grad, const double *amass, double tstp, double *xyz, double *vel, double *ke) {
    double k_e = 0.0;
    for (int k = 0; k < nat; ++k) {
        double mass = amass[k];
        for (int j = 0; j < 3; ++j) {
            double velold = vel[3 * k + j];
            vel[3 * k + j] = vel[3 * k + j] - (tstp * grad[3 * k + j] / mass);
            double velavg = 0.5 * (velold + vel[3 * k + j]);
            xyz[3 * k + j] = xyz[3 * k + j] + (tstp * vel[3 * k + j]);
            k_e = k_e + 0.5 * (mass * velavg * velavg);
        }
    }
    *ke = k_e;
}

/* reference src/mdinit.f90:57-76 */
void md_oracle_ekinet(int nat, const double *velo, const double *mass, double *e_kin, double *temp) {
    double e = 0.0;
    for (int i = 0; i < nat; ++i) {
        double ms = mass[i];
        e = e + ms * (velo[3 * i] * velo[3 * i] + velo[3 * i + 1] * velo[3 * i + 1] + velo[3 * i + 2] * velo[3 * i + 2]);
    }
    e = e * 0.5;
    *e_kin = e;
    *temp = e / (0.5 * 3.0 * nat * QC_KB);
}

/* reference src/impact.f90:12-55; 0.0002 and 0.001 are default-real (single precision) literals there */
int md_oracle_impactscale(int nuc, double *velo, const double *mass, const double *velof, double eimp, double ff, double e0) {
    double *v = malloc(3 * nuc * sizeof(double));
    double Esoll = eimp * ff + e0, scal = 0.0, E, T;
    int k = 0;
    for (;;) {
        k = k + 1;
        for (int i = 0; i < nuc; ++i)
            for (int j = 0; j < 3; ++j) v[3 * i + j] = velo[3 * i + j] * (1.0 + velof[i] * scal);
        md_oracle_ekinet(nuc, v, mass, &E, &T);
        scal = scal + (double)0.0002f;
        if (Esoll - E > (double)0.001f && k < 20000) continue;
        break;
    }
    for (int i = 0; i < nuc; ++i)
        for (int j = 0; j < 3; ++j) velo[3 * i + j] = velo[3 * i + j] * (1.0 + velof[i] * scal);
    free(v);
    return k >= 20000; /* reference: stop 'error in impactscale' */
}

static double qc_rad(int z) { return QC_AATOAU * QCXMS_RAD_AA[z]; }

/* reference src/fragments.f90:93-182 */
void md_oracle_fragment_structure(int nat, const int32_t *oz, const double *xyz, double rcut, int at1, int at2, int32_t *frag) {
    unsigned char *connect = calloc((size_t)nat * nat, 1);
    for (int i = 0; i < nat - 1; ++i)
        for (int j = i + 1; j < nat; ++j) {
            double dx = xyz[3 * i] - xyz[3 * j], dy = xyz[3 * i + 1] - xyz[3 * j + 1], dz = xyz[3 * i + 2] - xyz[3 * j + 2];
            double r = sqrt(dx * dx + dy * dy + dz * dz);
            double rcov = rcut * 0.5 * (qc_rad(oz[i]) + qc_rad(oz[j]));
            if (r < rcov) connect[i * nat + j] = connect[j * nat + i] = 1;
        }
    if (at1 == 0 && at2 == 0) {
        for (int i = 0; i < nat; ++i) frag[i] = 1;
        free(connect);
        return;
    }
    for (int i = 0; i < nat; ++i) frag[i] = 0;
    frag[at1 - 1] = 1;
    int attotal = 1;
    if (at2 != 0) connect[(at1 - 1) * nat + at2 - 1] = connect[(at2 - 1) * nat + at1 - 1] = 0;
    int finish = 0, currentfrag = 0;
    while (attotal != nat) {
        currentfrag += 1;
        while (!finish) {
            finish = 1;
            for (int i = 0; i < nat; ++i)
                if (frag[i] == currentfrag)
                    for (int j = 0; j < nat; ++j)
                        if (connect[i * nat + j] && frag[j] == 0) {
                            frag[j] = currentfrag;
                            attotal += 1;
                            finish = 0;
                        }
        }
        for (int i = 0; i < nat; ++i)
            if (frag[i] == 0) {
                frag[i] = currentfrag + 1;
                attotal += 1;
                break;
            }
        finish = 0;
    }
    free(connect);
}

/* reference src/fragments.f90:10-84 (formula strings omitted: they are derived from fragat) */
void md_oracle_fragmass(int nat, const int32_t *iat, const int32_t *list, const double *mass, const int32_t *imass, int32_t *nfrag,
                        double *fragx, int32_t *fragat) {
    double fragm[10];
    int fragel[10][200];
    memset(fragm, 0, sizeof fragm);
    memset(fragel, 0, sizeof fragel);
    for (int i = 0; i < nat; ++i) {
        int f = list[i] - 1;
        if (f < 0 || f >= 10) continue; /* the Fortran would write out of bounds here */
        fragm[f] += mass[i];
        int j = iat[i];
        if (imass && imass[i] > 0) j = 100 + imass[i];
        if (j >= 1 && j <= 200) fragel[f][j - 1] += 1;
    }
    int nf = 0;
    for (int i = 0; i < 10; ++i)
        if (fragm[i] > 0) {
            if (fragx) fragx[nf] = fragm[i] * QC_AUTOAMU;
            if (fragat) memcpy(fragat + 200 * nf, fragel[i], 200 * sizeof(int32_t));
            nf += 1;
        }
    *nfrag = nf;
}

/* reference src/md.f90:715-741 */
void md_oracle_intenergy(int nuc, const int32_t *list, const double *mass, const double *velo, int nfrag, double *T, double *e_int) {
    int n[10] = {0};
    for (int i = 0; i < 10; ++i) e_int[i] = 0.0;
    for (int i = 0; i < nuc; ++i) {
        int j = list[i] - 1;
        e_int[j] = e_int[j] + 0.5 * mass[i] * (velo[3 * i] * velo[3 * i] + velo[3 * i + 1] * velo[3 * i + 1] + velo[3 * i + 2] * velo[3 * i + 2]);
        n[j] += 1;
    }
    for (int i = 0; i < nfrag; ++i) T[i] = e_int[i] / (0.5 * 3 * n[i] * QC_KB);
}

/* reference src/iniqm.f90:684-732; gnorm counts g_y twice and never g_z (:727) */
int md_oracle_checkqc(int nuc, double *e, const double *grad, const double *qat, int mchrg) {
    if (fabs(*e) < 1e-8) return 0;
    double gn = 0.0;
    for (int i = 0; i < nuc; ++i) gn = gn + grad[3 * i] * grad[3 * i] + grad[3 * i + 1] * grad[3 * i + 1] + grad[3 * i + 1] * grad[3 * i + 1];
    gn = sqrt(gn);
    if (gn < 1e-8 || gn > 20.0) { *e = 0; return 0; }
    if (mchrg > 0) {
        double mx = qat[0];
        for (int i = 1; i < nuc; ++i) if (qat[i] > mx) mx = qat[i];
        if (fabs(mx) < 1e-5) { *e = 0; return 0; }
    }
    return 1;
}

/* reference src/utility.f90:69-86 */
double md_oracle_setetemp(int nfrag, double eimp, double ax, double ieetemp) {
    double etemp = 5000. + 20000. * ax;
    if (eimp > 0 && nfrag <= 1) {
        double tmp = eimp > 0.0 ? eimp : 0.0;
        etemp = etemp + tmp * ieetemp;
    }
    return etemp;
}

/* reference src/utility.f90:449-464 */
int md_oracle_getspin(int nat, const int32_t *ic, int chrg) {
    int j = 0;
    for (int i = 0; i < nat; ++i) j += ic[i];
    j -= abs(chrg);
    int isp = 1 + j % 2;
    if (j < 1) isp = -1;
    return isp;
}

/* reference src/utility.f90:541-562 */
void md_oracle_center_of_mass(int nat, const double *mass, const double *xyz, double *cm) {
    double tm = 0.0;
    cm[0] = cm[1] = cm[2] = 0.0;
    for (int i = 0; i < nat; ++i) {
        tm += mass[i];
        for (int j = 0; j < 3; ++j) cm[j] += mass[i] * xyz[3 * i + j];
    }
    for (int j = 0; j < 3; ++j) cm[j] /= tm;
}

/* reference src/iniqm.f90:457-497, 641-655 */
int md_oracle_egrad(int nuc, const double *xyz, const int32_t *iat, int mchrg, double etemp, int method_id, double *E, double *grad,
                    double *qat, int *niter) {
    for (int i = 0; i < 3 * nuc; ++i) grad[i] = 0.0;
    for (int i = 0; i < nuc; ++i) qat[i] = 0.0;
    *E = 0.0;
    int idum = md_oracle_getspin(nuc, iat, mchrg);
    xtb_oracle_detail_t d;
    memset(&d, 0, sizeof d);
    xtb_oracle_egrad(nuc, iat, xyz, mchrg, idum, method_id, etemp, qat, E, grad, &d);
    if (niter) *niter = d.niter;
    /* ok = stat == 0 is immediately overwritten by checkqc (src/iniqm.f90:646-651) */
    int ok = md_oracle_checkqc(nuc, E, grad, qat, mchrg);
    return !ok;
}

static void natf_count(int nuc, const int32_t *list, int nfrag, int *natf);

/* reference src/md.f90:34-708 restricted to it > 0, No_eTemp = .false., Temprun = .false., starting_md = .false.:
 *   icoll == 0: EI (global method 0)
 *   icoll >= 1: the mean-free-path MD between two collisions of a CID run (global method 3; called from src/main.F90:1860-1866):
 *               no IEE heating (:438), kinetic energy and temperature without the motion of the centre of mass (:246-255, :466-493),
 *               fragment-structure averaging over 50 steps after a fragmentation (:496-621) with max_steps = nstep + add_steps (:507),
 *               tmax as the only regular exit (:672), error threshold 0.2 (:325), axyz from the averaged fragments (:694-699).
 *               new_velo (m/s) is in/out (:252, :474). */
/*   esi_tsoll > 0: the pre-collision heating MD of an ESI/CID run (global method 3, icoll = 0, starting_md = .true.; called from
 *               src/main.F90:1357-1362 with Tsoll = tscale, tadd = pretadd, eimp = E_Scale): Berendsen scaling of the velocities towards
 *               Tsoll during the first nadd steps while the ion is intact (:428-434); everything else as in the mean-free-path mode
 *               except that the kinetic energy keeps the centre-of-mass motion (:254, :466 need icoll >= 1). */
static int md_core(const qcxms_b200_md_config_t *cfg, int nuc, const int32_t *iat, const double *mass, double *xyz, double *velo,
                   const double *velof, double eimp, double tadd, int icoll, double *new_velo_io, int step_limit, double *grad,
                   int32_t *list, double *achrg, double *axyz, qcxms_b200_md_result_t *res, double esi_tsoll) {
    const double tstep = cfg->tstep;
    const int com = icoll > 0;                      /* kinetic energy without the centre-of-mass motion */
    const int cid = icoll > 0 || esi_tsoll > 0;     /* global method == 3 */
    double Ekin, T, Epot, etemp;
    int mdok = 0, nfrag = 1, fragstate = 0, scc_total = 0, niter = 0;
    md_oracle_ekinet(nuc, velo, mass, &Ekin, &T);
    double Ekinstart = Ekin;
    if (cfg->etemp_in < 0) etemp = md_oracle_setetemp(nfrag, eimp, cfg->ax, cfg->ieetemp);
    else etemp = cfg->etemp_in;
    md_oracle_egrad(nuc, xyz, iat, cfg->mchrg, etemp, cfg->method_id, &Epot, grad, achrg, &niter);
    scc_total += niter;
    memset(res, 0, sizeof *res);
    for (int i = 0; i < nuc; ++i) list[i] = 1;
    if (Epot == 0) { res->status = 2; return 0; }
    const int more = 250, avdump = 50, cnt_steps = 50;
    double Tav = 0, Epav = 0, Ekav = 0, Edum = 0, Eerror = 0, aTlast = 0, dtime = 0, ttime = 0, Eav;
    int nstep = 0, morestep = 0, fconst = 0, kdump = avdump;
    double *avchrg = calloc(nuc, sizeof(double)), *avxyz = calloc(3 * nuc, sizeof(double));
    double *avxyz2 = calloc(3 * nuc, sizeof(double)), *store = calloc(3 * nuc, sizeof(double));
    int nadd = (int)((tadd + tstep) / tstep - 1);
    double fadd = tstep / (tadd + tstep);
    /* md.f90:209-235, 246-255, 283 */
    int check_fragmented = 1, cnt = 0, count_average = 0, add_steps = 0, natf[10], save_natf[10] = {0};
    if (nuc > 10) add_steps = (nuc / 10) * 500;
    if (nuc >= 40) add_steps = (nuc / 10) * 1000;
    double old_cm[3], cm[3], summass = 0, new_velo = new_velo_io ? *new_velo_io : 0.0, new_temp = 0;
    md_oracle_center_of_mass(nuc, mass, xyz, old_cm);
    for (int i = 0; i < nuc; ++i) summass = summass + mass[i];
    {
        double E_kin = 0.5 * summass * ((new_velo * QC_MSTOAU) * (new_velo * QC_MSTOAU));
        double E_kin_diff = Ekin - E_kin;
        new_temp = (2 * E_kin_diff) / (3 * QC_KB * nuc);
        if (com) Ekin = E_kin_diff;
    }
    int max_steps = cfg->nmax;
    for (;;) {
        if (step_limit > 0 && nstep >= step_limit) break; /* bounded run requested by the caller (not in the reference) */
        nstep = nstep + 1;
        T = Ekin / (0.5 * 3 * nuc * QC_KB);
        Tav = Tav + T; Epav = Epav + Epot; Ekav = Ekav + Ekin;
        if (nstep > nadd) { Edum = Edum + Epot + Ekin; Eav = Edum / (double)(float)(nstep - nadd); }
        else Eav = Epot + Ekin;
        Eerror = Eav - Epot - Ekin;
        int err1 = Epot == 0, err2 = fabs(Eerror) > (cid ? (double)0.2f : (double)0.1f);
        if (err1 || (err2 && cfg->exit_rules)) {
            mdok = ((nfrag > 1 && nfrag <= 4) || cfg->isec > 1);
            break;
        }
        if (kdump > avdump - 1) {
            kdump = 0;
            memset(avchrg, 0, nuc * sizeof(double));
            memset(avxyz, 0, 3 * nuc * sizeof(double));
            aTlast = 0;
        }
        for (int i = 0; i < nuc; ++i) avchrg[i] += achrg[i];
        for (int i = 0; i < 3 * nuc; ++i) avxyz[i] += xyz[i];
        aTlast = aTlast + (com ? new_temp : T);
        md_oracle_leapfrog(nuc, grad, mass, tstep, xyz, velo, &Ekin);
        ttime = ttime + tstep / QC_FSTOAU;
        md_oracle_egrad(nuc, xyz, iat, cfg->mchrg, etemp, cfg->method_id, &Epot, grad, achrg, &niter);
        scc_total += niter;
        kdump = kdump + 1;
        if (nfrag == 1) morestep = 0;
        if (nfrag > 1 && dtime < 1e-6) dtime = ttime / 1000.;
        if (cid && !com && nfrag == 1 && nstep <= nadd) {   /* Berendsen thermostat of the heating MD, md.f90:428-434 */
            double sca = sqrt(1.0 + ((tstep / QC_FSTOAU) / 150) * (esi_tsoll / T - 1.0));
            for (int i = 0; i < 3 * nuc; ++i) velo[i] = sca * velo[i];
        }
        if (!cid) {
            if (nstep <= nadd && nfrag == 1) {
                if (md_oracle_impactscale(nuc, velo, mass, velof, eimp, fadd * nstep, Ekinstart)) { res->status = 2; break; }
            }
            {   /* unconditional in the reference (src/md.f90:443-445): a user ETEMP only serves the first single point */
                double dum = eimp - eimp * (double)(float)nstep / (double)(float)nadd;
                etemp = md_oracle_setetemp(nfrag, dum, cfg->ax, cfg->ieetemp);
            }
        }
        md_oracle_fragment_structure(nuc, iat, xyz, 3.0, 1, 0, list);
        md_oracle_fragmass(nuc, iat, list, mass, NULL, &nfrag, NULL, NULL);
        if (cid) {
            if (nfrag > 6) break;
            if (com) {
                /* kinetic energy without the centre-of-mass motion, md.f90:466-493 */
                md_oracle_center_of_mass(nuc, mass, xyz, cm);
                double d0 = cm[0] - old_cm[0], d1 = cm[1] - old_cm[1], d2 = cm[2] - old_cm[2];
                double cm_out = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
                new_velo = (cm_out / tstep) / QC_MSTOAU;
                for (int k = 0; k < 3; ++k) old_cm[k] = cm[k];
                double E_kin = 0.5 * summass * ((new_velo * QC_MSTOAU) * (new_velo * QC_MSTOAU));
                double E_kin_diff = Ekin - E_kin;
                new_temp = (2.0 * E_kin_diff) / (3.0 * QC_KB * nuc);
                Ekin = E_kin_diff;
            }
            /* averaged fragment structures, md.f90:496-621 */
            if (nfrag > check_fragmented) { count_average = 1; check_fragmented = nfrag; max_steps = nstep + add_steps; }
            if (nfrag < check_fragmented && count_average) {
                cnt = 0; memset(avxyz2, 0, 3 * nuc * sizeof(double)); memset(store, 0, 3 * nuc * sizeof(double));
                count_average = 0; check_fragmented = 1;
            }
            if (count_average) {
                cnt = cnt + 1;
                for (int i = 0; i < 3 * nuc; ++i) { avxyz2[i] += xyz[i]; store[i] = avxyz2[i] / cnt; }
                natf_count(nuc, list, nfrag, natf);
                for (int i = 0; i < nfrag && i < 10; ++i) {
                    if (cnt == 1) save_natf[i] = natf[i];
                    if (natf[i] != save_natf[i]) {
                        cnt = 0; memset(store, 0, 3 * nuc * sizeof(double)); memset(avxyz2, 0, 3 * nuc * sizeof(double));
                        break;
                    }
                }
                if (cnt == cnt_steps) {
                    for (int i = 0; i < 3 * nuc; ++i) store[i] = avxyz2[i] / cnt;
                    memset(avxyz2, 0, 3 * nuc * sizeof(double)); cnt = 0; count_average = 0;
                }
            }
        } else if (cfg->exit_rules) {
            if (nfrag > 6) break;
            if (nfrag > cfg->nfragexit) { fragstate = 1; mdok = 1; break; }
            if (nfrag >= 2) fconst = fconst + 1; else fconst = 0;
            if (fconst > 1000) { fragstate = 2; mdok = 1; break; }
            if (nfrag >= cfg->nfragexit) {
                morestep = morestep + 1;
                if (morestep > more) { fragstate = 1; mdok = 1; break; }
            }
        }
        if (nstep >= max_steps) { fragstate = 1; mdok = 1; break; }
    }
    res->mdok = mdok; res->fragstate = fragstate; res->nstep = nstep; res->nfrag = nfrag;
    if (res->status == 0) res->status = 1;
    res->scc_iter_total = scc_total;
    res->Tav = Tav / nstep; res->Epav = Epav / nstep; res->Ekav = Ekav / nstep;
    for (int i = 0; i < nuc; ++i) achrg[i] = avchrg[i] / kdump;
    for (int i = 0; i < 3 * nuc; ++i) axyz[i] = (cid && check_fragmented > 1) ? store[i] : avxyz[i] / kdump;
    res->aTlast = aTlast / kdump; res->dtime = dtime; res->ttime = ttime; res->Epot = Epot; res->Ekin = Ekin;
    if (new_velo_io) *new_velo_io = new_velo;
    free(avchrg); free(avxyz); free(avxyz2); free(store);
    return 0;
}

/* md() for it = -1 (equilibration) and it = 0 (sampling) -- reference src/md.f90:34-708 with it <= 0: etemp = etempin on every step
 * (:290-297), nadd = 0 and velof = 1 (:270-274), Eerror = 0 for it < 0 (:318), qcxms.gs record before the step for it == 0
 * (:380-385), velocity rescaling towards Tsoll for it < 0 (:402-410), no fragment section (:416), tmax as the exit (:672).
 * gs (may be NULL): [nmax][nuc][6] records. */
int md_oracle_md_gs(const qcxms_b200_md_config_t *cfg, int it, double Tsoll, int nuc, const int32_t *iat, const double *mass, double *xyz,
                    double *velo, double *grad, double *achrg, double *gs, qcxms_b200_md_result_t *res) {
    const double tstep = cfg->tstep, etemp = cfg->etemp_in;
    double Ekin, T, Epot;
    int niter = 0, scc_total = 0, mdok = 0, nstep = 0;
    md_oracle_ekinet(nuc, velo, mass, &Ekin, &T);
    md_oracle_egrad(nuc, xyz, iat, cfg->mchrg, etemp, cfg->method_id, &Epot, grad, achrg, &niter);
    scc_total += niter;
    memset(res, 0, sizeof *res);
    if (Epot == 0) { res->status = 2; return 0; }
    double Tav = 0, Epav = 0, Ekav = 0, Edum = 0, Eav, Eerror, ttime = 0;
    const int nadd = 0;
    for (;;) {
        nstep = nstep + 1;
        T = Ekin / (0.5 * 3 * nuc * QC_KB);
        Tav = Tav + T; Epav = Epav + Epot; Ekav = Ekav + Ekin;
        if (nstep > nadd) { Edum = Edum + Epot + Ekin; Eav = Edum / (double)(float)(nstep - nadd); }
        else Eav = Epot + Ekin;
        Eerror = Eav - Epot - Ekin;
        if (it < 0) Eerror = 0;
        if (Epot == 0 || (fabs(Eerror) > (double)0.1f && cfg->exit_rules)) { mdok = cfg->isec > 1; break; }
        if (it == 0 && gs)
            for (int i = 0; i < nuc; ++i)
                for (int j = 0; j < 3; ++j) { gs[((size_t)(nstep - 1) * nuc + i) * 6 + j] = xyz[3 * i + j]; gs[((size_t)(nstep - 1) * nuc + i) * 6 + 3 + j] = velo[3 * i + j]; }
        md_oracle_leapfrog(nuc, grad, mass, tstep, xyz, velo, &Ekin);
        ttime = ttime + tstep / QC_FSTOAU;
        md_oracle_egrad(nuc, xyz, iat, cfg->mchrg, etemp, cfg->method_id, &Epot, grad, achrg, &niter);
        scc_total += niter;
        double dum = 100. * fabs(Tav / nstep - Tsoll) / Tsoll;
        if (dum > 5.0 && it < 0 && nstep > 50) {
            double f = sqrt(Tav / nstep / Tsoll);
            for (int i = 0; i < 3 * nuc; ++i) velo[i] = velo[i] / f;
        }
        if (nstep >= cfg->nmax) { mdok = 1; break; }
    }
    res->mdok = mdok; res->fragstate = mdok ? 1 : 0; res->nstep = nstep; res->nfrag = 1; res->status = 1; res->scc_iter_total = scc_total;
    res->Tav = Tav / nstep; res->Epav = Epav / nstep; res->Ekav = Ekav / nstep; res->ttime = ttime; res->Epot = Epot; res->Ekin = Ekin;
    return 0;
}

int md_oracle_md(const qcxms_b200_md_config_t *cfg, int nuc, const int32_t *iat, const double *mass, double *xyz, double *velo,
                 const double *velof, double eimp, double tadd, int step_limit, double *grad, int32_t *list, double *achrg, double *axyz,
                 qcxms_b200_md_result_t *res) {
    return md_core(cfg, nuc, iat, mass, xyz, velo, velof, eimp, tadd, 0, NULL, step_limit, grad, list, achrg, axyz, res, 0.0);
}

int md_oracle_md_mfp(const qcxms_b200_md_config_t *cfg, int nuc, const int32_t *iat, const double *mass, double *xyz, double *velo,
                     int icoll, double *new_velo, int step_limit, double *grad, int32_t *list, double *achrg, double *axyz,
                     qcxms_b200_md_result_t *res) {
    if (icoll < 1 || !new_velo) return 1;
    return md_core(cfg, nuc, iat, mass, xyz, velo, NULL, 0.0, 0.0, icoll, new_velo, step_limit, grad, list, achrg, axyz, res, 0.0);
}

/* md() as the heating MD before the first collision of an ESI/CID run (method 3, icoll = 0, starting_md; src/main.F90:1357-1362) */
int md_oracle_md_esi(const qcxms_b200_md_config_t *cfg, int nuc, const int32_t *iat, const double *mass, double *xyz, double *velo,
                     double tsoll, double eimp, double tadd, int step_limit, double *grad, int32_t *list, double *achrg, double *axyz,
                     qcxms_b200_md_result_t *res) {
    if (!(tsoll > 0.0)) return 1;
    double new_velo = 0.0;
    return md_core(cfg, nuc, iat, mass, xyz, velo, NULL, eimp, tadd, 0, &new_velo, step_limit, grad, list, achrg, axyz, res, tsoll);
}

/* ======================================================================================== CID
 * reference src/diag3x3.f90:85-260 (analytic eigen-decomposition of a symmetric 3x3 matrix; Fortran a(i,j) == a[i-1][j-1]) */
static void eigval3x3(double a[3][3], double w[3]) {
    const double twothirdpi = 8.0 * atan(1.0) / 3.0;
    double r = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    double q = (a[0][0] + a[1][1] + a[2][2]) / 3.0;
    w[0] = a[0][0] - q; w[1] = a[1][1] - q; w[2] = a[2][2] - q;
    double p = sqrt((w[0] * w[0] + w[1] * w[1] + w[2] * w[2] + 2 * r) / 6.0);
    r = (w[0] * (w[1] * w[2] - a[1][2] * a[1][2]) - a[0][1] * (a[0][1] * w[2] - a[1][2] * a[0][2]) +
         a[0][2] * (a[0][1] * a[1][2] - w[1] * a[0][2])) / (p * p * p) * 0.5;
    if (r <= -1.0) r = 0.5 * twothirdpi;
    else if (r >= 1.0) r = 0.0;
    else r = acos(r) / 3.0;
    w[2] = q + 2 * p * cos(r);
    w[0] = q + 2 * p * cos(r + twothirdpi);
    w[1] = 3 * q - w[0] - w[2];
}

void md_oracle_eigvec3x3(double a[3][3], double w[3], double q[3][3]) {
    const double eps = 2.220446049250313e-16;
    double norm, n1, n2, n3, precon;
    int i;
    w[0] = fmax(fabs(a[0][0]), fabs(a[0][1]));
    w[1] = fmax(fabs(a[0][2]), fabs(a[1][1]));
    w[2] = fmax(fabs(a[1][2]), fabs(a[2][2]));
    precon = fmax(w[0], fmax(w[1], w[2]));
    if (precon < eps) {
        w[0] = w[1] = w[2] = 0.0;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) q[r][c] = r == c;
        return;
    }
    norm = 1.0 / precon;
    a[0][0] *= norm; a[0][1] *= norm; a[1][1] *= norm; a[0][2] *= norm; a[1][2] *= norm; a[2][2] *= norm;
    eigval3x3(a, w);
    a[0][0] -= w[0]; a[1][1] -= w[0]; a[2][2] -= w[0];
    q[0][0] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
    q[1][0] = a[0][2] * a[0][1] - a[0][0] * a[1][2];
    q[2][0] = a[0][0] * a[1][1] - a[0][1] * a[0][1];
    q[0][1] = a[0][1] * a[2][2] - a[0][2] * a[1][2];
    q[1][1] = a[0][2] * a[0][2] - a[0][0] * a[2][2];
    q[2][1] = a[0][0] * a[1][2] - a[0][1] * a[0][2];
    q[0][2] = a[1][1] * a[2][2] - a[1][2] * a[1][2];
    q[1][2] = a[1][2] * a[0][2] - a[0][1] * a[2][2];
    q[2][2] = a[0][1] * a[1][2] - a[1][1] * a[0][2];
    n1 = q[0][0] * q[0][0] + q[1][0] * q[1][0] + q[2][0] * q[2][0];
    n2 = q[0][1] * q[0][1] + q[1][1] * q[1][1] + q[2][1] * q[2][1];
    n3 = q[0][2] * q[0][2] + q[1][2] * q[1][2] + q[2][2] * q[2][2];
    norm = n1; i = 1;
    if (n2 > norm) { i = 2; norm = n1; }   /* sic: the reference keeps norm = n1 here */
    if (n3 > norm) i = 3;
    if (i == 1) { norm = sqrt(1.0 / n1); q[0][0] *= norm; q[1][0] *= norm; q[2][0] *= norm; }
    else if (i == 2) { norm = sqrt(1.0 / n2); q[0][0] = q[0][1] * norm; q[1][0] = q[1][1] * norm; q[2][0] = q[2][1] * norm; }
    else { norm = sqrt(1.0 / n3); q[0][0] = q[0][2] * norm; q[1][0] = q[1][2] * norm; q[2][0] = q[2][2] * norm; }
    if (fabs(q[0][0]) > fabs(q[1][0])) {
        norm = sqrt(1.0 / (q[0][0] * q[0][0] + q[2][0] * q[2][0]));
        q[0][1] = -q[2][0] * norm; q[1][1] = 0.0; q[2][1] = +q[0][0] * norm;
    } else {
        norm = sqrt(1.0 / (q[1][0] * q[1][0] + q[2][0] * q[2][0]));
        q[0][1] = 0.0; q[1][1] = +q[2][0] * norm; q[2][1] = -q[1][0] * norm;
    }
    q[0][2] = q[1][0] * q[2][1] - q[2][0] * q[1][1];
    q[1][2] = q[2][0] * q[0][1] - q[0][0] * q[2][1];
    q[2][2] = q[0][0] * q[1][1] - q[1][0] * q[0][1];
    a[0][0] += w[0]; a[1][1] += w[0]; a[2][2] += w[0];
    n1 = a[0][0] * q[0][1] + a[0][1] * q[1][1] + a[0][2] * q[2][1];
    n2 = a[0][1] * q[0][1] + a[1][1] * q[1][1] + a[1][2] * q[2][1];
    n3 = a[0][2] * q[0][1] + a[1][2] * q[1][1] + a[2][2] * q[2][1];
    a[2][2] = a[0][2] * q[0][2] + a[1][2] * q[1][2] + a[2][2] * q[2][2];
    a[0][2] = a[0][0] * q[0][2] + a[0][1] * q[1][2] + a[0][2] * q[2][2];
    a[1][2] = a[0][1] * q[0][2] + a[1][1] * q[1][2] + a[1][2] * q[2][2];
    n1 = q[0][1] * n1 + q[1][1] * n2 + q[2][1] * n3 - w[1];
    n2 = q[0][1] * a[0][2] + q[1][1] * a[1][2] + q[2][1] * a[2][2];
    n3 = q[0][2] * a[0][2] + q[1][2] * a[1][2] + q[2][2] * a[2][2] - w[1];
    if (fabs(n1) >= fabs(n3)) {
        norm = fmax(fabs(n1), fabs(n2));
        if (norm > eps) {
            if (fabs(n1) >= fabs(n2)) { n2 = n2 / n1; n1 = sqrt(1.0 / (1.0 + n2 * n2)); n2 = n2 * n1; }
            else { n1 = n1 / n2; n2 = sqrt(1.0 / (1.0 + n1 * n1)); n1 = n1 * n2; }
            q[0][1] = n2 * q[0][1] - n1 * q[0][2];
            q[1][1] = n2 * q[1][1] - n1 * q[1][2];
            q[2][1] = n2 * q[2][1] - n1 * q[2][2];
        }
    } else {
        norm = fmax(fabs(n3), fabs(n2));
        if (norm > eps) {
            if (fabs(n3) >= fabs(n2)) { n2 = n2 / n3; n3 = sqrt(1.0 / (1.0 + n2 * n2)); n2 = n2 * n3; }
            else { n3 = n3 / n2; n2 = sqrt(1.0 / (1.0 + n3 * n3)); n3 = n3 * n2; }
            q[0][1] = n3 * q[0][1] - n2 * q[0][2];
            q[1][1] = n3 * q[1][1] - n2 * q[1][2];
            q[2][1] = n3 * q[2][1] - n2 * q[2][2];
        }
    }
    q[0][2] = q[1][0] * q[2][1] - q[2][0] * q[1][1];
    q[1][2] = q[2][0] * q[0][1] - q[0][0] * q[2][1];
    q[2][2] = q[0][0] * q[1][1] - q[1][0] * q[0][1];
    w[0] *= precon; w[1] *= precon; w[2] *= precon;
}

/* reference src/rotation.f90:11-88: R = R_alpha R_beta R_gamma applied to coordinates and velocities */
void md_oracle_euler_rotation(int nuc, double *xyz, double *velo, double a, double b, double c) {
    const double pi = 3.14159265358979323846264338327950288;
    double al = a * 2 * pi, be = b * 2 * pi, ga = c * pi;
    /* Fortran reshape fills column-major: rot(i,j) below is the mathematical element (row i, column j) */
    double ra[3][3] = {{1, 0, 0}, {0, cos(al), -sin(al)}, {0, sin(al), cos(al)}};
    double rb[3][3] = {{cos(be), 0, sin(be)}, {0, 1, 0}, {-sin(be), 0, cos(be)}};
    double rg[3][3] = {{cos(ga), -sin(ga), 0}, {sin(ga), cos(ga), 0}, {0, 0, 1}};
    double d[3][3], R[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { d[i][j] = 0; for (int k = 0; k < 3; ++k) d[i][j] += ra[i][k] * rb[k][j]; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { R[i][j] = 0; for (int k = 0; k < 3; ++k) R[i][j] += d[i][k] * rg[k][j]; }
    for (int i = 0; i < nuc; ++i) {
        double r[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, v[3] = {velo[3 * i], velo[3 * i + 1], velo[3 * i + 2]};
        for (int k = 0; k < 3; ++k) {
            xyz[3 * i + k] = R[k][0] * r[0] + R[k][1] * r[1] + R[k][2] * r[2];
            velo[3 * i + k] = R[k][0] * v[0] + R[k][1] * v[1] + R[k][2] * v[2];
        }
    }
}

/* reference src/rotation.f90:92-182 */
void md_oracle_rotation_velo(const double *xyz, int nuc, const double *mass, const double *velo, double *velo_rot, double *e_rot) {
    double mat[3][3] = {{0}}, ev[3], evec[3][3], w_new[3], om[3][3], Ekin, Tinit;
    for (int i = 0; i < nuc; ++i) {
        double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2], m = mass[i];
        mat[0][0] += (y * y + z * z) * m; mat[1][0] += (-x * y) * m; mat[2][0] += (-x * z) * m;
        mat[0][1] += (-y * x) * m; mat[1][1] += (x * x + z * z) * m; mat[2][1] += (-y * z) * m;
        mat[0][2] += (-z * x) * m; mat[1][2] += (-z * y) * m; mat[2][2] += (x * x + y * y) * m;
    }
    md_oracle_eigvec3x3(mat, ev, evec);
    md_oracle_ekinet(nuc, velo, mass, &Ekin, &Tinit);
    for (int k = 0; k < 3; ++k) w_new[k] = sqrt((QC_KB * Tinit) / ev[k]);
    for (int i = 0; i < 3; ++i) for (int r = 0; r < 3; ++r) om[r][i] = evec[r][i] * w_new[i];
    for (int i = 0; i < nuc; ++i) {
        double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2], vx = 0, vy = 0, vz = 0;
        for (int j = 0; j < 3; ++j) {
            vx = vx + (om[1][j] * z - om[2][j] * y);
            vy = vy + (om[2][j] * x - om[0][j] * z);
            vz = vz + (om[0][j] * y - om[1][j] * x);
        }
        velo_rot[3 * i] = vx; velo_rot[3 * i + 1] = vy; velo_rot[3 * i + 2] = vz;
    }
    double e = 0.0;
    for (int k = 0; k < 3; ++k) e = e + (0.5 * ev[k] * (w_new[k] * w_new[k]));
    if (e_rot) *e_rot = e;
}

/* reference src/boxmuller.f90:46-76 */
double md_oracle_vary_energies(double e_in, double e_distr, double dum, double dum2) {
    const double pi = 3.14159265358979323846264338327950288;
    double sigma = e_in * e_distr;
    double z0 = sqrt(-2.0 * log(dum)) * cos(2.0 * pi * dum2), z1 = sqrt(-2.0 * log(dum)) * sin(2.0 * pi * dum2);
    return dum > 0.5 ? z0 * sigma + e_in : z1 * sigma + e_in;
}

/* number of atoms per fragment as avg_frag_struc counts them (reference src/analyse.f90:452-502) */
static void natf_count(int nuc, const int32_t *list, int nfrag, int *natf) {
    for (int i = 0; i < 10; ++i) natf[i] = 0;
    for (int i = 0; i < nuc; ++i) if (list[i] >= 1 && list[i] <= nfrag && list[i] <= 10) natf[list[i] - 1] += 1;
}

/* reference src/cid.f90:24-1111 for mono-atomic gases, ConstVelo / MinPot / vScale off */
int md_oracle_cid(const qcxms_b200_cid_config_t *cfg, int nuc, const int32_t *iat, const double *mass, int icoll, double *xyz,
                  double *velo, const double *rnd, double velo_cm_in, double *direc, int32_t *collided_io, double *grad, double *achrg,
                  double *axyz, int32_t *list, qcxms_b200_cid_result_t *res) {
    const double time_step = cfg->tstep, autofs = 1.0 / QC_FSTOAU;
    const int ngas = cfg->gas_z == 7 ? 2 : 1;   /* N2: two atoms (reference src/cid.f90:179-180, 660-667) */
    const int nuc0 = nuc + ngas, dumpdist = 10, dumpavg = 50, cnt_steps = 50;
    const int ig = nuc0 - 1;                    /* the atom distArCOM looks at (src/cid.f90:1124-1139) */
    int ntot = cfg->ntot > 0 ? cfg->ntot : 15000;
    double etemp = cfg->etemp <= 0 ? 5000.0 : cfg->etemp;
    int add_steps = 0;
    if (nuc > 10) add_steps = (nuc / 10) * 500;
    if (nuc >= 40) add_steps = (nuc / 10) * 1000;
    int collided = *collided_io, fragmented = 0, count_average = 0, check_fragmented = 1, cnt = 0, nfrag = 1, scc_total = 0, niter = 0;
    double cm[3], old_cm[3];
    memset(res, 0, sizeof *res);
    double *xyz0 = calloc(3 * nuc0, sizeof(double)), *velo0 = calloc(3 * nuc0, sizeof(double)), *grad0 = calloc(3 * nuc0, sizeof(double));
    double *mass0 = calloc(nuc0, sizeof(double)), *achrg0 = calloc(nuc0, sizeof(double)), *velo_rot = calloc(3 * nuc, sizeof(double));
    double *avxyz = calloc(3 * nuc, sizeof(double)), *avxyz2 = calloc(3 * nuc, sizeof(double)), *store = calloc(3 * nuc, sizeof(double));
    int32_t *iat0 = calloc(nuc0, sizeof(int32_t));
    int natf[10], save_natf[10] = {0};

    if (icoll == 1) {
        md_oracle_center_of_mass(nuc, mass, xyz, cm);
        for (int i = 0; i < nuc; ++i) for (int k = 0; k < 3; ++k) xyz[3 * i + k] -= cm[k];
        md_oracle_euler_rotation(nuc, xyz, velo, rnd[0], rnd[1], rnd[2]);
        md_oracle_rotation_velo(xyz, nuc, mass, velo, velo_rot, NULL);
        for (int i = 0; i < 3 * nuc; ++i) velo[i] = velo[i] + velo_rot[i];
    }
    double summass = 0.0;
    for (int i = 0; i < nuc; ++i) summass = summass + mass[i];
    double beta = cfg->gas_mass / (cfg->gas_mass + summass), E_velo, fasti = 0.0;
    if (icoll == 1) {
        double Eimpact = cfg->ecom > 0.0 ? cfg->ecom / beta : cfg->elab;
        if (!cfg->eexact) E_velo = md_oracle_vary_energies(Eimpact, 0.1, rnd[3], rnd[4]) * QC_EVTOAU;
        else E_velo = Eimpact * QC_EVTOAU;
    } else
        E_velo = 0.5 * summass * ((velo_cm_in * QC_MSTOAU) * (velo_cm_in * QC_MSTOAU));
    double Ekin, Tinit, T;
    md_oracle_ekinet(nuc, velo, mass, &Ekin, &Tinit);
    if (icoll == 1) fasti = sqrt(2 * E_velo / summass);
    md_oracle_center_of_mass(nuc, mass, xyz, cm);
    double f = rnd[5], g = rnd[6], lmin = rnd[7], lpos = rnd[8];
    double lowestx = 1.7976931348623157e308, lowesty = lowestx, highestx = -lowestx, highesty = -lowestx;
    for (int i = 0; i < nuc; ++i) {
        double x = xyz[3 * i], y = xyz[3 * i + 1];
        if (x < lowestx) lowestx = x;
        if (x > highestx) highestx = x;
        if (y < lowesty) lowesty = y;
        if (y > highesty) highesty = y;
    }
    double diff1 = lmin < 0.5 ? lowesty * f : highesty * f;
    double diff2 = lpos < 0.5 ? lowestx * g : highestx * g;
    int step_dist = cfg->manual_dist == 0 ? (2 * nuc * 10 > 800 ? 800 : 2 * nuc * 10) : cfg->manual_dist;
    double start_dist, xyzAr[3], scale_velo[3];
    if (icoll == 1) {
        start_dist = fasti * (2 * time_step);
        start_dist = step_dist * start_dist * QC_AUTOAA;   /* Angstrom value added to bohr coordinates as is (sic) */
        if (start_dist < 10.0) start_dist = 10.0;
        double xs[3] = {cm[0], cm[1], cm[2] + start_dist};
        for (int k = 0; k < 3; ++k) direc[k] = xs[k];
        for (int k = 0; k < 3; ++k) direc[k] = direc[k] / sqrt(direc[0] * direc[0] + direc[1] * direc[1] + direc[2] * direc[2]);  /* sic: sequential */
        xyzAr[0] = xs[0] + diff2 * 0.8; xyzAr[1] = xs[1] + diff1 * 0.8; xyzAr[2] = xs[2];
        for (int k = 0; k < 3; ++k) scale_velo[k] = direc[k] * fasti;
    } else {
        start_dist = (velo_cm_in * QC_MSTOAU) * (2 * time_step);
        start_dist = step_dist * start_dist * QC_AUTOAA;
        if (start_dist < 17.0) start_dist = 17.0;
        for (int i = 0; i < nuc; ++i) for (int k = 0; k < 3; ++k) xyz[3 * i + k] -= cm[k];
        xyzAr[0] = cm[0] + direc[0] * start_dist + diff2 * 0.7;
        xyzAr[1] = cm[1] + direc[1] * start_dist + diff1 * 0.7;
        xyzAr[2] = cm[2] + direc[2] * start_dist;
        scale_velo[0] = scale_velo[1] = scale_velo[2] = 0.0;
    }
    for (int k = 0; k < 3; ++k) old_cm[k] = cm[k];
    md_oracle_center_of_mass(nuc, mass, xyz, cm);
    for (int i = 0; i < nuc; ++i) {
        for (int k = 0; k < 3; ++k) { velo0[3 * i + k] = velo[3 * i + k] + scale_velo[k]; xyz0[3 * i + k] = xyz[3 * i + k]; }
        mass0[i] = mass[i]; iat0[i] = iat[i];
    }
    for (int k = 0; k < 3; ++k) { xyz0[3 * ig + k] = xyzAr[k]; velo0[3 * ig + k] = 0.0; }
    mass0[ig] = cfg->gas_mass; iat0[ig] = cfg->gas_z;
    if (ngas == 2) {
        xyz0[3 * nuc] = xyzAr[0]; xyz0[3 * nuc + 1] = xyzAr[1]; xyz0[3 * nuc + 2] = xyzAr[2] + 1.09 * QC_AATOAU;
        velo0[3 * nuc] = velo0[3 * nuc + 1] = velo0[3 * nuc + 2] = 0.0;
        mass0[nuc] = cfg->gas_mass; iat0[nuc] = cfg->gas_z;
    }

    /* iniqm (one single point whose result is only checked) + initial egrad */
    double E;
    int gradfail = md_oracle_egrad(nuc0, xyz0, iat0, cfg->mchrg, etemp, cfg->method_id, &E, grad0, achrg0, &niter);
    scc_total += niter;
    if (gradfail) { res->stopcid = 1; res->status = 2; goto done; }

    {
        int nstep = 0, m = 0, step_counter = 0, distance_dump = 0, xyzavg_dump = 0, total_steps = ntot;
        double Tav = 0.0, new_velo = 0.0, new_dist, lowestCOM, ttime = 0.0, aTlast = 0.0, avgT = 0.0, ke;
        md_oracle_center_of_mass(nuc, mass, xyz, cm);
        {
            double d0 = xyz0[3 * ig] - cm[0], d1 = xyz0[3 * ig + 1] - cm[1], d2 = xyz0[3 * ig + 2] - cm[2];
            new_dist = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        }
        lowestCOM = new_dist;
        for (;;) {
            nstep = nstep + 1;
            if (xyzavg_dump == dumpavg) { xyzavg_dump = 0; memset(avxyz, 0, 3 * nuc * sizeof(double)); }
            ttime = ttime + time_step * autofs;
            distance_dump += 1; xyzavg_dump += 1;
            md_oracle_leapfrog(nuc0, grad0, mass0, time_step, xyz0, velo0, &ke);
            gradfail = md_oracle_egrad(nuc0, xyz0, iat0, cfg->mchrg, etemp, cfg->method_id, &E, grad0, achrg0, &niter);
            scc_total += niter;
            if (gradfail) { res->stopcid = 1; break; }
            md_oracle_center_of_mass(nuc, mass0, xyz0, cm);
            double dcm[3] = {cm[0] - old_cm[0], cm[1] - old_cm[1], cm[2] - old_cm[2]};
            double cm_out = sqrt(dcm[0] * dcm[0] + dcm[1] * dcm[1] + dcm[2] * dcm[2]);
            for (int k = 0; k < 3; ++k) old_cm[k] = cm[k];
            new_velo = nstep != 1 ? (cm_out / time_step) / QC_MSTOAU : 0.0;
            md_oracle_ekinet(nuc, velo0, mass0, &Ekin, &T);
            E_velo = 0.5 * summass * ((new_velo * QC_MSTOAU) * (new_velo * QC_MSTOAU));
            double new_temp = (2 * (Ekin - E_velo)) / (3 * QC_KB * nuc);
            if (nstep == 1) new_temp = Tinit;
            Tav = Tav + new_temp; m = m + 1; avgT = Tav / m;
            md_oracle_fragment_structure(nuc, iat0, xyz0, 3.0, 1, 0, list);
            md_oracle_fragmass(nuc, iat0, list, mass0, NULL, &nfrag, NULL, NULL);
            for (int i = 0; i < 3 * nuc; ++i) avxyz[i] += xyz0[i];
            if (nfrag > check_fragmented) { count_average = 1; check_fragmented = nfrag; }
            if (nfrag < check_fragmented && count_average) {
                cnt = 0; memset(avxyz2, 0, 3 * nuc * sizeof(double)); memset(store, 0, 3 * nuc * sizeof(double));
                count_average = 0; check_fragmented = 1;
            }
            if (count_average) {
                cnt = cnt + 1;
                for (int i = 0; i < 3 * nuc; ++i) { avxyz2[i] += xyz0[i]; store[i] = avxyz2[i] / cnt; }
                natf_count(nuc, list, nfrag, natf);
                for (int i = 0; i < nfrag && i < 10; ++i) {
                    if (cnt == 1) save_natf[i] = natf[i];
                    if (natf[i] != save_natf[i]) {
                        cnt = 0; memset(store, 0, 3 * nuc * sizeof(double)); memset(avxyz2, 0, 3 * nuc * sizeof(double));
                        break;
                    }
                }
                if (cnt == cnt_steps) {
                    for (int i = 0; i < 3 * nuc; ++i) store[i] = avxyz2[i] / cnt;
                    memset(avxyz2, 0, 3 * nuc * sizeof(double)); cnt = 0; count_average = 0;
                }
            }
            aTlast = avgT;
            if (distance_dump == dumpdist) {
                distance_dump = 0;
                md_oracle_center_of_mass(nuc, mass0, xyz0, cm);
                double d0 = xyz0[3 * ig] - cm[0], d1 = xyz0[3 * ig + 1] - cm[1], d2 = xyz0[3 * ig + 2] - cm[2];
                new_dist = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
                if (new_dist < lowestCOM) lowestCOM = new_dist;
                if (lowestCOM < new_dist) step_counter = step_counter + 1; else step_counter = 0;
                if (step_counter == 5) {
                    total_steps = nstep + (int)lround(800.0 * (2 * time_step * autofs));
                    collided = 1; Tav = 0; m = 0;
                }
            }
            if (nfrag > 1 && collided && !fragmented) { total_steps = nstep + add_steps; fragmented = 1; }
            if (nstep >= total_steps) { res->stopcid = 0; break; }
        }
        for (int i = 0; i < 3 * nuc; ++i) { xyz[i] = xyz0[i]; velo[i] = velo0[i]; grad[i] = grad0[i]; }
        for (int i = 0; i < nuc; ++i) achrg[i] = achrg0[i];
        for (int i = 0; i < 3 * nuc; ++i) axyz[i] = check_fragmented > 1 ? store[i] : avxyz[i] / xyzavg_dump;
        res->nstep = nstep; res->nfrag = nfrag; res->velo_cm = new_velo; res->aTlast = aTlast; res->ttime = ttime; res->epot = E;
        res->status = 1;
    }
done:
    res->collided = collided; res->scc_iter_total = scc_total;
    for (int k = 0; k < 3; ++k) res->direc[k] = direc[k];
    *collided_io = collided;
    free(xyz0); free(velo0); free(grad0); free(mass0); free(achrg0); free(velo_rot); free(avxyz); free(avxyz2); free(store); free(iat0);
    return 0;
}

/* ======================================================================================== fragment records
 * reference src/utility.f90:469-498 (units 2: eV) */
void md_oracle_boltz(int nfrag, double temp, const double *ip, double *pop) {
    double f = temp * QC_KB * QC_AUTOEV, esum = 0;
    for (int i = 0; i < nfrag; ++i) esum = esum + exp(-ip[i] / f);
    for (int i = 0; i < nfrag; ++i) pop[i] = exp(-ip[i] / f) / esum;
}

/* One record in the format '(F10.7,i3,2i5,2i2,2x,i3,2x,20(i4,i3))' (reference src/write_fragments.f90:402-441).
 * icoll < 0: EI item list (charge, mchrg, itrj, isec, j, l, pairs) -- one integer fewer than the descriptors expect. */
static int put_int(char *p, int v, int w) {
    char t[32];
    int n = snprintf(t, sizeof t, "%d", v);
    if (n > w) { memset(p, '*', w); return w; }
    memset(p, ' ', w - n); memcpy(p + w - n, t, n);
    return w;
}
int md_oracle_res_line(char *buf, double charge, int mchrg, int itrj, int icoll, int isec, int j, int ntypes, const int32_t *types,
                       const int32_t *counts) {
    int items[64], n = 0;
    items[n++] = mchrg; items[n++] = itrj;
    if (icoll >= 0) items[n++] = icoll;
    items[n++] = isec; items[n++] = j; items[n++] = ntypes;
    for (int m = 0; m < ntypes && n + 2 <= 64; ++m) { items[n++] = types[m]; items[n++] = counts[m]; }
    /* widths of the integer edit descriptors; negative = nX */
    int desc[8 + 40] = {3, 5, 5, 2, 2, -2, 3, -2};
    for (int k = 0; k < 20; ++k) { desc[8 + 2 * k] = 4; desc[9 + 2 * k] = 3; }
    char *p = buf;
    p += snprintf(p, 16, "%10.7f", charge);
    if (p - buf > 10) { memset(buf, '*', 10); p = buf + 10; }
    int it = 0, pad = 0;
    for (int d = 0; d < 48 && it < n; ++d) {
        if (desc[d] < 0) { pad += -desc[d]; continue; }
        memset(p, ' ', pad); p += pad; pad = 0;
        p += put_int(p, items[it++], desc[d]);
    }
    *p = 0;
    return (int)(p - buf);
}
