/* TEST INFRASTRUCTURE -- CPU restatement of the MD side of the QCxMS production-trajectory path.
 * Compiled with -ffp-contract=off so that every expression rounds like the reference's Fortran. */
#include "md_oracle.h"

#include <math.h>
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

/* reference src/md.f90:34-708 restricted to it > 0, method 0 (EI), icoll = 0, No_eTemp = .false. */
int md_oracle_md(const qcxms_b200_md_config_t *cfg, int nuc, const int32_t *iat, const double *mass, double *xyz, double *velo,
                 const double *velof, double eimp, double tadd, int step_limit, double *grad, int32_t *list, double *achrg, double *axyz,
                 qcxms_b200_md_result_t *res) {
    const double tstep = cfg->tstep;
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
    const int more = 250, avdump = 50;
    double Tav = 0, Epav = 0, Ekav = 0, Edum = 0, Eerror = 0, aTlast = 0, dtime = 0, ttime = 0, Eav;
    int nstep = 0, morestep = 0, fconst = 0, kdump = avdump;
    double *avchrg = calloc(nuc, sizeof(double)), *avxyz = calloc(3 * nuc, sizeof(double));
    int nadd = (int)((tadd + tstep) / tstep - 1);
    double fadd = tstep / (tadd + tstep);
    for (;;) {
        if (step_limit > 0 && nstep >= step_limit) break; /* bounded run requested by the caller (not in the reference) */
        nstep = nstep + 1;
        T = Ekin / (0.5 * 3 * nuc * QC_KB);
        Tav = Tav + T; Epav = Epav + Epot; Ekav = Ekav + Ekin;
        if (nstep > nadd) { Edum = Edum + Epot + Ekin; Eav = Edum / (double)(float)(nstep - nadd); }
        else Eav = Epot + Ekin;
        Eerror = Eav - Epot - Ekin;
        int err1 = Epot == 0, err2 = fabs(Eerror) > (double)0.1f;
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
        aTlast = aTlast + T;
        md_oracle_leapfrog(nuc, grad, mass, tstep, xyz, velo, &Ekin);
        ttime = ttime + tstep / QC_FSTOAU;
        md_oracle_egrad(nuc, xyz, iat, cfg->mchrg, etemp, cfg->method_id, &Epot, grad, achrg, &niter);
        scc_total += niter;
        kdump = kdump + 1;
        if (nfrag == 1) morestep = 0;
        if (nfrag > 1 && dtime < 1e-6) dtime = ttime / 1000.;
        if (nstep <= nadd && nfrag == 1) {
            if (md_oracle_impactscale(nuc, velo, mass, velof, eimp, fadd * nstep, Ekinstart)) { res->status = 2; break; }
        }
        if (cfg->etemp_in < 0) {
            double dum = eimp - eimp * (double)(float)nstep / (double)(float)nadd;
            etemp = md_oracle_setetemp(nfrag, dum, cfg->ax, cfg->ieetemp);
        }
        md_oracle_fragment_structure(nuc, iat, xyz, 3.0, 1, 0, list);
        md_oracle_fragmass(nuc, iat, list, mass, NULL, &nfrag, NULL, NULL);
        if (cfg->exit_rules) {
            if (nfrag > 6) break;
            if (nfrag > cfg->nfragexit) { fragstate = 1; mdok = 1; break; }
            if (nfrag >= 2) fconst = fconst + 1; else fconst = 0;
            if (fconst > 1000) { fragstate = 2; mdok = 1; break; }
            if (nfrag >= cfg->nfragexit) {
                morestep = morestep + 1;
                if (morestep > more) { fragstate = 1; mdok = 1; break; }
            }
        }
        if (nstep >= cfg->nmax) { fragstate = 1; mdok = 1; break; }
    }
    res->mdok = mdok; res->fragstate = fragstate; res->nstep = nstep; res->nfrag = nfrag;
    if (res->status == 0) res->status = 1;
    res->scc_iter_total = scc_total;
    res->Tav = Tav / nstep; res->Epav = Epav / nstep; res->Ekav = Ekav / nstep;
    for (int i = 0; i < nuc; ++i) achrg[i] = avchrg[i] / kdump;
    for (int i = 0; i < 3 * nuc; ++i) axyz[i] = avxyz[i] / kdump;
    res->aTlast = aTlast / kdump; res->dtime = dtime; res->ttime = ttime; res->Epot = Epot; res->Ekin = Ekin;
    free(avchrg); free(avxyz);
    return 0;
}
