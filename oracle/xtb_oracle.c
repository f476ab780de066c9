/* TEST INFRASTRUCTURE -- see xtb_oracle.h.  PARITY UNPINNED (no tblite on this machine).
 *
 * CPU restatement of the energy/gradient entry point of QCxMS,
 *     get_xtb_egrad  (reference src/tblite.f90:65-175)
 * for method id 2 (GFN2-xTB).  The call protocol follows the reference line by line:
 *   - new structure, charge = real(charge), uhf = min(multiplicity-1, 0)   (tblite.f90:111)
 *   - new calculator + NEW ZEROED wavefunction on every call, nspin = 1,
 *     kt = etemp * 3.166808578545117e-06                                   (tblite.f90:43,123,133)
 *   - xtb_singlepoint(..., accuracy = 1.0, ...)                            (tblite.f90:46,136)
 *   - qat = wfn%qat(:,1); failure -> stat = -1                             (tblite.f90:140-151)
 * What happens inside xtb_singlepoint is tblite v0.2.1 (external, un-vendored:
 * subprojects/tblite.wrap:1-4, dftd4 v3.4.0, mctc-lib v0.3.0); it is restated here from
 * the published GFN2-xTB / D4 method definitions: SURVEY.md 3.2 + Appendix B.  The D4
 * reference data and model follow the in-tree statement of the same model
 * (reference src/dftd4.f90:413-449 zeta, :451-497 quadrature, :524-595 reference
 * polarisabilities, :597-670 Gaussian weights; include/param_d4.fh), with the erf
 * covalent CN that dftd4 v3 uses.
 */
#include "xtb_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../qcxms_b200/csrc/params/constants.h"
#include "../qcxms_b200/csrc/params/d4_refdata.h"
#include "../qcxms_b200/csrc/params/elem_tables.h"
#include "../qcxms_b200/csrc/params/gfn2_params.h"
#include "../qcxms_b200/csrc/params/gfn1_params.h"
#include "../qcxms_b200/csrc/params/d3_refdata.h"
#include "../qcxms_b200/csrc/params/stong_table.h"

#define MAXPRIM 8 /* GFN1: the orthogonalised H 2s carries the 1s primitives as well (3 + 4) */
#define MAX_ITER 250
#define SQRT3 1.7320508075688772935
#define PI 3.14159265358979323846264338327950288

/* ------------------------------------------------------------------ basis --- */
static const int CART_EXP[3][6][3] = {
    {{0, 0, 0}},
    {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}},
    {{2, 0, 0}, {0, 2, 0}, {0, 0, 2}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}}};
static const int NCART[3] = {1, 3, 6};
static const int NSPH[3] = {1, 3, 5};
/* real solid harmonics in tblite order m = -l..l */
static const double TRAFO0[1][6] = {{1, 0, 0, 0, 0, 0}};
static const double TRAFO1[3][6] = {{0, 1, 0, 0, 0, 0}, {0, 0, 1, 0, 0, 0}, {1, 0, 0, 0, 0, 0}};
static const double TRAFO2[5][6] = {{0, 0, 0, SQRT3, 0, 0},
                                    {0, 0, 0, 0, 0, SQRT3},
                                    {-0.5, -0.5, 1.0, 0, 0, 0},
                                    {0, 0, 0, 0, SQRT3, 0},
                                    {0.5 * SQRT3, -0.5 * SQRT3, 0, 0, 0, 0}};
static const double *trafo_row(int l, int m) {
    return l == 0 ? TRAFO0[m] : (l == 1 ? TRAFO1[m] : TRAFO2[m]);
}

typedef struct {
    int nat, nsh, nao, charge;
    int method;                 /* QC_METHOD_GFN2 or QC_METHOD_GFN1 */
    const gfn2_elem_t *elem;    /* element table of the method */
    int *sh_val;                /* [nsh] GFN1: 0 marks a diffuse shell */
    double *at_gam3;            /* [nat] GFN1: atom-resolved third-order parameter */
    double kt;
    const int32_t *num;
    const double *xyz;
    int *at_sh0, *at_nsh, *sh_at, *sh_l, *sh_ao0, *sh_np, *ao_at, *ao_sh;
    double *sh_alpha, *sh_coef; /* [nsh][MAXPRIM] */
    double *sh_level, *sh_kcn, *sh_poly, *sh_refocc, *sh_hub, *sh_gam3, *sh_zeta;
    /* geometry dependent */
    double *cn, *dcndr;     /* dcndr[(k*nat+i)*3+c] = d cn_i / d R_k,c */
    double *cnd4, *dcnd4dr;
    double *S, *H0, *D, *Q; /* D [3][nao*nao], Q [6][nao*nao] */
    double *selfen;         /* [nsh] CN-dependent level */
} sys_t;

static void *xcalloc(size_t n, size_t sz) {
    void *p = calloc(n ? n : 1, sz);
    if (!p) {
        fprintf(stderr, "xtb_oracle: out of memory\n");
        abort();
    }
    return p;
}

static const stong_entry_t *stong_find(int n, int l, int ng) {
    for (int i = 0; i < STONG_NENTRY; ++i)
        if (STONG_TABLE[i].n == n && STONG_TABLE[i].l == l && STONG_TABLE[i].ng == ng) return &STONG_TABLE[i];
    return NULL;
}

int xtb_oracle_dims(int nat, const int32_t *num, int method_id, int *nsh, int *nao) {
    if (method_id != QC_METHOD_GFN2 && method_id != QC_METHOD_GFN1) return QC_STAT_UNKNOWN_METHOD;
    if (method_id == QC_METHOD_GFN1) gfn1_ensure_loaded();
    int ns = 0, na = 0;
    for (int i = 0; i < nat; ++i) {
        if (num[i] < 1 || num[i] > GFN2_MAXZ) return QC_STAT_FATAL;
        if (method_id == QC_METHOD_GFN1 && !GFN1_EXTRA[num[i]].supported) return QC_STAT_FATAL;
        const gfn2_elem_t *e = method_id == QC_METHOD_GFN1 ? &GFN1_ELEM[num[i]] : &GFN2_ELEM[num[i]];
        ns += e->nshell;
        for (int k = 0; k < e->nshell; ++k) na += NSPH[e->ang[k]];
    }
    *nsh = ns;
    *nao = na;
    return 0;
}

static int setup_basis(sys_t *s) {
    int nat = s->nat;
    if (xtb_oracle_dims(nat, s->num, s->method, &s->nsh, &s->nao)) return -1;
    const int gfn1 = s->method == QC_METHOD_GFN1;
    int nsh = s->nsh, nao = s->nao;
    s->at_sh0 = xcalloc(nat, sizeof(int));
    s->at_nsh = xcalloc(nat, sizeof(int));
    s->sh_at = xcalloc(nsh, sizeof(int));
    s->sh_l = xcalloc(nsh, sizeof(int));
    s->sh_ao0 = xcalloc(nsh, sizeof(int));
    s->sh_np = xcalloc(nsh, sizeof(int));
    s->ao_at = xcalloc(nao, sizeof(int));
    s->ao_sh = xcalloc(nao, sizeof(int));
    s->sh_alpha = xcalloc((size_t)nsh * MAXPRIM, sizeof(double));
    s->sh_coef = xcalloc((size_t)nsh * MAXPRIM, sizeof(double));
    s->sh_level = xcalloc(nsh, sizeof(double));
    s->sh_kcn = xcalloc(nsh, sizeof(double));
    s->sh_poly = xcalloc(nsh, sizeof(double));
    s->sh_refocc = xcalloc(nsh, sizeof(double));
    s->sh_hub = xcalloc(nsh, sizeof(double));
    s->sh_gam3 = xcalloc(nsh, sizeof(double));
    s->sh_zeta = xcalloc(nsh, sizeof(double));
    s->sh_val = xcalloc(nsh, sizeof(int));
    s->at_gam3 = xcalloc(nat, sizeof(double));
    int ish = 0, iao = 0;
    for (int i = 0; i < nat; ++i) {
        const gfn2_elem_t *e = &s->elem[s->num[i]];
        if (gfn1) s->at_gam3[i] = e->hubbard_deriv;
        s->at_sh0[i] = ish;
        s->at_nsh[i] = e->nshell;
        for (int k = 0; k < e->nshell; ++k, ++ish) {
            int l = e->ang[k];
            s->sh_at[ish] = i;
            s->sh_l[ish] = l;
            s->sh_ao0[ish] = iao;
            s->sh_np[ish] = e->nprim[k];
            s->sh_level[ish] = e->selfenergy[k] * GFN2_EVTOAU;
            s->sh_kcn[ish] = e->kcn[k] * GFN2_EVTOAU;
            s->sh_val[ish] = 1;
            if (gfn1) {   /* h = level (1 + kcn_l CN) written as level - kcn CN */
                s->sh_kcn[ish] = -e->selfenergy[k] * GFN2_EVTOAU * GFN1_KCN_L[l];
                s->sh_val[ish] = GFN1_EXTRA[s->num[i]].valence[k];
            }
            s->sh_poly[ish] = e->shpoly[k];
            s->sh_refocc[ish] = e->refocc[k];
            s->sh_hub[ish] = e->hubbard * e->shell_hubbard[l];
            s->sh_gam3[ish] = gfn1 ? 0.0 : e->hubbard_deriv * GFN2_KSHELL3[l];
            s->sh_zeta[ish] = e->slater[k];
            const stong_entry_t *t = stong_find(e->pqn[k], l, e->nprim[k]);
            if (!t) return -1;
            double dfact = (l == 2) ? 3.0 : 1.0;
            for (int p = 0; p < e->nprim[k]; ++p) {
                double a = t->alpha[p] * e->slater[k] * e->slater[k];
                s->sh_alpha[ish * MAXPRIM + p] = a;
                s->sh_coef[ish * MAXPRIM + p] =
                    t->coeff[p] * pow(2.0 * a / PI, 0.75) * pow(sqrt(4.0 * a), l) / sqrt(dfact);
            }
            for (int m = 0; m < NSPH[l]; ++m, ++iao) {
                s->ao_at[iao] = i;
                s->ao_sh[iao] = ish;
            }
            /* GFN1: a second shell of the same angular momentum (H 2s) is Schmidt-orthogonalised to the first one and renormalised:
             * chi' = (chi_2 - <1|2>/<1|1> chi_1) / |...|, so it carries the primitives of both (s functions only) */
            if (gfn1 && k > 0 && l == 0) {
                int first = -1;
                for (int kk = 0; kk < k; ++kk) if (e->ang[kk] == l) first = s->at_sh0[i] + kk;
                if (first >= 0) {
                    double *a1 = s->sh_alpha + first * MAXPRIM, *c1 = s->sh_coef + first * MAXPRIM;
                    double *a2 = s->sh_alpha + ish * MAXPRIM, *c2 = s->sh_coef + ish * MAXPRIM;
                    int n1 = s->sh_np[first], n2 = s->sh_np[ish];
                    if (n1 + n2 > MAXPRIM) return -1;
                    double s11 = 0, s12 = 0, s22 = 0;
                    for (int p = 0; p < n1; ++p) for (int q = 0; q < n1; ++q) s11 += c1[p] * c1[q] * pow(PI / (a1[p] + a1[q]), 1.5);
                    for (int p = 0; p < n1; ++p) for (int q = 0; q < n2; ++q) s12 += c1[p] * c2[q] * pow(PI / (a1[p] + a2[q]), 1.5);
                    for (int p = 0; p < n2; ++p) for (int q = 0; q < n2; ++q) s22 += c2[p] * c2[q] * pow(PI / (a2[p] + a2[q]), 1.5);
                    double f = s12 / s11, nrm = 1.0 / sqrt(s22 - s12 * s12 / s11);
                    for (int q = 0; q < n2; ++q) c2[q] *= nrm;
                    for (int p = 0; p < n1; ++p) { a2[n2 + p] = a1[p]; c2[n2 + p] = -f * nrm * c1[p]; }
                    s->sh_np[ish] = n1 + n2;
                }
            }
        }
    }
    return 0;
}

/* ---------------------------------------------------- coordination numbers --- */
static double d3_rcov(int z) { return 4.0 / 3.0 * COVRAD2009_AA[z] * TB_AATOAU; }

/* GFN2 double-exponential counting function: ka = 10, kb = 20, r_shift = 2, cutoff 25 bohr */
/* GFN1 / D3: exponential counting function, k1 = 16, covalent radii x 4/3 (reference src/dftd3.f90:607-642 ncoord, cn_thr = 1000 bohr^2) */
static void exp_cn(sys_t *s) {
    int nat = s->nat;
    const double k1 = 16.0, cn_thr = 1000.0;
    memset(s->cn, 0, nat * sizeof(double));
    memset(s->dcndr, 0, (size_t)nat * nat * 3 * sizeof(double));
    for (int i = 0; i < nat; ++i)
        for (int j = 0; j < i; ++j) {
            double v[3] = {s->xyz[3 * i] - s->xyz[3 * j], s->xyz[3 * i + 1] - s->xyz[3 * j + 1],
                           s->xyz[3 * i + 2] - s->xyz[3 * j + 2]};
            double r2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
            if (r2 > cn_thr) continue;
            double r = sqrt(r2), rco = D3_RCOV[s->num[i]] + D3_RCOV[s->num[j]];
            double ex = exp(-k1 * (rco / r - 1.0));
            double f = 1.0 / (1.0 + ex);
            double df = -k1 * rco / r2 * ex * f * f;
            s->cn[i] += f;
            s->cn[j] += f;
            for (int c = 0; c < 3; ++c) {
                double g = df * v[c] / r;
                s->dcndr[(i * nat + i) * 3 + c] += g;
                s->dcndr[(j * nat + i) * 3 + c] -= g;
                s->dcndr[(j * nat + j) * 3 + c] -= g;
                s->dcndr[(i * nat + j) * 3 + c] += g;
            }
        }
}

static void gfn_cn(sys_t *s) {
    int nat = s->nat;
    const double ka = 10.0, kb = 20.0, rshift = 2.0, cutoff2 = 25.0 * 25.0;
    memset(s->cn, 0, nat * sizeof(double));
    memset(s->dcndr, 0, (size_t)nat * nat * 3 * sizeof(double));
    for (int i = 0; i < nat; ++i)
        for (int j = 0; j < i; ++j) {
            double v[3] = {s->xyz[3 * i] - s->xyz[3 * j], s->xyz[3 * i + 1] - s->xyz[3 * j + 1],
                           s->xyz[3 * i + 2] - s->xyz[3 * j + 2]};
            double r2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
            if (r2 > cutoff2) continue;
            double r = sqrt(r2), rc = d3_rcov(s->num[i]) + d3_rcov(s->num[j]);
            double ea = exp(-ka * (rc / r - 1.0)), eb = exp(-kb * ((rc + rshift) / r - 1.0));
            double fa = 1.0 / (1.0 + ea), fb = 1.0 / (1.0 + eb);
            double dfa = -ka * rc / r2 * ea * fa * fa, dfb = -kb * (rc + rshift) / r2 * eb * fb * fb;
            double f = fa * fb, df = dfa * fb + fa * dfb;
            s->cn[i] += f;
            s->cn[j] += f;
            for (int c = 0; c < 3; ++c) {
                double g = df * v[c] / r;
                s->dcndr[(i * nat + i) * 3 + c] += g;
                s->dcndr[(j * nat + i) * 3 + c] -= g;
                s->dcndr[(j * nat + j) * 3 + c] -= g;
                s->dcndr[(i * nat + j) * 3 + c] += g;
            }
        }
}

/* D4 covalent CN: erf counting (kcn = 7.5) with EN-dependent bond-order factor */
static void d4_cn(sys_t *s) {
    int nat = s->nat;
    const double kcn = 7.5, k4 = 4.10451, k5 = 19.08857, k6 = 2.0 * 11.28174 * 11.28174, cutoff2 = 30.0 * 30.0;
    memset(s->cnd4, 0, nat * sizeof(double));
    memset(s->dcnd4dr, 0, (size_t)nat * nat * 3 * sizeof(double));
    for (int i = 0; i < nat; ++i)
        for (int j = 0; j < i; ++j) {
            double v[3] = {s->xyz[3 * i] - s->xyz[3 * j], s->xyz[3 * i + 1] - s->xyz[3 * j + 1],
                           s->xyz[3 * i + 2] - s->xyz[3 * j + 2]};
            double r2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
            if (r2 > cutoff2) continue;
            double r = sqrt(r2), rc = d3_rcov(s->num[i]) + d3_rcov(s->num[j]);
            double den = k4 * exp(-pow(fabs(PAULING_EN[s->num[i]] - PAULING_EN[s->num[j]]) + k5, 2) / k6);
            double arg = kcn * (r / rc - 1.0);
            double f = den * 0.5 * (1.0 + erf(-arg));
            double df = -den * kcn / rc / sqrt(PI) * exp(-arg * arg);
            s->cnd4[i] += f;
            s->cnd4[j] += f;
            for (int c = 0; c < 3; ++c) {
                double g = df * v[c] / r;
                s->dcnd4dr[(i * nat + i) * 3 + c] += g;
                s->dcnd4dr[(j * nat + i) * 3 + c] -= g;
                s->dcnd4dr[(j * nat + j) * 3 + c] -= g;
                s->dcnd4dr[(i * nat + j) * 3 + c] += g;
            }
        }
}

/* ---------------------------------------------------------------- integrals --- */
/* <J| O_I |I> for one contracted shell pair in the real-spherical basis.
 * vec = R_I - R_J; the multipole operators are centred on I (the ket).  Outputs are
 * [mj][mi] blocks (leading dimension nI).  Quadrupoles come back in the traceless
 * form 3/2 r_a r_b - 1/2 r^2 delta_ab, order xx,xy,yy,xz,yz,zz.  With want_deriv the
 * derivatives with respect to vec are returned as well. */
typedef struct {
    double S[25], D[3][25], Q[6][25];
    double dS[3][25], dD[3][3][25], dQ[3][6][25]; /* first index: derivative direction */
} pair_ints_t;

static void shell_pair(const sys_t *s, int jsh, int ish, const double vec[3], double r2, int want_deriv,
                       pair_ints_t *out) {
    const double intcut = 25.0; /* integral_cutoff(acc = 1) */
    int lj = s->sh_l[jsh], li = s->sh_l[ish];
    int ncj = NCART[lj], nci = NCART[li];
    double cS[36], cD[3][36], cQ[6][36], cdS[3][36], cdD[3][3][36], cdQ[3][6][36];
    memset(cS, 0, sizeof cS);
    memset(cD, 0, sizeof cD);
    memset(cQ, 0, sizeof cQ);
    if (want_deriv) {
        memset(cdS, 0, sizeof cdS);
        memset(cdD, 0, sizeof cdD);
        memset(cdQ, 0, sizeof cdQ);
    }
    for (int pj = 0; pj < s->sh_np[jsh]; ++pj) {
        double aj = s->sh_alpha[jsh * MAXPRIM + pj], cj = s->sh_coef[jsh * MAXPRIM + pj];
        for (int pi = 0; pi < s->sh_np[ish]; ++pi) {
            double ai = s->sh_alpha[ish * MAXPRIM + pi], ci = s->sh_coef[ish * MAXPRIM + pi];
            double gam = ai + aj, est = ai * aj * r2 / gam;
            if (est > intcut) continue;
            double pre = exp(-est) * pow(PI / gam, 1.5) * ci * cj;
            double oog = 0.5 / gam;
            /* 1-D tables t[d][a][b]: a on J up to lj+1, b on I up to li+2 */
            double t[3][4][5];
            for (int d = 0; d < 3; ++d) {
                double pa = ai / gam * vec[d];  /* P - R_J */
                double pb = -aj / gam * vec[d]; /* P - R_I */
                int amax = lj + 1, bmax = li + 2;
                t[d][0][0] = 1.0;
                for (int b = 0; b < bmax; ++b)
                    t[d][0][b + 1] = pb * t[d][0][b] + (b > 0 ? b * oog * t[d][0][b - 1] : 0.0);
                for (int a = 0; a < amax; ++a)
                    for (int b = 0; b <= bmax; ++b)
                        t[d][a + 1][b] = pa * t[d][a][b] + (a > 0 ? a * oog * t[d][a - 1][b] : 0.0) +
                                         (b > 0 ? b * oog * t[d][a][b - 1] : 0.0);
            }
            for (int kj = 0; kj < ncj; ++kj) {
                const int *ea = CART_EXP[lj][kj];
                for (int ki = 0; ki < nci; ++ki) {
                    const int *eb = CART_EXP[li][ki];
                    int ij = kj * nci + ki;
                    /* per-dimension factors with 0,1,2 extra powers on the ket */
                    double f[3][3];
                    for (int d = 0; d < 3; ++d)
                        for (int m = 0; m < 3; ++m) f[d][m] = t[d][ea[d]][eb[d] + m];
                    cS[ij] += pre * f[0][0] * f[1][0] * f[2][0];
                    cD[0][ij] += pre * f[0][1] * f[1][0] * f[2][0];
                    cD[1][ij] += pre * f[0][0] * f[1][1] * f[2][0];
                    cD[2][ij] += pre * f[0][0] * f[1][0] * f[2][1];
                    cQ[0][ij] += pre * f[0][2] * f[1][0] * f[2][0];
                    cQ[1][ij] += pre * f[0][1] * f[1][1] * f[2][0];
                    cQ[2][ij] += pre * f[0][0] * f[1][2] * f[2][0];
                    cQ[3][ij] += pre * f[0][1] * f[1][0] * f[2][1];
                    cQ[4][ij] += pre * f[0][0] * f[1][1] * f[2][1];
                    cQ[5][ij] += pre * f[0][0] * f[1][0] * f[2][2];
                    if (!want_deriv) continue;
                    /* d/dvec_k = -d/dR_J,k acting on the bra: -(2 aj t[a+1] - a t[a-1]) */
                    for (int k = 0; k < 3; ++k) {
                        double g[3][3];
                        for (int d = 0; d < 3; ++d)
                            for (int m = 0; m < 3; ++m) {
                                if (d == k) {
                                    double up = t[d][ea[d] + 1][eb[d] + m];
                                    double dn = ea[d] > 0 ? t[d][ea[d] - 1][eb[d] + m] : 0.0;
                                    g[d][m] = -(2.0 * aj * up - ea[d] * dn);
                                } else
                                    g[d][m] = f[d][m];
                            }
                        cdS[k][ij] += pre * g[0][0] * g[1][0] * g[2][0];
                        cdD[k][0][ij] += pre * g[0][1] * g[1][0] * g[2][0];
                        cdD[k][1][ij] += pre * g[0][0] * g[1][1] * g[2][0];
                        cdD[k][2][ij] += pre * g[0][0] * g[1][0] * g[2][1];
                        cdQ[k][0][ij] += pre * g[0][2] * g[1][0] * g[2][0];
                        cdQ[k][1][ij] += pre * g[0][1] * g[1][1] * g[2][0];
                        cdQ[k][2][ij] += pre * g[0][0] * g[1][2] * g[2][0];
                        cdQ[k][3][ij] += pre * g[0][1] * g[1][0] * g[2][1];
                        cdQ[k][4][ij] += pre * g[0][0] * g[1][1] * g[2][1];
                        cdQ[k][5][ij] += pre * g[0][0] * g[1][0] * g[2][2];
                    }
                }
            }
        }
    }
    /* cartesian -> spherical on both sides, then make quadrupoles traceless */
    int nsj = NSPH[lj], nsi = NSPH[li];
    int nblock = 10 + (want_deriv ? 30 : 0);
    for (int blk = 0; blk < nblock; ++blk) {
        const double *src;
        double *dst;
        if (blk == 0) { src = cS; dst = out->S; }
        else if (blk < 4) { src = cD[blk - 1]; dst = out->D[blk - 1]; }
        else if (blk < 10) { src = cQ[blk - 4]; dst = out->Q[blk - 4]; }
        else {
            int k = (blk - 10) / 10, c = (blk - 10) % 10;
            if (c == 0) { src = cdS[k]; dst = out->dS[k]; }
            else if (c < 4) { src = cdD[k][c - 1]; dst = out->dD[k][c - 1]; }
            else { src = cdQ[k][c - 4]; dst = out->dQ[k][c - 4]; }
        }
        for (int mj = 0; mj < nsj; ++mj) {
            const double *tj = trafo_row(lj, mj);
            for (int mi = 0; mi < nsi; ++mi) {
                const double *ti = trafo_row(li, mi);
                double acc = 0.0;
                for (int kj = 0; kj < ncj; ++kj) {
                    if (tj[kj] == 0.0) continue;
                    for (int ki = 0; ki < nci; ++ki) acc += tj[kj] * ti[ki] * src[kj * nci + ki];
                }
                dst[mj * nsi + mi] = acc;
            }
        }
    }
    int n = nsj * nsi;
    for (int ij = 0; ij < n; ++ij) {
        double tr = 0.5 * (out->Q[0][ij] + out->Q[2][ij] + out->Q[5][ij]);
        for (int c = 0; c < 6; ++c) out->Q[c][ij] *= 1.5;
        out->Q[0][ij] -= tr; out->Q[2][ij] -= tr; out->Q[5][ij] -= tr;
        if (want_deriv)
            for (int k = 0; k < 3; ++k) {
                double dtr = 0.5 * (out->dQ[k][0][ij] + out->dQ[k][2][ij] + out->dQ[k][5][ij]);
                for (int c = 0; c < 6; ++c) out->dQ[k][c][ij] *= 1.5;
                out->dQ[k][0][ij] -= dtr; out->dQ[k][2][ij] -= dtr; out->dQ[k][5][ij] -= dtr;
            }
    }
}

/* Move the multipole operator from centre I to centre J = I - vec (i.e. r - R_J = (r - R_I) + vec):
 * given s, d, q (traceless) about I returns dj, qj about J; with derivatives if ds != NULL. */
static void shift_operator(const double vec[3], double s, const double di[3], const double qi[6], double dj[3],
                           double qj[6], const double ds[3], double ddi[3][3], double dqi[3][6],
                           double ddj[3][3], double dqj[3][6]) {
    /* cartesian (non-traceless) second-moment shift: <(x+vx)(y+vy)> = <xy> + vx<y> + vy<x> + vx vy s.
       In the traceless form the correction c_ab = vec_a d_b + vec_b d_a + vec_a vec_b s enters as
       3/2 c_ab - 1/2 tr(c) delta_ab. */
    static const int A[6] = {0, 0, 1, 0, 1, 2}, B[6] = {0, 1, 1, 2, 2, 2};
    for (int c = 0; c < 3; ++c) dj[c] = di[c] + vec[c] * s;
    double corr[6], tr;
    for (int c = 0; c < 6; ++c) corr[c] = vec[A[c]] * di[B[c]] + vec[B[c]] * di[A[c]] + vec[A[c]] * vec[B[c]] * s;
    tr = 0.5 * (corr[0] + corr[2] + corr[5]);
    for (int c = 0; c < 6; ++c) qj[c] = qi[c] + 1.5 * corr[c];
    qj[0] -= tr; qj[2] -= tr; qj[5] -= tr;
    if (!ds) return;
    for (int k = 0; k < 3; ++k) {
        for (int c = 0; c < 3; ++c) ddj[k][c] = ddi[k][c] + vec[c] * ds[k] + (c == k ? s : 0.0);
        double dc[6];
        for (int c = 0; c < 6; ++c) {
            int a = A[c], b = B[c];
            dc[c] = vec[a] * ddi[k][b] + vec[b] * ddi[k][a] + vec[a] * vec[b] * ds[k] +
                    (a == k ? di[b] + vec[b] * s : 0.0) + (b == k ? di[a] + vec[a] * s : 0.0);
        }
        double dtr = 0.5 * (dc[0] + dc[2] + dc[5]);
        for (int c = 0; c < 6; ++c) dqj[k][c] = dqi[k][c] + 1.5 * dc[c];
        dqj[k][0] -= dtr; dqj[k][2] -= dtr; dqj[k][5] -= dtr;
    }
}

/* GFN2 shell-pair scaling of H0: zeta-weighting * K_ll' * (1 + enscale dEN^2) */
static double gfn2_hscale(const sys_t *s, int ish, int jsh) {
    static const double kdiag[3] = {GFN2_KDIAG_S, GFN2_KDIAG_P, GFN2_KDIAG_D};
    int li = s->sh_l[ish], lj = s->sh_l[jsh];
    if (s->method == QC_METHOD_GFN1) {
        static const double k1[3] = {GFN1_KDIAG_S, GFN1_KDIAG_P, GFN1_KDIAG_D};
        int zi = s->num[s->sh_at[ish]], zj = s->num[s->sh_at[jsh]];
        if (s->sh_val[ish] && s->sh_val[jsh]) {
            double kll = (li + lj == 1) ? GFN1_K_SP : 0.5 * (k1[li] + k1[lj]);
            double den = s->elem[zi].en - s->elem[zj].en;
            return gfn1_kpair(zi, zj) * kll * (1.0 + GFN1_ENSCALE * den * den);
        }
        if (s->sh_val[ish]) return 0.5 * (k1[li] + GFN1_KDIFF);
        if (s->sh_val[jsh]) return 0.5 * (k1[lj] + GFN1_KDIFF);
        return GFN1_KDIFF;
    }
    double k;
    if (li == lj) k = kdiag[li];
    else if (li == 2 || lj == 2) k = (li + lj == 2) ? GFN2_K_SD : GFN2_K_PD;
    else k = 0.5 * (kdiag[li] + kdiag[lj]);
    double zi = s->sh_zeta[ish], zj = s->sh_zeta[jsh];
    double zij = pow(2.0 * sqrt(zi * zj) / (zi + zj), GFN2_WEXP);
    double den = s->elem[s->num[s->sh_at[ish]]].en - s->elem[s->num[s->sh_at[jsh]]].en;
    return zij * k * (1.0 + GFN2_ENSCALE * den * den);
}

static void build_integrals(sys_t *s) {
    int nat = s->nat, nao = s->nao;
    size_t n2 = (size_t)nao * nao;
    memset(s->S, 0, n2 * sizeof(double));
    memset(s->H0, 0, n2 * sizeof(double));
    memset(s->D, 0, 3 * n2 * sizeof(double));
    memset(s->Q, 0, 6 * n2 * sizeof(double));
    for (int ish = 0; ish < s->nsh; ++ish) s->selfen[ish] = s->sh_level[ish] - s->sh_kcn[ish] * s->cn[s->sh_at[ish]];
    pair_ints_t pi;
    for (int iat = 0; iat < nat; ++iat)
        for (int jat = 0; jat <= iat; ++jat) {
            double vec[3] = {s->xyz[3 * iat] - s->xyz[3 * jat], s->xyz[3 * iat + 1] - s->xyz[3 * jat + 1],
                             s->xyz[3 * iat + 2] - s->xyz[3 * jat + 2]};
            double r2 = vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2];
            double radsum = (s->elem[s->num[iat]].atomic_rad + s->elem[s->num[jat]].atomic_rad) * TB_AATOAU;
            double rr = sqrt(sqrt(r2) / radsum);
            for (int is = 0; is < s->at_nsh[iat]; ++is)
                for (int js = 0; js < s->at_nsh[jat]; ++js) {
                    int ish = s->at_sh0[iat] + is, jsh = s->at_sh0[jat] + js;
                    shell_pair(s, jsh, ish, vec, r2, 0, &pi);
                    double hij;
                    if (iat == jat)
                        hij = 0.5 * (s->selfen[ish] + s->selfen[jsh]);
                    else
                        hij = 0.5 * (s->selfen[ish] + s->selfen[jsh]) * gfn2_hscale(s, ish, jsh) *
                              (1.0 + s->sh_poly[ish] * rr) * (1.0 + s->sh_poly[jsh] * rr);
                    int ni = NSPH[s->sh_l[ish]], nj = NSPH[s->sh_l[jsh]];
                    for (int mj = 0; mj < nj; ++mj)
                        for (int mi = 0; mi < ni; ++mi) {
                            int ij = mj * ni + mi;
                            size_t a = s->sh_ao0[jsh] + mj, b = s->sh_ao0[ish] + mi;
                            double di[3] = {pi.D[0][ij], pi.D[1][ij], pi.D[2][ij]}, qi[6], dj[3], qj[6];
                            for (int c = 0; c < 6; ++c) qi[c] = pi.Q[c][ij];
                            s->S[a * nao + b] = pi.S[ij];
                            s->H0[a * nao + b] = pi.S[ij] * hij;
                            for (int c = 0; c < 3; ++c) s->D[c * n2 + a * nao + b] = di[c];
                            for (int c = 0; c < 6; ++c) s->Q[c * n2 + a * nao + b] = qi[c];
                            if (iat != jat) {
                                shift_operator(vec, pi.S[ij], di, qi, dj, qj, NULL, NULL, NULL, NULL, NULL);
                                s->S[b * nao + a] = pi.S[ij];
                                s->H0[b * nao + a] = pi.S[ij] * hij;
                                for (int c = 0; c < 3; ++c) s->D[c * n2 + b * nao + a] = dj[c];
                                for (int c = 0; c < 6; ++c) s->Q[c * n2 + b * nao + a] = qj[c];
                            }
                        }
                }
        }
}

/* ------------------------------------------------------------ linear algebra --- */
/* Householder tridiagonalisation + implicit-shift QL; a (row-major, symmetric) is
 * overwritten by the eigenvectors: a[i*n+k] = i-th component of eigenvector k. */
int xtb_oracle_syev(int n, double *a, double *w) {
    double *e = xcalloc(n, sizeof(double));
    double *d = w;
#define A(i, j) a[(size_t)(i)*n + (j)]
    for (int i = n - 1; i > 0; --i) {
        int l = i - 1;
        double h = 0.0, scale = 0.0;
        if (l > 0) {
            for (int k = 0; k <= l; ++k) scale += fabs(A(i, k));
            if (scale == 0.0)
                e[i] = A(i, l);
            else {
                for (int k = 0; k <= l; ++k) {
                    A(i, k) /= scale;
                    h += A(i, k) * A(i, k);
                }
                double f = A(i, l);
                double g = f >= 0.0 ? -sqrt(h) : sqrt(h);
                e[i] = scale * g;
                h -= f * g;
                A(i, l) = f - g;
                f = 0.0;
                for (int j = 0; j <= l; ++j) {
                    A(j, i) = A(i, j) / h;
                    g = 0.0;
                    for (int k = 0; k <= j; ++k) g += A(j, k) * A(i, k);
                    for (int k = j + 1; k <= l; ++k) g += A(k, j) * A(i, k);
                    e[j] = g / h;
                    f += e[j] * A(i, j);
                }
                double hh = f / (h + h);
                for (int j = 0; j <= l; ++j) {
                    f = A(i, j);
                    e[j] = g = e[j] - hh * f;
                    for (int k = 0; k <= j; ++k) A(j, k) -= f * e[k] + g * A(i, k);
                }
            }
        } else
            e[i] = A(i, l);
        d[i] = h;
    }
    d[0] = 0.0;
    e[0] = 0.0;
    for (int i = 0; i < n; ++i) {
        int l = i - 1;
        if (d[i] != 0.0) {
            for (int j = 0; j <= l; ++j) {
                double g = 0.0;
                for (int k = 0; k <= l; ++k) g += A(i, k) * A(k, j);
                for (int k = 0; k <= l; ++k) A(k, j) -= g * A(k, i);
            }
        }
        d[i] = A(i, i);
        A(i, i) = 1.0;
        for (int j = 0; j <= l; ++j) A(j, i) = A(i, j) = 0.0;
    }
    /* QL with implicit shifts */
    for (int i = 1; i < n; ++i) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    int status = 0;
    for (int l = 0; l < n; ++l) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; ++m) {
                double dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) <= 2.3e-16 * dd) break;
            }
            if (m != l) {
                if (iter++ == 120) { status = -1; break; }
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? fabs(r) : -fabs(r)));
                double sn = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; --i) {
                    double f = sn * e[i], b = c * e[i];
                    e[i + 1] = (r = hypot(f, g));
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[m] = 0.0;
                        break;
                    }
                    sn = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * sn + 2.0 * c * b;
                    d[i + 1] = g + (p = sn * r);
                    g = c * r - b;
                    for (int k = 0; k < n; ++k) {
                        f = A(k, i + 1);
                        A(k, i + 1) = sn * A(k, i) + c * f;
                        A(k, i) = c * A(k, i) - sn * f;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    /* sort ascending */
    for (int i = 0; i < n - 1; ++i) {
        int k = i;
        double p = d[i];
        for (int j = i + 1; j < n; ++j)
            if (d[j] < p) { k = j; p = d[j]; }
        if (k != i) {
            d[k] = d[i];
            d[i] = p;
            for (int j = 0; j < n; ++j) {
                double t = A(j, i);
                A(j, i) = A(j, k);
                A(j, k) = t;
            }
        }
    }
#undef A
    free(e);
    return status;
}

/* generalised problem H C = S C eps by Cholesky reduction (what LAPACK dsygvd does).
 * Linv (lower-triangular inverse of the Cholesky factor of S) is computed once per geometry. */
static int cholesky_inverse(int n, const double *S, double *Linv) {
    double *L = xcalloc((size_t)n * n, sizeof(double));
    for (int j = 0; j < n; ++j) {
        double d = S[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) d -= L[(size_t)j * n + k] * L[(size_t)j * n + k];
        if (d <= 0.0) { free(L); return -1; }
        d = sqrt(d);
        L[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double v = S[(size_t)i * n + j];
            for (int k = 0; k < j; ++k) v -= L[(size_t)i * n + k] * L[(size_t)j * n + k];
            L[(size_t)i * n + j] = v / d;
        }
    }
    memset(Linv, 0, (size_t)n * n * sizeof(double));
    for (int j = 0; j < n; ++j) {
        Linv[(size_t)j * n + j] = 1.0 / L[(size_t)j * n + j];
        for (int i = j + 1; i < n; ++i) {
            double v = 0.0;
            for (int k = j; k < i; ++k) v -= L[(size_t)i * n + k] * Linv[(size_t)k * n + j];
            Linv[(size_t)i * n + j] = v / L[(size_t)i * n + i];
        }
    }
    free(L);
    return 0;
}

static int solve_gen(int n, const double *H, const double *Linv, double *C, double *eps, double *work) {
    /* work: T = Linv * H ; A = T * Linv^T */
    double *T = work, *A = work + (size_t)n * n;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double v = 0.0;
            for (int k = 0; k <= i; ++k) v += Linv[(size_t)i * n + k] * H[(size_t)k * n + j];
            T[(size_t)i * n + j] = v;
        }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double v = 0.0;
            for (int k = 0; k <= j; ++k) v += T[(size_t)i * n + k] * Linv[(size_t)j * n + k];
            A[(size_t)i * n + j] = A[(size_t)j * n + i] = v;
        }
    if (xtb_oracle_syev(n, A, eps)) return -1;
    /* C = Linv^T Y */
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < n; ++k) {
            double v = 0.0;
            for (int j = i; j < n; ++j) v += Linv[(size_t)j * n + i] * A[(size_t)j * n + k];
            C[(size_t)i * n + k] = v;
        }
    return 0;
}

/* ------------------------------------------------------------- Fermi smearing --- */
static void fermi_fill(int n, int homo, double kt, const double *emo, double *occ, double *efermi) {
    const double thr = sqrt(2.220446049250313e-16);
    double ef = 0.5 * (emo[(homo > 1 ? homo : 1) - 1] + emo[(homo + 1 < n ? homo + 1 : n) - 1]);
    double occt = homo;
    for (int cyc = 0; cyc < 200; ++cyc) {
        double total = 0.0, dtotal = 0.0;
        for (int i = 0; i < n; ++i) {
            double f = 0.0, df = 0.0, x = (emo[i] - ef) / kt;
            if (x < 50.0) {
                double ex = exp(x);
                f = 1.0 / (ex + 1.0);
                df = ex / (kt * (ex + 1.0) * (ex + 1.0));
            }
            occ[i] = f;
            total += f;
            dtotal += df;
        }
        double change = (occt - total) / dtotal;
        ef += change;
        if (fabs(occt - total) <= thr) break;
    }
    *efermi = ef;
}

static double electronic_entropy(int n, const double *occ, double kt) {
    const double thr = sqrt(2.220446049250313e-16);
    double sacc = 0.0;
    for (int i = 0; i < n; ++i)
        if (occ[i] > thr && 1.0 - occ[i] > thr) sacc += occ[i] * log(occ[i]) + (1.0 - occ[i]) * log(1.0 - occ[i]);
    return sacc * kt;
}

/* ---------------------------------------------------------------- D4 model --- */
typedef struct {
    int nref[ELEM_MAXZ + 1];
    int ngw[GFN2_MAXZ + 1][D4_MAXREF];
    double alpha[GFN2_MAXZ + 1][D4_MAXREF][D4_NFREQ];
    double *c6ref; /* [nat types collapsed: index by (zi, ri, zj, rj)] */
} d4_model_t;

static double d4_zeta(double a, double c, double qref, double qmod) {
    return qmod < 0.0 ? exp(a) : exp(a * (1.0 - exp(c * (1.0 - qref / qmod))));
}
static double d4_dzeta(double a, double c, double qref, double qmod) {
    return qmod < 0.0 ? 0.0 : -a * c * exp(c * (1.0 - qref / qmod)) * d4_zeta(a, c, qref, qmod) * qref / (qmod * qmod);
}
static double d4_trapz(const double *pol) {
    static const double freq[D4_NFREQ] = {0.000001, 0.050000, 0.100000, 0.200000, 0.300000, 0.400000,
                                          0.500000, 0.600000, 0.700000, 0.800000, 0.900000, 1.000000,
                                          1.200000, 1.400000, 1.600000, 1.800000, 2.000000, 2.500000,
                                          3.000000, 4.000000, 5.000000, 7.500000, 10.00000};
    double acc = 0.0;
    for (int k = 0; k < D4_NFREQ - 1; ++k) acc += 0.5 * (freq[k + 1] - freq[k]) * (pol[k + 1] + pol[k]);
    return acc;
}

/* reference polarisabilities and Gaussian-weight multiplicities for element z */
static void d4_ref_setup(int z, int *ngw, double alpha[D4_MAXREF][D4_NFREQ]) {
    int cnc[32];
    memset(cnc, 0, sizeof cnc);
    cnc[0] = 1;
    int nref = D4_REFN[z];
    for (int r = 0; r < nref; ++r) {
        int is = D4_REFSYS[z][r];
        double iz = D4_ZEFF[is];
        double zt = d4_zeta(GFN2_D4_GA, D4_GAM[is] * GFN2_D4_GC, D4_SECQ[is] + iz, D4_GFFH[z][r] + iz);
        for (int k = 0; k < D4_NFREQ; ++k) {
            double aiw = D4_SSCALE[is] * D4_SECAIW[is][k] * zt;
            double v = D4_ASCALE[z][r] * (D4_ALPHAIW[z][r][k] - D4_HCOUNT[z][r] * aiw);
            alpha[r][k] = v > 0.0 ? v : 0.0;
        }
        int icn = (int)lround(D4_REFCN[z][r]);
        cnc[icn] += 1;
    }
    for (int r = 0; r < nref; ++r) {
        int icn = cnc[(int)lround(D4_REFCN[z][r])];
        ngw[r] = icn * (icn + 1) / 2;
    }
}

/* Gaussian CN weights times charge scaling zeta for every (ref, atom) */
static void d4_weights(const sys_t *s, const int (*ngw)[D4_MAXREF], const double *cn, const double *q,
                       double *gw, double *gwdcn, double *gwdq) {
    for (int i = 0; i < s->nat; ++i) {
        int z = s->num[i], nref = D4_REFN[z];
        double zi = D4_ZEFF[z], gi = D4_GAM[z] * GFN2_D4_GC;
        double norm = 0.0, dnorm = 0.0;
        for (int r = 0; r < nref; ++r)
            for (int g = 1; g <= ngw[z][r]; ++g) {
                double wf = g * GFN2_D4_WF, dc = cn[i] - D4_REFCOVCN[z][r];
                double w = exp(-wf * dc * dc);
                norm += w;
                dnorm += 2.0 * wf * (D4_REFCOVCN[z][r] - cn[i]) * w;
            }
        norm = 1.0 / norm;
        double maxcn = -1.0;
        for (int r = 0; r < nref; ++r)
            if (D4_REFCOVCN[z][r] > maxcn) maxcn = D4_REFCOVCN[z][r];
        for (int r = 0; r < nref; ++r) {
            double expw = 0.0, expd = 0.0;
            for (int g = 1; g <= ngw[z][r]; ++g) {
                double wf = g * GFN2_D4_WF, dc = cn[i] - D4_REFCOVCN[z][r];
                double w = exp(-wf * dc * dc);
                expw += w;
                expd += 2.0 * wf * (D4_REFCOVCN[z][r] - cn[i]) * w;
            }
            double gwk = expw * norm;
            if (gwk != gwk || fabs(gwk) > 1e300) gwk = (maxcn == D4_REFCOVCN[z][r]) ? 1.0 : 0.0;
            double dgwk = norm * (expd - expw * dnorm * norm);
            if (dgwk != dgwk || fabs(dgwk) > 1e300) dgwk = 0.0;
            double zt = d4_zeta(GFN2_D4_GA, gi, D4_GFFQ[z][r] + zi, q[i] + zi);
            double dzt = d4_dzeta(GFN2_D4_GA, gi, D4_GFFQ[z][r] + zi, q[i] + zi);
            gw[i * D4_MAXREF + r] = gwk * zt;
            if (gwdcn) gwdcn[i * D4_MAXREF + r] = dgwk * zt;
            if (gwdq) gwdq[i * D4_MAXREF + r] = gwk * dzt;
        }
    }
}

/* ------------------------------------------------------------ the calculation --- */
typedef struct {
    /* second-order electrostatics */
    double *gamma;   /* [nsh*nsh] */
    /* multipole electrostatics */
    double *mrad, *dmrdcn;            /* [nat] */
    double *sd, *dd, *sq;             /* sd[(d*nat+c)*3+k], dd[((d*nat+c)*3+k)*3+l], sq[(d*nat+c)*6+k] */
    /* D4 */
    int ngw[GFN2_MAXZ + 1][D4_MAXREF];
    double alpha[GFN2_MAXZ + 1][D4_MAXREF][D4_NFREQ];
    double *c6ref;   /* [(i*MAXREF+ri)*(nat*MAXREF) + j*MAXREF+rj] reference C6 of atom pairs */
    double *dispmat; /* same layout: pairwise two-body dispersion kernel */
    double *gw, *gwdcn, *gwdq;
} cache_t;

static double bj_r0(int zi, int zj) { return GFN2_D4_A1 * sqrt(3.0 * D4_R4R2[zi] * D4_R4R2[zj]) + GFN2_D4_A2; }

static void d4_setup(const sys_t *s, cache_t *c) {
    int nat = s->nat, nd = nat * D4_MAXREF;
    int done[GFN2_MAXZ + 1];
    memset(done, 0, sizeof done);
    for (int i = 0; i < nat; ++i) {
        int z = s->num[i];
        if (!done[z]) {
            d4_ref_setup(z, c->ngw[z], c->alpha[z]);
            done[z] = 1;
        }
    }
    /* reference C6 by Casimir-Polder quadrature (3/pi) int alpha_i alpha_j */
    double c6z[GFN2_MAXZ + 1][GFN2_MAXZ + 1][D4_MAXREF][D4_MAXREF];
    for (int zi = 1; zi <= GFN2_MAXZ; ++zi)
        for (int zj = 1; zj <= zi; ++zj) {
            if (!done[zi] || !done[zj]) continue;
            for (int ri = 0; ri < D4_REFN[zi]; ++ri)
                for (int rj = 0; rj < D4_REFN[zj]; ++rj) {
                    double prod[D4_NFREQ];
                    for (int k = 0; k < D4_NFREQ; ++k) prod[k] = c->alpha[zi][ri][k] * c->alpha[zj][rj][k];
                    double v = 3.0 / PI * d4_trapz(prod);
                    c6z[zi][zj][ri][rj] = v;
                    c6z[zj][zi][rj][ri] = v;
                }
        }
    memset(c->c6ref, 0, (size_t)nd * nd * sizeof(double));
    memset(c->dispmat, 0, (size_t)nd * nd * sizeof(double));
    const double cutoff2 = 60.0 * 60.0;
    for (int i = 0; i < nat; ++i)
        for (int j = 0; j < nat; ++j) {
            int zi = s->num[i], zj = s->num[j];
            double v[3] = {s->xyz[3 * i] - s->xyz[3 * j], s->xyz[3 * i + 1] - s->xyz[3 * j + 1],
                           s->xyz[3 * i + 2] - s->xyz[3 * j + 2]};
            double r2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
            double edisp = 0.0;
            if (i != j && r2 <= cutoff2) {
                double r0 = bj_r0(zi, zj), rrij = 3.0 * D4_R4R2[zi] * D4_R4R2[zj];
                double t6 = 1.0 / (r2 * r2 * r2 + pow(r0, 6)), t8 = 1.0 / (r2 * r2 * r2 * r2 + pow(r0, 8));
                edisp = GFN2_D4_S6 * t6 + GFN2_D4_S8 * rrij * t8;
            }
            for (int ri = 0; ri < D4_REFN[zi]; ++ri)
                for (int rj = 0; rj < D4_REFN[zj]; ++rj) {
                    size_t idx = (size_t)(i * D4_MAXREF + ri) * nd + j * D4_MAXREF + rj;
                    c->c6ref[idx] = c6z[zi][zj][ri][rj];
                    c->dispmat[idx] = -c6z[zi][zj][ri][rj] * edisp;
                }
        }
}

/* atomic C6 and derivatives from weights */
static void d4_atomic_c6(const sys_t *s, const cache_t *c, const double *gw, const double *gwdcn, double *c6,
                         double *dc6dcn) {
    int nat = s->nat, nd = nat * D4_MAXREF;
    for (int i = 0; i < nat; ++i)
        for (int j = 0; j < nat; ++j) {
            double v = 0.0, dv = 0.0;
            for (int ri = 0; ri < D4_REFN[s->num[i]]; ++ri)
                for (int rj = 0; rj < D4_REFN[s->num[j]]; ++rj) {
                    double ref = c->c6ref[(size_t)(i * D4_MAXREF + ri) * nd + j * D4_MAXREF + rj];
                    v += gw[i * D4_MAXREF + ri] * gw[j * D4_MAXREF + rj] * ref;
                    dv += gwdcn[i * D4_MAXREF + ri] * gw[j * D4_MAXREF + rj] * ref;
                }
            c6[i * nat + j] = v;
            dc6dcn[i * nat + j] = dv; /* d c6(i,j) / d cn_i */
        }
}

/* Axilrod-Teller-Muto term with q = 0 weights; adds to gradient and dEdcn(D4) */
static double d4_atm(const sys_t *s, const cache_t *c, double *grad, double *dEdcn) {
    int nat = s->nat;
    double *q0 = xcalloc(nat, sizeof(double));
    double *gw = xcalloc((size_t)nat * D4_MAXREF, sizeof(double));
    double *gwdcn = xcalloc((size_t)nat * D4_MAXREF, sizeof(double));
    double *c6 = xcalloc((size_t)nat * nat, sizeof(double));
    double *dc6 = xcalloc((size_t)nat * nat, sizeof(double));
    d4_weights(s, c->ngw, s->cnd4, q0, gw, gwdcn, NULL);
    d4_atomic_c6(s, c, gw, gwdcn, c6, dc6);
    const double cutoff2 = 40.0 * 40.0, alp = GFN2_D4_ALP;
    double e = 0.0;
    for (int i = 0; i < nat; ++i)
        for (int j = 0; j < i; ++j) {
            double vij[3], r2ij = 0;
            for (int k = 0; k < 3; ++k) { vij[k] = s->xyz[3 * j + k] - s->xyz[3 * i + k]; r2ij += vij[k] * vij[k]; }
            if (r2ij > cutoff2) continue;
            double c6ij = c6[i * nat + j], r0ij = bj_r0(s->num[i], s->num[j]);
            for (int k = 0; k < j; ++k) {
                double vik[3], vjk[3], r2ik = 0, r2jk = 0;
                for (int d = 0; d < 3; ++d) {
                    vik[d] = s->xyz[3 * k + d] - s->xyz[3 * i + d];
                    vjk[d] = s->xyz[3 * k + d] - s->xyz[3 * j + d];
                    r2ik += vik[d] * vik[d];
                    r2jk += vjk[d] * vjk[d];
                }
                if (r2ik > cutoff2 || r2jk > cutoff2) continue;
                double c6ik = c6[i * nat + k], c6jk = c6[j * nat + k];
                double r0 = r0ij * bj_r0(s->num[i], s->num[k]) * bj_r0(s->num[j], s->num[k]);
                double c9 = -GFN2_D4_S9 * sqrt(fabs(c6ij * c6ik * c6jk));
                double r2 = r2ij * r2ik * r2jk, r1 = sqrt(r2), r3 = r2 * r1, r5 = r3 * r2;
                double rr0 = pow(r0 / r1, alp / 3.0);
                double fdmp = 1.0 / (1.0 + 6.0 * rr0);
                double ang = 0.375 * (r2ij + r2jk - r2ik) * (r2ij - r2jk + r2ik) * (-r2ij + r2jk + r2ik) / r5 + 1.0 / r3;
                double rr = ang * fdmp;
                double dE = rr * c9;
                e -= dE;
                /* derivatives with respect to the three squared distances */
                double dfdmp = -2.0 * alp * rr0 * fdmp * fdmp; /* d fdmp / d ln r1 * ... see below */
                /* d/d r_ij of ang (times r_ij -> expressed per r2): use explicit formulas */
                double dang_ij = -0.375 * (pow(r2ij, 3) + r2ij * r2ij * (r2jk + r2ik) +
                                           r2ij * (3.0 * r2jk * r2jk + 2.0 * r2jk * r2ik + 3.0 * r2ik * r2ik) -
                                           5.0 * (r2jk - r2ik) * (r2jk - r2ik) * (r2jk + r2ik)) / r5;
                double dang_ik = -0.375 * (pow(r2ik, 3) + r2ik * r2ik * (r2jk + r2ij) +
                                           r2ik * (3.0 * r2jk * r2jk + 2.0 * r2jk * r2ij + 3.0 * r2ij * r2ij) -
                                           5.0 * (r2jk - r2ij) * (r2jk - r2ij) * (r2jk + r2ij)) / r5;
                double dang_jk = -0.375 * (pow(r2jk, 3) + r2jk * r2jk * (r2ik + r2ij) +
                                           r2jk * (3.0 * r2ik * r2ik + 2.0 * r2ik * r2ij + 3.0 * r2ij * r2ij) -
                                           5.0 * (r2ik - r2ij) * (r2ik - r2ij) * (r2ik + r2ij)) / r5;
                /* dang_xy = r_xy * d ang / d r_xy ; fdmp depends on r1 only: r_xy d fdmp/d r_xy = -dfdmp... */
                double gij = c9 * (-dang_ij * fdmp + ang * dfdmp) / r2ij;
                double gik = c9 * (-dang_ik * fdmp + ang * dfdmp) / r2ik;
                double gjk = c9 * (-dang_jk * fdmp + ang * dfdmp) / r2jk;
                for (int d = 0; d < 3; ++d) {
                    grad[3 * i + d] += -gij * vij[d] - gik * vik[d];
                    grad[3 * j + d] += gij * vij[d] - gjk * vjk[d];
                    grad[3 * k + d] += gik * vik[d] + gjk * vjk[d];
                }
                dEdcn[i] -= dE * 0.5 * (dc6[i * nat + j] / c6ij + dc6[i * nat + k] / c6ik);
                dEdcn[j] -= dE * 0.5 * (dc6[j * nat + i] / c6ij + dc6[j * nat + k] / c6jk);
                dEdcn[k] -= dE * 0.5 * (dc6[k * nat + i] / c6ik + dc6[k * nat + j] / c6jk);
            }
        }
    free(q0); free(gw); free(gwdcn); free(c6); free(dc6);
    return e;
}

/* classical repulsion */
static double repulsion(const sys_t *s, double *grad) {
    double e = 0.0;
    for (int i = 0; i < s->nat; ++i)
        for (int j = 0; j < i; ++j) {
            const gfn2_elem_t *ei = &s->elem[s->num[i]], *ej = &s->elem[s->num[j]];
            double v[3], r2 = 0;
            for (int k = 0; k < 3; ++k) { v[k] = s->xyz[3 * i + k] - s->xyz[3 * j + k]; r2 += v[k] * v[k]; }
            double r = sqrt(r2);
            double kexp = s->method == QC_METHOD_GFN1 ? GFN1_REP_KEXP : ((s->num[i] <= 2 && s->num[j] <= 2) ? GFN2_REP_KEXP_LIGHT : GFN2_REP_KEXP);
            double alpha = sqrt(ei->rep_alpha * ej->rep_alpha), zz = ei->rep_zeff * ej->rep_zeff;
            double rk = pow(r, kexp);
            double ex = exp(-alpha * rk);
            double eij = zz * ex / r; /* rexp = 1 */
            e += eij;
            double dedr = -(alpha * rk * kexp + GFN2_REP_REXP) * eij / r2; /* (1/r) dE/dr */
            for (int k = 0; k < 3; ++k) {
                grad[3 * i + k] += dedr * v[k];
                grad[3 * j + k] -= dedr * v[k];
            }
        }
    return e;
}

static void coulomb_setup(const sys_t *s, cache_t *c) {
    int nat = s->nat, nsh = s->nsh;
    for (int a = 0; a < nsh; ++a)
        for (int b = 0; b < nsh; ++b) {
            int i = s->sh_at[a], j = s->sh_at[b];
            double gam = s->method == QC_METHOD_GFN1 ? 2.0 / (1.0 / s->sh_hub[a] + 1.0 / s->sh_hub[b])   /* harmonic average */
                                                     : 0.5 * (s->sh_hub[a] + s->sh_hub[b]);
            if (i == j)
                c->gamma[a * nsh + b] = gam;
            else {
                double r2 = 0;
                for (int k = 0; k < 3; ++k) { double d = s->xyz[3 * i + k] - s->xyz[3 * j + k]; r2 += d * d; }
                c->gamma[a * nsh + b] = 1.0 / sqrt(r2 + 1.0 / (gam * gam)); /* gexp = 2 */
            }
        }
    if (s->method == QC_METHOD_GFN1) return;   /* no multipole electrostatics */
    /* multipole damping radii from the GFN CN */
    for (int i = 0; i < nat; ++i) {
        const gfn2_elem_t *e = &s->elem[s->num[i]];
        double arg = s->cn[i] - e->mp_vcn - GFN2_MP_SHIFT;
        double t1 = exp(-GFN2_MP_KEXP * arg);
        double t2 = (GFN2_MP_RMAX - e->mp_rad) / (1.0 + t1);
        c->mrad[i] = e->mp_rad + t2;
        c->dmrdcn[i] = t2 * GFN2_MP_KEXP * t1 / (1.0 + t1);
    }
    memset(c->sd, 0, (size_t)nat * nat * 3 * sizeof(double));
    memset(c->dd, 0, (size_t)nat * nat * 9 * sizeof(double));
    memset(c->sq, 0, (size_t)nat * nat * 6 * sizeof(double));
    for (int d = 0; d < nat; ++d)      /* multipole site */
        for (int q = 0; q < nat; ++q) { /* partner site */
            if (d == q) continue;
            double v[3], r2 = 0;
            for (int k = 0; k < 3; ++k) { v[k] = s->xyz[3 * q + k] - s->xyz[3 * d + k]; r2 += v[k] * v[k]; }
            double r1 = sqrt(r2), g1 = 1.0 / r1, g3 = g1 * g1 * g1, g5 = g3 * g1 * g1;
            double rr = 0.5 * (c->mrad[d] + c->mrad[q]) * g1;
            double f3 = 1.0 / (1.0 + 6.0 * pow(rr, GFN2_MP_DMP3)), f5 = 1.0 / (1.0 + 6.0 * pow(rr, GFN2_MP_DMP5));
            size_t p = (size_t)d * nat + q;
            for (int k = 0; k < 3; ++k) c->sd[p * 3 + k] = v[k] * g3 * f3;
            for (int k = 0; k < 3; ++k)
                for (int l = 0; l < 3; ++l) c->dd[(p * 3 + k) * 3 + l] = ((k == l) ? g3 * f5 : 0.0) - 3.0 * v[k] * v[l] * g5 * f5;
            double tc[6] = {v[0] * v[0], 2 * v[0] * v[1], v[1] * v[1], 2 * v[0] * v[2], 2 * v[1] * v[2], v[2] * v[2]};
            for (int k = 0; k < 6; ++k) c->sq[p * 6 + k] = tc[k] * g5 * f5;
        }
}

static const double QSCALE[6] = {1.0, 2.0, 1.0, 2.0, 2.0, 1.0};

typedef struct {
    double *vsh, *vat, *vdp, *vqp, *vao;
} pot_t;

/* potentials from (input) populations */
static void build_potential(const sys_t *s, const cache_t *c, const double *qsh, const double *qat, const double *dpat,
                            const double *qpat, pot_t *p) {
    int nat = s->nat, nsh = s->nsh, nd = nat * D4_MAXREF;
    memset(p->vsh, 0, nsh * sizeof(double));
    memset(p->vat, 0, nat * sizeof(double));
    memset(p->vdp, 0, 3 * nat * sizeof(double));
    memset(p->vqp, 0, 6 * nat * sizeof(double));
    /* isotropic second order + onsite third order */
    for (int a = 0; a < nsh; ++a) {
        double v = 0.0;
        for (int b = 0; b < nsh; ++b) v += c->gamma[a * nsh + b] * qsh[b];
        p->vsh[a] += v + qsh[a] * qsh[a] * s->sh_gam3[a];
    }
    if (s->method == QC_METHOD_GFN1) {   /* atom-resolved third order; no multipoles, no self-consistent dispersion */
        for (int i = 0; i < nat; ++i) p->vat[i] += qat[i] * qat[i] * s->at_gam3[i];
        for (int mu = 0; mu < s->nao; ++mu) p->vao[mu] = p->vsh[s->ao_sh[mu]] + p->vat[s->ao_at[mu]];
        return;
    }
    /* anisotropic electrostatics + multipole XC kernels */
    for (int d = 0; d < nat; ++d) {
        const gfn2_elem_t *e = &s->elem[s->num[d]];
        for (int q = 0; q < nat; ++q) {
            if (q == d) continue;
            size_t pq = (size_t)d * nat + q;
            for (int k = 0; k < 3; ++k) {
                p->vdp[3 * d + k] += c->sd[pq * 3 + k] * qat[q];
                for (int l = 0; l < 3; ++l) p->vdp[3 * d + k] += c->dd[(pq * 3 + k) * 3 + l] * dpat[3 * q + l];
                p->vat[q] += c->sd[pq * 3 + k] * dpat[3 * d + k];
            }
            for (int k = 0; k < 6; ++k) {
                p->vqp[6 * d + k] += c->sq[pq * 6 + k] * qat[q];
                p->vat[q] += c->sq[pq * 6 + k] * qpat[6 * d + k];
            }
        }
        for (int k = 0; k < 3; ++k) p->vdp[3 * d + k] += 2.0 * e->dkernel * dpat[3 * d + k];
        for (int k = 0; k < 6; ++k) p->vqp[6 * d + k] += 2.0 * e->qkernel * qpat[6 * d + k] * QSCALE[k];
    }
    /* self-consistent D4 */
    d4_weights(s, c->ngw, s->cnd4, qat, c->gw, NULL, c->gwdq);
    for (int i = 0; i < nat; ++i)
        for (int ri = 0; ri < D4_REFN[s->num[i]]; ++ri) {
            double v = 0.0;
            const double *row = c->dispmat + (size_t)(i * D4_MAXREF + ri) * nd;
            for (int x = 0; x < nd; ++x) v += row[x] * c->gw[x];
            p->vat[i] += v * c->gwdq[i * D4_MAXREF + ri];
        }
    for (int mu = 0; mu < s->nao; ++mu) p->vao[mu] = p->vsh[s->ao_sh[mu]] + p->vat[s->ao_at[mu]];
}

/* energies of the charge-dependent terms at (output) populations */
static void scc_energies(const sys_t *s, cache_t *c, const double *qsh, const double *qat, const double *dpat,
                         const double *qpat, double *es2, double *es3, double *eaes, double *ed4) {
    int nat = s->nat, nsh = s->nsh, nd = nat * D4_MAXREF;
    double e2 = 0, e3 = 0, ea = 0, ed = 0;
    for (int a = 0; a < nsh; ++a) {
        double v = 0.0;
        for (int b = 0; b < nsh; ++b) v += c->gamma[a * nsh + b] * qsh[b];
        e2 += 0.5 * v * qsh[a];
        e3 += qsh[a] * qsh[a] * qsh[a] * s->sh_gam3[a] / 3.0;
    }
    if (s->method == QC_METHOD_GFN1) {
        for (int i = 0; i < nat; ++i) e3 += qat[i] * qat[i] * qat[i] * s->at_gam3[i] / 3.0;
        *es2 = e2; *es3 = e3; *eaes = 0.0; *ed4 = 0.0;
        return;
    }
    for (int d = 0; d < nat; ++d) {
        const gfn2_elem_t *e = &s->elem[s->num[d]];
        double vd[3] = {0, 0, 0}, vq[6] = {0, 0, 0, 0, 0, 0};
        for (int q = 0; q < nat; ++q) {
            if (q == d) continue;
            size_t pq = (size_t)d * nat + q;
            for (int k = 0; k < 3; ++k) {
                vd[k] += c->sd[pq * 3 + k] * qat[q];
                for (int l = 0; l < 3; ++l) vd[k] += 0.5 * c->dd[(pq * 3 + k) * 3 + l] * dpat[3 * q + l];
            }
            for (int k = 0; k < 6; ++k) vq[k] += c->sq[pq * 6 + k] * qat[q];
        }
        for (int k = 0; k < 3; ++k) ea += dpat[3 * d + k] * vd[k] + e->dkernel * dpat[3 * d + k] * dpat[3 * d + k];
        for (int k = 0; k < 6; ++k) ea += qpat[6 * d + k] * vq[k] + e->qkernel * qpat[6 * d + k] * qpat[6 * d + k] * QSCALE[k];
    }
    d4_weights(s, c->ngw, s->cnd4, qat, c->gw, NULL, NULL);
    for (int x = 0; x < nd; ++x) {
        double v = 0.0;
        const double *row = c->dispmat + (size_t)x * nd;
        for (int y = 0; y < nd; ++y) v += row[y] * c->gw[y];
        ed += 0.5 * v * c->gw[x];
    }
    *es2 = e2; *es3 = e3; *eaes = ea; *ed4 = ed;
}

/* modified Broyden mixer (damping 0.4, full history) */
typedef struct {
    int ndim, iter, memory;
    double damp;
    double *q_in, *qlast_in, *dq, *dqlast, *df, *u, *a, *omega;
} broyden_t;

static int lin_solve(int n, double *A, double *b) { /* Gaussian elimination with partial pivoting */
    for (int k = 0; k < n; ++k) {
        int piv = k;
        for (int i = k + 1; i < n; ++i)
            if (fabs(A[i * n + k]) > fabs(A[piv * n + k])) piv = i;
        if (A[piv * n + k] == 0.0) return -1;
        if (piv != k) {
            for (int j = 0; j < n; ++j) { double t = A[k * n + j]; A[k * n + j] = A[piv * n + j]; A[piv * n + j] = t; }
            double t = b[k]; b[k] = b[piv]; b[piv] = t;
        }
        for (int i = k + 1; i < n; ++i) {
            double f = A[i * n + k] / A[k * n + k];
            for (int j = k; j < n; ++j) A[i * n + j] -= f * A[k * n + j];
            b[i] -= f * b[k];
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        double v = b[i];
        for (int j = i + 1; j < n; ++j) v -= A[i * n + j] * b[j];
        b[i] = v / A[i * n + i];
    }
    return 0;
}

static int broyden_next(broyden_t *m) {
    int n = m->ndim, mem = m->memory;
    m->iter += 1;
    int iter = m->iter, itn = iter - 1;
    const double omega0 = 0.01, minw = 1.0, maxw = 100000.0, wfac = 0.01;
    if (iter == 1) {
        memcpy(m->dqlast, m->dq, n * sizeof(double));
        memcpy(m->qlast_in, m->q_in, n * sizeof(double));
        for (int i = 0; i < n; ++i) m->q_in[i] += m->damp * m->dq[i];
        return 0;
    }
    int it1 = (itn - 1) % mem; /* 0-based slot */
    int nb = itn < mem ? itn : mem;
    double *beta = xcalloc((size_t)nb * nb, sizeof(double)), *cvec = xcalloc(nb, sizeof(double));
    double nrm = 0.0;
    for (int i = 0; i < n; ++i) nrm += m->dq[i] * m->dq[i];
    nrm = sqrt(nrm);
    double om = nrm > wfac / maxw ? wfac / nrm : maxw;
    if (om < minw) om = minw;
    m->omega[it1] = om;
    double inv = 0.0;
    for (int i = 0; i < n; ++i) {
        double v = m->dq[i] - m->dqlast[i];
        m->df[(size_t)it1 * n + i] = v;
        inv += v * v;
    }
    inv = sqrt(inv);
    if (inv < 2.220446049250313e-16) inv = 2.220446049250313e-16;
    inv = 1.0 / inv;
    for (int i = 0; i < n; ++i) m->df[(size_t)it1 * n + i] *= inv;
    int j0 = itn - mem + 1 > 1 ? itn - mem + 1 : 1;
    for (int j = j0; j <= itn; ++j) {
        int i = (j - 1) % mem;
        double aij = 0.0, ci = 0.0;
        for (int k = 0; k < n; ++k) {
            aij += m->df[(size_t)i * n + k] * m->df[(size_t)it1 * n + k];
            ci += m->df[(size_t)i * n + k] * m->dq[k];
        }
        m->a[i * mem + it1] = aij;
        m->a[it1 * mem + i] = aij;
        cvec[i] = m->omega[i] * ci;
    }
    for (int j = j0; j <= itn; ++j) {
        int i = (j - 1) % mem;
        for (int k = 0; k < nb; ++k) beta[k * nb + i] = m->omega[k] * m->omega[i] * m->a[k * mem + i];
        beta[i * nb + i] += omega0 * omega0;
    }
    int info = lin_solve(nb, beta, cvec);
    for (int i = 0; i < n; ++i)
        m->u[(size_t)it1 * n + i] = m->damp * m->df[(size_t)it1 * n + i] + inv * (m->q_in[i] - m->qlast_in[i]);
    memcpy(m->dqlast, m->dq, n * sizeof(double));
    memcpy(m->qlast_in, m->q_in, n * sizeof(double));
    for (int i = 0; i < n; ++i) m->q_in[i] += m->damp * m->dq[i];
    for (int j = j0; j <= itn; ++j) {
        int i = (j - 1) % mem;
        for (int k = 0; k < n; ++k) m->q_in[k] -= m->omega[i] * cvec[i] * m->u[(size_t)i * n + k];
    }
    free(beta);
    free(cvec);
    return info;
}

/* The reference hard-codes accuracy = 1.0 (src/tblite.f90:46).  Tests may tighten it to
 * separate SCC-threshold noise from genuine disagreement (finite-difference checks). */
/* ------------------------------------------------------------ GFN1: D3(BJ) --- */
/* C6(i,j) interpolated over the reference systems with Gaussian weights in CN space and its derivatives with respect to both
 * coordination numbers (reference src/dftd3.f90:334-405 get_dC6_dCNij, k3 = -4). */
static void d3_c6(int zi, int zj, double cni, double cnj, double *c6, double *dc6i, double *dc6j) {
    const double k3 = -4.0;
    double c6mem = -1.e99, r_save = 9999.0, zaehler = 0, nenner = 0, dzi = 0, dni = 0, dzj = 0, dnj = 0;
    for (int a = 0; a < D3_MXC[zi]; ++a)
        for (int b = 0; b < D3_MXC[zj]; ++b) {
            double c6ref = D3_C6AB[zi][zj][a][b][0];
            if (c6ref > 0) {
                double cn_refi = D3_C6AB[zi][zj][a][b][1], cn_refj = D3_C6AB[zi][zj][a][b][2];
                double r = (cn_refi - cni) * (cn_refi - cni) + (cn_refj - cnj) * (cn_refj - cnj);
                if (r < r_save) { r_save = r; c6mem = c6ref; }
                double expterm = exp(k3 * r);
                zaehler += c6ref * expterm;
                nenner += expterm;
                expterm = expterm * 2.0 * k3;
                double term = expterm * (cni - cn_refi);
                dzi += c6ref * term; dni += term;
                term = expterm * (cnj - cn_refj);
                dzj += c6ref * term; dnj += term;
            }
        }
    if (nenner > 1.0e-99) {
        *c6 = zaehler / nenner;
        *dc6i = ((dzi * nenner) - (dni * zaehler)) / (nenner * nenner);
        *dc6j = ((dzj * nenner) - (dnj * zaehler)) / (nenner * nenner);
    } else {
        *c6 = c6mem; *dc6i = 0.0; *dc6j = 0.0;
    }
}

/* two-body D3 with Becke-Johnson damping (reference src/dftd3.f90:107-186, BJ variant, no three-body term):
 * E = - sum_{i>j} C6 (s6 / (r^6 + R0^6) + 3 s8 r42 / (r^8 + R0^8)), R0 = a1 sqrt(3 r42) + a2, r42 = r2r4_i r2r4_j.
 * Adds dE/dR at fixed CN to grad and dE/dCN to dEdcn (the CN is the exponential one the Hamiltonian uses). */
static double d3_bj(const sys_t *s, double *grad, double *dEdcn) {
    const double rthr = 4000.0;
    double disp = 0.0;
    for (int i = 0; i < s->nat; ++i)
        for (int j = 0; j < i; ++j) {
            double v[3], r2 = 0;
            for (int k = 0; k < 3; ++k) { v[k] = s->xyz[3 * i + k] - s->xyz[3 * j + k]; r2 += v[k] * v[k]; }
            if (r2 > rthr) continue;
            double c6, dc6i, dc6j;
            d3_c6(s->num[i], s->num[j], s->cn[i], s->cn[j], &c6, &dc6i, &dc6j);
            double r42 = D3_R2R4[s->num[i]] * D3_R2R4[s->num[j]];
            double r = sqrt(r2), r4 = r2 * r2, r6 = r4 * r2, r8 = r6 * r2;
            double R0 = GFN1_D3_A1 * sqrt(3.0 * r42) + GFN1_D3_A2;
            double t6 = r6 + pow(R0, 6), t8 = r8 + pow(R0, 8);
            double rest = GFN1_D3_S6 / t6 + 3.0 * GFN1_D3_S8 * r42 / t8;
            disp -= rest * c6;
            /* (1/r) dE/dr at fixed C6 */
            double dedr = c6 * (GFN1_D3_S6 * 6.0 * r4 * r / (t6 * t6) + GFN1_D3_S8 * 24.0 * r42 * r6 * r / (t8 * t8)) / r;
            for (int k = 0; k < 3; ++k) { grad[3 * i + k] += dedr * v[k]; grad[3 * j + k] -= dedr * v[k]; }
            dEdcn[i] -= rest * dc6i;
            dEdcn[j] -= rest * dc6j;
        }
    return disp;
}

/* ------------------------------------------------- GFN1: halogen-bond correction --- */
/* For every halogen X with its nearest neighbour K and every acceptor A (N, O, P, S) within 20 bohr:
 * E = c_X (1/2 - 1/4 cos(A-X-K))^6 (t^12 - damp t^6) / (1 + t^12), t = xbrad (R_A + R_X) / r_AX (atomic radii as in the
 * H0 polynomial).  Restated from the published GFN1-xTB definition (tblite classical/halogen.f90 is not available). */
static double halogen_bond(const sys_t *s, double *grad) {
    double exb = 0.0;
    for (int x = 0; x < s->nat; ++x) {
        if (!gfn1_xb_donor(s->num[x])) continue;
        int kn = -1;
        double best = 1e300;
        for (int k = 0; k < s->nat; ++k) {
            if (k == x) continue;
            double d2 = 0;
            for (int c = 0; c < 3; ++c) { double d = s->xyz[3 * k + c] - s->xyz[3 * x + c]; d2 += d * d; }
            if (d2 < best) { best = d2; kn = k; }
        }
        if (kn < 0) continue;
        for (int a = 0; a < s->nat; ++a) {
            if (a == x || a == kn || !gfn1_xb_acceptor(s->num[a])) continue;
            double u[3], w[3], ru2 = 0, rw2 = 0, uw = 0;
            for (int c = 0; c < 3; ++c) {
                u[c] = s->xyz[3 * a + c] - s->xyz[3 * x + c];
                w[c] = s->xyz[3 * kn + c] - s->xyz[3 * x + c];
                ru2 += u[c] * u[c]; rw2 += w[c] * w[c]; uw += u[c] * w[c];
            }
            if (ru2 > 400.0) continue;
            double ru = sqrt(ru2), rw = sqrt(rw2), cosv = uw / (ru * rw);
            double r0 = GFN1_XB_RAD * (s->elem[s->num[a]].atomic_rad + s->elem[s->num[x]].atomic_rad) * TB_AATOAU;
            double t = r0 / ru, t6 = pow(t, 6), t12 = t6 * t6;
            double lj = (t12 - GFN1_XB_DAMP * t6) / (1.0 + t12);
            double dt6 = -6.0 * t6 / ru, dt12 = -12.0 * t12 / ru;
            double dlj = ((dt12 - GFN1_XB_DAMP * dt6) * (1.0 + t12) - (t12 - GFN1_XB_DAMP * t6) * dt12) / ((1.0 + t12) * (1.0 + t12));
            double base = 0.5 - 0.25 * cosv, at = pow(base, 6), dat = 6.0 * pow(base, 5) * (-0.25);
            double cx = GFN1_EXTRA[s->num[x]].xbond;
            exb += cx * at * lj;
            for (int c = 0; c < 3; ++c) {
                double dcos_a = (w[c] / rw - cosv * u[c] / ru) / ru, dcos_k = (u[c] / ru - cosv * w[c] / rw) / rw;
                double ga = cx * (at * dlj * u[c] / ru + dat * lj * dcos_a), gk = cx * dat * lj * dcos_k;
                grad[3 * a + c] += ga;
                grad[3 * kn + c] += gk;
                grad[3 * x + c] -= ga + gk;
            }
        }
    }
    return exb;
}

static double g_accuracy = 1.0;
void xtb_oracle_set_accuracy(double acc) { g_accuracy = acc; }

int xtb_oracle_egrad(int nat, const int32_t *num, const double *xyz, int charge, int multiplicity, int method_id,
                     double etemp, double *qat_out, double *energy_out, double *grad_out,
                     xtb_oracle_detail_t *detail) {
    /* IPEA1 (method id 11) shares the GFN1 model with another element table that is not reconstructed here */
    if (method_id != QC_METHOD_GFN2 && method_id != QC_METHOD_GFN1) return QC_STAT_UNKNOWN_METHOD;
    const int gfn1 = method_id == QC_METHOD_GFN1;
    if (gfn1) gfn1_ensure_loaded();
    int stat = QC_STAT_OK;
    sys_t s;
    memset(&s, 0, sizeof s);
    s.method = method_id;
    s.elem = gfn1 ? GFN1_ELEM : GFN2_ELEM;
    s.nat = nat; s.num = num; s.xyz = xyz; s.charge = charge;
    s.kt = etemp * QC_KTOAU;
    if (setup_basis(&s)) return QC_STAT_FATAL;
    int nsh = s.nsh, nao = s.nao;
    size_t n2 = (size_t)nao * nao;
    s.cn = xcalloc(nat, sizeof(double));
    s.dcndr = xcalloc((size_t)nat * nat * 3, sizeof(double));
    s.cnd4 = xcalloc(nat, sizeof(double));
    s.dcnd4dr = xcalloc((size_t)nat * nat * 3, sizeof(double));
    s.S = xcalloc(n2, sizeof(double));
    s.H0 = xcalloc(n2, sizeof(double));
    s.D = xcalloc(3 * n2, sizeof(double));
    s.Q = xcalloc(6 * n2, sizeof(double));
    s.selfen = xcalloc(nsh, sizeof(double));
    double *grad = xcalloc(3 * nat, sizeof(double));
    double *dEdcn = xcalloc(nat, sizeof(double)), *dEdcn4 = xcalloc(nat, sizeof(double));

    cache_t c;
    memset(&c, 0, sizeof c);
    int nd = nat * D4_MAXREF;
    c.gamma = xcalloc((size_t)nsh * nsh, sizeof(double));
    c.mrad = xcalloc(nat, sizeof(double));
    c.dmrdcn = xcalloc(nat, sizeof(double));
    c.sd = xcalloc((size_t)nat * nat * 3, sizeof(double));
    c.dd = xcalloc((size_t)nat * nat * 9, sizeof(double));
    c.sq = xcalloc((size_t)nat * nat * 6, sizeof(double));
    c.c6ref = xcalloc((size_t)nd * nd, sizeof(double));
    c.dispmat = xcalloc((size_t)nd * nd, sizeof(double));
    c.gw = xcalloc(nd, sizeof(double));
    c.gwdcn = xcalloc(nd, sizeof(double));
    c.gwdq = xcalloc(nd, sizeof(double));

    /* --- geometry-dependent set-up ------------------------------------------------ */
    double e_rep, e_atm;   /* e_atm: the dispersion (+ halogen bond) energy that does not depend on the charges */
    if (gfn1) {
        exp_cn(&s);
        e_rep = repulsion(&s, grad);
        e_atm = d3_bj(&s, grad, dEdcn) + halogen_bond(&s, grad);
    } else {
        gfn_cn(&s);
        d4_cn(&s);
        e_rep = repulsion(&s, grad);
        d4_setup(&s, &c);
        e_atm = d4_atm(&s, &c, grad, dEdcn4);
    }
    coulomb_setup(&s, &c);
    build_integrals(&s);

    /* --- occupations (tblite get_occupation / get_alpha_beta_occupation) --------- */
    double nocc = -(double)charge;
    for (int a = 0; a < nsh; ++a) nocc += s.sh_refocc[a];
    int uhf = multiplicity - 1 < 0 ? multiplicity - 1 : 0; /* min(mult-1, 0), tblite.f90:111 */
    int nuhf = ((uhf % 2 + 2) % 2 == ((int)lround(nocc) % 2 + 2) % 2) ? uhf : ((int)lround(nocc) % 2 + 2) % 2;
    double diff = nuhf < nocc ? nuhf : nocc;
    double ntmp = nocc - diff;
    double nel[2] = {ntmp / 2 + diff, ntmp / 2};
    if (nocc <= 0.0 || nocc > 2.0 * nao) stat = QC_STAT_FATAL;

    /* --- SCC ------------------------------------------------------------------------ */
    double *qsh = xcalloc(nsh, sizeof(double)), *qat = xcalloc(nat, sizeof(double));
    double *dpat = xcalloc(3 * nat, sizeof(double)), *qpat = xcalloc(6 * nat, sizeof(double));
    pot_t pot = {xcalloc(nsh, sizeof(double)), xcalloc(nat, sizeof(double)), xcalloc(3 * nat, sizeof(double)),
                 xcalloc(6 * nat, sizeof(double)), xcalloc(nao, sizeof(double))};
    double *H1 = xcalloc(n2, sizeof(double)), *C = xcalloc(n2, sizeof(double)), *P = xcalloc(n2, sizeof(double));
    double *Linv = xcalloc(n2, sizeof(double)), *work = xcalloc(2 * n2, sizeof(double));
    double *emo = xcalloc(nao, sizeof(double)), *focc = xcalloc(nao, sizeof(double)), *ftmp = xcalloc(nao, sizeof(double));
    broyden_t mix;
    memset(&mix, 0, sizeof mix);
    mix.ndim = gfn1 ? nsh : nsh + 9 * nat; mix.memory = MAX_ITER; mix.damp = 0.4;   /* GFN1 mixes the shell charges only */
    mix.q_in = xcalloc(mix.ndim, sizeof(double)); mix.qlast_in = xcalloc(mix.ndim, sizeof(double));
    mix.dq = xcalloc(mix.ndim, sizeof(double)); mix.dqlast = xcalloc(mix.ndim, sizeof(double));
    mix.df = xcalloc((size_t)MAX_ITER * mix.ndim, sizeof(double)); mix.u = xcalloc((size_t)MAX_ITER * mix.ndim, sizeof(double));
    mix.a = xcalloc((size_t)MAX_ITER * MAX_ITER, sizeof(double)); mix.omega = xcalloc(MAX_ITER, sizeof(double));

    double eelec = 0.0, e_el = 0, e_es2 = 0, e_es3 = 0, e_aes = 0, e_d4 = 0, ts = 0;
    int iscf = 0, converged = 0;
    const double econv = 1e-6 * g_accuracy, pconv = 2e-5 * g_accuracy;
    if (stat == QC_STAT_OK && cholesky_inverse(nao, s.S, Linv)) stat = QC_STAT_FATAL;
    while (stat == QC_STAT_OK && !converged && iscf < MAX_ITER) {
        double elast = eelec;
        if (iscf > 0) {
            if (broyden_next(&mix)) { stat = QC_STAT_FATAL; break; }
            memcpy(qsh, mix.q_in, nsh * sizeof(double));
            if (!gfn1) {
                memcpy(dpat, mix.q_in + nsh, 3 * nat * sizeof(double));
                memcpy(qpat, mix.q_in + nsh + 3 * nat, 6 * nat * sizeof(double));
            }
            memset(qat, 0, nat * sizeof(double));
            for (int a = 0; a < nsh; ++a) qat[s.sh_at[a]] += qsh[a];
        }
        iscf += 1;
        build_potential(&s, &c, qsh, qat, dpat, qpat, &pot);
        for (int a = 0; a < nao; ++a)
            for (int b = 0; b < nao; ++b) {
                size_t ab = (size_t)a * nao + b, ba = (size_t)b * nao + a;
                double h = s.H0[ab] - 0.5 * s.S[ab] * (pot.vao[a] + pot.vao[b]);
                int ia = s.ao_at[a], ib = s.ao_at[b];
                for (int k = 0; k < 3; ++k)
                    h -= 0.5 * (s.D[k * n2 + ab] * pot.vdp[3 * ib + k] + s.D[k * n2 + ba] * pot.vdp[3 * ia + k]);
                for (int k = 0; k < 6; ++k)
                    h -= 0.5 * (s.Q[k * n2 + ab] * pot.vqp[6 * ib + k] + s.Q[k * n2 + ba] * pot.vqp[6 * ia + k]);
                H1[ab] = h;
            }
        /* mixer input of this cycle */
        memcpy(mix.q_in, qsh, nsh * sizeof(double));
        if (!gfn1) {
            memcpy(mix.q_in + nsh, dpat, 3 * nat * sizeof(double));
            memcpy(mix.q_in + nsh + 3 * nat, qpat, 6 * nat * sizeof(double));
        }
        {   /* test hook: XTB_ORACLE_DUMP=<file> appends [nao, iscf, H1(nao^2), S(nao^2)] of every SCC cycle (eigen-solver studies) */
            const char *dump = getenv("XTB_ORACLE_DUMP");
            if (dump) {
                FILE *fp = fopen(dump, "ab");
                if (fp) {
                    double hdr[2] = {(double)nao, (double)iscf};
                    fwrite(hdr, sizeof(double), 2, fp);
                    fwrite(H1, sizeof(double), n2, fp);
                    fwrite(s.S, sizeof(double), n2, fp);
                    fclose(fp);
                }
            }
        }
        if (solve_gen(nao, H1, Linv, C, emo, work)) { stat = QC_STAT_FATAL; break; }
        memset(focc, 0, nao * sizeof(double));
        ts = 0.0;
        for (int spin = 0; spin < 2; ++spin) {
            int homo = (int)floor(nel[spin]);
            if (fmod(nel[spin], 1.0) > 0.5) homo += 1;
            memset(ftmp, 0, nao * sizeof(double));
            double ef;
            if (homo > 0) fermi_fill(nao, homo, s.kt, emo, ftmp, &ef);
            ts += electronic_entropy(nao, ftmp, s.kt);
            for (int i = 0; i < nao; ++i) focc[i] += ftmp[i];
        }
        /* density P = C f C^T */
        for (int a = 0; a < nao; ++a)
            for (int b = 0; b <= a; ++b) {
                double v = 0.0;
                for (int k = 0; k < nao; ++k) v += C[(size_t)a * nao + k] * focc[k] * C[(size_t)b * nao + k];
                P[(size_t)a * nao + b] = P[(size_t)b * nao + a] = v;
            }
        /* Mulliken populations */
        for (int a = 0; a < nsh; ++a) qsh[a] = s.sh_refocc[a];
        memset(dpat, 0, 3 * nat * sizeof(double));
        memset(qpat, 0, 6 * nat * sizeof(double));
        e_el = 0.0;
        for (int a = 0; a < nao; ++a)
            for (int b = 0; b < nao; ++b) {
                size_t ab = (size_t)a * nao + b;
                double p = P[ab];
                qsh[s.ao_sh[b]] -= p * s.S[ab];
                int ib = s.ao_at[b];
                if (!gfn1) {
                    for (int k = 0; k < 3; ++k) dpat[3 * ib + k] -= p * s.D[k * n2 + ab];
                    for (int k = 0; k < 6; ++k) qpat[6 * ib + k] -= p * s.Q[k * n2 + ab];
                }
                e_el += p * s.H0[ab];
            }
        memset(qat, 0, nat * sizeof(double));
        for (int a = 0; a < nsh; ++a) qat[s.sh_at[a]] += qsh[a];
        /* mixer difference (output - input) */
        for (int i = 0; i < nsh; ++i) mix.dq[i] = qsh[i] - mix.q_in[i];
        if (!gfn1) {
            for (int i = 0; i < 3 * nat; ++i) mix.dq[nsh + i] = dpat[i] - mix.q_in[nsh + i];
            for (int i = 0; i < 6 * nat; ++i) mix.dq[nsh + 3 * nat + i] = qpat[i] - mix.q_in[nsh + 3 * nat + i];
        }
        scc_energies(&s, &c, qsh, qat, dpat, qpat, &e_es2, &e_es3, &e_aes, &e_d4);
        eelec = ts + e_el + e_es2 + e_es3 + e_aes + e_d4;
        double err = 0.0;
        for (int i = 0; i < mix.ndim; ++i) err += mix.dq[i] * mix.dq[i];
        err = sqrt(err / mix.ndim);
        if (detail && detail->e_iter) detail->e_iter[iscf - 1] = eelec;
        converged = fabs(eelec - elast) < econv && err < pconv;
    }
    if (stat == QC_STAT_OK && !converged) stat = QC_STAT_FATAL; /* "SCF not converged in 250 cycles" */

    double energy = e_rep + e_atm + eelec;

    /* --- gradient --------------------------------------------------------------------- */
    if (stat == QC_STAT_OK) {
        /* isotropic ES2: dE/dR = 1/2 sum q_a q_b d gamma_ab */
        for (int a = 0; a < nsh; ++a)
            for (int b = 0; b < nsh; ++b) {
                int i = s.sh_at[a], j = s.sh_at[b];
                if (i == j) continue;
                double g = c.gamma[a * nsh + b];
                double f = -qsh[a] * qsh[b] * g * g * g; /* (1/r) d gamma/dr; (a,b) and (b,a) both act on atom i */
                for (int k = 0; k < 3; ++k) grad[3 * i + k] += f * (xyz[3 * i + k] - xyz[3 * j + k]);
            }
        /* anisotropic ES */
        for (int i = 0; i < (gfn1 ? 0 : nat); ++i)
            for (int j = 0; j < i; ++j) {
                double v[3], r2 = 0;
                for (int k = 0; k < 3; ++k) { v[k] = xyz[3 * j + k] - xyz[3 * i + k]; r2 += v[k] * v[k]; }
                double r = sqrt(r2), g1 = 1.0 / r, g3 = g1 * g1 * g1, g5 = g3 * g1 * g1;
                double R0 = 0.5 * (c.mrad[i] + c.mrad[j]);
                double x3 = 6.0 * pow(R0 * g1, GFN2_MP_DMP3), x5 = 6.0 * pow(R0 * g1, GFN2_MP_DMP5);
                double f3 = 1.0 / (1.0 + x3), f5 = 1.0 / (1.0 + x5);
                double df3dr = f3 * f3 * GFN2_MP_DMP3 * x3 * g1, df5dr = f5 * f5 * GFN2_MP_DMP5 * x5 * g1;
                double df3dR0 = -f3 * f3 * GFN2_MP_DMP3 * x3 / R0, df5dR0 = -f5 * f5 * GFN2_MP_DMP5 * x5 / R0;
                const double *mi = dpat + 3 * i, *mj = dpat + 3 * j, *ti = qpat + 6 * i, *tj = qpat + 6 * j;
                double qi = qat[i], qj = qat[j];
                double miv = mi[0] * v[0] + mi[1] * v[1] + mi[2] * v[2], mjv = mj[0] * v[0] + mj[1] * v[1] + mj[2] * v[2];
                double mimj = mi[0] * mj[0] + mi[1] * mj[1] + mi[2] * mj[2];
                /* Theta v (full symmetric matrix) */
                double tiv[3] = {ti[0] * v[0] + ti[1] * v[1] + ti[3] * v[2], ti[1] * v[0] + ti[2] * v[1] + ti[4] * v[2],
                                 ti[3] * v[0] + ti[4] * v[1] + ti[5] * v[2]};
                double tjv[3] = {tj[0] * v[0] + tj[1] * v[1] + tj[3] * v[2], tj[1] * v[0] + tj[2] * v[1] + tj[4] * v[2],
                                 tj[3] * v[0] + tj[4] * v[1] + tj[5] * v[2]};
                double tivv = tiv[0] * v[0] + tiv[1] * v[1] + tiv[2] * v[2], tjvv = tjv[0] * v[0] + tjv[1] * v[1] + tjv[2] * v[2];
                double A = qj * miv - qi * mjv;        /* multiplies g3 f3 */
                double B = qj * tivv + qi * tjvv;      /* multiplies g5 f5 */
                double Cc = mimj;                      /* multiplies g3 f5 */
                double Dd = -3.0 * miv * mjv;          /* multiplies g5 f5 */
                double dg3 = -3.0 * g3 * g1, dg5 = -5.0 * g5 * g1;
                double radial = (dg3 * f3 + g3 * df3dr) * A + (dg5 * f5 + g5 * df5dr) * (B + Dd) + (dg3 * f5 + g3 * df5dr) * Cc;
                double dER0 = g3 * df3dR0 * A + g5 * df5dR0 * (B + Dd) + g3 * df5dR0 * Cc;
                for (int k = 0; k < 3; ++k) {
                    double dv = radial * v[k] * g1 + g3 * f3 * (qj * mi[k] - qi * mj[k]) +
                                g5 * f5 * (2.0 * qj * tiv[k] + 2.0 * qi * tjv[k] - 3.0 * (mi[k] * mjv + mj[k] * miv));
                    grad[3 * j + k] += dv;
                    grad[3 * i + k] -= dv;
                }
                dEdcn[i] += 0.5 * dER0 * c.dmrdcn[i];
                dEdcn[j] += 0.5 * dER0 * c.dmrdcn[j];
            }
        /* two-body D4 with the final charges */
        if (!gfn1) {
            double *c6 = xcalloc((size_t)nat * nat, sizeof(double)), *dc6 = xcalloc((size_t)nat * nat, sizeof(double));
            d4_weights(&s, c.ngw, s.cnd4, qat, c.gw, c.gwdcn, NULL);
            d4_atomic_c6(&s, &c, c.gw, c.gwdcn, c6, dc6);
            for (int i = 0; i < nat; ++i)
                for (int j = 0; j < i; ++j) {
                    double v[3], r2 = 0;
                    for (int k = 0; k < 3; ++k) { v[k] = xyz[3 * i + k] - xyz[3 * j + k]; r2 += v[k] * v[k]; }
                    if (r2 > 3600.0) continue;
                    double r0 = bj_r0(num[i], num[j]), rrij = 3.0 * D4_R4R2[num[i]] * D4_R4R2[num[j]];
                    double t6 = 1.0 / (r2 * r2 * r2 + pow(r0, 6)), t8 = 1.0 / (r2 * r2 * r2 * r2 + pow(r0, 8));
                    double d6 = -6.0 * r2 * r2 * t6 * t6, d8 = -8.0 * r2 * r2 * r2 * t8 * t8;
                    double edisp = GFN2_D4_S6 * t6 + GFN2_D4_S8 * rrij * t8;
                    double gdisp = GFN2_D4_S6 * d6 + GFN2_D4_S8 * rrij * d8;
                    for (int k = 0; k < 3; ++k) {
                        double dG = -c6[i * nat + j] * gdisp * v[k];
                        grad[3 * i + k] += dG;
                        grad[3 * j + k] -= dG;
                    }
                    dEdcn4[i] -= dc6[i * nat + j] * edisp;
                    dEdcn4[j] -= dc6[j * nat + i] * edisp;
                }
            free(c6); free(dc6);
        }
        /* Hamiltonian / overlap / multipole-integral derivatives */
        double *W = work; /* energy-weighted density */
        for (int a = 0; a < nao; ++a)
            for (int b = 0; b <= a; ++b) {
                double v = 0.0;
                for (int k = 0; k < nao; ++k) v += C[(size_t)a * nao + k] * focc[k] * emo[k] * C[(size_t)b * nao + k];
                W[(size_t)a * nao + b] = W[(size_t)b * nao + a] = v;
            }
        pair_ints_t pi;
        for (int iat = 0; iat < nat; ++iat) {
            /* on-site: only the CN dependence of the diagonal levels */
            for (int is = 0; is < s.at_nsh[iat]; ++is) {
                int ish = s.at_sh0[iat] + is;
                for (int m = 0; m < NSPH[s.sh_l[ish]]; ++m) {
                    size_t a = s.sh_ao0[ish] + m;
                    dEdcn[iat] += -s.sh_kcn[ish] * P[a * nao + a];
                }
            }
            for (int jat = 0; jat < iat; ++jat) {
                double vec[3] = {xyz[3 * iat] - xyz[3 * jat], xyz[3 * iat + 1] - xyz[3 * jat + 1], xyz[3 * iat + 2] - xyz[3 * jat + 2]};
                double r2 = vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2], r = sqrt(r2);
                double radsum = (s.elem[num[iat]].atomic_rad + s.elem[num[jat]].atomic_rad) * TB_AATOAU;
                double rr = sqrt(r / radsum);
                double dG[3] = {0, 0, 0};
                for (int is = 0; is < s.at_nsh[iat]; ++is)
                    for (int js = 0; js < s.at_nsh[jat]; ++js) {
                        int ish = s.at_sh0[iat] + is, jsh = s.at_sh0[jat] + js;
                        shell_pair(&s, jsh, ish, vec, r2, 1, &pi);
                        double pli = 1.0 + s.sh_poly[ish] * rr, plj = 1.0 + s.sh_poly[jsh] * rr;
                        double shpoly = pli * plj;
                        /* d shpoly / d r * (1/r): rr = sqrt(r/R) -> d rr/dr = rr/(2r) */
                        double dshpoly = (s.sh_poly[ish] * plj + s.sh_poly[jsh] * pli) * rr * 0.5 / r2;
                        double hs = gfn2_hscale(&s, ish, jsh);
                        double hav = 0.5 * (s.selfen[ish] + s.selfen[jsh]);
                        double hij = hav * hs * shpoly;
                        int ni = NSPH[s.sh_l[ish]], nj = NSPH[s.sh_l[jsh]];
                        for (int mj = 0; mj < nj; ++mj)
                            for (int mi = 0; mi < ni; ++mi) {
                                int ij = mj * ni + mi;
                                size_t a = s.sh_ao0[jsh] + mj, b = s.sh_ao0[ish] + mi;
                                double pij = P[a * nao + b], wij = W[a * nao + b];
                                double sval = 2.0 * pij * hij - 2.0 * wij - pij * (pot.vao[a] + pot.vao[b]);
                                double di[3] = {pi.D[0][ij], pi.D[1][ij], pi.D[2][ij]}, qi[6], dj[3], qj[6];
                                double ds[3] = {pi.dS[0][ij], pi.dS[1][ij], pi.dS[2][ij]};
                                double ddi[3][3], dqi[3][6], ddj[3][3], dqj[3][6];
                                for (int cc = 0; cc < 6; ++cc) qi[cc] = pi.Q[cc][ij];
                                for (int k = 0; k < 3; ++k) {
                                    for (int cc = 0; cc < 3; ++cc) ddi[k][cc] = pi.dD[k][cc][ij];
                                    for (int cc = 0; cc < 6; ++cc) dqi[k][cc] = pi.dQ[k][cc][ij];
                                }
                                shift_operator(vec, pi.S[ij], di, qi, dj, qj, ds, ddi, dqi, ddj, dqj);
                                for (int k = 0; k < 3; ++k) {
                                    double g = sval * ds[k] + 2.0 * pij * hav * hs * dshpoly * vec[k] * pi.S[ij];
                                    for (int cc = 0; cc < 3; ++cc)
                                        g -= pij * (ddi[k][cc] * pot.vdp[3 * iat + cc] + ddj[k][cc] * pot.vdp[3 * jat + cc]);
                                    for (int cc = 0; cc < 6; ++cc)
                                        g -= pij * (dqi[k][cc] * pot.vqp[6 * iat + cc] + dqj[k][cc] * pot.vqp[6 * jat + cc]);
                                    dG[k] += g;
                                }
                                /* H0 depends on CN through the levels: d hij/d cn = -1/2 kcn hs shpoly */
                                double t = pij * hs * shpoly * pi.S[ij];
                                dEdcn[iat] += -s.sh_kcn[ish] * t;
                                dEdcn[jat] += -s.sh_kcn[jsh] * t;
                            }
                    }
                for (int k = 0; k < 3; ++k) { grad[3 * iat + k] += dG[k]; grad[3 * jat + k] -= dG[k]; }
            }
        }
        /* chain rule through the coordination numbers */
        for (int k = 0; k < nat; ++k)
            for (int i = 0; i < nat; ++i)
                for (int cc = 0; cc < 3; ++cc)
                    grad[3 * k + cc] += s.dcndr[(k * nat + i) * 3 + cc] * dEdcn[i] + s.dcnd4dr[(k * nat + i) * 3 + cc] * dEdcn4[i];
    }

    *energy_out = energy;
    for (int i = 0; i < nat; ++i) qat_out[i] = qat[i];
    for (int i = 0; i < 3 * nat; ++i) grad_out[i] = grad[i];

    if (detail) {
        detail->nsh = nsh; detail->nao = nao; detail->niter = iscf; detail->converged = converged;
        detail->e_rep = e_rep; detail->e_disp_atm = e_atm; detail->e_disp_sc = e_d4; detail->e_el = e_el;
        detail->e_es2 = e_es2; detail->e_es3 = e_es3; detail->e_aes = e_aes; detail->e_ts = ts;
        if (detail->cn) memcpy(detail->cn, s.cn, nat * sizeof(double));
        if (detail->cn_d4) memcpy(detail->cn_d4, s.cnd4, nat * sizeof(double));
        if (detail->overlap) memcpy(detail->overlap, s.S, n2 * sizeof(double));
        if (detail->h0) memcpy(detail->h0, s.H0, n2 * sizeof(double));
        if (detail->dipole) memcpy(detail->dipole, s.D, 3 * n2 * sizeof(double));
        if (detail->quadrupole) memcpy(detail->quadrupole, s.Q, 6 * n2 * sizeof(double));
        if (detail->emo) memcpy(detail->emo, emo, nao * sizeof(double));
        if (detail->focc) memcpy(detail->focc, focc, nao * sizeof(double));
        if (detail->qsh) memcpy(detail->qsh, qsh, nsh * sizeof(double));
        if (detail->dpat) memcpy(detail->dpat, dpat, 3 * nat * sizeof(double));
        if (detail->qpat) memcpy(detail->qpat, qpat, 6 * nat * sizeof(double));
        if (detail->coeff) memcpy(detail->coeff, C, n2 * sizeof(double));
        if (detail->ao2at) for (int a = 0; a < nao; ++a) detail->ao2at[a] = s.ao_at[a];
        {
            int homo = (int)floor(nel[0]);
            if (fmod(nel[0], 1.0) > 0.5) homo += 1;
            detail->ihomo = homo > 1 ? homo : 1;
        }
    }

    free(s.at_sh0); free(s.at_nsh); free(s.sh_at); free(s.sh_l); free(s.sh_ao0); free(s.sh_np); free(s.ao_at); free(s.ao_sh);
    free(s.sh_alpha); free(s.sh_coef); free(s.sh_level); free(s.sh_kcn); free(s.sh_poly); free(s.sh_refocc);
    free(s.sh_hub); free(s.sh_gam3); free(s.sh_zeta); free(s.sh_val); free(s.at_gam3); free(s.cn); free(s.dcndr); free(s.cnd4); free(s.dcnd4dr);
    free(s.S); free(s.H0); free(s.D); free(s.Q); free(s.selfen); free(grad); free(dEdcn); free(dEdcn4);
    free(c.gamma); free(c.mrad); free(c.dmrdcn); free(c.sd); free(c.dd); free(c.sq); free(c.c6ref); free(c.dispmat);
    free(c.gw); free(c.gwdcn); free(c.gwdq);
    free(qsh); free(qat); free(dpat); free(qpat); free(pot.vsh); free(pot.vat); free(pot.vdp); free(pot.vqp); free(pot.vao);
    free(H1); free(C); free(P); free(Linv); free(work); free(emo); free(focc); free(ftmp);
    free(mix.q_in); free(mix.qlast_in); free(mix.dq); free(mix.dqlast); free(mix.df); free(mix.u); free(mix.a); free(mix.omega);
    return stat;
}

/* reference src/mo_energ.f90:31-54 (write_qmo): qmo(n, ia) = sum_{j on ia} sum_k c(j,n) c(k,n) S(j,k) + 1e-10, then every orbital
 * is normalised over the atoms */
void xtb_oracle_qmo(int nat, int nao, const int32_t *ao2at, const double *coeff, const double *overlap, double *qmo) {
    for (int ia = 0; ia < nat; ++ia)
        for (int n = 0; n < nao; ++n) {
            double qmo_sum = 0.0;
            for (int j = 0; j < nao; ++j)
                if (ao2at[j] == ia)
                    for (int k = 0; k < nao; ++k) qmo_sum = qmo_sum + coeff[(size_t)j * nao + n] * coeff[(size_t)k * nao + n] * overlap[(size_t)j * nao + k];
            qmo_sum = qmo_sum + 1.e-10;
            qmo[(size_t)n * nat + ia] = qmo_sum;
        }
    for (int k = 0; k < nao; ++k) {
        double summa = 0.0;
        for (int j = 0; j < nat; ++j) summa = summa + qmo[(size_t)k * nat + j];
        for (int j = 0; j < nat; ++j) qmo[(size_t)k * nat + j] = qmo[(size_t)k * nat + j] / summa;
    }
}
