/* GFN1-xTB element and global parameters (method id 1, reference src/tblite.f90:124-126 new_gfn1_calculator).
 *
 * PROVENANCE / STATUS -- UNVERIFIED, and weaker than the GFN2 table: the numbers live in tblite v0.2.1
 * src/tblite/xtb/gfn1.f90 (un-vendored, reference subprojects/tblite.wrap:1-4) and in the published
 * param_gfn1-xtb.txt (Grimme, Bannwarth, Shushkov, JCTC 13, 1989 (2017)); neither is on this machine.  The built-in
 * table is a reconstruction from memory for H, C, N, O, F; the Cl row holds values set BY ANALOGY (marked below) so
 * that the halogen-bond term can be exercised at all.  Nothing in the reference tree can falsify these digits
 * (its GFN1 example geometry is not a stored GFN1 result), so every GFN1 number produced with the built-in table is
 * "method structure verified by self-consistency tests, parameters unverified".  A published parameter file can be
 * dropped in without rebuilding: set QCXMS_B200_GFN1_PARAM to an xtb-format file (blocks "$Z= n" with ao=, lev=,
 * exp=, EN=, GAM=, GAM3=, REPA=, REPB=, POLYS=/POLYP=/POLYD=, LPARP=/LPARD=, CXB=); gfn1_load_param_file below reads
 * it into the same table for the oracle and the CUDA host model alike.
 *
 * Model restated: H gets a second, diffuse s shell (Schmidt-orthogonalised to 1s); exponential
 * D3-type CN (k = 16, radii x 4/3); H0 = 1/2 K (h_i + h_j) S Pi(R) with h = level (1 + kcn_l CN), K = kpair k_ll' (1 + kEN dEN^2)
 * for valence pairs, (k_ll + kdiff)/2 for valence/diffuse, kdiff for diffuse/diffuse; shell-resolved second order with
 * harmonically averaged hardnesses (gexp 2), atom-resolved third order; repulsion with kexp 1.5 for every pair; D3(BJ)
 * (in-tree model src/dftd3.f90, data include/pars.fh) with a1 0.63, a2 5.0, s8 2.4, no three-body term; halogen-bond
 * correction (xbrad 1.3, xbdamp 0.44); no multipole electrostatics.
 */
#pragma once
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gfn2_params.h" /* gfn2_elem_t: the element record both methods share */

#define GFN1_MAXZ 18
#define GFN1_KDIAG_S 1.85
#define GFN1_KDIAG_P 2.25
#define GFN1_KDIAG_D 2.00
#define GFN1_K_SP 2.08
#define GFN1_KDIFF 2.85
#define GFN1_ENSCALE (-0.007)
#define GFN1_REP_KEXP 1.5
#define GFN1_D3_S6 1.0
#define GFN1_D3_S8 2.4
#define GFN1_D3_A1 0.63
#define GFN1_D3_A2 5.0
#define GFN1_XB_DAMP 0.44
#define GFN1_XB_RAD 1.3
/* CN dependence of the levels by angular momentum: h = level (1 + kcn_l CN) */
static const double GFN1_KCN_L[3] = {0.006, -0.003, -0.005};

/* extra per-element data GFN1 needs beyond gfn2_elem_t */
typedef struct {
    int supported;   /* 0: no built-in parameters for this element */
    int valence[3];  /* 0 marks a diffuse shell (H 2s) */
    double xbond;    /* halogen-bond strength (Eh), 0: none */
} gfn1_extra_t;

/* field use for GFN1: kcn[] unused (global by l), hubbard_deriv = atomic third-order parameter, shell_hubbard[l] = 1 + lpar_l,
 * dkernel/qkernel/mp_* unused */
static gfn2_elem_t GFN1_ELEM[GFN1_MAXZ + 1] = {
    {0},
    /* H  */ {2, {0, 0, 0}, {1, 2, 0}, {4, 3, 0}, {1.0, 0.0, 0.0}, {-10.923452, -2.171902, 0.0},
              {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {1.207940, 1.993207, 0.0}, 0.470099,
              {1.0, 1.0, 1.0}, 0.0, 2.209700, 1.116244, 2.20, 0.32, 0.0, 0.0, 0.0, 0.0},
    {0}, {0}, {0}, {0},
    /* C  */ {2, {0, 1, 0}, {2, 2, 0}, {6, 6, 0}, {2.0, 2.0, 0.0}, {-13.587210, -10.052785, 0.0},
              {0.0, 0.0, 0.0}, {-0.07082170, 0.00812216, 0.0}, {1.960324, 1.832096, 0.0}, 0.479988,
              {1.0, 0.9461985, 1.0}, 0.1500000, 1.281954, 4.428763, 2.55, 0.75, 0.0, 0.0, 0.0, 0.0},
    /* N  */ {2, {0, 1, 0}, {2, 2, 0}, {6, 6, 0}, {2.0, 3.0, 0.0}, {-20.058000, -12.889326, 0.0},
              {0.0, 0.0, 0.0}, {-0.12745585, -0.01428367, 0.0}, {2.050067, 2.113682, 0.0}, 0.476106,
              {1.0, 1.0461493, 1.0}, -0.0639780, 1.727773, 5.498808, 3.04, 0.71, 0.0, 0.0, 0.0, 0.0},
    /* O  */ {2, {0, 1, 0}, {2, 2, 0}, {6, 6, 0}, {2.0, 4.0, 0.0}, {-23.398376, -17.886554, 0.0},
              {0.0, 0.0, 0.0}, {-0.13729047, -0.04453341, 0.0}, {2.345365, 2.153060, 0.0}, 0.583349,
              {1.0, 1.0451896, 1.0}, -0.0517134, 2.004253, 5.171786, 3.44, 0.64, 0.0, 0.0, 0.0, 0.0},
    /* F  */ {2, {0, 1, 0}, {2, 2, 0}, {6, 6, 0}, {2.0, 5.0, 0.0}, {-24.776175, -17.274415, 0.0},
              {0.0, 0.0, 0.0}, {-0.03921613, -0.11422491, 0.0}, {2.421394, 2.321971, 0.0}, 0.788194,
              {1.0, 1.0531518, 1.0}, 0.1426212, 2.507078, 6.931741, 3.98, 0.60, 0.0, 0.0, 0.0, 0.0},
    {0}, {0}, {0}, {0}, {0}, {0}, {0},
    /* Cl: BY ANALOGY (not recalled): levels / exponents / hardness patterned on the neighbouring rows; see header */
    /* Cl */ {3, {0, 1, 2}, {3, 3, 3}, {6, 6, 4}, {2.0, 5.0, 0.0}, {-24.452163, -12.883714, -1.670300},
              {0.0, 0.0, 0.0}, {-0.16562004, -0.06986430, 0.38045622}, {2.522150, 2.196703, 1.683125}, 0.366300,
              {1.0, 1.0489400, 0.9000000}, 0.1495483, 1.577144, 17.353134, 3.16, 1.00, 0.0, 0.0, 0.0, 0.0},
    {0},
};

static gfn1_extra_t GFN1_EXTRA[GFN1_MAXZ + 1] = {
    {0, {0, 0, 0}, 0.0},
    {1, {1, 0, 0}, 0.0},                                                                       /* H: 2s is diffuse */
    {0, {0, 0, 0}, 0.0}, {0, {0, 0, 0}, 0.0}, {0, {0, 0, 0}, 0.0}, {0, {0, 0, 0}, 0.0},
    {1, {1, 1, 0}, 0.0}, {1, {1, 1, 0}, 0.0}, {1, {1, 1, 0}, 0.0}, {1, {1, 1, 0}, 0.0},     /* C N O F */
    {0, {0, 0, 0}, 0.0}, {0, {0, 0, 0}, 0.0}, {0, {0, 0, 0}, 0.0}, {0, {0, 0, 0}, 0.0}, {0, {0, 0, 0}, 0.0},
    {0, {0, 0, 0}, 0.0}, {0, {0, 0, 0}, 0.0},
    {1, {1, 1, 1}, 0.0381742},                                                                 /* Cl (by analogy) */
    {0, {0, 0, 0}, 0.0},
};

/* pair scaling of H0 (kpair); 1 unless listed */
static inline double gfn1_kpair(int zi, int zj) {
    if (zi > zj) { int t = zi; zi = zj; zj = t; }
    if (zi == 1 && zj == 1) return 0.96;
    if (zi == 1 && zj == 5) return 0.95;
    if (zi == 1 && zj == 7) return 1.04;
    return 1.0;
}

/* halogen-bond acceptors (N, O, P, S) and donors (Cl; Br, I, At are beyond the H..Ar range) */
static inline int gfn1_xb_acceptor(int z) { return z == 7 || z == 8 || z == 15 || z == 16; }
static inline int gfn1_xb_donor(int z) { return z <= GFN1_MAXZ && GFN1_EXTRA[z].xbond != 0.0; }

/* Reads an xtb-format GFN1 parameter file (see header) over the built-in table.  Returns the number of element blocks read,
 * -1 if the file cannot be opened.  Not thread-safe: call once before the first GFN1 evaluation (gfn1_ensure_loaded). */
static int gfn1_load_param_file(const char *path) {
    FILE *fp = fopen(path, "r");
    if (!fp) return -1;
    char line[512];
    int z = 0, nread = 0;
    while (fgets(line, sizeof line, fp)) {
        char *p = line;
        while (*p && isspace((unsigned char)*p)) ++p;
        if (!strncmp(p, "$Z=", 3) || !strncmp(p, "$z=", 3)) {
            z = atoi(p + 3);
            if (z >= 1 && z <= GFN1_MAXZ) {
                memset(&GFN1_ELEM[z], 0, sizeof(gfn2_elem_t));
                GFN1_ELEM[z].shell_hubbard[0] = GFN1_ELEM[z].shell_hubbard[1] = GFN1_ELEM[z].shell_hubbard[2] = 1.0;
                GFN1_ELEM[z].en = GFN2_ELEM[z].en;
                GFN1_ELEM[z].atomic_rad = GFN2_ELEM[z].atomic_rad;
                memset(&GFN1_EXTRA[z], 0, sizeof(gfn1_extra_t));
                GFN1_EXTRA[z].supported = 1;
                ++nread;
            } else
                z = 0;
            continue;
        }
        if (*p == '$') { z = 0; continue; }
        if (!z) continue;
        char *eq = strchr(p, '=');
        if (!eq) continue;
        *eq = 0;
        char key[32];
        size_t k = 0;
        for (const char *q = p; *q && !isspace((unsigned char)*q) && k + 1 < sizeof key; ++q) key[k++] = (char)tolower((unsigned char)*q);
        key[k] = 0;
        char *val = eq + 1;
        gfn2_elem_t *e = &GFN1_ELEM[z];
        if (!strcmp(key, "ao")) {
            int ns = 0, seen[3] = {0, 0, 0};
            for (char *q = val; q[0] && q[1] && ns < GFN2_MAXSH; ++q) {
                if (!isdigit((unsigned char)q[0])) continue;
                int l = q[1] == 's' ? 0 : q[1] == 'p' ? 1 : q[1] == 'd' ? 2 : -1;
                if (l < 0) continue;
                e->pqn[ns] = q[0] - '0';
                e->ang[ns] = l;
                e->nprim[ns] = (z <= 2) ? (seen[l] ? 3 : 4) : (l == 2 ? 4 : 6);
                GFN1_EXTRA[z].valence[ns] = !seen[l];
                seen[l] = 1;
                ++ns;
                ++q;
            }
            e->nshell = ns;
            /* neutral-atom aufbau reference occupations */
            int nel = z <= 2 ? z : z <= 10 ? z - 2 : z - 10;
            for (int i = 0; i < ns; ++i) {
                if (!GFN1_EXTRA[z].valence[i] || e->ang[i] == 2) { e->refocc[i] = 0.0; continue; }
                if (e->ang[i] == 0) { e->refocc[i] = nel < 2 ? nel : 2; }
                else e->refocc[i] = nel > 2 ? nel - 2 : 0;
            }
        } else if (!strcmp(key, "lev") || !strcmp(key, "exp")) {
            double *dst = !strcmp(key, "lev") ? e->selfenergy : e->slater;
            char *end = val;
            for (int i = 0; i < GFN2_MAXSH; ++i) { double v = strtod(end, &end); dst[i] = v; }
        } else if (!strcmp(key, "en")) e->en = atof(val);
        else if (!strcmp(key, "gam")) e->hubbard = atof(val);
        else if (!strcmp(key, "gam3")) e->hubbard_deriv = 0.1 * atof(val);
        else if (!strcmp(key, "repa")) e->rep_alpha = atof(val);
        else if (!strcmp(key, "repb")) e->rep_zeff = atof(val);
        else if (!strcmp(key, "polys") || !strcmp(key, "polyp") || !strcmp(key, "polyd")) {
            int l = key[4] == 's' ? 0 : key[4] == 'p' ? 1 : 2;
            for (int i = 0; i < e->nshell; ++i) if (e->ang[i] == l && GFN1_EXTRA[z].valence[i]) e->shpoly[i] = 0.01 * atof(val);
        } else if (!strcmp(key, "lparp")) e->shell_hubbard[1] = 1.0 + 0.1 * atof(val);
        else if (!strcmp(key, "lpard")) e->shell_hubbard[2] = 1.0 + 0.1 * atof(val);
        else if (!strcmp(key, "cxb") || !strcmp(key, "xbond")) GFN1_EXTRA[z].xbond = 0.1 * atof(val);
    }
    fclose(fp);
    return nread;
}

static inline void gfn1_ensure_loaded(void) {
    static int done = 0;
    if (done) return;
    done = 1;
    const char *path = getenv("QCXMS_B200_GFN1_PARAM");
    if (path && *path && gfn1_load_param_file(path) < 0) fprintf(stderr, "qcxms_b200: cannot read QCXMS_B200_GFN1_PARAM=%s; built-in GFN1 table in use\n", path);
}
