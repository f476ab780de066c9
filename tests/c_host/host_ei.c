/* A host program in plain C that drives the C ABI the way the Fortran shim would (shim/qcxms_b200_shim.f90): it reads one
 * trajectory directory of the reference (start.xyz + qcxms.start, formats of src/utility.f90:363-422), runs md() through
 * qcxms_b200_ensemble_* and prints what the reference's md() hands back.  Test infrastructure (tests/test_c_host.py).
 *
 *   gcc -O2 -I include -o host_ei tests/c_host/host_ei.c -L qcxms_b200 -lqcxms_b200 -Wl,-rpath,$PWD/qcxms_b200 -lm
 *   host_ei <dir> <nmax>
 */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qcxms_b200.h"

static const char *SYM[] = {"", "h", "he", "li", "be", "b", "c", "n", "o", "f", "ne", "na", "mg", "al", "si", "p", "s", "cl", "ar"};
/* average atomic masses (amu) of the reference (src/atomic_masses.f90) for the elements of its examples */
static const double MASS[] = {0, 1.00794075, 4.00260193, 6.94003660, 9.01218307, 10.81102805, 12.0107359, 14.00670321, 15.99940492,
                              18.99840316, 20.18004638, 22.98976928, 24.30505162, 26.98153853, 28.08549871, 30.97376200, 32.06478741,
                              35.45293758, 39.94779856};
#define AMUTOAU 1822.8884850003578
#define AATOAU (1.0 / 0.52917726)

static double fortran_number(const char *field, int w) {
    char buf[64];
    int n = w < 63 ? w : 63;
    memcpy(buf, field, n);
    buf[n] = 0;
    for (char *p = buf; *p; ++p)
        if (*p == 'D' || *p == 'd') *p = 'E';
    return atof(buf);
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s <dir> <nmax>\n", argv[0]); return 2; }
    char path[1024], line[512];
    snprintf(path, sizeof path, "%s/start.xyz", argv[1]);
    FILE *fp = fopen(path, "r");
    if (!fp) { fprintf(stderr, " - Missing start.xyz file! -\n"); return 1; }
    int nat = 0;
    if (!fgets(line, sizeof line, fp) || sscanf(line, "%d", &nat) != 1 || nat < 1) return 1;
    if (!fgets(line, sizeof line, fp)) return 1;
    int32_t *num = malloc(nat * sizeof *num);
    double *xyz = malloc(3 * nat * sizeof *xyz), *velo = malloc(3 * nat * sizeof *velo), *velof = malloc(nat * sizeof *velof);
    double *mass = malloc(nat * sizeof *mass);
    for (int i = 0; i < nat; ++i) {
        char el[8];
        if (!fgets(line, sizeof line, fp) || sscanf(line, "%7s %lf %lf %lf", el, &xyz[3 * i], &xyz[3 * i + 1], &xyz[3 * i + 2]) != 4) return 1;
        for (char *p = el; *p; ++p) *p = (char)tolower((unsigned char)*p);
        num[i] = 0;
        for (int z = 1; z <= 18; ++z)
            if (!strcmp(el, SYM[z])) num[i] = z;
        if (!num[i]) { fprintf(stderr, "unknown element %s\n", el); return 1; }
        for (int c = 0; c < 3; ++c) xyz[3 * i + c] *= AATOAU;
        mass[i] = MASS[num[i]] * AMUTOAU;
    }
    fclose(fp);
    snprintf(path, sizeof path, "%s/qcxms.start", argv[1]);
    fp = fopen(path, "r");
    if (!fp) { fprintf(stderr, " - Missing qcxms.start file! -\n"); return 1; }
    int itrj = 0;
    double eimp, tadd;
    if (!fgets(line, sizeof line, fp) || sscanf(line, "%d", &itrj) != 1) return 1;
    if (!fgets(line, sizeof line, fp)) return 1;
    eimp = fortran_number(line, 22);
    if (!fgets(line, sizeof line, fp)) return 1;
    tadd = fortran_number(line, 22);
    for (int i = 0; i < nat; ++i) {
        if (!fgets(line, sizeof line, fp) || strlen(line) < 88) return 1;
        for (int c = 0; c < 3; ++c) velo[3 * i + c] = fortran_number(line + 22 * c, 22);
        velof[i] = fortran_number(line + 66, 22);
    }
    fclose(fp);

    qcxms_b200_md_config_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.method_id = QCXMS_B200_GFN2;
    cfg.mchrg = 1; cfg.nfragexit = 3; cfg.exit_rules = 1; cfg.nmax = atoi(argv[2]); cfg.isec = 1;
    cfg.tstep = 0.5 * 41.3413733365614;   /* tstep 0.5 fs */
    cfg.etemp_in = -1.0;
    qcxms_b200_ensemble_t *ens = NULL;
    int rc = qcxms_b200_ensemble_create(&cfg, 1, nat, num, mass, 0, &ens);
    if (!rc) rc = qcxms_b200_ensemble_set_trajectory(ens, 0, xyz, velo, velof, eimp, tadd);
    int64_t steps = 0;
    if (!rc) rc = qcxms_b200_ensemble_run_md(ens, 0, &steps);
    double *grad = malloc(3 * nat * sizeof *grad), *achrg = malloc(nat * sizeof *achrg), *axyz = malloc(3 * nat * sizeof *axyz);
    int32_t *list = malloc(nat * sizeof *list);
    qcxms_b200_md_result_t res;
    if (!rc) rc = qcxms_b200_ensemble_get_result(ens, 0, xyz, velo, grad, list, achrg, axyz, &res);
    if (rc) { fprintf(stderr, "qcxms_b200 error %d: %s\n", rc, qcxms_b200_last_error()); return 1; }
    printf("itrj %d nat %d steps %lld nstep %d mdok %d fragstate %d nfrag %d\n", itrj, nat, (long long)steps, res.nstep, res.mdok, res.fragstate, res.nfrag);
    printf("Epot %.14e Ekin %.14e\n", res.Epot, res.Ekin);
    for (int i = 0; i < nat; ++i)
        printf("%d %d %.14e %.14e %.14e %.10f\n", num[i], list[i], xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], achrg[i]);
    qcxms_b200_ensemble_destroy(ens);
    return 0;
}
