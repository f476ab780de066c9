/* TEST INFRASTRUCTURE -- CPU restatement ("oracle") of the QCxMS production-trajectory
 * hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (qcxms_b200/) never does.
 *
 * PARITY UNPINNED: the arithmetic of get_xtb_egrad lives in tblite v0.2.1, an
 * un-vendored dependency of the reference (subprojects/tblite.wrap:1-4) that is not on
 * this machine and has no golden vectors in the reference tree.  This
 * file restates the published GFN2-xTB method and tblite's call protocol as recorded
 * at the reference's call sites (src/tblite.f90:95-151).
 */
#ifndef XTB_ORACLE_H
#define XTB_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* optional detailed output for tests (any pointer may be NULL) */
typedef struct {
    int nsh, nao, niter, converged;
    double e_rep, e_disp_atm, e_disp_sc, e_el, e_es2, e_es3, e_aes, e_ts;
    double *cn;       /* [nat] GFN coordination number */
    double *cn_d4;    /* [nat] D4 covalent coordination number */
    double *overlap;  /* [nao*nao] */
    double *h0;       /* [nao*nao] */
    double *dipole;   /* [3*nao*nao]  component-major */
    double *quadrupole; /* [6*nao*nao] */
    double *emo;      /* [nao] */
    double *focc;     /* [nao] */
    double *qsh;      /* [nsh] */
    double *dpat;     /* [3*nat] atom-major */
    double *qpat;     /* [6*nat] atom-major */
    double *e_iter;   /* [250] electronic energy per SCC iteration */
    double *coeff;    /* [nao*nao] eigenvectors, coeff[a*nao + k] = AO a of orbital k (orbitals in ascending energy) */
    int32_t *ao2at;   /* [nao] atom (0-based) of every AO */
    int ihomo;        /* max(homo of the alpha channel, 1), 1-based (reference src/tblite.f90:160) */
} xtb_oracle_detail_t;

/* Restates get_xtb_egrad (reference src/tblite.f90:65-175): cold-start SCC with
 * accuracy 1.0, kt = etemp*ktoau, uhf = min(multiplicity-1, 0).
 * xyz [nat][3] bohr; gradient [nat][3] Eh/bohr; returns stat (0 ok, -1 fatal, 5 unknown method). */
int xtb_oracle_egrad(int nat, const int32_t *num, const double *xyz, int charge, int multiplicity,
                     int method_id, double etemp, double *qat, double *energy, double *gradient,
                     xtb_oracle_detail_t *detail);

/* write_qmo without the files (reference src/mo_energ.f90:31-54): Mulliken population of every orbital on every atom,
 * + 1e-10, normalised per orbital.  coeff/overlap as in the detail struct; qmo [nao][nat]. */
void xtb_oracle_qmo(int nat, int nao, const int32_t *ao2at, const double *coeff, const double *overlap, double *qmo);

/* test hook: scale the SCC thresholds (reference value 1.0, src/tblite.f90:46) */
void xtb_oracle_set_accuracy(double acc);

/* basis dimensions for a composition (returns 0 on success) */
int xtb_oracle_dims(int nat, const int32_t *num, int method_id, int *nsh, int *nao);

/* dense symmetric eigensolver used by the oracle (exposed for cross-checks against LAPACK):
 * a [n*n] symmetric in, eigenvectors (columns, row-major a[i*n+k] = component i of vector k) out */
int xtb_oracle_syev(int n, double *a, double *w);

#ifdef __cplusplus
}
#endif
#endif
