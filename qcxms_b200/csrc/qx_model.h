// Per-composition model tables of the GFN2-xTB ensemble kernels (host-built, device-resident).
// One DevModel describes ONE molecule composition; every trajectory of an ensemble shares it.
#pragma once
#include <cstdint>

#define QX_MAXPRIM 8   // GFN1: the orthogonalised H 2s carries the 1s primitives as well (3 + 4)
#define QX_MAXREF 7
#define QX_MAX_ITER 250   // tblite max_iter (SCC cycles and Broyden memory)
#ifndef QX_NT
#define QX_NT 320         // threads per CTA; one CTA == one trajectory (tu_*.cu may be built for other widths, see qx_kernels.h)
#endif
#ifndef QX_MINB
#define QX_MINB 2         // resident CTAs per SM the kernels are compiled for (register budget)
#endif
#ifndef QX_VARIANT
#define QX_VARIANT nt320
#endif

struct DevModel {
    int nat, nsh, nao, ntype, ld, ndim;  // ld: leading dimension of the shared-memory matrices (== 4 or 12 mod 16)
    double oa_kappa;                      // pairs with |coupling| > kappa |gap| go into exactly diagonalised clusters
    double oa_stop;                       // convergence threshold of the refinement (max |E_ij| of the last pass)
    int oa;                               // 1: GEMM-based eigenpair refinement (qx_oa.cuh) with three more shared-memory matrices (wide-CTA kernels)
    int method;                           // 2: GFN2-xTB, 1: GFN1-xTB (exp CN, D3(BJ), halogen bond, atomic third order, no multipoles)
    int polish;                           // 1: the last Jacobi sweep of a decomposition runs as three DMMA products (jacobi_polish; scratch at ScratchLayout::P)
    int jblock;                           // global-slab mode: rows per block of the shared-memory blocked Jacobi (0: none)
    int extras_off;                       // offset (doubles) of the MD/CID kernels' per-trajectory vectors in the CTA's shared memory
    int mat_in_global;                    // 1: the two SCC matrices do not fit shared memory and live in the per-CTA global slab
    int rows8;                            // rows of the shared-memory matrices (zero padded; multiple of 8 when the strip GEMMs apply)
    int ntask_int, ntask_grad;
    double nel[2];                        // alpha / beta electron numbers
    // per atom
    const int *num, *type, *at_sh0, *at_nsh, *at_ao0, *at_nao;
    const double *at_rcov, *at_rad, *at_repa, *at_repz, *at_en, *at_mprad, *at_mpvcn, *at_dk, *at_qk, *at_r4r2,
        *at_zeff, *at_gam, *at_qcrad, *mass;
    const int *at_nref;
    const double *at_refcn, *at_refq;  // [nat][QX_MAXREF]
    const int *at_ngw;                 // [nat][QX_MAXREF]
    // per shell
    const int *sh_at, *sh_l, *sh_ao0, *sh_np;
    const double *sh_alpha, *sh_coef;  // [nsh][QX_MAXPRIM]
    const double *sh_level, *sh_kcn, *sh_poly, *sh_refocc, *sh_hub, *sh_gam3;
    const double *hscale;              // [nsh][nsh]
    // per AO
    const int *ao_at, *ao_sh, *ao_m;
    // type-pair D4 reference C6 [ntype][ntype][QX_MAXREF][QX_MAXREF]
    const double *c6ref;
    // GFN1: atomic third-order parameter, halogen-bond strength (0: no donor), D3 sqrt(0.5 r4/r2 sqrt(Z)), number of D3 reference
    // systems per atom and the type-pair D3 reference table [ntype][ntype][5][5][3] = (C6, CN_ref(i), CN_ref(j)) (-1: none)
    const double *at_gam3, *at_xb, *at_r2r4d3;
    const int *at_mxc;
    const double *d3ref;
    // work lists: AO pairs (a on the bra atom, b on the ket atom; atom(a) <= atom(b))
    const int2 *task_int;   // all pairs incl. on-site ordered pairs
    // gradient reduction lists (CSR by atom over off-site tasks; sign +1 ket atom / -1 bra atom)
    const int *gr_ptr, *gr_task;
    // impactscale grid: scal_table[k] = sum of k additions of 0.0002f in double (reference src/impact.f90:37), k = 0..20000
    const double *scal_table;
};

// per-CTA scratch in global memory (stays L2 resident); offsets in doubles
struct ScratchLayout {
    size_t S, H0, Dt, Qt, T, P, W, matA, matC, gamma, dcnp, dcnp4, edisp, c6, dc6, taskout, br_df, br_u, br_a, br_vec, total;
};
