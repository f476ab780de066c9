// Device side of the GFN2-xTB energy/gradient evaluation: one CTA (QX_NT threads) per
// trajectory, matrices of the SCC staged in shared memory, integrals in a per-CTA
// global scratch slab that stays L2-resident.
//
// Replaces (for the batched ensemble) what the reference obtains from tblite through
// get_xtb_egrad (reference src/tblite.f90:65-175; call protocol :111,:123,:133,:136):
// zeroed wavefunction on every call, accuracy 1.0, Broyden-mixed SCC with Fermi smearing.
// Every summation runs in a fixed order (no floating-point atomics) so that a trajectory
// is bitwise reproducible from run to run.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "params/gfn2_params.h"
#include "params/gfn1_params.h"
#include "qx_model.h"

namespace qx {

#define QX_PI 3.14159265358979323846264338327950288
#define QX_SQRT3 1.7320508075688772935

struct SphTerm {
    int n;
    int ex[3][3];
    double c[3];
};
// real solid harmonics, index l*l + (m+l), order m = -l..l (p: y,z,x)
static __constant__ SphTerm c_sph[9] = {
    {1, {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, {1.0, 0, 0}},
    {1, {{0, 1, 0}, {0, 0, 0}, {0, 0, 0}}, {1.0, 0, 0}},
    {1, {{0, 0, 1}, {0, 0, 0}, {0, 0, 0}}, {1.0, 0, 0}},
    {1, {{1, 0, 0}, {0, 0, 0}, {0, 0, 0}}, {1.0, 0, 0}},
    {1, {{1, 1, 0}, {0, 0, 0}, {0, 0, 0}}, {QX_SQRT3, 0, 0}},
    {1, {{0, 1, 1}, {0, 0, 0}, {0, 0, 0}}, {QX_SQRT3, 0, 0}},
    {3, {{0, 0, 2}, {2, 0, 0}, {0, 2, 0}}, {1.0, -0.5, -0.5}},
    {1, {{1, 0, 1}, {0, 0, 0}, {0, 0, 0}}, {QX_SQRT3, 0, 0}},
    {2, {{2, 0, 0}, {0, 2, 0}, {0, 0, 0}}, {0.5 * QX_SQRT3, -0.5 * QX_SQRT3, 0}}};

// quadrupole component index pairs, order xx,xy,yy,xz,yz,zz
static __constant__ int c_qa[6] = {0, 0, 1, 0, 1, 2};
static __constant__ int c_qb[6] = {0, 1, 1, 2, 2, 2};
static __constant__ double c_qscale[6] = {1.0, 2.0, 1.0, 2.0, 2.0, 1.0};

// A pointer that is KNOWN to point into shared memory.  The phase functions are __noinline__ and take the CTA's working set
// through `Sm &`, so plain `double *` members reach them as generic pointers and every access becomes a generic LD/ST (ncu:
// long-scoreboard stalls on what should be LDS).  Keeping the 32-bit shared-window offset instead lets the compiler emit
// LDS/STS: cvta.shared->generic feeding a load is folded into ld.shared.
struct SmPtr {
    unsigned off;
    __device__ __forceinline__ SmPtr &operator=(double *p) { off = (unsigned)__cvta_generic_to_shared(p); return *this; }
    __device__ __forceinline__ double *ptr() const { return reinterpret_cast<double *>(__cvta_shared_to_generic(off)); }
    __device__ __forceinline__ double &operator[](int i) const { return *reinterpret_cast<double *>(__cvta_shared_to_generic(off + 8u * (unsigned)i)); }
    __device__ __forceinline__ double *operator+(int i) const { return reinterpret_cast<double *>(__cvta_shared_to_generic(off + 8u * (unsigned)i)); }
    __device__ __forceinline__ operator double *() const { return ptr(); }
};
#define QX_ASSUME_SHARED(p) __builtin_assume(__isShared(p))

#define QX_BSOL_N 13                              // largest Broyden system solved in shared memory
#define QX_BSOL (QX_BSOL_N * (QX_BSOL_N + 1) + 2)

struct Sm {
    double *A, *C;   // shared memory, or the CTA's global slab when the basis is too large (DevModel::mat_in_global)
    double *X3, *X4, *S5;   // DevModel::oa: work matrices of the eigenpair refinement (C^T H C / E, C^T S C) and the overlap, shared memory
    double *jblk;    // global-slab mode: shared-memory buffer of 2 * DevModel::jblock rows for the blocked Jacobi
    SmPtr xyz, cn, cn4, mrad, dmr, qat, vat, dpat, vdp, qpat, vqp;
    SmPtr qsh, vsh, selfen, vao, emo, focc, gw, gwd, dEdcn, dEdcn4, grad, red, jw, bsol, pop, d4u;
};

// doubles of the phase-local tail (bsol, pop, d4u) the global-slab mode's block buffer may overlay
__host__ __device__ inline size_t smem_tail_doubles(int nat, int nao, int ntype) { return QX_BSOL + 11 * (size_t)nao + (nao & 1) + 7 * (size_t)nat * ntype; }

__host__ __device__ inline size_t smem_doubles(int nat, int nsh, int nao, int ld, int rows8, int mat_in_global, int ntype, int oa = 0) {
    return (mat_in_global ? 0 : (oa ? 5 : 2) * (size_t)rows8 * ld) + 3 * nat + 6 * nat /*cn cn4 mrad dmr qat vat*/ + 6 * nat /*dpat vdp*/ + 12 * nat /*qpat vqp*/
           + 3 * nsh + 3 * nao + 8 + 14 * nat /*gw gwd*/ + 2 * nat + 3 * nat /*grad*/ + 64 + 3 * nao + 8 /*jw*/ + QX_BSOL /*Broyden solve*/ + 11 * nao + (nao & 1) /*pop*/ + 7 * nat * ntype /*d4u*/;
}

__device__ inline void carve(const DevModel &m, double *base, Sm &s, double *gmat = nullptr) {
    int nat = m.nat, nsh = m.nsh, nao = m.nao;
    double *p = base;
    if (m.mat_in_global) {   // large basis: matrices in the CTA's global slab (functional fallback, not the fast path)
        s.A = gmat; s.C = gmat + (size_t)m.rows8 * m.ld;
    } else {
        s.A = p; p += (size_t)m.rows8 * m.ld;
        s.C = p; p += (size_t)m.rows8 * m.ld;
        s.X3 = s.X4 = s.S5 = nullptr;
        if (m.oa) {
            s.X3 = p; p += (size_t)m.rows8 * m.ld;
            s.X4 = p; p += (size_t)m.rows8 * m.ld;
            s.S5 = p; p += (size_t)m.rows8 * m.ld;
        }
    }
    s.xyz = p; p += 3 * nat;
    s.cn = p; p += nat; s.cn4 = p; p += nat; s.mrad = p; p += nat; s.dmr = p; p += nat;
    s.qat = p; p += nat; s.vat = p; p += nat;
    s.dpat = p; p += 3 * nat; s.vdp = p; p += 3 * nat;
    s.qpat = p; p += 6 * nat; s.vqp = p; p += 6 * nat;
    s.qsh = p; p += nsh; s.vsh = p; p += nsh; s.selfen = p; p += nsh;
    s.vao = p; p += nao; s.emo = p; p += nao; s.focc = p; p += nao + 8;  // focc is read up to the padded dimension
    s.gw = p; p += 7 * nat; s.gwd = p; p += 7 * nat;
    s.dEdcn = p; p += nat; s.dEdcn4 = p; p += nat;
    s.grad = p; p += 3 * nat;
    s.red = p; p += 64;
    s.jw = p; p += 3 * nao + 8;
    s.bsol = p; p += QX_BSOL;
    s.pop = p; p += 11 * nao + (nao & 1);
    s.d4u = p;
    // global-slab mode: the block buffer of the blocked Jacobi (and the staging buffer of the staged GEMMs) starts ON TOP of the
    // phase-local scratch vectors bsol/pop/d4u -- none of them is live while the eigensolver or a GEMM runs -- and extends to the
    // MD/CID kernels' vectors at DevModel::extras_off
    s.jblk = m.jblock > 0 ? base + (((size_t)(s.bsol.ptr() - base) + 1) & ~(size_t)1) : nullptr;
}

__device__ inline double block_sum(double v, double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < QX_NT / 32; ++i) acc += red[i];
    return acc;
}

__device__ inline double block_max(double v, double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double acc = red[0];
#pragma unroll
    for (int i = 1; i < QX_NT / 32; ++i) acc = fmax(acc, red[i]);
    return acc;
}

// ------------------------------------------------------------------------------------
// coordination numbers (GFN double-exponential; D4 erf with EN weighting) + pair derivative
// tables dcnp[i*nat+j] = (1/r) d f(r_ij)/dr, so that d cn_i/d R_i = sum_j dcnp_ij (R_i - R_j).
static __device__ __noinline__ void phase_cn(const DevModel &m, Sm &s, double *dcnp, double *dcnp4) {
    const int nat = m.nat;
    if (m.method == 1) {   // GFN1 / D3: exponential counting function, k1 = 16 (reference src/dftd3.f90:607-642), cn_thr = 1000 bohr^2
        for (int i = threadIdx.x; i < nat; i += QX_NT) {
            double cn = 0.0;
            const double xi = s.xyz[3 * i], yi = s.xyz[3 * i + 1], zi = s.xyz[3 * i + 2], rci = m.at_rcov[i];
            for (int j = 0; j < nat; ++j) {
                double g = 0.0;
                if (j != i) {
                    const double vx = xi - s.xyz[3 * j], vy = yi - s.xyz[3 * j + 1], vz = zi - s.xyz[3 * j + 2];
                    const double r2 = vx * vx + vy * vy + vz * vz, r = sqrt(r2), rco = rci + m.at_rcov[j];
                    if (r2 <= 1000.0) {
                        const double ex = exp(-16.0 * (rco / r - 1.0)), f = 1.0 / (1.0 + ex);
                        cn += f;
                        g = -16.0 * rco / r2 * ex * f * f / r;
                    }
                }
                dcnp[i * nat + j] = g;
                dcnp4[i * nat + j] = 0.0;
            }
            s.cn[i] = cn;
            s.cn4[i] = 0.0;
        }
        __syncthreads();
        return;
    }
    for (int i = threadIdx.x; i < nat; i += QX_NT) {
        double cn = 0.0, cn4 = 0.0;
        double xi = s.xyz[3 * i], yi = s.xyz[3 * i + 1], zi = s.xyz[3 * i + 2];
        double rci = m.at_rcov[i], eni = m.at_en[i];
        for (int j = 0; j < nat; ++j) {
            double g = 0.0, g4 = 0.0;
            if (j != i) {
                double vx = xi - s.xyz[3 * j], vy = yi - s.xyz[3 * j + 1], vz = zi - s.xyz[3 * j + 2];
                double r2 = vx * vx + vy * vy + vz * vz, r = sqrt(r2), rc = rci + m.at_rcov[j];
                if (r2 <= 625.0) {
                    double ea = exp(-10.0 * (rc / r - 1.0)), eb = exp(-20.0 * ((rc + 2.0) / r - 1.0));
                    double fa = 1.0 / (1.0 + ea), fb = 1.0 / (1.0 + eb);
                    double dfa = -10.0 * rc / r2 * ea * fa * fa, dfb = -20.0 * (rc + 2.0) / r2 * eb * fb * fb;
                    cn += fa * fb;
                    g = (dfa * fb + fa * dfb) / r;
                }
                if (r2 <= 900.0) {
                    const double k4 = 4.10451, k5 = 19.08857, k6 = 2.0 * 11.28174 * 11.28174;
                    double den = fabs(eni - m.at_en[j]) + k5;
                    den = k4 * exp(-den * den / k6);
                    double arg = 7.5 * (r / rc - 1.0);
                    cn4 += den * 0.5 * erfc(arg);
                    g4 = -den * 7.5 / rc * 0.56418958354775628695 * exp(-arg * arg) / r;
                }
            }
            dcnp[i * nat + j] = g;
            dcnp4[i * nat + j] = g4;
        }
        s.cn[i] = cn;
        s.cn4[i] = cn4;
    }
    __syncthreads();
}

// classical repulsion: returns the CTA-wide energy; initialises s.grad
static __device__ __noinline__ double phase_repulsion(const DevModel &m, Sm &s) {
    const int nat = m.nat;
    double e = 0.0;
    for (int i = threadIdx.x; i < nat; i += QX_NT) {
        double gx = 0, gy = 0, gz = 0;
        for (int j = 0; j < nat; ++j) {
            if (j == i) continue;
            double vx = s.xyz[3 * i] - s.xyz[3 * j], vy = s.xyz[3 * i + 1] - s.xyz[3 * j + 1], vz = s.xyz[3 * i + 2] - s.xyz[3 * j + 2];
            double r2 = vx * vx + vy * vy + vz * vz, r = sqrt(r2);
            bool light = m.method != 1 && m.num[i] <= 2 && m.num[j] <= 2;   // GFN1: kexp = 1.5 for every pair
            double kexp = light ? GFN2_REP_KEXP_LIGHT : GFN2_REP_KEXP;
            double alpha = sqrt(m.at_repa[i] * m.at_repa[j]), zz = m.at_repz[i] * m.at_repz[j];
            double rk = light ? r : r * sqrt(r);
            double eij = zz * exp(-alpha * rk) / r;
            e += 0.5 * eij;
            double dedr = -(alpha * rk * kexp + 1.0) * eij / r2;
            gx += dedr * vx; gy += dedr * vy; gz += dedr * vz;
        }
        s.grad[3 * i] = gx; s.grad[3 * i + 1] = gy; s.grad[3 * i + 2] = gz;
        s.dEdcn[i] = 0.0;
        s.dEdcn4[i] = 0.0;
    }
    return block_sum(e, s.red);
}

// ------------------------------------------------------------------------------------ D4
__device__ inline double d4_zeta(double a, double c, double qref, double qmod) {
    return qmod < 0.0 ? exp(a) : exp(a * (1.0 - exp(c * (1.0 - qref / qmod))));
}
__device__ inline double d4_dzeta(double a, double c, double qref, double qmod) {
    return qmod < 0.0 ? 0.0 : -a * c * exp(c * (1.0 - qref / qmod)) * d4_zeta(a, c, qref, qmod) * qref / (qmod * qmod);
}

// Gaussian CN weights x charge scaling for atom i.  gw/gwdcn/gwdq point to 7-vectors (may be null).
__device__ inline void d4_weights_atom(const DevModel &m, int i, double cn, double q, double *gw, double *gwdcn, double *gwdq) {
    const int nref = m.at_nref[i];
    const double *refcn = m.at_refcn + i * QX_MAXREF, *refq = m.at_refq + i * QX_MAXREF;
    const int *ngw = m.at_ngw + i * QX_MAXREF;
    const double zi = m.at_zeff[i], gi = m.at_gam[i] * GFN2_D4_GC;
    double norm = 0.0, dnorm = 0.0, maxcn = -1.0;
    for (int r = 0; r < nref; ++r) {
        double dc = cn - refcn[r];
        for (int g = 1; g <= ngw[r]; ++g) {
            double wf = g * GFN2_D4_WF, w = exp(-wf * dc * dc);
            norm += w;
            dnorm += 2.0 * wf * (-dc) * w;
        }
        maxcn = fmax(maxcn, refcn[r]);
    }
    norm = 1.0 / norm;
    for (int r = 0; r < nref; ++r) {
        double dc = cn - refcn[r], expw = 0.0, expd = 0.0;
        for (int g = 1; g <= ngw[r]; ++g) {
            double wf = g * GFN2_D4_WF, w = exp(-wf * dc * dc);
            expw += w;
            expd += 2.0 * wf * (-dc) * w;
        }
        double gwk = expw * norm;
        if (gwk != gwk || fabs(gwk) > 1e300) gwk = (maxcn == refcn[r]) ? 1.0 : 0.0;
        double dgwk = norm * (expd - expw * dnorm * norm);
        if (dgwk != dgwk || fabs(dgwk) > 1e300) dgwk = 0.0;
        double zt = d4_zeta(GFN2_D4_GA, gi, refq[r] + zi, q + zi);
        if (gw) gw[r] = gwk * zt;
        if (gwdcn) gwdcn[r] = dgwk * zt;
        if (gwdq) gwdq[r] = gwk * d4_dzeta(GFN2_D4_GA, gi, refq[r] + zi, q + zi);
    }
    for (int r = nref; r < QX_MAXREF; ++r) {
        if (gw) gw[r] = 0.0;
        if (gwdcn) gwdcn[r] = 0.0;
        if (gwdq) gwdq[r] = 0.0;
    }
}

// ---- 8-lane groups: the O(nat), O(nsh) and O(nat^2) pieces of the SCC cycle have far fewer independent items than the CTA has
// threads, so an item (atom / shell) is shared by a group of 8 lanes and reduced with three shuffle stages; trip counts are
// the same for every lane of a warp (inactive groups work on a clamped index and do not store).
__device__ __forceinline__ double oct_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// d4_weights_atom with the reference systems of atom i spread over the 8 lanes of a group (QX_MAXREF = 7 <= 8)
__device__ __forceinline__ void d4_weights_oct(const DevModel &m, int i, bool active, double cn, double q, double *gw, double *gwdcn, double *gwdq) {
    const int r = threadIdx.x & 7;
    const bool on = r < m.at_nref[i];
    const int ir = i * QX_MAXREF + (on ? r : 0);
    const double refcn = m.at_refcn[ir], refq = m.at_refq[ir];
    const int ngw = on ? m.at_ngw[ir] : 0;
    const double zi = m.at_zeff[i], gi = m.at_gam[i] * GFN2_D4_GC;
    const double dc = cn - refcn;
    double expw = 0.0, expd = 0.0;
    for (int g = 1; g <= ngw; ++g) {
        const double wf = g * GFN2_D4_WF, w = exp(-wf * dc * dc);
        expw += w;
        expd += 2.0 * wf * (-dc) * w;
    }
    double norm = oct_sum(expw);
    const double dnorm = oct_sum(expd);
    double maxcn = on ? refcn : -1.0;
    maxcn = fmax(maxcn, __shfl_xor_sync(0xffffffffu, maxcn, 4));
    maxcn = fmax(maxcn, __shfl_xor_sync(0xffffffffu, maxcn, 2));
    maxcn = fmax(maxcn, __shfl_xor_sync(0xffffffffu, maxcn, 1));
    norm = 1.0 / norm;
    double gwk = expw * norm;
    if (gwk != gwk || fabs(gwk) > 1e300) gwk = (maxcn == refcn) ? 1.0 : 0.0;
    double dgwk = norm * (expd - expw * dnorm * norm);
    if (dgwk != dgwk || fabs(dgwk) > 1e300) dgwk = 0.0;
    if (active && r < QX_MAXREF) {
        const double zt = on ? d4_zeta(GFN2_D4_GA, gi, refq + zi, q + zi) : 0.0;
        if (gw) gw[r] = on ? gwk * zt : 0.0;
        if (gwdcn) gwdcn[r] = on ? dgwk * zt : 0.0;
        if (gwdq) gwdq[r] = on ? gwk * d4_dzeta(GFN2_D4_GA, gi, refq + zi, q + zi) : 0.0;
    }
}

// weights of every atom at the charges currently in s.qat (use_q) or at q = 0
__device__ __forceinline__ void d4_weights_all(const DevModel &m, Sm &s, bool use_q, double *gw, double *gwdcn, double *gwdq) {
    const int nat = m.nat, oct = threadIdx.x >> 3;
    for (int base = 0; base < nat; base += QX_NT / 8) {
        if (base + ((threadIdx.x >> 5) << 2) >= nat) continue;   // warp-uniform: none of this warp's four groups has an atom
        const bool active = base + oct < nat;
        const int i = active ? base + oct : 0;
        d4_weights_oct(m, i, active, s.cn4[i], use_q ? s.qat[i] : 0.0, gw ? gw + 7 * i : nullptr, gwdcn ? gwdcn + 7 * i : nullptr, gwdq ? gwdq + 7 * i : nullptr);
    }
}

__device__ inline double bj_r0(const DevModel &m, int i, int j) { return GFN2_D4_A1 * sqrt(3.0 * m.at_r4r2[i] * m.at_r4r2[j]) + GFN2_D4_A2; }

// atomic C6(i,j) and d C6(i,j)/d cn_i from the weights currently in s.gw / gwdcn (global tmp)
__device__ inline void d4_c6_tables(const DevModel &m, const double *gw, const double *gwdcn, double *c6, double *dc6) {
    const int nat = m.nat;
    for (int ij = threadIdx.x; ij < nat * nat; ij += QX_NT) {
        int i = ij / nat, j = ij - i * nat;
        const double *ref = m.c6ref + ((size_t)m.type[i] * m.ntype + m.type[j]) * QX_MAXREF * QX_MAXREF;
        double v = 0.0, dv = 0.0;
        int ni = m.at_nref[i], nj = m.at_nref[j];
        for (int ri = 0; ri < ni; ++ri) {
            double t = 0.0;
            for (int rj = 0; rj < nj; ++rj) t += ref[ri * QX_MAXREF + rj] * gw[j * QX_MAXREF + rj];
            v += gw[i * QX_MAXREF + ri] * t;
            dv += gwdcn[i * QX_MAXREF + ri] * t;
        }
        c6[ij] = v;
        dc6[ij] = dv;
    }
}

// non-self-consistent part of D4: ATM with q = 0 weights; also fills edisp[i*nat+j] (two-body BJ kernel).
// tmp: >= 14*nat + 5*nat*nat doubles of global scratch.
static __device__ __noinline__ double phase_d4_nonsc(const DevModel &m, Sm &s, double *edisp, double *c6, double *dc6, double *tmp) {
    const int nat = m.nat;
    double *gw0 = tmp, *gwdcn0 = tmp + 7 * nat, *part = tmp + 14 * nat;
    d4_weights_all(m, s, false, gw0, gwdcn0, nullptr);
    for (int ij = threadIdx.x; ij < nat * nat; ij += QX_NT) {
        int i = ij / nat, j = ij - i * nat;
        double e = 0.0;
        if (i != j) {
            double vx = s.xyz[3 * i] - s.xyz[3 * j], vy = s.xyz[3 * i + 1] - s.xyz[3 * j + 1], vz = s.xyz[3 * i + 2] - s.xyz[3 * j + 2];
            double r2 = vx * vx + vy * vy + vz * vz;
            if (r2 <= 3600.0) {
                double r0 = bj_r0(m, i, j), rrij = 3.0 * m.at_r4r2[i] * m.at_r4r2[j];
                double r02 = r0 * r0, r06 = r02 * r02 * r02, r6 = r2 * r2 * r2;
                e = GFN2_D4_S6 / (r6 + r06) + GFN2_D4_S8 * rrij / (r6 * r2 + r06 * r02);
            }
        }
        edisp[ij] = e;
    }
    __syncthreads();
    d4_c6_tables(m, gw0, gwdcn0, c6, dc6);
    __syncthreads();
    // triples: task (i, j != i), inner k > j, k != i; force and dE/dcn on vertex i only
    double e3 = 0.0;
    const double alp = GFN2_D4_ALP;
    for (int ij = threadIdx.x; ij < nat * nat; ij += QX_NT) {
        int i = ij / nat, j = ij - i * nat;
        double gx = 0, gy = 0, gz = 0, dcn = 0;
        if (i != j) {
            double vij[3] = {s.xyz[3 * j] - s.xyz[3 * i], s.xyz[3 * j + 1] - s.xyz[3 * i + 1], s.xyz[3 * j + 2] - s.xyz[3 * i + 2]};
            double r2ij = vij[0] * vij[0] + vij[1] * vij[1] + vij[2] * vij[2];
            if (r2ij <= 1600.0) {
                double c6ij = c6[i * nat + j], r0ij = bj_r0(m, i, j);
                for (int k = j + 1; k < nat; ++k) {
                    if (k == i) continue;
                    double vik[3] = {s.xyz[3 * k] - s.xyz[3 * i], s.xyz[3 * k + 1] - s.xyz[3 * i + 1], s.xyz[3 * k + 2] - s.xyz[3 * i + 2]};
                    double r2ik = vik[0] * vik[0] + vik[1] * vik[1] + vik[2] * vik[2];
                    double dx = vik[0] - vij[0], dy = vik[1] - vij[1], dz = vik[2] - vij[2];
                    double r2jk = dx * dx + dy * dy + dz * dz;
                    if (r2ik > 1600.0 || r2jk > 1600.0) continue;
                    double c6ik = c6[i * nat + k], c6jk = c6[j * nat + k];
                    double r0 = r0ij * bj_r0(m, i, k) * bj_r0(m, j, k);
                    double c9 = -GFN2_D4_S9 * sqrt(fabs(c6ij * c6ik * c6jk));
                    double r2 = r2ij * r2ik * r2jk, r1 = sqrt(r2), r3 = r2 * r1, r5 = r3 * r2;
                    double rr0 = pow(r0 / r1, alp / 3.0);
                    double fdmp = 1.0 / (1.0 + 6.0 * rr0);
                    double ang = 0.375 * (r2ij + r2jk - r2ik) * (r2ij - r2jk + r2ik) * (-r2ij + r2jk + r2ik) / r5 + 1.0 / r3;
                    double dE = ang * fdmp * c9;
                    e3 -= dE / 3.0;
                    double dfdmp = -2.0 * alp * rr0 * fdmp * fdmp;
                    double dang_ij = -0.375 * (r2ij * r2ij * r2ij + r2ij * r2ij * (r2jk + r2ik) +
                                               r2ij * (3.0 * r2jk * r2jk + 2.0 * r2jk * r2ik + 3.0 * r2ik * r2ik) -
                                               5.0 * (r2jk - r2ik) * (r2jk - r2ik) * (r2jk + r2ik)) / r5;
                    double dang_ik = -0.375 * (r2ik * r2ik * r2ik + r2ik * r2ik * (r2jk + r2ij) +
                                               r2ik * (3.0 * r2jk * r2jk + 2.0 * r2jk * r2ij + 3.0 * r2ij * r2ij) -
                                               5.0 * (r2jk - r2ij) * (r2jk - r2ij) * (r2jk + r2ij)) / r5;
                    double gij = c9 * (-dang_ij * fdmp + ang * dfdmp) / r2ij;
                    double gik = c9 * (-dang_ik * fdmp + ang * dfdmp) / r2ik;
                    gx += -gij * vij[0] - gik * vik[0];
                    gy += -gij * vij[1] - gik * vik[1];
                    gz += -gij * vij[2] - gik * vik[2];
                    dcn -= dE * 0.5 * (dc6[i * nat + j] / c6ij + dc6[i * nat + k] / c6ik);
                }
            }
        }
        part[4 * ij] = gx; part[4 * ij + 1] = gy; part[4 * ij + 2] = gz; part[4 * ij + 3] = dcn;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nat; i += QX_NT) {
        double gx = 0, gy = 0, gz = 0, dcn = 0;
        for (int j = 0; j < nat; ++j) {
            const double *p = part + 4 * (i * nat + j);
            gx += p[0]; gy += p[1]; gz += p[2]; dcn += p[3];
        }
        s.grad[3 * i] += gx; s.grad[3 * i + 1] += gy; s.grad[3 * i + 2] += gz;
        s.dEdcn4[i] += dcn;
    }
    return block_sum(e3, s.red);
}

// ------------------------------------------------------------------------------------ GFN1: D3(BJ) and the halogen-bond correction
// C6(i,j) from the reference systems with Gaussian weights in CN space, and dC6/dCN_i (reference src/dftd3.f90:334-405, k3 = -4)
__device__ inline void d3_c6_pair(const DevModel &m, int i, int j, double cni, double cnj, double &c6, double &dc6i) {
    const double *ref = m.d3ref + ((size_t)m.type[i] * m.ntype + m.type[j]) * 75;
    double c6mem = -1.e99, r_save = 9999.0, zaehler = 0.0, nenner = 0.0, dz = 0.0, dn = 0.0;
    for (int a = 0; a < m.at_mxc[i]; ++a)
        for (int b = 0; b < m.at_mxc[j]; ++b) {
            const double c6ref = ref[(a * 5 + b) * 3];
            if (c6ref > 0.0) {
                const double ci = ref[(a * 5 + b) * 3 + 1], cj = ref[(a * 5 + b) * 3 + 2];
                const double r = (ci - cni) * (ci - cni) + (cj - cnj) * (cj - cnj);
                if (r < r_save) { r_save = r; c6mem = c6ref; }
                double expterm = exp(-4.0 * r);
                zaehler += c6ref * expterm;
                nenner += expterm;
                expterm = expterm * 2.0 * -4.0;
                const double term = expterm * (cni - ci);
                dz += c6ref * term;
                dn += term;
            }
        }
    if (nenner > 1.0e-99) { c6 = zaehler / nenner; dc6i = ((dz * nenner) - (dn * zaehler)) / (nenner * nenner); }
    else { c6 = c6mem; dc6i = 0.0; }
}

// E = - sum_{i>j} C6 (s6 / (r^6 + R0^6) + 3 s8 r42 / (r^8 + R0^8)), R0 = a1 sqrt(3 r42) + a2 (reference src/dftd3.f90:107-186, BJ variant,
// no three-body term).  Thread i sums over all partners j: half the pair energy, the full gradient of atom i at fixed CN and
// dE/dCN_i.  Followed by the halogen-bond correction (a handful of X...A contacts: one thread, fixed order).
static __device__ __noinline__ double phase_d3_xb(const DevModel &m, Sm &s) {
    const int nat = m.nat;
    double e = 0.0;
    for (int i = threadIdx.x; i < nat; i += QX_NT) {
        double gx = 0, gy = 0, gz = 0, dcn = 0;
        for (int j = 0; j < nat; ++j) {
            if (j == i) continue;
            const double vx = s.xyz[3 * i] - s.xyz[3 * j], vy = s.xyz[3 * i + 1] - s.xyz[3 * j + 1], vz = s.xyz[3 * i + 2] - s.xyz[3 * j + 2];
            const double r2 = vx * vx + vy * vy + vz * vz;
            if (r2 > 4000.0) continue;
            double c6, dc6i;
            d3_c6_pair(m, i, j, s.cn[i], s.cn[j], c6, dc6i);
            const double r42 = m.at_r2r4d3[i] * m.at_r2r4d3[j], r = sqrt(r2), r4 = r2 * r2, r6 = r4 * r2, r8 = r6 * r2;
            const double R0 = GFN1_D3_A1 * sqrt(3.0 * r42) + GFN1_D3_A2, R02 = R0 * R0, R06 = R02 * R02 * R02;
            const double t6 = r6 + R06, t8 = r8 + R06 * R02;
            const double rest = GFN1_D3_S6 / t6 + 3.0 * GFN1_D3_S8 * r42 / t8;
            e -= 0.5 * rest * c6;
            const double dedr = c6 * (GFN1_D3_S6 * 6.0 * r4 * r / (t6 * t6) + GFN1_D3_S8 * 24.0 * r42 * r6 * r / (t8 * t8)) / r;
            gx += dedr * vx; gy += dedr * vy; gz += dedr * vz;
            dcn -= rest * dc6i;
        }
        s.grad[3 * i] += gx; s.grad[3 * i + 1] += gy; s.grad[3 * i + 2] += gz;
        s.dEdcn[i] += dcn;
    }
    __syncthreads();
    if (threadIdx.x == 0) {   // halogen bond: donors X (with strength), nearest neighbour K, acceptors A = N, O, P, S within 20 bohr
        for (int x = 0; x < nat; ++x) {
            const double cx = m.at_xb[x];
            if (cx == 0.0) continue;
            int kn = -1;
            double best = 1e300;
            for (int k = 0; k < nat; ++k) {
                if (k == x) continue;
                const double dx = s.xyz[3 * k] - s.xyz[3 * x], dy = s.xyz[3 * k + 1] - s.xyz[3 * x + 1], dz = s.xyz[3 * k + 2] - s.xyz[3 * x + 2];
                const double d2 = dx * dx + dy * dy + dz * dz;
                if (d2 < best) { best = d2; kn = k; }
            }
            if (kn < 0) continue;
            for (int a = 0; a < nat; ++a) {
                const int za = m.num[a];
                if (a == x || a == kn || !(za == 7 || za == 8 || za == 15 || za == 16)) continue;
                double u[3], w[3], ru2 = 0, rw2 = 0, uw = 0;
                for (int c = 0; c < 3; ++c) {
                    u[c] = s.xyz[3 * a + c] - s.xyz[3 * x + c];
                    w[c] = s.xyz[3 * kn + c] - s.xyz[3 * x + c];
                    ru2 += u[c] * u[c]; rw2 += w[c] * w[c]; uw += u[c] * w[c];
                }
                if (ru2 > 400.0) continue;
                const double ru = sqrt(ru2), rw = sqrt(rw2), cosv = uw / (ru * rw);
                const double r0 = GFN1_XB_RAD * (m.at_rad[a] + m.at_rad[x]);
                const double t = r0 / ru, t2 = t * t, t6 = t2 * t2 * t2, t12 = t6 * t6;
                const double lj = (t12 - GFN1_XB_DAMP * t6) / (1.0 + t12);
                const double dt6 = -6.0 * t6 / ru, dt12 = -12.0 * t12 / ru;
                const double dlj = ((dt12 - GFN1_XB_DAMP * dt6) * (1.0 + t12) - (t12 - GFN1_XB_DAMP * t6) * dt12) / ((1.0 + t12) * (1.0 + t12));
                const double base = 0.5 - 0.25 * cosv, b2 = base * base, b5 = b2 * b2 * base, at = b5 * base, dat = 6.0 * b5 * (-0.25);
                e += cx * at * lj;
                for (int c = 0; c < 3; ++c) {
                    const double dcos_a = (w[c] / rw - cosv * u[c] / ru) / ru, dcos_k = (u[c] / ru - cosv * w[c] / rw) / rw;
                    const double ga = cx * (at * dlj * u[c] / ru + dat * lj * dcos_a), gk = cx * dat * lj * dcos_k;
                    s.grad[3 * a + c] += ga;
                    s.grad[3 * kn + c] += gk;
                    s.grad[3 * x + c] -= ga + gk;
                }
            }
        }
    }
    return block_sum(e, s.red);
}

// ------------------------------------------------------------------------------------ Coulomb set-up
static __device__ __noinline__ void phase_coulomb_setup(const DevModel &m, Sm &s, double *gamma) {
    const int nat = m.nat, nsh = m.nsh;
    for (int ab = threadIdx.x; ab < nsh * nsh; ab += QX_NT) {
        int a = ab / nsh, b = ab - a * nsh, i = m.sh_at[a], j = m.sh_at[b];
        double gam = m.method == 1 ? 2.0 / (1.0 / m.sh_hub[a] + 1.0 / m.sh_hub[b])   // GFN1: harmonic average
                                   : 0.5 * (m.sh_hub[a] + m.sh_hub[b]);
        if (i != j) {
            double vx = s.xyz[3 * i] - s.xyz[3 * j], vy = s.xyz[3 * i + 1] - s.xyz[3 * j + 1], vz = s.xyz[3 * i + 2] - s.xyz[3 * j + 2];
            gam = 1.0 / sqrt(vx * vx + vy * vy + vz * vz + 1.0 / (gam * gam));
        }
        gamma[ab] = gam;
    }
    for (int i = threadIdx.x; i < nat; i += QX_NT) {
        double arg = s.cn[i] - m.at_mpvcn[i] - GFN2_MP_SHIFT;
        double t1 = exp(-GFN2_MP_KEXP * arg), t2 = (GFN2_MP_RMAX - m.at_mprad[i]) / (1.0 + t1);
        s.mrad[i] = m.at_mprad[i] + t2;
        s.dmr[i] = t2 * GFN2_MP_KEXP * t1 / (1.0 + t1);
    }
    for (int a = threadIdx.x; a < nsh; a += QX_NT) s.selfen[a] = m.sh_level[a] - m.sh_kcn[a] * s.cn[m.sh_at[a]];
    __syncthreads();
}

// ------------------------------------------------------------------------------------ integrals
// <a| O |b> with the multipole operator centred on atom(b); a on the bra atom J, b on the ket atom I,
// vec = R_I - R_J.  out: S, D(3), Q(6, traceless).  If grad != nullptr additionally returns
// grad[k] = d/dvec_k of  sum_c coef[c] * out_raw[c]  where out_raw are the raw (not traceless) moments
// -- used by the gradient kernel with pre-contracted coefficients.
__device__ inline void ao_pair_multipole(const DevModel &m, int sa, int ma, int sb, int mb, const double vec[3], double r2,
                                         double out[10], const double *coef, double *grad) {
    const int la = m.sh_l[sa], lb = m.sh_l[sb];
    const SphTerm &ta = c_sph[la * la + ma], &tb = c_sph[lb * lb + mb];
    for (int c = 0; c < 10; ++c) out[c] = 0.0;
    if (grad) grad[0] = grad[1] = grad[2] = 0.0;
    const int npa = m.sh_np[sa], npb = m.sh_np[sb];
    const int amax = la + (grad ? 1 : 0), bmax = lb + 2;
    for (int pa_ = 0; pa_ < npa; ++pa_) {
        const double aj = m.sh_alpha[sa * QX_MAXPRIM + pa_], cj = m.sh_coef[sa * QX_MAXPRIM + pa_];
        for (int pb_ = 0; pb_ < npb; ++pb_) {
            const double ai = m.sh_alpha[sb * QX_MAXPRIM + pb_], ci = m.sh_coef[sb * QX_MAXPRIM + pb_];
            const double gam = ai + aj, est = ai * aj * r2 / gam;
            if (est > 25.0) continue;
            const double pg = QX_PI / gam;
            const double pre = exp(-est) * pg * sqrt(pg) * ci * cj;
            const double oog = 0.5 / gam;
            double t[3][4][5];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const double pa = ai / gam * vec[d], pb = -aj / gam * vec[d];
                t[d][0][0] = 1.0;
                for (int b = 0; b < bmax; ++b) t[d][0][b + 1] = pb * t[d][0][b] + (b > 0 ? b * oog * t[d][0][b - 1] : 0.0);
                for (int a = 0; a < amax; ++a)
                    for (int b = 0; b <= bmax; ++b)
                        t[d][a + 1][b] = pa * t[d][a][b] + (a > 0 ? a * oog * t[d][a - 1][b] : 0.0) + (b > 0 ? b * oog * t[d][a][b - 1] : 0.0);
            }
            for (int ka = 0; ka < ta.n; ++ka) {
                const int *ea = ta.ex[ka];
                for (int kb = 0; kb < tb.n; ++kb) {
                    const int *eb = tb.ex[kb];
                    const double w = pre * ta.c[ka] * tb.c[kb];
                    double f[3][3];
#pragma unroll
                    for (int d = 0; d < 3; ++d)
#pragma unroll
                        for (int mm = 0; mm < 3; ++mm) f[d][mm] = t[d][ea[d]][eb[d] + mm];
                    out[0] += w * f[0][0] * f[1][0] * f[2][0];
                    out[1] += w * f[0][1] * f[1][0] * f[2][0];
                    out[2] += w * f[0][0] * f[1][1] * f[2][0];
                    out[3] += w * f[0][0] * f[1][0] * f[2][1];
                    out[4] += w * f[0][2] * f[1][0] * f[2][0];
                    out[5] += w * f[0][1] * f[1][1] * f[2][0];
                    out[6] += w * f[0][0] * f[1][2] * f[2][0];
                    out[7] += w * f[0][1] * f[1][0] * f[2][1];
                    out[8] += w * f[0][0] * f[1][1] * f[2][1];
                    out[9] += w * f[0][0] * f[1][0] * f[2][2];
                    if (grad) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            double g[3][3];
#pragma unroll
                            for (int d = 0; d < 3; ++d)
#pragma unroll
                                for (int mm = 0; mm < 3; ++mm) {
                                    if (d == k) {
                                        double up = t[d][ea[d] + 1][eb[d] + mm];
                                        double dn = ea[d] > 0 ? t[d][ea[d] - 1][eb[d] + mm] : 0.0;
                                        g[d][mm] = -(2.0 * aj * up - ea[d] * dn);
                                    } else
                                        g[d][mm] = f[d][mm];
                                }
                            double acc = coef[0] * g[0][0] * g[1][0] * g[2][0];
                            acc += coef[1] * g[0][1] * g[1][0] * g[2][0];
                            acc += coef[2] * g[0][0] * g[1][1] * g[2][0];
                            acc += coef[3] * g[0][0] * g[1][0] * g[2][1];
                            acc += coef[4] * g[0][2] * g[1][0] * g[2][0];
                            acc += coef[5] * g[0][1] * g[1][1] * g[2][0];
                            acc += coef[6] * g[0][0] * g[1][2] * g[2][0];
                            acc += coef[7] * g[0][1] * g[1][0] * g[2][1];
                            acc += coef[8] * g[0][0] * g[1][1] * g[2][1];
                            acc += coef[9] * g[0][0] * g[1][0] * g[2][2];
                            grad[k] += w * acc;
                        }
                    }
                }
            }
        }
    }
}

// ---- s / p specialisation of ao_pair_multipole.  The generic routine above indexes its Obara-Saika table with run-time
// exponents, which forces the table into local memory (ncu: the recursion lines are the hottest of the integral and gradient
// phases, all long-scoreboard stalls).  For s and p functions an AO is a single Cartesian monomial (p_m along axis (m+1) mod 3),
// so with the angular momenta as template parameters every table index is a compile-time constant and the only run-time
// choice -- "is this the axis of the p function?" -- is a register select.  Work lists are sorted by (la, lb), so the dispatch
// is warp-uniform except at class boundaries.  Same arithmetic as the generic routine.
template <int LA, int LB, bool GRAD>
__device__ __forceinline__ void ao_pair_multipole_sp(const DevModel &m, int sa, int ka, int sb, int kb, const double vec[3], double r2,
                                                     double out[10], const double *coef, double *grad) {
    constexpr int AMAX = LA + (GRAD ? 1 : 0), BMAX = LB + 2;
#pragma unroll
    for (int c = 0; c < 10; ++c) out[c] = 0.0;
    if (GRAD) grad[0] = grad[1] = grad[2] = 0.0;
    const int npa = m.sh_np[sa], npb = m.sh_np[sb];
    for (int pa_ = 0; pa_ < npa; ++pa_) {
        const double aj = m.sh_alpha[sa * QX_MAXPRIM + pa_], cj = m.sh_coef[sa * QX_MAXPRIM + pa_];
        for (int pb_ = 0; pb_ < npb; ++pb_) {
            const double ai = m.sh_alpha[sb * QX_MAXPRIM + pb_], ci = m.sh_coef[sb * QX_MAXPRIM + pb_];
            const double gam = ai + aj, est = ai * aj * r2 / gam;
            if (est > 25.0) continue;
            const double pg = QX_PI / gam;
            const double w = exp(-est) * pg * sqrt(pg) * ci * cj;
            const double oog = 0.5 / gam;
            double f[3][3], g[3][3];   // f: plain 1-D factors for mm = 0..2; g: the same with the bra function differentiated
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const double pa = ai / gam * vec[d], pb = -aj / gam * vec[d];
                double t[AMAX + 1][BMAX + 1];
                t[0][0] = 1.0;
#pragma unroll
                for (int b = 0; b < BMAX; ++b) t[0][b + 1] = pb * t[0][b] + (b > 0 ? b * oog * t[0][b - 1] : 0.0);
#pragma unroll
                for (int a = 0; a < AMAX; ++a)
#pragma unroll
                    for (int b = 0; b <= BMAX; ++b)
                        t[a + 1][b] = pa * t[a][b] + (a > 0 ? a * oog * t[a - 1][b] : 0.0) + (b > 0 ? b * oog * t[a][b - 1] : 0.0);
                const bool ia = LA == 1 && d == ka, ib = LB == 1 && d == kb;   // exponent 1 along this axis?
#pragma unroll
                for (int mm = 0; mm < 3; ++mm) {
                    // ket exponent (ib ? 1 : 0) + mm, bra exponent (ia ? 1 : 0)
                    const double r0 = LB == 1 ? (ib ? t[0][1 + mm] : t[0][mm]) : t[0][mm];
                    double r1 = 0.0, r2_ = 0.0;
                    if (AMAX >= 1) r1 = LB == 1 ? (ib ? t[AMAX >= 1 ? 1 : 0][1 + mm] : t[AMAX >= 1 ? 1 : 0][mm]) : t[AMAX >= 1 ? 1 : 0][mm];
                    if (AMAX >= 2) r2_ = LB == 1 ? (ib ? t[AMAX >= 2 ? 2 : 0][1 + mm] : t[AMAX >= 2 ? 2 : 0][mm]) : t[AMAX >= 2 ? 2 : 0][mm];
                    f[d][mm] = LA == 1 ? (ia ? r1 : r0) : r0;
                    if (GRAD) {
                        const double up = LA == 1 ? (ia ? r2_ : r1) : r1;
                        const double dn = LA == 1 ? (ia ? r0 : 0.0) : 0.0;     // e * t[e-1]: e = 1 only along the p axis
                        g[d][mm] = -(2.0 * aj * up - dn);
                    }
                }
            }
            out[0] += w * f[0][0] * f[1][0] * f[2][0];
            out[1] += w * f[0][1] * f[1][0] * f[2][0];
            out[2] += w * f[0][0] * f[1][1] * f[2][0];
            out[3] += w * f[0][0] * f[1][0] * f[2][1];
            out[4] += w * f[0][2] * f[1][0] * f[2][0];
            out[5] += w * f[0][1] * f[1][1] * f[2][0];
            out[6] += w * f[0][0] * f[1][2] * f[2][0];
            out[7] += w * f[0][1] * f[1][0] * f[2][1];
            out[8] += w * f[0][0] * f[1][1] * f[2][1];
            out[9] += w * f[0][0] * f[1][0] * f[2][2];
            if (GRAD) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double *x = k == 0 ? g[0] : f[0], *y = k == 1 ? g[1] : f[1], *z = k == 2 ? g[2] : f[2];
                    double acc = coef[0] * x[0] * y[0] * z[0];
                    acc += coef[1] * x[1] * y[0] * z[0];
                    acc += coef[2] * x[0] * y[1] * z[0];
                    acc += coef[3] * x[0] * y[0] * z[1];
                    acc += coef[4] * x[2] * y[0] * z[0];
                    acc += coef[5] * x[1] * y[1] * z[0];
                    acc += coef[6] * x[0] * y[2] * z[0];
                    acc += coef[7] * x[1] * y[0] * z[1];
                    acc += coef[8] * x[0] * y[1] * z[1];
                    acc += coef[9] * x[0] * y[0] * z[2];
                    grad[k] += w * acc;
                }
            }
        }
    }
}

// dispatch: s / p pairs through the specialisation, anything with a d function through the generic routine
template <bool GRAD>
__device__ __forceinline__ void ao_pair_dispatch(const DevModel &m, int sa, int ma, int sb, int mb, const double vec[3], double r2,
                                                 double out[10], const double *coef, double *grad) {
    const int la = m.sh_l[sa], lb = m.sh_l[sb];
    if (la <= 1 && lb <= 1) {
        const int ka = ma == 2 ? 0 : ma + 1, kb = mb == 2 ? 0 : mb + 1;   // p_m lies along axis (m + 1) mod 3 (order y, z, x)
        if (la == 0 && lb == 0) ao_pair_multipole_sp<0, 0, GRAD>(m, sa, ka, sb, kb, vec, r2, out, coef, grad);
        else if (la == 0) ao_pair_multipole_sp<0, 1, GRAD>(m, sa, ka, sb, kb, vec, r2, out, coef, grad);
        else if (lb == 0) ao_pair_multipole_sp<1, 0, GRAD>(m, sa, ka, sb, kb, vec, r2, out, coef, grad);
        else ao_pair_multipole_sp<1, 1, GRAD>(m, sa, ka, sb, kb, vec, r2, out, coef, grad);
    } else
        ao_pair_multipole(m, sa, ma, sb, mb, vec, r2, out, GRAD ? coef : nullptr, GRAD ? grad : nullptr);
}

__device__ inline void make_traceless(double q[6]) {
    double tr = 0.5 * (q[0] + q[2] + q[5]);
    for (int c = 0; c < 6; ++c) q[c] *= 1.5;
    q[0] -= tr; q[2] -= tr; q[5] -= tr;
}

// r - R_J = (r - R_I) + vec : move raw moments from centre I to centre J
__device__ inline void shift_raw(const double vec[3], const double in[10], double out[10]) {
    out[0] = in[0];
    for (int c = 0; c < 3; ++c) out[1 + c] = in[1 + c] + vec[c] * in[0];
    for (int c = 0; c < 6; ++c) {
        int a = c_qa[c], b = c_qb[c];
        out[4 + c] = in[4 + c] + vec[a] * in[1 + b] + vec[b] * in[1 + a] + vec[a] * vec[b] * in[0];
    }
}

__device__ inline double shpoly_pair(const DevModel &m, int sa, int sb, double rr) {
    return (1.0 + m.sh_poly[sa] * rr) * (1.0 + m.sh_poly[sb] * rr);
}

// Fills the per-CTA slab: S, H0 (symmetric), Dt/Qt in "operator on the FIRST index" layout:
//   Dt[c][b][a] = <a| (r - R_atom(b))_c |b>   (row b contiguous in a).
static __device__ __noinline__ void phase_integrals(const DevModel &m, Sm &s, double *S, double *H0, double *Dt, double *Qt) {
    const int nao = m.nao;
    const size_t n2 = (size_t)nao * nao;
    for (int t = threadIdx.x; t < m.ntask_int; t += QX_NT) {
        const int a = m.task_int[t].x, b = m.task_int[t].y;
        const int sa = m.ao_sh[a], sb = m.ao_sh[b], ja = m.ao_at[a], ib = m.ao_at[b];
        double vec[3] = {s.xyz[3 * ib] - s.xyz[3 * ja], s.xyz[3 * ib + 1] - s.xyz[3 * ja + 1], s.xyz[3 * ib + 2] - s.xyz[3 * ja + 2]};
        double r2 = vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2];
        double raw[10];
        ao_pair_dispatch<false>(m, sa, m.ao_m[a], sb, m.ao_m[b], vec, r2, raw, nullptr, nullptr);
        double hij = 0.5 * (s.selfen[sa] + s.selfen[sb]);
        if (ja != ib) {
            double rr = sqrt(sqrt(r2) / (m.at_rad[ja] + m.at_rad[ib]));
            hij *= m.hscale[sa * m.nsh + sb] * shpoly_pair(m, sa, sb, rr);
        }
        double q[6];
        for (int c = 0; c < 6; ++c) q[c] = raw[4 + c];
        make_traceless(q);
        size_t ba = (size_t)b * nao + a;
        S[ba] = raw[0];
        H0[ba] = raw[0] * hij;
        for (int c = 0; c < 3; ++c) Dt[c * n2 + ba] = raw[1 + c];
        for (int c = 0; c < 6; ++c) Qt[c * n2 + ba] = q[c];
        if (ja != ib) {
            double sh[10];
            shift_raw(vec, raw, sh);
            for (int c = 0; c < 6; ++c) q[c] = sh[4 + c];
            make_traceless(q);
            size_t ab = (size_t)a * nao + b;
            S[ab] = raw[0];
            H0[ab] = raw[0] * hij;
            for (int c = 0; c < 3; ++c) Dt[c * n2 + ab] = sh[1 + c];
            for (int c = 0; c < 6; ++c) Qt[c * n2 + ab] = q[c];
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------ dense kernels in shared memory
// Eigenvectors are kept TRANSPOSED: Ct[k][i] = C[i][k] (orbital k contiguous).  The leading dimension of the
// shared-memory matrices is == 4 or 12 (mod 16): with it both the DMMA fragment loads (8 rows x 4 columns)
// and the 128-bit row accesses of the Jacobi are free of bank conflicts.
//
// FP64 tensor-core GEMM  out(i,j) = sum_k A(i,k) B(k,j),  all n x n, operands given as element accessors.
// One warp owns a strip of 8 output rows and walks the column tiles, so the A fragment of a k-step is
// reused for every tile of the strip (mma.sync.m8n8k4.f64 == DMMA.8x8x4 on sm_100a).
template <int MAXT, class FA, class FB, class FS>
static __device__ __noinline__ void dmma_gemm(int n, FA loadA, FB loadB, FS store) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = QX_NT / 32;
    const int nt = (n + 7) >> 3, g = lane >> 2, tg = lane & 3;
    for (int ti = warp; ti < nt; ti += nwarp) {
        const int row = ti * 8 + g;
        for (int tb = 0; tb < nt; tb += MAXT) {   // column tiles in groups of MAXT (accumulators stay in registers)
            double acc[MAXT][2];
#pragma unroll
            for (int t = 0; t < MAXT; ++t) acc[t][0] = acc[t][1] = 0.0;
#pragma unroll 2
            for (int k0 = 0; k0 < n; k0 += 4) {
                const int k = k0 + tg;
                const double a = (row < n && k < n) ? loadA(row, k) : 0.0;
#pragma unroll
                for (int t = 0; t < MAXT; ++t) {
                    if (tb + t < nt) {
                        const int col = (tb + t) * 8 + g;
                        const double b = (k < n && col < n) ? loadB(k, col) : 0.0;
                        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                            : "+d"(acc[t][0]), "+d"(acc[t][1]) : "d"(a), "d"(b));
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < MAXT; ++t) {
                if (tb + t < nt) {
                    const int col = (tb + t) * 8 + 2 * tg;
                    if (row < n && col < n) store(row, col, acc[t][0]);
                    if (row < n && col + 1 < n) store(row, col + 1, acc[t][1]);
                }
            }
        }
    }
}

// Same product for operands that live in the global slab (large bases): the B operand is staged in shared memory one group of
// 8 MAXT columns at a time (all k), so that the inner loop reads B at shared-memory latency and B crosses L2 once instead of once
// per strip; the row stride 8 MAXT + 4 spreads the 4 x 8 doubles of a fragment load over all banks (MAXT odd).  Ends with a barrier.
template <int MAXT, class FA, class FB, class FS>
static __device__ __noinline__ void dmma_gemm_staged(int n, FA loadA, FB loadB, FS store, double *Bs) {
    QX_ASSUME_SHARED(Bs);
    constexpr int NC = 8 * MAXT, LDS = NC + 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = QX_NT / 32;
    const int nt = (n + 7) >> 3, g = lane >> 2, tg = lane & 3, kpad = (n + 3) & ~3;
    for (int tb = 0; tb < nt; tb += MAXT) {
        const int j0 = tb * 8;
        __syncthreads();   // the readers of the previous column group are done
        for (int base = 0; base < kpad * NC; base += 4 * QX_NT) {   // four independent L2 loads in flight per thread
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * QX_NT + threadIdx.x, k = idx / NC, j = j0 + idx - k * NC;
                v[u] = (idx < kpad * NC && k < n && j < n) ? loadB(k, j) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * QX_NT + threadIdx.x, k = idx / NC, jj = idx - k * NC;
                if (idx < kpad * NC) Bs[k * LDS + jj] = v[u];
            }
        }
        __syncthreads();
        const double *bp = Bs + tg * LDS + g;
        for (int ti = warp; ti < nt; ti += nwarp) {
            const int row = ti * 8 + g;
            double acc[MAXT][2];
#pragma unroll
            for (int t = 0; t < MAXT; ++t) acc[t][0] = acc[t][1] = 0.0;
            // The A operand comes straight from L2 (~600 cycles): its fragments are fetched KB k-steps ahead, one batch in flight while the
            // previous one feeds the tensor pipe (KB x MAXT DMMA of 16 cycles each per scheduler cover the round trip).
            constexpr int KB = 8;
            double a_cur[KB], a_nxt[KB];
#pragma unroll
            for (int u = 0; u < KB; ++u) { const int k = 4 * u + tg; a_cur[u] = (row < n && k < n) ? loadA(row, k) : 0.0; }
            for (int kb = 0; kb < kpad; kb += 4 * KB) {
#pragma unroll
                for (int u = 0; u < KB; ++u) { const int k = kb + 4 * KB + 4 * u + tg; a_nxt[u] = (row < n && k < n) ? loadA(row, k) : 0.0; }
#pragma unroll
                for (int u = 0; u < KB; ++u) {
                    const int k0 = kb + 4 * u;
                    if (k0 < kpad) {
#pragma unroll
                        for (int t = 0; t < MAXT; ++t) {
                            const double b = bp[k0 * LDS + 8 * t];
                            asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                : "+d"(acc[t][0]), "+d"(acc[t][1]) : "d"(a_cur[u]), "d"(b));
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < KB; ++u) a_cur[u] = a_nxt[u];
            }
#pragma unroll
            for (int t = 0; t < MAXT; ++t) {
                const int col = j0 + 8 * t + 2 * tg;
                if (row < n && col < n) store(row, col, acc[t][0]);
                if (row < n && col + 1 < n) store(row, col + 1, acc[t][1]);
            }
        }
    }
    __syncthreads();
}

// run-time dispatch on the number of column tiles (keeps the accumulators in registers); stage / stage_doubles: shared-memory
// buffer for the B operand when the operands are in the global slab (null: operands are read in place)
template <class FA, class FB, class FS>
__device__ inline void gemm_tc(int n, FA loadA, FB loadB, FS store, double *stage = nullptr, int stage_doubles = 0) {
    if (stage) {
        const int kpad = (n + 3) & ~3;
        if (kpad * (8 * 7 + 4) <= stage_doubles) { dmma_gemm_staged<7>(n, loadA, loadB, store, stage); return; }
        if (kpad * (8 * 5 + 4) <= stage_doubles) { dmma_gemm_staged<5>(n, loadA, loadB, store, stage); return; }
        if (kpad * (8 * 3 + 4) <= stage_doubles) { dmma_gemm_staged<3>(n, loadA, loadB, store, stage); return; }
    }
    const int nt = (n + 7) >> 3;
    if (nt <= 4) dmma_gemm<4>(n, loadA, loadB, store);
    else if (nt <= 9) dmma_gemm<9>(n, loadA, loadB, store);
    else dmma_gemm<16>(n, loadA, loadB, store);
}

// ---- guard-free strip GEMMs on zero-padded shared-memory matrices (rows and columns padded to 8*NT8, padding == 0).
// One warp owns one strip of 8 output rows (requires NT8 <= number of warps) and keeps the whole strip of the result in
// registers, so that results can replace an input in place after a single barrier and no global temporary is needed.
#define QX_DMMA(acc, a, b) asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"((acc)[0]), "+d"((acc)[1]) : "d"(a), "d"(b))

// Matrices have 8*NT8 rows (zero padded) and ld >= 4*ceil(n/4) columns (padding columns zero); column tiles that stick out
// of ld are predicated (only the last one does).
// A' = Ct * H * Ct^T  (H symmetric in `A`, result overwrites `A`); Ct in `Ct`.  The strip of T = Ct*H is parked in the
// warp's own rows of `A` (after a barrier: everybody has finished reading H) and read back as the A operand of the second product.
// WPS warps share one strip (each takes a contiguous range of its column tiles): 1 for the 320-thread kernels (10 warps, 9 strips),
// 2 for the wide CTAs, whose 18 warps would otherwise leave half of the tensor pipe's issue slots unused.
#define QX_WPS ((QX_NT / 32) >= 18 ? 2 : 1)
// Explicit shared-state-space accesses (32-bit shared-window addresses): with generic pointers the compiler emitted generic LD for
// some of the fragment loads of these kernels despite QX_ASSUME_SHARED (ncu: long-scoreboard stalls on the DMMA lines).
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds_f64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_v2f64(unsigned a, double x, double y) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(a), "d"(x), "d"(y) : "memory"); }

template <int NT8>
static __device__ __noinline__ void tc_transform(int n, const double *Ct, double *A, int ld) {
    constexpr int WPS = QX_WPS, TH = (NT8 + WPS - 1) / WPS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tg = lane & 3, kmax = (n + 3) & ~3;
    const int strip = warp / WPS, t0 = (warp % WPS) * TH;
    const bool act = strip < NT8;
    const unsigned ct = smem_addr(Ct), am = smem_addr(A), ld8 = 8u * (unsigned)ld;
    double acc[TH][2];
    bool okb[TH], oks[TH];
#pragma unroll
    for (int t = 0; t < TH; ++t) {
        okb[t] = t0 + t < NT8 && 8 * (t0 + t) + g < ld; oks[t] = t0 + t < NT8 && 8 * (t0 + t) + 2 * tg + 1 < ld;
        acc[t][0] = acc[t][1] = 0.0;
    }
    const unsigned myrow = am + (unsigned)(strip * 8 + g) * ld8;
    if (act) {
        const unsigned arow = ct + (unsigned)(strip * 8 + g) * ld8 + 8u * tg;   // A fragment: Ct[row][k0 + tg]
        const unsigned bcol = am + (unsigned)tg * ld8 + 8u * (g + 8 * t0);      // B fragment: H[k0 + tg][8 t + g]
#pragma unroll 2
        for (int k0 = 0; k0 < kmax; k0 += 4) {
            const double a = lds_f64(arow + 8u * k0);
#pragma unroll
            for (int t = 0; t < TH; ++t) { const double b = okb[t] ? lds_f64(bcol + (unsigned)k0 * ld8 + 64u * t) : 0.0; QX_DMMA(acc[t], a, b); }
        }
    }
    __syncthreads();   // every warp has finished reading H
    if (act) {
#pragma unroll
        for (int t = 0; t < TH; ++t) {
            if (oks[t]) sts_v2f64(myrow + 8u * (8 * (t0 + t) + 2 * tg), acc[t][0], acc[t][1]);
            acc[t][0] = acc[t][1] = 0.0;
        }
    }
    if (WPS == 1) __syncwarp(); else __syncthreads();   // the strip of T = Ct H is complete
    if (act) {
        // second product: A'[strip][col] = sum_k T[strip][k] Ct[col][k]
        const unsigned bct = ct + (unsigned)(g + 8 * t0) * ld8 + 8u * tg;      // B fragment: Ct[8 t + g][k0 + tg]
#pragma unroll 2
        for (int k0 = 0; k0 < kmax; k0 += 4) {
            const double a = lds_f64(myrow + 8u * (k0 + tg));
#pragma unroll
            for (int t = 0; t < TH; ++t) { const double b = t0 + t < NT8 ? lds_f64(bct + 8u * t * ld8 + 8u * k0) : 0.0; QX_DMMA(acc[t], a, b); }
        }
    }
    if (WPS == 1) __syncwarp(); else __syncthreads();   // everybody has read the strip of T before it is overwritten
    if (act) {
#pragma unroll
        for (int t = 0; t < TH; ++t)
            if (oks[t]) sts_v2f64(myrow + 8u * (8 * (t0 + t) + 2 * tg), acc[t][0], acc[t][1]);
    }
    __syncthreads();
}

// Ct <- X * Ct (in place), X in `X` (row-major; here the normalised rows of the Jacobi = J^T)
template <int NT8>
static __device__ __noinline__ void tc_left_apply(int n, const double *X, double *Ct, int ld) {
    constexpr int WPS = QX_WPS, TH = (NT8 + WPS - 1) / WPS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tg = lane & 3, kmax = (n + 3) & ~3;
    const int strip = warp / WPS, t0 = (warp % WPS) * TH;
    const bool act = strip < NT8;
    const unsigned xm = smem_addr(X), ct = smem_addr(Ct), ld8 = 8u * (unsigned)ld;
    double acc[TH][2];
    bool okb[TH], oks[TH];
#pragma unroll
    for (int t = 0; t < TH; ++t) {
        okb[t] = t0 + t < NT8 && 8 * (t0 + t) + g < ld; oks[t] = t0 + t < NT8 && 8 * (t0 + t) + 2 * tg + 1 < ld;
        acc[t][0] = acc[t][1] = 0.0;
    }
    if (act) {
        const unsigned arow = xm + (unsigned)(strip * 8 + g) * ld8 + 8u * tg;
        const unsigned bcol = ct + (unsigned)tg * ld8 + 8u * (g + 8 * t0);
#pragma unroll 2
        for (int k0 = 0; k0 < kmax; k0 += 4) {
            const double a = lds_f64(arow + 8u * k0);
#pragma unroll
            for (int t = 0; t < TH; ++t) { const double b = okb[t] ? lds_f64(bcol + (unsigned)k0 * ld8 + 64u * t) : 0.0; QX_DMMA(acc[t], a, b); }
        }
    }
    __syncthreads();
    if (act) {
        const unsigned orow = ct + (unsigned)(strip * 8 + g) * ld8 + 8u * (2 * tg + 8 * t0);
#pragma unroll
        for (int t = 0; t < TH; ++t)
            if (oks[t]) sts_v2f64(orow + 64u * t, acc[t][0], acc[t][1]);
    }
    __syncthreads();
}

// out = Ct^T diag(w) Ct = C diag(w) C^T.  out may alias Ct (result held in registers across a barrier).
template <int NT8>
static __device__ __noinline__ void tc_density(int n, const double *Ct, const double *w, double *out, int ld) {
    constexpr int WPS = QX_WPS, TH = (NT8 + WPS - 1) / WPS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tg = lane & 3, kmax = (n + 3) & ~3;
    const int strip = warp / WPS, t0 = (warp % WPS) * TH;
    const bool act = strip < NT8;
    const unsigned ct = smem_addr(Ct), wv = smem_addr(w), om = smem_addr(out), ld8 = 8u * (unsigned)ld;
    double acc[TH][2];
    bool okb[TH], oks[TH];
#pragma unroll
    for (int t = 0; t < TH; ++t) {
        okb[t] = t0 + t < NT8 && 8 * (t0 + t) + g < ld; oks[t] = t0 + t < NT8 && 8 * (t0 + t) + 2 * tg + 1 < ld;
        acc[t][0] = acc[t][1] = 0.0;
    }
    if (act) {
        const bool oka = strip * 8 + g < ld;
        const unsigned acol = ct + (unsigned)tg * ld8 + 8u * (oka ? strip * 8 + g : 0);   // A fragment: (Ct^T)[row][k] = Ct[k0 + tg][row]
        const unsigned bcol = ct + (unsigned)tg * ld8 + 8u * (g + 8 * t0);                // B fragment: Ct[k0 + tg][8 t + g]
#pragma unroll 2
        for (int k0 = 0; k0 < kmax; k0 += 4) {
            const double a = oka ? lds_f64(acol + (unsigned)k0 * ld8) * lds_f64(wv + 8u * (k0 + tg)) : 0.0;
#pragma unroll
            for (int t = 0; t < TH; ++t) { const double b = okb[t] ? lds_f64(bcol + (unsigned)k0 * ld8 + 64u * t) : 0.0; QX_DMMA(acc[t], a, b); }
        }
    }
    __syncthreads();
    if (act) {
        const unsigned orow = om + (unsigned)(strip * 8 + g) * ld8 + 8u * (2 * tg + 8 * t0);
#pragma unroll
        for (int t = 0; t < TH; ++t)
            if (oks[t]) sts_v2f64(orow + 64u * t, acc[t][0], acc[t][1]);
    }
    __syncthreads();
}

// padded dimension used by the strip GEMMs for a basis of n functions (0: not supported by the in-register path)
__host__ __device__ inline int tc_padded_dim(int n) {   // = 8 * NT8 of the instantiated strip kernels
    if (n <= 32) return 32;
    if (n <= 64) return 64;
    if (n <= 72) return 72;
    return 0;
}

// In-place Cholesky S = L L^T on the lower triangle of A (n x n, ld), then Ct = L^{-1} (lower triangular),
// i.e. C = L^{-T}: an S-orthonormal starting basis.  Returns false if S is not positive definite.
template <bool SH>
static __device__ __noinline__ bool cholesky_basis(int n, double *A, double *Ct, int ld, double *red) {
    QX_ASSUME_SHARED(red);
    if (SH) { QX_ASSUME_SHARED(A); QX_ASSUME_SHARED(Ct); }
    for (int j = 0; j < n; ++j) {
        double d = A[(size_t)j * ld + j];
        if (!(d > 0.0)) return false;  // uniform across the CTA (all threads read the same value)
        d = sqrt(d);
        __syncthreads();
        for (int i = j + threadIdx.x; i < n; i += QX_NT) A[(size_t)i * ld + j] = (i == j) ? d : A[(size_t)i * ld + j] / d;
        __syncthreads();
        // trailing update of the lower triangle: one warp per row, lanes over the columns j+1..i
        for (int i = j + 1 + (threadIdx.x >> 5); i < n; i += QX_NT / 32) {
            const double lij = A[(size_t)i * ld + j];
            for (int k = j + 1 + (threadIdx.x & 31); k <= i; k += 32) A[(size_t)i * ld + k] -= lij * A[(size_t)k * ld + j];
        }
        __syncthreads();
    }
    // X = L^{-1}, thread per column j (forward substitution); zero the strictly upper part
    for (int j = threadIdx.x; j < n; j += QX_NT) {
        for (int i = 0; i < j; ++i) Ct[(size_t)i * ld + j] = 0.0;
        Ct[(size_t)j * ld + j] = 1.0 / A[(size_t)j * ld + j];
        for (int i = j + 1; i < n; ++i) {
            double v = 0.0;
            for (int k = j; k < i; ++k) v -= A[(size_t)i * ld + k] * Ct[(size_t)k * ld + j];
            Ct[(size_t)i * ld + j] = v / A[(size_t)i * ld + i];
        }
    }
    __syncthreads();
    return true;
}

// plane-rotation parameters from the 2x2 Gram matrix (al, ga; ga, be): the angle is evaluated in single
// precision (it only steers convergence), c is refined to double so that c^2 + s^2 = 1 to rounding.
__device__ inline bool rotation_from_gram(double al, double be, double ga, float &ratio, double &c, double &sn) {
    const float gaf = (float)ga;
    ratio = fabsf(gaf) * rsqrtf((float)al * (float)be);
    if (!(ratio > 1e-15f)) return false;
    const float zeta = (float)(be - al) / (2.0f * gaf);
    const float tf = copysignf(1.0f, zeta) / (fabsf(zeta) + sqrtf(1.0f + zeta * zeta));
    const double t = (double)tf, w = 1.0 + t * t;
    double c0 = (double)rsqrtf((float)w);
    c0 = c0 * (1.5 - 0.5 * w * c0 * c0);
    c = c0 * (1.5 - 0.5 * w * c0 * c0);
    sn = c * t;
    return true;
}

// pair (p, q) of the round-robin tournament with mm players (mm even), slot k of round `round`
__device__ inline void tournament_pair(int mm, int round, int k, int &p, int &q) {
    if (k == 0) { p = mm - 1; q = round; }
    else { p = round + k; if (p >= mm - 1) p -= mm - 1; q = round - k; if (q < 0) q += mm - 1; }
    if (p > q) { int t = p; p = q; q = t; }
}

// One-sided (Hestenes) Jacobi on the rows of G = A' + sigma I (symmetric positive definite thanks to the
// Gershgorin shift): rows are rotated pairwise until mutually orthogonal.  At convergence row k equals
// (lambda_k + sigma) j_k, i.e. the eigenvectors are the normalised rows and the eigenvalues their norms, so
// no eigenvector matrix has to be dragged through the sweeps.
// Eight lanes own one pair per round and move the two rows through registers with 128-bit accesses; one
// __syncthreads per round.  Rotations are applied in the scaled ("fast Givens") form: row_k = d_k * stored_k,
// so an element update is a single DFMA; squared norms are tracked incrementally (recomputed every sweep),
// so only the cross product needs a reduction.  Control flow is warp-uniform (idle groups do dummy loads).
// jw: 4*n doubles of shared scratch (norms, scales, inverse scales).
// A sweep whose largest pre-rotation coupling |g_p.g_q| / (|g_p||g_q|) stays below QX_JACOBI_TOL is the last one: convergence
// is quadratic, so what it leaves behind is O(tol^2).  Measured on 592 distorted caffeine cations against tol = 1e-7
// (tools/phase_profile.py --dump): 3e-6 changes E by 2e-11 Eh, gradients by 2e-10, charges by 2e-10 and saves 8 % of the
// sweeps; 3e-5 would save 15 % but moves charges by 4e-7 (too close to the 1e-6 parity gate).
#ifndef QX_JACOBI_TOL
#define QX_JACOBI_TOL 3e-6f
#endif
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---- trimmed variant of the one-pair-per-round kernel (default).  The sub-partition pipes are what bounds the sweep
// (tools/microbench/lat.cu: ~2.1 cycles per double-precision warp instruction, ~8.2 per F2F / MUFU; ncu: >45 % of the issued
// instructions of jacobi_rows_lp8 are integer / control), so this version (1) walks the tournament incrementally
// (p and q advance by one player per round), (2) decides convergence with a double-precision compare instead of a
// single-precision ratio (flag + __syncthreads_or, no block-wide max), (3) gets the tangent from a two-MUFU chain
// t = 2 ga / (d + sign(d) sqrt(d^2 + 4 ga^2)) and c from the single-precision t: 4 F2F + 3 MUFU per rotation instead of
// 6 + 5, (4) is branch-free (t = 0 is an exact no-op), (5) keeps (scale, 1/scale) packed for 128-bit state accesses.
// jw: nrm2[n] | dd[n] as double2 (scale, inverse scale), 16-byte aligned
template <int R>
static __device__ __noinline__ int jacobi_rows_lp8t(int n, double *G, int ld, float tol, double *jw, float = 0.0f) {
    QX_ASSUME_SHARED(G); QX_ASSUME_SHARED(jw);   // n <= 72: matrices are in shared memory
    const int mm = (n + 1) & ~1, npair = mm >> 1, m1 = mm - 1;
    const int k = threadIdx.x >> 3, lsub = threadIdx.x & 7, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *nrm2 = jw;
    double2 *dd = reinterpret_cast<double2 *>(jw + ((n + 1) & ~1));
    for (int i = threadIdx.x; i < n; i += QX_NT) dd[i] = make_double2(1.0, 1.0);
    __syncthreads();
    const bool tail_ok = 2 * lsub + 16 * (R - 1) < n;
    const int tail_off = tail_ok ? 2 * lsub + 16 * (R - 1) : 0;
    // the only invalid pair is the one with the dummy player m1 == n (n odd), always in slot 0
    const bool valid = k < npair && !(k == 0 && (n & 1));
    const bool wact = (warp << 2) < npair;   // warp-uniform
    double *Gl = G + 2 * lsub, *Gt = G + tail_off;
    const double tol2 = (double)tol * (double)tol;
    int sweep = 0;
    for (; sweep < 60; ++sweep) {
        // fold the scales into the rows and refresh the norms
        for (int r = warp; r < n; r += QX_NT / 32) {
            const double d = dd[r].x;
            double acc = 0.0;
            for (int i = lane; i < n; i += 32) { const double x = G[(size_t)r * ld + i] * d; G[(size_t)r * ld + i] = x; acc = fma(x, x, acc); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            __syncwarp();
            if (lane == 0) { nrm2[r] = acc; dd[r] = make_double2(1.0, 1.0); }
        }
        __syncthreads();
        bool big = false;
        // slot 0 pairs the fixed player m1 with `round`; slot k pairs (round + k) with (round - k) (mod m1)
        int p = valid ? (k == 0 ? m1 : k) : 0, q = valid ? (k == 0 ? 0 : m1 - k) : 0;
        for (int round = 0; round < m1; ++round) {
            if (wact) {
                double *gp = Gl + p * ld, *gq = Gl + q * ld, *tp = Gt + p * ld, *tq = Gt + q * ld;
                const double2 sp = dd[p], sq = dd[q];
                const double al = nrm2[p], be = nrm2[q];
                double2 x[R], y[R];
#pragma unroll
                for (int r = 0; r < R - 1; ++r) {   // chunks r < R-1 are in range for every lane (n > 16 (R-1))
                    x[r] = *reinterpret_cast<const double2 *>(gp + 16 * r);
                    y[r] = *reinterpret_cast<const double2 *>(gq + 16 * r);
                }
                x[R - 1] = *reinterpret_cast<const double2 *>(tp);
                y[R - 1] = *reinterpret_cast<const double2 *>(tq);
                if (!tail_ok) { x[R - 1] = make_double2(0.0, 0.0); y[R - 1] = make_double2(0.0, 0.0); }
                double g0 = 0.0, g1 = 0.0;
#pragma unroll
                for (int r = 0; r < R; ++r) { g0 = fma(x[r].x, y[r].x, g0); g1 = fma(x[r].y, y[r].y, g1); }
                double gs = g0 + g1;
                gs += __shfl_xor_sync(0xffffffffu, gs, 4);
                gs += __shfl_xor_sync(0xffffffffu, gs, 2);
                gs += __shfl_xor_sync(0xffffffffu, gs, 1);
                const double ga = (sp.x * sq.x) * gs, ga2 = ga * ga, nn = al * be;
                big |= valid && ga2 > tol2 * nn;
                const bool rot = valid && ga2 > 1e-30 * nn;
                const float gf = (float)ga, df = (float)(be - al);
                const float g2 = gf + gf;
                const float hh = fmaf(df, df, g2 * g2);
                const float den = fabsf(df) + hh * rsqrt_approx(hh);          // |d| + sqrt(d^2 + 4 ga^2)
                float tf = g2 * rcp_approx(den);
                tf = __int_as_float(__float_as_int(tf) ^ (__float_as_int(df) & 0x80000000));
                tf = rot ? tf : 0.0f;    // also removes the NaN of a 0/0 pair; t == 0 leaves rows and scales untouched
                const double t = (double)tf;
                const double t1 = t * (sq.x * sp.y), t2 = t * (sp.x * sq.y);
#pragma unroll
                for (int r = 0; r < R - 1; ++r) {
                    double2 u, v;
                    u.x = fma(-t1, y[r].x, x[r].x); u.y = fma(-t1, y[r].y, x[r].y);
                    v.x = fma(t2, x[r].x, y[r].x); v.y = fma(t2, x[r].y, y[r].y);
                    if (valid) {
                        *reinterpret_cast<double2 *>(gp + 16 * r) = u;
                        *reinterpret_cast<double2 *>(gq + 16 * r) = v;
                    }
                }
                if (tail_ok && valid) {
                    double2 u, v;
                    u.x = fma(-t1, y[R - 1].x, x[R - 1].x); u.y = fma(-t1, y[R - 1].y, x[R - 1].y);
                    v.x = fma(t2, x[R - 1].x, y[R - 1].x); v.y = fma(t2, x[R - 1].y, y[R - 1].y);
                    *reinterpret_cast<double2 *>(gp + 16 * (R - 1)) = u;
                    *reinterpret_cast<double2 *>(gq + 16 * (R - 1)) = v;
                }
                {
                    const double w = fma(t, t, 1.0);
                    double c = (double)rsqrt_approx(fmaf(tf, tf, 1.0f));
                    c = c * fma(-0.5 * w * c, c, 1.5);
                    c = c * fma(-0.5 * w * c, c, 1.5);
                    const double wc = w * c, tg = t * ga;
                    if (lsub == 0 && valid) {
                        dd[p] = make_double2(c * sp.x, wc * sp.y); dd[q] = make_double2(c * sq.x, wc * sq.y);
                        nrm2[p] = al - tg; nrm2[q] = be + tg;
                    }
                }
                // next round: both players of a slot move on by one (the fixed player stays)
                if (k != 0) p = p + 1 == m1 ? 0 : p + 1;
                q = q + 1 == m1 ? 0 : q + 1;
                if (!valid) { p = 0; q = 0; }
            }
            __syncthreads();
        }
        if (!__syncthreads_or(big ? 1 : 0)) { ++sweep; break; }
    }
    // fold the remaining scales
    for (int r = warp; r < n; r += QX_NT / 32) {
        const double d = dd[r].x;
        for (int i = lane; i < n; i += 32) G[(size_t)r * ld + i] *= d;
    }
    __syncthreads();
    return sweep;
}

// predicated 128-bit shared-memory load into an existing value
__device__ __forceinline__ void lds_v2_if(bool p, double2 &v, const double *ptr) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(ptr);
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q ld.shared.v2.f64 {%0, %1}, [%2];\n\t}" : "+d"(v.x), "+d"(v.y) : "r"(a), "r"((int)p));
}
// predicated 64-bit / 128-bit shared-memory loads of the sweep's scale state.  Idle groups (no pair in this round) used to do dummy
// loads from row 0 while its owner rotated it -- harmless, their values are never used, but a read/write hazard for
// compute-sanitizer's racecheck; predicated loads touch nothing.
__device__ __forceinline__ double lds_f64_if(bool p, double dflt, const double *ptr) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(ptr);
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.shared.f64 %0, [%1];\n\t}" : "+d"(dflt) : "r"(a), "r"((int)p));
    return dflt;
}
__device__ __forceinline__ double2 lds_v2_or(bool p, double2 dflt, const double2 *ptr) {
    lds_v2_if(p, dflt, reinterpret_cast<const double *>(ptr));
    return dflt;
}
#ifndef QX_JACOBI_NOKEEP   // default; -DQX_JACOBI_NOKEEP selects jacobi_rows_lp8t
// ---- "kept row" variant of the trimmed kernel (default).  Group g plays slot k = (g - round) mod K, whose "plus" player (round + k)
// is the plus player of slot k - 1 in the next round, so the plus row stays in registers; only the "minus" rows (and both
// rows of slot 0, which wraps to slot K - 1) travel through shared memory: half the row traffic of jacobi_rows_lp8t, whose
// shared-memory pipe is 76 % busy at two CTAs per SM (ncu).  Same arithmetic in the same order: results are bitwise those of
// jacobi_rows_lp8t.  Register budget is the constraint (96 per thread): the slot bookkeeping is re-derived from k every round,
// the scale state of the kept row goes through shared memory like the other one's, and the kept row is (re)loaded with
// predicated loads -- a branch around plain assignments made the compiler keep the loop-carried array in local memory
// (65 M local loads per 1184 solves, 45 % slower than lp8t instead of 17 % faster).
template <int R>
// gate > 0: a sweep whose largest pre-rotation coupling stayed below `gate` (but not below tol) ends the call with -(sweeps): the
// caller then checks convergence on the Gram matrix and applies the last, tiny rotations as three DMMA products (jacobi_polish).
static __device__ __noinline__ int jacobi_rows_lp8r(int n, double *G, int ld, float tol, double *jw, float gate = 0.0f) {
    QX_ASSUME_SHARED(G); QX_ASSUME_SHARED(jw);
    const int mm = (n + 1) & ~1, K = mm >> 1, m1 = mm - 1;
    const int grp = threadIdx.x >> 3, lsub = threadIdx.x & 7;
    double *nrm2 = jw;
    double2 *dd = reinterpret_cast<double2 *>(jw + ((n + 1) & ~1));
    for (int i = threadIdx.x; i < n; i += QX_NT) dd[i] = make_double2(1.0, 1.0);
    __syncthreads();
    const bool tail_ok = 2 * lsub + 16 * (R - 1) < n;
    const bool gact = grp < K, wact = ((threadIdx.x >> 5) << 2) < K;   // wact is warp-uniform
    int sweep = 0;
    bool early = false;
    const double gate2 = (double)gate * (double)gate;
    for (; sweep < 60; ++sweep) {
        {   // fold the scales into the rows and refresh the norms
            const int lane = threadIdx.x & 31;
            for (int r = threadIdx.x >> 5; r < n; r += QX_NT / 32) {
                const double d = dd[r].x;
                double acc = 0.0;
                for (int i = lane; i < n; i += 32) { const double x = G[(size_t)r * ld + i] * d; G[(size_t)r * ld + i] = x; acc = fma(x, x, acc); }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                __syncwarp();
                if (lane == 0) { nrm2[r] = acc; dd[r] = make_double2(1.0, 1.0); }
            }
        }
        __syncthreads();
        bool big = false, mid = false;
        double2 x[R], y[R];
#pragma unroll
        for (int r = 0; r < R; ++r) x[r] = make_double2(0.0, 0.0);   // lanes beyond the row end keep zeros in the tail chunk
        int k = grp;                                     // slot of this group in the current round: (grp - round) mod K
        for (int round = 0; round < m1; ++round) {
            if (wact) {
                // rows of slot k: kept row ra = round + k (slot 0: round), other row rb = round - k (slot 0: the fixed player m1)
                int ra = round + k, rb = round - k;
                if (ra >= m1) ra -= m1;
                if (rb < 0) rb += m1;
                if (k == 0) rb = m1;
                const bool va = gact && ra < n, vb = gact && rb < n;
                if (!va) ra = 0;
                if (!vb) rb = 0;
                double *gp = G + ra * ld + 2 * lsub, *gq = G + rb * ld + 2 * lsub;
                {   // first round / freshly wrapped group: the kept row is new as well.  Predicated loads (not a branch around plain
                    // assignments): with a conditionally assigned loop-carried array the compiler keeps x[] in local memory.
                    const bool fresh = round == 0 || k == K - 1;
#pragma unroll
                    for (int r = 0; r < R - 1; ++r) lds_v2_if(fresh && va, x[r], gp + 16 * r);
                    lds_v2_if(fresh && va && tail_ok, x[R - 1], gp + 16 * (R - 1));
                }
#pragma unroll
                for (int r = 0; r < R; ++r) y[r] = make_double2(0.0, 0.0);
#pragma unroll
                for (int r = 0; r < R - 1; ++r) lds_v2_if(vb, y[r], gq + 16 * r);
                lds_v2_if(vb && tail_ok, y[R - 1], gq + 16 * (R - 1));
                const double2 one2 = make_double2(1.0, 1.0);
                const double2 sx = lds_v2_or(va, one2, dd + ra), sq = lds_v2_or(vb, one2, dd + rb);
                const double al = lds_f64_if(va, 1.0, nrm2 + ra), be = lds_f64_if(vb, 1.0, nrm2 + rb);
                double g0 = 0.0, g1 = 0.0;
#pragma unroll
                for (int r = 0; r < R; ++r) { g0 = fma(x[r].x, y[r].x, g0); g1 = fma(x[r].y, y[r].y, g1); }
                double gs = g0 + g1;
                gs += __shfl_xor_sync(0xffffffffu, gs, 4);
                gs += __shfl_xor_sync(0xffffffffu, gs, 2);
                gs += __shfl_xor_sync(0xffffffffu, gs, 1);
                const double ga = (sx.x * sq.x) * gs, ga2 = ga * ga, nn = al * be;
                const bool valid = va && vb;
                big |= valid && ga2 > ((double)tol * (double)tol) * nn;
                mid |= valid && ga2 > gate2 * nn;
                const bool rot = valid && ga2 > 1e-30 * nn;
                const float gf = (float)ga, df = (float)(be - al);
                const float g2 = gf + gf;
                const float hh = fmaf(df, df, g2 * g2);
                const float den = fabsf(df) + hh * rsqrt_approx(hh);
                float tf = g2 * rcp_approx(den);
                tf = __int_as_float(__float_as_int(tf) ^ (__float_as_int(df) & 0x80000000));
                tf = rot ? tf : 0.0f;
                const double t = (double)tf;
                const double t1 = t * (sq.x * sx.y), t2 = t * (sx.x * sq.y);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const double ux = fma(-t1, y[r].x, x[r].x), uy = fma(-t1, y[r].y, x[r].y);
                    y[r].x = fma(t2, x[r].x, y[r].x); y[r].y = fma(t2, x[r].y, y[r].y);
                    x[r].x = ux; x[r].y = uy;
                }
                if (vb) {
#pragma unroll
                    for (int r = 0; r < R - 1; ++r) *reinterpret_cast<double2 *>(gq + 16 * r) = y[r];
                    if (tail_ok) *reinterpret_cast<double2 *>(gq + 16 * (R - 1)) = y[R - 1];
                }
                if (va && (k == 0 || round == m1 - 1)) {   // slot 0 hands both rows on; everything goes back at the end of a sweep
#pragma unroll
                    for (int r = 0; r < R - 1; ++r) *reinterpret_cast<double2 *>(gp + 16 * r) = x[r];
                    if (tail_ok) *reinterpret_cast<double2 *>(gp + 16 * (R - 1)) = x[R - 1];
                }
                {
                    const double w = fma(t, t, 1.0);
                    double c = (double)rsqrt_approx(fmaf(tf, tf, 1.0f));
                    c = c * fma(-0.5 * w * c, c, 1.5);
                    c = c * fma(-0.5 * w * c, c, 1.5);
                    const double wc = w * c, tg = t * ga;
                    __syncwarp();   // the eight lanes of the group have read the scale state that lane 0 replaces
                    if (lsub == 0 && valid) {
                        dd[ra] = make_double2(c * sx.x, wc * sx.y); dd[rb] = make_double2(c * sq.x, wc * sq.y);
                        nrm2[ra] = al - tg; nrm2[rb] = be + tg;
                    }
                }
                k = k == 0 ? K - 1 : k - 1;
            }
            __syncthreads();
        }
        if (!__syncthreads_or(big ? 1 : 0)) { ++sweep; break; }
        if (gate > 0.0f && !__syncthreads_or(mid ? 1 : 0)) { ++sweep; early = true; break; }
    }
    {   // fold the remaining scales
        const int lane = threadIdx.x & 31;
        for (int r = threadIdx.x >> 5; r < n; r += QX_NT / 32) {
            const double d = dd[r].x;
            for (int i = lane; i < n; i += 32) G[(size_t)r * ld + i] *= d;
        }
    }
    __syncthreads();
    return early ? -sweep : sweep;
}
#define QX_JROWS jacobi_rows_lp8r
#else
#define QX_JROWS jacobi_rows_lp8t
#endif

// ---- multi-pass variant for bases between the one-pass limit (72 AOs: 36 groups of 8 lanes) and the shared-memory limit
// (~110 AOs, one CTA per SM): the ceil(n/2) pairs of a round are worked off in passes of QX_NT/8 groups.  Same rotation as
// jacobi_rows_lp8t; no kept row (a group would have to keep one per pass).
template <int R>
static __device__ __noinline__ int jacobi_rows_lp8m(int n, double *G, int ld, float tol, double *jw, float gate = 0.0f) {   // gate: as jacobi_rows_lp8r
    QX_ASSUME_SHARED(G); QX_ASSUME_SHARED(jw);
    const int mm = (n + 1) & ~1, npair = mm >> 1, m1 = mm - 1;
    const int nslot = QX_NT / 8, slot = threadIdx.x >> 3, lsub = threadIdx.x & 7, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npass = (npair + nslot - 1) / nslot;
    double *nrm2 = jw;
    double2 *dd = reinterpret_cast<double2 *>(jw + ((n + 1) & ~1));
    for (int i = threadIdx.x; i < n; i += QX_NT) dd[i] = make_double2(1.0, 1.0);
    __syncthreads();
    const bool tail_ok = 2 * lsub + 16 * (R - 1) < n;
    double *Gl = G + 2 * lsub;
    const double tol2 = (double)tol * (double)tol, gate2 = (double)gate * (double)gate;
    int sweep = 0;
    bool early = false;
    for (; sweep < 60; ++sweep) {
        for (int r = warp; r < n; r += QX_NT / 32) {
            const double d = dd[r].x;
            double acc = 0.0;
            for (int i = lane; i < n; i += 32) { const double x = G[(size_t)r * ld + i] * d; G[(size_t)r * ld + i] = x; acc = fma(x, x, acc); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            __syncwarp();
            if (lane == 0) { nrm2[r] = acc; dd[r] = make_double2(1.0, 1.0); }
        }
        __syncthreads();
        bool big = false, mid = false;
        for (int round = 0; round < m1; ++round) {
            for (int pass = 0; pass < npass; ++pass) {
                if ((warp << 2) + pass * nslot >= npair) continue;   // warp-uniform
                const int k = slot + pass * nslot;
                int p = round + k, q = round - k;
                if (p >= m1) p -= m1;
                if (q < 0) q += m1;
                if (k == 0) p = m1;
                const bool valid = k < npair && p < n && q < n;
                if (!valid) { p = 0; q = 0; }
                double *gp = Gl + p * ld, *gq = Gl + q * ld;
                const double2 one2 = make_double2(1.0, 1.0);
                const double2 sp = lds_v2_or(valid, one2, dd + p), sq = lds_v2_or(valid, one2, dd + q);
                const double al = lds_f64_if(valid, 1.0, nrm2 + p), be = lds_f64_if(valid, 1.0, nrm2 + q);
                double2 x[R], y[R];
#pragma unroll
                for (int r = 0; r < R; ++r) { x[r] = make_double2(0.0, 0.0); y[r] = make_double2(0.0, 0.0); }
#pragma unroll
                for (int r = 0; r < R - 1; ++r) {
                    lds_v2_if(valid, x[r], gp + 16 * r);
                    lds_v2_if(valid, y[r], gq + 16 * r);
                }
                lds_v2_if(valid && tail_ok, x[R - 1], gp + 16 * (R - 1));
                lds_v2_if(valid && tail_ok, y[R - 1], gq + 16 * (R - 1));
                double g0 = 0.0, g1 = 0.0;
#pragma unroll
                for (int r = 0; r < R; ++r) { g0 = fma(x[r].x, y[r].x, g0); g1 = fma(x[r].y, y[r].y, g1); }
                double gs = g0 + g1;
                gs += __shfl_xor_sync(0xffffffffu, gs, 4);
                gs += __shfl_xor_sync(0xffffffffu, gs, 2);
                gs += __shfl_xor_sync(0xffffffffu, gs, 1);
                const double ga = (sp.x * sq.x) * gs, ga2 = ga * ga, nn = al * be;
                big |= valid && ga2 > tol2 * nn;
                mid |= valid && ga2 > gate2 * nn;
                const bool rot = valid && ga2 > 1e-30 * nn;
                const float gf = (float)ga, df = (float)(be - al);
                const float g2 = gf + gf;
                const float hh = fmaf(df, df, g2 * g2);
                const float den = fabsf(df) + hh * rsqrt_approx(hh);
                float tf = g2 * rcp_approx(den);
                tf = __int_as_float(__float_as_int(tf) ^ (__float_as_int(df) & 0x80000000));
                tf = rot ? tf : 0.0f;
                const double t = (double)tf;
                const double t1 = t * (sq.x * sp.y), t2 = t * (sp.x * sq.y);
                if (valid) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        double2 u, v;
                        u.x = fma(-t1, y[r].x, x[r].x); u.y = fma(-t1, y[r].y, x[r].y);
                        v.x = fma(t2, x[r].x, y[r].x); v.y = fma(t2, x[r].y, y[r].y);
                        if (r < R - 1 || tail_ok) {
                            *reinterpret_cast<double2 *>(gp + 16 * r) = u;
                            *reinterpret_cast<double2 *>(gq + 16 * r) = v;
                        }
                    }
                }
                const double w = fma(t, t, 1.0);
                double c = (double)rsqrt_approx(fmaf(tf, tf, 1.0f));
                c = c * fma(-0.5 * w * c, c, 1.5);
                c = c * fma(-0.5 * w * c, c, 1.5);
                const double wc = w * c, tg = t * ga;
                __syncwarp();   // the eight lanes of the group have read the scale state that lane 0 replaces
                if (lsub == 0 && valid) {
                    dd[p] = make_double2(c * sp.x, wc * sp.y); dd[q] = make_double2(c * sq.x, wc * sq.y);
                    nrm2[p] = al - tg; nrm2[q] = be + tg;
                }
            }
            __syncthreads();
        }
        if (!__syncthreads_or(big ? 1 : 0)) { ++sweep; break; }
        if (gate > 0.0f && !__syncthreads_or(mid ? 1 : 0)) { ++sweep; early = true; break; }
    }
    for (int r = warp; r < n; r += QX_NT / 32) {
        const double d = dd[r].x;
        for (int i = lane; i < n; i += 32) G[(size_t)r * ld + i] *= d;
    }
    __syncthreads();
    return early ? -sweep : sweep;
}

// ---- large bases (matrices in the CTA's global slab, L2-resident): LP = 16 or 32 lanes per row pair, 128-bit coalesced row
// accesses (a group reads 16 LP contiguous bytes per request), the rows of the group's next pair of the round prefetched into L1
// while the current pair is rotated.  Same scaled rotation, norm tracking and stop rule as jacobi_rows_lp8m; n <= 2 LP R.
// The sweep is bound by the latency of the dependent chain load -> dot -> shuffles -> rotation parameters -> store with only the
// CTA's 9 warps to hide it: two pairs per warp (LP = 16) halve the number of passes of a round.
template <int R, int LP>
static __device__ __noinline__ int jacobi_rows_glob(int n, double *G, int ld, float tol, double *jw) {
    QX_ASSUME_SHARED(jw);
    const int mm = (n + 1) & ~1, npair = mm >> 1, m1 = mm - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = QX_NT / 32;
    const int nslot = QX_NT / LP, slot = threadIdx.x / LP, lsub = threadIdx.x & (LP - 1);
    const int npass = (npair + nslot - 1) / nslot;
    double *nrm2 = jw;
    double2 *dd = reinterpret_cast<double2 *>(jw + ((n + 1) & ~1));
    for (int i = threadIdx.x; i < n; i += QX_NT) dd[i] = make_double2(1.0, 1.0);
    __syncthreads();
    double *Gl = G + 2 * lsub;
    const bool tail_ok = 2 * lsub + 2 * LP * (R - 1) < n, tail_odd = 2 * lsub + 2 * LP * (R - 1) + 1 == n;
    const double tol2 = (double)tol * (double)tol;
    int sweep = 0;
    for (; sweep < 60; ++sweep) {
        for (int r = warp; r < n; r += nwarp) {
            const double d = dd[r].x;
            double acc = 0.0;
            for (int i = lane; i < n; i += 32) { const double x = G[(size_t)r * ld + i] * d; G[(size_t)r * ld + i] = x; acc = fma(x, x, acc); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            __syncwarp();
            if (lane == 0) { nrm2[r] = acc; dd[r] = make_double2(1.0, 1.0); }
        }
        __syncthreads();
        bool big = false;
        for (int round = 0; round < m1; ++round) {
            for (int pass = 0; pass < npass; ++pass) {
                if ((warp * 32) / LP + pass * nslot >= npair) continue;   // warp-uniform: no group of this warp has a pair
                const int k = slot + pass * nslot;
                int p = round + k, q = round - k;
                if (p >= m1) p -= m1;
                if (q < 0) q += m1;
                if (k == 0) p = m1;
                const bool valid = k < npair && p < n && q < n;
                if (!valid) { p = 0; q = 0; }
                {   // rows of this group's next pair of the round -> L1
                    const int k2 = k + nslot;
                    int p2 = round + k2, q2 = round - k2;
                    if (p2 >= m1) p2 -= m1;
                    if (q2 < 0) q2 += m1;
                    if (k2 < npair && p2 < n && q2 < n) {
#pragma unroll
                        for (int r = 0; r < R; ++r)
                            if (2 * lsub + 2 * LP * r < n) {
                                asm volatile("prefetch.global.L1 [%0];" ::"l"(Gl + (size_t)p2 * ld + 2 * LP * r));
                                asm volatile("prefetch.global.L1 [%0];" ::"l"(Gl + (size_t)q2 * ld + 2 * LP * r));
                            }
                    }
                }
                double *gp = Gl + (size_t)p * ld, *gq = Gl + (size_t)q * ld;
                const double2 sp = dd[p], sq = dd[q];
                const double al = nrm2[p], be = nrm2[q];
                double2 x[R], y[R];
                // every chunk but the last is in range for all lanes (2 LP (R - 1) < n for the R the dispatcher picks)
#pragma unroll
                for (int r = 0; r < R - 1; ++r) {
                    x[r] = *reinterpret_cast<const double2 *>(gp + 2 * LP * r);
                    y[r] = *reinterpret_cast<const double2 *>(gq + 2 * LP * r);
                }
                x[R - 1] = make_double2(0.0, 0.0); y[R - 1] = make_double2(0.0, 0.0);
                if (tail_ok) { x[R - 1] = *reinterpret_cast<const double2 *>(gp + 2 * LP * (R - 1)); y[R - 1] = *reinterpret_cast<const double2 *>(gq + 2 * LP * (R - 1)); }
                if (tail_odd) { x[R - 1].y = 0.0; y[R - 1].y = 0.0; }   // padding column of an odd dimension
                double g0 = 0.0, g1 = 0.0;
#pragma unroll
                for (int r = 0; r < R; ++r) { g0 = fma(x[r].x, y[r].x, g0); g1 = fma(x[r].y, y[r].y, g1); }
                double gs = g0 + g1;
#pragma unroll
                for (int o = LP >> 1; o > 0; o >>= 1) gs += __shfl_xor_sync(0xffffffffu, gs, o);
                const double ga = (sp.x * sq.x) * gs, ga2 = ga * ga, nn = al * be;
                big |= valid && ga2 > tol2 * nn;
                const bool rot = valid && ga2 > 1e-30 * nn;
                const float gf = (float)ga, df = (float)(be - al);
                const float g2 = gf + gf;
                const float hh = fmaf(df, df, g2 * g2);
                const float den = fabsf(df) + hh * rsqrt_approx(hh);
                float tf = g2 * rcp_approx(den);
                tf = __int_as_float(__float_as_int(tf) ^ (__float_as_int(df) & 0x80000000));
                tf = rot ? tf : 0.0f;
                const double t = (double)tf;
                const double t1 = t * (sq.x * sp.y), t2 = t * (sp.x * sq.y);
                if (valid) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        double2 u, v;
                        u.x = fma(-t1, y[r].x, x[r].x); u.y = fma(-t1, y[r].y, x[r].y);
                        v.x = fma(t2, x[r].x, y[r].x); v.y = fma(t2, x[r].y, y[r].y);
                        if (r < R - 1 || tail_ok) {
                            *reinterpret_cast<double2 *>(gp + 2 * LP * r) = u;
                            *reinterpret_cast<double2 *>(gq + 2 * LP * r) = v;
                        }
                    }
                }
                const double w = fma(t, t, 1.0);
                double c = (double)rsqrt_approx(fmaf(tf, tf, 1.0f));
                c = c * fma(-0.5 * w * c, c, 1.5);
                c = c * fma(-0.5 * w * c, c, 1.5);
                const double wc = w * c, tg = t * ga;
                if (lsub == 0 && valid) {
                    dd[p] = make_double2(c * sp.x, wc * sp.y); dd[q] = make_double2(c * sq.x, wc * sq.y);
                    nrm2[p] = al - tg; nrm2[q] = be + tg;
                }
            }
            __syncthreads();
        }
        if (!__syncthreads_or(big ? 1 : 0)) { ++sweep; break; }
    }
    for (int r = warp; r < n; r += nwarp) {
        const double d = dd[r].x;
        for (int i = lane; i < n; i += 32) G[(size_t)r * ld + i] *= d;
    }
    __syncthreads();
    return sweep;
}

#define QX_JB_LANES 16
// ---- large bases, blocked: the rows of G are cut into blocks of b <= bmax rows; a round-robin tournament over the blocks brings
// two blocks at a time into the shared-memory buffer B (2 bmax rows), where one full sweep over their rows runs at shared-memory
// latency (16 lanes per pair, same rotation as above), and writes them back.  An outer sweep visits every pair of blocks once and
// rotates every pair of rows exactly once.  L2 traffic per outer sweep: (nb - 1) reads and writes of G instead of n - 1.
// Columns n <= 32 R.
template <int R>
static __device__ __noinline__ int jacobi_rows_blocked(int n, double *G, int ld, float tol, double *jw, double *B, int bmax, float gate = 0.0f) {
    QX_ASSUME_SHARED(jw); QX_ASSUME_SHARED(B);
    // 512-thread CTAs have 128 registers per thread: the row pair stays in registers between dot product and rotation (two passes
    // over shared memory per rotation instead of three; at 96 registers this spills and is slower: 486 vs 672 peptide single points/s)
    constexpr bool HOLD = QX_NT == 512;
    constexpr int LP = QX_JB_LANES;   // lanes per row pair: a pass rotates QX_NT / 16 pairs
    const int nb = (n + bmax - 1) / bmax, b = (n + nb - 1) / nb, nbe = (nb + 1) & ~1, bm1 = nbe - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = QX_NT / 32;
    const int nslot = QX_NT / LP, slot = threadIdx.x / LP, lsub = threadIdx.x & (LP - 1);
    double *nrm2 = jw;
    double2 *dd = reinterpret_cast<double2 *>(jw + ((2 * b + 1) & ~1));
    double *Bl = B + 2 * lsub;
    const bool tail_ok = 2 * lsub + 2 * LP * (R - 1) < n;
    const double tol2 = (double)tol * (double)tol, gate2 = (double)gate * (double)gate;
    int sweep = 0;
    bool early = false;
    for (; sweep < 60; ++sweep) {
        bool big = false, mid = false;
        for (int bround = 0; bround < bm1; ++bround) {
            for (int bk = 0; bk < (nbe >> 1); ++bk) {
                int I = bround + bk, J = bround - bk;
                if (I >= bm1) I -= bm1;
                if (J < 0) J += bm1;
                if (bk == 0) I = bm1;
                // the first block round pairs every block exactly once: its visits rotate all pairs of their rows (inside the two
                // blocks and across), the later block rounds only the pairs across the two blocks.  The bye of an odd number of
                // blocks is a visit of one block alone in the first round and skipped later (uniform).
                const bool full = bround == 0;
                if (I >= nb) { const int t_ = I; I = J; J = t_; }
                if (I >= nb || (J >= nb && !full)) continue;
                const int r0I = I * b, r0J = J * b;
                const int nI = (n - r0I < b ? n - r0I : b), nJ = J >= nb ? 0 : (n - r0J < b ? n - r0J : b), nr = nI + nJ;
                // blocks -> shared memory, squared row norms on the way
                for (int lr = warp; lr < nr; lr += nwarp) {
                    const double *src = G + (size_t)(lr < nI ? r0I + lr : r0J + lr - nI) * ld;
                    double *dst = B + (size_t)lr * ld;
                    double acc = 0.0, xs[10];   // ld <= 320 (n <= 316): all loads of a row in flight at once, one L2 round trip per row
#pragma unroll
                    for (int c = 0; c < 10; ++c) { const int i = lane + 32 * c; xs[c] = i < n ? src[i] : 0.0; }
#pragma unroll
                    for (int c = 0; c < 10; ++c) { const int i = lane + 32 * c; if (i < ld) { dst[i] = xs[c]; acc = fma(xs[c], xs[c], acc); } }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    if (lane == 0) { nrm2[lr] = acc; dd[lr] = make_double2(1.0, 1.0); }
                }
                __syncthreads();
                const int mm = (nr + 1) & ~1, m1 = mm - 1, mx = nI > nJ ? nI : nJ;
                const int npair = full ? mm >> 1 : mx, nround = full ? m1 : mx, npass = (npair + nslot - 1) / nslot;
                for (int round = 0; round < nround; ++round) {
                    for (int pass = 0; pass < npass; ++pass) {
                        if ((warp * 32) / LP + pass * nslot >= npair) continue;   // warp-uniform
                        const int k = slot + pass * nslot;
                        int p, q;
                        bool valid;
                        if (full) {   // round-robin tournament over the nr rows
                            p = round + k; q = round - k;
                            if (p >= m1) p -= m1;
                            if (q < 0) q += m1;
                            if (k == 0) p = m1;
                            valid = k < npair && p < nr && q < nr;
                        } else {      // row k of the first block with row (k + round) mod mx of the second
                            int qq = k + round;
                            if (qq >= mx) qq -= mx;
                            p = k; q = nI + qq;
                            valid = k < nI && qq < nJ;
                        }
                        if (!valid) { p = 0; q = 0; }
                        double *gp = Bl + p * ld, *gq = Bl + q * ld;
                        const double2 one2 = make_double2(1.0, 1.0);
                        const double2 sp = lds_v2_or(valid, one2, dd + p), sq = lds_v2_or(valid, one2, dd + q);
                        const double al = lds_f64_if(valid, 1.0, nrm2 + p), be = lds_f64_if(valid, 1.0, nrm2 + q);
                        // the rows are read twice from shared memory (dot product, then rotation) instead of being held in 4 R
                        // registers across the rotation-parameter chain: as a __noinline__ callee this function only gets the registers
                        // its callers leave, and holding the rows spilled them to local memory (1.0 G LDL / 0.6 G STL warp instructions
                        // per 148 single points of C32H66, the hottest stalls of the phase)
                        double g0 = 0.0, g1 = 0.0;
                        double2 xh[HOLD ? R : 1], yh[HOLD ? R : 1];
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            if (r < R - 1 || tail_ok) {   // columns n .. ld-1 of B are zero
                                double2 x = make_double2(0.0, 0.0), y = x;   // (idle groups load nothing)
                                lds_v2_if(valid, x, gp + 2 * LP * r);
                                lds_v2_if(valid, y, gq + 2 * LP * r);
                                g0 = fma(x.x, y.x, g0); g1 = fma(x.y, y.y, g1);
                                if (HOLD) { xh[r] = x; yh[r] = y; }
                            }
                        }
                        double gs = g0 + g1;
#pragma unroll
                        for (int o = LP >> 1; o > 0; o >>= 1) gs += __shfl_xor_sync(0xffffffffu, gs, o);
                        const double ga = (sp.x * sq.x) * gs, ga2 = ga * ga, nn = al * be;
                        big |= valid && ga2 > tol2 * nn;
                        mid |= valid && ga2 > gate2 * nn;
                        const bool rot = valid && ga2 > 1e-30 * nn;
                        const float gf = (float)ga, df = (float)(be - al);
                        const float g2 = gf + gf;
                        const float hh = fmaf(df, df, g2 * g2);
                        const float den = fabsf(df) + hh * rsqrt_approx(hh);
                        float tf = g2 * rcp_approx(den);
                        tf = __int_as_float(__float_as_int(tf) ^ (__float_as_int(df) & 0x80000000));
                        tf = rot ? tf : 0.0f;
                        const double t = (double)tf;
                        const double t1 = t * (sq.x * sp.y), t2 = t * (sp.x * sq.y);
                        if (!HOLD) asm volatile("" ::: "memory");   // the rows are loaded again, not carried over from the dot product
                        if (valid) {
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                if (r < R - 1 || tail_ok) {
                                    const double2 x = HOLD ? xh[r] : *reinterpret_cast<const double2 *>(gp + 2 * LP * r);
                                    const double2 y = HOLD ? yh[r] : *reinterpret_cast<const double2 *>(gq + 2 * LP * r);
                                    double2 u, v;
                                    u.x = fma(-t1, y.x, x.x); u.y = fma(-t1, y.y, x.y);
                                    v.x = fma(t2, x.x, y.x); v.y = fma(t2, x.y, y.y);
                                    *reinterpret_cast<double2 *>(gp + 2 * LP * r) = u;
                                    *reinterpret_cast<double2 *>(gq + 2 * LP * r) = v;
                                }
                            }
                        }
                        const double w = fma(t, t, 1.0);
                        double c = (double)rsqrt_approx(fmaf(tf, tf, 1.0f));
                        c = c * fma(-0.5 * w * c, c, 1.5);
                        c = c * fma(-0.5 * w * c, c, 1.5);
                        const double wc = w * c, tg = t * ga;
                        __syncwarp();   // the lanes of the group have read the scale state that lane 0 replaces
                        if (lsub == 0 && valid) {
                            dd[p] = make_double2(c * sp.x, wc * sp.y); dd[q] = make_double2(c * sq.x, wc * sq.y);
                            nrm2[p] = al - tg; nrm2[q] = be + tg;
                        }
                    }
                    __syncthreads();
                }
                // blocks back to the slab, scales applied
                for (int lr = warp; lr < nr; lr += nwarp) {
                    double *dst = G + (size_t)(lr < nI ? r0I + lr : r0J + lr - nI) * ld;
                    const double *src = B + (size_t)lr * ld;
                    const double d = dd[lr].x;
                    for (int i = lane; i < n; i += 32) dst[i] = src[i] * d;
                }
                __syncthreads();
            }
        }
        if (!__syncthreads_or(big ? 1 : 0)) { ++sweep; break; }
        if (gate > 0.0f && !__syncthreads_or(mid ? 1 : 0)) { ++sweep; early = true; break; }
    }
    return early ? -sweep : sweep;
}

// generic fallback (any n): LP lanes per pair, scalar accesses
static __device__ __noinline__ int jacobi_rows_generic(int n, double *G, int ld, double *red, float tol) {
    const int mm = (n + 1) & ~1, npair = mm >> 1;
    int LP = 32;
    while (LP > 4 && npair * LP > QX_NT) LP >>= 1;
    const int nslot = QX_NT / LP, slot = threadIdx.x / LP, lsub = threadIdx.x & (LP - 1), lane = threadIdx.x & 31;
    const unsigned gmask = LP == 32 ? 0xffffffffu : (((1u << LP) - 1u) << (lane & ~(LP - 1)));
    int sweep = 0;
    for (; sweep < 60; ++sweep) {
        float smax = 0.0f;
        for (int round = 0; round < mm - 1; ++round) {
            for (int k = slot; k < npair; k += nslot) {
                int p, q;
                tournament_pair(mm, round, k, p, q);
                if (q >= n) continue;
                double *gp = G + (size_t)p * ld, *gq = G + (size_t)q * ld;
                double al = 0.0, be = 0.0, ga = 0.0;
                for (int i = lsub; i < n; i += LP) {
                    const double x = gp[i], y = gq[i];
                    al += x * x; be += y * y; ga += x * y;
                }
                for (int o = LP >> 1; o > 0; o >>= 1) {
                    al += __shfl_xor_sync(gmask, al, o);
                    be += __shfl_xor_sync(gmask, be, o);
                    ga += __shfl_xor_sync(gmask, ga, o);
                }
                float ratio;
                double c, sn;
                if (rotation_from_gram(al, be, ga, ratio, c, sn)) {
                    for (int i = lsub; i < n; i += LP) {
                        const double x = gp[i], y = gq[i];
                        gp[i] = c * x - sn * y;
                        gq[i] = sn * x + c * y;
                    }
                }
                smax = fmaxf(smax, ratio);
            }
            __syncthreads();
        }
        const float m = (float)block_max((double)smax, red);
        if (m < tol) { ++sweep; break; }
    }
    return sweep;
}

// The eigen-solver is three calls made from the SAME level as the other phases (not one wrapper that calls the sweep kernel):
// a __noinline__ callee only gets the registers its callers do not keep live across the call, and one more call level
// in between was enough to make the register-hungry sweep kernels spill inside the round loop.
// (1) Gershgorin shift: makes G positive definite, so that singular values == eigenvalues + sigma; sigma is parked in red[60]
template <bool SH>
static __device__ __noinline__ void jacobi_shift(int n, double *G, int ld, double *red) {
    QX_ASSUME_SHARED(red);
    if (SH) QX_ASSUME_SHARED(G);
    double rowsum = 0.0;
    for (int i = threadIdx.x >> 5; i < n; i += QX_NT / 32) {   // one warp per row
        double v = 0.0;
        for (int j = threadIdx.x & 31; j < n; j += 32) v += fabs(G[(size_t)i * ld + j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        rowsum = fmax(rowsum, v);
    }
    const double sigma = 1.0625 * block_max(rowsum, red) + 0.5;
    for (int i = threadIdx.x; i < n; i += QX_NT) G[(size_t)i * ld + i] += sigma;
    __syncthreads();
    if (threadIdx.x == 0) red[60] = sigma;
    __syncthreads();
}


// ------------------------------------------------------------------------------------ the last sweep as three DMMA products
// Every eigen-decomposition ends with a sweep whose pre-rotation couplings are all below QX_JACOBI_TOL: it confirms convergence and
// applies rotations of <= 1e-4 rad.  Both are cheaper on the tensor pipe: the couplings are the off-diagonal elements of the Gram
// matrix M = G G^T (one product), and the rotations, to second order in the antisymmetric Theta_ij = -t_ij (t_ij: the tangent the sweep
// would use, from M_ij, M_ii, M_jj), are G <- G + Theta (G + Theta G / 2) (two products; what they leave is O(theta^3)).  Pairs with a
// larger angle -- near-degenerate rows -- are left out of Theta and rotated exactly afterwards (a short list).  Numerical study on the
// SCC matrices of a distorted caffeine cation (all 12 cycles): eigenvalues to 1e-14, eigenvector residuals <= 5e-13, orthogonality
// <= 3e-12 -- the figures of the sweep it replaces.  Needs one shared-memory matrix besides G: the caller's C^T is parked in the
// CTA's global slab meanwhile.
#ifdef QX_PROFILE_PHASES
static __device__ unsigned long long g_sub_cycles[16];
#define QX_PSUB_BEGIN() long long psub_t0_ = clock64()
#define QX_PSUB(idx) do { if (threadIdx.x == 0) { long long t_ = clock64(); atomicAdd(&g_sub_cycles[idx], (unsigned long long)(t_ - psub_t0_)); psub_t0_ = t_; } } while (0)
#define QX_PCOUNT(idx) do { if (threadIdx.x == 0) atomicAdd(&g_sub_cycles[idx], 1ull); } while (0)
#else
#define QX_PSUB_BEGIN() do {} while (0)
#define QX_PSUB(idx) do {} while (0)
#define QX_PCOUNT(idx) do {} while (0)
#endif
#ifndef QX_POLISH_GATE
#define QX_POLISH_GATE 1e-2f   // try the Gram check after a sweep whose couplings all stayed below this (what it leaves is ~ 1e-4)
#endif
#define QX_POLISH_APPLY 1e-3   // couplings all below this: the correction is applied even if they are not yet below the tolerance
                               // (quadratic convergence: the next check finds ~ 1e-8) -- a polish iteration instead of a sweep
#define QX_POLISH_TBIG 1e-4    // tangents above this are rotated exactly, one after the other
#define QX_POLISH_CAP 64       // at most this many of them (otherwise: back to the sweeps)
#define QX_POLISH_T2 1e-6      // all tangents below this: the second-order term (<= 1e-12) is dropped, one product instead of two

// Mg (global, row stride ld) = G G^T
template <int NT8>
static __device__ __noinline__ void tc_gram(int n, const double *G, double *Mg, int ld) {
    constexpr int WPS = QX_WPS, TH = (NT8 + WPS - 1) / WPS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tg = lane & 3, kmax = (n + 3) & ~3;
    const int strip = warp / WPS, t0 = (warp % WPS) * TH;
    const bool act = strip < NT8;
    const unsigned gm = smem_addr(G), ld8 = 8u * (unsigned)ld;
    double acc[TH][2];
#pragma unroll
    for (int t = 0; t < TH; ++t) acc[t][0] = acc[t][1] = 0.0;
    if (act) {
        const unsigned arow = gm + (unsigned)(strip * 8 + g) * ld8 + 8u * tg;   // A fragment: G[row][k0 + tg]
        const unsigned brow = gm + (unsigned)(g + 8 * t0) * ld8 + 8u * tg;      // B fragment: G[8 t + g][k0 + tg]
#pragma unroll 2
        for (int k0 = 0; k0 < kmax; k0 += 4) {
            const double a = lds_f64(arow + 8u * k0);
#pragma unroll
            for (int t = 0; t < TH; ++t) { const double b = t0 + t < NT8 ? lds_f64(brow + 8u * t * ld8 + 8u * k0) : 0.0; QX_DMMA(acc[t], a, b); }
        }
        double *orow = Mg + (size_t)(strip * 8 + g) * ld + 2 * tg + 8 * t0;
#pragma unroll
        for (int t = 0; t < TH; ++t)
            if (t0 + t < NT8 && 8 * (t0 + t) + 2 * tg + 1 < ld) *reinterpret_cast<double2 *>(orow + 8 * t) = make_double2(acc[t][0], acc[t][1]);
    }
    __syncthreads();
}

// The Gram product with the tangents taken straight from the accumulator fragments: lane (g, tg) of the warp that owns strip s holds
// M[8 s + g][8 t + 2 tg + e]; the diagonal goes to shared memory (dg), then every lane turns its elements into Theta and stores them to
// the global slab (row stride ld; padding and diagonal as zeros) -- M itself never leaves the registers.  flags: bit 0 a coupling above
// tol, bit 1 above QX_POLISH_APPLY, bit 2 a tangent above QX_POLISH_T2 (this thread's elements; the caller combines them).
template <int NT8>
static __device__ __noinline__ int tc_gram_theta(int n, const double *G, double *Thg, int ld, double *dg, double tol2, int *cnt, int *list) {
    constexpr int WPS = QX_WPS, TH = (NT8 + WPS - 1) / WPS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tg = lane & 3, kmax = (n + 3) & ~3;
    const int strip = warp / WPS, t0 = (warp % WPS) * TH;
    const bool act = strip < NT8;
    const unsigned gm = smem_addr(G), ld8 = 8u * (unsigned)ld;
    double acc[TH][2];
#pragma unroll
    for (int t = 0; t < TH; ++t) acc[t][0] = acc[t][1] = 0.0;
    if (act) {
        const unsigned arow = gm + (unsigned)(strip * 8 + g) * ld8 + 8u * tg;   // A fragment: G[row][k0 + tg]
        const unsigned brow = gm + (unsigned)(g + 8 * t0) * ld8 + 8u * tg;      // B fragment: G[8 t + g][k0 + tg]
#pragma unroll 2
        for (int k0 = 0; k0 < kmax; k0 += 4) {
            const double a = lds_f64(arow + 8u * k0);
#pragma unroll
            for (int t = 0; t < TH; ++t) { const double b = t0 + t < NT8 ? lds_f64(brow + 8u * t * ld8 + 8u * k0) : 0.0; QX_DMMA(acc[t], a, b); }
        }
        // diagonal element (i, i), i = 8 strip + g: tile `strip`, held by the lane with 2 tg + e == g
#pragma unroll
        for (int t = 0; t < TH; ++t)
            if (t0 + t == strip && tg == (g >> 1) && strip * 8 + g < n) dg[strip * 8 + g] = (g & 1) ? acc[t][1] : acc[t][0];
    }
    __syncthreads();
    int flags = 0;
    if (act) {
        const int i = strip * 8 + g;
        const double di = i < n ? dg[i] : 0.0;
        double *orow = Thg + (size_t)i * ld + 2 * tg + 8 * t0;
#pragma unroll
        for (int t = 0; t < TH; ++t) {
            double th[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = 8 * (t0 + t) + 2 * tg + e;
                float tf = 0.0f;
                if (t0 + t < NT8 && i < n && j < n && i != j) {
                    const bool up = i < j;
                    const double gij = acc[t][e], dj = dg[j], a = up ? di : dj, b = up ? dj : di;
                    const double g2d = gij * gij, nn = a * b;
                    flags |= g2d > tol2 * nn ? 1 : 0;
                    flags |= g2d > (QX_POLISH_APPLY * QX_POLISH_APPLY) * nn ? 2 : 0;
                    // the sweep's tangent (single precision): t = 2 g / (|d| + sqrt(d^2 + 4 g^2)) with the sign of d = b - a
                    const float gf = (float)gij, df = (float)(b - a);
                    const float g2 = gf + gf;
                    const float hh = fmaf(df, df, g2 * g2);
                    const float den = fabsf(df) + hh * rsqrt_approx(hh);
                    tf = g2 * rcp_approx(den);
                    tf = __int_as_float(__float_as_int(tf) ^ (__float_as_int(df) & 0x80000000));
                    if (!(g2d > 1e-30 * nn)) tf = 0.0f;
                    if (fabsf(tf) > (float)QX_POLISH_TBIG) {
                        if (up) {
                            const int idx = atomicAdd(cnt, 1);
                            if (idx < QX_POLISH_CAP) list[idx] = (i << 16) | j;
                        }
                        tf = 0.0f;
                    }
                    flags |= fabsf(tf) > (float)QX_POLISH_T2 ? 4 : 0;
                    if (up) tf = -tf;   // row min gets -t row max, row max gets +t row min: the sweep's rotation to first order
                }
                th[e] = (double)tf;
            }
            if (t0 + t < NT8 && 8 * (t0 + t) + 2 * tg + 1 < ld) *reinterpret_cast<double2 *>(orow + 8 * t) = make_double2(th[0], th[1]);
        }
    }
    return flags;
}

// MODE 0: Out = G + (Th B) / 2 with B = G (Out: another matrix);  MODE 1: G += Th B with B = X;  MODE 2: G += Th G in place (first
// order only: the products of all warps are complete before anybody stores).  Th: global, row stride ld.
template <int NT8, int MODE>
static __device__ __noinline__ void tc_polish_apply(int n, const double *Th, const double *B, double *G, double *Out, int ld) {
    constexpr int WPS = QX_WPS, TH = (NT8 + WPS - 1) / WPS, KS = 2 * NT8;   // k-steps of the padded dimension
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tg = lane & 3, kmax = (n + 3) & ~3;
    const int strip = warp / WPS, t0 = (warp % WPS) * TH;
    const bool act = strip < NT8;
    const unsigned bm = smem_addr(B), gm = smem_addr(G), om = smem_addr(Out), ld8 = 8u * (unsigned)ld;
    double acc[TH][2];
    bool okb[TH], oks[TH];
#pragma unroll
    for (int t = 0; t < TH; ++t) {
        okb[t] = t0 + t < NT8 && 8 * (t0 + t) + g < ld; oks[t] = t0 + t < NT8 && 8 * (t0 + t) + 2 * tg + 1 < ld;
        acc[t][0] = acc[t][1] = 0.0;
    }
    if (act) {
        const double *arow = Th + (size_t)(strip * 8 + g) * ld + tg;            // A fragment: Th[row][k0 + tg]
        const unsigned bcol = bm + (unsigned)tg * ld8 + 8u * (g + 8 * t0);      // B fragment: B[k0 + tg][8 t + g]
        double av[KS];   // the warp's strip of Theta is 8 x n: all of its k-steps are fetched from L2 at once
#pragma unroll
        for (int u = 0; u < KS; ++u) av[u] = 4 * u < kmax ? __ldcg(arow + 4 * u) : 0.0;
#pragma unroll
        for (int u = 0; u < KS; ++u) {
            const int k0 = 4 * u;
            if (k0 < kmax) {
#pragma unroll
                for (int t = 0; t < TH; ++t) { const double b = okb[t] ? lds_f64(bcol + (unsigned)k0 * ld8 + 64u * t) : 0.0; QX_DMMA(acc[t], av[u], b); }
            }
        }
    }
    if (MODE == 2) __syncthreads();   // everybody has finished reading G
    if (act) {
        const unsigned grow = gm + (unsigned)(strip * 8 + g) * ld8 + 8u * (2 * tg + 8 * t0);
        const unsigned orow = om + (unsigned)(strip * 8 + g) * ld8 + 8u * (2 * tg + 8 * t0);
#pragma unroll
        for (int t = 0; t < TH; ++t)
            if (oks[t]) {
                const double g0 = lds_f64(grow + 64u * t), g1 = lds_f64(grow + 64u * t + 8u);
                if (MODE == 0) sts_v2f64(orow + 64u * t, fma(0.5, acc[t][0], g0), fma(0.5, acc[t][1], g1));
                else sts_v2f64(orow + 64u * t, g0 + acc[t][0], g1 + acc[t][1]);
            }
    }
    __syncthreads();
}

// G: rows after a sweep with couplings < QX_POLISH_GATE (scales folded in); X: the other shared-memory matrix (contents preserved);
// gs: 2 * rows8 * ld doubles of the CTA's global slab; jw: the sweep's (now idle) state vector.  Returns true if the rows are
// orthogonal to tol afterwards (the job of the last sweep is done), false if the couplings are not yet below tol (G unchanged).
template <int NT8>
static __device__ __noinline__ int jacobi_polish(int n, double *G, double *X, int ld, double *gs, double *jw, float tol) {
    QX_ASSUME_SHARED(G); QX_ASSUME_SHARED(X); QX_ASSUME_SHARED(jw);
    const int nfull = 8 * NT8 * ld;
    double *Mg = gs, *park = gs + nfull;
    int *cnt = reinterpret_cast<int *>(jw), *list = cnt + 2;
    double *dg = jw + 2 + QX_POLISH_CAP / 2;                   // diagonal of M (jw: 3 n + 8 doubles, n >= 16)
    if (threadIdx.x == 0) *cnt = 0;   // (tc_gram_theta has a barrier between its product and the tangents)
    QX_PSUB_BEGIN();
    const int flags = tc_gram_theta<NT8>(n, G, Mg, ld, dg, (double)tol * (double)tol, cnt, list);   // Mg <- Theta
    QX_PSUB(8);
    const bool above = flags & 1, far = flags & 2, second = flags & 4;
    if (__syncthreads_or(far ? 1 : 0)) { QX_PSUB(9); QX_PCOUNT(14); return 0; }   // too far for a simultaneous correction: sweep
    const int done = __syncthreads_or(above ? 1 : 0) ? 2 : 1;                      // 2: applied, to be checked again
    QX_PSUB(9);
    const int nbig = *cnt;
    if (nbig > QX_POLISH_CAP) return 0;
    if (!__syncthreads_or(second ? 1 : 0) && nbig == 0) {
        tc_polish_apply<NT8, 2>(n, Mg, G, G, G, ld);   // all angles below QX_POLISH_T2: G += Theta G is exact to 1e-12
        QX_PSUB(10); QX_PCOUNT(15);
        return done;
    }
    for (int t = threadIdx.x; t < nfull; t += QX_NT) park[t] = X[t];
    __syncthreads();
    QX_PSUB(11);
    tc_polish_apply<NT8, 0>(n, Mg, G, G, X, ld);   // X = G + Theta G / 2
    tc_polish_apply<NT8, 1>(n, Mg, X, G, G, ld);   // G += Theta X
    QX_PSUB(12);
    if (nbig > 0 && threadIdx.x < 32) {            // exact rotations of the near-degenerate pairs, one after the other
        const int lane = threadIdx.x;
        int last = -1;
        for (int e = 0; e < nbig; ++e) {
            // ascending (i, j): the list was filled through an atomic counter, its order is not reproducible -- the rotations' is
            int best = 0x7fffffff;
            for (int q = lane; q < nbig; q += 32) { const int kq = list[q]; if (kq > last && kq < best) best = kq; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const int other = __shfl_xor_sync(0xffffffffu, best, o); best = other < best ? other : best; }
            last = best;
            const int i = best >> 16, j = best & 0xffff;
            double a = 0.0, b = 0.0, c = 0.0;
            for (int k = lane; k < n; k += 32) { const double x = G[(size_t)i * ld + k], y = G[(size_t)j * ld + k]; a = fma(x, x, a); b = fma(y, y, b); c = fma(x, y, c); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
            if (c * c > 1e-30 * a * b) {
                const double d = b - a;
                double tt = 2.0 * c / (fabs(d) + sqrt(fma(d, d, 4.0 * c * c)));
                if (d < 0.0) tt = -tt;
                const double cs = 1.0 / sqrt(fma(tt, tt, 1.0)), sn = tt * cs;
                for (int k = lane; k < n; k += 32) {
                    const double x = G[(size_t)i * ld + k], y = G[(size_t)j * ld + k];
                    G[(size_t)i * ld + k] = cs * x - sn * y;
                    G[(size_t)j * ld + k] = sn * x + cs * y;
                }
            }
            __syncwarp();
        }
    }
    for (int t = threadIdx.x; t < nfull; t += QX_NT) X[t] = __ldcg(park + t);
    __syncthreads();
    QX_PSUB(13);
    return done;
}

// The same for bases that do not run the strip GEMMs (medium: G in shared memory; large: G in the CTA's global slab): the three
// products go through gemm_tc (staged through `stage` when given), M / Theta and Y = Theta G live in gs (2 n^2 doubles, row stride n).
template <bool SH>
static __device__ __noinline__ int jacobi_polish_gen(int n, double *G, int ld, double *gs, double *jw, float tol, double *stage, int stage_doubles) {
    if (SH) QX_ASSUME_SHARED(G);
    QX_ASSUME_SHARED(jw);
    double *Mg = gs, *Yg = gs + (size_t)n * n;
    int *cnt = reinterpret_cast<int *>(jw), *list = cnt + 2;
    double *dg = jw + 2 + QX_POLISH_CAP / 2;
    if (threadIdx.x == 0) *cnt = 0;
    {
        const double *Gc = G;
        gemm_tc(n, [=](int i, int k) { return Gc[(size_t)i * ld + k]; }, [=](int k, int j) { return Gc[(size_t)j * ld + k]; },
                [=](int i, int j, double v) { Mg[(size_t)i * n + j] = v; }, stage, stage_doubles);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += QX_NT) dg[i] = __ldcg(Mg + (size_t)i * n + i);
    __syncthreads();
    const double tol2 = (double)tol * (double)tol;
    bool above = false, second = false, far = false;
    for (int t0 = threadIdx.x; t0 < n * n; t0 += 4 * QX_NT) {   // four couplings in flight per thread
        double mv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int t = t0 + u * QX_NT; mv[u] = t < n * n ? __ldcg(Mg + t) : 0.0; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * QX_NT, i = t / n, j = t - i * n;
            if (t < n * n && i != j) {
                const bool up = i < j;
                const double gij = mv[u], a = dg[up ? i : j], b = dg[up ? j : i];
                const double g2d = gij * gij, nn = a * b;
                above |= g2d > tol2 * nn;
                far |= g2d > (QX_POLISH_APPLY * QX_POLISH_APPLY) * nn;
                const float gf = (float)gij, df = (float)(b - a);
                const float g2 = gf + gf;
                const float hh = fmaf(df, df, g2 * g2);
                const float den = fabsf(df) + hh * rsqrt_approx(hh);
                float tf = g2 * rcp_approx(den);
                tf = __int_as_float(__float_as_int(tf) ^ (__float_as_int(df) & 0x80000000));
                if (!(g2d > 1e-30 * nn)) tf = 0.0f;
                if (fabsf(tf) > (float)QX_POLISH_TBIG) {
                    if (up) {
                        const int idx = atomicAdd(cnt, 1);
                        if (idx < QX_POLISH_CAP) list[idx] = (i << 16) | j;
                    }
                    tf = 0.0f;
                }
                second |= fabsf(tf) > (float)QX_POLISH_T2;
                Mg[t] = up ? -(double)tf : (double)tf;
            }
        }
    }
    for (int i = threadIdx.x; i < n; i += QX_NT) Mg[(size_t)i * n + i] = 0.0;
    if (__syncthreads_or(far ? 1 : 0)) return 0;                   // too far for a simultaneous correction: sweep
    const int done = __syncthreads_or(above ? 1 : 0) ? 2 : 1;      // 2: applied, to be checked again
    const int nbig = *cnt;
    if (nbig > QX_POLISH_CAP) return 0;
    const bool need2 = __syncthreads_or(second ? 1 : 0) != 0;
    {
        const double *Gc = G;
        gemm_tc(n, [=](int i, int k) { return __ldcg(Mg + (size_t)i * n + k); }, [=](int k, int j) { return Gc[(size_t)k * ld + j]; },
                [=](int i, int j, double v) { Yg[(size_t)i * n + j] = v; }, stage, stage_doubles);
    }
    __syncthreads();
    if (!need2) {
        for (int t = threadIdx.x; t < n * n; t += QX_NT) { const int i = t / n, j = t - i * n; G[(size_t)i * ld + j] += __ldcg(Yg + t); }
    } else {   // G += Y + Theta Y / 2 (the product reads Theta and Y only)
        gemm_tc(n, [=](int i, int k) { return __ldcg(Mg + (size_t)i * n + k); }, [=](int k, int j) { return __ldcg(Yg + (size_t)k * n + j); },
                [=](int i, int j, double v) { G[(size_t)i * ld + j] += __ldcg(Yg + (size_t)i * n + j) + 0.5 * v; }, stage, stage_doubles);
    }
    __syncthreads();
    if (nbig > 0 && threadIdx.x < 32) {   // exact rotations of the near-degenerate pairs, one after the other
        const int lane = threadIdx.x;
        int last = -1;
        for (int e = 0; e < nbig; ++e) {
            // ascending (i, j): the list was filled through an atomic counter, its order is not reproducible -- the rotations' is
            int best = 0x7fffffff;
            for (int q = lane; q < nbig; q += 32) { const int kq = list[q]; if (kq > last && kq < best) best = kq; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const int other = __shfl_xor_sync(0xffffffffu, best, o); best = other < best ? other : best; }
            last = best;
            const int i = best >> 16, j = best & 0xffff;
            double a = 0.0, b = 0.0, c = 0.0;
            for (int k = lane; k < n; k += 32) { const double x = G[(size_t)i * ld + k], y = G[(size_t)j * ld + k]; a = fma(x, x, a); b = fma(y, y, b); c = fma(x, y, c); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
            if (c * c > 1e-30 * a * b) {
                const double d = b - a;
                double tt = 2.0 * c / (fabs(d) + sqrt(fma(d, d, 4.0 * c * c)));
                if (d < 0.0) tt = -tt;
                const double cs = 1.0 / sqrt(fma(tt, tt, 1.0)), sn = tt * cs;
                for (int k = lane; k < n; k += 32) {
                    const double x = G[(size_t)i * ld + k], y = G[(size_t)j * ld + k];
                    G[(size_t)i * ld + k] = cs * x - sn * y;
                    G[(size_t)j * ld + k] = sn * x + cs * y;
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    return done;
}

// (2) the sweeps
template <bool SH>
__device__ __forceinline__ int jacobi_sweeps(int n, double *G, int ld, double *red, double *jw, double *jblk = nullptr, int jblock = 0,
                                             double *Xc = nullptr, double *gpol = nullptr) {
    const float tol = 1e-7f;  // pre-rotation ratio of the last sweep; its rotations leave O(tol^2) couplings
    const int npair = (n + 1) >> 1;
    int sweeps;
    if (npair * 8 <= QX_NT && (ld & 1) == 0 && n <= 80) {
        const float tolr = QX_JACOBI_TOL;
        // Xc / gpol given (strip-GEMM sizes only): the verification sweep is replaced by jacobi_polish
        const int npad = tc_padded_dim(n);
        const float gate = SH && Xc && gpol && n >= 16 && npad != 0 && npad / 8 <= QX_NT / 32 ? QX_POLISH_GATE : 0.0f;   // (n >= 16: the pair list lives in jw)
        sweeps = 0;
        for (;;) {
            int r;
            switch ((n + 15) >> 4) {
                case 1: r = QX_JROWS<1>(n, G, ld, tolr, jw, gate); break;
                case 2: r = QX_JROWS<2>(n, G, ld, tolr, jw, gate); break;
                case 3: r = QX_JROWS<3>(n, G, ld, tolr, jw, gate); break;
                case 4: r = QX_JROWS<4>(n, G, ld, tolr, jw, gate); break;
                default: r = QX_JROWS<5>(n, G, ld, tolr, jw, gate); break;   // 8 lanes per pair and QX_NT threads: n <= 72
            }
            if (r >= 0) { sweeps += r; break; }
            sweeps -= r;
            int st = 2;   // polish iterations: 0 = back to the sweeps, 1 = converged, 2 = corrected, check again
            for (int it = 0; it < 4 && st == 2; ++it)
                st = npad == 32 ? jacobi_polish<4>(n, G, Xc, ld, gpol, jw, tolr)
                                : (npad == 64 ? jacobi_polish<8>(n, G, Xc, ld, gpol, jw, tolr) : jacobi_polish<9>(n, G, Xc, ld, gpol, jw, tolr));
            if (st == 1) break;
        }
    } else if ((ld & 1) == 0 && n <= 112 && SH) {   // matrices in shared memory, more pairs than 8-lane groups: several passes per round
        // (jacobi_polish_gen measured here: C14H30 -2 %, C17H36 -51 % -- the generic products with operands behind lambdas cost more
        // than the sweep they save at these sizes; the gate stays closed)
        const float gate = 0.0f;
        sweeps = 0;
        for (;;) {
            int r;
            switch ((n + 15) >> 4) {   // R = ceil(n / 16): every chunk but the last is in range for all lanes
                case 5: r = jacobi_rows_lp8m<5>(n, G, ld, QX_JACOBI_TOL, jw, gate); break;
                case 6: r = jacobi_rows_lp8m<6>(n, G, ld, QX_JACOBI_TOL, jw, gate); break;
                default: r = jacobi_rows_lp8m<7>(n, G, ld, QX_JACOBI_TOL, jw, gate); break;
            }
            if (r >= 0) { sweeps += r; break; }
            sweeps -= r;
            if (jacobi_polish_gen<SH>(n, G, ld, gpol, jw, QX_JACOBI_TOL, nullptr, 0) == 1) break;
        }
    } else if (!SH && jblock >= 8 && (ld & 1) == 0 && (reinterpret_cast<size_t>(G) & 15) == 0 && n <= 316) {   // global slab, blocked through shared memory
        if (n <= 96) return jacobi_rows_generic(n, G, ld, red, tol);
        const float gate = gpol ? QX_POLISH_GATE : 0.0f;
        sweeps = 0;
        for (;;) {
            int r;
            switch ((n + 2 * QX_JB_LANES - 1) / (2 * QX_JB_LANES)) {   // R = double2 chunks per lane and row
                case 2: r = jacobi_rows_blocked<2>(n, G, ld, QX_JACOBI_TOL, jw, jblk, jblock, gate); break;
                case 3: r = jacobi_rows_blocked<3>(n, G, ld, QX_JACOBI_TOL, jw, jblk, jblock, gate); break;
                case 4: r = jacobi_rows_blocked<4>(n, G, ld, QX_JACOBI_TOL, jw, jblk, jblock, gate); break;
                case 5: r = jacobi_rows_blocked<5>(n, G, ld, QX_JACOBI_TOL, jw, jblk, jblock, gate); break;
                case 6: r = jacobi_rows_blocked<6>(n, G, ld, QX_JACOBI_TOL, jw, jblk, jblock, gate); break;
                case 7: r = jacobi_rows_blocked<7>(n, G, ld, QX_JACOBI_TOL, jw, jblk, jblock, gate); break;
                case 8: r = jacobi_rows_blocked<8>(n, G, ld, QX_JACOBI_TOL, jw, jblk, jblock, gate); break;
                case 9: r = jacobi_rows_blocked<9>(n, G, ld, QX_JACOBI_TOL, jw, jblk, jblock, gate); break;
                default: r = jacobi_rows_blocked<10>(n, G, ld, QX_JACOBI_TOL, jw, jblk, jblock, gate); break;   // n <= 316 (ld <= 320)
            }
            if (r >= 0) { sweeps += r; break; }
            sweeps -= r;
            int st = 2;
            for (int it = 0; it < 4 && st == 2; ++it) st = jacobi_polish_gen<SH>(n, G, ld, gpol, jw, QX_JACOBI_TOL, jblk, 2 * jblock * ld);
            if (st == 1) break;
        }
    } else if (!SH && (ld & 1) == 0 && (reinterpret_cast<size_t>(G) & 15) == 0 && n <= 320) {   // global slab
        if (n <= 96) return jacobi_rows_generic(n, G, ld, red, tol);   // (never in practice: the slab mode starts above ~110 AOs)
        switch ((n + 31) >> 5) {   // R = ceil(n / 32) double2 per lane and row with 16 lanes per pair
            case 4: sweeps = jacobi_rows_glob<4, 16>(n, G, ld, QX_JACOBI_TOL, jw); break;
            case 5: sweeps = jacobi_rows_glob<5, 16>(n, G, ld, QX_JACOBI_TOL, jw); break;
            case 6: sweeps = jacobi_rows_glob<6, 16>(n, G, ld, QX_JACOBI_TOL, jw); break;
            case 7: sweeps = jacobi_rows_glob<7, 16>(n, G, ld, QX_JACOBI_TOL, jw); break;
            case 8: sweeps = jacobi_rows_glob<8, 16>(n, G, ld, QX_JACOBI_TOL, jw); break;
            default: sweeps = jacobi_rows_glob<5, 32>(n, G, ld, QX_JACOBI_TOL, jw); break;   // 256 < n <= 320: a full warp per pair
        }
    } else
        sweeps = jacobi_rows_generic(n, G, ld, red, tol);
    return sweeps;
}

// (3) eigenvalues from the row norms; normalise the rows.  On exit emo[k] = eigenvalue k and row k of G is the corresponding
// unit eigenvector (so G holds J^T).
template <bool SH>
static __device__ __noinline__ void jacobi_finish(int n, double *G, int ld, double *emo, const double *red) {
    QX_ASSUME_SHARED(emo); QX_ASSUME_SHARED(red);
    if (SH) QX_ASSUME_SHARED(G);
    const double sigma = red[60];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = warp; k < n; k += QX_NT / 32) {
        double acc = 0.0;
        for (int i = lane; i < n; i += 32) { const double x = G[(size_t)k * ld + i]; acc += x * x; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const double nrm = sqrt(acc), inv = 1.0 / nrm;
        for (int i = lane; i < n; i += 32) G[(size_t)k * ld + i] *= inv;
        if (lane == 0) emo[k] = nrm - sigma;
    }
    __syncthreads();
}

// Eigen-decomposition of the symmetric A' held in G (n x n, ld).  Returns the number of sweeps.
template <bool SH>
__device__ __forceinline__ int jacobi_eigh_rows(int n, double *G, int ld, double *emo, double *red, double *jw, double *jblk = nullptr, int jblock = 0,
                                                double *Xc = nullptr, double *gpol = nullptr) {
    jacobi_shift<SH>(n, G, ld, red);
    const int sweeps = jacobi_sweeps<SH>(n, G, ld, red, jw, jblk, jblock, Xc, gpol);
    jacobi_finish<SH>(n, G, ld, emo, red);
    return sweeps;
}

// ------------------------------------------------------------------------------------ Fermi smearing
// One warp runs the reference Newton iteration (tblite get_fermi_filling: start at the HOMO/LUMO
// midpoint, <= 200 cycles, threshold sqrt(eps)); lanes stride over the orbitals.  Returns the Fermi
// level the occupations of the final cycle were evaluated with (the reference updates e_fermi once
// more after filling).  e_lo / e_hi: the homo-th and (homo+1)-th smallest eigenvalue.
__device__ inline double fermi_level_warp(int n, int homo, double kt, const double *emo, double e_lo, double e_hi) {
    const double thr = 1.4901161193847656e-08;
    const int lane = threadIdx.x & 31;
    double ef = 0.5 * (e_lo + e_hi), ef_used = ef;
    const double occt = homo;
    for (int cyc = 0; cyc < 200; ++cyc) {
        double total = 0.0, dtotal = 0.0;
        for (int i = lane; i < n; i += 32) {
            double x = (emo[i] - ef) / kt;
            if (x < 50.0) {
                double ex = exp(x), den = 1.0 / (ex + 1.0);
                total += den;
                dtotal += ex * den * den / kt;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            total += __shfl_xor_sync(0xffffffffu, total, o);
            dtotal += __shfl_xor_sync(0xffffffffu, dtotal, o);
        }
        ef_used = ef;
        ef += (occt - total) / dtotal;
        if (fabs(occt - total) <= thr) break;
    }
    return ef_used;
}

}  // namespace qx
