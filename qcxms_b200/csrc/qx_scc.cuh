// SCC loop, Broyden mixer and analytic gradient of the GFN2-xTB ensemble kernel (one CTA per trajectory).
// Iteration protocol = tblite xtb_singlepoint as driven by the reference (src/tblite.f90:133-136):
// zero start, potential -> H1 -> eigenproblem -> Fermi filling per spin channel -> density ->
// Mulliken charges / atomic dipoles / quadrupoles -> energy -> modified Broyden (damping 0.4, full
// history); converged when |dE| < 1e-6 and rms(dq) < 2e-5 (accuracy = 1.0, src/tblite.f90:46).
#pragma once
#include "qx_device.cuh"

namespace qx {

// optional per-phase cycle accounting (profiling builds only: -DQX_PROFILE_PHASES)
#ifdef QX_PROFILE_PHASES
static __device__ unsigned long long g_phase_cycles[16];
// (g_sub_cycles: finer marks inside a phase, thread 0's clock, no extra barrier -- declared in qx_device.cuh, jacobi_polish uses it too)
#define QX_SUB_BEGIN() long long sub_t0_ = clock64()
#define QX_SUB(idx) do { if (threadIdx.x == 0) { long long t_ = clock64(); atomicAdd(&g_sub_cycles[idx], (unsigned long long)(t_ - sub_t0_)); sub_t0_ = t_; } } while (0)
static __device__ unsigned long long g_sweep_hist[64];  // [iteration index (<32)] -> sweeps, [32+..] -> count
#define QX_PH_BEGIN() long long ph_t0_ = clock64()
#define QX_PH(idx)                                                                  \
    do {                                                                            \
        __syncthreads();                                                            \
        if (threadIdx.x == 0) {                                                     \
            long long ph_t1_ = clock64();                                           \
            atomicAdd(&g_phase_cycles[idx], (unsigned long long)(ph_t1_ - ph_t0_)); \
            ph_t0_ = ph_t1_;                                                        \
        } else                                                                      \
            ph_t0_ = 0;                                                             \
    } while (0)
#else
#define QX_PH_BEGIN() do {} while (0)
#define QX_PH(idx) do {} while (0)
#define QX_SUB_BEGIN() do {} while (0)
#define QX_SUB(idx) do {} while (0)
#endif

}  // namespace qx
#include "qx_oa.cuh"   // (after the profiling macros it uses)
namespace qx {

struct EgradOut {
    double energy;
    double e_rep, e_atm, e_el, e_es, e_aes, e_d4, e_ts;
    int niter, stat, sweeps;
};

// damped multipole interaction kernels for the pair (i, j); v = R_j - R_i
__device__ inline void aes_pair(const Sm &s, int i, int j, double v[3], double &g3f3, double &g3f5, double &g5f5) {
    v[0] = s.xyz[3 * j] - s.xyz[3 * i]; v[1] = s.xyz[3 * j + 1] - s.xyz[3 * i + 1]; v[2] = s.xyz[3 * j + 2] - s.xyz[3 * i + 2];
    double r2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], g1 = rsqrt(r2);
    g1 = g1 * (1.5 - 0.5 * r2 * g1 * g1);  // one Newton step on top of rsqrt (full double accuracy)
    double g3 = g1 * g1 * g1, g5 = g3 * g1 * g1;
    double rr = 0.5 * (s.mrad[i] + s.mrad[j]) * g1, rr3 = rr * rr * rr;
    double f3 = 1.0 / (1.0 + 6.0 * rr3), f5 = 1.0 / (1.0 + 6.0 * rr3 * rr);
    g3f3 = g3 * f3; g3f5 = g3 * f5; g5f5 = g5 * f5;
}

__device__ inline double quad_contract(const double *q, const double v[3]) {
    return q[0] * v[0] * v[0] + 2.0 * q[1] * v[0] * v[1] + q[2] * v[1] * v[1] + 2.0 * q[3] * v[0] * v[2] + 2.0 * q[4] * v[1] * v[2] + q[5] * v[2] * v[2];
}

// D4 charge-dependent contraction in two steps (7x fewer multiply-adds than the direct triple loop):
//   u(j, ti, ri) = sum_rj c6ref(ti, type_j; ri, rj) gw(j, rj)          [s.d4u, nat * ntype * 7]
//   vvec(i, ri)  = - sum_j edisp_ij u(j, type_i, ri)
__device__ __forceinline__ void d4_u_table(const DevModel &m, Sm &s) {
    const int nat = m.nat, ntype = m.ntype;
    for (int t = threadIdx.x; t < nat * ntype * QX_MAXREF; t += QX_NT) {
        const int j = t / (ntype * QX_MAXREF), r = t - j * ntype * QX_MAXREF, ti = r / QX_MAXREF, ri = r - ti * QX_MAXREF;
        const double *ref = m.c6ref + (((size_t)ti * ntype + m.type[j]) * QX_MAXREF + ri) * QX_MAXREF;
        double acc = 0.0;
        const int nj = m.at_nref[j];
        for (int rj = 0; rj < nj; ++rj) acc += ref[rj] * s.gw[j * QX_MAXREF + rj];
        s.d4u[t] = acc;
    }
}
__device__ __forceinline__ double d4_vvec(const DevModel &m, const Sm &s, const double *edisp, int i, int ri) {
    const int nat = m.nat, stride = m.ntype * QX_MAXREF, off = m.type[i] * QX_MAXREF + ri;
    double acc = 0.0;
    for (int j = 0; j < nat; ++j) acc += edisp[i * nat + j] * s.d4u[j * stride + off];
    return -acc;
}

// v_a = sum_b gamma_ab q_b for shell a (group-wide result)
__device__ __forceinline__ double gamma_row_oct(const Sm &s, const double *gamma, int nsh, int a) {
    double v = 0.0;
    for (int b = threadIdx.x & 7; b < nsh; b += 8) v += gamma[a * nsh + b] * s.qsh[b];
    return oct_sum(v);
}

// potentials from the (input) populations in s.qsh/qat/dpat/qpat -> s.vsh, vat, vdp, vqp, vao
static __device__ __noinline__ void phase_potential(const DevModel &m, Sm &s, const double *gamma, const double *edisp, double *t7) {
    const int nat = m.nat, nsh = m.nsh, nao = m.nao;
    const int oct = threadIdx.x >> 3, l8 = threadIdx.x & 7, wbase = (threadIdx.x >> 5) << 2;
    if (m.method != 1) d4_weights_all(m, s, true, s.gw, nullptr, s.gwd);
    for (int base = 0; base < nsh; base += QX_NT / 8) {
        if (base + wbase >= nsh) continue;
        const bool active = base + oct < nsh;
        const int a = active ? base + oct : 0;
        const double v = gamma_row_oct(s, gamma, nsh, a);
        if (active && l8 == 0) s.vsh[a] = v + s.qsh[a] * s.qsh[a] * m.sh_gam3[a];
    }
    if (m.method == 1) {   // GFN1: atom-resolved third order; no multipole electrostatics, no self-consistent dispersion
        for (int i = threadIdx.x; i < nat; i += QX_NT) s.vat[i] = s.qat[i] * s.qat[i] * m.at_gam3[i];
        for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) s.vdp[i] = 0.0;
        for (int i = threadIdx.x; i < 6 * nat; i += QX_NT) s.vqp[i] = 0.0;
        __syncthreads();
        for (int mu = threadIdx.x; mu < nao; mu += QX_NT) s.vao[mu] = s.vsh[m.ao_sh[mu]] + s.vat[m.ao_at[mu]];
        __syncthreads();
        return;
    }
    // anisotropic electrostatics: one group per atom i, lanes over the partners j
    for (int base = 0; base < nat; base += QX_NT / 8) {
        if (base + wbase >= nat) continue;
        const bool active = base + oct < nat;
        const int i = active ? base + oct : 0;
        double acc[10];
#pragma unroll
        for (int c = 0; c < 10; ++c) acc[c] = 0.0;
        for (int j = l8; j < nat; j += 8) {
            if (j == i) continue;
            double v[3], g3f3, g3f5, g5f5;
            aes_pair(s, i, j, v, g3f3, g3f5, g5f5);
            const double qj = s.qat[j];
            const double *mj = s.dpat + 3 * j;
            const double mv = mj[0] * v[0] + mj[1] * v[1] + mj[2] * v[2];
#pragma unroll
            for (int k = 0; k < 3; ++k) acc[k] += v[k] * g3f3 * qj + g3f5 * mj[k] - 3.0 * g5f5 * v[k] * mv;
            const double tq = g5f5 * qj;
            acc[3] += tq * v[0] * v[0]; acc[4] += 2.0 * tq * v[0] * v[1]; acc[5] += tq * v[1] * v[1];
            acc[6] += 2.0 * tq * v[0] * v[2]; acc[7] += 2.0 * tq * v[1] * v[2]; acc[8] += tq * v[2] * v[2];
            acc[9] += -g3f3 * mv + g5f5 * quad_contract(s.qpat + 6 * j, v);
        }
#pragma unroll
        for (int c = 0; c < 10; ++c) acc[c] = oct_sum(acc[c]);
        if (active) {
#pragma unroll
            for (int c = 0; c < 10; ++c) {
                if (l8 == (c & 7)) {
                    if (c < 3) s.vdp[3 * i + c] = acc[c] + 2.0 * m.at_dk[i] * s.dpat[3 * i + c];
                    else if (c < 9) s.vqp[6 * i + c - 3] = acc[c] + 2.0 * m.at_qk[i] * s.qpat[6 * i + c - 3] * c_qscale[c - 3];
                    else s.vat[i] = acc[9];
                }
            }
        }
    }
    __syncthreads();
    d4_u_table(m, s);
    __syncthreads();
    for (int t = threadIdx.x; t < nat * QX_MAXREF; t += QX_NT) {
        int i = t / QX_MAXREF, ri = t - i * QX_MAXREF;
        t7[t] = ri < m.at_nref[i] ? d4_vvec(m, s, edisp, i, ri) * s.gwd[t] : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nat; i += QX_NT) {
        double v = 0.0;
        for (int r = 0; r < QX_MAXREF; ++r) v += t7[i * QX_MAXREF + r];
        s.vat[i] += v;
    }
    __syncthreads();
    for (int mu = threadIdx.x; mu < nao; mu += QX_NT) s.vao[mu] = s.vsh[m.ao_sh[mu]] + s.vat[m.ao_at[mu]];
    __syncthreads();
}

// energies of the charge-dependent terms at the (output) populations
static __device__ __noinline__ void phase_scc_energy(const DevModel &m, Sm &s, const double *gamma, const double *edisp, double &e_es, double &e_aes, double &e_d4) {
    const int nat = m.nat, nsh = m.nsh;
    const int oct = threadIdx.x >> 3, l8 = threadIdx.x & 7, wbase = (threadIdx.x >> 5) << 2;
    if (m.method != 1) d4_weights_all(m, s, true, s.gw, nullptr, nullptr);
    double es = 0.0, ea = 0.0, ed = 0.0;
    for (int base = 0; base < nsh; base += QX_NT / 8) {
        if (base + wbase >= nsh) continue;
        const bool active = base + oct < nsh;
        const int a = active ? base + oct : 0;
        const double v = gamma_row_oct(s, gamma, nsh, a);
        if (active && l8 == 0) es += 0.5 * v * s.qsh[a] + s.qsh[a] * s.qsh[a] * s.qsh[a] * m.sh_gam3[a] / 3.0;
    }
    if (m.method == 1) {   // GFN1: + atomic third order, nothing else depends on the charges
        for (int i = threadIdx.x; i < nat; i += QX_NT) es += s.qat[i] * s.qat[i] * s.qat[i] * m.at_gam3[i] / 3.0;
        e_es = block_sum(es, s.red);
        e_aes = 0.0;
        e_d4 = 0.0;
        return;
    }
    for (int base = 0; base < nat; base += QX_NT / 8) {
        if (base + wbase >= nat) continue;
        const bool active = base + oct < nat;
        const int i = active ? base + oct : 0;
        const double *mi = s.dpat + 3 * i;
        double e = 0.0;
        for (int j = l8; j < nat; j += 8) {
            if (j == i) continue;
            double v[3], g3f3, g3f5, g5f5;
            aes_pair(s, i, j, v, g3f3, g3f5, g5f5);
            const double qj = s.qat[j];
            const double *mj = s.dpat + 3 * j;
            const double mv = mj[0] * v[0] + mj[1] * v[1] + mj[2] * v[2];
            double vd0 = v[0] * g3f3 * qj + 0.5 * (g3f5 * mj[0] - 3.0 * g5f5 * v[0] * mv);
            double vd1 = v[1] * g3f3 * qj + 0.5 * (g3f5 * mj[1] - 3.0 * g5f5 * v[1] * mv);
            double vd2 = v[2] * g3f3 * qj + 0.5 * (g3f5 * mj[2] - 3.0 * g5f5 * v[2] * mv);
            e += mi[0] * vd0 + mi[1] * vd1 + mi[2] * vd2 + g5f5 * qj * quad_contract(s.qpat + 6 * i, v);
        }
        e = oct_sum(e);
        if (active && l8 == 0) {
            e += m.at_dk[i] * (mi[0] * mi[0] + mi[1] * mi[1] + mi[2] * mi[2]);
            for (int k = 0; k < 6; ++k) e += m.at_qk[i] * s.qpat[6 * i + k] * s.qpat[6 * i + k] * c_qscale[k];
            ea += e;
        }
    }
    __syncthreads();  // gw complete
    d4_u_table(m, s);
    __syncthreads();
    for (int t = threadIdx.x; t < nat * QX_MAXREF; t += QX_NT) {
        int i = t / QX_MAXREF, ri = t - i * QX_MAXREF;
        if (ri < m.at_nref[i]) ed += 0.5 * d4_vvec(m, s, edisp, i, ri) * s.gw[t];
    }
    e_es = block_sum(es, s.red);
    e_aes = block_sum(ea, s.red);
    e_d4 = block_sum(ed, s.red);
}

// H1 = H0 - 1/2 S (v_a + v_b) - 1/2 (D.vdp + D^T.vdp) - 1/2 (Q.vqp + ...) into s.A (symmetric)
// The integral slab (S, H0, 3 dipole, 6 quadrupole matrices: 11 nao^2 doubles) lives in the CTA's global scratch and is
// streamed from L2 twice per SCC cycle (here and in phase_mulliken).  Both phases are latency-bound, so what matters is the
// number of bytes in flight per round trip: 128-bit loads, all 11 matrices of an element pair issued back to back.
__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
template <bool SH>
static __device__ __noinline__ void phase_build_h1(const DevModel &m, Sm &s, const double *S, const double *H0, const double *Dt, const double *Qt) {
    const int nao = m.nao, ld = m.ld;
    double *const A = s.A;
    if (SH) QX_ASSUME_SHARED(A);
    const size_t n2 = (size_t)nao * nao;
    if ((nao & 1) == 0) {   // element pairs (t, t+1) of one row
        for (int u = threadIdx.x; u < nao * nao / 2; u += QX_NT) {
            const int t = 2 * u, b = t / nao, a = t - b * nao, ib = m.ao_at[b];
            const double2 h0 = ld2(H0 + t), sv = ld2(S + t);
            double2 d[3], q[6];
#pragma unroll
            for (int c = 0; c < 3; ++c) d[c] = ld2(Dt + c * n2 + t);
#pragma unroll
            for (int c = 0; c < 6; ++c) q[c] = ld2(Qt + c * n2 + t);
            const double vb = s.vao[b];
            const double *vd = s.vdp + 3 * ib, *vq = s.vqp + 6 * ib;
            double ax = 0.0, ay = 0.0;
#pragma unroll
            for (int c = 0; c < 3; ++c) { ax += d[c].x * vd[c]; ay += d[c].y * vd[c]; }
#pragma unroll
            for (int c = 0; c < 6; ++c) { ax += q[c].x * vq[c]; ay += q[c].y * vq[c]; }
            *reinterpret_cast<double2 *>(A + (size_t)b * ld + a) =
                make_double2(0.5 * h0.x - 0.5 * sv.x * vb - 0.5 * ax, 0.5 * h0.y - 0.5 * sv.y * vb - 0.5 * ay);
        }
    } else {
        for (int t = threadIdx.x; t < nao * nao; t += QX_NT) {
            int b = t / nao, a = t - b * nao, ib = m.ao_at[b];
            double g = 0.5 * H0[t] - 0.5 * S[t] * s.vao[b];
            const double *vd = s.vdp + 3 * ib, *vq = s.vqp + 6 * ib;
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < 3; ++c) acc += Dt[c * n2 + t] * vd[c];
#pragma unroll
            for (int c = 0; c < 6; ++c) acc += Qt[c * n2 + t] * vq[c];
            A[(size_t)b * ld + a] = g - 0.5 * acc;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nao * nao; t += QX_NT) {
        int b = t / nao, a = t - b * nao;
        if (a > b) continue;
        double v = A[(size_t)b * ld + a] + A[(size_t)a * ld + b];
        if (a == b) v = 2.0 * A[(size_t)b * ld + b];
        A[(size_t)b * ld + a] = v;
        A[(size_t)a * ld + b] = v;
    }
    __syncthreads();
}

// Mulliken populations from P (in s.A, symmetric): qsh, qat, dpat, qpat and tr(P H0); pop: 11*nao doubles of scratch
template <bool SH>
static __device__ __noinline__ double phase_mulliken(const DevModel &m, Sm &s, const double *S, const double *H0, const double *Dt, const double *Qt, double *pop) {
    const int nao = m.nao, ld = m.ld, nat = m.nat, nsh = m.nsh;
    const double *const A = s.A;
    if (SH) QX_ASSUME_SHARED(A);
    QX_ASSUME_SHARED(pop);
    const size_t n2 = (size_t)nao * nao;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // One warp per row b.  Even nao: lanes take element pairs; the pairs beyond the last full group of 32 are either one
    // more (partly idle) pass or -- when there are only a few of them, e.g. 1 of 33 for caffeine -- left to a second loop
    // over (row, matrix) items, so that a nearly empty pass does not cost a full memory round trip per row.
    const int npr = (nao & 1) ? 0 : nao >> 1, nfull = npr >> 5, rem = npr - (nfull << 5);
    const bool tail_items = rem > 0 && rem < 8 && nfull > 0;
    const int npass = (nao & 1) ? 0 : nfull + ((rem > 0 && !tail_items) ? 1 : 0);
    QX_SUB_BEGIN();
    for (int b = warp; b < nao; b += QX_NT / 32) {
        double acc[11];
#pragma unroll
        for (int c = 0; c < 11; ++c) acc[c] = 0.0;
        const size_t t0 = (size_t)b * nao;
        for (int j = 0; j < npass; ++j) {
            const int ap = lane + 32 * j;
            if (ap < npr) {
                const size_t t = t0 + 2 * ap;
                const double2 sv = ld2(S + t), h0 = ld2(H0 + t);
                double2 d[3], q[6];
#pragma unroll
                for (int c = 0; c < 3; ++c) d[c] = ld2(Dt + c * n2 + t);
#pragma unroll
                for (int c = 0; c < 6; ++c) q[c] = ld2(Qt + c * n2 + t);
                const double2 p = *reinterpret_cast<const double2 *>(A + (size_t)b * ld + 2 * ap);
                acc[0] += p.x * sv.x + p.y * sv.y;
                acc[10] += p.x * h0.x + p.y * h0.y;
#pragma unroll
                for (int c = 0; c < 3; ++c) acc[1 + c] += p.x * d[c].x + p.y * d[c].y;
#pragma unroll
                for (int c = 0; c < 6; ++c) acc[4 + c] += p.x * q[c].x + p.y * q[c].y;
            }
        }
        if (nao & 1) {
            for (int a = lane; a < nao; a += 32) {
                const double p = A[(size_t)b * ld + a];
                const size_t t = t0 + a;
                acc[0] += p * S[t];
                acc[10] += p * H0[t];
#pragma unroll
                for (int c = 0; c < 3; ++c) acc[1 + c] += p * Dt[c * n2 + t];
#pragma unroll
                for (int c = 0; c < 6; ++c) acc[4 + c] += p * Qt[c * n2 + t];
            }
        }
        // Transposed warp reduction of the 11 sums (padded to 16): every stage halves the number of values a lane carries
        // -- the half it gives away goes to the partner lane -- so 8 + 4 + 2 + 1 + 1 = 16 double shuffles instead of 55
        // (the shuffle unit is shared with the co-resident CTA's Jacobi and was the bottleneck of this loop).  Fixed order.
        double a8[8], a4[4], a2[2];
        {
            const bool up = lane & 16;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double lo = acc[i], hi = i + 8 < 11 ? acc[i + 8] : 0.0;
                const double got = __shfl_xor_sync(0xffffffffu, up ? lo : hi, 16);
                a8[i] = (up ? hi : lo) + got;
            }
        }
        {
            const bool up = lane & 8;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double got = __shfl_xor_sync(0xffffffffu, up ? a8[i] : a8[i + 4], 8);
                a4[i] = (up ? a8[i + 4] : a8[i]) + got;
            }
        }
        {
            const bool up = lane & 4;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double got = __shfl_xor_sync(0xffffffffu, up ? a4[i] : a4[i + 2], 4);
                a2[i] = (up ? a4[i + 2] : a4[i]) + got;
            }
        }
        double v;
        {
            const bool up = lane & 2;
            const double got = __shfl_xor_sync(0xffffffffu, up ? a2[0] : a2[1], 2);
            v = (up ? a2[1] : a2[0]) + got;
        }
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        {   // lane l now holds component 8 b4 + 4 b3 + 2 b2 + b1 (b_k: bit k of l), summed over the whole warp
            const int c = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            if ((lane & 1) == 0 && c < 11) pop[b * 11 + c] = v;
        }
    }
    QX_SUB(0);
    __syncthreads();
    QX_SUB(1);
    if (tail_items) {
        for (int it = threadIdx.x; it < nao * 11; it += QX_NT) {
            const int b = it / 11, c = it - 11 * b;
            const double *M = c == 0 ? S : (c == 10 ? H0 : (c < 4 ? Dt + (c - 1) * n2 : Qt + (c - 4) * n2));
            double v = 0.0;
            for (int ap = nfull << 5; ap < npr; ++ap) {
                const double2 x = ld2(M + (size_t)b * nao + 2 * ap);
                const double2 p = *reinterpret_cast<const double2 *>(A + (size_t)b * ld + 2 * ap);
                v += p.x * x.x + p.y * x.y;
            }
            pop[it] += v;
        }
        __syncthreads();
    }
    QX_SUB(2);
    for (int a = threadIdx.x; a < nsh; a += QX_NT) {
        double v = m.sh_refocc[a];
        int l = m.sh_l[a], ao0 = m.sh_ao0[a];
        for (int mu = ao0; mu < ao0 + 2 * l + 1; ++mu) v -= pop[mu * 11];
        s.qsh[a] = v;
    }
    double eel = 0.0;
    for (int i = threadIdx.x; i < nat; i += QX_NT) {
        double d[3] = {0, 0, 0}, q[6] = {0, 0, 0, 0, 0, 0};
        for (int mu = m.at_ao0[i]; mu < m.at_ao0[i] + m.at_nao[i]; ++mu) {
            for (int c = 0; c < 3; ++c) d[c] -= pop[mu * 11 + 1 + c];
            for (int c = 0; c < 6; ++c) q[c] -= pop[mu * 11 + 4 + c];
            eel += pop[mu * 11 + 10];
        }
        for (int c = 0; c < 3; ++c) s.dpat[3 * i + c] = d[c];
        for (int c = 0; c < 6; ++c) s.qpat[6 * i + c] = q[c];
    }
    eel = block_sum(eel, s.red);
    for (int i = threadIdx.x; i < nat; i += QX_NT) {
        double v = 0.0;
        for (int a = m.at_sh0[i]; a < m.at_sh0[i] + m.at_nsh[i]; ++a) v += s.qsh[a];
        s.qat[i] = v;
    }
    __syncthreads();
    QX_SUB(3);
    return eel;
}

// ------------------------------------------------------------------------------------ Broyden
struct Broyden {
    double *q_in, *qlast, *dq, *dqlast, *df, *u, *a, *omega, *beta, *cvec;
    int iter;
};

__device__ inline double warp_dot(const double *x, const double *y, int n) {
    double acc = 0.0;
    for (int i = threadIdx.x & 31; i < n; i += 32) acc += x[i] * y[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

// dense solve with partial pivoting on (beta[nb x nb], c[nb]) in global scratch; CTA-cooperative
static __device__ __noinline__ bool block_solve(int nb, double *beta, double *c, double *red) {
    QX_ASSUME_SHARED(red);
    __shared__ int s_piv;
    for (int k = 0; k < nb; ++k) {
        if (threadIdx.x == 0) {
            int piv = k;
            double best = fabs(beta[k * nb + k]);
            for (int i = k + 1; i < nb; ++i)
                if (fabs(beta[i * nb + k]) > best) { best = fabs(beta[i * nb + k]); piv = i; }
            s_piv = best == 0.0 ? -1 : piv;
        }
        __syncthreads();
        int piv = s_piv;
        if (piv < 0) return false;
        if (piv != k) {
            for (int j = threadIdx.x; j <= nb; j += QX_NT) {
                if (j < nb) { double t = beta[k * nb + j]; beta[k * nb + j] = beta[piv * nb + j]; beta[piv * nb + j] = t; }
                else { double t = c[k]; c[k] = c[piv]; c[piv] = t; }
            }
        }
        __syncthreads();
        const int rem = nb - k - 1;
        const double pivv = beta[k * nb + k];
        // column k multipliers are needed by every element of the row: compute on the fly
        for (int t = threadIdx.x; t < rem * (rem + 1); t += QX_NT) {
            int i = k + 1 + t / (rem + 1), jj = t % (rem + 1);
            double f = beta[i * nb + k] / pivv;
            if (jj < rem) beta[i * nb + k + 1 + jj] -= f * beta[k * nb + k + 1 + jj];
            else c[i] -= f * c[k];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        for (int i = nb - 1; i >= 0; --i) {
            double v = c[i];
            for (int j = i + 1; j < nb; ++j) v -= beta[i * nb + j] * c[j];
            c[i] = v / beta[i * nb + i];
        }
    }
    __syncthreads();
    return true;
}

// One mixer step: q_in <- next input.  dq = (output - input) of the cycle just finished must be set.
// bsol: QX_BSOL doubles of shared memory; systems up to QX_BSOL_N unknowns (the first 14 SCC cycles) are built and solved
// there -- the pivot search and back substitution are serial and were paying a global-memory round trip per element.
static __device__ __noinline__ bool broyden_next(Broyden &b, int n, double damp, double *red, double *bsol) {
    QX_ASSUME_SHARED(red); QX_ASSUME_SHARED(bsol);
    const int mem = QX_MAX_ITER;
    const double omega0 = 0.01, minw = 1.0, maxw = 100000.0, wfac = 0.01;
    b.iter += 1;
    const int iter = b.iter, itn = iter - 1;
    if (iter == 1) {
        for (int i = threadIdx.x; i < n; i += QX_NT) {
            b.dqlast[i] = b.dq[i];
            b.qlast[i] = b.q_in[i];
            b.q_in[i] += damp * b.dq[i];
        }
        __syncthreads();
        return true;
    }
    const int it1 = (itn - 1) % mem;
    const int nb = itn < mem ? itn : mem;
    const bool small = nb <= QX_BSOL_N;
    double *beta = small ? bsol : b.beta, *cvec = small ? bsol + QX_BSOL_N * QX_BSOL_N : b.cvec;
    QX_SUB_BEGIN();
    double nrm = 0.0, inv = 0.0;
    for (int i = threadIdx.x; i < n; i += QX_NT) {
        double d = b.dq[i], v = d - b.dqlast[i];
        nrm += d * d;
        inv += v * v;
    }
    nrm = sqrt(block_sum(nrm, red));
    inv = sqrt(block_sum(inv, red));
    double om = nrm > wfac / maxw ? wfac / nrm : maxw;
    if (om < minw) om = minw;
    if (inv < 2.220446049250313e-16) inv = 2.220446049250313e-16;
    inv = 1.0 / inv;
    for (int i = threadIdx.x; i < n; i += QX_NT) b.df[(size_t)it1 * n + i] = inv * (b.dq[i] - b.dqlast[i]);
    if (threadIdx.x == 0) b.omega[it1] = om;
    __syncthreads();
    QX_SUB(4);
    const int j0 = itn - mem + 1 > 1 ? itn - mem + 1 : 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = j0 + warp; j <= itn; j += QX_NT / 32) {
        int i = (j - 1) % mem;
        double aij = warp_dot(b.df + (size_t)i * n, b.df + (size_t)it1 * n, n);
        double ci = warp_dot(b.df + (size_t)i * n, b.dq, n);
        if (lane == 0) {
            b.a[i * mem + it1] = aij;
            b.a[it1 * mem + i] = aij;
            cvec[i] = b.omega[i] * ci;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nb * nb; t += QX_NT) {
        int k = t / nb, i = t - k * nb;
        double v = b.omega[k] * b.omega[i] * b.a[k * mem + i];
        if (k == i) v += omega0 * omega0;
        beta[k * nb + i] = v;
    }
    __syncthreads();
    QX_SUB(5);
    if (!block_solve(nb, beta, cvec, red)) return false;
    QX_SUB(6);
    for (int i = threadIdx.x; i < n; i += QX_NT) {
        b.u[(size_t)it1 * n + i] = damp * b.df[(size_t)it1 * n + i] + inv * (b.q_in[i] - b.qlast[i]);
        b.dqlast[i] = b.dq[i];
        b.qlast[i] = b.q_in[i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += QX_NT) {
        double v = b.q_in[i] + damp * b.dq[i];
        for (int j = j0; j <= itn; ++j) {
            int h = (j - 1) % mem;
            v -= b.omega[h] * cvec[h] * b.u[(size_t)h * n + i];
        }
        b.q_in[i] = v;
    }
    __syncthreads();
    QX_SUB(7);
    return true;
}

// ------------------------------------------------------------------------------------ gradient of the AO-pair terms
// s.A = P, s.C = W (energy weighted density); potentials in s.vao/vdp/vqp from the last SCC cycle.
template <bool SH>
static __device__ __noinline__ void phase_gradient_pairs(const DevModel &m, Sm &s, const int2 *tasks, int ntask, double *taskout) {
    const int ld = m.ld;
    const double *const A = s.A, *const Cw = s.C;
    if (SH) { QX_ASSUME_SHARED(A); QX_ASSUME_SHARED(Cw); }
    for (int t = threadIdx.x; t < ntask; t += QX_NT) {
        const int a = tasks[t].x, b = tasks[t].y;
        const int sa = m.ao_sh[a], sb = m.ao_sh[b], ja = m.ao_at[a], ib = m.ao_at[b];
        double *o = taskout + 5 * (size_t)t;
        if (ja == ib) { o[0] = o[1] = o[2] = o[3] = o[4] = 0.0; continue; }
        double vec[3] = {s.xyz[3 * ib] - s.xyz[3 * ja], s.xyz[3 * ib + 1] - s.xyz[3 * ja + 1], s.xyz[3 * ib + 2] - s.xyz[3 * ja + 2]};
        const double r2 = vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2];
        const double pij = A[(size_t)a * ld + b], wij = Cw[(size_t)a * ld + b];
        const double rr = sqrt(sqrt(r2) / (m.at_rad[ja] + m.at_rad[ib]));
        const double pla = 1.0 + m.sh_poly[sa] * rr, plb = 1.0 + m.sh_poly[sb] * rr;
        const double shp = pla * plb, dshp = (m.sh_poly[sa] * plb + m.sh_poly[sb] * pla) * rr * 0.5 / r2;
        const double hs = m.hscale[sa * m.nsh + sb], hav = 0.5 * (s.selfen[sa] + s.selfen[sb]);
        const double sval = 2.0 * pij * hav * hs * shp - 2.0 * wij - pij * (s.vao[a] + s.vao[b]);
        const double *vdI = s.vdp + 3 * ib, *vdJ = s.vdp + 3 * ja, *vqI = s.vqp + 6 * ib, *vqJ = s.vqp + 6 * ja;
        // weights of the raw second moments: 3/2 vq_c - 1/2 tr(vq) delta_c
        double wI[6], wJ[6];
        const double trI = 0.5 * (vqI[0] + vqI[2] + vqI[5]), trJ = 0.5 * (vqJ[0] + vqJ[2] + vqJ[5]);
        for (int c = 0; c < 6; ++c) { wI[c] = 1.5 * vqI[c]; wJ[c] = 1.5 * vqJ[c]; }
        wI[0] -= trI; wI[2] -= trI; wI[5] -= trI;
        wJ[0] -= trJ; wJ[2] -= trJ; wJ[5] -= trJ;
        double coef[10];
        coef[0] = sval - pij * (vec[0] * vdJ[0] + vec[1] * vdJ[1] + vec[2] * vdJ[2]);
        for (int k = 0; k < 3; ++k) coef[1 + k] = -pij * (vdI[k] + vdJ[k]);
        for (int c = 0; c < 6; ++c) {
            const int qa = c_qa[c], qb = c_qb[c];
            coef[4 + c] = -pij * (wI[c] + wJ[c]);
            coef[0] -= pij * wJ[c] * vec[qa] * vec[qb];
            coef[1 + qb] -= pij * wJ[c] * vec[qa];
            coef[1 + qa] -= pij * wJ[c] * vec[qb];
        }
        double raw[10], g[3];
        ao_pair_dispatch<true>(m, sa, m.ao_m[a], sb, m.ao_m[b], vec, r2, raw, coef, g);
        // explicit vec-dependence of the coefficients + distance dependence of H0
        for (int k = 0; k < 3; ++k) g[k] += -pij * vdJ[k] * raw[0] + 2.0 * pij * hav * hs * dshp * vec[k] * raw[0];
        for (int c = 0; c < 6; ++c) {
            const int qa = c_qa[c], qb = c_qb[c];
            g[qa] -= pij * wJ[c] * (raw[1 + qb] + vec[qb] * raw[0]);
            g[qb] -= pij * wJ[c] * (raw[1 + qa] + vec[qa] * raw[0]);
        }
        const double tcn = pij * hs * shp * raw[0];
        o[0] = g[0]; o[1] = g[1]; o[2] = g[2];
        o[3] = -m.sh_kcn[sb] * tcn;  // d/d cn of the ket atom
        o[4] = -m.sh_kcn[sa] * tcn;  // d/d cn of the bra atom
    }
    __syncthreads();
}

// Everything: s.xyz in, s.grad / s.qat out.  scratch: per-CTA global slab.
// qstart (may be null): OPT-IN warm start, not the reference protocol.  [2 ndim + 1]: converged populations (shell
// charges, atomic dipoles, quadrupoles) of the last and the last-but-one call and the number of valid entries; the SCC starts
// from their linear extrapolation instead of from zero, and the history is advanced when this call converges.
// spec (may be null): the reference's spec_calc output (src/tblite.f90:152-164, src/mo_energ.f90:31-43) -- [nao] orbital energies,
// [nao] occupations, [nao][nat] raw Mulliken population of every orbital on every atom (orbitals in solver order: the host sorts
// and normalises), [1] HOMO index of the alpha channel.
// eigseed (may be null; DevModel::oa only): [QX_OA_NSTORE nao^2 + 1] eigenvectors (transposed, dense nao x nao) of the first SCC cycles of
// the previous call for this trajectory + the number of valid entries.  They seed the eigenpair refinement of the same cycles of
// this call; the SCC protocol (zero start, iterates, stop test) is untouched -- only the eigen-solver's starting guess changes.
__device__ inline void egrad_cta(const DevModel &m, Sm &s, double *scratch, const ScratchLayout &L, double kt, EgradOut &out, double *qstart = nullptr,
                                 double *spec = nullptr, double *eigseed = nullptr) {
    const int nat = m.nat, nsh = m.nsh, nao = m.nao, ld = m.ld, ndim = m.ndim;
    // global-slab mode: the block buffer of the Jacobi doubles as the staging buffer of the GEMMs' B operand
    double *stg = m.mat_in_global && m.jblock > 0 ? s.jblk : nullptr;
    const int stg_cap = 2 * m.jblock * ld;
    double *S = scratch + L.S, *H0 = scratch + L.H0, *Dt = scratch + L.Dt, *Qt = scratch + L.Qt, *T = scratch + L.T;
    double *gamma = scratch + L.gamma, *dcnp = scratch + L.dcnp, *dcnp4 = scratch + L.dcnp4, *edisp = scratch + L.edisp;
    double *c6 = scratch + L.c6, *dc6 = scratch + L.dc6, *taskout = scratch + L.taskout;
    double *t7 = T;                 // [7*nat] temp (T is free whenever t7 is used)
    double *pop = s.pop;            // [11*nao] Mulliken partial sums (shared memory)
    Broyden br;
    {
        double *v = scratch + L.br_vec;
        br.q_in = v; br.qlast = v + ndim; br.dq = v + 2 * ndim; br.dqlast = v + 3 * ndim;
        br.omega = v + 4 * ndim; br.cvec = br.omega + QX_MAX_ITER;
        br.df = scratch + L.br_df; br.u = scratch + L.br_u; br.a = scratch + L.br_a;
        br.beta = br.a + (size_t)QX_MAX_ITER * QX_MAX_ITER;
        br.iter = 0;
    }
    out.stat = 0; out.niter = 0; out.sweeps = 0;

    QX_PH_BEGIN();
    phase_cn(m, s, dcnp, dcnp4);
    out.e_rep = phase_repulsion(m, s);
    QX_PH(0);
    out.e_atm = m.method == 1 ? phase_d3_xb(m, s) : phase_d4_nonsc(m, s, edisp, c6, dc6, taskout);
    QX_PH(1);
    phase_coulomb_setup(m, s, gamma);
    phase_integrals(m, s, S, H0, Dt, Qt);
    QX_PH(2);

    // S-orthonormal start basis: C = L^{-T}.  Padding columns of the shared matrices are zeroed once: the
    // 128-bit row accesses of the Jacobi read (and rewrite) them.
    for (int t = threadIdx.x; t < m.rows8 * ld; t += QX_NT) { s.A[t] = 0.0; s.C[t] = 0.0; }
    int nseed = 0;           // eigenvector seeds of the previous call (eigenpair refinement only)
    bool c_valid = false;    // s.C holds an S-orthonormal basis (Cholesky start basis or the eigenvectors of the previous cycle)
    if (m.oa) {
        for (int t = threadIdx.x; t < m.rows8 * ld; t += QX_NT) { s.X3[t] = 0.0; s.X4[t] = 0.0; s.S5[t] = 0.0; }
        if (eigseed) nseed = (int)__ldcg(eigseed + (size_t)QX_OA_NSTORE * nao * nao);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nao * nao; t += QX_NT) {
        const double v = S[t];
        s.A[(size_t)(t / nao) * ld + t % nao] = v;
        if (m.oa) s.S5[(size_t)(t / nao) * ld + t % nao] = v;
    }
    __syncthreads();
    if (!(m.oa && nseed > 0)) {   // (with a seed for the first cycle the Cholesky start basis is only built if the refinement fails)
        if (!(m.mat_in_global ? cholesky_basis<false>(nao, s.A, s.C, ld, s.red) : cholesky_basis<true>(nao, s.A, s.C, ld, s.red))) { out.stat = -2; out.energy = 0.0; return; }  // hard failure: S not positive definite
        c_valid = true;
    }

    if (qstart) {   // warm start (opt-in): linear extrapolation of the converged populations of the last two steps of this trajectory
        const bool two = __ldcg(qstart + 2 * ndim) >= 2.0;
        for (int i = threadIdx.x; i < ndim; i += QX_NT) {
            const double q0 = __ldcg(qstart + i);
            const double v = two ? 2.0 * q0 - __ldcg(qstart + ndim + i) : q0;
            if (i < nsh) s.qsh[i] = v;
            else if (i < nsh + 3 * nat) s.dpat[i - nsh] = v;
            else s.qpat[i - nsh - 3 * nat] = v;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nat; i += QX_NT) {
            double v = 0.0;
            for (int a = m.at_sh0[i]; a < m.at_sh0[i] + m.at_nsh[i]; ++a) v += s.qsh[a];
            s.qat[i] = v;
        }
    } else {        // reference protocol: zeroed wavefunction on every call (src/tblite.f90:133)
        for (int i = threadIdx.x; i < nsh; i += QX_NT) s.qsh[i] = 0.0;
        for (int i = threadIdx.x; i < nat; i += QX_NT) s.qat[i] = 0.0;
        for (int i = threadIdx.x; i < 3 * nat; i += QX_NT) s.dpat[i] = 0.0;
        for (int i = threadIdx.x; i < 6 * nat; i += QX_NT) s.qpat[i] = 0.0;
    }
    __syncthreads();
    QX_PH(3);

    double eelec = 0.0;
    bool converged = false;
    int iscf = 0;
    while (!converged && iscf < QX_MAX_ITER) {
        const double elast = eelec;
        if (iscf > 0) {
            if (!broyden_next(br, ndim, 0.4, s.red, s.bsol)) { out.stat = -2; break; }
            for (int i = threadIdx.x; i < ndim; i += QX_NT) {
                double v = br.q_in[i];
                if (i < nsh) s.qsh[i] = v;
                else if (i < nsh + 3 * nat) s.dpat[i - nsh] = v;
                else s.qpat[i - nsh - 3 * nat] = v;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < nat; i += QX_NT) {
                double v = 0.0;
                for (int a = m.at_sh0[i]; a < m.at_sh0[i] + m.at_nsh[i]; ++a) v += s.qsh[a];
                s.qat[i] = v;
            }
            __syncthreads();
        }
        QX_PH(4);
        iscf += 1;
        phase_potential(m, s, gamma, edisp, t7);
        for (int i = threadIdx.x; i < ndim; i += QX_NT)
            br.q_in[i] = i < nsh ? s.qsh[i] : (i < nsh + 3 * nat ? s.dpat[i - nsh] : s.qpat[i - nsh - 3 * nat]);
        QX_PH(5);
        if (m.mat_in_global) phase_build_h1<false>(m, s, S, H0, Dt, Qt); else phase_build_h1<true>(m, s, S, H0, Dt, Qt);
        QX_PH(6);
        // A' = C^T H1 C in the current S-orthonormal basis (s.C holds C transposed), then Jacobi (C <- C J)
        const int npad = tc_padded_dim(nao);
        const bool strip = npad != 0 && npad / 8 <= QX_NT / 32;
        bool refined = false;
        if (m.oa) {
            // ---- eigenpair refinement (qx_oa.cuh): seed = the same cycle of the previous call (first cycles) or the previous cycle
            const int kc = iscf - 1;
            const size_t n2 = (size_t)nao * nao;
            if (c_valid) {   // keep the orthonormal basis for the fall-back
                for (int t = threadIdx.x; t < nao * nao; t += QX_NT) T[t] = s.C[(size_t)(t / nao) * ld + t % nao];
            }
            const bool from_seed = eigseed && kc < QX_OA_NSTORE && kc < nseed;
            if (from_seed) {
                __syncthreads();
                for (int t = threadIdx.x; t < nao * nao; t += QX_NT) s.C[(size_t)(t / nao) * ld + t % nao] = __ldcg(eigseed + kc * n2 + t);
            }
            __syncthreads();
            int passes = 0;
            if (from_seed || (c_valid && kc > 0)) {
                const bool bid = !from_seed;
                passes = npad == 32 ? oa_refine<4>(m, s, bid) : (npad == 64 ? oa_refine<8>(m, s, bid) : oa_refine<9>(m, s, bid));
            }
            refined = passes > 0;
            out.sweeps += passes;
            if (!refined) {
                if (c_valid) {   // back to the orthonormal basis of the previous cycle (or the Cholesky basis)
                    __syncthreads();
                    for (int t = threadIdx.x; t < nao * nao; t += QX_NT) s.C[(size_t)(t / nao) * ld + t % nao] = T[t];
                    __syncthreads();
                } else {         // first cycle, seed not good enough: Cholesky start basis now (it needs s.A: H1 is rebuilt afterwards)
                    for (int t = threadIdx.x; t < m.rows8 * ld; t += QX_NT) { s.A[t] = 0.0; s.C[t] = 0.0; }
                    __syncthreads();
                    for (int t = threadIdx.x; t < nao * nao; t += QX_NT) s.A[(size_t)(t / nao) * ld + t % nao] = S[t];
                    __syncthreads();
                    if (!cholesky_basis<true>(nao, s.A, s.C, ld, s.red)) { out.stat = -2; break; }
                    for (int t = threadIdx.x; t < m.rows8 * ld; t += QX_NT) s.A[t] = 0.0;
                    __syncthreads();
                    phase_build_h1<true>(m, s, S, H0, Dt, Qt);
                }
            }
            c_valid = true;
        }
        if (refined) {
            // eigenvalues are in s.emo, eigenvectors in s.C
        } else if (strip) {
            if (npad == 32) tc_transform<4>(nao, s.C, s.A, ld);
            else if (npad == 64) tc_transform<8>(nao, s.C, s.A, ld);
            else tc_transform<9>(nao, s.C, s.A, ld);
        } else {
            const double *Ct = s.C, *Hm = s.A;
            // Tt = Ct H1  (global scratch), then A' = Ct Tt^T back into shared memory
            // (T is kept transposed so that the second product reads it along rows)
            gemm_tc(nao, [=](int i, int k) { return Ct[(size_t)i * ld + k]; }, [=](int k, int j) { return Hm[(size_t)k * ld + j]; },
                    [=](int i, int j, double v) { T[(size_t)j * nao + i] = v; }, stg, stg_cap);
            __syncthreads();
            double *Ap = s.A;
            gemm_tc(nao, [=](int i, int k) { return Ct[(size_t)i * ld + k]; }, [=](int k, int j) { return T[(size_t)k * nao + j]; },
                    [=](int i, int j, double v) { Ap[(size_t)i * ld + j] = v; }, stg, stg_cap);
            __syncthreads();
        }
        QX_PH(7);
        if (!refined) {
            double *gpol = m.polish ? scratch + L.P : nullptr;
            int sw_ = m.mat_in_global ? jacobi_eigh_rows<false>(nao, s.A, ld, s.emo, s.red, s.jw, s.jblk, m.jblock, nullptr, gpol)
                                      : jacobi_eigh_rows<true>(nao, s.A, ld, s.emo, s.red, s.jw, nullptr, 0, s.C, gpol);
            out.sweeps += sw_;
#ifdef QX_PROFILE_PHASES
            if (threadIdx.x == 0 && iscf <= 32) { atomicAdd(&g_sweep_hist[iscf - 1], (unsigned long long)sw_); atomicAdd(&g_sweep_hist[32 + iscf - 1], 1ull); }
#endif
            // rows of A now hold J^T: Ct_new = J^T Ct_old
            if (strip) {
                if (npad == 32) tc_left_apply<4>(nao, s.A, s.C, ld);
                else if (npad == 64) tc_left_apply<8>(nao, s.A, s.C, ld);
                else tc_left_apply<9>(nao, s.A, s.C, ld);
            } else {
                const double *Jt = s.A, *Ct = s.C;
                gemm_tc(nao, [=](int i, int k) { return Jt[(size_t)i * ld + k]; }, [=](int k, int j) { return Ct[(size_t)k * ld + j]; },
                        [=](int i, int j, double v) { T[(size_t)i * nao + j] = v; }, stg, stg_cap);
                __syncthreads();
                for (int t = threadIdx.x; t < nao * nao; t += QX_NT) { int i = t / nao; s.C[(size_t)i * ld + (t - i * nao)] = T[t]; }
                __syncthreads();
            }
        }
        if (m.oa && eigseed && iscf <= QX_OA_NSTORE) {   // seed for the same cycle of the next call
            const size_t n2 = (size_t)nao * nao;
            for (int t = threadIdx.x; t < nao * nao; t += QX_NT) eigseed[(iscf - 1) * n2 + t] = s.C[(size_t)(t / nao) * ld + t % nao];
        }
        QX_PH(8);
        // order statistics needed for the Fermi-level start value
        int homo[2];
        for (int sp = 0; sp < 2; ++sp) {
            double ne = m.nel[sp];
            homo[sp] = (int)floor(ne) + (fmod(ne, 1.0) > 0.5 ? 1 : 0);
        }
        for (int base = 0; base < nao; base += QX_NT / 4) {   // four lanes per orbital share the rank count
            const int k = base + (threadIdx.x >> 2), part = threadIdx.x & 3, kk = k < nao ? k : 0;
            const double ek = s.emo[kk];
            int rank = 0;
            for (int j = part; j < nao; j += 4) rank += (s.emo[j] < ek) || (s.emo[j] == ek && j < kk);
            rank += __shfl_xor_sync(0xffffffffu, rank, 1);
            rank += __shfl_xor_sync(0xffffffffu, rank, 2);
            if (k < nao && part == 0) {
                for (int sp = 0; sp < 2; ++sp) {
                    int lo = (homo[sp] > 1 ? homo[sp] : 1) - 1, hi = (homo[sp] + 1 < nao ? homo[sp] + 1 : nao) - 1;
                    if (rank == lo) s.red[32 + 2 * sp] = ek;
                    if (rank == hi) s.red[33 + 2 * sp] = ek;
                }
            }
        }
        __syncthreads();
        {
            const int warp = threadIdx.x >> 5;
            if (warp < 2) {
                double ef = 0.0;
                if (homo[warp] > 0) ef = fermi_level_warp(nao, homo[warp], kt, s.emo, s.red[32 + 2 * warp], s.red[33 + 2 * warp]);
                if ((threadIdx.x & 31) == 0) s.red[40 + warp] = ef;
            }
        }
        __syncthreads();
        double ts = 0.0;
        for (int k = threadIdx.x; k < nao; k += QX_NT) {
            double f = 0.0;
            for (int sp = 0; sp < 2; ++sp) {
                if (homo[sp] <= 0) continue;
                double x = (s.emo[k] - s.red[40 + sp]) / kt, occ = 0.0;
                if (x < 50.0) occ = 1.0 / (exp(x) + 1.0);
                f += occ;
                if (occ > 1.4901161193847656e-08 && 1.0 - occ > 1.4901161193847656e-08) ts += (occ * log(occ) + (1.0 - occ) * log(1.0 - occ)) * kt;
            }
            s.focc[k] = f;
        }
        ts = block_sum(ts, s.red);
        QX_PH(9);
        // density into A
        if (strip) {
            for (int k = nao + threadIdx.x; k < nao + 8; k += QX_NT) s.focc[k] = 0.0;   // padding orbitals carry no weight
            __syncthreads();
            if (npad == 32) tc_density<4>(nao, s.C, s.focc, s.A, ld);
            else if (npad == 64) tc_density<8>(nao, s.C, s.focc, s.A, ld);
            else tc_density<9>(nao, s.C, s.focc, s.A, ld);
        } else {
            const double *Ct = s.C, *f = s.focc;
            double *Pm = s.A;
            gemm_tc(nao, [=](int i, int k) { return Ct[(size_t)k * ld + i] * f[k]; }, [=](int k, int j) { return Ct[(size_t)k * ld + j]; },
                    [=](int i, int j, double v) { Pm[(size_t)i * ld + j] = v; }, stg, stg_cap);
            __syncthreads();
        }
        QX_PH(10);
        double eel = m.mat_in_global ? phase_mulliken<false>(m, s, S, H0, Dt, Qt, pop) : phase_mulliken<true>(m, s, S, H0, Dt, Qt, pop);
        QX_PH(11);
        double err = 0.0;
        for (int i = threadIdx.x; i < ndim; i += QX_NT) {
            double o = i < nsh ? s.qsh[i] : (i < nsh + 3 * nat ? s.dpat[i - nsh] : s.qpat[i - nsh - 3 * nat]);
            double d = o - br.q_in[i];
            br.dq[i] = d;
            err += d * d;
        }
        err = sqrt(block_sum(err, s.red) / ndim);
        double e_es, e_aes, e_d4;
        phase_scc_energy(m, s, gamma, edisp, e_es, e_aes, e_d4);
        eelec = ts + eel + e_es + e_aes + e_d4;
        QX_PH(12);
        out.e_el = eel; out.e_es = e_es; out.e_aes = e_aes; out.e_d4 = e_d4; out.e_ts = ts;
        converged = fabs(eelec - elast) < 1e-6 && err < 2e-5;
    }
    out.niter = iscf;
    out.energy = out.e_rep + out.e_atm + eelec;
    if (m.oa && eigseed && threadIdx.x == 0) eigseed[(size_t)QX_OA_NSTORE * nao * nao] = out.stat == -2 ? 0.0 : (double)(iscf < QX_OA_NSTORE ? iscf : QX_OA_NSTORE);
    if (out.stat == -2) return;
    if (qstart && converged) {   // hand the converged populations to the next step of this trajectory: [latest | previous | count]
        for (int i = threadIdx.x; i < ndim; i += QX_NT) {
            qstart[ndim + i] = __ldcg(qstart + i);
            qstart[i] = i < nsh ? s.qsh[i] : (i < nsh + 3 * nat ? s.dpat[i - nsh] : s.qpat[i - nsh - 3 * nat]);
        }
        if (threadIdx.x == 0) qstart[2 * ndim] = fmin(__ldcg(qstart + 2 * ndim) + 1.0, 2.0);
    }
    if (spec) {
        for (int k = threadIdx.x; k < nao; k += QX_NT) { spec[k] = s.emo[k]; spec[nao + k] = s.focc[k]; }
        for (int t = threadIdx.x; t < nao * nat; t += QX_NT) {
            const int k = t / nat, ia = t - k * nat;
            const double *ck = s.C + (size_t)k * ld;      // row k of C^T = orbital k
            double q = 0.0;
            for (int j = 0; j < nao; ++j) {
                if (m.ao_at[j] != ia) continue;
                double sc = 0.0;
                for (int l = 0; l < nao; ++l) sc += S[(size_t)j * nao + l] * ck[l];
                q += ck[j] * sc;
            }
            spec[2 * nao + t] = q;
        }
        if (threadIdx.x == 0) {
            const double ne = m.nel[0];
            const int homo = (int)floor(ne) + (fmod(ne, 1.0) > 0.5 ? 1 : 0);
            spec[2 * nao + nao * nat] = (double)(homo > 1 ? homo : 1);
        }
    }
    // "SCF not converged": flagged, but energy and gradient of the last cycle are still handed back -- the
    // reference's egrad overrides stat with checkqc (src/iniqm.f90:646-651)
    if (!converged) out.stat = -1;

    // ---------------- gradient ----------------
    // W = C diag(f e) C^T -> global T -> shared C (A holds P)
    for (int k = threadIdx.x; k < nao; k += QX_NT) s.focc[k] *= s.emo[k];
    __syncthreads();
    {
        const int npad = tc_padded_dim(nao);
        if (npad != 0 && npad / 8 <= QX_NT / 32) {   // W replaces C^T in place
            if (npad == 32) tc_density<4>(nao, s.C, s.focc, s.C, ld);
            else if (npad == 64) tc_density<8>(nao, s.C, s.focc, s.C, ld);
            else tc_density<9>(nao, s.C, s.focc, s.C, ld);
        } else {
            const double *Ct = s.C, *f = s.focc;
            gemm_tc(nao, [=](int i, int k) { return Ct[(size_t)k * ld + i] * f[k]; }, [=](int k, int j) { return Ct[(size_t)k * ld + j]; },
                    [=](int i, int j, double v) { T[(size_t)i * nao + j] = v; }, stg, stg_cap);
            __syncthreads();
            for (int t = threadIdx.x; t < nao * nao; t += QX_NT) s.C[(size_t)(t / nao) * ld + t % nao] = T[t];
            __syncthreads();
        }
    }
    QX_PH(13);
    if (m.mat_in_global) phase_gradient_pairs<false>(m, s, m.task_int, m.ntask_int, taskout); else phase_gradient_pairs<true>(m, s, m.task_int, m.ntask_int, taskout);
    QX_PH(14);
    // per-atom reduction of the task outputs: four lanes share one (atom, component) -- the lists have ~200 entries per atom
    for (int base = 0; base < 4 * nat; base += QX_NT / 4) {
        const int t = base + (threadIdx.x >> 2), part = threadIdx.x & 3;
        const bool on = t < 4 * nat;
        const int k = on ? t >> 2 : 0, c = t & 3;      // c = 0..2: gradient component, c = 3: d/d cn
        double g = 0.0;
        for (int e = m.gr_ptr[k] + part; e < m.gr_ptr[k + 1]; e += 4) {
            const int code = m.gr_task[e];
            const size_t o = 5 * (size_t)(code >> 1);
            if (c < 3) g += (code & 1) ? -taskout[o + c] : taskout[o + c];
            else g += taskout[o + ((code & 1) ? 4 : 3)];
        }
        g += __shfl_xor_sync(0xffffffffu, g, 1);
        g += __shfl_xor_sync(0xffffffffu, g, 2);
        if (on && part == 0) {
            if (c < 3) s.grad[3 * k + c] += g;
            else {
                for (int mu = m.at_ao0[k]; mu < m.at_ao0[k] + m.at_nao[k]; ++mu) g += -m.sh_kcn[m.ao_sh[mu]] * s.A[(size_t)mu * ld + mu];
                s.dEdcn[k] += g;
            }
        }
    }
    __syncthreads();
    // D4 two-body with the final charges
    const bool gfn1 = m.method == 1;
    if (!gfn1) {
        double *gwq = taskout, *gwdcnq = taskout + 7 * nat;
        d4_weights_all(m, s, true, gwq, gwdcnq, nullptr);
        __syncthreads();
        d4_c6_tables(m, gwq, gwdcnq, c6, dc6);
        __syncthreads();
    }
    for (int i = threadIdx.x >> 5; i < nat; i += QX_NT / 32) {   // one warp per atom, lanes over the partners
        const int lane = threadIdx.x & 31;
        double gx = 0, gy = 0, gz = 0, dcn = 0, dcn4 = 0;
        const double *mi = s.dpat + 3 * i, *ti = s.qpat + 6 * i;
        const double qi = s.qat[i];
        for (int j = lane; j < (gfn1 ? 0 : nat); j += 32) {   // GFN1: D3 and no multipoles -- both handled before the SCC / absent
            if (j == i) continue;
            // --- dispersion
            {
                double vx = s.xyz[3 * i] - s.xyz[3 * j], vy = s.xyz[3 * i + 1] - s.xyz[3 * j + 1], vz = s.xyz[3 * i + 2] - s.xyz[3 * j + 2];
                double r2 = vx * vx + vy * vy + vz * vz;
                if (r2 <= 3600.0) {
                    double r0 = bj_r0(m, i, j), rrij = 3.0 * m.at_r4r2[i] * m.at_r4r2[j];
                    double r02 = r0 * r0, r06 = r02 * r02 * r02, r6 = r2 * r2 * r2;
                    double t6 = 1.0 / (r6 + r06), t8 = 1.0 / (r6 * r2 + r06 * r02);
                    double gdisp = GFN2_D4_S6 * (-6.0 * r2 * r2 * t6 * t6) + GFN2_D4_S8 * rrij * (-8.0 * r6 * t8 * t8);
                    double f = -c6[i * nat + j] * gdisp;
                    gx += f * vx; gy += f * vy; gz += f * vz;
                    dcn4 -= dc6[i * nat + j] * edisp[i * nat + j];
                }
            }
            // --- anisotropic electrostatics, v = R_j - R_i
            {
                double v[3] = {s.xyz[3 * j] - s.xyz[3 * i], s.xyz[3 * j + 1] - s.xyz[3 * i + 1], s.xyz[3 * j + 2] - s.xyz[3 * i + 2]};
                double r2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], r = sqrt(r2), g1 = 1.0 / r, g3 = g1 * g1 * g1, g5 = g3 * g1 * g1;
                double R0 = 0.5 * (s.mrad[i] + s.mrad[j]), rr = R0 * g1, rr3 = rr * rr * rr;
                double x3 = 6.0 * rr3, x5 = 6.0 * rr3 * rr;
                double f3 = 1.0 / (1.0 + x3), f5 = 1.0 / (1.0 + x5);
                double df3dr = f3 * f3 * GFN2_MP_DMP3 * x3 * g1, df5dr = f5 * f5 * GFN2_MP_DMP5 * x5 * g1;
                double df3dR0 = -f3 * f3 * GFN2_MP_DMP3 * x3 / R0, df5dR0 = -f5 * f5 * GFN2_MP_DMP5 * x5 / R0;
                const double *mj = s.dpat + 3 * j, *tj = s.qpat + 6 * j;
                const double qj = s.qat[j];
                double miv = mi[0] * v[0] + mi[1] * v[1] + mi[2] * v[2], mjv = mj[0] * v[0] + mj[1] * v[1] + mj[2] * v[2];
                double mimj = mi[0] * mj[0] + mi[1] * mj[1] + mi[2] * mj[2];
                double tiv[3] = {ti[0] * v[0] + ti[1] * v[1] + ti[3] * v[2], ti[1] * v[0] + ti[2] * v[1] + ti[4] * v[2], ti[3] * v[0] + ti[4] * v[1] + ti[5] * v[2]};
                double tjv[3] = {tj[0] * v[0] + tj[1] * v[1] + tj[3] * v[2], tj[1] * v[0] + tj[2] * v[1] + tj[4] * v[2], tj[3] * v[0] + tj[4] * v[1] + tj[5] * v[2]};
                double tivv = tiv[0] * v[0] + tiv[1] * v[1] + tiv[2] * v[2], tjvv = tjv[0] * v[0] + tjv[1] * v[1] + tjv[2] * v[2];
                double Aa = qj * miv - qi * mjv, Bb = qj * tivv + qi * tjvv, Dd = -3.0 * miv * mjv;
                double dg3 = -3.0 * g3 * g1, dg5 = -5.0 * g5 * g1;
                double radial = (dg3 * f3 + g3 * df3dr) * Aa + (dg5 * f5 + g5 * df5dr) * (Bb + Dd) + (dg3 * f5 + g3 * df5dr) * mimj;
                double dER0 = g3 * df3dR0 * Aa + g5 * df5dR0 * (Bb + Dd) + g3 * df5dR0 * mimj;
                double dv[3];
                for (int k = 0; k < 3; ++k)
                    dv[k] = radial * v[k] * g1 + g3 * f3 * (qj * mi[k] - qi * mj[k]) +
                            g5 * f5 * (2.0 * qj * tiv[k] + 2.0 * qi * tjv[k] - 3.0 * (mi[k] * mjv + mj[k] * miv));
                gx -= dv[0]; gy -= dv[1]; gz -= dv[2];
                dcn += 0.5 * dER0 * s.dmr[i];
            }
        }
        // --- isotropic second order
        for (int a = m.at_sh0[i]; a < m.at_sh0[i] + m.at_nsh[i]; ++a)
            for (int b = lane; b < nsh; b += 32) {
                int j = m.sh_at[b];
                if (j == i) continue;
                double g = gamma[a * nsh + b];
                double f = -s.qsh[a] * s.qsh[b] * g * g * g;
                gx += f * (s.xyz[3 * i] - s.xyz[3 * j]); gy += f * (s.xyz[3 * i + 1] - s.xyz[3 * j + 1]); gz += f * (s.xyz[3 * i + 2] - s.xyz[3 * j + 2]);
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            gx += __shfl_xor_sync(0xffffffffu, gx, o); gy += __shfl_xor_sync(0xffffffffu, gy, o); gz += __shfl_xor_sync(0xffffffffu, gz, o);
            dcn += __shfl_xor_sync(0xffffffffu, dcn, o); dcn4 += __shfl_xor_sync(0xffffffffu, dcn4, o);
        }
        if (lane == 0) {
            s.grad[3 * i] += gx; s.grad[3 * i + 1] += gy; s.grad[3 * i + 2] += gz;
            s.dEdcn[i] += dcn;
            s.dEdcn4[i] += dcn4;
        }
    }
    __syncthreads();
    // chain rule through both coordination numbers
    for (int k = threadIdx.x; k < nat; k += QX_NT) {
        double gx = 0, gy = 0, gz = 0;
        for (int j = 0; j < nat; ++j) {
            if (j == k) continue;
            double f = (s.dEdcn[k] + s.dEdcn[j]) * dcnp[k * nat + j] + (s.dEdcn4[k] + s.dEdcn4[j]) * dcnp4[k * nat + j];
            gx += f * (s.xyz[3 * k] - s.xyz[3 * j]); gy += f * (s.xyz[3 * k + 1] - s.xyz[3 * j + 1]); gz += f * (s.xyz[3 * k + 2] - s.xyz[3 * j + 2]);
        }
        s.grad[3 * k] += gx; s.grad[3 * k + 1] += gy; s.grad[3 * k + 2] += gz;
    }
    __syncthreads();
    QX_PH(15);
}

}  // namespace qx
