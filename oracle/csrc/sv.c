/*
 * oracle/csrc/sv.c -- CPU state-vector sweeps for the parity ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under surface_17_iqt_b200/ may link, load or call this
 * file; it exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs have an independent CPU answer to compare the CUDA path with.
 *
 * What it restates: the execution model of ProjectQ's C++ `_cppsim` Simulator, the third-party
 * dependency in which the reference's arithmetic lives (reference README.md:4; call sites
 * compiled_surface_code_arbitrary_error_model.py:5 and every `Gate | qubit` line there).
 * ProjectQ is NOT vendored under /root/reference and its version is unpinned, so this follows
 * ProjectQ's published algorithm:
 *   - state = complex<double>[2^n], little-endian, k-th allocated qubit = bit k;
 *   - one full sweep over the state per gate (Simulator() default gate_fusion=False);
 *   - Measure = choose outcome, zero the disagreeing half while accumulating the norm,
 *     rescale by 1/sqrt(norm).
 * PARITY UNPINNED against real ProjectQ (absent from this image); pinned against the reference's
 * own Python control flow through oracle/pqshim (see oracle/make_golden.py).
 */
#include <complex.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef double complex cplx;

/* general 2x2 complex matrix on qubit q; m is row-major [m00 m01 m10 m11] */
void sv_apply_1q(cplx *psi, int n, int q, const cplx *m)
{
    const size_t dim = (size_t)1 << n, st = (size_t)1 << q;
    const cplx m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
    for (size_t base = 0; base < dim; base += 2 * st) {
        for (size_t i = base; i < base + st; ++i) {
            const cplx a0 = psi[i], a1 = psi[i + st];
            psi[i] = m00 * a0 + m01 * a1;
            psi[i + st] = m10 * a0 + m11 * a1;
        }
    }
}

/* general 4x4 complex matrix on (q0,q1); matrix index bit0 <-> q0, bit1 <-> q1 (ProjectQ order) */
void sv_apply_2q(cplx *psi, int n, int q0, int q1, const cplx *m)
{
    const size_t dim = (size_t)1 << n, s0 = (size_t)1 << q0, s1 = (size_t)1 << q1;
    for (size_t i = 0; i < dim; ++i) {
        if (i & (s0 | s1))
            continue;
        cplx a[4] = { psi[i], psi[i | s0], psi[i | s1], psi[i | s0 | s1] };
        for (int r = 0; r < 4; ++r) {
            cplx acc = 0;
            for (int c = 0; c < 4; ++c)
                acc += m[4 * r + c] * a[c];
            psi[i | ((r & 1) ? s0 : 0) | ((r & 2) ? s1 : 0)] = acc;
        }
    }
}

/* P(bit q == 1) */
double sv_prob_one(const cplx *psi, int n, int q)
{
    const size_t dim = (size_t)1 << n, st = (size_t)1 << q;
    double p = 0.0;
    for (size_t i = 0; i < dim; ++i)
        if (i & st)
            p += creal(psi[i]) * creal(psi[i]) + cimag(psi[i]) * cimag(psi[i]);
    return p;
}

/* ProjectQ-style outcome choice: walk cumulative |amp|^2 until it exceeds rnd, return that basis state */
uint64_t sv_sample_basis(const cplx *psi, int n, double rnd)
{
    const size_t dim = (size_t)1 << n;
    double P = 0.0;
    size_t pick = 0;
    for (size_t i = 0; i < dim; ++i) {
        P += creal(psi[i]) * creal(psi[i]) + cimag(psi[i]) * cimag(psi[i]);
        if (P > rnd) { pick = i; return pick; }
    }
    return dim - 1;
}

/* collapse qubit q onto `bit`, renormalise; returns the pre-normalisation probability */
double sv_collapse(cplx *psi, int n, int q, int bit)
{
    const size_t dim = (size_t)1 << n, st = (size_t)1 << q;
    const size_t want = bit ? st : 0;
    double N = 0.0;
    for (size_t i = 0; i < dim; ++i) {
        if ((i & st) == want)
            N += creal(psi[i]) * creal(psi[i]) + cimag(psi[i]) * cimag(psi[i]);
        else
            psi[i] = 0;
    }
    if (N > 0.0) {
        const double s = 1.0 / sqrt(N);
        for (size_t i = 0; i < dim; ++i)
            psi[i] *= s;
    }
    return N;
}

void sv_init_zero(cplx *psi, int n)
{
    memset(psi, 0, sizeof(cplx) << n);
    psi[0] = 1.0;
}
