// s17_diag.cuh -- one whole stabiliser cycle per shot in the frame where its gates are diagonal (k_cycle).
//
// Both halves of the compiled cycle are diagonal in a product basis of the nine data qubits:
//   * timesteps 1-4 (x_type_entangling, CS:291-309) contain only Rxx(+-pi/2) and Rx(-+pi/2): functions of X_d and
//     X_d X_a, diagonal after a Walsh-Hadamard transform of the data register;
//   * timesteps 5-8 (z_type_entangling, CS:312-344) bracket the same gates between Ry(+pi/2) ... Ry(-pi/2) on each data
//     qubit, and Ry(-pi/2) f(X) Ry(+pi/2) = f(Z): diagonal in the computational basis as they stand.
// With lambda = (-1)^{bit d of the basis label} an operation (d, a, s) is
//     Rxx(s pi/2)  ->  (1 - i s lambda X_a)/sqrt2  on ancilla a,       Rx(-s pi/2)  ->  the phase (1 + i s lambda)/sqrt2,
// so for a fixed label the four ancillas stay in a PRODUCT state that depends only on the <= 4 label bits of the data
// qubits the ancilla talks to.  Per shot and half cycle the kernel therefore builds, per ancilla, a 16-entry table of
// (amplitude for outcome 0, amplitude for outcome 1, their squared moduli) and the whole half cycle is
//     for each ancilla:  P(0), P(1) = sum_x |psi_x|^2 |T[idx(x)][.]|^2;  sample;  psi_x *= T[idx(x)][outcome].
// Pauli events (k_plan's net end-of-operation Pauli, conjugated into the frame) are label relabelings (flip mask),
// signs (sign mask, global phase) and ancilla X/Z applied while the tables are built; a leak-skipped Rxx is the
// identity.  The derivation is executable: tests/diag_model.py, checked against a brute-force simulation in
// tests/test_diag_frame_model.py.
//
// One WARP owns one shot: the 512 amplitudes live in registers (16 per lane) for the whole cycle, the Walsh-Hadamard
// transform is two in-register radix-16 passes, one shuffle stage and one exchange through shared memory.  HBM traffic
// per shot and cycle: 8 KiB read + 8 KiB written (+ ~150 B of metadata), prefetched one shot ahead by TMA.
// Exact same measurement semantics as k_half / k_measure (one Measure per ancilla, conditional on the earlier ones).
//
// Kernels in this file:
//   k_cycle   the algorithm with fixed label <-> (lane, register) maps (up to 16 table entries per lane and slot);
//   k_cycle1  the production form: maps searched by the host (make_diag1 in s17.cu) so that every slot indexes its
//             table with one register bit (two entries per lane), event-free tables shared per CTA, the data
//             read-out drawn from the registers, bulk store of the staged block;
//   k_prep    the memoised noiseless preparation cycle (replays k_cycle1's own decisions per preparation leaf);
//   k_quiet   the memoised event-free first noisy cycle (replays k_cycle1's cached result for the shot's leaf).
// Included by s17.cu after MeasArgs.
#pragma once

namespace s17 {

#ifndef S17_DG_WARPS
#define S17_DG_WARPS 12
#endif
#ifndef S17_DG_ST_UNROLL
#define S17_DG_ST_UNROLL 1      // the two-step loop of a cycle (1: the transform and the half-cycle code exist once)
#endif
#ifndef S17_DG_HP_UNROLL
#define S17_DG_HP_UNROLL 1      // the two passes of a transform
#endif
#ifndef S17_DG_XLANE
#define S17_DG_XLANE 1          // 1: the one lane-carried label bit of a transform rides in the exchange load (no shuffles)
#endif
constexpr int kDgStUnroll = S17_DG_ST_UNROLL, kDgHpUnroll = S17_DG_HP_UNROLL;
constexpr bool kDgXlane = S17_DG_XLANE != 0;
constexpr int kDgWarps = S17_DG_WARPS;                         // warps (= shots in flight) per SM
constexpr int kDgThreads = kDgWarps * 32;
constexpr int kDgFrmBytes = 16;                                // per-shot frame words (4 x u32, two used)
constexpr int kDgMetaBytes = kEvBytes + kDgFrmBytes;
constexpr int kDgTabEntries = 40;                              // table entries of one half cycle (sum over slots of 2^n_ops)
constexpr int kDgTabBytes = 3 * kDgTabEntries * 16;            // A0 | A1 | Q
constexpr int kDgSmallBytes = kDgTabBytes + 2 * kDgMetaBytes + 64;       // per warp: tables | metadata [2] | mbarriers
constexpr int kDgWarpBytes = 2 * kRunBytes + kDgSmallBytes;
constexpr size_t kDgSmem = (size_t)kDgWarps * kDgWarpBytes;
// k_cycle1 adds, per CTA, the tables of an event-free half cycle (both halves) and 128 zero bytes of event words
constexpr int kDgCleanBytes = 2 * kDgTabBytes + kEvBytes;
static_assert(kDgSmallBytes % 16 == 0, "per-warp shared memory must keep 16-byte alignment");
// The 8 KiB block buffers are aligned to their own size IN THE SHARED WINDOW, so that a lane's address is
// (buffer | lane part) ^ register part: one LOP3 with a constant-bank operand per access.  The dynamic segment starts
// wherever the driver puts it (1 KiB into the window today), so the kernel skips `gap` < 8 KiB bytes and uses the gap for
// the per-CTA tables when they fit; the launch always asks for the maximum.
constexpr size_t kDgSmem1 = 232448;
#ifndef S17_NO_SMEM_ASSERT
static_assert(kDgSmem + kDgCleanBytes + (kDgCleanBytes - 1) <= kDgSmem1 && kDgSmem + (kRunBytes - 1) <= kDgSmem1,
              "k_cycle1: shared memory per SM (either placement of the per-CTA tables)");
static_assert((kRunBytes & (kRunBytes - 1)) == 0, "block buffers are aligned to their size");
#endif

struct DiagOp {
    uint8_t d, ev_word;
    uint8_t flags;             // bit0: Rxx sign is -1, bit1: followed by Rx(-s pi/2), bit2: this slot owns label bit d for signs
    uint8_t pad;               // index of the operation within its half cycle, in time order (bit 20 + pad of the frame word)
};
struct DiagHalf {
    uint8_t frame_x;           // 1: Walsh-Hadamard frame (x-type half), 0: computational frame (z-type half)
    uint8_t final_part;        // 1: completes the cycle's All(Measure) | ancilla
    uint8_t n_entries;         // table entries in use
    uint8_t pad;
    uint8_t n_ops[4];          // operations per ancilla slot
    uint8_t aj[4];             // ancilla index (outcome bit) of slot k
    uint8_t tbase[4];          // first table entry of slot k
    DiagOp ops[4][4];          // per slot, in time order; table index bit j = label bit of ops[k][j].d
    uint8_t lane_bit[4][4];    // lane bit that feeds table index bit j (255: the bit comes from the register index)
    uint32_t roff[4][16];      // register-index part of the table index, as a byte offset (index << 4)
};
struct DiagProg { DiagHalf h[2]; };
constexpr int kDgCdfDoubles = 32 + 512;    // read-out weights of one cached block: inclusive lane sums + per-state weights

// k_cycle1: label <-> (lane, register) mappings chosen by the host so that every slot of a half cycle indexes its
// table with at most ONE register bit: a lane then needs two table entries per slot instead of sixteen.
struct DiagMap {
    uint8_t reg[4];            // label bit carried by register index bit j
    uint8_t lane[5];           // label bit carried by lane bit m
    uint8_t pad[3];
    uint32_t lin16[16];        // label part of register i, as a byte offset into a linear 512-amplitude buffer
    uint32_t swz16[16];        // the same through the exchange swizzle
    uint32_t swzp16[16];       // swz16 of the partner lane: the lane that differs in the label bit neither map keeps in registers
};
struct DiagProg1 {
    DiagHalf h[2];             // lane_bit refers to the map of the half (B for the Hadamard frame, A otherwise)
    DiagMap mapA, mapB;        // A: load / store and the computational-frame half; B: the Hadamard-frame half
    uint8_t sh_lane_bit;       // lane bit (map A) of the one label bit transformed by the shuffle stage
    uint8_t gvec[9];           // exchange swizzle: 16-byte column ^= gvec[b] for every set label bit b >= 3
    uint8_t sh_lane_bit_b;     // the same label bit as a lane bit of map B
    uint8_t pad[1];
    uint16_t own[2][4];        // label bits whose sign slot k carries (DiagOp flag bit 2), per half
    uint32_t off1[2][4];       // slot k's table index takes register index bit k: byte offset from the entry for that
                               // bit = 0 to the one for bit = 1 (0: the slot does not depend on a register bit)
    // per half and slot, for the private entries of a shot (dg_slot_entry): operation j at byte j
    uint32_t ev4[2][4];        // event word index of operation j (0 beyond n_ops)
    uint32_t pad4[2][4];       // its index within the half cycle, in time order
    uint32_t f1s[2][4];        // 0x01 where the Rxx sign is -1 (DiagOp flag bit 0)
    uint32_t f2s[2][4];        // 0x01 where an Rx follows (flag bit 1)
    uint32_t mm4[2][4];        // 0x01 per operation in use
};

// Coherent over-rotations of a noisy cycle (k_cycle1<true>; the `*_over` entries of an error model, SURVEY C-10).
//   * Rxx(s pi/2) Rxx(eps) and Rx(-s pi/2) Rx(eps) are still functions of X_d X_a / X_d: diagonal in the same frames,
//     their tables just stop being Gaussian integers.  rxx / rx hold sqrt2 (cos, sin) of the total half angle per
//     operation (Rxx sign folded in): (u, v) -> (c u - i lambda sn v, c v - i lambda sn u), phase (c - i lambda sn).
//   * Ry(+pi/2) Ry(eps_a) ... Ry(-pi/2) Ry(eps_b) = Ry(eps_b) [Ry(-pi/2) ... Ry(+pi/2)] Ry(eps_a): the bracket stays
//     diagonal, with one real rotation per data qubit before and after the computational-frame half.  The one before
//     is merged into the butterflies of the inverse Walsh-Hadamard transform that precedes it (Ry(eps) H =
//     [[a, b], [b, -a]], a = c - s, b = c + s), the one after is a pass of its own.
// Pauli events must stay Paulis in the frame: the host only selects this kernel for models whose stochastic entries
// are signs / ancilla Paulis where they act (s17.cu: coherent_events_ok).
struct DiagCoh {
    double2 rxx[2][4][4];      // per half, slot, operation
    double2 rx[2][4][4];       // (1, 0) for an operation without Rx
    double2 preB[4], preSh, preA[4];   // (a, b) per register bit of map B, the shuffle bit, per register bit of map A
    double2 postR[4], postL[5];        // (cos, sin) of eps_b / 2 per register bit / lane bit of map A
};

__device__ __forceinline__ double2 dg_cmul(const double2 a, const double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// shared-memory accesses by 32-bit window address (k_cycle1 forms them with one XOR, see kDgSmem1)
__device__ __forceinline__ double2 dg_lds(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t dg_lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void dg_sts(uint32_t a, const double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void dg_wht16(double2 (&p)[16]) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i & (1 << b)) continue;
            const double2 u = p[i], v = p[i | (1 << b)];
            p[i] = make_double2(u.x + v.x, u.y + v.y);
            p[i | (1 << b)] = make_double2(u.x - v.x, u.y - v.y);
        }
    }
}
// label bits a lane carries under a map
__device__ __forceinline__ uint32_t lane_label(const DiagMap &M, int lane) {
    uint32_t l = 0;
#pragma unroll
    for (int m = 0; m < 5; ++m)
        if ((lane >> m) & 1) l |= 1u << M.lane[m];
    return l;
}

// per-ancilla tables of one half cycle: up to two of the (slot, index) entries per lane.
// zmask: label bits whose value flips the sign of the amplitude (data Z events; data flips of the preceding
// Hadamard-frame half), applied by the slot that owns the bit; g: global phase i^g, applied by the last slot
// slots: bit k set = build the entries of slot k (the others keep using the event-free tables)
// ancz: bit i set = an idle Z hits the ancilla after operation i of the half cycle (sympathetic_cooling, CS:762-772)
__device__ __forceinline__ void dg_build_tables(const DiagHalf &H, const uint32_t *evw, double2 *tab, int lane,
                                                uint32_t zmask, uint32_t g, uint32_t ancz, uint32_t slots = 15u) {
#pragma unroll 1
    for (int e = lane; e < (int)H.n_entries; e += 32) {
        const int k = (e >= H.tbase[1]) + (e >= H.tbase[2]) + (e >= H.tbase[3]);
        if (!((slots >> k) & 1u)) continue;
        const int idx = e - H.tbase[k];
        double2 u = make_double2(1.0, 0.0), v = make_double2(0.0, 0.0), ph = make_double2(1.0, 0.0);
        const int n = H.n_ops[k];
        for (int j = 0; j < n; ++j) {
            const DiagOp op = H.ops[k][j];
            const uint32_t w = evw[op.ev_word];
            // sigma = s lambda, lambda evaluated on the label bit as relabeled by the data flips so far (bit 31)
            const uint32_t bit = (uint32_t)(idx >> j) & 1u;
            if ((op.flags & 4u) && ((zmask >> op.d) & 1u) && bit) ph = make_double2(-ph.x, -ph.y);
            const uint32_t neg = (uint32_t)(op.flags & 1u) ^ bit ^ (w >> 31);
            const double sg = neg ? -1.0 : 1.0;
            if (!((w >> 24) & 1u)) {                          // Rxx unless leak-skipped: (u, v) -> (u - i sg v, v - i sg u)
                const double2 nu = make_double2(fma(sg, v.y, u.x), fma(-sg, v.x, u.y));
                const double2 nv = make_double2(fma(sg, u.y, v.x), fma(-sg, u.x, v.y));
                u = nu; v = nv;
            }
            if (op.flags & 2u) ph = make_double2(fma(-sg, ph.y, ph.x), fma(sg, ph.x, ph.y));   // (1 + i sg) ph
            if ((w >> 28) & 1u) v = make_double2(-v.x, -v.y);                                  // net event: Z on the ancilla
            if ((w >> 27) & 1u) { const double2 t = u; u = v; v = t; }                         //            X on the ancilla
            if ((ancz >> op.pad) & 1u) v = make_double2(-v.x, -v.y);                           // idle Z after the operation
        }
        if (k == 3) {
            if (g == 1) ph = make_double2(-ph.y, ph.x);
            else if (g == 2) ph = make_double2(-ph.x, -ph.y);
            else if (g == 3) ph = make_double2(ph.y, -ph.x);
        }
        const double2 a0 = dg_cmul(ph, u), a1 = dg_cmul(ph, v);
        tab[e] = a0;
        tab[e + kDgTabEntries] = a1;
        tab[e + 2 * kDgTabEntries] = make_double2(fma(a0.x, a0.x, a0.y * a0.y), fma(a1.x, a1.x, a1.y * a1.y));
    }
}

// ------------------------------------------------------------------------------------------------------------------
// k_cycle1: the same cycle with host-chosen mappings (DiagProg1).  Per slot a lane needs two table entries (register
// bit k of the amplitude index <-> slot k), which makes the half cycle cheap:
//   * the four Measures only need the REAL weights |psi|^2: sixteen per lane, reduced once to marginal sums over the
//     register bits; slot k costs a handful of multiply-adds with the (|A0|^2, |A1|^2) pairs of its two entries and the
//     factors already selected by slots < k, then one warp reduction of (P0, P1);
//   * the complex amplitudes are updated ONCE per half cycle with the product of the four selected factors
//     (F01[b0 b1] F23[b2 b3]: 8 + 32 complex multiplications instead of 64);
//   * per-shot events touch few table entries: a data flip upstream of an operation is an XOR on the table index, a
//     label-dependent sign or the global phase multiplies the selected factors, and only a slot whose own operations
//     carry an ancilla Pauli / a leak skip / an idle Z gets private entries (exact integer arithmetic: every entry is a
//     Gaussian integer);
//   * the renormalisation rides in the last factor table; a preparation cycle starts from the Walsh-Hadamard transform
//     of its basis state, which is a sign pattern.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 dg_neg_if(const double2 a, uint32_t neg) {       // (-1)^neg a, sign-bit flip
    const int s = (int)(neg << 31);
    return make_double2(__hiloint2double(__double2hiint(a.x) ^ s, __double2loint(a.x)),
                        __hiloint2double(__double2hiint(a.y) ^ s, __double2loint(a.y)));
}
__device__ __forceinline__ double2 dg_times_i_pow(const double2 a, uint32_t g) {    // i^g a
    const double2 r = (g & 1u) ? make_double2(-a.y, a.x) : a;
    return dg_neg_if(r, (g >> 1) & 1u);
}

// Private entries of ONE slot (k) for one shot: lanes with idx < 2^n_ops[k] compute entry idx.  Events: bit 31 data flip
// seen by the operation, bit 24 leak skip, bits 27 / 28 net X / Z on the ancilla after the operation, ancz idle Z after
// it.  Signs and the global phase are NOT part of the entries (the caller applies them to the selected factors).
//
// Closed form of the recursion  (u, v) -> (u - i s v, v - i s u) [unless skipped],  ph *= (1 + i s) [Rx],  Z, X, Z  per
// operation: in the X_a eigenbasis a = u + v, b = u - v an Rxx multiplies a by (1 - i s) and b by (1 + i s), Z swaps a and
// b, X negates b.  With T_j = parity of the Z's before operation j the two accumulators are
//     A = +-(1 - i)^(n+ - n-) 2^min(n+, n-),   B = +-conj(.),   n+- = live operations with s_j (-1)^T_j = +-1,
// each X_j negating the one that is `b` at its time; a final swap if the number of Z's is odd; ph = (1 + i)^p+ (1 - i)^p-.
// Exact Gaussian-integer arithmetic, identical to the recursion on all 2 x 10^8 event patterns (tests/test_slot_entry.py).
__host__ __device__ __forceinline__ int dg_popc(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
// (1 - i)^d: real and imaginary part for |d| = 0..4 as signed nibbles (d < 0: the conjugate)
__host__ __device__ __forceinline__ void dg_pow1mi(int d, int &re, int &im) {
    const int ad = d < 0 ? -d : d;
    re = (int)(0x00CE011u << (28 - 4 * ad)) >> 28;          // 1, 1, 0, -2, -4
    im = (int)(0x00EEF0u << (28 - 4 * ad)) >> 28;           // 0, -1, -2, -2, 0
    im = d < 0 ? -im : im;
}
// W: byte j = event word of operation j >> 24; f1s / f2s / mm: DiagProg1; idxs: table index bits, cs: idle Z after the
// operation -- every mask with operation j at bit 8 j
__host__ __device__ __forceinline__ void dg_slot_entry(uint32_t W, uint32_t f1s, uint32_t f2s, uint32_t idxs, uint32_t cs,
                                                       uint32_t mm, int &a0r, int &a0i, int &a1r, int &a1i) {
    const uint32_t Fm = (W >> 7) & mm, Bm = (W >> 4) & mm, Xm = (W >> 3) & mm, Km = W & mm;
    const uint32_t S = (f1s ^ idxs ^ Fm) & mm;                               // sigma_j = -1
    const uint32_t Zm = (Bm ^ cs) & mm;
    const uint32_t T = ((Zm << 8) ^ (Zm << 16) ^ (Zm << 24)) & mm;           // parity of the ancilla Z's before operation j
    const uint32_t Sp = S ^ T, live = mm & ~Km;
    const int nminus = dg_popc(Sp & live), nplus = dg_popc(~Sp & live);
    const uint32_t tx = T ^ Bm;                                              // which accumulator is `b` when X_j acts
    const int sA = dg_popc(Xm & tx) & 1, sB = dg_popc(Xm & ~tx) & 1;
    int x, y;
    dg_pow1mi(nplus - nminus, x, y);
    const int mn = nplus < nminus ? nplus : nminus;
    x *= 1 << mn;
    y *= 1 << mn;
    int Ar = sA ? -x : x, Ai = sA ? -y : y, Br = sB ? -x : x, Bi = sB ? y : -y;
    if (dg_popc(Zm) & 1) { int t = Ar; Ar = Br; Br = t; t = Ai; Ai = Bi; Bi = t; }
    const int ur = (Ar + Br) >> 1, ui = (Ai + Bi) >> 1, vr = (Ar - Br) >> 1, vi = (Ai - Bi) >> 1;
    const int pp = dg_popc(~S & f2s & mm), pm = dg_popc(S & f2s & mm);
    int pr, pi;
    dg_pow1mi(pm - pp, pr, pi);
    const int mp = pp < pm ? pp : pm;
    pr *= 1 << mp;
    pi *= 1 << mp;
    a0r = pr * ur - pi * ui; a0i = pr * ui + pi * ur;
    a1r = pr * vr - pi * vi; a1i = pr * vi + pi * vr;
}
__device__ __forceinline__ void dg_build_slot(const DiagProg1 &P, int h, const uint32_t *evw, double2 *tab, int k, int idx,
                                              uint32_t ancz) {
    const uint32_t e4 = P.ev4[h][k];
    const uint32_t w0 = evw[e4 & 0xFFu], w1 = evw[(e4 >> 8) & 0xFFu], w2 = evw[(e4 >> 16) & 0xFFu], w3 = evw[e4 >> 24];
    const uint32_t W = __byte_perm(__byte_perm(w0, w1, 0x0073), __byte_perm(w2, w3, 0x0073), 0x5410);
    uint32_t cs = 0;
    if (ancz) {
        const uint32_t p4 = P.pad4[h][k];
#pragma unroll
        for (int j = 0; j < 4; ++j) cs |= ((ancz >> ((p4 >> (8 * j)) & 0xFFu)) & 1u) << (8 * j);
    }
    int a0r, a0i, a1r, a1i;
    dg_slot_entry(W, P.f1s[h][k], P.f2s[h][k], ((uint32_t)idx * 0x00204081u) & 0x01010101u, cs, P.mm4[h][k], a0r, a0i, a1r, a1i);
    const int e = P.h[h].tbase[k] + idx;
    tab[e] = make_double2((double)a0r, (double)a0i);
    tab[e + kDgTabEntries] = make_double2((double)a1r, (double)a1i);
    tab[e + 2 * kDgTabEntries] = make_double2((double)(a0r * a0r + a0i * a0i), (double)(a1r * a1r + a1i * a1i));
}
// one table entry (slot k, index idx) of a half cycle with coherent over-rotations: the same recursion in doubles.
// Events as dg_build_slot (data flip seen by the operation, leak skip, net ancilla Pauli, idle Z).
__device__ __forceinline__ void dg_build_entry_coh(const DiagHalf &H, const double2 (&rxx)[4][4], const double2 (&rx)[4][4],
                                                   const uint32_t *evw, double2 *tab, int k, int idx, uint32_t ancz) {
    double2 u = make_double2(1.0, 0.0), v = make_double2(0.0, 0.0), ph = make_double2(1.0, 0.0);
    const int n = H.n_ops[k];
    for (int j = 0; j < n; ++j) {
        const DiagOp op = H.ops[k][j];
        const uint32_t w = evw[op.ev_word];
        const double lam = ((((uint32_t)(idx >> j)) ^ (w >> 31)) & 1u) ? -1.0 : 1.0;
        if (!((w >> 24) & 1u)) {                              // Rxx(theta) unless leak-skipped (its over-rotation with it)
            const double c = rxx[k][j].x, sn = lam * rxx[k][j].y;
            const double2 nu = make_double2(fma(sn, v.y, c * u.x), fma(-sn, v.x, c * u.y));
            const double2 nv = make_double2(fma(sn, u.y, c * v.x), fma(-sn, u.x, c * v.y));
            u = nu; v = nv;
        }
        if (op.flags & 2u) {                                  // Rx(theta): phase (c - i lambda sn)
            const double c = rx[k][j].x, sn = lam * rx[k][j].y;
            ph = make_double2(fma(sn, ph.y, c * ph.x), fma(-sn, ph.x, c * ph.y));
        }
        if ((w >> 28) & 1u) v = make_double2(-v.x, -v.y);
        if ((w >> 27) & 1u) { const double2 t = u; u = v; v = t; }
        if ((ancz >> op.pad) & 1u) v = make_double2(-v.x, -v.y);
    }
    const double2 a0 = dg_cmul(ph, u), a1 = dg_cmul(ph, v);
    const int e = H.tbase[k] + idx;
    tab[e] = a0;
    tab[e + kDgTabEntries] = a1;
    tab[e + 2 * kDgTabEntries] = make_double2(fma(a0.x, a0.x, a0.y * a0.y), fma(a1.x, a1.x, a1.y * a1.y));
}
// radix-16 pass with a butterfly [[a, b], [b, -a]] per register bit (Ry(eps) H; a = b = 1 is dg_wht16)
__device__ __forceinline__ void dg_rot16(double2 (&p)[16], const double2 (&ab)[4]) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const double al = ab[b].x, be = ab[b].y;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i & (1 << b)) continue;
            const double2 u = p[i], v = p[i | (1 << b)];
            p[i] = make_double2(fma(al, u.x, be * v.x), fma(al, u.y, be * v.y));
            p[i | (1 << b)] = make_double2(fma(be, u.x, -al * v.x), fma(be, u.y, -al * v.y));
        }
    }
}
// plain rotation [[c, -s], [s, c]] per register bit
__device__ __forceinline__ void dg_ry16(double2 (&p)[16], const double2 (&cs)[4]) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const double c = cs[b].x, sn = cs[b].y;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i & (1 << b)) continue;
            const double2 u = p[i], v = p[i | (1 << b)];
            p[i] = make_double2(fma(c, u.x, -sn * v.x), fma(c, u.y, -sn * v.y));
            p[i | (1 << b)] = make_double2(fma(sn, u.x, c * v.x), fma(sn, u.y, c * v.y));
        }
    }
}

// warp all-reduce of four doubles at the price of ten shuffles (instead of twenty): two transposing butterfly stages leave
// one partial sum per lane, three plain stages finish it, four broadcasts hand every lane all four totals
__device__ __forceinline__ void dg_reduce4(double (&v)[4], int lane) {
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0;
    const double k0 = (b4 ? v[2] : v[0]) + __shfl_xor_sync(0xffffffffu, b4 ? v[0] : v[2], 16);
    const double k1 = (b4 ? v[3] : v[1]) + __shfl_xor_sync(0xffffffffu, b4 ? v[1] : v[3], 16);
    double s = (b3 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, b3 ? k0 : k1, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __shfl_sync(0xffffffffu, s, 8 * i);      // lane 8 i holds the total of v[i]
}

// the four Measures of one half cycle from the real weights, then one complex update of the amplitudes.
// tk0[k] / tk1[k]: this lane's table entries of slot k for register bit k = 0 / 1 (A0 at +0, A1 at +16 kDgTabEntries, Q
// at +32 kDgTabEntries).  lane_neg: (-1)^lane_neg multiplies this lane's amplitudes; reg_neg bit j: amplitudes with
// register bit j set change sign; g: global phase i^g.  scale_out: fold 1/sqrt(final weight) into the update.
// The Measures are taken two at a time: the joint weights J[o_a][o_b] of a pair of slots need one reduction, and
// P(o_a) = J[o_a][0] + J[o_a][1], P(o_b | o_a) = J[o_a][o_b] / P(o_a) are the sequential Measures' own probabilities.
__device__ __forceinline__ void dg_half1(double2 (&p)[16], const unsigned char *(&tk0)[4], const unsigned char *(&tk1)[4],
                                         const uint8_t (&aj)[4], bool forced, uint32_t fo, double uu, int u_lane0,
                                         uint32_t lane_neg, uint32_t reg_neg, uint32_t g, bool scale_out, uint32_t xs4,
                                         uint32_t &a_out, bool &bad, double &pfin, double *pp, int lane) {
    double w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = fma(p[i].x, p[i].x, p[i].y * p[i].y);
    uint32_t outs = 0;
    double csel[4][2];            // |A_out|^2 of the selected entries, per slot and register bit
    // slots (KA, KB) = (0, 1) / (2, 3); m[b_a + 2 b_b]: this lane's weight per value of the two register bits
    auto measure_pair = [&](int KA, const double (&m)[4]) {
        const int KB = KA + 1;
        const double2 qa0 = *reinterpret_cast<const double2 *>(tk0[KA] + 32 * kDgTabEntries);
        const double2 qa1 = *reinterpret_cast<const double2 *>(tk1[KA] + 32 * kDgTabEntries);
        const double2 qb0 = *reinterpret_cast<const double2 *>(tk0[KB] + 32 * kDgTabEntries);
        const double2 qb1 = *reinterpret_cast<const double2 *>(tk1[KB] + 32 * kDgTabEntries);
        // t[b_b][o_a], then J[o_a + 2 o_b]
        const double t00 = fma(m[0], qa0.x, m[1] * qa1.x), t01 = fma(m[0], qa0.y, m[1] * qa1.y);
        const double t10 = fma(m[2], qa0.x, m[3] * qa1.x), t11 = fma(m[2], qa0.y, m[3] * qa1.y);
        double J[4] = {fma(t00, qb0.x, t10 * qb1.x), fma(t01, qb0.x, t11 * qb1.x),
                       fma(t00, qb0.y, t10 * qb1.y), fma(t01, qb0.y, t11 * qb1.y)};
        dg_reduce4(J, lane);
        // J is indexed by table COLUMN (A0 / A1); a slot with an odd number of X events on its ancilla (xs) reports the
        // other column's outcome: X_a commutes with every gate of the slot, so it acts as if right before the Measure
        const bool xa = (xs4 >> (4 * KA)) & 1u, xb = (xs4 >> (4 * KB)) & 1u;
        const double c0 = J[0] + J[2], c1 = J[1] + J[3];                     // weight of column 0 / 1 of slot KA (unnormalised)
        uint32_t oa, ob;
        if (forced) {
            oa = (fo >> aj[KA]) & 1u;
            ob = (fo >> aj[KB]) & 1u;
        } else {
            // Measure: P(1) against one uniform.  Every lane holds the same weights; the lane that drew the uniform of
            // this Measure decides, the others read its bit from the ballot
            oa = (__ballot_sync(0xffffffffu, uu * (c0 + c1) < (xa ? c0 : c1)) >> (u_lane0 + KA)) & 1u;
        }
        const uint32_t ca = oa ^ (xa ? 1u : 0u);                              // the column that outcome selects
        const double d0 = ca ? J[1] : J[0], d1 = ca ? J[3] : J[2];            // joint weight with column 0 / 1 of slot KB
        if (!forced) ob = (__ballot_sync(0xffffffffu, uu * (d0 + d1) < (xb ? d0 : d1)) >> (u_lane0 + KB)) & 1u;
        const uint32_t cb = ob ^ (xb ? 1u : 0u);
        const double wsel = cb ? d1 : d0;                                     // weight of the two outcomes taken
        if (forced && ((ca ? c1 : c0) <= 0.0 || wsel <= 0.0)) bad = true;
        if (pp) {                                                             // (P(0), P(1)) per Measure; every lane holds the same pairs
            pp[2 * KA] = xa ? c1 : c0; pp[2 * KA + 1] = xa ? c0 : c1;
            pp[2 * KB] = xb ? d1 : d0; pp[2 * KB + 1] = xb ? d0 : d1;
        }
        csel[KA][0] = ca ? qa0.y : qa0.x;
        csel[KA][1] = ca ? qa1.y : qa1.x;
        csel[KB][0] = cb ? qb0.y : qb0.x;
        csel[KB][1] = cb ? qb1.y : qb1.x;
        outs |= (ca << KA) | (cb << KB);
        a_out |= (oa << aj[KA]) | (ob << aj[KB]);
        pfin = wsel;
    };
    {
        double m[4];                                                          // sum over register bits 2, 3
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = (w[j] + w[j + 4]) + (w[j + 8] + w[j + 12]);
        measure_pair(0, m);
    }
    {
        double c01[4], m[4];                                                  // weight of (b2, b3) given the first two outcomes
#pragma unroll
        for (int j = 0; j < 4; ++j) c01[j] = csel[0][j & 1] * csel[1][j >> 1];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            m[q] = fma(w[4 * q], c01[0], fma(w[4 * q + 1], c01[1], fma(w[4 * q + 2], c01[2], w[4 * q + 3] * c01[3])));
        measure_pair(2, m);
    }
    // the selected complex factors, with the shot's signs and phase
    double2 f[4][2];
#pragma unroll
    for (int K = 0; K < 4; ++K) {
        const uint32_t sel = ((outs >> K) & 1u) ? 16u * kDgTabEntries : 0u;
        f[K][0] = *reinterpret_cast<const double2 *>(tk0[K] + sel);
        f[K][1] = dg_neg_if(*reinterpret_cast<const double2 *>(tk1[K] + sel), (reg_neg >> K) & 1u);
    }
    f[0][0] = dg_neg_if(f[0][0], lane_neg);
    f[0][1] = dg_neg_if(f[0][1], lane_neg);
    f[3][0] = dg_times_i_pow(f[3][0], g);
    f[3][1] = dg_times_i_pow(f[3][1], g);
    if (scale_out) {
        const double inv = pfin > 0.0 ? 1.0 / sqrt(pfin) : 0.0;
        f[3][0] = make_double2(f[3][0].x * inv, f[3][0].y * inv);
        f[3][1] = make_double2(f[3][1].x * inv, f[3][1].y * inv);
    }
    double2 F01[4], F23[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        F01[j] = dg_cmul(f[0][j & 1], f[1][j >> 1]);
        F23[j] = dg_cmul(f[2][j & 1], f[3][j >> 1]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) p[i] = dg_cmul(dg_cmul(p[i], F01[i & 3]), F23[i >> 2]);
}

template <bool COH>
__global__ void __launch_bounds__(kDgThreads, 1)
k_cycle1(const __grid_constant__ DiagProg1 P, const __grid_constant__ DiagCoh C, const __grid_constant__ MeasArgs ma,
         double2 *__restrict__ state,
         size_t stride, const uint32_t *__restrict__ ev, const uint32_t *__restrict__ frm, const uint32_t *__restrict__ act_list,
         const uint32_t *__restrict__ n_act_ptr, ShotAux *__restrict__ aux, const uint8_t *__restrict__ forced,
         uint32_t *__restrict__ meas_out, uint32_t *__restrict__ dsamp, unsigned long long *__restrict__ counters,
         int init_basis, const double2 *__restrict__ alt_src, double *__restrict__ pp_out, double *__restrict__ cdf_out)
{
    // alt_src != nullptr: an entry of act_list is shot | own << 23 | leaf << 24 and the input block of the shot is
    // alt_src[leaf] unless `own` is set (the shot ran its own preparation cycle: a leak flag was set)
    // (the memoised outcome of the noiseless preparation cycle, s17.cu: build_prep_cache); pp_out != nullptr: record
    // the (P(0), P(1)) pair of every Measure, cdf_out != nullptr: the read-out's cumulative weights (used once, to
    // build those caches)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    int lane = threadIdx.x & 31;
#ifndef S17_DG_NO_PIN
    asm volatile("" : "+r"(lane));                // (otherwise re-read from the special register, with its latency, inside the cycle)
#endif
    // [gap][block buffers: warp x 2 x 8 KiB, 8 KiB aligned in the window][per-warp small blocks][per-CTA tables, if not in the gap]
    const uint32_t gap = (0u - smem_u32(smem_raw)) & (uint32_t)(kRunBytes - 1);
    uint32_t buf_off = gap + (uint32_t)warp * 2 * kRunBytes;
    uint32_t wb_off = gap + (uint32_t)kDgWarps * 2 * kRunBytes + (uint32_t)warp * kDgSmallBytes;
    uint32_t clean_off = gap >= (uint32_t)kDgCleanBytes ? 0u : gap + (uint32_t)kDgSmem;
#ifndef S17_DG_NO_PIN
    // keep the per-warp offsets in registers: under register pressure the compiler otherwise rebuilds them from the thread
    // index at every use (a dozen instructions each, several times per cycle)
    asm volatile("" : "+r"(buf_off), "+r"(wb_off), "+r"(clean_off));
#endif
    unsigned char *buf = smem_raw + buf_off;                                         // [2][8 KiB]
    unsigned char *wb = smem_raw + wb_off;
    unsigned char *tab = wb;                                                         // A0 | A1 | Q
    unsigned char *meta = wb + kDgTabBytes;                                          // [2][kDgMetaBytes]
    uint64_t *full = reinterpret_cast<uint64_t *>(meta + 2 * kDgMetaBytes);          // [2]
    uint64_t *mfull = full + 2;                                                      // [2]
    // per CTA: the tables of a half cycle without events (most slots of most shots), built once
    unsigned char *clean = smem_raw + clean_off;                                     // [2][kDgTabBytes]
    const uint32_t buf32 = smem_u32(buf);
    const uint32_t meta32_0 = smem_u32(meta);
    uint32_t *zero_ev = reinterpret_cast<uint32_t *>(clean + 2 * kDgTabBytes);       // [S17_EV_WORDS]: the event words of a shot without events
    const uint32_t zero32 = smem_u32(zero_ev);
    if (warp == 0) {
        zero_ev[lane] = 0u;
        __syncwarp();
        if (COH) {
#pragma unroll 1
            for (int h = 0; h < 2; ++h)
                for (int e = lane; e < (int)P.h[h].n_entries; e += 32) {
                    const int k = (e >= P.h[h].tbase[1]) + (e >= P.h[h].tbase[2]) + (e >= P.h[h].tbase[3]);
                    dg_build_entry_coh(P.h[h], C.rxx[h], C.rx[h], zero_ev, reinterpret_cast<double2 *>(clean + h * kDgTabBytes), k,
                                       e - P.h[h].tbase[k], 0u);
                }
        } else {
            dg_build_tables(P.h[0], zero_ev, reinterpret_cast<double2 *>(clean), lane, 0u, 0u, 0u);
            dg_build_tables(P.h[1], zero_ev, reinterpret_cast<double2 *>(clean + kDgTabBytes), lane, 0u, 0u, 0u);
        }
    }
    __syncthreads();

    // act_list == nullptr: every shot of the batch takes part and none has an event or a set leak flag (a preparation
    // cycle of a model without leakage): no list, no event / frame words to load
    const bool all_quiet = act_list == nullptr;
    const uint32_t n_items = all_quiet ? (uint32_t)ma.n_shots : *n_act_ptr;
    const uint32_t n_warps = gridDim.x * kDgWarps, gw = blockIdx.x * kDgWarps + warp;
    const uint32_t chunk = (n_items + n_warps - 1) / n_warps;
    const uint32_t first = gw * chunk, last = min(n_items, first + chunk);
    if (first >= last) return;                    // no CTA-wide barrier below: warps are independent

    if (lane == 0) {
        mbar_init(&full[0], 1); mbar_init(&full[1], 1);
        mbar_init(&mfull[0], 1); mbar_init(&mfull[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // which operation of which slot this lane checks for events (lanes 0-15: slot = lane / 4, op = lane % 4)
    uint32_t chk_word[2];                       // event word | index within the half << 8
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int ck = (lane >> 2) & 3, cj = lane & 3;
        chk_word[h] = (lane < 16 && cj < P.h[h].n_ops[ck])
                          ? ((uint32_t)P.h[h].ops[ck][cj].ev_word | ((uint32_t)P.h[h].ops[ck][cj].pad << 8)) : 0xFFu;
    }

    // lane parts (shot independent): label offsets of the two maps, linear and through the exchange swizzle, and this
    // lane's first table entry per slot (one byte per slot)
    uint32_t laneA = 0, laneB = 0;
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        if ((lane >> m) & 1) { laneA |= 1u << P.mapA.lane[m]; laneB |= 1u << P.mapB.lane[m]; }
    }
    auto swz = [&](uint32_t label) {
        uint32_t c = 0;
#pragma unroll
        for (int b = 3; b < 9; ++b)
            if ((label >> b) & 1u) c ^= P.gvec[b];
        return label ^ c;
    };
    const uint32_t linA = laneA << 4, swzA = swz(laneA) << 4, swzB = swz(laneB) << 4;
    // this lane's three access bases in buffer 0 (linear, exchange through map A / map B); buffer 1 is 8 KiB further
    uint32_t la0 = buf32 | linA, xa0 = buf32 | swzA, xc0 = buf32 | swzB;
#ifndef S17_DG_NO_PIN
    asm volatile("" : "+r"(la0), "+r"(xa0), "+r"(xc0));
#endif
    const double sh_sign = ((lane >> P.sh_lane_bit) & 1) ? -1.0 : 1.0;
    const int sh_mask = 1 << P.sh_lane_bit;
    // lane part of this lane's table index per slot (4 bits each; the register-fed index bit stays 0)
    uint32_t lidx[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        lidx[h] = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int lb = P.h[h].lane_bit[k][j];
                if (lb != 255) lidx[h] |= (uint32_t)((lane >> lb) & 1) << (4 * k + j);
            }
        }
    }

    auto prefetch = [&](uint32_t entry, int sb) {
        // plain shot index, or (preparation cache in use) shot | own-state flag << 23 | leaf << 24
        const uint32_t shot_ = alt_src ? (entry & 0x7FFFFFu) : entry;
        const bool cached = alt_src && !(entry & 0x800000u);
        unsigned char *m = meta + sb * kDgMetaBytes;
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the bulk store staged in this buffer has left it
        fence_proxy_async_smem();                 // the buffers were last touched through the generic proxy
        if (!all_quiet) {
            mbar_expect_tx(&mfull[sb], kDgMetaBytes);
            tma_load_1d(m, ev + (size_t)shot_ * S17_EV_WORDS, kEvBytes, &mfull[sb]);
            tma_load_1d(m + kEvBytes, frm + (size_t)shot_ * 4, kDgFrmBytes, &mfull[sb]);
        }
        if (init_basis < 0) {
            mbar_expect_tx(&full[sb], kRunBytes);
            tma_load_1d(buf + sb * kRunBytes, cached ? alt_src + (size_t)(entry >> 24) * kRun : state + (size_t)shot_ * stride,
                        kRunBytes, &full[sb]);
        }
    };
    auto lane_stage = [&](double2 (&q)[16]) {
#ifdef S17_DG_LS_GROUP
#pragma unroll
        for (int g0 = 0; g0 < 16; g0 += S17_DG_LS_GROUP) {
            double2 o[S17_DG_LS_GROUP];
#pragma unroll
            for (int i = 0; i < S17_DG_LS_GROUP; ++i)
                o[i] = make_double2(__shfl_xor_sync(0xffffffffu, q[g0 + i].x, sh_mask), __shfl_xor_sync(0xffffffffu, q[g0 + i].y, sh_mask));
#pragma unroll
            for (int i = 0; i < S17_DG_LS_GROUP; ++i)
                q[g0 + i] = make_double2(fma(sh_sign, q[g0 + i].x, o[i].x), fma(sh_sign, q[g0 + i].y, o[i].y));
        }
#else
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const double ox = __shfl_xor_sync(0xffffffffu, q[i].x, sh_mask), oy = __shfl_xor_sync(0xffffffffu, q[i].y, sh_mask);
            q[i] = make_double2(fma(sh_sign, q[i].x, ox), fma(sh_sign, q[i].y, oy));
        }
#endif
    };

    // list entries are read two shots ahead: the entry of shot j + 1 is needed at the top of shot j (its prefetch), a
    // global load issued there would be waited for at once
    uint32_t entry = all_quiet ? first : act_list[first];
    uint32_t entry1 = (first + 1 < last) ? (all_quiet ? first + 1 : act_list[first + 1]) : 0u;
    if (lane == 0) prefetch(entry, 0);
#ifdef S17_DG_STAGGER
    // the three warps of a scheduler start a third of a shot apart, so that they are not all in the same phase
    // (FP64 butterflies / shuffles / exchange) at the same time
    __nanosleep((uint32_t)(warp >> 2) * S17_DG_STAGGER);
#endif

    const uint32_t nout = 8u * (uint32_t)ma.stage;
    const bool is_forced = ma.forced != 0;
    uint32_t j = 0;
    for (uint32_t item = first; item < last; ++item, ++j) {
        const int sb = (int)(j & 1u);
        const uint32_t shot = alt_src ? (entry & 0x7FFFFFu) : entry;
        uint32_t entry2 = 0;
        if (item + 2 < last) entry2 = all_quiet ? item + 2 : act_list[item + 2];
        if (item + 1 < last && lane == 0) prefetch(entry1, sb ^ 1);
        // this shot's event words [S17_EV_WORDS] and frame words [4] by shared-window address (zeros for a cycle without a list)
        const uint32_t meta32 = all_quiet ? zero32 : meta32_0 + (uint32_t)sb * kDgMetaBytes;
        unsigned char *xb = buf + sb * kRunBytes;
        const uint32_t sboff = (uint32_t)sb * kRunBytes;                            // buffers are aligned to 8 KiB: the lane parts stay below
        const uint32_t la32 = la0 + sboff;

        double2 p[16];
        if (init_basis < 0) {
            mbar_wait(&full[sb], (j >> 1) & 1u);
#pragma unroll
            for (int i = 0; i < 16; ++i) p[i] = dg_lds(la32 ^ P.mapA.lin16[i]);
        } else {
            // first cycle of a shot: the (unnormalised) Walsh-Hadamard transform of the basis state |b> is the sign
            // pattern (-1)^{y.b} -- exactly what the transform below would compute from it -- written in map B
            // (bits 9-24 of init_basis: the parity of the basis state over the label part of register i, from the host)
            uint32_t lb = laneB;
            asm volatile("" : "+r"(lb));          // keeps the pattern inside this branch (the compiler otherwise builds it for every cycle and discards it)
            const uint32_t sg = ((uint32_t)init_basis >> 9) ^ ((__popc(lb & (uint32_t)init_basis & 511u) & 1) ? 0xFFFFu : 0u);
#pragma unroll
            for (int i = 0; i < 16; ++i)
                p[i] = make_double2(__hiloint2double((int)(0x3ff00000u | (((sg >> i) & 1u) << 31)), 0), 0.0);
        }
        if (!all_quiet) mbar_wait(&mfull[sb], (j >> 1) & 1u);
        __syncwarp();                             // every lane holds its amplitudes: the buffer becomes exchange scratch

        // Two steps, one per half cycle: [Walsh-Hadamard transform] [tables + four ancilla slots].  Half 0 works in the
        // Hadamard frame, half 1 in the computational frame (make_diag1 checks that), so the transform before half 0 goes
        // into the frame and the one before half 1 comes back.  The loop is not unrolled: the transform and the slot code
        // exist once in the instruction stream (the working set of twelve warps at different points of it has to stay
        // inside the instruction cache).
        //
        // ma.with_prep: two cycles per shot in this launch -- first the noiseless preparation cycle from the basis state
        // (no events, the streams and the outcome slots of stage 0), then the launch's own cycle on the block in registers.
        uint32_t perm = 0, prep_bits = 0, fo = 0;
        double uu = 0.0;
        const int n_stg = ma.with_prep ? 2 : 1;
#pragma unroll 1
        for (int stg = 0; stg < n_stg; ++stg) {
        const bool prep_stg = ma.with_prep && stg == 0;
        const bool quiet = all_quiet || prep_stg;
        const bool from_basis = init_basis >= 0 && stg == 0;
        if (is_forced) {
            const uint32_t b = lane < 8 ? (uint32_t)forced[(size_t)shot * S17_MAX_OUTCOMES + (prep_stg ? 0u : nout) + lane] : 0u;
            fo = __ballot_sync(0xffffffffu, (b & 1u) != 0u) & 0xFFu;
        } else if (lane < 9) {
            // lanes 0-7: the uniforms of the eight Measure gates (stream per half, counter block = slot); lane 8: the
            // read-out's uniform (k_readout's stream)
            const uint32_t strm = prep_stg ? ma.stream - 4u * (uint32_t)ma.stage : ma.stream;      // stage s draws from 16 + 4 s (+ 1)
            Philox rng(ma.seed, ma.first_shot + shot, lane == 8 ? ma.data_stream : strm + ((lane >> 2) ? 0u : 1u));
            rng.c3 = lane == 8 ? 0u : (uint32_t)(lane & 3);
            uu = rng.next();
        }
        uint32_t carry = 0;
        double pfin = 1.0;
        perm = 0;
#pragma unroll kDgStUnroll
        for (int h = 0; h < 2; ++h) {
            const DiagHalf &H = P.h[h];
            const uint32_t fw = quiet ? 0u : dg_lds32(meta32 + (uint32_t)kEvBytes + 4u * (uint32_t)h);
            if (h == 1 || !from_basis) {                          // (a cycle that starts from a basis state is in the frame already)
                // Walsh-Hadamard transform of the data register: the register bits of the current map, the exchange to the
                // other map, its register bits; the one remaining label bit is a lane bit of both maps (exchange load)
                const int dir = h;                                // 0: map A -> B (into the frame), 1: map B -> A
                const uint32_t xa32 = xa0 + sboff, xc32 = xc0 + sboff;
#pragma unroll kDgHpUnroll
                for (int half_pass = 0; half_pass < 2; ++half_pass) {
                    if (COH && dir) {
                        // back to the computational frame, with Ry(eps_a) of every data qubit merged into its butterfly
                        if (half_pass) {
                            if (!kDgXlane) {
                                const double al = ((lane >> P.sh_lane_bit) & 1) ? -C.preSh.x : C.preSh.x, be = C.preSh.y;
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    const double ox = __shfl_xor_sync(0xffffffffu, p[i].x, sh_mask), oy = __shfl_xor_sync(0xffffffffu, p[i].y, sh_mask);
                                    p[i] = make_double2(fma(al, p[i].x, be * ox), fma(al, p[i].y, be * oy));
                                }
                            }
                            dg_rot16(p, C.preA);
                        } else {
                            dg_rot16(p, C.preB);
                        }
                    } else {
                    if (!kDgXlane && half_pass == dir) lane_stage(p);          // in map A: before the first pass / before the last one
                    dg_wht16(p);
                    }
                    if (half_pass == 0) {
                        // the exchange, once per direction: the register parts are constant-bank operands of the XOR.
                        // The one label bit that is a lane bit under both maps is transformed on the way in (kDgXlane): a
                        // lane also loads the sixteen amplitudes of the lane that differs in that bit (the same addresses,
                        // permuted over the warp: as conflict-free as its own) -- no shuffles, no 64-bit repacking
                        // (coherent form, on the way back: the butterfly of that bit carries Ry(eps_a), [[a, b], [b, -a]])
                        const bool xl = kDgXlane;
                        __syncwarp();
                        if (dir) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) dg_sts(xc32 ^ P.mapB.swz16[i], p[i]);
                            __syncwarp();
                            if (xl) {
                                const double al = COH ? (((lane >> P.sh_lane_bit) & 1) ? -C.preSh.x : C.preSh.x) : sh_sign;
                                const double be = COH ? C.preSh.y : 1.0;
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    const double2 u = dg_lds(xa32 ^ P.mapA.swz16[i]), v = dg_lds(xa32 ^ P.mapA.swzp16[i]);
                                    p[i] = COH ? make_double2(fma(al, u.x, be * v.x), fma(al, u.y, be * v.y))
                                               : make_double2(fma(al, u.x, v.x), fma(al, u.y, v.y));
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 16; ++i) p[i] = dg_lds(xa32 ^ P.mapA.swz16[i]);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) dg_sts(xa32 ^ P.mapA.swz16[i], p[i]);
                            __syncwarp();
                            if (xl) {
                                const double sgb = ((lane >> P.sh_lane_bit_b) & 1) ? -1.0 : 1.0;
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    const double2 u = dg_lds(xc32 ^ P.mapB.swz16[i]), v = dg_lds(xc32 ^ P.mapB.swzp16[i]);
                                    p[i] = make_double2(fma(sgb, u.x, v.x), fma(sgb, u.y, v.y));
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 16; ++i) p[i] = dg_lds(xc32 ^ P.mapB.swz16[i]);
                            }
                        }
                    }
                }
            }
            __syncwarp();
            // This shot's events in this half cycle.  Data flips upstream of an operation (bit 31 of its event word) are
            // an XOR on the table index of its slot; signs (sign mask, flips of the preceding Hadamard-frame half) and
            // the global phase multiply the selected factors; only slots whose own operations carry an ancilla Pauli, a
            // leak skip or an idle Z get private entries.
            const uint32_t zmask = ((fw >> 9) ^ carry) & 511u, g = (fw >> 18) & 3u;
            const uint32_t cw = h ? chk_word[1] : chk_word[0];
            const uint32_t wv = (!quiet && cw != 0xFFu) ? dg_lds32(meta32 + 4u * (cw & 0xFFu)) : 0u;
            // (bit 24: leak skip, bit 28: net Z on the ancilla, frame bits 20+: idle Z after the operation; a net X on the
            //  ancilla alone -- bit 27, e.g. an XX error -- commutes with every gate of its slot and only swaps the two
            //  outcomes of that slot's Measure: xs, no private entries)
            // one nibble per slot (lanes 4 k .. 4 k + 3 look at the operations of slot k); dirty4 / xs4: bit 4 k = slot k
            uint32_t flipb = 0, dirty4 = 0, xs4 = 0;
            if (!quiet) {                           // (a cycle without a list, a preparation cycle: no events, nothing to look at)
                const uint32_t evd = __ballot_sync(0xffffffffu, cw != 0xFFu && ((wv & 0x11000000u) != 0u || ((fw >> (20 + (cw >> 8))) & 1u)));
                flipb = __ballot_sync(0xffffffffu, (wv >> 31) != 0u);
                const uint32_t xev = __ballot_sync(0xffffffffu, ((wv >> 27) & 1u) != 0u);
                uint32_t t = evd | (evd >> 1);
                dirty4 = (t | (t >> 2)) & 0x1111u;                                   // any event bit in the nibble
                flipb &= ~(dirty4 * 15u);                                            // flips and X's are part of private entries
                t = xev ^ (xev >> 1);
                xs4 = (t ^ (t >> 2)) & 0x1111u & ~dirty4;                            // parity of the nibble
            }
            if (dirty4) {
                const uint32_t *evw = reinterpret_cast<const uint32_t *>(meta + sb * kDgMetaBytes);    // (dirty: not all_quiet)
                uint32_t d = dirty4;
                while (d) {                                    // two slots per pass: lanes 0-15 / 16-31
                    const int k0 = (__ffs(d) - 1) >> 2;
                    d &= d - 1;
                    const int k1 = d ? (__ffs(d) - 1) >> 2 : -1;
                    d &= d - 1;
                    const int k = lane < 16 ? k0 : k1, idx = lane & 15;
                    if (k >= 0 && idx < (1 << H.n_ops[k])) {
                        if (COH) dg_build_entry_coh(H, C.rxx[h], C.rx[h], evw, reinterpret_cast<double2 *>(tab), k, idx, fw >> 20);
                        else dg_build_slot(P, h, evw, reinterpret_cast<double2 *>(tab), k, idx, fw >> 20);
                    }
                }
                __syncwarp();
            }
            const unsigned char *clean_h = clean + h * kDgTabBytes;
            const uint32_t li = (h ? lidx[1] : lidx[0]) ^ flipb;
            const unsigned char *tk0[4], *tk1[4];
            uint32_t reg_neg = 0;
            const DiagMap &M = h == 0 ? P.mapB : P.mapA;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t rmask = P.off1[h][k] >> 4;                          // index bit fed by register bit k (or 0)
                const uint32_t ix = (li >> (4 * k)) & 15u;
                const unsigned char *base = (((dirty4 >> (4 * k)) & 1u) ? tab : clean_h) + ((uint32_t)H.tbase[k] << 4);
                tk0[k] = base + (ix << 4);                                         // (ix & rmask: that operation sees its bit flipped)
                tk1[k] = base + ((ix ^ rmask) << 4);
                reg_neg |= ((zmask >> M.reg[k]) & 1u) << k;
            }
            const uint32_t lane_neg = __popc((h == 0 ? laneB : laneA) & zmask) & 1u;
            uint32_t a_out = 0;
            bool bad = false;
            dg_half1(p, tk0, tk1, H.aj, is_forced, fo, uu, 4 * h, lane_neg, reg_neg, g, h == 1, xs4, a_out, bad, pfin,
                     pp_out ? pp_out + ((size_t)shot * 8 + 4 * h) * 2 : nullptr, lane);
            if (h == 0) carry = fw & 511u;                        // data flips of the Hadamard-frame half are signs afterwards: half 1's sign mask
            else perm = fw & 511u;                                // data flips of the computational-frame half relabel the store
            // (the outcomes of a preparation cycle run ahead of this one: bits 16-24 of the same words, record_cycle)
            if (prep_stg) prep_bits |= a_out | (bad ? 0x100u : 0u);
            else if (lane == 0) meas_out[2 * shot + H.final_part] = a_out | (bad ? 0x100u : 0u) | (prep_bits << 16);
        }
        if (COH) {
            // Ry(eps_b) on every data qubit after its bracket: register bits in place, lane bits by shuffle
            dg_ry16(p, C.postR);
#pragma unroll 1
            for (int m = 0; m < 5; ++m) {
                const double c = C.postL[m].x, sn = ((lane >> m) & 1) ? C.postL[m].y : -C.postL[m].y;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const double ox = __shfl_xor_sync(0xffffffffu, p[i].x, 1 << m), oy = __shfl_xor_sync(0xffffffffu, p[i].y, 1 << m);
                    p[i] = make_double2(fma(sn, ox, c * p[i].x), fma(sn, oy, c * p[i].y));
                }
            }
        }
        }   // stage

        // All(Measure) | data of a read-out that may follow (CS:56): one basis state drawn from |psi|^2 -- one uniform
        // against the cumulative weights in (lane, register) order.  The decoder's X correction relabels it later.
        if (ma.sample_data) {
            double wsum[16], mine = 0.0;
#pragma unroll
            for (int i = 0; i < 16; ++i) { wsum[i] = fma(p[i].x, p[i].x, p[i].y * p[i].y); mine += wsum[i]; }
            double incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            if (cdf_out) {                        // kDgCdfDoubles per shot: 32 inclusive lane sums, then 16 weights per lane
                cdf_out[(size_t)shot * kDgCdfDoubles + lane] = incl;
#pragma unroll
                for (int i = 0; i < 16; ++i) cdf_out[(size_t)shot * kDgCdfDoubles + 32 + 16 * lane + i] = wsum[i];
            }
            const double target = __shfl_sync(0xffffffffu, uu, 8) * __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t hit = __ballot_sync(0xffffffffu, target < incl);
            const int sel = hit ? (__ffs(hit) - 1) : 31;
            // every lane walks its own sixteen weights, branch-free: pick = number of running sums <= target, starting
            // from the previous lane's total (<= target for the selected lane).  A state of weight zero repeats the
            // running sum before it, so it is counted past, never selected; only rounding at the very end can run off
            // the list, then the last state of positive weight is taken.
            double cum = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) cum = 0.0;
            int pick = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                cum += wsum[i];
                pick += (cum <= target) ? 1 : 0;
            }
            if (pick > 15) {                      // (rounding only: rare, and only the selected lane's value is used)
                pick = 0;
#pragma unroll
                for (int i = 0; i < 16; ++i) pick = wsum[i] > 0.0 ? i : pick;
            }
            uint32_t lab = (linA ^ P.mapA.lin16[pick]) >> 4;
            lab = __shfl_sync(0xffffffffu, lab, sel);
            if (lane == 0) dsamp[shot] = (lab ^ perm) & 511u;
        }

        // stage the (already renormalised) block in label order (data flips of the computational-frame half relabel it)
        // and hand it to one bulk store
        __syncwarp();
        {
            const uint32_t ls32 = la32 ^ (perm << 4);
#pragma unroll
            for (int i = 0; i < 16; ++i) dg_sts(ls32 ^ P.mapA.lin16[i], p[i]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            tma_store_1d(state + (size_t)shot * stride, xb, kRunBytes);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            ShotAux ax;
            ax.scale = 1.0;
            ax.anc_outcome = 0;
            ax.collapsed = 1;                     // the normalised block is stored at ancilla value 0
            aux[shot] = ax;
        }
        __syncwarp();                             // tables and the metadata slot are reused
        entry = entry1;
        entry1 = entry2;
    }
    if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        atomicAdd(&counters[4], (unsigned long long)(last - first) * (ma.with_prep ? 2ull : 1ull));     // C_SWEEPS (cycles)
        atomicAdd(&counters[8], (unsigned long long)(last - first) * ((init_basis < 0 && !alt_src) ? 2ull : 1ull));   // 8 KiB runs moved (HBM)
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Memoised preparation cycle.  logical_prep (CS:96-123) runs the NOISELESS cycle on |0...0>: without a leaked qubit its
// result depends only on the eight measurement outcomes.  build_prep_cache (s17.cu) runs k_cycle1 once per outcome
// pattern (256 forced shots) and keeps, per node of the outcome tree, the (P(0), P(1)) pair the kernel computed and, per
// leaf, the collapsed block.  k_prep then replays the kernel's own decisions per shot -- same Philox uniforms, same
// comparison, same doubles -- and the first noisy cycle reads its input block from the cache (L2 resident).
// ------------------------------------------------------------------------------------------------------------------
struct PrepTree {
    uint8_t aj[8];             // ancilla index of Measure step s (half s / 4, slot s % 4)
    uint8_t final_part[2];
    uint8_t pad[6];
};

__global__ void __launch_bounds__(256)
k_prep(const __grid_constant__ PrepTree T, const __grid_constant__ MeasArgs ma, const double *__restrict__ node_pp,
       const uint8_t *__restrict__ forced, const uint8_t *__restrict__ active, uint32_t *__restrict__ meas_out,
       uint8_t *__restrict__ leaf_out, ShotAux *__restrict__ aux, unsigned long long *__restrict__ counters,
       const uint32_t *__restrict__ leak /* nullptr, or: shots with a set leak flag run the real preparation cycle */,
       uint8_t *__restrict__ done)
{
    const int shot = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = shot < ma.n_shots && active[shot];
    if (live) {
        const bool own = leak && leak[shot] != 0;
        done[shot] = own ? 0 : 1;
        live = !own;
    }
    if (live) {
        uint32_t prefix = 0, r[2] = {0u, 0u};
        bool bad = false;
#pragma unroll 1
        for (int s = 0; s < 8; ++s) {
            // node (s, prefix): nodes of depth s start at 2^s - 1
            const double p0 = node_pp[2 * (((1u << s) - 1u) + prefix)], p1 = node_pp[2 * (((1u << s) - 1u) + prefix) + 1];
            uint32_t out;
            if (ma.forced) {
                out = forced[(size_t)shot * S17_MAX_OUTCOMES + T.aj[s]] & 1u;           // stage 0: the first eight outcomes
                if ((out ? p1 : p0) <= 0.0) bad = true;
            } else {
                Philox rng(ma.seed, ma.first_shot + shot, ma.stream + ((s >> 2) ? 0u : 1u));
                rng.c3 = (uint32_t)(s & 3);
                out = (rng.next() * (p0 + p1) < p1) ? 1u : 0u;
            }
            prefix |= out << s;
            r[s >> 2] |= out << T.aj[s];
        }
        meas_out[2 * shot + T.final_part[0]] = r[0] | (bad ? 0x100u : 0u);
        meas_out[2 * shot + T.final_part[1]] = r[1] | (bad ? 0x100u : 0u);
        leaf_out[shot] = (uint8_t)prefix;
        ShotAux ax;
        ax.scale = 1.0;
        ax.anc_outcome = 0;
        ax.collapsed = 1;
        aux[shot] = ax;
    }
    const uint32_t nl = __popc(__ballot_sync(0xffffffffu, live));
    if ((threadIdx.x & 31) == 0 && nl) atomicAdd(&counters[4], (unsigned long long)nl);                  // C_SWEEPS
}

// Memoised event-free first cycle.  A first noisy cycle in which no error slot fires and no qubit is leaked is the
// noiseless cycle applied to the preparation leaf the shot is in: it repeats the leaf's syndrome with certainty and
// returns the same block.  build_prep_cache runs k_cycle1 once per leaf for this case too, checks that the kernel
// itself finds every other outcome to have probability exactly zero, and keeps the block it stores and the read-out
// weights it computes.  One warp per such shot: copy the block (optional), replay the read-out draw bit for bit.
__global__ void __launch_bounds__(128)
k_quiet(const __grid_constant__ PrepTree T, const __grid_constant__ MeasArgs ma, const __grid_constant__ DiagMap mapA,
        const uint8_t *__restrict__ active, const uint8_t *__restrict__ quiet,
        const uint8_t *__restrict__ quiet_ok, const uint8_t *__restrict__ leaf_of, const double2 *__restrict__ blocks,
        const double *__restrict__ cdf /* per leaf: 32 inclusive lane sums, then per lane its 16 running sums */,
        const uint8_t *__restrict__ last_pos /* per leaf and lane: last state of positive weight */,
        const uint8_t *__restrict__ forced, double2 *__restrict__ state, size_t stride,
        uint32_t *__restrict__ meas_out, uint32_t *__restrict__ dsamp, unsigned long long *__restrict__ counters, int copy_block)
{
    // one thread per shot (copy_block: one warp per shot, lane 0 does the scalar part)
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int shot = copy_block ? (gid >> 5) : gid, lane = copy_block ? (threadIdx.x & 31) : 0;
    const bool live = shot < ma.n_shots && active[shot] && quiet[shot] && quiet_ok[leaf_of[shot]];
    if (live) {
        const uint32_t leaf = leaf_of[shot];
        if (copy_block) {
            const double2 *src = blocks + (size_t)leaf * kRun;
            double2 *dst = state + (size_t)shot * stride;
#pragma unroll
            for (int i = 0; i < 16; ++i) dst[lane + 32 * i] = src[lane + 32 * i];
        }
        if (lane == 0) {
            if (ma.forced) {
                // replay: the recorded outcomes of this cycle must be the leaf's (anything else has probability zero)
                uint32_t bad = 0;
                for (int s = 0; s < 8; ++s)
                    bad |= ((uint32_t)forced[(size_t)shot * S17_MAX_OUTCOMES + 8 + T.aj[s]] & 1u) ^ ((leaf >> s) & 1u);
                if (bad) { meas_out[2 * shot] |= 0x100u; meas_out[2 * shot + 1] |= 0x100u; }
            }
            if (ma.sample_data) {
                // the draw of k_cycle1 from the weights it computed for this block -- the same comparisons on the same
                // doubles: the first lane whose inclusive sum exceeds the target, then the number of that lane's running
                // sums that do not
                const double *c = cdf + (size_t)leaf * kDgCdfDoubles;
                Philox rng(ma.seed, ma.first_shot + shot, ma.data_stream);
                const double target = rng.next() * c[31];
                int sel = 31;                              // (a scan in another order may round a sum below its neighbour:
#pragma unroll 8
                for (int l = 30; l >= 0; --l)              //  take the first lane exactly as the kernel's ballot does)
                    if (target < c[l]) sel = l;
                int pick = 0;
#pragma unroll
                for (int i = 0; i < 16; ++i) pick += (c[32 + 16 * sel + i] <= target) ? 1 : 0;
                pick = pick > 15 ? (int)last_pos[leaf * 32 + sel] : pick;
                const uint32_t lab = (uint32_t)lane_label(mapA, sel) | (mapA.lin16[pick] >> 4);
                dsamp[shot] = lab & 511u;                  // a computational-frame half without events relabels nothing
            }
        }
    }
    const uint32_t nl = __popc(__ballot_sync(0xffffffffu, live && lane == 0));
    if ((threadIdx.x & 31) == 0 && nl) atomicAdd(&counters[4], (unsigned long long)nl);                  // C_SWEEPS
}

}  // namespace s17
