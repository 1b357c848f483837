// s17_layer2.cuh -- persistent, warp-specialised gate sweep (the production k_layer).
//
// One CTA per SM walks a contiguous range of (shot, tile) items.  Three 64 KiB shared-memory tile
// buffers form a ring:
//     loader warp  : TMA bulk loads of tile j+1, j+2 (8 runs x 8 KiB + the shot's 128 B of event words)
//     2 x 4 math warps : tiles j, j+1 -- two entangling operations per pass on 16 register-resident
//                    amplitudes; the two groups run out of phase (one in its LDS/STS phase, one in FP64)
//     storer warp  : TMA bulk stores of tile j-1, then hands the buffer back to the loader
// so HBM reads, FP64 math and HBM writes of different tiles overlap inside every SM.
// Algorithmic traffic per (shot, sweep) is unchanged: 2 MiB read + 2 MiB write.
#pragma once
#include <cuda.h>
#include "s17_kernels.cuh"

#ifndef S17_MATH_THREADS
#define S17_MATH_THREADS 256
#endif

namespace s17 {

constexpr int kMathThreads = S17_MATH_THREADS;
#ifndef S17_MATH_GROUPS
#define S17_MATH_GROUPS 2
#endif
constexpr int kGroups = S17_MATH_GROUPS;            // independent math groups, each on its own tile
constexpr int kGroupThreads = kMathThreads / kGroups;
constexpr int kL2Threads = kMathThreads + 64;      // + loader warp + storer warp
constexpr int kRing = 3;
constexpr int kEvBytes = S17_EV_WORDS * 4;
constexpr int kAuxBytes = 16;                       // sizeof(ShotAux)
constexpr size_t kL2Smem = (size_t)kRing * kTile * 16 + kRing * kEvBytes + kRing * kAuxBytes + 3 * kRing * 8 + 64;

struct OpC {                 // per-operation constants precomputed on the host
    uint8_t lb0, lb1;        // tile-local bit of the data / ancilla qubit
    uint8_t shape;           // bit0 Ry(+) before, bit1 Rx(-s) after, bit2 Ry(-) after   (standard ops)
    uint8_t generic;         // 1: interpret the uop list (coherent rotations, H, ...)
    int8_t s;                // sign of the Rxx angle
    uint8_t ev_word;
    uint8_t n_fixed, n_skip; // number of 1/sqrt2 gates always executed / executed unless skipped
};

struct PassC {
    uint8_t kind;            // 0: one op on 4 items x 4 amplitudes, 1: two ops on 1 item x 16 amplitudes, 2: idle only
    uint8_t o1, o2;
    uint8_t idle_word;       // 255: none; else event word holding a 17-bit Z mask applied after the pass
    uint8_t last;            // 1: last pass with gates -> applies the deferred scale
    uint8_t pad[3];
    // bank-conflict-free lane mapping for the 128B-swizzled tile layout (k_layer3): item bit k -> tile-local bit
    // pos[k]; the three lowest item bits (a quarter warp) land on bits whose swizzled 16-byte columns differ
    uint8_t pos[12];
    uint32_t idle_filter;    // qubits of the idle mask applied at this point (0x1FFFF: all); used by k_half
};

struct LayerArg2 {
    LayerArg A;
    OpC oc[S17_MAX_LAYER_OPS];
    PassC pass[S17_MAX_LAYER_OPS];
    uint32_t n_pass;
    // support of the state before this sweep (ancilla bits that may be non-zero): tiles / runs outside it are
    // known to be exactly zero and are neither read nor written
    uint8_t tile_shift;      // log2(number of tiles of a shot inside the support)
    uint8_t first;           // 1: first sweep after a measurement -> input is the surviving block (lazy collapse)
    uint16_t in_runs;        // bit r set: run r of a support tile may be non-zero on input
    uint8_t tile_ids[32];
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// mbarrier wait for the single-lane producer warps: back off so the spin does not steal issue slots
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t phase) {
    uint32_t ok = 0;
    for (;;) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
        if (ok) break;
        __nanosleep(200);
    }
}
__device__ __forceinline__ void group_barrier(int grp) {
    asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(kGroupThreads) : "memory");
}

// quad accessors: op 1 of a pass owns register bits 0,1; op 2 owns bits 2,3
template <bool OP2>
struct Quad {
    static __device__ __forceinline__ int r(int q, int k) { return OP2 ? (q + 4 * k) : (4 * q + k); }
};

#define S17_FOR_QUADS(body)                                                                                  \
    _Pragma("unroll") for (int q = 0; q < 4; ++q) {                                                          \
        double2 v[4] = {a[Quad<OP2>::r(q, 0)], a[Quad<OP2>::r(q, 1)], a[Quad<OP2>::r(q, 2)], a[Quad<OP2>::r(q, 3)]}; \
        body;                                                                                                \
        a[Quad<OP2>::r(q, 0)] = v[0]; a[Quad<OP2>::r(q, 1)] = v[1];                                          \
        a[Quad<OP2>::r(q, 2)] = v[2]; a[Quad<OP2>::r(q, 3)] = v[3];                                          \
    }

// the fixed gate template of x_type_entangling / z_type_entangling (CS:291-344)
template <bool OP2>
__device__ __forceinline__ void apply_standard(double2 (&a)[16], const OpC c, uint32_t w) {
    const double s = (double)c.s;
    const bool skip = (w >> 24) & 1u;
    const bool ev = (w & 0xFFFFFFu) != 0;
    if (c.shape & 1) {
        S17_FOR_QUADS(g_ry(v, 1.0));
        if (ev && (w & 63u)) { const uint32_t f = w & 63u; S17_FOR_QUADS(g_pauli(v, f)); }
    }
    if (!skip) {
        S17_FOR_QUADS(g_rxx(v, s));
        if (ev && ((w >> 6) & 63u)) { const uint32_t f = (w >> 6) & 63u; S17_FOR_QUADS(g_pauli(v, f)); }
    }
    if (c.shape & 2) {
        S17_FOR_QUADS(g_rx(v, -s));
        if (ev && ((w >> 12) & 63u)) { const uint32_t f = (w >> 12) & 63u; S17_FOR_QUADS(g_pauli(v, f)); }
    }
    if (c.shape & 4) {
        S17_FOR_QUADS(g_ry(v, -1.0));
        if (ev && ((w >> 18) & 63u)) { const uint32_t f = (w >> 18) & 63u; S17_FOR_QUADS(g_pauli(v, f)); }
    }
}

// interpreter for operations with non-standard micro-ops (coherent over-rotations, Hadamards)
template <bool OP2>
__device__ __forceinline__ void apply_generic(double2 (&a)[16], const s17_op &op, const double2 *prm, uint32_t w) {
    const bool skip = (w >> 24) & 1u;
    for (int u = 0; u < op.n_uops; ++u) {
        const s17_uop up = op.uops[u];
        if (up.skippable && skip) continue;
        const double sg = (double)up.sign;
        const double2 p = prm[up.param & 63];
        switch (up.type) {
        case S17_U_RY:     S17_FOR_QUADS(g_ry(v, sg)); break;
        case S17_U_RXX:    S17_FOR_QUADS(g_rxx(v, sg)); break;
        case S17_U_RX:     S17_FOR_QUADS(g_rx(v, sg)); break;
        case S17_U_H:      S17_FOR_QUADS(g_h(v)); break;
        case S17_U_ROT_X:  S17_FOR_QUADS(g_rot_x(v, p.x, p.y)); break;
        case S17_U_ROT_Y:  S17_FOR_QUADS(g_rot_y(v, p.x, p.y)); break;
        case S17_U_ROT_XX: S17_FOR_QUADS(g_rot_xx(v, p.x, p.y)); break;
        default: break;
        }
        if (up.ev_shift != 255) {
            const uint32_t f = (w >> up.ev_shift) & 63u;
            if (f) { S17_FOR_QUADS(g_pauli(v, f)); }
        }
    }
}

// Branch-free form of the same template for shots without Pauli events: an absent (or leak-skipped)
// gate is executed with coefficient 0, which is the exact identity in the unscaled add/sub form.
template <bool OP2>
__device__ __forceinline__ void apply_fast(double2 (&a)[16], double v1, double s, double sx, double v2) {
    S17_FOR_QUADS(g_ry(v, v1); g_rxx(v, s); g_rx(v, sx); g_ry(v, v2));
}

struct FastC { double v1, s, sx, v2; };
__device__ __forceinline__ FastC fast_coeffs(const OpC c, uint32_t w = 0) {
    FastC f;
    const double s = (double)c.s;
    f.v1 = (c.shape & 1) ? 1.0 : 0.0;
    f.s = ((w >> 24) & 1u) ? 0.0 : s;          // a leak-skipped Rxx is the identity
    f.sx = (c.shape & 2) ? -s : 0.0;
    f.v2 = (c.shape & 4) ? -1.0 : 0.0;
    return f;
}

template <bool OP2>
__device__ __forceinline__ void apply_op(double2 (&a)[16], const LayerArg2 &P, int o, uint32_t w) {
    if (P.oc[o].generic) apply_generic<OP2>(a, P.A.L.ops[o], P.A.prm, w);
    else apply_standard<OP2>(a, P.oc[o], w);
}

struct PassGeom { int p0, p1, p2, p3, m0, m1, m2, m3, n; };   // n = thread-items per pass (tile amplitudes / 16)

// 128-byte TMA swizzle expressed on amplitude (16-byte) indices: chunk bits 0-2 are XORed with row bits 3-5
__device__ __forceinline__ int swz(int idx) { return idx ^ ((idx >> 3) & 7); }

// swizzled-layout variant of item_offsets: deposits the item bits at PassC::pos and returns PHYSICAL offsets
template <int KIND>
__device__ __forceinline__ void item_offsets_swz(const PassGeom &g, const PassC &ps, int item, int (&off)[16]) {
    const int nb = 31 - __clz(g.n);                 // item bits (8 for 4096-amplitude tiles)
    int i0 = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k)
        if (k < nb) i0 |= ((item >> k) & 1) << ps.pos[k];
    i0 = swz(i0);
    if (KIND == 1) {
        const int q0 = swz(g.m0), q1 = swz(g.m1), q2 = swz(g.m2), q3 = swz(g.m3);
#pragma unroll
        for (int r = 0; r < 16; ++r)
            off[r] = i0 ^ ((r & 1) ? q0 : 0) ^ ((r & 2) ? q1 : 0) ^ ((r & 4) ? q2 : 0) ^ ((r & 8) ? q3 : 0);
    } else {
        const int q0 = swz(g.m0), q1 = swz(g.m1);
        const int e0 = swz(1 << ps.pos[nb]), e1 = swz(1 << ps.pos[nb + 1]);    // the two quad-index bits
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                off[4 * i + k] = i0 ^ ((i & 1) ? e0 : 0) ^ ((i & 2) ? e1 : 0) ^ ((k & 1) ? q0 : 0) ^ ((k & 2) ? q1 : 0);
    }
}

// tile offsets of the 16 amplitudes a thread owns in a pass.  KIND 1: one item, four register bits
// (op1 data, op1 ancilla, op2 data, op2 ancilla); KIND 0: four items, two register bits each.
template <int KIND>
__device__ __forceinline__ void item_offsets(const PassGeom &g, int item, int (&off)[16]) {
    if (KIND == 1) {
        const int i0 = insert0(insert0(insert0(insert0(item, g.p0), g.p1), g.p2), g.p3);
#pragma unroll
        for (int r = 0; r < 16; ++r)
            off[r] = i0 + ((r & 1) ? g.m0 : 0) + ((r & 2) ? g.m1 : 0) + ((r & 4) ? g.m2 : 0) + ((r & 8) ? g.m3 : 0);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int i0 = insert0(insert0(item + g.n * i, g.p0), g.p1);   // consecutive lanes -> consecutive quads
#pragma unroll
            for (int k = 0; k < 4; ++k) off[4 * i + k] = i0 + ((k & 1) ? g.m0 : 0) + ((k & 2) ? g.m1 : 0);
        }
    }
}

template <int KIND>
__device__ __forceinline__ void fast_item(double2 *tile, const PassGeom &g, int item, const FastC f1, const FastC f2,
                                          double sc) {
    double2 a[16];
    {
        int off[16];
        item_offsets<KIND>(g, item, off);
#pragma unroll
        for (int r = 0; r < 16; ++r) a[r] = tile[off[r]];
    }
    apply_fast<false>(a, f1.v1, f1.s, f1.sx, f1.v2);
    if (KIND == 1) apply_fast<true>(a, f2.v1, f2.s, f2.sx, f2.v2);
    {
        // recompute the offsets instead of keeping 16 address registers live across the FP64 block
        int item2 = item;
        asm volatile("" : "+r"(item2));
        int off[16];
        item_offsets<KIND>(g, item2, off);
#pragma unroll
        for (int r = 0; r < 16; ++r) tile[off[r]] = make_double2(a[r].x * sc, a[r].y * sc);
    }
}

// 2-D tensor-map TMA of one 8 KiB run (64 rows x 128 B) with the hardware 128-byte swizzle
__device__ __forceinline__ void tma_load_run(void *dst_smem, const CUtensorMap *tm, int row, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst_smem)),
        "l"(tm), "r"(0), "r"(row), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_run(const CUtensorMap *tm, int row, const void *src_smem) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(0),
                 "r"(row), "r"(smem_u32(src_smem))
                 : "memory");
}

// fast_item on the 128-byte-swizzled tile layout (tensor-map TMA): the item bits are deposited at PassC::pos so that a
// quarter warp's 16-byte accesses land on eight different bank groups whatever data bit the pass holds in registers
template <int KIND>
__device__ __forceinline__ void fast_item_swz(double2 *tile, const PassGeom &g, const PassC &ps, int item, const FastC f1,
                                              const FastC f2, double sc) {
    double2 a[16];
    {
        int off[16];
        item_offsets_swz<KIND>(g, ps, item, off);
#pragma unroll
        for (int r = 0; r < 16; ++r) a[r] = tile[off[r]];
    }
    apply_fast<false>(a, f1.v1, f1.s, f1.sx, f1.v2);
    if (KIND == 1) apply_fast<true>(a, f2.v1, f2.s, f2.sx, f2.v2);
    {
        // recompute the offsets instead of keeping 16 address registers live across the FP64 block
        int item2 = item;
        asm volatile("" : "+r"(item2));
        int off[16];
        item_offsets_swz<KIND>(g, ps, item2, off);
#pragma unroll
        for (int r = 0; r < 16; ++r) tile[off[r]] = make_double2(a[r].x * sc, a[r].y * sc);
    }
}

// everything else (Pauli events, leak-skipped Rxx, idle Z masks, coherent rotations): rare, kept out of line
__device__ __noinline__ void slow_item(double2 *tile, const LayerArg2 *Pp, const PassC ps, const PassGeom g, int item,
                                       uint32_t w1, uint32_t w2, uint32_t zm, uint32_t lm, uint32_t c0, double sc,
                                       bool swizzled = false) {
    const LayerArg2 &P = *Pp;
    double2 a[16];
    int off[16];
    if (swizzled) {
        if (ps.kind == 1) item_offsets_swz<1>(g, ps, item, off);
        else item_offsets_swz<0>(g, ps, item, off);
    } else {
        if (ps.kind == 1) item_offsets<1>(g, item, off);
        else item_offsets<0>(g, item, off);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) a[r] = tile[off[r]];
    if (ps.kind == 1) {
        apply_op<false>(a, P, ps.o1, w1);
        apply_op<true>(a, P, ps.o2, w2);
    } else if (ps.kind == 0) {
        apply_op<false>(a, P, ps.o1, w1);
    }
    if (zm) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int logical = swizzled ? swz(off[r]) : off[r];       // the swizzle is an involution
            if ((__popc(logical & lm) & 1u) ^ c0) a[r] = make_double2(-a[r].x, -a[r].y);
        }
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) tile[off[r]] = make_double2(a[r].x * sc, a[r].y * sc);
}

__device__ __forceinline__ void tile_coords(const LayerArg &A, int tile_id, uint32_t &base, uint32_t (&run_off)[8]) {
    base = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) base |= ((tile_id >> k) & 1u) << A.other_bits[k];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        uint32_t g = base;
#pragma unroll
        for (int k = 0; k < 3; ++k) g |= ((r >> k) & 1u) << A.L.anc_bits[k];
        run_off[r] = g;
    }
}

__global__ void __launch_bounds__(kL2Threads, 1)
k_layer2(const __grid_constant__ CUtensorMap tmap, double2 *__restrict__ state, const __grid_constant__ LayerArg2 P,
         const uint32_t *__restrict__ ev,
         const uint32_t *__restrict__ act_list, const uint32_t *__restrict__ n_act_ptr, double *__restrict__ hist,
         unsigned long long *__restrict__ counters, const ShotAux *__restrict__ aux, const double2 *__restrict__ zero_page,
         int init_basis)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double2 *tiles = reinterpret_cast<double2 *>(smem_raw);
    uint32_t *evbuf = reinterpret_cast<uint32_t *>(smem_raw + (size_t)kRing * kTile * 16);
    ShotAux *auxbuf = reinterpret_cast<ShotAux *>(smem_raw + (size_t)kRing * kTile * 16 + kRing * kEvBytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kRing * kTile * 16 + kRing * kEvBytes + kRing * kAuxBytes);
    uint64_t *done = full + kRing;
    uint64_t *free_ = done + kRing;

    const int t = threadIdx.x;
    const uint32_t n_items = (*n_act_ptr) << P.tile_shift;
    const uint32_t tile_mask = (1u << P.tile_shift) - 1u;
    const uint32_t chunk = (n_items + gridDim.x - 1) / gridDim.x;
    const uint32_t first = blockIdx.x * chunk;
    const uint32_t last = min(n_items, first + chunk);
    if (first >= last) return;
    const uint32_t n_local = last - first;

    if (t == 0) {
        for (int b = 0; b < kRing; ++b) {
            mbar_init(&full[b], 1);
            mbar_init(&done[b], kGroupThreads / 32);
            mbar_init(&free_[b], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int warp = t >> 5, lane = t & 31;

    if (warp == kMathThreads / 32) {
        // ------------------------------- loader ---------------------------------------------------
        if (lane == 0) {
            for (uint32_t j = 0; j < n_local; ++j) {
                const int b = j % kRing;
                if (j >= (uint32_t)kRing) mbar_wait_sleep(&free_[b], ((j / kRing) - 1) & 1);
                const uint32_t item = first + j;
                const uint32_t shot = act_list[item >> P.tile_shift];
                uint32_t base, run_off[8];
                tile_coords(P.A, P.tile_ids[item & tile_mask], base, run_off);
                const int row0 = (int)(shot * (uint32_t)(kDim / 8));       // tensor-map row (8 amplitudes = 128 B) of the shot
                double2 *tile = tiles + (size_t)b * kTile;
                if (init_basis < 0) {
                    mbar_expect_tx(&full[b], kTile * 16 + kEvBytes + kAuxBytes);
                    if (P.first) {
                        // lazy collapse + reset: the only non-zero input is the surviving 512-amplitude block
                        const ShotAux ax = aux[shot];
                        const uint32_t blk = ax.collapsed ? 0u : ax.anc_outcome;
                        tma_load_run(tile, &tmap, row0 + (int)blk * (kRun / 8), &full[b]);
#pragma unroll
                        for (int r = 1; r < 8; ++r) tma_load_1d(tile + r * kRun, zero_page, kRunBytes, &full[b]);
                    } else {
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            if ((P.in_runs >> r) & 1) tma_load_run(tile + r * kRun, &tmap, row0 + (int)(run_off[r] >> 3), &full[b]);
                            else tma_load_1d(tile + r * kRun, zero_page, kRunBytes, &full[b]);      // zeros in any layout
                        }
                    }
                } else {
                    mbar_expect_tx(&full[b], kEvBytes + kAuxBytes);
                }
                tma_load_1d(evbuf + b * S17_EV_WORDS, ev + (size_t)shot * S17_EV_WORDS, kEvBytes, &full[b]);
                tma_load_1d(auxbuf + b, aux + shot, kAuxBytes, &full[b]);
            }
        }
        return;
    }
    if (warp == kMathThreads / 32 + 1) {
        // ------------------------------- storer ---------------------------------------------------
        if (lane == 0) {
            for (uint32_t j = 0; j < n_local; ++j) {
                const int b = j % kRing;
                mbar_wait_sleep(&done[b], (j / kRing) & 1);
                const uint32_t item = first + j;
                const uint32_t shot = act_list[item >> P.tile_shift];
                uint32_t base, run_off[8];
                tile_coords(P.A, P.tile_ids[item & tile_mask], base, run_off);
                const int row0 = (int)(shot * (uint32_t)(kDim / 8));
                const double2 *tile = tiles + (size_t)b * kTile;
#pragma unroll
                for (int r = 0; r < 8; ++r) tma_store_run(&tmap, row0 + (int)(run_off[r] >> 3), tile + r * kRun);
                tma_store_commit_and_wait_read();
                mbar_arrive(&free_[b]);
            }
            // roofline accounting: (shot, sweep) pairs and 8 KiB run transfers that really touched HBM
            const uint32_t in_runs = init_basis >= 0 ? 0u : (P.first ? 1u : (uint32_t)__popc(P.in_runs));
            atomicAdd(&counters[4], (unsigned long long)((last + tile_mask) >> P.tile_shift) - ((first + tile_mask) >> P.tile_shift));
            atomicAdd(&counters[8], (unsigned long long)n_local * (in_runs + 8u));   // C_RUNS_MOVED
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        return;
    }

    // ----------------------------------- math warps ----------------------------------------------
    // two independent groups of 4 warps: group g owns tiles j = g, g+2, ... so that one group's
    // shared-memory phase overlaps the other group's FP64 phase
    const int grp = warp / (kGroupThreads / 32);
    const int tg = t - grp * kGroupThreads;                  // thread index inside the group
    const int gw = warp - grp * (kGroupThreads / 32);        // warp index inside the group
    for (uint32_t j = 0; j < n_local; ++j) {
        const int b = j % kRing;
        // Every group observes EVERY fill of every buffer, in order, also the fills the other group consumes: a
        // parity wait is only meaningful for the current or the immediately preceding phase, so a group that
        // skipped a fill could see "phase k-1 complete" while waiting for phase k+1 and start on stale data.
        mbar_wait(&full[b], (j / kRing) & 1);
        if ((int)(j % kGroups) != grp) continue;
        double2 *tile = tiles + (size_t)b * kTile;
        const uint32_t *evw = evbuf + b * S17_EV_WORDS;
        const uint32_t item = first + j;
        const int tile_id = P.tile_ids[item & tile_mask];
        uint32_t base = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) base |= ((tile_id >> k) & 1u) << P.A.other_bits[k];

        if (init_basis >= 0) {
#pragma unroll 4
            for (int i = 0; i < kTile / kGroupThreads; ++i) tile[tg + kGroupThreads * i] = make_double2(0.0, 0.0);
            group_barrier(grp);
            if (tile_id == 0 && tg == 0) tile[swz(init_basis)] = make_double2(1.0, 0.0);
            group_barrier(grp);
        }

        // deferred 2^(-n/2): count the 1/sqrt2 gates this shot really executes in this layer
        int nsc = 0;
        for (int o = 0; o < P.A.L.n_ops; ++o) {
            if (P.A.L.ops[o].kind != S17_OP_ENT) continue;
            const uint32_t skip = (evw[P.oc[o].ev_word] >> 24) & 1u;
            nsc += P.oc[o].n_fixed + (skip ? 0 : P.oc[o].n_skip);
        }
        double scale = ldexp((nsc & 1) ? 0.70710678118654752440 : 1.0, -(nsc >> 1));
        if (P.first && init_basis < 0 && !auxbuf[b].collapsed) scale *= auxbuf[b].scale;   // renormalisation of the lazy collapse

        for (uint32_t p = 0; p < P.n_pass; ++p) {
            const PassC ps = P.pass[p];
            uint32_t zm = 0, lm = 0, c0 = 0;
            if (ps.idle_word != 255) {
                // Z on every qubit whose bit is set in the 17-bit idle mask (sympathetic_cooling, CS:762-772)
                zm = evw[ps.idle_word];
                lm = zm & 0x1FFu;
#pragma unroll
                for (int k = 0; k < 3; ++k) lm |= ((zm >> P.A.L.anc_bits[k]) & 1u) << (9 + k);
                c0 = __popc(base & zm) & 1u;
            }
            const double sc = ps.last ? scale : 1.0;
            const OpC c1 = P.oc[ps.o1], c2 = P.oc[ps.kind == 1 ? ps.o2 : ps.o1];
            const uint32_t w1 = evw[c1.ev_word];
            const uint32_t w2 = ps.kind == 1 ? evw[c2.ev_word] : 0u;
            // fast path: standard ops, no Pauli event, no leak skip, no idle mask -> straight-line code whose
            // coefficients all come from the constant bank
            const bool fast = ps.kind != 2 && !(c1.generic | (ps.kind == 1 ? c2.generic : 0)) &&
                              !((w1 | w2) & 0x1FFFFFFu) && zm == 0;
            PassGeom g;
            g.n = kTile / 16;
            if (ps.kind == 1) {
                g.p0 = min(c1.lb0, c2.lb0); g.p1 = max(c1.lb0, c2.lb0);
                g.p2 = min(c1.lb1, c2.lb1); g.p3 = max(c1.lb1, c2.lb1);
                g.m0 = 1 << c1.lb0; g.m1 = 1 << c1.lb1; g.m2 = 1 << c2.lb0; g.m3 = 1 << c2.lb1;
            } else {
                const int lb0 = ps.kind == 0 ? c1.lb0 : 0, lb1 = ps.kind == 0 ? c1.lb1 : 9;
                g.p0 = lb0; g.p1 = lb1; g.p2 = g.p3 = 0;
                g.m0 = 1 << lb0; g.m1 = 1 << lb1; g.m2 = g.m3 = 0;
            }
            if (fast) {
                const FastC f1 = fast_coeffs(c1), f2 = fast_coeffs(c2);
#pragma unroll 1
                for (int it = 0; it < kTile / 16 / kGroupThreads; ++it) {
                    if (ps.kind == 1) fast_item_swz<1>(tile, g, ps, tg + kGroupThreads * it, f1, f2, sc);
                    else fast_item_swz<0>(tile, g, ps, tg + kGroupThreads * it, f1, f2, sc);
                }
            } else {
#pragma unroll 1
                for (int it = 0; it < kTile / 16 / kGroupThreads; ++it)
                    slow_item(tile, &P, ps, g, tg + kGroupThreads * it, w1, w2, zm, lm, c0, sc, true);
            }
            group_barrier(grp);
        }

        if (P.A.L.meas_mask) {
            // joint ancilla histogram: every warp of the group owns two runs (ancilla values) of the tile
#pragma unroll
            for (int rr = 0; rr < 8 / (kGroupThreads / 32); ++rr) {
                const int run = gw + (kGroupThreads / 32) * rr;
                double acc = 0.0;
#pragma unroll
                for (int i = 0; i < kRun / 32; ++i) {
                    const double2 v = tile[run * kRun + lane + 32 * i];
                    acc = fma(v.x, v.x, fma(v.y, v.y, acc));
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
                if (lane == 0) {
                    uint32_t g = base;
#pragma unroll
                    for (int k = 0; k < 3; ++k) g |= ((run >> k) & 1u) << P.A.L.anc_bits[k];
                    const uint32_t shot = act_list[item >> P.tile_shift];
                    hist[(size_t)shot * 256 + (g >> 9)] = acc;
                }
            }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[b]);
    }
}

}  // namespace s17
