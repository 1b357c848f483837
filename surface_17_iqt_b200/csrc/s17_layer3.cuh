// s17_layer3.cuh -- gate sweep for the support-aware schedule (the production path).
//
// After an ancilla reset only a small part of a shot's 2^17 amplitudes can be non-zero, so a sweep
// touches one or two 64 KiB tiles per shot, reads a single 8 KiB run from HBM and is bound by the
// in-SM work (FP64 pipe + shared memory), not by HBM.  This kernel therefore maximises math warps
// instead of pipelining HBM: three 128-thread CTAs per SM, each with ONE tile buffer and all four warps
// doing math; the TMA load / store latency of one CTA hides behind the math of the other two.
// (k_layer2 remains the kernel for dense, HBM-bound sweeps: S17_DENSE=1.)
#pragma once
#include <cuda.h>

#include "s17_layer2.cuh"

namespace s17 {

constexpr int kL3Threads = 128;
constexpr int kL3CtasPerSm = 3;
constexpr size_t kL3Smem = (size_t)kTile * 16 + 2 * (kEvBytes + kAuxBytes) + 64;

// x-type operations (timesteps 1-4) have no Ry: two gate slots instead of four
template <bool OP2>
__device__ __forceinline__ void apply_fast_x(double2 (&a)[16], double s, double sx) {
    S17_FOR_QUADS(g_rxx(v, s); g_rx(v, sx));
}

template <bool OP2>
__device__ __forceinline__ void apply_end_pauli(double2 (&a)[16], uint32_t e) {
    S17_FOR_QUADS(g_pauli(v, e));
}

// e1 / e2: net Pauli event of the operation, conjugated to its end by k_plan (0: none)
template <int KIND, bool XONLY, bool SCALE>
__device__ __forceinline__ void fast_item3(double2 *tile, const PassGeom &g, const PassC &ps, int item, const FastC f1,
                                           const FastC f2, double sc, uint32_t e1 = 0, uint32_t e2 = 0) {
    double2 a[16];
    int off[16];
    item_offsets_swz<KIND>(g, ps, item, off);
#pragma unroll
    for (int r = 0; r < 16; ++r) a[r] = tile[off[r]];
    if (XONLY) {
        apply_fast_x<false>(a, f1.s, f1.sx);
        if (e1) apply_end_pauli<false>(a, e1);
        if (KIND == 1) apply_fast_x<true>(a, f2.s, f2.sx);
    } else {
        apply_fast<false>(a, f1.v1, f1.s, f1.sx, f1.v2);
        if (e1) apply_end_pauli<false>(a, e1);
        if (KIND == 1) apply_fast<true>(a, f2.v1, f2.s, f2.sx, f2.v2);
    }
    if (KIND == 1 && e2) apply_end_pauli<true>(a, e2);
#pragma unroll
    for (int r = 0; r < 16; ++r) tile[off[r]] = SCALE ? make_double2(a[r].x * sc, a[r].y * sc) : a[r];
}

template <int KIND, bool XONLY>
__device__ __forceinline__ void fast_pass3(double2 *tile, const PassGeom &g, const PassC &ps, int t, const FastC f1,
                                           const FastC f2, double sc, bool scale, uint32_t e1 = 0, uint32_t e2 = 0) {
#pragma unroll 1
    for (int it = 0; it < kTile / 16 / kL3Threads; ++it) {
        if (scale) fast_item3<KIND, XONLY, true>(tile, g, ps, t + kL3Threads * it, f1, f2, sc, e1, e2);
        else fast_item3<KIND, XONLY, false>(tile, g, ps, t + kL3Threads * it, f1, f2, sc, e1, e2);
    }
}

__global__ void __launch_bounds__(kL3Threads, kL3CtasPerSm)
k_layer3(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ LayerArg2 P, const uint32_t *__restrict__ ev,
         const uint32_t *__restrict__ act_list, const uint32_t *__restrict__ n_act_ptr, double *__restrict__ hist,
         unsigned long long *__restrict__ counters, const ShotAux *__restrict__ aux, int init_basis)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double2 *tile = reinterpret_cast<double2 *>(smem_raw);
    uint32_t *evw2 = reinterpret_cast<uint32_t *>(smem_raw + (size_t)kTile * 16);                 // [2][32]
    ShotAux *auxs2 = reinterpret_cast<ShotAux *>(smem_raw + (size_t)kTile * 16 + 2 * kEvBytes);   // [2]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kTile * 16 + 2 * (kEvBytes + kAuxBytes));
    uint64_t *meta = full + 1;                                                                   // [2]

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t n_items = (*n_act_ptr) << P.tile_shift;
    const uint32_t tile_mask = (1u << P.tile_shift) - 1u;
    const uint32_t chunk = (n_items + gridDim.x - 1) / gridDim.x;
    const uint32_t first = blockIdx.x * chunk;
    const uint32_t last = min(n_items, first + chunk);
    if (first >= last) return;

    if (t == 0) {
        mbar_init(full, 1);
        mbar_init(&meta[0], 1);
        mbar_init(&meta[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // per-shot metadata (event words, collapse record) is prefetched one tile ahead
    auto prefetch_meta = [&](uint32_t shot_, int slot) {
        mbar_expect_tx(&meta[slot], kEvBytes + kAuxBytes);
        tma_load_1d(evw2 + slot * S17_EV_WORDS, ev + (size_t)shot_ * S17_EV_WORDS, kEvBytes, &meta[slot]);
        tma_load_1d(auxs2 + slot, aux + shot_, kAuxBytes, &meta[slot]);
    };
    uint32_t shot_next = act_list[first >> P.tile_shift];
    if (t == 0) prefetch_meta(shot_next, 0);

    // which runs of a tile are read from HBM; everything else is known to be zero on input
    const uint32_t load_mask = init_basis >= 0 ? 0u : (P.first ? 1u : (uint32_t)P.in_runs);
    const uint32_t n_load = __popc(load_mask);

    for (uint32_t item = first; item < last; ++item) {
        const uint32_t j = item - first;
        const uint32_t shot = shot_next;
        if (item + 1 < last) shot_next = act_list[(item + 1) >> P.tile_shift];     // consumed next iteration
        const int slot = j & 1;
        const uint32_t *evw = evw2 + slot * S17_EV_WORDS;
        const ShotAux *auxs = auxs2 + slot;
        const int tile_id = P.tile_ids[item & tile_mask];
        uint32_t base = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) base |= ((tile_id >> k) & 1u) << P.A.other_bits[k];
        const int row0 = (int)(shot * (uint32_t)(kDim / 8));       // tensor-map row (8 amplitudes = 128 B) of the shot

        mbar_wait(&meta[slot], (j >> 1) & 1);
        if (t == 0) {
            if (n_load) mbar_expect_tx(full, n_load * kRunBytes);
            if (P.first && init_basis < 0) {
                // lazy collapse + reset: the only non-zero input is the surviving 512-amplitude block
                const uint32_t blk = auxs->collapsed ? 0u : auxs->anc_outcome;
                tma_load_run(tile, &tmap, row0 + (int)blk * (kRun / 8), full);
            } else {
                for (int r = 0; r < 8; ++r) {
                    if (!((load_mask >> r) & 1u)) continue;
                    uint32_t g = base;
#pragma unroll
                    for (int k = 0; k < 3; ++k) g |= ((r >> k) & 1u) << P.A.L.anc_bits[k];
                    tma_load_run(tile + r * kRun, &tmap, row0 + (int)(g >> 3), full);
                }
            }
            if (item + 1 < last) prefetch_meta(shot_next, slot ^ 1);
        }
        // zero-fill the runs that are not loaded (disjoint from the TMA destinations)
        for (int r = 0; r < 8; ++r) {
            if ((load_mask >> r) & 1u) continue;
#pragma unroll
            for (int i = 0; i < kRun / kL3Threads; ++i) tile[r * kRun + t + kL3Threads * i] = make_double2(0.0, 0.0);
        }
        if (n_load) mbar_wait(full, j & 1);
        if (init_basis >= 0 && tile_id == 0 && t == 0) tile[swz(init_basis)] = make_double2(1.0, 0.0);
        __syncthreads();

        // deferred 2^(-n/2): count the 1/sqrt2 gates this shot really executes in this layer
        const uint32_t skips = __ballot_sync(0xffffffffu, (evw[lane] >> 24) & 1u);    // bit w: op word w is leak-skipped
        int nsc = 0;
        for (int o = 0; o < P.A.L.n_ops; ++o) {
            if (P.A.L.ops[o].kind != S17_OP_ENT) continue;
            nsc += P.oc[o].n_fixed + (((skips >> P.oc[o].ev_word) & 1u) ? 0 : P.oc[o].n_skip);
        }
        double scale = ldexp((nsc & 1) ? 0.70710678118654752440 : 1.0, -(nsc >> 1));
        if (P.first && init_basis < 0 && !auxs->collapsed) scale *= auxs->scale;   // renormalisation of the lazy collapse

        for (uint32_t p = 0; p < P.n_pass; ++p) {
            const PassC &ps = P.pass[p];
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
            // fast path: standard (Clifford) ops, no idle mask; Pauli events arrive as one net Pauli per op, a leak
            // skip as a zero coefficient
            const bool fast = ps.kind != 2 && !(c1.generic | (ps.kind == 1 ? c2.generic : 0)) && zm == 0;
            const uint32_t e1 = (w1 >> 25) & 63u, e2 = (w2 >> 25) & 63u;
            const bool xonly = !((c1.shape | c2.shape) & 5);
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
            // 256 items per pass, two per thread
            if (fast) {
                const FastC f1 = fast_coeffs(c1, w1), f2 = fast_coeffs(c2, w2);
                const bool scale = ps.last && sc != 1.0;
                if (ps.kind == 1) {
                    if (xonly) fast_pass3<1, true>(tile, g, ps, t, f1, f2, sc, scale, e1, e2);
                    else fast_pass3<1, false>(tile, g, ps, t, f1, f2, sc, scale, e1, e2);
                } else {
                    if (xonly) fast_pass3<0, true>(tile, g, ps, t, f1, f2, sc, scale, e1, 0);
                    else fast_pass3<0, false>(tile, g, ps, t, f1, f2, sc, scale, e1, 0);
                }
            } else {
#pragma unroll 1
                for (int it = 0; it < kTile / 16 / kL3Threads; ++it)
                    slow_item(tile, &P, ps, g, t + kL3Threads * it, w1, w2, zm, lm, c0, sc, true);
            }
            __syncthreads();
        }

        // hand the tile to the TMA store engine, then compute the histogram while it drains
        fence_proxy_async_smem();
        __syncthreads();
        if (t == 0) {
            for (int r = 0; r < 8; ++r) {
                uint32_t g = base;
#pragma unroll
                for (int k = 0; k < 3; ++k) g |= ((r >> k) & 1u) << P.A.L.anc_bits[k];
                tma_store_run(&tmap, row0 + (int)(g >> 3), tile + r * kRun);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (P.A.L.meas_mask) {
#pragma unroll
            for (int rr = 0; rr < 8 / (kL3Threads / 32); ++rr) {
                const int run = warp + (kL3Threads / 32) * rr;
                double acc4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int i = 0; i < kRun / 32; ++i) {
                    const double2 v = tile[run * kRun + lane + 32 * i];
                    acc4[i & 3] = fma(v.x, v.x, fma(v.y, v.y, acc4[i & 3]));
                }
                double acc = (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
                if (lane == 0) {
                    uint32_t g = base;
#pragma unroll
                    for (int k = 0; k < 3; ++k) g |= ((run >> k) & 1u) << P.A.L.anc_bits[k];
                    hist[(size_t)shot * 256 + (g >> 9)] = acc;
                }
            }
        }
        if (t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
    }
    if (t == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        atomicAdd(&counters[4], (unsigned long long)((last + tile_mask) >> P.tile_shift) - ((first + tile_mask) >> P.tile_shift));
        atomicAdd(&counters[8], (unsigned long long)(last - first) * (n_load + 8u));   // C_SWEEPS, C_RUNS_MOVED
    }
}

}  // namespace s17
