// s17_layer4.cuh -- gate sweep over 8192-amplitude tiles (9 data bits + 4 ancilla bits = 128 KiB): one
// sweep applies FOUR timesteps (12 entangling operations), i.e. a stabiliser cycle is two sweeps.
//
// Why a second kernel: with the support-aware schedule (only the part of the state that can be non-zero
// after the ancilla reset is touched) the sweeps are bound by the in-SM work, not by HBM, and larger
// fusion shrinks the work itself -- the X-ancilla half of the cycle runs on ONE tile per shot, the
// Z-ancilla half on 16 tiles that each read a single 8 KiB run.  A 128 KiB tile leaves room for one
// buffer per SM, so this kernel is not a TMA ring; instead all 8 warps do math with two independent
// items per thread in flight (the registers the producer warps of k_layer2 would have cost), and the
// TMA store of a tile overlaps the histogram pass.
#pragma once
#include "s17_layer2.cuh"

namespace s17 {

constexpr int kTile4 = 1 << 13;
constexpr int kRuns4 = 16;
constexpr int kL4Threads = 256;
constexpr size_t kL4Smem = (size_t)kTile4 * 16 + kEvBytes + kAuxBytes + 64;

template <int KIND>
__device__ __forceinline__ void fast_item2(double2 *tile, const PassGeom &g, int item0, int item1, const FastC f1,
                                           const FastC f2, double sc) {
    double2 a[16], b[16];
    {
        int off[16];
        item_offsets<KIND>(g, item0, off);
#pragma unroll
        for (int r = 0; r < 16; ++r) a[r] = tile[off[r]];
        item_offsets<KIND>(g, item1, off);
#pragma unroll
        for (int r = 0; r < 16; ++r) b[r] = tile[off[r]];
    }
    apply_fast<false>(a, f1.v1, f1.s, f1.sx, f1.v2);
    if (KIND == 1) apply_fast<true>(a, f2.v1, f2.s, f2.sx, f2.v2);
    {
        int it = item0;
        asm volatile("" : "+r"(it));
        int off[16];
        item_offsets<KIND>(g, it, off);
#pragma unroll
        for (int r = 0; r < 16; ++r) tile[off[r]] = make_double2(a[r].x * sc, a[r].y * sc);
    }
    apply_fast<false>(b, f1.v1, f1.s, f1.sx, f1.v2);
    if (KIND == 1) apply_fast<true>(b, f2.v1, f2.s, f2.sx, f2.v2);
    {
        int it = item1;
        asm volatile("" : "+r"(it));
        int off[16];
        item_offsets<KIND>(g, it, off);
#pragma unroll
        for (int r = 0; r < 16; ++r) tile[off[r]] = make_double2(b[r].x * sc, b[r].y * sc);
    }
}

__global__ void __launch_bounds__(kL4Threads, 1)
k_layer4(double2 *__restrict__ state, const __grid_constant__ LayerArg2 P, const uint32_t *__restrict__ ev,
         const uint32_t *__restrict__ act_list, const uint32_t *__restrict__ n_act_ptr, double *__restrict__ hist,
         unsigned long long *__restrict__ counters, const ShotAux *__restrict__ aux, int init_basis)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double2 *tile = reinterpret_cast<double2 *>(smem_raw);
    uint32_t *evw = reinterpret_cast<uint32_t *>(smem_raw + (size_t)kTile4 * 16);
    ShotAux *auxs = reinterpret_cast<ShotAux *>(smem_raw + (size_t)kTile4 * 16 + kEvBytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kTile4 * 16 + kEvBytes + kAuxBytes);

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t n_items = (*n_act_ptr) << P.tile_shift;
    const uint32_t tile_mask = (1u << P.tile_shift) - 1u;
    const uint32_t chunk = (n_items + gridDim.x - 1) / gridDim.x;
    const uint32_t first = blockIdx.x * chunk;
    const uint32_t last = min(n_items, first + chunk);
    if (first >= last) return;

    if (t == 0) {
        mbar_init(full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // which runs of a tile are read from HBM; everything else is known to be zero on input
    const uint32_t load_mask = init_basis >= 0 ? 0u : (P.first ? 1u : (uint32_t)P.in_runs);
    const uint32_t n_load = __popc(load_mask);

    for (uint32_t item = first; item < last; ++item) {
        const uint32_t j = item - first;
        const uint32_t shot = act_list[item >> P.tile_shift];
        const int tile_id = P.tile_ids[item & tile_mask];
        uint32_t base = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) base |= ((tile_id >> k) & 1u) << P.A.other_bits[k];
        double2 *sbase = state + (size_t)shot * kDim;

        if (t == 0) {
            mbar_expect_tx(full, n_load * kRunBytes + kEvBytes + kAuxBytes);
            if (P.first && init_basis < 0) {
                // lazy collapse + reset: the only non-zero input is the surviving 512-amplitude block
                const ShotAux ax = aux[shot];
                const uint32_t blk = ax.collapsed ? 0u : ax.anc_outcome;
                tma_load_1d(tile, sbase + (size_t)blk * kRun, kRunBytes, full);
            } else {
                for (int r = 0; r < kRuns4; ++r) {
                    if (!((load_mask >> r) & 1u)) continue;
                    uint32_t g = base;
#pragma unroll
                    for (int k = 0; k < 4; ++k) g |= ((r >> k) & 1u) << P.A.L.anc_bits[k];
                    tma_load_1d(tile + r * kRun, sbase + g, kRunBytes, full);
                }
            }
            tma_load_1d(evw, ev + (size_t)shot * S17_EV_WORDS, kEvBytes, full);
            tma_load_1d(auxs, aux + shot, kAuxBytes, full);
        }
        // zero-fill the runs that are not loaded (disjoint from the TMA destinations)
        for (int r = 0; r < kRuns4; ++r) {
            if ((load_mask >> r) & 1u) continue;
            tile[r * kRun + t] = make_double2(0.0, 0.0);
            tile[r * kRun + t + 256] = make_double2(0.0, 0.0);
        }
        mbar_wait(full, j & 1);
        if (init_basis >= 0 && tile_id == 0 && t == 0) tile[init_basis] = make_double2(1.0, 0.0);
        __syncthreads();

        // deferred 2^(-n/2): count the 1/sqrt2 gates this shot really executes in this layer
        int nsc = 0;
        for (int o = 0; o < P.A.L.n_ops; ++o) {
            if (P.A.L.ops[o].kind != S17_OP_ENT) continue;
            const uint32_t skip = (evw[P.oc[o].ev_word] >> 24) & 1u;
            nsc += P.oc[o].n_fixed + (skip ? 0 : P.oc[o].n_skip);
        }
        double scale = ldexp((nsc & 1) ? 0.70710678118654752440 : 1.0, -(nsc >> 1));
        if (P.first && init_basis < 0 && !auxs->collapsed) scale *= auxs->scale;   // renormalisation of the lazy collapse

        for (uint32_t p = 0; p < P.n_pass; ++p) {
            const PassC ps = P.pass[p];
            uint32_t zm = 0, lm = 0, c0 = 0;
            if (ps.idle_word != 255) {
                // Z on every qubit whose bit is set in the 17-bit idle mask (sympathetic_cooling, CS:762-772)
                zm = evw[ps.idle_word];
                lm = zm & 0x1FFu;
#pragma unroll
                for (int k = 0; k < 4; ++k) lm |= ((zm >> P.A.L.anc_bits[k]) & 1u) << (9 + k);
                c0 = __popc(base & zm) & 1u;
            }
            const double sc = ps.last ? scale : 1.0;
            const OpC c1 = P.oc[ps.o1], c2 = P.oc[ps.kind == 1 ? ps.o2 : ps.o1];
            const uint32_t w1 = evw[c1.ev_word];
            const uint32_t w2 = ps.kind == 1 ? evw[c2.ev_word] : 0u;
            const bool fast = ps.kind != 2 && !(c1.generic | (ps.kind == 1 ? c2.generic : 0)) &&
                              !((w1 | w2) & 0x1FFFFFFu) && zm == 0;
            PassGeom g;
            g.n = kTile4 / 16;
            if (ps.kind == 1) {
                g.p0 = min(c1.lb0, c2.lb0); g.p1 = max(c1.lb0, c2.lb0);
                g.p2 = min(c1.lb1, c2.lb1); g.p3 = max(c1.lb1, c2.lb1);
                g.m0 = 1 << c1.lb0; g.m1 = 1 << c1.lb1; g.m2 = 1 << c2.lb0; g.m3 = 1 << c2.lb1;
            } else {
                const int lb0 = ps.kind == 0 ? c1.lb0 : 0, lb1 = ps.kind == 0 ? c1.lb1 : 9;
                g.p0 = lb0; g.p1 = lb1; g.p2 = g.p3 = 0;
                g.m0 = 1 << lb0; g.m1 = 1 << lb1; g.m2 = g.m3 = 0;
            }
            // 512 items per pass, two per thread
            if (fast) {
                const FastC f1 = fast_coeffs(c1), f2 = fast_coeffs(c2);
                if (ps.kind == 1) fast_item2<1>(tile, g, t, t + kL4Threads, f1, f2, sc);
                else fast_item2<0>(tile, g, t, t + kL4Threads, f1, f2, sc);
            } else {
                slow_item(tile, &P, ps, g, t, w1, w2, zm, lm, c0, sc);
                slow_item(tile, &P, ps, g, t + kL4Threads, w1, w2, zm, lm, c0, sc);
            }
            __syncthreads();
        }

        // hand the tile to the TMA store engine, then compute the histogram while it drains
        fence_proxy_async_smem();
        __syncthreads();
        if (t == 0) {
            for (int r = 0; r < kRuns4; ++r) {
                uint32_t g = base;
#pragma unroll
                for (int k = 0; k < 4; ++k) g |= ((r >> k) & 1u) << P.A.L.anc_bits[k];
                tma_store_1d(sbase + g, tile + r * kRun, kRunBytes);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (P.A.L.meas_mask) {
#pragma unroll
            for (int rr = 0; rr < kRuns4 / (kL4Threads / 32); ++rr) {
                const int run = warp + (kL4Threads / 32) * rr;
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
                    for (int k = 0; k < 4; ++k) g |= ((run >> k) & 1u) << P.A.L.anc_bits[k];
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
        atomicAdd(&counters[8], (unsigned long long)(last - first) * (n_load + kRuns4));   // C_SWEEPS, C_RUNS_MOVED
    }
}

}  // namespace s17
