// s17_half.cuh -- one HALF stabiliser cycle per shot, fused into a single shared-memory resident kernel
// (the production path for the default schedule: support-aware sweeps + early X-ancilla measurement).
//
// A half cycle (timesteps 1-4 on the X-type ancillas, or 5-8 on the Z-type ones) starts from the 512 data
// amplitudes that survive the previous measurement and touches four ancillas.  Three of them ("main") plus the
// nine data bits are one 4096-amplitude tile; the operations on the fourth ("tail") ancilla act on qubits that
// no later operation of the half cycle uses, so they are moved to the end (host-side check in make_half).
// Per shot a CTA therefore
//   1. TMA-loads the 8 KiB surviving block (lazy collapse + reset) and zero-fills the rest of the tile,
//   2. runs the main operations as passes over the tile (same code as k_layer3),
//   3. applies the tail operations on registers while accumulating the 16-bin joint histogram of the four
//      ancillas -- the 8192-amplitude post-gate state is never written anywhere,
//   4. samples (or replays) the four outcomes, does the protocol bookkeeping of k_measure,
//   5. recomputes only the selected 512-amplitude block, renormalises it and writes 8 KiB back to HBM.
// HBM traffic per shot and half cycle: 8 KiB read + 8 KiB written (+ ~200 B of metadata).
// This file is included by s17.cu after MeasArgs / record_measurement.
#pragma once

namespace s17 {

struct HalfTail {
    uint8_t n_tail;            // number of tail operations (2)
    uint8_t e_bit;             // state-vector bit of the tail ancilla
    uint8_t op[2];             // their indices in LayerArg2::A.L.ops / oc
    uint8_t idle_word[3];      // idle masks applied before op 0, between the ops, after op 1 (255: none)
    uint8_t meas_mask;         // ancillas measured at the end of this half cycle
    uint32_t idle_filter[3];   // which qubits of those idle masks are applied here (the rest was applied in the main part)
    uint8_t final_part;        // 1: this half cycle completes the cycle's All(Measure) | ancilla
    uint8_t pad[3];
};

// one entangling operation on a single register quad (data + 2 * ancilla), all cases
__device__ __forceinline__ void op_quad(double2 (&v)[4], const LayerArg2 &P, int o, uint32_t w) {
    const bool skip = (w >> 24) & 1u;
    if (!P.oc[o].generic) {
        const OpC c = P.oc[o];
        const double s = (double)c.s;
        if (c.shape & 1) { g_ry(v, 1.0); if (w & 63u) g_pauli(v, w & 63u); }
        if (!skip) { g_rxx(v, s); if ((w >> 6) & 63u) g_pauli(v, (w >> 6) & 63u); }
        if (c.shape & 2) { g_rx(v, -s); if ((w >> 12) & 63u) g_pauli(v, (w >> 12) & 63u); }
        if (c.shape & 4) { g_ry(v, -1.0); if ((w >> 18) & 63u) g_pauli(v, (w >> 18) & 63u); }
        return;
    }
    const s17_op &op = P.A.L.ops[o];
    for (int u = 0; u < op.n_uops; ++u) {
        const s17_uop up = op.uops[u];
        if (up.skippable && skip) continue;
        const double sg = (double)up.sign;
        const double2 p = P.A.prm[up.param & 63];
        switch (up.type) {
        case S17_U_RY: g_ry(v, sg); break;
        case S17_U_RXX: g_rxx(v, sg); break;
        case S17_U_RX: g_rx(v, sg); break;
        case S17_U_H: g_h(v); break;
        case S17_U_ROT_X: g_rot_x(v, p.x, p.y); break;
        case S17_U_ROT_Y: g_rot_y(v, p.x, p.y); break;
        case S17_U_ROT_XX: g_rot_xx(v, p.x, p.y); break;
        default: break;
        }
        if (up.ev_shift != 255) {
            const uint32_t f = (w >> up.ev_shift) & 63u;
            if (f) g_pauli(v, f);
        }
    }
}

// tail of a half cycle on one register item: in[4] indexed bc + 2 bf (data bits of tail op 0 / 1, tail ancilla = 0)
// -> out[8] indexed bc + 2 bf + 4 be
__device__ __noinline__ void tail_slow(const double2 (&in)[4], double2 (&out)[8], const LayerArg2 &P, const HalfTail &T,
                                       const uint32_t *evw, uint32_t gidx0, uint32_t mC, uint32_t mF) {
#pragma unroll
    for (int r = 0; r < 4; ++r) { out[r] = in[r]; out[r + 4] = make_double2(0.0, 0.0); }
    auto idle = [&](int k) {
        if (T.idle_word[k] == 255) return;
        const uint32_t zm = evw[T.idle_word[k]] & T.idle_filter[k];
        if (!zm) return;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const uint32_t g = gidx0 | ((r & 1) ? mC : 0u) | ((r & 2) ? mF : 0u) | ((r & 4) ? (1u << T.e_bit) : 0u);
            if (__popc(g & zm) & 1) out[r] = make_double2(-out[r].x, -out[r].y);
        }
    };
    idle(0);
    const uint32_t w0 = evw[P.oc[T.op[0]].ev_word], w1 = evw[P.oc[T.op[1]].ev_word];
#pragma unroll
    for (int bf = 0; bf < 2; ++bf) {          // op 0 acts on (bc, be)
        double2 v[4] = {out[2 * bf], out[1 + 2 * bf], out[4 + 2 * bf], out[5 + 2 * bf]};
        op_quad(v, P, T.op[0], w0);
        out[2 * bf] = v[0]; out[1 + 2 * bf] = v[1]; out[4 + 2 * bf] = v[2]; out[5 + 2 * bf] = v[3];
    }
    idle(1);
#pragma unroll
    for (int bc = 0; bc < 2; ++bc) {          // op 1 acts on (bf, be)
        double2 v[4] = {out[bc], out[bc + 2], out[bc + 4], out[bc + 6]};
        op_quad(v, P, T.op[1], w1);
        out[bc] = v[0]; out[bc + 2] = v[1]; out[bc + 4] = v[2]; out[bc + 6] = v[3];
    }
    idle(2);
}

// keeps the caller's register arrays out of local memory: the out-of-line slow path gets private copies
__device__ __forceinline__ void tail_slow_copy(const double2 (&in)[4], double2 (&out)[8], const LayerArg2 &P,
                                               const HalfTail &T, const uint32_t *evw, uint32_t gidx0, uint32_t mC,
                                               uint32_t mF) {
    double2 in2[4], out2[8];
#pragma unroll
    for (int r = 0; r < 4; ++r) in2[r] = in[r];
    tail_slow(in2, out2, P, T, evw, gidx0, mC, mF);
#pragma unroll
    for (int r = 0; r < 8; ++r) out[r] = out2[r];
}

// the same without Pauli events, leak skips, idle masks or coherent rotations: straight-line, constant-bank coefficients
__device__ __forceinline__ void tail_fast(const double2 (&in)[4], double2 (&out)[8], const FastC f0, const FastC f1,
                                          uint32_t e0, uint32_t e1) {
#pragma unroll
    for (int bf = 0; bf < 2; ++bf) {          // op 0 acts on (bc, be); the tail ancilla starts in |0>
        double2 v[4] = {in[2 * bf], in[1 + 2 * bf], make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
        g_ry(v, f0.v1); g_rxx(v, f0.s); g_rx(v, f0.sx); g_ry(v, f0.v2);
        if (e0) g_pauli(v, e0);
        out[2 * bf] = v[0]; out[1 + 2 * bf] = v[1]; out[4 + 2 * bf] = v[2]; out[5 + 2 * bf] = v[3];
    }
#pragma unroll
    for (int bc = 0; bc < 2; ++bc) {          // op 1 acts on (bf, be)
        double2 v[4] = {out[bc], out[bc + 2], out[bc + 4], out[bc + 6]};
        g_ry(v, f1.v1); g_rxx(v, f1.s); g_rx(v, f1.sx); g_ry(v, f1.v2);
        if (e1) g_pauli(v, e1);
        out[bc] = v[0]; out[bc + 2] = v[1]; out[bc + 4] = v[2]; out[bc + 6] = v[3];
    }
}

__global__ void __launch_bounds__(kL3Threads, kL3CtasPerSm)
k_half(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ LayerArg2 P, const __grid_constant__ HalfTail T,
       const __grid_constant__ MeasArgs ma, double2 *__restrict__ state, const uint32_t *__restrict__ ev,
       const uint32_t *__restrict__ act_list, const uint32_t *__restrict__ n_act_ptr, ShotAux *__restrict__ aux,
       const uint8_t *__restrict__ forced, uint32_t *__restrict__ meas_out, unsigned long long *__restrict__ counters,
       int init_basis)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double2 *tile = reinterpret_cast<double2 *>(smem_raw);
    uint32_t *evw2 = reinterpret_cast<uint32_t *>(smem_raw + (size_t)kTile * 16);                 // [2][32]
    ShotAux *auxs2 = reinterpret_cast<ShotAux *>(smem_raw + (size_t)kTile * 16 + 2 * kEvBytes);   // [2]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kTile * 16 + 2 * (kEvBytes + kAuxBytes));
    uint64_t *meta = full + 1;                                                                   // [2]
    __shared__ double s_bins[kL3Threads / 32][16];
    __shared__ double s_h[16];
    __shared__ double s_inv;
    __shared__ uint32_t s_sel;

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t n_items = *n_act_ptr;
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
    auto prefetch_meta = [&](uint32_t shot_, int slot) {
        mbar_expect_tx(&meta[slot], kEvBytes + kAuxBytes);
        tma_load_1d(evw2 + slot * S17_EV_WORDS, ev + (size_t)shot_ * S17_EV_WORDS, kEvBytes, &meta[slot]);
        tma_load_1d(auxs2 + slot, aux + shot_, kAuxBytes, &meta[slot]);
    };
    uint32_t shot_next = act_list[first];
    if (t == 0) prefetch_meta(shot_next, 0);

    // tail geometry (warp-uniform): data bits of the two tail operations
    const int pC = P.oc[T.op[0]].lb0, pF = P.oc[T.op[1]].lb0;
    const int p_lo = min(pC, pF), p_hi = max(pC, pF);
    // ancilla index (bit position in the 8-bit outcome) of the three main ancillas and the tail ancilla
    const uint32_t aj0 = P.A.L.anc_bits[0] - 9, aj1 = P.A.L.anc_bits[1] - 9, aj2 = P.A.L.anc_bits[2] - 9, aje = T.e_bit - 9;

    for (uint32_t item = first; item < last; ++item) {
        const uint32_t j = item - first;
        const uint32_t shot = shot_next;
        if (item + 1 < last) shot_next = act_list[item + 1];
        const int slot = j & 1;
        const uint32_t *evw = evw2 + slot * S17_EV_WORDS;
        const ShotAux *auxs = auxs2 + slot;
        double2 *sbase = state + (size_t)shot * kDim;

        mbar_wait(&meta[slot], (j >> 1) & 1);
        if (t == 0) {
            if (init_basis < 0) {
                // lazy collapse + reset: the only non-zero input is the surviving 512-amplitude block
                const uint32_t blk = auxs->collapsed ? 0u : auxs->anc_outcome;
                mbar_expect_tx(full, kRunBytes);
                tma_load_run(tile, &tmap, (int)(shot * (uint32_t)(kDim / 8)) + (int)blk * (kRun / 8), full);
            }
            if (item + 1 < last) prefetch_meta(shot_next, slot ^ 1);
        }
        for (int r = (init_basis < 0 ? 1 : 0); r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < kRun / kL3Threads; ++i) tile[r * kRun + t + kL3Threads * i] = make_double2(0.0, 0.0);
        }
        if (init_basis < 0) mbar_wait(full, j & 1);
        __syncthreads();
        if (init_basis >= 0 && t == 0) tile[swz(init_basis)] = make_double2(1.0, 0.0);
        if (init_basis >= 0) __syncthreads();

        // ---- main operations: passes over the tile, no scaling (the final renormalisation absorbs every factor)
        for (uint32_t p = 0; p < P.n_pass; ++p) {
            const PassC &ps = P.pass[p];
            uint32_t zm = 0, lm = 0, c0 = 0;
            if (ps.idle_word != 255) {
                zm = evw[ps.idle_word] & ps.idle_filter;
                lm = zm & 0x1FFu;
#pragma unroll
                for (int k = 0; k < 3; ++k) lm |= ((zm >> P.A.L.anc_bits[k]) & 1u) << (9 + k);
            }
            const OpC c1 = P.oc[ps.o1], c2 = P.oc[ps.kind == 1 ? ps.o2 : ps.o1];
            const uint32_t w1 = evw[c1.ev_word];
            const uint32_t w2 = ps.kind == 1 ? evw[c2.ev_word] : 0u;
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
            if (fast) {
                const FastC f1 = fast_coeffs(c1, w1), f2 = fast_coeffs(c2, w2);
                if (ps.kind == 1) {
                    if (xonly) fast_pass3<1, true>(tile, g, ps, t, f1, f2, 1.0, false, e1, e2);
                    else fast_pass3<1, false>(tile, g, ps, t, f1, f2, 1.0, false, e1, e2);
                } else {
                    if (xonly) fast_pass3<0, true>(tile, g, ps, t, f1, f2, 1.0, false, e1, 0);
                    else fast_pass3<0, false>(tile, g, ps, t, f1, f2, 1.0, false, e1, 0);
                }
            } else {
#pragma unroll 1
                for (int it = 0; it < kTile / 16 / kL3Threads; ++it)
                    slow_item(tile, &P, ps, g, t + kL3Threads * it, w1, w2, zm, lm, c0, 1.0, true);
            }
            __syncthreads();
        }

        // ---- tail operations on registers + 16-bin histogram over (main ancilla value = i, tail ancilla bit)
        bool tail_is_fast;
        uint32_t te0, te1;
        FastC tf0, tf1;
        {
            const uint32_t w0 = evw[P.oc[T.op[0]].ev_word], w1 = evw[P.oc[T.op[1]].ev_word];
            bool idle_hit = false;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (T.idle_word[k] != 255 && (evw[T.idle_word[k]] & T.idle_filter[k])) idle_hit = true;
            tail_is_fast = !(P.oc[T.op[0]].generic | P.oc[T.op[1]].generic) && !idle_hit;
            te0 = (w0 >> 25) & 63u;
            te1 = (w1 >> 25) & 63u;
            tf0 = fast_coeffs(P.oc[T.op[0]], w0);
            tf1 = fast_coeffs(P.oc[T.op[1]], w1);
        }
        double acc[16];
#pragma unroll
        for (int b = 0; b < 16; ++b) acc[b] = 0.0;
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
            const int idx0 = insert0(insert0(t + kL3Threads * i, p_lo), p_hi);        // top three bits = i
            double2 in[4], out[8];
#pragma unroll
            for (int r = 0; r < 4; ++r) in[r] = tile[swz(idx0 | ((r & 1) << pC) | ((r >> 1) << pF))];
            const uint32_t gidx0 = (idx0 & 511u) | (((idx0 >> 9) & 1u) << P.A.L.anc_bits[0]) |
                                   (((idx0 >> 10) & 1u) << P.A.L.anc_bits[1]) | (((idx0 >> 11) & 1u) << P.A.L.anc_bits[2]);
            if (tail_is_fast) tail_fast(in, out, tf0, tf1, te0, te1);
            else tail_slow_copy(in, out, P, T, evw, gidx0, 1u << pC, 1u << pF);
            double n0 = 0.0, n1 = 0.0;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                n0 = fma(out[r].x, out[r].x, fma(out[r].y, out[r].y, n0));
                n1 = fma(out[r + 4].x, out[r + 4].x, fma(out[r + 4].y, out[r + 4].y, n1));
            }
#pragma unroll
            for (int b = 0; b < 8; ++b)
                if (b == i) { acc[b] += n0; acc[b + 8] += n1; }
        }
#pragma unroll
        for (int b = 0; b < 16; ++b) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], d);
        }
        if (lane == 0) {
#pragma unroll
            for (int b = 0; b < 16; ++b) s_bins[warp][b] = acc[b];
        }
        __syncthreads();
        if (t < 16) s_h[t] = (s_bins[0][t] + s_bins[1][t]) + (s_bins[2][t] + s_bins[3][t]);
        __syncthreads();

        // ---- measurement of the four ancillas (All(Measure), CS:848): one thread, sequential, conditional
        if (t == 0) {
            const double *h = s_h;
            // bin b: bits 0-2 = main ancillas (anc_bits order), bit 3 = tail ancilla -> ancilla indices
            const uint32_t ajs[4] = {aj0, aj1, aj2, aje};
            const uint32_t nout = 8u * (uint32_t)ma.stage;      // outcomes consumed by the earlier cycles of this shot
            uint32_t a = 0, chosen = 0;                 // a: ancilla-indexed outcome, chosen: bin-indexed
            bool bad = false;
            if (ma.forced) {
                // replay: the recorded outcome bits, in ascending ancilla index; each must have non-zero probability
                uint32_t seen = 0;
                for (uint32_t jj = 0; jj < 8; ++jj) {
                    int k = -1;
                    for (int q = 0; q < 4; ++q)
                        if (ajs[q] == jj) k = q;
                    if (k < 0 || !((ma.mask >> jj) & 1u)) continue;
                    const uint32_t out = forced[(size_t)shot * S17_MAX_OUTCOMES + nout + jj] & 1u;
                    double pc = 0.0;
                    for (uint32_t b = 0; b < 16; ++b)
                        if ((b & seen) == chosen && ((b >> k) & 1u) == out) pc += h[b];
                    if (pc <= 0.0) bad = true;
                    a |= out << jj;
                    chosen |= out << k;
                    seen |= 1u << k;
                }
            } else {
                // the four Measure gates sample the joint distribution of the four ancillas; one uniform against the
                // cumulative histogram draws from exactly that distribution (ProjectQ itself picks a basis state by a
                // cumulative walk) and keeps the serial section of the CTA short
                Philox rng(ma.seed, ma.first_shot + shot, ma.stream);
                double tot = 0.0;
                for (uint32_t b = 0; b < 16; ++b) tot += h[b];
                const double u = rng.next() * tot;
                double cum = 0.0;
                chosen = 15;
                for (uint32_t b = 0; b < 16; ++b) {
                    cum += h[b];
                    if (u < cum) { chosen = b; break; }
                }
                while (chosen > 0 && h[chosen] <= 0.0) --chosen;     // rounding guard: never select an empty bin
                for (int q = 0; q < 4; ++q) a |= ((chosen >> q) & 1u) << ajs[q];
            }
            const double pa = h[chosen];
            s_inv = pa > 0.0 ? 1.0 / sqrt(pa) : 0.0;        // every deferred gate factor cancels in the renormalisation
            s_sel = chosen;
            ShotAux ax;
            ax.scale = 1.0;
            ax.anc_outcome = 0;
            ax.collapsed = 1;                               // the normalised block is stored at ancilla value 0
            aux[shot] = ax;
            // the per-shot bookkeeping (outcome log, leak override, protocol branch, decode) is latency-bound global
            // traffic: it runs for all shots at once in k_record, not on this CTA's critical path
            meas_out[2 * shot + ma.final_part] = a | (bad ? 0x100u : 0u);
        }
        __syncthreads();

        // ---- recompute only the selected block (one item per thread), renormalise, store 8 KiB
        {
            const uint32_t sel = s_sel;
            const double inv = s_inv;
            const int idx0 = insert0(insert0(t, p_lo), p_hi) | (int)((sel & 7u) << 9);
            double2 in[4], out[8];
#pragma unroll
            for (int r = 0; r < 4; ++r) in[r] = tile[swz(idx0 | ((r & 1) << pC) | ((r >> 1) << pF))];
            const uint32_t gidx0 = (idx0 & 511u) | (((idx0 >> 9) & 1u) << P.A.L.anc_bits[0]) |
                                   (((idx0 >> 10) & 1u) << P.A.L.anc_bits[1]) | (((idx0 >> 11) & 1u) << P.A.L.anc_bits[2]);
            if (tail_is_fast) tail_fast(in, out, tf0, tf1, te0, te1);
            else tail_slow_copy(in, out, P, T, evw, gidx0, 1u << pC, 1u << pF);
            const int be = (sel >> 3) & 1;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double2 v = be ? out[r + 4] : out[r];
                sbase[(idx0 & 511) | ((r & 1) << pC) | ((r >> 1) << pF)] = make_double2(v.x * inv, v.y * inv);
            }
        }
        __syncthreads();       // the tile and the selection scratch are reused by the next shot
    }
    if (t == 0) {
        atomicAdd(&counters[4], (unsigned long long)(last - first));                 // C_SWEEPS: one fused sweep per shot
        atomicAdd(&counters[8], (unsigned long long)(last - first) * (init_basis < 0 ? 2ull : 1ull));   // 8 KiB runs moved
    }
}

}  // namespace s17
