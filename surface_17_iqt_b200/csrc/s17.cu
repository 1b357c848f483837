// s17.cu -- libs17.so: hand-written sm_100a kernels + the C ABI declared in include/s17.h.
//
// Reference behaviour being replaced (citations: CS = compiled_surface_code_arbitrary_error_model.py,
// CL = calc_log_e_rate_arbitrary_error_model.py of Alexowens94/Surface_17_IQT):
//   k_layer2/3 : the Rxx/Rx/Ry gate stream of stabiliser_timestep_1..8 (CS:365-650) plus the Pauli
//                gates insert_error applies (CS:198-217), ProjectQ's per-gate sweeps fused per layer
//   k_half, k_cycle1 : the same per half cycle / per whole cycle, with the measurements
//   place_errors : generate_error_locations / _non_uniform (CL:151-181)
//   k_plan     : insert_errors / insert_error / insert_2q_error / insert_leakage decision logic
//                (CS:127-288) and the leak skip predicate (CS:293, 322), one thread per shot
//   k_measure  : All(Measure) | ancilla, leak override, reset decision (CS:848-862) and the
//                s1/s2/single/run-til-fail branching (CL:240, 320-333, 415-424, 83-98) + lookup (CS:886)
//   k_collapse : ProjectQ's collapse + renormalise, and the ancilla reset X gates (CS:859-862)
//   k_pauli    : apply_correction (CS:901-907)
//   k_readout  : logical_measurement (CS:52-93) + failure tally (CL:329-330, 433-434)
// There is NO CPU fallback in this library.
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>      // header-only: ranges show up under nsys / ncu --nvtx, cost nothing otherwise

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <algorithm>
#include <vector>

#include "s17_kernels.cuh"
#include "s17_layer2.cuh"
#include "s17_layer3.cuh"
#include "s17_engine.cuh"

namespace s17 {

enum { C_FAILED = 0, C_NOT0 = 1, C_COUNTED = 2, C_ERRORS = 3, C_SWEEPS = 4, C_ROUNDS = 5, C_NEXT_RUN = 6, C_ALIVE = 7,
       C_RUNS_MOVED = 8, C_RUNS_DONE = 9, C_COLLAPSED = 10, C_N = 11 };

// =================================================================================================
// error placement + per-shot event planning
// =================================================================================================
struct NoiseDev {
    uint32_t n_lists;
    uint32_t list_len[S17_MAX_LISTS];
    uint32_t list_off[S17_MAX_LISTS];
    uint32_t eset[S17_MAX_LISTS];
    uint32_t weighted[S17_MAX_LISTS];   // 1: cum[] holds running sums of the weights for this list
};

__device__ __forceinline__ bool mask_get(const uint32_t *m, uint32_t bit) { return (m[bit >> 5] >> (bit & 31)) & 1u; }
__device__ __forceinline__ void mask_set(uint32_t *m, uint32_t bit) { m[bit >> 5] |= 1u << (bit & 31); }

// generate_error_locations (uniform, rejection) / generate_error_locations_non_uniform (CL:151-181) for one shot:
// sets the placed slots in the (zeroed) bitset m
__device__ __forceinline__ void place_errors(const NoiseDev &nd, const double *__restrict__ cum, uint32_t *m, uint64_t seed,
                                             uint64_t shot_id)
{
    Philox rng(seed, shot_id, 1u);
    for (uint32_t t = 0; t < nd.n_lists; ++t) {
        const uint32_t L = nd.list_len[t], off = nd.list_off[t];
        const double total = nd.weighted[t] ? cum[off + L - 1] : 0.0;
        for (uint32_t e = 0; e < nd.eset[t]; ++e) {
            for (;;) {
                uint32_t ind;
                if (!nd.weighted[t]) {
                    ind = (uint32_t)(rng.next() * (double)L);
                    if (ind >= L) ind = L - 1;
                } else {
                    const double r = rng.next() * total;      // generate_from_distr: first i with cum[i] > r
                    uint32_t lo = 0, hi = L - 1;
                    while (lo < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (cum[off + mid] > r) hi = mid; else lo = mid + 1;
                    }
                    ind = lo;
                }
                if (!mask_get(m, off + ind)) { mask_set(m, off + ind); break; }
            }
        }
    }
}

struct PlanArgs {
    int n_instr;
    int noise;        // S17_NOISE_*
    int noiseless;    // prep cycle: nothing fires, only the leak skip predicate is evaluated
    int second;       // 1: second (stab_ind=1) cycle -> use_round slots are offset by len/2 (CS:191-193)
    int rec_half;     // PROB mode: which half of the mask record receives the fired bits (0/1)
    int place;        // SUBSET mode, first noisy cycle: draw the shot's fixed-weight placement here (and record it)
    uint32_t stream;
    uint64_t seed, first_shot;
    int n_shots;
};

// multiply the 6-bit field (x0 z0 x1 z1 k k) on the left by Pauli `p` (1=X 2=Y 3=Z) on qubit q (0/1)
__device__ __forceinline__ uint32_t pauli_mul(uint32_t f, int q, int p) {
    uint32_t x = (f >> (2 * q)) & 1u, z = (f >> (2 * q + 1)) & 1u, k = (f >> 4) & 3u;
    if (p == 1) { x ^= 1u; }
    else if (p == 3) { k = (k + 2u * x) & 3u; z ^= 1u; }
    else { k = (k + 1u + 2u * x) & 3u; x ^= 1u; z ^= 1u; }
    f &= ~((3u << (2 * q)) | (3u << 4));
    return f | (x << (2 * q)) | (z << (2 * q + 1)) | (k << 4);
}

// product A.B of two 2-qubit Pauli fields (x0 z0 x1 z1 k k):  i^k (X^x0 Z^z0)_data (X^x1 Z^z1)_anc
__device__ __forceinline__ uint32_t pmul2(uint32_t A, uint32_t B) {
    const uint32_t sign = ((A >> 1) & B & 1u) + ((A >> 3) & (B >> 2) & 1u);       // Z_a X_b = -X_b Z_a per qubit
    const uint32_t k = ((A >> 4) + (B >> 4) + 2u * sign) & 3u;
    return ((A ^ B) & 15u) | (k << 4);
}
// U P U^dagger for one of the cycle's Clifford gates; img_* are the images of X_d, Z_d, X_a, Z_a
__device__ __forceinline__ uint32_t pconj(uint32_t P, uint32_t img_xd, uint32_t img_zd, uint32_t img_xa, uint32_t img_za) {
    uint32_t R = P & 0x30u;
    if (P & 1u) R = pmul2(R, img_xd);
    if (P & 2u) R = pmul2(R, img_zd);
    if (P & 4u) R = pmul2(R, img_xa);
    if (P & 8u) R = pmul2(R, img_za);
    return R;
}
__device__ __forceinline__ uint32_t pconj_rx(uint32_t P, int sigma) {     // Rx(sigma pi/2) on the data qubit: Z -> -sigma Y
    return pconj(P, 0x01u, 0x03u | ((sigma > 0 ? 3u : 1u) << 4), 0x04u, 0x08u);
}
__device__ __forceinline__ uint32_t pconj_ry(uint32_t P, int v) {         // Ry(v pi/2): X -> -v Z, Z -> v X
    return pconj(P, 0x02u | ((v > 0 ? 2u : 0u) << 4), 0x01u | ((v > 0 ? 0u : 2u) << 4), 0x04u, 0x08u);
}
__device__ __forceinline__ uint32_t pconj_rxx(uint32_t P, int s) {        // Rxx(s pi/2): Z_d -> -s Y_d X_a, Z_a -> -s X_d Y_a
    const uint32_t k = (s > 0 ? 3u : 1u) << 4;
    return pconj(P, 0x01u, 0x07u | k, 0x04u, 0x0Du | k);
}

// =================================================================================================
// measurement, collapse, correction, read-out
// =================================================================================================
struct MeasArgs {
    int stage;        // 0 prep, 1 first noisy cycle, 2 second cycle
    int protocol;
    int forced;
    uint32_t mask;    // ancillas measured by this call (bit j = ancilla j)
    int final_part;   // 1: completes the cycle's All(Measure) | ancilla
    int sample_data;  // k_cycle1: also draw the basis state a data read-out of the stored block would find (k_readout_s)
    uint32_t data_stream;
    uint32_t stream;
    uint64_t seed, first_shot;
    int n_shots;
    int with_prep;    // k_cycle1 ran the noiseless preparation cycle of every shot in the same launch, ahead of this cycle: its
                      // outcomes ride in bits 16-23 (zero-probability flag: bit 24) of the shot's meas_out words
};

__device__ __forceinline__ void reset_shot(int shot, uint8_t *__restrict__ active, uint8_t *__restrict__ ready,
                                           s17_record *__restrict__ rec) {
    active[shot] = 1;
    ready[shot] = 0;
    s17_record z;
    memset(&z, 0, sizeof(z));
    rec[shot] = z;
}

// bookkeeping after the ancillas in ma.mask of `shot` were measured with raw outcome bits `a`: outcome log, and on the
// final part of a cycle the leak override (CS:854-857), the syndrome comparison and the protocol's branch
// (CL:240, 320-333, 415-424, 83-98) plus the lookup decode (CS:886-892)
__device__ void record_measurement(const MeasArgs &ma, int shot, uint32_t a, uint32_t nout, bool bad,
                                   const uint32_t *__restrict__ lut, uint32_t *__restrict__ leak,
                                   uint8_t *__restrict__ active, uint8_t *__restrict__ ready, s17_record *__restrict__ rec,
                                   unsigned long long *__restrict__ counters)
{
    s17_record r = rec[shot];
    for (int j = 0; j < 8; ++j)
        if (((ma.mask >> j) & 1u) && nout + j < S17_MAX_OUTCOMES) r.outcomes[nout + j] = (uint8_t)((a >> j) & 1u);
    if (bad) { r.flags |= S17_F_ERROR; atomicAdd(&counters[C_ERRORS], 1ull); }
    const uint32_t all = r.partial | a;
    if (!ma.final_part) {
        r.partial = (uint16_t)all;
        rec[shot] = r;
        return;
    }
    r.partial = 0;
    r.n_outcomes = (uint8_t)min(nout + 8u, (uint32_t)S17_MAX_OUTCOMES);
    // a leaked ancilla always reads 1 and is reset (CS:854-857)
    uint32_t lk = leak[shot];
    const uint32_t synd = all | ((lk >> 9) & 0xFFu);
    lk &= 0x1FFu;
    leak[shot] = lk;
    r.leak = lk;

    bool decode = false;
    if (ma.stage == 0) {
        r.quiescent = (uint8_t)synd;
    } else if (ma.stage == 1) {
        r.synd1 = (uint8_t)synd;
        r.flips_a = r.quiescent ^ (uint8_t)synd;                  // (prev - synd) % 2, CL:240, 320, 415
        const bool not0 = r.flips_a != 0;
        if (not0) { r.flags |= S17_F_NOT0; atomicAdd(&counters[C_NOT0], 1ull); }
        if (ma.protocol == S17_PROTO_SINGLE) decode = true;
        else if (ma.protocol == S17_PROTO_S1) { decode = !not0; if (not0) active[shot] = 0; }      // CL:321, 332
        else if (ma.protocol == S17_PROTO_S2) { if (!not0) active[shot] = 0; }                    // CL:416, 437
        else { decode = !not0; if (!not0) active[shot] = 0; }                                     // CL:86-92
    } else {
        r.synd2 = (uint8_t)synd;
        r.flags |= S17_F_SYND2;
        decode = true;
    }
    if (decode) {
        r.ft_syndrome = r.quiescent ^ (uint8_t)synd;             // CL:424; equals (flips_a + flips_b) % 2, CL:98
        r.correction = lut[r.ft_syndrome];
        r.flags |= S17_F_COUNTED;
        ready[shot] = 1;
    }
    rec[shot] = r;
}

// the bookkeeping of a whole cycle from the two meas_out words of a shot (k_half / k_cycle1 / k_prep write them); with
// ma.with_prep the preparation cycle's first
__device__ __forceinline__ void record_cycle(const MeasArgs &ma, int shot, const uint32_t *__restrict__ meas_out,
                                             const uint32_t *__restrict__ lut, uint32_t *__restrict__ leak,
                                             uint8_t *__restrict__ active, uint8_t *__restrict__ ready,
                                             s17_record *__restrict__ rec, unsigned long long *__restrict__ counters)
{
    const uint32_t xz = meas_out[2 * shot] | meas_out[2 * shot + 1];
    if (ma.with_prep) {
        MeasArgs m0 = ma;
        m0.stage = 0;
        record_measurement(m0, shot, (xz >> 16) & 0xFFu, 0u, ((xz >> 24) & 1u) != 0, lut, leak, active, ready, rec, counters);
    }
    record_measurement(ma, shot, xz & 0xFFu, 8u * (uint32_t)ma.stage, (xz & 0x100u) != 0, lut, leak, active, ready, rec, counters);
}

// one warp per shot: sample the ancillas in `mask` from the joint histogram, then (on the final part of a
// cycle) apply the leak override and take the protocol's branch
__global__ void __launch_bounds__(128)
k_measure(MeasArgs ma, const double *__restrict__ hist, const uint8_t *__restrict__ forced,
          const uint32_t *__restrict__ lut, uint32_t *__restrict__ leak, uint8_t *__restrict__ active,
          uint8_t *__restrict__ ready, ShotAux *__restrict__ aux, s17_record *__restrict__ rec,
          unsigned long long *__restrict__ counters)
{
    const int shot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (shot >= ma.n_shots || !active[shot]) return;
    const double *h = hist + (size_t)shot * 256;
    // bins outside the measured set are not part of the support (their ancillas are reset to 0)
    double hv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t m = lane + 32 * i;
        hv[i] = (m & ~ma.mask) ? 0.0 : h[m];
    }
    const uint32_t nout = rec[shot].n_outcomes;
    Philox rng(ma.seed, ma.first_shot + shot, ma.stream);
    uint32_t a = 0, seen = 0;
    bool bad = false;
    // one Measure per ancilla in order (All(Measure) | ancilla, CS:848): conditional on earlier outcomes
    for (int j = 0; j < 8; ++j) {
        if (!((ma.mask >> j) & 1u)) continue;
        double p0 = 0.0, p1 = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t m = lane + 32 * i;
            if ((m & seen) != a) continue;
            if ((m >> j) & 1u) p1 += hv[i]; else p0 += hv[i];
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            p0 += __shfl_xor_sync(0xffffffffu, p0, d);
            p1 += __shfl_xor_sync(0xffffffffu, p1, d);
        }
        uint32_t out;
        if (ma.forced) {
            out = forced[(size_t)shot * S17_MAX_OUTCOMES + nout + j] & 1u;
            if ((out ? p1 : p0) <= 0.0) bad = true;
        } else {
            double u = rng.next();
            u = __shfl_sync(0xffffffffu, u, 0);
            out = (u * (p0 + p1) < p1) ? 1u : 0u;
        }
        a |= out << j;
        seen |= 1u << j;
    }
    if (lane != 0) return;
    const double pa = h[a];
    ShotAux ax;
    ax.scale = pa > 0.0 ? 1.0 / sqrt(pa) : 0.0;
    ax.anc_outcome = a;          // the surviving block; the ancillas measured earlier were already reset to 0
    ax.collapsed = 0;
    aux[shot] = ax;
    record_measurement(ma, shot, a, nout, bad, lut, leak, active, ready, rec, counters);
}

// bookkeeping for the fused half-cycle path: both halves of a cycle were sampled inside k_half
__global__ void k_record(MeasArgs ma, const uint32_t *__restrict__ meas_out, const uint32_t *__restrict__ lut,
                         uint32_t *__restrict__ leak, uint8_t *__restrict__ active, uint8_t *__restrict__ ready,
                         s17_record *__restrict__ rec, unsigned long long *__restrict__ counters)
{
    const int shot = blockIdx.x * blockDim.x + threadIdx.x;
    if (shot >= ma.n_shots || !active[shot]) return;
    record_cycle(ma, shot, meas_out, lut, leak, active, ready, rec, counters);
}

// =================================================================================================
// k_plan: one launch between two cycle kernels
// =================================================================================================
// One thread per shot; a warp stages its 32 shots' slot bitsets and event words through shared memory (row stride 33
// words) so that the global traffic is coalesced rows instead of 32 strided 4-byte accesses per instruction.  Per shot:
//   1. the bookkeeping of the cycle that just ended (StepArgs::do_record; the body of k_record: outcome log, leak
//      override, protocol branch, lookup decode), so that a step needs no separate launch for it;
//   2. error insertion for the cycle that starts (insert_errors / insert_error, CS:127-234) -> one event word per
//      operation + the frame words of the whole-cycle kernel.  The program is walked by GROUP (one entangling operation,
//      or the idle slots of one gap between timesteps, in program order).  A shot first finds which entries fire -- its
//      placed / replayed slot bits, or, with true probabilities, the position of the next firing entry sampled from the
//      survival products of the program (geometric skipping: P(first fire at i) = S_i p_i, S_i = prod_{j<i} (1 - p_j);
//      one uniform per event instead of one per entry) -- and marks them in a bitmap over the INSTRUCTIONS of the
//      program; an operation with a fired entry then visits just those instructions (in program order) and its end;
//      every other operation takes its default event word (the data-flip bit of the frame).  A set leak flag, or a
//      fired leakage entry, makes the shot walk every instruction of every group (the skip predicate of every later
//      operation may change), as does a program in which two entries share a slot;
//   3. the list of shots the cycle kernel will run (StepArgs::do_compact).
constexpr int kPlanThreads = 128;
constexpr int kPlanRow = 33;
constexpr int kPlanMaxInstr = 640;
constexpr int kPlanMaxGroups = 32;
constexpr int kPlanFiWords = kPlanMaxInstr / 32;       // fired-instruction bitmap of one shot
constexpr int kPlanFiRow = kPlanFiWords + 1;           // odd row stride: conflict-free
constexpr size_t kPlanSmem = (size_t)(2 * (kPlanThreads / 32) * 32 * kPlanRow + kPlanThreads * kPlanFiRow +
                                      kPlanMaxInstr * 3 + kPlanMaxInstr + 2 * kPlanMaxGroups) * 4 +
                             (size_t)(kPlanMaxInstr + 2) * 8;             // + the survival products (true-probability noise)
struct StepArgs {
    int do_record;                 // 1: first apply the previous cycle's measurement outcomes (rec_ma.stage) to the shot
    MeasArgs rec_ma;
    const uint32_t *meas_out;
    const uint32_t *lut;
    uint8_t *ready;
    s17_record *rec;
    unsigned long long *counters;
    int do_compact;                // 1: append the shots the cycle kernel runs to `list` (*n_act must be 0 on entry)
    uint32_t *list;
    const uint8_t *leaf;           // nullptr, or per shot: byte for bits 24-31 of its entry (k_cycle1 + prep cache)
    const uint8_t *quiet_ok;       // nullptr, or: event-free shots in a leaf with quiet_ok[leaf] are left out (k_quiet)
    const uint8_t *done;           // nullptr, or per shot: 1 = the memoised preparation handled it
    int skip_done;
    int walk_all;                  // 1: two entries share a slot somewhere in the program -> no per-entry shortcuts
    int do_reset;                  // 1: first launch of a call: fresh flags, record (and leak register unless keep_leak)
    int keep_leak;
    NoiseDev nd;                   // PlanArgs::place: the subset to place
    const double *cum;
};
// hot word of an instruction: bits 0-15 absolute slot, 16 in_skip, 17 entry (S17_P_ERR / S17_P_IDLE), 18-22 group,
// 23 leakage entry.  group words: [g] bits 0-9 first instruction, 10-19 one past the last; [32 + g] bit 0 gap (idle)
// group, 1-5 event word, 6-9 data bit, 10 first operation of its half cycle, 11 half index, 12 Hadamard frame, 13 the
// operation closes the Ry bracket of its data qubit (Ry(-)), 14 it opens it (Ry(+)), 15 it has an Rx, 16 Rxx sign is +,
// 17-22 offset of the operation's leak test (S17_P_SKIP_TEST) from its first instruction
__device__ __forceinline__ uint32_t plan_apply_error(uint32_t f, const s17_pinstr &in, uint32_t &lk, Philox &rng) {
    switch (in.err) {
    case S17_ERR_X: f = pauli_mul(f, 0, 1); break;
    case S17_ERR_Y: f = pauli_mul(f, 0, 2); break;
    case S17_ERR_Z: f = pauli_mul(f, 0, 3); break;
    case S17_ERR_ZC: f = pauli_mul(f, 0, 3); break;
    case S17_ERR_ZT: f = pauli_mul(f, 1, 3); break;
    case S17_ERR_XX: f = pauli_mul(pauli_mul(f, 0, 1), 1, 1); break;
    case S17_ERR_ZZ: f = pauli_mul(pauli_mul(f, 0, 3), 1, 3); break;
    case S17_ERR_LEAK_A: lk |= 1u << in.a; break;
    case S17_ERR_LEAK_AB: lk |= (1u << in.a) | (1u << in.b); break;
    case S17_ERR_DEPOL: {
        // insert_2q_error (CS:237-283): 15 equiprobable non-identity pairs, first letter -> qubits[0]
        int k = (int)(rng.next() * 15.0);
        if (k > 14) k = 14;
        const int code = k + 1, p0 = code >> 2, p1 = code & 3;   // IX,IY,IZ,XI,...,ZZ
        if (p0) f = pauli_mul(f, 0, p0);
        if (p1) f = pauli_mul(f, 1, p1);
    } break;
    default: break;
    }
    return f;
}
// net Pauli at the END of an operation (bits 25-30 of its event word): every event conjugated through the gates that
// follow it; exact because Rxx/Rx/Ry(+-pi/2) are Clifford.  The fast sweep path applies the gates branch-free and this
// single Pauli afterwards; the per-gate fields (bits 0-23) stay for the interpreter paths.
__device__ __forceinline__ uint32_t plan_net_pauli(uint32_t w, uint32_t skip, bool ry1, bool rx, bool ry2, int s) {
    if (!(w & 0xFFFFFFu)) return w;
    uint32_t N = 0;
    if (ry1) N = w & 63u;                                                 // after Ry(+)
    if (!skip) { N = pconj_rxx(N, s); N = pmul2((w >> 6) & 63u, N); }
    if (rx) { N = pconj_rx(N, -s); N = pmul2((w >> 12) & 63u, N); }
    if (ry2) { N = pconj_ry(N, -1); N = pmul2((w >> 18) & 63u, N); }
    return w | (N << 25);
}

__global__ void __launch_bounds__(kPlanThreads)
k_plan(const s17_pinstr *__restrict__ prog, PlanArgs pa, const __grid_constant__ StepArgs sa, const double *__restrict__ probs,
       uint32_t *__restrict__ ev, uint32_t *__restrict__ masks, uint32_t *__restrict__ leak,
       uint8_t *__restrict__ active, uint32_t *__restrict__ frm, uint32_t *__restrict__ n_act /* this cycle's list length */,
       uint32_t *__restrict__ n_act_next /* zeroed here for the launch after this one */,
       const uint32_t *__restrict__ hot, uint8_t *__restrict__ quiet /* nullptr, or per shot: 1 = no entry fired and no
       qubit is leaked in this cycle (event-free: k_quiet replays it from the cache) */,
       const uint32_t *__restrict__ grp, int n_grp, const uint16_t *__restrict__ bit2instr /* slot bit -> instruction */,
       const double *__restrict__ surv /* nullptr, or survival products S_0 .. S_n of this cycle's entries */)
{
    extern __shared__ uint32_t plan_smem[];
    uint32_t (*s_mask)[32][kPlanRow] = reinterpret_cast<uint32_t (*)[32][kPlanRow]>(plan_smem);
    uint32_t (*s_ev)[32][kPlanRow] = reinterpret_cast<uint32_t (*)[32][kPlanRow]>(plan_smem + (kPlanThreads / 32) * 32 * kPlanRow);
    uint32_t *s_fi = plan_smem + 2 * (kPlanThreads / 32) * 32 * kPlanRow;       // [kPlanThreads][kPlanFiRow]
    uint32_t *s_prog = s_fi + kPlanThreads * kPlanFiRow;   // the program of this cycle (slots resolved to absolute bit indices)
    uint32_t *s_hot = s_prog + kPlanMaxInstr * 3;
    uint32_t *s_grp = s_hot + kPlanMaxInstr;               // [0..31] instruction range, [32..63] operation word
    double *s_surv = reinterpret_cast<double *>(s_grp + 2 * kPlanMaxGroups);   // (the words before it are an even count)
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_act_next = 0;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(prog);
        for (int i = threadIdx.x; i < pa.n_instr * 3; i += blockDim.x) s_prog[i] = src[i];
        for (int i = threadIdx.x; i < pa.n_instr; i += blockDim.x) s_hot[i] = hot[i];
        if (threadIdx.x < 2 * kPlanMaxGroups) s_grp[threadIdx.x] = (threadIdx.x & (kPlanMaxGroups - 1)) < n_grp ? grp[threadIdx.x] : 0u;
        if (surv)
            for (int i = threadIdx.x; i <= pa.n_instr; i += blockDim.x) s_surv[i] = surv[i];
        __syncthreads();
    }
    const bool prob_mode = pa.noise == S17_NOISE_PROB;
    const int fi_words = (pa.n_instr + 31) >> 5;
    uint32_t gap_mask = 0;
    for (int gi = 0; gi < n_grp; ++gi) gap_mask |= (s_grp[kPlanMaxGroups + gi] & 1u) << gi;
    // persistent over groups of kPlanThreads shots: the program is staged once per CTA (no CTA-wide barrier below)
    for (int base = blockIdx.x * kPlanThreads; base < pa.n_shots; base += gridDim.x * kPlanThreads) {
    const int shot = base + threadIdx.x, shot0 = shot - lane;
    if (sa.do_reset && shot < pa.n_shots) {
        if (!sa.keep_leak) leak[shot] = 0;
        reset_shot(shot, active, sa.ready, sa.rec);
    }
    bool act = shot < pa.n_shots && (sa.do_reset || active[shot]);
    if (sa.do_record && act) {
        record_cycle(sa.rec_ma, shot, sa.meas_out, sa.lut, leak, active, sa.ready, sa.rec, sa.counters);
        act = active[shot] != 0;                       // the protocol may have retired the shot (CL:321, 416, 86)
    }
    // preparation stage next to the memoised one (k_prep): only the shots it left (a set leak flag) run a cycle
    if (act && sa.skip_done && sa.done && sa.done[shot]) act = false;
    const uint32_t actmask = __ballot_sync(0xffffffffu, act);
    if (!actmask) continue;
    // the shots' slot bitsets: placed / replayed ones are loaded, eight rows in flight; sampled or placed here: start empty
    const bool load_masks = !prob_mode && !pa.place && !pa.noiseless;
#pragma unroll 1
    for (int r0 = 0; r0 < 32; r0 += 8) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            v[k] = (load_masks && ((actmask >> (r0 + k)) & 1u)) ? masks[(size_t)(shot0 + r0 + k) * (2 * S17_MASK_WORDS) + lane] : 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) s_mask[wib][r0 + k][lane] = v[k];
    }
    __syncwarp();
    uint32_t *evs = s_ev[wib][lane];
    uint32_t *mk = s_mask[wib][lane];                  // placed slots (subset / replay) or the slots fired now (prob mode)
    uint32_t *fi = s_fi + threadIdx.x * kPlanFiRow;
    uint32_t lk = 0;
    Philox rng(pa.seed, pa.first_shot + shot, pa.stream);
    uint32_t f_flip = 0, f_sign = 0, f_phase = 0, f_ancz = 0;   // frame bookkeeping of the current half cycle (k_cycle)
    int f_half = -1;
    if (act) lk = leak[shot];
    bool any_event = lk != 0;
    if (act && pa.noiseless && lk == 0) {
        // noiseless preparation cycle without a leaked qubit: no event, no leak skip, identity frame
        for (int i = 0; i < (int)S17_EV_WORDS; ++i) evs[i] = 0;
        frm[(size_t)shot * 4] = 0;
        frm[(size_t)shot * 4 + 1] = 0;
    } else if (act) {
    for (int i = 24; i < (int)S17_EV_WORDS; ++i) evs[i] = 0;
    // ---- which entries fire ----
    if (pa.place) place_errors(sa.nd, sa.cum, mk, pa.seed, pa.first_shot + shot);
    if (prob_mode && !pa.noiseless) {
        if (surv) {
            int pos = 0;
            const double s_end = s_surv[pa.n_instr];
            for (;;) {
                const double target = rng.next() * s_surv[pos];
                if (!(s_end < target)) break;                        // nothing fires in the rest of the cycle
                int lo = pos, hi = pa.n_instr - 1;                   // smallest i >= pos with S_{i+1} < target
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (s_surv[mid + 1] < target) hi = mid; else lo = mid + 1;
                }
                mask_set(mk, s_hot[lo] & 0xFFFFu);
                pos = lo + 1;
            }
        } else {
            for (int i = 0; i < pa.n_instr; ++i) {                   // one Bernoulli draw per entry (CS:198)
                const uint32_t hw = s_hot[i];
                if ((hw & 0x20000u) && rng.next() < probs[hw & 0xFFFFu]) mask_set(mk, hw & 0xFFFFu);
            }
        }
    }
    // fired slots -> fired instructions, the groups they sit in, whether one of them is a leakage entry
    uint32_t flagged = 0, leak_event = 0;
    for (int i = 0; i < fi_words; ++i) fi[i] = 0;
    if (!pa.noiseless) {
        for (int wi = 0; wi < (int)S17_MASK_WORDS; ++wi) {
            uint32_t word = mk[wi];
            while (word) {
                const uint32_t ii = bit2instr[32 * wi + __ffs(word) - 1];
                word &= word - 1;
                if (ii == 0xFFFFu) continue;                         // a slot of the other cycle
                const uint32_t hw = s_hot[ii];
                flagged |= 1u << ((hw >> 18) & 31u);
                leak_event |= (hw >> 23) & 1u;
                fi[ii >> 5] |= 1u << (ii & 31u);
            }
        }
    }
    any_event = any_event || flagged != 0;
    const bool walk = lk != 0 || leak_event != 0 || sa.walk_all;
    const uint32_t todo = walk ? 0xFFFFFFFFu : flagged;
    if (!sa.walk_all) {
        // ---- pass 1: per operation, its fired instructions in program order, then its net Pauli.  Without a leaked
        // qubit only the operations with a fired entry are visited (each lane runs over ITS OWN operations); with one
        // (or with a leakage entry firing in this cycle) every operation is, because its leak test (evaluated right
        // before its Rxx, after the errors of a preceding Ry: CS:293, 322) may skip the Rxx block and its entries.
        uint32_t td = (walk ? (n_grp >= 32 ? 0xFFFFFFFFu : ((1u << n_grp) - 1u)) : flagged) & ~gap_mask;
        while (td) {
            const int gi = __ffs(td) - 1;
            td &= td - 1;
            const uint32_t gm = s_grp[gi], g2 = s_grp[kPlanMaxGroups + gi];
            const int first = (int)(gm & 1023u), last = (int)((gm >> 10) & 1023u);
            const int sk = first + (int)((g2 >> 17) & 63u);              // the operation's leak test
            uint32_t w = 0, skip = 0;
            bool tested = !walk;                                         // no leaked qubit: the test is negative
            for (int wi = first >> 5; wi <= (last - 1) >> 5; ++wi) {
                uint32_t bits = fi[wi];
                if (wi == (first >> 5)) bits &= 0xFFFFFFFFu << (first & 31);
                if (wi == ((last - 1) >> 5) && (last & 31)) bits &= 0xFFFFFFFFu >> (32 - (last & 31));
                while (bits) {
                    const int ii = 32 * wi + __ffs(bits) - 1;
                    bits &= bits - 1;
                    if (!tested && ii > sk) {
                        const s17_pinstr t = reinterpret_cast<const s17_pinstr *>(s_prog)[sk];
                        skip = ((lk >> t.a) | (lk >> t.b)) & 1u;
                        tested = true;
                    }
                    if (((s_hot[ii] >> 16) & 1u) & skip) continue;       // an entry inside a skipped Rxx block
                    const s17_pinstr in = reinterpret_cast<const s17_pinstr *>(s_prog)[ii];
                    const uint32_t f = plan_apply_error((w >> in.field) & 63u, in, lk, rng);
                    w = (w & ~(63u << in.field)) | (f << in.field);
                }
            }
            if (!tested) {
                const s17_pinstr t = reinterpret_cast<const s17_pinstr *>(s_prog)[sk];
                skip = ((lk >> t.a) | (lk >> t.b)) & 1u;
            }
            w |= skip << 24;
            evs[(g2 >> 1) & 31u] = plan_net_pauli(w, skip, (g2 >> 14) & 1u, (g2 >> 15) & 1u, (g2 >> 13) & 1u, ((g2 >> 16) & 1u) ? 1 : -1);
        }
    } else {
        // ---- pass 1 for a program in which two entries share a slot: every instruction of every group, in program
        // order (one instruction per trip of a single loop)
        uint32_t td = todo & (n_grp >= 32 ? 0xFFFFFFFFu : ((1u << n_grp) - 1u)) & ~gap_mask;
        int i = 0, end = 0;
        uint32_t w = 0, skip = 0;
        for (;;) {
            if (i == end) {
                if (!td) break;
                const uint32_t gm = s_grp[__ffs(td) - 1];
                td &= td - 1;
                i = (int)(gm & 1023u);
                end = (int)((gm >> 10) & 1023u);
            }
            const uint32_t hw = s_hot[i];
            const int ii = i++;
            if (hw & 0x20000u) {
                // a fired entry acts unless it sits inside a skipped Rxx block (CS:293, 322).  The slot is the absolute index
                // of the entry in this cycle (list offset + slot + half the list in a second cycle, CS:191-193).
                if ((((hw >> 16) & 1u) & skip) || pa.noiseless) continue;
                if (!mask_get(mk, hw & 0xFFFFu)) continue;
            }
            const s17_pinstr in = reinterpret_cast<const s17_pinstr *>(s_prog)[ii];
            if (in.kind == S17_P_OP_BEGIN) { w = 0; skip = 0; continue; }
            if (in.kind == S17_P_SKIP_TEST) {
                // evaluated AFTER the errors of a preceding Ry (which may just have leaked the data qubit)
                skip = ((lk >> in.a) | (lk >> in.b)) & 1u;
                w |= skip << 24;
                continue;
            }
            if (in.kind == S17_P_OP_END) {
                evs[in.word] = plan_net_pauli(w, skip, in.a & 1, in.a & 2, in.a & 4, in.b ? 1 : -1);
                continue;
            }
            const uint32_t f = plan_apply_error((w >> in.field) & 63u, in, lk, rng);
            w = (w & ~(63u << in.field)) | (f << in.field);
        }
    }
    // ---- pass 2: every group in program order (warp-uniform trip count): the net Pauli of each operation in the frame
    // where its half cycle is diagonal (s17_diag.cuh) -- a data X relabels the basis (flip mask; bit 31 tells the
    // operation whether its own label bit is flipped), a data Z is a label-dependent sign (sign mask), the rest is a
    // global phase -- and the idle entries of the gaps (sympathetic_cooling, CS:762-772).
    for (int gi = 0; gi < n_grp; ++gi) {
        const uint32_t gm = s_grp[gi], g2 = s_grp[kPlanMaxGroups + gi];
        const bool walked = (todo >> gi) & 1u;
        if (g2 & 1u) {                                 // gap between two timesteps: its fired idle entries
            if (!walked || pa.noiseless) continue;
            for (int i = (int)(gm & 1023u); i < (int)((gm >> 10) & 1023u); ++i) {
                const uint32_t hw = s_hot[i];
                if (!mask_get(mk, hw & 0xFFFFu)) continue;
                const s17_pinstr in = reinterpret_cast<const s17_pinstr *>(s_prog)[i];
                evs[in.word] |= 1u << in.a;            // the 17-bit idle masks of the sweep kernels / k_half
                if (in.b & 0x80u) {                    // the same Z in the frame of its half cycle (whole-cycle kernels)
                    if (in.a < 9u) {
                        const uint32_t bit = 1u << in.a;
                        if (in.b & 2u) f_flip ^= bit;                                   // H Z H = X
                        else if (in.b & 1u) { f_flip ^= bit; f_phase += 2u; }           // Ry(-pi/2) Z Ry(+pi/2) = -X
                        else { f_sign ^= bit; f_phase += 2u * ((f_flip >> in.a) & 1u); }
                    } else if (in.field != 255u) {
                        f_ancz ^= 1u << in.field;      // Z on the ancilla after that operation (and its own events)
                    }
                }
            }
            continue;
        }
        const uint32_t word = (g2 >> 1) & 31u, d = (g2 >> 6) & 15u;
        uint32_t w = walked ? evs[word] : 0u;
        if ((g2 >> 10) & 1u) {                         // first operation of a half: close the previous one (idle events may
            if (f_half >= 0)                           // follow its last operation), start from the identity frame
                frm[(size_t)shot * 4 + f_half] = f_flip | (f_sign << 9) | ((f_phase & 3u) << 18) | (f_ancz << 20);
            f_flip = 0; f_sign = 0; f_phase = 0; f_ancz = 0;
            f_half = (int)((g2 >> 11) & 1u);
        }
        w |= ((f_flip >> d) & 1u) << 31;
        const uint32_t N = (w >> 25) & 63u;
        if (N) {
            uint32_t F = N;
            if ((g2 >> 12) & 1u) {                     // H X^x Z^z H = (-1)^{xz} X^z Z^x
                const uint32_t x = N & 1u, z = (N >> 1) & 1u, k = (N >> 4) & 3u;
                F = (N & 0x0Cu) | z | (x << 1) | (((k + 2u * (x & z)) & 3u) << 4);
            } else if (!((g2 >> 13) & 1u)) {           // still inside the Ry bracket of its data qubit
                F = pconj_ry(N, -1);
            }
            if (F & 2u) { f_sign ^= 1u << d; f_phase += 2u * ((f_flip >> d) & 1u); }
            if (F & 1u) f_flip ^= 1u << d;
            f_phase += (F >> 4) & 3u;
        }
        evs[word] = w;
    }
    if (f_half >= 0) frm[(size_t)shot * 4 + f_half] = f_flip | (f_sign << 9) | ((f_phase & 3u) << 18) | (f_ancz << 20);
    leak[shot] = lk;
    }
    if (quiet && act) quiet[shot] = any_event ? 0 : 1;
    __syncwarp();
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
        if (!((actmask >> r) & 1u)) continue;
        ev[(size_t)(shot0 + r) * S17_EV_WORDS + lane] = s_ev[wib][r][lane];
        if ((prob_mode && !pa.noiseless) || pa.place)
            masks[(size_t)(shot0 + r) * (2 * S17_MASK_WORDS) + (pa.rec_half ? S17_MASK_WORDS : 0) + lane] = s_mask[wib][r][lane];
    }
    if (sa.do_compact) {
        // the shots that take part in the cycle kernel's sweep.  Warp-aggregated atomic append: the order of the list
        // is arbitrary (it only decides which warp processes which shot).
        bool flag = act;
        if (flag && quiet && !any_event && sa.quiet_ok[sa.leaf[shot]]) flag = false;          // k_quiet handles it
        const uint32_t ballot = __ballot_sync(0xffffffffu, flag);
        if (ballot) {
            uint32_t pos = 0;
            if (lane == 0) pos = atomicAdd(n_act, (uint32_t)__popc(ballot));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            if (flag) {
                uint32_t entry = (uint32_t)shot;
                if (sa.leaf) entry |= (sa.done && !sa.done[shot]) ? 0x800000u : ((uint32_t)sa.leaf[shot] << 24);   // own state / cached leaf
                sa.list[pos + __popc(ballot & ((1u << lane) - 1u))] = entry;
            }
        }
    }
    __syncwarp();                                      // the staging rows are reused by the next group
    }
}

}  // namespace s17
#include "s17_half.cuh"
#include "s17_diag.cuh"
namespace s17 {

// ProjectQ collapse + renormalise for the 8 measured ancillas, fused with the reset X gates (CS:859-862):
// the surviving 512-amplitude block moves to ancilla value 0, everything else becomes zero.
__global__ void __launch_bounds__(256)
k_collapse(double2 *__restrict__ state, ShotAux *__restrict__ aux, const uint8_t *__restrict__ flag,
           unsigned long long *__restrict__ counters)
{
    const int shot = blockIdx.x / 32, part = blockIdx.x % 32, t = threadIdx.x;
    if (!flag[shot]) return;
    const uint32_t a = aux[shot].anc_outcome;
    const double sc = aux[shot].scale;
    const int owner = a >> 3;
    double2 *s = state + (size_t)shot * kDim;
    double2 keep[2];
    if (part == owner) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const double2 v = s[a * kRun + t + 256 * i];
            keep[i] = make_double2(v.x * sc, v.y * sc);
        }
    }
    for (int b = 0; b < 8; ++b) {
        const int blk = part * 8 + b;
        if (blk == 0 && part != owner) continue;
#pragma unroll
        for (int i = 0; i < 2; ++i) s[blk * kRun + t + 256 * i] = make_double2(0.0, 0.0);
    }
    if (part == owner) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 2; ++i) s[t + 256 * i] = keep[i];
        if (t == 0) { aux[shot].collapsed = 1; atomicAdd(&counters[C_COLLAPSED], 1ull); }
    }
}

// apply_correction (CS:901-907): per data qubit X then Z  =>  out[m] = (-1)^{popc(x&z)} (-1)^{popc((m^x)&z)} in[m^x]
__global__ void __launch_bounds__(256)
k_pauli(double2 *__restrict__ state, const s17_record *__restrict__ rec, const uint8_t *__restrict__ flag)
{
    const int shot = blockIdx.x / 32, part = blockIdx.x % 32, t = threadIdx.x;
    if (!flag[shot]) return;
    const uint32_t c = rec[shot].correction;
    const uint32_t x = c & 0x1FFu, z = (c >> 9) & 0x1FFu;
    if (!(x | z)) return;
    double2 *s = state + (size_t)shot * kDim;
    const double g = (__popc(x & z) & 1) ? -1.0 : 1.0;
    if (x == 0) {
        for (int i = 0; i < kDim / 32 / 256; ++i) {
            const uint32_t m = part * (kDim / 32) + t + 256 * i;
            const double sg = (__popc(m & z) & 1) ? -g : g;
            const double2 v = s[m];
            s[m] = make_double2(sg * v.x, sg * v.y);
        }
        return;
    }
    const int p = __ffs(x) - 1;
    for (int i = 0; i < kDim / 64 / 256; ++i) {
        const uint32_t n = part * (kDim / 64) + t + 256 * i;
        const uint32_t m0 = insert0((int)n, p), m1 = m0 ^ x;
        const double2 v0 = s[m0], v1 = s[m1];
        const double s0 = (__popc(m1 & z) & 1) ? -g : g;   // out[m0] takes in[m1]
        const double s1 = (__popc(m0 & z) & 1) ? -g : g;
        s[m0] = make_double2(s0 * v1.x, s0 * v1.y);
        s[m1] = make_double2(s1 * v0.x, s1 * v0.y);
    }
}

struct ReadArgs {
    int forced;
    int folded;        // 1: the X part of the correction is folded into the read-out instead of the state
    int logical_state;
    uint32_t stream;
    uint64_t seed, first_shot;
    int n_shots;
};

// logical_measurement (CS:52-93): one warp per shot
__global__ void __launch_bounds__(128)
k_readout(ReadArgs ra, const double2 *__restrict__ state, size_t stride, const ShotAux *__restrict__ aux,
          const uint8_t *__restrict__ forced, const uint8_t *__restrict__ cw16, const uint32_t *__restrict__ leak,
          const uint8_t *__restrict__ ready, s17_record *__restrict__ rec, unsigned long long *__restrict__ counters)
{
    const int shot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (shot >= ra.n_shots || !ready[shot]) return;
    const ShotAux ax = aux[shot];
    const uint32_t sel = ax.collapsed ? 0u : ax.anc_outcome;
    const double sc2 = ax.collapsed ? 1.0 : ax.scale * ax.scale;
    const uint32_t corr = rec[shot].correction;
    const uint32_t xc = ra.folded ? (corr & 0x1FFu) : 0u;
    const double2 *blk = state + (size_t)shot * stride + (size_t)sel * kRun;
    double p[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t d = lane + 32 * i;
        const double2 v = blk[d ^ xc];       // probability of reading d from the corrected state
        p[i] = (v.x * v.x + v.y * v.y) * sc2;
    }
    // All(Measure) | data (CS:56): the nine Measure gates sample one basis state from |psi|^2 (ProjectQ itself picks a
    // basis state by a cumulative walk per Measure).  One uniform against the cumulative distribution in
    // the order (lane, register): a lane-local sum, an inclusive scan over the lanes, a walk inside the selected lane.
    uint32_t chosen = 0;
    const uint32_t nout = rec[shot].n_outcomes;
    bool bad = false;
    if (ra.forced) {
        const uint32_t b = lane < 9 ? (uint32_t)(forced[(size_t)shot * S17_MAX_OUTCOMES + nout + lane] & 1u) : 0u;
        chosen = __ballot_sync(0xffffffffu, b != 0u) & 0x1FFu;
        // every conditional probability along the nine Measure gates is positive iff the joint one is
        double pc = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if ((uint32_t)(lane + 32 * i) == chosen) pc = p[i];
        pc = __shfl_sync(0xffffffffu, pc, chosen & 31u);
        bad = pc <= 0.0;
    } else {
        double mine = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) mine += p[i];
        double incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        const double total = __shfl_sync(0xffffffffu, incl, 31);
        double u = 0.0;
        if (lane == 0) { Philox rng(ra.seed, ra.first_shot + shot, ra.stream); u = rng.next(); }
        const double target = __shfl_sync(0xffffffffu, u, 0) * total;
        const uint32_t hit = __ballot_sync(0xffffffffu, target < incl);          // non-empty: target < total = incl[31]
        const int sel = hit ? (__ffs(hit) - 1) : 31;
        uint32_t mine_d = 0;
        {
            double cum = incl - mine;
            int pick = -1, last_pos = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                cum += p[i];
                if (p[i] > 0.0) last_pos = i;
                if (pick < 0 && target < cum && p[i] > 0.0) pick = i;
            }
            if (pick < 0) pick = last_pos;            // rounding guard: never select an empty state
            mine_d = (uint32_t)(lane + 32 * pick);
        }
        chosen = __shfl_sync(0xffffffffu, mine_d, sel);
    }
    uint8_t outs[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) outs[j] = (uint8_t)((chosen >> j) & 1u);
    if (lane != 0) return;
    s17_record r = rec[shot];
    for (int j = 0; j < 9; ++j)
        if (r.n_outcomes < S17_MAX_OUTCOMES) r.outcomes[r.n_outcomes++] = outs[j];
    const uint32_t dm = chosen | (leak[shot] & 0x1FFu);                 // leaked data reads 1 (CS:61-63)
#define S17_B(i) ((dm >> (i)) & 1u)
    const uint32_t q = r.quiescent;
    const uint32_t c0 = (S17_B(0) ^ S17_B(1)) ^ ((q >> 0) & 1u);        // CS:71-77
    const uint32_t c1 = (S17_B(1) ^ S17_B(2) ^ S17_B(4) ^ S17_B(5)) ^ ((q >> 2) & 1u);
    const uint32_t c2 = (S17_B(3) ^ S17_B(4) ^ S17_B(6) ^ S17_B(7)) ^ ((q >> 5) & 1u);
    const uint32_t c3 = (S17_B(7) ^ S17_B(8)) ^ ((q >> 7) & 1u);
#undef S17_B
    const uint32_t weight = cw16[c0 | (c1 << 1) | (c2 << 2) | (c3 << 3)];
    const uint32_t logic = (__popc(dm) + weight) & 1u;                   // CS:86
    r.data_meas = (uint16_t)dm;
    r.logical_meas = (uint8_t)logic;
    atomicAdd(&counters[C_COUNTED], 1ull);
    if (logic != (uint32_t)ra.logical_state) { r.flags |= S17_F_FAIL; atomicAdd(&counters[C_FAILED], 1ull); }
    if (bad) { r.flags |= S17_F_ERROR; atomicAdd(&counters[C_ERRORS], 1ull); }
    rec[shot] = r;
}

// logical_measurement (CS:52-93) when k_cycle1 already drew the data register's basis state d' from the stored block:
// reading the corrected state gives d' ^ x-correction (a Pauli X relabels the computational basis).  One thread per
// shot, no state traffic.
__global__ void __launch_bounds__(256)
k_readout_s(ReadArgs ra, const uint32_t *__restrict__ dsamp, const uint8_t *__restrict__ cw16, uint32_t *__restrict__ leak,
            uint8_t *__restrict__ ready, s17_record *__restrict__ rec, unsigned long long *__restrict__ counters,
            int do_record /* 1: first the bookkeeping of the cycle that just ended (k_record's body) */, MeasArgs rec_ma,
            const uint32_t *__restrict__ meas_out, const uint32_t *__restrict__ lut, uint8_t *__restrict__ active)
{
    const int shot = blockIdx.x * blockDim.x + threadIdx.x;
    if (do_record && shot < ra.n_shots && active[shot]) {
        record_cycle(rec_ma, shot, meas_out, lut, leak, active, ready, rec, counters);
    }
    const bool live = shot < ra.n_shots && ready[shot];
    bool fail = false;
    if (live) {
        s17_record r = rec[shot];
        const uint32_t chosen = (dsamp[shot] ^ r.correction) & 0x1FFu;
        for (int j = 0; j < 9; ++j)
            if (r.n_outcomes < S17_MAX_OUTCOMES) r.outcomes[r.n_outcomes++] = (uint8_t)((chosen >> j) & 1u);
        const uint32_t dm = chosen | (leak[shot] & 0x1FFu);                 // leaked data reads 1 (CS:61-63)
#define S17_B(i) ((dm >> (i)) & 1u)
        const uint32_t q = r.quiescent;
        const uint32_t c0 = (S17_B(0) ^ S17_B(1)) ^ ((q >> 0) & 1u);        // CS:71-77
        const uint32_t c1 = (S17_B(1) ^ S17_B(2) ^ S17_B(4) ^ S17_B(5)) ^ ((q >> 2) & 1u);
        const uint32_t c2 = (S17_B(3) ^ S17_B(4) ^ S17_B(6) ^ S17_B(7)) ^ ((q >> 5) & 1u);
        const uint32_t c3 = (S17_B(7) ^ S17_B(8)) ^ ((q >> 7) & 1u);
#undef S17_B
        const uint32_t weight = cw16[c0 | (c1 << 1) | (c2 << 2) | (c3 << 3)];
        const uint32_t logic = (__popc(dm) + weight) & 1u;                   // CS:86
        r.data_meas = (uint16_t)dm;
        r.logical_meas = (uint8_t)logic;
        fail = logic != (uint32_t)ra.logical_state;
        if (fail) r.flags |= S17_F_FAIL;
        rec[shot] = r;
    }
    const uint32_t nl = __popc(__ballot_sync(0xffffffffu, live)), nf = __popc(__ballot_sync(0xffffffffu, fail));
    if ((threadIdx.x & 31) == 0) {
        if (nl) atomicAdd(&counters[C_COUNTED], (unsigned long long)nl);
        if (nf) atomicAdd(&counters[C_FAILED], (unsigned long long)nf);
    }
}

// start of a call: fresh per-shot flags and records, and (fixed-weight noise, run-til-fail) the placement of the shot's
// errors.  A plain s17_run folds both into k_plan: the reset into the preparation cycle's launch, the placement into the
// first noisy cycle's.
__global__ void k_reset(int n_shots, uint32_t *leak, uint8_t *active, uint8_t *ready, s17_record *rec, uint32_t *masks,
                        int clear_masks, int keep_leak, int place, NoiseDev nd, const double *__restrict__ cum, uint64_t seed,
                        uint64_t first_shot)
{
    const int shot = blockIdx.x * blockDim.x + threadIdx.x;
    if (shot >= n_shots) return;
    if (!keep_leak) leak[shot] = 0;
    reset_shot(shot, active, ready, rec);
    if (clear_masks)
        for (int i = 0; i < 2 * (int)S17_MASK_WORDS; ++i) masks[(size_t)shot * 2 * S17_MASK_WORDS + i] = 0;
    if (place) {
        uint32_t m[S17_MASK_WORDS];
#pragma unroll
        for (int i = 0; i < (int)S17_MASK_WORDS; ++i) m[i] = 0;
        place_errors(nd, cum, m, seed, first_shot + shot);
#pragma unroll
        for (int i = 0; i < (int)S17_MASK_WORDS; ++i) masks[(size_t)shot * (2 * S17_MASK_WORDS) + i] = m[i];
    }
}

// run-til-fail bookkeeping (CL:52-127): one slot = one run in flight; a failed run frees its slot for the next run id.
// alive: 0 idle, 1 running, 2 finished in the round that just ended (between k_rtf_finish and k_rtf_assign)
struct RtfSlot { uint32_t run_id, rounds, synd_b, alive; };
constexpr int kRtfThreads = 256;

__global__ void k_rtf_init(int n_slots, uint32_t total_runs, uint32_t chain, RtfSlot *slots, uint32_t *leak,
                           unsigned long long *counters)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    RtfSlot sl;
    sl.run_id = chain ? (uint32_t)s * chain : (uint32_t)s; sl.rounds = 0; sl.synd_b = 0; sl.alive = sl.run_id < total_runs;
    slots[s] = sl;
    leak[s] = 0;
    if (s == 0) { counters[C_NEXT_RUN] = (unsigned long long)min((uint32_t)n_slots, total_runs); }
}

__global__ void k_rtf_begin_round(int n_slots, const RtfSlot *slots, uint8_t *active, uint8_t *ready, s17_record *rec,
                                  uint32_t *leak, int reset_leak /* leak_reset = round */)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    active[s] = (uint8_t)(slots[s].alive != 0);
    ready[s] = 0;
    if (reset_leak) leak[s] = 0;
    s17_record z;
    memset(&z, 0, sizeof(z));
    rec[s] = z;
}

// End of a round, pass 1: count the round, close the runs that failed (or hit the cap), and count them per block.  The
// block that finishes last turns the per-block counts into exclusive offsets: the freed slots then take the next run ids
// in SLOT ORDER (k_rtf_assign), so the labelling of the runs is reproducible (no race on a shared counter).
__global__ void __launch_bounds__(kRtfThreads)
k_rtf_finish(int n_slots, uint32_t max_rounds, RtfSlot *slots, const s17_record *rec, uint32_t *rounds_out, uint32_t *syndb_out,
             uint32_t *block_cnt, uint32_t *block_off, unsigned long long *counters, uint32_t *ticket)
{
    __shared__ uint32_t s_w[kRtfThreads / 32], s_r[kRtfThreads / 32];
    __shared__ bool s_last;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    bool ran = false, fin = false;
    if (s < n_slots) {
        RtfSlot sl = slots[s];
        if (sl.alive) {
            ran = true;
            const s17_record r = rec[s];
            sl.rounds += 1;
            if (r.flags & S17_F_SYND2) sl.synd_b += 1;
            if ((r.flags & S17_F_FAIL) || (max_rounds && sl.rounds >= max_rounds)) {
                fin = true;
                rounds_out[sl.run_id] = sl.rounds;
                syndb_out[sl.run_id] = sl.synd_b;
                sl.alive = 2;
            }
            slots[s] = sl;
        }
    }
    const uint32_t bf = __ballot_sync(0xffffffffu, fin), br = __ballot_sync(0xffffffffu, ran);
    if ((threadIdx.x & 31) == 0) { s_w[threadIdx.x >> 5] = __popc(bf); s_r[threadIdx.x >> 5] = __popc(br); }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t nf = 0, nr = 0;
        for (int w = 0; w < kRtfThreads / 32; ++w) { nf += s_w[w]; nr += s_r[w]; }
        block_cnt[blockIdx.x] = nf;
        if (nr) atomicAdd(&counters[C_ROUNDS], (unsigned long long)nr);
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // exclusive scan of the block counts by the last block (<= a few thousand entries)
    __shared__ uint32_t s_sum[kRtfThreads];
    const int nb = gridDim.x, per = (nb + kRtfThreads - 1) / kRtfThreads;
    uint32_t mine = 0;
    for (int i = 0; i < per; ++i) {
        const int b = threadIdx.x * per + i;
        if (b < nb) mine += reinterpret_cast<volatile uint32_t *>(block_cnt)[b];
    }
    s_sum[threadIdx.x] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int t = 0; t < kRtfThreads; ++t) { const uint32_t v = s_sum[t]; s_sum[t] = run; run += v; }
        counters[C_RUNS_DONE] = counters[C_NEXT_RUN];            // base id of this round's re-assignments
        counters[C_NEXT_RUN] += run;
        *ticket = 0;
    }
    __syncthreads();
    uint32_t run = s_sum[threadIdx.x];
    for (int i = 0; i < per; ++i) {
        const int b = threadIdx.x * per + i;
        if (b < nb) { block_off[b] = run; run += reinterpret_cast<volatile uint32_t *>(block_cnt)[b]; }
    }
}

// End of a round, pass 2: freed slots take run ids base + rank (rank = position among the freed slots in slot order)
// while ids remain; count the slots that go on.
__global__ void __launch_bounds__(kRtfThreads)
k_rtf_assign(int n_slots, uint32_t total_runs, uint32_t chain, RtfSlot *slots, uint32_t *leak, const uint32_t *block_off,
             unsigned long long *counters, int reset_leak /* leak_reset = run */)
{
    __shared__ uint32_t s_w[kRtfThreads / 32];
    const int s = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    RtfSlot sl;
    sl.alive = 0;
    if (s < n_slots) sl = slots[s];
    const bool fin = sl.alive == 2;
    const uint32_t bf = __ballot_sync(0xffffffffu, fin);
    if (lane == 0) s_w[w] = __popc(bf);
    __syncthreads();
    if (fin) {
        uint32_t rank = block_off[blockIdx.x] + __popc(bf & ((1u << lane) - 1u));
        for (int i = 0; i < w; ++i) rank += s_w[i];
        unsigned long long id = counters[C_RUNS_DONE] + rank;
        if (chain) id = (sl.run_id % chain) + 1u < chain ? (unsigned long long)sl.run_id + 1u : (unsigned long long)total_runs;
        if (id < total_runs) {
            sl.run_id = (uint32_t)id; sl.rounds = 0; sl.synd_b = 0; sl.alive = 1;
            if (reset_leak) leak[s] = 0;
        } else {
            sl.alive = 0;
        }
        slots[s] = sl;
    }
    const uint32_t ba = __ballot_sync(0xffffffffu, sl.alive == 1);
    if (lane == 0 && ba) atomicAdd(&counters[C_ALIVE], (unsigned long long)__popc(ba));
}

}  // namespace s17

// =================================================================================================
// host side: context + C ABI
// =================================================================================================
using namespace s17;

struct TimedLaunch { int cat, sub; cudaEvent_t e0, e1; };

struct s17_ctx {
    int device = 0;
    uint64_t max_shots = 0;
    cudaStream_t stream = nullptr;
    double2 *state = nullptr;         // allocated on first use: full (2 MiB per shot) or compact (8 KiB per shot)
    bool state_full = false;
    size_t stride = kDim;             // amplitudes per shot in `state`
    void *encode_fn = nullptr;        // cuTensorMapEncodeTiled
    uint32_t *ev = nullptr, *masks = nullptr, *leak = nullptr, *lut = nullptr;
    uint8_t *active = nullptr, *ready = nullptr, *forced = nullptr, *cw16 = nullptr;
    double *hist = nullptr, *probs = nullptr, *cum = nullptr;
    ShotAux *aux = nullptr;
    s17_record *rec = nullptr;
    unsigned long long *counters = nullptr;
    s17_pinstr *plan = nullptr;
    uint32_t *plan_hot = nullptr;
    // k_plan's group tables, per cycle variant (first / second cycle): group words, slot bit -> group, survival products
    uint32_t *plan_grp = nullptr;
    uint16_t *plan_bit2grp = nullptr;   // slot bit -> instruction index of its entry, per cycle variant
    double *plan_surv = nullptr;
    int n_grp[2] = {0, 0};
    bool geo_ok[2] = {false, false}, plan_dup_slot[2] = {false, false}, no_geo = false;   // per cycle variant
    std::vector<s17_pinstr> plan_res;     // host copy of the resolved program (both variants)
    std::vector<double> h_probs;          // host copy of the slot probabilities
    RtfSlot *rtf_slots = nullptr;
    uint32_t *rtf_rounds = nullptr, *rtf_syndb = nullptr;
    uint32_t *rtf_block_cnt = nullptr, *rtf_block_off = nullptr, *rtf_ticket = nullptr;
    uint64_t rtf_capacity = 0;
    s17_run_args rtf_args{};          // the run-til-fail job in progress (s17_rtf_begin .. s17_rtf_end)
    uint64_t rtf_round_no = 0, rtf_alive = 0, rtf_rounds_total = 0;
    bool rtf_open = false;
    int n_plan = 0;
    std::vector<LayerArg> layers[3];   // prep, first noisy cycle, second noisy cycle
    std::vector<LayerArg2> layers2[3], layers2d[3];   // support-aware / dense variants
    double2 *zero_page = nullptr;
    CUtensorMap tmap;                 // the state as rows of 128 B, 8 KiB boxes, 128-byte swizzle (k_layer3)
    bool dense = false;               // option: full-state sweeps + explicit collapse (ProjectQ-equivalent traffic)
    // fused half-cycle programs (k_half) per cycle kind: [prep | stab 0 | stab 1][X half | Z half]
    LayerArg2 half_prog[3][2];
    HalfTail half_tail[3][2];
    bool half_ok[3] = {false, false, false};
    // whole-cycle diagonal-frame programs (k_cycle) per cycle kind, when the cycle has that structure
    DiagProg diag_prog[3];
    bool diag_ok[3] = {false, false, false};
    DiagProg1 diag_prog1[3];          // the same with host-chosen label maps (k_cycle1), when such maps exist
    bool diag1_ok[3] = {false, false, false};
    DiagCoh diag_coh[3];              // coherent over-rotations of a cycle that runs on k_cycle1<true>
    bool coh_ok[3] = {false, false, false};
    bool no_coh = false;              // S17_COH=0: coherent models keep the sweep kernels
    bool no_fuse_prep = false;        // S17_FUSE_PREP=0: the preparation cycle of an un-memoised run keeps its own launch
    bool fuse_prep = false;           // this call: the first noisy cycle's launch runs the preparation cycle of its shots first
    int fuse_basis = 0;               //            from this basis state
    bool no_trivial_prep = false;     // S17_TRIVIAL_PREP=0: plan the preparation cycle of a leak-free model like any other
    bool no_diag = false;
    bool last_collapsed = true;       // the previous cycle left every active shot collapsed to ancilla block 0
    uint32_t *frm = nullptr;          // per shot: frame words of the two half cycles (k_plan -> k_cycle)
    uint32_t *dsamp = nullptr;        // per shot: data basis state drawn by k_cycle1 for the read-out
    // memoised preparation cycle (k_prep): 256 leaf blocks, (P0, P1) per node of the outcome tree, leaf per shot
    double2 *prep_blocks = nullptr;
    double *prep_pp = nullptr;
    uint8_t *prep_leaf = nullptr, *prep_done = nullptr;
    bool prep_mixed = false;           // run-til-fail with a leakage model: shots with a leak flag run their own preparation
    int prep_basis = -1;              // basis index the cache was built for (-1: not built)
    bool no_prep_cache = false, plan_has_leak = false, use_prep = false;
    // memoised event-free first cycle (k_quiet): per leaf the block k_cycle1 stores and its read-out weights
    double2 *quiet_blocks = nullptr;
    double *quiet_cdf = nullptr;
    uint8_t *quiet_last = nullptr;    // per leaf and lane: the lane's last basis state of positive weight (k_quiet)
    uint8_t *quiet_ok = nullptr, *quiet_flag = nullptr;
    bool no_quiet = false, quiet_copy = false, use_quiet = false, any_quiet_leaf = false;
    bool quiet_lazy = false;          // the last run left event-free shots' blocks in the cache (s17_read_state looks there)
    PrepTree prep_tree;
    bool use_dsamp = false;
    bool no_fuse = false;
    bool debug_sync = false;
    std::string debug_note;
    bool force_ring = false;          // option: use the TMA-ring kernel (k_layer2) for support-aware sweeps as well
    uint32_t *act_list = nullptr, *n_act = nullptr, *meas_out = nullptr;
    uint32_t *n_act2 = nullptr;       // two list lengths used in turn: k_plan fills one and clears the other (n_act points at the current one)
    uint32_t plan_seq = 0;
    bool pend_reset = false;          // the per-shot reset of a call is still to be done: the preparation cycle's k_plan does it
    bool pend_rec = false;            // the bookkeeping (k_record) of the last cycle is still to be applied: the next k_plan /
    MeasArgs pend_ma;                 // k_readout_s does it, or flush_record
    int num_sms = 148;
    NoiseDev nd{};
    uint32_t total_slots = 0;
    bool have_program = false, have_tables = false, have_replay = false;
    bool two_round_lists = false;     // every use_round entry satisfies slot + len/2 < len (needed by a stab_ind=1 cycle)
    uint64_t last_n = 0;
    // timing
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    std::vector<TimedLaunch> launches;
    double t_gate = 0, t_meas = 0, t_other = 0, t_span = 0;
    double t_sub[S17_T_N] = {};       // per kernel class (S17_T_*)
    uint64_t n_sub[S17_T_N] = {}, u_sub[S17_T_N] = {};
    uint64_t n_gate = 0, n_all = 0;
    cudaEvent_t span0 = nullptr, span1 = nullptr;
    std::string err;
};

static thread_local std::string g_err;

static int fail(s17_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg;
    g_err = msg;
    return code;
}
#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(ctx, S17_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

static cudaEvent_t get_event(s17_ctx *c) {
    if (c->ev_used == c->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->ev_pool.push_back(e);
    }
    return c->ev_pool[c->ev_used++];
}
enum { CAT_GATE = 0, CAT_MEAS = 1, CAT_OTHER = 2 };
struct Scope {
    s17_ctx *c; TimedLaunch tl; int line;
    Scope(s17_ctx *c_, int cat, int line_ = 0, int sub = -1) : c(c_), line(line_) {
        tl.cat = cat; tl.sub = sub; tl.e0 = get_event(c); tl.e1 = get_event(c); cudaEventRecord(tl.e0, c->stream);
    }
    ~Scope() {
        cudaEventRecord(tl.e1, c->stream);
        c->launches.push_back(tl);
        if (c->debug_sync) {                   // S17_DEBUG_SYNC=1: find the launch that faults
            cudaError_t e = cudaStreamSynchronize(c->stream);
            if (e == cudaSuccess) e = cudaGetLastError();
            if (e != cudaSuccess && c->debug_note.empty())
                c->debug_note = std::string("first failing launch at s17.cu:") + std::to_string(line) + ": " + cudaGetErrorString(e);
        }
    }
};
// NVTX range around a phase of a call (tracing): s17_run > preparation / cycle 1 / cycle 2 / read-out
struct Nvtx {
    explicit Nvtx(const char *name) { nvtxRangePushA(name); }
    ~Nvtx() { nvtxRangePop(); }
};
#define SCOPE(cat) Scope s(ctx, cat, __LINE__)
#define SCOPE2(cat, sub) Scope s(ctx, cat, __LINE__, sub)

extern "C" int s17_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" const char *s17_last_error(const s17_ctx *ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

extern "C" int s17_create(s17_ctx **out, int device, uint64_t max_shots) {
    s17_ctx *ctx = nullptr;
    if (!out || max_shots == 0) return fail(nullptr, S17_E_ARG, "s17_create: bad arguments");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
        return fail(nullptr, S17_E_CUDA, "s17_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= n) return fail(nullptr, S17_E_ARG, "s17_create: device index out of range");
    ctx = new s17_ctx();
    ctx->device = device;
    ctx->max_shots = max_shots;
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    const size_t B = max_shots;
    CK(cudaMalloc(&ctx->ev, B * S17_EV_WORDS * 4));
    CK(cudaMalloc(&ctx->masks, B * 2 * S17_MASK_WORDS * 4));
    CK(cudaMalloc(&ctx->leak, B * 4));
    CK(cudaMalloc(&ctx->active, B));
    CK(cudaMalloc(&ctx->ready, B));
    CK(cudaMalloc(&ctx->forced, B * S17_MAX_OUTCOMES));
    CK(cudaMalloc(&ctx->aux, B * sizeof(ShotAux)));
    CK(cudaMalloc(&ctx->rec, B * sizeof(s17_record)));
    CK(cudaMalloc(&ctx->counters, C_N * sizeof(unsigned long long)));
    CK(cudaMalloc(&ctx->lut, 256 * 4));
    CK(cudaMalloc(&ctx->cw16, 16));
    CK(cudaMalloc(&ctx->probs, 2048 * sizeof(double)));
    CK(cudaMalloc(&ctx->cum, 2048 * sizeof(double)));
    CK(cudaMalloc(&ctx->rtf_slots, B * sizeof(RtfSlot)));
    CK(cudaMalloc(&ctx->rtf_block_cnt, (B / kRtfThreads + 1) * 4));
    CK(cudaMalloc(&ctx->rtf_block_off, (B / kRtfThreads + 1) * 4));
    CK(cudaMalloc(&ctx->rtf_ticket, 4));
    CK(cudaMemset(ctx->rtf_ticket, 0, 4));
    CK(cudaMemset(ctx->probs, 0, 2048 * sizeof(double)));
    CK(cudaMemset(ctx->leak, 0, B * 4));
    CK(cudaMemset(ctx->ev, 0, B * S17_EV_WORDS * 4));
    {
        // debugging aid: S17_MEMSET=<bitmask> fills selected buffers with S17_MEMSET_BYTE at creation
        const char *mm = getenv("S17_MEMSET");
        const char *mb = getenv("S17_MEMSET_BYTE");
        const int mask = mm ? atoi(mm) : 0, byte = mb ? atoi(mb) : 0;
        if (mask & 2) CK(cudaMemset(ctx->masks, byte, B * 2 * S17_MASK_WORDS * 4));
        if (mask & 4) { CK(cudaMemset(ctx->active, byte, B)); CK(cudaMemset(ctx->ready, byte, B)); }
        if (mask & 8) CK(cudaMemset(ctx->forced, byte, B * S17_MAX_OUTCOMES));
        if (mask & 32) CK(cudaMemset(ctx->aux, byte, B * sizeof(ShotAux)));
        if (mask & 64) CK(cudaMemset(ctx->rec, byte, B * sizeof(s17_record)));
    }
    CK(cudaFuncSetAttribute(k_layer2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kL2Smem));
    CK(cudaFuncSetAttribute(k_layer3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kL3Smem));
    CK(cudaFuncSetAttribute(k_half, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kL3Smem));
    CK(cudaMalloc(&ctx->act_list, B * 4));
    CK(cudaMalloc(&ctx->n_act2, 8));
    CK(cudaMemset(ctx->n_act2, 0, 8));
    ctx->n_act = ctx->n_act2;
    CK(cudaMalloc(&ctx->meas_out, B * 2 * 4));
    CK(cudaMalloc(&ctx->frm, B * 4 * 4));
    CK(cudaMalloc(&ctx->dsamp, B * 4));
    CK(cudaMalloc(&ctx->prep_leaf, B));
    CK(cudaMalloc(&ctx->prep_done, B));
    CK(cudaMalloc(&ctx->prep_blocks, 256 * (size_t)kRunBytes));
    CK(cudaMalloc(&ctx->prep_pp, 2 * 255 * sizeof(double)));
    CK(cudaMalloc(&ctx->quiet_blocks, 256 * (size_t)kRunBytes));
    CK(cudaMalloc(&ctx->quiet_cdf, 256 * (size_t)kDgCdfDoubles * sizeof(double)));
    CK(cudaMalloc(&ctx->quiet_last, 256 * 32));
    CK(cudaMalloc(&ctx->quiet_ok, 256));
    CK(cudaMemset(ctx->quiet_ok, 0, 256));
    CK(cudaMalloc(&ctx->quiet_flag, B));
    CK(cudaMemset(ctx->frm, 0, B * 4 * 4));
    CK(cudaFuncSetAttribute(k_plan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPlanSmem));
    CK(cudaFuncSetAttribute(k_cycle1<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDgSmem1));
    CK(cudaFuncSetAttribute(k_cycle1<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDgSmem1));
    {
        typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                     const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return fail(ctx, S17_E_CUDA, "cuTensorMapEncodeTiled is not available");
        ctx->encode_fn = fn;
    }
    CK(cudaMalloc(&ctx->zero_page, kRunBytes));
    CK(cudaMemset(ctx->zero_page, 0, kRunBytes));
    {
        const char *d = getenv("S17_DENSE");
        ctx->dense = d && d[0] == '1';
        const char *ds = getenv("S17_DEBUG_SYNC");
        ctx->debug_sync = ds && ds[0] == '1';
        const char *nf = getenv("S17_FUSE");
        ctx->no_fuse = nf && nf[0] == '0';
        const char *nd = getenv("S17_DIAG");
        ctx->no_diag = nd && nd[0] == '0';          // 0: k_half / the sweep kernels instead of the whole-cycle kernel
        const char *pc = getenv("S17_PREP_CACHE");
        ctx->no_prep_cache = pc && pc[0] == '0';
        const char *qc = getenv("S17_QUIET");            // 0: run event-free first cycles through k_cycle1 like any other
        ctx->no_quiet = qc && qc[0] == '0';
        const char *qcp = getenv("S17_QUIET_COPY");      // 1: always copy the cached block into the shot's state
        ctx->quiet_copy = qcp && qcp[0] == '1';
        const char *r = getenv("S17_RING");
        ctx->force_ring = r && r[0] == '1';
        const char *tp = getenv("S17_TRIVIAL_PREP");
        ctx->no_trivial_prep = tp && tp[0] == '0';
        const char *fp = getenv("S17_FUSE_PREP");
        ctx->no_fuse_prep = fp && fp[0] == '0';
        const char *ge = getenv("S17_GEO");               // 0: one Bernoulli draw per entry instead of geometric skipping
        ctx->no_geo = ge && ge[0] == '0';
    }
    {
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        ctx->num_sms = prop.multiProcessorCount;
    }
    CK(cudaEventCreate(&ctx->span0));
    CK(cudaEventCreate(&ctx->span1));
    *out = ctx;
    return S17_OK;
}

extern "C" int s17_destroy(s17_ctx *ctx) {
    if (!ctx) return S17_OK;
    cudaSetDevice(ctx->device);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        // the context is in a sticky error state (a kernel faulted): releasing resources would fault again
        delete ctx;
        return S17_E_CUDA;
    }
    void *ptrs[] = {ctx->state, ctx->ev, ctx->masks, ctx->leak, ctx->active, ctx->ready, ctx->forced, ctx->hist,
                    ctx->aux, ctx->rec, ctx->counters, ctx->lut, ctx->cw16, ctx->probs, ctx->cum, ctx->plan,
                    ctx->rtf_slots, ctx->rtf_rounds, ctx->rtf_syndb, ctx->rtf_block_cnt, ctx->rtf_block_off, ctx->rtf_ticket, ctx->act_list, ctx->n_act2, ctx->zero_page, ctx->meas_out,
                    ctx->frm, ctx->dsamp, ctx->prep_leaf, ctx->prep_done, ctx->prep_blocks, ctx->prep_pp, ctx->plan_hot,
                    ctx->plan_grp, ctx->plan_bit2grp, ctx->plan_surv,
                    ctx->quiet_blocks, ctx->quiet_cdf, ctx->quiet_last, ctx->quiet_ok, ctx->quiet_flag};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return S17_OK;
}

static int check_layer(s17_ctx *ctx, const s17_layer &L) {
    if (L.n_anc != 3) return fail(ctx, S17_E_ARG, "a sweep stages three ancilla bits (at most two timesteps per sweep)");
    if (L.n_ops > 8) return fail(ctx, S17_E_ARG, "too many ops in one layer");
    for (int k = 0; k < L.n_anc; ++k)
        if (L.anc_bits[k] < 9 || L.anc_bits[k] > 16 || (k && L.anc_bits[k] <= L.anc_bits[k - 1]))
            return fail(ctx, S17_E_ARG, "layer anc_bits must be ascending ancilla bits in 9..16");
    for (int o = 0; o < L.n_ops; ++o) {
        const s17_op &op = L.ops[o];
        if (op.ev_word >= S17_EV_WORDS) return fail(ctx, S17_E_ARG, "op ev_word out of range");
        if (op.kind == S17_OP_IDLE) continue;
        if (op.kind != S17_OP_ENT || op.data_bit > 8 || op.n_uops > 8) return fail(ctx, S17_E_ARG, "bad op");
        bool ok = false;
        for (int k = 0; k < L.n_anc; ++k) ok |= (L.anc_bits[k] == op.anc_bit);
        if (!ok) return fail(ctx, S17_E_ARG, "op ancilla bit is not staged by its layer");
    }
    return S17_OK;
}

// host-side lowering of one sweep for k_layer2: per-op constants and the pass schedule
static LayerArg2 make_layer2(const LayerArg &A, uint32_t support /* ancilla bits possibly non-zero */, bool first, bool dense,
                             int n_main = -1 /* ops that get passes (default: all) */, const uint32_t *idle_filter = nullptr) {
    LayerArg2 P;
    memset(&P, 0, sizeof(P));
    P.A = A;
    const int n_anc = A.L.n_anc, n_other = 8 - n_anc;
    {
        int n = 0;
        for (int tid = 0; tid < (1 << n_other); ++tid) {
            bool ok = true;
            for (int k = 0; k < n_other; ++k)
                if (((tid >> k) & 1) && !dense && !((support >> A.other_bits[k]) & 1u)) ok = false;
            if (ok) P.tile_ids[n++] = (uint8_t)tid;
        }
        int sh = 0;
        while ((1 << sh) < n) ++sh;
        P.tile_shift = (uint8_t)sh;
        for (int r = 0; r < (1 << n_anc); ++r) {
            bool ok = true;
            for (int k = 0; k < n_anc; ++k)
                if (((r >> k) & 1) && !dense && !((support >> A.L.anc_bits[k]) & 1u)) ok = false;
            if (ok) P.in_runs |= (uint16_t)(1u << r);
        }
        P.first = (first && !dense) ? 1 : 0;
    }
    const s17_layer &L = A.L;
    for (int o = 0; o < L.n_ops; ++o) {
        const s17_op &op = L.ops[o];
        OpC &c = P.oc[o];
        c.ev_word = op.ev_word;
        if (op.kind != S17_OP_ENT) continue;
        c.lb0 = op.data_bit;
        c.lb1 = 9;
        for (int k = 0; k < n_anc; ++k)
            if (L.anc_bits[k] == op.anc_bit) c.lb1 = (uint8_t)(9 + k);
        // standard template: [Ry(+) ev0] Rxx(s, skippable, ev6) [Rx(-s) ev12] [Ry(-) ev18]
        int u = 0;
        bool ok = true;
        uint8_t shape = 0;
        int8_t s = 0;
        if (u < op.n_uops && op.uops[u].type == S17_U_RY && op.uops[u].sign == 1 && !op.uops[u].skippable &&
            op.uops[u].ev_shift == 0) { shape |= 1; ++u; }
        if (u < op.n_uops && op.uops[u].type == S17_U_RXX && op.uops[u].skippable && op.uops[u].ev_shift == 6 &&
            (op.uops[u].sign == 1 || op.uops[u].sign == -1)) { s = op.uops[u].sign; ++u; } else ok = false;
        if (ok && u < op.n_uops && op.uops[u].type == S17_U_RX && op.uops[u].sign == -s && !op.uops[u].skippable &&
            op.uops[u].ev_shift == 12) { shape |= 2; ++u; }
        if (ok && u < op.n_uops && op.uops[u].type == S17_U_RY && op.uops[u].sign == -1 && !op.uops[u].skippable &&
            op.uops[u].ev_shift == 18) { shape |= 4; ++u; }
        if (u != op.n_uops || op.partial) ok = false;
        c.generic = ok ? 0 : 1;
        c.shape = shape;
        c.s = s;
        for (int k = 0; k < op.n_uops; ++k)
            if (op.uops[k].type <= S17_U_H) (op.uops[k].skippable ? c.n_skip : c.n_fixed)++;
    }
    // passes: consecutive operations on four distinct bits share one pass; an idle mask closes a run
    int np = 0, last_gate_pass = -1;
    std::vector<int> run;
    auto flush = [&]() {
        size_t i = 0;
        while (i < run.size()) {
            PassC &ps = P.pass[np];
            ps.idle_word = 255;
            const OpC &c1 = P.oc[run[i]];
            if (i + 1 < run.size() && P.oc[run[i + 1]].lb0 != c1.lb0 && P.oc[run[i + 1]].lb1 != c1.lb1) {
                ps.kind = 1; ps.o1 = (uint8_t)run[i]; ps.o2 = (uint8_t)run[i + 1];
                i += 2;
            } else {
                ps.kind = 0; ps.o1 = (uint8_t)run[i];
                i += 1;
            }
            last_gate_pass = np++;
        }
        run.clear();
    };
    const int n_sched = n_main < 0 ? (int)L.n_ops : n_main;
    for (int o = 0; o < n_sched; ++o) {
        if (L.ops[o].kind == S17_OP_ENT) { run.push_back(o); continue; }
        flush();
        const uint32_t filt = idle_filter ? idle_filter[o] : 0x1FFFFu;
        if (np == 0 || P.pass[np - 1].idle_word != 255) {      // nothing to attach to: idle-only pass
            P.pass[np].kind = 2;
            P.pass[np].idle_word = L.ops[o].ev_word;
            P.pass[np].idle_filter = filt;
            ++np;
        } else {
            P.pass[np - 1].idle_word = L.ops[o].ev_word;
            P.pass[np - 1].idle_filter = filt;
        }
    }
    flush();
    if (last_gate_pass >= 0) P.pass[last_gate_pass].last = 1;
    P.n_pass = (uint32_t)np;
    // lane mapping for the 128B-swizzled layout: a quarter warp must hit 8 different 16-byte columns.  Column bit k
    // of the swizzled index is idx[k] ^ idx[k+3], so the three lowest item bits go to positions k, or k+3 when k is
    // held in registers by this pass.
    const int tile_bits = 9 + n_anc;
    for (int q = 0; q < np; ++q) {
        PassC &ps = P.pass[q];
        bool removed[16] = {false};
        if (ps.kind != 2) { removed[P.oc[ps.o1].lb0] = removed[P.oc[ps.o1].lb1] = true; }
        else { removed[0] = removed[9] = true; }
        if (ps.kind == 1) { removed[P.oc[ps.o2].lb0] = removed[P.oc[ps.o2].lb1] = true; }
        bool used[16] = {false};
        int n = 0;
        for (int k = 0; k < 3; ++k) {
            int pos = !removed[k] ? k : (!removed[k + 3] && !used[k + 3] ? k + 3 : -1);
            if (pos < 0 || used[pos]) {                          // unavoidable conflict: take any free bit
                for (pos = 0; pos < tile_bits && (removed[pos] || used[pos]); ++pos) {}
            }
            used[pos] = true;
            ps.pos[n++] = (uint8_t)pos;
        }
        for (int b = 0; b < tile_bits; ++b)
            if (!removed[b] && !used[b] && n < 12) { used[b] = true; ps.pos[n++] = (uint8_t)b; }
    }
    return P;
}

// Lower one half cycle (the sweeps up to a measurement) to the fused k_half form.  Returns false when the half cycle
// does not have the required shape (then the generic sweep kernels are used).
static bool make_half(const std::vector<LayerArg> &layers, size_t begin, size_t end, LayerArg2 &P, HalfTail &T) {
    std::vector<s17_op> ops;
    for (size_t i = begin; i < end; ++i) {
        if (layers[i].L.n_anc != 3) return false;
        for (int o = 0; o < layers[i].L.n_ops; ++o) ops.push_back(layers[i].L.ops[o]);
    }
    const uint8_t mask = layers[end - 1].L.meas_mask;
    if (!mask) return false;
    // the four ancillas of the half cycle, the tail one = the one that is touched last for the first time
    std::vector<int> anc;
    for (const s17_op &op : ops)
        if (op.kind == S17_OP_ENT) {
            bool seen = false;
            for (int a : anc) seen |= (a == op.anc_bit);
            if (!seen) anc.push_back(op.anc_bit);
        }
    if (anc.size() != 4) return false;
    uint8_t set = 0;
    for (int a : anc) set |= (uint8_t)(1u << (a - 9));
    if (set != mask) return false;
    const int e = anc[3];
    std::vector<int> tail;
    for (size_t i = 0; i < ops.size(); ++i)
        if (ops[i].kind == S17_OP_ENT && ops[i].anc_bit == e) tail.push_back((int)i);
    if (tail.size() != 2 || ops[tail[0]].data_bit == ops[tail[1]].data_bit) return false;
    for (int ti : tail)                                   // a moved operation must commute with everything it jumps over
        for (size_t q = ti + 1; q < ops.size(); ++q)
            if (ops[q].kind == S17_OP_ENT && ops[q].anc_bit != e && ops[q].data_bit == ops[ti].data_bit) return false;
    for (const s17_op &op : ops)                          // uop parameters were re-indexed per sweep: coherent models keep the generic path
        if (op.kind == S17_OP_ENT)
            for (int u = 0; u < op.n_uops; ++u)
                if (op.uops[u].type >= S17_U_ROT_X) return false;
    LayerArg A;
    memset(&A, 0, sizeof(A));
    A.L.n_anc = 3;
    int na = 0;
    std::vector<int> main_anc(anc.begin(), anc.begin() + 3);
    std::sort(main_anc.begin(), main_anc.end());
    for (int a : main_anc) A.L.anc_bits[na++] = (uint8_t)a;
    int k = 0;
    for (int b = 9; b <= 16; ++b) {
        bool in = false;
        for (int a : main_anc) in |= (a == b);
        if (!in) A.other_bits[k++] = (uint8_t)b;
    }
    memset(&T, 0, sizeof(T));
    T.n_tail = 2;
    T.e_bit = (uint8_t)e;
    T.meas_mask = mask;
    for (int q = 0; q < 3; ++q) T.idle_word[q] = 255;
    uint32_t filt[S17_MAX_LAYER_OPS];
    int n = 0;
    uint32_t pend = 0;                                    // qubits whose (moved) operation precedes the current point
    int seg = 0;                                          // 0: before tail op 0, 1: between, 2: after tail op 1
    for (size_t i = 0; i < ops.size(); ++i) {
        const s17_op &op = ops[i];
        if (op.kind == S17_OP_ENT && op.anc_bit == e) {
            pend |= (1u << op.data_bit) | (1u << e);
            ++seg;
            continue;
        }
        if (n >= (int)S17_MAX_LAYER_OPS - 2) return false;
        if (op.kind == S17_OP_IDLE) {
            filt[n] = 0x1FFFFu & ~pend;
            if (pend) {
                if (T.idle_word[seg] != 255) return false;        // one idle mask per tail segment
                T.idle_word[seg] = op.ev_word;
                T.idle_filter[seg] = pend;
            }
        } else {
            filt[n] = 0x1FFFFu;
        }
        A.L.ops[n++] = op;
    }
    const int n_main = n;
    T.op[0] = (uint8_t)n;
    A.L.ops[n++] = ops[tail[0]];
    T.op[1] = (uint8_t)n;
    A.L.ops[n++] = ops[tail[1]];
    filt[n - 2] = filt[n - 1] = 0x1FFFFu;
    A.L.n_ops = (uint8_t)n;
    A.L.meas_mask = mask;
    P = make_layer2(A, 0, true, false, n_main, filt);
    return P.n_pass <= S17_MAX_LAYER_OPS;
}

// Lower a fused half cycle (make_half) to the diagonal-frame form of k_cycle.  Returns false when the half cycle is not
// of that form: an idle mask, a non-standard operation, or single-qubit gates that do not bracket the data qubits.
static bool make_diag_half(const LayerArg2 &P, const HalfTail &T, int half_index, const s17_pinstr *plan, uint32_t n_plan,
                           DiagHalf &H, uint32_t *covered /* data bits some slot of this half indexes */) {
    memset(&H, 0, sizeof(H));
    memset(H.lane_bit, 255, sizeof(H.lane_bit));
    const s17_layer &L = P.A.L;
    bool any_ry = false, all_x = true;
    int first_idx[9], last_idx[9], min_word = 255, max_word = -1;
    for (int d = 0; d < 9; ++d) first_idx[d] = last_idx[d] = -1;
    std::vector<int> order;                                   // operations in time order (= event word order)
    bool has_idle = false;
    for (int o = 0; o < L.n_ops; ++o) {
        if (L.ops[o].kind == S17_OP_IDLE) { has_idle = true; continue; }     // rides in the frame words (k_plan)
        if (L.ops[o].kind != S17_OP_ENT || P.oc[o].generic) return false;
        order.push_back(o);
    }
    if (has_idle)                                             // the plan must carry the frame information of idle events
        for (uint32_t i = 0; i < n_plan; ++i)
            if (plan[i].kind == S17_P_IDLE && !(plan[i].b & 0x80u)) return false;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return L.ops[a].ev_word < L.ops[b].ev_word; });
    for (int o : order) {
        const int d = L.ops[o].data_bit;
        if (P.oc[o].shape & 5) { any_ry = true; all_x = false; }
        if (first_idx[d] < 0) first_idx[d] = o;
        last_idx[d] = o;
        min_word = std::min(min_word, (int)L.ops[o].ev_word);
        max_word = std::max(max_word, (int)L.ops[o].ev_word);
    }
    if (any_ry) {
        // computational frame: on every data qubit the first operation opens the bracket with Ry(+), the last one
        // closes it with Ry(-), nothing in between carries an Ry
        for (int o : order) {
            const int d = L.ops[o].data_bit;
            const bool want1 = (o == first_idx[d]), want2 = (o == last_idx[d]);
            if (((P.oc[o].shape & 1) != 0) != want1 || ((P.oc[o].shape & 4) != 0) != want2) return false;
        }
    }
    H.frame_x = all_x ? 1 : 0;
    for (int o : order) {
        const s17_op &op = L.ops[o];
        int k;
        if (op.anc_bit == T.e_bit) k = 3;
        else {
            k = P.oc[o].lb1 - 9;
            if (k < 0 || k > 2) return false;
        }
        if (H.n_ops[k] >= 4) return false;
        for (int j = 0; j < H.n_ops[k]; ++j)
            if (H.ops[k][j].d == op.data_bit) return false;          // table index bits must be distinct label bits
        DiagOp &dop = H.ops[k][H.n_ops[k]++];
        dop.d = op.data_bit;
        dop.ev_word = op.ev_word;
        dop.flags = (uint8_t)((P.oc[o].s < 0 ? 1 : 0) | ((P.oc[o].shape & 2) ? 2 : 0));
        dop.pad = (uint8_t)(std::find(order.begin(), order.end(), o) - order.begin());     // index in time order
        if (dop.pad >= 12) return false;                     // 12 ancilla-Z bits in a frame word
        for (uint32_t i = 0; i < n_plan; ++i)                 // idle Z events that name this operation: same ancilla
            if (plan[i].kind == S17_P_IDLE && (plan[i].b & 0x80u) && plan[i].field != 255 && plan[i].pad == op.ev_word &&
                (plan[i].field != dop.pad || plan[i].a != op.anc_bit))
                return false;
        H.aj[k] = (uint8_t)(op.anc_bit - 9);
        // k_plan must have been told the same frame, data bit and half boundaries for this operation
        bool found = false;
        for (uint32_t i = 0; i < n_plan; ++i) {
            const s17_pinstr &in = plan[i];
            if (in.kind != S17_P_OP_END || in.word != op.ev_word) continue;
            found = true;
            if (in.slot != op.data_bit || ((in.a & 8) != 0) != all_x || ((in.a >> 5) & 1) != half_index) return false;
            if (((in.a & 16) != 0) != (op.ev_word == min_word) || ((in.a & 64) != 0) != (op.ev_word == max_word)) return false;
            if ((in.a & 7) != P.oc[o].shape || (in.b != 0) != (P.oc[o].s > 0)) return false;
        }
        if (!found) return false;
    }
    // where each table index bit comes from: label = i | lane << 4 in the Hadamard frame, lane | i << 5 otherwise
    for (int k = 0; k < 4; ++k)
        for (int j = 0; j < H.n_ops[k]; ++j) {
            const int d = H.ops[k][j].d;
            const int reg_lo = all_x ? 0 : 5, reg_n = 4, lane_lo = all_x ? 4 : 0;
            if (d >= reg_lo && d < reg_lo + reg_n) {
                for (int i = 0; i < 16; ++i) H.roff[k][i] |= (uint32_t)(((i >> (d - reg_lo)) & 1) << j) << 4;
            } else {
                H.lane_bit[k][j] = (uint8_t)(d - lane_lo);
            }
        }
    // label-dependent signs ride in the table of the lowest slot that indexes the label bit; every slot measures
    uint32_t owned = 0;
    int entries = 0;
    for (int k = 0; k < 4; ++k) {
        if (H.n_ops[k] == 0) return false;
        H.tbase[k] = (uint8_t)entries;
        entries += 1 << H.n_ops[k];
        for (int j = 0; j < H.n_ops[k]; ++j) {
            const uint32_t bit = 1u << H.ops[k][j].d;
            if (!(owned & bit)) { H.ops[k][j].flags |= 4; owned |= bit; }
        }
    }
    if (entries > kDgTabEntries) return false;
    H.n_entries = (uint8_t)entries;
    *covered = owned;
    return true;
}

// ---- k_cycle1: choose label <-> (lane, register) maps under which every slot indexes its table with <= 1 register bit ----
static bool gf2_invertible3(const uint8_t v[3]) {
    // three 3-bit vectors are linearly independent over GF(2) iff no non-empty subset XORs to zero
    for (int m = 1; m < 8; ++m) {
        uint8_t x = 0;
        for (int k = 0; k < 3; ++k)
            if ((m >> k) & 1) x ^= v[k];
        if (x == 0) return false;
    }
    return true;
}

static void fill_map(DiagMap &M, const std::vector<int> &reg, const std::vector<int> &lane, const uint8_t gvec[9]) {
    memset(&M, 0, sizeof(M));
    for (int j = 0; j < 4; ++j) M.reg[j] = (uint8_t)reg[j];
    for (int m = 0; m < 5; ++m) M.lane[m] = (uint8_t)lane[m];
    for (int i = 0; i < 16; ++i) {
        uint32_t label = 0, c = 0;
        for (int j = 0; j < 4; ++j)
            if ((i >> j) & 1) label |= 1u << reg[j];
        for (int b = 3; b < 9; ++b)
            if ((label >> b) & 1u) c ^= gvec[b];
        M.lin16[i] = label << 4;
        M.swz16[i] = (label ^ c) << 4;
    }
}

// Order the register bits of a map so that slot k of H is fed by register bit k (a slot takes at most one register bit
// and a register bit feeds at most one slot), then fill lane_bit / off1.  false: no such order.
static bool order_and_apply_map1(DiagHalf &H, std::vector<int> &reg, const std::vector<int> &lane, uint32_t off1[4]) {
    int slot_of_bit[9], bit_of_slot[4] = {-1, -1, -1, -1};
    for (int b = 0; b < 9; ++b) slot_of_bit[b] = -1;
    for (int k = 0; k < 4; ++k)
        for (int j = 0; j < H.n_ops[k]; ++j) {
            const int d = H.ops[k][j].d;
            if (std::find(reg.begin(), reg.end(), d) == reg.end()) continue;
            if (slot_of_bit[d] >= 0 || bit_of_slot[k] >= 0) return false;
            slot_of_bit[d] = k;
            bit_of_slot[k] = d;
        }
    std::vector<int> spare;
    for (int d : reg)
        if (slot_of_bit[d] < 0) spare.push_back(d);
    for (int k = 0; k < 4; ++k)
        if (bit_of_slot[k] < 0) { bit_of_slot[k] = spare.back(); spare.pop_back(); }
    for (int k = 0; k < 4; ++k) reg[k] = bit_of_slot[k];
    memset(H.lane_bit, 255, sizeof(H.lane_bit));
    for (int k = 0; k < 4; ++k) {
        off1[k] = 0;
        for (int j = 0; j < H.n_ops[k]; ++j) {
            const int d = H.ops[k][j].d;
            if (d == reg[k]) { off1[k] = (uint32_t)(1u << j) << 4; continue; }
            bool placed = false;
            for (int m = 0; m < 5; ++m)
                if (lane[m] == d) { H.lane_bit[k][j] = (uint8_t)m; placed = true; }
            if (!placed) return false;
        }
    }
    return true;
}

static bool make_diag1(const DiagProg &G, DiagProg1 &P1) {
    memset(&P1, 0, sizeof(P1));
    const DiagHalf &hx = G.h[0], &hz = G.h[1];
    if (!hx.frame_x || hz.frame_x) return false;            // k_cycle1: Hadamard-frame half first, computational-frame half second
    auto count_ok = [](const DiagHalf &H, uint32_t regmask) {
        for (int k = 0; k < 4; ++k) {
            int n = 0;
            for (int j = 0; j < H.n_ops[k]; ++j) n += (regmask >> H.ops[k][j].d) & 1u;
            if (n > 1) return false;
        }
        return true;
    };
    for (uint32_t rz = 0; rz < 512; ++rz) {
        if (__builtin_popcount(rz) != 4 || !count_ok(hz, rz)) continue;
        for (int sh = 0; sh < 9; ++sh) {
            if ((rz >> sh) & 1u) continue;
            const uint32_t rx = 0x1FFu & ~rz & ~(1u << sh);
            if (!count_ok(hx, rx)) continue;
            std::vector<int> regA, laneA, regB, laneB;
            for (int b = 0; b < 9; ++b) {
                ((rz >> b) & 1u ? regA : laneA).push_back(b);
                ((rx >> b) & 1u ? regB : laneB).push_back(b);
            }
            // exchange swizzle: the three lowest lane bits of either map must reach eight different 16-byte columns
            std::vector<int> hi;                               // label bits >= 3 among those six lane bits
            for (int m = 0; m < 3; ++m) {
                if (laneA[m] >= 3) hi.push_back(laneA[m]);
                if (laneB[m] >= 3 && std::find(hi.begin(), hi.end(), laneB[m]) == hi.end()) hi.push_back(laneB[m]);
            }
            uint8_t gvec[9] = {1, 2, 4, 0, 0, 0, 0, 0, 0};
            bool found = false;
            const uint32_t combos = 1u << (3 * hi.size());
            for (uint32_t c = 0; c < combos && !found; ++c) {
                for (size_t q = 0; q < hi.size(); ++q) gvec[hi[q]] = (uint8_t)((c >> (3 * q)) & 7u);
                uint8_t va[3], vb[3];
                for (int m = 0; m < 3; ++m) { va[m] = gvec[laneA[m]]; vb[m] = gvec[laneB[m]]; }
                found = gf2_invertible3(va) && gf2_invertible3(vb);
            }
            if (!found) continue;
            gvec[0] = gvec[1] = gvec[2] = 0;                   // bits 0-2 are the column itself
            P1.h[0] = hx;
            P1.h[1] = hz;
            if (!order_and_apply_map1(P1.h[0], regB, laneB, P1.off1[0])) continue;
            if (!order_and_apply_map1(P1.h[1], regA, laneA, P1.off1[1])) continue;
            fill_map(P1.mapA, regA, laneA, gvec);
            fill_map(P1.mapB, regB, laneB, gvec);
            memcpy(P1.gvec, gvec, 9);
            for (int h = 0; h < 2; ++h)
                for (int k = 0; k < 4; ++k) {
                    P1.own[h][k] = 0;
                    for (int j = 0; j < P1.h[h].n_ops[k]; ++j)
                        if (P1.h[h].ops[k][j].flags & 4) P1.own[h][k] |= (uint16_t)(1u << P1.h[h].ops[k][j].d);
                }
            for (int m = 0; m < 5; ++m) {
                if (laneA[m] == sh) P1.sh_lane_bit = (uint8_t)m;
                if (laneB[m] == sh) P1.sh_lane_bit_b = (uint8_t)m;
            }
            {
                // the partner lane of the exchange load: the one whose label differs in bit sh
                const uint32_t k = (((1u << sh) ^ (sh >= 3 ? (uint32_t)gvec[sh] : 0u)) << 4);
                for (int i = 0; i < 16; ++i) {
                    P1.mapA.swzp16[i] = P1.mapA.swz16[i] ^ k;
                    P1.mapB.swzp16[i] = P1.mapB.swz16[i] ^ k;
                }
            }
            for (int h = 0; h < 2; ++h)
                for (int k = 0; k < 4; ++k)
                    for (int j = 0; j < P1.h[h].n_ops[k]; ++j) {             // packed per-slot words of dg_slot_entry
                        const DiagOp &op = P1.h[h].ops[k][j];
                        P1.ev4[h][k] |= (uint32_t)op.ev_word << (8 * j);
                        P1.pad4[h][k] |= (uint32_t)op.pad << (8 * j);
                        P1.f1s[h][k] |= (uint32_t)(op.flags & 1u) << (8 * j);
                        P1.f2s[h][k] |= (uint32_t)((op.flags >> 1) & 1u) << (8 * j);
                        P1.mm4[h][k] |= 1u << (8 * j);
                    }
            return true;
        }
    }
    return false;
}

// Two cycle kinds lower to the same whole-cycle program apart from where their events live (event words, ev4): then the
// launch of the first noisy cycle can run the noiseless preparation cycle of its shots first (k_cycle1, MeasArgs::with_prep).
static bool diag1_same_circuit(const DiagProg1 &A, const DiagProg1 &B) {
    for (int h = 0; h < 2; ++h) {
        const DiagHalf &a = A.h[h], &b = B.h[h];
        if (a.frame_x != b.frame_x || a.final_part != b.final_part || a.n_entries != b.n_entries) return false;
        if (memcmp(a.n_ops, b.n_ops, 4) || memcmp(a.aj, b.aj, 4) || memcmp(a.tbase, b.tbase, 4) ||
            memcmp(a.lane_bit, b.lane_bit, sizeof(a.lane_bit)))
            return false;
        for (int k = 0; k < 4; ++k)
            for (int j = 0; j < a.n_ops[k]; ++j)
                if (a.ops[k][j].d != b.ops[k][j].d || a.ops[k][j].flags != b.ops[k][j].flags || a.ops[k][j].pad != b.ops[k][j].pad)
                    return false;
        if (memcmp(A.own[h], B.own[h], sizeof(A.own[h])) || memcmp(A.off1[h], B.off1[h], sizeof(A.off1[h])) ||
            memcmp(A.pad4[h], B.pad4[h], sizeof(A.pad4[h])) || memcmp(A.f1s[h], B.f1s[h], sizeof(A.f1s[h])) ||
            memcmp(A.f2s[h], B.f2s[h], sizeof(A.f2s[h])) || memcmp(A.mm4[h], B.mm4[h], sizeof(A.mm4[h])))
            return false;
    }
    return !memcmp(&A.mapA, &B.mapA, sizeof(DiagMap)) && !memcmp(&A.mapB, &B.mapB, sizeof(DiagMap)) &&
           A.sh_lane_bit == B.sh_lane_bit && A.sh_lane_bit_b == B.sh_lane_bit_b && !memcmp(A.gvec, B.gvec, 9);
}

// ---- coherent over-rotations on the whole-cycle kernel ------------------------------------------------------------------
// With the over-rotation micro-ops taken out a coherent cycle is the Clifford cycle; its diagonal-frame form and label
// maps carry over (the rotations are functions of the same X_d X_a / X_d, or real rotations of one data qubit around its
// Ry bracket: DiagCoh in s17_diag.cuh).  What does not carry over is k_plan's Clifford conjugation of a Pauli event to
// the end of its operation and into the frame: it stays exact only for events that are a sign of the basis label or an
// ancilla Pauli where they act, or that come after the last gate of the data qubit in this cycle.
static bool coherent_events_ok(const s17_pinstr *plan, uint32_t n_plan) {
    for (uint32_t i = 0; i < n_plan; ++i) {
        const s17_pinstr &in = plan[i];
        if (in.kind == S17_P_IDLE) return false;             // an idle Z on a data qubit is a flip in the Hadamard frame
        if (in.kind != S17_P_ERR) continue;
        if (in.err == S17_ERR_LEAK_A || in.err == S17_ERR_LEAK_AB) continue;      // classical flags
        const bool after_bracket = in.field == 18;           // after Ry(-pi/2) and its over-rotation
        const bool sign_like = in.err == S17_ERR_X || in.err == S17_ERR_XX || in.err == S17_ERR_ZT;
        const bool data_pauli = in.err == S17_ERR_X || in.err == S17_ERR_Y || in.err == S17_ERR_Z || in.err == S17_ERR_ZC;
        if (!(sign_like || (after_bracket && data_pauli))) return false;
    }
    return true;
}

// strip the rotation micro-ops of every operation (the event field of a group moves to the gate the rotation followed)
static bool strip_rotations(const std::vector<LayerArg> &in, std::vector<LayerArg> &out) {
    bool any = false;
    out = in;
    for (LayerArg &A : out)
        for (int o = 0; o < A.L.n_ops; ++o) {
            s17_op &op = A.L.ops[o];
            if (op.kind != S17_OP_ENT) continue;
            int n = 0;
            for (int u = 0; u < op.n_uops; ++u) {
                if (op.uops[u].type >= S17_U_ROT_X) {
                    any = true;
                    if (n == 0) return false;
                    if (op.uops[u].ev_shift != 255) op.uops[n - 1].ev_shift = op.uops[u].ev_shift;
                    continue;
                }
                op.uops[n++] = op.uops[u];
            }
            op.n_uops = (uint8_t)n;
        }
    return any;
}

// the rotation angles of the operations of `layers`, laid out for the slots / maps of P1
static bool fill_coherent(const std::vector<LayerArg> &layers, const DiagProg1 &P1, DiagCoh &C) {
    double2 pre[9], post[9];
    for (int d = 0; d < 9; ++d) { pre[d] = make_double2(1.0, 0.0); post[d] = make_double2(1.0, 0.0); }
    for (int h = 0; h < 2; ++h)
        for (int k = 0; k < 4; ++k)
            for (int j = 0; j < 4; ++j) { C.rxx[h][k][j] = make_double2(1.0, 1.0); C.rx[h][k][j] = make_double2(1.0, 0.0); }
    for (int h = 0; h < 2; ++h)
        for (int k = 0; k < 4; ++k)
            for (int j = 0; j < P1.h[h].n_ops[k]; ++j) {
                const DiagOp &dop = P1.h[h].ops[k][j];
                const s17_op *op = nullptr;
                const LayerArg *lay = nullptr;
                for (const LayerArg &A : layers)
                    for (int o = 0; o < A.L.n_ops; ++o)
                        if (A.L.ops[o].kind == S17_OP_ENT && A.L.ops[o].ev_word == dop.ev_word) { op = &A.L.ops[o]; lay = &A; }
                if (!op) return false;
                const double s = (dop.flags & 1u) ? -1.0 : 1.0;            // sign of the Rxx angle
                double2 exx = make_double2(1.0, 0.0), ex = make_double2(1.0, 0.0);
                int last = -1;                                             // 0 Ry(+), 1 Rxx, 2 Rx, 3 Ry(-)
                for (int u = 0; u < op->n_uops; ++u) {
                    const s17_uop &up = op->uops[u];
                    if (up.type == S17_U_RY) last = up.sign > 0 ? 0 : 3;
                    else if (up.type == S17_U_RXX) last = 1;
                    else if (up.type == S17_U_RX) last = 2;
                    else if (up.type == S17_U_ROT_Y && (last == 0 || last == 3)) (last == 0 ? pre : post)[dop.d] = lay->prm[up.param];
                    else if (up.type == S17_U_ROT_XX && last == 1) exx = lay->prm[up.param];
                    else if (up.type == S17_U_ROT_X && last == 2) ex = lay->prm[up.param];
                    else return false;
                }
                // sqrt2 (cos, sin) of the half angle of Rxx(s pi/2 + eps) and of Rx(-s pi/2 + eps)
                C.rxx[h][k][j] = make_double2(exx.x - s * exx.y, s * exx.x + exx.y);
                if (dop.flags & 2u) C.rx[h][k][j] = make_double2(ex.x + s * ex.y, -s * ex.x + ex.y);
            }
    auto ab = [&](int d) { return make_double2(pre[d].x - pre[d].y, pre[d].x + pre[d].y); };     // Ry(eps) H = [[a, b], [b, -a]]
    for (int j = 0; j < 4; ++j) {
        C.preB[j] = ab(P1.mapB.reg[j]);
        C.preA[j] = ab(P1.mapA.reg[j]);
        C.postR[j] = post[P1.mapA.reg[j]];
    }
    C.preSh = ab(P1.mapA.lane[P1.sh_lane_bit]);
    for (int m = 0; m < 5; ++m) C.postL[m] = post[P1.mapA.lane[m]];
    return true;
}

// Host-only lowering of the compiled cycle: per-sweep constants, the fused half-cycle form, the diagonal-frame form and
// its label maps.  No CUDA call: s17_describe_program runs it without a device.
static int lower_program(s17_ctx *ctx, const s17_layer *layers, const uint32_t n_layers[3], const double *params,
                         uint32_t n_params, const s17_pinstr *plan, uint32_t n_plan) {
    const s17_layer *src = layers;
    {
        const char *co = getenv("S17_COH");               // 0: coherent models keep the sweep kernels
        ctx->no_coh = co && co[0] == '0';
    }
    for (int c = 0; c < 3; ++c) {
        ctx->layers[c].clear();
        for (uint32_t i = 0; i < n_layers[c]; ++i, ++src) {
            if (int rc = check_layer(ctx, *src)) return rc;
            LayerArg A;
            memset(&A, 0, sizeof(A));
            A.L = *src;
            int np = 0;
            for (int o = 0; o < A.L.n_ops; ++o)
                for (int u = 0; u < A.L.ops[o].n_uops; ++u) {
                    s17_uop &up = A.L.ops[o].uops[u];
                    if (up.type >= S17_U_ROT_X) {
                        if (up.param >= n_params || !params) return fail(ctx, S17_E_ARG, "uop param out of range");
                        if (np >= 64) return fail(ctx, S17_E_ARG, "too many rotation parameters in one layer");
                        A.prm[np] = make_double2(params[2 * up.param], params[2 * up.param + 1]);
                        up.param = (uint8_t)np++;
                    }
                }
            int k = 0;
            for (int b = 9; b <= 16; ++b) {
                bool in = false;
                for (int j = 0; j < A.L.n_anc; ++j) in |= (A.L.anc_bits[j] == b);
                if (!in) A.other_bits[k++] = (uint8_t)b;
            }
            ctx->layers[c].push_back(A);
        }
        ctx->layers2[c].clear();
        ctx->layers2d[c].clear();
        uint32_t support = 0;
        bool first = true;
        for (size_t i = 0; i < ctx->layers[c].size(); ++i) {
            const LayerArg &A = ctx->layers[c][i];
            ctx->layers2[c].push_back(make_layer2(A, support, first, false));
            ctx->layers2d[c].push_back(make_layer2(A, 0x1FE00u, first, true));
            for (int k = 0; k < A.L.n_anc; ++k) support |= 1u << A.L.anc_bits[k];
            first = false;
            if (A.L.meas_mask) { support = 0; first = true; }     // measured + reset: the support collapses again
        }
        if (ctx->layers[c].empty() || !ctx->layers[c].back().L.meas_mask)
            return fail(ctx, S17_E_ARG, "the last sweep of a cycle must carry a measurement mask");
        {
            // fused half-cycle form: needs exactly two measurement points (early X-ancilla measurement)
            std::vector<size_t> cuts;
            for (size_t i = 0; i < ctx->layers[c].size(); ++i)
                if (ctx->layers[c][i].L.meas_mask) cuts.push_back(i + 1);
            ctx->half_ok[c] = cuts.size() == 2 &&
                              make_half(ctx->layers[c], 0, cuts[0], ctx->half_prog[c][0], ctx->half_tail[c][0]) &&
                              make_half(ctx->layers[c], cuts[0], cuts[1], ctx->half_prog[c][1], ctx->half_tail[c][1]);
            if (ctx->half_ok[c]) { ctx->half_tail[c][0].final_part = 0; ctx->half_tail[c][1].final_part = 1; }
            // whole-cycle diagonal-frame form: Hadamard-frame half followed by a computational-frame half
            // (the data flips of the first half become signs the second half's tables must be able to carry)
            uint32_t cov0 = 0, cov1 = 0;
            ctx->diag_ok[c] = ctx->half_ok[c] &&
                              make_diag_half(ctx->half_prog[c][0], ctx->half_tail[c][0], 0, plan, n_plan, ctx->diag_prog[c].h[0], &cov0) &&
                              make_diag_half(ctx->half_prog[c][1], ctx->half_tail[c][1], 1, plan, n_plan, ctx->diag_prog[c].h[1], &cov1) &&
                              ctx->diag_prog[c].h[0].frame_x == 1 && ctx->diag_prog[c].h[1].frame_x == 0 &&
                              (cov0 & ~cov1) == 0;
            if (ctx->diag_ok[c]) { ctx->diag_prog[c].h[0].final_part = 0; ctx->diag_prog[c].h[1].final_part = 1; }
            ctx->diag1_ok[c] = ctx->diag_ok[c] && make_diag1(ctx->diag_prog[c], ctx->diag_prog1[c]);
            // a cycle with coherent over-rotations: lower its Clifford skeleton, then attach the angles
            ctx->coh_ok[c] = false;
            std::vector<LayerArg> skel;
            if (!ctx->diag1_ok[c] && !ctx->no_coh && cuts.size() == 2 && strip_rotations(ctx->layers[c], skel) &&
                coherent_events_ok(plan, n_plan)) {
                LayerArg2 hp[2];
                HalfTail ht[2];
                DiagProg G;
                uint32_t c0 = 0, c1 = 0;
                if (make_half(skel, 0, cuts[0], hp[0], ht[0]) && make_half(skel, cuts[0], cuts[1], hp[1], ht[1]) &&
                    make_diag_half(hp[0], ht[0], 0, plan, n_plan, G.h[0], &c0) && make_diag_half(hp[1], ht[1], 1, plan, n_plan, G.h[1], &c1) &&
                    G.h[0].frame_x == 1 && G.h[1].frame_x == 0 && (c0 & ~c1) == 0) {
                    G.h[0].final_part = 0;
                    G.h[1].final_part = 1;
                    if (make_diag1(G, ctx->diag_prog1[c]) && fill_coherent(ctx->layers[c], ctx->diag_prog1[c], ctx->diag_coh[c]))
                        ctx->coh_ok[c] = ctx->diag1_ok[c] = true;
                }
            }
        }
    }
    return S17_OK;
}

extern "C" int s17_load_program(s17_ctx *ctx, const s17_layer *layers, const uint32_t n_layers[3], const double *params,
                                uint32_t n_params, const s17_pinstr *plan, uint32_t n_plan, const s17_lists *lists) {
    if (!ctx || !layers || !n_layers || !plan || !lists) return fail(ctx, S17_E_ARG, "s17_load_program: bad arguments");
    if (!n_layers[0] || !n_layers[1] || !n_layers[2]) return fail(ctx, S17_E_ARG, "s17_load_program: empty sweep list");
    if (lists->n_lists > S17_MAX_LISTS) return fail(ctx, S17_E_ARG, "too many probability lists");
    CK(cudaSetDevice(ctx->device));
    if (int rc = lower_program(ctx, layers, n_layers, params, n_params, plan, n_plan)) return rc;
    if (n_plan > (uint32_t)kPlanMaxInstr) return fail(ctx, S17_E_ARG, "injection program too long");
    ctx->n_plan = (int)n_plan;
    memset(&ctx->nd, 0, sizeof(ctx->nd));
    ctx->nd.n_lists = lists->n_lists;
    uint32_t off = 0;
    for (uint32_t t = 0; t < lists->n_lists; ++t) {
        ctx->nd.list_len[t] = lists->list_len[t];
        ctx->nd.list_off[t] = off;
        off += lists->list_len[t];
    }
    if (off > S17_MASK_WORDS * 32 || off > 2048) return fail(ctx, S17_E_ARG, "too many error slots");
    ctx->total_slots = off;
    for (uint32_t i = 0; i < n_plan; ++i) {
        const s17_pinstr &in = plan[i];
        if (in.kind == S17_P_ERR || in.kind == S17_P_IDLE) {
            if (in.list >= lists->n_lists) return fail(ctx, S17_E_ARG, "plan entry names an unknown list");
            if (in.slot >= lists->list_len[in.list]) return fail(ctx, S17_E_ARG, "plan slot beyond its list (list too short)");
        }
        if (in.word >= S17_EV_WORDS) return fail(ctx, S17_E_ARG, "plan word out of range");
    }
    {
        // device copies of the program with every slot resolved to its absolute bit / probability index: one for a first
        // cycle, one for a second cycle (stab_ind=1: slots offset by half the list, CS:191-193)
        std::vector<s17_pinstr> res(2 * (size_t)n_plan);
        for (int c2 = 0; c2 < 2; ++c2)
            for (uint32_t i = 0; i < n_plan; ++i) {
                s17_pinstr in = plan[i];
                if (in.kind == S17_P_ERR || in.kind == S17_P_IDLE) {
                    const uint32_t L = lists->list_len[in.list];
                    in.slot = (uint16_t)(ctx->nd.list_off[in.list] + in.slot + ((c2 && in.use_round) ? (L >> 1) : 0u));
                }
                res[(size_t)c2 * n_plan + i] = in;
            }
        std::vector<uint32_t> hot(res.size());
        for (size_t i = 0; i < res.size(); ++i) {
            const bool entry = res[i].kind == S17_P_ERR || res[i].kind == S17_P_IDLE;
            hot[i] = entry ? ((uint32_t)res[i].slot | ((uint32_t)(res[i].in_skip ? 1u : 0u) << 16) | 0x20000u) : 0u;
        }
        if (ctx->plan) { cudaFree(ctx->plan); ctx->plan = nullptr; }
        CK(cudaMalloc(&ctx->plan, res.size() * sizeof(s17_pinstr)));
        CK(cudaMemcpy(ctx->plan, res.data(), res.size() * sizeof(s17_pinstr), cudaMemcpyHostToDevice));
        // groups of the program (one entangling operation / the idle entries of one gap), in program order
        std::vector<uint32_t> grp(4 * kPlanMaxGroups, 0u);     // per variant: 32 instruction ranges, 32 operation words
        std::vector<uint16_t> b2i(2 * 2048, 0xFFFFu);          // slot bit -> instruction of its entry (per variant)
        ctx->plan_dup_slot[0] = ctx->plan_dup_slot[1] = false;
        for (int c2 = 0; c2 < 2; ++c2) {
            int ng = 0;
            uint32_t i = 0;
            while (i < n_plan) {
                const s17_pinstr *q = &res[(size_t)c2 * n_plan];
                uint32_t j = i, meta = 0;
                if (q[i].kind == S17_P_OP_BEGIN) {
                    while (j < n_plan && q[j].kind != S17_P_OP_END) ++j;
                    if (j == n_plan) return fail(ctx, S17_E_ARG, "plan: operation without an end");
                    const s17_pinstr &e = q[j];
                    uint32_t sk = 0;
                    for (uint32_t k = i; k < j; ++k)
                        if (q[k].kind == S17_P_SKIP_TEST) sk = k - i;
                    if (sk == 0 || sk > 63u) ctx->plan_dup_slot[c2] = true;   // no (or a far) leak test: plain walk
                    meta = (sk << 17) | ((uint32_t)e.word << 1) | ((uint32_t)(plan[j].slot & 15u) << 6) | (((e.a >> 4) & 1u) << 10) |
                           (((e.a >> 5) & 1u) << 11) | (((e.a >> 3) & 1u) << 12) | (((e.a >> 2) & 1u) << 13) |
                           ((e.a & 1u) << 14) | (((e.a >> 1) & 1u) << 15) | ((e.b ? 1u : 0u) << 16);
                    ++j;
                } else if (q[i].kind == S17_P_IDLE) {
                    while (j < n_plan && q[j].kind == S17_P_IDLE && q[j].word == q[i].word) ++j;
                    meta = 1u | ((uint32_t)q[i].word << 1);
                } else {
                    return fail(ctx, S17_E_ARG, "plan: entry outside an operation");
                }
                if (ng >= kPlanMaxGroups) return fail(ctx, S17_E_ARG, "plan: too many groups");
                grp[(size_t)c2 * 2 * kPlanMaxGroups + ng] = i | (j << 10);
                grp[(size_t)(c2 * 2 + 1) * kPlanMaxGroups + ng] = meta;
                for (uint32_t k = i; k < j; ++k)
                    if (q[k].kind == S17_P_ERR || q[k].kind == S17_P_IDLE) {
                        uint16_t &slot = b2i[(size_t)c2 * 2048 + q[k].slot];
                        const bool lk_kind = q[k].kind == S17_P_ERR && (q[k].err == S17_ERR_LEAK_A || q[k].err == S17_ERR_LEAK_AB);
                        if (slot != 0xFFFFu) ctx->plan_dup_slot[c2] = true;   // one slot, two entries: k_plan walks everything
                        slot = (uint16_t)k;
                        hot[(size_t)c2 * n_plan + k] |= ((uint32_t)ng << 18) | (lk_kind ? (1u << 23) : 0u);
                    }
                ++ng;
                i = j;
            }
            ctx->n_grp[c2] = ng;
        }
        if (ctx->plan_hot) { cudaFree(ctx->plan_hot); ctx->plan_hot = nullptr; }
        CK(cudaMalloc(&ctx->plan_hot, hot.size() * 4));
        CK(cudaMemcpy(ctx->plan_hot, hot.data(), hot.size() * 4, cudaMemcpyHostToDevice));
        for (void *q : {(void *)ctx->plan_grp, (void *)ctx->plan_bit2grp, (void *)ctx->plan_surv})
            if (q) cudaFree(q);
        ctx->plan_grp = nullptr; ctx->plan_bit2grp = nullptr; ctx->plan_surv = nullptr;
        CK(cudaMalloc(&ctx->plan_grp, grp.size() * 4));
        CK(cudaMemcpy(ctx->plan_grp, grp.data(), grp.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&ctx->plan_bit2grp, b2i.size() * 2));
        CK(cudaMemcpy(ctx->plan_bit2grp, b2i.data(), b2i.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&ctx->plan_surv, 2 * (size_t)(kPlanMaxInstr + 1) * sizeof(double)));
        ctx->plan_res = res;
        ctx->h_probs.assign(2048, 0.0);
        ctx->geo_ok[0] = ctx->geo_ok[1] = false;
    }
    CK(cudaMemset(ctx->probs, 0, 2048 * sizeof(double)));
    ctx->prep_basis = -1;
    ctx->plan_has_leak = false;
    ctx->two_round_lists = true;
    for (uint32_t i = 0; i < n_plan; ++i)
        if ((plan[i].kind == S17_P_ERR || plan[i].kind == S17_P_IDLE) && plan[i].use_round &&
            (uint32_t)plan[i].slot + (lists->list_len[plan[i].list] >> 1) >= lists->list_len[plan[i].list])
            ctx->two_round_lists = false;
    for (uint32_t i = 0; i < n_plan; ++i)
        if (plan[i].kind == S17_P_ERR && (plan[i].err == S17_ERR_LEAK_A || plan[i].err == S17_ERR_LEAK_AB)) ctx->plan_has_leak = true;
    ctx->have_program = true;
    return S17_OK;
}

extern "C" int s17_describe_program(const s17_layer *layers, const uint32_t n_layers[3], const double *params, uint32_t n_params,
                                    const s17_pinstr *plan, uint32_t n_plan, s17_program_info *out) {
    if (!layers || !n_layers || !plan || !out) return fail(nullptr, S17_E_ARG, "s17_describe_program: bad arguments");
    if (!n_layers[0] || !n_layers[1] || !n_layers[2]) return fail(nullptr, S17_E_ARG, "s17_describe_program: empty sweep list");
    s17_ctx *ctx = new s17_ctx();                      // host-side scratch only: no device, no stream
    const int rc = lower_program(ctx, layers, n_layers, params, n_params, plan, n_plan);
    if (rc) { g_err = ctx->err; delete ctx; return rc; }
    memset(out, 0, sizeof(*out));
    for (int c = 0; c < 3; ++c) {
        out->n_sweeps[c] = (uint32_t)ctx->layers[c].size();
        out->half_kernel[c] = ctx->half_ok[c] ? 1 : 0;
        out->cycle_kernel[c] = ctx->diag_ok[c] ? 1 : 0;
        out->cycle_kernel_maps[c] = ctx->diag1_ok[c] ? 1 : 0;
        out->coherent[c] = ctx->coh_ok[c] ? 1 : 0;
        if (c == 1)
            out->prep_rides_in_cycle1 = (ctx->diag1_ok[0] && ctx->diag1_ok[1] && !ctx->coh_ok[0] && !ctx->coh_ok[1] &&
                                         diag1_same_circuit(ctx->diag_prog1[0], ctx->diag_prog1[1])) ? 1 : 0;
        if (!ctx->diag1_ok[c]) continue;
        const DiagProg1 &P = ctx->diag_prog1[c];
        for (int j = 0; j < 4; ++j) { out->reg_bits_z[c][j] = P.mapA.reg[j]; out->reg_bits_x[c][j] = P.mapB.reg[j]; }
        out->shuffle_bit[c] = P.mapA.lane[P.sh_lane_bit];
        for (int h = 0; h < 2; ++h) {
            out->table_entries[c][h] = P.h[h].n_entries;
            for (int k = 0; k < 4; ++k) out->slot_ancilla[c][h][k] = P.h[h].aj[k];
        }
    }
    delete ctx;
    return S17_OK;
}

extern "C" int s17_diag_slot_entry(uint32_t W, uint32_t f1, uint32_t f2, uint32_t idx, uint32_t cz, uint32_t mm, int32_t out[4]) {
    if (!out) return fail(nullptr, S17_E_ARG, "s17_diag_slot_entry: bad arguments");
    int a0r, a0i, a1r, a1i;
    dg_slot_entry(W, f1, f2, idx, cz, mm, a0r, a0i, a1r, a1i);
    out[0] = a0r; out[1] = a0i; out[2] = a1r; out[3] = a1i;
    return S17_OK;
}

extern "C" int s17_load_tables(s17_ctx *ctx, const uint32_t *lut256, const uint8_t *cw16) {
    if (!ctx || !lut256 || !cw16) return fail(ctx, S17_E_ARG, "s17_load_tables: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(ctx->lut, lut256, 256 * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->cw16, cw16, 16, cudaMemcpyHostToDevice));
    ctx->have_tables = true;
    return S17_OK;
}

extern "C" int s17_set_slot_probs(s17_ctx *ctx, uint32_t list_id, const double *p, uint32_t len) {
    if (!ctx || !ctx->have_program || !p) return fail(ctx, S17_E_ARG, "s17_set_slot_probs: bad arguments");
    if (list_id >= ctx->nd.n_lists || len != ctx->nd.list_len[list_id])
        return fail(ctx, S17_E_ARG, "make sure you pass enough probabilities to specify the error model");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(ctx->probs + ctx->nd.list_off[list_id], p, len * sizeof(double), cudaMemcpyHostToDevice));
    for (uint32_t i = 0; i < len; ++i) ctx->h_probs[ctx->nd.list_off[list_id] + i] = p[i];
    // survival products of the program's entries, per cycle variant (k_plan's geometric skipping).  Only used while a
    // cycle is mostly event-free (S_n > 1/4): with more noise than that the one-draw-per-entry walk is as cheap, and it
    // copes with probabilities of 1; also not when two entries share a slot (they would need separate draws).
    const int n = ctx->n_plan;
    std::vector<double> surv(2 * (size_t)(kPlanMaxInstr + 1), 1.0);
    for (int c2 = 0; c2 < 2; ++c2) {
        double *S = &surv[(size_t)c2 * (kPlanMaxInstr + 1)];
        for (int i = 0; i < n; ++i) {
            const s17_pinstr &in = ctx->plan_res[(size_t)c2 * n + i];
            const bool entry = in.kind == S17_P_ERR || in.kind == S17_P_IDLE;
            double q = entry ? ctx->h_probs[in.slot] : 0.0;
            q = q < 0.0 ? 0.0 : (q > 1.0 ? 1.0 : q);
            S[i + 1] = S[i] * (1.0 - q);
        }
        ctx->geo_ok[c2] = !ctx->plan_dup_slot[c2] && !ctx->no_geo && S[n] > 0.25;
    }
    CK(cudaMemcpy(ctx->plan_surv, surv.data(), surv.size() * sizeof(double), cudaMemcpyHostToDevice));
    return S17_OK;
}

extern "C" int s17_set_subset(s17_ctx *ctx, uint32_t n_types, const uint32_t *eset, const uint32_t *n_loc,
                              const double *const *weights) {
    if (!ctx || !ctx->have_program || !eset || !n_loc) return fail(ctx, S17_E_ARG, "s17_set_subset: bad arguments");
    if (n_types != ctx->nd.n_lists)
        return fail(ctx, S17_E_ARG, "make sure you pass enough probabilities to specify the error model");
    CK(cudaSetDevice(ctx->device));
    std::vector<double> cum(2048, 0.0);
    for (uint32_t t = 0; t < n_types; ++t) {
        if (n_loc[t] != ctx->nd.list_len[t]) return fail(ctx, S17_E_ARG, "subset locations do not match the program's lists");
        if (eset[t] > n_loc[t]) return fail(ctx, S17_E_ARG, "more errors than locations (the reference would loop forever)");
        ctx->nd.eset[t] = eset[t];
        ctx->nd.weighted[t] = (weights && weights[t]) ? 1u : 0u;
        if (ctx->nd.weighted[t]) {
            double tot = 0.0;
            uint32_t positive = 0;
            for (uint32_t i = 0; i < n_loc[t]; ++i) {
                tot += weights[t][i];                      // same running sum as generate_from_distr (CL:165-168)
                cum[ctx->nd.list_off[t] + i] = tot;
                positive += weights[t][i] > 0.0;
            }
            if (eset[t] > positive) return fail(ctx, S17_E_ARG, "more errors than locations with non-zero weight");
        }
    }
    CK(cudaMemcpy(ctx->cum, cum.data(), 2048 * sizeof(double), cudaMemcpyHostToDevice));
    return S17_OK;
}

extern "C" int s17_set_replay(s17_ctx *ctx, uint64_t n, const uint8_t *masks, uint64_t mstride, const uint8_t *outcomes,
                              uint64_t ostride) {
    if (!ctx || !ctx->have_program || n > ctx->max_shots) return fail(ctx, S17_E_ARG, "s17_set_replay: bad arguments");
    CK(cudaSetDevice(ctx->device));
    if (masks) {
        if (mstride < ctx->total_slots) return fail(ctx, S17_E_ARG, "mask stride shorter than the program's slots");
        std::vector<uint32_t> bits(n * 2 * S17_MASK_WORDS, 0u);
        for (uint64_t s = 0; s < n; ++s)
            for (uint32_t i = 0; i < ctx->total_slots; ++i)
                if (masks[s * mstride + i]) bits[s * 2 * S17_MASK_WORDS + (i >> 5)] |= 1u << (i & 31);
        CK(cudaMemcpy(ctx->masks, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice));
    }
    if (outcomes) {
        if (ostride > S17_MAX_OUTCOMES) return fail(ctx, S17_E_ARG, "outcome stride too long");
        std::vector<uint8_t> o(n * S17_MAX_OUTCOMES, 0);
        for (uint64_t s = 0; s < n; ++s) memcpy(&o[s * S17_MAX_OUTCOMES], outcomes + s * ostride, ostride);
        CK(cudaMemcpy(ctx->forced, o.data(), o.size(), cudaMemcpyHostToDevice));
    }
    ctx->have_replay = true;
    return S17_OK;
}

static inline int blocks_for(uint64_t n, int per) { return (int)((n + per - 1) / per); }

// The state buffer is allocated on first use.  The whole-cycle kernel (k_cycle) only ever touches the surviving
// 512-amplitude block of a shot, so a context that runs nothing else keeps 8 KiB per shot ("compact"); any other path
// (full-state sweeps, explicit collapse, parity read-backs of intermediate states) needs all 2^17 amplitudes.
static int ensure_state(s17_ctx *ctx, bool need_full) {
    if (ctx->state && (ctx->state_full || !need_full)) return S17_OK;
    const size_t B = ctx->max_shots;
    if (ctx->state) { CK(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->state); ctx->state = nullptr; }
    ctx->state_full = need_full;
    ctx->stride = need_full ? (size_t)kDim : (size_t)kRun;
    if (cudaMalloc(&ctx->state, B * ctx->stride * sizeof(double2)) != cudaSuccess) {
        cudaGetLastError();
        ctx->state = nullptr;
        return fail(ctx, S17_E_NOMEM, "out of device memory for the state vectors (" + std::to_string(B) + " shots x " +
                                          std::to_string(ctx->stride * 16) + " bytes)");
    }
    {
        const char *mm = getenv("S17_MEMSET");
        const char *mb = getenv("S17_MEMSET_BYTE");
        if (mm && (atoi(mm) & 1)) CK(cudaMemset(ctx->state, mb ? atoi(mb) : 0, B * ctx->stride * sizeof(double2)));
    }
    if (need_full) {
        if (!ctx->hist) {
            if (cudaMalloc(&ctx->hist, B * 256 * sizeof(double)) != cudaSuccess) {
                cudaGetLastError();
                return fail(ctx, S17_E_NOMEM, "out of device memory for the ancilla histograms");
            }
        }
        typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                     const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        const cuuint64_t dims[2] = {16, (cuuint64_t)B * (kDim / 8)};       // 16 doubles = 8 amplitudes = 128 B per row
        const cuuint64_t strides[1] = {128};
        const cuuint32_t box[2] = {16, kRun / 8};                          // one run: 64 rows = 8 KiB
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = ((EncodeFn)ctx->encode_fn)(&ctx->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, ctx->state, dims, strides,
                                                      box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(ctx, S17_E_CUDA, "cuTensorMapEncodeTiled failed");
    }
    return S17_OK;
}

// the whole-cycle kernel of cycle kind `which` (0 preparation, 1 stab_ind=0, 2 stab_ind=1): Clifford or coherent form
static void launch_cycle1(s17_ctx *ctx, int which, const MeasArgs &ma, double2 *state, size_t stride, const uint32_t *ev,
                          const uint32_t *frm, const uint32_t *list, const uint32_t *n_act, ShotAux *aux, const uint8_t *forced,
                          uint32_t *meas_out, uint32_t *dsamp, unsigned long long *counters, int init_basis,
                          const double2 *alt_src, double *pp_out, double *cdf_out) {
    if (init_basis >= 0) {
        // a cycle that starts from a basis state: the kernel writes its Walsh-Hadamard transform (a sign pattern) straight
        // into map B; the register part of every sign rides in bits 9-24
        const DiagMap &B = ctx->diag_prog1[which].mapB;
        uint32_t regpar = 0;
        for (int i = 0; i < 16; ++i)
            regpar |= (uint32_t)(__builtin_popcount((B.lin16[i] >> 4) & (uint32_t)init_basis) & 1) << i;
        init_basis = (init_basis & 511) | (int)(regpar << 9);
    }
    if (ctx->coh_ok[which])
        k_cycle1<true><<<ctx->num_sms, kDgThreads, kDgSmem1, ctx->stream>>>(ctx->diag_prog1[which], ctx->diag_coh[which], ma, state,
                                                                           stride, ev, frm, list, n_act, aux, forced, meas_out,
                                                                           dsamp, counters, init_basis, alt_src, pp_out, cdf_out);
    else
        k_cycle1<false><<<ctx->num_sms, kDgThreads, kDgSmem1, ctx->stream>>>(ctx->diag_prog1[which], ctx->diag_coh[which], ma, state,
                                                                            stride, ev, frm, list, n_act, aux, forced, meas_out,
                                                                            dsamp, counters, init_basis, alt_src, pp_out, cdf_out);
}

// device scratch of a host-side set-up step: released when the scope ends, whichever way it ends
struct DevScratch {
    std::vector<void *> ptrs;
    ~DevScratch() { for (void *q : ptrs) cudaFree(q); }
    template <class T> cudaError_t alloc(T **out, size_t bytes) {
        const cudaError_t e = cudaMalloc(out, bytes);
        if (e == cudaSuccess) ptrs.push_back(*out);
        return e;
    }
};

// Build the memoised preparation cycle for `basis`: run the production kernel once per outcome pattern (256 forced shots
// on private buffers), keep the collapsed blocks and the probability pair of every Measure along the way.
static int build_prep_cache(s17_ctx *ctx, int basis) {
    if (ctx->prep_basis == basis) return S17_OK;
    const DiagProg1 &P = ctx->diag_prog1[0];
    const int N = 256;
    uint32_t *ev = nullptr, *frm = nullptr, *list = nullptr, *nact = nullptr, *mo = nullptr, *ds = nullptr;
    uint8_t *forced = nullptr;
    ShotAux *aux = nullptr;
    double *pp = nullptr;
    unsigned long long *cnt = nullptr;
    DevScratch scratch;                               // freed on every exit path (CK returns early on a CUDA error)
    CK(scratch.alloc(&ev, N * S17_EV_WORDS * 4));
    CK(scratch.alloc(&frm, N * 16));
    CK(scratch.alloc(&list, N * 4));
    CK(scratch.alloc(&nact, 4));
    CK(scratch.alloc(&mo, N * 8));
    CK(scratch.alloc(&ds, N * 4));
    CK(scratch.alloc(&forced, N * S17_MAX_OUTCOMES));
    CK(scratch.alloc(&aux, N * sizeof(ShotAux)));
    CK(scratch.alloc(&pp, N * 16 * sizeof(double)));
    CK(scratch.alloc(&cnt, C_N * sizeof(unsigned long long)));
    CK(cudaMemset(ev, 0, N * S17_EV_WORDS * 4));
    CK(cudaMemset(frm, 0, N * 16));
    CK(cudaMemset(cnt, 0, C_N * sizeof(unsigned long long)));
    std::vector<uint32_t> h_list(N);
    std::vector<uint8_t> h_forced((size_t)N * S17_MAX_OUTCOMES, 0);
    PrepTree &T = ctx->prep_tree;
    memset(&T, 0, sizeof(T));
    for (int s = 0; s < 8; ++s) T.aj[s] = P.h[s >> 2].aj[s & 3];
    T.final_part[0] = P.h[0].final_part;
    T.final_part[1] = P.h[1].final_part;
    for (int L = 0; L < N; ++L) {
        h_list[L] = (uint32_t)L;
        for (int s = 0; s < 8; ++s) h_forced[(size_t)L * S17_MAX_OUTCOMES + T.aj[s]] = (uint8_t)((L >> s) & 1);
    }
    const uint32_t n_items = N;
    CK(cudaMemcpy(list, h_list.data(), N * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(nact, &n_items, 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(forced, h_forced.data(), h_forced.size(), cudaMemcpyHostToDevice));
    MeasArgs ma;
    memset(&ma, 0, sizeof(ma));
    ma.stage = 0;
    ma.forced = 1;
    ma.mask = 0xFFu;
    ma.final_part = 1;
    ma.stream = 16u;
    ma.n_shots = N;
    launch_cycle1(ctx, 0, ma, ctx->prep_blocks, (size_t)kRun, ev, frm, list, nact, aux, forced, mo, ds, cnt, basis, nullptr, pp, nullptr);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    std::vector<double> h_pp((size_t)N * 16);
    CK(cudaMemcpy(h_pp.data(), pp, h_pp.size() * sizeof(double), cudaMemcpyDeviceToHost));
    // node (s, prefix) takes its pair from leaf = prefix: that leaf's path runs through the node, and if the node is
    // reachable every earlier step of the path was possible
    std::vector<double> node(2 * 255, 0.0);
    for (int s = 0; s < 8; ++s)
        for (int prefix = 0; prefix < (1 << s); ++prefix) {
            node[2 * (((1 << s) - 1) + prefix)] = h_pp[(size_t)prefix * 16 + 2 * s];
            node[2 * (((1 << s) - 1) + prefix) + 1] = h_pp[(size_t)prefix * 16 + 2 * s + 1];
        }
    CK(cudaMemcpy(ctx->prep_pp, node.data(), node.size() * sizeof(double), cudaMemcpyHostToDevice));

    // The event-free first cycle per leaf (k_quiet): run the first noisy cycle's program with no event word set on every
    // reachable leaf, its outcomes forced to the leaf's own.  The leaf is memoised only if the kernel itself finds
    // every other outcome to have probability exactly zero (then its decisions do not depend on the uniforms).
    std::vector<uint8_t> quiet_ok(N, 0);
    ctx->any_quiet_leaf = false;
    if (ctx->diag1_ok[1]) {
        std::vector<uint32_t> h_mo((size_t)N * 2);
        CK(cudaMemcpy(h_mo.data(), mo, h_mo.size() * 4, cudaMemcpyDeviceToHost));
        std::vector<uint8_t> reach(N, 0);
        for (int L = 0; L < N; ++L) reach[L] = !((h_mo[2 * L] | h_mo[2 * L + 1]) & 0x100u);    // leaf was not impossible
        for (int L = 0; L < N; ++L) {
            h_list[L] = (uint32_t)L | ((uint32_t)L << 24);                                        // input block = leaf L
            for (int s = 0; s < 8; ++s) h_forced[(size_t)L * S17_MAX_OUTCOMES + 8 + T.aj[s]] = (uint8_t)((L >> s) & 1);
        }
        CK(cudaMemcpy(list, h_list.data(), N * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(forced, h_forced.data(), h_forced.size(), cudaMemcpyHostToDevice));
        double *cdf = ctx->quiet_cdf;
        ma.stage = 1;
        ma.stream = 16u + 4u;
        ma.sample_data = 1;
        ma.data_stream = 32u;
        launch_cycle1(ctx, 1, ma, ctx->quiet_blocks, (size_t)kRun, ev, frm, list, nact, aux, forced, mo, ds, cnt, -1,
                      ctx->prep_blocks, pp, cdf);
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaGetLastError());
        CK(cudaMemcpy(h_pp.data(), pp, h_pp.size() * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h_mo.data(), mo, h_mo.size() * 4, cudaMemcpyDeviceToHost));
        {
            // k_quiet replays the kernel's read-out draw with one thread per shot: turn each leaf's per-state weights into
            // the running sums the kernel's lanes form (same additions in the same order: the previous lane's inclusive
            // sum, then the lane's sixteen weights one by one), and note each lane's last state of positive weight
            std::vector<double> h_cdf((size_t)N * kDgCdfDoubles);
            std::vector<uint8_t> h_last((size_t)N * 32, 0);
            CK(cudaMemcpy(h_cdf.data(), cdf, h_cdf.size() * sizeof(double), cudaMemcpyDeviceToHost));
            for (int L = 0; L < N; ++L) {
                double *c = &h_cdf[(size_t)L * kDgCdfDoubles];
                for (int lane = 0; lane < 32; ++lane) {
                    double cum = lane ? c[lane - 1] : 0.0;
                    for (int i = 0; i < 16; ++i) {
                        const double w = c[32 + 16 * lane + i];
                        if (w > 0.0) h_last[(size_t)L * 32 + lane] = (uint8_t)i;
                        cum += w;
                        c[32 + 16 * lane + i] = cum;
                    }
                }
            }
            CK(cudaMemcpy(cdf, h_cdf.data(), h_cdf.size() * sizeof(double), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(ctx->quiet_last, h_last.data(), h_last.size(), cudaMemcpyHostToDevice));
        }
        for (int L = 0; L < N; ++L) {
            bool ok = reach[L] && !((h_mo[2 * L] | h_mo[2 * L + 1]) & 0x100u);
            for (int s = 0; s < 8 && ok; ++s) {
                const int out = (L >> s) & 1;
                const double p_other = h_pp[(size_t)L * 16 + 2 * s + (out ? 0 : 1)], p_own = h_pp[(size_t)L * 16 + 2 * s + out];
                ok = p_other == 0.0 && p_own > 0.0;
            }
            quiet_ok[L] = ok ? 1 : 0;
            ctx->any_quiet_leaf |= ok;
        }
    }
    CK(cudaMemcpy(ctx->quiet_ok, quiet_ok.data(), N, cudaMemcpyHostToDevice));
    ctx->prep_basis = basis;
    return S17_OK;
}

// one stabiliser cycle: plan -> gate sweeps -> ancilla measurement decisions
static void collapse(s17_ctx *ctx, int n, const uint8_t *flag);

// the per-shot reset of a call, when its first kernel is not k_plan
static void flush_reset(s17_ctx *ctx, int n) {
    if (!ctx->pend_reset) return;
    ctx->pend_reset = false;
    SCOPE2(CAT_OTHER, S17_T_PLACE);
    k_reset<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(n, ctx->leak, ctx->active, ctx->ready, ctx->rec, ctx->masks, 0, 0, 0, ctx->nd,
                                                         ctx->cum, 0ull, 0ull);
}

// apply a deferred cycle bookkeeping now (something other than k_plan / k_readout_s is about to read the shot flags)
static void flush_record(s17_ctx *ctx, int n) {
    if (!ctx->pend_rec) return;
    ctx->pend_rec = false;
    SCOPE2(CAT_MEAS, S17_T_RECORD);
    k_record<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(ctx->pend_ma, ctx->meas_out, ctx->lut, ctx->leak, ctx->active, ctx->ready,
                                                          ctx->rec, ctx->counters);
}

static int run_cycle(s17_ctx *ctx, const s17_run_args &a, int stage, int second, int init_basis, uint32_t max_layers, bool dense,
                     bool mid_collapse) {
    const int n = (int)a.n_shots;
    if (stage == 0 && ctx->use_prep) {
        flush_reset(ctx, n);
        // memoised noiseless preparation cycle: replay the kernel's decisions per shot, no state traffic
        MeasArgs ma;
        memset(&ma, 0, sizeof(ma));
        ma.stage = 0;
        ma.protocol = (int)a.protocol;
        ma.forced = (int)a.forced;
        ma.seed = a.seed;
        ma.first_shot = a.first_shot_id;
        ma.n_shots = n;
        ma.mask = 0xFFu;
        ma.final_part = 1;
        ma.stream = 16u;
        {
            SCOPE(CAT_OTHER);      // (the gate category times the amplitude kernels only: it is the roofline's denominator)
            k_prep<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(ctx->prep_tree, ma, ctx->prep_pp, ctx->forced, ctx->active,
                                                                ctx->meas_out, ctx->prep_leaf, ctx->aux, ctx->counters,
                                                                ctx->prep_mixed ? ctx->leak : nullptr, ctx->prep_done);
        }
        if (!ctx->prep_mixed) {
            ctx->pend_rec = true;          // applied by the next cycle's k_plan
            ctx->pend_ma = ma;
            ctx->last_collapsed = true;
            return S17_OK;
        }
        // mixed: the shots with a leak flag go through the regular preparation cycle below (k_plan lists only them)
    }
    const int which0 = stage == 0 ? 0 : (second ? 2 : 1);
    // A preparation cycle of a plain run with a model that has no leakage entries: every shot takes part, none has an
    // event or a set leak flag -- nothing to plan, no list to build; the reset of the call waits for the next k_plan.
    const bool trivial_prep = stage == 0 && !ctx->plan_has_leak && a.protocol != S17_PROTO_RTF && ctx->diag1_ok[which0] &&
                              !ctx->no_diag && !dense && !mid_collapse && !max_layers && !ctx->force_ring && !ctx->no_fuse &&
                              !ctx->pend_rec && !ctx->no_trivial_prep;
    if (stage == 0) {
        // Un-memoised run: the preparation cycle needs no launch of its own -- the launch of the first noisy cycle starts every
        // shot from the basis state and runs the noiseless cycle first, block in registers (no 8 KiB store + load, no
        // second pass over the per-shot prologue); its outcomes reach the bookkeeping through the same meas_out words.
        ctx->fuse_prep = trivial_prep && !ctx->no_fuse_prep && !ctx->use_prep && !a.forced && !a.materialize &&
                         a.stop_stage == S17_STOP_NONE && ctx->diag1_ok[1] && !ctx->coh_ok[0] && !ctx->coh_ok[1] &&
                         diag1_same_circuit(ctx->diag_prog1[0], ctx->diag_prog1[1]);
        if (ctx->fuse_prep) {
            ctx->fuse_basis = init_basis;
            return S17_OK;
        }
    }
    PlanArgs pa;
    pa.n_instr = ctx->n_plan;
    pa.noise = (int)a.noise;
    pa.noiseless = (stage == 0);
    pa.second = second;
    pa.rec_half = (stage == 2) ? 1 : 0;
    // fixed-weight noise: the placement (generate_error_locations, CL:151-181) is drawn where the first noisy cycle is
    // planned; a run-til-fail job places once, when it begins (k_reset)
    pa.place = (a.noise == S17_NOISE_SUBSET && stage == 1 && a.protocol != S17_PROTO_RTF) ? 1 : 0;
    pa.stream = 8u + (uint32_t)stage;
    pa.seed = a.seed;
    pa.first_shot = a.first_shot_id;
    pa.n_shots = n;
    if (!trivial_prep) {
        StepArgs sa;
        memset(&sa, 0, sizeof(sa));
        sa.do_record = ctx->pend_rec ? 1 : 0;
        if (ctx->pend_rec) sa.rec_ma = ctx->pend_ma;
        ctx->pend_rec = false;
        sa.meas_out = ctx->meas_out;
        sa.lut = ctx->lut;
        sa.ready = ctx->ready;
        sa.rec = ctx->rec;
        sa.counters = ctx->counters;
        sa.do_compact = 1;
        sa.list = ctx->act_list;
        sa.leaf = (ctx->use_prep && stage == 1) ? ctx->prep_leaf : nullptr;
        sa.quiet_ok = ctx->quiet_ok;
        sa.done = (ctx->use_prep && stage <= 1) ? ctx->prep_done : nullptr;
        sa.skip_done = stage == 0 ? 1 : 0;
        sa.walk_all = ctx->plan_dup_slot[second ? 1 : 0] ? 1 : 0;
        sa.do_reset = ctx->pend_reset ? 1 : 0;
        sa.keep_leak = 0;
        ctx->pend_reset = false;
        sa.nd = ctx->nd;
        sa.cum = ctx->cum;
        // the list length of this cycle is the counter the previous k_plan cleared; this one clears the other
        ctx->n_act = ctx->n_act2 + (ctx->plan_seq & 1u);
        uint32_t *n_next = ctx->n_act2 + ((ctx->plan_seq + 1u) & 1u);
        ++ctx->plan_seq;
        SCOPE2(CAT_OTHER, S17_T_PLAN);
        const bool geo = a.noise == S17_NOISE_PROB && ctx->geo_ok[second ? 1 : 0];
        const size_t smem = kPlanSmem - (geo ? 0 : (size_t)(kPlanMaxInstr + 2) * 8);          // 59 / 54 KiB: 3 / 4 CTAs per SM
        k_plan<<<std::min(blocks_for(n, kPlanThreads), ctx->num_sms * (geo ? 3 : 4)), kPlanThreads, smem, ctx->stream>>>(
            ctx->plan + (second ? ctx->n_plan : 0), pa, sa, ctx->probs, ctx->ev, ctx->masks, ctx->leak, ctx->active, ctx->frm,
            ctx->n_act, n_next, ctx->plan_hot + (second ? ctx->n_plan : 0),
            (ctx->use_quiet && stage == 1) ? ctx->quiet_flag : nullptr, ctx->plan_grp + (second ? 2 * kPlanMaxGroups : 0),
            ctx->n_grp[second ? 1 : 0], ctx->plan_bit2grp + (second ? 2048 : 0),
            geo ? ctx->plan_surv + (second ? kPlanMaxInstr + 1 : 0) : nullptr);
    }
    const int which = stage == 0 ? 0 : (second ? 2 : 1);
    const std::vector<LayerArg> &Ls = ctx->layers[which];
    const std::vector<LayerArg2> &Ls2 = dense ? ctx->layers2d[which] : ctx->layers2[which];
    uint32_t nl = (uint32_t)Ls.size();
    if (max_layers && max_layers < nl) nl = max_layers;
    MeasArgs ma;
    memset(&ma, 0, sizeof(ma));
    ma.stage = stage;
    ma.protocol = (int)a.protocol;
    ma.forced = (int)a.forced;
    ma.seed = a.seed;
    ma.first_shot = a.first_shot_id;
    ma.n_shots = n;
    const bool fused_ok = !dense && !mid_collapse && !max_layers && !ctx->force_ring && !ctx->no_fuse;
    if (ctx->diag1_ok[which] && fused_ok && !ctx->no_diag) {
        const bool with_prep = stage == 1 && ctx->fuse_prep;
        ctx->fuse_prep = false;
        if (with_prep) init_basis = ctx->fuse_basis;
        if (init_basis < 0 && !ctx->last_collapsed) collapse(ctx, n, ctx->active);     // k_cycle reads the block at ancilla value 0
        ctx->last_collapsed = true;
        ma.with_prep = with_prep ? 1 : 0;
        // production path for Clifford cycles with Pauli / leakage noise: the whole cycle in the frame where it is
        // diagonal, one warp per shot, the state in registers (s17_diag.cuh)
        ma.mask = 0xFFu;
        ma.final_part = 1;
        ma.stream = 16u + (uint32_t)stage * 4u;
        // only cycles after which some shot is decoded: s2 decodes after its second cycle only (CL:424), s1 and the
        // single-round protocol after the first; run-til-fail after either
        const bool decodes = stage == 2 ? true : (stage == 1 && a.protocol != S17_PROTO_S2);
        ma.sample_data = (ctx->use_dsamp && decodes) ? 1 : 0;
        ma.data_stream = 32u;
        {
            SCOPE(CAT_GATE);
            launch_cycle1(ctx, which, ma, ctx->state, ctx->stride, ctx->ev, ctx->frm, trivial_prep ? nullptr : ctx->act_list,
                          ctx->n_act, ctx->aux,
                          ctx->forced, ctx->meas_out, ctx->dsamp, ctx->counters, init_basis,
                          (ctx->use_prep && stage == 1) ? ctx->prep_blocks : nullptr, nullptr, nullptr);
            ctx->n_gate++;
        }
        if (ctx->use_quiet && stage == 1) {
            // the shots k_plan left out of the list: no event in this cycle, memoised leaf.  meas_out still holds the leaf's
            // outcomes (k_prep wrote them), which are this cycle's outcomes too.
            // Their state is the cached block of their leaf: copied into the shot's slot only when something on the device
            // will read it (a replay / a state-reading read-out, or S17_QUIET_COPY=1); s17_read_state finds it either way.
            const int copy = (ctx->quiet_copy || a.forced || !ctx->use_dsamp) ? 1 : 0;
            ctx->quiet_lazy = !copy;
            SCOPE(CAT_OTHER);
            k_quiet<<<blocks_for(copy ? (uint64_t)n * 32 : (uint64_t)n, 128), 128, 0, ctx->stream>>>(
                ctx->prep_tree, ma, ctx->diag_prog1[1].mapA, ctx->active, ctx->quiet_flag, ctx->quiet_ok, ctx->prep_leaf, ctx->quiet_blocks,
                ctx->quiet_cdf, ctx->quiet_last, ctx->forced, ctx->state, ctx->stride, ctx->meas_out, ctx->dsamp, ctx->counters, copy);
        }
        // the cycle's bookkeeping (outcome log, leak override, protocol branch, decode) is applied per shot by whichever
        // kernel touches the shots next: the next cycle's k_plan, k_readout_s, or flush_record
        ctx->pend_rec = true;
        ctx->pend_ma = ma;
        return S17_OK;
    }
    if (ctx->half_ok[which] && fused_ok) {
        ctx->last_collapsed = true;
        // production path: one fused kernel per half cycle (gates + measurement + collapse), state stays on chip
        for (int h = 0; h < 2; ++h) {
            ma.mask = ctx->half_tail[which][h].meas_mask;
            ma.final_part = h;
            ma.stream = 16u + (uint32_t)stage * 4u + (h ? 0u : 1u);
            SCOPE(CAT_GATE);
            k_half<<<ctx->num_sms * kL3CtasPerSm, kL3Threads, kL3Smem, ctx->stream>>>(
                ctx->tmap, ctx->half_prog[which][h], ctx->half_tail[which][h], ma, ctx->state, ctx->ev, ctx->act_list,
                ctx->n_act, ctx->aux, ctx->forced, ctx->meas_out, ctx->counters, (h == 0) ? init_basis : -1);
            ctx->n_gate++;
        }
        {
            ma.mask = 0xFFu;
            ma.final_part = 1;
            SCOPE2(CAT_MEAS, S17_T_RECORD);
            k_record<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(ma, ctx->meas_out, ctx->lut, ctx->leak, ctx->active,
                                                                  ctx->ready, ctx->rec, ctx->counters);
        }
        return S17_OK;
    }
    ctx->last_collapsed = false;
    for (uint32_t i = 0; i < nl; ++i) {
        {
            SCOPE(CAT_GATE);
            if (!dense && !ctx->force_ring)
                k_layer3<<<ctx->num_sms * kL3CtasPerSm, kL3Threads, kL3Smem, ctx->stream>>>(
                    ctx->tmap, Ls2[i], ctx->ev, ctx->act_list, ctx->n_act, ctx->hist, ctx->counters, ctx->aux,
                    (i == 0) ? init_basis : -1);
            else
                k_layer2<<<ctx->num_sms, kL2Threads, kL2Smem, ctx->stream>>>(
                    ctx->tmap, ctx->state, Ls2[i], ctx->ev, ctx->act_list, ctx->n_act, ctx->hist, ctx->counters, ctx->aux,
                    ctx->zero_page, (i == 0) ? init_basis : -1);
            ctx->n_gate++;
        }
        if (!Ls[i].L.meas_mask) continue;
        const bool final_part = (i + 1 == Ls.size());
        ma.mask = Ls[i].L.meas_mask;
        ma.final_part = final_part ? 1 : 0;
        ma.stream = 16u + (uint32_t)stage * 4u + (final_part ? 0u : 1u);
        {
            SCOPE2(CAT_MEAS, S17_T_SAMPLE);
            k_measure<<<blocks_for((uint64_t)n * 32, 128), 128, 0, ctx->stream>>>(
                ma, ctx->hist, ctx->forced, ctx->lut, ctx->leak, ctx->active, ctx->ready, ctx->aux, ctx->rec, ctx->counters);
        }
        // a partial measurement inside the cycle: materialise its collapse when the caller wants exact states
        if (!final_part && mid_collapse) collapse(ctx, n, ctx->active);
    }
    return S17_OK;
}

static void collapse(s17_ctx *ctx, int n, const uint8_t *flag) {
    SCOPE2(CAT_MEAS, S17_T_COLLAPSE);
    k_collapse<<<n * 32, 256, 0, ctx->stream>>>(ctx->state, ctx->aux, flag, ctx->counters);
    ctx->last_collapsed = true;
}

// one full round for every active slot: prep, cycle A, optional cycle B, decode, read-out
static int run_round(s17_ctx *ctx, const s17_run_args &a) {
    const int n = (int)a.n_shots;
    const int basis_index = a.logical_state ? 0x1FF : 0;     // All(X) | data when state == 1 (CS:103-104)
    // dense: every sweep covers the full state and ProjectQ's collapse is materialised after each measurement.
    // otherwise the collapse + reset is lazy: the next sweep reads the surviving block, scaled, and sweeps only
    // cover the part of the state that can be non-zero.
    const bool dense = ctx->dense;
    const bool explicit_collapse = dense || a.materialize || a.stop_stage != S17_STOP_NONE;
    // both noisy cycles go through k_cycle1 and outcomes are sampled: the cycle kernel also draws the data read-out
    ctx->use_dsamp = ctx->diag1_ok[1] && ctx->diag1_ok[2] && !ctx->no_diag && !explicit_collapse &&
                     !ctx->force_ring && !ctx->no_fuse && !a.forced;
    // the noiseless preparation cycle is memoised when it and the first noisy cycle run on k_cycle1; where a leak flag can
    // be set when it starts (run-til-fail keeps leak flags across the rounds of a run) only for the shots without one
    ctx->use_prep = ctx->diag1_ok[0] && ctx->diag1_ok[1] && !ctx->no_diag && !explicit_collapse &&
                    !ctx->force_ring && !ctx->no_fuse && !ctx->no_prep_cache && ctx->max_shots < (1u << 23);
    ctx->prep_mixed = ctx->use_prep && a.protocol == S17_PROTO_RTF && ctx->plan_has_leak;
    if (ctx->use_prep)
        if (int rc = build_prep_cache(ctx, basis_index)) return rc;
    // an event-free first cycle of a shot that sits in a memoised preparation leaf is replayed from the cache as well
    ctx->use_quiet = ctx->use_prep && !ctx->no_quiet && ctx->any_quiet_leaf;
    ctx->pend_rec = false;
    ctx->fuse_prep = false;
    {
        Nvtx r("s17: preparation cycle");
        if (int rc = run_cycle(ctx, a, 0, 0, basis_index, 0, dense, explicit_collapse)) return rc;
    }
    if (explicit_collapse) { flush_record(ctx, n); collapse(ctx, n, ctx->active); }
    if (a.stop_stage == S17_STOP_AFTER_PREP) { flush_record(ctx, n); return S17_OK; }
    if (a.stop_stage == S17_STOP_CYCLE1_GATES) {
        const int rc = run_cycle(ctx, a, 1, 0, -1, a.stop_layers ? a.stop_layers : 1000, dense, explicit_collapse);
        flush_record(ctx, n);
        return rc;
    }
    {
        Nvtx r("s17: first noisy cycle");
        if (int rc = run_cycle(ctx, a, 1, 0, -1, 0, dense, explicit_collapse)) return rc;
    }
    if (a.stop_stage == S17_STOP_AFTER_CYCLE1) {
        flush_record(ctx, n);
        SCOPE(CAT_OTHER);
        cudaMemsetAsync(ctx->active, 1, n, ctx->stream);
        collapse(ctx, n, ctx->active);
        return S17_OK;
    }
    if (a.protocol == S17_PROTO_S2 || a.protocol == S17_PROTO_RTF) {
        if (explicit_collapse) { flush_record(ctx, n); collapse(ctx, n, ctx->active); }
        // s2: stab_ind=1 slots (CL:422); run-til-fail: cycle B reuses the stab_ind=0 probabilities (CL:92)
        Nvtx r("s17: second noisy cycle");
        if (int rc = run_cycle(ctx, a, 2, a.protocol == S17_PROTO_S2 ? 1 : 0, -1, 0, dense, explicit_collapse)) return rc;
    }
    if (a.materialize) {
        flush_record(ctx, n);
        collapse(ctx, n, ctx->ready);
        SCOPE(CAT_OTHER);
        k_pauli<<<n * 32, 256, 0, ctx->stream>>>(ctx->state, ctx->rec, ctx->ready);
    }
    Nvtx r_read("s17: decode + read-out");
    ReadArgs ra;
    ra.forced = (int)a.forced;
    ra.folded = a.materialize ? 0 : 1;
    ra.logical_state = (int)a.logical_state;
    ra.stream = 32u;
    ra.seed = a.seed;
    ra.first_shot = a.first_shot_id;
    ra.n_shots = n;
    if (ctx->use_dsamp) {
        const int do_record = ctx->pend_rec ? 1 : 0;          // the last cycle's bookkeeping rides in the read-out kernel
        ctx->pend_rec = false;
        SCOPE2(CAT_MEAS, S17_T_READOUT);
        k_readout_s<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(ra, ctx->dsamp, ctx->cw16, ctx->leak, ctx->ready, ctx->rec,
                                                                 ctx->counters, do_record, ctx->pend_ma, ctx->meas_out, ctx->lut,
                                                                 ctx->active);
    } else {
        flush_record(ctx, n);
        SCOPE2(CAT_MEAS, S17_T_READOUT);
        k_readout<<<blocks_for((uint64_t)n * 32, 128), 128, 0, ctx->stream>>>(ra, ctx->state, ctx->stride, ctx->aux, ctx->forced, ctx->cw16,
                                                                          ctx->leak, ctx->ready, ctx->rec, ctx->counters);
    }
    return S17_OK;
}

static int finish_timing(s17_ctx *ctx) {
    if (!ctx->debug_note.empty()) return fail(ctx, S17_E_CUDA, ctx->debug_note);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    for (const TimedLaunch &tl : ctx->launches) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, tl.e0, tl.e1);
        (tl.cat == CAT_GATE ? ctx->t_gate : tl.cat == CAT_MEAS ? ctx->t_meas : ctx->t_other) += ms;
        if (tl.sub >= 0 && tl.sub < S17_T_N) { ctx->t_sub[tl.sub] += ms; ctx->n_sub[tl.sub]++; }
    }
    ctx->n_all += ctx->launches.size();
    ctx->launches.clear();
    ctx->ev_used = 0;
    return S17_OK;
}

static int check_run_args(s17_ctx *ctx, const s17_run_args &a) {
    if (!ctx->have_program || !ctx->have_tables) return fail(ctx, S17_E_STATE, "s17_run: load the program and tables first");
    if (a.n_shots == 0 || a.n_shots > ctx->max_shots) return fail(ctx, S17_E_ARG, "s17_run: n_shots out of range");
    if (a.protocol > S17_PROTO_RTF || a.noise > S17_NOISE_REPLAY) return fail(ctx, S17_E_ARG, "s17_run: bad protocol/noise");
    if (a.leak_reset > S17_LEAK_RESET_ROUND) return fail(ctx, S17_E_ARG, "s17_run: bad leak_reset");
    if (a.basis_x) return fail(ctx, S17_E_ARG, "s17_run: X-basis preparation is not implemented (dead in the reference, CL:274)");
    if ((a.noise == S17_NOISE_REPLAY || a.forced) && !ctx->have_replay)
        return fail(ctx, S17_E_STATE, "s17_run: replay requested but s17_set_replay was not called");
    if (a.protocol == S17_PROTO_S2 && !ctx->two_round_lists)
        return fail(ctx, S17_E_ARG, "s17_run: the s2 protocol indexes the second half of every probability list (CS:191-193); "
                                    "load the program with two-round list lengths");
    if (ctx->rtf_open && a.protocol != S17_PROTO_RTF) return fail(ctx, S17_E_STATE, "s17_run: a run-til-fail job is open (s17_rtf_end)");
    return S17_OK;
}

static int begin_call(s17_ctx *ctx, const s17_run_args &a) {
    CK(cudaSetDevice(ctx->device));
    ctx->t_gate = ctx->t_meas = ctx->t_other = 0;
    for (int i = 0; i < S17_T_N; ++i) { ctx->t_sub[i] = 0; ctx->n_sub[i] = 0; }
    ctx->n_gate = ctx->n_all = 0;
    ctx->launches.clear();
    ctx->ev_used = 0;
    const bool all_diag = ctx->diag1_ok[0] && ctx->diag1_ok[1] && ctx->diag1_ok[2] && !ctx->no_diag && !ctx->dense &&
                          !ctx->force_ring && !ctx->no_fuse && !a.materialize &&
                          a.stop_stage == S17_STOP_NONE;
    if (int rc = ensure_state(ctx, !all_diag)) return rc;
    CK(cudaEventRecord(ctx->span0, ctx->stream));
    CK(cudaMemsetAsync(ctx->counters, 0, C_N * sizeof(unsigned long long), ctx->stream));
    const int n = (int)a.n_shots;
    // a plain run starts with the preparation cycle's k_plan, which resets the shots itself (flush_reset when the memoised
    // preparation runs instead); true-probability runs also clear both halves of the mask record (256 B per shot) and a
    // run-til-fail job places its fixed-weight errors once: k_reset
    ctx->pend_reset = a.protocol != S17_PROTO_RTF && a.noise != S17_NOISE_PROB;
    if (!ctx->pend_reset) {
        SCOPE2(CAT_OTHER, S17_T_PLACE);
        k_reset<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(n, ctx->leak, ctx->active, ctx->ready, ctx->rec, ctx->masks,
                                                             a.noise == S17_NOISE_PROB, 0,
                                                             a.noise == S17_NOISE_SUBSET && a.protocol == S17_PROTO_RTF, ctx->nd,
                                                             ctx->cum, a.seed, a.first_shot_id);
    }
    return S17_OK;
}

static int end_call(s17_ctx *ctx, const s17_run_args &a, s17_counts *out) {
    unsigned long long c[C_N];
    CK(cudaEventRecord(ctx->span1, ctx->stream));
    CK(cudaMemcpyAsync(c, ctx->counters, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->span0, ctx->span1));
        ctx->t_span = ms;
    }
    memset(out, 0, sizeof(*out));
    out->shots = a.n_shots;
    out->failed = c[C_FAILED];
    out->not0 = c[C_NOT0];
    out->counted = c[C_COUNTED];
    out->errors = c[C_ERRORS];
    out->sweeps = c[C_SWEEPS];
    out->rounds = c[C_ROUNDS];
    out->sweep_bytes = c[C_RUNS_MOVED] * (unsigned long long)kRunBytes;   // algorithmic HBM bytes moved by the gate sweeps
    ctx->u_sub[S17_T_COLLAPSE] = c[C_COLLAPSED];                         // shots whose collapse was materialised
    ctx->last_n = a.n_shots;
    return S17_OK;
}

// ---- run-til-fail (CL:33-148) as a job: begin, step (a number of rounds for every live slot), end ---------------------
// Every slot carries one run; rounds are fresh restarts (CL:70-73) that share only the leak register, which the
// reference never clears (CL:49; S17_LEAK_RESET_NEVER, per slot here), or which is cleared per run / per round.
extern "C" int s17_rtf_begin(s17_ctx *ctx, const s17_run_args *args) {
    if (!ctx || !args) return fail(ctx, S17_E_ARG, "s17_rtf_begin: bad arguments");
    s17_run_args a = *args;
    a.protocol = S17_PROTO_RTF;
    if (ctx->rtf_open) return fail(ctx, S17_E_STATE, "s17_rtf_begin: a run-til-fail job is already open");
    if (int rc = check_run_args(ctx, a)) return rc;
    if (a.rtf_runs == 0) return fail(ctx, S17_E_ARG, "s17_run: rtf_runs must be > 0");
    if (a.rtf_chain_runs && (uint64_t)a.rtf_chain_runs * a.n_shots != a.rtf_runs)
        return fail(ctx, S17_E_ARG, "s17_run: with rtf_chain_runs = K, rtf_runs must be n_shots * K");
    if (a.noise == S17_NOISE_REPLAY || a.forced) return fail(ctx, S17_E_ARG, "s17_rtf_begin: run-til-fail samples its outcomes");
    if (int rc = begin_call(ctx, a)) return rc;
    const int n = (int)a.n_shots;
    if (ctx->rtf_capacity < a.rtf_runs) {
        if (ctx->rtf_rounds) cudaFree(ctx->rtf_rounds);
        if (ctx->rtf_syndb) cudaFree(ctx->rtf_syndb);
        ctx->rtf_rounds = ctx->rtf_syndb = nullptr;
        ctx->rtf_capacity = 0;
        CK(cudaMalloc(&ctx->rtf_rounds, (size_t)a.rtf_runs * 4));
        CK(cudaMalloc(&ctx->rtf_syndb, (size_t)a.rtf_runs * 4));
        ctx->rtf_capacity = a.rtf_runs;
    }
    CK(cudaMemsetAsync(ctx->rtf_rounds, 0, (size_t)a.rtf_runs * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->rtf_syndb, 0, (size_t)a.rtf_runs * 4, ctx->stream));
    k_rtf_init<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(n, a.rtf_runs, a.rtf_chain_runs, ctx->rtf_slots, ctx->leak, ctx->counters);
    ctx->rtf_args = a;
    ctx->rtf_round_no = 0;
    ctx->rtf_rounds_total = 0;
    ctx->rtf_alive = std::min<uint64_t>(a.n_shots, a.rtf_runs);
    ctx->rtf_open = true;
    return S17_OK;
}

extern "C" int s17_rtf_step(s17_ctx *ctx, uint32_t n_rounds, uint64_t *alive, uint64_t *rounds_total) {
    if (!ctx || !ctx->rtf_open) return fail(ctx, S17_E_STATE, "s17_rtf_step: no run-til-fail job is open");
    CK(cudaSetDevice(ctx->device));
    const s17_run_args &a = ctx->rtf_args;
    const int n = (int)a.n_shots;
    for (uint32_t k = 0; k < n_rounds; ++k) {
        k_rtf_begin_round<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(n, ctx->rtf_slots, ctx->active, ctx->ready, ctx->rec,
                                                                       ctx->leak, a.leak_reset == S17_LEAK_RESET_ROUND);
        s17_run_args ar = a;
        ar.seed = a.seed + 0x9E3779B97F4A7C15ull * (ctx->rtf_round_no + 1);   // fresh Philox key per round
        if (int rc = run_round(ctx, ar)) return rc;
        CK(cudaMemsetAsync(ctx->counters + C_ALIVE, 0, sizeof(unsigned long long), ctx->stream));
        const int nb = blocks_for(n, kRtfThreads);
        k_rtf_finish<<<nb, kRtfThreads, 0, ctx->stream>>>(n, a.rtf_max_rounds, ctx->rtf_slots, ctx->rec, ctx->rtf_rounds,
                                                         ctx->rtf_syndb, ctx->rtf_block_cnt, ctx->rtf_block_off, ctx->counters,
                                                         ctx->rtf_ticket);
        k_rtf_assign<<<nb, kRtfThreads, 0, ctx->stream>>>(n, a.rtf_runs, a.rtf_chain_runs, ctx->rtf_slots, ctx->leak, ctx->rtf_block_off,
                                                         ctx->counters, a.leak_reset == S17_LEAK_RESET_RUN);
        ++ctx->rtf_round_no;
        // the timing pool is bounded: resolve it every few rounds (a synchronisation, but no host read of the result)
        if (ctx->launches.size() > 2048)
            if (int rc = finish_timing(ctx)) return rc;
    }
    if (int rc = finish_timing(ctx)) return rc;
    unsigned long long c[2] = {0, 0};
    CK(cudaMemcpy(&c[0], ctx->counters + C_ALIVE, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&c[1], ctx->counters + C_ROUNDS, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (n_rounds) ctx->rtf_alive = c[0];
    ctx->rtf_rounds_total = c[1];
    if (alive) *alive = ctx->rtf_alive;
    if (rounds_total) *rounds_total = ctx->rtf_rounds_total;
    return S17_OK;
}

extern "C" int s17_rtf_end(s17_ctx *ctx, s17_counts *out) {
    if (!ctx || !out || !ctx->rtf_open) return fail(ctx, S17_E_STATE, "s17_rtf_end: no run-til-fail job is open");
    CK(cudaSetDevice(ctx->device));
    ctx->rtf_open = false;
    if (int rc = end_call(ctx, ctx->rtf_args, out)) return rc;
    // failed_count of a run-til-fail job = runs that ended with a logical failure; the per-round tallies of the
    // read-out kernels count the same events
    return S17_OK;
}

extern "C" int s17_read_rtf_slots(s17_ctx *ctx, uint32_t *out /* n x {run_id, rounds, synd_b, alive} */, uint64_t n) {
    if (!ctx || !out || n > ctx->max_shots) return fail(ctx, S17_E_ARG, "s17_read_rtf_slots: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(out, ctx->rtf_slots, n * sizeof(RtfSlot), cudaMemcpyDeviceToHost));
    return S17_OK;
}

// Sum the counters of several contexts of ONE process (one context per GPU per host thread): the reduction the
// reference's loops do with `failed_count += 1` across the whole run.  Across processes (one per GPU, torchrun) the same
// sum is one all_reduce of these integers (surface_17_iqt_b200/distributed.py: NCCL over NVLink).
extern "C" int s17_allreduce_counts(const s17_counts *per_ctx, int n_ctx, s17_counts *out) {
    if (!per_ctx || !out || n_ctx <= 0) return fail(nullptr, S17_E_ARG, "s17_allreduce_counts: bad arguments");
    s17_counts t;
    memset(&t, 0, sizeof(t));
    for (int i = 0; i < n_ctx; ++i) {
        t.shots += per_ctx[i].shots; t.failed += per_ctx[i].failed; t.not0 += per_ctx[i].not0;
        t.counted += per_ctx[i].counted; t.errors += per_ctx[i].errors; t.sweeps += per_ctx[i].sweeps;
        t.rounds += per_ctx[i].rounds; t.sweep_bytes += per_ctx[i].sweep_bytes;
    }
    *out = t;
    return S17_OK;
}

extern "C" int s17_run(s17_ctx *ctx, const s17_run_args *args, s17_counts *out) {
    if (!ctx || !args || !out) return fail(ctx, S17_E_ARG, "s17_run: bad arguments");
    Nvtx r_run("s17_run");
    s17_run_args a = *args;
    if (a.protocol == S17_PROTO_RTF) {
        // poll the live-slot count every few rounds: the tail of a job (few live slots) is launch-bound, and a host
        // round trip per round would double its cost
        if (int rc = s17_rtf_begin(ctx, &a)) return rc;
        const uint32_t poll = a.rtf_poll ? a.rtf_poll : 4u;
        for (;;) {
            uint64_t alive = 0, rounds = 0;
            if (int rc = s17_rtf_step(ctx, poll, &alive, &rounds)) { ctx->rtf_open = false; return rc; }
            if (alive == 0) break;
            if (a.rtf_round_budget && rounds >= a.rtf_round_budget) break;      // the rest of the runs stay censored
        }
        return s17_rtf_end(ctx, out);
    }
    if (int rc = check_run_args(ctx, a)) return rc;
    if (int rc = begin_call(ctx, a)) return rc;
    if (int rc = run_round(ctx, a)) return rc;
    if (int rc = finish_timing(ctx)) return rc;
    return end_call(ctx, a, out);
}

extern "C" int s17_read_records(s17_ctx *ctx, s17_record *out, uint64_t n) {
    if (!ctx || !out || n > ctx->max_shots) return fail(ctx, S17_E_ARG, "s17_read_records: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(out, ctx->rec, n * sizeof(s17_record), cudaMemcpyDeviceToHost));
    return S17_OK;
}

extern "C" int s17_read_events(s17_ctx *ctx, uint32_t *out, uint64_t n) {
    if (!ctx || !out || n > ctx->max_shots) return fail(ctx, S17_E_ARG, "s17_read_events: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(out, ctx->ev, n * S17_EV_WORDS * 4, cudaMemcpyDeviceToHost));
    return S17_OK;
}

extern "C" int s17_read_masks(s17_ctx *ctx, uint32_t *out, uint64_t n) {
    if (!ctx || !out || n > ctx->max_shots) return fail(ctx, S17_E_ARG, "s17_read_masks: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(out, ctx->masks, n * 2 * S17_MASK_WORDS * 4, cudaMemcpyDeviceToHost));
    return S17_OK;
}

extern "C" int s17_read_state(s17_ctx *ctx, uint64_t shot, double *re_im) {
    if (!ctx || !re_im || shot >= ctx->max_shots) return fail(ctx, S17_E_ARG, "s17_read_state: bad arguments");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->state) return fail(ctx, S17_E_STATE, "s17_read_state: nothing has run yet");
    if (ctx->use_quiet && ctx->quiet_lazy) {
        // a shot whose first noisy cycle was event-free and that stopped there: its block is the memoised one of its leaf
        uint8_t q = 0, leaf = 0, ok = 0;
        CK(cudaMemcpy(&q, ctx->quiet_flag + shot, 1, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&leaf, ctx->prep_leaf + shot, 1, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&ok, ctx->quiet_ok + leaf, 1, cudaMemcpyDeviceToHost));
        if (q && ok) {
            memset(re_im, 0, kDim * sizeof(double2));
            CK(cudaMemcpy(re_im, ctx->quiet_blocks + (size_t)leaf * kRun, kRun * sizeof(double2), cudaMemcpyDeviceToHost));
            return S17_OK;
        }
    }
    if (ctx->state_full) {
        CK(cudaMemcpy(re_im, ctx->state + shot * kDim, kDim * sizeof(double2), cudaMemcpyDeviceToHost));
    } else {
        // compact context: every ancilla is measured and reset, the state is its 512-amplitude data block
        memset(re_im, 0, kDim * sizeof(double2));
        CK(cudaMemcpy(re_im, ctx->state + shot * kRun, kRun * sizeof(double2), cudaMemcpyDeviceToHost));
    }
    return S17_OK;
}

extern "C" int s17_read_rtf(s17_ctx *ctx, uint32_t *rounds, uint32_t *syndb, uint64_t n_runs) {
    if (!ctx || !rounds || n_runs > ctx->rtf_capacity) return fail(ctx, S17_E_ARG, "s17_read_rtf: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(rounds, ctx->rtf_rounds, n_runs * 4, cudaMemcpyDeviceToHost));
    if (syndb) CK(cudaMemcpy(syndb, ctx->rtf_syndb, n_runs * 4, cudaMemcpyDeviceToHost));
    return S17_OK;
}

extern "C" int s17_last_timing_detail(s17_ctx *ctx, double *ms, uint64_t *launches, uint64_t *units) {
    if (!ctx || !ms || !launches) return fail(ctx, S17_E_ARG, "s17_last_timing_detail: bad arguments");
    for (int i = 0; i < S17_T_N; ++i) {
        ms[i] = ctx->t_sub[i];
        launches[i] = ctx->n_sub[i];
        if (units) units[i] = ctx->u_sub[i];
    }
    return S17_OK;
}

extern "C" int s17_last_timing(s17_ctx *ctx, double *gate_ms, double *measure_ms, double *other_ms, uint64_t *gate_launches,
                               uint64_t *all_launches, double *span_ms) {
    if (!ctx) return fail(ctx, S17_E_ARG, "s17_last_timing: bad arguments");
    if (span_ms) *span_ms = ctx->t_span;
    if (gate_ms) *gate_ms = ctx->t_gate;
    if (measure_ms) *measure_ms = ctx->t_meas;
    if (other_ms) *other_ms = ctx->t_other;
    if (gate_launches) *gate_launches = ctx->n_gate;
    if (all_launches) *all_launches = ctx->n_all;
    return S17_OK;
}
