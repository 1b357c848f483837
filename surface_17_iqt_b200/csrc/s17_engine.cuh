// s17_engine.cuh -- the engine-level face of libs17.so: ProjectQ's Simulator backend, batched.
//
// What the reference asks its backend to do, one call per gate per shot:
//   `Gate | qubit(s)`  (CS:291-344 and every insert_error Pauli, CS:198-217)  -> one sweep over the 2^n amplitudes
//   `Measure | qubit`  (CS:56, 848)   -> P(1) norm reduction, one uniform, zero the other half, rescale by 1/sqrt(P)
// Here the same two primitives run for a BATCH of state vectors per call: `state[shot][2^n]` complex128 in HBM, every
// gate one in-place sweep (algorithmic traffic 2^n x 16 B read + written per shot = 4 194 304 B at n = 17, the "gate
// layer" of SURVEY 8d), every measurement a reduction pass (2 097 152 B read) plus a collapse pass (the surviving half
// rescaled in place, the other half zeroed).  Both are HBM-bound: a thread streams amplitude pairs / quadruples with
// 16-byte accesses that are contiguous across the warp, several independent pairs in flight per thread; the grid is a
// multiple of the SM count and walks the batch with a grid-stride loop.
//
// A per-shot predicate lets the host apply a gate to a subset of the batch (conditional reset `if int(a) == 1: X | a`,
// CS:859-862; per-shot error insertion).  Included by s17.cu.
#pragma once

namespace s17 {

struct Mat2 { double2 m[2][2]; };
struct Mat4 { double2 m[4][4]; };

constexpr int kEngThreads = 256;
constexpr int kEngUnroll = 4;             // independent pairs / quadruples per thread and loop trip
constexpr int kEngChunks = 16;            // partial sums per shot in the probability reduction

__device__ __forceinline__ double2 eng_cfma(const double2 a, const double2 b, const double2 c) {   // a * b + c
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
__device__ __forceinline__ double2 eng_cmul(const double2 a, const double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}

// one-qubit gate on state-vector bit `bit` of every (selected) shot
__global__ void __launch_bounds__(kEngThreads)
k_eng_gate1(double2 *__restrict__ state, uint32_t n_bits, uint64_t n_shots, uint32_t bit, const __grid_constant__ Mat2 M,
            const uint8_t *__restrict__ pred)
{
    const uint64_t pairs_per_shot = 1ull << (n_bits - 1);
    const uint64_t total = n_shots * pairs_per_shot;
    const uint64_t stride = (uint64_t)gridDim.x * kEngThreads;
    const uint64_t step = 1ull << bit;
    for (uint64_t k0 = (uint64_t)blockIdx.x * kEngThreads + threadIdx.x; k0 < total; k0 += stride * kEngUnroll) {
        double2 a[kEngUnroll], b[kEngUnroll];
        uint64_t idx[kEngUnroll];
        bool on[kEngUnroll];
#pragma unroll
        for (int u = 0; u < kEngUnroll; ++u) {
            const uint64_t k = k0 + (uint64_t)u * stride;
            on[u] = k < total;
            if (on[u]) {
                const uint64_t shot = k >> (n_bits - 1), p = k & (pairs_per_shot - 1);
                if (pred && !pred[shot]) { on[u] = false; continue; }
                idx[u] = (shot << n_bits) | ((p >> bit) << (bit + 1)) | (p & (step - 1));
                a[u] = state[idx[u]];
                b[u] = state[idx[u] + step];
            }
        }
#pragma unroll
        for (int u = 0; u < kEngUnroll; ++u) {
            if (!on[u]) continue;
            state[idx[u]] = eng_cfma(M.m[0][1], b[u], eng_cmul(M.m[0][0], a[u]));
            state[idx[u] + step] = eng_cfma(M.m[1][1], b[u], eng_cmul(M.m[1][0], a[u]));
        }
    }
}

// two-qubit gate; matrix index = (value of bit0) + 2 * (value of bit1), ProjectQ's ordering for `Gate | (q0, q1)`
__global__ void __launch_bounds__(kEngThreads)
k_eng_gate2(double2 *__restrict__ state, uint32_t n_bits, uint64_t n_shots, uint32_t bit0, uint32_t bit1,
            const __grid_constant__ Mat4 M, const uint8_t *__restrict__ pred)
{
    const uint64_t quads_per_shot = 1ull << (n_bits - 2);
    const uint64_t total = n_shots * quads_per_shot;
    const uint64_t stride = (uint64_t)gridDim.x * kEngThreads;
    const uint32_t lo = min(bit0, bit1), hi = max(bit0, bit1);
    const uint64_t s0 = 1ull << bit0, s1 = 1ull << bit1;
    for (uint64_t k0 = (uint64_t)blockIdx.x * kEngThreads + threadIdx.x; k0 < total; k0 += stride * 2) {
        double2 a[2][4];
        uint64_t idx[2];
        bool on[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint64_t k = k0 + (uint64_t)u * stride;
            on[u] = k < total;
            if (on[u]) {
                const uint64_t shot = k >> (n_bits - 2);
                uint64_t p = k & (quads_per_shot - 1);
                if (pred && !pred[shot]) { on[u] = false; continue; }
                p = ((p >> lo) << (lo + 1)) | (p & ((1ull << lo) - 1));
                p = ((p >> hi) << (hi + 1)) | (p & ((1ull << hi) - 1));
                idx[u] = (shot << n_bits) | p;
                a[u][0] = state[idx[u]];
                a[u][1] = state[idx[u] + s0];
                a[u][2] = state[idx[u] + s1];
                a[u][3] = state[idx[u] + s0 + s1];
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!on[u]) continue;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                double2 acc = eng_cmul(M.m[r][0], a[u][0]);
                acc = eng_cfma(M.m[r][1], a[u][1], acc);
                acc = eng_cfma(M.m[r][2], a[u][2], acc);
                acc = eng_cfma(M.m[r][3], a[u][3], acc);
                state[idx[u] + ((r & 1) ? s0 : 0) + ((r & 2) ? s1 : 0)] = acc;
            }
        }
    }
}

// Measure, step 1: per shot and chunk the squared norm of the amplitudes with `bit` clear / set.
// grid = n_shots * kEngChunks; partial[shot][chunk] = (P0, P1) in a fixed summation order (reproducible).
__global__ void __launch_bounds__(kEngThreads)
k_eng_prob(const double2 *__restrict__ state, uint32_t n_bits, uint32_t bit, double2 *__restrict__ partial)
{
    __shared__ double2 s_part[kEngThreads / 32];
    const uint64_t shot = blockIdx.x / kEngChunks;
    const uint32_t chunk = blockIdx.x % kEngChunks;
    const uint32_t per_chunk = (1u << n_bits) / kEngChunks;
    const double2 *s = state + (shot << n_bits) + (uint64_t)chunk * per_chunk;
    double p0 = 0.0, p1 = 0.0;
#pragma unroll 8
    for (uint32_t i = threadIdx.x; i < per_chunk; i += kEngThreads) {
        const double2 v = s[i];
        const double w = fma(v.x, v.x, v.y * v.y);
        if (((chunk * per_chunk + i) >> bit) & 1u) p1 += w; else p0 += w;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        p0 += __shfl_xor_sync(0xffffffffu, p0, d);
        p1 += __shfl_xor_sync(0xffffffffu, p1, d);
    }
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = make_double2(p0, p1);
    __syncthreads();
    if (threadIdx.x == 0) {
        double2 t = make_double2(0.0, 0.0);
        for (int w = 0; w < kEngThreads / 32; ++w) { t.x += s_part[w].x; t.y += s_part[w].y; }
        partial[shot * kEngChunks + chunk] = t;
    }
}

// Measure, step 2: one thread per shot -- outcome from P(1) against one uniform (statistically ProjectQ's cumulative
// walk), or the forced / lock-step outcome; 1/sqrt(P(outcome)) for the collapse.
// mode 0: independent Philox draw per shot; 1: lock-step (every shot takes the outcome drawn for shot 0: the batch is
// N copies of one trajectory); 2: outcomes given per shot in `forced`.
__global__ void __launch_bounds__(256)
k_eng_sample(uint64_t n_shots, const double2 *__restrict__ partial, int mode, const uint8_t *__restrict__ forced,
             uint64_t seed, uint64_t first_shot, uint32_t draw_index, uint8_t *__restrict__ out, double *__restrict__ scale,
             double *__restrict__ prob, uint32_t *__restrict__ n_bad)
{
    const uint64_t shot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (shot >= n_shots) return;
    auto pair_of = [&](uint64_t s_) {
        double2 t = make_double2(0.0, 0.0);
        for (int c = 0; c < kEngChunks; ++c) { const double2 v = partial[s_ * kEngChunks + c]; t.x += v.x; t.y += v.y; }
        return t;
    };
    const double2 mine = pair_of(shot);
    uint32_t o;
    if (mode == 2) {
        o = forced[shot] & 1u;
    } else {
        const uint64_t src = mode == 1 ? 0 : shot;
        const double2 pp = mode == 1 ? pair_of(0) : mine;
        Philox rng(seed, first_shot + src, 64u);
        rng.c3 = draw_index;
        o = (rng.next() * (pp.x + pp.y) < pp.y) ? 1u : 0u;
    }
    const double p_sel = o ? mine.y : mine.x;
    if (p_sel <= 0.0) atomicAdd(n_bad, 1u);            // a forced / lock-step outcome this shot cannot produce
    out[shot] = (uint8_t)o;
    scale[shot] = p_sel > 0.0 ? 1.0 / sqrt(p_sel) : 0.0;
    prob[shot] = (mine.x + mine.y) > 0.0 ? p_sel / (mine.x + mine.y) : 0.0;
}

// Measure, step 3: collapse + renormalise (the half that disagrees with the outcome becomes zero)
__global__ void __launch_bounds__(kEngThreads)
k_eng_collapse(double2 *__restrict__ state, uint32_t n_bits, uint64_t n_shots, uint32_t bit, const uint8_t *__restrict__ out,
               const double *__restrict__ scale)
{
    const uint64_t pairs_per_shot = 1ull << (n_bits - 1);
    const uint64_t total = n_shots * pairs_per_shot;
    const uint64_t stride = (uint64_t)gridDim.x * kEngThreads;
    const uint64_t step = 1ull << bit;
    for (uint64_t k0 = (uint64_t)blockIdx.x * kEngThreads + threadIdx.x; k0 < total; k0 += stride * kEngUnroll) {
        double2 keep[kEngUnroll];
        uint64_t ik[kEngUnroll], iz[kEngUnroll];
        double sc[kEngUnroll];
        bool on[kEngUnroll];
#pragma unroll
        for (int u = 0; u < kEngUnroll; ++u) {
            const uint64_t k = k0 + (uint64_t)u * stride;
            on[u] = k < total;
            if (on[u]) {
                const uint64_t shot = k >> (n_bits - 1), p = k & (pairs_per_shot - 1);
                const uint64_t i0 = (shot << n_bits) | ((p >> bit) << (bit + 1)) | (p & (step - 1));
                const uint32_t o = out[shot];
                ik[u] = i0 + (o ? step : 0);
                iz[u] = i0 + (o ? 0 : step);
                sc[u] = scale[shot];
                keep[u] = state[ik[u]];
            }
        }
#pragma unroll
        for (int u = 0; u < kEngUnroll; ++u) {
            if (!on[u]) continue;
            state[ik[u]] = make_double2(keep[u].x * sc[u], keep[u].y * sc[u]);
            state[iz[u]] = make_double2(0.0, 0.0);
        }
    }
}

// every shot back to |0...0>
__global__ void __launch_bounds__(kEngThreads)
k_eng_reset(double2 *__restrict__ state, uint32_t n_bits, uint64_t n_shots)
{
    const uint64_t total = n_shots << n_bits;
    const uint64_t stride = (uint64_t)gridDim.x * kEngThreads;
    for (uint64_t k = (uint64_t)blockIdx.x * kEngThreads + threadIdx.x; k < total; k += stride)
        state[k] = make_double2((k & ((1ull << n_bits) - 1)) == 0 ? 1.0 : 0.0, 0.0);
}

}  // namespace s17

// =================================================================================================
// host side: engine context + C ABI (include/s17.h, "engine-level face")
// =================================================================================================
using namespace s17;

struct EngTimed { int cat; cudaEvent_t e0, e1; };

struct s17_eng {
    int device = 0;
    uint32_t n_bits = 0;
    uint64_t batch = 0, seed = 0;
    uint32_t draws = 0;                   // Measure commands so far: Philox counter block of the next draw
    cudaStream_t stream = nullptr;
    double2 *state = nullptr, *partial = nullptr;
    uint8_t *out = nullptr, *pred = nullptr, *forced = nullptr;
    double *scale = nullptr, *prob = nullptr;
    uint32_t *n_bad = nullptr;
    int num_sms = 148;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    std::vector<EngTimed> pending;
    double t_ms[3] = {0, 0, 0};           // gate sweeps | probability reductions | collapses
    uint64_t n_calls[3] = {0, 0, 0};
    std::string err;
};

static thread_local std::string g_eng_err;
static int eng_fail(s17_eng *e, int code, const std::string &msg) {
    if (e) e->err = msg;
    g_eng_err = msg;
    return code;
}
#define ECK(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return eng_fail(eng, S17_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
    } while (0)

static int eng_resolve(s17_eng *eng) {
    ECK(cudaStreamSynchronize(eng->stream));
    ECK(cudaGetLastError());
    for (const EngTimed &t : eng->pending) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t.e0, t.e1);
        eng->t_ms[t.cat] += ms;
        eng->n_calls[t.cat] += 1;
    }
    eng->pending.clear();
    eng->ev_used = 0;
    return S17_OK;
}
struct EngScope {
    s17_eng *e; EngTimed t;
    EngScope(s17_eng *e_, int cat) : e(e_) {
        while (e->ev_pool.size() < e->ev_used + 2) { cudaEvent_t ev; cudaEventCreate(&ev); e->ev_pool.push_back(ev); }
        t.cat = cat; t.e0 = e->ev_pool[e->ev_used++]; t.e1 = e->ev_pool[e->ev_used++];
        cudaEventRecord(t.e0, e->stream);
    }
    ~EngScope() { cudaEventRecord(t.e1, e->stream); e->pending.push_back(t); }
};

extern "C" const char *s17_eng_last_error(const s17_eng *eng) { return eng ? eng->err.c_str() : g_eng_err.c_str(); }

extern "C" int s17_eng_create(s17_eng **out, int device, uint32_t n_bits, uint64_t batch, uint64_t seed) {
    s17_eng *eng = nullptr;
    if (!out || batch == 0 || n_bits < 2 || n_bits > 26) return eng_fail(nullptr, S17_E_ARG, "s17_eng_create: bad arguments");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
        return eng_fail(nullptr, S17_E_CUDA, "s17_eng_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= n) return eng_fail(nullptr, S17_E_ARG, "s17_eng_create: device index out of range");
    eng = new s17_eng();
    eng->device = device; eng->n_bits = n_bits; eng->batch = batch; eng->seed = seed;
    ECK(cudaSetDevice(device));
    ECK(cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking));
    if (cudaMalloc(&eng->state, (batch << n_bits) * sizeof(double2)) != cudaSuccess) {
        cudaGetLastError();
        const int rc = eng_fail(nullptr, S17_E_NOMEM, "s17_eng_create: out of device memory for " + std::to_string(batch) +
                                                          " state vectors of 2^" + std::to_string(n_bits) + " amplitudes");
        cudaStreamDestroy(eng->stream);
        delete eng;
        return rc;
    }
    ECK(cudaMalloc(&eng->partial, batch * kEngChunks * sizeof(double2)));
    ECK(cudaMalloc(&eng->out, batch));
    ECK(cudaMalloc(&eng->pred, batch));
    ECK(cudaMalloc(&eng->forced, batch));
    ECK(cudaMalloc(&eng->scale, batch * sizeof(double)));
    ECK(cudaMalloc(&eng->prob, batch * sizeof(double)));
    ECK(cudaMalloc(&eng->n_bad, 4));
    cudaDeviceProp prop;
    ECK(cudaGetDeviceProperties(&prop, device));
    eng->num_sms = prop.multiProcessorCount;
    *out = eng;
    return s17_eng_reset(eng);
}

extern "C" int s17_eng_destroy(s17_eng *eng) {
    if (!eng) return S17_OK;
    cudaSetDevice(eng->device);
    const bool ok = cudaStreamSynchronize(eng->stream) == cudaSuccess;
    if (ok) {
        for (void *p : {(void *)eng->state, (void *)eng->partial, (void *)eng->out, (void *)eng->pred, (void *)eng->forced,
                        (void *)eng->scale, (void *)eng->prob, (void *)eng->n_bad})
            if (p) cudaFree(p);
        for (auto ev : eng->ev_pool) cudaEventDestroy(ev);
        cudaStreamDestroy(eng->stream);
    }
    delete eng;
    return ok ? S17_OK : S17_E_CUDA;
}

static int eng_grid(const s17_eng *eng, uint64_t items, int per_thread) {
    const uint64_t want = (items + (uint64_t)kEngThreads * per_thread - 1) / ((uint64_t)kEngThreads * per_thread);
    const uint64_t cap = (uint64_t)eng->num_sms * 16;          // a multiple of the SM count; grid-stride beyond it
    return (int)std::max<uint64_t>(1, std::min(want, cap));
}

extern "C" int s17_eng_reset(s17_eng *eng) {
    if (!eng) return eng_fail(eng, S17_E_ARG, "s17_eng_reset: bad arguments");
    ECK(cudaSetDevice(eng->device));
    k_eng_reset<<<eng_grid(eng, eng->batch << eng->n_bits, 4), kEngThreads, 0, eng->stream>>>(eng->state, eng->n_bits, eng->batch);
    eng->draws = 0;
    ECK(cudaGetLastError());
    return S17_OK;
}

static int eng_pred(s17_eng *eng, const uint8_t *pred, const uint8_t **dev) {
    *dev = nullptr;
    if (!pred) return S17_OK;
    ECK(cudaMemcpyAsync(eng->pred, pred, eng->batch, cudaMemcpyHostToDevice, eng->stream));
    ECK(cudaStreamSynchronize(eng->stream));          // the caller's buffer may be reused right after the call
    *dev = eng->pred;
    return S17_OK;
}

extern "C" int s17_eng_gate1(s17_eng *eng, uint32_t bit, const double *m, const uint8_t *pred) {
    if (!eng || !m || bit >= eng->n_bits) return eng_fail(eng, S17_E_ARG, "s17_eng_gate1: bad arguments");
    ECK(cudaSetDevice(eng->device));
    Mat2 M;
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 2; ++c) M.m[r][c] = make_double2(m[2 * (2 * r + c)], m[2 * (2 * r + c) + 1]);
    const uint8_t *dp;
    if (int rc = eng_pred(eng, pred, &dp)) return rc;
    if (eng->pending.size() > 4096) if (int rc = eng_resolve(eng)) return rc;
    {
        EngScope s(eng, 0);
        k_eng_gate1<<<eng_grid(eng, eng->batch << (eng->n_bits - 1), kEngUnroll), kEngThreads, 0, eng->stream>>>(
            eng->state, eng->n_bits, eng->batch, bit, M, dp);
    }
    ECK(cudaGetLastError());
    return S17_OK;
}

extern "C" int s17_eng_gate2(s17_eng *eng, uint32_t bit0, uint32_t bit1, const double *m, const uint8_t *pred) {
    if (!eng || !m || bit0 >= eng->n_bits || bit1 >= eng->n_bits || bit0 == bit1)
        return eng_fail(eng, S17_E_ARG, "s17_eng_gate2: bad arguments");
    ECK(cudaSetDevice(eng->device));
    Mat4 M;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) M.m[r][c] = make_double2(m[2 * (4 * r + c)], m[2 * (4 * r + c) + 1]);
    const uint8_t *dp;
    if (int rc = eng_pred(eng, pred, &dp)) return rc;
    if (eng->pending.size() > 4096) if (int rc = eng_resolve(eng)) return rc;
    {
        EngScope s(eng, 0);
        k_eng_gate2<<<eng_grid(eng, eng->batch << (eng->n_bits - 2), 2), kEngThreads, 0, eng->stream>>>(
            eng->state, eng->n_bits, eng->batch, bit0, bit1, M, dp);
    }
    ECK(cudaGetLastError());
    return S17_OK;
}

// P(bit = 1) per shot, normalised (host array of `batch` doubles): what ProjectQ checks before it lets a qubit go
extern "C" int s17_eng_prob1(s17_eng *eng, uint32_t bit, double *p1) {
    if (!eng || !p1 || bit >= eng->n_bits) return eng_fail(eng, S17_E_ARG, "s17_eng_prob1: bad arguments");
    ECK(cudaSetDevice(eng->device));
    {
        EngScope s(eng, 1);
        k_eng_prob<<<(unsigned)(eng->batch * kEngChunks), kEngThreads, 0, eng->stream>>>(eng->state, eng->n_bits, bit, eng->partial);
    }
    std::vector<double2> h(eng->batch * kEngChunks);
    ECK(cudaMemcpyAsync(h.data(), eng->partial, h.size() * sizeof(double2), cudaMemcpyDeviceToHost, eng->stream));
    if (int rc = eng_resolve(eng)) return rc;
    for (uint64_t s = 0; s < eng->batch; ++s) {
        double a = 0.0, b = 0.0;
        for (int c = 0; c < kEngChunks; ++c) { a += h[s * kEngChunks + c].x; b += h[s * kEngChunks + c].y; }
        p1[s] = (a + b) > 0.0 ? b / (a + b) : 0.0;
    }
    return S17_OK;
}

extern "C" int s17_eng_measure(s17_eng *eng, uint32_t bit, uint32_t mode, const uint8_t *forced, uint8_t *out, double *prob,
                               uint32_t *n_impossible) {
    if (!eng || !out || bit >= eng->n_bits || mode > 2 || (mode == 2 && !forced))
        return eng_fail(eng, S17_E_ARG, "s17_eng_measure: bad arguments");
    ECK(cudaSetDevice(eng->device));
    if (mode == 2) ECK(cudaMemcpyAsync(eng->forced, forced, eng->batch, cudaMemcpyHostToDevice, eng->stream));
    ECK(cudaMemsetAsync(eng->n_bad, 0, 4, eng->stream));
    {
        EngScope s(eng, 1);
        k_eng_prob<<<(unsigned)(eng->batch * kEngChunks), kEngThreads, 0, eng->stream>>>(eng->state, eng->n_bits, bit, eng->partial);
    }
    k_eng_sample<<<(unsigned)((eng->batch + 255) / 256), 256, 0, eng->stream>>>(eng->batch, eng->partial, (int)mode, eng->forced,
                                                                             eng->seed, 0, eng->draws, eng->out, eng->scale,
                                                                             eng->prob, eng->n_bad);
    eng->draws += 1;
    {
        EngScope s(eng, 2);
        k_eng_collapse<<<eng_grid(eng, eng->batch << (eng->n_bits - 1), kEngUnroll), kEngThreads, 0, eng->stream>>>(
            eng->state, eng->n_bits, eng->batch, bit, eng->out, eng->scale);
    }
    uint32_t bad = 0;
    ECK(cudaMemcpyAsync(out, eng->out, eng->batch, cudaMemcpyDeviceToHost, eng->stream));
    if (prob) ECK(cudaMemcpyAsync(prob, eng->prob, eng->batch * sizeof(double), cudaMemcpyDeviceToHost, eng->stream));
    ECK(cudaMemcpyAsync(&bad, eng->n_bad, 4, cudaMemcpyDeviceToHost, eng->stream));
    if (int rc = eng_resolve(eng)) return rc;
    if (n_impossible) *n_impossible = bad;
    return S17_OK;
}

extern "C" int s17_eng_flush(s17_eng *eng) {
    if (!eng) return eng_fail(eng, S17_E_ARG, "s17_eng_flush: bad arguments");
    ECK(cudaSetDevice(eng->device));
    return eng_resolve(eng);
}

extern "C" int s17_eng_read_state(s17_eng *eng, uint64_t shot, double *re_im) {
    if (!eng || !re_im || shot >= eng->batch) return eng_fail(eng, S17_E_ARG, "s17_eng_read_state: bad arguments");
    ECK(cudaSetDevice(eng->device));
    if (int rc = eng_resolve(eng)) return rc;
    ECK(cudaMemcpy(re_im, eng->state + (shot << eng->n_bits), (sizeof(double2)) << eng->n_bits, cudaMemcpyDeviceToHost));
    return S17_OK;
}

// accumulated device time (CUDA events on the launch stream) and call counts per kernel class since the last call with
// clear != 0: [0] gate sweeps, [1] probability reductions, [2] collapses
extern "C" int s17_eng_timing(s17_eng *eng, double ms[3], uint64_t calls[3], int clear) {
    if (!eng || !ms || !calls) return eng_fail(eng, S17_E_ARG, "s17_eng_timing: bad arguments");
    ECK(cudaSetDevice(eng->device));
    if (int rc = eng_resolve(eng)) return rc;
    for (int i = 0; i < 3; ++i) { ms[i] = eng->t_ms[i]; calls[i] = eng->n_calls[i]; }
    if (clear) for (int i = 0; i < 3; ++i) { eng->t_ms[i] = 0; eng->n_calls[i] = 0; }
    return S17_OK;
}
