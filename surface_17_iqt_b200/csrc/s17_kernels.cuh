// s17_kernels.cuh -- device code of libs17.so (sm_100a only).
//
// One shot = one 2^17-amplitude complex128 state vector resident in HBM (2 MiB).  The gate
// kernel sweeps the whole batch once per *layer*: a CTA stages a 4096-amplitude tile (all 9 data
// bits + 3 ancilla bits = 8 contiguous 8 KiB runs) into shared memory with 1-D TMA bulk copies,
// applies every entangling operation of the layer on register-resident amplitude quadruples with
// per-shot predicated Pauli events, optionally emits the joint ancilla histogram, and streams the
// tile back with TMA bulk stores.  Algorithmic traffic per (shot, layer): 2 MiB read + 2 MiB write.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/s17.h"

namespace s17 {

constexpr int kDim = 1 << 17;
constexpr int kTileBits = 12;
constexpr int kTile = 1 << kTileBits;          // amplitudes per tile (64 KiB)
constexpr int kTilesPerShot = kDim / kTile;    // 32
constexpr int kRun = 512;                      // contiguous amplitudes per ancilla value (all data bits)
constexpr int kRunBytes = kRun * 16;
constexpr int kLayerThreads = 256;

struct LayerArg {
    s17_layer L;
    double2 prm[64];          // (cos, sin) of theta/2 for S17_U_ROT_* uops, layer-local indices
    uint8_t other_bits[5];    // the ancilla bits NOT staged in the tile, ascending
    uint8_t pad[3];
};

struct ShotAux {
    double scale;             // 1/sqrt(P(outcome)) of the last ancilla measurement
    uint32_t anc_outcome;     // raw 8-bit ancilla outcome of the last measurement
    uint32_t collapsed;       // 1: state already collapsed + reset to ancilla block 0
};

__device__ __forceinline__ int insert0(int x, int pos) { return (x & ((1 << pos) - 1)) | ((x >> pos) << (pos + 1)); }

// ------------------------------------------------------------------------------------------------
// Philox4x32-10, key = seed, counter = (global shot id, stream, block)
// ------------------------------------------------------------------------------------------------
struct Philox {
    uint32_t k0, k1;
    uint32_t c0, c1, c2, c3;
    uint32_t buf[4];
    int have;
    __device__ Philox(uint64_t seed, uint64_t shot, uint32_t stream)
        : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), c0((uint32_t)shot), c1((uint32_t)(shot >> 32)),
          c2(stream), c3(0), have(0) {}
    __device__ void refill() {
        uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3, a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            const uint32_t y0 = hi1 ^ x1 ^ a, y1 = lo1, y2 = hi0 ^ x3 ^ b, y3 = lo0;
            x0 = y0; x1 = y1; x2 = y2; x3 = y3;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        buf[0] = x0; buf[1] = x1; buf[2] = x2; buf[3] = x3;
        ++c3;
        have = 4;
    }
    // uniform double in [0,1) with 53 random bits
    __device__ double next() {
        if (have < 2) refill();
        const uint32_t a = buf[4 - have], b = buf[5 - have];
        have -= 2;
        return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
    }
};

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D TMA bulk copies
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store_1d(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// register-level gate arithmetic on a[4]; index = data_bit_value + 2 * ancilla_bit_value.
// The +-pi/2 gates are applied WITHOUT their 1/sqrt(2) factor (pure add/sub); the layer multiplies
// by 2^(-n/2) once, n = number of such gates the shot really executed.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void g_rxx(double2 (&a)[4], double s) {
    // Rxx(s*pi/2) * sqrt2 : a_k' = a_k - i s a_{3-k};  -i s (x + i y) = s y - i s x
    const double2 n0 = {fma(s, a[3].y, a[0].x), fma(-s, a[3].x, a[0].y)};
    const double2 n3 = {fma(s, a[0].y, a[3].x), fma(-s, a[0].x, a[3].y)};
    const double2 n1 = {fma(s, a[2].y, a[1].x), fma(-s, a[2].x, a[1].y)};
    const double2 n2 = {fma(s, a[1].y, a[2].x), fma(-s, a[1].x, a[2].y)};
    a[0] = n0; a[1] = n1; a[2] = n2; a[3] = n3;
}
__device__ __forceinline__ void g_rx(double2 (&a)[4], double s) {
    // Rx(s*pi/2) * sqrt2 on the data bit: a0' = a0 - i s a1, a1' = a1 - i s a0
#pragma unroll
    for (int h = 0; h < 4; h += 2) {
        const double2 n0 = {fma(s, a[h + 1].y, a[h].x), fma(-s, a[h + 1].x, a[h].y)};
        const double2 n1 = {fma(s, a[h].y, a[h + 1].x), fma(-s, a[h].x, a[h + 1].y)};
        a[h] = n0; a[h + 1] = n1;
    }
}
__device__ __forceinline__ void g_ry(double2 (&a)[4], double v) {
    // Ry(v*pi/2) * sqrt2 on the data bit: a0' = a0 - v a1, a1' = v a0 + a1
#pragma unroll
    for (int h = 0; h < 4; h += 2) {
        const double2 n0 = {fma(-v, a[h + 1].x, a[h].x), fma(-v, a[h + 1].y, a[h].y)};
        const double2 n1 = {fma(v, a[h].x, a[h + 1].x), fma(v, a[h].y, a[h + 1].y)};
        a[h] = n0; a[h + 1] = n1;
    }
}
__device__ __forceinline__ void g_h(double2 (&a)[4]) {
#pragma unroll
    for (int h = 0; h < 4; h += 2) {
        const double2 n0 = {a[h].x + a[h + 1].x, a[h].y + a[h + 1].y};
        const double2 n1 = {a[h].x - a[h + 1].x, a[h].y - a[h + 1].y};
        a[h] = n0; a[h + 1] = n1;
    }
}
__device__ __forceinline__ void g_rot_x(double2 (&a)[4], double c, double s) {
#pragma unroll
    for (int h = 0; h < 4; h += 2) {
        const double2 n0 = {fma(s, a[h + 1].y, c * a[h].x), fma(-s, a[h + 1].x, c * a[h].y)};
        const double2 n1 = {fma(s, a[h].y, c * a[h + 1].x), fma(-s, a[h].x, c * a[h + 1].y)};
        a[h] = n0; a[h + 1] = n1;
    }
}
__device__ __forceinline__ void g_rot_y(double2 (&a)[4], double c, double s) {
#pragma unroll
    for (int h = 0; h < 4; h += 2) {
        const double2 n0 = {fma(-s, a[h + 1].x, c * a[h].x), fma(-s, a[h + 1].y, c * a[h].y)};
        const double2 n1 = {fma(s, a[h].x, c * a[h + 1].x), fma(s, a[h].y, c * a[h + 1].y)};
        a[h] = n0; a[h + 1] = n1;
    }
}
__device__ __forceinline__ void g_rot_xx(double2 (&a)[4], double c, double s) {
    const double2 n0 = {fma(s, a[3].y, c * a[0].x), fma(-s, a[3].x, c * a[0].y)};
    const double2 n3 = {fma(s, a[0].y, c * a[3].x), fma(-s, a[0].x, c * a[3].y)};
    const double2 n1 = {fma(s, a[2].y, c * a[1].x), fma(-s, a[2].x, c * a[1].y)};
    const double2 n2 = {fma(s, a[1].y, c * a[2].x), fma(-s, a[1].x, c * a[2].y)};
    a[0] = n0; a[1] = n1; a[2] = n2; a[3] = n3;
}
// event field f: bit0 x(data) bit1 z(data) bit2 x(anc) bit3 z(anc) bits4-5 k;  P = i^k X^x Z^z (Z first)
__device__ __forceinline__ void g_pauli(double2 (&a)[4], uint32_t f) {
    if (f & 2u) { a[1].x = -a[1].x; a[1].y = -a[1].y; a[3].x = -a[3].x; a[3].y = -a[3].y; }
    if (f & 8u) { a[2].x = -a[2].x; a[2].y = -a[2].y; a[3].x = -a[3].x; a[3].y = -a[3].y; }
    if (f & 1u) { double2 t = a[0]; a[0] = a[1]; a[1] = t; t = a[2]; a[2] = a[3]; a[3] = t; }
    if (f & 4u) { double2 t = a[0]; a[0] = a[2]; a[2] = t; t = a[1]; a[1] = a[3]; a[3] = t; }
    const uint32_t k = (f >> 4) & 3u;
    if (k) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double2 t = a[i];
            if (k == 1) a[i] = make_double2(-t.y, t.x);
            else if (k == 2) a[i] = make_double2(-t.x, -t.y);
            else a[i] = make_double2(t.y, -t.x);
        }
    }
}

}  // namespace s17
