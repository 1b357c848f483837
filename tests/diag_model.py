"""Executable derivation of the diagonal-frame half cycle (k_cycle in s17_diag.cuh), in numpy.

Not product code and not an oracle: a CPU model of the *algorithm* the CUDA kernel implements, checked in
tests/test_diag_frame_model.py against a brute-force 2^13-amplitude simulation of the same half cycle (gate by gate,
Pauli events gate by gate).  It pins the sign / phase conventions of

* k_plan's net end-of-operation Pauli (pmul2 / pconj_* in s17.cu) and its frame bookkeeping (flip mask, sign mask,
  global phase, per-op flip bit), and
* the per-ancilla 16-entry tables, the Walsh-Hadamard frame of the X-type half and the final index / sign fix-up.

Event field (6 bits): bit0 x_d, bit1 z_d, bit2 x_a, bit3 z_a, bits 4-5 k:  P = i^k (X^x Z^z)_d (X^x Z^z)_a, Z first.
"""
import numpy as np

SQ = 1.0 / np.sqrt(2.0)


# ---- Pauli algebra on 6-bit fields (mirrors s17.cu) -------------------------------------------------
def pmul2(A, B):
    sign = ((A >> 1) & B & 1) + ((A >> 3) & (B >> 2) & 1)
    k = ((A >> 4) + (B >> 4) + 2 * sign) & 3
    return ((A ^ B) & 15) | (k << 4)


def pconj(P, img_xd, img_zd, img_xa, img_za):
    R = P & 0x30
    if P & 1:
        R = pmul2(R, img_xd)
    if P & 2:
        R = pmul2(R, img_zd)
    if P & 4:
        R = pmul2(R, img_xa)
    if P & 8:
        R = pmul2(R, img_za)
    return R


def pconj_rx(P, sigma):
    return pconj(P, 0x01, 0x03 | ((3 if sigma > 0 else 1) << 4), 0x04, 0x08)


def pconj_ry(P, v):
    return pconj(P, 0x02 | ((2 if v > 0 else 0) << 4), 0x01 | ((0 if v > 0 else 2) << 4), 0x04, 0x08)


def pconj_rxx(P, s):
    k = (3 if s > 0 else 1) << 4
    return pconj(P, 0x01, 0x07 | k, 0x04, 0x0D | k)


def net_pauli(w, shape, s, skip):
    """k_plan, S17_P_OP_END: every per-gate event conjugated to the end of the operation."""
    N = 0
    if shape & 1:
        N = w & 63
    if not skip:
        N = pconj_rxx(N, s)
        N = pmul2((w >> 6) & 63, N)
    if shape & 2:
        N = pconj_rx(N, -s)
        N = pmul2((w >> 12) & 63, N)
    if shape & 4:
        N = pconj_ry(N, -1)
        N = pmul2((w >> 18) & 63, N)
    return N


def frame_pauli(N, frame_x, shape):
    """The net Pauli as it acts in the frame of the half cycle (data part only changes)."""
    if frame_x:                      # H (X^x Z^z) H = Z^x X^z = (-1)^{xz} X^z Z^x
        x, z, k = N & 1, (N >> 1) & 1, (N >> 4) & 3
        return (N & 0x0C) | z | (x << 1) | (((k + 2 * x * z) & 3) << 4)
    if not (shape & 4):              # still inside the Ry bracket of its data qubit: conjugate out by Ry(-pi/2)
        return pconj_ry(N, -1)
    return N


# ---- brute force on 9 data + 4 ancilla-slot bits ------------------------------------------------------
def _pair_view(psi, bit):
    n = psi.size
    hi = n >> (bit + 1)
    return psi.reshape(hi, 2, 1 << bit)


def apply_1q(psi, bit, M):
    v = _pair_view(psi, bit)
    a0, a1 = v[:, 0, :].copy(), v[:, 1, :].copy()
    v[:, 0, :] = M[0, 0] * a0 + M[0, 1] * a1
    v[:, 1, :] = M[1, 0] * a0 + M[1, 1] * a1


def apply_xx(psi, bd, ba, s):
    idx = np.arange(psi.size)
    flipped = psi[idx ^ ((1 << bd) | (1 << ba))]
    psi[:] = (psi - 1j * s * flipped) * SQ


X = np.array([[0, 1], [1, 0]], dtype=complex)
Z = np.array([[1, 0], [0, -1]], dtype=complex)


def apply_field(psi, f, bd, ba):
    if f & 2:
        apply_1q(psi, bd, Z)
    if f & 8:
        apply_1q(psi, ba, Z)
    if f & 1:
        apply_1q(psi, bd, X)
    if f & 4:
        apply_1q(psi, ba, X)
    psi *= 1j ** ((f >> 4) & 3)


def ry(v):
    return np.array([[1, -v], [v, 1]], dtype=complex) * SQ


def rx(s):
    return np.array([[1, -1j * s], [-1j * s, 1]], dtype=complex) * SQ


def brute_half(psi512, ops, words, idle=None):
    """ops: list of dict(d, slot, s, shape[, ts]) in time order; words[i]: event word of op i; idle: {ts: [qubits]} = Z
    on data qubit q < 9 / ancilla slot q - 9 in the gap after timestep ts.  Returns the 2^13 state."""
    psi = np.zeros(1 << 13, dtype=complex)
    psi[:512] = psi512
    idle = idle or {}
    for n, (op, w) in enumerate(zip(ops, words)):
        if n and ops[n - 1].get("ts") != op.get("ts"):
            for q in idle.get(ops[n - 1].get("ts"), []):
                apply_1q(psi, q, Z)
        bd, ba, s, shape = op["d"], 9 + op["slot"], op["s"], op["shape"]
        skip = (w >> 24) & 1
        if shape & 1:
            apply_1q(psi, bd, ry(+1))
            apply_field(psi, w & 63, bd, ba)
        if not skip:
            apply_xx(psi, bd, ba, s)
            apply_field(psi, (w >> 6) & 63, bd, ba)
        if shape & 2:
            apply_1q(psi, bd, rx(-s))
            apply_field(psi, (w >> 12) & 63, bd, ba)
        if shape & 4:
            apply_1q(psi, bd, ry(-1))
            apply_field(psi, (w >> 18) & 63, bd, ba)
    for q in idle.get(ops[-1].get("ts"), []):
        apply_1q(psi, q, Z)
    return psi


# ---- the diagonal-frame algorithm ------------------------------------------------------------------------
def wht(v):
    v = v.copy()
    n = v.size
    h = 1
    while h < n:
        w = v.reshape(n // (2 * h), 2, h)
        a, b = w[:, 0, :].copy(), w[:, 1, :].copy()
        w[:, 0, :] = a + b
        w[:, 1, :] = a - b
        h *= 2
    return v


def plan_frame(ops, words, frame_x, idle=None):
    """k_plan's frame bookkeeping for one half: per-op flip bit, final flip mask, sign mask, global phase, and the
    operations after which an idle Z hits their ancilla."""
    Fm = ZACC = g = 0
    sflip, nets = [], []
    ancz = set()
    idle = idle or {}

    def idle_gap(ts, upto):
        nonlocal Fm, ZACC, g
        for q in idle.get(ts, []):
            if q < 9:
                mine = [o["ts"] for o in ops if o["d"] == q]
                if frame_x:
                    Fm ^= 1 << q                               # H Z H = X
                elif mine and min(mine) <= ts < max(mine):
                    Fm ^= 1 << q                               # Ry(-pi/2) Z Ry(+pi/2) = -X
                    g += 2
                else:
                    ZACC ^= 1 << q
                    g += 2 * ((Fm >> q) & 1)
            else:
                done = [i for i in range(upto) if ops[i]["slot"] == q - 9]
                if done:
                    ancz.symmetric_difference_update({done[-1]})

    for n, (op, w) in enumerate(zip(ops, words)):
        if n and ops[n - 1].get("ts") != op.get("ts"):
            idle_gap(ops[n - 1].get("ts"), n)
        d = op["d"]
        skip = (w >> 24) & 1
        N = net_pauli(w, op["shape"], op["s"], skip) if (w & 0xFFFFFF) else 0
        nets.append(N)
        sflip.append((Fm >> d) & 1)
        F = frame_pauli(N, frame_x, op["shape"])
        xf, zf, kf = F & 1, (F >> 1) & 1, (F >> 4) & 3
        if zf:
            ZACC ^= 1 << d
            g += 2 * ((Fm >> d) & 1)
        if xf:
            Fm ^= 1 << d
        g += kf
    idle_gap(ops[-1].get("ts"), len(ops))
    return sflip, nets, Fm, ZACC, g & 3, ancz


def diag_half(psi512, ops, words, frame_x, idle=None):
    """Returns out[a] (512 amplitudes, unnormalised) for every joint outcome a of the 4 ancilla slots, and the factor
    that relates it to the brute-force projection."""
    sflip, nets, Fm, ZACC, g, ancz = plan_frame(ops, words, frame_x, idle)
    S = wht(psi512) if frame_x else psi512.copy()
    p = np.arange(512)
    amp = {a: S.copy() for a in range(16)}
    n_factors = 0
    for k in range(4):
        mine = [i for i, op in enumerate(ops) if op["slot"] == k]
        assert len({ops[i]["d"] for i in mine}) == len(mine)
        T = np.zeros((1 << len(mine), 2), dtype=complex)
        for idx in range(1 << len(mine)):
            u, v, ph = 1.0 + 0j, 0j, 1.0 + 0j
            for j, i in enumerate(mine):
                op, w = ops[i], words[i]
                bit = (idx >> j) & 1
                sig = op["s"] * (-1) ** (bit ^ sflip[i])
                if not ((w >> 24) & 1):
                    u, v = u - 1j * sig * v, v - 1j * sig * u
                if op["shape"] & 2:
                    ph *= 1 + 1j * sig
                N = nets[i]
                if N & 8:
                    v = -v
                if N & 4:
                    u, v = v, u
                if i in ancz:
                    v = -v
            T[idx] = (ph * u, ph * v)
        for i in mine:
            n_factors += (0 if (words[i] >> 24) & 1 else 1) + (1 if ops[i]["shape"] & 2 else 0)
        idx_of_p = np.zeros(512, dtype=int)
        for j, i in enumerate(mine):
            idx_of_p |= ((p >> ops[i]["d"]) & 1) << j
        for a in range(16):
            amp[a] = amp[a] * T[idx_of_p, (a >> k) & 1]
    sign = (-1.0) ** np.array([bin(int(q) & ZACC).count("1") for q in p])
    out = {}
    for a in range(16):
        Sa = amp[a] * sign * (1j ** g)
        if frame_x:
            W = wht(Sa)
            fix = (-1.0) ** np.array([bin(int(z) & Fm).count("1") for z in p])
            out[a] = W * fix
        else:
            o = np.zeros(512, dtype=complex)
            o[p ^ Fm] = Sa
            out[a] = o
    factor = SQ ** n_factors / (512.0 if frame_x else 1.0)
    return out, factor
