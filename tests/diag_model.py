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


# ---- coherent over-rotations (the `*_over` entries; k_cycle1<true>) -----------------------------------------------------
def rot(name, theta):
    c, s = np.cos(0.5 * theta), np.sin(0.5 * theta)
    if name == "x":
        return np.array([[c, -1j * s], [-1j * s, c]], dtype=complex)
    return np.array([[c, -s], [s, c]], dtype=complex)          # "y"


def apply_xx_angle(psi, bd, ba, theta):
    idx = np.arange(psi.size)
    flipped = psi[idx ^ ((1 << bd) | (1 << ba))]
    psi[:] = np.cos(0.5 * theta) * psi - 1j * np.sin(0.5 * theta) * flipped


def brute_half_coh(psi512, ops, words, eps):
    """Gate by gate with over-rotations: every gate is followed by the same-axis rotation eps[i][gate] (gate in
    'ry1', 'rxx', 'rx', 'ry2'), then by its event field -- the order of the `*_over` entries and the stochastic entries in
    an error model (coherent first).  A leak-skipped Rxx skips its over-rotation and its field too."""
    psi = np.zeros(1 << 13, dtype=complex)
    psi[:512] = psi512
    for op, w, e in zip(ops, words, eps):
        bd, ba, s, shape = op["d"], 9 + op["slot"], op["s"], op["shape"]
        skip = (w >> 24) & 1
        if shape & 1:
            apply_1q(psi, bd, ry(+1))
            apply_1q(psi, bd, rot("y", e["ry1"]))
            apply_field(psi, w & 63, bd, ba)
        if not skip:
            apply_xx(psi, bd, ba, s)
            apply_xx_angle(psi, bd, ba, e["rxx"])
            apply_field(psi, (w >> 6) & 63, bd, ba)
        if shape & 2:
            apply_1q(psi, bd, rx(-s))
            apply_1q(psi, bd, rot("x", e["rx"]))
            apply_field(psi, (w >> 12) & 63, bd, ba)
        if shape & 4:
            apply_1q(psi, bd, ry(-1))
            apply_1q(psi, bd, rot("y", e["ry2"]))
            apply_field(psi, (w >> 18) & 63, bd, ba)
    return psi


def diag_half_coh(psi512, ops, words, frame_x, eps):
    """The diagonal-frame half cycle with over-rotations and Pauli events of ANY kind, every event treated where it acts
    (no conjugation to the end of the operation: a Pauli pushed through a non-Clifford rotation is not a Pauli):

    * a frame-diagonal data part (sign type) commutes with every gate of the half: collected in the sign mask;
    * a frame-flip data part relabels the basis for every LATER gate on that qubit (per-gate flip state) and ends in the
      flip mask; a sign met after a flip costs a factor -1;
    * an ancilla part acts on the ancilla's (u, v) pair where it occurs;
    * Ry(eps) after Ry(+pi/2) / Ry(-pi/2): a real rotation of the data qubit before / after its bracket; flips that
      happened inside the bracket invert the angle of the one after it (X Ry(e) X = Ry(-e)), and the signs of events that
      come after it are applied after it.

    Returns out[a] and the factor to the brute-force projection, as diag_half."""
    nq = 9
    first = {}
    last = {}
    for i, op in enumerate(ops):
        first.setdefault(op["d"], i)
        last[op["d"]] = i
    S = psi512.astype(complex).copy()
    if frame_x:
        S = wht(S)
    else:
        for q in range(nq):                                   # pre-layer: Ry(eps_a) of the operation that opens the bracket
            if q in first and ops[first[q]]["shape"] & 1:
                apply_1q(S, q, rot("y", eps[first[q]]["ry1"]))
    Fm = ZACC = g = 0
    zacc_out = 0                                              # signs of events after a closed bracket (after its Ry(eps))
    flip_in = 0                                               # flips that happened inside a (still open or closed) bracket
    per_op = []
    for i, (op, w) in enumerate(zip(ops, words)):
        d, shape = op["d"], op["shape"]
        skip = (w >> 24) & 1
        rec = dict(pre=(0, 0), post=[], fl_rxx=0, fl_rx=0)
        inside = True                                         # x-type half, or inside the Ry bracket

        def take(field, inside_bracket):
            nonlocal Fm, ZACC, g, flip_in, zacc_out
            x, z, xa, za, k = field & 1, (field >> 1) & 1, (field >> 2) & 1, (field >> 3) & 1, (field >> 4) & 3
            F = frame_pauli(field & 0x33, frame_x, 0 if inside_bracket else 4)
            xf, zf, kf = F & 1, (F >> 1) & 1, (F >> 4) & 3
            if zf:
                if inside_bracket:
                    ZACC ^= 1 << d
                else:
                    zacc_out ^= 1 << d
                g += 2 * ((Fm >> d) & 1)
            if xf:
                Fm ^= 1 << d
                if inside_bracket:
                    flip_in ^= 1 << d
            g += kf
            return xa, za
        if shape & 1:
            rec["pre"] = take(w & 63, True)
        rec["fl_rxx"] = (Fm >> d) & 1
        if not skip:
            rec["post"].append(take((w >> 6) & 63, True))
        rec["fl_rx"] = (Fm >> d) & 1
        if shape & 2:
            rec["post"].append(take((w >> 12) & 63, True))
        if shape & 4:
            rec["post"].append(take((w >> 18) & 63, False))
        per_op.append(rec)
    p = np.arange(512)
    amp = {a: S.copy() for a in range(16)}
    for k in range(4):
        mine = [i for i, op in enumerate(ops) if op["slot"] == k]
        T = np.zeros((1 << len(mine), 2), dtype=complex)
        for idx in range(1 << len(mine)):
            u, v, ph = 1.0 + 0j, 0j, 1.0 + 0j
            for j, i in enumerate(mine):
                op, w, rec, e = ops[i], words[i], per_op[i], eps[i]
                bit = (idx >> j) & 1
                xa, za = rec["pre"]
                if za:
                    v = -v
                if xa:
                    u, v = v, u
                if not ((w >> 24) & 1):
                    lam = (-1) ** (bit ^ rec["fl_rxx"])
                    th = 0.5 * (op["s"] * 0.5 * np.pi + e["rxx"])
                    c, sn = np.cos(th), lam * np.sin(th)
                    u, v = c * u - 1j * sn * v, c * v - 1j * sn * u
                if op["shape"] & 2:
                    lam = (-1) ** (bit ^ rec["fl_rx"])
                    th = 0.5 * (-op["s"] * 0.5 * np.pi + e["rx"])
                    ph *= np.cos(th) - 1j * lam * np.sin(th)
                for xa, za in rec["post"]:
                    if za:
                        v = -v
                    if xa:
                        u, v = v, u
            T[idx] = (ph * u, ph * v)
        idx_of_p = np.zeros(512, dtype=int)
        for j, i in enumerate(mine):
            idx_of_p |= ((p >> ops[i]["d"]) & 1) << j
        for a in range(16):
            amp[a] = amp[a] * T[idx_of_p, (a >> k) & 1]
    sign = (-1.0) ** np.array([bin(int(q) & ZACC).count("1") for q in p])
    out = {}
    for a in range(16):
        Sa = amp[a] * sign * (1j ** (g & 3))
        if frame_x:
            W = wht(Sa)
            fix = (-1.0) ** np.array([bin(int(z) & Fm).count("1") for z in p])
            out[a] = W * fix
        else:
            for q in range(nq):                               # post-layer, before the relabelling by the flips
                if q in last and ops[last[q]]["shape"] & 4:
                    e = eps[last[q]]["ry2"]
                    apply_1q(Sa, q, rot("y", -e if (flip_in >> q) & 1 else e))
            Sa = Sa * (-1.0) ** np.array([bin(int(q) & zacc_out).count("1") for q in p])
            o = np.zeros(512, dtype=complex)
            o[p ^ Fm] = Sa
            out[a] = o
    return out, (1.0 / 512.0 if frame_x else 1.0)
