"""The closed form k_cycle1 uses for the private table entries of an ancilla slot (s17_diag.cuh: dg_slot_entry, reached
through the host-only hook s17_diag_slot_entry) against the gate-by-gate recursion it replaces: per operation of the slot
Rxx(s pi/2) unless leak-skipped (CS:293, 322), the Rx phase, then the net Z / X event on the ancilla (insert_error,
CS:188-234) and an idle Z (sympathetic_cooling, CS:762-772).  Exact integer arithmetic: equality, not a tolerance."""
import ctypes as C
import itertools
import random

from surface_17_iqt_b200 import _lib


def spread(m):
    return (m * 0x00204081) & 0x01010101


def recursion(n, f1, idx, flip, skip, zb, xb, zc, f2):
    u, v, ph = 1 + 0j, 0j, 1 + 0j
    for j in range(n):
        sg = -1 if ((f1 >> j) ^ (idx >> j) ^ (flip >> j)) & 1 else 1
        if not (skip >> j) & 1:
            u, v = u - 1j * sg * v, v - 1j * sg * u
        if (f2 >> j) & 1:
            ph *= 1 + 1j * sg
        if (zb >> j) & 1:
            v = -v
        if (xb >> j) & 1:
            u, v = v, u
        if (zc >> j) & 1:
            v = -v
    return ph * u, ph * v


def closed(L, n, f1, idx, flip, skip, zb, xb, zc, f2, junk=0):
    W = 0
    for j in range(4):
        if j < n:
            t = (junk & 0x66) | (((flip >> j) & 1) << 7) | (((zb >> j) & 1) << 4) | (((xb >> j) & 1) << 3) | ((skip >> j) & 1)
        else:
            t = 0xFF                      # bytes beyond the slot's operations must not matter
        W |= t << (8 * j)
    mm = spread((1 << n) - 1)
    out = (C.c_int32 * 4)()
    # bits beyond the slot's operations are set in f1 / idx / cz as well
    extra = 0x01010101 & ~mm
    assert L.s17_diag_slot_entry(W, spread(f1) | extra, spread(f2), spread(idx) | extra, spread(zc) | extra, mm, out) == 0
    return complex(out[0], out[1]), complex(out[2], out[3])


def test_slot_entry_closed_form_equals_the_recursion():
    L = _lib.load()
    rnd = random.Random(17)
    n_cases = 0
    # every event pattern of slots with one to three operations, all sign / index / phase patterns
    for n in (1, 2, 3):
        R = range(1 << n)
        for f1, idx, flip, skip, zb, xb, zc in itertools.product(R, R, R, R, R, R, R):
            if n == 3 and (f1 & 2 or zc & 5):
                continue                  # (keeps the run in seconds; the four-operation sample below covers the rest)
            f2 = rnd.randrange(1 << n)
            assert closed(L, n, f1, idx, flip, skip, zb, xb, zc, f2, rnd.randrange(256)) == \
                recursion(n, f1, idx, flip, skip, zb, xb, zc, f2), (n, f1, idx, flip, skip, zb, xb, zc, f2)
            n_cases += 1
    for _ in range(60000):
        a = [rnd.randrange(16) for _ in range(8)]
        assert closed(L, 4, *a, rnd.randrange(256)) == recursion(4, *a), a
        n_cases += 1
    assert n_cases > 100000


def test_slot_entry_without_events_is_the_clean_table():
    """No event: (A0, A1) = ph (u, v) of the plain product of (1 - i s X_a) factors -- e.g. two operations of opposite sign
    cancel to 2 |0>, four equal ones give -4 |0>."""
    L = _lib.load()
    assert closed(L, 2, 0b01, 0, 0, 0, 0, 0, 0, 0) == (2 + 0j, 0j)
    assert closed(L, 4, 0, 0, 0, 0, 0, 0, 0, 0) == (-4 + 0j, 0j)
    assert closed(L, 1, 0, 0, 0, 0, 0, 0, 0, 0) == (1 + 0j, -1j)
    assert closed(L, 1, 0, 0, 0, 1, 0, 0, 0, 0) == (1 + 0j, 0j)          # leak-skipped: identity


def test_ancilla_x_events_only_swap_the_outcomes():
    """What k_cycle1 does instead of building private entries for a slot whose only events are net X's on its ancilla (an
    XX error, say): X_a commutes with every Rxx of the slot, so with an odd number of them the two outcome amplitudes of
    EVERY table entry swap -- data flips seen by the operations (bit 31: an XOR on the table index in the kernel) included --
    and with an even number nothing changes.  Checked on the recursion (the reference's gate order) and on the closed form."""
    L = _lib.load()
    rnd = random.Random(5)
    for n in (1, 2, 3, 4):
        R = range(1 << n)
        for f1, idx, xb in itertools.product(R, R, R):
            flip, f2 = rnd.randrange(1 << n), rnd.randrange(1 << n)
            clean = recursion(n, f1, idx ^ flip, 0, 0, 0, 0, 0, f2)          # the clean table at the flipped index
            want = clean[::-1] if bin(xb).count("1") & 1 else clean
            assert recursion(n, f1, idx, flip, 0, 0, xb, 0, f2) == want, (n, f1, idx, flip, xb, f2)
            assert closed(L, n, f1, idx, flip, 0, 0, xb, 0, f2) == want, (n, f1, idx, flip, xb, f2)


def test_ancilla_z_event_is_an_index_flip_and_a_phase():
    """The next shortcut (not in the kernel yet): one net Z on the ancilla after operation j of a slot swaps the two
    X_a eigen-components, which later operations then rotate the other way -- the entry equals the CLEAN entry at the index
    with the bits of the later operations flipped, the outcome-1 amplitude negated, times i s_k lambda_k for every later
    operation k that carries an Rx."""
    for n in (2, 3, 4):
        R = range(1 << n)
        for f1, idx, f2 in itertools.product(R, R, R):
            for j in range(n):
                later = ((1 << n) - 1) & ~((1 << (j + 1)) - 1)
                u, v = recursion(n, f1, idx ^ later, 0, 0, 0, 0, 0, f2)
                ph = 1 + 0j
                for k in range(j + 1, n):
                    if (f2 >> k) & 1:
                        ph *= 1j * (-1 if ((f1 >> k) ^ (idx >> k)) & 1 else 1)
                assert recursion(n, f1, idx, 0, 0, 1 << j, 0, 0, f2) == (ph * u, -ph * v), (n, f1, idx, f2, j)
