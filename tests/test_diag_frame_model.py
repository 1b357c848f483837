"""The diagonal-frame half cycle (tests/diag_model.py, the algorithm of k_cycle) against a brute-force simulation."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import diag_model as dm                                     # noqa: E402
from surface_17_iqt_b200 import circuit                      # noqa: E402


def halves():
    ops = circuit.build_cycle()
    out = []
    for lo, frame_x in ((0, True), (12, False)):
        sel = ops[lo:lo + 12]
        anc = []
        for o in sel:
            if o.ancilla not in anc:
                anc.append(o.ancilla)
        assert len(anc) == 4
        out.append(([dict(d=o.data, slot=anc.index(o.ancilla), s=o.s, ts=o.timestep,
                          shape=int(o.has_ry1) | (int(o.has_rx) << 1) | (int(o.has_ry2) << 2)) for o in sel], frame_x))
    return out


def random_words(rng, ops, density, skips):
    words = []
    for op in ops:
        w = 0
        skip = int(skips and rng.random() < 0.25)
        w |= skip << 24
        for shift, present in ((0, op["shape"] & 1), (6, not skip), (12, op["shape"] & 2), (18, op["shape"] & 4)):
            if present and rng.random() < density:
                w |= int(rng.integers(0, 64)) << shift
        words.append(w)
    return words


@pytest.mark.parametrize("density,skips", [(0.0, False), (0.15, False), (0.5, True), (1.0, True)])
def test_diag_half_equals_brute_force(density, skips):
    rng = np.random.default_rng(int(density * 100) + skips)
    for ops, frame_x in halves():
        for _ in range(3):
            psi = rng.normal(size=512) + 1j * rng.normal(size=512)
            psi /= np.linalg.norm(psi)
            words = random_words(rng, ops, density, skips)
            brute = dm.brute_half(psi, ops, words)
            out, factor = dm.diag_half(psi, ops, words, frame_x)
            for a in range(16):
                np.testing.assert_allclose(out[a] * factor, brute[512 * a:512 * (a + 1)], atol=1e-12, rtol=0)


@pytest.mark.parametrize("density", [0.0, 0.3])
def test_diag_half_with_idle_z_events(density):
    """sympathetic_cooling: Z on any data qubit / ancilla in the gaps after the timesteps (incl. the one before the
    measurement of the x-type half)."""
    rng = np.random.default_rng(7 + int(10 * density))
    for ops, frame_x in halves():
        steps = sorted({op["ts"] for op in ops})
        gaps = steps if frame_x else steps[:-1]            # no idle slots after timestep 8 (CS:790-836)
        for _ in range(4):
            psi = rng.normal(size=512) + 1j * rng.normal(size=512)
            psi /= np.linalg.norm(psi)
            words = random_words(rng, ops, density, density > 0)
            idle = {ts: [q for q in range(13) if rng.random() < 0.25] for ts in gaps}
            brute = dm.brute_half(psi, ops, words, idle)
            out, factor = dm.diag_half(psi, ops, words, frame_x, idle)
            for a in range(16):
                np.testing.assert_allclose(out[a] * factor, brute[512 * a:512 * (a + 1)], atol=1e-12, rtol=0)


def random_eps(rng, ops, scale):
    return [dict(ry1=rng.uniform(-scale, scale), rxx=rng.uniform(-scale, scale), rx=rng.uniform(-scale, scale),
                 ry2=rng.uniform(-scale, scale)) for _ in ops]


@pytest.mark.parametrize("density,skips", [(0.0, False), (0.3, False), (0.6, True)])
def test_coherent_half_with_any_pauli_event_equals_brute_force(density, skips):
    """Over-rotations of +-0.3 rad after every gate, Pauli events of every kind between the gates, leak skips: the
    diagonal-frame algorithm with every event treated where it acts (per-gate flip state, signs collected, ancilla
    parts on the (u, v) pair, the rotation after a bracket inverted by the flips inside it) against the gate-by-gate
    simulation, every joint outcome, 1e-12."""
    rng = np.random.default_rng(40 + int(density * 10) + skips)
    for ops, frame_x in halves():
        for _ in range(4):
            psi = rng.normal(size=512) + 1j * rng.normal(size=512)
            psi /= np.linalg.norm(psi)
            words = random_words(rng, ops, density, skips)
            eps = random_eps(rng, ops, 0.3)
            brute = dm.brute_half_coh(psi, ops, words, eps)
            out, factor = dm.diag_half_coh(psi, ops, words, frame_x, eps)
            for a in range(16):
                np.testing.assert_allclose(out[a] * factor, brute[512 * a:512 * (a + 1)], atol=1e-12, rtol=0)
