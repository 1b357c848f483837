"""GPU parity tests for coherent over-rotations (BASELINE config 3; the `*_over` entries, SURVEY C-10) on the whole-cycle
kernel's coherent form (k_cycle1<true>, s17_diag.cuh: DiagCoh) -- against the oracle (one CPU sweep per gate), against the
sweep kernels (S17_COH=0), and statistically against oracle sampling.  All through the C ABI."""
import random

import numpy as np
import pytest

from oracle import s17_oracle as so
from oracle import sv
from surface_17_iqt_b200 import _lib as L
from surface_17_iqt_b200 import api, circuit

pytestmark = pytest.mark.gpu

AMP_TOL = 1e-10
FIELDS = ("quiescent", "synd1", "synd2", "flips_a", "ft_syndrome", "correction", "data_meas", "logical_meas", "fail", "not0")
MODEL = "coherent_model"


def make_sim(models, ct, bt, list_len, n, angles, fusion=2):
    return api.Simulation(MODEL, list_len, max_shots=n, fusion=fusion, correction_table=ct, bitflip_table=bt,
                          error_models=models, over_angles=angles)


def random_angles(rounds, seed, scale):
    """A different angle in every slot (both signs), so that a slot mix-up cannot cancel."""
    rng = random.Random(seed)
    return {"eps_1q": [rng.uniform(-scale, scale) for _ in range(30 * rounds)],
            "eps_2q": [rng.uniform(-scale, scale) for _ in range(24 * rounds)]}


def xx_masks(n, rounds, k, seed):
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        m = [0] * (24 * rounds)
        for j in rng.sample(range(24 * rounds), k):
            m[j] = 1
        out.append(m)
    return out


@pytest.mark.parametrize("protocol,rounds", [("single", 1), ("s1", 1), ("s2", 2)])
def test_coherent_cycles_on_the_whole_cycle_kernel_against_oracle(error_models, correction_table, bitflip_table, protocol,
                                                                   rounds, monkeypatch):
    """Slot-dependent over-rotations of +-0.2 rad on every gate and 3 stochastic XX errors per shot, outcomes sampled on
    the device; the oracle replays masks and outcomes gate by gate.  Records bit-exact; the 512-amplitude block the
    kernel stores after the last executed cycle equals the oracle's state there within 1e-10 per amplitude (phase
    included), and nothing leaks out of the block."""
    monkeypatch.setenv("S17_DIAG", "1")
    n = 24
    angles = random_angles(rounds, 7, 0.2)
    lens = [30 * rounds, 24 * rounds, 24 * rounds]
    sim = make_sim(error_models, correction_table, bitflip_table, lens, n, angles)
    assert [i["coherent"] for i in sim.backend.program_info] == [False, True, True]
    masks = xx_masks(n, rounds, 3, 11)
    sim.backend.set_replay([[0] * (54 * rounds) + m for m in masks], [[0] * 40 for _ in range(n)])
    c = sim.backend.run(protocol, n, L.NOISE_REPLAY, seed=5, forced=False)
    assert c.errors == 0
    assert c.sweeps <= 3 * n                               # one whole-cycle sweep per shot and cycle: really k_cycle1
    recs = sim.backend.records(n)
    orc = so.ShotOracle(error_models, correction_table, bitflip_table)
    gates = error_models[MODEL]["gates"]
    zero = {"eps_1q": [0.0] * (30 * rounds), "eps_2q": [0.0] * (24 * rounds), "p_XX_ctrl": [0] * (24 * rounds)}
    fails = 0
    for i in range(n):
        p_set = [angles["eps_1q"], angles["eps_2q"], masks[i]]
        ref = orc.run_shot(MODEL, p_set, protocol, rnd=random.Random(0), meas=so.ForcedMeasure(recs[i]["outcomes"]))
        for f in FIELDS:
            assert recs[i].get(f) == ref.get(f), (i, f, recs[i].get(f), ref.get(f))
        fails += ref["fail"]
        # the state after the last executed cycle, before lookup / apply_correction
        st = sv.StateVector(17)
        probs = dict(zip(error_models[MODEL]["probabilities"], p_set))
        meas = so.ForcedMeasure(recs[i]["outcomes"])
        leaked = 17 * [0]
        orc.logical_prep(st, "Z", 0, leaked, gates, zero, random.Random(0), meas)
        orc.stabilizer_cycle(st, leaked, gates, probs, random.Random(0), meas, stab_ind=0)
        if protocol == "s2" and ref["not0"]:
            orc.stabilizer_cycle(st, leaked, gates, probs, random.Random(0), meas, stab_ind=1)
        psi = sim.backend.state(i)
        assert np.max(np.abs(psi[:512] - st.psi[:512])) < AMP_TOL, (i, np.max(np.abs(psi[:512] - st.psi[:512])))
        assert np.sum(np.abs(psi[512:]) ** 2) == 0.0 and np.sum(np.abs(st.psi[512:]) ** 2) < 1e-20
    assert c.failed == fails
    sim.close()


def test_coherent_form_matches_the_sweep_kernels(error_models, correction_table, bitflip_table, monkeypatch):
    """The same replayed trajectories (masks and outcomes) through k_cycle1<true> and through the general sweep kernels
    (S17_COH=0): identical records and counters."""
    n = 32
    angles = random_angles(2, 3, 0.1)
    lens = [60, 48, 48]
    masks = [[0] * 108 + m for m in xx_masks(n, 2, 2, 5)]
    monkeypatch.setenv("S17_COH", "1")
    fast = make_sim(error_models, correction_table, bitflip_table, lens, n, angles)
    fast.backend.set_replay(masks, [[0] * 40 for _ in range(n)])
    cf = fast.backend.run("s2", n, L.NOISE_REPLAY, seed=9, forced=False)
    rf = fast.backend.records(n)
    fast.close()
    monkeypatch.setenv("S17_COH", "0")
    slow = make_sim(error_models, correction_table, bitflip_table, lens, n, angles)
    assert not any(i["coherent"] for i in slow.backend.program_info)
    slow.backend.set_replay(masks, [r["outcomes"] for r in rf])
    cs = slow.backend.run("s2", n, L.NOISE_REPLAY, seed=9, forced=True, materialize=True)
    rs = slow.backend.records(n)
    assert cs.sweeps > 6 * n and cf.sweeps <= 3 * n
    assert (cf.failed, cf.not0, cf.errors) == (cs.failed, cs.not0, 0) and cs.errors == 0
    for i in range(n):
        for f in FIELDS:
            assert rf[i].get(f) == rs[i].get(f), (i, f)
    slow.close()


def test_coherent_rates_against_oracle_sampling(error_models, correction_table, bitflip_table):
    """True-probability mode (config 3's shape at a noise level where the rates are measurable): eps = 0.12 rad on every
    gate, p_XX = 0.03, single-round protocol.  2^18 device shots against 240 oracle shots: detection and failure rates
    within 3.5 binomial sigma of each other."""
    angles = {"eps_1q": [0.12] * 30, "eps_2q": [0.12] * 24}
    p_xx = [0.03] * 24
    n = 1 << 18
    sim = make_sim(error_models, correction_table, bitflip_table, [30, 24, 24], n, angles)
    sim.backend.set_slot_probs([[0.0] * 30, [0.0] * 24, p_xx])
    c = sim.backend.run("single", n, L.NOISE_PROB, seed=77)
    assert c.errors == 0 and c.sweeps <= 2 * n
    sim.close()
    orc = so.ShotOracle(error_models, correction_table, bitflip_table)
    rnd, mr = random.Random(1), random.Random(2)
    m = 240
    fails = not0 = 0
    for _ in range(m):
        r = orc.run_shot(MODEL, [angles["eps_1q"], angles["eps_2q"], p_xx], "single", rnd=rnd, meas=so.SampledMeasure(mr))
        fails += r["fail"]
        not0 += r["not0"]
    for got, ref in ((c.failed / n, fails / m), (c.not0 / n, not0 / m)):
        p = got
        assert abs(got - ref) < 3.5 * np.sqrt(p * (1 - p) / m) + 1e-9, (got, ref)
    assert 0.01 < c.not0 / n < 0.99


def test_coherent_config3_through_the_drop_in(error_models, correction_table, bitflip_table):
    """BASELINE config 3 through the public entry point: 10^5 shots, eps = 0.02 rad, p_XX = 5e-3, on the coherent form of
    the whole-cycle kernel (at most two sweeps per shot)."""
    import surface_17_iqt_b200.calc_log_e_rate_arbitrary_error_model as drv
    drv.VERBOSE = False
    p_set = [[0.02] * 30, [0.02] * 24, [5e-3] * 24]
    rate, failed, counted, not0 = drv.calculate_log_e_rate_true_probability(
        100000, correction_table, MODEL, p_set, bitflip_table, protocol="s1",
        over_angles={"eps_1q": [0.02] * 30, "eps_2q": [0.02] * 24})
    assert counted + not0 == 100000 and 0 < not0 < 50000 and failed <= counted and rate == failed / counted
    sim = next(v for k, v in drv._SIMS.items() if k[0] == MODEL)
    assert [i["coherent"] for i in sim.backend.program_info] == [False, True, True]
    drv.release()
