"""Python face of one libs17.so context: a batch of Surface-17 shots resident on one B200.

Replaces what the reference builds per shot -- ``MainEngine(Simulator())`` plus two
``allocate_qureg`` calls (calc_log_e_rate_arbitrary_error_model.py:299-303) and a re-parse of
``error_models.json`` (compiled_surface_code_arbitrary_error_model.py:31-49) -- with one device
context that is programmed once and then runs whole shot protocols per call.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from . import circuit

PROTOCOLS = {"single": L.PROTO_SINGLE, "s1": L.PROTO_S1, "s2": L.PROTO_S2, "rtf": L.PROTO_RTF}


def pack_correction_table(table):
    """``{"b0 ... b7": [[x0..x8, z0..z8], weight]}`` -> uint32[256], index bit j = syndrome bit of ancilla j."""
    lut = np.zeros(256, dtype=np.uint32)
    if len(table) != 256:
        raise ValueError("correction table must have 256 entries")
    for key, (vec, _w) in table.items():
        bits = [int(b) for b in key.split()]
        if len(bits) != 8 or len(vec) != 18:
            raise ValueError("bad correction table entry {!r}".format(key))
        idx = sum(b << j for j, b in enumerate(bits))
        lut[idx] = sum(int(v) << i for i, v in enumerate(vec))
    return lut


def pack_bitflip_table(table):
    """``{"c0 c1 c2 c3": [[...], weight]}`` -> uint8[16] of weights, index = c0 | c1<<1 | c2<<2 | c3<<3."""
    w = np.zeros(16, dtype=np.uint8)
    if len(table) != 16:
        raise ValueError("classical bit-flip table must have 16 entries")
    for key, (_vec, weight) in table.items():
        bits = [int(b) for b in key.split()]
        w[sum(b << j for j, b in enumerate(bits))] = int(weight)
    return w


def describe_program(prog):
    """Which kernels the library's host lowering selects for `prog` (circuit.Program), per cycle kind, and the label
    maps of the whole-cycle kernel.  Needs no GPU (s17_describe_program)."""
    lib = L.load()
    flat = [lay for cyc in prog.layers for lay in cyc]
    layers = (L.Layer * len(flat))(*flat)
    counts = (C.c_uint32 * 3)(*[len(cyc) for cyc in prog.layers])
    params = (C.c_double * max(1, len(prog.params)))(*prog.params)
    plan = (L.PInstr * len(prog.plan))(*prog.plan)
    info = L.ProgramInfo()
    rc = lib.s17_describe_program(layers, counts, params, len(prog.params) // 2, plan, len(plan), C.byref(info))
    if rc != 0:
        raise L.S17Error(lib.s17_last_error(None).decode())
    out = []
    for c, name in enumerate(("preparation", "stab_ind=0", "stab_ind=1")):
        out.append({
            "cycle": name, "sweeps": int(info.n_sweeps[c]), "k_half": bool(info.half_kernel[c]),
            "k_cycle": bool(info.cycle_kernel[c]), "k_cycle1": bool(info.cycle_kernel_maps[c]),
            "coherent": bool(info.coherent[c]),
            "shuffle_bit": int(info.shuffle_bit[c]),
            "reg_bits_z": [int(v) for v in info.reg_bits_z[c]], "reg_bits_x": [int(v) for v in info.reg_bits_x[c]],
            "table_entries": [int(v) for v in info.table_entries[c]],
            "slot_ancilla": [[int(v) for v in info.slot_ancilla[c][h]] for h in range(2)],
            # (stab_ind=0 entry) the preparation cycle lowers to the same whole-cycle program: an un-memoised run
            # computes it inside the launch of the first noisy cycle
            "prep_rides_along": bool(info.prep_rides_in_cycle1) if c == 1 else False})
    return out


def bits_to_list(value, n):
    return [(int(value) >> i) & 1 for i in range(n)]


class Backend:
    def __init__(self, max_shots, device=0):
        self._lib = L.load()
        self._h = C.c_void_p()
        self.max_shots = int(max_shots)
        self.device = int(device)
        rc = self._lib.s17_create(C.byref(self._h), self.device, self.max_shots)
        if rc != 0:
            msg = self._lib.s17_last_error(None).decode()
            self._h = None
            raise L.S17Error(msg)
        self.program = None
        self.last_counts = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.s17_destroy(self._h)
            self._h = None

    __del__ = close

    def _ck(self, rc):
        if rc != 0:
            raise L.S17Error(self._lib.s17_last_error(self._h).decode())

    # ---- programming ----------------------------------------------------------------------------
    def load_program(self, prog: circuit.Program):
        flat = [lay for cyc in prog.layers for lay in cyc]
        layers = (L.Layer * len(flat))(*flat)
        counts = (C.c_uint32 * 3)(*[len(cyc) for cyc in prog.layers])
        params = (C.c_double * max(1, len(prog.params)))(*prog.params)
        plan = (L.PInstr * len(prog.plan))(*prog.plan)
        lists = L.Lists()
        lists.n_lists = len(prog.list_len)
        for i, n in enumerate(prog.list_len):
            lists.list_len[i] = n
        self._ck(self._lib.s17_load_program(self._h, layers, counts, params, len(prog.params) // 2, plan,
                                            len(plan), C.byref(lists)))
        self.program = prog

    @property
    def program_info(self):
        """Which kernels the lowering selects for the loaded program (describe_program: the same host code path)."""
        return describe_program(self.program)

    def load_tables(self, correction_table, bitflip_table):
        lut = pack_correction_table(correction_table)
        cw = pack_bitflip_table(bitflip_table)
        self._ck(self._lib.s17_load_tables(self._h, lut.ctypes.data_as(C.POINTER(C.c_uint32)),
                                           cw.ctypes.data_as(C.POINTER(C.c_uint8))))

    def set_slot_probs(self, p_set):
        for i, plist in enumerate(p_set):
            arr = np.ascontiguousarray(plist, dtype=np.float64)
            self._ck(self._lib.s17_set_slot_probs(self._h, i, arr.ctypes.data_as(C.POINTER(C.c_double)), len(arr)))

    def set_subset(self, eset, locations, weights=None):
        n = len(eset)
        e = (C.c_uint32 * n)(*[int(v) for v in eset])
        loc = (C.c_uint32 * n)(*[int(v) for v in locations])
        keep, ptrs = [], (C.POINTER(C.c_double) * n)()
        for i in range(n):
            wi = None if weights is None else weights[i]
            if wi is None or len(wi) == 1:
                ptrs[i] = None
            else:
                if len(wi) != locations[i]:
                    raise ValueError("weights for error type {} must have one entry per location".format(i))
                arr = np.ascontiguousarray(wi, dtype=np.float64)
                keep.append(arr)
                ptrs[i] = arr.ctypes.data_as(C.POINTER(C.c_double))
        self._ck(self._lib.s17_set_subset(self._h, n, e, loc, ptrs))

    def set_replay(self, masks=None, outcomes=None):
        """masks: [n_shots][total_slots] 0/1 (lists concatenated); outcomes: [n_shots][<=40] bits in program order."""
        n = len(masks) if masks is not None else len(outcomes)
        m = o = None
        ms = os_ = 0
        if masks is not None:
            m = np.ascontiguousarray(masks, dtype=np.uint8)
            ms = m.shape[1]
        if outcomes is not None:
            width = max(len(r) for r in outcomes)
            o = np.zeros((n, width), dtype=np.uint8)
            for i, r in enumerate(outcomes):
                o[i, :len(r)] = r
            os_ = width
        self._ck(self._lib.s17_set_replay(self._h, n, None if m is None else m.ctypes.data, ms,
                                          None if o is None else o.ctypes.data, os_))

    # ---- execution ------------------------------------------------------------------------------
    def _args(self, protocol, n_shots, noise, seed=0, first_shot_id=0, logical_state=0, materialize=False,
              stop_stage=L.STOP_NONE, stop_layers=0, forced=False, rtf_runs=0, rtf_max_rounds=0, leak_reset="run",
              rtf_poll=0, rtf_round_budget=0, rtf_chain_runs=0):
        a = L.RunArgs()
        a.protocol = PROTOCOLS[protocol] if isinstance(protocol, str) else int(protocol)
        a.noise = int(noise)
        a.first_shot_id, a.n_shots, a.seed = int(first_shot_id), int(n_shots), int(seed) & (2 ** 64 - 1)
        a.logical_state, a.basis_x = int(logical_state), 0
        a.materialize, a.stop_stage, a.stop_layers = int(materialize), int(stop_stage), int(stop_layers)
        a.forced, a.rtf_runs, a.rtf_max_rounds = int(forced), int(rtf_runs), int(rtf_max_rounds)
        a.leak_reset = L.LEAK_RESET[leak_reset] if isinstance(leak_reset, str) else int(leak_reset)
        a.rtf_poll, a.rtf_round_budget, a.rtf_chain_runs = int(rtf_poll), int(rtf_round_budget), int(rtf_chain_runs)
        return a

    def run(self, protocol, n_shots, noise, **kw):
        """One batch of shots of `protocol` ('single' | 's1' | 's2' | 'rtf'); keyword arguments as in `_args`."""
        a = self._args(protocol, n_shots, noise, **kw)
        c = L.Counts()
        self._ck(self._lib.s17_run(self._h, C.byref(a), C.byref(c)))
        self.last_counts = c
        return c

    # run-til-fail as a job (s17_rtf_begin / _step / _end): lets a caller look at every round
    def rtf_begin(self, n_slots, noise, **kw):
        a = self._args("rtf", n_slots, noise, **kw)
        self._ck(self._lib.s17_rtf_begin(self._h, C.byref(a)))

    def rtf_step(self, n_rounds=1):
        alive, rounds = C.c_uint64(), C.c_uint64()
        self._ck(self._lib.s17_rtf_step(self._h, int(n_rounds), C.byref(alive), C.byref(rounds)))
        return alive.value, rounds.value

    def rtf_end(self):
        c = L.Counts()
        self._ck(self._lib.s17_rtf_end(self._h, C.byref(c)))
        self.last_counts = c
        return c

    def rtf_slots(self, n):
        """[n][4] uint32: run id, rounds so far, second syndromes so far, alive."""
        buf = np.zeros((n, 4), dtype=np.uint32)
        self._ck(self._lib.s17_read_rtf_slots(self._h, buf.ctypes.data_as(C.POINTER(C.c_uint32)), n))
        return buf

    def timing(self):
        g, m, o, sp = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        ng, na = C.c_uint64(), C.c_uint64()
        self._ck(self._lib.s17_last_timing(self._h, C.byref(g), C.byref(m), C.byref(o), C.byref(ng), C.byref(na),
                                           C.byref(sp)))
        ms = (C.c_double * len(L.T_CLASSES))()
        nl = (C.c_uint64 * len(L.T_CLASSES))()
        un = (C.c_uint64 * len(L.T_CLASSES))()
        self._ck(self._lib.s17_last_timing_detail(self._h, ms, nl, un))
        return {"gate_ms": g.value, "measure_ms": m.value, "other_ms": o.value, "gate_launches": ng.value,
                "launches": na.value, "span_ms": sp.value,
                "classes": {k: {"ms": ms[i], "launches": int(nl[i]), "units": int(un[i])} for i, k in enumerate(L.T_CLASSES)}}

    def records(self, n):
        buf = (L.Record * n)()
        self._ck(self._lib.s17_read_records(self._h, buf, n))
        out = []
        for r in buf:
            f = r.flags
            d = {"quiescent": bits_to_list(r.quiescent, 8), "synd1": bits_to_list(r.synd1, 8),
                 "flips_a": bits_to_list(r.flips_a, 8), "not0": int(bool(f & L.F_NOT0)),
                 "counted": int(bool(f & L.F_COUNTED)), "fail": int(bool(f & L.F_FAIL)),
                 "error": int(bool(f & L.F_ERROR)), "leaked": bits_to_list(r.leak, 17),
                 "outcomes": [int(b) for b in r.outcomes[:r.n_outcomes]]}
            if f & L.F_SYND2:
                d["synd2"] = bits_to_list(r.synd2, 8)
            if f & L.F_COUNTED:
                d["ft_syndrome"] = bits_to_list(r.ft_syndrome, 8)
                d["correction"] = bits_to_list(r.correction, 18)
                d["data_meas"] = bits_to_list(r.data_meas, 9)
                d["logical_meas"] = int(r.logical_meas)
            out.append(d)
        return out

    def masks(self, n):
        """Fired/placed slot bits per shot: [n][2][total_slots] (second index: cycle-1 / cycle-2 record)."""
        buf = np.zeros((n, 2 * L.MASK_WORDS), dtype=np.uint32)
        self._ck(self._lib.s17_read_masks(self._h, buf.ctypes.data_as(C.POINTER(C.c_uint32)), n))
        total = sum(self.program.list_len)
        bits = ((buf[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).astype(np.uint8)
        bits = bits.reshape(n, 2, L.MASK_WORDS * 32)
        return bits[:, :, :total]

    def masks_packed(self, n):
        """The same as bit sets: uint32 [n][2 * MASK_WORDS], bit b of word w of a half = slot 32 w + b."""
        buf = np.zeros((n, 2 * L.MASK_WORDS), dtype=np.uint32)
        self._ck(self._lib.s17_read_masks(self._h, buf.ctypes.data_as(C.POINTER(C.c_uint32)), n))
        return buf

    def events(self, n):
        buf = np.zeros((n, L.EV_WORDS), dtype=np.uint32)
        self._ck(self._lib.s17_read_events(self._h, buf.ctypes.data_as(C.POINTER(C.c_uint32)), n))
        return buf

    def split_mask(self, flat):
        out, off = [], 0
        for n in self.program.list_len:
            out.append([int(b) for b in flat[off:off + n]])
            off += n
        return out

    def state(self, shot):
        buf = np.zeros(2 * L.DIM, dtype=np.float64)
        self._ck(self._lib.s17_read_state(self._h, int(shot), buf.ctypes.data_as(C.POINTER(C.c_double))))
        return buf.view(np.complex128)

    def rtf_results(self, n_runs):
        r = np.zeros(n_runs, dtype=np.uint32)
        b = np.zeros(n_runs, dtype=np.uint32)
        self._ck(self._lib.s17_read_rtf(self._h, r.ctypes.data_as(C.POINTER(C.c_uint32)),
                                        b.ctypes.data_as(C.POINTER(C.c_uint32)), n_runs))
        return r, b
