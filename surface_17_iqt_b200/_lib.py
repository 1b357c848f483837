"""ctypes binding of libs17.so (include/s17.h).  Fails loudly when the CUDA library is missing:
there is no CPU fallback in this package."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("S17_LIB") or os.path.join(_HERE, "libs17.so")

N_QUBITS = 17
DIM = 1 << 17
EV_WORDS = 32
MASK_WORDS = 32
MAX_LISTS = 8
MAX_OUTCOMES = 40

U_RY, U_RXX, U_RX, U_H, U_ROT_X, U_ROT_Y, U_ROT_XX = 1, 2, 3, 4, 5, 6, 7
OP_ENT, OP_IDLE = 0, 1
P_OP_BEGIN, P_ERR, P_OP_END, P_IDLE, P_SKIP_TEST = 1, 2, 3, 4, 5
(ERR_X, ERR_Y, ERR_Z, ERR_ZC, ERR_ZT, ERR_XX, ERR_ZZ, ERR_LEAK_A, ERR_LEAK_AB, ERR_DEPOL) = range(1, 11)
PROTO_SINGLE, PROTO_S1, PROTO_S2, PROTO_RTF = 0, 1, 2, 3
NOISE_PROB, NOISE_SUBSET, NOISE_REPLAY = 0, 1, 2
STOP_NONE, STOP_AFTER_PREP, STOP_CYCLE1_GATES, STOP_AFTER_CYCLE1, STOP_AFTER_CORRECTION = 0, 1, 2, 3, 4
F_NOT0, F_COUNTED, F_FAIL, F_ERROR, F_SYND2 = 1, 2, 4, 8, 16
LEAK_RESET = {"never": 0, "run": 1, "round": 2}
MEAS_SAMPLE, MEAS_LOCKSTEP, MEAS_FORCED = 0, 1, 2


class Uop(C.Structure):
    _fields_ = [("type", C.c_uint8), ("sign", C.c_int8), ("skippable", C.c_uint8), ("ev_shift", C.c_uint8),
                ("param", C.c_uint8), ("pad", C.c_uint8 * 3)]


class Op(C.Structure):
    _fields_ = [("kind", C.c_uint8), ("data_bit", C.c_uint8), ("anc_bit", C.c_uint8), ("n_uops", C.c_uint8),
                ("ev_word", C.c_uint8), ("partial", C.c_uint8), ("pad", C.c_uint8 * 2), ("uops", Uop * 8)]


class Layer(C.Structure):
    _fields_ = [("n_ops", C.c_uint8), ("n_anc", C.c_uint8), ("anc_bits", C.c_uint8 * 4), ("meas_mask", C.c_uint8),
                ("pad", C.c_uint8), ("ops", Op * 16)]


class PInstr(C.Structure):
    _fields_ = [("kind", C.c_uint8), ("err", C.c_uint8), ("list", C.c_uint8), ("field", C.c_uint8),
                ("slot", C.c_uint16), ("a", C.c_uint8), ("b", C.c_uint8), ("word", C.c_uint8),
                ("in_skip", C.c_uint8), ("use_round", C.c_uint8), ("pad", C.c_uint8)]


class Lists(C.Structure):
    _fields_ = [("n_lists", C.c_uint32), ("list_len", C.c_uint32 * MAX_LISTS)]


class RunArgs(C.Structure):
    _fields_ = [("protocol", C.c_uint32), ("noise", C.c_uint32), ("first_shot_id", C.c_uint64),
                ("n_shots", C.c_uint64), ("seed", C.c_uint64), ("logical_state", C.c_uint32),
                ("basis_x", C.c_uint32), ("materialize", C.c_uint32), ("stop_stage", C.c_uint32),
                ("stop_layers", C.c_uint32), ("forced", C.c_uint32), ("rtf_runs", C.c_uint32),
                ("rtf_max_rounds", C.c_uint32), ("leak_reset", C.c_uint32), ("rtf_poll", C.c_uint32),
                ("rtf_round_budget", C.c_uint64), ("rtf_chain_runs", C.c_uint32), ("pad0", C.c_uint32)]


class Counts(C.Structure):
    _fields_ = [("shots", C.c_uint64), ("failed", C.c_uint64), ("not0", C.c_uint64), ("counted", C.c_uint64),
                ("errors", C.c_uint64), ("sweeps", C.c_uint64), ("rounds", C.c_uint64), ("sweep_bytes", C.c_uint64)]


class Record(C.Structure):
    _fields_ = [("quiescent", C.c_uint8), ("synd1", C.c_uint8), ("synd2", C.c_uint8), ("flips_a", C.c_uint8),
                ("ft_syndrome", C.c_uint8), ("flags", C.c_uint8), ("logical_meas", C.c_uint8),
                ("n_outcomes", C.c_uint8), ("correction", C.c_uint32), ("leak", C.c_uint32),
                ("data_meas", C.c_uint16), ("partial", C.c_uint16), ("outcomes", C.c_uint8 * MAX_OUTCOMES)]


class ProgramInfo(C.Structure):
    _fields_ = [("n_sweeps", C.c_uint32 * 3), ("half_kernel", C.c_uint8 * 3), ("cycle_kernel", C.c_uint8 * 3),
                ("cycle_kernel_maps", C.c_uint8 * 3), ("shuffle_bit", C.c_uint8 * 3), ("reg_bits_z", (C.c_uint8 * 4) * 3),
                ("reg_bits_x", (C.c_uint8 * 4) * 3), ("table_entries", (C.c_uint8 * 2) * 3),
                ("slot_ancilla", ((C.c_uint8 * 4) * 2) * 3), ("coherent", C.c_uint8 * 3),
                ("prep_rides_in_cycle1", C.c_uint8)]


assert C.sizeof(Uop) == 8 and C.sizeof(Op) == 72 and C.sizeof(Layer) == 1160
assert C.sizeof(PInstr) == 12 and C.sizeof(Record) == 60 and C.sizeof(RunArgs) == 88

EXPORTS = ("s17_create", "s17_destroy", "s17_last_error", "s17_device_count", "s17_load_program", "s17_describe_program",
           "s17_diag_slot_entry",
           "s17_load_tables", "s17_set_slot_probs", "s17_set_subset", "s17_set_replay", "s17_run",
           "s17_read_records", "s17_read_masks", "s17_read_events", "s17_read_state", "s17_read_rtf", "s17_last_timing",
           "s17_last_timing_detail", "s17_allreduce_counts",
           "s17_rtf_begin", "s17_rtf_step", "s17_rtf_end", "s17_read_rtf_slots",
           "s17_eng_create", "s17_eng_destroy", "s17_eng_last_error", "s17_eng_reset", "s17_eng_gate1", "s17_eng_gate2",
           "s17_eng_measure", "s17_eng_prob1", "s17_eng_flush", "s17_eng_read_state", "s17_eng_timing")

# kernel classes of s17_last_timing_detail (include/s17.h: S17_T_*)
T_CLASSES = ("collapse", "sample", "plan", "compact", "record", "readout", "place", "reserved")

_LIB = None


class S17Error(RuntimeError):
    pass


def load():
    """Load libs17.so; raise if the CUDA extension has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise S17Error("libs17.so is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(surface_17_iqt_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    L.s17_create.argtypes = [C.POINTER(vp), i32, u64]
    L.s17_destroy.argtypes = [vp]
    L.s17_last_error.argtypes = [vp]
    L.s17_last_error.restype = C.c_char_p
    L.s17_device_count.argtypes = []
    L.s17_load_program.argtypes = [vp, C.POINTER(Layer), C.POINTER(u32), C.POINTER(C.c_double), u32,
                                   C.POINTER(PInstr), u32, C.POINTER(Lists)]
    L.s17_describe_program.argtypes = [C.POINTER(Layer), C.POINTER(u32), C.POINTER(C.c_double), u32, C.POINTER(PInstr), u32,
                                       C.POINTER(ProgramInfo)]
    L.s17_diag_slot_entry.argtypes = [u32, u32, u32, u32, u32, u32, C.POINTER(C.c_int32)]
    L.s17_load_tables.argtypes = [vp, C.POINTER(u32), C.POINTER(C.c_uint8)]
    L.s17_set_slot_probs.argtypes = [vp, u32, C.POINTER(C.c_double), u32]
    L.s17_set_subset.argtypes = [vp, u32, C.POINTER(u32), C.POINTER(u32), C.POINTER(C.POINTER(C.c_double))]
    L.s17_set_replay.argtypes = [vp, u64, C.c_void_p, u64, C.c_void_p, u64]
    L.s17_run.argtypes = [vp, C.POINTER(RunArgs), C.POINTER(Counts)]
    L.s17_read_records.argtypes = [vp, C.POINTER(Record), u64]
    L.s17_read_masks.argtypes = [vp, C.POINTER(u32), u64]
    L.s17_read_events.argtypes = [vp, C.POINTER(u32), u64]
    L.s17_read_state.argtypes = [vp, u64, C.POINTER(C.c_double)]
    L.s17_read_rtf.argtypes = [vp, C.POINTER(u32), C.POINTER(u32), u64]
    L.s17_last_timing.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                  C.POINTER(u64), C.POINTER(u64), C.POINTER(C.c_double)]
    L.s17_last_timing_detail.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(u64), C.POINTER(u64)]
    L.s17_allreduce_counts.argtypes = [C.POINTER(Counts), C.c_int, C.POINTER(Counts)]
    L.s17_rtf_begin.argtypes = [vp, C.POINTER(RunArgs)]
    L.s17_rtf_step.argtypes = [vp, u32, C.POINTER(u64), C.POINTER(u64)]
    L.s17_rtf_end.argtypes = [vp, C.POINTER(Counts)]
    L.s17_read_rtf_slots.argtypes = [vp, C.POINTER(u32), u64]
    dp, bp = C.POINTER(C.c_double), C.c_void_p
    L.s17_eng_create.argtypes = [C.POINTER(vp), i32, u32, u64, u64]
    L.s17_eng_destroy.argtypes = [vp]
    L.s17_eng_last_error.argtypes = [vp]
    L.s17_eng_last_error.restype = C.c_char_p
    L.s17_eng_reset.argtypes = [vp]
    L.s17_eng_gate1.argtypes = [vp, u32, dp, bp]
    L.s17_eng_gate2.argtypes = [vp, u32, u32, dp, bp]
    L.s17_eng_measure.argtypes = [vp, u32, u32, bp, bp, dp, C.POINTER(u32)]
    L.s17_eng_prob1.argtypes = [vp, u32, dp]
    L.s17_eng_flush.argtypes = [vp]
    L.s17_eng_read_state.argtypes = [vp, u64, dp]
    L.s17_eng_timing.argtypes = [vp, dp, C.POINTER(u64), i32]
    for name in EXPORTS:
        if name not in ("s17_last_error", "s17_eng_last_error"):
            getattr(L, name).restype = i32
    _LIB = L
    return L
