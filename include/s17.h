/*
 * s17.h -- C ABI of libs17.so, the B200 (sm_100a) batched Surface-17 trajectory backend.
 *
 * This is the drop-in boundary for the reference's hot path.  Citations are into the reference
 * repository (Alexowens94/Surface_17_IQT):
 *   CS = compiled_surface_code_arbitrary_error_model.py, CL = calc_log_e_rate_arbitrary_error_model.py
 *
 * In the reference the boundary is the ProjectQ engine object: `MainEngine(Simulator())`
 * (CL:63, 221, 299, 392), `eng.allocate_qureg` (CL:302-303), `Gate | qubits` (every gate line of
 * CS:291-344), `All(Measure) | ancilla; eng.flush(); int(q)` (CS:848-851) -- one pybind call per
 * gate per shot.  Here the boundary is one call per BATCH of shots: the host compiles the cycle
 * and the error model once (s17_load_program), then s17_run executes whole shot protocols
 * (CL:183 single, CL:261 s1, CL:355 s2, CL:33 run-til-fail) on device and returns the counters
 * those functions return.
 *
 * Conventions: every function returns 0 on success, a negative S17_E_* code otherwise;
 * s17_last_error gives the message.  No C++ types, exceptions or torch types cross the line.  The
 * caller owns every buffer it passes; the library owns device memory behind the opaque handle.
 * A context is bound to one CUDA device and is not thread-safe.  All entry points are synchronous.
 * There is no CPU fallback: s17_create fails if no CUDA device is usable.
 */
#ifndef S17_H
#define S17_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S17_N_QUBITS 17u
#define S17_DIM (1u << 17)
#define S17_EV_WORDS 32u       /* per-shot event words per cycle: 24 op words + 7 idle masks + 1 spare */
#define S17_MASK_WORDS 32u     /* per-shot error-slot bitset (<= 1024 slots over both rounds) */
#define S17_MAX_LISTS 8u
#define S17_MAX_OUTCOMES 40u   /* 3 x 8 ancilla + 9 data, padded */

enum { S17_OK = 0, S17_E_ARG = -1, S17_E_CUDA = -2, S17_E_STATE = -3, S17_E_NOMEM = -4 };

/* ---- gate stream ------------------------------------------------------------------------- */

/* micro-ops executed on the 4 register-resident amplitudes of one (data, ancilla) pair */
enum {
    S17_U_RY = 1,      /* Ry(sign*pi/2) on the data bit          (CS:315, 339) */
    S17_U_RXX = 2,     /* Rxx(sign*pi/2) on (data, ancilla)      (CS:294, 323) */
    S17_U_RX = 3,      /* Rx(sign*pi/2) on the data bit          (CS:304, 333) */
    S17_U_H = 4,       /* Hadamard on the data bit               (CS:54, 106)  */
    S17_U_ROT_X = 5,   /* Rx(theta), cos/sin from the parameter table (coherent over-rotation) */
    S17_U_ROT_Y = 6,   /* Ry(theta) */
    S17_U_ROT_XX = 7   /* Rxx(theta) */
};

typedef struct {
    uint8_t type;       /* S17_U_* */
    int8_t sign;        /* +1 / -1 for the +-pi/2 forms */
    uint8_t skippable;  /* 1: not executed when the op's leak predicate holds (CS:293, 322) */
    uint8_t ev_shift;   /* bit offset of the 6-bit Pauli event field applied AFTER this uop; 255 = none */
    uint8_t param;      /* index into the (cos, sin) parameter table for S17_U_ROT_* */
    uint8_t pad[3];
} s17_uop;

enum { S17_OP_ENT = 0, S17_OP_IDLE = 1 };

/* one entangling operation (x_type_entangling CS:291 / z_type_entangling CS:312), or an idle z-mask */
typedef struct {
    uint8_t kind;       /* S17_OP_ENT or S17_OP_IDLE */
    uint8_t data_bit;   /* state-vector bit of the data qubit (0..8) */
    uint8_t anc_bit;    /* state-vector bit of the ancilla qubit (9..16) */
    uint8_t n_uops;
    uint8_t ev_word;    /* index of this op's per-shot event word */
    uint8_t partial;    /* 1: holds only some gates of its entangling operation (one-gate-per-sweep mode): Pauli events
                           must be applied gate by gate, not as the operation's net end-of-op Pauli */
    uint8_t pad[2];
    s17_uop uops[8];
} s17_op;

/* one sweep over the state: every op's ancilla bit must be one of anc_bits */
#define S17_MAX_LAYER_OPS 16u
typedef struct {
    uint8_t n_ops;
    uint8_t n_anc;         /* 3: 4096-amplitude tiles (<= 8 ops), 4: 8192-amplitude tiles (<= 16 ops) */
    uint8_t anc_bits[4];   /* ancilla bits staged in shared memory next to the 9 data bits, ascending */
    uint8_t meas_mask;     /* != 0: after this sweep the ancillas j with bit j set are measured (the sweep also emits the
                              joint ancilla histogram).  0xFF on the last sweep = All(Measure) | ancilla (CS:848); a
                              partial mask measures qubits no later gate touches earlier, which commutes */
    uint8_t pad;
    s17_op ops[S17_MAX_LAYER_OPS];
} s17_layer;

/* ---- error-injection program (restates insert_errors / insert_error, CS:127-234) ---------------- */

enum { S17_P_OP_BEGIN = 1, S17_P_ERR = 2, S17_P_OP_END = 3, S17_P_IDLE = 4,
       S17_P_SKIP_TEST = 5 /* evaluate `leaked[a] or leaked[b]` right before the Rxx (CS:293, 322) */ };
enum {
    S17_ERR_X = 1, S17_ERR_Y = 2, S17_ERR_Z = 3,   /* on qubits[0]            (CS:201-206) */
    S17_ERR_ZC = 4, S17_ERR_ZT = 5,                /* Z on qubits[0] / [1]    (CS:207-211) */
    S17_ERR_XX = 6, S17_ERR_ZZ = 7,                /* both qubits             (CS:212-217) */
    S17_ERR_LEAK_A = 8,                            /* set leak[a]             (CS:218-219, 225-232) */
    S17_ERR_LEAK_AB = 9,                           /* set leak[a], leak[b]    (CS:220-223) */
    S17_ERR_DEPOL = 10                             /* uniform 15-way 2q Pauli (CS:237-283) */
};

typedef struct {
    uint8_t kind;      /* S17_P_* */
    uint8_t err;       /* S17_ERR_* */
    uint8_t list;      /* probability-list id (position in the model's "probabilities") */
    uint8_t field;     /* event-field bit offset inside the op word (0, 6, 12, 18) */
    uint16_t slot;     /* slot index inside ONE round of the list */
    uint8_t a, b;      /* SKIP_TEST: leak predicate indices; ERR leak kinds: indices to set; IDLE: a = qubit */
    uint8_t word;      /* event word index (op word, or 24 + timestep for IDLE) */
    uint8_t in_skip;   /* 1: entry sits inside the `if not leaked` block and is not drawn when skipped */
    uint8_t use_round; /* 1: slot is offset by half the list in the second cycle (CS:191-193); 0 for Idle */
    uint8_t pad;
} s17_pinstr;

typedef struct {
    uint32_t n_lists;
    uint32_t list_len[S17_MAX_LISTS];   /* total length the caller passes per list (1 or 2 rounds) */
} s17_lists;

/* ---- run control ------------------------------------------------------------------------------ */

enum { S17_PROTO_SINGLE = 0, S17_PROTO_S1 = 1, S17_PROTO_S2 = 2, S17_PROTO_RTF = 3 };
enum { S17_NOISE_PROB = 0, S17_NOISE_SUBSET = 1, S17_NOISE_REPLAY = 2 };
enum { S17_STOP_NONE = 0, S17_STOP_AFTER_PREP = 1, S17_STOP_CYCLE1_GATES = 2, S17_STOP_AFTER_CYCLE1 = 3,
       S17_STOP_AFTER_CORRECTION = 4 };

typedef struct {
    uint32_t protocol;        /* S17_PROTO_* */
    uint32_t noise;           /* S17_NOISE_* */
    uint64_t first_shot_id;   /* global id of shot 0 of this batch: Philox counter => results independent of sharding */
    uint64_t n_shots;
    uint64_t seed;
    uint32_t logical_state;   /* 0 / 1  (CS:103-104) */
    uint32_t basis_x;         /* 1: prepare / read out in the X basis (CS:53-54, 105-106) */
    uint32_t materialize;     /* 1: apply collapse + correction to the stored state (parity mode) */
    uint32_t stop_stage;      /* S17_STOP_* (parity/debug) */
    uint32_t stop_layers;     /* with S17_STOP_CYCLE1_GATES: run only this many sweeps of cycle 1 (0 = all) */
    uint32_t forced;          /* 1: measurement outcomes come from s17_set_replay */
    uint32_t rtf_runs;        /* S17_PROTO_RTF: number of runs to complete */
    uint32_t rtf_max_rounds;  /* S17_PROTO_RTF: safety cap on rounds per run (0 = none) */
    uint32_t leak_reset;      /* S17_PROTO_RTF: S17_LEAK_RESET_* -- when the leak register of a slot is cleared */
    uint32_t rtf_poll;        /* S17_PROTO_RTF: rounds between two host reads of the live-slot count (0 = default) */
    uint64_t rtf_round_budget; /* S17_PROTO_RTF: stop once this many rounds were simulated in total (0 = run every run to
                                  its failure); runs still in flight then stay censored (0 in s17_read_rtf) */
    uint32_t rtf_chain_runs;  /* S17_PROTO_RTF: 0 = a freed slot takes the next unassigned run id (in slot order);
                                  K > 0 = every slot is a chain of exactly K consecutive runs, ids slot*K .. slot*K+K-1
                                  (rtf_runs must be n_shots*K): with S17_LEAK_RESET_NEVER a slot then IS one call of the
                                  reference's calculate_log_e_rate(K, ...) (CL:33-148) */
    uint32_t pad0;
} s17_run_args;
/* leaked_q_reg of calculate_log_e_rate is created once for all runs (CL:49): NEVER is the reference's literal behaviour
 * (per slot here: a slot is one sequential chain of runs); RUN clears it when a slot starts a new run, ROUND at every
 * round.  The library's callers default to RUN (SURVEY C-6). */
enum { S17_LEAK_RESET_NEVER = 0, S17_LEAK_RESET_RUN = 1, S17_LEAK_RESET_ROUND = 2 };

typedef struct {
    uint64_t shots;        /* shots executed */
    uint64_t failed;       /* failed_count        (CL:330, 433, 251) */
    uint64_t not0;         /* flipsa_not0_count   (CL:333, 417) */
    uint64_t counted;      /* shots that reached decode + read-out */
    uint64_t errors;       /* shots with a forced outcome of zero probability */
    uint64_t sweeps;       /* (shot, sweep) pairs executed by the gate kernel */
    uint64_t rounds;       /* S17_PROTO_RTF: total rounds simulated */
    uint64_t sweep_bytes;  /* algorithmic HBM bytes the gate sweeps moved (8 KiB runs read + written) */
} s17_counts;

typedef struct {
    uint8_t quiescent, synd1, synd2, flips_a;  /* bit j = ancilla j */
    uint8_t ft_syndrome, flags, logical_meas, n_outcomes;
    uint32_t correction;      /* bits 0-8: X on data i; bits 9-17: Z on data i (CS:901-907) */
    uint32_t leak;            /* 17 leak flags (CS:286-288) */
    uint16_t data_meas;       /* after the leak override (CS:61-63) */
    uint16_t partial;         /* scratch: outcome bits of a cycle measured so far */
    uint8_t outcomes[S17_MAX_OUTCOMES];  /* raw measurement outcomes in program order */
} s17_record;
enum { S17_F_NOT0 = 1, S17_F_COUNTED = 2, S17_F_FAIL = 4, S17_F_ERROR = 8, S17_F_SYND2 = 16 };

typedef struct s17_ctx s17_ctx;

/* replaces MainEngine(Simulator()) + allocate_qureg(9)/(8) for a whole batch (CL:299-303) */
int s17_create(s17_ctx **out, int device, uint64_t max_shots);
int s17_destroy(s17_ctx *ctx);
const char *s17_last_error(const s17_ctx *ctx);
int s17_device_count(void);

/* the compiled cycle: what stabilizer_cycle_error_index (CS:773-845) emits, grouped into sweeps.
 * `layers` concatenates three sweep lists of n_layers[0..2] entries: the noiseless preparation cycle
 * (logical_prep, CS:96-123), the first noisy cycle (stab_ind=0) and the second one (stab_ind=1). */
int s17_load_program(s17_ctx *ctx, const s17_layer *layers, const uint32_t n_layers[3],
                     const double *params_cos_sin, uint32_t n_params,
                     const s17_pinstr *plan, uint32_t n_plan, const s17_lists *lists);
/* Which kernels the host lowering selects for a compiled program, per cycle kind (0 preparation, 1 stab_ind=0,
 * 2 stab_ind=1), and the label maps of the whole-cycle kernel.  Host-only (no device needed): the same code path
 * s17_load_program runs before it uploads anything. */
typedef struct {
    uint32_t n_sweeps[3];
    uint8_t half_kernel[3];         /* 1: the fused half-cycle kernel (k_half) applies */
    uint8_t cycle_kernel[3];        /* 1: the cycle has the Clifford diagonal-frame form (tables of Gaussian integers) */
    uint8_t cycle_kernel_maps[3];   /* 1: label maps with one register bit per ancilla slot exist (k_cycle1) */
    uint8_t shuffle_bit[3];         /* data bit that is a lane bit under both label maps: transformed between lanes (in
                                       the exchange load of the Walsh-Hadamard transform; a shuffle stage before round 2) */
    uint8_t reg_bits_z[3][4];       /* data bits carried by the register index in the computational frame; slot k uses [k] */
    uint8_t reg_bits_x[3][4];       /* ... in the Hadamard frame */
    uint8_t table_entries[3][2];    /* table entries per half cycle */
    uint8_t slot_ancilla[3][2][4];  /* ancilla index measured by slot k of half h */
    uint8_t coherent[3];            /* 1: the cycle carries coherent over-rotations and runs on the whole-cycle kernel's
                                       coherent form (its Clifford skeleton fixed the maps above) */
    uint8_t prep_rides_in_cycle1;   /* 1: the preparation cycle and the first noisy cycle lower to the same whole-cycle
                                       program, so an un-memoised run computes both in one launch per shot */
} s17_program_info;
int s17_describe_program(const s17_layer *layers, const uint32_t n_layers[3], const double *params_cos_sin,
                         uint32_t n_params, const s17_pinstr *plan, uint32_t n_plan, s17_program_info *out);
/* Host-side check hook (no device): the closed form the whole-cycle kernel uses for one private table entry of an ancilla
 * slot (s17_diag.cuh: dg_slot_entry) -- the amplitudes (A0, A1) the ancilla ends in for outcome 0 / 1 after the operations
 * of the slot (x_type_entangling / z_type_entangling, CS:291-344, with the events insert_error put there, CS:188-234).
 * Every mask has operation j at bit 8 j: W = event word >> 24 (bit 7 data flip seen by the operation, 4 / 3 net Z / X on the
 * ancilla after it, 0 leak skip), f1 = Rxx sign is -1, f2 = an Rx follows, idx = table index bits, cz = idle Z after the
 * operation, mm = operations in use.  out = {Re A0, Im A0, Re A1, Im A1} (Gaussian integers). */
int s17_diag_slot_entry(uint32_t W, uint32_t f1, uint32_t f2, uint32_t idx, uint32_t cz, uint32_t mm, int32_t out[4]);
/* decoder tables: lookup (CS:886-892) and return_weight_classical_lookup (CS:895-898) */
int s17_load_tables(s17_ctx *ctx, const uint32_t *lut256, const uint8_t *classical_weight16);
/* true-probability mode: the lists instantiate_error_model binds (CS:31-49) */
int s17_set_slot_probs(s17_ctx *ctx, uint32_t list_id, const double *p, uint32_t len);
/* fixed-weight mode: generate_error_locations[_non_uniform] (CL:151-181); weights[t] NULL => uniform */
int s17_set_subset(s17_ctx *ctx, uint32_t n_types, const uint32_t *eset, const uint32_t *n_loc,
                   const double *const *weights);
/* replay mode: explicit masks (one byte per slot, lists concatenated) and forced outcomes per shot */
int s17_set_replay(s17_ctx *ctx, uint64_t n_shots, const uint8_t *masks, uint64_t mask_stride,
                   const uint8_t *outcomes, uint64_t outcome_stride);
/* the shot loops of CL:183 / 261 / 355 / 33 */
int s17_run(s17_ctx *ctx, const s17_run_args *args, s17_counts *out);
/* Sum of the counters of several contexts of one process (one context per GPU per host thread).  Across processes
 * (one per GPU under torchrun, the layout bench.py and the drop-in drivers use) the same sum is one all_reduce of these
 * integers over NCCL: surface_17_iqt_b200/distributed.py. */
int s17_allreduce_counts(const s17_counts *per_ctx, int n_ctx, s17_counts *out);
int s17_read_records(s17_ctx *ctx, s17_record *out, uint64_t n);
int s17_read_masks(s17_ctx *ctx, uint32_t *out /* n x 2 x S17_MASK_WORDS: cycle-1 / cycle-2 record */, uint64_t n);
/* per-shot event words of the most recent cycle (debug / parity): n x S17_EV_WORDS */
int s17_read_events(s17_ctx *ctx, uint32_t *out, uint64_t n);
int s17_read_state(s17_ctx *ctx, uint64_t shot, double *re_im /* 2 * 2^17 doubles */);
int s17_read_rtf(s17_ctx *ctx, uint32_t *rounds_til_fail, uint32_t *syndrome_b_count, uint64_t n_runs);
/* run-til-fail (CL:33-148) as a job, for callers that want to look at the slots between rounds (parity tests replay
 * every round through the oracle): s17_run with S17_PROTO_RTF is begin + step until no slot is live + end.  A slot is
 * one run in flight; a slot whose run failed takes the next run id, in slot order.  After a step the records / masks of
 * the last round are readable with s17_read_records / s17_read_masks (cycle-A slots in the first half, cycle-B slots in
 * the second). */
int s17_rtf_begin(s17_ctx *ctx, const s17_run_args *args);
int s17_rtf_step(s17_ctx *ctx, uint32_t n_rounds, uint64_t *alive_slots, uint64_t *rounds_total);
int s17_rtf_end(s17_ctx *ctx, s17_counts *out);
int s17_read_rtf_slots(s17_ctx *ctx, uint32_t *out /* n x {run_id, rounds so far, second syndromes so far, alive} */, uint64_t n);
/* device time of the last s17_run, CUDA events on the launch stream (ms): per kernel class, and the span from
 * the first to the last device operation of the call */
int s17_last_timing(s17_ctx *ctx, double *gate_ms, double *measure_ms, double *other_ms,
                    uint64_t *gate_launches, uint64_t *all_launches, double *span_ms);
/* the same per kernel class: ms[S17_T_N] and launches[S17_T_N], indexed by S17_T_* (the measurement layer of a
 * full-state run is S17_T_SAMPLE + S17_T_COLLAPSE: the probability reduction itself is fused into the gate sweep that
 * precedes the Measure) */
enum { S17_T_COLLAPSE = 0,   /* k_collapse: ProjectQ's collapse + renormalise + reset, materialised */
       S17_T_SAMPLE = 1,     /* k_measure: the Measure gates of one All(Measure) from the joint weights */
       S17_T_PLAN = 2,       /* k_plan: error insertion (insert_errors / insert_error) -> per-operation events; in the
                                whole-cycle path also the previous cycle's bookkeeping, the shot list, the call's
                                reset and the fixed-weight placement (one launch between two cycle kernels) */
       S17_T_COMPACT = 3,    /* (unused since the shot list is built inside k_plan) */
       S17_T_RECORD = 4,     /* k_record: leak override, protocol branch, lookup decode -- only where it runs on its own */
       S17_T_READOUT = 5,    /* k_readout / k_readout_s: logical_measurement */
       S17_T_PLACE = 6,      /* k_reset: per-shot reset (+ generate_error_locations(_non_uniform) for a run-til-fail job) */
       S17_T_N = 8 };
int s17_last_timing_detail(s17_ctx *ctx, double *ms, uint64_t *launches,
                           uint64_t *units /* NULL, or [S17_T_COLLAPSE] = shots whose collapse was materialised */);

/* ---- engine-level face: ProjectQ's Simulator backend for a batch of state vectors -------------------------------
 *
 * The reference drives its backend gate by gate: `MainEngine(Simulator())` (CL:63, 221, 299, 392), `Gate | qubits`
 * (CS:291-344; the Paulis of insert_error CS:198-217), `All(Measure) | qureg; eng.flush(); int(q)` (CS:56-58, 848-851).
 * These entry points are that surface for `batch` state vectors of 2^n_bits complex128 amplitudes at once (one sweep per
 * gate, like ProjectQ's default gate_fusion=False): surface_17_iqt_b200/pq_compat/projectq binds them behind ProjectQ's
 * own class names, so the reference's unmodified circuit functions run on the GPU.  Bit k = k-th allocated qubit.
 * Gate calls are asynchronous (stream order); measure / flush / read_state / timing synchronise. */
typedef struct s17_eng s17_eng;
enum { S17_MEAS_SAMPLE = 0,     /* one Philox uniform per shot against P(1) */
       S17_MEAS_LOCKSTEP = 1,   /* the outcome drawn for shot 0 is imposed on every shot (the batch holds N copies of
                                   one trajectory: `int(qubit)` has one value) */
       S17_MEAS_FORCED = 2 };   /* outcomes given per shot (replays) */
int s17_eng_create(s17_eng **out, int device, uint32_t n_bits, uint64_t batch, uint64_t seed);
int s17_eng_destroy(s17_eng *eng);
const char *s17_eng_last_error(const s17_eng *eng);
int s17_eng_reset(s17_eng *eng);                               /* every shot back to |0...0> */
/* m: row-major complex matrix as (re, im) pairs, 8 / 32 doubles; two-qubit index = bit0 value + 2 * bit1 value.
 * pred: NULL, or one byte per shot (host): the gate is applied to the shots with a non-zero byte only
 * (`if int(a) == 1: X | a`, CS:859-862). */
int s17_eng_gate1(s17_eng *eng, uint32_t bit, const double *m, const uint8_t *pred);
int s17_eng_gate2(s17_eng *eng, uint32_t bit0, uint32_t bit1, const double *m, const uint8_t *pred);
/* Measure | qubit: P(1) reduction, outcome, collapse + renormalise.  out: one byte per shot (host); prob: NULL or the
 * probability each shot's outcome had; n_impossible: shots whose imposed outcome had probability zero. */
int s17_eng_measure(s17_eng *eng, uint32_t bit, uint32_t mode, const uint8_t *forced, uint8_t *out, double *prob,
                    uint32_t *n_impossible);
int s17_eng_prob1(s17_eng *eng, uint32_t bit, double *p1 /* per shot */);   /* ProjectQ's check before deallocation */
int s17_eng_flush(s17_eng *eng);
int s17_eng_read_state(s17_eng *eng, uint64_t shot, double *re_im /* 2 * 2^n_bits doubles */);
/* device time (CUDA events on the launch stream) and launch counts since the last call with clear != 0:
 * [0] gate sweeps, [1] probability reductions, [2] collapses */
int s17_eng_timing(s17_eng *eng, double ms[3], uint64_t calls[3], int clear);

#ifdef __cplusplus
}
#endif
#endif
