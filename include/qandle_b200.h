/* qandle_b200 -- C ABI of the B200-native state-vector engine that replaces QANDLE's hot path.
 *
 * The reference (gstenzel/qandle v0.1.8) is pure Python and has no FFI; the interface this library
 * replaces is its de-facto operator protocol:
 *   - UnsplittedCircuit.forward's gate loop          reference src/qandle/qcircuit.py:163-174
 *   - BuiltParametrizedOperator.get_matrix/forward   reference src/qandle/operators.py:265-298
 *   - BuiltU / BuiltCNOT / BuiltCZ / BuiltSWAP.forward  operators.py:125-126, 557-558, 590-591, 673-674
 *   - AngleEmbeddingBuilt.forward                    embeddings.py:148-164 (== rotations on |0..0>)
 *   - MeasureProbabilityBuilt.forward, MeasureAllAbsolute.forward   measurements.py:120-123, 78-79
 *   - torch's autograd tape over all of the above -> adjoint-state backward here
 *
 * Conventions are the reference's: qubit 0 is the MOST significant bit of the state index
 * (operators.py:549); a state is B x 2^n interleaved complex (re, im) values, sample-major;
 * MeasureProbability yields P(qubit = 0).
 *
 * All entry points are plain C: pointers + sizes, no torch types.  "dev" entry points take CUDA device
 * pointers and a cudaStream_t (passed as void*); they never synchronise the host.  "host" entry points
 * take host pointers and do their own transfers.  Every function returns 0 on success; on failure a
 * non-zero code is returned and qb_last_error() describes it (thread-local).  No entry point has a CPU
 * fallback: without a CUDA device the compute calls fail.
 */
#ifndef QANDLE_B200_H
#define QANDLE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- gate-program IR: int32 [n_gates][4] rows of (opcode | flags, q0, q1, slot) -------------------
 * RX/RY/RZ: q0 = qubit, slot = column of shared_angles (or of batch_angles when QB_FLAG_BATCH is set).
 *           The angle is the final gate angle: remapping(theta) (operators.py:271) or the raw named
 *           input (operators.py:267); the engine halves it internally like the reference.
 * U:        q0 = qubit, slot = index into fixed_mats; the 2x2 is applied as M psi.  (To reproduce the
 *           reference's `state @ kron(.., matrix, ..)`, operators.py:103-104/125-126, pass matrix^T.)
 * CNOT/CZ:  q0 = control, q1 = target.   SWAP: q0, q1. */
#define QB_OP_RX 1
#define QB_OP_RY 2
#define QB_OP_RZ 3
#define QB_OP_U 4
#define QB_OP_CNOT 5
#define QB_OP_CZ 6
#define QB_OP_SWAP 7
#define QB_OP_MASK 0xFF
#define QB_FLAG_BATCH 0x100

#define QB_C64 0  /* float  re/im, angles float  */
#define QB_C128 1 /* double re/im, angles double */

#define QB_MEASURE_STATE 0 /* MeasureState       measurements.py:85-91  */
#define QB_MEASURE_PROBS 1 /* MeasureProbability measurements.py:94-123 */
#define QB_MEASURE_JOINT 2 /* MeasureJointProbability measurements.py:73-82 */

#define QB_INIT_ZERO 0  /* |0...0>  (qcircuit.py:148-149) */
#define QB_INIT_STATE 1 /* caller-provided state already in the state buffer */

#define QB_STEP_SWEEP 0
#define QB_STEP_EXCHANGE 1 /* amplitude sharding: swap the top-g local index bits with the g rank bits */

typedef struct qb_plan qb_plan;

typedef struct qb_plan_opts {
  int32_t tile_bits;    /* log2(amplitudes) staged in shared memory per CTA; 0 = default for dtype */
  int32_t low_bits;     /* lowest index bits always staged (size of one contiguous HBM chunk); 0 = default */
  int32_t fuse;         /* 0 = default (fuse runs of 1-qubit gates on a qubit into one 2x2), -1 = off */
  int32_t n_local;      /* amplitude sharding: log2(amplitudes per rank); 0 = n_qubits (not sharded) */
  int32_t host_only;    /* 1 = build the plan without touching CUDA (introspection / CPU tests) */
  int32_t swap_relabel; /* 0 = default (SWAP is a relabelling of index bits), -1 = move data */
  int32_t final_layout; /* 0 = restore the identity qubit->bit layout at the end, 1 = leave permuted */
  int32_t max_ops_per_sweep; /* 0 = default */
  int32_t staged;       /* 0 = default (register-blocked staged sweep kernels), -1 = generic kernels only */
  int32_t packed;       /* 0 = default (complex64: packed FFMA2 kernel, planar shared memory), -1 = scalar staged kernel */
  int32_t flat;         /* 0 = default (complex64 packed kernel with straight-line "flat" stage bodies), -1 = interpreted stage bodies */
  int32_t narrow_sync;  /* 0 = default (flat stages: warp / sub-CTA named barriers where the data flow allows), -1 = CTA barriers only */
  int32_t exchange_any_bit; /* amplitude sharding: 0 = exchanges swap the rank bits with the TOP local bits (all-to-all over contiguous
                               chunks: NCCL / push exchange), 1 = the planner picks the local bits per exchange (fewer exchanges; needs
                               qb_exchange_p2p_dev) */
  int32_t sweep_search; /* 0 = default (single-GPU plans: search over where each sweep ends, a few sweeps ahead; fewer sweeps and
                           stages), -1 = plain greedy fill */
  int32_t reserved[2];
} qb_plan_opts;

/* Compile a gate program into a plan (fused gate groups, shared-memory sweeps, exchange steps). */
int qb_plan_create(const int32_t* program, int32_t n_gates, int32_t n_qubits, int32_t dtype,
                   const qb_plan_opts* opts, qb_plan** out);
void qb_plan_destroy(qb_plan* plan);

/* Introspection. */
int32_t qb_plan_num_steps(const qb_plan* plan);
int32_t qb_plan_num_sweeps(const qb_plan* plan);
int32_t qb_plan_num_groups(const qb_plan* plan);
int32_t qb_plan_step_type(const qb_plan* plan, int32_t step);
/* final physical bit of logical qubit q (n_qubits entries) */
int qb_plan_final_pos(const qb_plan* plan, int32_t* pos_out);
/* Serialise the plan as int64 words (format documented in qandle_b200/csrc/plan.h); returns the number of
 * words needed; writes at most cap words. */
int64_t qb_plan_dump(const qb_plan* plan, int64_t* buf, int64_t cap);
/* bytes of device scratch the compute entry points need for a batch of B states */
int64_t qb_workspace_bytes(const qb_plan* plan, int64_t batch);
/* algorithmic HBM bytes moved by forward / backward for a batch of B states */
int64_t qb_plan_algorithmic_bytes(const qb_plan* plan, int64_t batch, int32_t backward);
/* number of kernel launches issued by qb_forward_dev / qb_backward_dev */
int32_t qb_plan_num_launches(const qb_plan* plan, int32_t backward, int32_t measure);

/* ---- device-pointer entry points (the torch.library layer binds these) -----------------------------
 * INTERNAL LAYOUT: between the step-level calls below a complex128 state is interleaved (re, im); a complex64 state is
 * "pack-planar": 16-byte units (re[2k], re[2k+1], im[2k], im[2k+1]).  qb_convert_layout_dev converts in place
 * (an involution; no-op for complex128).  The one-call entry points (qb_forward_dev, qb_backward_dev, qb_run_host) take
 * and return ordinary interleaved complex data wherever the caller sees amplitudes: init_state, the MeasureState result,
 * grad_out for MeasureState; the `lambda` left by qb_backward_dev is internal (convert it before reading it as dL/dpsi0*).
 * state:  [batch][2^n_local] complex, updated in place.   shared_angles: [n_shared] real.
 * batch_angles: [batch][n_batch_cols] real (row stride = n_batch_cols).   fixed_mats: [n_mats][2][2] complex.
 * workspace: qb_workspace_bytes(plan, batch) bytes, 256-byte aligned.  rank: this rank's index when
 * amplitude-sharded (0 otherwise). */
int qb_prepare_dev(const qb_plan* plan, int64_t batch, const void* shared_angles, const void* batch_angles,
                   int32_t n_batch_cols, const void* fixed_mats, void* workspace, void* stream);
int qb_convert_layout_dev(const qb_plan* plan, int64_t batch, void* state, void* stream);
int qb_init_zero_dev(const qb_plan* plan, int64_t batch, void* state, int32_t rank, void* stream);
/* run steps [step_begin, step_end) forward (only sweep steps; the caller performs exchange steps) */
int qb_apply_forward_dev(const qb_plan* plan, int32_t step_begin, int32_t step_end, int64_t batch, void* state,
                         void* workspace, int32_t rank, void* stream);
/* probs_out: [batch][n_qubits] real, P(q=0) partial sums of this rank (sum over ranks when sharded) */
int qb_measure_probs_dev(const qb_plan* plan, int64_t batch, const void* state, void* probs_out, void* workspace,
                         int32_t rank, void* stream);
int qb_measure_joint_dev(const qb_plan* plan, int64_t batch, const void* state, void* joint_out, void* stream);
/* adjoint seeds: lambda = dL/dpsi*   */
int qb_seed_probs_dev(const qb_plan* plan, int64_t batch, const void* state, const void* grad_probs, void* lambda,
                      int32_t rank, void* stream);
int qb_seed_joint_dev(const qb_plan* plan, int64_t batch, const void* state, const void* grad_joint, void* lambda,
                      void* stream);
int qb_seed_state_dev(const qb_plan* plan, int64_t batch, const void* grad_state, void* lambda, void* stream);
/* zero the gradient accumulators in the workspace (call once before the first backward sweep) */
int qb_backward_begin_dev(const qb_plan* plan, int64_t batch, void* workspace, void* stream);
/* run steps [step_begin, step_end) in REVERSE order: psi <- G^+ psi, lambda <- G^+ lambda, accumulate <lambda|dG|psi> */
int qb_apply_backward_dev(const qb_plan* plan, int32_t step_begin, int32_t step_end, int64_t batch, void* state,
                          void* lambda, void* workspace, int32_t rank, void* stream);
/* grad_shared: [n_shared], grad_batch: [batch][n_batch_cols] (both are overwritten) */
int qb_finalize_grads_dev(const qb_plan* plan, int64_t batch, const void* shared_angles, const void* batch_angles,
                          int32_t n_batch_cols, const void* fixed_mats, void* workspace, void* grad_shared,
                          int32_t n_shared, void* grad_batch, void* stream);

/* Amplitude sharding: the local index bits exchange step `step` swaps with the rank bits (rank bit j <-> local bit pos_out[j]);
 * returns their number g = log2(world), or -1 if `step` is not an exchange step. */
int32_t qb_plan_exchange_bits(const qb_plan* plan, int32_t step, int32_t* pos_out);

/* Amplitude sharding: perform exchange step `step` IN PLACE over NVLink peer memory.  peer_state_ptrs[i] = device pointer of rank
 * i's state shard mapped into this process (CUDA IPC / symmetric memory; peer_state_ptrs[rank] is the local shard).  All ranks must
 * call it between two cross-rank barriers.  Replaces an NCCL all-to-all + pack/unpack for QB_STEP_EXCHANGE steps, and handles any
 * exchanged bit positions (plan option exchange_any_bit). */
int qb_exchange_p2p_dev(const qb_plan* plan, int64_t batch, const void* const* peer_state_ptrs, int32_t rank,
                        int32_t world, int32_t step, void* stream);

/* The same exchange in PUSH form (NVLink carries only posted writes; qandle_b200/csrc/exchange.cuh): the chunk is cut into `pieces`;
 * for each piece all ranks run  barrier, phase 0 (push my chunk c -> slot `rank` of peer c's staging buffer, by TMA bulk copies),
 * barrier, phase 1 (unpack slot c of my staging buffer -> my chunk c).  peer_staging_ptrs[i] = rank i's staging buffer mapped into this
 * process, world * batch * chunk_bytes / pieces bytes each. */
int qb_exchange_push_dev(const qb_plan* plan, int64_t batch, void* state, const void* const* peer_staging_ptrs, int32_t rank,
                         int32_t world, int32_t piece, int32_t pieces, int32_t phase, void* stream);

/* One-call forward for an unsharded plan: prepare + (init) + all sweeps + measurement.
 * measure_out: STATE -> ignored (the result is `state`); PROBS -> [batch][n] real; JOINT -> [batch][2^n] real. */
int qb_forward_dev(const qb_plan* plan, int64_t batch, const void* shared_angles, const void* batch_angles,
                   int32_t n_batch_cols, const void* fixed_mats, int32_t init_kind, void* state, int32_t measure,
                   void* measure_out, void* workspace, void* stream);
/* One-call adjoint backward for an unsharded plan.  `state` must hold the forward's final state and is
 * un-computed in place back to the initial state; `lambda` is scratch of the same size and on return holds
 * dL/dpsi0* (torch's gradient w.r.t. the initial state is 2*lambda).  grad_out matches measure_out. */
int qb_backward_dev(const qb_plan* plan, int64_t batch, const void* shared_angles, const void* batch_angles,
                    int32_t n_batch_cols, const void* fixed_mats, void* state, void* lambda, int32_t measure,
                    const void* grad_out, void* grad_shared, int32_t n_shared, void* grad_batch, void* workspace,
                    void* stream);

/* The same with the forward's final state left INTACT: it is read from state_in, the un-computed copy is written to `state`
 * (scratch of the same size; state_in == state gives qb_backward_dev).  When the first adjoint sweep is the seed-fused complex64
 * sweep it reads state_in directly -- no extra pass; otherwise one device-to-device copy comes first.  This is what the
 * torch.library layer binds: autograd may run the backward of one forward more than once. */
int qb_backward_from_dev(const qb_plan* plan, int64_t batch, const void* shared_angles, const void* batch_angles,
                         int32_t n_batch_cols, const void* fixed_mats, const void* state_in, void* state, void* lambda,
                         int32_t measure, const void* grad_out, void* grad_shared, int32_t n_shared, void* grad_batch,
                         void* workspace, void* stream);

/* ---- host-buffer entry point (what a non-torch host binds: ctypes / cgo / JNI) --------------------
 * All pointers are HOST memory.  Runs forward (+ backward when grad_out != NULL) on the current CUDA
 * device, including the host<->device copies; blocks until the results are in the host buffers.
 * init_state may be NULL (|0..0>).  Any output pointer may be NULL. */
int qb_run_host(const qb_plan* plan, int64_t batch, const void* shared_angles, int32_t n_shared,
                const void* batch_angles, int32_t n_batch_cols, const void* fixed_mats, int32_t n_mats,
                const void* init_state, int32_t measure, void* measure_out, void* final_state_out,
                const void* grad_out, void* grad_shared, void* grad_batch, void* grad_init_state);

const char* qb_last_error(void);
const char* qb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QANDLE_B200_H */
