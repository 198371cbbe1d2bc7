/* qcsim_b200.h -- C ABI of the B200-native statevector engine (libqcsim_b200.so).
 *
 * This is the drop-in boundary for QCSim's statevector hot path.  The reference has no FFI
 * seam of its own -- the seam is the C++ class pair QC::QubitRegister / QC::QubitRegisterCalculator
 * -- so each entry point below cites the reference member it replaces (file:line relative to
 * /root/reference/QCSim/).  The C++ facade in qcsim_b200/cpp/QubitRegister.h and the Python
 * mirror in qcsim_b200/register.py bind exactly these symbols; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C types only: pointers, sizes, doubles.  No torch / Eigen / STL types.
 *   - an amplitude is two doubles (re, im), i.e. std::complex<double>; host buffers are
 *     interleaved (re, im) arrays.
 *   - qubit 0 is the least significant bit of the basis-state index (QubitRegister.h:7).
 *   - gate matrices are ROW-major (re, im) pairs, 2^nq x 2^nq, with row/col bit0 = `q`
 *     (target), bit1 = `c1`, bit2 = `c2` (QubitRegisterCalculator.h:427,749).  Eigen's default
 *     storage is column-major: the facade transposes on the way in.
 *   - every call returns 0 on success or a negative QCSIM_ERR_* code; qcsim_last_error()
 *     returns a thread-local message for the last failure.
 *   - a handle may be used from any host thread, one thread at a time.  Work is queued on the
 *     handle's CUDA stream; calls that return a value synchronise that stream.
 *   - there is NO CPU fallback: without a CUDA device every entry point that needs one fails
 *     with QCSIM_ERR_CUDA.
 */
#ifndef QCSIM_B200_H
#define QCSIM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QCSIM_ABI_VERSION 2

/* error codes ------------------------------------------------------------------------------ */
#define QCSIM_OK 0
#define QCSIM_ERR_QUBIT_TOO_HIGH -1   /* "Qubit number is too high"              QubitRegister.h:680 */
#define QCSIM_ERR_CTRL_TOO_HIGH -2    /* "Controlling qubit number is too high"  QubitRegister.h:682,687 */
#define QCSIM_ERR_SAME_QUBITS -3      /* "Qubit and controlling qubit are the same" / "Qubits must be different" :683,688 */
#define QCSIM_ERR_BAD_ARG -4          /* null pointer, nq outside 1..3, bad range ... */
#define QCSIM_ERR_BAD_STATE -5        /* basis state >= 2^n where the reference silently ignores it */
#define QCSIM_ERR_CUDA -6
#define QCSIM_ERR_NCCL -7
#define QCSIM_ERR_OOM -8
#define QCSIM_ERR_UNSUPPORTED -9

/* structural hints == the reference's virtual gate flags (SimpleGates.h:27-60).  With flags
 * set, the engine reads only the matrix block the reference kernel of that kind reads
 * (QubitRegisterCalculator.h:39-227).  flags == 0 ("flag-less", what AppliedGate / Compute /
 * Uncompute produce, QubitRegister.h:488-497,554-590) makes the engine classify the matrix
 * itself (identity blocks -> controls, diagonal, pair-swap) so it still gets the cheapest kernel. */
#define QCSIM_GATE_CONTROLLED 1    /* isControlled()                     */
#define QCSIM_GATE_TWO_CONTROLS 2  /* isControlQubit(1): c1 and c2 both control (Toffoli, CCZ) */
#define QCSIM_GATE_DIAGONAL 4      /* isDiagonal()  (of the controlled block) */
#define QCSIM_GATE_ANTIDIAGONAL 8  /* isAntidiagonal()                   */
#define QCSIM_GATE_SWAP 16         /* isSwapGate()  (2q: SWAP, 3q: Fredkin) */
#define QCSIM_GATE_ISWAP 32        /* IsISwapGate()                      */
#define QCSIM_GATE_ISWAPDAG 64     /* IsISwapDagGate()                   */

typedef struct qcsim_sv qcsim_sv; /* opaque register handle */

/* one gate application, the element type of qcsim_sv_apply_batch (an AppliedGate,
 * SimpleGates.h:439-578, plus the flags) */
typedef struct qcsim_gate {
  int32_t nq;    /* 1, 2 or 3 */
  int32_t flags; /* QCSIM_GATE_* or 0 */
  uint64_t q, c1, c2;
  double m[128]; /* row-major (re, im); first 8 * 4^nq / 4 ... i.e. 2 * 4^nq doubles used */
} qcsim_gate;

/* counters since creation / last reset; every field is a plain count */
typedef struct qcsim_stats {
  uint64_t gates_applied;   /* gate applications requested */
  uint64_t kernel_launches; /* CUDA kernels launched by this handle */
  uint64_t state_passes;    /* kernels that stream the (touched part of the) state through HBM */
  uint64_t bytes_moved;     /* algorithmic HBM bytes of those passes (read + write) */
  uint64_t exchange_calls;  /* global<->local qubit exchanges (sharded registers) */
  uint64_t exchange_bytes;  /* bytes sent over NVLink by this rank */
  double exchange_ms;       /* device time spent in exchanges (CUDA events) */
  uint64_t fused_rounds;    /* register rounds executed inside fused gate-block passes */
  uint64_t fused_ops;       /* gate applications executed inside fused gate-block passes */
} qcsim_stats;

const char* qcsim_last_error(void);
int qcsim_abi_version(void);
int qcsim_device_count(int* count);

/* ---- lifecycle: QubitRegister ctor / dtor / Clone (QubitRegister.h:17-57, 662-674) -------- */
/* state starts as |0...0> like the reference constructor */
int qcsim_sv_create(qcsim_sv** out, int n_qubits, int device);
/* sharded register: this process holds the slice whose top log2(world) index bits == rank.
 * nccl_id is the 128-byte ncclUniqueId produced by qcsim_nccl_unique_id on rank 0 and
 * broadcast by the host (torch.distributed / MPI / files -- the engine does not care). */
int qcsim_nccl_unique_id(void* out_128_bytes);
int qcsim_sv_create_sharded(qcsim_sv** out, int n_qubits, int device, int rank, int world,
                            const void* nccl_id_128_bytes);
/* One host thread, several GPUs (the form QC::QubitRegister needs: one object, one caller): the register is sharded
 * over `n_devices` (a power of two, <= 8) devices of this process on its top log2(n_devices) qubits.  Every call on
 * the returned handle runs on all shards (one worker thread per device inside the library); global<->local qubit
 * exchanges go through peer memory (cudaDeviceEnablePeerAccess), scalars through NCCL.  device_ids == NULL: devices
 * 0 .. n_devices-1.  n_devices == 1 is qcsim_sv_create.  Not supported on such a handle: clone, device_ptr,
 * apply_operator. */
int qcsim_sv_create_multi(qcsim_sv** out, int n_qubits, int n_devices, const int* device_ids);
int qcsim_sv_destroy(qcsim_sv* h);
int qcsim_sv_clone(const qcsim_sv* src, qcsim_sv** out);
int qcsim_sv_sync(qcsim_sv* h);  /* flush queued gates and wait for the stream */
int qcsim_sv_flush(qcsim_sv* h); /* submit queued gates to the stream without waiting */
int qcsim_sv_n_qubits(const qcsim_sv* h, int* n_qubits, int* n_local_qubits);
/* raw device pointer / stream of the local slice, for zero-copy interop (e.g. torch.from_blob) */
int qcsim_sv_device_ptr(qcsim_sv* h, void** dptr, void** cuda_stream);

/* ---- state setters / getters (QubitRegister.h:62-130, 507-524) ----------------------------- */
int qcsim_sv_set_basis_state(qcsim_sv* h, uint64_t state);      /* setToBasisState :74 */
int qcsim_sv_fill(qcsim_sv* h, double re, double im);           /* setConstant, Clear :108,121 */
int qcsim_sv_set_amplitude(qcsim_sv* h, uint64_t state, double re, double im); /* setRawAmplitude :112 */
int qcsim_sv_get_amplitude(qcsim_sv* h, uint64_t state, double* re_im);        /* getBasisStateAmplitude :62 */
/* global index range [first, first+count); on a sharded register the range must lie inside
 * the local slice (each rank moves its own part) */
int qcsim_sv_upload(qcsim_sv* h, const double* host, uint64_t first, uint64_t count);   /* setRegisterStorage* :512-524 */
int qcsim_sv_download(qcsim_sv* h, double* host, uint64_t first, uint64_t count);       /* getRegisterStorage :507 */
int qcsim_sv_norm2(qcsim_sv* h, double* out);                   /* squared norm, Normalize :126 */
int qcsim_sv_scale(qcsim_sv* h, double factor);                 /* registerStorage *= s :129 */
int qcsim_sv_normalize(qcsim_sv* h);                            /* Normalize :124-130 (no-op if norm < 1e-20) */
int qcsim_sv_save_state(qcsim_sv* h);                           /* SaveState :600 */
int qcsim_sv_restore_state(qcsim_sv* h, int destructive);       /* RestoreState / RestoreStateDestructive :605-616 */
/* <saved|psi> style inner product of another register with this one: conj(a) . b  (:527-534, :655) */
int qcsim_sv_inner_product(qcsim_sv* a, qcsim_sv* b, double* re_im);

/* ---- gates (QubitRegister.h:434-497; kernels QubitRegisterCalculator.h:39-939) ------------- */
int qcsim_sv_apply(qcsim_sv* h, int nq, const double* m, int flags, uint64_t q, uint64_t c1, uint64_t c2);
/* ApplyGates / Compute (QubitRegister.h:493-497, 554-569): same result as `count` calls of
 * qcsim_sv_apply, executed as fused shared-memory gate blocks (several gates per HBM pass) */
int qcsim_sv_apply_batch(qcsim_sv* h, const qcsim_gate* gates, uint64_t count);
/* 0 = every qcsim_sv_apply runs immediately as its own pass (reference behaviour);
 * 1 = qcsim_sv_apply only queues; the queue is fused and flushed by the next call that
 *     observes the state (measure, download, norm, sync ...).  Results are identical to 1e-15. */
int qcsim_sv_set_fusion(qcsim_sv* h, int enabled);
/* ApplyOperatorMatrix (QubitRegister.h:499-505): psi = M psi for a dense 2^n x 2^n operator, ROW-major (re, im)
 * pairs.  The reference's teaching path (Shor / dense-oracle Grover / phase estimation, and Compute / Uncompute of
 * recorded gates on more than 3 qubits, :563,581): a device GEMV, registers of at most QCSIM_MAX_OPERATOR_QUBITS
 * qubits (the matrix alone is 16 * 4^n bytes), unsharded. */
#define QCSIM_MAX_OPERATOR_QUBITS 13
int qcsim_sv_apply_operator(qcsim_sv* h, const double* m);
/* QuantumFourierTransform::QFT / IQFT on qubits [sq, eq] (QuantumFourierTransform.h:35-87),
 * including QubitsSwapper::Swap when do_swap (QubitsSwapper.h:23-34) */
int qcsim_sv_qft(qcsim_sv* h, uint64_t sq, uint64_t eq, int do_swap, int inverse);

/* ---- circuit files ------------------------------------------------------------------------------
 * A recorded gate stream (what ComputeStart/ComputeEnd record and ApplyGates / Compute replay, QubitRegister.h:493-497,
 * 536-590) as a file both sides of the parity harness read: the engine through qcsim_sv_apply_circuit_file (=
 * qcsim_sv_apply_batch on its records), the compiled reference through oracle/ref_driver.cpp::ref_apply_circuit_file.
 * Layout (little endian): char magic[8] = "QCSIMC1\0"; uint32 n_qubits; uint32 reserved; uint64 count; then per gate
 *   int32 nq, flags, gate_id, reserved; uint64 q, c1, c2; double params[4]; double m[2 * 4^nq] (row-major re, im).
 * gate_id / params name the reference gate class (oracle/ref_driver.cpp: makeGate) or are -1 / 0 for a plain matrix;
 * the engine only reads the matrix and the flags. */
typedef struct qcsim_circuit_gate {
  int32_t nq, flags, gate_id, reserved;
  uint64_t q, c1, c2;
  double params[4];
  double m[128];
} qcsim_circuit_gate;
int qcsim_circuit_save(const char* path, uint32_t n_qubits, const qcsim_circuit_gate* gates, uint64_t count);
/* *gates is malloc'ed by the library: release it with qcsim_circuit_free */
int qcsim_circuit_load(const char* path, uint32_t* n_qubits, qcsim_circuit_gate** gates, uint64_t* count);
void qcsim_circuit_free(qcsim_circuit_gate* gates);
int qcsim_sv_apply_circuit_file(qcsim_sv* h, const char* path);

/* ---- measurement (QubitRegister.h:169-224, 592-598, 619-642; Calculator :948-1254) --------- */
/* `prob` is the reference's `1. - uniformZeroOne(rng)`; the RNG stays with the caller */
int qcsim_sv_measure_all(qcsim_sv* h, double prob, uint64_t* outcome);                 /* MeasureAll :169 */
int qcsim_sv_measure(qcsim_sv* h, uint64_t first, uint64_t last, double prob, uint64_t* outcome); /* Measure/MeasureQubit :198-224 */
int qcsim_sv_measure_all_nocollapse(qcsim_sv* h, double prob, uint64_t* outcome);      /* MeasureNoCollapse() :619 */
int qcsim_sv_measure_nocollapse(qcsim_sv* h, uint64_t first, uint64_t last, double prob, uint64_t* outcome); /* :705 */
int qcsim_sv_qubit_probability(qcsim_sv* h, uint64_t q, double* p);                    /* GetQubitProbability :592 */
/* RepeatedMeasure (QubitRegister.h:227-429): `count` draws against ONE cumulative table -- the reference's sequential
 * running sum, built once on the device (two passes over the state, whatever `count` is).  outcomes[i] is what the
 * reference's std::lower_bound over its table gives for probs[i], including its quirks: the table is cut at the first
 * entry above 1 - DBL_EPSILON (:250-254) and a draw above the last stored entry yields the table size (:268). */
int qcsim_sv_sample(qcsim_sv* h, const double* probs, uint64_t count, uint64_t* outcomes);
/* Kept for ABI compatibility: every measurement reproduces the reference's strictly sequential fp64 running sum
 * bit for bit (csrc/reduce_kernels.cuh), so there is no non-strict mode any more; the call is a no-op. */
int qcsim_sv_set_strict_measure(qcsim_sv* h, int enabled);

/* ---- introspection --------------------------------------------------------------------------- */
int qcsim_sv_get_stats(const qcsim_sv* h, qcsim_stats* out);
int qcsim_sv_reset_stats(qcsim_sv* h);

#ifdef __cplusplus
}
#endif
#endif /* QCSIM_B200_H */
