/*
 * aqs_engine.h — C ABI of the B200-native state-vector engine (libaqs_engine.so).
 *
 * This is the drop-in boundary for afQuantumSim's hot path.  The reference has
 * no FFI: its gate classes call ArrayFire directly (af::matmul on an explicit
 * 2^n x 2^n operator, src/quantum.cpp:277-291 and every QGate::operator()).
 * The entry points below are what a maintainer binds in place of those
 * ArrayFire call sites; each one cites the reference code it replaces
 * (paths relative to the reference repository root).  INTEGRATION.md shows the
 * reference-side patch.
 *
 * Conventions
 *   - amplitudes: complex64, interleaved (re, im)         include/quantum.h:107-111
 *   - qubit numbering: API qubit 0 is the MOST significant index bit;
 *     qubit q lives at bit n-1-q of the amplitude index    src/quantum.cpp:546
 *   - every function returns 0 on success, a negative aqs_status otherwise and
 *     never throws; aqs_last_error() gives the thread-local message.  The host
 *     layer performs the reference's argument checks first and rethrows the
 *     reference's exception types.
 *   - gate calls are asynchronous on the state's CUDA stream; functions that
 *     return data to the host synchronise that stream.
 *   - host buffers are caller-owned and only touched during the call.
 *   - a state handle is not thread-safe; distinct handles are independent.
 */
#ifndef AQS_ENGINE_H
#define AQS_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AQS_ENGINE_ABI_VERSION 1
#define AQS_MAX_QUBITS 36 /* engine limit; the drop-in C++ API keeps max_qubit_count = 30 (include/quantum.h:122) */

typedef struct { float re, im; } aqs_c32;

typedef struct aqs_state_s* aqs_state_t; /* one 2^n complex64 state vector resident in HBM */
typedef struct aqs_plan_s*  aqs_plan_t;  /* a compiled (optionally fused) gate list */
typedef struct aqs_timer_s* aqs_timer_t; /* CUDA-event pair on a state's stream */

typedef enum {
    AQS_OK = 0,
    AQS_ERR_INVALID = -1,   /* bad argument */
    AQS_ERR_CUDA = -2,      /* CUDA runtime failure */
    AQS_ERR_NOMEM = -3,     /* device or host allocation failed */
    AQS_ERR_STATE = -4,     /* engine not initialised / handle destroyed */
    AQS_ERR_COMM = -5       /* NCCL failure (sharded states) */
} aqs_status;

/* Primitive operations every reference gate flattens to. */
typedef enum {
    AQS_OP_U2 = 0,   /* general 2x2 on `target`: H :817-829, RotX/RotY :683-693/:730-740, Y :586-606,
                        CH :1355-1397, CY :1054-1085, CRotX/CRotY :1433/:1470 (all src/quantum.cpp) */
    AQS_OP_DIAG = 1, /* diag(m[0], m[3]) on `target`: Z :633-650, Phase :855-874, RotZ :777-787,
                        CZ :1128-1149, CPhase :1192-1215, CRotZ :1507 */
    AQS_OP_X = 2,    /* bit flip of `target`: X :546-559, CX :992-1011, CCNot :1558-1577, Or :1630-1650 */
    AQS_OP_SWAP = 3  /* exchange `target` and `target2`: Swap :925-949, CSwap :1284-1310 */
} aqs_op_kind;

/* One primitive, in API qubit numbering.  The op acts only on amplitudes whose
 * index satisfies, for every qubit q with bit q of ctrl_mask set:
 *     (value of qubit q) == bit q of ctrl_value.
 * ctrl_value == ctrl_mask is the ordinary "all controls are |1>" case;
 * ControlGate nests (src/quantum.cpp:1888-1950) flatten by OR-ing their control
 * into the mask.  Or(a,b,t) is X(t) followed by X(t) with mask {a,b}, value 0. */
typedef struct aqs_op {
    int32_t  kind;       /* aqs_op_kind */
    int32_t  target;     /* API qubit */
    int32_t  target2;    /* second qubit of AQS_OP_SWAP, otherwise -1 */
    int32_t  reserved;
    uint64_t ctrl_mask;
    uint64_t ctrl_value;
    aqs_c32  m[4];       /* row-major 2x2 (m00 m01 m10 m11); DIAG reads m[0], m[3]; X/SWAP ignore it */
} aqs_op;

/* ---- engine lifetime --------------------------------------------------------
 * replaces aqs::initialize: af::setBackend / af::setDevice / constant upload
 * (src/quantum.cpp:69-86).  `device` has the meaning of argv[1] there. */
int aqs_engine_init(int device);
int aqs_engine_shutdown(void);
int aqs_engine_device(int* device, int* sm_count, size_t* hbm_bytes);
int aqs_engine_abi_version(void);
const char* aqs_last_error(void);

/* ---- state vector -----------------------------------------------------------
 * aqs_state_create: af::constant(0, 2^n, c32) of the QSimulator constructors
 * (src/quantum.cpp:212-218); the state is left at |0...0>.  */
int aqs_state_create(int n_qubits, aqs_state_t* out);
/* same, on caller-owned device memory (2^n complex64, 16-byte aligned): the engine
 * neither allocates nor frees it and leaves its contents untouched.  Lets a host that
 * owns the allocation (torch tensors for the NCCL half-shard exchanges) run kernels on it. */
int aqs_state_wrap(int n_qubits, void* device_ptr, aqs_state_t* out);
int aqs_state_destroy(aqs_state_t s);
/* copy construction of QSimulator (af::array value semantics; used e.g. by
 * examples/quantum_teleportation.cpp) */
int aqs_state_clone(aqs_state_t src, aqs_state_t* out);
int aqs_state_qubits(aqs_state_t s, int* n_qubits);
/* statevector_(k) = 1 on a zero vector: src/quantum.cpp:219-222 and measure_all :366-367 */
int aqs_state_set_basis(aqs_state_t s, uint64_t index);
/* generate_statevector: Kronecker product of n single-qubit states, qubit 0 first
 * (src/quantum.cpp:261-275, src/utils.cpp:169-195).  q is n x 2 complex64. */
int aqs_state_set_product(aqs_state_t s, const aqs_c32* q);
/* identity "matrix" I(2^m) laid out column-major as a state of n = 2m qubits:
 * used to materialise QCircuit::circuit() (af::identity, src/quantum.cpp:159-164) */
int aqs_state_set_identity(aqs_state_t s);
/* af::array(2^n, host ptr) / .host(): src/quantum.cpp:242-259, src/quantum_visuals.cpp:35 */
int aqs_state_upload(aqs_state_t s, const aqs_c32* host, uint64_t offset, uint64_t count);
int aqs_state_download(aqs_state_t s, aqs_c32* host, uint64_t offset, uint64_t count);
/* statevector_(k).scalar<af::cfloat>(): include/quantum.h:658-661 */
int aqs_state_get_amp(aqs_state_t s, uint64_t index, aqs_c32* out);
/* raw device pointer + stream, for zero-copy interop (torch / NCCL plumbing) */
int aqs_state_device_ptr(aqs_state_t s, void** dptr);
int aqs_state_set_stream(aqs_state_t s, void* cuda_stream);
int aqs_state_get_stream(aqs_state_t s, void** cuda_stream);
int aqs_sync(aqs_state_t s);

/* ---- gate application -------------------------------------------------------
 * aqs_apply_op: one reference gate application, i.e. what
 * `circuit = af::matmul(M_gate, circuit)` does to a state vector inside
 * QSimulator::simulate's loop (src/quantum.cpp:287-289).  One kernel launch that
 * touches exactly the amplitudes the op changes. */
int aqs_apply_op(aqs_state_t s, const aqs_op* op);
int aqs_apply_ops(aqs_state_t s, const aqs_op* ops, uint64_t n_ops);

/* An opaque 2^k x 2^k matrix (row-major, m[r * 2^k + c]) on k distinct qubits, 1 <= k <= 6, under the
 * same control convention as aqs_op: what Gate::operator() / ControlGate::operator() do with a
 * compiled inner circuit's matrix (src/quantum.cpp:1760-1814, 1888-1950; bit layout src/utils.cpp:137-167),
 * for arbitrary target qubits.  qubits[0] is the MOST significant bit of the matrix index (inner qubit 0
 * of the reference's Gate).  One launch, in place, FP32. */
int aqs_apply_dense(aqs_state_t s, const int* qubits, int k, uint64_t ctrl_mask, uint64_t ctrl_value, const aqs_c32* m);

/* ---- compiled circuits ------------------------------------------------------
 * replaces QCircuit::compile (src/quantum.cpp:199-210): instead of a dense
 * 2^n x 2^n unitary the engine builds a launch plan.  With AQS_PLAN_FUSE the
 * planner groups consecutive ops into passes that stage a tile of the state in
 * shared memory and apply the whole group in place. */
#define AQS_PLAN_FUSE  1u   /* gate-fusion pass on */
#define AQS_PLAN_GRAPH 2u   /* capture the launch sequence in a CUDA graph */
#define AQS_PLAN_JIT   4u   /* (with AQS_PLAN_FUSE) specialise every fused pass: a straight-line sm_100a kernel per pass
                             * SHAPE, compiled at run time and cached by shape for the life of the process; the
                             * coefficients stay kernel parameters, so circuits that differ only in their angles
                             * share kernels.  aqs_plan_build returns when the kernels are compiled. */
#define AQS_PLAN_JIT_ASYNC 8u /* same, but compilation runs on background host threads: a pass runs on the generic
                             * (interpreting) tile kernel until its specialised kernel is ready */
typedef struct aqs_plan_info {
    uint64_t n_ops;            /* primitive ops (= gate applications) */
    uint64_t n_launches;       /* kernel launches per run */
    uint64_t n_fused_passes;   /* launches of the tile kernel */
    uint64_t n_single_ops;     /* ops left on the per-gate kernels */
    double   bytes_unfused;    /* algorithmic HBM bytes if every op ran alone */
    double   bytes_planned;    /* algorithmic HBM bytes of this plan */
    int32_t  n_qubits;
    int32_t  tile_bits;        /* log2 amplitudes per shared-memory tile (0 if unfused) */
} aqs_plan_info;
int aqs_plan_build(int n_qubits, const aqs_op* ops, uint64_t n_ops, uint32_t flags, aqs_plan_t* out);
int aqs_plan_run(aqs_state_t s, aqs_plan_t p);
int aqs_plan_get_info(aqs_plan_t p, aqs_plan_info* info);
int aqs_plan_destroy(aqs_plan_t p);
/* Introspection (tests, tools): the launch descriptors of fused pass `index`, exactly as the tile
 * kernel receives them (record layout: afquantumsim_b200/csrc/plan.cu).  *needed gets the record
 * size; the record is copied when cap is large enough.  No reference counterpart. */
int aqs_plan_export_pass(aqs_plan_t p, uint64_t index, void* buf, uint64_t cap, uint64_t* needed);
/* Specialised passes (afquantumsim_b200/csrc/specialize.cu; no reference counterpart: QCircuit::compile's product is
 * a dense matrix, src/quantum.cpp:199-210).  aqs_plan_pass_source: the generated CUDA C++ of fused pass `index`
 * (also valid host C++ under -DAQS_HOST_EMU, which is how the CPU tests execute it) and aqs_plan_pass_coefs: its
 * table of packed 64-bit coefficient operands; both work without a GPU and without AQS_PLAN_JIT.  *geom gets
 * {threads per CTA, dynamic shared-memory bytes, CTAs of a full launch}.  aqs_plan_jit_ready: how many fused passes
 * of the plan would run specialised if the plan ran now.  aqs_jit_wait blocks until no compilation is pending. */
int aqs_plan_pass_source(aqs_plan_t p, uint64_t index, char* buf, uint64_t cap, uint64_t* needed, uint64_t* geom);
int aqs_plan_pass_coefs(aqs_plan_t p, uint64_t index, uint64_t* buf, uint64_t cap, uint64_t* needed);
int aqs_plan_jit_ready(aqs_plan_t p, uint64_t* n_ready);
int aqs_jit_wait(void);
typedef struct aqs_jit_info {
    uint64_t compiled;         /* kernels compiled by this process */
    uint64_t cache_hits;       /* passes that found their shape already compiled (or compiling) */
    uint64_t failed;           /* compilations that failed (those passes stay on the generic kernel) */
    uint64_t pending;          /* compilations queued or running */
    double   compile_seconds;  /* host-thread seconds spent compiling (summed over threads) */
} aqs_jit_info;
int aqs_jit_get_info(aqs_jit_info* out);

/* ---- probabilities and measurement -----------------------------------------
 * Exact-sum contract: p_k = fl32(fl32(re*re)+fl32(im*im)),
 * F_k = trunc(p_k * 2^62) as uint64, all sums are integer sums of F_k, so every
 * result below is independent of summation order and bit-identical between the
 * CPU oracle, one GPU and R GPUs.  Requires ||psi||^2 < 4. */

/* af::norm(statevector): src/quantum.cpp:253-258 (plain double reduction) */
int aqs_norm2(aqs_state_t s, double* out);
/* statevector / norm: src/quantum.cpp:258 */
int aqs_scale(aqs_state_t s, float factor);
/* sum of F_k over indices whose qubits in `mask` (API numbering) read `value`, as 2^-62 units */
int aqs_prob_fixed(aqs_state_t s, uint64_t qubit_mask, uint64_t qubit_value, uint64_t* out);
/* qubit_probability_true: src/quantum.cpp:372-391 */
int aqs_qubit_prob1(aqs_state_t s, int qubit, double* out);
/* probabilities(): src/quantum.cpp:404-414 (f32 |a|^2 to a host buffer) */
int aqs_probabilities(aqs_state_t s, float* host_out, uint64_t offset, uint64_t count);
/* the collapse of measure(): keep the half where `qubit` == outcome, zero the
 * rest, divide by sqrtf(p): src/quantum.cpp:336-339 */
int aqs_collapse_qubit(aqs_state_t s, int qubit, int outcome, float p);
/* peek_measure_all / profile_measure_all: for each draw u_i in [0,1) the first
 * index k whose inclusive cumulative probability exceeds u_i, 0 if none
 * (src/quantum.cpp:344-359, 467-501).  Draws are an INPUT so that results are
 * reproducible; the host layer owns the RNG (src/quantum.cpp:57-62). */
int aqs_sample(aqs_state_t s, const float* u_host, uint64_t n_draws, uint64_t* out_index_host);
/* same rule with the draws already in the fixed-point domain (U = trunc(u * 2^62), possibly
 * minus the probability mass of lower-ranked shards): the sharded sampler's local step.
 * out = local index of the first S_k > U, or UINT64_MAX if U >= the shard's total. */
int aqs_sample_fixed(aqs_state_t s, const uint64_t* u_fixed_host, uint64_t n_draws, uint64_t* out_index_host);
/* same, as the dense histogram profile_measure_all returns (std::vector<uint32_t>(2^n),
 * src/quantum.cpp:470,498); hist_host has 2^n entries.  Only the n_draws outcome indices cross the bus
 * (8 bytes per draw); the dense vector is filled on the host. */
int aqs_sample_hist(aqs_state_t s, const float* u_host, uint64_t n_draws, uint32_t* hist_host);
/* the same histogram as sorted (index, count) pairs: what the reference's sort + countByKey produce
 * before they are scattered into the dense vector (src/quantum.cpp:490-498).  *n_bins gets the number of
 * distinct outcomes (<= n_draws); the pairs are written when index / count are non-null and cap >= *n_bins.
 * This is the histogram API for states above 30 qubits, where a dense vector is not an option. */
int aqs_sample_hist_sparse(aqs_state_t s, const float* u_host, uint64_t n_draws, uint64_t* index_host, uint32_t* count_host,
                           uint64_t cap, uint64_t* n_bins);

/* ---- peer memory: sharded states on one NVLink / NVSwitch node -------------------
 * No reference counterpart (the reference is single-device).  A state of
 * n qubits is sharded over 2^g GPUs, one process each; the top g index bits are the rank.
 * aqs_state_ipc_export / aqs_ipc_open hand an engine-allocated shard to the other processes
 * (CUDA IPC; opened mappings are cached until aqs_ipc_close_all / aqs_engine_shutdown).
 * aqs_peer_bitswap swaps k (global bit, local bit) pairs of the sharded index IN PLACE, reading
 * and writing the partners' shards directly over NVLink: members[v] is the shard of the group
 * member whose k selected rank bits read v (members[my_value] is ignored), local_bits[i] is the
 * INDEX-BIT position (not an API qubit number; >= 1) traded with selected rank bit i.  Every
 * member of the group must call it with the same k and local_bits, between two cross-rank
 * barriers on the stream (afquantumsim_b200/sharded.py: stream-ordered NCCL all_reduce). */
#define AQS_IPC_HANDLE_BYTES 64
int aqs_state_ipc_export(aqs_state_t s, void* handle_out /* AQS_IPC_HANDLE_BYTES */);
int aqs_ipc_open(const void* handle, void** peer_ptr);
int aqs_ipc_close_all(void);
int aqs_peer_bitswap(aqs_state_t s, void* const* members, int k, const int* local_bits, uint32_t my_value);

/* ---- flat multi-GPU address space ----------------------------------------------
 * No reference counterpart.  Every rank allocates its shard with the CUDA virtual memory
 * management API, exports it as a POSIX file descriptor (*fd_out; the host layer passes it to
 * the other processes, afquantumsim_b200/sharded.py), imports the others' (aqs_flat_attach)
 * and maps all `world` shards back to back: amplitude k of the whole 2^n state is at
 * base + 8k on every GPU.  aqs_state_wrap(n, base) then gives a state handle on the WHOLE
 * state, and aqs_plan_run_shard runs this rank's share (1/world of the tiles) of passes
 * [first, first + count) of a fused plan on it: a pass whose tile contains rank bits reads and
 * writes peer memory over NVLink while it computes.  aqs_plan_pass_span reports how many rank
 * bits a pass's tile contains; the caller puts a cross-rank barrier on the stream between a
 * pass that spans GPUs and its neighbours.  shard_bytes must be a multiple of the device's
 * mapping granularity (2 MiB). */
typedef struct aqs_flat_s* aqs_flat_t;
int aqs_flat_create(uint64_t shard_bytes, int world, int rank, aqs_flat_t* out, int* fd_out);
int aqs_flat_attach(aqs_flat_t f, int peer_rank, int fd);
int aqs_flat_ptr(aqs_flat_t f, void** base, void** own_shard);
int aqs_flat_destroy(aqs_flat_t f);
int aqs_plan_run_shard(aqs_state_t s, aqs_plan_t p, uint64_t first, uint64_t count, int rank, int log2_world);
int aqs_plan_pass_span(aqs_plan_t p, uint64_t index, int log2_world, int* rank_bits_in_tile);
/* the index-bit positions of the tile of fused pass `index`, ascending (pos: 16 entries of storage) */
int aqs_plan_pass_tile(aqs_plan_t p, uint64_t index, uint8_t* pos, int* tile_bits);
/* introspection (tests): which tiles of pass `index` rank `rank` runs — the tile numbers whose bits
 * fix_pos[0 .. *fix_n) (ascending, 8 entries of storage) equal those of *fix_or */
int aqs_plan_shard_cut(aqs_plan_t p, uint64_t index, int rank, int log2_world, uint32_t* fix_n, uint32_t* fix_or, uint8_t* fix_pos);
/* Staged passes (flat.cu, sharded.py).  Scattered 256-byte reads of peer HBM run at about half the NVLink rate, large
 * contiguous copies and peer writes at the full rate, so a pass whose tiles contain rank bits is run in CHUNKS of tiles:
 * the copy engines fetch the next chunk's remote blocks into local staging memory (aqs_memcpy_async, to the view's own
 * addresses) while the kernel computes the current chunk, reading a VIEW of the state in which those blocks are backed by
 * local staging allocations (aqs_flat_view_create) and writing its results directly into the peers' HBM (aqs_plan_run_tiles: one pass,
 * restricted to the tiles whose number has the listed bits pinned, loads from `load_base`, stores in place). */
typedef struct aqs_flat_block { uint64_t state_offset, bytes; } aqs_flat_block;   /* in bytes; multiples of the 2 MiB mapping granularity */
int aqs_flat_view_create(aqs_flat_t f, const aqs_flat_block* blocks, uint64_t n_blocks, void** view_base);
int aqs_memcpy_async(void* dst, const void* src, uint64_t bytes, void* stream);
int aqs_plan_run_tiles(aqs_state_t s, aqs_plan_t p, uint64_t index, const void* load_base, uint32_t fix_n, const uint8_t* fix_pos,
                       uint32_t fix_or, void* stream);

/* ---- timing (CUDA events on the state's stream) ---------------------------- */
int aqs_timer_create(aqs_timer_t* out);
int aqs_timer_start(aqs_timer_t t, aqs_state_t s);
int aqs_timer_stop(aqs_timer_t t, aqs_state_t s);
int aqs_timer_elapsed_ms(aqs_timer_t t, double* ms); /* synchronises on the stop event */
int aqs_timer_destroy(aqs_timer_t t);

/* ---- counters -------------------------------------------------------------- */
typedef struct aqs_counters {
    uint64_t kernel_launches;  /* engine kernels launched since init / last reset */
    uint64_t gate_ops;         /* primitive ops applied */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
} aqs_counters;
int aqs_counters_get(aqs_counters* out);
int aqs_counters_reset(void);
/* give every cached device buffer back to the driver (the engine recycles state-sized buffers between states;
 * other allocators of the process — torch, NCCL, the flat address space — do not see that cache) */
int aqs_pool_trim(void);

#ifdef __cplusplus
}
#endif
#endif /* AQS_ENGINE_H */
