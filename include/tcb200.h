/*
 * tcb200.h -- C ABI of the B200-native statevector engine behind TensorCircuit's
 * circuit-evaluation hot path.
 *
 * The reference (tencent-quantum-lab/tensorcircuit) has no FFI of its own: the path is pure
 * Python that hands a tensor-network node graph to a contractor
 * (tensorcircuit/cons.py:523-631) which executes numpy/jax tensordot calls.  Each entry point
 * below names the reference interface whose O(2^n) work it replaces; INTEGRATION.md shows the
 * ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain C: pointers, sizes, ints.  No torch / C++ types cross this boundary.
 *   - every function returns 0 on success, a negative code on failure; the message is
 *     retrievable (thread-local) through tcb200_last_error().  Nothing throws.
 *   - `state` is a DEVICE pointer to `batch` contiguous vectors of 2^nbits amplitudes,
 *     interleaved (re, im); dtype TCB200_C64 = float2, TCB200_C128 = double2.  Caller-owned,
 *     never freed or reallocated by the library; kernels update it in place.
 *   - amplitude index bit b of a vector <-> TensorCircuit qubit (nbits-1-b): qubit 0 is the
 *     most significant bit (tensorcircuit/quantum.py:1439, quantum.py:2104-2119).
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered and the calls
 *     are asynchronous unless stated otherwise.  Re-entrant; no hidden global state besides a
 *     per-process cache of device attributes.
 *   - there is no CPU fallback: every call fails with TCB200_ERR_CUDA when no device/context
 *     is usable.
 */
#ifndef TCB200_H
#define TCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TCB200_C64 0
#define TCB200_C128 1

#define TCB200_OK 0
#define TCB200_ERR_ARG (-1)         /* bad argument (null pointer, bit out of range, duplicate bit ...) */
#define TCB200_ERR_UNSUPPORTED (-2) /* e.g. k > TCB200_MAX_K */
#define TCB200_ERR_WORKSPACE (-3)   /* workspace too small */
#define TCB200_ERR_CAPACITY (-4)    /* a gate pass holds more rounds / matrices than one launch carries: split it */
#define TCB200_ERR_CUDA (-1000)     /* -(1000 + cudaError_t) */

#define TCB200_MAX_K 5          /* widest dense fused block (2^5 x 2^5) */
#define TCB200_MAX_DIAG_K 12    /* widest diagonal fused block (2^12 entries) */
#define TCB200_MAX_PASS_OPS 16  /* fused blocks executed by one staged pass */
#define TCB200_MAX_PASS_K 4     /* widest block inside a staged multi-block pass */
#define TCB200_MAX_TERMS 8      /* Pauli strings per expectation launch */
#define TCB200_MAX_GATE_PASS_OPS 256 /* gates handed to one structure-aware gate pass */

/* Library / build identification: "tcb200 <version> sm_100a". */
const char* tcb200_version(void);

/* Message of the last failing call made by this thread ("" if none). */
const char* tcb200_last_error(void);

/*
 * |0...0> for every batch element.
 * Replaces BaseCircuit.all_zero_nodes (tensorcircuit/basecircuit.py:46-60) + the contraction
 * that materialises it.
 */
int tcb200_init_zero(void* state, int nbits, int dtype, int64_t batch, void* stream);

/*
 * All-zero vector (a shard of a distributed |0...0> that does not hold amplitude 0).
 */
int tcb200_set_zero(void* state, int nbits, int dtype, int64_t batch, void* stream);

/*
 * Device-to-device copy of `nrows` runs of `row_bytes` bytes (pitches in bytes): the pack /
 * unpack step of a chunked global<->local qubit remap.  Runs on the copy path of `stream`.
 */
int tcb200_copy_rows(void* dst, size_t dst_pitch, const void* src, size_t src_pitch,
                     size_t row_bytes, size_t nrows, void* stream);

/*
 * Cast/copy an initial state into the engine buffer (device -> device, complex128 source):
 * Circuit(n, inputs=...) (tensorcircuit/circuit.py:86-96).  `src_c128` holds 2^nbits double2.
 */
int tcb200_load_c128(void* state, int nbits, int dtype, const void* src_c128, void* stream);

/*
 * In-place application of one fused dense block  psi <- U psi  on k <= TCB200_MAX_K bits.
 *
 *   bits[k]   HOST array, strictly ascending amplitude-index bit positions.
 *   mat       HOST array of 2^k x 2^k complex128 (re, im interleaved), row-major U[out][in];
 *             bit j of the row/column index is the value of amplitude bit bits[j].
 *             It is cast to the state's dtype and travels in the kernel-parameter constant
 *             bank.  U need not be unitary (gates.py:225-232 allows complex parameters).
 *
 * One HBM read + one HBM write of the state.  Replaces the per-path-step
 * tn.contract_between -> backend.tensordot of tensorcircuit/cons.py:605-623 and the final
 * reorder_edges transpose (cons.py:629-630) for the gates merged into U.
 */
int tcb200_apply_dense(void* state, int nbits, int dtype, int k, const int* bits,
                       const double* mat, int64_t batch, void* stream);

/*
 * Batched variant for backend.vmap (tensorcircuit/backends/jax_backend.py:718-730): batch
 * element b is multiplied by its own matrix.
 *   mats_dev  DEVICE array [batch][2^k][2^k] in the state's dtype.
 */
int tcb200_apply_dense_batched(void* state, int nbits, int dtype, int k, const int* bits,
                               const void* mats_dev, int64_t batch, void* stream);

/*
 * In-place multiplication by a diagonal fused block on k <= TCB200_MAX_DIAG_K bits
 * (rz / phase / rzz / cz / cphase runs): psi_r <- d[idx(r)] psi_r with idx bit j = bit bits[j]
 * of r.  `diag` is a HOST array of 2^k complex128.  `diag_dev` (optional, may be NULL) is the
 * batched form, DEVICE [batch][2^k] in the state's dtype; when given, `diag` is ignored.
 */
int tcb200_apply_diag(void* state, int nbits, int dtype, int k, const int* bits,
                      const double* diag, const void* diag_dev, int64_t batch, void* stream);

/*
 * Staged multi-block pass: the tile {low bits} U {tile_hi bits} is brought into shared memory
 * once, `nops` dense blocks whose bits all lie inside the tile are applied back to back, and
 * the tile is written back -- one HBM read + write for the whole run of blocks.
 *
 *   ops_k[nops]         HOST, k of each block (<= TCB200_MAX_PASS_K)
 *   ops_bits[sum k]     HOST, concatenated ascending bit lists
 *   ops_mats_dev        DEVICE, concatenated matrices in the state's dtype, block i holding
 *                       [batch_mats][2^k_i][2^k_i]; batch_mats is 1 (shared) or `batch`
 *   tile_hi[n_hi]       HOST, ascending bit positions >= the contiguous low part that the
 *                       tile gathers (every block bit must be < nbits_low or in tile_hi)
 */
int tcb200_apply_pass(void* state, int nbits, int dtype, int nops, const int* ops_k,
                      const int* ops_bits, const void* ops_mats_dev, int64_t batch_mats,
                      int n_hi, const int* tile_hi, int64_t batch, void* stream);

/*
 * Same staged multi-block pass for matrices shared by all batch elements, given on the HOST
 * (complex128, concatenated, 4^k_i entries each, at most 12 KiB in the state's dtype): they are
 * cast and travel in the kernel-parameter constant bank, so the FMAs read them as constant
 * operands.  This is the kernel the pass planner (fusion.plan_passes) drives: a run of fused
 * blocks costs one HBM read + write however many blocks it holds.
 */
int tcb200_apply_pass_host(void* state, int nbits, int dtype, int nops, const int* ops_k,
                           const int* ops_bits, const double* ops_mats, int n_hi,
                           const int* tile_hi, int64_t batch, void* stream);

/*
 * Register-tile pass: the staged pass above, with one more level of blocking.  Each of the
 * `nrt` (<= TCB200_MAX_PASS_OPS) register tiles is a set of <= 4 bits inside the shared-memory
 * tile; a thread loads the 2^kt amplitudes of a group once, applies the tile's `rt_nsub` gates
 * (1 or 2 bits each, all bits inside the register tile) back to back in registers, and stores
 * the group once -- one shared-memory round trip per register tile instead of one per gate.
 *
 *   rt_k[nrt], rt_bits[sum rt_k]   HOST: bits of each register tile, ascending
 *   rt_nsub[nrt]                    HOST: gates per register tile, applied in order
 *   sub_k[], sub_bits[sum sub_k]    HOST: per gate, its ascending bits (subset of its tile)
 *   sub_mats                        HOST complex128, concatenated 4^k entries per gate
 *                                   (at most 12 KiB in the state's dtype, 96 gates per pass)
 */
int tcb200_apply_rpass_host(void* state, int nbits, int dtype, int nrt, const int* rt_k,
                            const int* rt_bits, const int* rt_nsub, const int* sub_k,
                            const int* sub_bits, const double* sub_mats, int n_hi,
                            const int* tile_hi, int64_t batch, void* stream);

/*
 * Structure-aware staged pass ("gate pass"): the production kernel behind Circuit gate application.
 * Same tile and same argument meaning as tcb200_apply_pass_host -- `nops` gates in program order,
 * every bit inside the tile {low bits} U {tile_hi} -- but the library classifies each matrix
 * (exact zeros) and treats the three classes of tensorcircuit/gates.py differently:
 *   - permutation matrices of an affine bit map (x, cnot, swap, cx chains; any monomial gate on
 *     <= 2 bits such as y, cy, iswap after its phases are split off): absorbed into the tile's
 *     GF(2)-affine index map on the host -- no device instruction, undone by the write-back;
 *   - diagonal gates on <= 4 bits (z, s, t, rz, phase, cz, cphase, rzz, crz ...): one 16-entry
 *     table multiply, consecutive ones merged into one table;
 *   - everything else on <= 3 bits: dense 2x2 / 4x4 / 8x8 in registers.
 * Gates are scheduled into rounds: per round a thread loads 16 amplitudes (4 tile bits) once,
 * applies every ready gate that lives on those bits (dependencies respected, gates on disjoint
 * bits reordered freely), and stores them once.  Matrices travel in the kernel-parameter
 * constant bank and reach the FMAs as uniform-register operands.
 *
 *   ops_k[nops] (1..4), ops_bits[sum k] ascending, ops_mats HOST complex128 [sum 4^k]
 *   info8       optional HOST double[8]: rounds, gates absorbed into the index map, diagonal gates,
 *               dense gates, rounds with shared-memory bank conflicts, rounds using 16-byte
 *               accesses, real FMAs per amplitude, parameter-bank elements used
 * Returns TCB200_ERR_CAPACITY when the gates need more than one launch can carry (48 rounds,
 * 1280 matrix elements): the caller splits the run (same tile) and calls again.
 * TCB200_ERR_UNSUPPORTED: dense gate wider than 3 bits, diagonal wider than 4, state below 4 bits.
 * Replaces tensorcircuit/cons.py:605-623 for the run, as tcb200_apply_pass_host does.
 */
int tcb200_apply_gate_pass(void* state, int nbits, int dtype, int nops, const int* ops_k,
                           const int* ops_bits, const double* ops_mats, int n_hi,
                           const int* tile_hi, int64_t batch, double* info8, void* stream);

/*
 * The same gate pass for backend.vmap (tensorcircuit/backends/jax_backend.py:718-730): gate i
 * carries either one matrix shared by the batch (ops_batched[i] == 0) or one per batch element
 * (ops_batched[i] != 0: ops_mats holds [batch][4^k] for it).  Classification uses the union of
 * the non-zero patterns over the batch, so one schedule serves every element.  The per-element
 * matrices travel through pinned staging into `workspace` (DEVICE,
 * tcb200_gate_pass_batched_workspace_bytes) on `stream`; in the production shape the pass is then cut
 * into chunks of as many batch elements as fit 60 KiB of matrices, each chunk is copied device to
 * device into the library's __constant__ array and batch element b (grid.y) reads its matrices at a
 * warp-uniform offset of the constant bank (uniform-register operands, like the unbatched kernel).
 * The constant array is one per process: concurrent batched passes from several streams or host
 * threads are serialised by the library.  TCB200_CBANK=0 (and states of one tile or less) read the
 * matrices from `workspace` directly.
 */
size_t tcb200_gate_pass_batched_workspace_bytes(int dtype, int64_t batch);
int tcb200_apply_gate_pass_batched(void* state, int nbits, int dtype, int nops, const int* ops_k,
                                   const int* ops_bits, const double* ops_mats, const int* ops_batched,
                                   int n_hi, const int* tile_hi, int64_t batch, void* workspace,
                                   size_t ws_bytes, double* info8, void* stream);

/* Host-only dry run of the gate-pass scheduler (no device work): fills info8 as above. */
int tcb200_gate_pass_info(int nbits, int dtype, int nops, const int* ops_k, const int* ops_bits,
                          const double* ops_mats, int n_hi, const int* tile_hi, double* info8);

/* Geometry the pass planner needs: log2 of the tile size (amplitudes) used by
 * tcb200_apply_pass / tcb200_apply_pass_host / tcb200_apply_rpass_host for `dtype`. */
int tcb200_pass_tile_bits(int dtype);

/*
 * sum_r |psi_r|^2 per batch element -> out_dev[batch] (DEVICE, double).  Deterministic
 * (fixed reduction order).  workspace: tcb200_reduce_workspace_bytes(nbits, batch).
 */
int tcb200_norm2(const void* state, int nbits, int dtype, int64_t batch, double* out_dev,
                 void* workspace, size_t ws_bytes, void* stream);
size_t tcb200_reduce_workspace_bytes(int nbits, int64_t batch);

/*
 * p_r = |psi_r|^2 into a real DEVICE array of the matching real dtype (float / double):
 * BaseCircuit.probability (tensorcircuit/basecircuit.py:510-523).
 */
int tcb200_probability(const void* state, int nbits, int dtype, void* prob_dev, int64_t batch,
                       void* stream);

/*
 * Pauli-string expectations  <psi| P_t |psi>, t < nterms <= TCB200_MAX_TERMS, for every
 * batch element, in one read of the state:
 *     sum_r conj(psi_r) (-1)^{popc(r & sign[t])} (-i)^{ny[t]} psi_{r ^ flip[t]}
 * (closed form of tensorcircuit/quantum.py:1461-1482; replaces expectation_before +
 * contractor, tensorcircuit/basecircuit.py:267-319, circuit.py:914-990, once per term).
 *
 *   flip/sign/ny   HOST arrays (amplitude-index bit masks; ny = number of Y factors)
 *   tile_hi[n_hi]  HOST ascending bit positions gathered into the tile; every bit of every
 *                  flip mask must be below the contiguous low part or in tile_hi
 *   out_dev        DEVICE double [batch][nterms][2] (re, im); not normalised
 *   workspace      tcb200_expect_workspace_bytes(nbits, batch)
 * Accumulation is float64 for both dtypes, reduction order fixed (run-to-run deterministic).
 */
int tcb200_expect_pauli(const void* state, int nbits, int dtype, int nterms,
                        const uint64_t* flip, const uint64_t* sign, const int* ny, int n_hi,
                        const int* tile_hi, double* out_dev, int64_t batch, void* workspace,
                        size_t ws_bytes, void* stream);
size_t tcb200_expect_workspace_bytes(int nbits, int64_t batch);
/* log2 tile size (amplitudes) used by tcb200_expect_pauli for `dtype`. */
int tcb200_expect_tile_bits(int dtype);

/*
 * CDF sampler driven by caller-supplied uniforms, the rule of
 * ExtendedBackend.probability_sample (tensorcircuit/backends/abstract_backend.py:1145-1157):
 *     r = total * (1 - u);  index = first i with CDF[i] >= r   (searchsorted side="left")
 * evaluated with a two-level float64 CDF (block sums + in-block scan); the 2^n-entry p / CDF
 * arrays of the reference are never materialised.
 *
 *   uniforms_dev  DEVICE double[shots], values in [0, 1)
 *   out_idx_dev   DEVICE int64[shots]
 *   total_dev     DEVICE double[1] (optional, may be NULL): receives sum |psi|^2
 *   workspace     tcb200_sample_workspace_bytes(nbits)
 * The state must be a single vector (batch 1).  `cdf_offset`/`cdf_total` support a state
 * sharded over ranks: pass this shard's exclusive prefix of the per-shard masses and the
 * global mass (or 0 and a negative total for "single shard"); shots whose r falls outside
 * this shard's interval get index -1.
 */
int tcb200_sample(const void* state, int nbits, int dtype, const double* uniforms_dev,
                  int64_t shots, int64_t* out_idx_dev, double* total_dev, double cdf_offset,
                  double cdf_total, void* workspace, size_t ws_bytes, void* stream);
size_t tcb200_sample_workspace_bytes(int nbits);

/*
 * Host-buffer convenience entry point (what a reference-side binding would call for the whole
 * path): uploads nothing but the small per-block matrices, runs `npasses` dense blocks on the
 * device-resident state, then -- if shots > 0 -- draws samples for HOST uniforms into a HOST
 * index buffer.  Blocking (synchronises `stream` before returning).
 *   pass_k[npasses], pass_bits[sum k], pass_mats[sum 4^k complex128] : HOST
 */
int tcb200_run_circuit_host(void* state, int nbits, int dtype, int init_zero, int npasses,
                            const int* pass_k, const int* pass_bits, const double* pass_mats,
                            int64_t shots, const double* uniforms_host, int64_t* out_idx_host,
                            void* workspace, size_t ws_bytes, void* stream);

/* Single-flip Pauli strings -- exactly one X or Y, any number of Z's: the transverse-field /
 * kinetic half of TFIM- and Heisenberg-type Hamiltonians (quantum.py:1461-1482 with a one-bit flip
 * mask).  Up to tcb200_expect_single_flip_max_terms() = 12 strings per read of the state (12
 * distinct flip bits, one string per flip bit and launch), all flip bits inside the tile {low bits} U
 * {tile_hi} (<= 9 gathered bits): the pair sums are formed in registers, 16 amplitudes per
 * shared-memory access, instead of one partner load per amplitude and string.
 *   flip_bit[t]  amplitude-index bit of the X / Y;  sign[t]: bits carrying Z or Y;  ny[t]: 0 or 1
 *   out_dev      DEVICE double [batch][nterms][2], as tcb200_expect_pauli
 * The state must be larger than one 64 KiB tile (14 bits complex64 / 13 bits complex128); smaller
 * states go through tcb200_expect_pauli.  Float64 accumulation across tiles, fixed order. */
int tcb200_expect_single_flip_max_terms(void);
size_t tcb200_expect_single_flip_workspace_bytes(int nbits, int64_t batch);
int tcb200_expect_single_flip(const void* state, int nbits, int dtype, int nterms, const int* flip_bit,
                              const uint64_t* sign, const int* ny, int n_hi, const int* tile_hi,
                              double* out_dev, int64_t batch, void* workspace, size_t ws_bytes, void* stream);

/* Diagonal Pauli strings (only I and Z): expectation_ps(z=[...]) / the cost function of an Ising
 * or QAOA Hamiltonian, quantum.py:1461-1482 with an empty flip mask.  Up to
 * tcb200_expect_z_max_terms(dtype) strings (32 complex64 / 16 complex128) are evaluated in ONE
 * streaming read of the state -- no tile, no shared memory, one FFMA per (amplitude, string) -- against
 * 8 strings per read in tcb200_expect_pauli.  sign[t]: amplitude-index bits carrying a Z.
 * out_dev: [batch][nterms][2] doubles (imaginary parts are 0), device memory.  The state must have at
 * least tcb200_expect_z_min_bits(dtype) bits (13 / 12); smaller states go through tcb200_expect_pauli. */
int tcb200_expect_z_max_terms(int dtype);
int tcb200_expect_z_min_bits(int dtype);
size_t tcb200_expect_z_workspace_bytes(int nbits, int64_t batch);
int tcb200_expect_z(const void* state, int nbits, int dtype, int nterms, const uint64_t* sign,
                    double* out_dev, int64_t batch, void* workspace, size_t ws_bytes, void* stream);

/* Probability mass of a partial measurement record: sum of |psi_e|^2 over the amplitudes with
 * (e & mask) == value.  This is the reduced-density element rho[0,0] that measure_jit /
 * perfect_sampling contract qubit by qubit (basecircuit.py:359-443): with `mask` = the bits measured
 * so far plus the current one and `value` = their outcomes (current bit 0).  One streaming read,
 * float64 accumulation, fixed-order reduction.  out_dev: one double (device). */
size_t tcb200_masked_norm2_workspace_bytes(void);
int tcb200_masked_norm2(const void* state, int nbits, int dtype, uint64_t mask, uint64_t value, double* out_dev,
                        void* workspace, size_t ws_bytes, void* stream);

/* Readout error (basecircuit.py:587-596, 760-803): the reference maps p = |psi|^2 through the tensor
 * product of per-qubit 2x2 stochastic matrices before it samples.  Here the probabilities are kept
 * in a complex buffer of the state's dtype so that the existing kernels do all the work:
 *   mode 0: out_e = (|in_e|^2, 0)            -- then the readout matrices are applied as ordinary
 *                                               (real, non-unitary) 1-bit blocks with tcb200_apply_*
 *   mode 1: out_e = (sqrt(max(Re in_e,0)),0) -- afterwards |out_e|^2 = p'_e, so tcb200_sample and
 *                                               tcb200_expect_z read the noisy distribution
 * in and out may be the same buffer. */
int tcb200_probability_state(const void* in, void* out, int nbits, int dtype, int mode, void* stream);

/* <psi| H |psi> for a sparse operator kept on the DEVICE in COO form (rows / cols: int64 amplitude
 * indices in the reference's basis order, vals: complex128), one value per batch element.
 * Replaces templates/measurements.py:173-188 (sparse_expectation: backend.sparse_dense_matmul +
 * adjoint(state) @ tmp) and the dense branch of operator_expectation (measurements.py:156-170).
 * Pauli-sum Hamiltonians do not take this route: quantum.PauliStringSum2COO keeps the strings and
 * they go through tcb200_expect_*.  out_dev: DEVICE double [batch][2] (re, im).  Indices must be
 * below 2^nbits (checked by the caller when the operator is uploaded).  Float64 accumulation,
 * fixed-order reduction. */
size_t tcb200_coo_expectation_workspace_bytes(int64_t nnz, int64_t batch);
int tcb200_coo_expectation(const void* state, int nbits, int dtype, int64_t nnz, const int64_t* rows_dev, const int64_t* cols_dev,
                           const void* vals_dev, double* out_dev, int64_t batch, void* workspace, size_t ws_bytes, void* stream);

/* dst = sum_t coef_t P_t src for Pauli strings P_t given as (flip, sign) masks, coef_t = w_t (-i)^{ny_t}
 * (HOST arrays; coef: [nterms][2] doubles): row r of quantum.py:1461-1482,
 *     dst_r = sum_t coef_t (-1)^{popc(r & sign_t)} src_{r ^ flip_t}.
 * This is lambda = H psi, the seed of the adjoint-state gradient sweep that stands in for the
 * backend autodiff of backends/jax_backend.py:668-776.  src and dst are distinct device buffers of
 * `batch` states. */
size_t tcb200_apply_pauli_sum_workspace_bytes(int nterms);
int tcb200_apply_pauli_sum(const void* src, void* dst, int nbits, int dtype, int nterms, const uint64_t* flip, const uint64_t* sign,
                           const double* coef, int64_t batch, void* workspace, size_t ws_bytes, void* stream);

/* dst = coef * H src  (accumulate == 0)  or  dst += coef * H src  for a sparse operator in CSR form on the DEVICE
 * (indptr: int64 [2^nbits + 1], indices: int64 [nnz], vals: complex128 [nnz]); one thread per row, the entries of a
 * row are summed in storage order.  This is lambda = H psi for a generic sparse Hamiltonian -- the seed of the
 * adjoint sweep when the loss goes through templates/measurements.py:173-188 (sparse_expectation) instead of Pauli
 * strings.  src and dst are distinct device buffers of one state each. */
int tcb200_csr_matvec(const void* src, void* dst, int nbits, int dtype, const int64_t* indptr_dev, const int64_t* indices_dev,
                      const void* vals_dev, double coef_re, double coef_im, int accumulate, void* stream);

/* <bra| G_j |ket> for up to tcb200_transition_local_max_ops() local operators G_j (1 or 2 bits each;
 * ops_bits ascending amplitude-index bits, ops_mats row-major complex128 with matrix index bit i <->
 * bit i of the operator, HOST arrays) between two device states, all in one launch.  In the adjoint
 * sweep G_j = dM_j/dtheta M_j^+, bra = lambda_j, ket = psi_j, and 2 Re of the result is the
 * derivative of the energy through gate j (jax_backend.py:668-776).  out_dev: DEVICE double [nops][2].
 * The operators are packed into groups of four amplitude-index bits; a thread loads the 16 + 16 amplitudes
 * that differ in those bits once and evaluates every operator of the group from registers (one read of the
 * two states per bit group instead of one per operator), and groups whose bits all lie inside an L2-sized
 * chunk (2^22 complex64 / 2^21 complex128 amplitudes) are evaluated chunk by chunk. */
int tcb200_transition_local_max_ops(void);
size_t tcb200_transition_local_workspace_bytes(int nops, int nbits, int dtype);
int tcb200_transition_local(const void* bra, const void* ket, int nbits, int dtype, int nops, const int* ops_k, const int* ops_bits,
                            const double* ops_mats, double* out_dev, void* workspace, size_t ws_bytes, void* stream);

/* Number of kernel launches issued by this process through the library so far. */
int64_t tcb200_launch_count(void);

/* How many of those were the persistent TMA pipeline (tpass_kernel) rather than the LDGSTS-staged
 * cpass_kernel.  The pipeline is OPT-IN (TCB200_TMA=1: it measured slower than the staged kernels,
 * profiles/README.md); with it enabled tcb200_apply_pass_host uses it whenever the pass has the
 * production 64 KiB tile, the state is larger than one tile and the driver exports
 * cuTensorMapEncodeTiled (TCB200_TMA_STRICT=1 turns a failed tensor-map encode into an error
 * instead of a fallback). */
int64_t tcb200_tma_pass_count(void);

#ifdef __cplusplus
}
#endif
#endif /* TCB200_H */
