/*
 * flatdd_b200.h — C-ABI of the B200-native FlatDD array-phase hot path.
 *
 * The reference (IDEA-CUHK/FlatDD) has no FFI layer: its boundary for this path is a set of
 * C++ member functions that take a decision-diagram edge plus raw `double*` arrays.  Every
 * entry point below names the reference interface it replaces (file:line under the reference
 * tree).  Signatures use only plain pointers and sizes, so cgo / JNI / ctypes / a C++ host can
 * bind them alike.  All functions return 0 on success and a negative FDD_ERR_* code on
 * failure; `fdd_last_error()` gives the message of the last failure on the calling thread.
 *
 * Index convention (reference: include/dd/SwitchPackage.hpp:3597):
 * qubit q is bit q of the amplitude index, qubit 0 = least significant bit, and the top
 * decision-diagram node carries level n-1.
 */
#ifndef FLATDD_B200_H
#define FLATDD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDD_OK 0
#define FDD_ERR_INVALID (-1)  /* malformed table / bad argument                      */
#define FDD_ERR_CUDA (-2)     /* a CUDA runtime call failed (message has the detail) */
#define FDD_ERR_TOO_DENSE (-3) /* gate has more non-zeros per row than one launch supports; split it */
#define FDD_ERR_STATE (-4)    /* call not valid in the current state of the context  */
#define FDD_ERR_COMM (-5)     /* multi-GPU exchange failed                            */

/* Terminal / zero child index inside the flat tables. */
#define FDD_TERMINAL (-1)

/*
 * Flat vector DD (the state DD, reference types vNode/vEdge: include/dd/Node.hpp:17-27,
 * include/dd/Edge.hpp:12-14).  Node i has level `level[i]`, successor `child[2*i+b]` for
 * bit value b of its qubit and edge weight (weight[2*(2*i+b)], weight[2*(2*i+b)+1]) = (re, im).
 * A zero edge has child == FDD_TERMINAL and weight exactly (0,0) (the reference's
 * `w.exactlyZero()` predicate).  Children are exactly one level below their parent
 * (reference makeDDNode assert, include/dd/SwitchPackage.hpp:1251); level-0 nodes point to
 * FDD_TERMINAL.  Nodes may come in any order.
 */
typedef struct fdd_vecdd {
    int32_t n_qubits;
    int32_t n_nodes;
    int32_t root;          /* index of the root node (level n_qubits-1) */
    int32_t reserved;
    double root_weight[2]; /* (re, im) of the root edge */
    const int32_t* level;  /* [n_nodes]   */
    const int32_t* child;  /* [2*n_nodes] */
    const double* weight;  /* [4*n_nodes] */
} fdd_vecdd;

/*
 * Flat matrix DD (a gate, reference types mNode/mEdge: include/dd/Node.hpp:35-83).
 * Successor index is 2*row_bit + col_bit exactly like `mNode::e[4]`; the DD is full depth
 * (every level n-1..0 has a node on every non-zero path; identity levels are explicit),
 * which is what makeGateDD / makeIdent / multiply produce
 * (include/dd/SwitchPackage.hpp:694-775, 2922-2960) and what DDArrMultiplyRecurIP asserts (:2251).
 */
typedef struct fdd_matdd {
    int32_t n_qubits;
    int32_t n_nodes;
    int32_t root;
    int32_t reserved;
    double root_weight[2];
    const int32_t* level;  /* [n_nodes]   */
    const int32_t* child;  /* [4*n_nodes] */
    const double* weight;  /* [8*n_nodes] */
} fdd_matdd;

typedef struct fdd_ctx fdd_ctx;   /* one simulator state living on one GPU (or one shard of it) */
typedef struct fdd_gate fdd_gate; /* a gate compiled for the device (tables resident in HBM)     */

/* ---- library ------------------------------------------------------------------------- */
const char* fdd_version(void);
const char* fdd_last_error(void);
/* Number of CUDA devices visible; fails (FDD_ERR_CUDA) when there is none: there is no CPU fallback. */
int fdd_device_count(int* count);

/* ---- context = the state buffers -------------------------------------------------------
 * Replaces the SwitchSimulator constructor / destructor, which malloc and zero four
 * double[2^n] arrays (include/SwitchSimulator.hpp:26-49), and the double-buffer protocol of
 * singleShot (src/SwitchSimulator.cpp:143-151, 398-403).  The device state is two
 * interleaved complex<double> buffers of 2^n_local amplitudes (ping-pong); gate kernels
 * overwrite their destination, so there is no memset contract.
 *
 * world_size > 1 creates one shard of a state distributed over `world_size` contexts (one
 * per process / GPU): rank r holds the amplitudes whose top log2(world_size) *physical* index
 * bits equal r.  world_size must be a power of two. */
int fdd_create(int n_qubits, int device, fdd_ctx** out);
int fdd_create_sharded(int n_qubits, int device, int rank, int world_size, fdd_ctx** out);
int fdd_destroy(fdd_ctx* ctx);
int fdd_n_qubits(const fdd_ctx* ctx);
int fdd_n_local_qubits(const fdd_ctx* ctx);
int fdd_synchronize(fdd_ctx* ctx);
/* Tunables for experiments: "dmavm_variant" (2 tile kernel where the gate allows, 1 / 0 walk kernels, 9 chunk kernel),
 * "warps_per_cta", "ctas_per_sm", "prefetch", "tile_mode", "dense_slots", "dmma", "flat_table", "pdl", "context_table", "exchange_unroll", "exchange_ctas_per_sm";
 * the dense-block path: "block_kernel" (0: older kernels only), "block_tile_bits", "block_max_tile_bits", "block_max_per_pass" (blocks that may share a pass),
 * "block_buffers", "block_ws", "block_tables_shared" (matrix tables of a shared pass in shared memory), "block_reorder" (blocks move up over gates they
 * commute with to share a pass; 0 keeps the caller's order). */
int fdd_set_option(fdd_ctx* ctx, const char* key, long value);
/* Reads a tunable back, or a launch counter: "launches", "tensor_core_launches" (DMAVM launches that ran the
 * FP64 tensor-core path), "flat_table_launches", "context_table_launches", "exchanges", "block_launches" (passes of the dense-block
 * kernel), "blocks_applied" (fused gates those passes applied). */
int fdd_get_option(const fdd_ctx* ctx, const char* key, long* value);

/* ---- multi-GPU (one process per GPU) ----------------------------------
 * The reference has no distributed code; the partition is its thread partition
 * (seg_size = nDim / n_thread, include/dd/SwitchPackage.hpp:2146) lifted to devices.
 * fdd_comm_unique_id fills a 128-byte NCCL unique id on rank 0; the host broadcasts it
 * (torch.distributed / MPI / a file) and every rank calls fdd_comm_init with it.  fdd_comm_init
 * also maps the state buffers of all peers through CUDA IPC (NVLink peer access).
 *
 * Gates are applied shard-locally: a gate may be non-diagonal only on LOCAL physical qubits
 * (controls and diagonal gates on global qubits are fine).  When a gate needs a global qubit the
 * host first calls fdd_exchange_qubits(global physical bit, local physical bit) on every rank:
 * SWAP of the two index bits = each rank trades half of its shard (8 * 2^n / G bytes each way)
 * with the rank that differs in that global bit.  method 0: one kernel that reads the partner's
 * half straight over NVLink peer memory; method 1: NCCL send/recv.  The host tracks the
 * logical->physical qubit map (the reference's qc::Permutation, include/Permutation.hpp:9-25)
 * and builds later gate DDs in physical order (dd::getDD(op, dd, permutation),
 * include/dd/Operations.hpp:591-678); the context mirrors that map for fdd_get_permutation. */
int fdd_comm_unique_id(void* id128);
int fdd_comm_init(fdd_ctx* ctx, const void* id128);
int fdd_exchange_qubits(fdd_ctx* ctx, int global_physical_bit, int local_physical_bit, int method);
/* A stretch of the schedule followed by the exchange the next gate needs, as if by fdd_apply_many / fdd_gate_apply_many and then
 * fdd_exchange_qubits(..., 0) (the executor loop src/SwitchSimulator.cpp:386-412 up to a point where the partition
 * include/dd/SwitchPackage.hpp:2146 has to change).  In one call the library can FUSE the two: when the stretch ends in a pass of
 * the tile-resident kernel and the local bit is >= 5, that pass stores the half of its result that changes owner straight into
 * the partner shard's buffer (peer memory over NVLink, per 512-byte segment) and the exchange costs no pass of its own
 * ("block_fuse_exchange" = 0: never fuse; "fused_exchanges" counts them).  Every rank makes the same call. */
int fdd_apply_many_exchange(fdd_ctx* ctx, const fdd_matdd* gates, int count, int global_physical_bit, int local_physical_bit);
int fdd_gate_apply_many_exchange(fdd_ctx* ctx, const fdd_gate* const* gates, int count, int global_physical_bit, int local_physical_bit);
/* Bookkeeping only: the logical qubits sitting at two physical bits trade names (an uncontrolled
 * SWAP gate absorbed into the layout, include/dd/Operations.hpp:611-620).  No data moves. */
int fdd_relabel_qubits(fdd_ctx* ctx, int physical_bit_a, int physical_bit_b);
int fdd_barrier(fdd_ctx* ctx);

/* ---- DD -> array conversion -------------------------------------------------------------
 * Replaces SwitchSimulator::getVectorFromDDSwitch1 (include/SwitchSimulator.hpp:169-352) and
 * getVectorFromDD (:66-82) + getValueByPathPar (include/dd/SwitchPackage.hpp:3605-3634):
 * amplitude(i) = w_root * prod_{v=n-1..0} w(node_v.e[bit_v(i)]), multiplied root first, leaf
 * last, with un-fused IEEE multiplies/adds (bit-identical to the reference's serial walk).
 * Every amplitude is written (zeros included).  The result becomes the current state. */
int fdd_convert(fdd_ctx* ctx, const fdd_vecdd* dd);

/* ---- DMAVM: gate (matrix DD) x state (array) --------------------------------------------
 * Replaces SwitchPackage::DDArrMultiplyIP (include/dd/SwitchPackage.hpp:1897-1925,
 * 2132-2261) and DDArrMultiplyOP (:1928-1956, 2265-2508): z[r] = sum_c M[r][c] * y[c] with
 * M[r][c] = w_root * prod_v w(node_v.e[2*bit_v(r)+bit_v(c)]), summed over c ascending.  The
 * result becomes the current state (the context flips its ping-pong buffers, which replaces
 * the caller-side memset of the old buffer, src/SwitchSimulator.cpp:150-151). */
int fdd_apply(fdd_ctx* ctx, const fdd_matdd* gate);
/* `count` gates in order, as if by `count` calls of fdd_apply (the executor loop of src/SwitchSimulator.cpp:386-412 in one
 * call).  Handing the library the whole stretch lets it keep the state tile-resident across gates: consecutive gates that are
 * dense blocks (at most four non-diagonal qubits, at most ten qubits they depend on diagonally) are applied in ONE pass over
 * the state while their target qubits fit one shared-memory tile — 32 bytes of HBM traffic per amplitude for the group
 * instead of per gate.  A block may be applied earlier than its place in the list when it commutes with every gate it
 * passes (no target of one among the targets or context qubits of the other), so that it can share a pass: the product of
 * the gates is the same, the order of commuting factors is not ("block_reorder" = 0 keeps the list order). */
int fdd_apply_many(fdd_ctx* ctx, const fdd_matdd* gates, int count);
/* Host only, no device needed: the gate as a DENSE BLOCK (north_star: "gate DDs are flattened to dense 2^k x 2^k blocks for
 * the fused qubit set") — `targets`: its non-diagonal qubits (ascending, at most 4), `controls`: the other qubits its matrix
 * depends on (diagonally: controls, phases; ascending, at most max_controls <= 10), `matrices`: [2^n_controls][2^k][2^k]
 * complex (re, im) entries, row-major, index bit i <-> targets[i] / controls[i]; every entry is the product of the edge weights
 * along its DD path, root first (the reference's order, include/dd/SwitchPackage.hpp:2221-2236).
 * FDD_ERR_TOO_DENSE: the gate is not such a block.  FDD_ERR_INVALID with n_targets / n_controls filled in: the buffer is too
 * small (2 * 4^k * 2^c doubles are needed).  The fusion pass of the host driver (flatdd_b200/host/block_fusion.hpp) uses this
 * once per distinct circuit operation. */
int fdd_block_from_matdd(const fdd_matdd* gate, int max_controls, int32_t* n_targets, int32_t* targets, int32_t* n_controls,
                         int32_t* controls, double* matrices, size_t capacity_doubles);
/* Same in two steps, so a schedule can be compiled once and replayed. */
int fdd_gate_compile(fdd_ctx* ctx, const fdd_matdd* gate, fdd_gate** out);
int fdd_gate_apply(fdd_ctx* ctx, const fdd_gate* gate);
/* Applies `count` compiled gates back to back (one call per schedule segment instead of one per gate). */
int fdd_gate_apply_many(fdd_ctx* ctx, const fdd_gate* const* gates, int count);
int fdd_gate_free(fdd_gate* gate);
/* Facts about a compiled gate: key in {"kind", "max_paths", "max_sub_k", "upper_nodes", "upper_depth", "sub_tables",
 * "nnz_per_row_max", "nnz", "top_level", "stack_cap", "tileable", "uniform", "sub_tile_bits", "non_diag_upper",
 * "non_diag_mask", "tile_mask", "fill_mask"}; -1 for an unknown key. */
long fdd_gate_info(const fdd_gate* gate, const char* key);

/* Literal drop-in for one DDArrMultiplyIP call on HOST arrays (SoA re/im, nDim = 2^n):
 * uploads y, multiplies, downloads z (z is overwritten, not accumulated). */
int fdd_ddarr_multiply(const fdd_matdd* gate, const double* y_real, const double* y_imag,
                       double* z_real, double* z_imag, size_t n_dim, int device);

/* ---- cost model (host only) ---------------------------------------------------------------
 * fdd_mac_count = DMAVMACCountIP of the root (include/dd/SwitchPackage.hpp:3285-3311): the
 * number of non-zero root-to-terminal paths = nnz of the matrix.
 * fdd_cost_ip   = DMAVMACStatIP (:3012-3017) = nnz / 2^n_thread_exp.
 * fdd_cost_op1  = DMAVMACStatOP1 (:3006-3010, 3202-3283).
 * fdd_cost_gpu  = estimated device nanoseconds of one launch under the HBM / fp64 roofline
 *                 (the re-tuned cost that drives GPU-aware fusion). */
int fdd_mac_count(const fdd_matdd* gate, uint64_t* nnz);
int fdd_cost_ip(const fdd_matdd* gate, unsigned n_thread_exp, uint64_t* cost);
int fdd_cost_op1(const fdd_matdd* gate, unsigned n_thread_exp, uint64_t* cost);
int fdd_cost_gpu(const fdd_matdd* gate, double hbm_gbs, double fp64_gflops, double* nanoseconds);
/* Host-only structural facts of a gate, same keys as fdd_gate_info (no device needed). */
int fdd_matdd_info(const fdd_matdd* gate, const char* key, long* value);

/* ---- state access ---------------------------------------------------------------------------
 * fdd_get_state replaces SwitchSimulator::getVector (include/SwitchSimulator.hpp:55-63): it
 * synchronises and copies the current state into host SoA arrays of 2^n_local doubles each
 * (the reference's state_real / state_imag layout).  Sharded contexts return their shard in
 * logical order only after fdd_canonicalize (which undoes the global/local qubit remap). */
int fdd_get_state(fdd_ctx* ctx, double* real, double* imag);
int fdd_set_state(fdd_ctx* ctx, const double* real, const double* imag);
/* |0...0> without a DD (reference: makeZeroState + conversion). */
int fdd_set_zero_state(fdd_ctx* ctx);
/* Read `count` amplitudes starting at local index `first` (interleaved re,im pairs). */
int fdd_get_amplitudes(fdd_ctx* ctx, uint64_t first, uint64_t count, double* interleaved);
/* Read the amplitudes at `count` arbitrary LOCAL indices (interleaved re,im pairs), gathered on the device: the sampled
 * comparison of states that are too large to download (reference getVector, include/SwitchSimulator.hpp:55-63, read at a
 * few indices).  Sharded states: local index = global index without the top log2(world) bits, valid after fdd_canonicalize. */
int fdd_get_amplitudes_at(fdd_ctx* ctx, const uint64_t* local_indices, uint64_t count, double* interleaved);
/* sum |amp|^2 over the local shard, computed on the device. */
int fdd_norm2(fdd_ctx* ctx, double* out);
/* Measurement sampling on the device (the reference skips measurements,
 * src/SwitchSimulator.cpp:108-113, so there is no parity target): draws `n_shots` basis states of the local shard
 * with probability |amplitude|^2 / (shard norm) and writes their LOCAL PHYSICAL indices.  Deterministic in `seed`.
 * Sharded states: draw the rank of every shot from the shard norms (fdd_norm2) first, and map physical to logical
 * bits with fdd_get_permutation. */
int fdd_sample(fdd_ctx* ctx, uint64_t n_shots, uint64_t seed, uint64_t* local_indices);
/* Device pointer of the current state (interleaved complex<double>), for zero-copy hosts. */
int fdd_state_device_ptr(fdd_ctx* ctx, void** ptr);
/* Sharded contexts: logical->physical qubit map (length n_qubits) and its undo. */
int fdd_get_permutation(const fdd_ctx* ctx, int32_t* logical_to_physical);
int fdd_canonicalize(fdd_ctx* ctx);

/* ---- measurement hooks ----------------------------------------------------------------------
 * Device time (CUDA events on the context's stream) of the last convert / gate launch, and
 * the number of kernels the library has launched on this context since creation. */
int fdd_last_kernel_ms(fdd_ctx* ctx, float* ms);
int fdd_set_timing(fdd_ctx* ctx, int enabled);
uint64_t fdd_launch_count(const fdd_ctx* ctx);
/* The CUDA stream (cudaStream_t) kernels are launched on. */
int fdd_stream(fdd_ctx* ctx, void** stream);

#ifdef __cplusplus
}
#endif
#endif /* FLATDD_B200_H */
