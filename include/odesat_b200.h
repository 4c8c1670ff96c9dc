/* odesat_b200.h — C ABI of the B200-native digital-memcomputing (DMM) ODE integrator.
 *
 * Drop-in boundary for ONE hot path of AHartNtkn/odesat: the public functions of the Rust
 * module `odesat::system` (reference `src/lib.rs:3`, `src/system.rs`).  The reference has no
 * FFI layer; the entry points below are what a ~100-line `ffi.rs` would bind (see
 * INTEGRATION.md and odesat_b200/host/ffi.rs).  Every function cites the reference interface it
 * replaces as `file:line` relative to the reference root.
 *
 * Conventions
 *  - plain pointers and sizes only; all buffers are HOST memory owned by the caller unless a
 *    parameter is documented as a device pointer.  Host state layout is the reference's: one
 *    contiguous vector per replica, `v[R][N]`, `xs[R][M]`, `xl[R][M]` (R = 1 for single-state
 *    calls).  Element type is `double` for the unsuffixed / `_f64` functions (the reference
 *    is f64-only) and `float` for `_f32`.
 *  - every function returns an odesat_status (0 = ok) and never throws, aborts or exits.
 *    `odesat_last_error()` returns a thread-local message for the last non-zero status.
 *  - the reference's `Option<T>` arguments are passed as sentinels: NaN for floating-point
 *    options, a negative value for `Option<usize>`.
 *  - handles are not thread-safe.  A formula handle lives on the CUDA device that was current when it was
 *    created; the replica-batch calls (odesat_simulate_batch*, odesat_simulate_inter) drive `params->n_gpus`
 *    devices of the process from that one handle (contiguous replica shards, the formula replicated on each
 *    device, one stream per shard, no data-path collective — see odesat_params::n_gpus).  Calls are
 *    synchronous: results are on the host when they return.  The caller's current device is restored.
 *  - there is NO CPU fallback.  Without a CUDA device every compute entry point returns
 *    ODESAT_ECUDA.
 */
#ifndef ODESAT_B200_H
#define ODESAT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ODESAT_B200_ABI_VERSION 2   /* 2: odesat_params gained n_gpus / sub_batches; async batch entry points */

typedef enum odesat_status {
    ODESAT_OK = 0,
    ODESAT_EINVAL = 1,       /* bad argument (the reference would panic / index out of bounds) */
    ODESAT_ECUDA = 2,        /* CUDA runtime failure or no device                              */
    ODESAT_ENOMEM = 3,
    ODESAT_EUNSUPPORTED = 4  /* requested engine cannot run this formula                       */
} odesat_status;

typedef enum odesat_precision { ODESAT_F64 = 0, ODESAT_F32 = 1 } odesat_precision;

/* Which kernel family integrates a replica batch.
 *  GATHER: general path — any clause length, any N/M/R, fixed and adaptive steps; replica-major
 *          state [row][replica]; a clause phase materialises the per-literal contributions, a
 *          variable phase adds them over the variable→clause transpose in the reference's order
 *          (ascending clause, then literal position) ⇒ bit-identical sums.  Instances whose whole
 *          state fits in the shared memory of one SM (the reference's fixtures) run on a persistent
 *          one-CTA-per-replica kernel with the step loop inside the kernel.
 *  TILE  : throughput path — up to ≈ 13 000 variables (two f32 replicas / one f64 replica per CTA) or ≈ 27 000
 *          (one f32 replica per CTA, uniform 3-literal formulas): variables of a replica tile resident in shared
 *          memory, clauses streamed in conflict-free levels.  Fixed steps on any formula: uniform 3-literal clauses
 *          with distinct variables run on the packed kernels (tile_engine.cuh, tile_ws.cuh); ragged clause lengths,
 *          unit / empty clauses and variables repeated inside a clause on k_tile_ragged (tile_ragged.cuh).  Adaptive
 *          steps (system.rs:111-139, per-replica dt): k_tile_adaptive (tile_adaptive.cuh), uniform and ragged formulas
 *          (except formulas with 4..32-literal clauses under the EXACT schedule).  Larger instances can be forced onto a thread-block cluster (rows in distributed
 *          shared memory) by requesting TILE explicitly; AUTO does not, because the general engine is faster there
 *          (DESIGN.md §5b).
 *  SLAB  : throughput path for fixed steps on uniform 3-literal formulas that do NOT fit in shared memory (up to
 *          2^20 variables): the gather engine's deterministic two-phase RHS (bit-identical sums) run by ONE
 *          persistent kernel per chunk of steps over slabs of 32 bytes of replicas, ordered so that a slab's
 *          per-literal contributions are consumed while they are still in L2 (csrc/slab_engine.cuh).
 *          Experimental: bit-identical, HBM traffic 13 GB per step against the gather engine's 22.7 GB at N = 50 000,
 *          but latency-bound and slower (4.8 ms against 3.56 ms per step), so AUTO never picks it.
 *  AUTO  : TILE when it applies and the batch has at least 8 replicas, else GATHER.  (Adaptive steps: TILE only where
 *          it has the adaptive kernel.  Formulas with clauses of more than three literals under the EXACT schedule:
 *          GATHER, which is as fast there — DESIGN.md §5d.) */
typedef enum odesat_engine { ODESAT_ENGINE_AUTO = 0, ODESAT_ENGINE_GATHER = 1, ODESAT_ENGINE_TILE = 2, ODESAT_ENGINE_SLAB = 3 } odesat_engine;

/* Clause schedule of the TILE engine.
 *  EXACT   : order-preserving levels (list scheduling) — each variable receives its clause
 *            contributions in ascending clause index, i.e. the reference's summation order
 *            (bit-identical dv).
 *  BALANCED: equal-size colour classes (fewer barriers); per-variable summation order differs
 *            from the reference's, results agree to rounding (deterministic run to run). */
typedef enum odesat_schedule { ODESAT_SCHED_EXACT = 0, ODESAT_SCHED_BALANCED = 1 } odesat_schedule;

typedef enum odesat_mode {
    ODESAT_MODE_BATCH = 0,   /* main.rs:254-323: independent replicas, each stops at its own flag;
                                winner = lowest replica whose thresholded state satisfies the CNF */
    ODESAT_MODE_INTER = 1    /* system.rs:241-359: all replicas step together, stop at the first
                                step on which any replica flags; winner = lowest flagged index   */
} odesat_mode;

/* Options of `simulate` / `simulate_inter` (system.rs:156-163, 241-248). */
typedef struct odesat_params {
    double tolerance;    /* Option<f64>: NaN → 1e-3            (system.rs:174)                 */
    double step_size;    /* Option<f64>: NaN → adaptive, dt0 = 0.01 (system.rs:190, 205)       */
    int64_t steps;       /* Option<usize>: < 0 → unbounded     (system.rs:191, 198)            */
    double learning_rate;/* Option<f64> zeta: NaN → density rule (system.rs:164-173)           */
    int32_t precision;   /* odesat_precision of the device arithmetic                           */
    int32_t engine;      /* odesat_engine                                                       */
    int32_t schedule;    /* odesat_schedule                                                     */
    int32_t chunk;       /* Euler steps between early-exit polls; <= 0 → 32 (1024 on the persistent
                            small-instance kernel)                                               */
    int32_t n_gpus;      /* replica batches only: CUDA devices to shard the replicas over (replica r of R goes
                            to shard r·G/R, main.rs:278-308 / system.rs:279-289 have no cross-replica data flow);
                            <= 0 → 1; more than odesat_device_count() → ODESAT_EINVAL.  The only exchange is
                            the early-exit word: every shard posts (first flagged step << 32 | replica) and its
                            count of unflagged replicas to pinned host memory after each chunk, the host takes
                            the minimum one chunk late while the next chunk is already running, and that chunk's
                            kernels read the previous key on the device and do nothing once it is set.        */
    int32_t sub_batches; /* shards per device with their own stream, so that the host→device copy of one
                            shard's initial state overlaps the integration of the previous one; <= 0 → auto
                            (one per 512 replicas of a device, at most 8, when v0 comes from the host; else 1) */
} odesat_params;

typedef struct odesat_formula odesat_formula;   /* device-resident CSR + variable→clause transpose */
typedef struct odesat_batch odesat_batch;       /* device-resident replica-major state of R replicas */

const char* odesat_last_error(void);
int odesat_abi_version(void);
/* Number of visible CUDA devices (0 when there is none); never fails. */
int odesat_device_count(void);

/* ---- formula: the boundary type `CNFFormula` (cnf.rs:53-57) flattened ------------------------
 * `lits[j]` = ±(index+1) over normalised variables 0..varnum-1 (negative = is_negated, cnf.rs:5-9);
 * clause m owns lits[clause_off[m] .. clause_off[m+1]).  Indices are validated here (the
 * reference panics on out-of-bounds at system.rs:48). */
int odesat_formula_create(int64_t varnum, int64_t n_clauses, const int64_t* clause_off,
                          const int32_t* lits, odesat_formula** out);
void odesat_formula_destroy(odesat_formula* f);
int odesat_formula_info(const odesat_formula* f, int64_t* varnum, int64_t* n_clauses,
                        int64_t* n_literals, int32_t* uniform_k);
/* zeta chosen by the density rule (system.rs:164-173). */
int odesat_formula_default_zeta(const odesat_formula* f, double* zeta);

/* ---- single-state mirrors of odesat::system (one replica, host buffers in and out) ---------- */

/* system.rs:362-372 init_short_term_memory(&CNFFormula) -> Array1<f64>; xs0[M]. */
int odesat_init_short_term_memory(const odesat_formula* f, double* xs0);
int odesat_init_short_term_memory_f32(const odesat_formula* f, float* xs0);

/* system.rs:25-31 compute_derivatives(&State, &mut State, &CNFFormula, zeta, &mut SlabState) -> bool */
int odesat_compute_derivatives(const odesat_formula* f, const double* v, const double* xs,
                               const double* xl, double zeta, double* dv, double* dxs, double* dxl,
                               int* allsat);
int odesat_compute_derivatives_f32(const odesat_formula* f, const float* v, const float* xs,
                                   const float* xl, double zeta, float* dv, float* dxs, float* dxl,
                                   int* allsat);

/* system.rs:93 update_state(&mut State, &State, dt, clause_nums); clause_nums = formula's M. */
int odesat_update_state(const odesat_formula* f, double* v, double* xs, double* xl,
                        const double* dv, const double* dxs, const double* dxl, double dt);
int odesat_update_state_f32(const odesat_formula* f, float* v, float* xs, float* xl,
                            const float* dv, const float* dxs, const float* dxl, double dt);

/* system.rs:101 max_error(&State, &State) -> f64 (NaN-ignoring max of |a-b| over v, xs, xl). */
int odesat_max_error(const odesat_formula* f, const double* av, const double* axs,
                     const double* axl, const double* bv, const double* bxs, const double* bxl,
                     double* err);
int odesat_max_error_f32(const odesat_formula* f, const float* av, const float* axs,
                         const float* axl, const float* bv, const float* bxs, const float* bxl,
                         double* err);

/* system.rs:141-148 euler_step_fixed(...) -> bool: state updated in place, *allsat is the
 * flag of the PRE-update state. */
int odesat_euler_step_fixed(const odesat_formula* f, double* v, double* xs, double* xl, double dt,
                            double zeta, int* allsat);
int odesat_euler_step_fixed_f32(const odesat_formula* f, float* v, float* xs, float* xl, double dt,
                                double zeta, int* allsat);

/* system.rs:111-119 euler_step(..., tolerance, &mut dt, zeta, ...) -> bool. */
int odesat_euler_step(const odesat_formula* f, double* v, double* xs, double* xl, double tolerance,
                      double* dt, double zeta, int* allsat);
int odesat_euler_step_f32(const odesat_formula* f, float* v, float* xs, float* xl, double tolerance,
                          double* dt, double zeta, int* allsat);

/* system.rs:156-163 simulate(&mut State, &CNFFormula, tolerance, step_size, steps, learning_rate)
 * -> Vec<bool>.  State is updated in place; assignment[i] = v[i] > 0 (system.rs:238).
 * steps_taken = loop iterations executed; *allsat = 1 when the loop ended on the flag;
 * *final_dt (may be NULL) = the adaptive step size at exit.  `params->precision` selects the
 * device arithmetic; host buffers are f64 for this entry point and f32 for `_f32`. */
int odesat_simulate(const odesat_formula* f, double* v, double* xs, double* xl,
                    const odesat_params* params, uint8_t* assignment, int64_t* steps_taken,
                    int* allsat, double* final_dt);
int odesat_simulate_f32(const odesat_formula* f, float* v, float* xs, float* xl,
                        const odesat_params* params, uint8_t* assignment, int64_t* steps_taken,
                        int* allsat, double* final_dt);

/* ---- replica batches: `batch` (main.rs:278-308) and `simulate_inter` (system.rs:241-248) ----
 * One call = upload R host states, integrate on the device with per-replica flags and early
 * exit, verify every thresholded state exactly (cnf.rs:246-264), return the winner.
 *  v/xs/xl      : host [R][N] / [R][M] / [R][M] initial states; any of them may be NULL, in
 *                 which case it is generated on the device (v0 from `seed` with the documented
 *                 counter-based generator, xs0 by init_short_term_memory, xl0 = 1).
 *  write_back   : non-zero → final states are copied back into v/xs/xl (the reference mutates
 *                 its `&mut` states); zero → they are left untouched.
 *  solved_step  : [R] first step index (0-based) whose pre-update state was all-satisfied, -1 if none
 *  verified     : [R] 1 when the replica's thresholded final state satisfies the CNF exactly
 *  winner       : BATCH: lowest verified replica; INTER: lowest replica flagged at the earliest
 *                 flagged step (system.rs:353); -1 when none (assignment then comes from
 *                 replica R-1 for BATCH — the last one the reference's loop ran — or 0 for INTER,
 *                 system.rs:357)
 *  assignment   : [N] thresholded state of the winner
 *  steps_run    : outer Euler steps executed by the device loop
 * INTER with fixed steps and write_back != 0 leaves EVERY replica after exactly steps_run steps, like the reference's
 * lock-step loop (system.rs:279-293): a chunk that ran past the winning step is replayed from a device-side snapshot
 * of the chunk's start.  Without write_back only the winner's state is defined (the others may be up to two chunks
 * ahead), and solved_step reports flags up to the winning step only.
 * R == 0: nothing runs; winner = -1, steps_run = 0, assignment zero-filled.
 * Adaptive INTER (params.step_size = NaN) runs the reference's loop literally: the replicas take their steps
 * one after the other inside an outer step and share ONE dt (system.rs:314), and the loop
 * stops right after the outer step in which the first replica flags — a sequential dependency replica →
 * replica, so it is correct but not fast (one CTA walks the replicas on instances that fit in shared memory;
 * otherwise the general engine steps one replica at a time). */
int odesat_simulate_batch(const odesat_formula* f, int64_t R, double* v, double* xs, double* xl,
                          uint64_t seed, int64_t replica_offset, const odesat_params* params,
                          int32_t mode, int32_t write_back, int64_t* solved_step,
                          uint8_t* verified, int64_t* winner, uint8_t* assignment,
                          int64_t* steps_run);
int odesat_simulate_batch_f32(const odesat_formula* f, int64_t R, float* v, float* xs, float* xl,
                              uint64_t seed, int64_t replica_offset, const odesat_params* params,
                              int32_t mode, int32_t write_back, int64_t* solved_step,
                              uint8_t* verified, int64_t* winner, uint8_t* assignment,
                              int64_t* steps_run);
/* system.rs:241-248 simulate_inter(&mut Vec<State>, ...) -> Vec<bool>: INTER mode of the above
 * with caller-supplied states, written back. */
int odesat_simulate_inter(const odesat_formula* f, int64_t R, double* v, double* xs, double* xl,
                          const odesat_params* params, uint8_t* assignment, int64_t* winner,
                          int64_t* steps_taken);

/* ---- stochastic local search: src/stoch.rs (a separate algorithm with the same data layout) ----
 * State (stoch.rs:8-12): v[N] bool as uint8, xl[M] uint64.  The reference draws its flips from an OS-seeded ThreadRng
 * (stoch.rs:68, :81 — unreproducible by construction); here r = 1 + mulhi64(SplitMix64(seed, replica, step, variable),
 * total), the same function as the CPU oracle.  A variable that occurs in no clause never flips (the reference panics
 * on the empty range 1..=0).
 *
 * stoch.rs:26-31 step(&mut State, &CNFFormula, &mut SlabState, &mut ThreadRng) -> bool: one step on a caller-owned
 * state, in place; step_index selects the random draws; *allsat = the step's return value. */
int odesat_stoch_step(const odesat_formula* f, uint8_t* v, uint64_t* xl, uint64_t seed, int64_t replica,
                      int64_t step_index, int* allsat);
/* stoch.rs:80 search(&CNFFormula, Option<usize>) -> Vec<bool> for R independent replicas on the device.
 *  v / xl      : host [R][N] uint8 / [R][M] uint64 initial states, or NULL for the reference's (false / 1, :84-87);
 *                written back when write_back != 0
 *  steps       : < 0 → run until some replica's step returns true (stoch.rs:101-105)
 *  The loop ends after the first chunk of `chunk` steps (<= 0 → 64) in which some replica flagged; a flagged replica
 *  stops stepping (search's `break`).  winner = lowest replica flagged at the earliest step, -1 when none;
 *  assignment[N] = its v (replica 0's when none); verified[r] = replica r's v satisfies the CNF (cnf.rs:246-264). */
int odesat_stoch_search(const odesat_formula* f, int64_t R, uint8_t* v, uint64_t* xl, uint64_t seed,
                        int64_t replica_offset, int64_t steps, int32_t chunk, int32_t write_back,
                        int64_t* solved_step, uint8_t* verified, int64_t* winner, uint8_t* assignment,
                        int64_t* steps_run);

/* ---- diagnostics (host only, no device needed) -------------------------------------------------
 * Compiles the TILE engine's clause schedule for a formula and reports it: out[0] = levels,
 * out[1] = items, out[2] = clause slots (incl. padding), out[3] = 1000 x the average number of
 * shared-memory wavefronts per quarter-warp row access (1000 = conflict-free).  perm (optional,
 * capacity >= out[2]) receives slot -> clause index (-1 = padding); items (optional, capacity >=
 * out[1]) the packed item words (slot base | count << 20 | last-of-level << 31). */
int odesat_tile_schedule_stats(int64_t varnum, int64_t n_clauses, const int64_t* clause_off,
                               const int32_t* lits, int32_t schedule, int32_t threads, int32_t depth,
                               int32_t* perm, int64_t perm_capacity, uint32_t* items,
                               int64_t items_capacity, int64_t* out);

/* ---- device-resident batch object (what the calls above are built from; used by the
 *      multi-GPU host layer and the benchmark to keep state in HBM between calls) ------------- */
int odesat_batch_create(const odesat_formula* f, int64_t R, int32_t precision, int32_t engine,
                        int32_t schedule, odesat_batch** out);
void odesat_batch_destroy(odesat_batch* b);
/* Engine actually selected (AUTO resolved), launches issued so far, bytes of HBM held. */
int odesat_batch_info(const odesat_batch* b, int32_t* engine, int64_t* kernel_launches,
                      int64_t* device_bytes);
/* main.rs:283-289 on the device: v0[r][i] = 2u-1 with u from SplitMix64(seed, replica_offset+r, i),
 * xs0 by init_short_term_memory, xl0 = 1; resets flags, step counter and dt (0.01). */
int odesat_batch_init(odesat_batch* b, uint64_t seed, int64_t replica_offset);
/* Host [R][N]/[R][M] in the batch's precision (double* or float*); resets flags/step/dt. */
int odesat_batch_upload(odesat_batch* b, const void* v, const void* xs, const void* xl);
int odesat_batch_download(odesat_batch* b, void* v, void* xs, void* xl);
/* n fixed Euler steps (system.rs:141-154) on every replica.  freeze != 0: a replica stops after
 * the update of its flagged step (simulate's `break`, system.rs:193-195).  *device_ms (may be
 * NULL) = CUDA-event time of the step loop on the library's stream. */
int odesat_batch_run_fixed(odesat_batch* b, double dt, double zeta, int64_t n, int32_t freeze,
                           float* device_ms);
/* n adaptive steps (system.rs:111-139), per-replica dt; flagged replicas stay untouched. */
int odesat_batch_run_adaptive(odesat_batch* b, double tolerance, double zeta, int64_t n,
                              float* device_ms);
/* solved_step[R] (−1 = not flagged yet); *steps_done = Euler steps issued since init/upload. */
int odesat_batch_status(odesat_batch* b, int64_t* solved_step, int64_t* steps_done);
/* min over replicas of (solved_step << 32 | replica) — the early-exit key the multi-GPU layer
 * all-reduces (MIN); INT64_MAX when no replica has flagged. */
int odesat_batch_first_solved(odesat_batch* b, int64_t* key);
/* ---- asynchronous pieces of the step loop (what odesat_simulate_batch is built from; they let a multi-process
 *      host layer put its own collective — e.g. an NCCL MIN all-reduce of the key — on the batch's stream) ----
 * The CUDA stream every kernel and copy of this batch is issued on (a cudaStream_t). */
int odesat_batch_stream(const odesat_batch* b, void** stream);
/* Enqueue n fixed Euler steps and return at once.  stop_key (may be NULL) is a DEVICE pointer to one uint64 early-exit
 * key; when the launches start and find *stop_key != INT64_MAX they do nothing (a chunk issued speculatively after the
 * chunk in which some replica, on any rank, flagged). */
int odesat_batch_run_fixed_async(odesat_batch* b, double dt, double zeta, int64_t n, int32_t freeze,
                                 const uint64_t* stop_key);
/* Enqueue the reduction of the flags: key_out[0] = min over replicas of (solved_step << 32 | replica_offset + r)
 * (INT64_MAX: none), key_out[1] = number of replicas not flagged yet.  key_out is a DEVICE pointer to two uint64. */
int odesat_batch_post_key(odesat_batch* b, int64_t replica_offset, uint64_t* key_out);
/* Wait for everything enqueued on the batch's stream. */
int odesat_batch_sync(odesat_batch* b);
/* cnf.rs:246-264 on the device: verified[r] = thresholded state of replica r satisfies the CNF. */
int odesat_batch_verify(odesat_batch* b, uint8_t* verified);
/* system.rs:238 for one replica: assignment[i] = v[i] > 0. */
int odesat_batch_assignment(odesat_batch* b, int64_t replica, uint8_t* assignment);
/* Current adaptive step size of every replica (f64 copy). */
int odesat_batch_dt(odesat_batch* b, double* dt);

#ifdef __cplusplus
}
#endif
#endif /* ODESAT_B200_H */
