/*
 * cfem.h -- C ABI of one compiled filter-error model (B200 / sm_100a).
 *
 * Every model shared library produced by colloc_fem_code_b200.codegen exports
 * exactly these entry points (extern "C", plain pointers and sizes, no C++ or
 * torch types).  One library = one model structure (class composition + dims);
 * one handle = one problem (or one batch of same-shaped problems) bound to one
 * GPU.  All functions return 0 on success or a negative CFEM_E* code; the
 * message of the last failure is available from cfem_last_error().  Handles
 * are not re-entrant; distinct handles may be driven from distinct threads.
 *
 * What each entry point replaces in the reference stack
 * (/root/reference; the glue package ceacoest.optim is third party and absent,
 * so the citations are the reference's call sites of that glue):
 *
 *   cfem_create              fem.py:11-57      Problem(model, y, u): layout of the
 *                                              decision / constraint vectors and the
 *                                              per-sample data y, u (copied to HBM once)
 *   cfem_set_dvec            fem.py:59-62      Problem.variables(dvec): binds a new
 *                                              decision vector ("new_x" of IPOPT)
 *   cfem_eval_f              symfem.py:61-65   objective  sum_k loglikelihood
 *   cfem_eval_grad_f         adfem.py:99-120   dense objective gradient, parameter
 *                                              entries summed over the samples
 *   cfem_eval_g              symfem.py:50-59   dynamics defects + innovations
 *                                              (+ parameter-only constraints)
 *   cfem_eval_jac_values     adfem.py:88-96    COO values of the constraint Jacobian
 *   cfem_eval_hess_values    adfem.py:65-79    COO values of the Lagrangian Hessian
 *                                              (lower triangle), obj_factor and
 *                                              multipliers as in IPOPT's eval_h
 *   cfem_layout / cfem_model_json              the sparsity bookkeeping the glue
 *                                              derives from fem.py:36-57 (offsets of
 *                                              every variable, constraint and COO block)
 *
 * The five cfem_eval_* calls map 1:1 onto IPOPT's eval_f / eval_grad_f / eval_g /
 * eval_jac_g(values) / eval_h(values) callbacks (attas_sp_ml.py:153-159 reaches
 * them through problem.ipopt()).  Index (structure) arrays are produced on the
 * host from cfem_layout; see INTEGRATION.md.
 *
 * Array conventions: every array is FP64, C order.  The decision vector, the
 * constraint / multiplier vector and the COO value arrays use the problem's own
 * (IPOPT-facing) order described in DESIGN.md; with batch > 1 they are
 * [batch][...] with dense per-problem stride (ndec, ncons, nnz_jac, nnz_hess).
 */
#ifndef CFEM_H
#define CFEM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFEM_ABI_VERSION 1

/* status codes */
#define CFEM_OK            0
#define CFEM_EINVAL      (-1)   /* bad argument                      */
#define CFEM_ECUDA       (-2)   /* CUDA runtime / launch failure     */
#define CFEM_ENOMEM      (-3)   /* host or device allocation failure */
#define CFEM_ESTATE      (-4)   /* call sequence error (e.g. no dvec) */

/* evaluation selector bits for cfem_eval() / cfem_fetch() */
#define CFEM_F     1u    /* objective value                          */
#define CFEM_GRAD  2u    /* dense objective gradient                 */
#define CFEM_G     4u    /* constraint values                        */
#define CFEM_JAC   8u    /* constraint Jacobian COO values           */
#define CFEM_HESS 16u    /* Lagrangian Hessian COO values            */
#define CFEM_ALL  31u
/* input selectors of cfem_upload_pieces() */
#define CFEM_X      32u  /* decision vector                          */
#define CFEM_LAMBDA 64u  /* constraint multipliers                   */

typedef struct cfem_problem cfem_problem;   /* opaque */

/* ---- model introspection (no GPU needed) ---------------------------------- */
int          cfem_abi_version(void);
/* JSON description of the compiled structure: variables, data arrays, scalars,
 * functions, Jacobian / Hessian blocks, tile size, instantiated kernels. */
const char*  cfem_model_json(void);
/* Counts for n_samples samples (halo as in cfem_create); any pointer may be NULL. */
int          cfem_model_sizes(int64_t n_samples, int32_t halo, int64_t* ndec,
                              int64_t* ncons, int64_t* nnz_jac,
                              int64_t* nnz_hess);

/* ---- life cycle ------------------------------------------------------------ */
/*
 * n_samples  number of samples N owned by this handle (per problem)
 * batch      number of independent same-shaped problems evaluated together
 * halo       0: stand-alone problem (functions with N-1 rows have N-1 rows);
 *            1: left/inner time shard -- per-sample variables that are read one
 *               sample ahead carry one extra (halo) row and the N-1 row
 *               functions have N rows
 * data       n_data host pointers, one per per-sample data array of the model
 *            (order of "data" in cfem_model_json), each [batch][rows][core]
 * scalars    n_scalars scalar model inputs (order of "scalars" in the JSON)
 * device     CUDA device ordinal
 */
int          cfem_create(cfem_problem** out, int64_t n_samples, int32_t batch,
                         int32_t halo, const double* const* data,
                         int32_t n_data, const double* scalars,
                         int32_t n_scalars, int32_t device);
void         cfem_destroy(cfem_problem* p);
/* p may be NULL: returns the message of the last failed cfem_create. */
const char*  cfem_last_error(const cfem_problem* p);
/* Use an existing cudaStream_t (e.g. torch's current stream); NULL = own stream. */
int          cfem_set_stream(cfem_problem* p, void* cuda_stream);
int          cfem_sizes(const cfem_problem* p, int64_t* ndec, int64_t* ncons,
                        int64_t* nnz_jac, int64_t* nnz_hess);
/* Offsets (in doubles) of every variable in dvec, every constraint function in
 * g / lambda, every Jacobian block and every Hessian block in the value arrays,
 * plus the row count of every function; array lengths as listed in the JSON. */
int          cfem_layout(const cfem_problem* p, int64_t* var_offset,
                         int64_t* cons_offset, int64_t* jac_offset,
                         int64_t* hess_offset, int64_t* fun_rows);

/* ---- inputs ---------------------------------------------------------------- */
int          cfem_set_dvec(cfem_problem* p, const double* dvec_host);
int          cfem_set_dvec_device(cfem_problem* p, const double* dvec_dev);
int          cfem_set_multipliers(cfem_problem* p, double obj_factor,
                                  const double* lambda_host);
int          cfem_set_multipliers_device(cfem_problem* p, double obj_factor,
                                         const double* lambda_dev);

/* ---- evaluation ------------------------------------------------------------ */
/* Launch the kernels that produce the selected outputs into the handle's device
 * buffers (asynchronous on the handle's stream). */
int          cfem_eval(cfem_problem* p, uint32_t what);
/* Copy ONE result (a single CFEM_* bit) to host memory and synchronise. */
int          cfem_fetch(cfem_problem* p, uint32_t which, double* host_out);
/* Same copy without the synchronisation (host_out should be pinned memory);
 * complete after cfem_synchronize(). */
int          cfem_fetch_async(cfem_problem* p, uint32_t which, double* host_out);
/* A whole callback set with ONE copy per direction.  Inputs [dvec | lambda]
 * and results [f | grad | g | jac | hess] each live in one device allocation;
 * cfem_io_layout gives the segment offsets (in doubles, 256-byte aligned) and
 * the total sizes of host blocks that mirror them.  cfem_set_inputs copies
 * the CFEM_X and/or CFEM_LAMBDA segment(s) of such a host block (and sets
 * obj_factor with CFEM_LAMBDA); cfem_fetch_results_async copies the range
 * from the first to the last requested result (a mask of CFEM_F..CFEM_HESS,
 * all evaluated) into the host block.  Asynchronous on the handle's stream;
 * the blocks should be page-locked (cfem_host_alloc). */
int          cfem_io_layout(const cfem_problem* p, int64_t* in_off /*[2]*/,
                            int64_t* in_total, int64_t* res_off /*[5]*/,
                            int64_t* res_total);
int          cfem_set_inputs(cfem_problem* p, uint32_t which, double obj_factor,
                             const double* host_inputs);
int          cfem_fetch_results_async(cfem_problem* p, uint32_t which,
                                      double* host_results);
/* One whole callback set (eval_f .. eval_h of IpStdCInterface.h at a new x and
 * new multipliers) between page-locked host blocks laid out as above, with
 * the two directions of the bus overlapped: x goes up first and the kernels
 * of f | grad | g | Jacobian start; lambda goes up on a second stream WHILE the
 * first results come down, then the Hessian kernels and their copy follow on
 * that stream.  Returns when everything is in host_results. */
int          cfem_eval_callback_set(cfem_problem* p, double obj_factor,
                                    const double* host_inputs,
                                    double* host_results);
/* Time-sharded problems behind ONE solver process: the decision vector, the
 * multipliers and the results live in shared page-locked host vectors in the
 * GLOBAL order; every rank moves only its own pieces, straight between that
 * memory and its device arrays over its own PCIe link (asynchronous on the
 * handle's stream; offsets and lengths in doubles).  cfem_upload_pieces marks
 * the input as set (new x / new lambda); cfem_set_obj_factor sets sigma. */
int          cfem_upload_pieces(cfem_problem* p, uint32_t which,
                                const double* host_base, int32_t n,
                                const int64_t* dev_off, const int64_t* host_off,
                                const int64_t* len);
int          cfem_fetch_pieces(cfem_problem* p, uint32_t which, double* host_base,
                               int32_t n, const int64_t* dev_off,
                               const int64_t* host_off, const int64_t* len);
int          cfem_set_obj_factor(cfem_problem* p, double obj_factor);
/* IPOPT-shaped conveniences: evaluate if stale, copy to host, synchronise. */
int          cfem_eval_f(cfem_problem* p, double* f);
int          cfem_eval_grad_f(cfem_problem* p, double* grad);
int          cfem_eval_g(cfem_problem* p, double* g);
int          cfem_eval_jac_values(cfem_problem* p, double* values);
int          cfem_eval_hess_values(cfem_problem* p, double obj_factor,
                                   const double* lambda, double* values);
/* Device-resident results (valid until cfem_destroy); any pointer may be NULL.
 * f is [batch]; reduce is the [batch][n_reduce] vector (objective first, then
 * the parameter-gradient entries listed under "reduce" in the JSON) that a
 * time-sharded run all-reduces across ranks. */
int          cfem_device_ptrs(cfem_problem* p, double** dvec, double** lambda,
                              double** f, double** grad, double** g,
                              double** jac, double** hess, double** reduce);
/* Time-sharded runs: add the all-reduced [batch][n_reduce] vector back into f
 * and the gradient (device pointer, same stream). */
int          cfem_apply_reduced(cfem_problem* p, const double* reduce_dev);
/* Fused alternative to all-reduce + cfem_apply_reduced: every rank's kernel
 * exchanges its [n_reduce] partial sums with the peers through NVLink-mapped
 * memory and finishes with the global objective / parameter gradient itself.
 * inbox_ptrs[r] / flag_ptrs[r]: rank r's inbox (cfem_peer_layout doubles) and
 * flag words (uint64), zero-initialised, mapped into THIS process (CUDA IPC /
 * symmetric memory); entry `rank` is the local one.  All ranks must then issue
 * the same sequence of cfem_eval calls.  world <= 1 switches it off. */
int          cfem_set_peers(cfem_problem* p, int32_t rank, int32_t world,
                            void* const* inbox_ptrs, void* const* flag_ptrs);
int          cfem_peer_layout(const cfem_problem* p, int32_t world,
                              int64_t* inbox_doubles, int64_t* flag_words);
/* pipelined != 0: the per-sample kernel POSTS its partial sums to the peers
 * and, instead of waiting for theirs, finishes the sums of the PREVIOUS launch
 * (posted a whole kernel duration ago); ranks no longer rendezvous at every
 * step and a rank can run one evaluation ahead of the slowest one.  The sums
 * of the latest launch are finished on demand by a small collect kernel:
 * cfem_fetch*, the cfem_eval_* conveniences and cfem_synchronize launch it;
 * readers of the raw device pointers must call cfem_synchronize first.
 * Default 0: the exchange completes inside the per-sample kernel. */
int          cfem_set_peer_mode(cfem_problem* p, int32_t pipelined);
int          cfem_synchronize(cfem_problem* p);
/* enabled != 0: cfem_eval launches ONE CUDA graph per evaluation (the
 * parameter-only kernel and the per-sample kernel as parallel nodes, captured
 * once per kernel variant, kernel arguments updated in place) instead of two
 * kernel launches with fork/join events -- for the launch-bound native
 * trajectory lengths.  Ignored while cfem_set_kernel_timing is on (the timing
 * events live between the launches).  Default: environment CFEM_GRAPH, else 0. */
int          cfem_set_graph_mode(cfem_problem* p, int32_t enabled);

/* ---- measurement helpers --------------------------------------------------- */
/* CUDA events on the handle's stream (slot 0..15). */
int          cfem_event_record(cfem_problem* p, int32_t slot);
int          cfem_event_elapsed_ms(cfem_problem* p, int32_t start_slot,
                                   int32_t stop_slot, float* ms);
/* When enabled, every cfem_eval brackets its per-sample kernel (the dominant,
 * HBM-bound launch) with CUDA events on the handle's stream;
 * cfem_last_sample_kernel_ms returns the duration of the latest such launch. */
int          cfem_set_kernel_timing(cfem_problem* p, int32_t enabled);
int          cfem_last_sample_kernel_ms(cfem_problem* p, float* ms);
/* Durations of the last n (<= 64) timed launches, oldest first: lets a timed
 * loop run without any host synchronisation inside it. */
int          cfem_sample_kernel_ms_history(cfem_problem* p, float* ms, int32_t n);
/* Number of kernels this handle has launched so far. */
int64_t      cfem_launch_count(const cfem_problem* p);
/* Write `bytes` of zeros to a scratch buffer (L2 flush between timed iterations). */
int          cfem_flush_l2(cfem_problem* p, size_t bytes);

/* ---- pinned host memory for the IPOPT-facing buffers ----------------------- */
/* Host arrays passed to cfem_set_dvec / cfem_set_multipliers / cfem_fetch* may
 * be ordinary pageable memory (the arrays IPOPT owns, IpStdCInterface.h:
 * Eval_F_CB .. Eval_H_CB): arrays above 1 MiB are then pipelined through
 * page-locked bounce buffers with a multi-threaded host copy
 * (CFEM_COPY_THREADS, CFEM_COPY_CHUNK_MB) and the call returns when the data
 * has arrived.  Page-locked arrays (cfem_host_alloc, or cfem_host_register of
 * an allocation the caller keeps in place) are DMA sources / targets
 * themselves; such copies are asynchronous on the handle's stream. */
void*        cfem_host_alloc(size_t bytes);
void         cfem_host_free(void* ptr);
int          cfem_host_register(void* ptr, size_t bytes);
int          cfem_host_unregister(void* ptr);
/* Release store / acquire load of a control word in host memory shared by the
 * solver process and the ranks of a time-sharded problem (no CUDA involved). */
void         cfem_store_release_i64(int64_t* ptr, int64_t value);
int64_t      cfem_load_acquire_i64(const int64_t* ptr);

#ifdef __cplusplus
}
#endif
#endif /* CFEM_H */
