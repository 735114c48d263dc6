/* sqaod_b200.h -- C ABI of libsqaod_b200.so, the B200-native back end for sqaod's CUDA solvers.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Each entry point replaces one
 * function of the reference's Python C-extension glue (sqaodc/pyglue/annealer.inc, bf_searcher.inc, formulas.inc,
 * sqaodpy/sqaod/cuda/src/cuda_device.cpp) -- the layer through which every `sqaod.cuda` call reaches
 * libsqaodc_cuda.so -- and forwards to the same C++ virtual interface (include/sqaod_b200/sqaod_api.hpp, a
 * source-compatible restatement of sqaodc/common/Solver.h and sqaodc/cuda/api.h).  INTEGRATION.md shows the binding
 * a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success; on failure a non-zero code, and sqb_last_error() returns the message the
 *     reference would have put into its std::runtime_error -> Python RuntimeError (pyglue.h:394-399).
 *   - `dtype` selects the solver precision like the trailing numpy dtype argument of the reference glue
 *     (pyglue.h:252-254): SQB_F32 or SQB_F64.  Real-valued buffers are float or double accordingly.
 *   - matrices are row-major with `stride` in elements (pyglue.h:62-69 takes it from the numpy row stride).
 *   - bit / spin buffers are signed char (numpy int8), contiguous, one row per trotter / solution.
 *   - inputs are borrowed for the duration of the call only; outputs are written into caller-allocated buffers.
 *   - handles are opaque; the reference passes the same raw pointers through Python as numpy.uint64 scalars
 *     (annealer.inc:6-10).
 * There is no CPU fallback: without a CUDA device every solver call fails with an error.
 */
#ifndef SQAOD_B200_H
#define SQAOD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SQB_F32 0
#define SQB_F64 1
#define SQB_MINIMIZE 0 /* sqaodc/common/Solver.h:9-14 "should sync with python bind" */
#define SQB_MAXIMIZE 1

typedef void *sqb_handle;

/* ---- library ---- */
int sqb_version(void);
const char *sqb_last_error(void);
/* the reference's own C symbol, probed by sqaodpy/sqaod/common/envcheck.py:77-99 (api_cuda.cpp:5-9) */
void sqaodc_cuda_version(int *version, int *cuda_version);
int sqb_device_count(int *count);

/* ---- device: sqaodpy/sqaod/cuda/src/cuda_device.cpp:5-72 (new / initialize / finalize / delete) ---- */
int sqb_device_new(sqb_handle *dev);
int sqb_device_initialize(sqb_handle dev, int devNo);
int sqb_device_finalize(sqb_handle dev);
int sqb_device_delete(sqb_handle dev);
int sqb_device_synchronize(sqb_handle dev);
/* run this device's work on a caller-owned cudaStream_t (NULL restores the device's own stream) */
int sqb_device_set_stream(sqb_handle dev, void *cuda_stream);
/* number of kernels this library launched on the device (optionally reset) */
int sqb_device_launch_count(sqb_handle dev, unsigned long long *count, int reset);
int sqb_device_num_sms(sqb_handle dev, int *num_sms);

/* ---- dense-graph annealer: pyglue/annealer.inc (line of the PyArg_ParseTuple site replaced) ---- */
int sqb_dg_annealer_new(sqb_handle *ann, int dtype);                                           /* :17  */
int sqb_dg_annealer_delete(sqb_handle ann, int dtype);                                         /* :35  */
int sqb_dg_annealer_assign_device(sqb_handle ann, sqb_handle dev, int dtype);                  /* :56  */
int sqb_dg_annealer_seed(sqb_handle ann, unsigned long long seed, int dtype);                  /* :80  */
int sqb_dg_annealer_set_qubo(sqb_handle ann, const void *W, int N, int stride, int optimize, int dtype);        /* :111 */
int sqb_dg_annealer_set_hamiltonian(sqb_handle ann, const void *h, const void *J, int N, int strideJ, double c, int dtype); /* :146 */
int sqb_dg_annealer_get_hamiltonian(sqb_handle ann, void *h, void *J, int strideJ, void *c, int dtype);          /* :472 */
int sqb_dg_annealer_get_problem_size(sqb_handle ann, int *N, int dtype);                       /* :165 */
/* one preference: `name` as in Preference.cpp:61-100; string-valued ones (algorithm) use `str`, the rest `value` */
int sqb_dg_annealer_set_preference(sqb_handle ann, const char *name, const char *str, long value, int dtype);   /* :281 */
/* "algorithm=coloring;n_trotters=32;precision=float;device=cuda" */
int sqb_dg_annealer_get_preferences(sqb_handle ann, char *buf, int buflen, int dtype);          /* :304 */
int sqb_dg_annealer_get_num_trotters(sqb_handle ann, int *m, int dtype);
int sqb_dg_annealer_get_E(sqb_handle ann, void *E, int capacity, int dtype);                    /* :334 (m values) */
int sqb_dg_annealer_get_x(sqb_handle ann, signed char *x, int dtype);                           /* :369 (m x N) */
int sqb_dg_annealer_get_q(sqb_handle ann, signed char *q, int dtype);                           /* :508 (m x N) */
int sqb_dg_annealer_set_q(sqb_handle ann, const signed char *q, int N, int dtype);              /* :397 */
int sqb_dg_annealer_set_qset(sqb_handle ann, const signed char *q, int m, int N, int dtype);    /* :436 */
int sqb_dg_annealer_randomize_spin(sqb_handle ann, int dtype);                                  /* :749 */
int sqb_dg_annealer_calculate_E(sqb_handle ann, int dtype);                                     /* :768 */
int sqb_dg_annealer_prepare(sqb_handle ann, int dtype);                                         /* :788 */
int sqb_dg_annealer_make_solution(sqb_handle ann, int dtype);                                   /* :807 */
int sqb_dg_annealer_get_system_E(sqb_handle ann, double G, double beta, double *E, int dtype);  /* :838 */
int sqb_dg_annealer_anneal_one_step(sqb_handle ann, double G, double beta, int dtype);          /* :866 */
/* extras (no reference counterpart): accepted flips and inter-CTA flag waits since prepare(); spins without
 * building solution lists (device int8 matrix <-> host m x N buffer) */
int sqb_dg_annealer_get_stats(sqb_handle ann, unsigned long long *accepted, unsigned long long *waits, int dtype);
int sqb_dg_annealer_get_spins(sqb_handle ann, signed char *q, int dtype);
/* SM cycles dot warp 0 / the chain warp spent working inside the look-ahead windows, summed over CTAs (profiling aid:
 * whichever is larger paces the sweep) */
int sqb_dg_annealer_get_barrier_cycles(sqb_handle ann, unsigned long long *dot, unsigned long long *chain, int dtype);
/* raw sweep counters summed over CTAs since prepare(): [0] accepted flips, [1] flag polls, [2] dot-warp-0 busy cycles,
 * [3] chain-warp busy cycles, [5] helper-warp busy cycles (snapshots, conflict masks), [6] prep-warp busy cycles (Philox
 * tables), [4] / [7] chain-warp cycles waiting for dot products / for neighbour data */
int sqb_dg_annealer_get_counters(sqb_handle ann, unsigned long long *out8, int dtype);
/* field mode with carried fields (field_refresh > 1): H[y][j] = h[j] + 2 sum_i J[j][i] q[y][i] as the sweep left it, rows of ldH
 * elements; *valid = 0 when the solver carries no fields (classic mode, per-step recomputation, spins just written) */
int sqb_dg_annealer_get_fields(sqb_handle ann, void *H, int ldH, int *valid, int dtype);
/* the eight counters above plus the field-mode chain profile of chain warp 0, summed over CTAs: [8] cycles waiting for
 * cross-term gathers, [9] cycles idle (blocked on another warp / CTA), [10] cycles in the per-window barrier of the chain warps,
 * [11] evaluation passes, [12] / [13] gather waits forced by an uncertain attempt / by a second commit, [14] passes that
 * ended on a blocked attempt */
int sqb_dg_annealer_get_profile(sqb_handle ann, unsigned long long *out16, int dtype);
/* field mode, last launch, chain warp 0 of every CTA (16 words per CTA; words 8-11: the neighbour warp): trotters, cycles waiting for fields / for neighbour CTAs,
 * cycles of chain work, cycles of the whole sweep loop, end time (globaltimer ns), accepted flips of that warp, evaluation passes */
int sqb_dg_annealer_get_cta_profile(sqb_handle ann, unsigned long long *out, int max_ctas, int *n, int dtype);
/* how annealOneStep obtains h_x + sum_j J_xj q_j (no reference counterpart; both modes run the same Markov chain):
 *   0 "classic": one J row streamed per attempt (N*sizeof(real) bytes of traffic per attempt);
 *   1 "field":   the local fields of every trotter are computed once per step (J.q spin GEMM, tensor cores for fp32), kept in
 *                shared memory and updated with one J row per ACCEPTED flip;
 *  -1 automatic (default): field mode whenever the field rows fit in shared memory.
 * field_refresh > 0: number of steps between two recomputations of J.q (0: automatic).  Takes effect at the next prepare(). */
int sqb_dg_annealer_set_sweep_mode(sqb_handle ann, int mode, int field_refresh, int dtype);
int sqb_dg_annealer_get_sweep_mode(sqb_handle ann, int *mode, int dtype); /* the mode prepare() chose: 0 or 1 */

/* replica batch (no reference counterpart): R independent replicas of the problem, replica r seeded
 * seed + r, annealed side by side in one launch.  Spin and energy buffers then hold R x n_trotters rows, replica major. */
/* a batch of n_problems DIFFERENT QUBOs of the same size (W: n_problems x N x N, row stride ldW elements) annealed side by side
 * in one cooperative launch per step; problem r uses seed + r; spins / energies are returned as n_problems*m rows (SURVEY 8f-2:
 * the caller loop of sqaodpy/sqaod/common/common.py:144-160 run for many problems at once) */
/* extras for benchmarks and multi-GPU tests: a synthetic symmetric W ~ U(-0.5, 0.5) generated on the device from
 * Philox(seed, min(i,j), max(i,j)) (optionally on the reference tests' 2^-14 grid), so that large problems (N = 32768: 4 GiB)
 * need no host generation or upload; get_qubo_random returns the same matrix to the host (for oracle checks) */
int sqb_dg_annealer_set_qubo_random(sqb_handle ann, int N, unsigned long long seed, int quantize, int optimize, int dtype);
int sqb_dg_annealer_get_qubo_random(sqb_handle ann, void *W, int N, int ldW, unsigned long long seed, int quantize, int dtype);
int sqb_dg_annealer_set_qubo_batch(sqb_handle ann, const void *W, int n_problems, int N, int ldW, int optimize, int dtype);
int sqb_dg_annealer_set_num_replicas(sqb_handle ann, int n_replicas, int dtype);
int sqb_dg_annealer_get_num_replicas(sqb_handle ann, int *n_replicas, int dtype);

/* ring sharding over several GPUs (no reference counterpart): the solver anneals trotters
 * [rank*m/world, (rank+1)*m/world) of ONE ring of m trotters; J and h are replicated.  Call order: set_qubo,
 * ring_configure, prepare, ring_export -> exchange the 64-byte handles -> ring_attach(left, right), set/randomize spins,
 * ring_push_halos, then anneal_one_step (which ends with a halo push over NVLink). */
int sqb_dg_annealer_ring_configure(sqb_handle ann, int rank, int world, int m_global, int dtype);
int sqb_dg_annealer_ring_export(sqb_handle ann, unsigned char *handle64, int dtype);
int sqb_dg_annealer_ring_attach(sqb_handle ann, const unsigned char *left_handle64, const unsigned char *right_handle64, int dtype);
int sqb_dg_annealer_ring_push_halos(sqb_handle ann, int dtype);

/* ---- bipartite-graph annealer: pyglue/annealer.inc ---- */
int sqb_bg_annealer_new(sqb_handle *ann, int dtype);
int sqb_bg_annealer_delete(sqb_handle ann, int dtype);
int sqb_bg_annealer_assign_device(sqb_handle ann, sqb_handle dev, int dtype);
int sqb_bg_annealer_seed(sqb_handle ann, unsigned long long seed, int dtype);
int sqb_bg_annealer_set_qubo(sqb_handle ann, const void *b0, const void *b1, const void *W, int N0, int N1, int stride,
                             int optimize, int dtype);                                          /* :202 */
int sqb_bg_annealer_set_hamiltonian(sqb_handle ann, const void *h0, const void *h1, const void *J, int N0, int N1,
                                    int strideJ, double c, int dtype);                          /* :240 */
int sqb_bg_annealer_get_hamiltonian(sqb_handle ann, void *h0, void *h1, void *J, int strideJ, void *c, int dtype); /* :727 */
int sqb_bg_annealer_get_problem_size(sqb_handle ann, int *N0, int *N1, int dtype);              /* :259 */
int sqb_bg_annealer_set_preference(sqb_handle ann, const char *name, const char *str, long value, int dtype);
int sqb_bg_annealer_get_preferences(sqb_handle ann, char *buf, int buflen, int dtype);
int sqb_bg_annealer_get_num_trotters(sqb_handle ann, int *m, int dtype);
int sqb_bg_annealer_get_E(sqb_handle ann, void *E, int capacity, int dtype);
int sqb_bg_annealer_get_x(sqb_handle ann, signed char *x0, signed char *x1, int dtype);         /* :553 (m x N0, m x N1) */
int sqb_bg_annealer_get_q(sqb_handle ann, signed char *q0, signed char *q1, int dtype);         /* :679 */
int sqb_bg_annealer_set_q(sqb_handle ann, const signed char *q0, const signed char *q1, int N0, int N1, int dtype); /* :586 */
int sqb_bg_annealer_set_qset(sqb_handle ann, const signed char *q0, const signed char *q1, int m, int N0, int N1, int dtype); /* :636 */
int sqb_bg_annealer_randomize_spin(sqb_handle ann, int dtype);
int sqb_bg_annealer_calculate_E(sqb_handle ann, int dtype);
int sqb_bg_annealer_prepare(sqb_handle ann, int dtype);
int sqb_bg_annealer_make_solution(sqb_handle ann, int dtype);
int sqb_bg_annealer_get_system_E(sqb_handle ann, double G, double beta, double *E, int dtype);
int sqb_bg_annealer_anneal_one_step(sqb_handle ann, double G, double beta, int dtype);

/* ---- dense-graph brute-force searcher: pyglue/bf_searcher.inc ---- */
int sqb_dg_bf_searcher_new(sqb_handle *s, int dtype);                                           /* :16  */
int sqb_dg_bf_searcher_delete(sqb_handle s, int dtype);                                         /* :34  */
int sqb_dg_bf_searcher_assign_device(sqb_handle s, sqb_handle dev, int dtype);                  /* :55  */
int sqb_dg_bf_searcher_set_qubo(sqb_handle s, const void *W, int N, int stride, int optimize, int dtype); /* :89 */
int sqb_dg_bf_searcher_get_problem_size(sqb_handle s, int *N, int dtype);                       /* :116 */
int sqb_dg_bf_searcher_set_preference(sqb_handle s, const char *name, const char *str, long value, int dtype); /* :198 */
int sqb_dg_bf_searcher_get_preferences(sqb_handle s, char *buf, int buflen, int dtype);         /* :221 */
int sqb_dg_bf_searcher_prepare(sqb_handle s, int dtype);                                        /* :351 */
int sqb_dg_bf_searcher_calculate_E(sqb_handle s, int dtype);                                    /* :370 */
int sqb_dg_bf_searcher_make_solution(sqb_handle s, int dtype);                                  /* :389 */
int sqb_dg_bf_searcher_search_range(sqb_handle s, int *done, unsigned long long *cur_x, int dtype); /* :410-424 */
int sqb_dg_bf_searcher_search(sqb_handle s, int dtype);                                         /* :457 */
int sqb_dg_bf_searcher_get_num_solutions(sqb_handle s, int *n, int dtype);
int sqb_dg_bf_searcher_get_x(sqb_handle s, signed char *x, int capacity, int dtype);            /* :262 (n x N) */
int sqb_dg_bf_searcher_get_E(sqb_handle s, void *E, int capacity, int dtype);                   /* :334 (n values) */
/* extras for sharded search: restrict the searcher to x in [begin, end); local minimum so far; drop
 * the solution list when another shard found a lower minimum; packed solutions for gathering */
int sqb_dg_bf_searcher_set_range(sqb_handle s, unsigned long long x_begin, unsigned long long x_end, int dtype);
int sqb_dg_bf_searcher_get_Emin(sqb_handle s, double *Emin, int dtype);
int sqb_dg_bf_searcher_get_packed_x(sqb_handle s, unsigned long long *x, int capacity, int *n, int dtype);
int sqb_dg_bf_searcher_set_packed_solutions(sqb_handle s, double Emin, const unsigned long long *x, int n, int dtype);

/* ---- bipartite-graph brute-force searcher: pyglue/bf_searcher.inc ---- */
int sqb_bg_bf_searcher_new(sqb_handle *s, int dtype);
int sqb_bg_bf_searcher_delete(sqb_handle s, int dtype);
int sqb_bg_bf_searcher_assign_device(sqb_handle s, sqb_handle dev, int dtype);
int sqb_bg_bf_searcher_set_qubo(sqb_handle s, const void *b0, const void *b1, const void *W, int N0, int N1, int stride,
                                int optimize, int dtype);                                       /* :151 */
int sqb_bg_bf_searcher_get_problem_size(sqb_handle s, int *N0, int *N1, int dtype);             /* :178 */
int sqb_bg_bf_searcher_set_preference(sqb_handle s, const char *name, const char *str, long value, int dtype);
int sqb_bg_bf_searcher_get_preferences(sqb_handle s, char *buf, int buflen, int dtype);
int sqb_bg_bf_searcher_prepare(sqb_handle s, int dtype);
int sqb_bg_bf_searcher_calculate_E(sqb_handle s, int dtype);
int sqb_bg_bf_searcher_make_solution(sqb_handle s, int dtype);
int sqb_bg_bf_searcher_search_range(sqb_handle s, int *done, unsigned long long *cur_x0, unsigned long long *cur_x1, int dtype); /* :434-448 */
int sqb_bg_bf_searcher_search(sqb_handle s, int dtype);
int sqb_bg_bf_searcher_get_num_solutions(sqb_handle s, int *n, int dtype);
int sqb_bg_bf_searcher_get_x(sqb_handle s, signed char *x0, signed char *x1, int capacity, int dtype); /* :306 */
int sqb_bg_bf_searcher_get_E(sqb_handle s, void *E, int capacity, int dtype);

/* ---- formulas: pyglue/formulas.inc (outputs first, as in the reference glue) ---- */
int sqb_dg_formulas_new(sqb_handle *f, int dtype);
int sqb_dg_formulas_delete(sqb_handle f, int dtype);                                            /* :41  */
int sqb_dg_formulas_assign_device(sqb_handle f, sqb_handle dev, int dtype);                     /* :61  */
/* E[nBatch] = x_b^T W x_b ; nBatch == 1 is dense_graph_calculate_E (:94), otherwise batch_calculate_E (:127) */
int sqb_dg_formulas_calculate_E(sqb_handle f, void *E, const void *W, int N, int strideW, const signed char *x, int nBatch, int dtype);
int sqb_dg_formulas_calculate_hamiltonian(sqb_handle f, void *h, void *J, int strideJ, void *c, const void *W, int N, int strideW, int dtype); /* :163 */
int sqb_dg_formulas_calculate_E_from_spin(sqb_handle f, void *E, const void *h, const void *J, int N, int strideJ, double c,
                                          const signed char *q, int nBatch, int dtype);         /* :207, :246 */
int sqb_bg_formulas_new(sqb_handle *f, int dtype);
int sqb_bg_formulas_delete(sqb_handle f, int dtype);                                            /* :294 */
int sqb_bg_formulas_assign_device(sqb_handle f, sqb_handle dev, int dtype);                     /* :315 */
int sqb_bg_formulas_calculate_E(sqb_handle f, void *E, const void *b0, const void *b1, const void *W, int N0, int N1, int strideW,
                                const signed char *x0, const signed char *x1, int nBatch, int dtype); /* :350, :387 */
int sqb_bg_formulas_calculate_E_2d(sqb_handle f, void *E, const void *b0, const void *b1, const void *W, int N0, int N1, int strideW,
                                   const signed char *x0, int n0, const signed char *x1, int n1, int dtype); /* :426 */
int sqb_bg_formulas_calculate_hamiltonian(sqb_handle f, void *h0, void *h1, void *J, int strideJ, void *c, const void *b0,
                                          const void *b1, const void *W, int N0, int N1, int strideW, int dtype); /* :468 */
int sqb_bg_formulas_calculate_E_from_spin(sqb_handle f, void *E, const void *h0, const void *h1, const void *J, int N0, int N1,
                                          int strideJ, double c, const signed char *q0, const signed char *q1, int nBatch, int dtype); /* :516, :562 */

#ifdef __cplusplus
}
#endif
#endif /* SQAOD_B200_H */
