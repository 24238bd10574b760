/* rmt_b200.h — C ABI of librmtb200.so, the B200-native engine behind PyREMOT's
 * `rmtExe` for the pseudo-homogeneous packed-bed models N1 and N2.
 *
 * The reference is pure Python and has no FFI; the boundary it exposes for this
 * path is (file:line relative to /root/reference/PyREMOT/)
 *
 *   rmtExe(modelInput)                                       rmt.py:21-80
 *     -> rmtCoreClass.modExe -> N1Init / N2Init              docs/rmtCore.py:63-127, :393-413
 *     -> PackedBedHomoReactorClass.runN1 / runN2             docs/pbHomoReactor.py:2694, :3319
 *          setup of per-solve constants                      :2744-2852 / :3370-3507
 *          solve_ivp(fun=modelEquationN1|N2, ...)            :2931 / :3609   <- hot loop
 *          un-scaling sortResult4/5 + mole fractions         solvers/solResultAnalysis.py:191-301
 *   reactionRateExe(loopVars, VARS, RATES)                   docs/rmtReaction.py:11-61
 *
 * A maintainer binds these entry points with `ctypes`;
 * rmt_app_b200/capi.py is that binding.  Conventions:
 *   - every function returns 0 on success, non-zero on error; the message is
 *     available (per thread) from rmt_last_error();
 *   - `d_*` arguments are DEVICE pointers owned by the caller (e.g. the
 *     data_ptr() of a torch CUDA tensor), `h_*` / unprefixed arrays are host
 *     memory; the library allocates only internal scratch, released by
 *     rmt_module_free / rmt_shutdown;
 *   - `stream` is a CUstream / cudaStream_t handle (NULL = default stream);
 *     device-pointer entry points are asynchronous on it;
 *   - a per-instance solver failure is reported in status[], never as a call
 *     failure:  0 ok, 1 max_steps reached, 2 step size underflow, 3 non-finite
 *     state (the Python reference raises on those);
 *   - all arrays are FP64, structure-of-arrays, instance index fastest.
 *
 * There is no CPU fallback: without a CUDA driver rmt_init fails.
 */
#ifndef RMT_B200_H
#define RMT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t rmt_blob_t;     /* host-side compiled module image (cubin)   */
typedef uint64_t rmt_module_t;   /* module loaded on the current device       */
typedef uint64_t rmt_comm_t;     /* communicator over the GPUs of one job     */
#define RMT_COMM_ID_BYTES 128    /* size of a communicator id (ncclUniqueId)  */

/* model description read back from a loaded module */
typedef struct rmt_module_info {
    int32_t model;        /* 1 = N1, 2 = N2, 7 = M7 (dimensional twin of N1;    */
                          /* served by the rmt_n1_* entry points), 9 = M9       */
                          /* (dimensional twin of N2; rmt_n2_* entry points)    */
    int32_t n;            /* unknowns per axial point                          */
    int32_t nc;           /* species                                           */
    int32_t nr;           /* reactions                                         */
    int32_t nin;          /* primary input rows (rmt_setup)                    */
    int32_t nconst;       /* derived constant rows                             */
    int32_t nkp;          /* kinetic parameter slots                           */
    int32_t stages;       /* Rosenbrock stages                                 */
    int32_t block;        /* threads per block of the integrator               */
    int32_t iso;          /* 1 = iso-thermal                                   */
    int32_t flops_rhs_alg, flops_rhs_wt;     /* per RHS evaluation (per node)  */
    int32_t flops_jac_alg, flops_jac_wt;     /* per RHS+Jacobian evaluation    */
    int32_t m;            /* unknowns of the integrator's linear systems: n, or */
                          /* nr + (n - nc) when it works in reaction extents    */
    int32_t lanes;        /* N2/M9: threads per reactor (nodes handled in        */
                          /* parallel); 0 = stage-pipelined mapping (one thread */
                          /* per reactor and pair of Rosenbrock stages)         */
} rmt_module_info;

const char* rmt_last_error(void);
const char* rmt_version(void);

/* Bind the calling thread to CUDA device `device` (primary context, shared with
 * the CUDA runtime / PyTorch).  Fails when no driver or device is present. */
int rmt_init(int device);
int rmt_device_count(int* count);
int rmt_shutdown(void);

/* ---- code generation back end: NVRTC, sm_100a --------------------------------
 * Replaces the reference's per-call `eval()` of Cp strings (docs/rmtThermo.py:37)
 * and Python-lambda kinetics (docs/rmtReaction.py:44-58).  `model_src` is the
 * generated "rmt_model.cuh", `kernels_src` the hand-written rmt_kernels.cu.
 * Works without a GPU (used by the build check). */
int rmt_nvrtc_compile(const char* model_src, const char* kernels_src, const char* arch,
                      int block, const char* const* extra_opts, int n_extra_opts, rmt_blob_t* blob_out);
int rmt_blob_data(rmt_blob_t blob, const void** data, size_t* size);
const char* rmt_blob_log(rmt_blob_t blob);
int rmt_blob_ptx(rmt_blob_t blob, const char** ptx, size_t* size);
int rmt_blob_free(rmt_blob_t blob);

int rmt_module_load(const void* cubin, size_t size, rmt_module_t* module_out);
int rmt_module_get_info(rmt_module_t m, rmt_module_info* info);
int rmt_module_free(rmt_module_t m);

/* ---- per-solve constants: runN1 :2744-2852 / runN2 :3370-3507 ------------------
 * d_rows [n_rows][B] holds the inputs that vary per instance; row_map[q] (q <
 * nin) is the row of input q in d_rows or -1, in which case uniform[q] is used
 * for every instance.  Input order: temperature, pressure, concentration[nc],
 * volumetric-flowrate, ReInDi, ReLe, PaDi, BeVoFr, OvHeTrCo, MeTe,
 * mixture-viscosity, EfHeTrAr (read by the dimensional models M7 / M9 only), CaDe,
 * CaSpHeCa (M9 only), then the scalar VARS entries in VARS order.
 * d_consts [nconst][B]. */
int rmt_setup(rmt_module_t m, int64_t B, const double* d_rows, int32_t n_rows, const int32_t* row_map,
              const double* uniform, double* d_consts, void* stream);

/* ---- N1: modelEquationN1 (:3017-3314) ------------------------------------------
 * d_y, d_f [n][B];  d_J [n*n][B] with d f_r / d y_c at row r*n + c. */
int rmt_n1_rhs(rmt_module_t m, int64_t B, const double* d_consts, const double* d_y, double* d_f, void* stream);
int rmt_n1_jac(rmt_module_t m, int64_t B, const double* d_consts, const double* d_y, double* d_f, double* d_J,
               void* stream);

/* The integrator's own form of the same equations (no reference counterpart; parity
 * hook for the reduced Jacobian).  When nr < nc the species balances f_C = c(y) nu^T R
 * are integrated in reaction extents: unknowns (xi_1..xi_nr, P, T), y_C = y_C(0) +
 * nu^T xi.  d_g [m][B] = (c R_j, f_P, f_T), d_A [m*m][B] = dg/d(unknowns), so that
 * E d_A == d_J E and E d_g == d_f with E = [[nu^T, 0], [0, I]].  m == n: d_g == d_f,
 * d_A == d_J. */
int rmt_n1_sys(rmt_module_t m, int64_t B, const double* d_consts, const double* d_y, double* d_g, double* d_A,
               void* stream);

/* ---- N1: solve_ivp call of runN1 (:2931) + post-processing (:2949-2983) ---------
 * z_eval[n_eval] (host): increasing output positions in [0, 1], the last one is
 * the end of the integration (the reference uses linspace(0, 1, 101)).
 * out_mode 0: raw scaled states (sol.y), rows = n;  1: dataYs rows (y_i, P [Pa],
 * T [K]), rows = n;  2: everything runN1 packs, rows = 2n + nc: raw | C_i
 * [mol/m^3] (dataYCons2) | dataYs rows.
 * dense 1: Rosenbrock dense output at z_eval;  0: steps land on every z_eval.
 * d_out [n_eval][rows][B], d_status [B], d_stats [4][B] = accepted steps, rejected
 * steps, RHS evaluations, Jacobian(+RHS) evaluations.
 * obj_ref (host, [n], may be NULL) + d_obj [B]: fused least-squares objective of
 * the outlet against obj_ref (parameter-estimation populations).
 * ctrl (host, [6], may be NULL = defaults): step-size controller {safety, max
 * shrink factor, max growth factor, kappa, PI beta, initial-step factor}; the error test is
 * rms(err / (kappa*(atol + rtol*max(|y|,|y_new|)))) <= 1. */
int rmt_n1_solve(rmt_module_t m, int64_t B, const double* d_consts, int32_t n_eval, const double* z_eval,
                 double rtol, double atol, int32_t max_steps, int32_t dense, int32_t out_mode,
                 double* d_out, int32_t* d_status, int32_t* d_stats,
                 const double* obj_ref, double* d_obj, const double* ctrl, void* stream);

/* Parameter-estimation populations: rmt_n1_solve for the outlet only (z_end = 1 for N1, ReLe
 * for M7; out_mode 1: d_out [n][B]) with the fused objective AND its reduction: the last block of the integrator kernel to finish folds
 * d_obj[0..B) and d_status[0..B) — in a fixed order, deterministic — into d_red[0..3] = {sum, min, argmin +
 * index_offset (exact as a double), number of reactors with status != 0}.  No second kernel and no host round trip
 * before the cross-GPU step: d_red may point into the buffer that is handed to rmt_comm_allgather.
 * (no reference counterpart: the reference never implemented the estimation loop its README names, README.md:5) */
int rmt_n1_solve_population(rmt_module_t m, int64_t B, const double* d_consts, double z_end, double rtol, double atol,
                            int32_t max_steps, double* d_out, int32_t* d_status, int32_t* d_stats,
                            const double* obj_ref, double* d_obj, double* d_red, int64_t index_offset,
                            const double* ctrl, void* stream);

/* Same path with HOST buffers: copies inputs in, runs setup + solve, copies
 * results back, synchronises.  h_rows [n_rows][B], h_out [n_eval][rows][B].
 * From 2^18 reactors on the library cuts the ensemble into three chunks on two
 * streams of its own, so that the copies run under the integrator kernel (pinned
 * host buffers make the copies asynchronous); the results are the same bits. */
int rmt_n1_solve_host(rmt_module_t m, int64_t B, const double* h_rows, int32_t n_rows, const int32_t* row_map,
                      const double* uniform, int32_t n_eval, const double* z_eval,
                      double rtol, double atol, int32_t max_steps, int32_t dense, int32_t out_mode,
                      double* h_out, int32_t* h_status, int32_t* h_stats,
                      const double* obj_ref, double* h_obj, const double* ctrl);

/* ---- N2: modelEquationN2 (:3706-4134) and the slab loop of runN2 (:3589-3685) ---
 * States are variable-major like the reference's reshape (:3873): d_y, d_f
 * [n][zNo][B].  rmt_n2_solve integrates t in [0, period] and writes the state at
 * the end of each of the tNo slabs: d_out [tNo][rows][zNo][B] (out_mode as for N1;
 * mode 1 rows = y_i, T [K]; mode 2 rows = raw | C_i | (y_i, T)).  The reference
 * restarts solve_ivp at every slab (:3589-3685); here the integration continues and
 * a step is clipped to end on each slab boundary.  d_work is caller-provided scratch
 * of rmt_n2_work_doubles(m, B, zNo) doubles.  ctrl as in rmt_n1_solve. */
int rmt_n2_rhs(rmt_module_t m, int64_t B, int32_t zNo, const double* d_consts, const double* d_y, double* d_f,
               void* stream);
int64_t rmt_n2_work_doubles(rmt_module_t m, int64_t B, int32_t zNo);
int rmt_n2_solve(rmt_module_t m, int64_t B, int32_t zNo, int32_t tNo, double period, const double* d_consts,
                 double rtol, double atol, int32_t max_steps, int32_t out_mode,
                 double* d_out, int32_t* d_status, int32_t* d_stats, double* d_work, const double* ctrl,
                 void* stream);

/* ---- ensemble reductions (per GPU; the cross-GPU step is an NCCL all-reduce of
 * these three numbers, done by the caller's torch.distributed group) ------------- */
int rmt_reduce_objective(rmt_module_t m, int64_t n, const double* d_obj, int64_t index_offset,
                         double* h_sum, double* h_min, int64_t* h_argmin, void* stream);

/* ---- sharded ensembles: the cross-GPU step -------------------------
 * Reactor instances are independent, so an ensemble is split into contiguous blocks, one per GPU (one process per
 * GPU, rmt_init(device) in each), and nothing is exchanged on the data path.  These calls serve the END of a run:
 * gather of results / objectives and the all-reduce of O(1) objective statistics, with NCCL over NVLink
 * (libnccl.so.2 is loaded on first use; when the process already carries an NCCL — PyTorch's — that one is used).
 * The reference has no counterpart: it solves one reactor per rmtExe call (docs/rmtCore.py:393-413).
 *   rank 0: rmt_comm_unique_id(id) -> ship the RMT_COMM_ID_BYTES bytes to the other ranks by any means (file, pipe,
 *           MPI, torch.distributed store) -> every rank: rmt_comm_init(nranks, rank, id) -> collectives -> rmt_comm_free.
 * Buffers are DEVICE pointers, counts are in doubles; the calls are asynchronous on `stream`.
 * rmt_comm_allgather: d_recv [nranks][count].  rmt_comm_allreduce: op 0 = sum, 1 = min, 2 = max. */
int rmt_comm_unique_id(void* id_out, size_t id_bytes);
int rmt_comm_init(int32_t nranks, int32_t rank, const void* id, size_t id_bytes, rmt_comm_t* comm_out);
int rmt_comm_info(rmt_comm_t comm, int32_t* nranks, int32_t* rank, int32_t* nccl_version);
int rmt_comm_allgather(rmt_comm_t comm, const double* d_send, double* d_recv, int64_t count, void* stream);
int rmt_comm_allreduce(rmt_comm_t comm, const double* d_send, double* d_recv, int64_t count, int32_t op, void* stream);
int rmt_comm_free(rmt_comm_t comm);

/* ---- diagnostics: accuracy probe of the branch-free device math used by the generated kinetics:
 * d_out [5][n] = exp(x), log(x), sqrt(x), 10^x, 1/x for x = d_x[0..n). ----------------------------- */
int rmt_math_probe(rmt_module_t m, int32_t n, const double* d_x, double* d_out, void* stream);

/* ---- diagnostics: log the step sequence of one instance of the next N1 solves:
 * d_trace [cap][4] = (t, h, err, accepted) per attempt; NULL switches it off. ------ */
int rmt_debug_trace(double* d_trace, int32_t cap, int64_t instance);

/* ---- measurement helper: FP64 FMA throughput of the device (TFLOP/s) ------------ */
int rmt_fp64_peak(rmt_module_t m, int32_t iters, int32_t repeats, double* tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* RMT_B200_H */
