/*
 * rbq.h — C ABI of the B200-native batched propagation engine for rbqoc's hot path.
 *
 * Every entry point is what a Julia `ccall` (or Python `ctypes`) stub binds in place of
 * the reference function cited beside it.  Citations are `file:line` relative to the
 * reference tree (SchusterLab/rbqoc).  No torch / C++ types cross this boundary: plain
 * pointers, sizes and POD structs only.
 *
 * Conventions
 *  - All floating point data is IEEE binary64 ("double").
 *  - Augmented states / controls are dense vectors laid out exactly like the reference's
 *    `astate` / `acontrol` SVectors (0-based here, 1-based there).
 *  - Jacobians are column-major n x (n+m) blocks `[A_k B_k]`, the layout
 *    RobotDynamics' `discrete_jacobian!(::Type{RK3}, ∇f, model, z)` fills.
 *  - Trajectories are knot-major: X[k*n + i], U[k*m + j]  (what `permutedims` at
 *    spin13.jl:180-183 produces, seen from C).
 *  - Every function returns int: 0 = ok, <0 = argument error (RBQ_E_*), >0 = CUDA error
 *    code.  Nothing throws or aborts across the boundary.  `rbq_last_error` gives text.
 *  - Host-pointer entry points are synchronous: results are in host memory on return.
 *    `_dev` entry points take device pointers, enqueue on the handle's stream and
 *    return without synchronising.
 *  - A handle is bound to one CUDA device and is not thread-safe; distinct handles are
 *    independent.
 *  - There is no CPU fallback: without a usable CUDA device `rbq_create` fails.
 */
#ifndef RBQ_H
#define RBQ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RBQ_VERSION 100

/* ---- error codes (negative = caller error) ---------------------------------------- */
#define RBQ_OK 0
#define RBQ_E_NULL (-1)      /* a required pointer was NULL                        */
#define RBQ_E_MODEL (-2)     /* unknown model_id / inconsistent model parameters   */
#define RBQ_E_SIZE (-3)      /* a size/count argument is out of range              */
#define RBQ_E_GATE (-4)      /* unknown gate id                                    */
#define RBQ_E_ALGO (-5)      /* unknown propagator algorithm                       */
#define RBQ_E_NODEVICE (-6)  /* no CUDA device / wrong architecture                */
#define RBQ_E_ALIGN (-7)     /* device pointer not suitably aligned                */

/* ---- enums mirrored from the reference --------------------------------------------- */
/* GateType, spin.jl:29-34 */
#define RBQ_GATE_ZPIBY2 1
#define RBQ_GATE_YPIBY2 2
#define RBQ_GATE_XPIBY2 3
#define RBQ_GATE_XPI 4

/* model ids = the number in the reference file name that defines the Model */
#define RBQ_MODEL_SPIN11 11 /* derivative robust, d/dfq      spin11.jl:51-117  */
#define RBQ_MODEL_SPIN12 12 /* sampling robust, fq +- sigma  spin12.jl:52-56,165-192 */
#define RBQ_MODEL_SPIN13 13 /* nominal                       spin13.jl:37-58   */
#define RBQ_MODEL_SPIN15 15 /* decay aware / time optimal    spin15.jl:40-73   */
#define RBQ_MODEL_SPIN17 17 /* derivative robust, d/da       spin17.jl:48-114  */
#define RBQ_MODEL_SPIN18 18 /* sampling robust, a +- namp    spin18.jl:50-55,162-190 */
#define RBQ_MODEL_SPIN25 25 /* unscented sampling, amplitude noise through the pink filter (time dependent)
                               spin25.jl:67-239 */
#define RBQ_MODEL_SPIN30 30 /* unscented sampling, fq        spin30.jl:60-202  */

/* propagator algorithm for exp(dt*(f*H0 + a*H1)) */
#define RBQ_ALGO_SU2 0  /* closed form: (dt H)^2 = -theta^2 I  =>  cos + sinc series  */
#define RBQ_ALGO_PADE 1 /* reference algorithm: generic 4x4 Pade-3/5/7/9/13 + LU solve
                           (StaticArrays 0.12.5 `exp`, the call at spin13.jl:46)      */

/* integrator of the augmented state.  EXP: the discrete_dynamics of the model's reference file (matrix exponential + Euler
 * chain), registered there under the tag RK3 (rbqoc.jl:49-54).  RK2 / RK3 / RK4: RobotDynamics' explicit Runge-Kutta steps
 * (RobotDynamics 0.2.1, external; rbqoc.jl:41-54 IT_RDI) of the CONTINUOUS dynamics RD.dynamics(model, x, u, t) as the reference's
 * RK variants define them (spin15_standalone.jl:263-312; derivative states as spin19.jl:66-129 / spin21.jl:63-109):
 *     psi' = (fq H0 + a H1) psi,  (int a)' = a,  a' = da,  da' = u1,  [decay aware] (int gamma)' = 1 / T1(a) (T1 in ns,
 *     amp_t1_spline),  [derivative order o] (d^o psi)' = o H_d d^(o-1) psi + (fq H0 + a H1) d^o psi  (H_d = H0: spin11, H1: spin17),
 *     [time optimal] everything times u2^2 with the solver's dt (spin15_standalone.jl:283-287).
 * Models: spin13, spin15, spin11, spin17.  Open-loop rollouts, per-knot steps and Jacobians; the closed-loop rollout kernels
 * implement the exponential step only. */
#define RBQ_INTEGRATOR_EXP 0
#define RBQ_INTEGRATOR_RK2 2
#define RBQ_INTEGRATOR_RK3 3
#define RBQ_INTEGRATOR_RK4 4

/* Lindblad channels (bit mask) */
#define RBQ_LINDBLAD_T1 1 /* amplitude damping up+down at gamma_1 = 1/T1(a)  spin.jl:529-531 */
#define RBQ_LINDBLAD_T2 2 /* pure dephasing Gamma(t) = gamma_c + 2 gamma_f^2 t  spin.jl:542-554 */

#define RBQ_MAX_SAMPLE_STATES 4
#define RBQ_HIST_BINS 256

/*
 * Plain-old-data mirror of the reference's immutable `Model` structs:
 *   spin12.jl:52-54  Model(h0_samples)          -> fq_samples[2]
 *   spin18.jl:50-52  Model(namp)                -> namp
 *   spin30.jl:60-66  Model(S, nominal_idxs, fq_cov, alpha, sample_state_count)
 *   spin25.jl:67-74  Model(S, nominal_idxs, namp, wnamp_cov, alpha, sample_state_count)
 *                    -> namp, fq_cov (= wnamp_cov), alpha, S_diag, sample_state_count
 *   spin15.jl:40-42  Model{DA,TO}               -> decay_aware, time_optimal
 *   spin11.jl:51-52  Model{DO} / spin17.jl      -> derivative_order
 * `fq` replaces the global constant FQ (spin.jl:148) so that batches of detuned
 * problems can share one code path; set it to 1.4e-2 for the reference behaviour.
 */
typedef struct rbq_model_params {
    int32_t model_id;
    int32_t derivative_order;   /* spin11/17: 0,1,2                                   */
    int32_t decay_aware;        /* spin15 DA                                          */
    int32_t time_optimal;       /* spin15 TO                                          */
    int32_t sample_state_count; /* spin30: 1..RBQ_MAX_SAMPLE_STATES                   */
    int32_t algo;               /* RBQ_ALGO_*                                         */
    int32_t nominal_offset[RBQ_MAX_SAMPLE_STATES]; /* spin30: 0-based first index of
                                   the nominal state of sample state i
                                   (= nominal_idxs[i][1]-1, spin30.jl:62,168)         */
    int32_t integrator;         /* RBQ_INTEGRATOR_*: how x_{k+1} is formed (default 0 = the exponential step of the model file) */
    int32_t reserved;
    double fq;                  /* GHz                                                */
    double fq_samples[2];       /* spin12: (FQ+fq_cov), (FQ-fq_cov)  spin12.jl:204-205 */
    double namp;                /* spin18                                             */
    double fq_cov;              /* spin30 (used as a variance, spin30.jl:102,138); spin25: wnamp_cov */
    double alpha;               /* spin30                                             */
    double S_diag[4];           /* spin30 penalty weights, spin30.jl:215              */
} rbq_model_params;

/* Infidelity statistics of a Monte-Carlo sweep.
 * Additive across shards: count/nonfinite/hist add, sum/sumsq add, min/max combine. */
typedef struct rbq_stats {
    int64_t count;      /* finite samples accumulated                                 */
    int64_t nonfinite;  /* samples whose infidelity was NaN/Inf (returned, not hidden) */
    double sum;         /* sum of infidelities 1-F                                    */
    double sumsq;       /* sum of squares                                             */
    double min;
    double max;
    int64_t hist[RBQ_HIST_BINS]; /* log10 histogram over [1e-16,1): bin =
                                    floor((log10(e)+16)*16) clamped to [0,255]         */
} rbq_stats;

typedef struct rbq_handle rbq_handle;

/* ---- lifetime ------------------------------------------------------------------------ */
int rbq_version(void);
/* device = CUDA ordinal.  Fails with RBQ_E_NODEVICE when no sm_100 device is present. */
int rbq_create(int device, rbq_handle** out);
int rbq_destroy(rbq_handle* h);
const char* rbq_last_error(const rbq_handle* h);
/* cuda_stream = a cudaStream_t (NULL = the legacy default stream). */
int rbq_set_stream(rbq_handle* h, void* cuda_stream);
int rbq_synchronize(rbq_handle* h);
/* Performance hint for rbq_mc_static(_dev): the point (qubit frequency, amplitude offset) the explicit samples scatter
 * around.  Default (FQ, 0) = (1.4e-2 GHz, 0), the reference's FQ (spin.jl:148).  Samples near the centre take the
 * 14-instruction knot-table step, the others the polynomial step; both meet the same accuracy bound, so the centre moves
 * results by rounding only (~1e-15).  Philox sweeps are centred on their own fq0.  Non-finite values are rejected. */
int rbq_set_sweep_centre(rbq_handle* h, double fq0, double da0);
/* Page-locked host memory for the caller's buffers (e.g. the nabla-f storage a Julia caller hands to rbq_jacobians):
 * host-pointer entry points accept any host memory, but copies from/to pinned memory run at full PCIe rate and let the
 * chunked sweeps overlap upload, compute and download.  Free with rbq_host_free (h may be NULL there: page-locked memory
 * is process wide and may outlive the handle that allocated it). */
int rbq_host_alloc(rbq_handle* h, int64_t bytes, void** out);
int rbq_host_free(rbq_handle* h, void* ptr);
/* Page-lock (and map for the device) memory the caller already owns -- e.g. a Julia Array that is reused across calls --
 * so that rbq_mc_static can read it in place (zero-copy form) and the other host-pointer calls copy at full PCIe rate.
 * cudaHostRegister / cudaHostUnregister underneath; the range must stay allocated until it is unregistered. */
int rbq_host_register(rbq_handle* h, void* ptr, int64_t bytes);
int rbq_host_unregister(rbq_handle* h, void* ptr);
/* Tuning / A-B switches.  The library reads its RBQ_* environment variables ONCE, in rbq_create, into the handle; no call
 * reads the environment afterwards.  rbq_set_option changes one switch of one handle later on (same names and value
 * syntax as the environment variables; value NULL or "" restores the default).  Names: RBQ_NO_TABLE_CLASS, RBQ_MC_ZEROCOPY,
 * RBQ_MC_CHUNK_WAVES ("first,rest"), RBQ_MC_VARIANT, RBQ_PEAK_MODE, RBQ_NOISE_GATE_PARALLEL, RBQ_NOISE_MATERIALISE, RBQ_LINDBLAD_NO_SHARED_MAP,
 * RBQ_LINDBLAD_TABLE, RBQ_LINDBLAD_SPT, RBQ_ROLLOUT_PACKED, RBQ_ROLLOUT_SCAN, RBQ_JAC_NT_UNSCENTED, RBQ_JAC_STRUCTURED, RBQ_JAC_STAGE_KB,
 * RBQ_JAC_STAGED, RBQ_BP_TMA, RBQ_BP_GENERIC, RBQ_BP_WARPS, RBQ_BP_CLUSTER.  Every switch selects between kernels that meet the same parity
 * bound; none changes results beyond rounding.  Unknown names / malformed values return RBQ_E_SIZE. */
int rbq_set_option(rbq_handle* h, const char* name, const char* value);
/* number of kernels this handle has launched since creation (bench `gpu_launches`). */
int64_t rbq_launch_count(const rbq_handle* h);

/* ---- model API: RD.state_dim / RD.control_dim (spin13.jl:39-40, spin30.jl:67-70) ------ */
int rbq_state_dim(const rbq_model_params* p);   /* <0 on error */
int rbq_control_dim(const rbq_model_params* p); /* <0 on error */

/* ---- per-knot: RD.discrete_dynamics(::Type{RK3}, model, x, u, t, dt)  (spin13.jl:44-58,
 *      spin12.jl:165-192, spin30.jl:177-202, spin15.jl:50-73, spin11.jl:64-117,
 *      spin17.jl:61-114, spin18.jl:162-190).  Host pointers. */
int rbq_discrete_dynamics(rbq_handle* h, const rbq_model_params* p, const double* x,
                          const double* u, double t, double dt, double* xnext);
/* RD.discrete_jacobian!(::Type{RK3}, ∇f, model, z) — default is ForwardDiff.jacobian over
 * [x;u] (RobotDynamics 0.2.1); AB is column-major n x (n+m). */
int rbq_discrete_jacobian(rbq_handle* h, const rbq_model_params* p, const double* x,
                          const double* u, double t, double dt, double* AB);

/* ---- trajectory level (what Altro's `for k = 1:N-1` loops become) ------------------------
 * rollout: B independent trajectories.  x0[B*n], U[B*(N-1)*m], dt[N-1] (shared by the
 * batch; ignored by time-optimal spin15 which takes dt = u2^2), X[B*N*n].  Knot times start at 0 and accumulate
 * (t[k+1] = t[k] + dt[k]), as the KnotPoints of a TO.Traj do. */
int rbq_rollout(rbq_handle* h, const rbq_model_params* p, int64_t B, int64_t N,
                const double* x0, const double* U, const double* dt, double* X);
/* closed-loop rollout used by the iLQR line search: for each candidate alpha_c
 *   du = K_k (x_k - Xref_k) + alpha_c d_k ;  u_k = Uref_k + du ;  x_{k+1} = f(x_k,u_k)
 * K[(N-1)*m*n] column-major m x n per knot, d[(N-1)*m].  X[nalpha*N*n], Uout[nalpha*(N-1)*m]. */
int rbq_rollout_closedloop(rbq_handle* h, const rbq_model_params* p, int64_t N,
                           const double* Xref, const double* Uref, const double* K,
                           const double* d, const double* dt, int64_t nalpha,
                           const double* alphas, double* X, double* Uout);
/* device-pointer variant (asynchronous on the handle's stream): with rbq_jacobians_dev and rbq_backward_pass_dev an
 * iLQR iteration keeps X, U, the Jacobians and the gains on the device */
int rbq_rollout_closedloop_dev(rbq_handle* h, const rbq_model_params* p, int64_t N,
                           const double* d_Xref, const double* d_Uref, const double* K,
                           const double* d, const double* dt, int64_t nalpha,
                           const double* alphas, double* X, double* Uout);
/* jacobians: K independent knots (a whole trajectory, or a batch of them, flattened).
 * X[K*n], U[K*m], dt[K], AB[K*n*(n+m)] (each block column-major).  t[K] = the knot times z.t; may be NULL (= 0) for
 * every model but spin25, the only one whose dynamics read their time argument (spin25.jl:196). */
int rbq_jacobians(rbq_handle* h, const rbq_model_params* p, int64_t K, const double* X,
                  const double* U, const double* t, const double* dt, double* AB);
int rbq_rollout_dev(rbq_handle* h, const rbq_model_params* p, int64_t B, int64_t N,
                    const double* d_x0, const double* d_U, const double* d_dt, double* d_X);
int rbq_jacobians_dev(rbq_handle* h, const rbq_model_params* p, int64_t K, const double* d_X,
                      const double* d_U, const double* d_t, const double* d_dt, double* d_AB);

/* ---- cost-side primitives of the sampling objectives (first "next" row of the scope table):
 *      gate_error_iso2, jacobian_gate_error_iso2, hessian_gate_error_iso2  (spin.jl:1040-1092), the per-knot terms of
 *      the custom Cost of spin12.jl:59-161 / spin18.jl:58-160.
 *  s1: K states of 4 doubles, consecutive states s1_stride doubles apart (so the caller can point into an augmented
 *      trajectory X at a sample state's offset with stride n);  s2: K targets, s2_stride apart (0 = one shared target).
 *  err[K]      1 - a^2 - b^2
 *  grad[K*4]   d err / d s1            (may be NULL)
 *  hess[K*16]  the constant-in-s1 4x4 matrix the reference returns (row-major; may be NULL) */
int rbq_gate_error_expansion(rbq_handle* h, int64_t K, const double* s1, int64_t s1_stride, const double* s2,
                             int64_t s2_stride, double* err, double* grad, double* hess);
int rbq_gate_error_expansion_dev(rbq_handle* h, int64_t K, const double* d_s1, int64_t s1_stride, const double* d_s2,
                                 int64_t s2_stride, double* d_err, double* d_grad, double* d_hess);

/* TO.stage_cost and TO.gradient! of the sampling objective `Cost` (spin12.jl:59-161, spin18.jl:58-160) for K knots of the
 * 43-state spin12/spin18 layout (11 base states + 8 sample states of 4):
 *   J[k]       = 0.5 x'Qx + q'x + c + sum_{i<8} q_ss[i%4] gate_error_iso2(x[11+4i..], target[i%4])  (+ 0.5 u'Ru if U)
 *   gx[k*43..] = Qx + q + [0 (11); q_ss[i%4] jacobian_gate_error_iso2(x[11+4i..], target[i%4])]
 *   gu[k]      = R u                                   (NULL when U is NULL: the terminal cost)
 * with q = -Q xf and c = 0.5 xf'Q xf as the constructor at spin12.jl:77-79 sets them.  The Hessian of this cost is constant
 * (hess_astate, spin12.jl:80-100) and stays with the caller.  targets[16] = 4 iso targets, q_ss[4] their weights. */
int rbq_sampling_cost_expansion(rbq_handle* h, int64_t K, const double* X, const double* U, const double* Qdiag,
                                const double* Rdiag, const double* xf, const double* targets, const double* q_ss,
                                double* J, double* gx, double* gu);

/* ---- Monte-Carlo robustness evaluation: run_sim_prop(...; dynamics_type=schroed)
 *      (spin.jl:1553-1608 -> integrate_prop_schroed! spin.jl:438-452 -> compute_fidelities
 *      spin.jl:1198-1240), batched over S samples instead of one call per (parameter, seed).
 *
 *  controls[nknots]   pulse amplitudes a_j (GHz), nknots = control_knot_count
 *  dt                 1/controls_dt_inv
 *  fq[S]              per-sample qubit frequency: negi_h0 = fq[s]*NEGI_H0_ISO  (figures.jl:270)
 *  da[S] or NULL      per-sample static amplitude offset added to every knot
 *  psi0[S*4]          per-sample initial iso state (gen_rand_state_iso, spin.jl:1119-1138)
 *  fid[S*(gate_count+1)] or NULL   fidelities after 0..gate_count gates, sample-major
 *  stats or NULL      statistics of the infidelity after the last gate
 *
 *  Host buffers may be any host memory.  If fq, da, psi0 and fid are all page-locked (rbq_host_alloc, cudaHostRegister,
 *  torch pin_memory) and gate_count <= 3, the sweep runs zero-copy: one kernel reads the samples and writes the
 *  fidelities over PCIe itself; otherwise the samples are streamed through the device in chunks.  Same results, bit
 *  for bit (RBQ_MC_ZEROCOPY=0 forces the streamed form).
 */
int rbq_mc_static(rbq_handle* h, const double* controls, int64_t nknots, double dt, int gate,
                  int64_t gate_count, int64_t S, const double* fq, const double* da,
                  const double* psi0, int algo, double* fid, rbq_stats* stats);
/* device-pointer variant; d_stats points to a device rbq_stats that is ACCUMULATED into
 * (zero it with rbq_stats_reset_dev first).  d_fid may be NULL. */
int rbq_mc_static_dev(rbq_handle* h, const double* d_controls, int64_t nknots, double dt,
                      int gate, int64_t gate_count, int64_t S, const double* d_fq,
                      const double* d_da, const double* d_psi0, int algo, double* d_fid,
                      rbq_stats* d_stats);
/* Samples drawn on the device by a counter-based generator keyed by the GLOBAL sample
 * index (Philox4x32-10; key = seed, counter = index), so any sharding of
 * [first, first+S) over GPUs reproduces the same samples:
 *   fq   = fq0*(1 + clip(fq_rel_sigma*N(0,1), +-fq_rel_clip))      (figures.jl:261,386)
 *   da   = da_sigma*N(0,1)
 *   psi0 = (u1 + i u2, u3 + i u4)/norm                              (spin.jl:1132-1133)
 */
int rbq_mc_static_philox_dev(rbq_handle* h, const double* d_controls, int64_t nknots, double dt,
                             int gate, int64_t gate_count, int64_t first, int64_t S,
                             uint64_t seed, double fq0, double fq_rel_sigma,
                             double fq_rel_clip, double da_sigma, int algo, double* d_fid,
                             rbq_stats* d_stats);
/* materialise the samples the generator above draws (for parity tests / the oracle) */
/* host-pointer form of the above: controls on the host, fid (or NULL) and stats returned to the host */
int rbq_mc_static_philox(rbq_handle* h, const double* controls, int64_t nknots, double dt, int gate,
                         int64_t gate_count, int64_t first, int64_t S, uint64_t seed, double fq0,
                         double fq_rel_sigma, double fq_rel_clip, double da_sigma, int algo,
                         double* fid, rbq_stats* stats);
/* ---- single-process multi-device sweeps ("rbq_create(device_ids[], n) with the communicator inside the handle").
 * A group owns one handle per listed CUDA device (a device may be listed more than once).  A group sweep shards the global
 * sample range [first, first+S) contiguously over the members, launches every member's share before waiting for any (one host
 * thread drives all GPUs), and merges the members' statistics in member order on the host -- the same numbers any sharding
 * gives, because the samples are keyed by their global index.  No device-to-device traffic: the only exchange is the 2.1 KB
 * rbq_stats per member.  (One process per GPU + NCCL, rbqoc_b200/dist.py, is the other supported shape.) */
typedef struct rbq_group rbq_group;
int rbq_group_create(const int* devices, int n, rbq_group** out);
int rbq_group_destroy(rbq_group* g);
int rbq_group_size(const rbq_group* g);
rbq_handle* rbq_group_handle(rbq_group* g, int member);
const char* rbq_group_last_error(const rbq_group* g);
/* rbq_mc_static_philox over the group; fid[S*(gate_count+1)] (host, or NULL) is filled in global sample order */
int rbq_group_mc_static_philox(rbq_group* g, const double* controls, int64_t nknots, double dt, int gate, int64_t gate_count,
                               int64_t first, int64_t S, uint64_t seed, double fq0, double fq_rel_sigma, double fq_rel_clip,
                               double da_sigma, int algo, double* fid, rbq_stats* stats);
int rbq_philox_samples(rbq_handle* h, int64_t first, int64_t S, uint64_t seed, double fq0,
                       double fq_rel_sigma, double fq_rel_clip, double da_sigma, double* fq,
                       double* da, double* psi0);
int rbq_stats_reset_dev(rbq_handle* h, rbq_stats* d_stats);
/* stats helpers on the host: merge b into a (what the NCCL all-reduce computes) */
int rbq_stats_merge(rbq_stats* a, const rbq_stats* b);

/* ---- time-dependent amplitude noise: run_sim_prop(...; dynamics_type=schroedda)
 *      (integrate_prop_schroedda! spin.jl:469-484).
 *  noise[S*noise_len] per-sample noise offsets (already scaled by namp, spin.jl:1571),
 *  looked up as noise[floor(time*noise_dt_inv)] with `time` accumulated by repeated
 *  `time += dt` exactly like spin.jl:476-480. */
int rbq_mc_noise(rbq_handle* h, const double* controls, int64_t nknots, double dt, int gate,
                 int64_t gate_count, int64_t S, double fq, const double* noise,
                 int64_t noise_len, double noise_dt_inv, const double* psi0, int algo,
                 double* fid, rbq_stats* stats);
int rbq_mc_noise_dev(rbq_handle* h, const double* d_controls, int64_t nknots, double dt,
                     int gate, int64_t gate_count, int64_t S, double fq, const double* d_noise,
                     int64_t noise_len, double noise_dt_inv, const double* d_psi0, int algo,
                     double* d_fid, rbq_stats* d_stats);
/* noise generated on the fly: white N(0,1) from Philox(seed, sample index) through the
 * 3-pole pink filter of gen_pink_noise (spin.jl:1178-1192), times noise_dt_inv*namp. */
int rbq_mc_pink_philox_dev(rbq_handle* h, const double* d_controls, int64_t nknots, double dt,
                           int gate, int64_t gate_count, int64_t first, int64_t S,
                           uint64_t seed, double fq, double namp, double noise_dt_inv,
                           int algo, double* d_fid, rbq_stats* d_stats);
int rbq_mc_pink_philox(rbq_handle* h, const double* controls, int64_t nknots, double dt, int gate,
                       int64_t gate_count, int64_t first, int64_t S, uint64_t seed, double fq,
                       double namp, double noise_dt_inv, int algo, double* fid, rbq_stats* stats);
/* materialise that noise: noise[S*noise_len] */
int rbq_pink_noise(rbq_handle* h, int64_t first, int64_t S, uint64_t seed, double namp,
                   double noise_dt_inv, int64_t noise_len, double* noise);

/* ---- Lindblad T1 rollout: run_sim_prop(...; dynamics_type=lindbladt1)
 *      (integrate_prop_lindbladt1! spin.jl:521-539, fidelity_mat spin.jl:1101-1104).
 *  rho0[S*8]: per-sample 2x2 complex density, column-major (vec(rho), spin.jl:524), each
 *  entry (re,im); the Hermitian part is used.  da[S] or NULL: static amplitude offset per
 *  realisation (enters both the Hamiltonian and gamma_1(a)).  channels: RBQ_LINDBLAD_T1 for the
 *  reference's lindbladt1 path, RBQ_LINDBLAD_T2 for the pure dephasing
 *  Gamma(t) = t2_gamma_c + 2 t2_gamma_f^2 t of dynamics_lindbladt2_deqjl (spin.jl:542-554),
 *  evaluated at knot mid-points, or both.
 *  fid[S*(gate_count+1)] or NULL: Uhlmann sqrt-fidelity per gate.  rho_out[S*8] or NULL:
 *  final density.  stats: of 1-fid after the last gate. */
int rbq_lindblad(rbq_handle* h, const double* controls, int64_t nknots, double dt, int gate,
                 int64_t gate_count, int64_t S, double fq, const double* da,
                 const double* rho0, int channels, double t2_gamma_c, double t2_gamma_f,
                 double* fid, double* rho_out, rbq_stats* stats);
int rbq_lindblad_dev(rbq_handle* h, const double* d_controls, int64_t nknots, double dt,
                     int gate, int64_t gate_count, int64_t S, double fq, const double* d_da,
                     const double* d_rho0, int channels, double t2_gamma_c, double t2_gamma_f,
                     double* d_fid, double* d_rho_out, rbq_stats* d_stats);

/* ---- analytic gates: piecewise-constant pulses with PER-SEGMENT durations
 *      integrate_prop_zpiby2nodis!/zpiby2t1!, ypiby2nodis!/ypiby2t1!, xpiby2nodis!/xpiby2t1!
 *      (spin.jl:562-594, 651-670, 721-727, 765-785): the gate propagator is the ordered product of exp(segment generator *
 *      segment duration), applied gate_count times.  amps[nseg] (GHz), durs[nseg] (ns); everything else as in
 *      rbq_mc_static / rbq_lindblad (da[S] or NULL shifts every segment's amplitude).  Host pointers. */
int rbq_mc_static_segments(rbq_handle* h, const double* amps, const double* durs, int64_t nseg, int gate,
                           int64_t gate_count, int64_t S, const double* fq, const double* da, const double* psi0,
                           double* fid, rbq_stats* stats);
/* integrate_prop_xpiby2da! (spin.jl:819-905): segment j uses amplitude amps[j] + noise[s][noise_idx[j]]; the same
 * indices are read in every gate, as in the reference.  noise[S*noise_len], noise_idx[nseg] (0-based, int32). */
int rbq_mc_noise_segments(rbq_handle* h, const double* amps, const double* durs, const int32_t* noise_idx, int64_t nseg,
                          int gate, int64_t gate_count, int64_t S, double fq, const double* noise, int64_t noise_len,
                          const double* psi0, double* fid, rbq_stats* stats);
/* run_sim_deqjl(...; dynamics_type = schroed / schroedda) (spin.jl:1277-1380): the ODE the reference hands to Vern9 has a
 * right-hand side that is constant between the breakpoints of the control grid (index floor(t dt_inv) mod knot count,
 * spin.jl:425, 456), the noise grid (floor(t noise_dt_inv), spin.jl:457) and the save times (multiples of the gate time), so
 * its exact solution is the ordered product of segment exponentials.  The caller supplies that segment list: amplitude,
 * duration and (optional) noise index per segment, and save_after[g] = number of segments before save point g+1 (ascending).
 * fid[S*(nsave+1)]: fidelity against the 4-cyclic targets at t = 0 and at every save point; da[S] or NULL is a static
 * offset; noise[S*noise_len] with noise_idx[nseg], or both NULL. */
int rbq_mc_timeline(rbq_handle* h, const double* amps, const double* durs, const int32_t* noise_idx, int64_t nseg,
                    const int64_t* save_after, int64_t nsave, int gate, int64_t S, const double* fq, const double* da,
                    const double* noise, int64_t noise_len, const double* psi0, double* fid, rbq_stats* stats);
/* the same with the iso states themselves at t = 0 and at every save point: states[S*(nsave+1)*4] (or NULL) -- what
 * run_sim_fine_deqjl (spin.jl:1389-1478) stores as "states" on its save_step grid */
int rbq_mc_timeline_states(rbq_handle* h, const double* amps, const double* durs, const int32_t* noise_idx, int64_t nseg,
                           const int64_t* save_after, int64_t nsave, int gate, int64_t S, const double* fq, const double* da,
                           const double* noise, int64_t noise_len, const double* psi0, double* fid, double* states, rbq_stats* stats);
/* run_sim_deqjl / run_sim_fine_deqjl for the density dynamics: dynamics_lindbladnodis_deqjl (channels = 0),
 * dynamics_lindbladt1_deqjl (RBQ_LINDBLAD_T1) and dynamics_lindbladt2_deqjl (RBQ_LINDBLAD_T2, Gamma(t) = t2_gamma_c +
 * 2 t2_gamma_f^2 t with t running from the start of the timeline) -- spin.jl:488-518, 542-554 -- over the same kind of
 * segment timeline (amplitude constant inside a segment).  Exact for the time-independent channels; one fourth-order Magnus
 * step per segment for the running rate.  rho0[S*8] as in rbq_lindblad; fid[S*(nsave+1)] Uhlmann sqrt-fidelity against the
 * 4-cyclic targets (or NULL); rho_saves[S*(nsave+1)*8] the densities (column-major vec, (re,im)) at t = 0 and every save point
 * (or NULL). */
int rbq_lindblad_timeline(rbq_handle* h, const double* amps, const double* durs, int64_t nseg, const int64_t* save_after,
                          int64_t nsave, int gate, int64_t S, double fq, const double* da, const double* rho0, int channels,
                          double t2_gamma_c, double t2_gamma_f, double* fid, double* rho_saves, rbq_stats* stats);
int rbq_lindblad_segments(rbq_handle* h, const double* amps, const double* durs, int64_t nseg, int gate,
                          int64_t gate_count, int64_t S, double fq, const double* da, const double* rho0, int channels,
                          double t2_gamma_c, double* fid, double* rho_out, rbq_stats* stats);

/* ---- iLQR backward pass -------------------------------------------------------------------------------
 * Altro.jl 0.2.0 `backwardpass!` (external: SchusterLab/rbqoc-stable fork, Manifest.toml:26-32) for B independent problems
 * of N knots, state dimension n <= 64, control dimension m <= 4, one thread block per problem.  It consumes the Jacobian
 * blocks exactly as rbq_jacobians(_dev) writes them, so they never leave the device between the expansion and the pass.
 *   AB  [B][N-1][n*(n+m)]  column-major n x (n+m) blocks [A_k B_k]
 *   Exx [B][N][n*n]        cost Hessian blocks, column-major, symmetric with BOTH triangles stored (even n copies whole
 *                          columns by TMA; odd n and the terminal block read the lower triangle); knot N-1 = terminal
 *   Ex  [B][N][n]          cost gradients
 *   Euu [B][N-1][m*m], Eux [B][N-1][m*n] (column-major m x n), Eu [B][N-1][m]
 *   rho, reg_type          Quu_reg = Quu + rho I (RBQ_REG_CONTROL) or Quu + rho B'B, Qux_reg = Qux + rho B'A (RBQ_REG_STATE)
 *   K   [B][N-1][m*n] (column-major m x n), d [B][N-1][m], dV [B][2] = (sum d'Qu, sum d'Quu d / 2)
 *   fail[B]                -1, or the 0-based knot at which Quu_reg was not positive definite (the pass stops there, as the
 *                          reference does before it raises rho and restarts; gains of later knots are valid, earlier not) */
#define RBQ_REG_CONTROL 0
#define RBQ_REG_STATE 1
int rbq_backward_pass_dev(rbq_handle* h, int n, int m, int64_t N, int64_t B, const double* d_AB, const double* d_Exx,
                          const double* d_Ex, const double* d_Euu, const double* d_Eux, const double* d_Eu, double rho,
                          int reg_type, double* d_K, double* d_d, double* d_dV, int64_t* d_fail);
int rbq_backward_pass(rbq_handle* h, int n, int m, int64_t N, int64_t B, const double* AB, const double* Exx, const double* Ex,
                      const double* Euu, const double* Eux, const double* Eu, double rho, int reg_type, double* K, double* d,
                      double* dV, int64_t* fail);

/* ---- augmented-Lagrangian cost expansion (constraint side) ---------------------------------------------
 * What Altro / TrajectoryOptimization (external, Manifest.toml:26-32, 1523-1529) do per knot on the host in
 * cost_expansion!(E, ::ALObjective, Z), for the problem class the reference sets up (spin13.jl:122-153 and siblings):
 * a diagonal LQRObjective (Q, R, Qf, xf) plus BoundConstraint / GoalConstraint / NormConstraint, all knots in one launch,
 * written where rbq_backward_pass_dev reads:
 *   J[k]   = w_k (1/2 (x-xf)'Q(x-xf) + 1/2 u'Ru)  [Qf, no control term at k = N-1]  + sum_con lambda'c + 1/2 c' I_mu c
 *   Exx_k  = w_k Q + sum_con cx' I_mu cx,      Ex_k = w_k Q (x-xf) + sum_con cx'(lambda + I_mu c)      (same for Euu, Eu; Eux = 0)
 * with I_mu = diag(a .* mu), a_j = 1 for equalities and (c_j >= 0) | (lambda_j > 0) for inequalities (first-order
 * constraint expansion, as the reference solver's).  w (N per-knot weights, NULL = 1) carries the solver's dt scaling of
 * stage costs where its version applies one.
 * A constraint covers knots [k0, k1) (0-based) and owns p entries per knot of lambda / mu / c at dual_off + (k - k0) p:
 *   RBQ_CON_BOUND  p = 2(n+m): z - zmax <= 0 (z = [x; u]) then zmin - z <= 0; val_pool[val_off..] = zmax[n+m], zmin[n+m],
 *                  +-infinity = no bound (c = -inf is written); control entries are skipped at the last knot
 *   RBQ_CON_GOAL   p = nidx: x[idx_j] - val_pool[val_off + j] = 0               (indices distinct)
 *   RBQ_CON_NORM   p = 1:    sum_j x[idx_j]^2 - val_pool[val_off] = 0
 * d_Exx = NULL: values only (J and c), e.g. the merit function of line-search candidates.
 * B trajectories at once (B <= 65535): X[B][N][n], U[B][N-1][m], every output with a leading batch dimension (c: the table's
 * total dual length per trajectory; the table's rows must be in increasing dual_off order); lambda / mu are shared
 * (dual_stride = 0: candidates of one problem) or per trajectory, dual_stride doubles apart (independent problems).
 * All pointers are device pointers; asynchronous on the handle's stream. */
#define RBQ_CON_BOUND 0
#define RBQ_CON_GOAL 1
#define RBQ_CON_NORM 2
typedef struct rbq_constraint {
    int32_t kind, p;
    int64_t k0, k1;
    int32_t nidx, idx_off;
    int64_t val_off, dual_off;
} rbq_constraint;
int rbq_al_cost_expansion_dev(rbq_handle* h, int n, int m, int64_t N, int64_t B, const double* d_X, const double* d_U,
                              const double* d_Qdiag, const double* d_Rdiag, const double* d_Qfdiag, const double* d_xf,
                              const double* d_w, int ncon, const rbq_constraint* d_cons, const int32_t* d_idx_pool,
                              const double* d_val_pool, const double* d_lambda, const double* d_mu, int64_t dual_stride,
                              double* d_Exx, double* d_Ex, double* d_Euu, double* d_Eux, double* d_Eu, double* d_c, double* d_J);
/* Outer-loop updates of the same table on the device (Altro dual_update!, penalty_update!, max_violation; external):
 *   lambda <- clamp(lambda + mu c, lo, lambda_max), lo = 0 for RBQ_CON_BOUND (inequalities), -lambda_max for equalities;
 *   mu <- min(phi mu, mu_max);   *d_max_violation = max over entries of |c| (equalities) or max(c, 0) (inequalities).
 * c as rbq_al_cost_expansion_dev wrote it (entries of absent bounds, -inf, are skipped); update = 0: violation only.
 * max_p = the largest p of the table (sizes the grid); with dual_stride = 0 and B > 1 the violation covers all B
 * trajectories and the shared multipliers are updated from trajectory 0. */
int rbq_al_dual_update_dev(rbq_handle* h, int64_t N, int64_t B, int ncon, const rbq_constraint* d_cons, int64_t max_p,
                           const double* d_c, double* d_lambda, double* d_mu, int64_t dual_stride, double phi, double mu_max,
                           double lambda_max, int update, double* d_max_violation);
/* host-pointer form (synchronous): the same with every array in host memory; nidx / nval = lengths of the two pools. */
int rbq_al_cost_expansion(rbq_handle* h, int n, int m, int64_t N, int64_t B, const double* X, const double* U,
                          const double* Qdiag, const double* Rdiag, const double* Qfdiag, const double* xf, const double* w,
                          int ncon, const rbq_constraint* cons, const int32_t* idx_pool, int64_t nidx, const double* val_pool,
                          int64_t nval, const double* lambda, const double* mu, int64_t dual_stride, double* Exx, double* Ex,
                          double* Euu, double* Eux, double* Eu, double* c, double* J);

/* ---- measurement aid: saturating DFMA loop, returns achieved FP64 TFLOP/s -------------- */
int rbq_fp64_peak(rbq_handle* h, int iters, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* RBQ_H */
