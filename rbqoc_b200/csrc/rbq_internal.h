// rbq_internal.h — handle layout and kernel launchers shared by the translation units of librbq.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rbq.h"

#define RBQ_NBUF 16

struct rbq_handle {
    int device;
    int sm_count;
    cudaStream_t stream;
    cudaStream_t copy_stream[2];   // H2D / D2H overlap for the chunked host-pointer paths
    cudaEvent_t ev[8];
    int64_t launches;
    char err[512];
    void* dbuf[RBQ_NBUF];          // grow-only device scratch, indexed by role
    size_t dcap[RBQ_NBUF];
    // cached noise-index table of integrate_prop_schroedda! (host copy + key); the device copy is dbuf[B_IDX]
    int32_t* idx_host; size_t idx_cap;
    int64_t idx_nk, idx_gates, idx_len; double idx_dt, idx_ndi;
    // centre (qubit frequency, amplitude offset) the static sweep's knot table is expanded around: rbq_set_sweep_centre
    double centre_fq, centre_da;
    uint64_t table_gen;            // counts the knot tables this handle has built (owner tag of the constant-memory copy)
    unsigned stpos_mask;           // which slices of the structured-Jacobian position table this handle has uploaded
    // tuning / A-B switches: read from the environment ONCE in rbq_create (RBQ_* names), changed afterwards only through
    // rbq_set_option; never read from the environment inside a call.  -1 = "library default / heuristic".
    struct {
        int no_table_class;        // RBQ_NO_TABLE_CLASS      1: polynomial step classes only
        int mc_zerocopy;           // RBQ_MC_ZEROCOPY         0: streamed host sweep even for page-locked buffers
        int chunk_first, chunk_rest;   // RBQ_MC_CHUNK_WAVES  "<first>,<rest>" chunk sizes of the streamed sweep, in waves
        int mc_variant;            // RBQ_MC_VARIANT          <spt><minb><threads/128> of mc_static_kernel
        int peak_mode;             // RBQ_PEAK_MODE           operand form of the DFMA saturation loop
        int noise_gate_parallel;   // RBQ_NOISE_GATE_PARALLEL 0: never the gate-parallel pair
        int noise_materialise;     // RBQ_NOISE_MATERIALISE   0: seed-in noise sweeps generate their pink noise inside the sweep kernel
        int lindblad_shared_map;   // RBQ_LINDBLAD_NO_SHARED_MAP (inverted): 0 = every realisation composes its own map
        int lindblad_table;        // RBQ_LINDBLAD_TABLE      0: series step only (no per-knot Frechet table)
        int lindblad_spt;          // RBQ_LINDBLAD_SPT        realisations per thread of the table sweep (1 or 2)
        int rollout_packed;        // RBQ_ROLLOUT_PACKED      0 / 1 force
        int rollout_scan;          // RBQ_ROLLOUT_SCAN        0 / 1 force the prefix-product open-loop rollout
        int jac_nt_unscented;      // RBQ_JAC_NT_UNSCENTED    generic unscented Jacobian kernel: tangents per lane (1 or 2)
        int jac_structured;        // RBQ_JAC_STRUCTURED      0: generic forward mode for the unscented models too
        int jac_stage_kb;          // RBQ_JAC_STAGE_KB
        int jac_staged;            // RBQ_JAC_STAGED          0 / 1 force
        int bp_tma;                // RBQ_BP_TMA              staging mode mask of the backward pass
        int bp_generic;            // RBQ_BP_GENERIC          1: any-shape kernels only
        int bp_warps;              // RBQ_BP_WARPS            1 / 8 force warp-per-problem or block-per-problem
        int bp_cluster;            // RBQ_BP_CLUSTER          1 / 2 / 4 CTAs per problem (reference shapes n = 43, 56, 58 with one control)
    } opt;
};
// option parsing shared by rbq_create (environment) and rbq_set_option; value NULL or "" restores the default
int rbq_apply_option(rbq_handle* h, const char* name, const char* value);

// grow-only device scratch of the handle for the launchers (role = one of the RBQ_SCR_* slots); returns 0 or an error code
#define RBQ_SCR_JACREC 15
int rbq_scratch(rbq_handle* h, int role, size_t bytes, void** out);

namespace rbq {

// per-block partial statistics, reduced in block order by stats_finalize_kernel (deterministic)
struct BlockPartial {
    double sum, sumsq, mn, mx;
    long long cnt, nonfinite;
};

// gate tables passed by value to the MC kernels
struct GateIso { double g[3][16]; };   // G, G^2, G^3 as row-major 4x4 iso matrices (spin.jl:314-323)
struct GateRot { double r[3][9]; };    // the same gates as SO(3) rotations of the Bloch vector

struct McStaticArgs {
    const double* table; int64_t nk; double dt;   // knot table (q_j, c0_j, c1_j, c2_j), see static_table_kernel
    const double* table_m; uint64_t table_gen;    // rows of the factored form, or NULL; serial number of this table (per handle)
    double* quat;                                 // scratch of the constant-memory form: 4 doubles per sample of this launch, or NULL
    int host_io;                                  // sample / fidelity pointers are page-locked host memory (zero-copy call)
    int64_t win_first; int win_knots;             // constant-memory form: the window of knots this launch covers (set by the launcher)
    int64_t gate_count; int64_t first; int64_t S;
    const double* fq; const double* da; const double* psi0;   // explicit samples
    int philox; uint64_t seed; double fq0, fq_rel_sigma, fq_rel_clip, da_sigma;
    double* fid;                 // S*(gate_count+1) or NULL
    BlockPartial* partials;      // one per block, or NULL
    unsigned long long* hist;    // RBQ_HIST_BINS global bins, or NULL
    const double* scal;          // device scalars: max|a|, f0, a0, max|q_j|, table_ok, max x_j
    int no_table_class;          // RBQ_NO_TABLE_CLASS=1: polynomial classes only (A/B measurements, parity of both)
    GateIso gate;
};
struct McNoiseArgs {
    const double* table; int64_t nk; double dt;    // knot table centred on (fq, 0), see static_table_kernel
    int64_t gate_count; int64_t first; int64_t S; double fq;
    const double* noise; int64_t noise_len; double noise_dt_inv; const double* psi0;
    const int32_t* idx_table; int64_t idx_stride;   // noise-sample index per (gate, knot), rows padded to 4 entries
    int philox; uint64_t seed; double namp;
    double* fid; BlockPartial* partials; unsigned long long* hist;
    const double* scal;          // device scalars of the table: max|a|, f0, a0, max|q_j|, table_ok, max x_j
    int no_table_class;
    GateIso gate;
};
struct LindbladArgs {
    const double* controls; int64_t nk; double dt;
    int64_t gate_count; int64_t S; double fq;
    const double* da; const double* rho0; double t2_gamma_c, t2_gamma_f;
    double* fid; double* rho_out; BlockPartial* partials; unsigned long long* hist; const double* amax;
    const double* shared_map;   // 3x3 one-gate Bloch map common to all realisations (column images), or NULL
    GateRot gate;
};

struct SegmentArgs {
    const double* amps; const double* durs; int64_t nseg;
    int64_t gate_count, S;
    const double* fq; double fq_scalar; const double* da; const double* psi0; const double* rho0;
    const double* noise; const int32_t* nidx; int64_t noise_len;   // optional: amplitude += noise[s][nidx[j]]
    int channels; double t2_gamma_c, t2_gamma_f;
    double* states_out;          // timeline: iso states at every save point, S x (nsave+1) x 4, or NULL
    double* rho_saves;           // density timeline: vec(rho) at every save point, S x (nsave+1) x 8, or NULL
    double* fid; double* rho_out; BlockPartial* partials; unsigned long long* hist;
    GateIso gate; GateRot rot;
};

// every launcher enqueues on h->stream, bumps h->launches and returns a cudaError_t as int
int launch_pulse_prep(rbq_handle* h, const double* d_controls, int64_t nk, double* d_amax);
// knot table of the static sweep, expanded around the sweep centre (f0, a0)
int launch_static_table(rbq_handle* h, const double* d_controls, int64_t nk, double dt, double f0, double a0, double* d_table,
                        double* d_scal, double* d_table_m = nullptr);
void ctab_forget(const rbq_handle* h);   // a handle that goes away gives up the constant-memory table (rbq_destroy)
int launch_mc_static(rbq_handle* h, const McStaticArgs& a, int algo, int* nblocks_out);
int launch_mc_noise(rbq_handle* h, const McNoiseArgs& a, int algo, int* nblocks_out);
int launch_mc_noise_gate_parallel(rbq_handle* h, const McNoiseArgs& a, double* d_quat, int* nblocks_out);
int launch_segments(rbq_handle* h, const SegmentArgs& a, bool lindblad, int* nblocks_out);
int launch_timeline(rbq_handle* h, const SegmentArgs& a, const int64_t* d_save_after, int64_t nsave, int* nblocks_out);
int launch_lindblad_shared_map(rbq_handle* h, const LindbladArgs& a, int channels, double* d_map);
int launch_lindblad(rbq_handle* h, const LindbladArgs& a, int channels, int* nblocks_out);
// per-knot Frechet table (d_table: nk rows of 48 doubles) + the sweep kernel that reads it; needs a.da != NULL
int launch_lindblad_table(rbq_handle* h, const LindbladArgs& a, int channels, const double* d_controls, double* d_table,
                          int* nblocks_out);
int mc_static_blocks(const rbq_handle* h, int64_t S, int algo);
int64_t mc_static_wave_samples(const rbq_handle* h, int algo);
int mc_noise_blocks(const rbq_handle* h, int64_t S);
int lindblad_blocks(const rbq_handle* h, int64_t S);
int launch_stats_finalize(rbq_handle* h, const BlockPartial* partials, int nblocks, rbq_stats* d_stats);
int launch_stats_reset(rbq_handle* h, rbq_stats* d_stats);
int launch_philox_samples(rbq_handle* h, int64_t first, int64_t S, uint64_t seed, double fq0, double fq_rel_sigma,
                          double fq_rel_clip, double da_sigma, double* d_fq, double* d_da, double* d_psi0);
int launch_pink_noise(rbq_handle* h, int64_t first, int64_t S, uint64_t seed, double namp, double noise_dt_inv,
                      int64_t noise_len, double* d_noise);
int launch_fp64_peak(rbq_handle* h, int iters, double* d_sink, int* nthreads_out, int* flops_per_thread_iter);

int launch_rollout(rbq_handle* h, const rbq_model_params& p, int64_t B, int64_t N, double t0, const double* d_x0,
                   const double* d_U, const double* d_dt, double* d_X);
int launch_closedloop(rbq_handle* h, const rbq_model_params& p, int64_t N, const double* d_Xref, const double* d_Uref,
                      const double* d_K, const double* d_d, const double* d_dt, int64_t nalpha, const double* d_alphas,
                      double* d_X, double* d_Uout);
int launch_gate_error(rbq_handle* h, int64_t K, const double* d_s1, int64_t st1, const double* d_s2, int64_t st2, double* d_err,
                      double* d_grad, double* d_hess);
int launch_sampling_cost(rbq_handle* h, int64_t K, const double* d_X, const double* d_U, const double* Qdiag, double R,
                         const double* xf, const double* targets, const double* q_ss, double* d_J, double* d_gx, double* d_gu);
int launch_backward_pass(rbq_handle* h, int n, int m, int64_t N, int64_t B, const double* d_AB, const double* d_Exx,
                         const double* d_Ex, const double* d_Euu, const double* d_Eux, const double* d_Eu, double rho,
                         int reg_type, double* d_K, double* d_d, double* d_dV, long long* d_fail);
int launch_al_cost(rbq_handle* h, int n, int m, int64_t N, int64_t B, int64_t dual_stride, const double* d_X, const double* d_U, const double* d_Qd,
                   const double* d_Rd, const double* d_Qfd, const double* d_xf, const double* d_w, int ncon,
                   const rbq_constraint* d_cons, const int32_t* d_idx_pool, const double* d_val_pool, const double* d_lambda,
                   const double* d_mu, double* d_Exx, double* d_Ex, double* d_Euu, double* d_Eux, double* d_Eu, double* d_c,
                   double* d_J);
int launch_al_dual_update(rbq_handle* h, int ncon, const rbq_constraint* d_cons, int64_t max_len, int64_t B, int64_t dual_stride,
                          const double* d_c, double* d_lambda, double* d_mu, double phi, double mu_max, double lambda_max,
                          int update, double* d_viol);
int launch_jacobians(rbq_handle* h, const rbq_model_params& p, int64_t K, const double* d_X, const double* d_U,
                     const double* d_t, const double* d_dt, double* d_AB);
// structured Jacobians of the unscented models (rbq_jac_unscented.cu): one sample state, d_AB 16-byte aligned
size_t unscented_record_bytes(int64_t K);
int launch_jacobians_unscented(rbq_handle* h, const rbq_model_params& p, int64_t K, const double* d_X, const double* d_U,
                               const double* d_t, const double* d_dt, double* d_rec, double* d_AB);

}  // namespace rbq
