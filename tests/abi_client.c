/* abi_client.c — a plain-C process (no Python, no torch) driving librbq.so exactly the way a Julia `ccall` does:
 * opaque handle, POD model struct by pointer, caller-owned host arrays, integer return codes.  Prints results as text;
 * tests/test_abi_client_gpu.py compares them with the oracle.  Build: gcc abi_client.c -I include -L rbqoc_b200 -lrbq */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rbq.h"

#define CHECK(call)                                                                         \
    do {                                                                                    \
        int rc_ = (call);                                                                   \
        if (rc_ != RBQ_OK) {                                                                \
            fprintf(stderr, "%s -> %d (%s)\n", #call, rc_, h ? rbq_last_error(h) : "");     \
            return 2;                                                                       \
        }                                                                                   \
    } while (0)

static void print_vec(const char* name, const double* v, int n) {
    printf("%s", name);
    for (int i = 0; i < n; ++i) printf(" %.17g", v[i]);
    printf("\n");
}

int main(void) {
    rbq_handle* h = NULL;
    printf("version %d\n", rbq_version());
    CHECK(rbq_create(0, &h));

    /* spin12.jl model: n = 43, m = 1 */
    rbq_model_params p;
    memset(&p, 0, sizeof(p));
    p.model_id = RBQ_MODEL_SPIN12;
    p.fq = 1.4e-2;
    p.fq_samples[0] = 1.4e-2 * 1.01;
    p.fq_samples[1] = 1.4e-2 * 0.99;
    p.alpha = 1.0;
    const int n = rbq_state_dim(&p), m = rbq_control_dim(&p);
    printf("dims %d %d\n", n, m);
    double x[43], u[1] = {0.05}, xn[43];
    for (int i = 0; i < n; ++i) x[i] = 0.1 * ((i * 7) % 11) - 0.5;
    x[9] = 0.37;
    CHECK(rbq_discrete_dynamics(h, &p, x, u, 0.0, 0.1, xn));
    print_vec("step", xn, n);

    /* pinned nabla-f storage, as INTEGRATION.md recommends */
    double* AB = NULL;
    CHECK(rbq_host_alloc(h, (int64_t)sizeof(double) * n * (n + m), (void**)&AB));
    CHECK(rbq_discrete_jacobian(h, &p, x, u, 0.0, 0.1, AB));
    print_vec("jac", AB, n * (n + m));
    CHECK(rbq_host_free(h, AB));

    /* run_sim_prop-style sweep: 4 samples, 50 knots, 2 gates */
    enum { NK = 50, S = 4, G = 2 };
    double controls[NK], fq[S], da[S], psi0[4 * S], fid[S * (G + 1)];
    for (int j = 0; j < NK; ++j) controls[j] = 0.3 * (j % 7) / 7.0 - 0.1;
    for (int s = 0; s < S; ++s) {
        fq[s] = 1.4e-2 * (1.0 + 0.01 * (s - 1.5));
        da[s] = 1e-5 * s;
        psi0[4 * s] = 0.5; psi0[4 * s + 1] = 0.5; psi0[4 * s + 2] = s % 2 ? 0.5 : -0.5; psi0[4 * s + 3] = 0.5;
    }
    rbq_stats st;
    CHECK(rbq_mc_static(h, controls, NK, 0.1, RBQ_GATE_XPIBY2, G, S, fq, da, psi0, RBQ_ALGO_SU2, fid, &st));
    print_vec("fid", fid, S * (G + 1));
    printf("stats %lld %lld %.17g\n", (long long)st.count, (long long)st.nonfinite, st.sum);

    /* argument errors come back as codes, never as crashes */
    printf("bad_gate %d\n", rbq_mc_static(h, controls, NK, 0.1, 99, G, S, fq, da, psi0, RBQ_ALGO_SU2, fid, &st));
    printf("null_ptr %d\n", rbq_discrete_dynamics(h, &p, NULL, u, 0.0, 0.1, xn));
    printf("launches %lld\n", (long long)rbq_launch_count(h));
    CHECK(rbq_destroy(h));
    return 0;
}
