/* A host with no Python and no CUDA headers: plain C99 against include/rst_b200.h and librst_b200.so.
 * Solves the reference's two-variable linear BVP  y' = z, z' = 0, y(0) = 0, z(1) = 1  (closed form y = x, z = 1; forward Euler
 * is exact for it) through the same calls a Rust host would bind:
 *   rst_b200_create -> rst_b200_host_alloc -> rst_b200_set_state -> rst_b200_newton_solve -> rst_b200_get_state.
 * Exit status 0 with "NO DEVICE" on a box without a GPU (the library must refuse: there is no CPU fallback),
 * 0 with "OK ..." when the solve matches the closed form, 1 otherwise. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "rst_b200.h"

int main(int argc, char **argv) {
    const long long n = argc > 1 ? atoll(argv[1]) : 20000;
    int nv = 0, np = 0;
    const char *hash = NULL;
    if (rst_b200_problem_info("linear2", &nv, &np, &hash) != RST_B200_OK || nv != 2 || np != 0) {
        printf("FAIL problem registry: nv %d np %d\n", nv, np);
        return 1;
    }
    const int bc_var[2] = {0, 1}, bc_side[2] = {0, 1};
    const double bc_val[2] = {0.0, 1.0};
    const double lo[2] = {-0.2, -2.0}, hi[2] = {1.2, 2.0}, rtol[2] = {1e-6, 1e-6};
    rst_b200_desc d;
    memset(&d, 0, sizeof(d));
    d.problem_key = "linear2";
    d.n_steps = n;
    d.scheme = RST_B200_SCHEME_FORWARD;
    d.t0 = 0.0;
    d.t_end = 1.0;
    d.n_bc = 2;
    d.bc_var = bc_var; d.bc_side = bc_side; d.bc_val = bc_val;
    d.lo = lo; d.hi = hi; d.rtol = rtol;
    d.world = 1;
    rst_b200_solver *s = NULL;
    const int rc = rst_b200_create(&d, &s);
    if (rc == RST_B200_ERR_CUDA && rst_b200_device_count() <= 0) {
        printf("NO DEVICE: %s\n", rst_b200_last_error());
        return s == NULL ? 0 : 1;
    }
    if (rc != RST_B200_OK) { printf("FAIL create: %d %s\n", rc, rst_b200_last_error()); return 1; }
    const size_t len = (size_t)(2 * n);                 /* reduced unknowns: 2 (N + 1) values minus the 2 pinned ones */
    double *y = NULL;
    if (rst_b200_host_alloc(len * sizeof(double), (void **)&y) != RST_B200_OK) { printf("FAIL host_alloc\n"); return 1; }
    for (size_t i = 0; i < len; ++i) y[i] = 0.25;       /* a constant start, far from the answer */
    rst_b200_newton_opts o = {40, 3, 5, 0.5, 1e-8};
    rst_b200_newton_stats st;
    int ok = rst_b200_set_state(s, y, len) == RST_B200_OK && rst_b200_newton_solve(s, &o, &st) == RST_B200_OK && st.status == 1 &&
             rst_b200_get_state(s, y, len) == RST_B200_OK;
    double *full = (double *)malloc((size_t)(2 * (n + 1)) * sizeof(double)), err = 0.0;
    ok = ok && rst_b200_get_full_solution(s, full, (size_t)(2 * (n + 1))) == RST_B200_OK;
    if (ok)
        for (long long j = 0; j <= n; ++j) {
            const double ey = fabs(full[2 * j] - (double)j / (double)n), ez = fabs(full[2 * j + 1] - 1.0);
            if (ey > err) err = ey;
            if (ez > err) err = ez;
        }
    printf("%s iterations %d linear_solves %d (reused %d) max error %.3e\n", ok && err < 1e-9 ? "OK" : "FAIL", st.iterations,
           st.linear_solves, st.solves_reused, err);
    free(full);
    rst_b200_host_free(y);
    rst_b200_destroy(s);
    return ok && err < 1e-9 ? 0 : 1;
}
