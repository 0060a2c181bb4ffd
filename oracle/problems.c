/*
 * problems.c -- hand-written continuous right-hand sides f(t,y,p) and Jacobians df/dy for the
 * oracle.  TEST INFRASTRUCTURE (see rst_oracle.h).  Deliberately written by hand with loops, NOT
 * produced by rustedscithe_b200/emit, so that the CUDA emitter and the oracle are two independent
 * derivations of the same mathematics.
 *
 * Problem definitions follow (paths relative to /root/reference/):
 *  linear2      examples/bvp_damped_aot_guide.rs:67-72       y' = z, z' = 0              (names y,z)
 *  damp1        src/symbolic/codegen/tests/codegen_test_support.rs:136-143 as *observed* in the
 *               checked-in generated fixture src/symbolic/codegen/testing_fixtures/
 *               test_codegen_generated_bvp_fixtures.rs:13-66: z' = y - z, y' = z^3 (names z,y).
 *               (The source string is "-z^3"; the reference parser drops that leading unary minus,
 *               parse_expr.rs:187-197, and the fixture pins the resulting +z^3.)
 *  combustion6  src/numerical/BVP_Damp/tests/backend_compare.rs:72-149
 *  flame12      SURVEY.md section 8(d) C3: the same template with 5 species pairs
 *  coupled64    SURVEY.md section 8(d) C5: 31 species pairs, every rate depends on all C_k
 */
#include "rst_oracle.h"
#include <math.h>
#include <string.h>

/* ---------------- linear2 ---------------- */
static void linear2_rhs(double t, const double *y, const double *p, double *f) {
    (void)t; (void)p;
    f[0] = y[1];
    f[1] = 0.0;
}
static void linear2_jac(double t, const double *y, const double *p, double *fy) {
    (void)t; (void)p; (void)y;
    fy[0] = 0.0; fy[1] = 1.0; fy[2] = 0.0; fy[3] = 0.0;
}

/* ---------------- damp1 (names z, y) ---------------- */
static void damp1_rhs(double t, const double *y, const double *p, double *f) {
    (void)t; (void)p;
    const double z = y[0], yy = y[1];
    f[0] = yy - z;
    f[1] = pow(z, 3.0);
}
static void damp1_jac(double t, const double *y, const double *p, double *fy) {
    (void)t; (void)p;
    const double z = y[0];
    fy[0] = -1.0; fy[1] = 1.0;
    fy[2] = 3.0 * z * z; fy[3] = 0.0;
}

/* ---------------- flame template (combustion6 = 2 species, flame12 = 5 species) ---------------- */
/* constants: backend_compare.rs:81-112 */
#define Q_HEAT (3000.0 * 1e3 * 0.034)
#define DT 600.0
#define T_SCALE 600.0
#define L_LEN 3e-4
#define M0 (34.2 / 1000.0)
#define LAMBDA 0.07
#define P_PRESS 2e6
#define TM 1500.0
#define PE_Q 0.0090168
#define D_RO 2.88e-4
#define PE_D 1.50e-3
#define A_PRE 1.3e5
#define E_ACT (5000.0 * 4.184)
#define R_G 8.314
#define M_REAG 0.342

/* product weights: w_k = 2 (nsp-k) / (nsp (nsp-1)), k = 1..nsp-1 ; nsp = 2 gives w_1 = 1 (reference) */
static double flame_weight(int nsp, int k) {
    return 2.0 * (double)(nsp - k) / ((double)nsp * (double)(nsp - 1));
}

static void flame_rhs_n(int nsp, const double *y, double *f) {
    const double ro_m = M0 * P_PRESS / (R_G * TM);
    const double qm = pow(L_LEN, 2.0) / T_SCALE;
    const double qs = pow(L_LEN, 2.0);
    const double teta = y[0], q = y[1], c0 = y[2];
    const double rate = A_PRE * exp(-E_ACT / (R_G * (teta * T_SCALE + DT))) * c0 * (ro_m / M_REAG);
    f[0] = q / LAMBDA;
    f[1] = q * PE_Q - Q_HEAT * rate * qm;
    for (int k = 0; k < nsp; ++k) {
        const double jk = y[3 + 2 * k];
        f[2 + 2 * k] = jk / D_RO;
        if (k == 0)
            f[3] = jk * PE_D + rate * ro_m * qs;
        else
            f[3 + 2 * k] = jk * PE_D - flame_weight(nsp, k) * rate * ro_m * qs;
    }
}
static void flame_jac_n(int nsp, const double *y, double *fy) {
    const int nv = 2 + 2 * nsp;
    const double ro_m = M0 * P_PRESS / (R_G * TM);
    const double qm = pow(L_LEN, 2.0) / T_SCALE;
    const double qs = pow(L_LEN, 2.0);
    const double teta = y[0], c0 = y[2];
    const double T = teta * T_SCALE + DT;
    const double ex = A_PRE * exp(-E_ACT / (R_G * T)) * (ro_m / M_REAG);
    const double rate = ex * c0;
    const double drate_dteta = rate * (E_ACT / (R_G * T * T)) * T_SCALE;
    const double drate_dc0 = ex;
    memset(fy, 0, sizeof(double) * (size_t)nv * nv);
    fy[0 * nv + 1] = 1.0 / LAMBDA;
    fy[1 * nv + 1] = PE_Q;
    fy[1 * nv + 0] = -Q_HEAT * qm * drate_dteta;
    fy[1 * nv + 2] = -Q_HEAT * qm * drate_dc0;
    for (int k = 0; k < nsp; ++k) {
        const int rc = 2 + 2 * k, rj = 3 + 2 * k;
        const double s = (k == 0) ? 1.0 : -flame_weight(nsp, k);
        fy[rc * nv + rj] = 1.0 / D_RO;
        fy[rj * nv + rj] = PE_D;
        fy[rj * nv + 0] = s * ro_m * qs * drate_dteta;
        fy[rj * nv + 2] = s * ro_m * qs * drate_dc0;
    }
}
static void combustion6_rhs(double t, const double *y, const double *p, double *f) { (void)t; (void)p; flame_rhs_n(2, y, f); }
static void combustion6_jac(double t, const double *y, const double *p, double *fy) { (void)t; (void)p; flame_jac_n(2, y, fy); }
static void flame12_rhs(double t, const double *y, const double *p, double *f) { (void)t; (void)p; flame_rhs_n(5, y, f); }
static void flame12_jac(double t, const double *y, const double *p, double *fy) { (void)t; (void)p; flame_jac_n(5, y, fy); }

/* ---------------- coupled template (coupled64 = 31 species): dense C-coupling ----------------
 * rate_k = A exp(-E_k/(R T)) (ro_m/M) (e0 C_0 + eps sum_m a_km C_m),  a_km = 1/(1+|k-m|),  E_k = E (1 + 0.01 k)
 * coupled8: (eps, e0) = (1, 0);  coupled64: (0.01, 1) -- weak dense coupling on top of the flame rate law so that the
 * reference's damped loop converges from the flat guess
 * Teta' = q/lambda ; q' = Pe_q q - Q qm sum_k w_k rate_k ; C_k' = J_k/(ro D) ;
 * J_k' = Pe_D J_k + s_k rate_k ro_m qs,  s_0 = +1, s_k = -w_k,  w_k = 1/nsp */
static void coupled_rates(int nsp, double eps, double e0, const double *y, double *rate, double *ex) {
    const double ro_m = M0 * P_PRESS / (R_G * TM);
    const double T = y[0] * T_SCALE + DT;
    for (int k = 0; k < nsp; ++k) {
        const double Ek = E_ACT * (1.0 + 0.01 * (double)k);
        double s = 0.0;
        for (int m = 0; m < nsp; ++m) s += eps * y[2 + 2 * m] / (1.0 + (double)(k > m ? k - m : m - k));
        s += e0 * y[2];
        ex[k] = A_PRE * exp(-Ek / (R_G * T)) * (ro_m / M_REAG);
        rate[k] = ex[k] * s;
    }
}
static void coupled_rhs_n(int nsp, double eps, double e0, const double *y, double *f) {
    const double ro_m = M0 * P_PRESS / (R_G * TM);
    const double qm = pow(L_LEN, 2.0) / T_SCALE;
    const double qs = pow(L_LEN, 2.0);
    double rate[64], ex[64];
    coupled_rates(nsp, eps, e0, y, rate, ex);
    const double w = 1.0 / (double)nsp;
    double heat = 0.0;
    for (int k = 0; k < nsp; ++k) heat += w * rate[k];
    f[0] = y[1] / LAMBDA;
    f[1] = y[1] * PE_Q - Q_HEAT * qm * heat;
    for (int k = 0; k < nsp; ++k) {
        const double jk = y[3 + 2 * k];
        const double s = (k == 0) ? 1.0 : -w;
        f[2 + 2 * k] = jk / D_RO;
        f[3 + 2 * k] = jk * PE_D + s * rate[k] * ro_m * qs;
    }
}
static void coupled_jac_n(int nsp, double eps, double e0, const double *y, double *fy) {
    const int nv = 2 + 2 * nsp;
    const double ro_m = M0 * P_PRESS / (R_G * TM);
    const double qm = pow(L_LEN, 2.0) / T_SCALE;
    const double qs = pow(L_LEN, 2.0);
    const double T = y[0] * T_SCALE + DT;
    double rate[64], ex[64];
    coupled_rates(nsp, eps, e0, y, rate, ex);
    const double w = 1.0 / (double)nsp;
    memset(fy, 0, sizeof(double) * (size_t)nv * nv);
    fy[0 * nv + 1] = 1.0 / LAMBDA;
    fy[1 * nv + 1] = PE_Q;
    for (int k = 0; k < nsp; ++k) {
        const double Ek = E_ACT * (1.0 + 0.01 * (double)k);
        const double drdteta = rate[k] * (Ek / (R_G * T * T)) * T_SCALE;
        const int rc = 2 + 2 * k, rj = 3 + 2 * k;
        const double s = (k == 0) ? 1.0 : -w;
        fy[rc * nv + rj] = 1.0 / D_RO;
        fy[rj * nv + rj] = PE_D;
        fy[rj * nv + 0] = s * ro_m * qs * drdteta;
        fy[1 * nv + 0] += -Q_HEAT * qm * w * drdteta;
        for (int m = 0; m < nsp; ++m) {
            const double a = eps / (1.0 + (double)(k > m ? k - m : m - k)) + (m == 0 ? e0 : 0.0);
            fy[rj * nv + 2 + 2 * m] = s * ro_m * qs * ex[k] * a;
            fy[1 * nv + 2 + 2 * m] += -Q_HEAT * qm * w * ex[k] * a;
        }
    }
}
static void coupled64_rhs(double t, const double *y, const double *p, double *f) { (void)t; (void)p; coupled_rhs_n(31, 0.01, 1.0, y, f); }
static void coupled64_jac(double t, const double *y, const double *p, double *fy) { (void)t; (void)p; coupled_jac_n(31, 0.01, 1.0, y, fy); }
static void coupled8_rhs(double t, const double *y, const double *p, double *f) { (void)t; (void)p; coupled_rhs_n(3, 1.0, 0.0, y, f); }
static void coupled8_jac(double t, const double *y, const double *p, double *fy) { (void)t; (void)p; coupled_jac_n(3, 1.0, 0.0, y, fy); }

/* ---------------- parameterised test problem: z' = y - z, y' = p0 * z^3 + p1 * t ---------------- */
static void damp1p_rhs(double t, const double *y, const double *p, double *f) {
    f[0] = y[1] - y[0];
    f[1] = p[0] * pow(y[0], 3.0) + p[1] * t;
}
static void damp1p_jac(double t, const double *y, const double *p, double *fy) {
    (void)t;
    fy[0] = -1.0; fy[1] = 1.0;
    fy[2] = p[0] * 3.0 * y[0] * y[0]; fy[3] = 0.0;
}

static const orc_problem PROBLEMS[] = {
    {"linear2", 2, 0, linear2_rhs, linear2_jac},
    {"damp1", 2, 0, damp1_rhs, damp1_jac},
    {"damp1p", 2, 2, damp1p_rhs, damp1p_jac},
    {"combustion6", 6, 0, combustion6_rhs, combustion6_jac},
    {"flame12", 12, 0, flame12_rhs, flame12_jac},
    {"coupled8", 8, 0, coupled8_rhs, coupled8_jac},
    {"coupled64", 64, 0, coupled64_rhs, coupled64_jac},
};

const orc_problem *orc_problem_get(const char *name) {
    for (size_t i = 0; i < sizeof(PROBLEMS) / sizeof(PROBLEMS[0]); ++i)
        if (strcmp(PROBLEMS[i].name, name) == 0) return &PROBLEMS[i];
    return NULL;
}

/* Structural pattern = union of non-zeros of df/dy over a few generic probe points (the symbolic
 * route prunes structurally, View/jacobian.rs:111-128; probing at generic points reproduces that
 * for these smooth right-hand sides). */
void orc_problem_pattern(const orc_problem *pb, unsigned char *pattern) {
    const int nv = pb->nv;
    double y[64], fy[64 * 64], p[8];
    memset(pattern, 0, (size_t)nv * nv);
    for (int probe = 0; probe < 3; ++probe) {
        for (int i = 0; i < nv; ++i) y[i] = 0.37 + 0.11 * probe + 0.013 * i;
        for (int i = 0; i < 8; ++i) p[i] = 0.71 + 0.05 * i;
        pb->jac(0.3 + 0.2 * probe, y, p, fy);
        for (int i = 0; i < nv * nv; ++i)
            if (fy[i] != 0.0) pattern[i] = 1;
    }
}
