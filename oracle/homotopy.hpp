// ORACLE (test infrastructure, NOT product code): the homotopy combinators.
//
// Follows (reference file:line):
//   src/homotopies/straight_line_homotopy.jl:81-154   StraightLineHomotopy
//   src/homotopies/parameter_homotopy.jl:66-101       ParameterHomotopy
//   src/homotopies/coefficient_homotopy.jl:89-141     CoefficientHomotopy
//   src/homotopies/toric_homotopy.jl:66-278           ToricHomotopy (real t only)
//   src/systems/fixed_parameter_system.jl:28-33       FixedParameterSystem (params baked in)
//   src/model_kit/abstract_system_homotopy.jl:96-120  operator API
#pragma once
#include "tape.hpp"

namespace orc {

enum HKind : int32_t { H_STRAIGHT_LINE = 0, H_PARAMETER = 1, H_COEFFICIENT = 2, H_TORIC = 3 };

// Immutable description shared by all tracker copies.
struct HomotopyDef {
    HKind kind = H_PARAMETER;
    const System* F = nullptr;  // target system (SL) or the parametrised system
    const System* G = nullptr;  // start system (SL only)
    cplx gamma;                 // SL
    std::vector<cplx> G_params; // SL: fixed parameters of G (`scaling`, total_degree.jl:102-104)
    std::vector<cplx> F_params; // SL: fixed parameters of F (coefficients, solve.jl:101-106)
    std::vector<cplx> p, q;     // PARAMETER/COEFFICIENT: start p (t=1), target q (t=0)
    std::vector<cplx> system_coeffs;  // TORIC: u_i
    int m() const { return F->m(); }
    int n() const { return F->n(); }
};

// Mutable per-tracker instance (caches, scratch).
struct Homotopy {
    const HomotopyDef* D = nullptr;
    SystemWS F, G;
    int m = 0, n = 0, P = 0;
    std::vector<cplx> p, q;          // PARAMETER/COEFFICIENT (per-path overridable: solve.jl:856)
    std::vector<cplx> pt, dpt;       // p(t), p - q
    std::vector<double> weights, t_weights;  // TORIC
    std::vector<cplx> coeffs, dt_coeffs, tc; // TORIC: coeffs, d/dt, taylor coeffs (5 x P)
    std::vector<cplx> u2, U2, tay1, tay2, tpbuf;
    std::vector<cdd> vdd, udd;

    void init(const HomotopyDef* d) {
        D = d; m = d->m(); n = d->n();
        F.init(d->F);
        P = d->F->eval.P;
        if (d->kind == H_STRAIGHT_LINE) {
            G.init(d->G);
            u2.assign(m, cplx()); U2.assign((size_t)m * n, cplx());
            vdd.assign(m, cdd()); udd.assign(m, cdd());
        }
        p = d->p; q = d->q;
        pt.assign(P, cplx()); dpt.assign(P, cplx());
        if (d->kind == H_TORIC) {
            weights.assign(P, 0.0); t_weights.assign(P, 0.0);
            coeffs.assign(P, cplx()); dt_coeffs.assign(P, cplx()); tc.assign((size_t)5 * P, cplx());
        }
        tay1.assign((size_t)TMAX * m, cplx()); tay2.assign((size_t)TMAX * m, cplx());
        tpbuf.assign((size_t)2 * P, cplx());
    }

    // parameter_homotopy.jl:66-87 tp! / coefficient_homotopy.jl:89-107 coeffs!
    void linear_params(cplx t) {
        if (t.im == 0.0) {
            double s = t.re, s1 = 1.0 - s;
            for (int i = 0; i < P; ++i) { pt[i] = s * p[i] + s1 * q[i]; dpt[i] = p[i] - q[i]; }
        } else {
            cplx t1 = cplx(1.0) - t;
            for (int i = 0; i < P; ++i) { pt[i] = t * p[i] + t1 * q[i]; dpt[i] = p[i] - q[i]; }
        }
    }
    // toric_homotopy.jl:114-121, 145-177 (real t >= 0 only; Appendix B.21 of SURVEY.md)
    void toric_coeffs(double t) {
        const auto& u = D->system_coeffs;
        if (t == 0.0) {
            for (int i = 0; i < P; ++i) coeffs[i] = weights[i] == 0 ? u[i] : cplx();
        } else {
            double s = std::log(t);
            for (int i = 0; i < P; ++i) { t_weights[i] = std::exp(weights[i] * s); coeffs[i] = u[i] * t_weights[i]; }
        }
    }
    void toric_dt_coeffs(double t) {  // :180-204
        const auto& u = D->system_coeffs;
        if (t == 0.0) {
            for (int i = 0; i < P; ++i) dt_coeffs[i] = weights[i] == 1 ? u[i] : cplx();
        } else {
            toric_coeffs(t);
            double sinv = 1.0 / t;
            for (int i = 0; i < P; ++i) dt_coeffs[i] = (weights[i] * coeffs[i]) * sinv;
        }
    }
    void toric_taylor_coeffs(double t) {  // :220-264; tc[r*P + i] = coefficient r
        const auto& u = D->system_coeffs;
        if (t == 0.0) {
            for (int i = 0; i < P; ++i) {
                double w = weights[i];
                for (int r = 0; r < 5; ++r) tc[(size_t)r * P + i] = cplx();
                if (w < 1e-12) tc[i] = u[i];
                else if (std::fabs(w - 1.0) <= std::sqrt(EPS) * std::fmax(std::fabs(w), 1.0)) tc[(size_t)P + i] = u[i];
            }
        } else {
            double s = std::log(t);
            for (int i = 0; i < P; ++i) t_weights[i] = std::exp(weights[i] * s);
            double tinv = 1.0 / t;
            for (int i = 0; i < P; ++i) {
                double w = weights[i], tw = t_weights[i];
                tc[i] = u[i] * tw;
                double tw1 = w * tw * tinv;
                tc[(size_t)P + i] = u[i] * tw1;
                double tw2 = 0.5 * (w - 1) * tw1 * tinv;
                tc[(size_t)2 * P + i] = u[i] * tw2;
                double tw3 = (w - 2) * tw2 * tinv / 3;
                tc[(size_t)3 * P + i] = u[i] * tw3;
                double tw4 = 0.25 * (w - 3) * tw3 * tinv;
                tc[(size_t)4 * P + i] = u[i] * tw4;
            }
        }
    }

    const cplx* params_at(cplx t) {
        switch (D->kind) {
            case H_PARAMETER: case H_COEFFICIENT: linear_params(t); return pt.data();
            case H_TORIC: toric_coeffs(t.re); return coeffs.data();
            default: return nullptr;
        }
    }

    void evaluate(cplx* u, const cplx* x, cplx t) {
        if (D->kind == H_STRAIGHT_LINE) {  // straight_line_homotopy.jl:96-104
            G.evaluate(u, x, D->G_params.data());
            F.evaluate(u2.data(), x, D->F_params.data());
            cplx ts = D->gamma * t, tt = cplx(1.0) - t;
            for (int i = 0; i < m; ++i) u[i] = ts * u[i] + tt * u2[i];
        } else {
            F.evaluate(u, x, params_at(t));
        }
    }
    void evaluate_dd(cplx* u, const cdd* x, cplx t) {
        if (D->kind == H_STRAIGHT_LINE) {  // :81-94, combine in DD then round
            G.evaluate_dd_native(vdd.data(), x, D->G_params.data());
            F.evaluate_dd_native(udd.data(), x, D->F_params.data());
            cdd ts(D->gamma * t), tt(cplx(1.0) - t);
            for (int i = 0; i < m; ++i) u[i] = to_cplx(ts * vdd[i] + tt * udd[i]);
        } else {
            F.evaluate_dd(u, x, params_at(t));
        }
    }
    void evaluate_and_jacobian(cplx* u, cplx* U, const cplx* x, cplx t) {
        if (D->kind == H_STRAIGHT_LINE) {  // :106-124
            G.evaluate_and_jacobian(u, U, x, D->G_params.data());
            F.evaluate_and_jacobian(u2.data(), U2.data(), x, D->F_params.data());
            cplx ts = D->gamma * t, tt = cplx(1.0) - t;
            for (int i = 0; i < m; ++i) u[i] = ts * u[i] + tt * u2[i];
            for (int j = 0; j < m * n; ++j) U[j] = ts * U[j] + tt * U2[j];
        } else {
            F.evaluate_and_jacobian(u, U, x, params_at(t));
        }
    }
    // taylor!(u, Val(K), H, tx, t): K-th Taylor coefficient of lambda -> H(x(lambda), t + lambda)
    // tx: K rows (x^0 .. x^{K-1}) of n.
    void taylor(int K, cplx* u, const cplx* tx, cplx t) {
        switch (D->kind) {
            case H_STRAIGHT_LINE: {
                if (K == 1) {  // :130-137
                    G.evaluate(u, tx, D->G_params.data());
                    F.evaluate(u2.data(), tx, D->F_params.data());
                    for (int i = 0; i < m; ++i) u[i] = D->gamma * u[i] - u2[i];
                } else {       // :138-154
                    G.taylor(K, tay1.data(), tx, K, D->G_params.data(), 1);
                    F.taylor(K, tay2.data(), tx, K, D->F_params.data(), 1);
                    for (int i = 0; i < m; ++i) {
                        cplx start = D->gamma * (tay1[(size_t)(K - 1) * m + i] + t * tay1[(size_t)K * m + i]);
                        cplx target = (cplx(1.0) - t) * tay2[(size_t)K * m + i] - tay2[(size_t)(K - 1) * m + i];
                        u[i] = start + target;
                    }
                }
            } break;
            case H_PARAMETER: {  // parameter_homotopy.jl:98-101 (all orders through the Taylor tape)
                linear_params(t);
                for (int i = 0; i < P; ++i) { tpbuf[i] = pt[i]; tpbuf[P + i] = dpt[i]; }
                F.taylor(K, tay1.data(), tx, K, tpbuf.data(), 2);
                for (int i = 0; i < m; ++i) u[i] = tay1[(size_t)K * m + i];
            } break;
            case H_COEFFICIENT: {
                if (K == 1) {  // coefficient_homotopy.jl:117-122: F(x; c - d)
                    for (int i = 0; i < P; ++i) dpt[i] = p[i] - q[i];
                    F.evaluate(u, tx, dpt.data());
                } else {       // :139-141
                    linear_params(t);
                    for (int i = 0; i < P; ++i) { tpbuf[i] = pt[i]; tpbuf[P + i] = dpt[i]; }
                    F.taylor(K, tay1.data(), tx, K, tpbuf.data(), 2);
                    for (int i = 0; i < m; ++i) u[i] = tay1[(size_t)K * m + i];
                }
            } break;
            case H_TORIC: {
                if (K == 1) {  // toric_homotopy.jl:216-218
                    toric_dt_coeffs(t.re);
                    F.evaluate(u, tx, dt_coeffs.data());
                } else {       // :267-278 (tc2 = 3 rows, tc3 = 4 rows, full = 5 rows)
                    toric_taylor_coeffs(t.re);
                    F.taylor(K, tay1.data(), tx, K, tc.data(), K + 1);
                    for (int i = 0; i < m; ++i) u[i] = tay1[(size_t)K * m + i];
                }
            } break;
        }
    }
};

}  // namespace orc
