// flow_test.cpp -- C++ host test of skelly_b200/flow.hpp: FiberContainer::flow semantics (trapezoid weights,
// all-pairs Stokeslet / eta, self-term subtraction) and the apply_matvec target masks, against straightforward host
// loops written from the reference's statements (fcfd.cpp:172-214, kernels.cpp:146-195, periphery.cpp:68-74).
#include <skelly_b200/flow.hpp>

#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

using skelly_b200::Matrix;

static void oseen_pair(const double *t, const double *s, const double *f, double factor, double reg, double eps,
                       bool regularise, double *u) {
    const double dx = t[0] - s[0], dy = t[1] - s[1], dz = t[2] - s[2];
    const double r2 = dx * dx + dy * dy + dz * dz;
    if (r2 == 0.0)
        return;
    double r = std::sqrt(r2);
    if (regularise && r <= eps)
        r = std::sqrt(r * r + reg * reg);
    const double fr = factor / r, gr = factor / (r * r * r);
    const double dot = gr * (dx * f[0] + dy * f[1] + dz * f[2]);
    u[0] += fr * f[0] + dx * dot;
    u[1] += fr * f[1] + dy * dot;
    u[2] += fr * f[2] + dz * dot;
}

int main() {
    const int n_fibers = 12, n = 16, n_shell = 150;
    const double eta = 0.9, L = 1.0;
    std::mt19937_64 gen(5);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    Matrix r_fib(3, n_fibers * n), forces(3, n_fibers * n), r_shell(3, n_shell), n_shell_m(3, n_shell), rho(3, n_shell);
    for (int f = 0; f < n_fibers; ++f) {
        double x0[3] = {3 * U(gen), 3 * U(gen), 3 * U(gen)}, d[3] = {U(gen), U(gen), U(gen)};
        const double nd = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k) {
                r_fib(k, f * n + i) = x0[k] + L * i / (n - 1.0) * d[k] / nd;
                forces(k, f * n + i) = U(gen);
            }
    }
    for (int i = 0; i < n_shell; ++i) {
        double d[3] = {U(gen), U(gen), U(gen)};
        const double nd = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        for (int k = 0; k < 3; ++k) {
            r_shell(k, i) = 6.0 * d[k] / nd;
            n_shell_m(k, i) = -d[k] / nd;
            rho(k, i) = U(gen);
        }
    }
    Matrix empty;
    try {
        skelly_b200::FlowEngine eng(0), eng_copy = eng;
        eng.set_fibers(r_fib, std::vector<int>(n_fibers, n), std::vector<double>(n_fibers, L));
        eng.set_periphery(r_shell, n_shell_m);
        eng.set_bodies(empty, empty, empty);
        Matrix v = eng_copy.matvec_flow(forces, rho, empty, empty, eta);

        // host reference
        const int nf = n_fibers * n, n_all = nf + n_shell;
        std::vector<double> ref(3 * (size_t)n_all, 0.0), w(nf);
        for (int f = 0; f < n_fibers; ++f)
            for (int i = 0; i < n; ++i)
                w[f * n + i] = 0.5 * L * ((i == 0 || i == n - 1) ? 1.0 : 2.0) / (n - 1.0);
        const double factor = 1.0 / (8.0 * M_PI * eta);
        for (int t = 0; t < n_all; ++t) {
            const double *xt = t < nf ? &r_fib.data()[3 * t] : &r_shell.data()[3 * (t - nf)];
            for (int s = 0; s < nf; ++s) {
                if (t < nf && t / n == s / n)
                    continue; // own fiber: all-pairs minus the (unregularised here: spacing >> eps) self block
                double wf[3] = {w[s] * forces(0, s), w[s] * forces(1, s), w[s] * forces(2, s)};
                oseen_pair(xt, &r_fib.data()[3 * s], wf, factor, 5e-3, 1e-5, false, &ref[3 * t]);
            }
            if (t < nf) // shell double layer acts on fiber targets only (system.cpp:301-315)
                for (int s = 0; s < n_shell; ++s) {
                    const double dx = xt[0] - r_shell(0, s), dy = xt[1] - r_shell(1, s), dz = xt[2] - r_shell(2, s);
                    const double r2 = dx * dx + dy * dy + dz * dz, rinv = 1.0 / std::sqrt(r2);
                    const double dn = dx * n_shell_m(0, s) + dy * n_shell_m(1, s) + dz * n_shell_m(2, s);
                    const double dr = dx * rho(0, s) + dy * rho(1, s) + dz * rho(2, s);
                    const double c = -3.0 / (4.0 * M_PI) * dn * dr * rinv * rinv * rinv * rinv * rinv;
                    ref[3 * t] += c * dx;
                    ref[3 * t + 1] += c * dy;
                    ref[3 * t + 2] += c * dz;
                }
        }
        double dmax = 0, umax = 0;
        for (size_t i = 0; i < ref.size(); ++i) {
            dmax = std::fmax(dmax, std::fabs(ref[i] - v.data()[i]));
            umax = std::fmax(umax, std::fabs(ref[i]));
        }
        printf("flow_test: max rel err %.3e (gate 1e-11)\n", dmax / umax);
        if (dmax / umax > 1e-11)
            return 1;

        // ---- per-fiber dense operators through the wrapper (fcfd.cpp:272-287, ffd.cpp:276-312) ----
        const int n4 = 4 * n, n3 = 3 * n, bc = n4 - 14;
        Matrix D(n, n), P(bc, n4), xs(3, nf), x(4 * nf, 1), link(7, n_fibers);
        for (long i = 0; i < D.size(); ++i) D.data()[i] = U(gen);
        for (long i = 0; i < P.size(); ++i) P.data()[i] = U(gen) / n4;
        for (long i = 0; i < xs.size(); ++i) xs.data()[i] = U(gen);
        for (long i = 0; i < x.size(); ++i) x.data()[i] = U(gen);
        for (long i = 0; i < link.size(); ++i) link.data()[i] = U(gen);
        std::vector<Matrix> A(n_fibers, Matrix(n4, n4)), F(n_fibers, Matrix(n3, n4));
        std::vector<const double *> Ap, Fp;
        std::vector<double> lprev(n_fibers);
        std::vector<int> plus(n_fibers);
        for (int f = 0; f < n_fibers; ++f) {
            for (long i = 0; i < A[f].size(); ++i) A[f].data()[i] = U(gen) / n4;
            for (long i = 0; i < F[f].size(); ++i) F[f].data()[i] = U(gen) / n4;
            Ap.push_back(A[f].data());
            Fp.push_back(F[f].data());
            lprev[f] = 1.0 + 0.05 * f;
            plus[f] = f % 2;
        }
        eng.set_fiber_class(n, D, P);
        eng.set_fiber_operators(Ap, Fp, xs, lprev, plus);
        eng.set_cross(-1); // the switches of the matvec at their defaults (compile + link + argument checks)
        eng.set_self_exclusion(false);
        eng.set_overlap(true);
        Matrix fw = eng.apply_fiber_force(x), v_s, v_b;
        Matrix res = eng_copy.apply_matvec(x, rho, empty, empty, link, eta, v_s, v_b);
        Matrix v_all = eng.matvec_flow(fw, rho, empty, empty, eta);
        double e_fw = 0, m_fw = 0, e_res = 0, m_res = 0;
        for (int f = 0; f < n_fibers; ++f) {
            const double *xf = x.data() + (size_t)4 * f * n;
            for (int r = 0; r < n3; ++r) { // fw(k, off + i) = (force_operator_ * x)(k n + i)
                double s = 0;
                for (int j = 0; j < n4; ++j)
                    s += F[f](r, j) * xf[j];
                const double got = fw(r / n, f * n + r % n);
                e_fw = std::fmax(e_fw, std::fabs(s - got));
                m_fw = std::fmax(m_fw, std::fabs(s));
            }
            std::vector<double> vT(n4, 0.0);
            for (int i = 0; i < n; ++i)
                for (int k = 0; k < 3; ++k)
                    vT[k * n + i] = v_all(k, f * n + i);
            for (int j = 0; j < n; ++j) // (xsDs vx + ysDs vy + zsDs vz)(j) = sum_i D_1(i,j) xs_i . v_i
                for (int i = 0; i < n; ++i) {
                    double dot = 0;
                    for (int k = 0; k < 3; ++k)
                        dot += xs(k, f * n + i) * v_all(k, f * n + i);
                    vT[n3 + j] += D(i, j) * (2.0 / lprev[f]) * dot;
                }
            for (int r = 0; r < n4; ++r) {
                double s = 0;
                for (int j = 0; j < n4; ++j)
                    s += A[f](r, j) * xf[j];
                if (r < bc)
                    for (int j = 0; j < n4; ++j)
                        s -= P(r, j) * vT[j];
                if (r >= bc && r < bc + 7)
                    s += link(r - bc, f);
                if (r == bc + 3 || (r == bc + 10 && plus[f])) {
                    const int node = f * n + (r == bc + 3 ? 0 : n - 1);
                    for (int k = 0; k < 3; ++k)
                        s += v_all(k, node) * xs(k, node);
                }
                e_res = std::fmax(e_res, std::fabs(s - res.data()[(size_t)4 * f * n + r]));
                m_res = std::fmax(m_res, std::fabs(s));
            }
        }
        double e_vs = 0;
        for (long i = 0; i < v_s.size(); ++i)
            e_vs = std::fmax(e_vs, std::fabs(v_s.data()[i] - v_all.data()[3 * nf + i]));
        printf("flow_test: fiber operators  fw %.3e  res %.3e  v_shell %.3e (gate 1e-12, exact)\n", e_fw / m_fw,
               e_res / m_res, e_vs);
        return (e_fw / m_fw > 1e-12 || e_res / m_res > 1e-12 || e_vs != 0.0) ? 2 : 0;
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 3;
    }
}
