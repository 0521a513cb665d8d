// kernel_test.cpp -- re-creation of SkellySim's tests/core/kernel_test.cpp for the drop-in library:
// n_src = 1229, n_trg = 743, eta = 1.3, U[-1,1] inputs (kernel_test.cpp:25-37, seeded here), candidate =
// the GPU evaluators, reference = a 1-thread direct sum written as the reference states the math
// (kernels.cu:57-77 / :24-55).  Pass iff ||ref - other||_2 <= 5e-9 (kernel_test.cpp:92) AND our own gate
// max|d|/max|u| <= 1e-12.   usage: kernel_test --kernel=[stokeslet|stresslet] --driver=[gpu|gpu_cached|impl]
#include <skelly_b200/kernels.hpp>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>

// the reference-named C++ entry points exported by libskelly_b200.so (include/kernels.hpp:17-20 of SkellySim)
namespace kernels {
void stokeslet_direct_gpu_impl(const double *r_src, const double *f_src, int n_src, const double *r_trg,
                               double *u_trg, int n_trg);
void stresslet_direct_gpu_impl(const double *r_src, const double *f_src, int n_src, const double *r_trg,
                               double *u_trg, int n_trg);
} // namespace kernels

using skelly_b200::Matrix;

static Matrix random_matrix(long rows, long cols, std::mt19937_64 &gen) {
    std::uniform_real_distribution<double> d(-1.0, 1.0);
    Matrix m(rows, cols);
    for (long j = 0; j < cols; ++j)
        for (long i = 0; i < rows; ++i)
            m(i, j) = d(gen);
    return m;
}

static Matrix direct_single_thread(const Matrix &r_src, const Matrix &r_trg, const Matrix &f, bool stresslet,
                                   double eta) {
    Matrix u(3, r_trg.cols());
    const double scale = 1.0 / 8.0 / M_PI;
    for (long t = 0; t < r_trg.cols(); ++t) {
        double a[3] = {0, 0, 0};
        for (long s = 0; s < r_src.cols(); ++s) {
            const double dx = r_trg(0, t) - r_src(0, s), dy = r_trg(1, t) - r_src(1, s), dz = r_trg(2, t) - r_src(2, s);
            const double r2 = dx * dx + dy * dy + dz * dz;
            const double rinv = r2 == 0.0 ? 0.0 : 1.0 / std::sqrt(r2);
            const double rinv2 = rinv * rinv;
            if (!stresslet) {
                const double ip = (f(0, s) * dx + f(1, s) * dy + f(2, s) * dz) * rinv2;
                a[0] += rinv * (f(0, s) + dx * ip);
                a[1] += rinv * (f(1, s) + dy * ip);
                a[2] += rinv * (f(2, s) + dz * ip);
            } else {
                double c = f(0, s) * dx * dx + f(4, s) * dy * dy + f(8, s) * dz * dz;
                c += (f(1, s) + f(3, s)) * dx * dy;
                c += (f(2, s) + f(6, s)) * dx * dz;
                c += (f(5, s) + f(7, s)) * dy * dz;
                c *= -3.0 * rinv * rinv2 * rinv2;
                a[0] += dx * c;
                a[1] += dy * c;
                a[2] += dz * c;
            }
        }
        for (int k = 0; k < 3; ++k)
            u(k, t) = a[k] * scale / eta;
    }
    return u;
}

int main(int argc, char **argv) {
    std::string kernel = "stokeslet", driver = "gpu";
    for (int i = 1; i < argc; ++i) {
        if (!std::strncmp(argv[i], "--kernel=", 9)) kernel = argv[i] + 9;
        if (!std::strncmp(argv[i], "--driver=", 9)) driver = argv[i] + 9;
    }
    constexpr int n_src = 1229, n_trg = 743;
    constexpr double eta = 1.3;
    std::mt19937_64 gen(1);
    Matrix r_src = random_matrix(3, n_src, gen), r_trg = random_matrix(3, n_trg, gen), nullmat;
    const bool stresslet = kernel == "stresslet";
    if (!stresslet && kernel != "stokeslet") {
        fprintf(stderr, "Invalid kernel supplied \"%s\"\n", kernel.c_str());
        return 2;
    }
    Matrix f_src = random_matrix(stresslet ? 9 : 3, n_src, gen);
    Matrix ref = direct_single_thread(r_src, r_trg, f_src, stresslet, eta);
    Matrix other;
    try {
        if (driver == "gpu") {
            other = stresslet ? skelly_b200::stresslet_direct_gpu(nullmat, r_src, r_trg, nullmat, f_src, eta)
                              : skelly_b200::stokeslet_direct_gpu(r_src, nullmat, r_trg, f_src, nullmat, eta);
        } else if (driver == "gpu_cached") {
            skelly_b200::GPUEvaluator ev(1);
            skelly_b200::GPUEvaluator ev_copy = ev; // containers copy evaluators by value
            for (int rep = 0; rep < 2; ++rep)  // second call hits the position cache
                other = stresslet ? ev_copy(nullmat, r_src, r_trg, nullmat, f_src, eta)
                                  : ev_copy(r_src, nullmat, r_trg, f_src, nullmat, eta);
        } else if (driver == "impl") {
            other = Matrix::Zero(3, n_trg);
            if (stresslet)
                kernels::stresslet_direct_gpu_impl(r_src.data(), f_src.data(), n_src, r_trg.data(), other.data(), n_trg);
            else
                kernels::stokeslet_direct_gpu_impl(r_src.data(), f_src.data(), n_src, r_trg.data(), other.data(), n_trg);
            other /= eta;
        } else {
            fprintf(stderr, "Invalid driver supplied \"%s\"\n", driver.c_str());
            return 2;
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 3;
    }
    double err2 = 0, dmax = 0, umax = 0;
    for (long i = 0; i < ref.size(); ++i) {
        const double d = ref.data()[i] - other.data()[i];
        err2 += d * d;
        dmax = std::fmax(dmax, std::fabs(d));
        umax = std::fmax(umax, std::fabs(ref.data()[i]));
    }
    const double err = std::sqrt(err2);
    printf("%s %s: l2 err %.3e (gate 5e-9), max rel err %.3e (gate 1e-12)\n", kernel.c_str(), driver.c_str(), err,
           dmax / umax);
    return (err > 5e-9) || (dmax / umax > 1e-12);
}
