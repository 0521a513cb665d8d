// skelly_b200/kernels.hpp -- C++ host-side mirror of SkellySim's pair-evaluator interface on top of the C ABI
// (skelly_b200.h).  Header-only; works with Eigen (`Eigen::MatrixXd`, `Eigen::Ref<const MatrixXd>`) or with the
// minimal column-major `skelly_b200::Matrix` below when Eigen is not available.
//
//   reference (include/kernels.hpp)                         here (namespace skelly_b200)
//   ------------------------------------------------------  --------------------------------------------
//   using Evaluator = std::function<MatrixXd(r_sl, r_dl,    template <class M> using EvaluatorT = std::function<...>
//        r_trg, f_sl, f_dl, eta)>                  :14-15
//   stokeslet_direct_gpu / stresslet_direct_gpu    :34-38   stokeslet_direct_gpu<M> / stresslet_direct_gpu<M>
//        (src/core/kernels.cpp:354-366)                        (stateless: upload everything, like the reference)
//   class GPUEvaluator (declared only)             :136-191 class GPUEvaluatorT<M>: positions cached on the
//   class FMM<stkfmm_type>  (position-cache logic) :56-134     device, re-uploaded only when they change
//
// Error behaviour: every failure of the C ABI becomes std::runtime_error, which SkellySim's main() catches
// and turns into MPI_Abort (src/skelly_sim.cpp:57-64).
#pragma once
#include "../skelly_b200.h"

#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace skelly_b200 {

/// Minimal column-major dense matrix with the slice of the Eigen::MatrixXd interface the evaluators use
/// (`rows()`, `cols()`, `size()`, `data()`, `(i,j)`); a default-constructed Matrix is the "empty dummy"
/// the reference passes for the unused source class (fiber_container_finite_difference.cpp:196).
class Matrix {
  public:
    Matrix() = default;
    Matrix(long rows, long cols) : rows_(rows), cols_(cols), v_((size_t)(rows * cols), 0.0) {}
    static Matrix Zero(long rows, long cols) { return Matrix(rows, cols); }
    long rows() const { return rows_; }
    long cols() const { return cols_; }
    long size() const { return rows_ * cols_; }
    double *data() { return v_.data(); }
    const double *data() const { return v_.data(); }
    double &operator()(long i, long j) { return v_[(size_t)(j * rows_ + i)]; }
    double operator()(long i, long j) const { return v_[(size_t)(j * rows_ + i)]; }
    Matrix &operator/=(double s) {
        for (double &x : v_)
            x /= s;
        return *this;
    }

  private:
    long rows_ = 0, cols_ = 0;
    std::vector<double> v_;
};

inline void check(int rc, const char *what) {
    if (rc != SKB_OK)
        throw std::runtime_error(std::string("skelly_b200: ") + what + ": " + skb_last_error_string());
}

template <class M>
using EvaluatorT = std::function<M(const M &r_sl, const M &r_dl, const M &r_trg, const M &f_sl, const M &f_dl, double eta)>;

/// kernels::stokeslet_direct_gpu (src/core/kernels.cpp:361-366): zero-initialised 3 x n_trg result, / eta.
template <class M> M stokeslet_direct_gpu(const M &r_sl, const M &, const M &r_trg, const M &f_sl, const M &, double eta) {
    M u = M::Zero(3, r_trg.cols());
    check(skb_stokeslet_direct(r_sl.data(), f_sl.data(), (int)r_sl.cols(), r_trg.data(), u.data(), (int)r_trg.cols()),
          "stokeslet_direct_gpu");
    u /= eta;
    return u;
}

/// kernels::stresslet_direct_gpu (src/core/kernels.cpp:354-359)
template <class M> M stresslet_direct_gpu(const M &, const M &r_dl, const M &r_trg, const M &, const M &f_dl, double eta) {
    M u = M::Zero(3, r_trg.cols());
    check(skb_stresslet_direct(r_dl.data(), f_dl.data(), (int)r_dl.cols(), r_trg.data(), u.data(), (int)r_trg.cols()),
          "stresslet_direct_gpu");
    u /= eta;
    return u;
}

/// What kernels::GPUEvaluator (include/kernels.hpp:136-191) was meant to be: an Evaluator-shaped functor that keeps
/// source/target positions on the device(s) and re-sends them only when they change -- the same full compare the
/// FMM functor uses to decide on a tree rebuild (kernels.hpp:81-83).  Copies share the device context
/// (containers copy their evaluators by value, body_container.cpp:577-600), like FMM::fmmPtr_.
template <class M> class GPUEvaluatorT {
  public:
    explicit GPUEvaluatorT(int n_gpus = 1) : st_(std::make_shared<State>()) {
        check(skb_ctx_create(n_gpus, &st_->ctx), "skb_ctx_create");
    }
    /// force the next call to re-upload positions (FMM::force_setup_tree, kernels.hpp:66)
    void force_device_sync() { st_->force = true; }

    M operator()(const M &r_sl, const M &r_dl, const M &r_trg, const M &f_sl, const M &f_dl, double eta) {
        State &s = *st_;
        sync(s.sl, r_sl, [&](const double *p, long n) { check(skb_set_sources(s.ctx, SKB_STOKESLET, p, n), "set_sources"); });
        sync(s.dl, r_dl, [&](const double *p, long n) { check(skb_set_sources(s.ctx, SKB_STRESSLET, p, n), "set_sources"); });
        sync(s.trg, r_trg, [&](const double *p, long n) { check(skb_set_targets(s.ctx, p, n), "set_targets"); });
        s.force = false;
        if (f_sl.cols() != r_sl.cols() || f_dl.cols() != r_dl.cols())
            throw std::runtime_error("skelly_b200: strength / position column mismatch");
        M u = M::Zero(3, r_trg.cols());
        check(skb_eval_fused(s.ctx, r_sl.cols() ? f_sl.data() : nullptr, r_dl.cols() ? f_dl.data() : nullptr, u.data()),
              "skb_eval_fused");
        u /= eta; // kernels.hpp:121 / kernels.cpp:358,365
        return u;
    }

  private:
    struct Cache {
        std::vector<double> last;
        bool valid = false;
    };
    struct State {
        skb_ctx *ctx = nullptr;
        Cache sl, dl, trg;
        bool force = true;
        ~State() { skb_ctx_destroy(ctx); }
    };
    template <class F> void sync(Cache &c, const M &r, F upload) {
        const size_t n = (size_t)r.size();
        const bool same = !st_->force && c.valid && c.last.size() == n &&
                          (n == 0 || std::memcmp(c.last.data(), r.data(), n * sizeof(double)) == 0);
        if (same)
            return;
        upload(r.data(), (long)r.cols());
        c.last.assign(r.data(), r.data() + n);
        c.valid = true;
    }
    std::shared_ptr<State> st_;
};

using Evaluator = EvaluatorT<Matrix>;
using GPUEvaluator = GPUEvaluatorT<Matrix>;

} // namespace skelly_b200
