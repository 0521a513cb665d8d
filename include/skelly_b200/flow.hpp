// skelly_b200/flow.hpp -- C++ host-side mirror of the three container flows and of the hydrodynamic part of
// System::apply_matvec, on top of the C ABI in skelly_b200_flow.h.  Header-only, works with Eigen::MatrixXd or the
// minimal skelly_b200::Matrix (kernels.hpp).  Names and argument meaning follow the reference:
//
//   FiberContainerFiniteDifference::flow(r_trg, fib_forces, eta, subtract_self)   fcfd.cpp:172-214
//   Periphery::flow(r_trg, density, eta)                                           periphery.cpp:55-79
//   BodyContainer::flow(r_trg, body densities, forces_torques, eta)                body_container.cpp:269-477
//   System::apply_matvec  (v_all over [fibers | shell | bodies])                   system.cpp:284-316
//   System::velocity_at_targets                                                    system.cpp:330-384
//   FiberContainerFiniteDifference::apply_fiber_force / ::matvec                   fcfd.cpp:272-287, 216-232
//
// Every failure of the C ABI becomes std::runtime_error (caught in skelly_sim.cpp:57-64).
#pragma once
#include "../skelly_b200_dense.h"
#include "../skelly_b200_flow.h"
#include "kernels.hpp"

#include <algorithm>
#include <memory>
#include <vector>

namespace skelly_b200 {

template <class M> class FlowEngineT {
  public:
    explicit FlowEngineT(int device = 0) {
        st_ = std::make_shared<State>();
        check(skb_flow_create(device, &st_->fl), "skb_flow_create");
    }

    // ---- geometry: call after positions change (System::step), not per GMRES iteration ----
    /// r_fib: 3 x N_f node positions in container order; n_nodes / lengths per fiber
    void set_fibers(const M &r_fib, const std::vector<int> &n_nodes, const std::vector<double> &lengths) {
        if (n_nodes.size() != lengths.size())
            throw std::runtime_error("skelly_b200: n_nodes / lengths size mismatch");
        check(skb_flow_set_fibers(st_->fl, r_fib.data(), n_nodes.data(), lengths.data(), (int)n_nodes.size()),
              "skb_flow_set_fibers");
        st_->n_fib = r_fib.cols();
        st_->n_nodes = n_nodes;
    }
    void set_periphery(const M &node_pos, const M &node_normal) {
        check(skb_flow_set_periphery(st_->fl, node_pos.data(), node_normal.data(), node_pos.cols()),
              "skb_flow_set_periphery");
        st_->n_shell = node_pos.cols();
    }
    void set_bodies(const M &node_pos, const M &node_normal, const M &centers) {
        check(skb_flow_set_bodies(st_->fl, node_pos.data(), node_normal.data(), node_pos.cols(), centers.data(),
                                  (int)centers.cols()),
              "skb_flow_set_bodies");
        st_->n_body = node_pos.cols();
    }

    // ---- the reference's flow() functions ----
    M fiber_flow(const M &r_trg, const M &fib_forces, double eta, bool subtract_self = true) const {
        M vel = M::Zero(3, r_trg.cols());
        check(skb_flow_fibers(st_->fl, r_trg.data(), r_trg.cols(), fib_forces.data(), eta, subtract_self ? 1 : 0,
                              vel.data()),
              "skb_flow_fibers");
        return vel;
    }
    M periphery_flow(const M &r_trg, const M &density, double eta) const {
        M vel = M::Zero(3, r_trg.cols());
        check(skb_flow_periphery(st_->fl, r_trg.data(), r_trg.cols(), density.data(), eta, vel.data()),
              "skb_flow_periphery");
        return vel;
    }
    /// forces_torques: 6 x n_bodies (body_container.cpp:128-135)
    M body_flow(const M &r_trg, const M &densities, const M &forces_torques, double eta) const {
        M vel = M::Zero(3, r_trg.cols());
        check(skb_flow_bodies(st_->fl, r_trg.data(), r_trg.cols(), densities.data(), forces_torques.data(), eta,
                              vel.data()),
              "skb_flow_bodies");
        return vel;
    }
    /// hydrodynamic part of System::apply_matvec: 3 x (N_f + N_s + N_b), targets [fibers | shell | bodies]
    M matvec_flow(const M &fib_forces, const M &shell_density, const M &body_densities, const M &forces_torques,
                  double eta) const {
        M v = M::Zero(3, st_->n_fib + st_->n_shell + st_->n_body);
        check(skb_flow_matvec(st_->fl, fib_forces.data(), shell_density.data(), body_densities.data(),
                              forces_torques.data(), eta, v.data()),
              "skb_flow_matvec");
        return v;
    }
    M velocity_at_targets(const M &r_trg, const M &fib_forces, const M &shell_density, const M &body_densities,
                          const M &forces_torques, double eta) const {
        M vel = M::Zero(3, r_trg.cols());
        check(skb_flow_velocity_at_targets(st_->fl, r_trg.data(), r_trg.cols(), fib_forces.data(),
                                           shell_density.data(), body_densities.data(), forces_torques.data(), eta,
                                           vel.data()),
              "skb_flow_velocity_at_targets");
        return vel;
    }

    // ---- per-fiber dense operators on the device ----
    /// FiberFiniteDifference::matrices_.at(n_nodes): D_1_0 (n x n), P_downsample_bc ((4n-14) x 4n)  ffd.cpp:537-555
    void set_fiber_class(int n_nodes, const M &D_1_0, const M &P_downsample_bc) {
        check(skb_flow_set_fiber_class(st_->fl, n_nodes, D_1_0.data(), P_downsample_bc.data()),
              "skb_flow_set_fiber_class");
    }
    /// Once per timestep after set_fibers: A[f] / force_operator[f] point at fib.A_.data() / fib.force_operator_.data()
    /// (column-major 4n x 4n / 3n x 4n); xs = 3 x N_f tangents; length_prev / plus_bc_velocity per fiber.
    void set_fiber_operators(const std::vector<const double *> &A, const std::vector<const double *> &force_operator,
                             const M &xs, const std::vector<double> &length_prev,
                             const std::vector<int> &plus_bc_velocity) {
        const size_t nf = st_->n_nodes.size();
        if (A.size() != nf || force_operator.size() != nf || length_prev.size() != nf || plus_bc_velocity.size() != nf)
            throw std::runtime_error("skelly_b200: set_fiber_operators needs one entry per fiber of set_fibers");
        size_t na = 0, nfo = 0;
        for (int n : st_->n_nodes) {
            na += (size_t)16 * n * n;
            nfo += (size_t)12 * n * n;
        }
        std::vector<double> a(na), fo(nfo);
        size_t oa = 0, of = 0;
        for (size_t f = 0; f < nf; ++f) {
            const size_t n = (size_t)st_->n_nodes[f];
            std::copy(A[f], A[f] + 16 * n * n, a.begin() + oa);
            std::copy(force_operator[f], force_operator[f] + 12 * n * n, fo.begin() + of);
            oa += 16 * n * n;
            of += 12 * n * n;
        }
        check(skb_flow_set_fiber_operators(st_->fl, a.data(), fo.data(), xs.data(), length_prev.data(),
                                           plus_bc_velocity.data()),
              "skb_flow_set_fiber_operators");
    }
    /// A_inv[f] points at the column-major 4n x 4n inverse of fiber f's A_ (e.g. MatrixXd(fib.A_LU_.inverse()).data())
    void set_fiber_preconditioner(const std::vector<const double *> &A_inv) {
        const size_t nf = st_->n_nodes.size();
        if (A_inv.size() != nf)
            throw std::runtime_error("skelly_b200: set_fiber_preconditioner needs one entry per fiber of set_fibers");
        size_t na = 0;
        for (int n : st_->n_nodes)
            na += (size_t)16 * n * n;
        std::vector<double> a(na);
        size_t oa = 0;
        for (size_t f = 0; f < nf; ++f) {
            const size_t n = (size_t)st_->n_nodes[f];
            std::copy(A_inv[f], A_inv[f] + 16 * n * n, a.begin() + oa);
            oa += 16 * n * n;
        }
        check(skb_flow_set_fiber_preconditioner(st_->fl, a.data()), "skb_flow_set_fiber_preconditioner");
    }
    /// FiberContainerFiniteDifference::apply_preconditioner (fcfd.cpp:331-339): y = A_^-1 x per fiber
    M apply_fiber_preconditioner(const M &x_fibers) const {
        M y = M::Zero(4 * st_->n_fib, 1);
        check(skb_flow_apply_fiber_preconditioner(st_->fl, x_fibers.data(), y.data()),
              "skb_flow_apply_fiber_preconditioner");
        return y;
    }
    /// FiberContainerFiniteDifference::apply_fiber_force (fcfd.cpp:272-287): x_fibers (4 N_f) -> 3 x N_f
    M apply_fiber_force(const M &x_fibers) const {
        M fw = M::Zero(3, st_->n_fib);
        check(skb_flow_apply_fiber_force(st_->fl, x_fibers.data(), fw.data()), "skb_flow_apply_fiber_force");
        return fw;
    }
    /// FiberContainerFiniteDifference::matvec (fcfd.cpp:216-232); v_fib_boundary 7 x n_fibers (size 0: none)
    M fiber_matvec(const M &x_fibers, const M &v_fibers, const M &v_fib_boundary) const {
        M res = M::Zero(4 * st_->n_fib, 1);
        check(skb_flow_fiber_matvec(st_->fl, x_fibers.data(), v_fibers.data(),
                                    v_fib_boundary.size() > 0 ? v_fib_boundary.data() : nullptr, res.data()),
              "skb_flow_fiber_matvec");
        return res;
    }
    /// System::apply_matvec (system.cpp:298-318) with fw and v_fibers kept on the device: returns res_fibers (4 N_f)
    /// and fills v_shell (3 x N_s) / v_bodies (3 x N_b) for shell.matvec / bc.matvec.
    M apply_matvec(const M &x_fibers, const M &shell_density, const M &body_densities, const M &forces_torques,
                   const M &fiber_link_conditions, double eta, M &v_shell, M &v_bodies) const {
        M res = M::Zero(4 * st_->n_fib, 1);
        v_shell = M::Zero(3, st_->n_shell);
        v_bodies = M::Zero(3, st_->n_body);
        check(skb_flow_apply_matvec(st_->fl, x_fibers.data(), shell_density.data(), body_densities.data(),
                                    forces_torques.data(),
                                    fiber_link_conditions.size() > 0 ? fiber_link_conditions.data() : nullptr, eta,
                                    res.data(), v_shell.data(), v_bodies.data()),
              "skb_flow_apply_matvec");
        return res;
    }
    /// The same with shell.matvec on the device: `dn` holds stresslet_plus_complementary_ (skelly_b200_dense.h);
    /// res_shell (3 N_s) = stresslet_plus_complementary_ * x_shell + v_shell  (system.cpp:319, periphery.cpp:38-47)
    M apply_matvec(skb_dense *dn, const M &x_fibers, const M &x_shell, const M &body_densities,
                   const M &forces_torques, const M &fiber_link_conditions, double eta, M &res_shell,
                   M &v_bodies) const {
        M res = M::Zero(4 * st_->n_fib, 1);
        res_shell = M::Zero(3 * st_->n_shell, 1);
        v_bodies = M::Zero(3, st_->n_body);
        check(skb_flow_apply_matvec_dense(st_->fl, dn, x_fibers.data(), x_shell.data(), body_densities.data(),
                                          forces_torques.data(),
                                          fiber_link_conditions.size() > 0 ? fiber_link_conditions.data() : nullptr,
                                          eta, res.data(), res_shell.data(), v_bodies.data()),
              "skb_flow_apply_matvec_dense");
        return res;
    }
    // ---- switches of the matvec (all optional; defaults are the fast paths with the reference's semantics) ----
    /// fibers <-> periphery pairs in one geometry pass: -1 auto (default), 0 never, 1 whenever applicable
    void set_cross(int mode) { check(skb_flow_set_cross(st_->fl, mode), "skb_flow_set_cross"); }
    /// the pair kernels skip intra-fiber pairs instead of the reference's compute-then-subtract (fcfd.cpp:203-210)
    void set_self_exclusion(bool fused) {
        check(skb_flow_set_self_exclusion(st_->fl, fused ? 1 : 0), "skb_flow_set_self_exclusion");
    }
    /// Periphery::matvec's dense operator beside the pair kernels on a side stream (default) or after them
    void set_overlap(bool on) { check(skb_flow_set_overlap(st_->fl, on ? 1 : 0), "skb_flow_set_overlap"); }
    skb_flow *handle() const { return st_->fl; }

  private:
    struct State { // copies of the engine share the device state and the geometry sizes, like the evaluators
        skb_flow *fl = nullptr;
        long n_fib = 0, n_shell = 0, n_body = 0;
        std::vector<int> n_nodes;
        ~State() { skb_flow_destroy(fl); }
    };
    std::shared_ptr<State> st_;
};

using FlowEngine = FlowEngineT<Matrix>;

} // namespace skelly_b200
