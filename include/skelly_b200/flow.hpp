// skelly_b200/flow.hpp -- C++ host-side mirror of the three container flows and of the hydrodynamic part of
// System::apply_matvec, on top of the C ABI in skelly_b200_flow.h.  Header-only, works with Eigen::MatrixXd or the
// minimal skelly_b200::Matrix (kernels.hpp).  Names and argument meaning follow the reference:
//
//   FiberContainerFiniteDifference::flow(r_trg, fib_forces, eta, subtract_self)   fcfd.cpp:172-214
//   Periphery::flow(r_trg, density, eta)                                           periphery.cpp:55-79
//   BodyContainer::flow(r_trg, body densities, forces_torques, eta)                body_container.cpp:269-477
//   System::apply_matvec  (v_all over [fibers | shell | bodies])                   system.cpp:284-316
//   System::velocity_at_targets                                                    system.cpp:330-384
//
// Every failure of the C ABI becomes std::runtime_error (caught in skelly_sim.cpp:57-64).
#pragma once
#include "../skelly_b200_flow.h"
#include "kernels.hpp"

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
    skb_flow *handle() const { return st_->fl; }

  private:
    struct State { // copies of the engine share the device state and the geometry sizes, like the evaluators
        skb_flow *fl = nullptr;
        long n_fib = 0, n_shell = 0, n_body = 0;
        ~State() { skb_flow_destroy(fl); }
    };
    std::shared_ptr<State> st_;
};

using FlowEngine = FlowEngineT<Matrix>;

} // namespace skelly_b200
