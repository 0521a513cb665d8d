/*
 * skelly_b200_flow.h -- C ABI of the device-resident flow() layer: the three container flows that
 * System::apply_matvec evaluates every GMRES iteration, and their fused sum.
 *
 *   reference (SkellySim)                                                       here
 *   --------------------------------------------------------------------------  -----------------------
 *   FiberContainerFiniteDifference::flow   fiber_container_finite_difference.cpp:172-214   skb_flow_fibers
 *   Periphery::flow                        periphery.cpp:55-79                             skb_flow_periphery
 *   BodyContainer::flow (spherical/ellipsoidal)  body_container.cpp:269-411,471-477        skb_flow_bodies
 *   System::apply_matvec, hydrodynamic part      system.cpp:284-316                        skb_flow_matvec
 *
 * A skb_flow object holds the geometry of ONE timestep on the device (node positions, normals, quadrature
 * weights, body centres); positions only change in System::step (system.cpp:486-489), so every GMRES
 * iteration ships densities/forces only.  Layouts are the reference's: Eigen column-major 3 x n == AoS xyz.
 * Velocities returned here are the reference's flow() return values, i.e. they INCLUDE the 1/eta of the
 * evaluator wrappers (kernels.cpp:358,365).
 */
#ifndef SKELLY_B200_FLOW_H
#define SKELLY_B200_FLOW_H

#include "skelly_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct skb_flow skb_flow;

SKB_API int skb_flow_create(int device, skb_flow **out);
SKB_API int skb_flow_destroy(skb_flow *fl);

/* ---- geometry, once per timestep --------------------------------------------------------------------- */
/* r_fib: 3 x N_f node positions of all fibers, concatenated in container order; n_nodes[f] nodes and
 * length[f] per fiber (trapezoid weights 0.5 L weights_0 are formed on the device,
 * fiber_container_finite_difference.cpp:185-193).  n_fibers == 0 is allowed (flow returns zeros, :178-179). */
SKB_API int skb_flow_set_fibers(skb_flow *fl, const double *r_fib, const int *n_nodes, const double *length,
                                int n_fibers);
/* periphery quadrature nodes and (inward) normals, periphery.hpp node_pos_ / node_normal_ */
SKB_API int skb_flow_set_periphery(skb_flow *fl, const double *node_pos, const double *node_normal, int64_t n_nodes);
/* all body surface nodes / normals concatenated [spherical..., ellipsoidal...] (body_container.cpp:286-290),
 * and the body centres (get_local_center_positions) */
SKB_API int skb_flow_set_bodies(skb_flow *fl, const double *node_pos, const double *node_normal, int64_t n_nodes,
                                const double *centers, int n_bodies);

/* ---- the three flows at arbitrary targets (velocity_at_targets / listener path, system.cpp:330-384) ------
 * Targets are cached on the device and re-uploaded only when they differ from the previous call's (the same
 * full compare the reference's FMM functor does, kernels.hpp:81-83). */
SKB_API int skb_flow_fibers(skb_flow *fl, const double *r_trg, int64_t n_trg, const double *fib_forces, double eta,
                            int subtract_self, double *vel);
SKB_API int skb_flow_periphery(skb_flow *fl, const double *r_trg, int64_t n_trg, const double *density, double eta,
                               double *vel);
/* densities: 3 x N_b body-node densities; forces_torques: 6 x n_bodies (net force then torque per body,
 * body_container.cpp:128-135) */
SKB_API int skb_flow_bodies(skb_flow *fl, const double *r_trg, int64_t n_trg, const double *densities,
                            const double *forces_torques, double eta, double *vel);

/* System::velocity_at_targets (system.cpp:330-384), hydrodynamic part: fiber flow WITHOUT self-term subtraction +
 * body flow + periphery flow at arbitrary targets (system.cpp:355-359; point/background sources are analytic
 * add-ons outside this library).  This is what the listener and the streamline integrator evaluate.  Any class may
 * be empty; one upload of the strengths, one download of the velocities. */
SKB_API int skb_flow_velocity_at_targets(skb_flow *fl, const double *r_trg, int64_t n_trg, const double *fib_forces,
                                         const double *shell_density, const double *body_densities,
                                         const double *body_forces_torques, double eta, double *vel);

/* Optional analytic add-ons of velocity_at_targets (system.cpp:358-359): point forces/torques
 * (PointSourceContainer::flow, point_source.cpp:16-54: regularised Oseen contraction + rotlet; the caller passes
 * the points that are alive at the current time) and the linear background flow (BackgroundSource::flow,
 * background_source.cpp:15-24: u_j = uniform_j + r[components_j] * scale_factor_j).  n_points == 0 / NULL clears. */
SKB_API int skb_flow_set_point_sources(skb_flow *fl, const double *positions, const double *forces,
                                       const double *torques, int n_points);
SKB_API int skb_flow_set_background(skb_flow *fl, const int *components, const double *scale_factor,
                                    const double *uniform);

/* ---- fused matvec flow -------------------------------------------------------------------------------
 * v_all over targets [fibers | periphery | bodies] (get_node_maps, system.cpp:234-241):
 *   v_all  = fiber flow (all targets, self term subtracted)
 *   v_fib, v_body += periphery flow   (shell sources never act on shell targets, system.cpp:301-315)
 *   v_all += body flow
 * Any class may be empty.  One H2D of the four strength arrays, one D2H of v_all; everything in between stays
 * on the device. */
SKB_API int skb_flow_matvec(skb_flow *fl, const double *fib_forces, const double *shell_density,
                            const double *body_densities, const double *body_forces_torques, double eta,
                            double *v_all);

/* The (fiber node, periphery node) pairs of a matvec are visited twice -- the fibers' Stokeslets act on the periphery
 * (system.cpp:299) and the periphery's stresslets on the fibers (system.cpp:304,313-315).  By default (mode -1: when both
 * classes are large enough and all fiber rows are in the target list) one kernel evaluates both directions from one pass
 * over the pairs (37 instead of 49 FP64 instructions per visited pair); results agree with the two separate evaluator
 * calls to rounding.  0: never (two calls, as the reference issues them); 1: whenever applicable. */
SKB_API int skb_flow_set_cross(skb_flow *fl, int mode);

/* Opt-in: the matvec's fiber self term.  fused == 0 (default): the reference's way -- all pairs,
 * then `vel -= fib.stokeslet_ * wf` per fiber (fiber_container_finite_difference.cpp:203-210) on the device.
 * fused != 0: the pair kernels skip every intra-fiber pair (per-node fiber id compared in the integer pipe), nothing is
 * subtracted: no cancellation of the O(1/ds) near-neighbour terms, and N_f * n fewer pairs.  The two agree except where
 * the reference's regularised branch acts (distinct nodes of one fiber closer than 1e-5, kernels.cpp:176-184).  Applies
 * to skb_flow_matvec / skb_flow_apply_matvec with all fiber nodes in the target list. */
SKB_API int skb_flow_set_self_exclusion(skb_flow *fl, int fused);
/* Overlap of the HBM-bound dense operator with the FP64-bound pair kernels (default on): skb_flow_apply_matvec_dense /
 * _device hand stresslet_plus_complementary * x_shell (Periphery::matvec, periphery.cpp:38-47) to the background row
 * streamer (skb_dense_apply_background_device) on a side stream as soon as x_shell is complete on the device, and add
 * v_shell when both are there.  0: one GEMV kernel at the end of the stream instead (the round-1 order).  The results
 * differ in the last bits only (another summation order inside a row). */
SKB_API int skb_flow_set_overlap(skb_flow *fl, int on);

/* Restrict the matvec to rows [begin, end) of the target list [fibers | periphery | bodies] (end < 0: all).
 * For one-rank-per-GPU hosts: every rank loads the full geometry, all-gathers the strengths, and evaluates its own
 * block of targets; v_all / d_v_window then has (end - begin) rows.  The self-term subtraction is applied to the
 * fiber rows inside the window. */
SKB_API int skb_flow_set_target_window(skb_flow *fl, int64_t begin, int64_t end);

/* The reference's own MPI decomposition as three ranges (whole fibers [fiber_begin, fiber_end) of the container,
 * fiber_container_finite_difference.cpp:102-120; periphery rows, periphery.cpp:387-400; body-node rows -- all on
 * rank 0 in the reference): the target list of this flow becomes [own fiber nodes | own shell rows | own body rows]
 * and v_all / d_v_window has that many rows in that order.  Replaces a window set earlier (and vice versa). */
SKB_API int skb_flow_set_target_ranges(skb_flow *fl, int fiber_begin, int fiber_end, int64_t shell_begin,
                                       int64_t shell_end, int64_t body_begin, int64_t body_end);

/* Device-pointer form of skb_flow_matvec: all inputs and the output already on this flow's device; asynchronous on
 * `stream` (cudaStream_t as void*, used verbatim).  d_body_forces / d_body_torques are 3 x n_bodies each. */
SKB_API int skb_flow_matvec_device(skb_flow *fl, const double *d_fib_forces, const double *d_shell_density,
                                   const double *d_body_densities, const double *d_body_forces,
                                   const double *d_body_torques, double eta, double *d_v_window, void *stream);

/* ---- multi-GPU groups: one skb_flow per GPU, exchange through peer memory (NVLink / NVSwitch) -----------------
 * Every member holds the full geometry (skb_flow_set_fibers / _periphery / _bodies with identical arguments) and owns
 * whole fibers, periphery rows and body rows (skb_flow_set_target_ranges).  Per matvec each member packs the strengths
 * of ITS fibers / periphery rows once and stores them straight into all members' windows (pack + all-gather in one
 * kernel), evaluates its block rows of the symmetric fiber-fiber interaction and its own remainder rows from local
 * memory, and finally adds up the partial velocities of its own fiber rows from all windows (reduce-scatter fused
 * with the accumulation).  No library collective is on the path; ordering is by epoch flags in the windows.
 *   same process (what the reference's one-rank rule for direct evaluators admits, system.cpp:618-623): create one
 *     flow per device, skb_flow_group_init on each, then skb_flow_group_connect every ordered pair -- or simply use
 *     skb_mflow_* below, which does all of that;
 *   one process per GPU: skb_flow_group_init, skb_flow_group_export -> exchange the 64-byte handles by any means
 *     (MPI_Allgather, torch.distributed) -> skb_flow_group_import for every peer.
 * skb_flow_group_init must follow the geometry calls (the window is laid out for their node counts; positions may
 * change later, counts may not).  In a group, the *_device entry points take the OWN slices: fiber forces / x_fibers of
 * the own fibers, periphery density of the own rows; body inputs are complete on every member.  All members must issue
 * the same sequence of matvec calls (like any collective). */
#define SKB_FLOW_IPC_HANDLE_BYTES 64
SKB_API int skb_flow_group_init(skb_flow *fl, int rank, int size);
SKB_API int skb_flow_group_export(skb_flow *fl, void *handle_64_bytes);
SKB_API int skb_flow_group_import(skb_flow *fl, int peer_rank, const void *handle_64_bytes);
SKB_API int skb_flow_group_connect(skb_flow *fl, int peer_rank, skb_flow *peer);
/* one matvec of this member alone (no flags, zero strengths) so that every buffer reaches its final size before the
 * first real matvec: device-wide synchronisation by cudaMalloc / cudaFree must not happen while a peer ON THE SAME
 * DEVICE waits on a flag.  Call after the ranges are set and all peers are connected; skb_mflow does it itself. */
SKB_API int skb_flow_group_warmup(skb_flow *fl);
/* profiling aid: from now on evaluate this member's share alone (no flags, own window only): the per-rank device time
 * of an n-way group measured on ONE GPU.  The results are the member's partial view, not the group's. */
SKB_API int skb_flow_group_set_solo(skb_flow *fl, int solo);
/* after synchronising: *missing_peer = rank of a member whose flag never arrived within the time-out (a flag wait
 * gives up after ~10 s instead of hanging the GPU), or -1 */
SKB_API int skb_flow_group_error(skb_flow *fl, int *missing_peer);

/* ---- per-fiber dense operators on the device ------------------------------------------
 * The O(N_f n^2) host work of one GMRES iteration -- FiberContainerFiniteDifference::apply_fiber_force
 * (fiber_container_finite_difference.cpp:272-287) and ::matvec (:216-232 -> FiberFiniteDifference::matvec,
 * fiber_finite_difference.cpp:276-312) -- as batched GEMVs over operators that stay resident for the timestep.
 * All matrices are COLUMN-MAJOR, exactly as Eigen stores the reference's members, so `.data()` goes in as is.
 *
 * skb_flow_set_fiber_class: the per-node-count constants FiberFiniteDifference::matrices_.at(n)
 *   (fiber_finite_difference.cpp:537-555): D_1_0 (n x n) and P_downsample_bc ((4n-14) x 4n).  Call once per
 *   distinct n (n >= 4), any time before skb_flow_set_fiber_operators.
 * skb_flow_set_fiber_operators: once per timestep, after skb_flow_set_fibers (which defines the fibers and their
 *   node counts): A = concatenated A_ (4n x 4n each), force_operator = concatenated force_operator_ (3n x 4n
 *   each), xs = 3 x N_f tangents xs_, length_prev[f] = length_prev_, plus_bc_velocity[f] = (bc_plus_.first ==
 *   Velocity). */
SKB_API int skb_flow_set_fiber_class(skb_flow *fl, int n_nodes, const double *D_1_0, const double *P_downsample_bc);
SKB_API int skb_flow_set_fiber_operators(skb_flow *fl, const double *A, const double *force_operator,
                                         const double *xs, const double *length_prev, const int *plus_bc_velocity);
/* fw (3 x N_f) = apply_fiber_force(x_fibers); x_fibers = per fiber [x; y; z; T] of 4n (fcfd.cpp:272-287) */
SKB_API int skb_flow_apply_fiber_force(skb_flow *fl, const double *x_fibers, double *fw);
/* res (4 N_f) = fc.matvec(x_fibers, v_fibers, v_fib_boundary); v_fibers 3 x N_f, v_fib_boundary 7 x n_fibers
 * (fiber_link_conditions of BodyContainer::calculate_link_conditions) or NULL for none (fcfd.cpp:216-232) */
SKB_API int skb_flow_fiber_matvec(skb_flow *fl, const double *x_fibers, const double *v_fibers,
                                  const double *v_fib_boundary, double *res);
/* One rank per GPU: under skb_flow_set_target_ranges (or a window that cuts no fiber) the resident operators are those
 * of the rank's OWN fibers -- skb_flow_set_fiber_operators, skb_flow_apply_fiber_force and skb_flow_fiber_matvec then
 * take / return the arrays of the own fibers only (x: 4 per own node, fw / v: 3 per own node, link conditions: 7 per
 * own fiber), as the reference's local solution vectors do.  Device-pointer forms, asynchronous on `stream`:
 *   d_fw_own = apply_fiber_force(d_x_own)  ->  [caller all-gathers fw]  ->  skb_flow_matvec_device  ->
 *   d_res_own = fc.matvec(d_x_own, first own-fiber rows of d_v_window, d_link_own) */
SKB_API int skb_flow_apply_fiber_force_device(skb_flow *fl, const double *d_x_fibers, double *d_fw, void *stream);
SKB_API int skb_flow_fiber_matvec_device(skb_flow *fl, const double *d_x_fibers, const double *d_v_fibers,
                                         const double *d_v_fib_boundary, double *d_res, void *stream);

/* The fiber block of System::apply_preconditioner (system.cpp:248-262): FiberContainerFiniteDifference::
 * apply_preconditioner (fcfd.cpp:331-339) is `A_LU_.solve(x)` per fiber; here the solve is a GEMV over the explicit
 * inverse, `fib.A_LU_.inverse()` (column-major 4n x 4n per fiber, concatenated in the order of
 * skb_flow_set_fiber_operators), uploaded once per timestep.  As a preconditioner it is a fixed linear operator, so
 * GMRES is indifferent to the rounding difference between LU substitution and the inverse.  (The shell block is
 * skb_dense_apply(SKB_DENSE_M_INV); bodies stay on the host.) */
SKB_API int skb_flow_set_fiber_preconditioner(skb_flow *fl, const double *A_inv);
SKB_API int skb_flow_apply_fiber_preconditioner(skb_flow *fl, const double *x_fibers, double *y);
SKB_API int skb_flow_apply_fiber_preconditioner_device(skb_flow *fl, const double *d_x_fibers, double *d_y,
                                                       void *stream);

/* System::apply_matvec (system.cpp:269-324) with the fiber operators on the device:
 *   fw = apply_fiber_force(x_fibers); v_all = the fused flow of skb_flow_matvec;
 *   res_fibers = fc.matvec(x_fibers, v_fibers, fiber_link_conditions)
 * and the periphery / body rows of v_all returned for shell.matvec / bc.matvec (periphery.cpp:38-47 ->
 * skb_dense_apply; body_container.cpp matvec stays host glue).  Needs the full target window.  One H2D of the
 * solution-sized inputs, one D2H of the solution-sized outputs; fw and v_fibers never leave the device. */
SKB_API int skb_flow_apply_matvec(skb_flow *fl, const double *x_fibers, const double *shell_density,
                                  const double *body_densities, const double *body_forces_torques,
                                  const double *fiber_link_conditions, double eta, double *res_fibers,
                                  double *v_shell, double *v_bodies);

/* The same with Periphery::matvec on the device as well (skelly_b200_dense.h, single-device handle holding
 * SKB_DENSE_STRESSLET_PLUS_COMPLEMENTARY of size 3 N_s): returns res_shell = stresslet_plus_complementary_ * x_shell +
 * v_shell (system.cpp:319, periphery.cpp:38-47) instead of v_shell; x_shell and v_shell never leave the device. */
struct skb_dense;
SKB_API int skb_flow_apply_matvec_dense(skb_flow *fl, struct skb_dense *dn, const double *x_fibers,
                                        const double *x_shell, const double *body_densities,
                                        const double *body_forces_torques, const double *fiber_link_conditions,
                                        double eta, double *res_fibers, double *res_shell, double *v_bodies);

/* Device-pointer form of skb_flow_apply_matvec(_dense): everything already on this flow's device, asynchronous on
 * `stream`.  dn == NULL: d_out_shell = v_shell; else res_shell (the handle holds the own periphery rows x all columns).
 * Works for a whole system and for a group member (OWN slices: d_x_fibers / d_res_fibers 4 per own fiber node,
 * d_x_shell / d_out_shell 3 per own periphery row, d_fiber_link_conditions 7 per own fiber or NULL, d_v_bodies 3 per own
 * body row; body densities / forces / torques complete).  This is one GMRES iteration's operator apply with nothing
 * crossing PCIe (a12: solver_hydro.cpp:42-48 calls exactly this per iteration). */
SKB_API int skb_flow_apply_matvec_device(skb_flow *fl, struct skb_dense *dn, const double *d_x_fibers,
                                         const double *d_x_shell, const double *d_body_densities,
                                         const double *d_body_forces, const double *d_body_torques,
                                         const double *d_fiber_link_conditions, double eta, double *d_res_fibers,
                                         double *d_out_shell, double *d_v_bodies, void *stream);

typedef struct skb_flow_stats {
    double device_ms;   /* CUDA-event time of the last call, first launch to last kernel (copies excluded) */
    double total_ms;    /* including H2D / D2H */
    int64_t n_pairs;    /* source x target pairs evaluated by the pair kernels */
    int32_t launches;
} skb_flow_stats;
SKB_API int skb_flow_last_stats(const skb_flow *fl, skb_flow_stats *out);
/* the matvec's dominant kernel (symmetric fiber-fiber block): duration of its last launch and the ordered pairs it
 * covered (skb_ctx_last_sym_kernel of the matvec's fiber evaluator) -- the roofline line of bench.py */
SKB_API int skb_flow_last_sym_kernel(const skb_flow *fl, double *ms, int64_t *pairs);

/* ---- skb_mflow: ONE process driving n GPUs ---------------------------------------------------------------------
 * The multi-device form the reference's rule "direct evaluators need a single MPI rank" (system.cpp:618-623) admits.
 * Same calls and array conventions as skb_flow, complete host arrays in and out; internally one group member per
 * device (see the group section above), each driven by its own host thread.  Whole fibers (balanced by node count),
 * periphery rows and body rows are block-partitioned over the devices; the periphery's dense operators are stored by
 * row blocks (periphery.cpp:387-417).  devices == NULL: 0..n-1.  (A device may be listed more than once: that
 * exercises the exchange protocol on a single GPU and is meant for tests.) */
typedef struct skb_mflow skb_mflow;
SKB_API int skb_mflow_create(const int *devices, int n, skb_mflow **out);
SKB_API int skb_mflow_destroy(skb_mflow *mf);
SKB_API int skb_mflow_n_devices(const skb_mflow *mf, int *n);
SKB_API int skb_mflow_set_fibers(skb_mflow *mf, const double *r_fib, const int *n_nodes, const double *length,
                                 int n_fibers);
SKB_API int skb_mflow_set_periphery(skb_mflow *mf, const double *node_pos, const double *node_normal, int64_t n_nodes);
SKB_API int skb_mflow_set_bodies(skb_mflow *mf, const double *node_pos, const double *node_normal, int64_t n_nodes,
                                 const double *centers, int n_bodies);
SKB_API int skb_mflow_set_self_exclusion(skb_mflow *mf, int fused);
SKB_API int skb_mflow_set_cross(skb_mflow *mf, int mode);
SKB_API int skb_mflow_set_overlap(skb_mflow *mf, int on); /* skb_flow_set_overlap on every member */
/* which fibers / periphery rows / body rows device `member` owns (any pointer may be NULL) */
SKB_API int skb_mflow_partition(skb_mflow *mf, int member, int *fiber_begin, int *fiber_end, int64_t *shell_begin,
                                int64_t *shell_end, int64_t *body_begin, int64_t *body_end);
/* the same partition without a GPU or a handle (rank-per-GPU hosts): out6 = the six arguments of
 * skb_flow_set_target_ranges for `member` of `n_members` */
SKB_API int skb_partition_query(const int *n_nodes, int n_fibers, int64_t n_shell, int64_t n_body, int n_members,
                                int member, int64_t *out6);
SKB_API int skb_mflow_set_fiber_class(skb_mflow *mf, int n_nodes, const double *D_1_0, const double *P_downsample_bc);
/* complete arrays of ALL fibers, as skb_flow_set_fiber_operators without a window; every device keeps its own slice */
SKB_API int skb_mflow_set_fiber_operators(skb_mflow *mf, const double *A, const double *force_operator,
                                          const double *xs, const double *length_prev, const int *plus_bc_velocity);
SKB_API int skb_mflow_set_fiber_preconditioner(skb_mflow *mf, const double *A_inv);
/* op = skb_dense_op (skelly_b200_dense.h); A row-major (3 N_s) x (3 N_s) */
SKB_API int skb_mflow_set_dense(skb_mflow *mf, int op, const double *A_rowmajor, int64_t n_rows, int64_t n_cols);
SKB_API int skb_mflow_matvec(skb_mflow *mf, const double *fib_forces, const double *shell_density,
                             const double *body_densities, const double *body_forces_torques, double eta,
                             double *v_all);
/* out_shell = res_shell when skb_mflow_set_dense(SKB_DENSE_STRESSLET_PLUS_COMPLEMENTARY, ...) was called, else v_shell */
SKB_API int skb_mflow_apply_matvec(skb_mflow *mf, const double *x_fibers, const double *x_shell,
                                   const double *body_densities, const double *body_forces_torques,
                                   const double *fiber_link_conditions, double eta, double *res_fibers,
                                   double *out_shell, double *v_bodies);
/* System::velocity_at_targets (system.cpp:330-384): targets block-partitioned over the devices */
SKB_API int skb_mflow_velocity_at_targets(skb_mflow *mf, const double *r_trg, int64_t n_trg, const double *fib_forces,
                                          const double *shell_density, const double *body_densities,
                                          const double *body_forces_torques, double eta, double *vel);
SKB_API int skb_mflow_last_stats(const skb_mflow *mf, skb_flow_stats *out);

#ifdef __cplusplus
}
#endif
#endif
