// skb_internal.hpp -- declarations shared by the translation units of libskelly_b200.so
#pragma once
#include "../../include/skelly_b200.h"

#include <cuda_runtime.h>
#include <functional>
#include <stddef.h>
#include <vector>

namespace skb {

struct DeviceInfo {
    int dev = 0;
    int num_sms = 0;
    int cc_major = 0, cc_minor = 0;
    int occupancy[2][4] = {{1, 1, 1, 1}, {1, 1, 1, 1}}; // [kind][T index: 1,2,4,8] resident CTAs / SM
};

struct LaunchPlan {
    int T = 4;               // targets per consumer thread
    int n_splits = 1;        // source splits (gridDim.y)
    int tiles_per_split = 1; // source tiles per split
    unsigned grid_x = 1;     // target tiles
};

struct DevBuf {
    void *ptr = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes); // grow-only; contents are NOT preserved
    void release();
};

int set_error(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
const char *last_error();
void count_launch(int n);
long long launch_count();

// Large host -> device copy from PAGEABLE caller memory (per-timestep operator uploads, hundreds of MB): the bytes are
// moved into pinned staging buffers by several host threads while the previous chunk is on the PCIe bus, instead of
// the driver's single-threaded staging path.  Synchronous for the host buffer (it may be reused on return); the device
// side is ordered on `st`.
int upload_pageable(void *d_dst, const void *h_src, size_t bytes, cudaStream_t st);

DeviceInfo query_device(int dev);
LaunchPlan plan_launch(const DeviceInfo &di, int kind, long long n_trg, int n_src_tiles, int force_T, int force_S);
int launch_pair_sum(const DeviceInfo &di, int kind, const double *d_r_src, const double *d_f_packed, long long n_src,
                    long long n_src_pad, const double *d_r_trg, long long n_trg, double *d_partial,
                    const LaunchPlan &plan, cudaStream_t st, const int *d_src_fid = nullptr,
                    const int *d_trg_fid = nullptr);
int launch_reduce(const double *d_partial, double *d_u, long long n_trg, int n_splits, double scale, int accumulate,
                  cudaStream_t st);

// kRaw: strengths as the caller ships them (3 or 9 per source); kNormalDensity: density, stresslet formed on the device;
// kPacked: already in the kernels' packed, padded layout ([n_pad*3] weighted Stokeslet strengths / [n_pad*6] sym6
// stresslet combinations), e.g. the landing zone peers pushed into -- no pack kernel runs
enum StrengthMode { kRaw = 0, kNormalDensity = 1, kPacked = 2 };

struct EvalOpts {
    double *d_u_sym = nullptr; // symmetric path: write the leading (self-interaction) rows here instead of d_u_out
    int sym_accumulate = -1;   // -1: same as `accumulate`
    double *d_u_rem = nullptr; // symmetric path: write the rows beyond the self-interaction block here instead of
                               // d_u_out + 3 n_src (group mode: the leading rows are partial sums that live elsewhere)
};

} // namespace skb

// ---- context internals (shared by skb_runtime.cu and skb_matvec.cu) ----
struct SourceSet {
    long long n = -1; // -1 = never set
    long long n_pad = 0;
    bool has_normals = false;
    skb::DevBuf r;       // padded positions
    bool has_weights = false;
    skb::DevBuf normals; // optional (double layer formed on device)
    skb::DevBuf weights; // optional per-source quadrature weight folded into the Stokeslet strengths
    skb::DevBuf f_raw;   // strengths as shipped by the caller (3 or 9 per source; all-gather landing zone)
    skb::DevBuf f_packed;
    const double *f_cur = nullptr; // packed strengths of the evaluation in flight (f_packed, or the caller's kPacked buffer)
    // opt-in fused self-exclusion: Stokeslet sources and the leading targets carry an id (the fiber
    // index); pairs with equal ids contribute 0.  excl_ids: [n_pad] source ids; excl_trg: [n_trg] target ids (-1 beyond
    // the leading n targets), built on demand for the plain kernel
    bool excl = false;
    skb::DevBuf excl_ids, excl_trg;
    long long excl_trg_n = -1;
    // symmetric (Newton's third law) path of the Stokeslet self-interaction, sym_kernels.cuh
    int self_state = -1;       // -1 unknown, 0 targets do not start with these sources, 1 they do
    bool sym_plan_valid = false;
    int sym_T = 0, sym_nb = 0, sym_items = 0, sym_part = 0, sym_parts = 1, sym_owned = 0;
    skb::DevBuf sym_item_buf, sym_row_begin, sym_P, sym_F, sym_row_ids /* block row I of every row of sym_P */, sym_flag;
    long long sym_pairs = 0; // ordered (target, source) pairs one launch of the symmetric kernel covers (both directions)
};

struct DeviceState {
    skb::DeviceInfo info;
    cudaStream_t stream = nullptr;
    cudaStream_t aux_stream = nullptr;              // remainder targets run beside the symmetric kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_fork2 = nullptr, ev_join2 = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
    cudaEvent_t ev_s0 = nullptr, ev_s1 = nullptr; // around the symmetric kernel's launch (its live duration inside a step)
    bool sym_timed = false;
    long long trg_begin = 0, n_trg = 0; // this device's block of the global target list
    // symmetric multi-device layout: targets = [all n_self leading targets | remainder rows [rem_begin, +rem_count)]
    long long rem_begin = 0, rem_count = 0;
    int sym_part = 0, sym_parts = 1;    // block rows of the self-interaction this device evaluates
    skb::DevBuf u_rs;                   // reduce-scatter output (this device's chunk of the leading rows)
    skb::DevBuf r_trg, u, partial, scratch;
    SourceSet src[2];
};

struct skb_ctx {
    std::vector<DeviceState> devs;
    long long n_trg = -1;
    int force_T = 0, force_S = 0;
    int sym_mode = -1; // -1 auto, 0 never, 1 whenever the sources are the leading targets
    bool last_was_sym = false;
    // single-process multi-device symmetric layout (host-pointer API): host copies of the positions decide it
    std::vector<double> h_trg, h_src[2];
    bool sym_layout = false;    // devices currently hold the targets in the symmetric layout
    bool layout_dirty = false;  // positions changed since the devices' target layout was decided
    long long n_self = 0;       // leading targets that are the Stokeslet sources
    skb_eval_stats stats{};
    bool kernel_events_pending = false; // device-pointer path: kernel_ms is read back lazily
    void *nccl = nullptr; // NcclGroup*, multi-device contexts only
};


namespace skb {
// pack -> pair sums -> split reduction for one device; strengths already resident at d_f_raw.
// result (=|+=) scale_mul * [1 | -3]/(8 pi) * sum.
int eval_on_device(skb_ctx *ctx, DeviceState &d, int kind, StrengthMode mode, const double *d_f_raw, double two_eta,
                   double *d_u_out, int accumulate, cudaStream_t st, bool record_events, int *launches,
                   LaunchPlan *plan_out, double scale_mul, const EvalOpts &opts = EvalOpts());

// strengths -> the kernels' packed, padded layout in s.f_packed (what eval_on_device does first); sets s.f_cur
int pack_on_device(DeviceState &d, int kind, StrengthMode mode, const double *d_f_raw, double two_eta, cudaStream_t st,
                   int *launches);

// fiber <-> periphery in one geometry pass (cross_kernels.cuh): plan and partial buffers of one (fiber rows, periphery)
// shape
struct CrossState {
    bool valid = false;
    long long node0 = -1, n_rows = -1, n_sh = -1, n_sh_pad = -1;
    int n_blocks = 0, n_items = 0, num_sms = 0;
    DevBuf items, row_begin, P, F;
    void release() {
        items.release(), row_begin.release(), P.release(), F.release();
        valid = false;
    }
};
long long cross_block_nodes();
// u_fib[n_rows x 3] (=|+=) scale_dl * (stresslet sums of ALL periphery nodes at fiber nodes [node0, node0 + n_rows));
// u_shell[n_sh x 3] (=|+=) scale_sl * (Stokeslet sums of those fiber nodes at every periphery node).
// r_fib / h: positions and packed Stokeslet strengths of all fiber nodes (padded); r_sh / s6: periphery, padded.
int cross_eval(CrossState &cs, const DeviceInfo &di, const double *d_r_fib, const double *d_h, long long node0,
               long long n_rows, const double *d_r_sh, const double *d_s6, long long n_sh, long long n_sh_pad,
               double scale_dl, double scale_sl, double *d_u_fib, int acc_fib, double *d_u_shell, int acc_shell,
               cudaStream_t st, int *launches);

// background row streamer (skb_dense.cu, stream_kernels.cuh): load the kernel and fix its attributes on `dev`
int dense_stream_preload(int dev);

struct SymItem;
long long sym_block_nodes(); // nodes per block of the symmetric kernel (its I side)
void build_sym_items(int nb, int part, int parts, int num_sms, std::vector<SymItem> &order,
                     std::vector<int> &row_begin);

// NCCL is bound at run time (dlopen "libnccl.so.2"): a single-GPU user never needs it, and a host that
// already loaded NCCL (PyTorch) shares that copy instead of getting a second one.
int nccl_group_create(size_t n, const std::function<int(int)> &device_of, void **out);
int nccl_group_allgather_inplace(void *group, void *const *bufs, size_t count_per_rank_doubles,
                                 const cudaStream_t *streams);
// recv[g] (count doubles) = sum over ranks of send[r][g*count : (g+1)*count]
int nccl_group_reduce_scatter(void *group, void *const *send, void *const *recv, size_t count_per_rank_doubles,
                              const cudaStream_t *streams);
void nccl_group_destroy(void *group);

} // namespace skb
