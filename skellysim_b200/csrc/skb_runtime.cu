// skb_runtime.cu -- host side of libskelly_b200.so: evaluator contexts, launch planning, the C ABI
// declared in include/skelly_b200.h, and the reference-named C++ entry points
// kernels::stokeslet_direct_gpu_impl / kernels::stresslet_direct_gpu_impl (SkellySim
// include/kernels.hpp:17-20) so the library links in place of src/core/kernels.cu.
//
// No CPU fallback exists here: every path ends in a CUDA launch or an error code.
#include "pair_kernels.cuh"
#include "sym_kernels.cuh"
#include "cross_kernels.cuh"
#include "skb_internal.hpp"

#ifndef SKB_CARVEOUT_MAX
#define SKB_CARVEOUT_MAX 1
#endif

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace skb {

// ------------------------------------------------------------------------------------------------
// errors / bookkeeping
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
const char *last_error() { return g_last_error.c_str(); }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------------
// launch planning
// ------------------------------------------------------------------------------------------------
template <int KIND, int T, int MINB, bool EXCL = false>
static cudaError_t launch_variant(const PairArgs &a, dim3 grid, cudaStream_t st) {
    using L = SmemLayout<KIND, T, EXCL>;
    auto kern = pair_sum_kernel<KIND, T, MINB, EXCL>;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total_bytes);
        if (e != cudaSuccess)
            return e;
        attr_set[dev] = true;
    }
    kern<<<grid, kCtaThreads, L::total_bytes, st>>>(a);
    return cudaGetLastError();
}

template <int KIND, int T, int MINB> static int occupancy_variant() {
    using L = SmemLayout<KIND, T>;
    auto kern = pair_sum_kernel<KIND, T, MINB>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total_bytes);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kCtaThreads, L::total_bytes) != cudaSuccess)
        nb = 1;
    return nb < 1 ? 1 : nb;
}

// min-blocks-per-SM hints (register caps at 128 threads/CTA): T=1:8 (64)  T=2:6 (85)  T=4:4 (128)  T=8:2 (255)
#define SKB_DISPATCH_T(KIND, T, CALL)                                                                                 \
    switch (T) {                                                                                                      \
    case 1: CALL(KIND, 1, 8); break;                                                                                  \
    case 2: CALL(KIND, 2, 6); break;                                                                                  \
    case 4: CALL(KIND, 4, 4); break;                                                                                  \
    default: CALL(KIND, 8, 2); break;                                                                                 \
    }

DeviceInfo query_device(int dev) {
    DeviceInfo di;
    di.dev = dev;
    cudaDeviceProp p;
    cudaSetDevice(dev);
    cudaGetDeviceProperties(&p, dev);
    di.num_sms = p.multiProcessorCount;
    di.cc_major = p.major;
    di.cc_minor = p.minor;
    const int Ts[4] = {1, 2, 4, 8};
    for (int k = 0; k < 2; ++k)
        for (int ti = 0; ti < 4; ++ti) {
            int occ = 1;
#define OCC_CALL(KIND, T, MINB) occ = occupancy_variant<KIND, T, MINB>()
            if (k == 0) {
                SKB_DISPATCH_T(kStokeslet, Ts[ti], OCC_CALL)
            } else {
                SKB_DISPATCH_T(kStresslet, Ts[ti], OCC_CALL)
            }
#undef OCC_CALL
            di.occupancy[k][ti] = occ;
        }
    return di;
}

// Pick targets-per-thread T and the number of source splits S.
//   cost(T,S) ~ (CTAs on the busiest SM) x (source tiles per CTA) x T x penalty(T, resident chains)
// The busiest SM runs ceil(n_ctas / num_sms) CTAs; a CTA's time is proportional to the pairs it owns
// (T targets/thread x tiles) because the FP64 pipe is shared by the co-resident CTAs.
LaunchPlan plan_launch(const DeviceInfo &di, int kind, long long n_trg, int n_src_tiles, int force_T, int force_S) {
    LaunchPlan best;
    double best_cost = 1e300;
    const int Ts[4] = {1, 2, 4, 8};
    for (int ti = 0; ti < 4; ++ti) {
        const int T = Ts[ti];
        if (force_T > 0 && T != force_T)
            continue;
        const int occ = di.occupancy[kind][ti];
        const long long tile_t = (long long)kCtaThreads * T;
        const long long n_tiles_t = (n_trg + tile_t - 1) / tile_t;
        // efficiency of the inner loop vs T (LDS + loop overhead amortised over T pairs); calibrated on B200
        // measured Stokeslet rates on B200: T=8 712, T=4 691, T=2 ~614 Gpairs/s (profiles/r1_probe_perf_v3.json,
        // profiles/r2_launches_bench.md); T=1 extrapolated
        // and the few-target sweep profiles/r2_small_targets.md (T=1 wins below ~2000 targets)
        const double t_pen = (T == 1) ? 1.18 : (T == 2) ? 1.11 : (T == 4) ? 0.985 : 0.955;
        const int s_max = std::min(n_src_tiles, kMaxSplits);
        for (int S = 1; S <= s_max; ++S) {
            if (force_S > 0 && S != std::min(force_S, s_max))
                continue;
            const int per = (n_src_tiles + S - 1) / S;
            const int S_eff = (n_src_tiles + per - 1) / per;
            if (S_eff != S && force_S <= 0)
                continue; // duplicate of a smaller S
            const long long n_ctas = n_tiles_t * S_eff;
            const long long per_sm = (n_ctas + di.num_sms - 1) / di.num_sms;
            // CTAs beyond the resident limit run in later waves; same total either way
            const double resident = (double)std::min<long long>(per_sm, occ);
            // latency hiding: want >= ~12 independent pair chains per scheduler (1 warp/CTA/SMSP, T chains each)
            // (T <= 2 evaluates two sources per iteration: 2 T chains per warp, stokeslet_two_sources)
            const double chains = resident * (T <= 2 ? 2 * T : T);
            // (and CTA slots left empty on the busiest SM cost latency hiding too: 441 CTAs of T=2 on 888 slots measured
            // slower than 875, profiles/r2_small_targets.md)
            const double fill = std::min(1.0, (double)per_sm / (double)occ);
            const double lat_pen = (chains >= 12.0 ? 1.0 : (12.0 / chains) * 0.5 + 0.5) * (1.0 + 0.25 * (1.0 - fill));
            // fixed per-CTA cost (barrier init, target staging, partial write) in units of source tiles
            const double cta_overhead = 0.5;
            double cost = (double)per_sm * ((double)per + cta_overhead) * T * t_pen * lat_pen;
            // reduction cost of S partial slabs, in the same units (tiny, breaks ties toward small S)
            cost += 1e-3 * S_eff;
            if (cost < best_cost) {
                best_cost = cost;
                best.T = T;
                best.n_splits = S_eff;
                best.tiles_per_split = per;
                best.grid_x = (unsigned)n_tiles_t;
            }
        }
    }
    return best;
}

int launch_pair_sum(const DeviceInfo &di, int kind, const double *d_r_src, const double *d_f_packed, long long n_src,
                    long long n_src_pad, const double *d_r_trg, long long n_trg, double *d_partial,
                    const LaunchPlan &plan, cudaStream_t st, const int *d_src_fid, const int *d_trg_fid) {
    PairArgs a;
    a.src_fid = d_src_fid;
    a.trg_fid = d_trg_fid;
    a.r_src = d_r_src;
    a.f_src = d_f_packed;
    a.r_trg = d_r_trg;
    a.partial = d_partial;
    a.n_trg = n_trg;
    a.n_src = n_src;
    a.n_src_tiles = (int)((n_src + kSrcTile - 1) / kSrcTile);
    a.tiles_per_split = plan.tiles_per_split;
    (void)n_src_pad;
    dim3 grid(plan.grid_x, plan.n_splits, 1);
    cudaError_t e = cudaSuccess;
#define LAUNCH_CALL(KIND, T, MINB) e = launch_variant<KIND, T, MINB>(a, grid, st)
#define LAUNCH_CALL_EXCL(KIND, T, MINB) e = launch_variant<KIND, T, MINB, true>(a, grid, st)
    if (kind == kStokeslet && d_src_fid) { // fused self-exclusion variant (same-id pairs contribute 0)
        if (!d_trg_fid)
            return set_error(SKB_ERR_INVALID, "launch_pair_sum: exclusion ids without target ids");
        SKB_DISPATCH_T(kStokeslet, plan.T, LAUNCH_CALL_EXCL)
    } else if (kind == kStokeslet) {
        SKB_DISPATCH_T(kStokeslet, plan.T, LAUNCH_CALL)
    } else {
        SKB_DISPATCH_T(kStresslet, plan.T, LAUNCH_CALL)
    }
#undef LAUNCH_CALL
#undef LAUNCH_CALL_EXCL
    if (e != cudaSuccess)
        return set_error(SKB_ERR_CUDA, "pair_sum_kernel launch failed: %s", cudaGetErrorString(e));
    count_launch(1);
    return SKB_OK;
}

int launch_reduce(const double *d_partial, double *d_u, long long n_trg, int n_splits, double scale, int accumulate,
                  cudaStream_t st) {
    const long long n3 = 3 * n_trg;
    const int bs = 256;
    if (n_splits >= 4 * kReduceWideWarps && n3 <= 32LL * 4096) // few targets, many splits
        reduce_partials_wide_kernel<<<(unsigned)((n3 + 31) / 32), 32 * kReduceWideWarps, 0, st>>>(
            d_partial, d_u, n3, n_splits, scale, accumulate);
    else
        reduce_partials_kernel<<<(unsigned)((n3 + bs - 1) / bs), bs, 0, st>>>(d_partial, d_u, n3, n_splits, scale,
                                                                              accumulate);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return set_error(SKB_ERR_CUDA, "reduce_partials_kernel launch failed: %s", cudaGetErrorString(e));
    count_launch(1);
    return SKB_OK;
}

// ------------------------------------------------------------------------------------------------
// device buffers
// ------------------------------------------------------------------------------------------------
int DevBuf::ensure(size_t bytes) {
    if (bytes <= cap)
        return SKB_OK;
    if (ptr)
        cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) {
        ptr = nullptr;
        return set_error(SKB_ERR_ALLOC, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    cap = want;
    return SKB_OK;
}
void DevBuf::release() {
    if (ptr)
        cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
}

// ------------------------------------------------------------------------------------------------
// pageable -> device upload through a pinned, multi-threaded staging ring
// ------------------------------------------------------------------------------------------------
namespace {
struct StagingRing {
    static constexpr int kBufs = 4;
    static constexpr size_t kChunk = 16u << 20;
    static constexpr int kMaxDev = 64;
    void *buf[kBufs] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[kMaxDev][kBufs] = {}; // events belong to a device: one set per device that uploads
    bool ev_ok[kMaxDev] = {};
    bool ok = false;
    std::mutex mu;
    bool init(int dev) {
        if (dev < 0 || dev >= kMaxDev)
            return false;
        if (!ok) {
            for (int i = 0; i < kBufs; ++i)
                if (cudaMallocHost(&buf[i], kChunk) != cudaSuccess) {
                    (void)cudaGetLastError();
                    return false;
                }
            ok = true;
        }
        if (!ev_ok[dev]) {
            for (int i = 0; i < kBufs; ++i)
                if (cudaEventCreateWithFlags(&ev[dev][i], cudaEventDisableTiming) != cudaSuccess) {
                    (void)cudaGetLastError();
                    return false;
                }
            ev_ok[dev] = true;
        }
        return true;
    }
};
StagingRing g_ring; // pinned host memory is not tied to a device: one ring per process (uploads are serialised)
} // namespace

int upload_pageable(void *d_dst, const void *h_src, size_t bytes, cudaStream_t st) {
    if (bytes == 0)
        return SKB_OK;
    std::lock_guard<std::mutex> lock(g_ring.mu);
    int dev = -1;
    (void)cudaGetDevice(&dev); // the caller made the stream's device current
    if (bytes < (8u << 20) || !g_ring.init(dev)) { // small, or no pinned memory to be had: the driver's own staging
        cudaError_t e = cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess)
            return set_error(SKB_ERR_CUDA, "cudaMemcpyAsync H2D (%zu bytes): %s", bytes, cudaGetErrorString(e));
        e = cudaStreamSynchronize(st);
        return e == cudaSuccess ? SKB_OK : set_error(SKB_ERR_CUDA, "H2D copy: %s", cudaGetErrorString(e));
    }
    const unsigned hw = std::thread::hardware_concurrency();
    const int n_thr = (int)std::max(1u, std::min(8u, hw ? hw / 2 : 4u));
    size_t done = 0;
    int k = 0;
    cudaError_t e = cudaSuccess;
    while (done < bytes && e == cudaSuccess) {
        const int b = k % StagingRing::kBufs;
        const size_t len = std::min(StagingRing::kChunk, bytes - done);
        if (k >= StagingRing::kBufs)
            e = cudaEventSynchronize(g_ring.ev[dev][b]); // the copy that last used this buffer has left it
        if (e != cudaSuccess)
            break;
        const char *src = (const char *)h_src + done;
        char *dst = (char *)g_ring.buf[b];
        const size_t piece = (len + n_thr - 1) / n_thr;
        std::vector<std::thread> th;
        for (int t = 1; t < n_thr; ++t) {
            const size_t o = std::min(len, (size_t)t * piece), l = std::min(piece, len - o);
            if (l)
                th.emplace_back([=] { std::memcpy(dst + o, src + o, l); });
        }
        std::memcpy(dst, src, std::min(piece, len));
        for (auto &t : th)
            t.join();
        e = cudaMemcpyAsync((char *)d_dst + done, dst, len, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess)
            e = cudaEventRecord(g_ring.ev[dev][b], st);
        done += len;
        ++k;
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(st);
    if (e != cudaSuccess)
        return set_error(SKB_ERR_CUDA, "staged H2D copy (%zu bytes): %s", bytes, cudaGetErrorString(e));
    return SKB_OK;
}

} // namespace skb

using namespace skb;

#define CUDA_TRY(expr)                                                                                                \
    do {                                                                                                              \
        cudaError_t _e = (expr);                                                                                      \
        if (_e != cudaSuccess)                                                                                        \
            return set_error(SKB_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                      \
                             cudaGetErrorString(_e));                                                                 \
    } while (0)
#define SKB_TRY(expr)                                                                                                 \
    do {                                                                                                              \
        int _rc = (expr);                                                                                             \
        if (_rc != SKB_OK)                                                                                            \
            return _rc;                                                                                               \
    } while (0)

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
static int check_kind(int kind) {
    if (kind != SKB_STOKESLET && kind != SKB_STRESSLET)
        return set_error(SKB_ERR_INVALID, "unknown kernel kind %d", kind);
    return SKB_OK;
}

static int ctx_create_impl(const int *ids, int n, skb_ctx **out) {
    if (!out)
        return set_error(SKB_ERR_INVALID, "skb_ctx_create: out == NULL");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return set_error(SKB_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                         e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (n < 1 || n > n_dev)
        return set_error(SKB_ERR_INVALID, "skb_ctx_create: n_gpus=%d but %d device(s) visible", n, n_dev);
    std::unique_ptr<skb_ctx> ctx(new skb_ctx);
    if (const char *e = getenv("SKB_SYMMETRIC")) // default of skb_ctx_set_symmetric for contexts created from now on
        ctx->sym_mode = std::max(-1, std::min(1, atoi(e)));
    ctx->devs.resize(n);
    for (int g = 0; g < n; ++g) {
        DeviceState &d = ctx->devs[g];
        const int dev = ids ? ids[g] : g;
        if (dev < 0 || dev >= n_dev)
            return set_error(SKB_ERR_INVALID, "device id %d out of range", dev);
        CUDA_TRY(cudaSetDevice(dev));
        d.info = query_device(dev);
        if (d.info.cc_major < 10)
            return set_error(SKB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", dev,
                             d.info.cc_major, d.info.cc_minor);
        CUDA_TRY(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
        // (d.aux_stream is created on first use by the symmetric path: most contexts of a flow never need one, and a
        // device has a finite number of hardware queues for independent streams)
        CUDA_TRY(cudaEventCreateWithFlags(&d.ev_fork, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&d.ev_join, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&d.ev_fork2, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&d.ev_join2, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreate(&d.ev_t0));
        CUDA_TRY(cudaEventCreate(&d.ev_t1));
        CUDA_TRY(cudaEventCreate(&d.ev_k0));
        CUDA_TRY(cudaEventCreate(&d.ev_k1));
        CUDA_TRY(cudaEventCreate(&d.ev_s0));
        CUDA_TRY(cudaEventCreate(&d.ev_s1));
    }
    if (n > 1)
        SKB_TRY(nccl_group_create(ctx->devs.size(), [&](int g) { return ctx->devs[g].info.dev; }, &ctx->nccl));
    *out = ctx.release();
    return SKB_OK;
}

extern "C" {

const char *skb_version(void) { return "skelly_b200 0.1.0 (sm_100a)"; }
const char *skb_last_error_string(void) { return last_error(); }
int64_t skb_launch_count(void) { return launch_count(); }

int skb_device_count(int *n) {
    if (!n)
        return set_error(SKB_ERR_INVALID, "skb_device_count: NULL");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        *n = 0;
        return set_error(SKB_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *n = c;
    return SKB_OK;
}

int skb_ctx_create(int n_gpus, skb_ctx **out) { return ctx_create_impl(nullptr, n_gpus, out); }
int skb_ctx_create_on(const int *device_ids, int n_gpus, skb_ctx **out) {
    if (!device_ids)
        return set_error(SKB_ERR_INVALID, "skb_ctx_create_on: device_ids == NULL");
    return ctx_create_impl(device_ids, n_gpus, out);
}

int skb_ctx_destroy(skb_ctx *ctx) {
    if (!ctx)
        return SKB_OK;
    for (auto &d : ctx->devs) {
        cudaSetDevice(d.info.dev);
        if (d.stream)
            cudaStreamSynchronize(d.stream);
        d.r_trg.release();
        d.u.release();
        d.partial.release();
        d.scratch.release();
        d.u_rs.release();
        for (auto &s : d.src) {
            s.r.release();
            s.normals.release();
            s.weights.release();
            s.sym_item_buf.release();
            s.sym_row_begin.release();
            s.sym_P.release();
            s.sym_F.release();
            s.sym_row_ids.release();
            s.sym_flag.release();
            s.f_raw.release();
            s.f_packed.release();
            s.excl_ids.release();
            s.excl_trg.release();
        }
        if (d.ev_t0) cudaEventDestroy(d.ev_t0);
        if (d.ev_t1) cudaEventDestroy(d.ev_t1);
        if (d.ev_k0) cudaEventDestroy(d.ev_k0);
        if (d.ev_k1) cudaEventDestroy(d.ev_k1);
        if (d.ev_s0) cudaEventDestroy(d.ev_s0);
        if (d.ev_s1) cudaEventDestroy(d.ev_s1);
        if (d.ev_fork) cudaEventDestroy(d.ev_fork);
        if (d.ev_join) cudaEventDestroy(d.ev_join);
        if (d.ev_fork2) cudaEventDestroy(d.ev_fork2);
        if (d.ev_join2) cudaEventDestroy(d.ev_join2);
        if (d.aux_stream) cudaStreamDestroy(d.aux_stream);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    if (ctx->nccl)
        nccl_group_destroy(ctx->nccl);
    delete ctx;
    return SKB_OK;
}

int skb_ctx_n_gpus(const skb_ctx *ctx, int *n) {
    if (!ctx || !n)
        return set_error(SKB_ERR_INVALID, "skb_ctx_n_gpus: NULL");
    *n = (int)ctx->devs.size();
    return SKB_OK;
}

int skb_ctx_set_tuning(skb_ctx *ctx, int T, int S) {
    if (!ctx)
        return set_error(SKB_ERR_INVALID, "skb_ctx_set_tuning: NULL ctx");
    if (!(T == 0 || T == 1 || T == 2 || T == 4 || T == 8) || S < 0)
        return set_error(SKB_ERR_INVALID, "tuning: targets_per_thread must be 0/1/2/4/8, source_splits >= 0");
    ctx->force_T = T;
    ctx->force_S = S;
    return SKB_OK;
}

int skb_ctx_set_symmetric(skb_ctx *ctx, int mode) {
    if (!ctx || mode < -1 || mode > 1)
        return set_error(SKB_ERR_INVALID, "skb_ctx_set_symmetric: mode must be -1 (auto), 0 (off) or 1 (on)");
    ctx->sym_mode = mode;
    return SKB_OK;
}

int skb_plan_query(int kind, int64_t n_trg, int64_t n_src, int num_sms, const int *occupancy4, int force_T,
                   int force_S, int *T, int *n_splits, int *tiles_per_split, int *grid_x) {
    if ((kind != 0 && kind != 1) || n_trg < 1 || n_src < 1 || num_sms < 1 || !occupancy4 || !T || !n_splits ||
        !tiles_per_split || !grid_x)
        return set_error(SKB_ERR_INVALID, "skb_plan_query: bad arguments");
    DeviceInfo di;
    di.num_sms = num_sms;
    for (int i = 0; i < 4; ++i)
        di.occupancy[kind][i] = std::max(1, occupancy4[i]);
    const LaunchPlan p = plan_launch(di, kind, n_trg, (int)((n_src + kSrcTile - 1) / kSrcTile), force_T, force_S);
    *T = p.T;
    *n_splits = p.n_splits;
    *tiles_per_split = p.tiles_per_split;
    *grid_x = (int)p.grid_x;
    return SKB_OK;
}

int skb_sym_plan_query(int n_blocks, int part, int n_parts, int num_sms, int max_items, int *items4, int *n_items,
                       int *row_begin) {
    if (n_blocks < 1 || n_parts < 1 || part < 0 || part >= n_parts || num_sms < 1 || !n_items)
        return set_error(SKB_ERR_INVALID, "skb_sym_plan_query: bad arguments");
    std::vector<SymItem> order;
    std::vector<int> rb;
    build_sym_items(n_blocks, part, n_parts, num_sms, order, rb);
    *n_items = (int)order.size();
    if (items4)
        for (int i = 0; i < (int)order.size() && i < max_items; ++i) {
            items4[4 * i + 0] = order[i].I;
            items4[4 * i + 1] = order[i].g0;
            items4[4 * i + 2] = order[i].g1;
            items4[4 * i + 3] = order[i].slot;
        }
    if (row_begin)
        for (int b = 0; b <= n_blocks; ++b)
            row_begin[b] = rb[b];
    return SKB_OK;
}

int skb_sym_groups_per_block(void) { return kSymGroupsPerBlock; }

int skb_ctx_last_eval_was_symmetric(const skb_ctx *ctx, int *yes) {
    if (!ctx || !yes)
        return set_error(SKB_ERR_INVALID, "skb_ctx_last_eval_was_symmetric: NULL");
    *yes = ctx->last_was_sym ? 1 : 0;
    return SKB_OK;
}

int skb_ctx_last_sym_kernel(const skb_ctx *ctx, double *ms, int64_t *pairs) {
    if (!ctx || !ms || !pairs)
        return set_error(SKB_ERR_INVALID, "skb_ctx_last_sym_kernel: NULL");
    *ms = 0;
    *pairs = 0;
    const DeviceState &d = ctx->devs[0];
    if (!d.sym_timed)
        return SKB_OK;
    float t = 0;
    cudaSetDevice(d.info.dev);
    if (cudaEventElapsedTime(&t, d.ev_s0, d.ev_s1) == cudaSuccess) {
        *ms = t;
        *pairs = d.src[SKB_STOKESLET].sym_pairs;
    } else {
        (void)cudaGetLastError(); // not finished yet
    }
    return SKB_OK;
}

int skb_ctx_set_sym_partition(skb_ctx *ctx, int part, int n_parts) {
    if (!ctx || n_parts < 1 || part < 0 || part >= n_parts)
        return set_error(SKB_ERR_INVALID, "skb_ctx_set_sym_partition: need 0 <= part < n_parts");
    if (ctx->devs.size() != 1)
        return set_error(SKB_ERR_INVALID, "skb_ctx_set_sym_partition: single-GPU contexts only");
    ctx->devs[0].sym_part = part;
    ctx->devs[0].sym_parts = n_parts;
    return SKB_OK;
}

int skb_last_eval_stats(const skb_ctx *ctx, skb_eval_stats *out) {
    if (!ctx || !out)
        return set_error(SKB_ERR_INVALID, "skb_last_eval_stats: NULL");
    *out = ctx->stats;
    if (ctx->kernel_events_pending) {
        // asynchronous evaluation: the events are valid once the caller's stream has passed them
        const DeviceState &d = ctx->devs[0];
        float ms = 0;
        cudaSetDevice(d.info.dev);
        if (cudaEventElapsedTime(&ms, d.ev_k0, d.ev_k1) == cudaSuccess)
            out->kernel_ms = ms;
        else
            (void)cudaGetLastError(); // not ready yet: leave 0
    }
    return SKB_OK;
}

} // extern "C"

// ------------------------------------------------------------------------------------------------
// positions
// ------------------------------------------------------------------------------------------------
static void partition_targets(skb_ctx *ctx, long long n_trg) {
    const long long P = (long long)ctx->devs.size();
    const long long chunk = (n_trg + P - 1) / P;
    for (long long g = 0; g < P; ++g) {
        DeviceState &d = ctx->devs[g];
        d.trg_begin = std::min(n_trg, g * chunk);
        d.n_trg = std::min(n_trg, (g + 1) * chunk) - d.trg_begin;
    }
    ctx->n_trg = n_trg;
}

static int set_targets_impl(skb_ctx *ctx, const double *r_trg, long long n_trg, bool on_device, cudaStream_t user) {
    // on_device: `user` is used verbatim (0 = CUDA's legacy default stream); host path: the context's stream
    if (!ctx)
        return set_error(SKB_ERR_INVALID, "set_targets: NULL ctx");
    if (n_trg < 0 || (n_trg > 0 && !r_trg))
        return set_error(SKB_ERR_INVALID, "set_targets: bad arguments (n_trg=%lld)", n_trg);
    if (on_device && ctx->devs.size() != 1)
        return set_error(SKB_ERR_INVALID, "device-pointer entry points need a single-GPU context");
    if (!on_device && ctx->devs.size() > 1) {
        // multi-device context: the device layout of the targets (plain block partition, or the symmetric layout
        // when they start with the Stokeslet sources) is decided at the next evaluation, see update_layout()
        ctx->h_trg.assign(r_trg, r_trg + 3 * n_trg);
        ctx->n_trg = n_trg;
        ctx->layout_dirty = true;
        return SKB_OK;
    }
    partition_targets(ctx, n_trg);
    for (auto &d : ctx->devs) {
        d.src[0].self_state = -1;
        d.src[1].self_state = -1;
        d.src[0].excl_trg_n = -1;
        if (d.n_trg == 0)
            continue;
        CUDA_TRY(cudaSetDevice(d.info.dev));
        SKB_TRY(d.r_trg.ensure((size_t)d.n_trg * 24));
        SKB_TRY(d.u.ensure((size_t)d.n_trg * 24));
        cudaStream_t st = on_device ? user : d.stream;
        CUDA_TRY(cudaMemcpyAsync(d.r_trg.ptr, r_trg + 3 * d.trg_begin, (size_t)d.n_trg * 24,
                                 on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    }
    if (!on_device)
        for (auto &d : ctx->devs) {
            CUDA_TRY(cudaSetDevice(d.info.dev));
            CUDA_TRY(cudaStreamSynchronize(d.stream));
        }
    return SKB_OK;
}

namespace skb {
static long long sym_mem_budget();
static int sym_owned_rows(int nb, int part, int parts);
}

// Multi-device contexts (host-pointer API): put the targets on the devices.  Plain layout: contiguous blocks.
// Symmetric layout (the targets start with the Stokeslet sources, no stresslet sources in this context): every
// device holds ALL n_self leading targets -- it evaluates its serpentine share of the self-interaction's block rows
// and produces partial sums for all of them -- followed by its block of the remaining targets.
static int update_layout(skb_ctx *ctx) {
    if (ctx->devs.size() < 2 || !ctx->layout_dirty)
        return SKB_OK;
    const long long P = (long long)ctx->devs.size();
    const long long n_trg = ctx->n_trg;
    const long long n_sl = ctx->devs[0].src[SKB_STOKESLET].n, n_dl = ctx->devs[0].src[SKB_STRESSLET].n;
    const long long block = (long long)kSymThreads * kSymT;
    bool want = ctx->sym_mode != 0 && n_dl <= 0 && n_sl >= (ctx->sym_mode == 1 ? 2 * block : 4096) && n_trg >= n_sl &&
                (long long)ctx->h_src[SKB_STOKESLET].size() == 3 * n_sl &&
                std::memcmp(ctx->h_trg.data(), ctx->h_src[SKB_STOKESLET].data(), (size_t)n_sl * 24) == 0;
    if (want) {
        const long long n_pad = ctx->devs[0].src[SKB_STOKESLET].n_pad;
        if (((n_pad / block + P - 1) / P) * n_pad * 24 > sym_mem_budget())
            want = false;
    }
    ctx->sym_layout = want;
    ctx->n_self = want ? n_sl : 0;
    if (!want) {
        partition_targets(ctx, n_trg);
        for (auto &d : ctx->devs) {
            d.sym_part = 0;
            d.sym_parts = 1;
            d.rem_begin = d.rem_count = 0;
        }
    } else {
        const long long n_rem = n_trg - n_sl, chunk = (n_rem + P - 1) / P;
        for (long long g = 0; g < P; ++g) {
            DeviceState &d = ctx->devs[g];
            d.rem_begin = std::min(n_rem, g * chunk);
            d.rem_count = std::min(n_rem, (g + 1) * chunk) - d.rem_begin;
            d.trg_begin = 0;
            d.n_trg = n_sl + d.rem_count;
            d.sym_part = (int)g;
            d.sym_parts = (int)P;
        }
    }
    for (auto &d : ctx->devs) {
        d.src[0].self_state = want ? 1 : 0; // decided on the host copies
        d.src[1].self_state = 0;
        if (d.n_trg == 0)
            continue;
        CUDA_TRY(cudaSetDevice(d.info.dev));
        const long long chunk_f = want ? (ctx->n_self + P - 1) / P : 0;
        const long long u_rows = std::max(d.n_trg, chunk_f * P);
        SKB_TRY(d.r_trg.ensure((size_t)d.n_trg * 24));
        SKB_TRY(d.u.ensure((size_t)u_rows * 24));
        if (want) {
            SKB_TRY(d.u_rs.ensure((size_t)chunk_f * 24));
            CUDA_TRY(cudaMemcpyAsync(d.r_trg.ptr, ctx->h_trg.data(), (size_t)ctx->n_self * 24, cudaMemcpyHostToDevice,
                                     d.stream));
            if (d.rem_count > 0)
                CUDA_TRY(cudaMemcpyAsync((double *)d.r_trg.ptr + 3 * ctx->n_self,
                                         ctx->h_trg.data() + 3 * (ctx->n_self + d.rem_begin), (size_t)d.rem_count * 24,
                                         cudaMemcpyHostToDevice, d.stream));
        } else {
            CUDA_TRY(cudaMemcpyAsync(d.r_trg.ptr, ctx->h_trg.data() + 3 * d.trg_begin, (size_t)d.n_trg * 24,
                                     cudaMemcpyHostToDevice, d.stream));
        }
    }
    for (auto &d : ctx->devs) {
        CUDA_TRY(cudaSetDevice(d.info.dev));
        CUDA_TRY(cudaStreamSynchronize(d.stream));
    }
    ctx->layout_dirty = false;
    return SKB_OK;
}

static int set_sources_impl(skb_ctx *ctx, int kind, const double *r_src, long long n_src, bool on_device,
                            cudaStream_t user) {
    if (!ctx)
        return set_error(SKB_ERR_INVALID, "set_sources: NULL ctx");
    SKB_TRY(check_kind(kind));
    if (n_src < 0 || (n_src > 0 && !r_src))
        return set_error(SKB_ERR_INVALID, "set_sources: bad arguments (n_src=%lld)", n_src);
    if (n_src > (1LL << 31) - 1 - kSrcTile)
        return set_error(SKB_ERR_INVALID, "set_sources: n_src=%lld exceeds the supported range", n_src);
    if (on_device && ctx->devs.size() != 1)
        return set_error(SKB_ERR_INVALID, "device-pointer entry points need a single-GPU context");
    const long long kPadTo = 1024; // multiple of every source tile / symmetric block size in use
    const long long n_pad = ((n_src + kPadTo - 1) / kPadTo) * kPadTo;
    const int fdim_raw = kind == SKB_STOKESLET ? 3 : 9;
    const int fdim_packed = kind == SKB_STOKESLET ? 3 : 6;
    const long long P = (long long)ctx->devs.size();
    const long long chunk = (n_src + P - 1) / P; // all-gather slot per device
    if (!on_device && P > 1) {
        ctx->h_src[kind].assign(r_src, r_src + 3 * n_src);
        ctx->layout_dirty = true;
    }
    for (auto &d : ctx->devs) {
        SourceSet &s = d.src[kind];
        s.n = n_src;
        s.n_pad = n_pad;
        s.has_normals = false;
        s.has_weights = false;
        s.excl = false;    // ids belong to one set of sources
        s.excl_trg_n = -1;
        s.self_state = -1; // (the symmetric plan depends on the block count only and survives position updates)
        if (n_src == 0)
            continue;
        CUDA_TRY(cudaSetDevice(d.info.dev));
        cudaStream_t st = on_device ? user : d.stream;
        SKB_TRY(s.r.ensure((size_t)n_pad * 24));
        SKB_TRY(s.f_raw.ensure((size_t)chunk * P * fdim_raw * 8));
        SKB_TRY(s.f_packed.ensure((size_t)n_pad * fdim_packed * 8));
        const double *d_in = r_src;
        if (!on_device) {
            SKB_TRY(d.scratch.ensure((size_t)n_src * 24));
            CUDA_TRY(cudaMemcpyAsync(d.scratch.ptr, r_src, (size_t)n_src * 24, cudaMemcpyHostToDevice, st));
            d_in = (const double *)d.scratch.ptr;
        }
        const int bs = 256;
        pad_positions_kernel<<<(unsigned)((n_pad * 3 + bs - 1) / bs), bs, 0, st>>>(d_in, (double *)s.r.ptr, n_src,
                                                                                   n_pad);
        CUDA_TRY(cudaGetLastError());
        count_launch(1);
    }
    if (!on_device)
        for (auto &d : ctx->devs) {
            CUDA_TRY(cudaSetDevice(d.info.dev));
            CUDA_TRY(cudaStreamSynchronize(d.stream));
        }
    return SKB_OK;
}

// ------------------------------------------------------------------------------------------------
// evaluation
// ------------------------------------------------------------------------------------------------

// device-side part for one device: pack -> pair kernel -> reduce into d_u_out.  f_raw already resident.

// ------------------------------------------------------------------------------------------------
// symmetric Stokeslet self-interaction (sym_kernels.cuh)
// ------------------------------------------------------------------------------------------------
namespace skb {

__global__ void self_check_kernel(const double *__restrict__ a, const double *__restrict__ b, long long n,
                                  int *__restrict__ differs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && __double_as_longlong(a[i]) != __double_as_longlong(b[i]))
        *differs = 1;
}

template <int T, int MINB, bool EXCL>
static cudaError_t launch_sym(const SymArgs &a, int n_items, cudaStream_t st) {
    using L = SymSmem<T>;
    auto kern = pair_sym_kernel<T, MINB, EXCL>;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total_bytes);
        if (e != cudaSuccess)
            return e;
#if SKB_CARVEOUT_MAX
        // one shared-memory configuration for this kernel and the background row streamer (stream_kernels.cuh): an SM
        // switches configuration only when idle, so a different preference would keep the two from sharing it
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess)
            return e;
#endif
        attr_set[dev] = true;
    }
    kern<<<n_items, kSymThreads, L::total_bytes, st>>>(a);
    return cudaGetLastError();
}

// bytes the reverse partials of the symmetric kernel may occupy: SKB_SYM_MAX_BYTES, else 30 % of the device memory
// (at least 8 GiB): 47 GB at 1e6 nodes on one 180 GB B200, 1/P of that per rank with the row partition
static long long sym_mem_budget() {
    static long long cached = -1; // all devices of a box are alike; the environment is read once
    if (cached >= 0)
        return cached;
    if (const char *e = getenv("SKB_SYM_MAX_BYTES"))
        return cached = atoll(e);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess)
        return cached = 8LL << 30;
    return cached = std::max<long long>(8LL << 30, (long long)(0.30 * (double)total_b));
}
long long sym_block_nodes() { return (long long)kSymThreads * kSymT; }
static int sym_owned_rows(int nb, int part, int parts) {
    int n = 0;
    for (int I = 0; I < nb; ++I)
        n += sym_row_owner(I, parts) == part;
    return n;
}

// Work items of the symmetric kernel: (I, groups [g0, g1)) over the upper triangle of nb blocks INCLUDING the block
// diagonal (row I starts at its own first group), restricted to the block rows owned by `part`.  Units are 32-node groups (kSymGroupsPerBlock per block).  Guided sizes: most of the
// work goes out in large items (about a quarter of a CTA slot's share each), the last ~30 % in items a quarter of that
// size and the last ~8 % in single stages (4 groups = 17 us of one CTA slot), so that all CTA slots drain together: the
// hardware hands the next item to whichever slot frees up first, in `order` (large items first).  item.slot is the
// item's row-major position, row_begin[b]..row_begin[b+1] the slots of row b (their forward partials are summed in
// that order by sym_reduce_kernel).
void build_sym_items(int nb, int part, int parts, int num_sms, std::vector<SymItem> &order,
                     std::vector<int> &row_begin) {
    const int gpb = kSymGroupsPerBlock, sg = kSymStageGroups;
    const int occ = kSymMinB;
    auto env_int = [](const char *name, int dflt) {
        const char *e = getenv(name);
        return e && atoi(e) > 0 ? atoi(e) : dflt;
    };
    static const int waves = env_int("SKB_SYM_WAVES", 4);      // large items per CTA slot (if all work were large items)
    static const int pct_mid = env_int("SKB_SYM_PCT_MID", 30);  // last x % of the work: quarter-size items
    static const int pct_fine = env_int("SKB_SYM_PCT_FINE", 8); // last x % of the work: single stages
    long long work = 0; // groups x blocks owned by this part
    for (int I = 0; I < nb; ++I)
        if (sym_row_owner(I, parts) == part)
            work += (long long)gpb * (nb - I);
    const long long slots = (long long)num_sms * occ;
    long long big = work / std::max<long long>(1, slots * waves);
    big = std::max<long long>(sg, std::min<long long>(big / sg * sg, 1024LL * sg));
    const long long mid = std::max<long long>(sg, big / 4 / sg * sg);
    struct Piece {
        int I, g0, g1, prow;
    };
    std::vector<Piece> pieces;
    int prow = 0;
    for (int I = 0; I < nb; ++I) {
        if (sym_row_owner(I, parts) != part)
            continue;
        const int my_prow = prow++; // every owned row has a P row, in increasing I (sym_reduce_kernel counts the same way)
        const int first = gpb * I, len = gpb * (nb - I); // from the block's own groups (the block diagonal) onwards
        const int n_chunks = (int)((len + big - 1) / big);
        for (int c = 0; c < n_chunks; ++c) { // near-equal chunks, cut at stage boundaries
            const int a = (int)((long long)(len / sg) * c / n_chunks) * sg;
            const int b = c + 1 == n_chunks ? len : (int)((long long)(len / sg) * (c + 1) / n_chunks) * sg;
            if (b > a)
                pieces.push_back(Piece{I, first + a, first + b, my_prow});
        }
    }
    // the tail of the launch order is cut finer
    std::stable_sort(pieces.begin(), pieces.end(),
                     [](const Piece &x, const Piece &y) { return (x.g1 - x.g0) > (y.g1 - y.g0); });
    std::vector<Piece> fine;
    long long done = 0;
    // (at most ~6 items of a tier per CTA slot: more would only add forward-partial slabs)
    const long long fine_work = std::min<long long>(work * pct_fine / 100, slots * 6 * sg);
    const long long mid_work = std::min<long long>(work * pct_mid / 100, fine_work + slots * 6 * mid);
    for (const Piece &pc : pieces) {
        const long long len = pc.g1 - pc.g0;
        const long long left = work - done;
        const long long unit = left <= fine_work ? sg : left <= mid_work ? mid : len;
        for (long long g = pc.g0; g < pc.g1; g += unit)
            fine.push_back(Piece{pc.I, (int)g, (int)std::min<long long>(pc.g1, g + unit), pc.prow});
        done += len;
    }
    // slots: row-major
    std::sort(fine.begin(), fine.end(),
              [](const Piece &x, const Piece &y) { return x.I != y.I ? x.I < y.I : x.g0 < y.g0; });
    std::vector<SymItem> items(fine.size());
    row_begin.assign(nb + 1, 0);
    for (size_t i = 0; i < fine.size(); ++i) {
        items[i] = SymItem{fine[i].I, fine[i].g0, fine[i].g1, (int)i, fine[i].prow, 0};
        row_begin[fine[i].I + 1] = (int)i + 1;
    }
    for (int b = 0; b < nb; ++b) // rows without items inherit the running count
        row_begin[b + 1] = std::max(row_begin[b + 1], row_begin[b]);
    order = items;
    std::stable_sort(order.begin(), order.end(),
                     [](const SymItem &x, const SymItem &y) { return (x.g1 - x.g0) > (y.g1 - y.g0); });
}

// Are the first n_src targets bit-identical to the Stokeslet sources?  (decided once per set of positions)
static int ensure_self_state(DeviceState &d, SourceSet &s, cudaStream_t st) {
    if (s.self_state >= 0)
        return SKB_OK;
    if (d.n_trg < s.n || s.n <= 0) {
        s.self_state = 0;
        return SKB_OK;
    }
    SKB_TRY(s.sym_flag.ensure(sizeof(int)));
    CUDA_TRY(cudaMemsetAsync(s.sym_flag.ptr, 0, sizeof(int), st));
    const long long n3 = 3 * s.n;
    self_check_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, st>>>((const double *)d.r_trg.ptr, (const double *)s.r.ptr,
                                                                    n3, (int *)s.sym_flag.ptr);
    CUDA_TRY(cudaGetLastError());
    count_launch(1);
    int differs = 1;
    CUDA_TRY(cudaMemcpyAsync(&differs, s.sym_flag.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    s.self_state = differs ? 0 : 1;
    return SKB_OK;
}

// Decide whether the symmetric path applies and make sure its plan / buffers exist.
static int sym_prepare(skb_ctx *ctx, DeviceState &d, cudaStream_t st, int *use) {
    *use = 0;
    SourceSet &s = d.src[SKB_STOKESLET];
    const int T = kSymT;
    const long long block = (long long)kSymThreads * T;
    if (s.n < 4096 && ctx->sym_mode != 1)
        return SKB_OK; // too small to matter
    if (s.n < 2 * block || d.n_trg < s.n)
        return SKB_OK;
    const long long nb = s.n_pad / block; // n_pad is a multiple of 1024
    if ((long long)sym_owned_rows((int)nb, d.sym_part, d.sym_parts) * s.n_pad * 24 > sym_mem_budget())
        return SKB_OK;
    SKB_TRY(ensure_self_state(d, s, st));
    if (s.self_state != 1)
        return SKB_OK;
    if (!s.sym_plan_valid || s.sym_T != T || s.sym_nb != (int)nb || s.sym_part != d.sym_part ||
        s.sym_parts != d.sym_parts) {
        std::vector<SymItem> order;
        std::vector<int> row_begin;
        build_sym_items((int)nb, d.sym_part, d.sym_parts, d.info.num_sms, order, row_begin);
        SKB_TRY(s.sym_item_buf.ensure(order.size() * sizeof(SymItem) + 16));
        SKB_TRY(s.sym_row_begin.ensure(row_begin.size() * sizeof(int)));
        if (s.sym_P.ensure((size_t)sym_owned_rows((int)nb, d.sym_part, d.sym_parts) * (size_t)s.n_pad * 24) !=
            SKB_OK) { // not enough free memory for the reverse partials: the plain kernel serves this geometry
            (void)cudaGetLastError();
            if (ctx->sym_layout) // every device holds ALL leading targets there and the caller reduce-scatters the
                                 // partial sums: a silent switch to the plain kernel would count them P times
                return set_error(SKB_ERR_ALLOC, "symmetric multi-device layout: no memory for the reverse partials "
                                                "(%zu bytes); free memory or skb_ctx_set_symmetric(ctx, 0)",
                                 (size_t)sym_owned_rows((int)nb, d.sym_part, d.sym_parts) * (size_t)s.n_pad * 24);
            s.self_state = 0;
            return SKB_OK;
        }
        SKB_TRY(s.sym_F.ensure(order.size() * (size_t)block * 24 + 16));
        std::vector<int> owned;
        for (int I = 0; I < (int)nb; ++I)
            if (sym_row_owner(I, d.sym_parts) == d.sym_part)
                owned.push_back(I);
        SKB_TRY(s.sym_row_ids.ensure(owned.size() * sizeof(int) + 16)); // (block rows of P's rows, for sym_reduce_kernel)
        CUDA_TRY(cudaMemcpyAsync(s.sym_row_ids.ptr, owned.data(), owned.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        s.sym_owned = (int)owned.size();
        CUDA_TRY(cudaMemcpyAsync(s.sym_item_buf.ptr, order.data(), order.size() * sizeof(SymItem),
                                 cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(s.sym_row_begin.ptr, row_begin.data(), row_begin.size() * sizeof(int),
                                 cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st)); // host vectors go out of scope
        // pairs of one launch: every ordered pair of real nodes in (owned row I, any block J > I), both directions
        long long pairs = 0;
        for (long long I = 0; I < nb; ++I) {
            if (sym_row_owner((int)I, d.sym_parts) != d.sym_part)
                continue;
            const long long nI = std::max<long long>(0, std::min<long long>(s.n, (I + 1) * block) - I * block);
            const long long after = std::max<long long>(0, s.n - (I + 1) * block);
            pairs += 2 * nI * after + nI * nI; // both directions beyond the block + the block diagonal
        }
        s.sym_pairs = pairs;
        s.sym_T = T;
        s.sym_nb = (int)nb;
        s.sym_items = (int)order.size();
        s.sym_part = d.sym_part;
        s.sym_parts = d.sym_parts;
        s.sym_plan_valid = true;
    }
    *use = 1;
    return SKB_OK;
}

// u[0 : 3 n_src] (=|+=) scale * sum over the square block, using the plan above.  Strengths are packed already.
static int sym_eval(skb_ctx *ctx, DeviceState &d, double *d_u_out, int accumulate, cudaStream_t st, double scale_mul,
                    int *launches) {
    SourceSet &s = d.src[SKB_STOKESLET];
    const double *f_packed = s.f_cur;
    if (!d.aux_stream)
        CUDA_TRY(cudaStreamCreateWithFlags(&d.aux_stream, cudaStreamNonBlocking));
    const int T = s.sym_T;
    const long long block = (long long)kSymThreads * T;
    SymArgs a;
    a.r = (const double *)s.r.ptr;
    a.f = f_packed;
    a.items = (const SymItem *)s.sym_item_buf.ptr;
    a.P = (double *)s.sym_P.ptr;
    a.F = (double *)s.sym_F.ptr;
    a.n_pad = s.n_pad;
    a.nb = s.sym_nb;
    cudaError_t e = cudaSuccess;
    if (s.sym_items > 0) {
        a.fid = s.excl ? (const int *)s.excl_ids.ptr : nullptr;
        CUDA_TRY(cudaEventRecord(d.ev_s0, st));
        e = s.excl ? launch_sym<kSymT, kSymMinB, true>(a, s.sym_items, st)
                   : launch_sym<kSymT, kSymMinB, false>(a, s.sym_items, st);
        if (e == cudaSuccess)
            CUDA_TRY(cudaEventRecord(d.ev_s1, st));
        d.sym_timed = true;
        if (e != cudaSuccess)
            return set_error(SKB_ERR_CUDA, "pair_sym_kernel launch failed: %s", cudaGetErrorString(e));
        count_launch(1);
    }
    const long long n3 = 3 * s.n;
    const double scale = scale_mul / (8.0 * M_PI);
    sym_reduce_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, st>>>(
        (const double *)s.sym_P.ptr, (const double *)s.sym_F.ptr, (const int *)s.sym_row_begin.ptr,
        (const int *)s.sym_row_ids.ptr, s.sym_owned, (int)block, s.n_pad, n3, scale, accumulate, d_u_out);
    e = cudaGetLastError();
    if (e != cudaSuccess)
        return set_error(SKB_ERR_CUDA, "sym_reduce_kernel launch failed: %s", cudaGetErrorString(e));
    count_launch(1);
    if (launches)
        *launches += 2;
    (void)ctx;
    return SKB_OK;
}

} // namespace skb

namespace skb {

long long cross_block_nodes() { return kCrossBlock; }

// Work items of the cross kernel: rectangles (fiber block I, periphery groups [g0, g1)), every block meets every group;
// guided sizes as for the symmetric kernel (large items first, a fine tail).
static void build_cross_items(int n_blocks, int n_groups, int num_sms, std::vector<SymItem> &order,
                              std::vector<int> &row_begin) {
    const int sg = kSymStageGroups;
    const long long work = (long long)n_blocks * n_groups;
    const long long slots = (long long)num_sms * kCrossMinB;
    long long big = work / std::max<long long>(1, slots * 4);
    big = std::max<long long>(sg, std::min<long long>(big / sg * sg, 1024LL * sg));
    const long long mid = std::max<long long>(sg, big / 4 / sg * sg);
    struct Piece {
        int I, g0, g1;
    };
    std::vector<Piece> pieces;
    for (int I = 0; I < n_blocks; ++I) {
        const int n_chunks = (int)((n_groups + big - 1) / big);
        for (int c = 0; c < n_chunks; ++c) {
            const int a = (int)((long long)(n_groups / sg) * c / n_chunks) * sg;
            const int b = c + 1 == n_chunks ? n_groups : (int)((long long)(n_groups / sg) * (c + 1) / n_chunks) * sg;
            if (b > a)
                pieces.push_back(Piece{I, a, b});
        }
    }
    std::stable_sort(pieces.begin(), pieces.end(),
                     [](const Piece &x, const Piece &y) { return (x.g1 - x.g0) > (y.g1 - y.g0); });
    std::vector<Piece> fine;
    long long done = 0;
    const long long fine_work = std::min<long long>(work * 8 / 100, slots * 6 * sg);
    const long long mid_work = std::min<long long>(work * 30 / 100, fine_work + slots * 6 * mid);
    for (const Piece &pc : pieces) {
        const long long len = pc.g1 - pc.g0, left = work - done;
        const long long unit = left <= fine_work ? sg : left <= mid_work ? mid : len;
        for (long long g = pc.g0; g < pc.g1; g += unit)
            fine.push_back(Piece{pc.I, (int)g, (int)std::min<long long>(pc.g1, g + unit)});
        done += len;
    }
    std::sort(fine.begin(), fine.end(),
              [](const Piece &x, const Piece &y) { return x.I != y.I ? x.I < y.I : x.g0 < y.g0; });
    std::vector<SymItem> items(fine.size());
    row_begin.assign(n_blocks + 1, 0);
    for (size_t i = 0; i < fine.size(); ++i) {
        items[i] = SymItem{fine[i].I, fine[i].g0, fine[i].g1, (int)i, fine[i].I, 0};
        row_begin[fine[i].I + 1] = (int)i + 1;
    }
    for (int b = 0; b < n_blocks; ++b)
        row_begin[b + 1] = std::max(row_begin[b + 1], row_begin[b]);
    order = items;
    std::stable_sort(order.begin(), order.end(),
                     [](const SymItem &x, const SymItem &y) { return (x.g1 - x.g0) > (y.g1 - y.g0); });
}

int cross_eval(CrossState &cs, const DeviceInfo &di, const double *d_r_fib, const double *d_h, long long node0,
               long long n_rows, const double *d_r_sh, const double *d_s6, long long n_sh, long long n_sh_pad,
               double scale_dl, double scale_sl, double *d_u_fib, int acc_fib, double *d_u_shell, int acc_shell,
               cudaStream_t st, int *launches) {
    if (n_rows <= 0 || n_sh <= 0)
        return SKB_OK;
    const long long block = kCrossBlock;
    const int n_blocks = (int)((n_rows + block - 1) / block);
    const int n_groups = (int)((n_sh + kSymGroup - 1) / kSymGroup);
    if (!cs.valid || cs.node0 != node0 || cs.n_rows != n_rows || cs.n_sh != n_sh || cs.n_sh_pad != n_sh_pad ||
        cs.num_sms != di.num_sms) {
        std::vector<SymItem> order;
        std::vector<int> row_begin;
        build_cross_items(n_blocks, n_groups, di.num_sms, order, row_begin);
        SKB_TRY(cs.items.ensure(order.size() * sizeof(SymItem) + 16));
        SKB_TRY(cs.row_begin.ensure(row_begin.size() * sizeof(int)));
        SKB_TRY(cs.P.ensure((size_t)n_blocks * (size_t)n_sh_pad * 24));
        SKB_TRY(cs.F.ensure(order.size() * (size_t)block * 24 + 16));
        CUDA_TRY(cudaMemcpyAsync(cs.items.ptr, order.data(), order.size() * sizeof(SymItem), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(cs.row_begin.ptr, row_begin.data(), row_begin.size() * sizeof(int),
                                 cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st)); // host vectors go out of scope
        cs.node0 = node0, cs.n_rows = n_rows, cs.n_sh = n_sh, cs.n_sh_pad = n_sh_pad;
        cs.n_blocks = n_blocks, cs.n_items = (int)order.size(), cs.num_sms = di.num_sms;
        cs.valid = true;
    }
    CrossArgs a;
    a.r_fib = d_r_fib;
    a.h = d_h;
    a.r_sh = d_r_sh;
    a.s6 = d_s6;
    a.items = (const SymItem *)cs.items.ptr;
    a.P = (double *)cs.P.ptr;
    a.F = (double *)cs.F.ptr;
    a.n_sh_pad = n_sh_pad;
    a.node0 = node0;
    a.n_rows = n_rows;
    using L = CrossSmem<kCrossT>;
    auto kern = pair_cross_kernel<kCrossT, kCrossMinB>;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total_bytes));
#if SKB_CARVEOUT_MAX
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
#endif
        attr_set[dev] = true;
    }
    kern<<<cs.n_items, kSymThreads, L::total_bytes, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    const long long nr3 = 3 * n_rows, ns3 = 3 * n_sh;
    cross_reduce_fib_kernel<<<(unsigned)((nr3 + 255) / 256), 256, 0, st>>>(
        (const double *)cs.F.ptr, (const int *)cs.row_begin.ptr, (int)block, nr3, scale_dl, acc_fib, d_u_fib);
    cross_reduce_shell_kernel<<<(unsigned)((ns3 + 255) / 256), 256, 0, st>>>((const double *)cs.P.ptr, n_blocks,
                                                                              n_sh_pad, ns3, scale_sl, acc_shell,
                                                                              d_u_shell);
    CUDA_TRY(cudaGetLastError());
    count_launch(3);
    if (launches)
        *launches += 3;
    return SKB_OK;
}

} // namespace skb

namespace skb {
__global__ void excl_target_ids_kernel(const int *__restrict__ src_ids, long long n_src, long long trg_begin,
                                       long long n_trg, int *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_trg) {
        const long long g = trg_begin + i; // global target row: the leading n_src rows are the sources themselves
        out[i] = g < n_src ? src_ids[g] : -1;
    }
}

int pack_on_device(DeviceState &d, int kind, StrengthMode mode, const double *d_f_raw, double two_eta, cudaStream_t st,
                   int *launches) {
    SourceSet &s = d.src[kind];
    const int bs = 256;
    s.f_cur = (const double *)s.f_packed.ptr;
    if (s.n <= 0)
        return SKB_OK;
    if (kind == SKB_STOKESLET) {
        pack_sl_kernel<<<(unsigned)((s.n_pad * 3 + bs - 1) / bs), bs, 0, st>>>(
            d_f_raw, s.has_weights ? (const double *)s.weights.ptr : nullptr, (double *)s.f_packed.ptr, s.n, s.n_pad);
    } else if (mode == kRaw) {
        pack_dl9_kernel<<<(unsigned)((s.n_pad + bs - 1) / bs), bs, 0, st>>>(d_f_raw, (double *)s.f_packed.ptr, s.n,
                                                                             s.n_pad);
    } else {
        pack_dl_normal_density_kernel<<<(unsigned)((s.n_pad + bs - 1) / bs), bs, 0, st>>>(
            (const double *)s.normals.ptr, d_f_raw, two_eta, (double *)s.f_packed.ptr, s.n, s.n_pad);
    }
    CUDA_TRY(cudaGetLastError());
    count_launch(1);
    if (launches)
        *launches += 1;
    return SKB_OK;
}

int eval_on_device(skb_ctx *ctx, DeviceState &d, int kind, StrengthMode mode, const double *d_f_raw, double two_eta,
                   double *d_u_out, int accumulate, cudaStream_t st, bool record_events, int *launches,
                   LaunchPlan *plan_out, double scale_mul, const EvalOpts &opts) {
    SourceSet &s = d.src[kind];
    const int bs = 256;
    if (d.n_trg == 0)
        return SKB_OK;
    if (s.n == 0) {
        if (!accumulate)
            CUDA_TRY(cudaMemsetAsync(d_u_out, 0, (size_t)d.n_trg * 24, st));
        return SKB_OK;
    }
    // 1. strengths -> packed, padded layout
    if (mode == kPacked)
        s.f_cur = d_f_raw;
    else
        SKB_TRY(pack_on_device(d, kind, mode, d_f_raw, two_eta, st, launches));
    if (record_events)
        CUDA_TRY(cudaEventRecord(d.ev_k0, st));
    // fused self-exclusion (opt-in): ids only make sense when the sources are the leading targets
    const bool excl = kind == SKB_STOKESLET && s.excl;
    if (excl && ctx->devs.size() == 1) { // (multi-device contexts check their host copies of the positions, eval_host)
        SKB_TRY(ensure_self_state(d, s, st));
        if (s.self_state != 1)
            return set_error(SKB_ERR_STATE, "exclusion ids are set, but the targets do not start with the Stokeslet "
                                            "sources (skb_set_source_exclusion_ids)");
    }
    // 2. pair sums
    // ---- symmetric path: sources are the leading targets (fiber -> fiber block of apply_matvec) ----
    long long n_sym = 0;
    if (kind == SKB_STOKESLET && (ctx->devs.size() == 1 || ctx->sym_layout) && ctx->sym_mode != 0) {
        int use = 0;
        SKB_TRY(sym_prepare(ctx, d, st, &use));
        if (use)
            n_sym = s.n;
    }
    const long long n_trg_std = d.n_trg - n_sym; // targets beyond the square block go through the plain kernel
    const double *d_r_trg_std = (const double *)d.r_trg.ptr + 3 * n_sym;
    double *d_u_std = (n_sym > 0 && opts.d_u_rem) ? opts.d_u_rem : d_u_out + 3 * n_sym;
    const double scale = scale_mul * (kind == SKB_STOKESLET ? 1.0 : -3.0) / (8.0 * M_PI);
    cudaStream_t st_std = st;
    if (n_sym > 0 && !d.aux_stream)
        CUDA_TRY(cudaStreamCreateWithFlags(&d.aux_stream, cudaStreamNonBlocking));
    if (n_sym > 0 && n_trg_std > 0) { // fork: the remainder fills the SMs the symmetric kernel's tail leaves idle
        CUDA_TRY(cudaEventRecord(d.ev_fork, st));
        CUDA_TRY(cudaStreamWaitEvent(d.aux_stream, d.ev_fork, 0));
        st_std = d.aux_stream;
    }
    LaunchPlan plan{};
    // the long symmetric kernel is launched first so that the remainder's CTAs are dispatched into its tail
    if (n_sym > 0)
        SKB_TRY(sym_eval(ctx, d, opts.d_u_sym ? opts.d_u_sym : d_u_out,
                         opts.sym_accumulate >= 0 ? opts.sym_accumulate : accumulate, st, scale_mul, launches));
    if (n_trg_std > 0) {
        const int *d_src_ids = nullptr, *d_trg_ids = nullptr;
        if (excl && n_sym == 0) { // the whole target list goes through the plain kernel: it needs per-target ids
            if (s.excl_trg_n != d.n_trg) {
                SKB_TRY(s.excl_trg.ensure((size_t)d.n_trg * sizeof(int)));
                excl_target_ids_kernel<<<(unsigned)((d.n_trg + bs - 1) / bs), bs, 0, st_std>>>(
                    (const int *)s.excl_ids.ptr, s.n, d.trg_begin, d.n_trg, (int *)s.excl_trg.ptr);
                CUDA_TRY(cudaGetLastError());
                count_launch(1);
                s.excl_trg_n = d.n_trg;
            }
            d_src_ids = (const int *)s.excl_ids.ptr;
            d_trg_ids = (const int *)s.excl_trg.ptr;
        } // (with the symmetric path the remainder targets are not sources: nothing to exclude there)
        plan = plan_launch(d.info, kind, n_trg_std, (int)((s.n + kSrcTile - 1) / kSrcTile), ctx->force_T, ctx->force_S);
        SKB_TRY(d.partial.ensure((size_t)plan.n_splits * (size_t)n_trg_std * 24));
        SKB_TRY(launch_pair_sum(d.info, kind, (const double *)s.r.ptr, s.f_cur, s.n, s.n_pad, d_r_trg_std, n_trg_std,
                                (double *)d.partial.ptr, plan, st_std, d_src_ids, d_trg_ids));
        // 3. combine splits, scale: 1/(8 pi) (kernels.cu:59) or -3/(8 pi) (kernels.cu:26,51)
        SKB_TRY(launch_reduce((const double *)d.partial.ptr, d_u_std, n_trg_std, plan.n_splits, scale, accumulate,
                              st_std));
        if (launches)
            *launches += 2;
    }
    if (kind == SKB_STOKESLET)
        ctx->last_was_sym = n_sym > 0;
    if (n_sym > 0) {
        if (n_trg_std > 0) { // join
            CUDA_TRY(cudaEventRecord(d.ev_join, d.aux_stream));
            CUDA_TRY(cudaStreamWaitEvent(st, d.ev_join, 0));
        }
        plan.T = s.sym_T;
        plan.n_splits = 1;
        plan.grid_x = (unsigned)s.sym_items;
        plan.tiles_per_split = 0;
    }
    if (record_events)
        CUDA_TRY(cudaEventRecord(d.ev_k1, st));
    if (plan_out)
        *plan_out = plan;
    return SKB_OK;
}
} // namespace skb

// host-pointer evaluation over all devices of the context
static int eval_host(skb_ctx *ctx, int kind, StrengthMode mode, const double *f_src, double eta, double *u_trg,
                     int accumulate) {
    if (!ctx)
        return set_error(SKB_ERR_INVALID, "eval: NULL ctx");
    SKB_TRY(check_kind(kind));
    if (ctx->n_trg < 0)
        return set_error(SKB_ERR_STATE, "eval: skb_set_targets has not been called");
    const long long n_src = ctx->devs[0].src[kind].n;
    if (n_src < 0)
        return set_error(SKB_ERR_STATE, "eval: skb_set_sources(kind=%d) has not been called", kind);
    if ((n_src > 0 && !f_src) || (ctx->n_trg > 0 && !u_trg))
        return set_error(SKB_ERR_INVALID, "eval: NULL strength or output pointer");
    if (mode == kNormalDensity && n_src > 0 && !ctx->devs[0].src[kind].has_normals)
        return set_error(SKB_ERR_STATE, "eval_double_layer: skb_set_source_normals has not been called");
    SKB_TRY(update_layout(ctx));
    if (kind == SKB_STOKESLET && ctx->devs.size() > 1 && ctx->devs[0].src[kind].excl &&
        !(ctx->n_trg >= n_src && (long long)ctx->h_src[kind].size() == 3 * n_src &&
          std::memcmp(ctx->h_trg.data(), ctx->h_src[kind].data(), (size_t)n_src * 24) == 0))
        return set_error(SKB_ERR_STATE, "exclusion ids are set, but the targets do not start with the Stokeslet sources "
                                        "(skb_set_source_exclusion_ids)");
    if (n_src == 0) { // an empty class contributes nothing, whatever layout the devices hold the targets in
        if (!accumulate && ctx->n_trg > 0)
            std::memset(u_trg, 0, (size_t)ctx->n_trg * 24);
        ctx->stats = skb_eval_stats{};
        return SKB_OK;
    }
    const bool sym_layout = ctx->sym_layout && kind == SKB_STOKESLET; // (no stresslet sources exist in that layout)
    const long long n_self = sym_layout ? ctx->n_self : 0;
    const int fdim = (kind == SKB_STOKESLET || mode == kNormalDensity) ? 3 : 9;
    const long long P = (long long)ctx->devs.size();
    const long long chunk = (n_src + P - 1) / P;
    int launches = 0;
    LaunchPlan plan{};

    // stage 1: strengths to the devices.  P == 1: one H2D.  P > 1: each device receives its 1/P slice over
    // PCIe, then ONE NCCL all-gather per evaluation distributes the slices over NVLink.
    for (long long g = 0; g < P; ++g) {
        DeviceState &d = ctx->devs[g];
        CUDA_TRY(cudaSetDevice(d.info.dev));
        CUDA_TRY(cudaEventRecord(d.ev_t0, d.stream));
        if (n_src == 0)
            continue;
        SourceSet &s = d.src[kind];
        const long long b = std::min(n_src, g * chunk), e = std::min(n_src, (g + 1) * chunk);
        if (e > b)
            CUDA_TRY(cudaMemcpyAsync((double *)s.f_raw.ptr + (size_t)g * chunk * fdim, f_src + (size_t)b * fdim,
                                     (size_t)(e - b) * fdim * 8, cudaMemcpyHostToDevice, d.stream));
    }
    if (P > 1 && n_src > 0) {
        std::vector<void *> bufs(P);
        std::vector<cudaStream_t> sts(P);
        for (long long g = 0; g < P; ++g) {
            bufs[g] = ctx->devs[g].src[kind].f_raw.ptr;
            sts[g] = ctx->devs[g].stream;
        }
        SKB_TRY(nccl_group_allgather_inplace(ctx->nccl, bufs.data(), (size_t)chunk * fdim, sts.data()));
    }
    // stage 2: every device evaluates its targets against all sources
    const long long chunk_f = sym_layout ? (n_self + P - 1) / P : 0; // reduce-scatter slot (leading rows per device)
    for (long long g = 0; g < P; ++g) {
        DeviceState &d = ctx->devs[g];
        if (d.n_trg == 0)
            continue;
        CUDA_TRY(cudaSetDevice(d.info.dev));
        if (!sym_layout) {
            if (accumulate)
                CUDA_TRY(cudaMemcpyAsync(d.u.ptr, u_trg + 3 * d.trg_begin, (size_t)d.n_trg * 24,
                                         cudaMemcpyHostToDevice, d.stream));
        } else {
            // leading rows are partial sums over the devices: the caller's values enter once (device 0)
            if (chunk_f * P > n_self)
                CUDA_TRY(cudaMemsetAsync((double *)d.u.ptr + 3 * n_self, 0, (size_t)(chunk_f * P - n_self) * 24,
                                         d.stream));
            if (accumulate) {
                if (g == 0)
                    CUDA_TRY(cudaMemcpyAsync(d.u.ptr, u_trg, (size_t)n_self * 24, cudaMemcpyHostToDevice, d.stream));
                else
                    CUDA_TRY(cudaMemsetAsync(d.u.ptr, 0, (size_t)n_self * 24, d.stream));
                if (d.rem_count > 0)
                    CUDA_TRY(cudaMemcpyAsync((double *)d.u.ptr + 3 * n_self, u_trg + 3 * (n_self + d.rem_begin),
                                             (size_t)d.rem_count * 24, cudaMemcpyHostToDevice, d.stream));
            }
        }
        SKB_TRY(eval_on_device(ctx, d, kind, mode, (const double *)d.src[kind].f_raw.ptr, 2.0 * eta,
                               (double *)d.u.ptr, accumulate, d.stream, true, &launches, g == 0 ? &plan : nullptr, 1.0));
        if (!sym_layout) {
            CUDA_TRY(cudaMemcpyAsync(u_trg + 3 * d.trg_begin, d.u.ptr, (size_t)d.n_trg * 24, cudaMemcpyDeviceToHost,
                                     d.stream));
            CUDA_TRY(cudaEventRecord(d.ev_t1, d.stream));
        }
    }
    if (sym_layout) {
        // one reduce-scatter of the leading rows: device g ends up with rows [g*chunk_f, (g+1)*chunk_f)
        std::vector<void *> send(P), recv(P);
        std::vector<cudaStream_t> sts(P);
        for (long long g = 0; g < P; ++g) {
            send[g] = ctx->devs[g].u.ptr;
            recv[g] = ctx->devs[g].u_rs.ptr;
            sts[g] = ctx->devs[g].stream;
        }
        SKB_TRY(nccl_group_reduce_scatter(ctx->nccl, send.data(), recv.data(), (size_t)chunk_f * 3, sts.data()));
        for (long long g = 0; g < P; ++g) {
            DeviceState &d = ctx->devs[g];
            CUDA_TRY(cudaSetDevice(d.info.dev));
            const long long b = std::min(n_self, g * chunk_f), e = std::min(n_self, (g + 1) * chunk_f);
            if (e > b)
                CUDA_TRY(cudaMemcpyAsync(u_trg + 3 * b, d.u_rs.ptr, (size_t)(e - b) * 24, cudaMemcpyDeviceToHost,
                                         d.stream));
            if (d.rem_count > 0)
                CUDA_TRY(cudaMemcpyAsync(u_trg + 3 * (n_self + d.rem_begin), (double *)d.u.ptr + 3 * n_self,
                                         (size_t)d.rem_count * 24, cudaMemcpyDeviceToHost, d.stream));
            CUDA_TRY(cudaEventRecord(d.ev_t1, d.stream));
        }
    }
    // stage 3: wait, collect timings
    double k_ms = 0, t_ms = 0;
    for (long long g = 0; g < P; ++g) {
        DeviceState &d = ctx->devs[g];
        CUDA_TRY(cudaSetDevice(d.info.dev));
        CUDA_TRY(cudaStreamSynchronize(d.stream));
        if (d.n_trg == 0)
            continue;
        float ms = 0;
        if (n_src > 0 && cudaEventElapsedTime(&ms, d.ev_k0, d.ev_k1) == cudaSuccess)
            k_ms = std::max(k_ms, (double)ms);
        if (cudaEventElapsedTime(&ms, d.ev_t0, d.ev_t1) == cudaSuccess)
            t_ms = std::max(t_ms, (double)ms);
    }
    ctx->kernel_events_pending = false;
    ctx->stats.kernel_ms = k_ms;
    ctx->stats.total_ms = t_ms;
    ctx->stats.n_pairs = n_src * ctx->n_trg;
    ctx->stats.launches = launches;
    ctx->stats.targets_per_thread = plan.T;
    ctx->stats.source_splits = plan.n_splits;
    ctx->stats.grid_ctas = (int)(plan.grid_x * plan.n_splits);
    return SKB_OK;
}

// process-wide context behind the stateless reference-shaped entry points
static std::mutex g_default_mu;
static skb_ctx *g_default_ctx = nullptr;
static int default_ctx(skb_ctx **out) {
    if (!g_default_ctx) {
        int rc = skb_ctx_create(1, &g_default_ctx);
        if (rc != SKB_OK)
            return rc;
    }
    *out = g_default_ctx;
    return SKB_OK;
}

static int direct_impl(int kind, const double *r_src, const double *f_src, int n_src, const double *r_trg,
                       double *u_trg, int n_trg) {
    if (n_src < 0 || n_trg < 0)
        return set_error(SKB_ERR_INVALID, "negative count");
    std::lock_guard<std::mutex> lock(g_default_mu);
    skb_ctx *ctx = nullptr;
    SKB_TRY(default_ctx(&ctx));
    SKB_TRY(skb_set_sources(ctx, kind, r_src, n_src));
    SKB_TRY(skb_set_targets(ctx, r_trg, n_trg));
    return skb_eval(ctx, kind, f_src, u_trg, 0);
}

extern "C" {

int skb_set_targets(skb_ctx *ctx, const double *r_trg, int64_t n_trg) {
    return set_targets_impl(ctx, r_trg, n_trg, false, nullptr);
}
int skb_set_targets_device(skb_ctx *ctx, const double *d_r_trg, int64_t n_trg, void *stream) {
    return set_targets_impl(ctx, d_r_trg, n_trg, true, (cudaStream_t)stream);
}
int skb_set_sources(skb_ctx *ctx, int kind, const double *r_src, int64_t n_src) {
    return set_sources_impl(ctx, kind, r_src, n_src, false, nullptr);
}
int skb_set_sources_device(skb_ctx *ctx, int kind, const double *d_r_src, int64_t n_src, void *stream) {
    return set_sources_impl(ctx, kind, d_r_src, n_src, true, (cudaStream_t)stream);
}

int skb_set_source_normals(skb_ctx *ctx, const double *normals, int64_t n_src) {
    if (!ctx)
        return set_error(SKB_ERR_INVALID, "set_source_normals: NULL ctx");
    const long long n = ctx->devs[0].src[SKB_STRESSLET].n;
    if (n < 0)
        return set_error(SKB_ERR_STATE, "set_source_normals: call skb_set_sources(SKB_STRESSLET, ...) first");
    if (n_src != n || (n > 0 && !normals))
        return set_error(SKB_ERR_INVALID, "set_source_normals: n_src=%lld does not match the %lld stresslet sources",
                         (long long)n_src, n);
    for (auto &d : ctx->devs) {
        SourceSet &s = d.src[SKB_STRESSLET];
        s.has_normals = true;
        if (n == 0)
            continue;
        CUDA_TRY(cudaSetDevice(d.info.dev));
        SKB_TRY(s.normals.ensure((size_t)n * 24));
        CUDA_TRY(cudaMemcpyAsync(s.normals.ptr, normals, (size_t)n * 24, cudaMemcpyHostToDevice, d.stream));
        CUDA_TRY(cudaStreamSynchronize(d.stream));
    }
    return SKB_OK;
}

int skb_set_source_exclusion_ids(skb_ctx *ctx, const int32_t *ids, int64_t n_src) {
    if (!ctx)
        return set_error(SKB_ERR_INVALID, "skb_set_source_exclusion_ids: NULL ctx");
    const long long n = ctx->devs[0].src[SKB_STOKESLET].n;
    if (n < 0)
        return set_error(SKB_ERR_STATE, "skb_set_source_exclusion_ids: call skb_set_sources(SKB_STOKESLET, ...) first");
    if (ids && n_src != n)
        return set_error(SKB_ERR_INVALID, "skb_set_source_exclusion_ids: n_src=%lld does not match the %lld Stokeslet "
                                          "sources", (long long)n_src, n);
    for (auto &d : ctx->devs) {
        SourceSet &s = d.src[SKB_STOKESLET];
        s.excl = false;
        s.excl_trg_n = -1;
        if (!ids || n == 0)
            continue;
        CUDA_TRY(cudaSetDevice(d.info.dev));
        SKB_TRY(s.excl_ids.ensure((size_t)s.n_pad * sizeof(int)));
        CUDA_TRY(cudaMemsetAsync(s.excl_ids.ptr, 0xff, (size_t)s.n_pad * sizeof(int), d.stream)); // pads: id -1
        CUDA_TRY(cudaMemcpyAsync(s.excl_ids.ptr, ids, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, d.stream));
        CUDA_TRY(cudaStreamSynchronize(d.stream));
        s.excl = true;
    }
    return SKB_OK;
}

int skb_eval(skb_ctx *ctx, int kind, const double *f_src, double *u_trg, int accumulate) {
    return eval_host(ctx, kind, kRaw, f_src, 0.0, u_trg, accumulate);
}

int skb_eval_double_layer(skb_ctx *ctx, const double *density, double eta, double *u_trg, int accumulate) {
    return eval_host(ctx, SKB_STRESSLET, kNormalDensity, density, eta, u_trg, accumulate);
}

int skb_eval_fused(skb_ctx *ctx, const double *f_sl, const double *f_dl, double *u_trg) {
    if (!ctx)
        return set_error(SKB_ERR_INVALID, "eval_fused: NULL ctx");
    int acc = 0;
    skb_eval_stats total{};
    if (f_sl) {
        SKB_TRY(eval_host(ctx, SKB_STOKESLET, kRaw, f_sl, 0.0, u_trg, acc));
        acc = 1;
        total = ctx->stats;
    }
    if (f_dl) {
        SKB_TRY(eval_host(ctx, SKB_STRESSLET, kRaw, f_dl, 0.0, u_trg, acc));
        acc = 1;
        total.kernel_ms += ctx->stats.kernel_ms;
        total.total_ms += ctx->stats.total_ms;
        total.n_pairs += ctx->stats.n_pairs;
        total.launches += ctx->stats.launches;
        total.targets_per_thread = ctx->stats.targets_per_thread;
        total.source_splits = ctx->stats.source_splits;
        total.grid_ctas = ctx->stats.grid_ctas;
        ctx->stats = total;
    }
    if (!acc && ctx->n_trg > 0) {
        if (!u_trg)
            return set_error(SKB_ERR_INVALID, "eval_fused: NULL output");
        std::memset(u_trg, 0, (size_t)ctx->n_trg * 24);
    }
    return SKB_OK;
}

int skb_eval_device(skb_ctx *ctx, int kind, const double *d_f_src, double *d_u_trg, int accumulate, void *stream) {
    if (!ctx)
        return set_error(SKB_ERR_INVALID, "eval_device: NULL ctx");
    SKB_TRY(check_kind(kind));
    if (ctx->devs.size() != 1)
        return set_error(SKB_ERR_INVALID, "device-pointer entry points need a single-GPU context");
    DeviceState &d = ctx->devs[0];
    if (ctx->n_trg < 0 || d.src[kind].n < 0)
        return set_error(SKB_ERR_STATE, "eval_device: targets / sources not set");
    if ((d.src[kind].n > 0 && !d_f_src) || (ctx->n_trg > 0 && !d_u_trg))
        return set_error(SKB_ERR_INVALID, "eval_device: NULL pointer");
    CUDA_TRY(cudaSetDevice(d.info.dev));
    cudaStream_t st = (cudaStream_t)stream; // verbatim: 0 is CUDA's legacy default stream
    int launches = 0;
    LaunchPlan plan{};
    SKB_TRY(eval_on_device(ctx, d, kind, kRaw, d_f_src, 0.0, d_u_trg, accumulate, st, true, &launches, &plan, 1.0));
    ctx->kernel_events_pending = d.src[kind].n > 0 && d.n_trg > 0;
    ctx->stats.kernel_ms = 0;
    ctx->stats.total_ms = 0;
    ctx->stats.n_pairs = d.src[kind].n * ctx->n_trg;
    ctx->stats.launches = launches;
    ctx->stats.targets_per_thread = plan.T;
    ctx->stats.source_splits = plan.n_splits;
    ctx->stats.grid_ctas = (int)(plan.grid_x * plan.n_splits);
    return SKB_OK;
}

int skb_sync(skb_ctx *ctx) {
    if (!ctx)
        return set_error(SKB_ERR_INVALID, "skb_sync: NULL ctx");
    for (auto &d : ctx->devs) {
        CUDA_TRY(cudaSetDevice(d.info.dev));
        CUDA_TRY(cudaStreamSynchronize(d.stream));
    }
    return SKB_OK;
}

int skb_stokeslet_direct(const double *r_src, const double *f_src, int n_src, const double *r_trg, double *u_trg,
                         int n_trg) {
    return direct_impl(SKB_STOKESLET, r_src, f_src, n_src, r_trg, u_trg, n_trg);
}
int skb_stresslet_direct(const double *r_src, const double *f_src, int n_src, const double *r_trg, double *u_trg,
                         int n_trg) {
    return direct_impl(SKB_STRESSLET, r_src, f_src, n_src, r_trg, u_trg, n_trg);
}

int skb_measure_fp64_peak(skb_ctx *ctx, double *flops_per_s) {
    if (!ctx || !flops_per_s)
        return set_error(SKB_ERR_INVALID, "skb_measure_fp64_peak: NULL");
    DeviceState &d = ctx->devs[0];
    CUDA_TRY(cudaSetDevice(d.info.dev));
    const int blocks = d.info.num_sms * 8, threads = 256, iters = 1 << 15;
    SKB_TRY(d.scratch.ensure((size_t)blocks * threads * 8));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(d.ev_k0, d.stream));
        dfma_probe_kernel<<<blocks, threads, 0, d.stream>>>((double *)d.scratch.ptr, iters, 0.999999, 1e-9);
        CUDA_TRY(cudaGetLastError());
        count_launch(1);
        CUDA_TRY(cudaEventRecord(d.ev_k1, d.stream));
        CUDA_TRY(cudaStreamSynchronize(d.stream));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, d.ev_k0, d.ev_k1));
        const double fl = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3);
        if (rep > 0)
            best = std::max(best, fl);
    }
    *flops_per_s = best;
    return SKB_OK;
}

} // extern "C"

// ------------------------------------------------------------------------------------------------
// Reference-named C++ entry points (SkellySim include/kernels.hpp:17-20).  Linking this library in
// place of src/core/kernels.cu keeps kernels::stokeslet_direct_gpu / stresslet_direct_gpu
// (src/core/kernels.cpp:354-366) and `pair_evaluator = "GPU"` working unchanged.  Error behaviour
// follows the reference's CUDA path: message on stderr, exit(EXIT_FAILURE) (kernels.cu:8-15).
// ------------------------------------------------------------------------------------------------
namespace kernels {
__attribute__((visibility("default"))) void stokeslet_direct_gpu_impl(const double *r_src, const double *f_src,
                                                                      int n_src, const double *r_trg, double *u_trg,
                                                                      int n_trg) {
    if (skb_stokeslet_direct(r_src, f_src, n_src, r_trg, u_trg, n_trg) != SKB_OK) {
        fprintf(stderr, "skelly_b200: stokeslet_direct_gpu_impl: %s\n", skb_last_error_string());
        exit(EXIT_FAILURE);
    }
}
__attribute__((visibility("default"))) void stresslet_direct_gpu_impl(const double *r_src, const double *f_src,
                                                                      int n_src, const double *r_trg, double *u_trg,
                                                                      int n_trg) {
    if (skb_stresslet_direct(r_src, f_src, n_src, r_trg, u_trg, n_trg) != SKB_OK) {
        fprintf(stderr, "skelly_b200: stresslet_direct_gpu_impl: %s\n", skb_last_error_string());
        exit(EXIT_FAILURE);
    }
}
} // namespace kernels
