// sym_kernels.cuh -- Stokeslet self-interaction with Newton's third law (sm_100a).
//
// When the sources ARE the first n_src targets -- the fiber->fiber block that is ~90 % of the pairs of
// System::apply_matvec (r_all starts with the fiber nodes, src/core/system.cpp:284-299) -- the Oseen tensor is
// symmetric in the pair: G(x_i - x_j) = G(x_j - x_i).  d, r^2, 1/r and 1/r^2 (12 of the 22 FP64 instructions of a
// pair) can then serve both directions: 32 instructions per pair-of-pairs instead of 44, and under the measured
// register-file read model (profiles/r1_fp64_ubench.md) 82 instead of 107 FP64-pipe cycles.
//
// Layout of the work: nodes are cut into blocks of kSymThreads*T; a work item is (target block I, source blocks
// [J0, J1)), all J > I (strict upper triangle; the block diagonal and the non-square remainder go through
// pair_sum_kernel).  Each thread keeps T nodes of I in registers: position, strength and the forward accumulators
// (u_I += G f_J), for the whole item.  The J block is staged in shared memory by TMA; within a warp the 32 lanes walk
// a group of 32 J-nodes as a ring: at step s lane l meets node (l + s) mod 32, and that node's reverse accumulator
// (u_J += G f_I) travels with it from lane to lane by one warp shuffle per step, so after 32 steps it is back home
// holding the sum over the warp's 32*T targets -- no cross-lane reduction tree.  The four warps visit the 16 groups in a
// staggered order and add their sums into one shared-memory slab in a fixed order (bitwise reproducible); the slab is
// the (I -> J) reverse partial written once to P[I][J-nodes].  The final combination is a fixed-order sum.
#pragma once
#include "pair_kernels.cuh"

namespace skb {

constexpr int kSymThreads = 128;
constexpr int kSymStages = 2;

struct SymItem {
    int I, J0, J1, slot; // target block, source blocks [J0, J1) (all > I), forward-partial slab index
    int prow, pad;       // row of P this item writes: index of I among the block rows owned by this part
};

struct SymArgs {
    const double *r;      // [n_pad*3] node positions, padded to a multiple of the block (pads replicate the last node)
    const double *f;      // [n_pad*3] packed Stokeslet strengths, zero padded
    const SymItem *items; // [gridDim.x]
    double *P;            // [owned rows][n_pad*3]  reverse partials: P[prow(I)][node of J] = sum over targets in I
    double *F;            // [n_items][block*3] forward partials of each item
    long long n_pad;
    int nb;
};

template <int T> struct SymSmem {
    static constexpr int block = kSymThreads * T;
    static constexpr int stage_bytes = block * 48; // positions + strengths of one J block
    static constexpr int rev_bytes = block * 24;
    static constexpr int bar_offset = kSymStages * stage_bytes + rev_bytes;
    static constexpr int total_bytes = bar_offset + 2 * kSymStages * 8;
};

// T pair-of-pairs: targets (tx,ty,tz | strength hx,hy,hz) against ONE record (rx,ry,rz | strength gx,gy,gz).
// forward:  uf_t += y (g + d (g.d) y^2)      reverse:  ur += y (h_t + d (h_t.d) y^2)       d = x_t - x_rec
// (the reverse displacement is -d; the two sign changes cancel).  32 FP64 instructions per pair-of-pairs.
template <int T>
__device__ __forceinline__ void stokeslet_pairpairs(const double (&tx)[T], const double (&ty)[T],
                                                    const double (&tz)[T], const double (&hx)[T],
                                                    const double (&hy)[T], const double (&hz)[T], double rx, double ry,
                                                    double rz, double gx, double gy, double gz, double (&ufx)[T],
                                                    double (&ufy)[T], double (&ufz)[T], double &urx, double &ury,
                                                    double &urz) {
    double dx[T], dy[T], dz[T], r2[T], y[T], q[T], fr[T], hr[T];
#pragma unroll
    for (int c = 0; c < T; ++c)
        dx[c] = tx[c] - rx;
#pragma unroll
    for (int c = 0; c < T; ++c)
        dy[c] = ty[c] - ry;
#pragma unroll
    for (int c = 0; c < T; ++c)
        dz[c] = tz[c] - rz;
#pragma unroll
    for (int c = 0; c < T; ++c)
        r2[c] = dx[c] * dx[c];
#pragma unroll
    for (int c = 0; c < T; ++c)
        r2[c] = fma(dy[c], dy[c], r2[c]);
#pragma unroll
    for (int c = 0; c < T; ++c)
        r2[c] = fma(dz[c], dz[c], r2[c]);
#pragma unroll
    for (int c = 0; c < T; ++c) {
        double y0;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(r2[c]));
        if (__double2hiint(r2[c]) < 0x00100000)
            y0 = 0.0;
        y[c] = y0;
    }
#pragma unroll
    for (int c = 0; c < T; ++c)
        fr[c] = gx * dx[c];
#pragma unroll
    for (int c = 0; c < T; ++c)
        hr[c] = hx[c] * dx[c];
#pragma unroll
    for (int c = 0; c < T; ++c)
        fr[c] = fma(gy, dy[c], fr[c]);
#pragma unroll
    for (int c = 0; c < T; ++c)
        hr[c] = fma(hy[c], dy[c], hr[c]);
#pragma unroll
    for (int c = 0; c < T; ++c)
        fr[c] = fma(gz, dz[c], fr[c]);
#pragma unroll
    for (int c = 0; c < T; ++c)
        hr[c] = fma(hz[c], dz[c], hr[c]);
#pragma unroll
    for (int c = 0; c < T; ++c)
        r2[c] = r2[c] * y[c];
#pragma unroll
    for (int c = 0; c < T; ++c)
        r2[c] = fma(-r2[c], y[c], 1.0);
#pragma unroll
    for (int c = 0; c < T; ++c)
        q[c] = fma(0.375, r2[c], 0.5);
#pragma unroll
    for (int c = 0; c < T; ++c)
        q[c] = r2[c] * q[c];
#pragma unroll
    for (int c = 0; c < T; ++c)
        y[c] = fma(y[c], q[c], y[c]); // 1/|d|
#pragma unroll
    for (int c = 0; c < T; ++c)
        q[c] = y[c] * y[c];
#pragma unroll
    for (int c = 0; c < T; ++c)
        fr[c] = fr[c] * q[c];
#pragma unroll
    for (int c = 0; c < T; ++c)
        hr[c] = hr[c] * q[c];
#pragma unroll
    for (int c = 0; c < T; ++c) {
        ufx[c] = fma(y[c], fma(dx[c], fr[c], gx), ufx[c]);
        ufy[c] = fma(y[c], fma(dy[c], fr[c], gy), ufy[c]);
        ufz[c] = fma(y[c], fma(dz[c], fr[c], gz), ufz[c]);
    }
#pragma unroll
    for (int c = 0; c < T; ++c) {
        urx = fma(y[c], fma(dx[c], hr[c], hx[c]), urx);
        ury = fma(y[c], fma(dy[c], hr[c], hy[c]), ury);
        urz = fma(y[c], fma(dz[c], hr[c], hz[c]), urz);
    }
}

template <int T, int MINB>
__global__ void __launch_bounds__(kSymThreads, MINB) pair_sym_kernel(const SymArgs a) {
    using L = SymSmem<T>;
    constexpr int kBlock = L::block;
    constexpr int kGroups = kBlock / 32;
    extern __shared__ __align__(128) unsigned char smem[];
    double *rev = reinterpret_cast<double *>(smem + kSymStages * L::stage_bytes);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + L::bar_offset);
    uint64_t *empty_bar = full_bar + kSymStages;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const SymItem item = a.items[blockIdx.x];
    const int nJ = item.J1 - item.J0;

    auto issue_block = [&](int k) { // k-th J block of the item -> ring slot k % kSymStages
        const int s = k % kSymStages;
        unsigned char *dst = smem + s * L::stage_bytes;
        const size_t off = (size_t)(item.J0 + k) * kBlock * 24;
        mbar_arrive_expect_tx(&full_bar[s], L::stage_bytes);
        tma_bulk_g2s(dst, reinterpret_cast<const char *>(a.r) + off, kBlock * 24, &full_bar[s]);
        tma_bulk_g2s(dst + kBlock * 24, reinterpret_cast<const char *>(a.f) + off, kBlock * 24, &full_bar[s]);
    };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kSymStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kSymThreads / 32);
        }
        mbar_fence_init();
        issue_block(0);
    }
    // this thread's T nodes of block I: position, strength, forward accumulators
    double tx[T], ty[T], tz[T], hx[T], hy[T], hz[T], ufx[T], ufy[T], ufz[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const size_t i = (size_t)item.I * kBlock + t * kSymThreads + tid;
        tx[t] = __ldg(a.r + 3 * i + 0), ty[t] = __ldg(a.r + 3 * i + 1), tz[t] = __ldg(a.r + 3 * i + 2);
        hx[t] = __ldg(a.f + 3 * i + 0), hy[t] = __ldg(a.f + 3 * i + 1), hz[t] = __ldg(a.f + 3 * i + 2);
        ufx[t] = ufy[t] = ufz[t] = 0.0;
    }
    __syncthreads();

    for (int k = 0; k < nJ; ++k) {
        const int s = k % kSymStages;
        if (tid == 0 && k + 1 < nJ) { // prefetch the next J block into the other slot
            if (k + 1 >= kSymStages)
                mbar_wait(&empty_bar[(k + 1) % kSymStages], (((k + 1) / kSymStages) - 1) & 1);
            issue_block(k + 1);
        }
        for (int i = tid; i < kBlock * 3; i += kSymThreads)
            rev[i] = 0.0;
        mbar_wait(&full_bar[s], (k / kSymStages) & 1);
        __syncthreads();
        const double *ps = reinterpret_cast<const double *>(smem + s * L::stage_bytes);
        const double *fs = ps + kBlock * 3;
#pragma unroll 1
        for (int it = 0; it < kGroups; ++it) {
            // staggered so the 4 warps are always on 4 different groups; every warp visits every group once
            const int g = (it + warp * (kGroups / 4)) % kGroups;
            const int base = g * 32;
            double urx = 0.0, ury = 0.0, urz = 0.0;
#pragma unroll 1
            for (int st = 0; st < 32; ++st) {
                const int idx = base + ((lane + st) & 31);
                // per-lane record from shared memory: 24 B stride is bank-conflict free for 64-bit loads
                const double rx = ps[3 * idx + 0], ry = ps[3 * idx + 1], rz = ps[3 * idx + 2];
                const double gx = fs[3 * idx + 0], gy = fs[3 * idx + 1], gz = fs[3 * idx + 2];
                stokeslet_pairpairs<T>(tx, ty, tz, hx, hy, hz, rx, ry, rz, gx, gy, gz, ufx, ufy, ufz, urx, ury, urz);
                // the record moves to lane - 1 for the next step; its accumulator goes with it
                const int from = (lane + 1) & 31;
                urx = __shfl_sync(0xffffffffu, urx, from);
                ury = __shfl_sync(0xffffffffu, ury, from);
                urz = __shfl_sync(0xffffffffu, urz, from);
            }
            // after 32 steps lane l holds the finished sum of node base + l over this warp's 32*T targets
            rev[3 * (base + lane) + 0] += urx;
            rev[3 * (base + lane) + 1] += ury;
            rev[3 * (base + lane) + 2] += urz;
            __syncthreads(); // fixed order of the 4 warps' additions per group -> reproducible
        }
        // reverse partial of (I -> block J0+k): one coalesced write
        double *out = a.P + ((size_t)item.prow * a.n_pad + (size_t)(item.J0 + k) * kBlock) * 3;
        for (int i = tid; i < kBlock * 3; i += kSymThreads)
            out[i] = rev[i];
        __syncthreads();
        if (lane == 0)
            mbar_arrive(&empty_bar[s]);
    }
    double *fo = a.F + (size_t)item.slot * kBlock * 3;
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int i = t * kSymThreads + tid;
        fo[3 * i + 0] = ufx[t];
        fo[3 * i + 1] = ufy[t];
        fo[3 * i + 2] = ufz[t];
    }
}

// Fixed-order combination for the symmetric path, one thread per velocity component of node `n` (block b):
//   u = (acc ? u : 0) + scale * ( diag[n] + sum_{owned I < b} P[prow(I)][n] + sum_{items of row b} F[item][n - b*block] )
__global__ void sym_reduce_kernel(const double *__restrict__ diag, const double *__restrict__ P,
                                  const double *__restrict__ F, const int *__restrict__ row_item_begin, int block,
                                  long long n_pad, long long n_valid3, double scale, int accumulate,
                                  double *__restrict__ u, int part, int n_parts, int n_diag, long long n_valid) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; // component index 3*n + k
    if (i >= n_valid3)
        return;
    const long long node = i / 3;
    const int b = (int)(node / block);
    // only rows owned by this part were evaluated here (n_parts == 1: all of them)
    double acc = 0.0;
    if (sym_row_owner(b, n_parts) == part)
        for (int y = 0; y < n_diag; ++y) // diagonal block, one slab per source tile of the block
            acc += diag[(size_t)y * n_valid * 3 + i];
    int prow = 0; // P holds only the rows this part owns, in increasing I
    for (int I = 0; I < b; ++I)
        if (sym_row_owner(I, n_parts) == part) {
            acc += P[((size_t)prow * n_pad) * 3 + i];
            ++prow;
        }
    const long long local = i - (long long)b * block * 3;
    for (int it = row_item_begin[b]; it < row_item_begin[b + 1]; ++it)
        acc += F[(size_t)it * block * 3 + local];
    acc *= scale;
    u[i] = accumulate ? u[i] + acc : acc;
}

} // namespace skb
