// stream_kernels.cuh -- the BACKGROUND row streamer: y = A x for a row-major FP64 matrix, built to run BESIDE the
// FP64-bound pair kernels instead of after them.
//
// Why: the periphery's dense operator (Periphery::matvec, SkellySim src/core/periphery.cpp:38-47) reads 8 B per FMA -- it
// is bound by HBM and leaves the FP64 pipe idle; pair_sym_kernel is bound by the FP64 pipe and leaves HBM idle.  Run one
// after the other they cost 10.2 + 0.43 ms of the C3 matvec.  The classic GEMV (dense_gemv_kernel: 256 threads, all
// warps resident) cannot share an SM with the symmetric kernel, whose two CTAs held the whole register file.  This
// kernel can: ONE CTA of three warps x 32 registers per SM (a producer warp and two consumer warps).  The register file
// is split over the SM's four sub-partitions (16 384 registers each); pair_sym_kernel is capped at 240 registers, so its
// two CTAs (one warp per sub-partition each) leave 1 024 registers per sub-partition free -- exactly one 32-register
// warp each.  The matrix is moved by the TMA engine, not by threads --
//   * five lanes of the producer warp post 1-D bulk copies (cp.async.bulk, SASS UBLKCP): 4 rows x 512 columns of A plus
//     the 512 matching entries of x (L2 resident) = 20 KB per stage into a 5-stage ring, 4 stages (64 KB of A) ahead of
//     use: HBM latency x bandwidth per SM is ~45 KB, so the ring keeps the SM's share of the 6.5 TB/s in flight with
//     no warp waiting on a load instruction (a first version read x with LDG: every stage then waited one L2 latency,
//     1.8 TB/s);
//   * the 64 consumer threads consume a stage with 20 LDS.128 + 32 DFMA each -- 4 % of the SM's FP64 issue slots while
//     it runs, 0.3 % of the symmetric kernel's total;
//   * the bulk copies carry an L2 evict-first policy: 2.6 GB of read-once matrix must not evict the 4.6 MB of node
//     records the pair kernels stream from L2.
// Measured (profiles/r2_overlap.md): alone 3.84 TB/s; beside pair_sym_kernel the 2.59 GB of the C3 operator arrive in
// 0.95 ms while the pair kernel loses 0.09 ms -- the matvec 12.56 -> 12.25 ms.
// Rows are handed out in groups of 4 by a ticket counter (first group static); a row's sum is accumulated in a fixed
// thread -> column mapping and a fixed reduction tree, so results are bitwise reproducible whatever the ticket order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tma_ptx.cuh"

namespace skb {

constexpr int kStreamConsumers = 64;                 // two consumer warps
constexpr int kStreamThreads = kStreamConsumers + 32; // + the producer warp (five lanes active)
constexpr int kStreamRows = 4;     // rows per group (x is re-used 4 times per load)
#ifndef SKB_STREAM_COLS
#define SKB_STREAM_COLS 512
#endif
#ifndef SKB_STREAM_HINT
#define SKB_STREAM_HINT 1
#endif
constexpr int kStreamCols = SKB_STREAM_COLS; // columns per stage (a multiple of 128)
#ifndef SKB_STREAM_STAGES
#define SKB_STREAM_STAGES 5
#endif
constexpr int kStreamStages = SKB_STREAM_STAGES;
constexpr int kStreamRowBytes = kStreamCols * 8;
constexpr int kStreamStageBytes = (kStreamRows + 1) * kStreamRowBytes; // rows of A, then the segment of x
// ring | full + empty barriers | (group, chunk) of every stage | cross-warp sums
constexpr int kStreamBarOffset = kStreamStages * kStreamStageBytes;
constexpr int kStreamMetaOffset = kStreamBarOffset + 2 * kStreamStages * 8;
constexpr int kStreamRedOffset = kStreamMetaOffset + kStreamStages * 8;
constexpr int kStreamSmemBytes = kStreamRedOffset + 2 * kStreamRows * 8;

struct StreamArgs {
    const double *A; // [n_rows][n_cols] row-major, n_cols even (every row 16-byte aligned)
    const double *x; // [n_cols], 16-byte aligned
    double *y;       // [n_rows]
    long long n_rows, n_cols;
    unsigned long long *next_group; // ticket counter, zero at launch
};

#ifndef SKB_STREAM_MAXNREG
#define SKB_STREAM_MAXNREG 32
#endif
__global__ void __maxnreg__(SKB_STREAM_MAXNREG) dense_stream_kernel(const StreamArgs a) {
    extern __shared__ __align__(128) unsigned char st_smem[];
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(st_smem + kStreamBarOffset);
    uint64_t *empty_bar = full_bar + kStreamStages;
    int2 *meta = reinterpret_cast<int2 *>(st_smem + kStreamMetaOffset);
    double *red = reinterpret_cast<double *>(st_smem + kStreamRedOffset);
    const int t = threadIdx.x;
    const int n_chunks = (int)((a.n_cols + kStreamCols - 1) / kStreamCols);
    if (t == 0) {
        for (int s = 0; s < kStreamStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kStreamConsumers / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (t >= kStreamConsumers) {
        // ---- producer: lanes 0..3 own one row of the group each, lane 4 the segment of x; every lane keeps its own source
        // pointer, so a stage costs one short instruction sequence per lane, not five address computations in one lane
        // (that first version spent ~150 cycles per copy and was bound by it: 3 TB/s whatever the ring depth) ----
        const int lane = t - kStreamConsumers;
        if (lane > kStreamRows)
            return;
        constexpr unsigned kMask = (2u << kStreamRows) - 1u;
        const long long n_groups = (a.n_rows + kStreamRows - 1) / kStreamRows;
        const uint64_t policy = l2_policy_evict_first();
        long long group = blockIdx.x;
        int s = 0;
        uint32_t empty_parity = 1; // parity of the phase BEFORE the first use: the wait falls through on a fresh barrier
        for (;;) {
            if (group >= n_groups) { // end marker: completes on the arrive alone
                if (lane == 0) {
                    mbar_wait(&empty_bar[s], empty_parity);
                    meta[s] = make_int2(-1, 0);
                    mbar_arrive(&full_bar[s]);
                }
                return;
            }
            // the ticket of the FOLLOWING group is drawn a whole group ahead of its use
            unsigned long long ticket = 0;
            if (lane == 0)
                ticket = atomicAdd(a.next_group, 1ULL);
            const long long row0 = group * kStreamRows;
            const int rows = (int)min((long long)kStreamRows, a.n_rows - row0);
            const double *src = lane < kStreamRows ? a.A + (row0 + lane) * a.n_cols : a.x;
            const bool active = lane == kStreamRows || lane < rows;
            long long left = a.n_cols;
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(&empty_bar[s], empty_parity); // the consumers are done with the stage's previous use
                const uint32_t bytes = (uint32_t)(min((long long)kStreamCols, left) * 8);
                unsigned char *dst = st_smem + s * kStreamStageBytes + lane * kStreamRowBytes;
                if (lane == 0) {
                    meta[s] = make_int2((int)group, c);
                    mbar_arrive_expect_tx(&full_bar[s], bytes * (rows + 1));
                }
                if (active) {
#if SKB_STREAM_HINT
                    if (lane < kStreamRows)
                        tma_bulk_g2s_hint(dst, src, bytes, &full_bar[s], policy);
                    else
#endif
                        tma_bulk_g2s(dst, src, bytes, &full_bar[s]);
                }
                src += kStreamCols;
                left -= kStreamCols;
                if (++s == kStreamStages) {
                    s = 0;
                    empty_parity ^= 1u;
                }
            }
            group = (long long)gridDim.x + (long long)__shfl_sync(kMask, ticket, 0);
        }
    }

    // ---- consumers ----
    double acc[kStreamRows];
#pragma unroll
    for (int i = 0; i < kStreamRows; ++i)
        acc[i] = 0.0;
    const int n_pairs_last = (int)((a.n_cols - (long long)(n_chunks - 1) * kStreamCols) >> 1); // pairs of the last chunk
    int s = 0;
    uint32_t full_parity = 0;
    for (;; s = (s + 1 == kStreamStages) ? 0 : s + 1, full_parity ^= (s == 0)) {
        mbar_wait(&full_bar[s], full_parity);
        const int2 m = meta[s];
        if (m.x < 0)
            break;
        const int n_pairs = m.y == n_chunks - 1 ? n_pairs_last : kStreamCols / 2;
        const unsigned char *stage = st_smem + s * kStreamStageBytes;
        // thread t owns the column pairs t and 64 + t of the chunk: consecutive lanes read consecutive 16-byte words
#pragma unroll
        for (int h = 0; h < kStreamCols / 2 / kStreamConsumers; ++h) {
            const int p = h * kStreamConsumers + t;
            if (p < n_pairs) {
                const double2 xv = *reinterpret_cast<const double2 *>(stage + kStreamRows * kStreamRowBytes + p * 16);
#pragma unroll
                for (int i = 0; i < kStreamRows; ++i) {
                    // (rows beyond n_rows of the last group hold stale ring data: their sums are never stored)
                    const double2 av = *reinterpret_cast<const double2 *>(stage + i * kStreamRowBytes + p * 16);
                    acc[i] = fma(av.x, xv.x, acc[i]);
                    acc[i] = fma(av.y, xv.y, acc[i]);
                }
            }
        }
        __syncwarp();
        if ((t & 31) == 0)
            mbar_arrive(&empty_bar[s]); // this warp is done with the stage
        if (m.y == n_chunks - 1) { // the group is complete: fixed tree over the 64 consumer threads
#pragma unroll
            for (int i = 0; i < kStreamRows; ++i) {
                double v = acc[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    v += __shfl_xor_sync(0xffffffffu, v, o);
                if ((t & 31) == 0)
                    red[(t >> 5) * kStreamRows + i] = v;
                acc[i] = 0.0;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kStreamConsumers) : "memory");
            const long long row = (long long)m.x * kStreamRows + t;
            if (t < kStreamRows && row < a.n_rows)
                a.y[row] = red[t] + red[kStreamRows + t];
            asm volatile("bar.sync 1, %0;" ::"n"(kStreamConsumers) : "memory"); // red is free again
        }
    }
}

} // namespace skb
