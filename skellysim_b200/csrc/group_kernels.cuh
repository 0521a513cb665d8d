// group_kernels.cuh -- the exchange steps of a multi-GPU matvec, written against PEER MEMORY (NVLink / NVSwitch
// loads and stores from inside kernels) instead of library collectives.
//
// A group is one skb_flow per GPU -- all in one process (skb_mflow: the shape the reference's "direct evaluators need
// a single rank" rule admits, src/core/system.cpp:618-623) or one per process (CUDA IPC, rank-per-GPU hosts).  Every
// member owns a WINDOW of device memory that all members can address:
//
//   flags    [4][kMaxGroup]  epoch counters, one slot per (phase, writer)                  -- written by peers
//   f_sl     [2][n_pad_fib*3]   packed (trapezoid-weighted) Stokeslet strengths of ALL fiber nodes   -- by peers
//   f_shell  [2][n_pad_shell*6] packed sym6 stresslet strengths of ALL periphery nodes                -- by peers
//   x_shell  [2][n_shell*3]     periphery density of ALL nodes (input of the dense operator)          -- by peers
//   u_part   [n_fib*3]          this member's partial fiber velocities (its rows of the symmetric block) -- by itself
//
// One matvec = (1) PUSH: every member packs the strengths of ITS fibers / periphery rows once and stores them into
// all windows (the pack kernel and the all-gather are one kernel; 2.3 MB per member at 1e5 nodes); (2) flag barrier;
// (3) pair kernels on local memory only (TMA sources come from the member's own window); (4) flag barrier;
// (5) PULL: every member adds up the partial velocities of its own fiber rows from all windows (the reduce-scatter,
// fused with the accumulation into the velocity vector).  [2] = double buffering by epoch parity: a peer may already
// push the next matvec's strengths while this member still reads the current ones.
#pragma once
#include <cuda_runtime.h>

namespace skb {

constexpr int kMaxGroup = 16;
constexpr int kGroupPhases = 4;

struct GroupPushArgs {
    // own fibers: raw forces -> weighted, into rows [fa, fa + n_f) of every member's f_sl
    const double *fw;     // [n_f*3]
    const double *weight; // [n_fib] trapezoid weights of ALL fiber nodes (nullptr: 1)
    long long fa, n_f;
    // own periphery rows: density -> sym6 (2 eta n (x) rho) into rows [sa, sa + n_s) of f_shell, raw into x_shell
    const double *density; // [n_s*3]
    const double *normal;  // [n_shell*3] normals of ALL periphery nodes
    long long sa, n_s;
    double two_eta;
    int size;
    double *f_sl[kMaxGroup];
    double *f_shell[kMaxGroup];
    double *x_shell[kMaxGroup];
};

// pack + all-gather in one pass: every value is formed once and stored to `size` windows (coalesced per window)
__global__ void group_push_kernel(const GroupPushArgs a) {
    const long long nf3 = a.n_f * 3;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nf3 + a.n_s; i += stride) {
        if (i < nf3) {
            const long long node = a.fa + i / 3;
            double v = a.fw[i];
            if (a.weight)
                v *= a.weight[node];
            const long long o = 3 * a.fa + i;
            for (int m = 0; m < a.size; ++m)
                a.f_sl[m][o] = v;
        } else {
            const long long s = i - nf3, g = a.sa + s;
            // products formed exactly as the host does (2*eta*n_a*rho_b, then the symmetric sums): bit-identical to
            // pack_dl_normal_density_kernel / pack_dl9 of the host-formed f_dl (periphery.cpp:68-71)
            const double n0 = a.two_eta * a.normal[3 * g + 0], n1 = a.two_eta * a.normal[3 * g + 1],
                         n2 = a.two_eta * a.normal[3 * g + 2];
            const double r0 = a.density[3 * s + 0], r1 = a.density[3 * s + 1], r2 = a.density[3 * s + 2];
            double o6[6];
            o6[0] = n0 * r0;
            o6[1] = n1 * r1;
            o6[2] = n2 * r2;
            o6[3] = __dadd_rn(__dmul_rn(n0, r1), __dmul_rn(n1, r0));
            o6[4] = __dadd_rn(__dmul_rn(n0, r2), __dmul_rn(n2, r0));
            o6[5] = __dadd_rn(__dmul_rn(n1, r2), __dmul_rn(n2, r1));
            for (int m = 0; m < a.size; ++m) {
                double *d6 = a.f_shell[m] + 6 * g;
#pragma unroll
                for (int k = 0; k < 6; ++k)
                    d6[k] = o6[k];
                double *dx = a.x_shell[m] + 3 * g;
                dx[0] = r0, dx[1] = r1, dx[2] = r2;
            }
        }
    }
}

struct GroupFlagArgs {
    unsigned long long *flags[kMaxGroup]; // every member's flag array [kGroupPhases][kMaxGroup]
    int rank, size, phase;
    unsigned long long epoch;
    int do_signal, do_wait;
    unsigned long long timeout_cycles; // a peer that never arrives must not hang the GPU: give up and flag an error
};

// flags[kGroupPhases * kMaxGroup] of a window is the error word
__global__ void group_flag_kernel(const GroupFlagArgs a) {
    const int m = threadIdx.x;
    if (m >= a.size)
        return;
    if (a.do_signal) {
        __threadfence_system(); // everything this member stored before (earlier kernels of the stream included)
        volatile unsigned long long *p = a.flags[m] + a.phase * kMaxGroup + a.rank;
        *p = a.epoch;
    }
    if (a.do_wait) {
        volatile unsigned long long *q = a.flags[a.rank] + a.phase * kMaxGroup + m;
        const long long t0 = clock64();
        while (*q < a.epoch) {
            if ((unsigned long long)(clock64() - t0) > a.timeout_cycles) {
                a.flags[a.rank][kGroupPhases * kMaxGroup] = 1ULL + (unsigned long long)m; // who was missing
                break;
            }
            __nanosleep(200);
        }
        __threadfence_system();
    }
}

struct GroupPullArgs {
    const double *u_part[kMaxGroup]; // every member's partial fiber velocities [n_fib*3]
    int size;
    long long fa, n_f; // own fiber rows
    double *v;         // [n_f*3] (+)= sum over members, fixed order
    int accumulate;
};

// reduce-scatter fused with the accumulation: own rows only, members added in rank order (bitwise reproducible)
__global__ void group_pull_kernel(const GroupPullArgs a) {
    const long long n3 = a.n_f * 3;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3)
        return;
    double acc = a.accumulate ? a.v[i] : 0.0;
    for (int m = 0; m < a.size; ++m)
        acc += __ldcv(a.u_part[m] + 3 * a.fa + i); // peer memory: never from a stale cache line
    a.v[i] = acc;
}

} // namespace skb
