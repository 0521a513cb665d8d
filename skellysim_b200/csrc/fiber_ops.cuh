// fiber_ops.cuh -- per-fiber dense operators of the GMRES matvec, device resident.
//
//   fiber_gemv_kernel      y_f = M_f x_f for every fiber f; M_f column-major as Eigen stores
//                          FiberFiniteDifference::A_ (4n x 4n) and ::force_operator_ (3n x 4n).  HBM-bound: every
//                          matrix element is read once per matvec, 8 B / FMA.
//   fiber_velocity_kernel  the velocity / boundary part of FiberFiniteDifference::matvec
//                          (src/core/fiber_finite_difference.cpp:276-312): -P_downsample_bc * vT + xs_vT + y_BC
//
// Work decomposition of the GEMV: one CTA of 256 threads per (fiber, 64-row block); thread (lr, q) accumulates row
// row0+lr over the columns j == q (mod 4); a warp reads 32 consecutive rows of one column = 256 contiguous bytes.
#pragma once
#include <cuda_runtime.h>

namespace skb {

constexpr int kFiberGemvRows = 64;   // rows per CTA
constexpr int kFiberGemvSlices = 4;  // column slices per row
constexpr int kFiberGemvThreads = kFiberGemvRows * kFiberGemvSlices;

struct FiberGemvItem {
    long long mat_off; // element offset of M_f in the concatenated operator buffer
    long long x_off;   // element offset of x_f (4 * node offset)
    long long out_off; // MODE 0: element offset of y_f;  MODE 1: node offset of the fiber
    int rows, cols;    // shape of M_f
    int row0;          // first row of this CTA
    int n_nodes;       // nodes of the fiber
};

// MODE 0: out[out_off + row] = (M x)[row]
// MODE 1: force layout of FiberContainerFiniteDifference::apply_fiber_force (fcfd.cpp:272-287): row = k*n + i of
//         force_operator_ * x goes to fw(k, node_off + i), i.e. AoS out[3*(out_off + i) + k]
template <int MODE>
__global__ void __launch_bounds__(kFiberGemvThreads)
    fiber_gemv_kernel(const FiberGemvItem *__restrict__ items, const double *__restrict__ mats,
                      const double *__restrict__ x, double *__restrict__ out) {
    extern __shared__ double fg_smem[];
    const FiberGemvItem it = items[blockIdx.x];
    double *xs = fg_smem;                        // cols
    double *part = fg_smem + ((it.cols + 1) & ~1); // kFiberGemvSlices x kFiberGemvRows
    for (int j = threadIdx.x; j < it.cols; j += kFiberGemvThreads)
        xs[j] = x[it.x_off + j];
    __syncthreads();
    const int lr = threadIdx.x & (kFiberGemvRows - 1), q = threadIdx.x / kFiberGemvRows;
    const int row = it.row0 + lr;
    double acc = 0.0;
    if (row < it.rows) {
        const double *m = mats + it.mat_off + row;
        const long long ld = it.rows;
        int j = q;
        // 8 independent loads in flight per thread
        for (; j + 7 * kFiberGemvSlices < it.cols; j += 8 * kFiberGemvSlices) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                v[u] = __ldg(m + (long long)(j + u * kFiberGemvSlices) * ld);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                acc = fma(v[u], xs[j + u * kFiberGemvSlices], acc);
        }
        for (; j < it.cols; j += kFiberGemvSlices)
            acc = fma(__ldg(m + (long long)j * ld), xs[j], acc);
    }
    part[q * kFiberGemvRows + lr] = acc;
    __syncthreads();
    if (q == 0 && row < it.rows) {
        double s = part[lr];
#pragma unroll
        for (int u = 1; u < kFiberGemvSlices; ++u)
            s += part[u * kFiberGemvRows + lr];
        if (MODE == 0) {
            out[it.out_off + row] = s;
        } else {
            const int k = row / it.n_nodes, i = row - k * it.n_nodes;
            out[3 * (it.out_off + i) + k] = s;
        }
    }
}

// One CTA per fiber of the launch (fiber_offset points at the first of them; xs, v, res, v_boundary, length_prev,
// plus_velocity and the class offsets are indexed from that fiber).  res[4*off + r] += -(P vT)[r] (r < 4n-14) + xs_vT[r] + y_BC[r]   (ffd.cpp:276-312)
//   vT = [v_x; v_y; v_z; D_1^T (xs_x v_x + xs_y v_y + xs_z v_z)],  D_1 = D_1_0 * 2 / length_prev   (:280-293)
//   xs_vT[bc+3] = v_0 . xs_0;  xs_vT[bc+10] = v_{n-1} . xs_{n-1} when the plus end has a velocity BC  (:298-309)
//   y_BC[bc .. bc+7) = v_boundary(:, fiber)                                                           (:303-306)
// with bc = 4n - 14.  D_1_0 (n x n) and P_downsample_bc ((4n-14) x 4n) are column-major, shared by every fiber with
// the same node count; class_D / class_P give their element offsets in `class_mats` per fiber.
__global__ void __launch_bounds__(256)
    fiber_velocity_kernel(const long long *__restrict__ fiber_offset, const double *__restrict__ xs_all,
                          const double *__restrict__ v_all, const double *__restrict__ length_prev,
                          const int *__restrict__ plus_velocity, const double *__restrict__ class_mats,
                          const long long *__restrict__ class_D, const long long *__restrict__ class_P,
                          const double *__restrict__ v_boundary, double *__restrict__ res) {
    extern __shared__ double fv_smem[];
    const int f = blockIdx.x;
    const long long off = fiber_offset[f] - fiber_offset[0]; // node offset among the fibers of this launch
    const int n = (int)(fiber_offset[f + 1] - fiber_offset[f]);
    double *vT = fv_smem;      // 4n
    double *s = fv_smem + 4 * n; // n
    const double *xs = xs_all + 3 * off, *v = v_all + 3 * off;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double vx = v[3 * i], vy = v[3 * i + 1], vz = v[3 * i + 2];
        vT[i] = vx;
        vT[n + i] = vy;
        vT[2 * n + i] = vz;
        s[i] = xs[3 * i] * vx + xs[3 * i + 1] * vy + xs[3 * i + 2] * vz;
    }
    __syncthreads();
    const double scale = 2.0 / length_prev[f];
    const double *D = class_mats + class_D[f];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int j = warp; j < n; j += n_warps) { // one warp per entry: (D_1^T s)[j] = sum_i D_1(i, j) s_i, column j contiguous
        const double *col = D + (long long)j * n;
        double acc = 0.0;
        for (int i = lane; i < n; i += 32)
            acc = fma(col[i], s[i], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0)
            vT[3 * n + j] = scale * acc;
    }
    __syncthreads();
    // -(P vT): row r of P_downsample_bc by thread lr of column slice sl; S slices when the rows leave threads idle
    const int bc = 4 * n - 14, n4 = 4 * n;
    const double *P = class_mats + class_P[f];
    const int rp32 = (bc + 31) & ~31;
    const int S = 2 * rp32 <= (int)blockDim.x ? (int)blockDim.x / rp32 : 1; // column slices per row
    const int rp = S == 1 ? (int)blockDim.x : rp32;                         // threads per slice
    const int sl = threadIdx.x / rp, lr = threadIdx.x - sl * rp;
    double *part = fv_smem + 5 * n;                        // S x rp <= blockDim.x partial sums
    double *out = res + 4 * off;
    if (S == 1) {
        for (int r = lr; r < bc; r += rp) {
            const double *p = P + r;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            for (int j = 0; j + 3 < n4; j += 4) {
                a0 = fma(p[(long long)j * bc], vT[j], a0);
                a1 = fma(p[(long long)(j + 1) * bc], vT[j + 1], a1);
                a2 = fma(p[(long long)(j + 2) * bc], vT[j + 2], a2);
                a3 = fma(p[(long long)(j + 3) * bc], vT[j + 3], a3);
            }
            out[r] -= (a0 + a1) + (a2 + a3);
        }
    } else {
        double a0 = 0.0, a1 = 0.0;
        if (sl < S && lr < bc) {
            const double *p = P + lr;
            int j = sl;
            for (; j + S < n4; j += 2 * S) {
                a0 = fma(p[(long long)j * bc], vT[j], a0);
                a1 = fma(p[(long long)(j + S) * bc], vT[j + S], a1);
            }
            if (j < n4)
                a0 = fma(p[(long long)j * bc], vT[j], a0);
        }
        if (sl < S)
            part[sl * rp + lr] = a0 + a1;
        __syncthreads();
        if (sl == 0 && lr < bc) {
            double sum = part[lr];
            for (int u = 1; u < S; ++u)
                sum += part[u * rp + lr];
            out[lr] -= sum;
        }
    }
    // the 14 boundary rows: xs_vT and y_BC
    if (threadIdx.x < 14) {
        const int t = threadIdx.x;
        double val = 0.0;
        if (t == 3)
            val += v[0] * xs[0] + v[1] * xs[1] + v[2] * xs[2];
        if (v_boundary && t < 7)
            val += v_boundary[7 * (long long)f + t];
        if (t == 10 && plus_velocity[f]) {
            const int e = 3 * (n - 1);
            val += v[e] * xs[e] + v[e + 1] * xs[e + 1] + v[e + 2] * xs[e + 2];
        }
        out[bc + t] += val;
    }
}

} // namespace skb
