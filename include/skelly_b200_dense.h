/*
 * skelly_b200_dense.h -- C ABI of the periphery's dense operators on the GPU.
 *
 *   reference (SkellySim)                                           here
 *   --------------------------------------------------------------  -------------------------------
 *   Periphery::matvec               src/core/periphery.cpp:38-47     skb_dense_apply(SKB_DENSE_STRESSLET_PLUS_COMPLEMENTARY, x, v)
 *       stresslet_plus_complementary_ * x_shell + v_local
 *   Periphery::apply_preconditioner src/core/periphery.cpp:21-30     skb_dense_apply(SKB_DENSE_M_INV, x, NULL)
 *       M_inv_ * x_shell
 *   row scatter of the precompute   src/core/periphery.cpp:404-417   skb_dense_set_matrix (rows block-partitioned over GPUs)
 *
 * Both operators are (3 N_s) x (3 N_s) FP64 matrices from the precompute .npz (row-major as numpy stores them,
 * src/skelly_sim/precompute.py:113-148); they are uploaded once, x is shipped per application.  The apply is a GEMV
 * bounded by HBM bandwidth: 8 * rows * cols bytes per application.
 */
#ifndef SKELLY_B200_DENSE_H
#define SKELLY_B200_DENSE_H

#include "skelly_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct skb_dense skb_dense;

enum skb_dense_op { SKB_DENSE_STRESSLET_PLUS_COMPLEMENTARY = 0, SKB_DENSE_M_INV = 1 };

/* n_gpus devices (0..n-1) of this process; rows are split in contiguous blocks like the reference's MPI_Scatterv */
SKB_API int skb_dense_create(int n_gpus, skb_dense **out);
/* the same on an explicit device list (one rank per GPU: a single-device handle on the rank's own GPU holding the
 * rank's row block of the operators, periphery.cpp:387-417) */
SKB_API int skb_dense_create_on(const int *device_ids, int n_gpus, skb_dense **out);
SKB_API int skb_dense_device(const skb_dense *dn, int index, int *device);
SKB_API int skb_dense_destroy(skb_dense *dn);

/* A: row-major n_rows x n_cols (the numpy layout of the precompute file).  Copied to the device(s). */
SKB_API int skb_dense_set_matrix(skb_dense *dn, int op, const double *A_rowmajor, int64_t n_rows, int64_t n_cols);

/* y[n_rows] = A x (+ v_add if not NULL).  Host pointers, synchronous. */
SKB_API int skb_dense_apply(skb_dense *dn, int op, const double *x, const double *v_add, double *y);

/* Device-pointer form for a single-device handle: x, v_add (or NULL) and y already on the handle's device (device 0);
 * asynchronous on `stream` (cudaStream_t as void*, used verbatim).  This is what keeps x_shell / v_shell on the
 * device inside skb_flow_apply_matvec_dense (skelly_b200_flow.h). */
SKB_API int skb_dense_apply_device(skb_dense *dn, int op, const double *d_x, const double *d_v_add, double *d_y,
                                   void *stream);
/* y = A x with the BACKGROUND streamer (csrc/stream_kernels.cuh): one 64-thread CTA per SM, the matrix moved by TMA bulk
 * copies -- small enough to share every SM with the FP64-bound pair kernels, so that an HBM-bound operator launched on
 * a side stream costs (almost) no time of the matvec.  Same contract as skb_dense_apply_device without v_add; falls
 * back to the classic kernel when n_cols is odd or d_x is not 16-byte aligned.  skb_flow_apply_matvec_* use it for
 * Periphery::matvec (periphery.cpp:38-47) unless skb_flow_set_overlap(fl, 0). */
SKB_API int skb_dense_apply_background_device(skb_dense *dn, int op, const double *d_x, double *d_y, void *stream);
/* shape given to skb_dense_set_matrix (n_rows = -1 before it was called) */
SKB_API int skb_dense_shape(const skb_dense *dn, int op, int64_t *n_rows, int64_t *n_cols);

typedef struct skb_dense_stats {
    double kernel_ms; /* CUDA-event time of the GEMV kernel (max over devices) */
    double total_ms;  /* including H2D of x and D2H of y */
    int64_t bytes;    /* algorithmic HBM bytes: 8 * rows * cols */
} skb_dense_stats;
SKB_API int skb_dense_last_stats(const skb_dense *dn, skb_dense_stats *out);

#ifdef __cplusplus
}
#endif
#endif
