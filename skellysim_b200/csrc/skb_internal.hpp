// skb_internal.hpp -- declarations shared by the translation units of libskelly_b200.so
#pragma once
#include "../../include/skelly_b200.h"

#include <cuda_runtime.h>
#include <functional>
#include <stddef.h>

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

DeviceInfo query_device(int dev);
LaunchPlan plan_launch(const DeviceInfo &di, int kind, long long n_trg, int n_src_tiles, int force_T, int force_S);
int launch_pair_sum(const DeviceInfo &di, int kind, const double *d_r_src, const double *d_f_packed, long long n_src,
                    long long n_src_pad, const double *d_r_trg, long long n_trg, double *d_partial,
                    const LaunchPlan &plan, cudaStream_t st);
int launch_reduce(const double *d_partial, double *d_u, long long n_trg, int n_splits, double scale, int accumulate,
                  cudaStream_t st);

// NCCL is bound at run time (dlopen "libnccl.so.2"): a single-GPU user never needs it, and a host that
// already loaded NCCL (PyTorch) shares that copy instead of getting a second one.
int nccl_group_create(size_t n, const std::function<int(int)> &device_of, void **out);
int nccl_group_allgather_inplace(void *group, void *const *bufs, size_t count_per_rank_doubles,
                                 const cudaStream_t *streams);
void nccl_group_destroy(void *group);

} // namespace skb
