// diral_launch.h -- host-side launch entry points of the CUDA kernels (internal to libdiral_env.so)
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

namespace diral {

struct Params;
struct ShapingArgs;

// N <= 32: one lane group per environment, table keys in registers (diral_step_group.cu)
constexpr int GROUP_MAX_N = 32;
int group_width(int N);
size_t step_group_smem_bytes(const Params &p);
cudaError_t prepare_step_group(const Params &p);
cudaError_t launch_step_group(const Params &p, cudaStream_t stream);

// any N <= 256: persistent CTAs, one environment at a time, table keys in shared memory or scratch (diral_step_block.cu)
constexpr int BLOCK_MAX_N = 256;
int key_src_bits(int N);
size_t step_block_smem_bytes(const Params &p, bool keys_in_smem);
bool step_block_keys_fit_smem(const Params &p);
size_t step_block_scratch_bytes(long long E, int N);
size_t step_block_scratch_words_per_env(int N);
cudaError_t prepare_step_block(const Params &p);
cudaError_t launch_step_block(const Params &p, cudaStream_t stream);

// 33 <= N <= 64: one warp per environment, two table rows per lane (diral_step_pair.cu); subject-major layout like the group kernel
bool step_pair_supported(const Params &p);
size_t step_pair_smem_bytes(const Params &p);
cudaError_t prepare_step_pair(const Params &p);
cudaError_t launch_step_pair(const Params &p, cudaStream_t stream);

// 32 < N <= 256, ROW layout: observer-major tables, positions in a ring, receiver-centric merges (diral_step_row.cu)
bool step_row_supported(const Params &p);
int step_row_stride(int N);             // T: padded row stride of the tables
int step_row_ring_depth(int N);         // H: ticks the position ring holds
size_t step_row_smem_bytes(const Params &p);
size_t step_row_scratch_bytes(long long E, int N);
cudaError_t prepare_step_row(const Params &p);
cudaError_t launch_step_row(const Params &p, cudaStream_t stream);

// standalone kernels (diral_aux.cu)
cudaError_t launch_obtain_state(const Params &p, const float *obs, const int32_t *actions, const float *rews,
                                float *out, cudaStream_t stream);
cudaError_t launch_reset(const Params &p, const double *x0, const double *y0, const double *v0, cudaStream_t stream);
cudaError_t launch_sample(const Params &p, int32_t *out, cudaStream_t stream);
cudaError_t launch_update_velocity(const Params &p, double *vel, const int8_t *draws, long long episode,
                                   cudaStream_t stream);
cudaError_t launch_information_age(const Params &p, int32_t *out, cudaStream_t stream);
cudaError_t launch_episode_metrics(const Params &p, double *out110, cudaStream_t stream);
cudaError_t launch_materialize_x(const Params &p, double *out, cudaStream_t stream);
cudaError_t launch_shape_rewards(const Params &p, const ShapingArgs &s, cudaStream_t stream);
cudaError_t launch_ring_gather(const void *ring, long long capacity, long long agents, long long width, int elem_bytes,
                               const long long *start, int batch, int step, void *out, cudaStream_t stream);


// RealNeS wire-format positional distribution and the SPS baseline policy (diral_wire.cu)
int wire_max_entries();
int wire_max_bins();
cudaError_t launch_wire_vpd(const void *tables, const int32_t *observer, long long M, int N, int type, int B, double W,
                            int age_limit, const double *host_edges, float *out, cudaStream_t stream);
cudaError_t launch_sps_step(long long A, int Wn, const double *window, double rssi_threshold, double inc_db,
                            double prob_keep, double min_sa, const double *draws, unsigned long long seed, long long t,
                            int32_t *prev_action, int32_t *counter, int32_t *actions, int32_t *flags, cudaStream_t stream);

}  // namespace diral
