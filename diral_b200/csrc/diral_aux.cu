// diral_aux.cu -- the kernels around the fused slot: standalone observation build, reset, action
// sampling, velocity jitter, information age and the end-of-episode metric reduction.
#include "diral_dev.cuh"
#include "diral_launch.h"

#include <algorithm>
#include <cstdint>

namespace diral {

namespace {

constexpr int SORT_MAX_N = 256;      // per-thread local arrays of the sorted variants
constexpr int HIST_MAX_B = 256;

// A thread's scratch array of the sorted State variants: element k of thread t sits at base[k * stride + t] in SHARED
// memory (conflict-free: consecutive threads, consecutive words).  Round 1 kept 256 doubles per thread in local memory
// whatever the vehicle count, which made the sorted variants 3-4x the cost of the whole slot kernel.
struct StridedBuf {
    double *base; int stride;
    __device__ __forceinline__ double &operator[](int k) const { return base[k * stride]; }
};

__device__ __forceinline__ void insertion_sort(const StridedBuf &a, int n)
{
    for (int i = 1; i < n; ++i) {
        const double v = a[i];
        int k = i - 1;
        while (k >= 0 && a[k] > v) { a[k + 1] = a[k]; --k; }
        a[k + 1] = v;
    }
}

// Up to 32 values per thread sort in REGISTERS: a bitonic network with compile-time indices (240 compare-exchanges, no
// shared-memory traffic, no data-dependent branch); unused slots hold +inf and end up behind the values.  Equal values
// are bit-identical here (a zero signed distance is always -0.0), so the network's instability cannot show.
__device__ __forceinline__ void bitonic32(double (&v)[32])
{
#pragma unroll
    for (int ks = 1; ks <= 5; ++ks)
#pragma unroll
        for (int js = ks - 1; js >= 0; --js)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int k = 1 << ks, j = 1 << js, l = i ^ j;
                if (l > i) {
                    const double a = v[i], b = v[l];
                    const bool sw = ((i & k) == 0) ? (a > b) : (a < b);
                    v[i] = sw ? b : a; v[l] = sw ? a : b;
                }
            }
}

// the first n (<= 32) values of a thread's scratch array, sorted in place through registers (one copy of the network
// for every caller: the unrolled code is ~1 200 instructions)
__device__ __noinline__ void sort32(const StridedBuf buf, int n)
{
    double v[32];
#pragma unroll
    for (int t = 0; t < 32; ++t) v[t] = t < n ? buf[t] : INFINITY;
    bitonic32(v);
#pragma unroll
    for (int t = 0; t < 32; ++t) if (t < n) buf[t] = v[t];
}

// TestEnv.obtain_state (reference envs/test_env.py:527-583) on caller-supplied obs / actions /
// rewards, one thread per (env, vehicle).  Every State flag is honoured, including the two variants
// the fused kernels leave to this one: the direct sorted positional distribution
// (Network.get_positional_dist, network.py:409-430) and VPD type 1
// (Network.get_positional_dist_piggy, network.py:432-471).
// KIND: 0 no sorted block, 1 the direct distribution only, 2 VPD type 1 only (one inlined copy of the sorting network
// each, values never leave registers before they are sorted), 3 both (they share the out-of-line sort32).
template <int KIND>
__global__ void __launch_bounds__(128) obtain_state_kernel(const Params p, const float *__restrict__ obs,
                                                           const int32_t *__restrict__ actions,
                                                           const float *__restrict__ rews, float *__restrict__ out,
                                                           const double *__restrict__ edges1)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.E * p.N) return;
    const int N = p.N, R = p.R, B = p.B;
    const long long e = gid / N;
    const int u = (int)(gid - e * N);
    const long long vbase = e * N;
    const double xu = p.pos_x[gid], yu = p.pos_y[gid];
    float *row = out + gid * p.S;
    int k = 0;
    const int a = actions[gid];
    if (p.add_action) {
        if (p.action_binary) { for (int r = 0; r < R; ++r) row[k++] = (a == r) ? 1.0f : 0.0f; }
        else row[k++] = (float)a;
    }
    if (p.add_channel_obs) { for (int r = 0; r < R; ++r) row[k++] = obs[gid * R + r]; }

    constexpr bool SORTED = KIND != 0;
    extern __shared__ double sort_scratch[];
    const StridedBuf buf{sort_scratch + threadIdx.x, (int)blockDim.x};      // (only touched by the SORTED instantiation)
    if ((KIND & 1) && p.add_positional_dist) {      // network.py:409-430
        int m = 0; double max_dist = 0.0;
        if (N <= 32) {
          if constexpr (KIND == 1) {
            double v[32];
#pragma unroll
            for (int t = 0; t < 32; ++t) {
                double sv = INFINITY;
                if (t < N && t != u) {
                    const double xt = p.pos_x[vbase + t], yt = p.pos_y[vbase + t];
                    const double d = dist2d(xt, yt, xu, yu);
                    if (d > max_dist) max_dist = d;
                    sv = (__dsub_rn(xt, xu) > 0.0) ? d : -d;
                }
                v[t] = sv;
            }
            bitonic32(v);
#pragma unroll
            for (int t = 0; t < 32; ++t) if (t < N) buf[t] = v[t];
            m = N - 1;
          } else {
            for (int t = 0; t < N; ++t) {
                if (t == u) continue;
                const double xt = p.pos_x[vbase + t], yt = p.pos_y[vbase + t];
                const double d = dist2d(xt, yt, xu, yu);
                if (d > max_dist) max_dist = d;
                buf[m++] = (__dsub_rn(xt, xu) > 0.0) ? d : -d;
            }
            sort32(buf, m);
          }
        } else {
            for (int t = 0; t < N; ++t) {
                if (t == u) continue;
                const double xt = p.pos_x[vbase + t], yt = p.pos_y[vbase + t];
                const double d = dist2d(xt, yt, xu, yu);
                if (d > max_dist) max_dist = d;
                buf[m++] = (__dsub_rn(xt, xu) > 0.0) ? d : -d;
            }
            insertion_sort(buf, m);
        }
        for (int q = 0; q < m; ++q) row[k++] = (float)__ddiv_rn(buf[q], max_dist);
    }
    if (p.piggy) {
        // (i = this vehicle, j = subject) through the layout-independent accessors of diral_dev.cuh
        auto seq_at = [&](int j) { return p.tab_seq[tab_index(p, e, u, j)]; };
        auto lu_at = [&](int j) { return p.tab_lu[tab_index(p, e, u, j)]; };
        auto x_at = [&](int j) { return tab_xpos(p, e, u, j, p.tab_seq[tab_index(p, e, u, j)]); };
        if (p.pos_dist_type == 2 || !p.vpd_enabled) {        // network.py:473-513
            unsigned short hist[HIST_MAX_B];
            for (int b = 0; b < B; ++b) hist[b] = 0;
            int m = 0;
            if (p.vpd_enabled) {
                for (int j = 0; j < N; ++j) {
                    if (j == u || lu_at(j) >= p.age_threshold) continue;
                    const double x1 = x_at(j);
                    const double y1 = seq_at(j) > 0 ? p.pos_y[vbase + j] : 0.0;
                    const double d = dist2d(x1, y1, xu, yu);
                    if (d < p.W) {
                        const double s = (__dsub_rn(x1, xu) > 0.0) ? d : -d;
                        hist[vpd_bin(s, p.W, p.inv_binw, B, p.edges)] += 1;
                        ++m;
                    }
                }
            }
            const float den = (float)m;
            for (int b = 0; b < B; ++b) row[k++] = m > 0 ? __fdiv_rn((float)hist[b], den) : 0.0f;
        } else if (KIND & 2) {                                // network.py:432-471 (type 1)
            int m = 0;
            if (N <= 32 && !p.layout) {
              if constexpr (KIND == 2) {
                // subject-major tables: entry (u, j) of all three arrays sits at base + j * N; the loads of eight
                // entries are issued before the first dependent use, the values sort in registers
                const long long base = e * (long long)N * N + u;
                double v[32];
#pragma unroll
                for (int j0 = 0; j0 < 32; j0 += 8) {
                    int sq[8], lu8[8]; double x8[8], y8[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int j = j0 + q;
                        const bool have = j < N && j != u;
                        const long long idx = base + (long long)j * N;
                        sq[q] = have ? p.tab_seq[idx] : 0;
                        lu8[q] = have ? p.tab_lu[idx] : p.age_threshold;
                        x8[q] = have ? p.tab_x[idx] : 0.0;
                        y8[q] = have ? p.pos_y[vbase + j] : 0.0;
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        double sv = INFINITY;
                        if (lu8[q] < p.age_threshold) {
                            const double d = dist2d(x8[q], sq[q] > 0 ? y8[q] : 0.0, xu, yu);
                            sv = (__dsub_rn(x8[q], xu) > 0.0) ? d : -d;
                            ++m;
                        }
                        v[j0 + q] = sv;
                    }
                }
                if (m > 0) {
                    bitonic32(v);
#pragma unroll
                    for (int t = 0; t < 32; ++t) if (t < N) buf[t] = v[t];
                }
              } else {
                const long long base = e * (long long)N * N + u;
                for (int j0 = 0; j0 < N; j0 += 8) {
                    int sq[8], lu8[8]; double x8[8], y8[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int j = j0 + q;
                        const bool have = j < N && j != u;
                        const long long idx = base + (long long)j * N;
                        sq[q] = have ? p.tab_seq[idx] : 0;
                        lu8[q] = have ? p.tab_lu[idx] : p.age_threshold;
                        x8[q] = have ? p.tab_x[idx] : 0.0;
                        y8[q] = have ? p.pos_y[vbase + j] : 0.0;
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (lu8[q] >= p.age_threshold) continue;
                        const double d = dist2d(x8[q], sq[q] > 0 ? y8[q] : 0.0, xu, yu);
                        buf[m++] = (__dsub_rn(x8[q], xu) > 0.0) ? d : -d;
                    }
                }
                if (m > 0) sort32(buf, m);
              }
            } else {
                for (int j = 0; j < N; ++j) {
                    if (j == u || lu_at(j) >= p.age_threshold) continue;
                    const double x1 = x_at(j);
                    const double y1 = seq_at(j) > 0 ? p.pos_y[vbase + j] : 0.0;
                    const double d = dist2d(x1, y1, xu, yu);
                    buf[m++] = (__dsub_rn(x1, xu) > 0.0) ? d : -d;
                }
                if (m > 0) insertion_sort(buf, m);
            }
            if (m == 0) { for (int b = 0; b < B; ++b) row[k++] = 0.0f; }
            else {
                double nrm = 0.0;
                for (int q = 0; q < m; ++q) nrm = fmax(nrm, fabs(buf[q]));
                for (int q = 0; q < m; ++q) buf[q] = __ddiv_rn(buf[q], nrm);
                // np.histogram(values, explicit edges, weights=values): cumulative sums of the sorted
                // weights, searchsorted 'left' for every edge but the last ('right'), differences
                int idx = 0; double cum = 0.0, prev = 0.0;
                for (int b = 0; b <= B; ++b) {
                    const double edge = edges1[b];
                    if (b < B) { while (idx < m && buf[idx] < edge)  { cum = __dadd_rn(cum, buf[idx]); ++idx; } }
                    else       { while (idx < m && buf[idx] <= edge) { cum = __dadd_rn(cum, buf[idx]); ++idx; } }
                    if (b > 0) row[k++] = (float)__dsub_rn(cum, prev);
                    prev = cum;
                }
            }
        }
    }
    if (p.add_reward) row[k++] = rews[gid];
    if (p.add_index) row[k++] = (float)(u + 1);
    if (p.add_position) { row[k++] = (float)__ddiv_rn(xu, p.L); row[k++] = (float)__ddiv_rn(yu, 2.0); }
    if (p.add_velocity) row[k++] = (float)p.vel[gid];
    if (p.fingerprint) { row[k++] = (float)p.episode; row[k++] = (float)p.epsilon; }
}

// Network.initialize_mobility_topology* (network.py:69-79,92-112) on the counter-based generator
__global__ void reset_kernel(const Params p, const double *x0, const double *y0, const double *v0,
                             double *pos_y, double *vel)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.E * p.N) return;
    const long long e = gid / p.N;
    const int u = (int)(gid - e * p.N);
    double x, y, v;
    if (x0) { x = x0[gid]; y = y0[gid]; v = v0[gid]; }
    else if (p.design_topology) { x = 195.0 * u; y = u < 2 ? 1.0 : 2.0; v = 1.0; }   // network.py:74-79
    else {
        const uint4 o = philox_draw(p.seed, STREAM_TOPOLOGY, (uint32_t)u, p.env0 + e, 0);
        x = (double)__umulhi(o.x, (uint32_t)p.L);                  // integer-valued uniform on [0, L)
        y = 0.0;                                                   // randint(0, highway_height/2) == 0
        const double r53 = __dmul_rn(__dadd_rn(__dmul_rn((double)(o.y >> 5), 67108864.0), (double)(o.z >> 6)),
                                     1.0 / 9007199254740992.0);
        v = p.mobility_vary ? 1.7 : __dadd_rn(1.1, __dmul_rn(2.7 - 1.1, r53));   // random.uniform(1.1, 2.7)
    }
    p.pos_x[gid] = x; pos_y[gid] = y; vel[gid] = v;
}

__global__ void sample_kernel(const Params p, int32_t *out)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.E * p.N) return;
    const long long e = gid / p.N;
    out[gid] = philox_action(p.seed, (int)(gid - e * p.N), p.env0 + e, p.timestep, p.R);
}

// Network.update_velocity (network.py:208-222)
__global__ void update_velocity_kernel(const Params p, double *vel, const int8_t *draws, long long episode)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.E * p.N) return;
    const long long e = gid / p.N;
    int d;
    if (draws) d = draws[gid];
    else d = 1 + (int)__umulhi(philox_draw(p.seed, STREAM_VELOCITY, (uint32_t)(gid - e * p.N), p.env0 + e, episode).x, 3u);
    double v = vel[gid];
    if (d == 1) { v = __dadd_rn(v, 0.55); if (v > 2.77) v = 2.77; }
    else if (d == 2) { v = __dsub_rn(v, 0.55); if (v < 1.1) v = 1.1; }
    vel[gid] = v;
}

__device__ __forceinline__ void ia_accumulate(int *h, int lat, long long timestep)
{
    if (lat == -1) return;                                         // network.py:569
    long long ia = timestep - lat;
    if (ia >= IA_BINS) return;
    if (ia < 0) ia += IA_BINS;                                     // Python negative index
    if (ia >= 0) atomicAdd(&h[ia], 1);
}

// Network.get_information_age (network.py:560-574), one CTA per env
__global__ void information_age_kernel(const Params p, int32_t *out)
{
    __shared__ int h[IA_BINS];
    const long long e = blockIdx.x;
    for (int i = threadIdx.x; i < IA_BINS; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const int NN = p.N * p.N;
    const int32_t *lat = p.lat + e * (long long)NN;
    for (int i = threadIdx.x; i < NN; i += blockDim.x) {
        const int t = i / p.N, r = i - t * p.N;
        if (t != r) ia_accumulate(h, lat[i], p.timestep);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < IA_BINS; i += blockDim.x) out[e * IA_BINS + i] = h[i];
}

// End-of-episode metric vector; single CTA, fixed reduction order => bit-reproducible.
__global__ void __launch_bounds__(1024) episode_metrics_kernel(const Params p, double *out)
{
    __shared__ double red[5][1024];
    __shared__ int h[IA_BINS];
    const int T = blockDim.x, tid = threadIdx.x;
    for (int i = tid; i < IA_BINS; i += T) h[i] = 0;
    double s[5] = {0, 0, 0, 0, 0};
    for (long long e = tid; e < p.E; e += T) {
        s[0] += p.acc_reward[e];
        long long *c = p.acc_count + e * ACC_COUNTS;
        s[1] += (double)c[0]; s[2] += (double)c[1]; s[3] += (double)c[2]; s[4] += (double)c[3];
        p.acc_reward[e] = 0.0; c[0] = 0; c[1] = 0; c[2] = 0; c[3] = 0;
    }
    for (int q = 0; q < 5; ++q) red[q][tid] = s[q];
    __syncthreads();
    if (p.track_lat && p.lat) {
        const long long tot = p.E * (long long)p.N * p.N;
        for (long long i = tid; i < tot; i += T) {
            const long long w = i % ((long long)p.N * p.N);
            const int t = (int)(w / p.N), r = (int)(w - (long long)t * p.N);
            if (t != r) ia_accumulate(h, p.lat[i], p.timestep);
        }
    }
    for (int o = T / 2; o > 0; o >>= 1) {
        if (tid < o) { for (int q = 0; q < 5; ++q) red[q][tid] += red[q][tid + o]; }
        __syncthreads();
    }
    if (tid == 0) {
        const double slots = red[4][0];
        out[0] = red[0][0];
        out[1] = slots * p.R - red[0][0];          // main_test.py:178 summed over env-slots
        out[2] = red[1][0]; out[3] = red[2][0];
        out[4] = slots * p.N; out[5] = red[3][0];
        out[6] = slots; out[7] = 0.0; out[8] = 0.0; out[9] = 0.0;
    }
    for (int i = tid; i < IA_BINS; i += T) out[10 + i] = (double)h[i];
}

// xpos of every table entry in the reference's own indexing, out[e][i][j] = vehicles[i].pos_of_neighbors[j]["xpos"]
// (vehicle.py:30): a transpose for the subject-major layout, ring / spill look-ups for the row layout
__global__ void materialize_x_kernel(const Params p, double *out)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long NN = (long long)p.N * p.N;
    if (gid >= p.E * NN) return;
    const long long e = gid / NN;
    const int i = (int)((gid - e * NN) / p.N), j = (int)(gid - e * NN - (long long)i * p.N);
    out[gid] = tab_xpos(p, e, i, j, p.tab_seq[tab_index(p, e, i, j)]);
}

inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

cudaError_t launch_obtain_state(const Params &p, const float *obs, const int32_t *actions, const float *rews,
                                float *out, cudaStream_t stream)
{
    const bool sorted = p.add_positional_dist || (p.piggy && p.pos_dist_type == 1);
    // edges1 (linspace(-1, 1, B+1)) lives right behind edges in the same device allocation
    const double *edges1 = p.edges + (p.B + 1);
    if (!sorted) {
        obtain_state_kernel<0><<<blocks_for(p.E * p.N, 128), 128, 0, stream>>>(p, obs, actions, rews, out, edges1);
        return cudaGetLastError();
    }
    // N doubles of shared-memory scratch per thread: as many threads per CTA as 200 KB allow, at most 128
    int threads = 128;
    while (threads > 32 && (size_t)threads * p.N * sizeof(double) > (size_t)200 * 1024) threads -= 32;
    const size_t smem = (size_t)threads * p.N * sizeof(double);
    const int kind = (p.add_positional_dist ? 1 : 0) | ((p.piggy && p.pos_dist_type == 1) ? 2 : 0);
    auto launch = [&](auto kernel) -> cudaError_t {
        if (smem > 48 * 1024) {
            cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (err != cudaSuccess) return err;
        }
        kernel<<<blocks_for(p.E * p.N, threads), threads, smem, stream>>>(p, obs, actions, rews, out, edges1);
        return cudaGetLastError();
    };
    if (kind == 1) return launch(obtain_state_kernel<1>);
    if (kind == 2) return launch(obtain_state_kernel<2>);
    return launch(obtain_state_kernel<3>);
}

cudaError_t launch_reset(const Params &p, const double *x0, const double *y0, const double *v0, cudaStream_t stream)
{
    reset_kernel<<<blocks_for(p.E * p.N, 256), 256, 0, stream>>>(p, x0, y0, v0, const_cast<double *>(p.pos_y),
                                                                 const_cast<double *>(p.vel));
    return cudaGetLastError();
}

cudaError_t launch_sample(const Params &p, int32_t *out, cudaStream_t stream)
{
    sample_kernel<<<blocks_for(p.E * p.N, 256), 256, 0, stream>>>(p, out);
    return cudaGetLastError();
}

cudaError_t launch_update_velocity(const Params &p, double *vel, const int8_t *draws, long long episode,
                                   cudaStream_t stream)
{
    update_velocity_kernel<<<blocks_for(p.E * p.N, 256), 256, 0, stream>>>(p, vel, draws, episode);
    return cudaGetLastError();
}

cudaError_t launch_information_age(const Params &p, int32_t *out, cudaStream_t stream)
{
    information_age_kernel<<<(unsigned)p.E, 128, 0, stream>>>(p, out);
    return cudaGetLastError();
}

cudaError_t launch_materialize_x(const Params &p, double *out, cudaStream_t stream)
{
    materialize_x_kernel<<<blocks_for(p.E * (long long)p.N * p.N, 256), 256, 0, stream>>>(p, out);
    return cudaGetLastError();
}

cudaError_t launch_episode_metrics(const Params &p, double *out110, cudaStream_t stream)
{
    episode_metrics_kernel<<<1, 1024, 0, stream>>>(p, out110);
    return cudaGetLastError();
}

}  // namespace diral

// ---- per-slot caller epilogue (reference main_test.py:150-206, utils/misc.py:1-12) -------------------
namespace diral {

namespace {

// One CTA per environment: information-age histogram of this slot (Network.get_information_age,
// network.py:560-574), its weighted sum (calculate_ia_penalty, utils/misc.py:1-12), then the reward
// shaping of main_test.py:153-206 in place.  Rewards arrive as the float32 the slot kernel wrote; all
// arithmetic here is float64 like the reference's, rounded to float32 once at the store.
__global__ void shape_rewards_kernel(const Params p, const ShapingArgs s)
{
    __shared__ int h[IA_BINS];
    __shared__ double s_sum[32];
    __shared__ long long s_ia;
    __shared__ int s_penalty;
    const long long e = blockIdx.x;
    const int N = p.N, T = blockDim.x, tid = threadIdx.x;
    for (int i = tid; i < IA_BINS; i += T) h[i] = 0;
    __syncthreads();
    if (p.lat && p.track_lat) {
        const int32_t *lat = p.lat + e * (long long)N * N;
        for (int i = tid; i < N * N; i += T) {
            const int t = i / N, r = i - t * N;
            if (t != r) ia_accumulate(h, lat[i], p.timestep);
        }
    }
    // sum of the raw rewards (main_test.py:174), fixed order: lane-strided partials, then a tree
    double part = 0.0;
    for (int i = tid; i < N; i += T) part += (double)s.rewards[e * N + i];
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) s_sum[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
        double sum_r = 0.0;
        for (int w = 0; w < (T + 31) / 32; ++w) sum_r += s_sum[w];
        s_sum[0] = sum_r;
        long long ia_sum = 0;                                      // utils/misc.py:7-10
        for (int i = 0; i < IA_BINS; ++i) if (h[i] > 0) ia_sum += (long long)(i + 1) * h[i];
        s_ia = ia_sum;
        int penalty = 0;
        if (s.ia_averaging) {                                      // main_test.py:153-160
            const long long prev = s.sum_ia_prev[e];
            penalty = ia_sum > prev ? -1 : (ia_sum < prev ? 1 : 0);
            s.sum_ia_prev[e] = ia_sum;
        }
        s_penalty = penalty;
        if (s.slot_sums) {
            s.slot_sums[e * 3 + 0] = sum_r;
            s.slot_sums[e * 3 + 1] = (double)p.R - sum_r;          // collision, main_test.py:178
            s.slot_sums[e * 3 + 2] = (double)ia_sum;
        }
    }
    __syncthreads();
    if (s.ia_out) for (int i = tid; i < IA_BINS; i += T) s.ia_out[e * IA_BINS + i] = h[i];
    const double sum_r = s_sum[0];
    for (int i = tid; i < N; i += T) {                             // main_test.py:188-206
        const long long k = e * N + i;
        double r = (double)s.rewards[k];
        if (s.ia_averaging) r += (double)s_penalty;
        if (s.ia_penalty_enable) {
            const int a = s.actions[k];
            int c = s.ia_counter[k];
            c = (r < 1.0 && a == s.prev_actions[k]) ? c + 1 : 0;
            if (c > s.ia_penalty_threshold) r = s.ia_penalty_value;
            s.ia_counter[k] = c;
            s.prev_actions[k] = a;
        }
        if (s.global_reward_avg) r = r + sum_r / (double)N;
        s.rewards[k] = (float)r;
    }
}

}  // namespace

cudaError_t launch_shape_rewards(const Params &p, const ShapingArgs &s, cudaStream_t stream)
{
    const int threads = p.N <= 32 ? 32 : (p.N <= 128 ? 128 : 256);
    shape_rewards_kernel<<<(unsigned)p.E, threads, 0, stream>>>(p, s);
    return cudaGetLastError();
}

}  // namespace diral

// ---- device-resident replay ring (reference utils/memory.py:162-194 + drl_drqn.py:294-377) ------------
namespace diral {

namespace {

// out[(a * batch + b) * step + k][0..width) = ring[(start[b] + k) mod capacity][a][0..width)
// One warp per (a, b, k) row group; `width` elements of `VEC` bytes move as coalesced vectors.
template <typename V>
__global__ void ring_gather_kernel(const V *__restrict__ ring, long long capacity, long long agents, int width,
                                   const long long *__restrict__ start, int batch, int step, V *__restrict__ out)
{
    const long long rows = agents * batch * step;
    const int lane = threadIdx.x & 31;
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
         row += (long long)gridDim.x * (blockDim.x >> 5)) {
        const int k = (int)(row % step);
        const long long ab = row / step;
        const int b = (int)(ab % batch);
        const long long a = ab / batch;
        long long slot = (start[b] + k) % capacity;
        if (slot < 0) slot += capacity;
        const V *src = ring + (slot * agents + a) * width;
        V *dst = out + row * width;
        for (int i = lane; i < width; i += 32) dst[i] = src[i];
    }
}

}  // namespace

cudaError_t launch_ring_gather(const void *ring, long long capacity, long long agents, long long width, int elem_bytes,
                               const long long *start, int batch, int step, void *out, cudaStream_t stream)
{
    const long long rows = agents * batch * step;
    if (rows <= 0 || width <= 0) return cudaSuccess;
    const long long row_bytes = width * elem_bytes;
    const int threads = 256;
    const unsigned grid = (unsigned)std::min<long long>((rows + 7) / 8, 148LL * 32);
    const bool a16 = (row_bytes % 16 == 0) && ((uintptr_t)ring % 16 == 0) && ((uintptr_t)out % 16 == 0);
    if (a16)
        ring_gather_kernel<uint4><<<grid, threads, 0, stream>>>(static_cast<const uint4 *>(ring), capacity, agents,
                                                                (int)(row_bytes / 16), start, batch, step, static_cast<uint4 *>(out));
    else if (row_bytes % 4 == 0)
        ring_gather_kernel<uint32_t><<<grid, threads, 0, stream>>>(static_cast<const uint32_t *>(ring), capacity, agents,
                                                                   (int)(row_bytes / 4), start, batch, step, static_cast<uint32_t *>(out));
    else
        ring_gather_kernel<uint8_t><<<grid, threads, 0, stream>>>(static_cast<const uint8_t *>(ring), capacity, agents,
                                                                  (int)row_bytes, start, batch, step, static_cast<uint8_t *>(out));
    return cudaGetLastError();
}

}  // namespace diral
