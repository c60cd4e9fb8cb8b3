// diral_wire.cu -- the last "next" row of SURVEY.md 8(f): the view-based positional distribution on
// neighbour tables in the RealNeS wire format, and the 3GPP semi-persistent-scheduling baseline policy.
//
//   wire_vpd_kernel   RealnessEnv.get_neighbor_dist2 / get_neighbor_dist (reference envs/realness_env.py:87-118 /
//                     :52-85) on tables of MA_NeighborTableEntry {pos_x f32, pos_y f32, seq_num i32, last_update i32}
//                     (envs/ma_messages_pb2.py): 16 B entries, so one warp reads a table as coalesced 16 B loads.
//                     Differences from the test simulator's VPD (network.py:473-513) that are kept on purpose:
//                     the age test is `last_updated > limit` (not >=), the observer's own position is its own
//                     table entry, there is no range pre-filter (numpy.histogram drops samples outside
//                     [-W, W] but the divisor counts them), and positions are float32 widened to float64.
//   sps_step_kernel   SemiPersistentScheduling.step / choose_new_resource (reference algorithms/v2x_sps.py:34-104),
//                     one thread per agent; the three random draws of a step are inputs (or Philox).
#include "diral_dev.cuh"
#include "diral_launch.h"

#include <cstdint>

namespace diral {

namespace {

constexpr int WIRE_MAX_N = 1024;     // entries per table (sorted variant keeps them in shared memory)
constexpr int WIRE_MAX_B = 256;

struct WireEntry { float x, y; int32_t seq, lu; };
struct EdgeTable { double e[WIRE_MAX_B + 1]; };   // numpy.linspace edges, passed by value (constant bank)

// RealnessEnv.dist (realness_env.py:193-207): distance and the side the neighbour is on
__device__ __forceinline__ double wire_signed_dist(double x1, double y1, double x2, double y2)
{
    const double dx = __dsub_rn(x2, x1), dy = __dsub_rn(y2, y1);
    const double d = (dy == 0.0) ? fabs(dx) : __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    return (__dsub_rn(x1, x2) > 0.0) ? d : -d;
}

// one warp per table; shared memory per warp: B counters (type 2) or N samples (type 1)
template <int TYPE>
__global__ void __launch_bounds__(128) wire_vpd_kernel(const WireEntry *__restrict__ tables, const int32_t *__restrict__ observer,
                                                       long long M, int N, int B, double W, int age_limit,
                                                       const __grid_constant__ EdgeTable et, float *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const double *edges = et.e;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long m = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (m >= M) return;
    const WireEntry *tab = tables + m * N;
    const int tx = observer[m];
    if (tx < 0 || tx >= N) {             // an observer id outside the table: NaN row instead of an out-of-bounds read
        for (int b = lane; b < B; b += 32) out[m * B + b] = __int_as_float(0x7fc00000);
        return;
    }
    const float4 own = *reinterpret_cast<const float4 *>(tab + tx);          // uniform address: one broadcast load
    const double xo = (double)own.x, yo = (double)own.y;
    float *row = out + m * B;

    if (TYPE == 2) {
        unsigned *hist = reinterpret_cast<unsigned *>(smem_raw) + warp * B;
        for (int b = lane; b < B; b += 32) hist[b] = 0u;
        __syncwarp();
        int cnt = 0;
        const double inv_binw = __ddiv_rn((double)B, __dmul_rn(2.0, W));
        for (int j = lane; j < N; j += 32) {
            const float4 raw = reinterpret_cast<const float4 *>(tab)[j];
            const int lu = __float_as_int(raw.w);
            if (j == tx || lu > age_limit) continue;                          // realness_env.py:98-101
            const double s = wire_signed_dist((double)raw.x, (double)raw.y, xo, yo);
            ++cnt;
            // numpy.histogram(range=(-W, W)) keeps -W <= s <= W and closes the last bin.  s == W is binned here,
            // not by vpd_bin: the other kernels only ever pass |s| < W, and for the clamped first guess of s == W
            // ptxas 12.9 derives `k != B - 1` from the predicate output of VIMNMX.RELU and lets the increment fire
            // (measured: bin B; profiles/README.md)
            if (s >= -W && s <= W) atomicAdd(&hist[s == W ? B - 1 : vpd_bin(s, W, inv_binw, B, edges)], 1u);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        __syncwarp();
        for (int b = lane; b < B; b += 32)
            row[b] = cnt > 0 ? (float)__ddiv_rn((double)hist[b], (double)cnt) : 0.0f;   // realness_env.py:113-116
    } else {
        // type 1: sorted samples / max |sample|, numpy.histogram(values, edges, weights=values) -- the
        // cumulative-sum path of NumPy for explicit edges (lane 0 walks the sorted list; N is small)
        double *buf = reinterpret_cast<double *>(smem_raw) + (size_t)warp * N;
        int cnt = 0;
        for (int j0 = 0; j0 < N; j0 += 32) {
            const int j = j0 + lane;
            bool keep = false; double s = 0.0;
            if (j < N) {
                const float4 raw = reinterpret_cast<const float4 *>(tab)[j];
                keep = j != tx && __float_as_int(raw.w) <= age_limit;
                if (keep) s = wire_signed_dist((double)raw.x, (double)raw.y, xo, yo);
            }
            const unsigned bm = __ballot_sync(0xffffffffu, keep);
            if (keep) buf[cnt + __popc(bm & ((1u << lane) - 1u))] = s;
            cnt += __popc(bm);
        }
        __syncwarp();
        if (lane == 0) {
            if (cnt == 0) { for (int b = 0; b < B; ++b) row[b] = 0.0f; }
            else {
                for (int i = 1; i < cnt; ++i) {                 // sorted() (realness_env.py:78)
                    const double v = buf[i]; int k = i - 1;
                    while (k >= 0 && buf[k] > v) { buf[k + 1] = buf[k]; --k; }
                    buf[k + 1] = v;
                }
                double nrm = 0.0;
                for (int q = 0; q < cnt; ++q) nrm = fmax(nrm, fabs(buf[q]));
                for (int q = 0; q < cnt; ++q) buf[q] = __ddiv_rn(buf[q], nrm);
                int idx = 0; double cum = 0.0, prev = 0.0;
                for (int b = 0; b <= B; ++b) {
                    const double edge = edges[b];
                    if (b < B) { while (idx < cnt && buf[idx] < edge)  { cum = __dadd_rn(cum, buf[idx]); ++idx; } }
                    else       { while (idx < cnt && buf[idx] <= edge) { cum = __dadd_rn(cum, buf[idx]); ++idx; } }
                    if (b > 0) row[b - 1] = (float)__dsub_rn(cum, prev);
                    prev = cum;
                }
            }
        }
    }
}

constexpr uint32_t STREAM_SPS = 4;

// SemiPersistentScheduling.step (v2x_sps.py:76-104), one thread per agent.
// draws: [A][3] float64 = (reselection counter in [5, 16], keep-uniform in [0, 1), choice index >= 0)
__global__ void sps_step_kernel(long long A, int Wn, const double *__restrict__ window, double rssi_threshold,
                                double inc_db, double prob_keep, double min_sa, const double *__restrict__ draws,
                                unsigned long long seed, long long t, int32_t *__restrict__ prev_action,
                                int32_t *__restrict__ counter, int32_t *__restrict__ actions, int32_t *__restrict__ flags)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= A) return;
    int prev = prev_action[g], cnt = counter[g];
    int action = prev;
    if (cnt != 0) { cnt -= 1; }                                                  // v2x_sps.py:87-90
    else {
        double d0, u1, d2;
        if (draws) { d0 = draws[g * 3]; u1 = draws[g * 3 + 1]; d2 = draws[g * 3 + 2]; }
        else {
            const uint4 o = philox_draw(seed, STREAM_SPS, (uint32_t)(g & 0xffffffffll), g >> 32, t);
            d0 = 5.0 + (double)__umulhi(o.x, 12u);                               // random.randint(5, 16)
            u1 = __dmul_rn((double)o.y, 1.0 / 4294967296.0);
            d2 = (double)o.z;
        }
        cnt = (int)d0;                                                           // v2x_sps.py:92
        if (!(u1 < prob_keep)) {                                                 // v2x_sps.py:94-99
            // choose_new_resource (v2x_sps.py:24-74): raise the threshold by inc_db until at least min_sa
            // subframes (other than the previous one) sense below it; the candidates are the
            // max(1, ceil(min(min_sa, |sA|))) weakest of them (stable by subframe); one is drawn uniformly
            const double *w = window + g * Wn;
            double thr = rssi_threshold, next_thr = rssi_threshold; int n_sa = 0; int guard = 0;
            for (;;) {
                thr = next_thr;                      // the threshold this scan uses
                n_sa = 0;
                for (int s = 0; s < Wn; ++s) n_sa += (s != prev && w[s] < thr) ? 1 : 0;
                next_thr = __dadd_rn(thr, inc_db);
                if (!((double)n_sa < min_sa)) break;
                if (++guard > 4096) break;           // the reference would loop forever: flagged below
            }
            // min_sa <= 0 (an empty candidate list) makes the reference raise; never-enough candidates make it spin
            if (!(min_sa > 0.0) || (double)n_sa < min_sa || n_sa == 0) { if (flags) flags[g] = 1; }
            else {
                const double min_len = fmin(min_sa, (double)n_sa);
                int n_sb = 1;
                while ((double)n_sb < min_len) ++n_sb;                            // len(sB) >= min_len stops the loop
                const int pick = (int)fmod(d2, (double)n_sb);                     // random.choice(sB)
                // the pick-th entry of sA in (rssi, subframe) order: selection by repeated minimum
                double last_v = -INFINITY; int last_s = -1; int chosen = prev;
                for (int k = 0; k <= pick; ++k) {
                    double best_v = INFINITY; int best_s = -1;
                    for (int s = 0; s < Wn; ++s) {
                        if (s == prev || !(w[s] < thr)) continue;
                        const double v = w[s];
                        const bool after = v > last_v || (v == last_v && s > last_s);
                        if (after && (v < best_v || best_s < 0)) { best_v = v; best_s = s; }
                    }
                    last_v = best_v; last_s = best_s; chosen = best_s;
                }
                action = chosen; prev = chosen;                                   // v2x_sps.py:98-99
            }
        }
    }
    prev_action[g] = prev; counter[g] = cnt; actions[g] = action;
}

}  // namespace

cudaError_t launch_wire_vpd(const void *tables, const int32_t *observer, long long M, int N, int type, int B, double W,
                            int age_limit, const double *edges, float *out, cudaStream_t stream)
{
    EdgeTable et;
    for (int i = 0; i <= B; ++i) et.e[i] = edges[i];      // host array, B + 1 entries
    const int warps = 4;
    const unsigned grid = (unsigned)((M + warps - 1) / warps);
    const WireEntry *t = static_cast<const WireEntry *>(tables);
    if (type == 2) {
        wire_vpd_kernel<2><<<grid, warps * 32, sizeof(unsigned) * warps * B, stream>>>(t, observer, M, N, B, W, age_limit, et, out);
    } else {
        const size_t smem = sizeof(double) * warps * (size_t)N;
        wire_vpd_kernel<1><<<grid, warps * 32, smem, stream>>>(t, observer, M, N, B, W, age_limit, et, out);
    }
    return cudaGetLastError();
}

int wire_max_entries() { return WIRE_MAX_N; }
int wire_max_bins() { return WIRE_MAX_B; }

cudaError_t launch_sps_step(long long A, int Wn, const double *window, double rssi_threshold, double inc_db,
                            double prob_keep, double min_sa, const double *draws, unsigned long long seed, long long t,
                            int32_t *prev_action, int32_t *counter, int32_t *actions, int32_t *flags, cudaStream_t stream)
{
    const int threads = 128;
    sps_step_kernel<<<(unsigned)((A + threads - 1) / threads), threads, 0, stream>>>(
        A, Wn, window, rssi_threshold, inc_db, prob_keep, min_sa, draws, seed, t, prev_action, counter, actions, flags);
    return cudaGetLastError();
}

}  // namespace diral
