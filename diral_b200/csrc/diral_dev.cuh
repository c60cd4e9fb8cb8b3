// diral_dev.cuh -- kernel parameter block and device helpers shared by the step / state kernels.
//
// Arithmetic rules (they are what makes the integer state bit-exact against the reference):
//   * positions, velocities and every distance are float64, rounded operation by operation
//     (__dmul_rn / __dadd_rn / __dsqrt_rn are never contracted into FMAs), mirroring
//     Network.dist (reference envs/network.py:318-332): sqrt((x2-x1)**2 + (y2-y1)**2);
//   * when dy == 0 the distance is |dx| -- bit-identical to sqrt(fl(dx*dx) + 0) under IEEE
//     round-to-nearest, and it skips the multi-instruction fp64 sqrt on the common y == 0 highway;
//   * outputs (obs, rewards, state) are rounded to float32 once, at the store.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace diral {

constexpr int MODE_STEP = 0, MODE_DESIGN = 1, MODE_CH = 2;
constexpr int IA_BINS = 100;          // network.py:566
constexpr int ACC_COUNTS = 4;         // acc_count columns: received, in-range pairs, bad actions, slots

// RNG stream tags (specification shared with oracle/diral_oracle.c, not code)
constexpr uint32_t STREAM_ACTIONS = 1, STREAM_TOPOLOGY = 2, STREAM_VELOCITY = 3;

struct Params;

struct Params {
    // sizes
    long long E, env0;
    int N, R, B, S;
    // geometry
    double L, C, C2, W, sentinel, inv_binw;
    int age_threshold;
    // behaviour switches (all warp-uniform)
    int reward_design, state_type, toy, mobility, mobility_vary, design_topology, piggy;
    int add_action, action_binary, add_channel_obs, add_reward, add_index, add_velocity, add_position,
        add_positional_dist, pos_dist_type, fingerprint;
    int vpd_enabled;          // piggy && (mobility || design_topology)  (network.py:545)
    int fast_nearest;         // C <= sentinel: a lone in-range transmitter is the nearest without comparing
    // per-call
    int mode, track_lat, build_state, gen_actions;
    int prefetch_ahead;       // envs between this CTA's env and the one that will reuse its SM slot
    int tail_split;           // lane-group kernel, 32 lanes: 0 never split an environment over warps, 1 the tail of a batch, 2 always
    int n_slots;              // consecutive slots one launch of the lane-group kernel runs (fused rollout), >= 1
    long long timestep;
    int tick;                 // table ticks since the reset including this slot (= every vehicle's own seq)
    double episode, epsilon;
    unsigned long long seed;
    // device memory
    const int32_t *actions;
    int32_t *actions_out;
    double *pos_x; const double *pos_y; const double *vel;
    int32_t *tab_seq, *tab_lu; double *tab_x;
    int32_t *lat;
    float *obs, *rews, *state;
    uint8_t *vpd_counts;      // compact host record per agent (diral_step_host) or NULL: B bin counts (one byte each, when
                              // piggy), padded to 4 bytes, then the float32 reward; rec_stride bytes apart
    int rec_stride;
    // streamed host records (diral_step_host, host_format 3): ONE launch covers the batch, the records go straight to
    // mapped host memory and the kernel tells the host which chunk of environments is complete (see env_records_done)
    unsigned *chunk_count;    // [chunks] device counters, zero between slots
    unsigned *chunk_flag;     // [chunks] mapped host words: = chunk_epoch once the chunk's records are in host memory
    int chunk_envs;           // environments per chunk (the last chunk may be shorter)
    unsigned chunk_epoch;     // this slot's flag value
    double *acc_reward; long long *acc_count;
    uint32_t *scratch;
    const double *trace; long long trace_len;
    const double *edges;      // [B+1] numpy.linspace(-W, W, B+1), computed on the host in float64
    // table layout: 0 = SUBJECT-major [E][N][N] with a dense xpos table (lane-group kernel, round-1 block kernel);
    // 1 = ROW layout (diral_step_row.cu): tab_seq / tab_lu OBSERVER-major [E][N][T], positions in ring [E][H][T] by
    // tick mod H, tab_x = two spill halves [2][E][N][T] (spill_half elements apart) for entries older than H ticks
    int layout, T, H;
    double *ring;
    long long spill_half;
};

// ---- table accessors that work for both layouts (standalone kernels) -------------------------------------------
__device__ __forceinline__ long long tab_index(const Params &p, long long e, int i /*observer*/, int j /*subject*/)
{
    return p.layout ? (e * p.N + i) * (long long)p.T + j : (e * p.N + j) * (long long)p.N + i;
}
// xpos of vehicle i's entry about j (sequence number sn) as of the last completed slot (p.tick)
__device__ __forceinline__ double tab_xpos(const Params &p, long long e, int i, int j, int sn)
{
    if (!p.layout) return p.tab_x[tab_index(p, e, i, j)];
    if (sn <= 0) return 0.0;
    if (p.tick - sn < p.H) return p.ring[(e * p.H + (sn & (p.H - 1))) * (long long)p.T + j];
    return p.tab_x[((p.tick & 1) ? p.spill_half : 0) + tab_index(p, e, i, j)];
}

// Streamed host records: called by ONE thread of environment e after every thread's record stores were ordered before it
// (__syncwarp / __syncthreads).  The system-scope fence makes the records visible to the host before the count moves;
// whoever completes a chunk re-arms its counter and raises the chunk's flag in host memory, which the host threads that
// assemble the state rows poll -- no copy engine, no event, no second launch between the kernel and the consumer.
__device__ __forceinline__ void env_records_done(const Params &p, long long e)
{
#ifdef DIRAL_FENCE_SYS_PER_ENV
    __threadfence_system();
#else
    // Device scope is enough here: the environment that completes the chunk observes every other one through the counter
    // (release / acquire at device scope) and issues the one system-scope fence before the flag -- causality order is
    // transitive across the two scopes, so the host that acquires the flag sees every record of the chunk.  A system
    // fence per environment made each of them wait for the whole PCIe write queue (first flag at 68 us, not 40).
    __threadfence();
#endif
    const int k = (int)(e / p.chunk_envs);
    const long long left = p.E - (long long)k * p.chunk_envs;
    const unsigned size = (unsigned)(left < p.chunk_envs ? left : p.chunk_envs);
    if (atomicAdd(p.chunk_count + k, 1u) + 1u == size) {
        p.chunk_count[k] = 0u;               // the last environment of the chunk: nobody else touches it this slot
        __threadfence_system();
        *reinterpret_cast<volatile unsigned *>(p.chunk_flag + k) = p.chunk_epoch;
    }
}

// per-slot caller epilogue (main_test.py:150-206): device pointers and switches of one call
struct ShapingArgs {
    int ia_averaging, ia_penalty_enable, ia_penalty_threshold, global_reward_avg;
    double ia_penalty_value;
    const int32_t *actions;       // [E][N]
    float *rewards;               // [E][N] in/out
    long long *sum_ia_prev;       // [E]      (ia_averaging)
    int32_t *ia_counter;          // [E][N]   (ia_penalty_enable)
    int32_t *prev_actions;        // [E][N]   (ia_penalty_enable)
    double *slot_sums;            // [E][3] = {sum of raw rewards, collisions, weighted information age} or NULL
    int32_t *ia_out;              // [E][100] or NULL
};

// ---- Philox4x32-10 (Salmon et al., SC'11; published round constants) -------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ uint4 philox_draw(unsigned long long seed, uint32_t stream, uint32_t agent,
                                             long long env, long long t)
{
    return philox4x32_10(make_uint4(agent, (uint32_t)env, (uint32_t)t, (uint32_t)((unsigned long long)t >> 32)),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ stream));
}

// TestEnv.sample (test_env.py:116-122) on the counter-based generator
__device__ __forceinline__ int philox_action(unsigned long long seed, int agent, long long env, long long t, int R)
{
    return (int)__umulhi(philox_draw(seed, STREAM_ACTIONS, (uint32_t)agent, env, t).x, (uint32_t)R);
}

// ---- distances ----------------------------------------------------------------------------------
// Network.dist (network.py:318-332)
__device__ __forceinline__ double dist2d(double x1, double y1, double x2, double y2)
{
    const double dx = __dsub_rn(x2, x1), dy = __dsub_rn(y2, y1);
    if (dy == 0.0) return fabs(dx);
    return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// Python float `%` for a positive modulus (network.py:202): fmod, then fold a negative remainder
__device__ __forceinline__ double py_mod_pos(double a, double m)
{
    double r = fmod(a, m);
    if (r != 0.0 && r < 0.0) r = __dadd_rn(r, m);
    return r;
}

// Network.update_positions (network.py:189-206)
__device__ __forceinline__ double mobility_step_at(const Params &p, double x, double v, int u, long long timestep)
{
    if (!p.mobility) return x;
    if (p.trace) {
        long long t = timestep % p.trace_len;
        if (t < 0) t += p.trace_len;
        return p.trace[t * p.N + u];
    }
    return py_mod_pos(__dadd_rn(__dadd_rn(x, v), p.L), p.L);
}
__device__ __forceinline__ double mobility_step(const Params &p, double x, double v, int u)
{
    return mobility_step_at(p, x, v, u, p.timestep);
}

// numpy.histogram(.., bins=B, range=(-W, W)) bin of a sample with |s| < W: NumPy's equal-width fast
// path followed by its two edge corrections lands in the bin that contains s with respect to
// linspace(-W, W, B+1) (half-open, last bin closed) -- numpy/lib/_histograms_impl.py.
__device__ __forceinline__ int vpd_bin(double s, double W, double inv_binw, int B, const double *edges)
{
    int k = (int)(__dmul_rn(__dadd_rn(s, W), inv_binw));
    k = max(0, min(k, B - 1));
    if (s < edges[k]) { if (k > 0) --k; }
    else if (k != B - 1 && s >= edges[k + 1]) ++k;
    return k;
}

// CPython >= 3.12 builtin sum() over floats: Neumaier-compensated (Python/bltinmodule.c)
struct PySum {
    double f = 0.0, c = 0.0; bool first = true;
    __device__ __forceinline__ void add(double x)
    {
        if (first) { f = __dadd_rn(0.0, x); first = false; return; }
        const double t = __dadd_rn(f, x);
        if (fabs(f) >= fabs(x)) c = __dadd_rn(c, __dadd_rn(__dsub_rn(f, t), x));
        else                    c = __dadd_rn(c, __dadd_rn(__dsub_rn(x, t), f));
        f = t;
    }
    __device__ __forceinline__ double result() const
    {
        return (c != 0.0 && isfinite(c)) ? __dadd_rn(f, c) : f;
    }
};

// Reward of a transmitter that shares its resource with tot-1 others under TestEnv.my_step
// (test_env.py:163-199); w = calculate_reward_weights(..)[0] where the design consults it.
__device__ __forceinline__ double collision_reward_step(int design, int tot, int w)
{
    switch (design) {
    case 1: return -1.0 * (1.0 - (double)w / (double)tot);
    case 2: return tot == 2 ? (double)(2 * w) - 2.0 : 0.0 - (double)tot;
    case 3: return -1.0 * exp(1.0 - 1.0 / (double)tot);
    case 4: return 1.0 / (double)tot;
    case 5: return (tot == 2 && w == 1) ? 0.0 : -1.0;
    default: return 0.0;
    }
}
__device__ __forceinline__ bool design_needs_weight(int design, int tot)
{
    return design == 1 || ((design == 2 || design == 5) && tot == 2);
}

// Reward of a transmitter under TestEnv.my_step_ch (test_env.py:402-429).  in_range == 0 makes the
// reference use the *int* 1, so design 2 yields +0.0 there but -0.0 when received == in_range.
__device__ __forceinline__ double channel_reward(int design, int tot, int received, int in_range)
{
    if (tot == 1) return design == 3 ? 1.0 : design == 4 ? exp(1.0) : design == 2 ? 1.0 : 0.0;
    const bool int_one = in_range == 0;
    const double prr = int_one ? 1.0 : (double)received / (double)in_range;
    if (design == 3) return 1.0 - exp(1.0 - prr);
    if (design == 4) return -1.0 * exp(1.0 - prr);
    if (design == 2) return int_one ? 0.0 : -1.0 * (1.0 - prr);
    return 0.0;
}

}  // namespace diral
