// diral_api.cu -- the C ABI declared in include/diral_env.h: configuration checks, kernel-variant
// choice, per-call parameter blocks and launches.  No torch, no CPU compute path: every entry point
// either launches the sm_100a kernels or fails with a message.
#include "../../include/diral_env.h"
#include "diral_dev.cuh"
#include "diral_host.h"
#include "diral_launch.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <immintrin.h>
#include <memory>
#include <mutex>
#include <new>
#include <sched.h>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_last_error = buf;
    return code;
}

#define DIRAL_CUDA(expr)                                                                       \
    do {                                                                                       \
        cudaError_t err__ = (expr);                                                            \
        if (err__ != cudaSuccess)                                                              \
            return fail(DIRAL_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
    } while (0)

enum Variant { VARIANT_AUTO = 0, VARIANT_GROUP = 1, VARIANT_BLOCK = 2, VARIANT_BLOCK_V1 = 3, VARIANT_ROW = 4, VARIANT_PAIR = 5 };
// Measured on B200 (profiles/README.md): the row-layout kernel wins from ~100 vehicles on (1.07x at 128, 1.67x at 256);
// below that its padded rows (T = 128 for 65..96 vehicles) and per-environment fixed costs lose to round 1's kernel.
constexpr int ROW_MIN_N = 97;
constexpr int MAX_HOST_CHUNKS = 64;

struct Handle {
    diral_cfg cfg{};
    diral_buffers bufs{};
    diral::Params base{};
    bool bound = false;
    bool group_ok = false, block_ok = false;   // which slot kernels this configuration's shared-memory carve-up fits
    bool row_ok = false;            // the row-layout kernel (diral_step_row.cu) takes this configuration
    bool pair_ok = false;           // the two-rows-per-lane warp kernel (diral_step_pair.cu, 33..64 vehicles) does
    int device = 0;
    int variant = VARIANT_AUTO;
    int force_track_lat = 0;
    bool lat_live = false;          // a my_step_ch call has stamped last_arrival_time since the reset
    long long ticks = 0;            // table ticks since the reset (= every vehicle's own seq number)
    long long launches = 0;
    double *d_edges = nullptr;      // [2*(B+1)]: linspace(-W, W, B+1) then linspace(-1, 1, B+1)
    int32_t *d_actions = nullptr;   // staging for generated / host-side actions
    cudaStream_t pipe[2] = {nullptr, nullptr};   // diral_step_host: env chunks alternate between these
    cudaEvent_t pipe_ev[3] = {nullptr, nullptr, nullptr};
    // compact host format of diral_step_host (see diral_host.h)
    int host_format = 0;            // 0 = full rows over PCIe, 1 = compact record + host-side row assembly,
                                    // 2 = the same, records written by the kernel straight into mapped host memory,
                                    // 3 = streamed: one launch, records into mapped host memory, per-chunk completion
                                    //     flags raised by the kernel (lane-group kernel; other kernels run as format 1)
    int stream_split = 1;           // format 3: launch the split-environment instantiation (Params::tail_split = 2): half as
                                    // many environments in flight, each twice as fast, so records -- and the PCIe writes
                                    // and the row assembly behind them -- start 25 us earlier and arrive spread out
    int stream_chunks = 16;         // format 3: chunks the assembly threads are released by
    int actions_direct = 1;         // the kernel reads the caller's actions in place when they are pinned: 0 never, 1 format 3
                                    // only (measured: -12 us there, +15 us for the chunked formats), 2 always
    unsigned *d_chunk_count = nullptr;    // [MAX_HOST_CHUNKS] device counters (zero between slots)
    unsigned *h_chunk_flag = nullptr;     // [MAX_HOST_CHUNKS] mapped pinned flags
    unsigned *d_chunk_flag = nullptr;     // device address of the same
    unsigned chunk_epoch = 0;
    bool async_pending = false;     // diral_step_host_begin without its diral_step_host_wait yet
    cudaStream_t async_stream = nullptr;
    uint8_t *d_counts_mapped = nullptr;   // device address of h_counts (mapped pinned allocation)
    int host_threads = 0;           // 0 = pick from the CPUs this process may run on
    int host_chunks = 8;
    int tail_split = 0;             // see Params::tail_split (opt-in: it shortens a small tail wave, not a saturated device)
    int host_nt = -1;               // output-row stores: 1 non-temporal, 0 ordinary, -1 by output size per thread
    diral::HostPool *pool = nullptr;
    bool pool_shared = false;       // `pool` is the process-wide one (not owned)
    int want_shared_pool = 0;       // option "host_pool_shared"
    unsigned long long job_id = 0;  // the pool job of the slot in flight
    uint8_t *d_counts = nullptr;    // [E][N][B] device
    uint8_t *h_counts = nullptr;    // pinned staging of the same
    float *h_obs_stage = nullptr;   // pinned [E][N][R] when the State block wants obs and the caller passes no h_obs
    double *h_kin = nullptr;        // pinned [3][E][N]: post-mobility x, y, velocity (add_position / add_velocity)
    cudaEvent_t chunk_ev[MAX_HOST_CHUNKS] = {};
    cudaStream_t chunk_stream[MAX_HOST_CHUNKS] = {};
    double trace_us[MAX_HOST_CHUNKS + 4] = {};   // timeline of the last compact diral_step_host call (diral_host_trace)
    int trace_n = 0;
};

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int state_space(const diral_cfg &c)
{
    int s = 0;                                              // test_env.py:49-85
    if (c.add_action) s += c.action_binary ? c.R : 1;
    if (c.add_channel_obs) s += c.R;
    if (c.add_reward) s += 1;
    if (c.add_index) s += 1;
    if (c.add_velocity) s += 1;
    if (c.add_position) s += 2;
    if (c.add_positional_dist) s += c.N - 1;
    if (c.fingerprint) s += 2;
    if (c.add_piggy) s += c.B;
    return s;
}

int check_cfg(const diral_cfg *c)
{
    if (!c) return fail(DIRAL_ERR_ARG, "cfg is NULL");
    if (c->E < 1) return fail(DIRAL_ERR_ARG, "E must be >= 1 (got %lld)", (long long)c->E);
    if (c->N < 1 || c->N > diral::BLOCK_MAX_N) return fail(DIRAL_ERR_ARG, "num_users must be in [1, %d] (got %d)", diral::BLOCK_MAX_N, c->N);
    if (c->R < 1 || c->R > 4096) return fail(DIRAL_ERR_ARG, "num_channels must be in [1, 4096] (got %d)", c->R);
    if (c->B < 1 || c->B > 256) return fail(DIRAL_ERR_ARG, "num_bins must be in [1, 256] (got %d)", c->B);
    if (!(c->L > 0.0) || c->L >= 4294967296.0) return fail(DIRAL_ERR_ARG, "highway_length must be in (0, 2^32)");
    if (!(c->W > 0.0)) return fail(DIRAL_ERR_ARG, "bin_range must be > 0");
    if (c->add_piggy && c->pos_dist_type != 1 && c->pos_dist_type != 2)
        return fail(DIRAL_ERR_ARG, "add_positional_dist_type must be 1 or 2 (the reference raises for %d)", c->pos_dist_type);
    if ((c->add_positional_dist || (c->add_piggy && c->pos_dist_type == 1)) && c->N > 256)
        return fail(DIRAL_ERR_UNSUPPORTED, "sorted positional-distribution variants support num_users <= 256");
    if (state_space(*c) < 1) return fail(DIRAL_ERR_ARG, "State selects an empty state vector");
    return DIRAL_OK;
}

bool use_group(const Handle *h)
{
    if (h->variant == VARIANT_BLOCK || h->variant == VARIANT_BLOCK_V1 || h->variant == VARIANT_ROW || h->variant == VARIANT_PAIR) return false;
    if (h->variant == VARIANT_AUTO && !h->block_ok) return true;
    return h->group_ok;
}

// the row-layout kernel is the one-CTA-per-env kernel of choice wherever it applies; VARIANT_BLOCK_V1 pins round 1's
bool use_pair(const Handle *h)
{
    if (!h->pair_ok || use_group(h)) return false;
    return h->variant == VARIANT_PAIR || h->variant == VARIANT_AUTO || h->variant == VARIANT_BLOCK;
}

bool use_row(const Handle *h)
{
    if (!h->row_ok || use_group(h) || h->variant == VARIANT_BLOCK_V1 || h->variant == VARIANT_PAIR) return false;
    if (h->variant == VARIANT_ROW) return true;
    return !use_pair(h) && (h->cfg.N >= ROW_MIN_N || !h->block_ok);
}

bool fused_state_ok(const diral_cfg &c)
{
    return !c.add_positional_dist && !(c.add_piggy && c.pos_dist_type == 1);
}

// the compact host format carries the positional distribution as one byte per bin (<= 255 samples per observer)
bool compact_ok(const diral_cfg &c)
{
    return fused_state_ok(c) && c.N <= 256;
}

diral::HostLayout host_layout(const diral_cfg &c)
{
    diral::HostLayout l{};
    l.N = c.N; l.R = c.R; l.B = c.B; l.S = state_space(c);
    l.add_action = c.add_action != 0; l.action_binary = c.action_binary != 0; l.add_channel_obs = c.add_channel_obs != 0;
    l.piggy = c.add_piggy != 0; l.add_reward = c.add_reward != 0; l.add_index = c.add_index != 0;
    l.add_position = c.add_position != 0; l.add_velocity = c.add_velocity != 0; l.fingerprint = c.fingerprint != 0;
    l.L = c.L;
    return l;
}

void fill_base(Handle *h)
{
    const diral_cfg &c = h->cfg;
    diral::Params &p = h->base;
    p = diral::Params{};
    p.n_slots = 1; p.tail_split = h->tail_split;
    p.E = c.E; p.env0 = c.env0; p.N = c.N; p.R = c.R; p.B = c.B; p.S = state_space(c);
    p.L = c.L; p.C = c.C; p.C2 = 2 * c.C; p.W = c.W; p.sentinel = c.sentinel;
    p.inv_binw = (double)c.B / (2.0 * c.W);
    p.age_threshold = c.age_threshold;
    p.reward_design = c.reward_design; p.state_type = c.state_type; p.toy = c.toy != 0;
    p.mobility = c.mobility != 0; p.mobility_vary = c.mobility_vary != 0; p.design_topology = c.design_topology != 0; p.piggy = c.add_piggy != 0;
    p.add_action = c.add_action != 0; p.action_binary = c.action_binary != 0;
    p.add_channel_obs = c.add_channel_obs != 0; p.add_reward = c.add_reward != 0; p.add_index = c.add_index != 0;
    p.add_velocity = c.add_velocity != 0; p.add_position = c.add_position != 0;
    p.add_positional_dist = c.add_positional_dist != 0; p.pos_dist_type = c.pos_dist_type;
    p.fingerprint = c.fingerprint != 0;
    p.vpd_enabled = p.piggy && (p.mobility || p.design_topology);
    p.fast_nearest = c.C <= c.sentinel;
    p.edges = h->d_edges;
}

// table layout of the kernel that will run (fixed once buffers are bound)
void set_layout(Handle *h)
{
    diral::Params &p = h->base;
    const bool row = use_row(h);
    p.layout = row ? 1 : 0;
    p.T = row ? diral::step_row_stride(h->cfg.N) : h->cfg.N;
    p.H = row ? diral::step_row_ring_depth(h->cfg.N) : 0;
    p.spill_half = row ? (long long)h->cfg.E * h->cfg.N * p.T : 0;
}

size_t handle_scratch_bytes(const Handle *h)
{
    if (!h->cfg.add_piggy || use_group(h) || use_pair(h)) return 0;
    if (use_row(h)) return diral::step_row_scratch_bytes(h->cfg.E, h->cfg.N);
    return diral_scratch_bytes(&h->cfg);
}

cudaError_t launch_slot(const Handle *h, const diral::Params &p, cudaStream_t s)
{
    if (use_group(h)) return diral::launch_step_group(p, s);
    if (use_pair(h)) return diral::launch_step_pair(p, s);
    return use_row(h) ? diral::launch_step_row(p, s) : diral::launch_step_block(p, s);
}

void bind_params(Handle *h)
{
    diral::Params &p = h->base;
    const diral_buffers &b = h->bufs;
    p.ring = b.ring;
    p.pos_x = b.pos_x; p.pos_y = b.pos_y; p.vel = b.vel;
    p.tab_seq = b.tab_seq; p.tab_lu = b.tab_lu; p.tab_x = b.tab_x; p.lat = b.lat;
    p.obs = b.obs; p.rews = b.rews; p.state = b.state;
    p.acc_reward = b.acc_reward; p.acc_count = reinterpret_cast<long long *>(b.acc_count);
    p.scratch = b.scratch; p.trace = b.trace; p.trace_len = b.trace ? b.trace_len : 0;
}

// numpy.linspace(start, stop, num) as numpy/_core/function_base.py evaluates it (float64,
// arange(num) * step + start, last element forced to stop).  Built without FMA contraction.
void np_linspace(double start, double stop, int num, double *y)
{
    const int div = num - 1;
    const double delta = stop - start;
    volatile double step = delta / div;
    for (int k = 0; k < num; ++k) {
        volatile double t = (double)k;
        if (step == 0) { t = t / div; t = t * delta; } else { t = t * step; }
        y[k] = t + start;
    }
    if (num > 1) y[num - 1] = stop;
}

Handle *as_handle(void *h) { return static_cast<Handle *>(h); }

// fresh neighbour tables (vehicle.py:24-33) in whichever layout the handle uses
int zero_tables(Handle *h, cudaStream_t s)
{
    const diral_cfg &c = h->cfg;
    const diral::Params &p = h->base;
    const size_t ent = (size_t)c.E * c.N * (p.layout ? p.T : c.N);
    if (h->bufs.tab_seq) DIRAL_CUDA(cudaMemsetAsync(h->bufs.tab_seq, 0, ent * 4, s));
    if (h->bufs.tab_lu) DIRAL_CUDA(cudaMemsetAsync(h->bufs.tab_lu, 0, ent * 4, s));
    if (h->bufs.tab_x) DIRAL_CUDA(cudaMemsetAsync(h->bufs.tab_x, 0, ent * 8 * (p.layout ? 2 : 1), s));
    if (p.layout && h->bufs.ring) DIRAL_CUDA(cudaMemsetAsync(h->bufs.ring, 0, (size_t)c.E * p.H * p.T * 8, s));
    return DIRAL_OK;
}

int require_bound(Handle *h)
{
    if (!h) return fail(DIRAL_ERR_ARG, "handle is NULL");
    if (!h->bound) return fail(DIRAL_ERR_UNBOUND, "diral_bind() has not been called on this handle");
    // (every stateful entry point comes through here: nothing may touch the environment while a begun slot is in flight)
    if (h->async_pending) return fail(DIRAL_ERR_ARG, "a begun slot is pending: call diral_step_host_wait first");
    return DIRAL_OK;
}

// the same parameter block restricted to envs [e0, e0 + n): every per-env array is contiguous per env
diral::Params env_range(const diral::Params &p, long long e0, long long n)
{
    diral::Params q = p;
    const long long N = p.N, NN = N * (p.layout ? p.T : N);
    q.E = n; q.env0 = p.env0 + e0;
    if (q.actions) q.actions += e0 * N;
    if (q.actions_out) q.actions_out += e0 * N;
    q.pos_x += e0 * N; q.pos_y += e0 * N; q.vel += e0 * N;
    if (q.tab_seq) { q.tab_seq += e0 * NN; q.tab_lu += e0 * NN; q.tab_x += e0 * NN; }
    if (q.ring) q.ring += e0 * (long long)p.H * p.T;
    if (q.lat) q.lat += e0 * N * N;
    q.obs += e0 * N * p.R; q.rews += e0 * N; q.state += e0 * N * p.S;
    if (q.vpd_counts) q.vpd_counts += e0 * N * p.rec_stride;
    q.acc_reward += e0; q.acc_count += e0 * diral::ACC_COUNTS;
    if (q.scratch) q.scratch += e0 * (p.layout ? (long long)p.T * p.T : (long long)diral::step_block_scratch_words_per_env(p.N));
    return q;
}

int ensure_actions_staging(Handle *h)
{
    if (!h->d_actions) DIRAL_CUDA(cudaMalloc(&h->d_actions, sizeof(int32_t) * (size_t)h->cfg.E * h->cfg.N));
    return DIRAL_OK;
}

// CPUs this process may run on (cgroup / affinity aware), for the default size of the row-assembly pool
int usable_cpus()
{
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof set, &set) == 0) { const int n = CPU_COUNT(&set); if (n > 0) return n; }
    const unsigned hc = std::thread::hardware_concurrency();
    return hc ? (int)hc : 1;
}

// ends a row-assembly job on every exit path: a worker must never be left spinning on a chunk that will not come
struct PoolJobGuard {
    diral::HostPool *pool; unsigned long long id; int chunks; bool closed = false;
    void close() { if (!closed) { pool->publish(id, chunks - 1); pool->finish(id); closed = true; } }
    ~PoolJobGuard() { if (!closed) { pool->abort(id); close(); } }     // error path: the missing chunks are skipped
};

// one process-wide pool for the handles that ask for it ("host_pool_shared"): groups of environments stepped with
// diral_step_host_begin / _wait then take turns on ALL assembly threads instead of idling on a private share
diral::HostPool *shared_pool(int threads)
{
    static std::mutex mu;
    static std::unique_ptr<diral::HostPool> pool;
    std::lock_guard<std::mutex> lock(mu);
    if (!pool) pool.reset(new (std::nothrow) diral::HostPool(threads));
    return pool.get();
}

// diral_step_host, compact host format: only the information of a slot crosses PCIe -- per agent one record of VPD
// bin counts (one byte per bin) and the float32 reward, plus obs / positions / velocities when the State block
// carries them -- and the [E][N][S] rows are assembled in the caller's buffer by the handle's host threads while later
// env chunks are still in flight.  A chunk is latency-bound on the device (copy in, one slot kernel, copy out: ~45 us
// whatever its size), so every chunk runs on its own stream and all of them overlap.
int step_host_compact(Handle *h, int mode, const int32_t *h_actions, int64_t timestep, double episode, double epsilon,
                      float *h_state, float *h_rews, float *h_obs, cudaStream_t s, bool async_begin = false)
{
    const auto t_entry = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_entry).count(); };
    const diral_cfg &c = h->cfg;
    const long long E = c.E, N = c.N, R = c.R, A = E * N;
    const long long nb = c.add_piggy ? c.B : 0, rec = ((nb + 3) & ~3ll) + 4;     // counts, padding, float32 reward
    const bool group = use_group(h);
    const int src_bits = group ? diral::key_src_bits(diral::group_width(c.N)) : diral::key_src_bits(c.N);
    if (c.add_piggy && h->ticks + 1 >= (1ll << (32 - src_bits)))
        return fail(DIRAL_ERR_SEQ_RANGE, "slot %lld since reset exceeds the %d-bit sequence field of the packed table keys",
                    h->ticks + 1, 32 - src_bits);
    const bool want_obs = c.add_channel_obs != 0, want_kin = c.add_position || c.add_velocity;
    if (!h->d_counts) {
        DIRAL_CUDA(cudaMalloc(&h->d_counts, (size_t)(A * rec)));
        DIRAL_CUDA(cudaHostAlloc(&h->h_counts, (size_t)(A * rec), cudaHostAllocMapped));
        DIRAL_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&h->d_counts_mapped), h->h_counts, 0));
    }
    // zero-copy: the lane-group kernel stages an environment's records in shared memory and writes them out as one
    // coalesced stream, so they can go over PCIe as the kernel runs instead of through a copy engine afterwards
    const bool streamed = h->host_format == 3 && group && h->d_counts_mapped != nullptr && !want_obs && !h_obs && !want_kin;
    if (async_begin && !streamed)
        return fail(DIRAL_ERR_UNSUPPORTED, "diral_step_host_begin needs host_format 3 on a configuration the lane-group kernel "
                                           "serves (N <= 32, fused State block without obs / positions / velocities)");
    const bool zero_copy = (h->host_format == 2 || streamed) && group && h->d_counts_mapped != nullptr;
    if (want_obs && !h_obs && !h->h_obs_stage) DIRAL_CUDA(cudaMallocHost(&h->h_obs_stage, sizeof(float) * (size_t)(A * R)));
    if (want_kin && !h->h_kin) DIRAL_CUDA(cudaMallocHost(&h->h_kin, sizeof(double) * (size_t)(3 * A)));
    int chunks = E >= 1024 ? h->host_chunks : 1;
    long long chunk_envs = 0;
    if (streamed) {
        chunk_envs = (E + h->stream_chunks - 1) / h->stream_chunks;
        chunks = (int)((E + chunk_envs - 1) / chunk_envs);
        if (!h->d_chunk_count) {
            DIRAL_CUDA(cudaMalloc(&h->d_chunk_count, sizeof(unsigned) * MAX_HOST_CHUNKS));
            DIRAL_CUDA(cudaMemset(h->d_chunk_count, 0, sizeof(unsigned) * MAX_HOST_CHUNKS));
            DIRAL_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&h->h_chunk_flag), sizeof(unsigned) * MAX_HOST_CHUNKS, cudaHostAllocMapped));
            memset(h->h_chunk_flag, 0, sizeof(unsigned) * MAX_HOST_CHUNKS);
            DIRAL_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&h->d_chunk_flag), h->h_chunk_flag, 0));
        }
    }
    if (!h->pipe_ev[2]) DIRAL_CUDA(cudaEventCreateWithFlags(&h->pipe_ev[2], cudaEventDisableTiming));
    for (int k = 0; k < chunks && !streamed; ++k) {
        if (!h->chunk_stream[k]) DIRAL_CUDA(cudaStreamCreateWithFlags(&h->chunk_stream[k], cudaStreamNonBlocking));
        if (!h->chunk_ev[k]) DIRAL_CUDA(cudaEventCreateWithFlags(&h->chunk_ev[k], cudaEventDisableTiming));
    }
    if (!h->pool) {
        int n = h->host_threads > 0 ? h->host_threads : std::min(std::max(usable_cpus() - 2, 1), 32);
        h->pool_shared = h->want_shared_pool != 0;
        h->pool = h->pool_shared ? shared_pool(n) : new (std::nothrow) diral::HostPool(n);
        if (!h->pool) return fail(DIRAL_ERR_ARG, "out of host memory");
    }

    diral::Params p = h->base;
    p.mode = mode; p.timestep = timestep; p.episode = episode; p.epsilon = epsilon; p.seed = 0;
    p.tick = (int)(h->ticks + 1);
    p.build_state = 1; p.actions = h->d_actions; p.gen_actions = 0; p.actions_out = nullptr;
    p.vpd_counts = zero_copy ? h->d_counts_mapped : h->d_counts; p.rec_stride = (int)rec;
    if (mode == DIRAL_MY_STEP_CH && h->bufs.lat) h->lat_live = true;
    p.track_lat = (h->bufs.lat && (h->lat_live || h->force_track_lat)) ? 1 : 0;

    float *obs_dst = h_obs ? h_obs : h->h_obs_stage;
    diral::HostJob job{};
    job.actions = h_actions; job.counts = h->h_counts; job.count_stride = rec;
    job.rews = reinterpret_cast<const float *>(h->h_counts + (rec - 4)); job.rew_stride = rec; job.rews_out = h_rews;
    job.obs = obs_dst;
    job.pos_x = h->h_kin; job.pos_y = h->h_kin ? h->h_kin + A : nullptr; job.vel = h->h_kin ? h->h_kin + 2 * A : nullptr;
    job.episode = episode; job.epsilon = epsilon; job.out = h_state;
    long long bounds[MAX_HOST_CHUNKS + 1];
    for (int k = 0; k <= chunks; ++k) bounds[k] = (streamed ? std::min<long long>(E, chunk_envs * k) : E * k / chunks) * N;
    diral::HostLayout lay = host_layout(c);
    // a worker's share of the rows stays in its private L2 from call to call when it is small enough
    lay.nt_stores = h->host_nt >= 0 ? h->host_nt
                                    : ((size_t)A * lay.S * sizeof(float) / (size_t)h->pool->threads() > (size_t)(1 << 20) + (1 << 19));
    unsigned epoch = 0;
    if (streamed) epoch = ++h->chunk_epoch ? h->chunk_epoch : ++h->chunk_epoch;         // never 0: flags start there
    // (asynchronous begin: the assembly threads watch the device-raised flags themselves)
    h->job_id = h->pool->begin(lay, job, bounds, chunks, async_begin ? h->h_chunk_flag : nullptr, epoch);
    PoolJobGuard guard{h->pool, h->job_id, chunks};
    h->trace_us[0] = since();                                      // workers woken

    // Pinned caller memory is read in place by the kernels (unified addressing: one PCIe read of 4 N bytes per environment
    // at the start of its decision phase) -- no host-to-device copy to enqueue; pageable actions go through the staging copy.
    bool direct = false;
    if (h->actions_direct == 2 || (h->actions_direct == 1 && streamed)) {
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, h_actions) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer) {
            p.actions = static_cast<const int32_t *>(attr.devicePointer); direct = true;
        }
        cudaGetLastError();                                        // (an unregistered pointer is not an error here)
    }

    if (streamed) {
        // One launch on the caller's stream.  The kernel writes every environment's records into mapped host memory and
        // raises a chunk's flag when its last environment is through (env_records_done); this thread forwards the flags
        // to the assembly workers.
        if (!direct) DIRAL_CUDA(cudaMemcpyAsync(h->d_actions, h_actions, (size_t)A * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        p.chunk_count = h->d_chunk_count; p.chunk_flag = h->d_chunk_flag; p.chunk_envs = (int)chunk_envs; p.chunk_epoch = epoch;
        if (h->stream_split) p.tail_split = 2;
        DIRAL_CUDA(launch_slot(h, p, s));
        h->launches += 1;
        if (c.add_piggy) h->ticks += 1;
        h->trace_us[1] = since();                                  // everything enqueued
        if (async_begin) {                                         // diral_step_host_wait picks it up from here
            guard.closed = true;
            h->async_pending = true; h->async_stream = s;
            h->trace_n = 2;
            return DIRAL_OK;
        }
        volatile unsigned *flags = h->h_chunk_flag;
        for (int k = 0; k < chunks; ++k) {
            for (unsigned spins = 1; flags[k] != epoch; ++spins) {
                _mm_pause();
                if ((spins & 0xfff) == 0) {                        // every few tens of microseconds: is the launch still alive?
                    const cudaError_t q = cudaStreamQuery(s);
                    if (q != cudaSuccess && q != cudaErrorNotReady)
                        return fail(DIRAL_ERR_CUDA, "slot kernel failed: %s", cudaGetErrorString(q));
                    if (q == cudaSuccess && flags[k] != epoch)
                        return fail(DIRAL_ERR_CUDA, "slot kernel finished without completing chunk %d", k);
                }
            }
            std::atomic_thread_fence(std::memory_order_acquire);
            h->pool->publish(h->job_id, k);
            h->trace_us[2 + k] = since();                          // chunk k's records are in host memory
        }
        guard.close();
        h->trace_us[2 + chunks] = since();                         // every row assembled
        h->trace_n = 3 + chunks;
        DIRAL_CUDA(cudaStreamSynchronize(s));                      // the launch itself retires (tables, accumulators)
        return DIRAL_OK;
    }

    // the chunk streams start after whatever the caller queued on its stream -- nothing to order when that stream is idle
    const bool caller_idle = cudaStreamQuery(s) == cudaSuccess;
    if (!caller_idle) { cudaGetLastError(); DIRAL_CUDA(cudaEventRecord(h->pipe_ev[2], s)); }
    int published = 0;
    for (int k = 0; k < chunks; ++k) {
        const long long a0 = bounds[k], n = bounds[k + 1] - a0, e0 = a0 / N;
        cudaStream_t ps = h->chunk_stream[k];
        if (!caller_idle) DIRAL_CUDA(cudaStreamWaitEvent(ps, h->pipe_ev[2], 0));
        if (!direct) DIRAL_CUDA(cudaMemcpyAsync(h->d_actions + a0, h_actions + a0, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ps));
        const diral::Params q = env_range(p, e0, n / N);
        DIRAL_CUDA(launch_slot(h, q, ps));
        h->launches += 1;
        if (!zero_copy)
            DIRAL_CUDA(cudaMemcpyAsync(h->h_counts + a0 * rec, h->d_counts + a0 * rec, (size_t)(n * rec), cudaMemcpyDeviceToHost, ps));
        if (h_obs || want_obs)
            DIRAL_CUDA(cudaMemcpyAsync(obs_dst + a0 * R, h->bufs.obs + a0 * R, (size_t)(n * R) * sizeof(float), cudaMemcpyDeviceToHost, ps));
        if (c.add_position) {
            DIRAL_CUDA(cudaMemcpyAsync(h->h_kin + a0, h->bufs.pos_x + a0, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ps));
            DIRAL_CUDA(cudaMemcpyAsync(h->h_kin + A + a0, h->bufs.pos_y + a0, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ps));
        }
        if (c.add_velocity)
            DIRAL_CUDA(cudaMemcpyAsync(h->h_kin + 2 * A + a0, h->bufs.vel + a0, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ps));
        DIRAL_CUDA(cudaEventRecord(h->chunk_ev[k], ps));
        // earlier chunks may have landed while this one was being enqueued: release their rows now
        while (published < k && cudaEventQuery(h->chunk_ev[published]) == cudaSuccess) {
            h->pool->publish(h->job_id, published);
            h->trace_us[2 + published++] = since();
        }
        cudaGetLastError();                                        // (cudaErrorNotReady is not an error)
    }
    // (the caller's stream needs no wait on the chunk streams: this call returns only after every chunk event has been
    //  synchronised on the host, so whatever the caller enqueues next is ordered after the whole slot)
    if (c.add_piggy) h->ticks += 1;
    h->trace_us[1] = since();                                      // everything enqueued
    for (int k = published; k < chunks; ++k) {                     // rows of chunk k are assembled while k+1.. are in flight
        DIRAL_CUDA(cudaEventSynchronize(h->chunk_ev[k]));
        h->pool->publish(h->job_id, k);
        h->trace_us[2 + k] = since();                              // chunk k's records are in host memory
    }
    guard.close();
    h->trace_us[2 + chunks] = since();                             // every row assembled
    h->trace_n = 3 + chunks;
    return DIRAL_OK;
}

}  // namespace

extern "C" {

int32_t diral_abi_version(void) { return DIRAL_ABI_VERSION; }

const char *diral_last_error(void) { return g_last_error.c_str(); }

int32_t diral_state_space(const diral_cfg *cfg) { return cfg ? state_space(*cfg) : 0; }

size_t diral_state_bytes(const diral_cfg *c)
{
    if (!c) return 0;
    const size_t E = (size_t)c->E, N = (size_t)c->N, NN = N * N;
    return E * N * 8 * 3 + E * NN * (4 + 4 + 8 + 4) + E * N * 4 * ((size_t)c->R + 1 + (size_t)state_space(*c))
         + E * (8 + 8 * diral::ACC_COUNTS) + diral_scratch_bytes(c);
}

size_t diral_scratch_bytes(const diral_cfg *c)
{
    if (!c || !c->add_piggy) return 0;
    diral::Params p{};
    p.N = c->N; p.R = c->R; p.B = c->B; p.vpd_enabled = c->add_piggy && (c->mobility || c->design_topology);
    // the group kernel keeps keys in registers; the block kernel needs scratch only past its smem budget
    if (diral::step_block_keys_fit_smem(p)) return 0;
    return diral::step_block_scratch_bytes(c->E, c->N);
}

int diral_create(const diral_cfg *cfg, void **handle)
{
    if (!handle) return fail(DIRAL_ERR_ARG, "handle out-pointer is NULL");
    *handle = nullptr;
    if (int rc = check_cfg(cfg)) return rc;
    int dev = 0;
    DIRAL_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    DIRAL_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10)
        return fail(DIRAL_ERR_CUDA, "device %d is sm_%d%d; libdiral_env.so carries sm_100a code only", dev, prop.major, prop.minor);
    Handle *h = new (std::nothrow) Handle();
    if (!h) return fail(DIRAL_ERR_ARG, "out of host memory");
    h->cfg = *cfg; h->device = dev;
    if (const char *v = getenv("DIRAL_HOST_NT")) h->host_nt = atoi(v) < 0 ? -1 : (atoi(v) != 0);     // measurement knob
    std::vector<double> edges(2 * (cfg->B + 1));
    np_linspace(-cfg->W, cfg->W, cfg->B + 1, edges.data());
    np_linspace(-1.0, 1.0, cfg->B + 1, edges.data() + cfg->B + 1);
    cudaError_t err = cudaMalloc(&h->d_edges, sizeof(double) * edges.size());
    if (err == cudaSuccess) err = cudaMemcpy(h->d_edges, edges.data(), sizeof(double) * edges.size(), cudaMemcpyHostToDevice);
    if (err != cudaSuccess) { delete h; return fail(DIRAL_ERR_CUDA, "edge table upload: %s", cudaGetErrorString(err)); }
    fill_base(h);
    // make the (possibly > 48 KB) dynamic shared memory carve-ups legal once, up front.  The lane-group kernel
    // keeps one observation row and one merge-script row per resource in shared memory: when that carve-up does
    // not fit (R beyond ~1300 at 32 vehicles) the one-CTA-per-env kernel, which chunks resources, takes over.
    diral::Params q = h->base; q.build_state = 1;
    bool group_fits = cfg->N <= diral::GROUP_MAX_N && diral::step_group_smem_bytes(q) <= (size_t)prop.sharedMemPerBlockOptin;
    const size_t need = diral::step_block_smem_bytes(q, diral::step_block_keys_fit_smem(q));
    const bool block_fits = need <= (size_t)prop.sharedMemPerBlockOptin;
    if (!group_fits && !block_fits) {
        cudaFree(h->d_edges); delete h;
        return fail(DIRAL_ERR_UNSUPPORTED, "N=%d R=%d B=%d needs %zu B of shared memory per CTA (limit %zu)",
                    cfg->N, cfg->R, cfg->B, need, (size_t)prop.sharedMemPerBlockOptin);
    }
    if (group_fits) err = diral::prepare_step_group(h->base);
    else h->variant = VARIANT_BLOCK;
    if (err == cudaSuccess && block_fits) err = diral::prepare_step_block(h->base);
    h->group_ok = group_fits; h->block_ok = block_fits;
    h->row_ok = diral::step_row_supported(h->base) && diral::step_row_smem_bytes(q) <= (size_t)prop.sharedMemPerBlockOptin;
    if (err == cudaSuccess && h->row_ok) err = diral::prepare_step_row(h->base);
    h->pair_ok = diral::step_pair_supported(h->base) && diral::step_pair_smem_bytes(h->base) <= (size_t)prop.sharedMemPerBlockOptin;
    if (err == cudaSuccess && h->pair_ok) err = diral::prepare_step_pair(h->base);
    set_layout(h);
    if (err != cudaSuccess) { cudaFree(h->d_edges); delete h; return fail(DIRAL_ERR_CUDA, "kernel attribute setup: %s", cudaGetErrorString(err)); }
    *handle = h;
    return DIRAL_OK;
}

int diral_destroy(void *handle)
{
    Handle *h = as_handle(handle);
    if (!h) return DIRAL_OK;
    if (h->async_pending) diral_step_host_wait(handle);        // a slot in flight still reads and writes what is freed below
    DeviceGuard g(h->device);
    cudaFree(h->d_edges);
    cudaFree(h->d_actions);
    for (auto &st : h->pipe) if (st) cudaStreamDestroy(st);
    for (auto &ev : h->pipe_ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : h->chunk_ev) if (ev) cudaEventDestroy(ev);
    for (auto &st : h->chunk_stream) if (st) cudaStreamDestroy(st);
    if (!h->pool_shared) delete h->pool;
    cudaFree(h->d_counts);
    cudaFree(h->d_chunk_count);
    if (h->h_chunk_flag) cudaFreeHost(h->h_chunk_flag);
    if (h->h_counts) cudaFreeHost(h->h_counts);
    if (h->h_obs_stage) cudaFreeHost(h->h_obs_stage);
    if (h->h_kin) cudaFreeHost(h->h_kin);
    delete h;
    return DIRAL_OK;
}

int diral_set_option(void *handle, const char *name, int64_t value)
{
    Handle *h = as_handle(handle);
    if (!h || !name) return fail(DIRAL_ERR_ARG, "handle/name is NULL");
    if (!strcmp(name, "variant")) {
        if (value < 0 || value > 5)
            return fail(DIRAL_ERR_ARG, "variant must be 0 (auto), 1 (group), 2 (block), 3 (round-1 block kernel), 4 (row layout) or 5 (pair)");
        if (value == VARIANT_PAIR && !h->pair_ok)
            return fail(DIRAL_ERR_UNSUPPORTED, "the two-rows-per-lane kernel takes 33..64 vehicles");
        if (value == VARIANT_ROW && !h->row_ok)
            return fail(DIRAL_ERR_UNSUPPORTED, "the row-layout kernel takes 33..256 vehicles with neighbour tables and a fused State block");
        if (h->bound) return fail(DIRAL_ERR_ARG, "the kernel variant fixes the table layout: choose it before diral_bind()");
        if (value == VARIANT_GROUP && h->cfg.N > diral::GROUP_MAX_N)
            return fail(DIRAL_ERR_ARG, "the group kernel handles num_users <= %d", diral::GROUP_MAX_N);
        if (value == VARIANT_GROUP && !h->group_ok)
            return fail(DIRAL_ERR_UNSUPPORTED, "the group kernel's shared-memory carve-up does not fit at R=%d", h->cfg.R);
        if ((value == VARIANT_BLOCK && !h->block_ok && !h->row_ok && !h->pair_ok) || (value == VARIANT_BLOCK_V1 && !h->block_ok))
            return fail(DIRAL_ERR_UNSUPPORTED, "the one-CTA-per-env kernel's shared-memory carve-up does not fit this configuration");
        if (value == VARIANT_AUTO && !h->group_ok) value = VARIANT_BLOCK;
        h->variant = (int)value;
        set_layout(h);
        return DIRAL_OK;
    }
    if (!strcmp(name, "track_lat")) { h->force_track_lat = value != 0; return DIRAL_OK; }
    if (!strcmp(name, "host_format")) {
        if (value < 0 || value > 3)
            return fail(DIRAL_ERR_ARG, "host_format must be 0 (full rows), 1 (compact), 2 (compact, zero-copy records) or 3 (streamed records)");
        h->host_format = (int)value;
        return DIRAL_OK;
    }
    if (!strcmp(name, "host_threads")) {
        if (value < 0 || value > 256) return fail(DIRAL_ERR_ARG, "host_threads must be in [0, 256]");
        if (h->pool && !h->pool_shared && h->pool->threads() != (int)value) { delete h->pool; h->pool = nullptr; }
        h->host_threads = (int)value;
        return DIRAL_OK;
    }
    if (!strcmp(name, "tail_split")) {
        if (value < 0 || value > 2) return fail(DIRAL_ERR_ARG, "tail_split must be 0 (off), 1 (auto) or 2 (always)");
        h->tail_split = (int)value; h->base.tail_split = (int)value;
        return DIRAL_OK;
    }
    if (!strcmp(name, "host_nt")) {
        if (value < -1 || value > 1) return fail(DIRAL_ERR_ARG, "host_nt must be -1 (auto), 0 or 1");
        h->host_nt = (int)value;
        return DIRAL_OK;
    }
    if (!strcmp(name, "host_chunks")) {
        if (value < 1 || value > 32) return fail(DIRAL_ERR_ARG, "host_chunks must be in [1, 32]");
        h->host_chunks = (int)value;
        return DIRAL_OK;
    }
    if (!strcmp(name, "host_pool_shared")) {
        if (h->pool && (value != 0) != h->pool_shared) { if (!h->pool_shared) delete h->pool; h->pool = nullptr; }
        h->want_shared_pool = value != 0;
        return DIRAL_OK;
    }
    if (!strcmp(name, "stream_chunks")) {
        if (value < 1 || value > MAX_HOST_CHUNKS) return fail(DIRAL_ERR_ARG, "stream_chunks must be in [1, %d]", MAX_HOST_CHUNKS);
        h->stream_chunks = (int)value;
        return DIRAL_OK;
    }
    if (!strcmp(name, "stream_split")) { h->stream_split = value != 0; return DIRAL_OK; }
    if (!strcmp(name, "actions_direct")) {
        if (value < 0 || value > 2) return fail(DIRAL_ERR_ARG, "actions_direct must be 0 (never), 1 (streamed format only) or 2 (always)");
        h->actions_direct = (int)value;
        return DIRAL_OK;
    }
    // checkpoint restore (TestEnv.load_state_dict): the slot counters the kernels derive keys and stamps from
    if (!strcmp(name, "ticks")) {
        if (value < 0) return fail(DIRAL_ERR_ARG, "ticks must be >= 0");
        h->ticks = value;
        return DIRAL_OK;
    }
    if (!strcmp(name, "lat_live")) { h->lat_live = value != 0; return DIRAL_OK; }
    return fail(DIRAL_ERR_ARG, "unknown option '%s'", name);
}

int64_t diral_get_option(void *handle, const char *name)
{
    Handle *h = as_handle(handle);
    if (!h || !name) return -1;
    if (!strcmp(name, "variant")) return use_group(h) ? VARIANT_GROUP : VARIANT_BLOCK;
    if (!strcmp(name, "kernel")) return use_group(h) ? 1 : (use_pair(h) ? 4 : (use_row(h) ? 3 : 2));   // 1 lane-group, 2 round-1 block, 3 row layout, 4 pair
    if (!strcmp(name, "layout")) return h->base.layout;
    if (!strcmp(name, "row_stride")) return h->base.T;
    if (!strcmp(name, "ring_depth")) return h->base.H;
    if (!strcmp(name, "scratch_bytes")) return (int64_t)handle_scratch_bytes(h);
    if (!strcmp(name, "track_lat")) return h->force_track_lat;
    if (!strcmp(name, "host_format")) return h->host_format;
    if (!strcmp(name, "host_threads")) return h->pool ? h->pool->threads() : h->host_threads;
    if (!strcmp(name, "host_chunks")) return h->host_chunks;
    if (!strcmp(name, "stream_chunks")) return h->stream_chunks;
    if (!strcmp(name, "stream_split")) return h->stream_split;
    if (!strcmp(name, "host_pool_shared")) return h->want_shared_pool;
    if (!strcmp(name, "actions_direct")) return h->actions_direct;
    if (!strcmp(name, "host_nt")) return h->host_nt;
    if (!strcmp(name, "tail_split")) return h->tail_split;
    if (!strcmp(name, "ticks")) return h->ticks;
    if (!strcmp(name, "lat_live")) return h->lat_live ? 1 : 0;
    if (!strcmp(name, "compact_ok")) return compact_ok(h->cfg) ? 1 : 0;
    return -1;
}

int diral_bind(void *handle, const diral_buffers *b)
{
    Handle *h = as_handle(handle);
    if (!h || !b) return fail(DIRAL_ERR_ARG, "handle/buffers is NULL");
    if (!b->pos_x || !b->pos_y || !b->vel || !b->obs || !b->rews || !b->state || !b->acc_reward || !b->acc_count)
        return fail(DIRAL_ERR_ARG, "pos_x/pos_y/vel/obs/rews/state/acc_* must all be bound");
    if (h->cfg.add_piggy && (!b->tab_seq || !b->tab_lu || !b->tab_x))
        return fail(DIRAL_ERR_ARG, "add_positional_dist_piggy needs tab_seq/tab_lu/tab_x");
    if (handle_scratch_bytes(h) > 0 && !b->scratch)
        return fail(DIRAL_ERR_ARG, "this configuration needs %zu B of scratch", handle_scratch_bytes(h));
    if (h->base.layout && !b->ring) return fail(DIRAL_ERR_ARG, "the row layout needs the position ring (diral_buffers.ring)");
    if (b->trace && b->trace_len < 1) return fail(DIRAL_ERR_ARG, "trace_len must be >= 1 when a trace is bound");
    h->bufs = *b;
    bind_params(h);
    h->bound = true;
    return DIRAL_OK;
}

int diral_reset(void *handle, const double *x0, const double *y0, const double *v0, uint64_t seed, void *stream)
{
    Handle *h = as_handle(handle);
    if (int rc = require_bound(h)) return rc;
    if (x0 && (!y0 || !v0)) return fail(DIRAL_ERR_ARG, "x0, y0 and v0 must be given together");
    DeviceGuard g(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const diral_cfg &c = h->cfg;
    const size_t EN = (size_t)c.E * c.N, ENN = EN * c.N;
    if (int rc = zero_tables(h, s)) return rc;
    if (h->bufs.lat) DIRAL_CUDA(cudaMemsetAsync(h->bufs.lat, 0xFF, ENN * 4, s));           // network.py:39-42
    DIRAL_CUDA(cudaMemsetAsync(h->bufs.acc_reward, 0, (size_t)c.E * 8, s));
    DIRAL_CUDA(cudaMemsetAsync(h->bufs.acc_count, 0, (size_t)c.E * 8 * diral::ACC_COUNTS, s));
    diral::Params p = h->base;
    p.seed = seed;
    DIRAL_CUDA(diral::launch_reset(p, x0, y0, v0, s));
    h->launches += 1;
    h->ticks = 0; h->lat_live = false;
    return DIRAL_OK;
}

int diral_reset_topology(void *handle, const double *x0, const double *y0, const double *v0, uint64_t seed, void *stream)
{
    Handle *h = as_handle(handle);
    if (int rc = require_bound(h)) return rc;
    if (x0 && (!y0 || !v0)) return fail(DIRAL_ERR_ARG, "x0, y0 and v0 must be given together");
    DeviceGuard g(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (int rc = zero_tables(h, s)) return rc;
    diral::Params p = h->base;
    p.seed = seed;
    DIRAL_CUDA(diral::launch_reset(p, x0, y0, v0, s));
    h->launches += 1;
    h->ticks = 0;
    return DIRAL_OK;
}

int diral_sample(void *handle, uint64_t seed, int64_t t, int32_t *out, void *stream)
{
    Handle *h = as_handle(handle);
    if (int rc = require_bound(h)) return rc;
    if (!out) return fail(DIRAL_ERR_ARG, "out is NULL");
    DeviceGuard g(h->device);
    diral::Params p = h->base;
    p.seed = seed; p.timestep = t;
    DIRAL_CUDA(diral::launch_sample(p, out, static_cast<cudaStream_t>(stream)));
    h->launches += 1;
    return DIRAL_OK;
}

int diral_obtain_state(void *handle, const float *obs, const int32_t *actions, const float *rews, double episode,
                       double epsilon, float *out, void *stream)
{
    Handle *h = as_handle(handle);
    if (int rc = require_bound(h)) return rc;
    if (!obs || !actions || !rews || !out) return fail(DIRAL_ERR_ARG, "obs/actions/rews/out must not be NULL");
    DeviceGuard g(h->device);
    diral::Params p = h->base;
    p.episode = episode; p.epsilon = epsilon; p.tick = (int)h->ticks;
    DIRAL_CUDA(diral::launch_obtain_state(p, obs, actions, rews, out, static_cast<cudaStream_t>(stream)));
    h->launches += 1;
    return DIRAL_OK;
}

int diral_step(void *handle, int mode, const int32_t *actions, int64_t timestep, int build_state, double episode,
               double epsilon, uint64_t seed, int32_t *actions_out, void *stream)
{
    Handle *h = as_handle(handle);
    if (int rc = require_bound(h)) return rc;
    if (mode < DIRAL_MY_STEP || mode > DIRAL_MY_STEP_CH) return fail(DIRAL_ERR_ARG, "mode must be 0, 1 or 2 (got %d)", mode);
    const bool group = use_group(h);
    const int src_bits = group ? diral::key_src_bits(diral::group_width(h->cfg.N)) : diral::key_src_bits(h->cfg.N);
    if (h->cfg.add_piggy && h->ticks + 1 >= (1ll << (32 - src_bits)))
        return fail(DIRAL_ERR_SEQ_RANGE, "slot %lld since reset exceeds the %d-bit sequence field of the packed table keys",
                    h->ticks + 1, 32 - src_bits);
    DeviceGuard g(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool fused = build_state && fused_state_ok(h->cfg);
    diral::Params p = h->base;
    p.mode = mode; p.timestep = timestep; p.episode = episode; p.epsilon = epsilon; p.seed = seed;
    p.tick = (int)(h->ticks + 1);
    p.build_state = fused ? 1 : 0;
    p.actions = actions; p.gen_actions = actions == nullptr; p.actions_out = actions_out;
    if (mode == DIRAL_MY_STEP_CH && h->bufs.lat) h->lat_live = true;
    p.track_lat = (h->bufs.lat && (h->lat_live || h->force_track_lat)) ? 1 : 0;
    if (build_state && !fused && !actions && !actions_out) {
        if (int rc = ensure_actions_staging(h)) return rc;
        p.actions_out = h->d_actions;
    }
    DIRAL_CUDA(launch_slot(h, p, s));
    h->launches += 1;
    if (h->cfg.add_piggy) h->ticks += 1;
    if (build_state && !fused) {
        const int32_t *acts = actions ? actions : p.actions_out;
        DIRAL_CUDA(diral::launch_obtain_state(p, h->bufs.obs, acts, h->bufs.rews, h->bufs.state, s));
        h->launches += 1;
    }
    return DIRAL_OK;
}

int diral_rollout(void *handle, int mode, int32_t T, int64_t t0, uint64_t seed, void *stream)
{
    Handle *h = as_handle(handle);
    if (int rc = require_bound(h)) return rc;
    if (mode < DIRAL_MY_STEP || mode > DIRAL_MY_STEP_CH) return fail(DIRAL_ERR_ARG, "mode must be 0, 1 or 2 (got %d)", mode);
    if (T < 1) return fail(DIRAL_ERR_ARG, "T must be >= 1 (got %d)", T);
    // The lane-group kernel (N <= 32) runs all T slots of an environment in ONE launch: small configurations are
    // launch- and latency-bound per slot, and no slot needs anything another environment wrote.
    if (use_group(h) && fused_state_ok(h->cfg) && T > 1) {
        const int src_bits = diral::key_src_bits(diral::group_width(h->cfg.N));
        if (h->cfg.add_piggy && h->ticks + T >= (1ll << (32 - src_bits)))
            return fail(DIRAL_ERR_SEQ_RANGE, "slot %lld since reset exceeds the %d-bit sequence field of the packed table keys",
                        h->ticks + T, 32 - src_bits);
        DeviceGuard g(h->device);
        diral::Params p = h->base;
        p.mode = mode; p.timestep = t0; p.episode = 0.0; p.epsilon = 1.0; p.seed = seed;
        p.tick = (int)(h->ticks + 1); p.n_slots = T;
        p.build_state = 1; p.actions = nullptr; p.gen_actions = 1; p.actions_out = nullptr;
        if (mode == DIRAL_MY_STEP_CH && h->bufs.lat) h->lat_live = true;
        p.track_lat = (h->bufs.lat && (h->lat_live || h->force_track_lat)) ? 1 : 0;
        DIRAL_CUDA(diral::launch_step_group(p, static_cast<cudaStream_t>(stream)));
        h->launches += 1;
        if (h->cfg.add_piggy) h->ticks += T;
        return DIRAL_OK;
    }
    for (int32_t k = 0; k < T; ++k)
        if (int rc = diral_step(handle, mode, nullptr, t0 + k, 1, 0.0, 1.0, seed, nullptr, stream)) return rc;
    return DIRAL_OK;
}

int diral_update_velocity(void *handle, const int8_t *draws, uint64_t seed, int64_t episode, void *stream)
{
    Handle *h = as_handle(handle);
    if (int rc = require_bound(h)) return rc;
    if (!h->cfg.mobility_vary) return DIRAL_OK;                         // test_env.py:498-504
    DeviceGuard g(h->device);
    diral::Params p = h->base;
    p.seed = seed;
    DIRAL_CUDA(diral::launch_update_velocity(p, h->bufs.vel, draws, episode, static_cast<cudaStream_t>(stream)));
    h->launches += 1;
    return DIRAL_OK;
}

int diral_information_age(void *handle, int64_t timestep, int32_t *out, void *stream)
{
    Handle *h = as_handle(handle);
    if (int rc = require_bound(h)) return rc;
    if (!out) return fail(DIRAL_ERR_ARG, "out is NULL");
    if (!h->bufs.lat) return fail(DIRAL_ERR_ARG, "last_arrival_time (lat) is not bound");
    DeviceGuard g(h->device);
    diral::Params p = h->base;
    p.timestep = timestep;
    DIRAL_CUDA(diral::launch_information_age(p, out, static_cast<cudaStream_t>(stream)));
    h->launches += 1;
    return DIRAL_OK;
}

int diral_episode_metrics(void *handle, int64_t timestep, double *out110, void *stream)
{
    Handle *h = as_handle(handle);
    if (int rc = require_bound(h)) return rc;
    if (!out110) return fail(DIRAL_ERR_ARG, "out110 is NULL");
    DeviceGuard g(h->device);
    diral::Params p = h->base;
    p.timestep = timestep;
    p.track_lat = (h->bufs.lat && (h->lat_live || h->force_track_lat)) ? 1 : 0;
    DIRAL_CUDA(diral::launch_episode_metrics(p, out110, static_cast<cudaStream_t>(stream)));
    h->launches += 1;
    return DIRAL_OK;
}

int diral_shape_rewards(void *handle, const diral_shaping *cfg, const int32_t *actions, int64_t timestep, float *rewards,
                        int64_t *sum_ia_prev, int32_t *ia_counter, int32_t *prev_actions, double *slot_sums,
                        int32_t *ia_out, void *stream)
{
    Handle *h = as_handle(handle);
    if (int rc = require_bound(h)) return rc;
    if (!cfg || !rewards) return fail(DIRAL_ERR_ARG, "cfg/rewards must not be NULL");
    if (cfg->ia_averaging && !sum_ia_prev) return fail(DIRAL_ERR_ARG, "ia_averaging needs sum_ia_prev");
    if (cfg->ia_penalty_enable && (!actions || !ia_counter || !prev_actions))
        return fail(DIRAL_ERR_ARG, "ia_penalty_enable needs actions, ia_counter and prev_actions");
    DeviceGuard g(h->device);
    diral::Params p = h->base;
    p.timestep = timestep;
    p.track_lat = (h->bufs.lat && (h->lat_live || h->force_track_lat)) ? 1 : 0;
    diral::ShapingArgs a{};
    a.ia_averaging = cfg->ia_averaging != 0; a.ia_penalty_enable = cfg->ia_penalty_enable != 0;
    a.ia_penalty_threshold = cfg->ia_penalty_threshold; a.global_reward_avg = cfg->global_reward_avg != 0;
    a.ia_penalty_value = cfg->ia_penalty_value;
    a.actions = actions; a.rewards = rewards; a.sum_ia_prev = reinterpret_cast<long long *>(sum_ia_prev);
    a.ia_counter = ia_counter; a.prev_actions = prev_actions; a.slot_sums = slot_sums; a.ia_out = ia_out;
    DIRAL_CUDA(diral::launch_shape_rewards(p, a, static_cast<cudaStream_t>(stream)));
    h->launches += 1;
    return DIRAL_OK;
}

int diral_ring_gather(const void *ring, int64_t capacity, int64_t agents, int64_t width, int32_t elem_bytes,
                      const int64_t *start, int32_t batch, int32_t step, void *out, void *stream)
{
    if (!ring || !start || !out) return fail(DIRAL_ERR_ARG, "ring/start/out must not be NULL");
    if (capacity < 1 || agents < 1 || width < 1 || batch < 1 || step < 1 || elem_bytes < 1)
        return fail(DIRAL_ERR_ARG, "capacity/agents/width/batch/step/elem_bytes must be >= 1");
    DIRAL_CUDA(diral::launch_ring_gather(ring, capacity, agents, width, elem_bytes, reinterpret_cast<const long long *>(start),
                                         batch, step, out, static_cast<cudaStream_t>(stream)));
    return DIRAL_OK;
}

int diral_wire_vpd(const diral_wire_entry *tables, const int32_t *observer, int64_t M, int32_t N, int32_t pos_dist,
                   int32_t bins, double range, int32_t age_limit, float *out, void *stream)
{
    if (!tables || !observer || !out) return fail(DIRAL_ERR_ARG, "tables/observer/out must not be NULL");
    if (M < 1 || N < 1 || N > diral::wire_max_entries()) return fail(DIRAL_ERR_ARG, "M must be >= 1 and N in [1, %d]", diral::wire_max_entries());
    if (bins < 1 || bins > diral::wire_max_bins()) return fail(DIRAL_ERR_ARG, "state_bins must be in [1, %d] (got %d)", diral::wire_max_bins(), bins);
    if (pos_dist != 1 && pos_dist != 2) return fail(DIRAL_ERR_ARG, "pos_dist must be 1 or 2 (got %d)", pos_dist);
    if (pos_dist == 2 && !(range > 0.0)) return fail(DIRAL_ERR_ARG, "state_range must be > 0");
    std::vector<double> edges(bins + 1);
    if (pos_dist == 2) np_linspace(-range, range, bins + 1, edges.data());     // numpy.histogram(.., bins, range=(-W, W))
    else np_linspace(-1.0, 1.0, bins + 1, edges.data());                       // realness_env.py:77
    DIRAL_CUDA(diral::launch_wire_vpd(tables, observer, M, N, pos_dist, bins, range, age_limit, edges.data(), out,
                                      static_cast<cudaStream_t>(stream)));
    return DIRAL_OK;
}

int diral_sps_step(int64_t agents, int32_t window_len, const double *selection_window, const diral_sps_cfg *cfg,
                   const double *draws, uint64_t seed, int64_t t, int32_t *prev_action, int32_t *reselection_counter,
                   int32_t *actions, int32_t *flags, void *stream)
{
    if (!selection_window || !cfg || !prev_action || !reselection_counter || !actions)
        return fail(DIRAL_ERR_ARG, "selection_window/cfg/prev_action/reselection_counter/actions must not be NULL");
    if (agents < 1 || window_len < 1) return fail(DIRAL_ERR_ARG, "agents and window_len must be >= 1");
    DIRAL_CUDA(diral::launch_sps_step(agents, window_len, selection_window, cfg->rssi_threshold, cfg->inc_db, cfg->prob_resource_keep,
                                      cfg->min_candidates, draws, seed, t, prev_action, reselection_counter, actions, flags,
                                      static_cast<cudaStream_t>(stream)));
    return DIRAL_OK;
}

int diral_step_host(void *handle, int mode, const int32_t *h_actions, int64_t timestep, double episode, double epsilon,
                    float *h_state, float *h_rews, float *h_obs, void *stream)
{
    Handle *h = as_handle(handle);
    if (int rc = require_bound(h)) return rc;
    if (!h_actions || !h_state || !h_rews) return fail(DIRAL_ERR_ARG, "h_actions/h_state/h_rews must not be NULL");
    if (mode < DIRAL_MY_STEP || mode > DIRAL_MY_STEP_CH) return fail(DIRAL_ERR_ARG, "mode must be 0, 1 or 2 (got %d)", mode);
    if (int rc = ensure_actions_staging(h)) return rc;
    DeviceGuard g(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (h->host_format >= 1 && compact_ok(h->cfg))
        return step_host_compact(h, mode, h_actions, timestep, episode, epsilon, h_state, h_rews, h_obs, s);
    const long long E = h->cfg.E, N = h->cfg.N, R = h->cfg.R, S = h->base.S;
    // Envs are independent, so the batch is cut into chunks that alternate between two internal
    // streams: chunk c's results travel over PCIe while chunk c+1 computes and chunk c+2's actions arrive.
    const int chunks = (fused_state_ok(h->cfg) && E >= 1024) ? 4 : 1;
    if (chunks == 1) {
        DIRAL_CUDA(cudaMemcpyAsync(h->d_actions, h_actions, (size_t)(E * N) * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        if (int rc = diral_step(handle, mode, h->d_actions, timestep, 1, episode, epsilon, 0, nullptr, stream)) return rc;
        DIRAL_CUDA(cudaMemcpyAsync(h_state, h->bufs.state, (size_t)(E * N * S) * sizeof(float), cudaMemcpyDeviceToHost, s));
        DIRAL_CUDA(cudaMemcpyAsync(h_rews, h->bufs.rews, (size_t)(E * N) * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (h_obs) DIRAL_CUDA(cudaMemcpyAsync(h_obs, h->bufs.obs, (size_t)(E * N * R) * sizeof(float), cudaMemcpyDeviceToHost, s));
        DIRAL_CUDA(cudaStreamSynchronize(s));
        return DIRAL_OK;
    }
    const bool group = use_group(h);
    const int src_bits = group ? diral::key_src_bits(diral::group_width(h->cfg.N)) : diral::key_src_bits(h->cfg.N);
    if (h->cfg.add_piggy && h->ticks + 1 >= (1ll << (32 - src_bits)))
        return fail(DIRAL_ERR_SEQ_RANGE, "slot %lld since reset exceeds the %d-bit sequence field of the packed table keys",
                    h->ticks + 1, 32 - src_bits);
    for (auto &st : h->pipe) if (!st) DIRAL_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (auto &ev : h->pipe_ev) if (!ev) DIRAL_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    diral::Params p = h->base;
    p.mode = mode; p.timestep = timestep; p.episode = episode; p.epsilon = epsilon; p.seed = 0;
    p.tick = (int)(h->ticks + 1);
    p.build_state = 1; p.actions = h->d_actions; p.gen_actions = 0; p.actions_out = nullptr;
    if (mode == DIRAL_MY_STEP_CH && h->bufs.lat) h->lat_live = true;
    p.track_lat = (h->bufs.lat && (h->lat_live || h->force_track_lat)) ? 1 : 0;
    DIRAL_CUDA(cudaEventRecord(h->pipe_ev[2], s));                 // everything queued on the caller's stream so far
    for (int c = 0; c < chunks; ++c) {
        const long long e0 = E * c / chunks, n = E * (c + 1) / chunks - e0;
        cudaStream_t ps = h->pipe[c & 1];
        if (c < 2) DIRAL_CUDA(cudaStreamWaitEvent(ps, h->pipe_ev[2], 0));
        DIRAL_CUDA(cudaMemcpyAsync(h->d_actions + e0 * N, h_actions + e0 * N, (size_t)(n * N) * sizeof(int32_t),
                                   cudaMemcpyHostToDevice, ps));
        const diral::Params q = env_range(p, e0, n);
        DIRAL_CUDA(launch_slot(h, q, ps));
        h->launches += 1;
        DIRAL_CUDA(cudaMemcpyAsync(h_state + e0 * N * S, h->bufs.state + e0 * N * S, (size_t)(n * N * S) * sizeof(float),
                                   cudaMemcpyDeviceToHost, ps));
        DIRAL_CUDA(cudaMemcpyAsync(h_rews + e0 * N, h->bufs.rews + e0 * N, (size_t)(n * N) * sizeof(float),
                                   cudaMemcpyDeviceToHost, ps));
        if (h_obs) DIRAL_CUDA(cudaMemcpyAsync(h_obs + e0 * N * R, h->bufs.obs + e0 * N * R, (size_t)(n * N * R) * sizeof(float),
                                              cudaMemcpyDeviceToHost, ps));
    }
    if (h->cfg.add_piggy) h->ticks += 1;
    for (int k = 0; k < 2; ++k) {                                  // the caller's stream sees the step as done
        DIRAL_CUDA(cudaEventRecord(h->pipe_ev[k], h->pipe[k]));
        DIRAL_CUDA(cudaStreamWaitEvent(s, h->pipe_ev[k], 0));
    }
    DIRAL_CUDA(cudaStreamSynchronize(s));
    return DIRAL_OK;
}

int diral_step_host_begin(void *handle, int mode, const int32_t *h_actions, int64_t timestep, double episode, double epsilon,
                          float *h_state, float *h_rews, void *stream)
{
    Handle *h = as_handle(handle);
    if (int rc = require_bound(h)) return rc;
    if (!h_actions || !h_state || !h_rews) return fail(DIRAL_ERR_ARG, "h_actions/h_state/h_rews must not be NULL");
    if (mode < DIRAL_MY_STEP || mode > DIRAL_MY_STEP_CH) return fail(DIRAL_ERR_ARG, "mode must be 0, 1 or 2 (got %d)", mode);
    if (h->host_format != 3 || !compact_ok(h->cfg))
        return fail(DIRAL_ERR_UNSUPPORTED, "diral_step_host_begin needs host_format 3 and a State block with a compact host format");
    if (int rc = ensure_actions_staging(h)) return rc;
    DeviceGuard g(h->device);
    return step_host_compact(h, mode, h_actions, timestep, episode, epsilon, h_state, h_rews, nullptr,
                             static_cast<cudaStream_t>(stream), true);
}

int diral_step_host_wait(void *handle)
{
    Handle *h = as_handle(handle);
    if (!h) return fail(DIRAL_ERR_ARG, "handle is NULL");
    if (!h->async_pending) return DIRAL_OK;
    DeviceGuard g(h->device);
    h->async_pending = false;
    unsigned retired_polls = 0;                                    // polls since the launch was first seen retired
    for (unsigned spins = 1; !h->pool->done(h->job_id); ++spins) {
        _mm_pause();
        if ((spins & 0xfff) == 0) {                                // every few tens of microseconds: is the launch still alive?
            const cudaError_t q = cudaStreamQuery(h->async_stream);
            if (q == cudaErrorNotReady) { cudaGetLastError(); continue; }
            // retired: every flag is up, the rows follow within microseconds -- unless something is badly wrong
            if (q == cudaSuccess && ++retired_polls < 100000u) continue;
            h->pool->abort(h->job_id); h->pool->finish(h->job_id);
            if (q == cudaSuccess) return fail(DIRAL_ERR_CUDA, "slot kernel retired, but the row assembly did not complete");
            return fail(DIRAL_ERR_CUDA, "slot kernel failed: %s", cudaGetErrorString(q));
        }
    }
    h->pool->finish(h->job_id);
    h->pool->timeline(h->job_id, h->trace_us + 2);                 // [2..4]: first / last chunk released, rows written
    h->trace_n = 5;
    // the launch itself retires (tables, accumulators): it usually has by now -- the last rows were assembled after its
    // last flag -- so ask before blocking
    if (cudaStreamQuery(h->async_stream) != cudaSuccess) {
        cudaGetLastError();
        DIRAL_CUDA(cudaStreamSynchronize(h->async_stream));
    }
    return DIRAL_OK;
}

int diral_expand_state_host(const diral_cfg *cfg, int64_t agents, const int32_t *actions, const uint8_t *counts,
                            const float *rews, const float *obs, const double *pos_x, const double *pos_y,
                            const double *vel, double episode, double epsilon, int32_t threads, float *out)
{
    if (int rc = check_cfg(cfg)) return rc;
    if (!compact_ok(*cfg)) return fail(DIRAL_ERR_UNSUPPORTED, "this State block has no compact host format");
    if (!actions || !out || agents < 0) return fail(DIRAL_ERR_ARG, "actions/out must not be NULL");
    if (cfg->add_piggy && !counts) return fail(DIRAL_ERR_ARG, "add_positional_dist_piggy needs counts");
    if (cfg->add_reward && !rews) return fail(DIRAL_ERR_ARG, "add_reward needs rews");
    if (cfg->add_channel_obs && !obs) return fail(DIRAL_ERR_ARG, "add_channel_obs needs obs");
    if (cfg->add_position && (!pos_x || !pos_y)) return fail(DIRAL_ERR_ARG, "add_position needs pos_x and pos_y");
    if (cfg->add_velocity && !vel) return fail(DIRAL_ERR_ARG, "add_velocity needs vel");
    diral::HostJob job{};
    job.actions = actions; job.counts = counts; job.count_stride = cfg->B; job.rews = rews; job.rew_stride = sizeof(float);
    job.rews_out = nullptr; job.obs = obs; job.pos_x = pos_x; job.pos_y = pos_y;
    job.vel = vel; job.episode = episode; job.epsilon = epsilon; job.out = out;
    diral::HostLayout lay = host_layout(*cfg);
    lay.nt_stores = threads < 0;             // (bench knob: a negative thread count selects non-temporal stores)
    if (threads < 0) threads = -threads;
    if (threads <= 1) { diral::expand_rows(lay, job, 0, agents); return DIRAL_OK; }
    // one pool per thread count, kept for the life of the process (this entry point has no handle to own it)
    static std::mutex mu;
    static std::vector<std::unique_ptr<diral::HostPool>> pools;
    std::lock_guard<std::mutex> lock(mu);
    diral::HostPool *pool = nullptr;
    bool pool_shared = false;       // `pool` is the process-wide one (not owned)
    int want_shared_pool = 0;       // option "host_pool_shared"
    unsigned long long job_id = 0;  // the pool job of the slot in flight
    for (auto &q : pools) if (q->threads() == threads) pool = q.get();
    if (!pool) { pools.emplace_back(new diral::HostPool(threads)); pool = pools.back().get(); }
    const long long bounds[3] = {0, (agents / 2) & ~3ll, agents};      // two chunks: exercises the chunk hand-over too
    const unsigned long long id = pool->begin(lay, job, bounds, 2);
    pool->publish(id, 0); pool->publish(id, 1);
    pool->finish(id);
    return DIRAL_OK;
}

int32_t diral_host_trace(void *handle, double *out_us, int32_t n)
{
    Handle *h = as_handle(handle);
    if (!h || !out_us) return 0;
    const int m = std::min<int>(n, h->trace_n);
    for (int i = 0; i < m; ++i) out_us[i] = h->trace_us[i];
    return m;
}

int diral_materialize_x(void *handle, double *out, void *stream)
{
    Handle *h = as_handle(handle);
    if (int rc = require_bound(h)) return rc;
    if (!out) return fail(DIRAL_ERR_ARG, "out is NULL");
    if (!h->cfg.add_piggy) return fail(DIRAL_ERR_ARG, "this configuration keeps no neighbour tables");
    DeviceGuard g(h->device);
    diral::Params p = h->base;
    p.tick = (int)h->ticks;
    DIRAL_CUDA(diral::launch_materialize_x(p, out, static_cast<cudaStream_t>(stream)));
    h->launches += 1;
    return DIRAL_OK;
}

int diral_ring_put(void *ring, int64_t capacity, int64_t slot, int64_t row_bytes, const void *src, void *stream)
{
    if (!ring || !src) return fail(DIRAL_ERR_ARG, "ring/src must not be NULL");
    if (capacity < 1 || slot < 0 || slot >= capacity || row_bytes < 1)
        return fail(DIRAL_ERR_ARG, "slot must be in [0, capacity) and row_bytes >= 1");
    DIRAL_CUDA(cudaMemcpyAsync(static_cast<char *>(ring) + slot * row_bytes, src, (size_t)row_bytes, cudaMemcpyDeviceToDevice,
                               static_cast<cudaStream_t>(stream)));
    return DIRAL_OK;
}

int64_t diral_launch_count(void *handle)
{
    Handle *h = as_handle(handle);
    return h ? h->launches : 0;
}

}  // extern "C"
