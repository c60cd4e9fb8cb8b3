// diral_host.h -- host side of the compact host format of diral_step_host (internal to libdiral_env.so).
//
// TestEnv.obtain_state (reference envs/test_env.py:527-583) emits, per agent, blocks that are either known to the
// host already (the one-hot of the action it just sent, the agent index, episode / epsilon) or a few bytes of
// information (B small integers behind the B float32 of the positional distribution).  In the compact format only
// the information crosses PCIe -- bin counts as one byte each, rewards, and the float64 positions / velocities /
// channel observations when the State block asks for them -- and the [E][N][S] float32 rows the caller expects are
// assembled here by a pool of host threads, bit for bit what the device writes into bufs.state.
#pragma once
#include <cstdint>

namespace diral {

struct HostLayout {                 // the State block (test_env.py:27-41) as the expander needs it
    int N, R, B, S;
    int add_action, action_binary, add_channel_obs, piggy, add_reward, add_index, add_position, add_velocity, fingerprint;
    double L;
    int nt_stores;                  // non-temporal stores for the output rows (else ordinary, cache-resident ones)
};

struct HostJob {                    // one call: per-agent inputs (index a = env * N + vehicle) and the output rows
    const int32_t *actions;         // [A]          what the caller sent (clamped to [0, R) like the kernels do)
    const uint8_t *counts;          // [A] x count_stride bytes: B VPD bin counts per agent (piggy)
    const float *rews;              // [A] x rew_stride bytes: rewards
    long long count_stride, rew_stride;
    float *rews_out;                // [A] or NULL: rewards copied out densely (the caller's h_rews)
    const float *obs;               // [A][R]       (add_channel_obs)
    const double *pos_x, *pos_y;    // [A]          post-mobility x, y (add_position)
    const double *vel;              // [A]          (add_velocity)
    double episode, epsilon;
    float *out;                     // [A][S]
};

// rows [a0, a1) of one job, on the calling thread
void expand_rows(const HostLayout &lay, const HostJob &job, long long a0, long long a1);

// Persistent worker threads with a small FIFO of jobs, so that several handles (groups of environments stepped with
// diral_step_host_begin / _wait) can share one pool: begin() queues a job split into `nchunks` consecutive agent ranges
// [bounds[c], bounds[c+1]) and returns its id; a chunk is released either by publish(id, c) ("chunk c's inputs have
// landed in host memory") or, when `flags` is given, by the device itself (flags[c] == epoch: words the slot kernel
// raises in mapped host memory); finish(id) waits until every row of that job has been written.  The workers take
// the jobs in the order they were queued, sleep between bursts and spin (pause) inside one.
class HostPool {
public:
    static constexpr int MAX_CHUNKS = 64;
    explicit HostPool(int threads);
    ~HostPool();
    int threads() const;
    unsigned long long begin(const HostLayout &lay, const HostJob &job, const long long *bounds, int nchunks,
                             const volatile unsigned *flags = nullptr, unsigned epoch = 0);
    void publish(unsigned long long id, int chunk);
    void finish(unsigned long long id);
    bool done(unsigned long long id) const;     // every worker through with that job (non-blocking)
    void abort(unsigned long long id);          // the chunks still missing will never come: workers skip them
    // measurement aid: microseconds from begin(id) until worker 0 was released on the job's first chunk / its last chunk /
    // had written its last row (valid once done(id) and until the ring slot is reused)
    void timeline(unsigned long long id, double out_us[3]) const;
private:
    struct Impl;
    Impl *impl;
};

}  // namespace diral
