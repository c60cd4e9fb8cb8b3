/*
 * diral_env.h -- C ABI of libdiral_env.so: the per-time-slot body of the DIRAL V2V "test
 * simulator" environment as hand-written sm_100a CUDA kernels over a batch of E independent
 * environments x N vehicles.
 *
 * The reference (gundoganalperen/DIRAL, pure Python) has no FFI layer; its boundary is the
 * duck-typed env object (envs/test_env.py:6-595) that main_test.py:46-236 drives and the learners
 * query (algorithms/drl_drqn.py:30,38,39).  Each entry point below cites the reference method it
 * replaces; the Python mirror of that object (diral_b200/env.py) is a thin ctypes caller of this
 * file and nothing else.  No torch types cross this boundary: plain pointers, sizes and a
 * CUstream/cudaStream_t passed as void*.
 *
 * Conventions
 *   - every entry point returns 0 on success or a negative DIRAL_ERR_* code; the message for the
 *     calling thread's last failure is diral_last_error().
 *   - "device pointer" arguments must be valid on the handle's CUDA device; launches go to the
 *     stream the caller passes and never synchronise, except where stated.
 *   - batched arrays carry a leading env axis E.  Neighbour tables are stored SUBJECT-major:
 *     tab_*[e][j][i] is vehicle i's belief about vehicle j (the reference keeps
 *     vehicles[i].pos_of_neighbors[j], envs/vehicle.py:20-33, i.e. the transpose).
 */
#ifndef DIRAL_ENV_H
#define DIRAL_ENV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIRAL_ABI_VERSION 3

enum {
    DIRAL_OK = 0,
    DIRAL_ERR_ARG = -1,        /* bad configuration / null pointer / shape             */
    DIRAL_ERR_CUDA = -2,       /* a CUDA runtime call failed (message has the detail)   */
    DIRAL_ERR_UNBOUND = -3,    /* diral_bind() has not been called                      */
    DIRAL_ERR_SEQ_RANGE = -4,  /* slot counter beyond what the packed table keys hold   */
    DIRAL_ERR_UNSUPPORTED = -5 /* a reference variant SURVEY.md section 8(a) lists as out of scope */
};

/* step modes: TestEnv.my_step / my_step_design / my_step_ch (envs/test_env.py:124,269,351) */
enum { DIRAL_MY_STEP = 0, DIRAL_MY_STEP_DESIGN = 1, DIRAL_MY_STEP_CH = 2 };

/* POD mirror of the reference's EnvironmentTest + State kwargs (envs/test_env.py:12-48) */
typedef struct diral_cfg {
    int64_t E;                 /* number of independent environments on this device            */
    int64_t env0;              /* global index of env 0 (keys the counter-based RNG, so results
                                  do not depend on how the batch is sharded across GPUs)       */
    int32_t N, R, B;           /* num_users, num_channels, State.num_bins    (test_env.py:12,13,40) */
    double  L, C, W;           /* highway_length, communication_range, bin_range (:18,21,24)   */
    int32_t reward_design;     /* :20 */
    int32_t state_type;        /* State.type :27 */
    int32_t toy;               /* congestion_test -> Network.toy_example (network.py:36)       */
    int32_t mobility;          /* :14 */
    int32_t mobility_vary;     /* :15 */
    int32_t design_topology;   /* enable_design_topology :16 */
    int32_t add_action, action_binary, add_channel_obs, add_reward, add_index, add_velocity,
            add_position, add_positional_dist, add_piggy, pos_dist_type, fingerprint; /* :28-41,19 */
    int32_t age_threshold;     /* hard-coded 20 in network.py:547 */
    double  sentinel;          /* hard-coded 100000 in network.py:385 */
} diral_cfg;

/* Device memory the caller (the Python layer, through torch.empty) owns and binds once.
 * Element counts use N, R and S = diral_state_space(cfg). */
typedef struct diral_buffers {
    double  *pos_x;            /* [E][N]   Vehicle.pos_x   (vehicle.py:10)                      */
    double  *pos_y;            /* [E][N]   Vehicle.pos_y   (vehicle.py:11)                      */
    double  *vel;              /* [E][N]   Vehicle.velocity(vehicle.py:14)                      */
    int32_t *tab_seq;          /* [E][N][N] seq_number,   subject-major (vehicle.py:32)         */
    int32_t *tab_lu;           /* [E][N][N] last_updated, subject-major (vehicle.py:33)         */
    double  *tab_x;            /* [E][N][N] xpos,         subject-major (vehicle.py:30)         */
                               /* ypos is not stored: pos_y never changes between resets, so
                                  ypos[i][j] == (seq[i][j] > 0 ? pos_y[j] : 0)                 */
    int32_t *lat;              /* [E][N(tx)][N(rx)] Network.last_arrival_time (network.py:39-42) */
    float   *obs;              /* [E][N][R] channel observations returned by my_step*           */
    float   *rews;             /* [E][N]    rewards returned by my_step*                        */
    float   *state;            /* [E][N][S] obtain_state output                                 */
    double  *acc_reward;       /* [E]       running sum of rewards since the last episode_metrics */
    int64_t *acc_count;        /* [E][4]    running {packets received, (tx,rx) pairs in range,
                                             out-of-range actions, slots}                       */
    uint32_t *scratch;         /* [diral_scratch_bytes/4] work space (may be NULL if that is 0) */
    const double *trace;       /* [trace_len][N] Network.x_positions (network.py:171-178) or NULL */
    int64_t trace_len;
    double  *ring;             /* ROW layout only (diral_get_option "layout" == 1, see below): [E][H][T] position of
                                  every vehicle at tick mod H; NULL otherwise                                        */
} diral_buffers;

/* Table layouts.  diral_get_option(handle, "layout") says which one the handle's kernel uses; it is fixed by the
 * configuration and the "variant" option, so query it after diral_create / diral_set_option and before allocating.
 *   0  SUBJECT-major, dense xpos (lane-group kernel N <= 32, round-1 one-CTA-per-env kernel): tab_seq / tab_lu / tab_x
 *      are [E][N][N] with tab_*[e][j][i] = vehicle i's belief about vehicle j.
 *   1  ROW layout (diral_step_row.cu, 32 < N <= 256): tab_seq / tab_lu are OBSERVER-major [E][N][T], T =
 *      diral_get_option("row_stride") (N rounded up to a multiple of 64; columns >= N are padding).  xpos is not stored
 *      per entry: an entry's position is pos_x[subject] at the tick that produced its sequence number, kept in
 *      ring [E][H][T] (H = "ring_depth"); entries older than H ticks live in tab_x, here two spill halves
 *      [2][E][N][T] used alternately.  diral_materialize_x rebuilds the reference's xpos table from either layout. */

/* Bytes of device memory behind diral_buffers for this configuration (tables + outputs). */
size_t diral_state_bytes(const diral_cfg *cfg);
/* Bytes the `scratch` member must hold (0 when the table keys fit in shared memory). */
size_t diral_scratch_bytes(const diral_cfg *cfg);
/* TestEnv.get_state_space (test_env.py:49-85,492). */
int32_t diral_state_space(const diral_cfg *cfg);
int32_t diral_abi_version(void);

/* TestEnv.__init__ (test_env.py:7-107): validates the configuration, picks the kernel variant. */
int diral_create(const diral_cfg *cfg, void **handle);
int diral_destroy(void *handle);
int diral_bind(void *handle, const diral_buffers *bufs);

/* Network.__init__ + initialize_mobility_topology* (network.py:15-119): zero tables, lat = -1,
 * topology from x0/y0/v0 (device, [E][N] float64) or -- when x0 is NULL -- from the counter-based
 * generator (Philox4x32-10 keyed by seed and the GLOBAL env index). */
int diral_reset(void *handle, const double *x0, const double *y0, const double *v0, uint64_t seed,
                void *stream);

/* Network.reset_positions (network.py:181-187), reached through TestEnv.reset_mobility_env (test_env.py:479-484):
 * fresh vehicles -- zero tables and a new topology as in diral_reset -- while last_arrival_time, the episode
 * accumulators and everything the caller counts (slot, episode) keep their values. */
int diral_reset_topology(void *handle, const double *x0, const double *y0, const double *v0, uint64_t seed,
                         void *stream);

/* TestEnv.sample (test_env.py:116-122): uniform actions on [0,R), out [E][N] int32 (device). */
int diral_sample(void *handle, uint64_t seed, int64_t t, int32_t *out, void *stream);

/* TestEnv.my_step / my_step_design / my_step_ch (test_env.py:124-266, 269-316, 351-443), batched:
 * table tick, per-resource collision histogram, reward model, nearest-transmitter search, table
 * merges in ascending resource order, last_arrival_time, mobility.  Writes bufs.obs and bufs.rews.
 * When build_state != 0 it also writes bufs.state = obtain_state(obs, actions, rews, episode,
 * epsilon) (test_env.py:527-583) from the same pass over the tables (one read, one write).
 * actions: [E][N] int32 device pointer, or NULL to draw them on device as diral_sample(seed,
 * timestep) would (then they are also written to actions_out if that is not NULL). */
int diral_step(void *handle, int mode, const int32_t *actions, int64_t timestep, int build_state,
               double episode, double epsilon, uint64_t seed, int32_t *actions_out, void *stream);

/* TestEnv.obtain_state (test_env.py:527-583) on caller-supplied obs/acts/rewards (device). */
int diral_obtain_state(void *handle, const float *obs, const int32_t *actions, const float *rews,
                       double episode, double epsilon, float *out, void *stream);

/* T consecutive slots of diral_step(mode, NULL, t0 + k, build_state=1, ...) with on-device actions; the buffers
 * hold the outputs of the last slot.  For N <= 32 (lane-group kernel, fused state) all T slots run in ONE launch. */
int diral_rollout(void *handle, int mode, int32_t T, int64_t t0, uint64_t seed, void *stream);

/* Network.update_velocity (network.py:208-222); draws [E][N] int8 in {1,2,3} or NULL -> Philox. */
int diral_update_velocity(void *handle, const int8_t *draws, uint64_t seed, int64_t episode,
                          void *stream);

/* Network.get_information_age (network.py:560-574); out [E][100] int32 (device). */
int diral_information_age(void *handle, int64_t timestep, int32_t *out, void *stream);

/* End-of-episode metric vector (the only cross-GPU exchange: one all-reduce(sum) of out110):
 * [0] sum of rewards, [1] sum of collisions = slots*R - sum rewards (main_test.py:178), [2] packets
 * received, [3] (tx,rx) pairs in range, [4] agent-steps, [5] out-of-range actions, [6..9] reserved,
 * [10..109] information-age histogram at `timestep`.  Deterministic (fixed-order) reduction over
 * this device's envs; clears the per-env accumulators.  out110: device, float64[110]. */
int diral_episode_metrics(void *handle, int64_t timestep, double *out110, void *stream);

/* Per-slot caller epilogue of the reference's driver loop (main_test.py:150-206 + utils/misc.py:1-12),
 * the first "next" row of SURVEY.md 8(f): information age of this slot, its weighted sum
 * (calculate_ia_penalty), and the reward shaping -- ia_averaging (+-1 by the direction of the weighted
 * sum), ia_penalty_enable (stuck-on-a-bad-resource penalty), global_reward_avg (+ sum_r / N) -- applied in
 * place to rewards [E][N] (float32, device).  The caller owns the shaping state: sum_ia_prev [E] int64
 * (initially 0), ia_counter [E][N] int32 (0), prev_actions [E][N] int32 (-1); arrays of disabled options
 * may be NULL.  slot_sums [E][3] float64 = {sum of raw rewards, collisions = R - sum, weighted information
 * age} and ia_out [E][100] int32 are optional outputs. */
typedef struct diral_shaping {
    int32_t ia_averaging, ia_penalty_enable, ia_penalty_threshold, global_reward_avg;   /* main_test.py:34-50 */
    double  ia_penalty_value;
} diral_shaping;
int diral_shape_rewards(void *handle, const diral_shaping *cfg, const int32_t *actions, int64_t timestep,
                        float *rewards, int64_t *sum_ia_prev, int32_t *ia_counter, int32_t *prev_actions,
                        double *slot_sums, int32_t *ia_out, void *stream);

/* Device-resident replay ring, the second "next" row of SURVEY.md 8(f): the window gather behind
 * Memory.sample (utils/memory.py:177-194) fused with the learners' regrouping loops
 * (algorithms/drl_drqn.py:294-377, which turn batch x step x user tuples into [user][batch][step][...]).
 * ring:  [capacity][agents][width] elements of elem_bytes (states: width = S float32; actions / rewards:
 *        width = 1); a batched env folds its env axis into the agent axis (agents = E * N).
 * start: device int64 [batch], first ring slot of every sampled window (np.random.choice(len - step) on the host).
 * out:   [agents * batch][step][width], row (a * batch + b) = agent a in window b -- the layout
 *        drl_drqn.py:235-238 reshapes to.  No handle: it touches no environment state. */
int diral_ring_gather(const void *ring, int64_t capacity, int64_t agents, int64_t width, int32_t elem_bytes,
                      const int64_t *start, int32_t batch, int32_t step, void *out, void *stream);

/* The last "next" row of SURVEY.md 8(f), part 1: the view-based positional distribution of the RealNeS
 * environment, RealnessEnv.get_neighbor_dist (pos_dist 1, envs/realness_env.py:52-85) and get_neighbor_dist2
 * (pos_dist 2, :87-118), on neighbour tables in the wire layout of MA_NeighborTableEntry (envs/ma_messages_pb2.py;
 * filled by RealNeSZmqBridge.get_observation_syn_dist, envs/realness_bridge.py:168-191).
 * tables:   device, [M][N] entries; observer: device int32 [M], the 0-based user_id - 1 each table is seen from
 *           (realness_env.py:371-373); out: device float32 [M][bins].
 * Entries with last_update > age_limit (20 in the reference) are skipped; pos_dist 2 bins the signed distances
 * into `bins` equal bins over [-range, range] and divides by the number of fresh entries (out-of-range ones
 * included, as the reference does); pos_dist 1 is the weighted histogram of the sorted, max-normalised distances.
 * No handle: it touches no environment state. */
typedef struct diral_wire_entry {
    float   pos_x, pos_y;          /* MA_NeighborTableEntry.pos_x / pos_y  (float32 on the wire) */
    int32_t seq_num, last_update;  /* MA_NeighborTableEntry.seq_num / last_update */
} diral_wire_entry;
int diral_wire_vpd(const diral_wire_entry *tables, const int32_t *observer, int64_t M, int32_t N, int32_t pos_dist,
                   int32_t bins, double range, int32_t age_limit, float *out, void *stream);

/* ... part 2: the 3GPP semi-persistent-scheduling baseline, SemiPersistentScheduling.step and
 * choose_new_resource (algorithms/v2x_sps.py:76-104, :24-74), one state machine per agent.
 * selection_window: device float64 [agents][window_len] averaged RSSI per subframe (the reference receives
 *   them as protobuf doubles, envs/realness_bridge.py:195-208); prev_action / reselection_counter: device
 *   int32 [agents], the per-agent state (v2x_sps.py:13-18), updated in place; actions: device int32 [agents].
 * draws: device float64 [agents][3] = {new reselection counter (random.randint(5, 16)), keep-uniform
 *   (random.random()), choice index (random.choice picks candidates[index mod len])}, consumed only by agents
 *   whose counter is 0; NULL draws them from the counter-based generator keyed by (seed, agent, t).
 * flags (optional, device int32 [agents]) is set to 1 where the reference would raise or never return
 *   (no candidate list can reach cfg.min_candidates); that agent keeps its previous subframe. */
typedef struct diral_sps_cfg {
    double rssi_threshold;         /* v2x_sps.py:11 */
    double inc_db;                 /* 3 dB per widening round, v2x_sps.py:19 */
    double prob_resource_keep;     /* 0.8, v2x_sps.py:22 */
    double min_candidates;         /* len(selection_window) / 5 as the host interpreter evaluates it, v2x_sps.py:40 */
} diral_sps_cfg;
int diral_sps_step(int64_t agents, int32_t window_len, const double *selection_window, const diral_sps_cfg *cfg,
                   const double *draws, uint64_t seed, int64_t t, int32_t *prev_action, int32_t *reselection_counter,
                   int32_t *actions, int32_t *flags, void *stream);

/* Host-buffer convenience for callers that keep data on the CPU (the reference's learners do):
 * copies h_actions in, runs diral_step(build_state=1), copies state/rews (and obs if not NULL) out,
 * synchronises the stream.  Host pointers should be pinned for full PCIe bandwidth. */
int diral_step_host(void *handle, int mode, const int32_t *h_actions, int64_t timestep,
                    double episode, double epsilon, float *h_state, float *h_rews, float *h_obs,
                    void *stream);

/* diral_step_host in two halves ("host_format" 3 only): _begin enqueues the slot (one launch on `stream`, a few
 * microseconds) and returns; _wait returns once every state row and reward of that slot is in the caller's buffers
 * and the launch has retired.  A caller that keeps two or more handles (groups of environments) in flight overlaps
 * one group's kernel and PCIe records with the row assembly of another.  The buffers passed to _begin must stay
 * untouched until _wait; no other call on the handle in between.  DIRAL_ERR_UNSUPPORTED where host_format 3 does not
 * apply (use diral_step_host). */
int diral_step_host_begin(void *handle, int mode, const int32_t *h_actions, int64_t timestep,
                          double episode, double epsilon, float *h_state, float *h_rews, void *stream);
int diral_step_host_wait(void *handle);

/* Compact host format of diral_step_host ("host_format" = 1, opt-in): the state rows of TestEnv.obtain_state
 * (test_env.py:527-583) are mostly known to the host already (the one-hot of the action it sent, index, episode,
 * epsilon) or a few bytes of information per agent (B bin counts behind the B float32 of the positional
 * distribution).  With host_format 1 only that information crosses PCIe (counts as bytes, rewards, and obs /
 * positions / velocities when the State block carries them) and the library's host threads ("host_threads", default
 * = the CPUs the process may run on, minus one) assemble the [E][N][S] float32 rows in h_state, bit for bit what
 * host_format 0 delivers.  State blocks without a fused build (add_positional_dist, VPD type 1) keep format 0.
 *
 * host_format 3 ("streamed", what diral_b200.TestEnv selects) moves the same records without a copy engine and with
 * ONE launch per slot: the lane-group kernel reads pinned caller actions in place ("actions_direct"), stages an
 * environment's records in shared memory, writes them into mapped host memory as one bulk store and -- after a
 * system-scope fence -- counts the environment into its chunk; the environment that completes a chunk raises that
 * chunk's flag in host memory, which releases the assembly threads ("stream_chunks" flags per slot, default 16).
 * Configurations the lane-group kernel does not serve, and calls that also ask for obs, run as host_format 1.
 *
 * diral_expand_state_host is that row assembly alone (no device involved): agents = E*N records in, rows out. */
int diral_expand_state_host(const diral_cfg *cfg, int64_t agents, const int32_t *actions, const uint8_t *counts,
                            const float *rews, const float *obs, const double *pos_x, const double *pos_y,
                            const double *vel, double episode, double epsilon, int32_t threads, float *out);

/* Timeline of the last compact-format diral_step_host call, microseconds since its entry: [0] host threads woken,
 * [1] all chunks enqueued, [2 + k] chunk k's record landed in host memory, [2 + chunks] every row assembled.
 * After diral_step_host_begin + _wait: [0] host threads woken, [1] launch enqueued, then -- as seen by assembly thread 0 --
 * [2] its first chunk released, [3] its last chunk released, [4] its last row written.
 * Returns the number of values written (<= n).  Measurement aid: says what bounds the end-to-end slot. */
int32_t diral_host_trace(void *handle, double *out_us, int32_t n);

/* vehicles[i].pos_of_neighbors[j]["xpos"] (vehicle.py:30) for every entry, out [E][N][N] float64 (device), indexed
 * [e][i][j] like the reference, from whichever layout the handle uses. */
int diral_materialize_x(void *handle, double *out, void *stream);

/* One slot of a device-resident replay ring (Memory.add, utils/memory.py:169-175): ring[slot] = src, row_bytes
 * bytes device to device on `stream`. */
int diral_ring_put(void *ring, int64_t capacity, int64_t slot, int64_t row_bytes, const void *src, void *stream);

/* Options.  Test/bench knobs: "variant" = 0 auto | 1 lane-group kernel (N <= 32) | 2 one-CTA-per-env kernel (the
 * row-layout kernel from 97 vehicles on, else round 1's) | 3 round 1's one-CTA-per-env kernel | 4 the row-layout
 * kernel (33..256 vehicles, fused State block); set before diral_bind;
 * "track_lat" = 1 keeps last_arrival_time bookkeeping on even before the first my_step_ch call.
 * diral_step_host: "host_format" (0 full rows | 1 compact, records through the copy engine | 2 compact, records written
 * by the kernels into mapped host memory | 3 streamed: one launch + per-chunk flags), "host_threads", "host_chunks" (env
 * chunks pipelined per call, formats 1 and 2), "stream_chunks" (flags per slot, format 3), "stream_split" (format 3: 1 = launch
 * the split-environment instantiation, two warps per environment, so that records start arriving earlier; default 1), "actions_direct" (pinned
 * actions read in place: 0 never | 1 format 3 only | 2 always), "host_nt" (row stores: -1 auto | 0 ordinary | 1
 * non-temporal), "host_pool_shared" (1: this handle's rows are assembled by the process-wide pool, sized by the first handle
 * that uses it -- what handles pipelined with diral_step_host_begin / _wait should share).  Checkpoint restore: "ticks" (table ticks since the reset = every vehicle's own sequence number) and
 * "lat_live" (a my_step_ch has stamped last_arrival_time); diral_get_option reads any of them back (-1: unknown),
 * plus "compact_ok" (1 when this State block has a compact host format), "kernel" (1 lane-group, 2 round-1
 * one-CTA-per-env, 3 row layout), "layout", "row_stride", "ring_depth" and "scratch_bytes" (what `scratch` must hold
 * for the handle's kernel). */
int diral_set_option(void *handle, const char *name, int64_t value);
int64_t diral_get_option(void *handle, const char *name);

/* Number of kernels launched through this handle since creation (bench bookkeeping). */
int64_t diral_launch_count(void *handle);

const char *diral_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
