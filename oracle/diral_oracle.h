/*
 * diral_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, float64 like the reference) of the per-slot body of the
 * DIRAL "test simulator" environment.  It exists to check the CUDA path; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors of its own
 * (SURVEY.md section 4), so the oracle is pinned against outputs of the reference itself,
 * executed unmodified in the build container by tests/golden/make_golden.py and committed
 * as the .npz fixtures under tests/golden/ (tests/test_oracle_golden.py replays each one).
 *
 * Layout: all arrays carry a leading env axis E.  Tables are stored the way the reference
 * stores them (row i = vehicle i's belief about everyone, reference envs/vehicle.py:20-33),
 * i.e. [E][N(observer)][N(subject)] -- deliberately NOT the subject-major layout the CUDA
 * path uses, so the two implementations share no indexing code.
 */
#ifndef DIRAL_ORACLE_H
#define DIRAL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t N, R, B;          /* num_users, num_channels, num_bins  (test_env.py:12,13,40) */
    double  L, C, W;          /* highway_length, communication_range, bin_range (:18,21,24) */
    int32_t reward_design;    /* test_env.py:20 */
    int32_t state_type;       /* State.type, test_env.py:27 */
    int32_t toy;              /* congestion_test -> Network.toy_example, network.py:36 */
    int32_t mobility;         /* test_env.py:14 */
    int32_t mobility_vary;    /* test_env.py:15 */
    int32_t design_topology;  /* enable_design_topology, test_env.py:16 */
    int32_t add_action, action_binary, add_channel_obs, add_reward, add_index,
            add_velocity, add_position, add_positional_dist, add_piggy,
            pos_dist_type, fingerprint;                 /* test_env.py:28-41,19 */
    int32_t age_threshold;    /* hard-coded 20, network.py:547 */
    double  sentinel;         /* hard-coded 100000, network.py:385 */
} orc_cfg;

typedef struct {
    int64_t E;
    double  *pos_x, *pos_y, *vel;                 /* [E][N] */
    double  *tab_x, *tab_y;                       /* [E][N][N] observer-major */
    int32_t *tab_seq, *tab_lu;                    /* [E][N][N] */
    int32_t *lat;                                 /* [E][N(tx)][N(rx)] last_arrival_time */
    const double *trace;                          /* [trace_len][N] shared by all envs, or NULL */
    int64_t trace_len;
} orc_batch;

enum { ORC_MY_STEP = 0, ORC_MY_STEP_DESIGN = 1, ORC_MY_STEP_CH = 2 };

int  orc_state_space(const orc_cfg *c);
void orc_set_threads(int n);
int  orc_get_threads(void);

/* TestEnv.__init__/Network.__init__ with caller-supplied topology (zero tables, lat=-1) */
void orc_reset(const orc_cfg *c, orc_batch *b, const double *x0, const double *y0, const double *v0);

/* my_step / my_step_design / my_step_ch ; obs [E][N][R], rews [E][N]; counts (nullable)
 * [E][2] = {packets received, (tx,rx) pairs in range} this slot */
void orc_step(const orc_cfg *c, orc_batch *b, int mode, const int32_t *actions, int64_t timestep,
              double *obs, double *rews, int64_t *counts);

/* TestEnv.obtain_state; out [E][N][S] */
void orc_obtain_state(const orc_cfg *c, const orc_batch *b, const double *obs, const int32_t *acts,
                      const double *rews, double episode, double epsilon, double *out);

/* Network.get_information_age; out [E][100] */
void orc_information_age(const orc_cfg *c, const orc_batch *b, int64_t timestep, int32_t *out);

/* Network.update_velocity with explicit draws in {1,2,3}; draws [E][N] */
void orc_update_velocity(const orc_cfg *c, orc_batch *b, const int8_t *draws);

/* Counter-based RNG shared (by specification, not by code) with the CUDA path:
 * Philox4x32-10, key = (seed_lo, seed_hi ^ stream), counter = (agent, env, t_lo, t_hi). */
void orc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                uint32_t out[4]);
void orc_philox_actions(uint64_t seed, int64_t env0, int64_t E, int32_t N, int32_t R, int64_t t,
                        int32_t *out);
void orc_philox_topology(const orc_cfg *c, uint64_t seed, int64_t env0, int64_t E,
                         double *x0, double *y0, double *v0);
void orc_philox_draws(uint64_t seed, int64_t env0, int64_t E, int32_t N, int64_t episode,
                      int8_t *out);

#ifdef __cplusplus
}
#endif
#endif
