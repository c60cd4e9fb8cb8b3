// diral_step_group.cu -- fused time-slot kernel for N <= 32 vehicles per environment.
//
// One group of G lanes (G = 4, 8, 16 or 32, the power of two >= N) owns one environment; lane u is
// vehicle u both as a receiver and as the owner of row u of the neighbour table.  The slot splits
// into a DECISION phase that never touches the table and a TABLE phase that streams it once:
//
//   A  load actions / kinematics; prefetch this env's whole table (and the next env's inputs) to L2
//   C1 decisions.  __match_any_sync on the actions gives every lane its collision set (the
//      per-resource collision histogram, reference envs/test_env.py:149-157) in one instruction, so all
//      three reward models are lane-local.  An in-range bitmask over all G candidates is built once; then
//      for r = 0..R-1 in order (test_env.py:147), two resources in flight: nearest in-range transmitter
//      (Network.find_closest_tx, network.py:378-398) = lowest candidate bit, second lowest compared
//      branch-free, further ones in a rare loop; channel observation; last_arrival_time; every
//      non-empty pass appends one byte per lane ("which row does lane u merge") to a shared-memory script
//   D  mobility (Network.update_positions, network.py:189-206)
//   C2/E per slab of 8 table columns (subject-major storage => every access is one coalesced G*4 or
//      G*8 byte segment; the next slab's seq column and this slab's last_updated / xpos are requested
//      before the replay loop, so their latency hides behind it):
//        tick (Vehicle.periodic_update, vehicle.py:56-70) and pack  key = seq << log2 G | origin-row
//        (two columns per register -- 16-bit "freshness" keys -- unless the slab holds entries older
//        than 2^(16 - log2 G) slots), replay the script IN PASS ORDER -- the passes are a true
//        sequential dependency, but table COLUMNS are independent, so a slab can run all
//        passes by itself:   key[q] = max(key[q], shfl(key[q], script[pass][u]))
//        (Vehicle.received_update, vehicle.py:35-47: one shuffle + one integer max per entry or entry
//        pair).  (xpos, ypos) of an entry is a pure function of (subject, seq), so the max-by-seq join
//        never moves positions: the key's low bits remember which row held that version when the slot
//        began.  Then gather xpos from the origin row by shuffle, last_updated bookkeeping, write
//        seq / last_updated / xpos back, and bin the view-based positional distribution
//        (Network.get_positional_dist_2_piggy + dist_piggy, network.py:473-513,538-558) with
//        fire-and-forget shared-memory reductions
//   F  TestEnv.obtain_state (test_env.py:527-583): assemble [E][N][S] float32 rows in shared memory
//      and write obs / rewards / state with coalesced float4 stores.
//
// Every hot loop is written as ONE basic block (compile-time FLAT variants for the "all vehicles on one
// lane" case, unconditional reductions into a dummy histogram row, predicated selects instead of
// branches): the kernel is bound by instruction latency, and ptxas only interleaves independent
// columns / resources inside a basic block.
//
// HBM traffic per env-slot is the algorithmic minimum SURVEY.md 8(d) states: the table is read once
// and written once (16 B per entry each way), everything else is O(N).  Only one slab of keys lives
// in registers at a time (16 warps resident per SM at ~118 registers).
//
// FULL (N == G) instantiations drop every "is this lane / column live" predicate and turn all
// table addresses into compile-time offsets from one base pointer.
#include "diral_dev.cuh"
#include "diral_launch.h"

#include <algorithm>
#include <mutex>
#include <type_traits>

namespace diral {

namespace {

template <int G> struct Log2;
template <> struct Log2<4>  { static constexpr int v = 2; };
template <> struct Log2<8>  { static constexpr int v = 3; };
template <> struct Log2<16> { static constexpr int v = 4; };
template <> struct Log2<32> { static constexpr int v = 5; };

__host__ __device__ inline int align16i(int x) { return (x + 15) & ~15; }

// shared-memory carve-up of one group (bytes); the host computes the same numbers
// State rows leave as float4 stores straight from the registers that hold them (no shared-memory staging,
// which is what lets 28 one-warp CTAs share an SM at C3) when every feature block starts 16 B aligned.
__host__ __device__ inline bool group_direct_rows(const Params &p)
{
    const int n_act = p.add_action ? (p.action_binary ? p.R : 1) : 0;
    return (p.S & 3) == 0 && (n_act & 3) == 0 && (!p.add_channel_obs || (p.R & 3) == 0) && (!p.piggy || (p.B & 3) == 0);
}

struct GroupSmem {
    int off_sx, off_sy, off_script, off_txm, off_recv, off_mcnt, off_obs, off_hist, off_st, off_rec, bytes;
    __host__ __device__ GroupSmem(int G, int R, int B, int S, bool state, bool vpd, bool direct, int rec_stride)
    {
        int o = 0;
        off_script = o; o += align16i(G * (R + 1));    // merge script: one byte per (pass, lane), + 1 spare row
        off_txm = o;    o += align16i(4 * R);          // transmitter mask of every resource
        off_recv = o;   o += align16i(4 * G);          // packets received per transmitter (my_step_ch)
        off_mcnt = o;   o += align16i(4 * G);          // VPD sample counts summed over the warps of a split environment
        off_sx = o;   o += align16i(8 * G);
        off_sy = o;   o += align16i(8 * G);
        off_obs = o;  o += align16i(4 * G * R);
        off_hist = o; o += (state && vpd) ? align16i(4 * G * (B + 1)) : 0;   // + 1 dummy row
        off_st = o;   o += (state && !direct) ? align16i(4 * G * S) : 0;
        off_rec = o;  o += align16i(G * rec_stride);   // compact host records of this environment (one coalesced copy-out)
        bytes = o;
    }
};

// general distance, kept out of line so that the common dy == 0 highway never executes (or even
// schedules around) the fp64 square-root sequence
__device__ __noinline__ double dist_slow(double dx, double dy)
{
    return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// Network.dist (network.py:318-332).  FLAT promises dy == 0 (every vehicle on one lane of the
// highway), where the result is |dx|; it is a compile-time flag in the hot loops so that they contain
// no branch and no call, and the scheduler can interleave independent columns / resources.
template <bool FLAT>
__device__ __forceinline__ double dist_t(double x1, double y1, double x2, double y2)
{
    const double dx = __dsub_rn(x2, x1);
    if (FLAT) return fabs(dx);
    const double dy = __dsub_rn(y2, y1);
    if (dy == 0.0) return fabs(dx);
    return dist_slow(dx, dy);
}

__device__ __forceinline__ double dist_uni(bool flat, double x1, double y1, double x2, double y2)
{
    return flat ? dist_t<true>(x1, y1, x2, y2) : dist_t<false>(x1, y1, x2, y2);
}

// Network.calculate_reward_weights / calculate_avg_distance (network.py:273-316):
// mean of dist over itertools.combinations(transmitters, 2), Python sum() semantics
__device__ __noinline__ int reward_weight(const Params &p, bool flat, const double *sx, const double *sy,
                                          unsigned txm, double norm)
{
    PySum s; int pairs = 0;
    for (unsigned mi = txm; mi; mi &= mi - 1) {
        const int i = __ffs(mi) - 1;
        for (unsigned mj = mi & (mi - 1); mj; mj &= mj - 1) {
            const int j = __ffs(mj) - 1;
            s.add(dist_uni(flat, sx[i], sy[i], sx[j], sy[j]));
            ++pairs;
        }
    }
    const double m = __ddiv_rn(s.result(), (double)pairs);
    return p.toy ? (m == norm) : (m > p.C);
}

// hist[..] += 1 in shared memory without waiting for (or depending on) the old value
__device__ __forceinline__ void smem_red_inc(unsigned *addr)
{
    // no "memory" clobber on purpose: the surrounding column code must stay free to interleave; every
    // plain access to the histogram is separated from these reductions by a __syncwarp
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(addr)));
}

// exact bin of a sample against the linspace edges -- only reached within 1e-6 of a bin boundary
__device__ __noinline__ int vpd_bin_edges(double s, double W, double inv_binw, int B, const double *edges)
{
    return vpd_bin(s, W, inv_binw, B, edges);
}

#ifndef DIRAL_MIN_BLOCKS
#define DIRAL_MIN_BLOCKS 1       // tuning knob: minimum resident WARPS per SM for the G >= 16 instantiations
#endif
#ifndef DIRAL_GROUP_WARPS
#define DIRAL_GROUP_WARPS 1      // tuning knob: warps (= environments at G == 32) per CTA for G >= 16
#endif

// SP ("split"): the WARPS warps of the CTA share ONE environment.  Every warp runs the decision phase for itself (no
// cross-warp hazards: the merge script, observation rows and masks are per-warp copies), then takes every WARPS-th
// slab of table columns; the positional-distribution histogram is shared (it is filled by reductions anyway), and
// warp 0 alone writes rewards, state rows and the episode accumulators.  Latency of one environment drops to the
// decision phase plus 1 / WARPS of the table phase -- what the launch uses for the tail of a batch that does not
// fill the device a whole number of times (launch_k).
template <int G, bool FULL, int WARPS, int MODE, bool LAT, bool ROLL, bool CNT, bool SP = false>
__global__ void __launch_bounds__(WARPS * 32, ((ROLL || G < 16) ? 1 : (DIRAL_MIN_BLOCKS + WARPS - 1) / WARPS))
step_group_kernel(const Params p)
{
    static_assert(!SP || (G == 32 && !ROLL), "split environments: 32-lane groups, single slot");
    constexpr int EPW = 32 / G;              // environments per warp
    constexpr int SB = Log2<G>::v;           // low key bits holding the origin row
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int u = lane & (G - 1), sub = lane / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (sub * G));
    const long long e = SP ? (long long)blockIdx.x : ((long long)blockIdx.x * WARPS + warp) * EPW + sub;
    const bool lead = !SP || warp == 0;      // the warp that owns an environment's outputs
    const int N = FULL ? G : p.N;
    const int R = p.R, B = p.B, S = p.S;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_edges = reinterpret_cast<double *>(smem_raw);
    for (int i = threadIdx.x; i <= B; i += blockDim.x) s_edges[i] = p.edges[i];
    __syncthreads();
    if (e >= p.E) return;                    // whole groups leave; only group-masked syncs below

    const bool want_state = p.build_state != 0;
    const bool vpd = want_state && p.vpd_enabled;
    const bool direct = group_direct_rows(p);
    const GroupSmem lay(G, R, B, S, want_state, p.vpd_enabled, direct, (CNT && p.vpd_counts) ? p.rec_stride : 0);
    unsigned char *gbase = smem_raw + align16i(8 * (B + 1)) + (size_t)(warp * EPW + sub) * lay.bytes;
    double *sx = reinterpret_cast<double *>(gbase + lay.off_sx);
    double *sy = reinterpret_cast<double *>(gbase + lay.off_sy);
    float *obsS = reinterpret_cast<float *>(gbase + lay.off_obs);      // [N][R], same layout as global
    unsigned char *script = gbase + lay.off_script;                    // [passes][G]
    unsigned *txm_s = reinterpret_cast<unsigned *>(gbase + lay.off_txm);
    unsigned *recv_s = reinterpret_cast<unsigned *>(gbase + lay.off_recv);
    unsigned char *gbase0 = SP ? smem_raw + align16i(8 * (B + 1)) : gbase;   // split: histogram / counts live in warp 0's area
    unsigned *hist = reinterpret_cast<unsigned *>(gbase0 + lay.off_hist);
    unsigned *mcnt_s = reinterpret_cast<unsigned *>(gbase0 + lay.off_mcnt);
    float *st = reinterpret_cast<float *>(gbase + lay.off_st);         // [N][S], rows rotated (see F)
    unsigned char *recS = gbase + lay.off_rec;                         // [N][rec_stride] (CNT)

    const bool act = FULL ? true : (u < N);
    const long long vbase = e * N;           // first vehicle of this env in the [E][N] arrays
    const long long tbase = e * (long long)N * N;

    // (tuning knob, measured at C3: 4 columns = 80 registers, 24 warps / SM: 45.8 us; 8 = 114 registers, 16 warps: 44.3 us;
    //  16 = 194 registers, 10 warps: 57.5 us -- warps traded for per-warp instruction-level parallelism, same throughput)
#ifndef DIRAL_SLAB
#define DIRAL_SLAB 8
#endif
    constexpr int SL = G < DIRAL_SLAB ? G : DIRAL_SLAB;        // columns per slab
    // Fused rollout: one launch runs n_slots consecutive slots of this environment.  Every global word a slot
    // reads was written by the same lane one slot earlier (table columns, own position) or never changes, so the
    // slots of an environment need no more than the group-wide __syncwarp at the end of the loop body.
    // (a separate instantiation: the single-slot kernel keeps its register allocation)
    const int n_slots = ROLL ? p.n_slots : 1;
#pragma unroll 1
    for (int slot = 0; slot < n_slots; ++slot) {
    const long long timestep = p.timestep + slot;
    const int tick = p.tick + slot;
    // ---- A: start every independent global access before the first dependent use ------------------
    int32_t *seqp = p.tab_seq + tbase + u, *lup = p.tab_lu + tbase + u;   // column j of this lane: [j * N]
    double *xp = p.tab_x + tbase + u;
    int a = -1; double x = 0.0, y = 0.0, v = 0.0; int bad = 0;
    if (act) {
        if (!p.gen_actions) a = p.actions[vbase + u];
        x = p.pos_x[vbase + u]; y = p.pos_y[vbase + u]; v = p.vel[vbase + u];
    }
    int s_next[SL];                          // seq column of the upcoming slab (loaded one slab ahead)
    constexpr int SLAB_STEP = SP ? WARPS * SL : SL;      // a split environment's warps interleave slabs
    const int slab0 = SP ? warp * SL : 0;
#pragma unroll
    for (int q = 0; q < SL; ++q) s_next[q] = (p.piggy && (FULL || (slab0 + q < N && act))) ? seqp[(slab0 + q) * N] : 0;
    {
        // the CTA that will take over this slot most likely handles env e + (resident CTAs); start its
        // per-vehicle inputs towards L2 now so that its first dependent instruction does not wait on HBM
        const long long en = e + (long long)p.prefetch_ahead;
        if (u == 0 && lead && en < p.E) {
            if (!p.gen_actions) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.actions + en * N));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.pos_x + en * N));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.pos_y + en * N));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.vel + en * N));
            if (N * 8 > 128) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.pos_x + en * N + 16));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.pos_y + en * N + 16));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.vel + en * N + 16));
            }
        }
    }
    if (p.piggy) {
        // pull this environment's whole table (16 * N * N bytes, contiguous per array) from HBM into L2
        // now; the decision phase below runs while it arrives and the slab loads then hit L2
        const char *b0 = reinterpret_cast<const char *>(p.tab_seq + tbase);
        const char *b1 = reinterpret_cast<const char *>(p.tab_lu + tbase);
        const char *b2 = reinterpret_cast<const char *>(p.tab_x + tbase);
        const int bytes4 = N * N * 4;
        for (int o = (SP ? warp * G + u : u) * 128; o < bytes4; o += (SP ? WARPS : 1) * G * 128) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(b0 + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(b1 + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + bytes4 + o));
        }
    }
    if (act) {
        if (p.gen_actions) a = philox_action(p.seed, u, p.env0 + e, timestep, R);
        if (a < 0 || a >= R) { bad = 1; a = min(max(a, 0), R - 1); }
        if (p.gen_actions && p.actions_out && lead) p.actions_out[vbase + u] = a;
    }
    sx[u] = x; sy[u] = y;

    // every vehicle on the same lane of the highway (dy == 0 for every pair)?  warp-uniform per group
    const double y0 = __shfl_sync(gmask, y, 0, G);
    const bool flat = (__ballot_sync(gmask, act && y != y0) & gmask) == 0u;
    const bool flat0 = flat && y0 == 0.0;    // ... and phantom (never heard, ypos = 0) entries too
    __syncwarp(gmask);

    // toy reward: distance between the first-min-x and first-max-x vehicle (network.py:225-246)
    double norm = 0.0;
    if (MODE == MODE_STEP && p.toy) {
        double xmin = act ? x : INFINITY, xmax = act ? x : -INFINITY; int imin = u, imax = u;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            const double ox = __shfl_xor_sync(gmask, xmin, o, G); const int oi = __shfl_xor_sync(gmask, imin, o, G);
            if (ox < xmin || (ox == xmin && oi < imin)) { xmin = ox; imin = oi; }
            const double px = __shfl_xor_sync(gmask, xmax, o, G); const int pi = __shfl_xor_sync(gmask, imax, o, G);
            if (px > xmax || (px == xmax && pi < imax)) { xmax = px; imax = pi; }
        }
        norm = dist_uni(flat, sx[imin], sy[imin], sx[imax], sy[imax]);
    }

    // ---- C1: decisions (no table access) ---------------------------------------------------------------
    double rew = 0.0;
    int n_recv = 0, n_pairs = 0;
    int32_t *latp = LAT ? p.lat + tbase + u : nullptr;               // lat[t][u] = latp[t * N]
    const bool merge_mode = p.piggy && (MODE != MODE_STEP || p.state_type == 1 || p.state_type == 2);
    const double Cr = p.C, sentinel = p.sentinel;
    float *obs_row = obsS + u * R;
    int npass = 0;                           // passes with at least one transmitter (group-uniform)

    // who is within communication range of this vehicle (network.py:595-607), all G candidates at once
    unsigned inr_mask = 0u;
    if (flat) {
#pragma unroll
        for (int t = 0; t < G; ++t) if (dist_t<true>(sx[t], sy[t], x, y) < Cr) inr_mask |= 1u << t;
    } else {
#pragma unroll
        for (int t = 0; t < G; ++t) if (dist_t<false>(sx[t], sy[t], x, y) < Cr) inr_mask |= 1u << t;
    }
    const unsigned live_mask = FULL ? 0xffffffffu >> (32 - G) : ((1u << N) - 1u);

    // per-resource collision histogram (test_env.py:149-157): `own` = the lanes sharing this lane's
    // resource, in one instruction; its popcount is tot_actions of that resource
    const unsigned own_all = (__match_any_sync(gmask, a) & gmask) >> (sub * G);    // every lane takes part
    const unsigned own = act ? own_all : 0u;
    const int my_tot = __popc(own);
    for (int r = u; r < R; r += G) txm_s[r] = 0u;
    if (MODE == MODE_CH) recv_s[u] = 0u;
    __syncwarp(gmask);
    if (act) txm_s[a] = own;                 // every transmitter of a resource writes the same word
    __syncwarp(gmask);

    // One resource: nearest in-range transmitter of this lane (Network.find_closest_tx,
    // network.py:378-398) and its channel observation.  The candidates are the in-range bits among the
    // transmitters, visited in ascending id with strict '<' (first wins ties): the two lowest are
    // compared branch-free, a third and later ones are rare.  Pure (no stores), so two resources can be
    // in flight at once.
    const bool ov_best = MODE == MODE_STEP && p.state_type == 2;                       // observation = distance to the nearest
    const float ov_const = (MODE != MODE_STEP || p.state_type == 1) ? 1.0f : 0.0f;     // ... or a constant
    auto decide = [&](auto flat_c, int r, unsigned txm, int &tstar, float &o, unsigned &more) {
        constexpr bool FL = decltype(flat_c)::value;
        const bool is_tx = (a == r);
        const unsigned cand = (act && !is_tx) ? (inr_mask & txm) : 0u;
        const unsigned rest = cand & (cand - 1u);
        const int t1 = (__ffs(cand) - 1) & (G - 1);      // cand == 0: any valid lane, the result is discarded
        const int f2 = __ffs(rest) - 1;
        const int t2 = f2 < 0 ? t1 : f2;                 // no second candidate: compare t1 with itself
        double best = dist_t<FL>(sx[t1], sy[t1], x, y);
        const double d2 = dist_t<FL>(sx[t2], sy[t2], x, y);
        tstar = t1;
        if (d2 < best) { best = d2; tstar = t2; }
        more = rest & (rest - 1u);                       // a third, fourth .. candidate: finished by the caller
        if (more) {                                      // (rare; kept out of line of the two-resource schedule)
            for (unsigned m = more; m; m &= m - 1u) {
                const int t = __ffs(m) - 1;
                const double d = dist_t<FL>(sx[t], sy[t], x, y);
                if (d < best) { best = d; tstar = t; }
            }
        }
        if (!cand || !(best < sentinel)) { tstar = -1; best = sentinel; }              // network.py:385
        n_recv += tstar >= 0 ? 1 : 0;
        // channel observation (test_env.py:203-240 / :305-306 / :431)
        const float ov = ov_best ? (float)best : ov_const;
        o = (!is_tx && txm) ? ov : 0.0f;
    };
    // side effects of one resource, in resource order
    auto commit = [&](int r, unsigned txm, int tstar, float o) {
        obs_row[r] = o;
        if (LAT || MODE == MODE_CH) {
            if (txm != 0u) {
                if (LAT) {                                                            // network.py:394
                    const bool is_rx = act && a != r;
                    for (unsigned m = txm; m; m &= m - 1) {
                        const int t = __ffs(m) - 1;
                        if (is_rx && !((inr_mask >> t) & 1u)) latp[t * N] = -1;
                    }
                    if (MODE == MODE_CH && tstar >= 0) latp[tstar * N] = (int32_t)timestep;      // test_env.py:436
                }
                if (MODE == MODE_CH && tstar >= 0) smem_red_inc(&recv_s[tstar]);      // test_env.py:396-397
            }
        }
        // table merge (vehicle.py:35-47) is deferred: log which row this lane merges in this pass.  The row
        // of an empty pass is written too but not counted, so the next pass overwrites it.
        if (merge_mode) {
            script[npass * G + u] = (unsigned char)(tstar >= 0 ? tstar : u);
            npass += txm != 0u ? 1 : 0;
        }
    };
    auto run_decisions = [&](auto flat_c) {
        int r = 0;
        for (; r + 2 <= R; r += 2) {
            const unsigned m0 = txm_s[r], m1 = txm_s[r + 1];
            int ts0, ts1; float o0, o1; unsigned mo0, mo1;
            decide(flat_c, r, m0, ts0, o0, mo0);
            decide(flat_c, r + 1, m1, ts1, o1, mo1);
            commit(r, m0, ts0, o0);
            commit(r + 1, m1, ts1, o1);
        }
        if (r < R) {
            const unsigned m0 = txm_s[r];
            int ts0; float o0; unsigned mo0;
            decide(flat_c, r, m0, ts0, o0, mo0);
            commit(r, m0, ts0, o0);
        }
    };
    if (flat) run_decisions(std::true_type{}); else run_decisions(std::false_type{});
    // candidate (receiver, transmitter) pairs of the slot: the transmitter masks of the resources partition the live
    // vehicles, so summed over the resources a lane does not transmit on they are its in-range vehicles outside its own
    // collision set
    if (act) n_pairs += __popc(inr_mask & live_mask & ~own);

    // rewards (test_env.py:159-199 / :294-302 / :408-429), all lane-local
    if (MODE == MODE_STEP) {
        if (my_tot <= 1) rew = 1.0;
        else {
            int w = 0;
            if (design_needs_weight(p.reward_design, my_tot)) {
                if (my_tot == 2) {   // one pair: the mean is that pair's distance (sum([d]) / 1 == d)
                    const int o = __ffs(own & ~(1u << u)) - 1;
                    const double m = dist_uni(flat, x, y, sx[o], sy[o]);
                    w = p.toy ? (m == norm) : (m > Cr);
                } else w = reward_weight(p, flat, sx, sy, own, norm);
            }
            rew = collision_reward_step(p.reward_design, my_tot, w);
        }
    } else if (MODE == MODE_DESIGN) {
        if (my_tot <= 1) rew = 1.0;
        else {   // TestEnv.calculate_reward_design (test_env.py:319-349)
            int k = 1, last = u;
            for (unsigned m = own; m; m &= m - 1) {
                const int t = __ffs(m) - 1;
                if (t != u && dist_uni(flat, x, y, sx[t], sy[t]) < p.C2) { ++k; last = t; }
            }
            if (k == 1) rew = 1.0;
            else if (k == 2) rew = (dist_uni(flat, x, y, sx[last], sy[last]) > p.C2) ? 0.0 : -2.0;
            else rew = -(double)k;
        }
    } else {
        // PRR (test_env.py:384-405): receivers in range of this transmitter = its own in-range bits
        // outside its collision set (distance is symmetric bit for bit); packets received were counted
        // by the receivers into recv_s
        __syncwarp(gmask);
        const int in_range = __popc(inr_mask & ~own & live_mask);
        rew = channel_reward(p.reward_design, max(my_tot, 1), (int)recv_s[u], in_range);
    }

    // ---- D: mobility -------------------------------------------------------------------------------
    const double x_new = act ? mobility_step_at(p, x, v, u, timestep) : 0.0;

    // ---- C2/E: per slab -- tick, replay the merge script, gather xpos, age, write back, VPD ---------
    int m_cnt = 0;
    if (vpd && lead) {
        for (int k = 0; k < B; ++k) hist[k * G + u] = 0u;
        if (SP) mcnt_s[u] = 0u;
    }
    if (SP) __syncthreads();                 // every warp has read the old positions; the shared histogram is clear
    if (act && p.mobility && lead) p.pos_x[vbase + u] = x_new;
    __syncwarp(gmask);                       // script rows written by every lane are read by its own lane only,
                                             // but the histogram zeroing above must precede the slab updates
    const double W = p.W, inv_binw = p.inv_binw;
    const int age_thr = p.age_threshold;
    auto do_slab = [&](int jbase) {
        // loads first: last_updated / xpos of this slab are consumed only after the replay loop, and
        // the NEXT slab's seq column is requested now, so neither latency is exposed
        int sb[SL], lb[SL]; double xb[SL];
#pragma unroll
        for (int q = 0; q < SL; ++q) {
            const int j = jbase + q;
            sb[q] = s_next[q];
            if (FULL || (j < N && act)) { lb[q] = lup[j * N]; xb[q] = xp[j * N]; }
            else { lb[q] = 0; xb[q] = 0.0; }
        }
        if (jbase + SLAB_STEP < G) {
#pragma unroll
            for (int q = 0; q < SL; ++q) {
                const int j = jbase + SLAB_STEP + q;
                s_next[q] = (FULL || (j < N && act)) ? seqp[j * N] : 0;
            }
        }
        int snv[SL], org[SL];                    // after the replay: sequence number and origin row of every entry
        // Two columns per register when every sequence number of the slab is either 0 (never heard) or
        // within FMAX slots of the newest possible one: then  fresh = seq - base  fits 16 - SB bits, the
        // order of the packed halves equals the order of the 32-bit keys, and one shuffle + one
        // VIMNMX.U16x2 merge two entries.  Slabs holding older information take the 32-bit path.
        constexpr int FMAX = (1 << (16 - SB)) - 1;
        const int base = tick - FMAX;
        // No sequence number exceeds the tick (a vehicle's own), so "0 or within FMAX of the newest" is
        // "every non-zero one is above base": one unsigned minimum of seq - 1 (0 wraps to the maximum) per column
        unsigned oldest = 0xffffffffu;
#pragma unroll
        for (int q = 0; q < SL; ++q) {
            if (jbase + q == u && act) sb[q] += 1;                                // vehicle.py:58 (tick)
            oldest = min(oldest, (unsigned)(sb[q] - 1));
        }
        // (lanes beyond N hold no entries: they must not veto the packed form for the whole group)
        const bool narrow = !act || base <= 0 || oldest >= (unsigned)base;
        if ((SL & 1) == 0 && (__ballot_sync(gmask, !narrow) & gmask) == 0u) {
            unsigned k2[SL / 2 > 0 ? SL / 2 : 1];
#pragma unroll
            for (int i = 0; i < SL / 2; ++i) {
                const unsigned f0 = sb[2 * i] ? (unsigned)(sb[2 * i] - base) : 0u;
                const unsigned f1 = sb[2 * i + 1] ? (unsigned)(sb[2 * i + 1] - base) : 0u;
                k2[i] = ((f0 << SB) | (unsigned)u) | (((f1 << SB) | (unsigned)u) << 16);
            }
            for (int pp = 0; pp < npass; ++pp) {
                const int srcl = script[pp * G + u];
#pragma unroll
                for (int i = 0; i < SL / 2; ++i) k2[i] = __vmaxu2(k2[i], __shfl_sync(gmask, k2[i], srcl, G));
            }
#pragma unroll
            for (int q = 0; q < SL; ++q) {       // straight to (new sequence number, origin row): no 32-bit key is formed
                const unsigned hk = (q & 1) ? (k2[q / 2] >> 16) : (k2[q / 2] & 0xffffu);
                const unsigned f = hk >> SB;
                snv[q] = f ? (int)f + base : 0; org[q] = (int)(hk & (unsigned)(G - 1));
            }
        } else {
            unsigned key[SL];
#pragma unroll
            for (int q = 0; q < SL; ++q) key[q] = ((unsigned)sb[q] << SB) | (unsigned)u;
            // replay the passes in order on this slab's columns
            for (int pp = 0; pp < npass; ++pp) {
                const int srcl = script[pp * G + u];
#pragma unroll
                for (int q = 0; q < SL; ++q) key[q] = max(key[q], __shfl_sync(gmask, key[q], srcl, G));
            }
#pragma unroll
            for (int q = 0; q < SL; ++q) { snv[q] = (int)(key[q] >> SB); org[q] = (int)(key[q] & (unsigned)(G - 1)); }
        }
        // gather xpos from the origin row, age, write back (independent per column => ILP)
#pragma unroll
        for (int q = 0; q < SL; ++q) {
            if (jbase + q == u) { lb[q] = 0; xb[q] = x; } else lb[q] += 1;        // vehicle.py:59-70 (tick)
            const int sn = snv[q];
            const bool changed = sn != sb[q];                  // strictly newer version merged in
            const int src = changed ? org[q] : u;
            xb[q] = __shfl_sync(gmask, xb[q], src, G);
            if (changed) lb[q] = 0;                            // vehicle.py:47
            sb[q] = sn;
        }
#pragma unroll
        for (int q = 0; q < SL; ++q) {
            const int j = jbase + q;
            if (FULL || (j < N && act)) { seqp[j * N] = sb[q]; lup[j * N] = lb[q]; xp[j * N] = xb[q]; }
        }
        if (vpd) {
            // numpy.histogram bin (network.py:500): t = (s + W) * B / 2W is a few ulp from the
            // real-valued bin coordinate, so trunc(t) is NumPy's edge-corrected bin unless t sits within
            // 1e-6 of an integer; those (rare) samples are re-binned against the edges themselves.
            // Straight-line per column; the histogram update is a fire-and-forget shared-memory
            // reduction (lane-private column => conflict-free), so columns do not serialise on it.
            unsigned fix = 0u;
            auto vpd_cols = [&](auto flat_c) {
                constexpr bool FL0 = decltype(flat_c)::value;
#pragma unroll
                for (int q = 0; q < SL; ++q) {
                    const int j = jbase + q;
                    bool in = (FULL || (j < N && act)) && j != u && lb[q] < age_thr;   // network.py:547
                    double sv;
                    if (FL0) {             // dy == 0: signed distance is exactly xpos - own x
                        sv = __dsub_rn(xb[q], x_new);
                        in = in && fabs(sv) < W;                                // network.py:487
                    } else {
                        const double y1 = sb[q] > 0 ? sy[j] : 0.0;
                        const double d = dist_t<false>(xb[q], y1, x_new, y);
                        in = in && d < W;
                        sv = (__dsub_rn(xb[q], x_new) > 0.0) ? d : -d;
                    }
                    const double t = __dmul_rn(__dadd_rn(sv, W), inv_binw);
                    const double rt = __dadd_rn(__dadd_rn(t, 6755399441055744.0), -6755399441055744.0);
                    const bool near = fabs(__dsub_rn(t, rt)) < 1e-6;
                    const int kb = (in && !near) ? min(max((int)t, 0), B - 1) : B;
                    smem_red_inc(&hist[kb * G + u]);
                    if (in && near) fix |= 1u << q;
                    m_cnt += in ? 1 : 0;
                    xb[q] = sv;                                // keep the sample for the exact re-binning
                }
            };
            if (flat0) vpd_cols(std::true_type{}); else vpd_cols(std::false_type{});
            if (fix) {
#pragma unroll
                for (int q = 0; q < SL; ++q)
                    if ((fix >> q) & 1u) smem_red_inc(&hist[vpd_bin_edges(xb[q], W, inv_binw, B, s_edges) * G + u]);
            }
        }
    };
    if (p.piggy) {
#pragma unroll 1
        for (int jbase = slab0; jbase < G; jbase += SLAB_STEP) do_slab(jbase);
    }
    if (SP) {                                // the other warps' columns are in the histogram once the CTA has met
        if (vpd) atomicAdd(&mcnt_s[u], (unsigned)m_cnt);
        __syncthreads();
        if (!lead) return;
        if (vpd) m_cnt = (int)mcnt_s[u];
    }

    // ---- per-env metric accumulators -------------------------------------------------------------
    {
        double rs = act ? rew : 0.0;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) rs += __shfl_xor_sync(gmask, rs, o, G);
        const int nr = __reduce_add_sync(gmask, n_recv), np = __reduce_add_sync(gmask, n_pairs), nb = __reduce_add_sync(gmask, bad);
        if (u == 0) {    // reductions, not read-modify-writes: nothing waits for the old values
            atomicAdd(p.acc_reward + e, rs);
            unsigned long long *c = reinterpret_cast<unsigned long long *>(p.acc_count + e * ACC_COUNTS);
            atomicAdd(c + 0, (unsigned long long)nr); atomicAdd(c + 1, (unsigned long long)np);
            atomicAdd(c + 2, (unsigned long long)nb); atomicAdd(c + 3, 1ull);
        }
    }
    if (act) p.rews[vbase + u] = (float)rew;
    if (CNT && act) *reinterpret_cast<float *>(recS + (u + 1) * p.rec_stride - 4) = (float)rew;

    // ---- F: state rows (TestEnv.obtain_state) in shared memory -----------------------------------
    // Row u sits at st[u*S .. u*S+S) exactly as in global memory, so the copy-out is a plain
    // contiguous float4 stream.  (The scalar row writes bank-conflict for some S; they are few.)
    __syncwarp(gmask);                       // all histogram reductions of this group have landed
    if (want_state && act) {
        float *wp = direct ? p.state + (vbase + u) * S : st + u * S;
        if (p.add_action) {
            if (p.action_binary) {
                int r = 0;
                if ((S & 3) == 0) {          // rows are 16-byte aligned: four one-hot entries per store
                    for (; r + 4 <= R; r += 4, wp += 4)
                        *reinterpret_cast<float4 *>(wp) = make_float4(a == r ? 1.0f : 0.0f, a == r + 1 ? 1.0f : 0.0f,
                                                                      a == r + 2 ? 1.0f : 0.0f, a == r + 3 ? 1.0f : 0.0f);
                }
                for (; r < R; ++r) *wp++ = (a == r) ? 1.0f : 0.0f;
            } else *wp++ = (float)a;
        }
        if (p.add_channel_obs) { for (int r = 0; r < R; ++r) *wp++ = obs_row[r]; }
        if (p.piggy) {
            // counts / len with one correctly rounded reciprocal and two FMAs per bin: exact
            // (== RN(c / m)) for all 0 <= c <= m < 1024, checked exhaustively (tests/test_host.py)
            const float den = (float)m_cnt, rcp = __frcp_rn(den);
            const bool have = vpd && m_cnt > 0;
            for (int b0 = 0; b0 < B; b0 += 8) {      // loads first: the row stores below may alias them
                unsigned hv[8]; float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) hv[i] = (have && b0 + i < B) ? hist[(b0 + i) * G + u] : 0u;
                if (CNT) {                           // compact host format: the counts themselves, one byte per bin
                    unsigned char *cp = recS + u * p.rec_stride + b0;
                    if ((B & 3) == 0) {
                        *reinterpret_cast<unsigned *>(cp) = hv[0] | (hv[1] << 8) | (hv[2] << 16) | (hv[3] << 24);
                        if (b0 + 4 < B) *reinterpret_cast<unsigned *>(cp + 4) = hv[4] | (hv[5] << 8) | (hv[6] << 16) | (hv[7] << 24);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) if (b0 + i < B) cp[i] = (unsigned char)hv[i];
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float c = (float)hv[i];
                    const float q0 = __fmul_rn(c, rcp);
                    f[i] = have ? __fmaf_rn(__fmaf_rn(-q0, den, c), rcp, q0) : 0.0f;
                }
                if (direct) {                        // B % 4 == 0: whole float4 groups
                    *reinterpret_cast<float4 *>(wp) = make_float4(f[0], f[1], f[2], f[3]); wp += 4;
                    if (b0 + 4 < B) { *reinterpret_cast<float4 *>(wp) = make_float4(f[4], f[5], f[6], f[7]); wp += 4; }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) if (b0 + i < B) *wp++ = f[i];
                }
            }
        }
        if (p.add_reward) *wp++ = (float)rew;
        if (p.add_index) *wp++ = (float)(u + 1);
        if (p.add_position) { *wp++ = (float)__ddiv_rn(x_new, p.L); *wp++ = (float)__ddiv_rn(y, 2.0); }
        if (p.add_velocity) *wp++ = (float)v;
        if (p.fingerprint) { *wp++ = (float)p.episode; *wp++ = (float)p.epsilon; }
    }
    __syncwarp(gmask);

    // coalesced copy-out of the [N][R] observation block and the [N][S] state block
    auto copy_out = [&](float *dst, const float *src, int n) {
        if ((n & 3) == 0) {
            for (int i = u; i < n / 4; i += G) reinterpret_cast<float4 *>(dst)[i] = reinterpret_cast<const float4 *>(src)[i];
        } else {
            for (int i = u; i < n; i += G) dst[i] = src[i];
        }
    };
    copy_out(p.obs + vbase * R, obsS, N * R);
    if (want_state && !direct) copy_out(p.state + vbase * S, st, N * S);
    // the environment's host records leave as one contiguous stream (they may sit in mapped host memory: 16-byte pieces)
    if (CNT) {
        const unsigned rec_bytes = (unsigned)(N * p.rec_stride);
        if ((rec_bytes & 15u) == 0u) {
            // one bulk store per environment (TMA): the whole record block leaves shared memory as a single asynchronous
            // copy -- towards mapped host memory that is one stream of full-line PCIe writes instead of 16-byte lane stores
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // this lane's recS stores -> async proxy
            __syncwarp(gmask);
            if (u == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(p.vpd_counts + vbase * p.rec_stride), "r"((unsigned)__cvta_generic_to_shared(recS)), "r"(rec_bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // written, not merely read out of shared memory
            }
        } else {
            copy_out(reinterpret_cast<float *>(p.vpd_counts + vbase * p.rec_stride), reinterpret_cast<const float *>(recS), N * (p.rec_stride >> 2));
        }
    }
    __syncwarp(gmask);                       // the next slot reuses the shared-memory staging
    if (CNT) { if (p.chunk_flag && u == 0) env_records_done(p, e); }
    }   // slot
}

template <int G>
size_t smem_bytes(const Params &p, int warps)
{
    const GroupSmem lay(G, p.R, p.B, p.S, p.build_state != 0, p.vpd_enabled != 0, group_direct_rows(p), p.vpd_counts ? p.rec_stride : 0);
    return (size_t)align16i(8 * (p.B + 1)) + (size_t)warps * (32 / G) * lay.bytes;
}

// one warp per CTA keeps the tail of the last wave short when few groups share a warp (G >= 16);
// small groups pack 4 warps so that a CTA still carries a useful number of environments
template <int G> struct WarpsFor { static constexpr int v = G >= 16 ? DIRAL_GROUP_WARPS : 4; };

// (2 warps per split environment: measured for the streamed host records, where spreading the completions over the
//  launch matters more than the launch time -- blocking diral_step_host 137-141 -> 114-118 us per C3 slot, 4 warps
//  126-133 us; see profiles/README.md)
#ifndef DIRAL_SPLIT_WARPS
#define DIRAL_SPLIT_WARPS 2
#endif
constexpr size_t SPLIT_SMEM_LIMIT = 227 * 1024;
constexpr int SPLIT_WARPS = DIRAL_SPLIT_WARPS;      // warps that share one split environment (they interleave the 8-column slabs)

// the same parameter block restricted to envs [e0, e0 + n) (every per-env array the lane-group kernel touches)
Params env_slice(const Params &p, long long e0, long long n)
{
    Params q = p;
    const long long N = p.N, NN = N * N;
    q.E = n; q.env0 = p.env0 + e0;
    if (q.actions) q.actions += e0 * N;
    if (q.actions_out) q.actions_out += e0 * N;
    q.pos_x += e0 * N; q.pos_y += e0 * N; q.vel += e0 * N;
    if (q.tab_seq) { q.tab_seq += e0 * NN; q.tab_lu += e0 * NN; q.tab_x += e0 * NN; }
    if (q.lat) q.lat += e0 * NN;
    q.obs += e0 * N * p.R; q.rews += e0 * N; q.state += e0 * N * p.S;
    if (q.vpd_counts) q.vpd_counts += e0 * N * p.rec_stride;
    q.acc_reward += e0; q.acc_count += e0 * ACC_COUNTS;
    return q;
}

template <int G, bool FULL, int MODE, bool LAT>
cudaError_t prepare_k(const Params &p)
{
    Params q = p; q.build_state = 1;         // the largest carve-up this configuration can ask for (host records included)
    q.vpd_counts = reinterpret_cast<uint8_t *>(1); q.rec_stride = (((p.piggy ? p.B : 0) + 3) & ~3) + 4;
    const size_t smem = smem_bytes<G>(q, WarpsFor<G>::v);
    if (smem <= 48 * 1024) return cudaSuccess;
    cudaError_t err = cudaFuncSetAttribute(step_group_kernel<G, FULL, WarpsFor<G>::v, MODE, LAT, false, false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    err = cudaFuncSetAttribute(step_group_kernel<G, FULL, WarpsFor<G>::v, MODE, LAT, false, true>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    err = cudaFuncSetAttribute(step_group_kernel<G, FULL, WarpsFor<G>::v, MODE, LAT, true, false>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if constexpr (G == 32) {
        const size_t smem_sp = smem_bytes<G>(q, SPLIT_WARPS);          // (beyond the limit: launch_k never splits)
        if (err == cudaSuccess && smem_sp > 48 * 1024 && smem_sp <= SPLIT_SMEM_LIMIT)
            err = cudaFuncSetAttribute(step_group_kernel<G, FULL, SPLIT_WARPS, MODE, LAT, false, false, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sp);
        if (err == cudaSuccess && smem_sp > 48 * 1024 && smem_sp <= SPLIT_SMEM_LIMIT)
            err = cudaFuncSetAttribute(step_group_kernel<G, FULL, SPLIT_WARPS, MODE, LAT, false, true, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sp);
    }
    return err;
}

// CTAs of this instantiation the current device holds at once (cached per instantiation, device and smem size)
template <int G, bool FULL, int W, int MODE, bool LAT, bool ROLL, bool CNT, bool SP = false>
long long resident_ctas(size_t smem)
{
    constexpr int MAX_DEV = 64;
    static std::mutex mu;
    static size_t cached_smem[MAX_DEV]; static long long cached[MAX_DEV];      // zero-initialised: 0 = not cached
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = dev >= 0 && dev < MAX_DEV ? dev : 0;
    std::lock_guard<std::mutex> lock(mu);
    if (cached[slot] == 0 || smem != cached_smem[slot] || slot != dev) {
        int sms = 148, per_sm = 16;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_group_kernel<G, FULL, W, MODE, LAT, ROLL, CNT, SP>, W * 32, smem) != cudaSuccess)
            per_sm = 16;
        cached[slot] = (long long)sms * std::max(per_sm, 1); cached_smem[slot] = smem;
    }
    return cached[slot];
}

template <int G, bool FULL, int MODE, bool LAT>
cudaError_t launch_k(const Params &p, cudaStream_t stream)
{
    constexpr int W = WarpsFor<G>::v;
    const long long envs_per_cta = (long long)W * (32 / G);
    const long long grid = (p.E + envs_per_cta - 1) / envs_per_cta;
    size_t smem = smem_bytes<G>(p, W);
    Params q = p;
    const bool roll = p.n_slots > 1, cnt = p.vpd_counts != nullptr;     // (the fused rollout has no host-record output)
    if (roll) {
        q.prefetch_ahead = (int)(resident_ctas<G, FULL, W, MODE, LAT, true, false>(smem) * envs_per_cta);
        step_group_kernel<G, FULL, W, MODE, LAT, true, false><<<(unsigned)grid, W * 32, smem, stream>>>(q);
        return cudaGetLastError();
    }
    auto launch_one = [&](const Params &r, auto cnt_c) {
        constexpr bool C = decltype(cnt_c)::value;
        Params t = r;
        t.prefetch_ahead = (int)(resident_ctas<G, FULL, W, MODE, LAT, false, C>(smem) * envs_per_cta);
        const long long g = (r.E + envs_per_cta - 1) / envs_per_cta;
        step_group_kernel<G, FULL, W, MODE, LAT, false, C><<<(unsigned)g, W * 32, smem, stream>>>(t);
    };
    if constexpr (G == 32) if (p.tail_split != 0 && smem_bytes<G>(p, SPLIT_WARPS) <= SPLIT_SMEM_LIMIT) {
        // Tail splitting.  One warp per environment makes a slot cost `waves x (latency of one environment)`: a batch
        // that fills the device 1.x times pays for 2.  The whole waves run one warp per environment; the remainder
        // runs SPLIT_WARPS warps per environment (shorter latency) when those fit the device at once.
        const long long slots = cnt ? resident_ctas<G, FULL, W, MODE, LAT, false, true>(smem)
                                    : resident_ctas<G, FULL, W, MODE, LAT, false, false>(smem);
        const long long full = p.tail_split == 2 ? 0 : (p.E / slots) * slots, rem = p.E - full;
        const size_t smem_sp = smem_bytes<G>(p, SPLIT_WARPS);
        const long long slots_sp = cnt ? resident_ctas<G, FULL, SPLIT_WARPS, MODE, LAT, false, true, true>(smem_sp)
                                       : resident_ctas<G, FULL, SPLIT_WARPS, MODE, LAT, false, false, true>(smem_sp);
        if (p.tail_split != 0 && rem > 0 && (p.tail_split == 2 || (full > 0 && rem <= slots_sp))) {
            if (full > 0) { const Params a = env_slice(p, 0, full); if (cnt) launch_one(a, std::true_type{}); else launch_one(a, std::false_type{}); }
            Params b = env_slice(p, full, rem);
            b.prefetch_ahead = 0;
            if (cnt) step_group_kernel<G, FULL, SPLIT_WARPS, MODE, LAT, false, true, true><<<(unsigned)rem, SPLIT_WARPS * 32, smem_sp, stream>>>(b);
            else step_group_kernel<G, FULL, SPLIT_WARPS, MODE, LAT, false, false, true><<<(unsigned)rem, SPLIT_WARPS * 32, smem_sp, stream>>>(b);
            return cudaGetLastError();
        }
    }
    if (cnt) launch_one(q, std::true_type{}); else launch_one(q, std::false_type{});
    return cudaGetLastError();
}

// run `prepare` (launch == false) or `launch` for the instantiation p selects
template <int G, bool FULL>
cudaError_t dispatch_mode(const Params &p, cudaStream_t stream, bool launch, int mode, bool lat)
{
#define DIRAL_CASE(M, L) return launch ? launch_k<G, FULL, M, L>(p, stream) : prepare_k<G, FULL, M, L>(p)
    if (mode == MODE_STEP)   { if (lat) DIRAL_CASE(MODE_STEP, true);   DIRAL_CASE(MODE_STEP, false); }
    if (mode == MODE_DESIGN) { if (lat) DIRAL_CASE(MODE_DESIGN, true); DIRAL_CASE(MODE_DESIGN, false); }
    if (lat) DIRAL_CASE(MODE_CH, true);
    DIRAL_CASE(MODE_CH, false);
#undef DIRAL_CASE
}

template <int G>
cudaError_t dispatch_g(const Params &p, cudaStream_t stream, bool launch, int mode, bool lat)
{
    return p.N == G ? dispatch_mode<G, true>(p, stream, launch, mode, lat)
                    : dispatch_mode<G, false>(p, stream, launch, mode, lat);
}

cudaError_t dispatch(const Params &p, cudaStream_t stream, bool launch, int mode, bool lat)
{
    switch (group_width(p.N)) {
    case 4: return dispatch_g<4>(p, stream, launch, mode, lat);
    case 8: return dispatch_g<8>(p, stream, launch, mode, lat);
    case 16: return dispatch_g<16>(p, stream, launch, mode, lat);
    default: return dispatch_g<32>(p, stream, launch, mode, lat);
    }
}

}  // namespace

int group_width(int N)
{
    return N <= 4 ? 4 : N <= 8 ? 8 : N <= 16 ? 16 : 32;
}

size_t step_group_smem_bytes(const Params &p)
{
    switch (group_width(p.N)) {
    case 4: return smem_bytes<4>(p, WarpsFor<4>::v);
    case 8: return smem_bytes<8>(p, WarpsFor<8>::v);
    case 16: return smem_bytes<16>(p, WarpsFor<16>::v);
    default: return smem_bytes<32>(p, WarpsFor<32>::v);
    }
}

cudaError_t prepare_step_group(const Params &p)
{
    for (int mode = 0; mode < 3; ++mode)
        for (int lat = 0; lat < 2; ++lat) {
            cudaError_t err = dispatch(p, nullptr, false, mode, lat != 0);
            if (err != cudaSuccess) return err;
        }
    return cudaSuccess;
}

cudaError_t launch_step_group(const Params &p, cudaStream_t stream)
{
    return dispatch(p, stream, true, p.mode, p.track_lat != 0 && p.lat != nullptr);
}

}  // namespace diral
