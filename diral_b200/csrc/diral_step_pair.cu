// diral_step_pair.cu -- fused time-slot kernel for 33 <= N <= 64 vehicles: ONE WARP PER ENVIRONMENT, TWO TABLE ROWS PER LANE.
//
// The lane-group kernel (diral_step_group.cu, N <= 32) keeps a table row per lane and replays the slot's merges with one
// shuffle + one integer max per entry; round 1 switched to the one-CTA-per-environment kernel at 33 vehicles and lost a
// factor of four there (barriers per resource pass, keys in shared memory).  This kernel keeps the lane-group design up
// to 64 vehicles: lane u owns vehicles u and u + 32 (rows u and u + 32 of the neighbour table).
//
//   A  inputs; every vehicle's in-range mask (64 bits) and the transmitter mask of every resource (two words)
//   B  DECISIONS, no table access, by walking the in-range vehicles of each owned row: a vehicle transmits on exactly one
//      resource, so an in-range t is the nearest transmitter of its resource (Network.find_closest_tx,
//      network.py:378-398) unless another in-range vehicle shares that resource -- then the lowest such candidate's
//      iteration runs the first-minimum search once.  Every reception writes one byte of the merge script
//      ("in pass p, row i takes row t"; identity elsewhere), the channel observation and, in PRR mode, the counters
//   C  the table streams once in slabs of 4 subject columns x 2 rows per lane (subject-major storage: every access is a
//      coalesced 128-byte segment per row half): tick, pack two columns per register as 16-bit keys
//      fresh << 6 | origin-row, replay the script in pass order -- per register and pass two shuffles per owned row
//      (the source lane's two rows) + a select + one VIMNMX.U16x2 -- decode, fetch the position of merged entries from
//      the origin row through a 2 KB shared-memory column buffer, age, write back, and bin the positional distribution
//      (network.py:473-513) with fire-and-forget shared-memory reductions.  Slabs holding entries older than 1023 slots
//      take 32-bit keys (one column per register)
//   D  rewards (lane-local from the collision masks), mobility, state rows written by the lane that owns them.
//
// (Tried and dropped: the whole 16-bit key table in shared memory with row-wise LDS.128 / VIMNMX merges instead of the shuffle
// replay -- 21 K instead of 27 K warp-instructions per environment, but 0.66 MB of shared-memory traffic per environment and a
// dependent LDS -> max -> STS chain per pass: 566-586 us against 414 us, profiles/README.md.)
//
// Same arithmetic rules as the other slot kernels (diral_dev.cuh); same HBM layout as the lane-group kernel (layout 0).
#include "diral_dev.cuh"
#include "diral_launch.h"

#include <algorithm>
#include <type_traits>

namespace diral {

namespace {

constexpr int PV = 64;                        // vehicles per environment, padded
constexpr int PSL = 4;                        // subject columns per slab
constexpr int PSB = 6;                        // low key bits: the origin row
constexpr unsigned PFULL = 0xffffffffu;

__host__ __device__ inline int align16p(int x) { return (x + 15) & ~15; }

struct PairSmem {
    int off_sx, off_sy, off_sa, off_txm, off_passof, off_recv, off_script, off_hist, off_colx, off_edges, bytes;
    __host__ __device__ PairSmem(int R, int B, bool vpd)
    {
        int o = 0;
        off_sx = o;     o += 8 * PV;
        off_sy = o;     o += 8 * PV;
        off_colx = o;   o += 8 * PSL * PV;                 // xpos of one slab, all 64 rows (origin-row gather)
        off_edges = o;  o += align16p(8 * (B + 1));
        off_sa = o;     o += 4 * PV;
        off_recv = o;   o += 4 * PV;
        off_txm = o;    o += align16p(8 * R);              // two words per resource
        off_passof = o; o += align16p(2 * R);              // index of a resource among the non-empty ones
        off_script = o; o += align16p(PV * R);             // [pass][row]
        off_hist = o;   o += vpd ? align16p(4 * PV * (B + 1)) : 0;     // [bin][row], + 1 dummy bin
        bytes = o;
    }
};

__device__ __noinline__ double pair_dist_slow(double dx, double dy)
{
    return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}
template <bool FLAT>
__device__ __forceinline__ double pdist(double x1, double y1, double x2, double y2)
{
    const double dx = __dsub_rn(x2, x1);
    if (FLAT) return fabs(dx);
    const double dy = __dsub_rn(y2, y1);
    if (dy == 0.0) return fabs(dx);
    return pair_dist_slow(dx, dy);
}

// Network.calculate_reward_weights / calculate_avg_distance (network.py:273-316), ascending ids, Python sum() semantics
__device__ __noinline__ int pair_reward_weight(const Params &p, const double *sx, const double *sy, unsigned m0, unsigned m1, double norm)
{
    PySum s; int pairs = 0;
    const unsigned long long m = (unsigned long long)m0 | ((unsigned long long)m1 << 32);
    for (unsigned long long mi = m; mi; mi &= mi - 1) {
        const int i = __ffsll((long long)mi) - 1;
        for (unsigned long long mj = mi & (mi - 1); mj; mj &= mj - 1) {
            const int j = __ffsll((long long)mj) - 1;
            s.add(dist2d(sx[i], sy[i], sx[j], sy[j]));
            ++pairs;
        }
    }
    const double mean = __ddiv_rn(s.result(), (double)pairs);
    return p.toy ? (mean == norm) : (mean > p.C);
}

__device__ __forceinline__ void pair_red_inc(unsigned *addr)
{
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(addr)));
}

// FULL (N == 64) drops every "does this row / column exist" predicate and turns the table strides into constants.
// (tuning knob, measured at 64 x 32 / 8192 envs: 18 resident warps = 96 registers 443 us, 24 = 80 registers with
//  spills 503 us, unconstrained = 128 registers 413 us -- like the lane-group kernel, fewer registers cost more than
//  the extra warps hide)
//  (an explicit minimum of 1 is not the same as none: ptxas then takes 185 registers, 10 warps / SM: 538 us)
#ifdef DIRAL_PAIR_MIN_BLOCKS
#define DIRAL_PAIR_BOUNDS __launch_bounds__(32, DIRAL_PAIR_MIN_BLOCKS)
#else
#define DIRAL_PAIR_BOUNDS __launch_bounds__(32)
#endif
template <int MODE, bool FULL>
__global__ void DIRAL_PAIR_BOUNDS step_pair_kernel(const Params p)
{
    const int u = threadIdx.x;
    const long long e = blockIdx.x;
    const int N = FULL ? PV : p.N;
    const int R = p.R, B = p.B, S = p.S;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const bool want_state = p.build_state != 0;
    const bool vpd = want_state && p.vpd_enabled;
    const PairSmem lay(R, B, p.vpd_enabled != 0);
    double *sx = reinterpret_cast<double *>(smem_raw + lay.off_sx);
    double *sy = reinterpret_cast<double *>(smem_raw + lay.off_sy);
    double *colx = reinterpret_cast<double *>(smem_raw + lay.off_colx);
    double *s_edges = reinterpret_cast<double *>(smem_raw + lay.off_edges);
    int *sa = reinterpret_cast<int *>(smem_raw + lay.off_sa);
    unsigned *recv_s = reinterpret_cast<unsigned *>(smem_raw + lay.off_recv);
    unsigned *txm_s = reinterpret_cast<unsigned *>(smem_raw + lay.off_txm);               // [R][2]
    unsigned short *passof = reinterpret_cast<unsigned short *>(smem_raw + lay.off_passof);
    unsigned char *script = smem_raw + lay.off_script;                                    // [pass][64]
    float *obsS = p.obs + e * (long long)N * R;                                           // [row][R], written in place (L1 / L2 resident)
    unsigned *hist = reinterpret_cast<unsigned *>(smem_raw + lay.off_hist);               // [B + 1][64]

    const bool lat_on = p.track_lat != 0 && p.lat != nullptr;
    const long long vbase = e * N, tbase = e * (long long)N * N;
    const bool act1 = FULL ? true : (u + 32 < N);   // row u always exists (N > 32)
    const long long timestep = p.timestep;
    const int tick = p.tick;
    const double Cr = p.C, sentinel = p.sentinel;
    const bool merge_mode = p.piggy && (MODE != MODE_STEP || p.state_type == 1 || p.state_type == 2);

    // ---- A: inputs ---------------------------------------------------------------------------------------------------
    for (int i = u; i <= B; i += 32) s_edges[i] = p.edges[i];
    int a[2] = {-1, -1}; double x[2] = {0.0, 0.0}, y[2] = {0.0, 0.0}, v[2] = {0.0, 0.0}; int bad = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = u + 32 * h;
        if (h == 0 || act1) {
            if (!p.gen_actions) a[h] = p.actions[vbase + i];
            x[h] = p.pos_x[vbase + i]; y[h] = p.pos_y[vbase + i]; v[h] = p.vel[vbase + i];
        }
    }
    if (p.piggy) {         // this environment's whole table towards L2 while the decisions run
        const char *b0 = reinterpret_cast<const char *>(p.tab_seq + tbase);
        const char *b1 = reinterpret_cast<const char *>(p.tab_lu + tbase);
        const char *b2 = reinterpret_cast<const char *>(p.tab_x + tbase);
        const int bytes4 = N * N * 4;
        for (int o = u * 128; o < bytes4; o += 32 * 128) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(b0 + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(b1 + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + bytes4 + o));
        }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = u + 32 * h;
        if (h == 0 || act1) {
            if (p.gen_actions) a[h] = philox_action(p.seed, i, p.env0 + e, timestep, R);
            if (a[h] < 0 || a[h] >= R) { bad += 1; a[h] = min(max(a[h], 0), R - 1); }
            if (p.gen_actions && p.actions_out) p.actions_out[vbase + i] = a[h];
        }
        sx[i] = x[h]; sy[i] = y[h]; sa[i] = (h == 0 || act1) ? a[h] : -1;
        recv_s[i] = 0u;
    }
    for (int i = u; i < 2 * R; i += 32) txm_s[i] = 0u;
    const double y0 = __shfl_sync(PFULL, y[0], 0);
    const bool flat = __ballot_sync(PFULL, y[0] != y0 || (act1 && y[1] != y0)) == 0u;
    const bool flat0 = flat && y0 == 0.0;
    __syncwarp();
    atomicOr(&txm_s[a[0] * 2], 1u << u);                               // test_env.py:149-157
    if (act1) atomicOr(&txm_s[a[1] * 2 + 1], 1u << u);
    __syncwarp();

    // toy reward: distance between the first-min-x and first-max-x vehicle (network.py:225-246)
    double norm = 0.0;
    if (MODE == MODE_STEP && p.toy) {
        double xmin = x[0], xmax = x[0]; int imin = u, imax = u;
        if (act1) { if (x[1] < xmin) { xmin = x[1]; imin = u + 32; } if (x[1] > xmax) { xmax = x[1]; imax = u + 32; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ox = __shfl_xor_sync(PFULL, xmin, o); const int oi = __shfl_xor_sync(PFULL, imin, o);
            if (ox < xmin || (ox == xmin && oi < imin)) { xmin = ox; imin = oi; }
            const double px = __shfl_xor_sync(PFULL, xmax, o); const int pi = __shfl_xor_sync(PFULL, imax, o);
            if (px > xmax || (px == xmax && pi < imax)) { xmax = px; imax = pi; }
        }
        norm = dist2d(sx[imin], sy[imin], sx[imax], sy[imax]);
    }

    // index of every resource among the non-empty ones (= its pass of the merge replay)
    int npass = 0;
    for (int r0 = 0; r0 < R; r0 += 32) {
        const int r = r0 + u;
        const bool busy = r < R && (txm_s[2 * r] | txm_s[2 * r + 1]) != 0u;
        const unsigned bal = __ballot_sync(PFULL, busy);
        if (r < R) passof[r] = (unsigned short)(npass + __popc(bal & ((1u << u) - 1u)));
        npass += __popc(bal);
    }
    // who is within communication range of my two vehicles (Network.check_communicaiton_range, network.py:595-607)
    unsigned inr[2][2] = {{0u, 0u}, {0u, 0u}};
    const unsigned live1 = (FULL || N >= 64) ? PFULL : ((1u << (N - 32)) - 1u);
    auto in_range_masks = [&](auto flat_c) {
        constexpr bool FL = decltype(flat_c)::value;
#pragma unroll 4
        for (int t = 0; t < 32; ++t) {
            const double xt0 = sx[t], yt0 = sy[t], xt1 = sx[t + 32], yt1 = sy[t + 32];
            inr[0][0] |= (pdist<FL>(xt0, yt0, x[0], y[0]) < Cr ? 1u : 0u) << t;
            inr[0][1] |= (pdist<FL>(xt1, yt1, x[0], y[0]) < Cr ? 1u : 0u) << t;
            inr[1][0] |= (pdist<FL>(xt0, yt0, x[1], y[1]) < Cr ? 1u : 0u) << t;
            inr[1][1] |= (pdist<FL>(xt1, yt1, x[1], y[1]) < Cr ? 1u : 0u) << t;
        }
    };
    if (flat) in_range_masks(std::true_type{}); else in_range_masks(std::false_type{});
    inr[0][1] &= live1; inr[1][1] &= live1;
    if (!act1) { inr[1][0] = 0u; inr[1][1] = 0u; }

    // channel observations before any reception (test_env.py:203-240 / :305-306 / :431), merge script = identity
    {
        const float basev = (MODE != MODE_STEP || p.state_type == 1) ? 1.0f : (p.state_type == 2 ? (float)sentinel : 0.0f);
        for (int r0 = 0; r0 < R; r0 += 32) {
            const int r = r0 + u;
            if (r < R) {
                const bool busy = (txm_s[2 * r] | txm_s[2 * r + 1]) != 0u;
                for (int i = 0; i < N; ++i) obsS[i * R + r] = (busy && sa[i] != r) ? basev : 0.0f;
            }
        }
        if (merge_mode)
            for (int pp = 0; pp < npass; ++pp) { script[pp * PV + u] = (unsigned char)u; script[pp * PV + u + 32] = (unsigned char)(u + 32); }
    }
    __syncwarp();

    // ---- B: decisions ---------------------------------------------------------------------------------------------------
    int n_recv = 0, n_pairs = 0;
    int32_t *latg = lat_on ? p.lat + tbase : nullptr;                   // lat[t][rx] = latg[t * N + rx]
    auto decisions = [&](auto flat_c) {
        constexpr bool FL = decltype(flat_c)::value;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h == 1 && !act1) continue;
            const int i = u + 32 * h, ai = a[h];
            const double xi = x[h], yi = y[h];
            auto reception = [&](int t, int at, double d) {             // row i hears t on resource at
                ++n_recv;
                if (MODE == MODE_STEP && p.state_type == 2) obsS[i * R + at] = (float)d;
                if (MODE == MODE_CH) {
                    atomicAdd(&recv_s[t], 1u);                                          // test_env.py:396-397
                    if (lat_on) latg[t * N + i] = (int32_t)timestep;                    // test_env.py:436
                }
                if (merge_mode) script[passof[at] * PV + i] = (unsigned char)t;
            };
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                unsigned m = inr[h][w];
                if (w == h) m &= ~(1u << u);                                            // not myself
                for (; m; m &= m - 1) {
                    const int t = w * 32 + __ffs(m) - 1, at = sa[t];
                    if (at == ai) continue;                                             // i transmits there itself (half duplex)
                    const unsigned c0 = inr[h][0] & txm_s[2 * at], c1 = inr[h][1] & txm_s[2 * at + 1];
                    const int nc = __popc(c0) + __popc(c1);
                    ++n_pairs;
                    if (nc == 1) {
                        const double d = pdist<FL>(sx[t], sy[t], xi, yi);
                        if (d < sentinel) reception(t, at, d);                          // best starts at the sentinel (network.py:380)
                    } else if ((c0 ? __ffs(c0) - 1 : 32 + __ffs(c1) - 1) == t) {
                        // several candidates: ascending ids, strict '<' -- the first minimum wins (network.py:384-391)
                        double best = sentinel; int tstar = -1;
                        for (unsigned c = c0; c; c &= c - 1) {
                            const int t2 = __ffs(c) - 1;
                            const double d2 = pdist<FL>(sx[t2], sy[t2], xi, yi);
                            if (d2 < best) { best = d2; tstar = t2; }
                        }
                        for (unsigned c = c1; c; c &= c - 1) {
                            const int t2 = 32 + __ffs(c) - 1;
                            const double d2 = pdist<FL>(sx[t2], sy[t2], xi, yi);
                            if (d2 < best) { best = d2; tstar = t2; }
                        }
                        if (tstar >= 0) reception(tstar, at, best);
                    }
                }
                if (lat_on) {                                                           // network.py:394
                    const unsigned lv = w == 0 ? PFULL : live1;
                    for (unsigned c = ~inr[h][w] & lv; c; c &= c - 1) {
                        const int t = w * 32 + __ffs(c) - 1;
                        if (sa[t] != ai) latg[t * N + i] = -1;
                    }
                }
            }
        }
    };
    if (flat) decisions(std::true_type{}); else decisions(std::false_type{});
    __syncwarp();

    // ---- rewards (test_env.py:159-199 / :294-302 / :408-429), lane-local from the collision masks ------------------------
    double rew[2] = {0.0, 0.0};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (h == 1 && !act1) continue;
        const int i = u + 32 * h;
        const unsigned o0 = txm_s[2 * a[h]], o1 = txm_s[2 * a[h] + 1];
        const int my_tot = __popc(o0) + __popc(o1);
        if (MODE == MODE_STEP) {
            if (my_tot <= 1) rew[h] = 1.0;
            else {
                int wgt = 0;
                if (design_needs_weight(p.reward_design, my_tot)) wgt = pair_reward_weight(p, sx, sy, o0, o1, norm);
                rew[h] = collision_reward_step(p.reward_design, my_tot, wgt);
            }
        } else if (MODE == MODE_DESIGN) {
            if (my_tot <= 1) rew[h] = 1.0;
            else {   // TestEnv.calculate_reward_design (test_env.py:319-349)
                int k = 1, last = i;
                const unsigned long long m = (unsigned long long)o0 | ((unsigned long long)o1 << 32);
                for (unsigned long long mm = m; mm; mm &= mm - 1) {
                    const int t = __ffsll((long long)mm) - 1;
                    if (t != i && dist2d(x[h], y[h], sx[t], sy[t]) < p.C2) { ++k; last = t; }
                }
                if (k == 1) rew[h] = 1.0;
                else if (k == 2) rew[h] = (dist2d(x[h], y[h], sx[last], sy[last]) > p.C2) ? 0.0 : -2.0;
                else rew[h] = -(double)k;
            }
        } else {
            // PRR (test_env.py:384-405): receivers in range = own in-range bits outside the collision set
            const int in_range = __popc(inr[h][0] & ~o0) + __popc(inr[h][1] & ~o1 & live1);
            rew[h] = channel_reward(p.reward_design, max(my_tot, 1), (int)recv_s[i], in_range);
        }
    }

    // ---- mobility (Network.update_positions, network.py:189-206) ------------------------------------------------------------
    double x_new[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = u + 32 * h;
        const bool on = h == 0 || act1;
        x_new[h] = on ? mobility_step_at(p, x[h], v[h], i, timestep) : 0.0;
        if (on && p.mobility) p.pos_x[vbase + i] = x_new[h];
    }

    // ---- C: the table, slab by slab ---------------------------------------------------------------------------------------
    int m_cnt[2] = {0, 0};
    if (vpd) { for (int k = 0; k <= B; ++k) { hist[k * PV + u] = 0u; hist[k * PV + u + 32] = 0u; } }
    __syncwarp();
    if (p.piggy) {
        int32_t *seqg = p.tab_seq + tbase, *lug = p.tab_lu + tbase;
        double *xg = p.tab_x + tbase;
        const double W = p.W, inv_binw = p.inv_binw;
        const int age_thr = p.age_threshold;
        constexpr int FMAX = (1 << (16 - PSB)) - 1;
        const int base = tick - FMAX;
        // sequence numbers are loaded one slab ahead (the key formation needs them first); a slab's ages and positions are
        // requested when the slab starts and consumed after the replay loop, whose latency chain hides them
        int s_next[2][PSL];
        auto load_seq = [&](int jb) {
#pragma unroll
            for (int q = 0; q < PSL; ++q) {
                const int j = jb + q;
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    s_next[h][q] = (FULL || (j < N && (h == 0 || act1))) ? seqg[j * N + u + 32 * h] : 0;
            }
        };
        auto do_slab = [&](int jbase) {
            int sb[2][PSL], lb[2][PSL]; double xb[2][PSL];
            unsigned oldest = 0xffffffffu;
#pragma unroll
            for (int q = 0; q < PSL; ++q) {
                const int j = jbase + q;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = u + 32 * h;
                    sb[h][q] = s_next[h][q];
                    if (FULL || (j < N && (h == 0 || act1))) { lb[h][q] = lug[j * N + i]; xb[h][q] = xg[j * N + i]; }
                    else { lb[h][q] = 0; xb[h][q] = 0.0; }
                }
            }
            if (jbase + PSL < N) load_seq(jbase + PSL);
#pragma unroll
            for (int q = 0; q < PSL; ++q)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (jbase + q == u + 32 * h && (FULL || h == 0 || act1)) sb[h][q] += 1;        // vehicle.py:58 (tick)
                    oldest = min(oldest, (unsigned)(sb[h][q] - 1));
                }
            const bool narrow = base <= 0 || oldest >= (unsigned)base;
            int snv[2][PSL], org[2][PSL];
            if (__ballot_sync(PFULL, !narrow) == 0u) {
                // two columns per register: fresh = seq - base fits 10 bits, the order of the packed halves is the order of
                // the 32-bit keys; per pass and register the source lane's two rows arrive by shuffle, a select picks the half
                unsigned k2[2][PSL / 2];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int w = 0; w < PSL / 2; ++w) {
                        const unsigned i = (unsigned)(u + 32 * h);
                        const unsigned f0 = sb[h][2 * w] ? (unsigned)(sb[h][2 * w] - base) : 0u;
                        const unsigned f1 = sb[h][2 * w + 1] ? (unsigned)(sb[h][2 * w + 1] - base) : 0u;
                        k2[h][w] = ((f0 << PSB) | i) | (((f1 << PSB) | i) << 16);
                    }
                for (int pp = 0; pp < npass; ++pp) {
                    const int s0 = script[pp * PV + u], s1 = script[pp * PV + u + 32];
#pragma unroll
                    for (int w = 0; w < PSL / 2; ++w) {
                        const unsigned a0 = __shfl_sync(PFULL, k2[0][w], s0), a1 = __shfl_sync(PFULL, k2[1][w], s0);
                        const unsigned b0 = __shfl_sync(PFULL, k2[0][w], s1), b1 = __shfl_sync(PFULL, k2[1][w], s1);
                        k2[0][w] = __vmaxu2(k2[0][w], (s0 & 32) ? a1 : a0);
                        k2[1][w] = __vmaxu2(k2[1][w], (s1 & 32) ? b1 : b0);
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int q = 0; q < PSL; ++q) {
                        const unsigned hk = (q & 1) ? (k2[h][q / 2] >> 16) : (k2[h][q / 2] & 0xffffu), f = hk >> PSB;
                        snv[h][q] = f ? (int)f + base : 0; org[h][q] = (int)(hk & (PV - 1));
                    }
            } else {
                unsigned key[2][PSL];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int q = 0; q < PSL; ++q) key[h][q] = ((unsigned)sb[h][q] << PSB) | (unsigned)(u + 32 * h);
                for (int pp = 0; pp < npass; ++pp) {
                    const int s0 = script[pp * PV + u], s1 = script[pp * PV + u + 32];
#pragma unroll
                    for (int q = 0; q < PSL; ++q) {
                        const unsigned a0 = __shfl_sync(PFULL, key[0][q], s0), a1 = __shfl_sync(PFULL, key[1][q], s0);
                        const unsigned b0 = __shfl_sync(PFULL, key[0][q], s1), b1 = __shfl_sync(PFULL, key[1][q], s1);
                        key[0][q] = max(key[0][q], (s0 & 32) ? a1 : a0);
                        key[1][q] = max(key[1][q], (s1 & 32) ? b1 : b0);
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int q = 0; q < PSL; ++q) { snv[h][q] = (int)(key[h][q] >> PSB); org[h][q] = (int)(key[h][q] & (PV - 1)); }
            }
            // own entries (vehicle.py:59-70), then the slab's positions go through shared memory: a merged entry takes the
            // position its origin row held when the slot began (an entry's position is a pure function of (subject, seq))
#pragma unroll
            for (int q = 0; q < PSL; ++q)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (jbase + q == u + 32 * h) { lb[h][q] = 0; xb[h][q] = x[h]; } else lb[h][q] += 1;
                    colx[q * PV + u + 32 * h] = xb[h][q];
                }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < PSL; ++q)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (snv[h][q] != sb[h][q]) { xb[h][q] = colx[q * PV + org[h][q]]; lb[h][q] = 0; sb[h][q] = snv[h][q]; }   // vehicle.py:41-47
                }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < PSL; ++q) {
                const int j = jbase + q;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = u + 32 * h;
                    if (FULL || (j < N && (h == 0 || act1))) { seqg[j * N + i] = sb[h][q]; lug[j * N + i] = lb[h][q]; xg[j * N + i] = xb[h][q]; }
                }
            }
            if (vpd) {
                unsigned fix = 0u;
                auto vpd_cols = [&](auto flat_c) {
                    constexpr bool FL0 = decltype(flat_c)::value;
#pragma unroll
                    for (int q = 0; q < PSL; ++q) {
                        const int j = jbase + q;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int i = u + 32 * h;
                            bool in = (FULL || (j < N && (h == 0 || act1))) && j != i && lb[h][q] < age_thr;      // network.py:547
                            double sv;
                            if (FL0) { sv = __dsub_rn(xb[h][q], x_new[h]); in = in && fabs(sv) < W; }  // network.py:487
                            else {
                                const double y1 = sb[h][q] > 0 ? sy[min(j, N - 1)] : 0.0;
                                const double d = pdist<false>(xb[h][q], y1, x_new[h], y[h]);
                                in = in && d < W;
                                sv = (__dsub_rn(xb[h][q], x_new[h]) > 0.0) ? d : -d;
                            }
                            const double t = __dmul_rn(__dadd_rn(sv, W), inv_binw);
                            const double rt = __dadd_rn(__dadd_rn(t, 6755399441055744.0), -6755399441055744.0);
                            const bool near = fabs(__dsub_rn(t, rt)) < 1e-6;
                            const int kb = (in && !near) ? min(max((int)t, 0), B - 1) : B;
                            pair_red_inc(&hist[kb * PV + i]);
                            if (in && near) fix |= 1u << (2 * q + h);
                            m_cnt[h] += in ? 1 : 0;
                            xb[h][q] = sv;
                        }
                    }
                };
                if (flat0) vpd_cols(std::true_type{}); else vpd_cols(std::false_type{});
                if (fix) {
#pragma unroll
                    for (int q = 0; q < PSL; ++q)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            if ((fix >> (2 * q + h)) & 1u)
                                pair_red_inc(&hist[vpd_bin(xb[h][q], W, inv_binw, B, s_edges) * PV + u + 32 * h]);
                }
            }
        };
        load_seq(0);
#pragma unroll 1
        for (int jbase = 0; jbase < N; jbase += PSL) do_slab(jbase);
    }

    // ---- per-env metric accumulators ----------------------------------------------------------------------------------------
    {
        double rs = rew[0] + (act1 ? rew[1] : 0.0); int nr = n_recv, np = n_pairs, nb = bad;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            rs += __shfl_xor_sync(PFULL, rs, o);
            nr += __shfl_xor_sync(PFULL, nr, o);
            np += __shfl_xor_sync(PFULL, np, o);
            nb += __shfl_xor_sync(PFULL, nb, o);
        }
        if (u == 0) {
            atomicAdd(p.acc_reward + e, rs);
            unsigned long long *c = reinterpret_cast<unsigned long long *>(p.acc_count + e * ACC_COUNTS);
            atomicAdd(c + 0, (unsigned long long)nr); atomicAdd(c + 1, (unsigned long long)np);
            atomicAdd(c + 2, (unsigned long long)nb); atomicAdd(c + 3, 1ull);
        }
    }
    __syncwarp();                             // every histogram reduction has landed

    // ---- state rows (TestEnv.obtain_state, test_env.py:527-583), by the lane that owns the row --------------------------------
    const int n_act = p.add_action ? (p.action_binary ? R : 1) : 0;
    const bool vec = (S & 3) == 0 && (n_act & 3) == 0 && (!p.add_channel_obs || (R & 3) == 0) && (!p.piggy || (B & 3) == 0);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (h == 1 && !act1) continue;
        const int i = u + 32 * h;
        p.rews[vbase + i] = (float)rew[h];
        if (p.vpd_counts) *reinterpret_cast<float *>(p.vpd_counts + (vbase + i + 1) * p.rec_stride - 4) = (float)rew[h];
        if (!want_state) continue;
        float *wp = p.state + (vbase + i) * S;
        const int ai = a[h];
        if (p.add_action) {
            if (p.action_binary) {
                int r = 0;
                if (vec) for (; r + 4 <= R; r += 4, wp += 4)
                    *reinterpret_cast<float4 *>(wp) = make_float4(ai == r ? 1.0f : 0.0f, ai == r + 1 ? 1.0f : 0.0f,
                                                                  ai == r + 2 ? 1.0f : 0.0f, ai == r + 3 ? 1.0f : 0.0f);
                for (; r < R; ++r) *wp++ = (ai == r) ? 1.0f : 0.0f;
            } else *wp++ = (float)ai;
        }
        if (p.add_channel_obs) { for (int r = 0; r < R; ++r) *wp++ = __ldcg(obsS + i * R + r); }
        if (p.piggy) {
            const float den = (float)m_cnt[h], rcp = __frcp_rn(den);
            const bool have = vpd && m_cnt[h] > 0;
            unsigned char *cp = p.vpd_counts ? p.vpd_counts + (vbase + i) * p.rec_stride : nullptr;
            for (int b = 0; b < B; ++b) {
                const unsigned cnt = have ? hist[b * PV + i] : 0u;
                const float cf = (float)cnt, q0 = __fmul_rn(cf, rcp);
                *wp++ = have ? __fmaf_rn(__fmaf_rn(-q0, den, cf), rcp, q0) : 0.0f;
                if (cp) cp[b] = (unsigned char)cnt;
            }
        }
        if (p.add_reward) *wp++ = (float)rew[h];
        if (p.add_index) *wp++ = (float)(i + 1);
        if (p.add_position) { *wp++ = (float)__ddiv_rn(x_new[h], p.L); *wp++ = (float)__ddiv_rn(y[h], 2.0); }
        if (p.add_velocity) *wp++ = (float)v[h];
        if (p.fingerprint) { *wp++ = (float)p.episode; *wp++ = (float)p.epsilon; }
    }
}

size_t pair_smem(const Params &p) { return (size_t)PairSmem(p.R, p.B, p.vpd_enabled != 0).bytes; }

template <int MODE>
cudaError_t pair_prepare(size_t smem)
{
    if (smem <= 48 * 1024) return cudaSuccess;
    cudaError_t err = cudaFuncSetAttribute(step_pair_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err == cudaSuccess) err = cudaFuncSetAttribute(step_pair_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return err;
}

template <int MODE>
void pair_launch(const Params &p, size_t smem, cudaStream_t stream)
{
    if (p.N == PV) step_pair_kernel<MODE, true><<<(unsigned)p.E, 32, smem, stream>>>(p);
    else step_pair_kernel<MODE, false><<<(unsigned)p.E, 32, smem, stream>>>(p);
}

}  // namespace

size_t step_pair_smem_bytes(const Params &p) { return pair_smem(p); }

// 33..64 vehicles with neighbour tables or without; the fused State blocks only (the un-fused ones run the standalone
// obtain_state kernel afterwards, which reads the same layout)
bool step_pair_supported(const Params &p)
{
    return p.N > GROUP_MAX_N && p.N <= PV && pair_smem(p) <= (size_t)226 * 1024;
}

cudaError_t prepare_step_pair(const Params &p)
{
    const size_t smem = pair_smem(p);
    cudaError_t err = pair_prepare<MODE_STEP>(smem);
    if (err == cudaSuccess) err = pair_prepare<MODE_DESIGN>(smem);
    if (err == cudaSuccess) err = pair_prepare<MODE_CH>(smem);
    return err;
}

cudaError_t launch_step_pair(const Params &p, cudaStream_t stream)
{
    const size_t smem = pair_smem(p);
    if (p.mode == MODE_STEP) pair_launch<MODE_STEP>(p, smem, stream);
    else if (p.mode == MODE_DESIGN) pair_launch<MODE_DESIGN>(p, smem, stream);
    else pair_launch<MODE_CH>(p, smem, stream);
    return cudaGetLastError();
}

}  // namespace diral
