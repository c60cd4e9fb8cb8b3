// diral_step_block.cu -- fused time-slot kernel for 32 < N <= 256 vehicles (one CTA per environment).
//
// Same slot semantics and the same ideas as diral_step_group.cu, re-mapped for rows that no longer
// fit one warp.  Thread u is vehicle u; the packed table keys  seq << SB | origin-row  live in shared
// memory as K[u][j] with an odd row stride (or in the scratch buffer when 4*N*(N+1) bytes do not fit).
//
//   A   inputs, whole-table L2 prefetch, transmitter masks per resource (the per-resource collision
//       histogram, reference envs/test_env.py:149-157) by shared-memory atomicOr, keys from the seq
//       columns (coalesced), in-range bitmask of every vehicle (NW = N/32 words per thread)
//   C   for r = 0..R-1 in order:
//         every thread finds its nearest in-range transmitter on r (Network.find_closest_tx,
//         network.py:378-398) from  inr & txm[r]  and appends (receiver, transmitter) to the pass list;
//         ONE barrier; then the CTA turns around -- thread j owns table COLUMN j and applies the
//         pass's merges to it sequentially (Vehicle.received_update, vehicle.py:35-47):
//             K[rx][j] = max(K[rx][j], K[tx][j])
//         Columns are independent and a pass never modifies a transmitter's row, so no second barrier
//         is needed (the pass list is double-buffered) and the accesses are conflict-free.
//       channel observations leave through a per-warp 32x32 transpose tile as coalesced 128 B rows
//   D   mobility
//   E   thread u = observer again: stream the columns CB at a time -- gather xpos from the origin row
//       through a double-buffered shared-memory column buffer (one barrier per CB columns), age, write
//       back, bin the positional distribution (network.py:473-513) with shared-memory reductions
//   F   state rows are staged in the (now free) key region and written as contiguous float4s
//
// Reward models are lane-local exactly as in the group kernel: a thread's collision set is txm[a].
#include "diral_dev.cuh"
#include "diral_launch.h"

#include <algorithm>
#include <type_traits>

namespace diral {

namespace {

constexpr int CB = 2;                        // table columns per thread per epilogue barrier
constexpr unsigned short PAIR_NONE = 0xFFFFu;

__host__ __device__ inline size_t align16z(size_t x) { return (x + 15) & ~(size_t)15; }

struct BlockSmem {
    size_t off_sx, off_sy, off_sxn, off_edges, off_txm, off_pairs, off_misc, off_recv, off_red, off_union, off_keys, bytes;
    size_t union_bytes, keys_bytes;
    __host__ __device__ BlockSmem(int N, int R, int B, int TN, int H, bool vpd_state, bool keys_in_smem)
    {
        const int NW = TN / 32, T = TN;
        size_t o = 0;
        off_sx = o;    o += align16z(8 * (size_t)N);
        off_sy = o;    o += align16z(8 * (size_t)N);
        off_sxn = o;   o += align16z(8 * (size_t)N);          // post-mobility x, for the helper threads
        off_edges = o; o += align16z(8 * (size_t)(B + 1));
        off_txm = o;   o += align16z(4 * (size_t)R * NW);
        off_pairs = o; o += align16z(2 * 8 * (size_t)N);         // two pass lists of (rx row offset, tx row offset)
        off_misc = o;  o += 16;                                  // list lengths of passes pc, pc+1, pc+2, pc+3
        off_recv = o;  o += align16z(4 * (size_t)N);
        off_red = o;   o += align16z(8 * 4 * 32);
        // phase-disjoint: the obs transpose tiles (phase C) share space with histogram + column buffer (E)
        const size_t tiles = 4 * (size_t)NW * 32 * 33;
        const size_t epi = (vpd_state ? align16z(4 * (size_t)(B + 1) * T) : 0) + 2 * (size_t)H * CB * 8 * (size_t)N;
        union_bytes = align16z(tiles > epi ? tiles : epi);
        off_union = o; o += union_bytes;
        keys_bytes = keys_in_smem ? align16z(4 * (size_t)N * (N + 1)) : 0;
        off_keys = o;  o += keys_bytes;
        bytes = o;
    }
};

constexpr size_t SMEM_BUDGET = 220 * 1024;

// Network.calculate_reward_weights / calculate_avg_distance (network.py:273-316) over the
// transmitters whose bits are set in m[0..NW), ascending ids, Python sum() semantics
template <int NW>
__device__ __noinline__ int block_reward_weight(const Params &p, const double *sx, const double *sy,
                                                const unsigned *m, double norm)
{
    PySum s; int pairs = 0;
    for (int wi = 0; wi < NW; ++wi)
        for (unsigned mi = m[wi]; mi; mi &= mi - 1) {
            const int i = wi * 32 + __ffs(mi) - 1;
            for (int wj = wi; wj < NW; ++wj)
                for (unsigned mj = (wj == wi) ? (mi & (mi - 1)) : m[wj]; mj; mj &= mj - 1) {
                    const int j = wj * 32 + __ffs(mj) - 1;
                    s.add(dist2d(sx[i], sy[i], sx[j], sy[j]));
                    ++pairs;
                }
        }
    const double mean = __ddiv_rn(s.result(), (double)pairs);
    return p.toy ? (mean == norm) : (mean > p.C);
}

// H helper copies of the N-thread team: thread (h, u).  Team 0 makes the decisions; all teams share the
// column merges (receptions k = h, h+H, .. of a pass), the key loads and the epilogue column blocks.
#ifndef DIRAL_BLOCK_HELPERS
// measured on B200 (profiles/README.md): no helpers up to 64 vehicles, 4 teams up to 128, 2 beyond
template <int NW> struct Helpers { static constexpr int v = NW <= 2 ? 1 : (NW <= 4 ? 4 : 2); };
#else
template <int NW> struct Helpers { static constexpr int v = (NW * 32 * DIRAL_BLOCK_HELPERS <= 1024) ? DIRAL_BLOCK_HELPERS : 1024 / (NW * 32); };
#endif

template <int NW>
__global__ void __launch_bounds__(NW * 32 * Helpers<NW>::v, (NW * 32 * Helpers<NW>::v <= 512 ? 2 : 1))
step_block_kernel(const Params p, const int SB, const int keys_in_smem)
{
    constexpr int T = NW * 32;               // team size (threads that map to vehicles / columns)
    constexpr int H = Helpers<NW>::v;
    constexpr int TT = T * H;                // CTA size
    const int N = p.N, R = p.R, B = p.B, S = p.S;
    const int tid = threadIdx.x, h = tid / T, u = tid - h * T, lane = tid & 31, warp = u >> 5;
    const bool col = u < N;                  // a valid vehicle / column index
    const bool act = col && h == 0;          // ... and the thread that decides for it
    const long long e = blockIdx.x;
    const long long vbase = e * N, tbase = e * (long long)N * N;
    const bool want_state = p.build_state != 0;
    const bool vpd = want_state && p.vpd_enabled;
    const int ld = N + 1;
    const int mode = p.mode;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const BlockSmem lay(N, R, B, T, H, p.vpd_enabled != 0, keys_in_smem != 0);
    double *sx = reinterpret_cast<double *>(smem_raw + lay.off_sx);
    double *sy = reinterpret_cast<double *>(smem_raw + lay.off_sy);
    double *sxn = reinterpret_cast<double *>(smem_raw + lay.off_sxn);
    double *s_edges = reinterpret_cast<double *>(smem_raw + lay.off_edges);
    unsigned *txm_s = reinterpret_cast<unsigned *>(smem_raw + lay.off_txm);         // [R][NW]
    uint2 *pairs = reinterpret_cast<uint2 *>(smem_raw + lay.off_pairs);               // [2][N] byte offsets of (rx row, tx row)
    int *npairs = reinterpret_cast<int *>(smem_raw + lay.off_misc);                 // [4], indexed by pass & 3
    unsigned *recv_s = reinterpret_cast<unsigned *>(smem_raw + lay.off_recv);
    double *s_red = reinterpret_cast<double *>(smem_raw + lay.off_red);
    float *tile = reinterpret_cast<float *>(smem_raw + lay.off_union) + warp * 32 * 33;     // phase C
    unsigned *hist = reinterpret_cast<unsigned *>(smem_raw + lay.off_union);                // phase E
    double *colbuf = reinterpret_cast<double *>(smem_raw + lay.off_union + (p.vpd_enabled ? align16z(4 * (size_t)(B + 1) * T) : 0));
    unsigned *K = keys_in_smem ? reinterpret_cast<unsigned *>(smem_raw + lay.off_keys)
                               : p.scratch + (size_t)e * N * ld;
    const unsigned srcmask = (1u << SB) - 1u;

    // ---- A: inputs ---------------------------------------------------------------------------------
    int a = -1; double x = 0.0, y = 0.0, v = 0.0; int bad = 0;
    if (act) {
        if (!p.gen_actions) a = p.actions[vbase + u];
        x = p.pos_x[vbase + u]; y = p.pos_y[vbase + u]; v = p.vel[vbase + u];
    }
    if (p.piggy) {    // this environment's whole table towards L2 while the decisions run
        const char *b0 = reinterpret_cast<const char *>(p.tab_seq + tbase);
        const char *b1 = reinterpret_cast<const char *>(p.tab_lu + tbase);
        const char *b2 = reinterpret_cast<const char *>(p.tab_x + tbase);
        const int bytes4 = N * N * 4;
        for (int o = tid * 128; o < bytes4; o += TT * 128) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(b1 + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + bytes4 + o));
        }
        (void)b0;     // the seq columns are read right below
    }
    for (int i = tid; i <= B; i += TT) s_edges[i] = p.edges[i];
    for (int i = tid; i < R * NW; i += TT) txm_s[i] = 0u;
    if (tid < 4) npairs[tid] = 0;
    if (act) {
        if (p.gen_actions) a = philox_action(p.seed, u, p.env0 + e, p.timestep, R);
        if (a < 0 || a >= R) { bad = 1; a = min(max(a, 0), R - 1); }
        if (p.gen_actions && p.actions_out) p.actions_out[vbase + u] = a;
        sx[u] = x; sy[u] = y; recv_s[u] = 0u;
    }
    __syncthreads();
    if (act) atomicOr(&txm_s[a * NW + warp], 1u << lane);            // test_env.py:149-157
    if (p.piggy && col) {
        const int32_t *seqp = p.tab_seq + tbase + u;
#pragma unroll 8
        for (int j = h; j < N; j += H) {
            int s = seqp[j * N];
            if (j == u) s += 1;                                      // vehicle.py:58 (tick)
            K[u * ld + j] = ((unsigned)s << SB) | (unsigned)u;
        }
    }
    // every vehicle on the same lane of the highway?  (dy == 0 for every pair => dist == |dx| exactly)
    const double y0 = sy[0];
    const bool flat = __syncthreads_and(!act || y == y0) != 0;       // also publishes txm_s and K
    const bool flat0 = flat && y0 == 0.0;

    // who is within communication range of this vehicle (network.py:595-607)
    unsigned inr[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) inr[w] = 0u;
    const double Cr = p.C, sentinel = p.sentinel;
    if (act) {
        if (flat) {
#pragma unroll
            for (int w = 0; w < NW; ++w)
                for (int b = 0; b < 32 && w * 32 + b < N; ++b)
                    if (fabs(__dsub_rn(x, sx[w * 32 + b])) < Cr) inr[w] |= 1u << b;
        } else {
#pragma unroll
            for (int w = 0; w < NW; ++w)
                for (int b = 0; b < 32 && w * 32 + b < N; ++b)
                    if (dist2d(sx[w * 32 + b], sy[w * 32 + b], x, y) < Cr) inr[w] |= 1u << b;
        }
    }
    unsigned own[NW]; int my_tot = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) { own[w] = act ? txm_s[a * NW + w] : 0u; my_tot += __popc(own[w]); }

    // toy reward: first-min-x / first-max-x vehicle (network.py:225-246)
    double norm = 0.0;
    if (p.toy && mode == MODE_STEP && act && my_tot > 1 && design_needs_weight(p.reward_design, my_tot)) {
        double xmin = p.L + 1.0, xmax = -p.L - 1.0; int imin = 0, imax = 0;
        for (int t = 0; t < N; ++t) {
            if (sx[t] < xmin) { xmin = sx[t]; imin = t; }
            if (sx[t] > xmax) { xmax = sx[t]; imax = t; }
        }
        norm = dist2d(sx[imin], sy[imin], sx[imax], sy[imax]);
    }

    // ---- C: resources in ascending order -----------------------------------------------------------
    int n_recv = 0, n_pairs = 0;
    int32_t *latp = p.track_lat ? p.lat + tbase + u : nullptr;       // lat[t][u] = latp[t * N]
    const bool merge_mode = p.piggy && (mode != MODE_STEP || p.state_type == 1 || p.state_type == 2);
    float *og = p.obs + vbase * R;
    int pc = 0;                               // non-empty passes so far (uniform)
    auto flush_tile = [&](int r_end) {       // rows of the 32x32 tile -> coalesced 128 B segments of obs
        if (h != 0) return;
        const int r0 = (r_end - 1) & ~31, nr = r_end - r0;
        __syncwarp();
        for (int i = 0; i < 32; ++i) {
            const int uu = warp * 32 + i;
            if (uu < N && lane < nr) og[(long long)uu * R + r0 + lane] = tile[i * 33 + lane];
        }
        __syncwarp();
    };
    for (int r = 0; r < R; ++r) {
        unsigned txm[NW]; unsigned any = 0u;
#pragma unroll
        for (int w = 0; w < NW; ++w) { txm[w] = txm_s[r * NW + w]; any |= txm[w]; }
        float o = 0.0f;
        if (any != 0u) {
            const bool is_rx = act && a != r;
            // candidates = in-range transmitters; nearest in ascending id with strict '<' (first wins)
            double best = sentinel; int tstar = -1;
            if (is_rx) {
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    unsigned c = inr[w] & txm[w];
                    n_pairs += __popc(c);
                    for (; c; c &= c - 1) {
                        const int t = w * 32 + __ffs(c) - 1;
                        const double d = flat ? fabs(__dsub_rn(x, sx[t])) : dist2d(sx[t], sy[t], x, y);
                        if (d < best) { best = d; tstar = t; }
                    }
                }
                if (latp) {                                                          // network.py:394
#pragma unroll
                    for (int w = 0; w < NW; ++w)
                        for (unsigned c = txm[w] & ~inr[w]; c; c &= c - 1) latp[(w * 32 + __ffs(c) - 1) * N] = -1;
                    if (mode == MODE_CH && tstar >= 0) latp[tstar * N] = (int32_t)p.timestep;   // test_env.py:436
                }
                if (tstar >= 0) {
                    ++n_recv;
                    if (mode == MODE_CH) atomicAdd(&recv_s[tstar], 1u);             // test_env.py:396-397
                }
                if (mode == MODE_STEP) o = p.state_type == 2 ? (float)best : (p.state_type == 1 ? 1.0f : 0.0f);
                else o = 1.0f;
            }
            // pass list, warp-aggregated append (order inside a pass is irrelevant)
            if (merge_mode) {
                const bool has = tstar >= 0;
                const int parity = pc & 1, slot = pc & 3;
                const unsigned bm = __ballot_sync(0xffffffffu, has);
                int basep = 0;
                if (lane == 0 && bm) basep = atomicAdd(&npairs[slot], __popc(bm));
                basep = __shfl_sync(0xffffffffu, basep, 0);
                if (has) pairs[parity * N + basep + __popc(bm & ((1u << lane) - 1u))] = make_uint2((unsigned)(u * ld * 4), (unsigned)(tstar * ld * 4));
                // the counter of pass pc+2: its last readers (pass pc-2) are all past barrier pc-1, and its
                // next writers come after barrier pc+1
                if (tid == 0) npairs[(pc + 2) & 3] = 0;
                __syncthreads();
                const int np = npairs[slot];
                if (col) {           // team h applies receptions h, h+H, .. of this pass to column j = u
                    const uint2 *pl = pairs + parity * N;
                    char *Kj = reinterpret_cast<char *>(K + u);
                    auto at = [&](unsigned off) -> unsigned & { return *reinterpret_cast<unsigned *>(Kj + off); };
                    int k = h;
                    for (; k + 3 * H < np; k += 4 * H) {     // receptions of one pass are independent
                        const uint2 p0 = pl[k], p1 = pl[k + H], p2 = pl[k + 2 * H], p3 = pl[k + 3 * H];
                        const unsigned a0 = at(p0.x), b0 = at(p0.y), a1 = at(p1.x), b1 = at(p1.y);
                        const unsigned a2 = at(p2.x), b2 = at(p2.y), a3 = at(p3.x), b3 = at(p3.y);
                        at(p0.x) = max(a0, b0); at(p1.x) = max(a1, b1);
                        at(p2.x) = max(a2, b2); at(p3.x) = max(a3, b3);
                    }
                    for (; k < np; k += H) {
                        const uint2 p0 = pl[k];
                        at(p0.x) = max(at(p0.x), at(p0.y));
                    }
                }
                ++pc;
            }
        }
        if (h == 0) tile[lane * 33 + (r & 31)] = o;
        if ((r & 31) == 31 || r == R - 1) flush_tile(r + 1);
    }
    __syncthreads();                          // keys final; recv_s complete; tiles free for the epilogue

    // rewards (test_env.py:159-199 / :294-302 / :408-429), all lane-local
    double rew = 0.0;
    if (act) {
        if (mode == MODE_STEP) {
            if (my_tot <= 1) rew = 1.0;
            else {
                int w = 0;
                if (design_needs_weight(p.reward_design, my_tot)) w = block_reward_weight<NW>(p, sx, sy, own, norm);
                rew = collision_reward_step(p.reward_design, my_tot, w);
            }
        } else if (mode == MODE_DESIGN) {
            if (my_tot <= 1) rew = 1.0;
            else {   // TestEnv.calculate_reward_design (test_env.py:319-349)
                int k = 1, last = u;
                for (int w = 0; w < NW; ++w)
                    for (unsigned m = own[w]; m; m &= m - 1) {
                        const int t = w * 32 + __ffs(m) - 1;
                        if (t != u && dist2d(x, y, sx[t], sy[t]) < p.C2) { ++k; last = t; }
                    }
                if (k == 1) rew = 1.0;
                else if (k == 2) rew = (dist2d(x, y, sx[last], sy[last]) > p.C2) ? 0.0 : -2.0;
                else rew = -(double)k;
            }
        } else {
            // PRR (test_env.py:384-405): receivers in range = own in-range bits outside the collision set
            int in_range = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const unsigned livew = (w * 32 + 32 <= N) ? 0xffffffffu : ((w * 32 < N) ? ((1u << (N - w * 32)) - 1u) : 0u);
                in_range += __popc(inr[w] & ~own[w] & livew);
            }
            rew = channel_reward(p.reward_design, max(my_tot, 1), (int)recv_s[u], in_range);
        }
        p.rews[vbase + u] = (float)rew;
    }

    // ---- D: mobility -------------------------------------------------------------------------------
    double x_new = act ? mobility_step(p, x, v, u) : 0.0;
    if (act) { if (p.mobility) p.pos_x[vbase + u] = x_new; sxn[u] = x_new; recv_s[u] = 0u; }   // recv_s now sums m_cnt
    __syncthreads();
    if (col && h != 0) { x = sx[u]; y = sy[u]; x_new = sxn[u]; }

    // ---- E: stream the columns (thread u = observer) -----------------------------------------------
    int m_cnt = 0;
    if (vpd && h == 0) { for (int k = 0; k <= B; ++k) hist[k * T + u] = 0u; }
    __syncthreads();
    if (p.piggy) {
        int32_t *seqp = p.tab_seq + tbase + u, *lup = p.tab_lu + tbase + u;
        double *xp = p.tab_x + tbase + u;
        const double W = p.W, inv_binw = p.inv_binw;
        int buf = 0;
        for (int jb0 = 0; jb0 < N; jb0 += H * CB) {
            const int jb = jb0 + h * CB;         // this team's columns of the block
            int s0[CB], lu[CB], sn[CB]; double xo[CB]; unsigned key[CB];
            double *cb = colbuf + ((size_t)buf * H + h) * CB * N;
#pragma unroll
            for (int c = 0; c < CB; ++c) {
                const int j = jb + c;
                s0[c] = 0; lu[c] = 0; xo[c] = 0.0; key[c] = 0u;
                if (col && j < N) {
                    s0[c] = seqp[j * N]; lu[c] = lup[j * N]; xo[c] = xp[j * N];
                    if (j == u) { s0[c] += 1; lu[c] = 0; xo[c] = x; } else lu[c] += 1;     // vehicle.py:58-70
                    key[c] = K[u * ld + j];
                    cb[c * N + u] = xo[c];
                }
            }
            __syncthreads();                  // column buffer complete (double-buffered: one barrier per CB)
#pragma unroll
            for (int c = 0; c < CB; ++c) {
                const int j = jb + c;
                if (col && j < N) {
                    sn[c] = (int)(key[c] >> SB);
                    double xn = xo[c];
                    if (sn[c] != s0[c]) { xn = cb[c * N + (int)(key[c] & srcmask)]; lu[c] = 0; }   // vehicle.py:41-47
                    seqp[j * N] = sn[c]; lup[j * N] = lu[c]; xp[j * N] = xn;
                    if (vpd) {
                        bool in = j != u && lu[c] < p.age_threshold;                      // network.py:547
                        double sv;
                        if (flat0) { sv = __dsub_rn(xn, x_new); in = in && fabs(sv) < W; }
                        else {
                            const double d = dist2d(xn, sn[c] > 0 ? sy[j] : 0.0, x_new, y);
                            in = in && d < W;                                             // network.py:487
                            sv = (__dsub_rn(xn, x_new) > 0.0) ? d : -d;
                        }
                        if (in) {
                            // trunc(t) is NumPy's edge-corrected bin unless t is within 1e-6 of an edge
                            const double t = __dmul_rn(__dadd_rn(sv, W), inv_binw);
                            const double rt = __dadd_rn(__dadd_rn(t, 6755399441055744.0), -6755399441055744.0);
                            const int kb = (fabs(__dsub_rn(t, rt)) < 1e-6) ? vpd_bin(sv, W, inv_binw, B, s_edges)
                                                                           : min(max((int)t, 0), B - 1);
                            atomicAdd(&hist[kb * T + u], 1u);
                            ++m_cnt;
                        }
                    }
                }
            }
            buf ^= 1;
        }
    }
    if (col && m_cnt) atomicAdd(&recv_s[u], (unsigned)m_cnt);
    __syncthreads();                          // every thread is done with the keys: the region becomes staging
    if (act) m_cnt = (int)recv_s[u];

    // ---- F: state rows (TestEnv.obtain_state, test_env.py:527-583) ---------------------------------
    if (want_state) {
        const bool staged = keys_in_smem && (size_t)N * S * 4 <= lay.keys_bytes;
        float *st = reinterpret_cast<float *>(smem_raw + lay.off_keys);
        float *wp = staged ? st + u * S : p.state + (vbase + u) * S;
        if (act) {
            const float *orow = og + (long long)u * R;   // written by this thread's own warp; visible after the barriers
            if (p.add_action) {
                if (p.action_binary) { for (int r = 0; r < R; ++r) *wp++ = (a == r) ? 1.0f : 0.0f; }
                else *wp++ = (float)a;
            }
            if (p.add_channel_obs) { for (int r = 0; r < R; ++r) *wp++ = __ldcg(orow + r); }
            if (p.piggy) {
                const float den = (float)m_cnt, rcp = __frcp_rn(den);
                const bool have = vpd && m_cnt > 0;
                for (int b = 0; b < B; ++b) {
                    const float c = have ? (float)hist[b * T + u] : 0.0f;
                    const float q0 = __fmul_rn(c, rcp);
                    *wp++ = have ? __fmaf_rn(__fmaf_rn(-q0, den, c), rcp, q0) : 0.0f;
                }
            }
            if (p.add_reward) *wp++ = (float)rew;
            if (p.add_index) *wp++ = (float)(u + 1);
            if (p.add_position) { *wp++ = (float)__ddiv_rn(x_new, p.L); *wp++ = (float)__ddiv_rn(y, 2.0); }
            if (p.add_velocity) *wp++ = (float)v;
            if (p.fingerprint) { *wp++ = (float)p.episode; *wp++ = (float)p.epsilon; }
        }
        if (staged) {
            __syncthreads();
            float *sg = p.state + vbase * S;
            const int n = N * S;
            if ((n & 3) == 0) {
                for (int i = tid; i < n / 4; i += TT) reinterpret_cast<float4 *>(sg)[i] = reinterpret_cast<const float4 *>(st)[i];
            } else {
                for (int i = tid; i < n; i += TT) sg[i] = st[i];
            }
        }
    }

    // ---- per-env metric accumulators (fixed-order block reduction) ----------------------------------
    {
        double rs = act ? rew : 0.0; int nr = n_recv, np = n_pairs, nb = bad;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            rs += __shfl_xor_sync(0xffffffffu, rs, o);
            nr += __shfl_xor_sync(0xffffffffu, nr, o);
            np += __shfl_xor_sync(0xffffffffu, np, o);
            nb += __shfl_xor_sync(0xffffffffu, nb, o);
        }
        if (lane == 0 && h == 0) { s_red[warp * 4 + 0] = rs; s_red[warp * 4 + 1] = nr; s_red[warp * 4 + 2] = np; s_red[warp * 4 + 3] = nb; }
        __syncthreads();
        if (tid == 0) {
            double trs = 0.0, tnr = 0.0, tnp = 0.0, tnb = 0.0;
            for (int i = 0; i < NW; ++i) { trs += s_red[i * 4]; tnr += s_red[i * 4 + 1]; tnp += s_red[i * 4 + 2]; tnb += s_red[i * 4 + 3]; }
            atomicAdd(p.acc_reward + e, trs);
            unsigned long long *c = reinterpret_cast<unsigned long long *>(p.acc_count + e * ACC_COUNTS);
            atomicAdd(c + 0, (unsigned long long)tnr); atomicAdd(c + 1, (unsigned long long)tnp);
            atomicAdd(c + 2, (unsigned long long)tnb); atomicAdd(c + 3, 1ull);
        }
    }
}

int block_threads(int N) { return ((N + 31) / 32) * 32; }
int helpers_for(int N)
{
    switch (block_threads(N) / 32) {
    case 1: return Helpers<1>::v; case 2: return Helpers<2>::v; case 3: return Helpers<3>::v; case 4: return Helpers<4>::v;
    case 5: return Helpers<5>::v; case 6: return Helpers<6>::v; case 7: return Helpers<7>::v; default: return Helpers<8>::v;
    }
}

template <int NW>
cudaError_t prepare_nw(const Params &p, size_t smem)
{
    if (smem <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(step_block_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

template <int NW>
cudaError_t launch_nw(const Params &p, size_t smem, int SB, int fit, cudaStream_t stream)
{
    step_block_kernel<NW><<<(unsigned)p.E, NW * 32 * Helpers<NW>::v, smem, stream>>>(p, SB, fit);
    return cudaGetLastError();
}

}  // namespace

int key_src_bits(int N)
{
    int b = 1;
    while ((1 << b) < N) ++b;
    return b;
}

bool step_block_keys_fit_smem(const Params &p)
{
    const BlockSmem lay(p.N, p.R, p.B, block_threads(p.N), helpers_for(p.N), p.vpd_enabled != 0, true);
    return lay.bytes <= SMEM_BUDGET;
}

size_t step_block_smem_bytes(const Params &p, bool keys_in_smem)
{
    const BlockSmem lay(p.N, p.R, p.B, block_threads(p.N), helpers_for(p.N), p.vpd_enabled != 0, keys_in_smem);
    return lay.bytes;
}

size_t step_block_scratch_bytes(long long E, int N)
{
    return (size_t)E * N * (N + 1) * sizeof(unsigned);
}

cudaError_t prepare_step_block(const Params &p)
{
    const size_t smem = step_block_smem_bytes(p, step_block_keys_fit_smem(p));
    switch (block_threads(p.N) / 32) {
    case 1: return prepare_nw<1>(p, smem);
    case 2: return prepare_nw<2>(p, smem);
    case 3: return prepare_nw<3>(p, smem);
    case 4: return prepare_nw<4>(p, smem);
    case 5: return prepare_nw<5>(p, smem);
    case 6: return prepare_nw<6>(p, smem);
    case 7: return prepare_nw<7>(p, smem);
    default: return prepare_nw<8>(p, smem);
    }
}

cudaError_t launch_step_block(const Params &p, cudaStream_t stream)
{
    const bool fit = step_block_keys_fit_smem(p);
    const size_t smem = step_block_smem_bytes(p, fit);
    const int SB = key_src_bits(p.N);
    switch (block_threads(p.N) / 32) {
    case 1: return launch_nw<1>(p, smem, SB, fit, stream);
    case 2: return launch_nw<2>(p, smem, SB, fit, stream);
    case 3: return launch_nw<3>(p, smem, SB, fit, stream);
    case 4: return launch_nw<4>(p, smem, SB, fit, stream);
    case 5: return launch_nw<5>(p, smem, SB, fit, stream);
    case 6: return launch_nw<6>(p, smem, SB, fit, stream);
    case 7: return launch_nw<7>(p, smem, SB, fit, stream);
    default: return launch_nw<8>(p, smem, SB, fit, stream);
    }
}

}  // namespace diral
