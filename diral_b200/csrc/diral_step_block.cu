// diral_step_block.cu -- fused time-slot kernel for 32 < N <= 256 vehicles (one CTA per environment,
// persistent CTAs that walk the environments of the launch).
//
// Same slot semantics as diral_step_group.cu, re-mapped for rows that no longer fit one warp.  The table keys
// live in shared memory as K[observer i][subject j] (rows 16 B aligned), per environment in one of two forms:
// packed 16-bit keys  fresh << SB | origin-row  (fresh = seq - (tick - FMAX), two columns per VIMNMX.U16x2) while
// every sequence number is 0 or within FMAX slots of the newest one, else 32-bit keys  seq << SB | origin-row
// (in shared memory up to 128 vehicles, in an L2-resident scratch slice per CTA beyond).  See KM below.
//
//   A   inputs -> shared memory; packed keys from the seq columns (one warp per PAIR of subject columns,
//       coalesced, one 32-bit store carries both keys) and a block vote on whether they are exact, 32-bit keys
//       otherwise; in-range bit mask of every vehicle (Network.check_communicaiton_range, network.py:595-607),
//       thread (u, w) forms word w of vehicle u against warp-uniform candidate positions
//   B   DECISIONS, no table access.  Resources are taken RC at a time.  Transmitter masks per resource (the
//       per-resource collision histogram, envs/test_env.py:149-157) by shared-memory atomicOr, copied to a
//       per-vehicle "who shares my resource" mask.  Channel observations start as coalesced rows of the
//       no-reception value.  Then thread (u, w, part) walks the in-range vehicles t of (a part of) word w: every
//       t transmits on exactly one resource, so t is u's nearest in-range transmitter there
//       (Network.find_closest_tx, network.py:378-398) unless another in-range vehicle shares t's resource -- the
//       common case costs a few mask words, no search.  Winners store the distance and append (u, t) to the
//       list of pass a[t]
//   C   MERGES, resource passes in ascending order (they are a true dependency: Vehicle.periodic_update
//       aliases the transmitted table, vehicle.py:61).  One sub-warp per reception applies
//       Vehicle.received_update (vehicle.py:35-47) to the whole row with 16 B accesses:
//           K[rx][:] = max(K[rx][:], K[tx][:])
//       conflict-free, exactly the merges that happen; one barrier per non-empty pass (a pass never modifies
//       a transmitter's row and a receiver appears once per pass)
//   D   rewards (lane-local: a vehicle's collision set is txm[a]) and mobility
//   E   one warp per subject column: xpos of a merged entry is the old position held by the origin row (an
//       entry's position is a pure function of (subject, seq)), gathered through a per-warp column buffer
//       (N <= 128) or from the column itself in global memory before any lane writes it back; ages,
//       write-back, and the positional distribution (network.py:473-513) binned with shared-memory reductions
//       -- no CTA barrier inside
//   F   state rows (TestEnv.obtain_state, test_env.py:527-583): one warp per row, coalesced stores
#include "diral_dev.cuh"
#include "diral_launch.h"

#include <algorithm>

namespace diral {

namespace {

// compile-time geometry of the NW-warp instantiation (N <= 32 NW vehicles)
__host__ __device__ constexpr int geo_lpr(int nw) { return nw <= 1 ? 4 : nw <= 2 ? 8 : nw <= 4 ? 16 : 32; }  // lanes per reception in a row merge (8 key words each)
__host__ __device__ constexpr int geo_ld(int nw) { return 8 * geo_lpr(nw) + 4; }                              // key row stride (words)
__host__ __device__ constexpr int geo_ld16(int nw) { return 8 * geo_lpr(nw) + 8; }                            // packed (16-bit) key row stride, in keys
#ifndef DIRAL_NW2_THREADS
#define DIRAL_NW2_THREADS 256
#endif
__host__ __device__ constexpr int geo_threads(int nw) { return nw <= 1 ? 128 : nw <= 2 ? DIRAL_NW2_THREADS : nw <= 4 ? 512 : 1024; }
__host__ __device__ constexpr int geo_nwp(int nw) { return nw | 1; }                                         // odd stride of the bit-mask rows
__host__ __device__ constexpr int geo_rc(int nw) { return nw <= 4 ? ((256 / nw) & ~31) : 64; }              // resources per decision chunk (reception lists <= 16 KB, 32 KB beyond 128 vehicles)

// Shared-memory carve-up.  Everything the hot loops touch sits at an offset that depends on the
// instantiation only (a compile-time constant inside the kernel: addresses fold into the instructions
// instead of being recomputed under register pressure); the arrays sized by the runtime bin count follow.
__host__ __device__ constexpr size_t align16c(size_t x) { return (x + 15) & ~(size_t)15; }
struct BlockSmem {
    size_t off_sx, off_sy, off_sxn, off_rewd, off_sa, off_aux, off_rew, off_recv, off_txm, off_cnt, off_inr, off_own, off_red, off_union,
        off_keys, off_edges, off_hist, bytes;
    size_t keys_bytes, list_bytes;
    // keys_mode: 2 = 32-bit keys in shared memory (the packed layout shares the region), 1 = packed keys only
    // (32-bit fallback in the scratch slice), 0 = no keys in shared memory at all
    __host__ __device__ constexpr BlockSmem(int B, int nw, bool vpd_state, int keys_mode)
        : off_sx(0), off_sy(0), off_sxn(0), off_rewd(0), off_sa(0), off_aux(0), off_rew(0), off_recv(0), off_txm(0), off_cnt(0), off_inr(0), off_own(0),
          off_red(0), off_union(0), off_keys(0), off_edges(0), off_hist(0), bytes(0), keys_bytes(0), list_bytes(0)
    {
        const int T = nw * 32, NWP = geo_nwp(nw), RC = geo_rc(nw), NWARPS = geo_threads(nw) / 32;
        size_t o = 0;
        off_sx = o;    o += align16c(8 * (size_t)T);
        off_sy = o;    o += align16c(8 * (size_t)T);
        off_sxn = o;   o += align16c(8 * (size_t)T);          // post-mobility x
        off_rewd = o;  o += align16c(8 * (size_t)T);          // rewards (float64, for the episode sums)
        off_sa = o;    o += align16c(4 * (size_t)T);          // actions
        off_aux = o;   o += align16c(4 * (size_t)T);          // collision-set size | receivers in range << 16
        off_rew = o;   o += align16c(4 * (size_t)T);          // rewards (float)
        off_recv = o;  o += align16c(4 * (size_t)T);          // receptions per transmitter, later VPD sample counts
        off_txm = o;   o += align16c(4 * (size_t)RC * NWP);
        off_cnt = o;   o += align16c(4 * (size_t)RC);
        off_inr = o;   o += align16c(4 * (size_t)T * NWP);
        off_own = o;   o += align16c(4 * (size_t)T * NWP);    // txm[a[t]] per vehicle t
        off_red = o;   o += align16c(8 * 4 * 16 + 16);
        // phase-disjoint: reception lists of one resource chunk (B, C) and -- up to 128 vehicles -- the per-warp
        // column buffers of phase E (beyond that the merged positions are gathered from global memory instead,
        // which leaves room for 1024-thread CTAs)
        list_bytes = align16c(2 * (size_t)RC * T);
        const size_t colx = nw <= 4 ? 8 * (size_t)NWARPS * T : 0;
        off_union = o; o += align16c(list_bytes > colx ? list_bytes : colx);
        // 32-bit keys when they fit (the packed 16-bit layout then uses the first half), else room for the packed
        // layout only and the 32-bit fallback lives in the L2 scratch slice; sized by the padded vehicle count
        keys_bytes = keys_mode == 2 ? align16c(4 * (size_t)T * geo_ld(nw)) : keys_mode == 1 ? align16c(2 * (size_t)T * geo_ld16(nw)) : 0;
        off_keys = o;  o += keys_bytes;
        off_hist = o;  o += vpd_state ? align16c(4 * (size_t)B * T) : 0;
        off_edges = o; o += align16c(8 * (size_t)(B + 1));        // (only the near-edge path of the binning reads them)
        bytes = o;
    }
};

constexpr size_t SMEM_BUDGET = 226 * 1024;

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

// one lane's share of a key row: two 16 B pieces, LPR * 16 B apart (sub = lane within its reception group)
template <int LPR, bool KS>
__device__ __forceinline__ void row_load(const unsigned *row, int sub, uint4 &lo, uint4 &hi)
{
    const uint4 *q = reinterpret_cast<const uint4 *>(row) + sub;
    if (KS) { lo = q[0]; hi = q[LPR]; } else { lo = __ldcg(q); hi = __ldcg(q + LPR); }
}
template <int LPR, bool KS>
__device__ __forceinline__ void row_store(unsigned *row, int sub, const uint4 &lo, const uint4 &hi)
{
    uint4 *q = reinterpret_cast<uint4 *>(row) + sub;
    if (KS) { q[0] = lo; q[LPR] = hi; } else { __stcg(q, lo); __stcg(q + LPR, hi); }
}
__device__ __forceinline__ uint4 max4(const uint4 &a, const uint4 &b)
{
    return make_uint4(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z), max(a.w, b.w));
}

template <int NW, int KM>
__global__ void __launch_bounds__(geo_threads(NW), 1024 / geo_threads(NW))
step_block_kernel(const Params p, const int SB)
{
    constexpr bool KS = KM == 2;                  // 32-bit keys in shared memory (else in the scratch slice)
    // packed 16-bit keys in shared memory: whenever there is no room for the 32-bit ones, and from 65 vehicles on
    // (below that a pass has fewer receptions than one round of 32-bit sub-warps takes: nothing to halve)
    constexpr bool HAS16 = KM == 1 || (KM == 2 && NW >= 3);
    constexpr int T = NW * 32;                    // padded vehicle count
    constexpr int LPR = geo_lpr(NW), LD = geo_ld(NW), TT = geo_threads(NW), NWARPS = TT / 32;
    constexpr int G = 32 / LPR;                   // receptions per warp and merge round
    constexpr int LD16 = geo_ld16(NW), LPR16 = LPR / 2, G16 = 32 / LPR16;   // the same for packed 16-bit keys
    constexpr int NWP = geo_nwp(NW), RC = geo_rc(NW);
    const int N = p.N, R = p.R, B = p.B, S = p.S;
    int tid;      // read once through an opaque asm: ptxas otherwise re-reads SR_TID.X inside the hot loops
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int lane = tid & 31, warp = tid >> 5;
    const bool act = tid < N;                     // thread u = vehicle u for the per-vehicle work
    const bool want_state = p.build_state != 0;
    const bool vpd = want_state && p.vpd_enabled;
    const int mode = p.mode;
    const bool merge_mode = p.piggy && (mode != MODE_STEP || p.state_type == 1 || p.state_type == 2);
    const double Cr = p.C, sentinel = p.sentinel;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr BlockSmem fix(0, NW, false, KM);    // offsets up to the histogram do not depend on the bin count
    const BlockSmem lay(B, NW, p.vpd_enabled != 0, KM);
    double *sx = reinterpret_cast<double *>(smem_raw + fix.off_sx);
    double *sy = reinterpret_cast<double *>(smem_raw + fix.off_sy);
    double *sxn = reinterpret_cast<double *>(smem_raw + fix.off_sxn);
    double *s_edges = reinterpret_cast<double *>(smem_raw + lay.off_edges);
    double *s_rewd = reinterpret_cast<double *>(smem_raw + fix.off_rewd);
    int *sa = reinterpret_cast<int *>(smem_raw + fix.off_sa);
    int *s_aux = reinterpret_cast<int *>(smem_raw + fix.off_aux);
    float *s_rew = reinterpret_cast<float *>(smem_raw + fix.off_rew);
    unsigned *recv_s = reinterpret_cast<unsigned *>(smem_raw + fix.off_recv);
    unsigned *txm_s = reinterpret_cast<unsigned *>(smem_raw + fix.off_txm);            // [RC][NWP]
    int *cnt_s = reinterpret_cast<int *>(smem_raw + fix.off_cnt);                      // [RC] receptions of a pass
    unsigned *inr_s = reinterpret_cast<unsigned *>(smem_raw + fix.off_inr);            // [N][NWP]
    unsigned *own_s = reinterpret_cast<unsigned *>(smem_raw + fix.off_own);            // [N][NWP]
    double *s_red = reinterpret_cast<double *>(smem_raw + fix.off_red);                // [16][4]
    unsigned *s_tot = reinterpret_cast<unsigned *>(smem_raw + fix.off_red + 8 * 4 * 16); // received, pairs, bad
    unsigned *hist = reinterpret_cast<unsigned *>(smem_raw + fix.off_hist);            // [B][T]
    unsigned short *list = reinterpret_cast<unsigned short *>(smem_raw + fix.off_union);   // [RC][T]  rx << 8 | tx
    constexpr bool COLX = NW <= 4;                // merged positions through a per-warp shared-memory column buffer
    double *colx = reinterpret_cast<double *>(smem_raw + fix.off_union) + warp * T;    // phase E (COLX only)
    unsigned *K;
    if constexpr (KS) K = reinterpret_cast<unsigned *>(smem_raw + fix.off_keys);
    else K = p.scratch + (size_t)blockIdx.x * N * LD;
    // Packed keys (always in shared memory): fresh << SB | origin with fresh = seq - (tick - FMAX), 0 for "never
    // heard".  Their order equals the order of the 32-bit keys whenever every sequence number of the environment is 0
    // or within FMAX slots of the newest one; environments holding older entries take the 32-bit path.
    unsigned short *K16 = reinterpret_cast<unsigned short *>(smem_raw + fix.off_keys);
    const unsigned srcmask = (1u << SB) - 1u;
    const int FMAX = (1 << (16 - SB)) - 1, kbase = p.tick - FMAX;

    for (int i = tid; i <= B; i += TT) s_edges[i] = p.edges[i];

    // HBM -> L2 ahead of use: an environment's ages and positions while its decisions and merges run
    // (they are first touched in phase E), the NEXT environment's sequence numbers during phase E
    auto prefetch_rest = [&](long long ee) {
        const char *b1 = reinterpret_cast<const char *>(p.tab_lu + ee * N * N);
        const char *b2 = reinterpret_cast<const char *>(p.tab_x + ee * N * N);
        const int bytes4 = N * N * 4;
        for (int o = tid * 128; o < bytes4; o += TT * 128) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(b1 + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + bytes4 + o));
        }
    };
    auto prefetch_seq = [&](long long ee) {
        const char *b0 = reinterpret_cast<const char *>(p.tab_seq + ee * N * N);
        for (int o = tid * 128; o < N * N * 4; o += TT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(b0 + o));
    };

    // prefetching pays only while the tables of all resident environments fit the L2 next to everything else
    const bool pf_ok = p.piggy && (long long)gridDim.x * N * N * 16 <= (80ll << 20);
    for (long long e = blockIdx.x; e < p.E; e += gridDim.x) {
        const long long vbase = e * N, tbase = e * (long long)N * N;

        // ---- A: inputs -----------------------------------------------------------------------------
        // (per-vehicle values live in shared memory, not in registers, across the long phases)
        if (tid < 3) s_tot[tid] = 0u;
        bool y_same = true;
        if (act) {
            int a;
            if (p.gen_actions) a = philox_action(p.seed, tid, p.env0 + e, p.timestep, R);
            else a = p.actions[vbase + tid];
            const double y = p.pos_y[vbase + tid];
            sx[tid] = p.pos_x[vbase + tid]; sy[tid] = y;
            y_same = y == p.pos_y[vbase];
            if (a < 0 || a >= R) { a = min(max(a, 0), R - 1); s_aux[tid] = -1; } else s_aux[tid] = 0;
            if (p.gen_actions && p.actions_out) p.actions_out[vbase + tid] = a;
            sa[tid] = a; recv_s[tid] = 0u; s_rewd[tid] = 0.0;
        }
        if (pf_ok) prefetch_rest(e);
        if (vpd) for (int i = tid; i < B * T; i += TT) hist[i] = 0u;
        // every vehicle on the same lane of the highway?  (dy == 0 for every pair => dist == |dx| exactly)
        const bool flat = __syncthreads_and(y_same) != 0;              // also publishes sx, sy, sa
        const bool flat0 = flat && sy[0] == 0.0;
        if (act && s_aux[tid] < 0) { atomicAdd(&s_tot[2], 1u); s_aux[tid] = 0; }   // out-of-range actions are counted

        // keys: K[i][j] = (seq[i][j] (+1 on the diagonal: the tick, vehicle.py:58)) << SB | i, packed first
        bool wide_needed = !HAS16;
        if (HAS16 && p.piggy) {       // two subject columns per warp pass: one 32-bit store carries both packed keys
            const int32_t *seqg = p.tab_seq + tbase;
            // no sequence number exceeds the tick, so the packed form is exact iff every non-zero one is above
            // kbase: one unsigned minimum of seq - 1 per entry (0 wraps to the maximum)
            unsigned oldest = 0xffffffffu;
            for (int j = 2 * warp; j < N; j += 2 * NWARPS) {
                const bool two = j + 1 < N;
                const int32_t *sc = seqg + (long long)j * N + lane;
#pragma unroll
                for (int q = 0; q < NW; ++q) {
                    const int i = q * 32 + lane;
                    if (i < N) {
                        int s0 = sc[q * 32], s1 = two ? sc[N + q * 32] : 0;
                        if (i == j) s0 += 1;
                        if (i == j + 1) s1 += 1;
                        oldest = min(oldest, min((unsigned)(s0 - 1), (unsigned)(s1 - 1)));
                        const unsigned f0 = s0 ? (unsigned)(s0 - kbase) : 0u, f1 = s1 ? (unsigned)(s1 - kbase) : 0u;   // (garbage when
                        const unsigned k0 = ((f0 << SB) | (unsigned)i) & 0xffffu, k1 = ((f1 << SB) | (unsigned)i) & 0xffffu;  // out of range)
                        if (two) *reinterpret_cast<unsigned *>(K16 + i * LD16 + j) = k0 | (k1 << 16);
                        else K16[i * LD16 + j] = (unsigned short)k0;
                    }
                }
            }
            wide_needed = kbase > 0 && oldest < (unsigned)kbase;
        }
        // who is within communication range of whom (self included; it never counts as a candidate):
        // thread (u, w) forms word w of vehicle u, the candidate positions are warp-uniform broadcasts
        for (int it = tid; it < T * NW; it += TT) {
            const int w = it / T, u = it - w * T;
            if (u < N) {
                const double xu = sx[u], yu = sy[u];
                const int nb = min(32, N - w * 32);
                unsigned m = 0u;
                if (flat) {
#pragma unroll 8
                    for (int b = 0; b < nb; ++b) m |= (fabs(__dsub_rn(xu, sx[w * 32 + b])) < Cr ? 1u : 0u) << b;
                } else {
                    for (int b = 0; b < nb; ++b) m |= (dist2d(sx[w * 32 + b], sy[w * 32 + b], xu, yu) < Cr ? 1u : 0u) << b;
                }
                inr_s[u * NWP + w] = m;
            }
        }

        const bool narrow = __syncthreads_or(wide_needed) == 0 && HAS16;   // uniform; also publishes K16 and inr_s
        if (p.piggy && !narrow) {     // (rare) entries older than the packed range: 32-bit keys
            const int32_t *seqg = p.tab_seq + tbase;
#pragma unroll 2
            for (int j = warp; j < N; j += NWARPS) {
#pragma unroll
                for (int q = 0; q < NW; ++q) {
                    const int i = q * 32 + lane;
                    if (i < N) {
                        int s = seqg[j * N + i];
                        if (i == j) s += 1;
                        const unsigned key = ((unsigned)s << SB) | (unsigned)i;
                        if (KS) K[i * LD + j] = key; else __stcg(K + i * LD + j, key);
                    }
                }
            }
        }

        // ---- B + C: resources, RC at a time --------------------------------------------------------
        float *og = p.obs + vbase * R;
        for (int r0 = 0; r0 < R; r0 += RC) {
            const int rend = min(R, r0 + RC), nres = rend - r0;
            for (int i = tid; i < nres * NWP; i += TT) txm_s[i] = 0u;
            for (int i = tid; i < nres; i += TT) cnt_s[i] = 0;
            __syncthreads();
            const int a = act ? sa[tid] : -1;
            const bool mine = act && a >= r0 && a < rend;
            if (mine) atomicOr(&txm_s[(a - r0) * NWP + warp], 1u << lane);             // test_env.py:149-157
            __syncthreads();

            // rewards that need nothing but the collision set (test_env.py:159-199 / :294-302)
            if (mine) {
                unsigned own[NW], inr[NW];
                int my_tot = 0, in_range = 0; double rew = 0.0;
                const double x = sx[tid], y = sy[tid];
#pragma unroll
                for (int w = 0; w < NW; ++w) { own[w] = txm_s[(a - r0) * NWP + w]; inr[w] = inr_s[tid * NWP + w]; my_tot += __popc(own[w]); }
                if (mode == MODE_STEP) {
                    if (my_tot <= 1) rew = 1.0;
                    else {
                        int wgt = 0;
                        if (design_needs_weight(p.reward_design, my_tot)) {
                            double norm = 0.0;
                            if (p.toy) {   // first-min-x / first-max-x vehicle (network.py:225-246)
                                double xmin = p.L + 1.0, xmax = -p.L - 1.0; int imin = 0, imax = 0;
                                for (int t = 0; t < N; ++t) {
                                    if (sx[t] < xmin) { xmin = sx[t]; imin = t; }
                                    if (sx[t] > xmax) { xmax = sx[t]; imax = t; }
                                }
                                norm = dist2d(sx[imin], sy[imin], sx[imax], sy[imax]);
                            }
                            wgt = block_reward_weight<NW>(p, sx, sy, own, norm);
                        }
                        rew = collision_reward_step(p.reward_design, my_tot, wgt);
                    }
                } else if (mode == MODE_DESIGN) {
                    if (my_tot <= 1) rew = 1.0;
                    else {   // TestEnv.calculate_reward_design (test_env.py:319-349)
                        int k = 1, last = tid;
                        for (int w = 0; w < NW; ++w)
                            for (unsigned m = own[w]; m; m &= m - 1) {
                                const int t = w * 32 + __ffs(m) - 1;
                                if (t != tid && dist2d(x, y, sx[t], sy[t]) < p.C2) { ++k; last = t; }
                            }
                        if (k == 1) rew = 1.0;
                        else if (k == 2) rew = (dist2d(x, y, sx[last], sy[last]) > p.C2) ? 0.0 : -2.0;
                        else rew = -(double)k;
                    }
                } else {
                    // PRR (test_env.py:384-405): receivers in range = own in-range bits outside the collision set
#pragma unroll
                    for (int w = 0; w < NW; ++w) in_range += __popc(inr[w] & ~own[w]);
                }
                s_rewd[tid] = rew; s_aux[tid] = my_tot | (in_range << 16);
            }

            // per-vehicle collision set; channel observations before any reception (coalesced rows)
            if (mine) {
#pragma unroll
                for (int w = 0; w < NW; ++w) own_s[tid * NWP + w] = txm_s[(a - r0) * NWP + w];
            }
            {
                const float basev = (mode != MODE_STEP || p.state_type == 1) ? 1.0f : (p.state_type == 2 ? (float)sentinel : 0.0f);
                const int rcw = (nres + 31) >> 5;
                for (int g = 0; g < rcw; ++g) {
                    const int rl = g * 32 + lane, r = r0 + rl;
                    unsigned any = 0u;
                    if (rl < nres) {
#pragma unroll
                        for (int w = 0; w < NW; ++w) any |= txm_s[rl * NWP + w];
                    }
                    if (rl < nres)
                        for (int u = warp; u < N; u += NWARPS)
                            og[(long long)u * R + r] = (any != 0u && sa[u] != r) ? basev : 0.0f;
                }
            }
            __syncthreads();

            // decisions: thread (u, w, part) walks the in-range vehicles t of one part of word w (a CTA has more
            // threads than (vehicle, word) pairs below 128 vehicles: SUBS parts per word keep all of them busy)
            constexpr int SUBS = TT > T * NW ? TT / (T * NW) : 1;
            int n_recv = 0, n_pairs = 0;
            for (int it = tid; it < T * NW * SUBS; it += TT) {
                const int part = it / (T * NW), wu = it - part * (T * NW);
                const int w = wu / T, u = wu - w * T;
                if (u >= N) continue;
                const unsigned part_mask = SUBS == 1 ? 0xffffffffu : (((1u << (32 / SUBS)) - 1u) << (part * (32 / SUBS)));
                const int au = sa[u];
                const double xu = sx[u], yu = sy[u];
                unsigned inr[NW];
#pragma unroll
                for (int w2 = 0; w2 < NW; ++w2) inr[w2] = inr_s[u * NWP + w2];
                unsigned mine_w = 0u;
#pragma unroll
                for (int w2 = 0; w2 < NW; ++w2) if (w2 == w) mine_w = inr[w2];
                for (unsigned c = mine_w & part_mask; c; c &= c - 1) {
                    const int t = w * 32 + __ffs(c) - 1;
                    const int at = sa[t];
                    if (at == au || at < r0 || at >= rend) continue;          // u transmits there itself (half duplex)
                    // in-range vehicles that share t's resource: candidates of find_closest_tx
                    int ncand = 0;
#pragma unroll
                    for (int w2 = 0; w2 < NW; ++w2) ncand += __popc(inr[w2] & own_s[t * NWP + w2]);
                    const double d = flat ? fabs(__dsub_rn(xu, sx[t])) : dist2d(sx[t], sy[t], xu, yu);
                    bool win = d < sentinel;          // best starts at the sentinel (network.py:380)
                    if (ncand > 1) {        // ascending ids, strict '<': the first minimum wins (network.py:384-391)
                        for (int w2 = 0; w2 < NW; ++w2)
                            for (unsigned c2 = inr_s[u * NWP + w2] & own_s[t * NWP + w2]; c2; c2 &= c2 - 1) {
                                const int t2 = w2 * 32 + __ffs(c2) - 1;
                                const double d2 = flat ? fabs(__dsub_rn(xu, sx[t2])) : dist2d(sx[t2], sy[t2], xu, yu);
                                if (d2 < d || (d2 == d && t2 < t)) win = false;
                            }
                    }
                    n_pairs += 1;
                    if (win) {
                        ++n_recv;
                        const int rl = at - r0;
                        if (mode == MODE_CH) {
                            atomicAdd(&recv_s[t], 1u);                                     // test_env.py:396-397
                            if (p.track_lat) p.lat[tbase + (long long)t * N + u] = (int32_t)p.timestep;   // test_env.py:436
                        }
                        if (mode == MODE_STEP && p.state_type == 2) og[(long long)u * R + at] = (float)d;
                        if (merge_mode) {
                            const int slot = atomicAdd(&cnt_s[rl], 1);
                            list[rl * T + slot] = (unsigned short)((u << 8) | t);
                        }
                    }
                }
                if (p.track_lat) {                                                       // network.py:394
                    const unsigned live = (w * 32 + 32 <= N) ? 0xffffffffu : ((1u << (N - w * 32)) - 1u);
                    for (unsigned c = ~mine_w & live & part_mask; c; c &= c - 1) {
                        const int t = w * 32 + __ffs(c) - 1;
                        const int at = sa[t];
                        if (at != au && at >= r0 && at < rend) p.lat[tbase + (long long)t * N + u] = -1;
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                n_recv += __shfl_xor_sync(0xffffffffu, n_recv, o);
                n_pairs += __shfl_xor_sync(0xffffffffu, n_pairs, o);
            }
            if (lane == 0) {
                if (n_recv) atomicAdd(&s_tot[0], (unsigned)n_recv);
                if (n_pairs) atomicAdd(&s_tot[1], (unsigned)n_pairs);
            }
            __syncthreads();

            // merges, pass by pass: G receptions per warp and round, LPR lanes per row (packed keys: twice as many
            // receptions per round, two columns per VIMNMX.U16x2)
            if (merge_mode && narrow) {
                const int grp = lane / LPR16, sub = lane - grp * LPR16;
                for (int rl = 0; rl < nres; ++rl) {
                    const int np = cnt_s[rl];
                    if (np == 0) continue;
                    const unsigned short *pl = list + rl * T;
                    for (int k0 = warp * G16; k0 < np; k0 += NWARPS * G16) {   // receptions of one pass are independent
                        const int k = k0 + grp;
                        if (k < np) {
                            const unsigned en = pl[k];
                            uint4 *ra = reinterpret_cast<uint4 *>(K16 + (en >> 8) * LD16) + sub;
                            const uint4 *rb = reinterpret_cast<const uint4 *>(K16 + (en & 255u) * LD16) + sub;
                            const uint4 alo = ra[0], ahi = ra[LPR16], blo = rb[0], bhi = rb[LPR16];
                            ra[0] = make_uint4(__vmaxu2(alo.x, blo.x), __vmaxu2(alo.y, blo.y), __vmaxu2(alo.z, blo.z), __vmaxu2(alo.w, blo.w));
                            ra[LPR16] = make_uint4(__vmaxu2(ahi.x, bhi.x), __vmaxu2(ahi.y, bhi.y), __vmaxu2(ahi.z, bhi.z), __vmaxu2(ahi.w, bhi.w));
                        }
                    }
                    __syncthreads();
                }
            } else if (merge_mode) {
                const int grp = lane / LPR, sub = lane - grp * LPR;
                for (int rl = 0; rl < nres; ++rl) {
                    const int np = cnt_s[rl];
                    if (np == 0) continue;
                    const unsigned short *pl = list + rl * T;
                    for (int k0 = warp * G; k0 < np; k0 += NWARPS * G) {   // receptions of one pass are independent
                        const int k = k0 + grp;
                        if (k < np) {
                            const unsigned en = pl[k];
                            unsigned *ra = K + (en >> 8) * LD, *rb = K + (en & 255u) * LD;
                            uint4 alo, ahi, blo, bhi;
                            row_load<LPR, KS>(ra, sub, alo, ahi); row_load<LPR, KS>(rb, sub, blo, bhi);
                            row_store<LPR, KS>(ra, sub, max4(alo, blo), max4(ahi, bhi));
                        }
                    }
                    __syncthreads();
                }
            }
        }
        __syncthreads();                          // recv_s complete; txm/cnt/list free

        // ---- D: rewards out, mobility ----------------------------------------------------------------
        if (act) {
            double rew = s_rewd[tid];
            if (mode == MODE_CH) {
                const int aux = s_aux[tid];
                rew = channel_reward(p.reward_design, max(aux & 0xffff, 1), (int)recv_s[tid], aux >> 16);
                s_rewd[tid] = rew;
            }
            p.rews[vbase + tid] = (float)rew;
            if (p.vpd_counts) *reinterpret_cast<float *>(p.vpd_counts + (vbase + tid + 1) * p.rec_stride - 4) = (float)rew;
            s_rew[tid] = (float)rew;
            const double x_new = mobility_step(p, sx[tid], p.vel[vbase + tid], tid);
            if (p.mobility) p.pos_x[vbase + tid] = x_new;
            sxn[tid] = x_new;
        }
        if (pf_ok && e + gridDim.x < p.E) prefetch_seq(e + gridDim.x);
        __syncthreads();

        // ---- E: one warp per subject column ------------------------------------------------------------
        if (p.piggy) {
            int32_t *seqg = p.tab_seq + tbase, *lug = p.tab_lu + tbase;
            double *xg = p.tab_x + tbase;
            const double W = p.W, inv_binw = p.inv_binw;
            const int age_thr = p.age_threshold;
            for (int j = warp; j < N; j += NWARPS) {
                int s0[NW], lu[NW]; double xo[NW];
                const double xj = sx[j], yj = sy[j];
                // this lane's entries of column j: one pointer per array, the 32-row steps are immediates
                const long long cb = (long long)j * N + lane;
                int32_t *sc = seqg + cb, *lc = lug + cb; double *xc = xg + cb;
                const unsigned short *kc16 = K16 + lane * LD16 + j;
#pragma unroll
                for (int q = 0; q < NW; ++q) {
                    const int i = q * 32 + lane;
                    s0[q] = 0; lu[q] = 0; xo[q] = 0.0;
                    if (i < N) { s0[q] = __ldcs(sc + q * 32); lu[q] = __ldcs(lc + q * 32); xo[q] = __ldcs(xc + q * 32); }
                }
                // tick, key decode, and the position of every merged entry: it is the OLD position held by the
                // origin row (a pure function of (subject, seq)), read from the per-warp column buffer or (N > 128)
                // straight from this column in global memory -- the warp loaded it a moment ago and nobody has
                // written it yet (the __syncwarp below)
                const double *xcol = xg + (long long)j * N;
                if constexpr (COLX) {
#pragma unroll
                    for (int q = 0; q < NW; ++q) colx[q * 32 + lane] = (q * 32 + lane == j) ? xj : xo[q];
                    __syncwarp();
                }
#pragma unroll
                for (int q = 0; q < NW; ++q) {
                    const int i = q * 32 + lane;
                    if (i == j) { s0[q] += 1; lu[q] = 0; xo[q] = xj; } else lu[q] += 1;       // vehicle.py:58-70
                    if (i < N) {
                        int sn; unsigned origin;
                        if (narrow) {
                            const unsigned hk = kc16[q * 32 * LD16], f = hk >> SB;
                            sn = f ? (int)f + kbase : 0; origin = hk & srcmask;
                        } else {
                            const unsigned key = KS ? K[i * LD + j] : __ldcg(K + i * LD + j);
                            sn = (int)(key >> SB); origin = key & srcmask;
                        }
                        if (sn != s0[q]) {                                                     // vehicle.py:41-47
                            xo[q] = COLX ? colx[origin] : (((int)origin == j) ? xj : xcol[origin]);
                            lu[q] = 0; s0[q] = sn;
                        }
                    }
                }
                __syncwarp();
#pragma unroll
                for (int q = 0; q < NW; ++q) {
                    const int i = q * 32 + lane;
                    if (i < N) {
                        const int sn = s0[q];
                        const double xn = xo[q];
                        __stcs(sc + q * 32, sn); __stcs(lc + q * 32, lu[q]); __stcs(xc + q * 32, xn);
                        if (vpd) {
                            bool in = j != i && lu[q] < age_thr;                                // network.py:547
                            const double xi = sxn[i];
                            double sv;
                            if (flat0) { sv = __dsub_rn(xn, xi); in = in && fabs(sv) < W; }
                            else {
                                const double d = dist2d(xn, sn > 0 ? yj : 0.0, xi, sy[i]);
                                in = in && d < W;                                               // network.py:487
                                sv = (__dsub_rn(xn, xi) > 0.0) ? d : -d;
                            }
                            // trunc(t) is NumPy's edge-corrected bin unless t is within 1e-6 of an edge; formed
                            // unconditionally (clamped garbage when !in), only the reduction is predicated
                            const double t = __dmul_rn(__dadd_rn(sv, W), inv_binw);
                            const double rt = __dadd_rn(__dadd_rn(t, 6755399441055744.0), -6755399441055744.0);
                            int kb = min(max(__double2int_rz(t), 0), B - 1);
                            if (in && fabs(__dsub_rn(t, rt)) < 1e-6) kb = vpd_bin(sv, W, inv_binw, B, s_edges);
                            if (in) atomicAdd(&hist[kb * T + i], 1u);
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (vpd && act) {     // samples per observer (the divisor of network.py:501)
            unsigned m = 0u;
            for (int b = 0; b < B; ++b) m += hist[b * T + tid];
            recv_s[tid] = m;
        }

        // ---- F: state rows (TestEnv.obtain_state, test_env.py:527-583), one warp per row ---------------
        if (want_state) {
            __syncthreads();
            const int n_act = p.add_action ? (p.action_binary ? R : 1) : 0;
            const int o_vpd = n_act + (p.add_channel_obs ? R : 0);
            const int o_tail = o_vpd + (p.piggy ? B : 0);
            for (int u = warp; u < N; u += NWARPS) {
                const int au = sa[u];
                const float den = vpd ? (float)recv_s[u] : 0.0f, rcp = __frcp_rn(den);
                const bool have = vpd && den > 0.0f;
                const float *orow = og + (long long)u * R;     // written by this CTA before the barriers above
                float *srow = p.state + (vbase + u) * S;
                if (p.add_action) {
                    if (p.action_binary) { for (int s = lane; s < R; s += 32) srow[s] = (au == s) ? 1.0f : 0.0f; }
                    else if (lane == 0) srow[0] = (float)au;
                }
                if (p.add_channel_obs) for (int s = lane; s < R; s += 32) srow[n_act + s] = __ldcg(orow + s);
                if (p.piggy)
                    for (int b = lane; b < B; b += 32) {
                        float val = 0.0f;
                        if (have) {
                            const float c = (float)hist[b * T + u];
                            const float q0 = __fmul_rn(c, rcp);
                            val = __fmaf_rn(__fmaf_rn(-q0, den, c), rcp, q0);
                        }
                        srow[o_vpd + b] = val;
                        if (p.vpd_counts) p.vpd_counts[(vbase + u) * p.rec_stride + b] = (unsigned char)(have ? hist[b * T + u] : 0u);
                    }
                if (lane < S - o_tail) {
                    float val = 0.0f; int k = lane;
                    if (p.add_reward)   { if (k == 0) val = s_rew[u]; --k; }
                    if (p.add_index)    { if (k == 0) val = (float)(u + 1); --k; }
                    if (p.add_position) { if (k == 0) val = (float)__ddiv_rn(sxn[u], p.L); if (k == 1) val = (float)__ddiv_rn(sy[u], 2.0); k -= 2; }
                    if (p.add_velocity) { if (k == 0) val = (float)p.vel[vbase + u]; --k; }
                    if (p.fingerprint)  { if (k == 0) val = (float)p.episode; if (k == 1) val = (float)p.epsilon; k -= 2; }
                    srow[o_tail + lane] = val;
                }
            }
        }

        // ---- per-env metric accumulators (fixed-order block reduction for the reward sum) ---------------
        {
            double rs = act ? s_rewd[tid] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
            if (lane == 0 && warp < NW) s_red[warp] = rs;
            __syncthreads();
            if (tid == 0) {
                double trs = 0.0;
                for (int i = 0; i < NW; ++i) trs += s_red[i];
                atomicAdd(p.acc_reward + e, trs);
                unsigned long long *c = reinterpret_cast<unsigned long long *>(p.acc_count + e * ACC_COUNTS);
                atomicAdd(c + 0, (unsigned long long)s_tot[0]); atomicAdd(c + 1, (unsigned long long)s_tot[1]);
                atomicAdd(c + 2, (unsigned long long)s_tot[2]); atomicAdd(c + 3, 1ull);
            }
            __syncthreads();                      // shared memory is recycled by the next environment
        }
    }
}

int block_warps(int N) { return std::max(1, (N + 31) / 32); }

template <int NW, int KM>
cudaError_t prepare_nw(size_t smem)
{
    if (smem <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(step_block_kernel<NW, KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

template <int NW, int KM>
cudaError_t launch_nw(const Params &p, size_t smem, int SB, cudaStream_t stream)
{
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err == cudaSuccess) err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (err == cudaSuccess) err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_block_kernel<NW, KM>, geo_threads(NW), smem);
    if (err != cudaSuccess) return err;
    const long long resident = (long long)sms * std::max(per_sm, 1);
    const unsigned grid = (unsigned)std::min<long long>(p.E, resident);
    step_block_kernel<NW, KM><<<grid, geo_threads(NW), smem, stream>>>(p, SB);
    return cudaGetLastError();
}

// where the table keys live (see BlockSmem): 32-bit keys in shared memory up to 128 vehicles (two CTAs per SM);
// beyond that the packed 16-bit keys only, with the 32-bit fallback in the scratch slice; nothing if even those
// do not fit next to the histogram
int key_mode(const Params &p)
{
    const int nw = block_warps(p.N);
    const bool vpd = p.vpd_enabled != 0;
    if (nw <= 4 && BlockSmem(p.B, nw, vpd, 2).bytes <= SMEM_BUDGET) return 2;
    if (BlockSmem(p.B, nw, vpd, 1).bytes <= SMEM_BUDGET) return 1;
    return 0;
}

template <int NW>
cudaError_t prepare_both(const Params &p)
{
    const int km = key_mode(p);
    const size_t smem = BlockSmem(p.B, NW, p.vpd_enabled != 0, km).bytes;
    if constexpr (NW <= 4) { if (km == 2) return prepare_nw<NW, 2>(smem); }
    return km == 1 ? prepare_nw<NW, 1>(smem) : prepare_nw<NW, 0>(smem);
}

template <int NW>
cudaError_t launch_both(const Params &p, cudaStream_t stream)
{
    const int km = key_mode(p);
    const size_t smem = BlockSmem(p.B, NW, p.vpd_enabled != 0, km).bytes;
    const int SB = key_src_bits(p.N);
    if constexpr (NW <= 4) { if (km == 2) return launch_nw<NW, 2>(p, smem, SB, stream); }
    return km == 1 ? launch_nw<NW, 1>(p, smem, SB, stream) : launch_nw<NW, 0>(p, smem, SB, stream);
}

}  // namespace

int key_src_bits(int N)
{
    int b = 1;
    while ((1 << b) < N) ++b;
    return b;
}

bool step_block_keys_fit_smem(const Params &p) { return key_mode(p) == 2; }

size_t step_block_smem_bytes(const Params &p, bool /*keys_in_smem: derived from p*/)
{
    return BlockSmem(p.B, block_warps(p.N), p.vpd_enabled != 0, key_mode(p)).bytes;
}

size_t step_block_scratch_words_per_env(int N) { return (size_t)N * geo_ld(block_warps(N)); }

size_t step_block_scratch_bytes(long long E, int N)
{
    return (size_t)E * step_block_scratch_words_per_env(N) * sizeof(unsigned);
}

cudaError_t prepare_step_block(const Params &p)
{
    switch (block_warps(p.N)) {
    case 1: return prepare_both<1>(p);
    case 2: return prepare_both<2>(p);
    case 3: return prepare_both<3>(p);
    case 4: return prepare_both<4>(p);
    case 5: return prepare_both<5>(p);
    case 6: return prepare_both<6>(p);
    case 7: return prepare_both<7>(p);
    default: return prepare_both<8>(p);
    }
}

cudaError_t launch_step_block(const Params &p, cudaStream_t stream)
{
    switch (block_warps(p.N)) {
    case 1: return launch_both<1>(p, stream);
    case 2: return launch_both<2>(p, stream);
    case 3: return launch_both<3>(p, stream);
    case 4: return launch_both<4>(p, stream);
    case 5: return launch_both<5>(p, stream);
    case 6: return launch_both<6>(p, stream);
    case 7: return launch_both<7>(p, stream);
    default: return launch_both<8>(p, stream);
    }
}

}  // namespace diral
