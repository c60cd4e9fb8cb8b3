// diral_step_row.cu -- fused time-slot kernel for 32 < N <= 256 vehicles, ROW layout (one CTA per environment,
// persistent CTAs; replaces diral_step_block.cu wherever the configuration allows, see step_row_supported()).
//
// What is different from the one-CTA-per-env kernel of round 1 (same slot semantics, same packed keys):
//
//   * Tables are stored OBSERVER-major with a padded row stride T (a multiple of 64): tab_seq / tab_lu [E][N][T].
//     Row i -- everything vehicle i believes (reference envs/vehicle.py:20-33) -- is contiguous, so one warp owns a
//     row from its first load to its last store: 16-byte loads, keys formed in registers, merged in registers, decoded,
//     aged, binned and written back without ever being transposed through shared memory.
//   * xpos is not a table any more.  An entry's position is a pure function of (subject, sequence number) -- it is
//     pos_x[subject] at the tick that made that sequence number (vehicle.py:58-60) -- so the environment keeps a
//     position RING ring[e][tick mod H][j] (8 N bytes written per slot) instead of 8 N^2 bytes of xpos read and
//     written per slot.  The ring (H*T*8 bytes, contiguous) is staged into shared memory with ONE cp.async.bulk per
//     environment, completion on an mbarrier that is only waited for when the epilogue starts; the slot-major layout
//     makes the per-entry lookup bank-conflict free.  Entries older than H ticks (sparse highways) take their position
//     from a ping-pong spill table (tab_x, two [E][N][T] halves: read last slot's half, write this slot's), which dense
//     highways never touch.  diral_materialize_x() rebuilds the reference's xpos table for inspection / parity.
//   * Merges are RECEIVER-centric and need no per-pass barrier.  A vehicle transmits on exactly one resource, so its
//     row is read by others in exactly one pass, a[t]: what they must see is SNAPSHOT(t) = row t after t's receptions
//     in passes < a[t] (Vehicle.periodic_update aliases the live table, vehicle.py:61).  Phase 1 forms the snapshots
//     in transmit order -- vehicles sorted by (resource, id), one warp per row, a warp waits on a shared-memory flag
//     per source row, sources always sort earlier so the wait chain is acyclic -- and publishes them in shared memory.
//     Phase 2 is embarrassingly parallel: row u = SNAPSHOT(u) joined with the snapshots heard in passes > a[u]; nobody
//     reads row u after its own pass, so the result stays in registers and goes straight into the epilogue.
//     One CTA barrier separates the phases; the round-1 kernel needed one per non-empty resource (56 at 128 x 64).
//   * Decisions: the common case (t is the only in-range transmitter on its resource) is decided in a branch-free
//     pass over the in-range pairs; contested (receiver, resource) pairs are only marked there and resolved in a
//     second, short pass, so no warp drags the nearest-of-several search through every iteration.
//   * The epilogue and TestEnv.obtain_state are one loop per row: decode, age, look the position up, bin the
//     positional distribution into a per-warp histogram, write seq / last_updated back with 16-byte stores and emit the
//     state row -- no whole-environment histogram, no second pass over the vehicles.
//
// HBM traffic per entry is 16 B (seq and last_updated, read + written) instead of 32 B.
#include "diral_dev.cuh"
#include "diral_launch.h"

#include <algorithm>
#include <type_traits>

namespace diral {

namespace {

__host__ __device__ constexpr size_t align16r(size_t x) { return (x + 15) & ~(size_t)15; }

// Shared-memory carve-up (bytes) of the T-vehicle instantiation; the host computes the same numbers.
struct RowSmem {
    size_t off_sx, off_sy, off_sxn, off_rewd, off_sa, off_aux, off_rew, off_recv, off_flag, off_order, off_base, off_txm, off_inr,
        off_rmask, off_cmask, off_any, off_src, off_keys, off_ring, off_whist, off_edges, off_misc, bytes;
    __host__ __device__ RowSmem(int T, int R, int B, int H, int nwarps)
    {
        const int NW = T / 32, NWP = NW | 1, RW = (R + 31) / 32;
        size_t o = 0;
        off_sx = o;    o += align16r(8 * (size_t)T);
        off_sy = o;    o += align16r(8 * (size_t)T);
        off_sxn = o;   o += align16r(8 * (size_t)T);
        off_rewd = o;  o += align16r(8 * (size_t)T);
        off_sa = o;    o += align16r(4 * (size_t)T);
        off_aux = o;   o += align16r(4 * (size_t)T);
        off_rew = o;   o += align16r(4 * (size_t)T);
        off_recv = o;  o += align16r(4 * (size_t)T);
        off_flag = o;  o += align16r(4 * (size_t)T);
        off_order = o; o += align16r(2 * (size_t)T);
        off_base = o;  o += align16r(4 * (size_t)(R + 1));
        off_txm = o;   o += align16r(4 * (size_t)R * NWP);
        off_inr = o;   o += align16r(4 * (size_t)T * NWP);
        off_rmask = o; o += align16r(4 * (size_t)T * RW);
        off_cmask = o; o += align16r(4 * (size_t)T * RW);
        off_any = o;   o += align16r(4 * (size_t)RW);
        off_src = o;   o += align16r((size_t)T * R);
        off_keys = o;  o += align16r(2 * (size_t)T * T);
        off_ring = o;  o += align16r(8 * (size_t)H * T);
        off_whist = o; o += align16r(4 * (size_t)nwarps * (B + 1));     // + 1: the bin of samples that do not count
        off_edges = o; o += align16r(8 * (size_t)(B + 1));
        off_misc = o;  o += 256;
        bytes = o;
    }
};

constexpr size_t ROW_SMEM_BUDGET = 226 * 1024;

// Network.calculate_reward_weights / calculate_avg_distance (network.py:273-316) over the transmitters whose bits are
// set in m[0..NW), ascending ids, Python sum() semantics
template <int NW>
__device__ __noinline__ int row_reward_weight(const Params &p, const double *sx, const double *sy, const unsigned *m, double norm)
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

__device__ __forceinline__ unsigned smem_addr(const void *ptr) { return (unsigned)__cvta_generic_to_shared(ptr); }

__device__ __forceinline__ int ld_acquire_smem(const int *ptr)
{
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_addr(ptr)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_smem(int *ptr, int v)
{
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_addr(ptr)), "r"(v) : "memory");
}

// lane-contiguous pieces of a row: NI 32-bit words at `ptr` (8- or 16-byte vectors where the count allows)
template <int NI, bool GLOBAL_CG>
__device__ __forceinline__ void load_words(const unsigned *ptr, unsigned (&v)[NI])
{
    if constexpr (NI % 4 == 0) {
#pragma unroll
        for (int q = 0; q < NI / 4; ++q) {
            const uint4 t = GLOBAL_CG ? __ldcg(reinterpret_cast<const uint4 *>(ptr) + q) : reinterpret_cast<const uint4 *>(ptr)[q];
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    } else if constexpr (NI % 2 == 0) {
#pragma unroll
        for (int q = 0; q < NI / 2; ++q) {
            const uint2 t = GLOBAL_CG ? __ldcg(reinterpret_cast<const uint2 *>(ptr) + q) : reinterpret_cast<const uint2 *>(ptr)[q];
            v[2 * q] = t.x; v[2 * q + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int q = 0; q < NI; ++q) v[q] = GLOBAL_CG ? __ldcg(ptr + q) : ptr[q];
    }
}
template <int NI, bool GLOBAL_CG>
__device__ __forceinline__ void store_words(unsigned *ptr, const unsigned (&v)[NI])
{
    if constexpr (NI % 4 == 0) {
#pragma unroll
        for (int q = 0; q < NI / 4; ++q) {
            const uint4 t = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            if (GLOBAL_CG) __stcg(reinterpret_cast<uint4 *>(ptr) + q, t); else reinterpret_cast<uint4 *>(ptr)[q] = t;
        }
    } else if constexpr (NI % 2 == 0) {
#pragma unroll
        for (int q = 0; q < NI / 2; ++q) {
            const uint2 t = make_uint2(v[2 * q], v[2 * q + 1]);
            if (GLOBAL_CG) __stcg(reinterpret_cast<uint2 *>(ptr) + q, t); else reinterpret_cast<uint2 *>(ptr)[q] = t;
        }
    } else {
#pragma unroll
        for (int q = 0; q < NI; ++q) { if (GLOBAL_CG) __stcg(ptr + q, v[q]); else ptr[q] = v[q]; }
    }
}

#ifndef DIRAL_ROW_THREADS_64
#define DIRAL_ROW_THREADS_64 256      // tuning knob: threads per CTA of the 64-vehicle instantiation
#endif
#ifndef DIRAL_ROW_THREADS_128
#define DIRAL_ROW_THREADS_128 512     // tuning knob: threads per CTA of the 128-vehicle instantiation
#endif
__host__ __device__ constexpr int row_threads(int nw2) { return nw2 == 1 ? DIRAL_ROW_THREADS_64 : (nw2 == 2 ? DIRAL_ROW_THREADS_128 : 256 * nw2); }
__host__ __device__ constexpr int row_min_ctas(int nw2) { return nw2 == 1 ? 1024 / DIRAL_ROW_THREADS_64 : (nw2 == 2 ? (DIRAL_ROW_THREADS_128 == 512 ? 2 : 3) : 1); }

template <int NW2>
__global__ void __launch_bounds__(row_threads(NW2), row_min_ctas(NW2))
step_row_kernel(const Params p, const int SB)
{
    constexpr int T = 64 * NW2;                   // padded vehicle count = row stride of the tables
    constexpr int NW = 2 * NW2, NWP = NW | 1;     // 32-bit words of a vehicle bit mask (odd stride in shared memory)
    constexpr int TT = row_threads(NW2), NWARPS = TT / 32, RPW = T / NWARPS;   // rows per warp (8; 16 with 128-thread CTAs)
    constexpr int KPL = NW;                       // table columns per lane of a row warp (T / 32)
    constexpr int WPL = NW2;                      // packed key words per lane (two 16-bit keys per word)
    constexpr int T2 = T / 2;                     // packed key words per row
    constexpr unsigned FULL = 0xffffffffu;
    const int N = p.N, R = p.R, B = p.B, S = p.S, H = p.H;
    const int RW = (R + 31) >> 5;
    int tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int lane = tid & 31, warp = tid >> 5;
    const bool act = tid < N;
    const bool want_state = p.build_state != 0;
    const bool vpd = want_state && p.vpd_enabled;
    const int mode = p.mode;
    const bool merge_mode = mode != MODE_STEP || p.state_type == 1 || p.state_type == 2;
    const double Cr = p.C, sentinel = p.sentinel;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const RowSmem lay(T, R, B, H, NWARPS);
    double *sx = reinterpret_cast<double *>(smem_raw + lay.off_sx);
    double *sy = reinterpret_cast<double *>(smem_raw + lay.off_sy);
    double *sxn = reinterpret_cast<double *>(smem_raw + lay.off_sxn);
    double *s_rewd = reinterpret_cast<double *>(smem_raw + lay.off_rewd);
    int *sa = reinterpret_cast<int *>(smem_raw + lay.off_sa);
    int *s_aux = reinterpret_cast<int *>(smem_raw + lay.off_aux);
    float *s_rew = reinterpret_cast<float *>(smem_raw + lay.off_rew);
    unsigned *recv_s = reinterpret_cast<unsigned *>(smem_raw + lay.off_recv);
    int *flag_s = reinterpret_cast<int *>(smem_raw + lay.off_flag);
    unsigned short *order_s = reinterpret_cast<unsigned short *>(smem_raw + lay.off_order);
    int *base_s = reinterpret_cast<int *>(smem_raw + lay.off_base);
    unsigned *txm_s = reinterpret_cast<unsigned *>(smem_raw + lay.off_txm);       // [R][NWP]
    unsigned *inr_s = reinterpret_cast<unsigned *>(smem_raw + lay.off_inr);       // [T][NWP]
    unsigned *rmask_s = reinterpret_cast<unsigned *>(smem_raw + lay.off_rmask);   // [T][RW] resources heard
    unsigned *cmask_s = reinterpret_cast<unsigned *>(smem_raw + lay.off_cmask);   // [T][RW] resources with >= 2 candidates
    unsigned char *src_s = smem_raw + lay.off_src;                                // [T][R] whom u hears on r
    unsigned *K = reinterpret_cast<unsigned *>(smem_raw + lay.off_keys);          // [T][T2] packed 16-bit keys
    double *ring_s = reinterpret_cast<double *>(smem_raw + lay.off_ring);         // [H][T]
    unsigned *whist = reinterpret_cast<unsigned *>(smem_raw + lay.off_whist) + warp * (B + 1);
    unsigned *any_s = reinterpret_cast<unsigned *>(smem_raw + lay.off_any);       // [RW] resources with a transmitter
    double *s_edges = reinterpret_cast<double *>(smem_raw + lay.off_edges);
    double *s_red = reinterpret_cast<double *>(smem_raw + lay.off_misc);          // [8]
    unsigned *s_tot = reinterpret_cast<unsigned *>(smem_raw + lay.off_misc + 64);   // received, pairs, bad
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_raw + lay.off_misc + 96);

    const unsigned srcmask = (1u << SB) - 1u;
    const int FMAX = (1 << (16 - SB)) - 1, kbase = p.tick - FMAX, tick = p.tick;
    const long long envN = (long long)N * T;      // table elements per environment
    const unsigned ring_bytes = (unsigned)(8u * (unsigned)H * T);

    for (int i = tid; i <= B; i += TT) s_edges[i] = p.edges[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned ring_phase = 0;

    const bool pf_ok = (long long)gridDim.x * envN * 8 <= (80ll << 20);
    for (long long e = blockIdx.x; e < p.E; e += gridDim.x) {
        const long long vbase = e * N, tbase = e * envN;
        int32_t *seqg = p.tab_seq + tbase, *lug = p.tab_lu + tbase;

        // ---- ring -> shared memory: one bulk copy, waited for when the epilogue starts ------------------------------
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // last environment's reads of ring_s are done
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(mbar)), "r"(ring_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_addr(ring_s)), "l"(p.ring + e * (long long)H * T), "r"(ring_bytes), "r"(smem_addr(mbar)) : "memory");
        }
        if (pf_ok) {         // ages arrive in L2 while decisions and merges run
            const char *b1 = reinterpret_cast<const char *>(lug);
            for (int o = tid * 128; o < (int)(envN * 4); o += TT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(b1 + o));
        }

        // ---- A: inputs ---------------------------------------------------------------------------------------------
        if (tid < 4) s_tot[tid] = 0u;
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
        if (tid < T) flag_s[tid] = 0;
        for (int i = tid; i < R * NWP; i += TT) txm_s[i] = 0u;
        for (int i = tid; i < T * RW; i += TT) { rmask_s[i] = 0u; cmask_s[i] = 0u; }
        const bool flat = __syncthreads_and(y_same) != 0;              // also publishes sx, sy, sa
        const bool flat0 = flat && sy[0] == 0.0;
        if (act && s_aux[tid] < 0) { atomicAdd(&s_tot[2], 1u); s_aux[tid] = 0; }

        // ---- keys: one warp per row, packed in registers ------------------------------------------------------------
        // K[i][j] = (seq[i][j] (+1 on the diagonal: the tick, vehicle.py:58)) - kbase << SB | i, 0 stays 0 ("never heard")
        unsigned oldest = 0xffffffffu;
#pragma unroll 2
        for (int k = 0; k < RPW; ++k) {
            const int i = warp + k * NWARPS;
            if (i >= N) break;
            unsigned sq[KPL];
            load_words<KPL, false>(reinterpret_cast<const unsigned *>(seqg + (long long)i * T + lane * KPL), sq);
            unsigned kw[WPL];
#pragma unroll
            for (int q = 0; q < KPL; ++q) {
                int s = (int)sq[q];
                if (lane * KPL + q == i) s += 1;
                oldest = min(oldest, (unsigned)(s - 1));
                const unsigned f = s ? (unsigned)(s - kbase) : 0u;
                const unsigned key = ((f << SB) | (unsigned)i) & 0xffffu;
                if (q & 1) kw[q >> 1] |= key << 16; else kw[q >> 1] = key;
            }
            store_words<WPL, false>(K + i * T2 + lane * WPL, kw);
        }
        // who is within communication range of whom (Network.check_communicaiton_range, network.py:595-607)
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
        // per-resource collision histogram (test_env.py:149-157)
        const int a_me = act ? sa[tid] : -1;
        if (act) atomicOr(&txm_s[a_me * NWP + warp], 1u << lane);
        const bool wide = __syncthreads_or(kbase > 0 && oldest < (unsigned)kbase) != 0;   // uniform; publishes K, inr_s, txm_s
        unsigned *Kw = p.scratch + (size_t)e * T * T;                  // 32-bit keys of a "wide" environment (rare)
        if (wide) {
            for (int k = 0; k < RPW; ++k) {
                const int i = warp + k * NWARPS;
                if (i >= N) break;
                unsigned sq[KPL];
                load_words<KPL, false>(reinterpret_cast<const unsigned *>(seqg + (long long)i * T + lane * KPL), sq);
#pragma unroll
                for (int q = 0; q < KPL; ++q) {
                    int s = (int)sq[q];
                    if (lane * KPL + q == i) s += 1;
                    sq[q] = ((unsigned)s << SB) | (unsigned)i;
                }
                store_words<KPL, true>(Kw + (size_t)i * T + lane * KPL, sq);
            }
        }

        // ---- B: rewards that need nothing but the collision set, transmit order ------------------------------------
        if (act) {
            unsigned own[NW], inr[NW];
            int my_tot = 0, in_range = 0; double rew = 0.0;
            const double x = sx[tid], y = sy[tid];
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                own[w] = txm_s[a_me * NWP + w]; inr[w] = inr_s[tid * NWP + w]; my_tot += __popc(own[w]);
            }
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
                        wgt = row_reward_weight<NW>(p, sx, sy, own, norm);
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
        // channel observations before any reception (test_env.py:203-240 / :305-306 / :431): one warp per vehicle,
        // coalesced rows; a reception then overwrites its own entry with the distance (after the barrier below)
        float *og = p.obs + vbase * R;
        const bool want_d = mode == MODE_STEP && p.state_type == 2;
        {
            const float basev = (mode != MODE_STEP || p.state_type == 1) ? 1.0f : (p.state_type == 2 ? (float)sentinel : 0.0f);
            for (int g = 0; g < RW; ++g) {
                const int r = g * 32 + lane;
                unsigned busy = 0u;
                if (r < R) {
#pragma unroll
                    for (int w = 0; w < NW; ++w) busy |= txm_s[r * NWP + w];
                }
                if (r < R)
                    for (int u = warp; u < N; u += NWARPS) og[(long long)u * R + r] = (busy != 0u && sa[u] != r) ? basev : 0.0f;
            }
        }
        if (warp == 0) {     // exclusive scan of the transmitters per resource: where a resource's vehicles start in transmit order
            int carry = 0;
            for (int r0 = 0; r0 < R; r0 += 32) {
                const int r = r0 + lane;
                int c = 0;
                if (r < R) {
#pragma unroll
                    for (int w = 0; w < NW; ++w) c += __popc(txm_s[r * NWP + w]);
                }
                int inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += v; }
                if (r < R) base_s[r] = carry + inc - c;
                const unsigned busy = __ballot_sync(FULL, c > 0);
                if (lane == 0) any_s[r0 >> 5] = busy;
                carry += __shfl_sync(FULL, inc, 31);
            }
        }
        __syncthreads();
        if (act) {           // my position in transmit order: vehicles sorted by (resource, id)
            int pos = base_s[a_me];
            for (int w = 0; w < warp; ++w) pos += __popc(txm_s[a_me * NWP + w]);
            pos += __popc(txm_s[a_me * NWP + warp] & ((1u << lane) - 1u));
            order_s[pos] = (unsigned short)tid;
        }

        // ---- B: decisions, no table access.  Thread (u, w, part) walks the in-range vehicles t of one part of word w: every
        // t transmits on exactly one resource, so t is u's nearest in-range transmitter there (Network.find_closest_tx,
        // network.py:378-398) unless another in-range vehicle shares t's resource; those (u, resource) pairs are marked
        // and resolved below.
        int n_recv = 0, n_pairs = 0;
        auto reception = [&](int u, int t, int at, double d) {       // u hears t on resource at, d metres away
            ++n_recv;
            if (want_d) og[(long long)u * R + at] = (float)d;
            if (mode == MODE_CH) {
                atomicAdd(&recv_s[t], 1u);                                                 // test_env.py:396-397
                if (p.track_lat) p.lat[e * (long long)N * N + (long long)t * N + u] = (int32_t)p.timestep;   // test_env.py:436
            }
            src_s[u * R + at] = (unsigned char)t;
            atomicOr(&rmask_s[u * RW + (at >> 5)], 1u << (at & 31));
        };
        {
            constexpr int SUBS = TT > T * NW ? TT / (T * NW) : 1;
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
                    const int bit = __ffs(c) - 1, t = w * 32 + bit;
                    const int at = sa[t];
                    if (at == au) continue;                          // u transmits there itself (half duplex)
                    int ncand = 0, lower = 0;
#pragma unroll
                    for (int w2 = 0; w2 < NW; ++w2) {
                        const unsigned cw = inr[w2] & txm_s[at * NWP + w2];      // in-range vehicles sharing t's resource
                        ncand += __popc(cw);
                        lower += w2 < w ? __popc(cw) : (w2 == w ? __popc(cw & ((1u << bit) - 1u)) : 0);
                    }
                    n_pairs += 1;
                    if (ncand == 1) {
                        const double d = flat ? fabs(__dsub_rn(xu, sx[t])) : dist2d(sx[t], sy[t], xu, yu);
                        if (d < sentinel) reception(u, t, at, d);    // best starts at the sentinel (network.py:380)
                    } else if (lower == 0) {
                        atomicOr(&cmask_s[u * RW + (at >> 5)], 1u << (at & 31));   // several candidates: resolved once, below
                    }
                }
                if (p.track_lat) {                                                       // network.py:394
                    const int nlive = N - w * 32;               // (T is padded to 64: whole words may lie beyond N)
                    const unsigned live = nlive >= 32 ? 0xffffffffu : (nlive > 0 ? (1u << nlive) - 1u : 0u);
                    for (unsigned c = ~mine_w & live & part_mask; c; c &= c - 1) {
                        const int t = w * 32 + __ffs(c) - 1;
                        if (sa[t] != au) p.lat[e * (long long)N * N + (long long)t * N + u] = -1;
                    }
                }
            }
        }
        __syncthreads();
        // contested (receiver, resource) pairs: ascending ids, strict '<': the first minimum wins (network.py:384-391)
        for (int it = tid; it < N * RW; it += TT) {
            const int u = it / RW, g = it - u * RW;
            for (unsigned cm = cmask_s[u * RW + g]; cm; cm &= cm - 1) {
                const int r = g * 32 + __ffs(cm) - 1;
                const double xu = sx[u], yu = sy[u];
                double best = sentinel; int tstar = -1;
                for (int w2 = 0; w2 < NW; ++w2)
                    for (unsigned c2 = inr_s[u * NWP + w2] & txm_s[r * NWP + w2]; c2; c2 &= c2 - 1) {
                        const int t2 = w2 * 32 + __ffs(c2) - 1;
                        const double d2 = flat ? fabs(__dsub_rn(xu, sx[t2])) : dist2d(sx[t2], sy[t2], xu, yu);
                        if (d2 < best) { best = d2; tstar = t2; }
                    }
                if (tstar >= 0) reception(u, tstar, r, best);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_recv += __shfl_xor_sync(FULL, n_recv, o);
            n_pairs += __shfl_xor_sync(FULL, n_pairs, o);
        }
        if (lane == 0) {
            if (n_recv) atomicAdd(&s_tot[0], (unsigned)n_recv);
            if (n_pairs) atomicAdd(&s_tot[1], (unsigned)n_pairs);
        }
        __syncthreads();

        // ---- D: rewards out, mobility (Network.update_positions, network.py:189-206) ---------------------------------
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

        // ---- ring: this tick's row is the pre-mobility position of every vehicle (vehicle.py:58-60) ---------------------
        {
            unsigned done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(smem_addr(mbar)), "r"(ring_phase) : "memory");
            ring_phase ^= 1u;
            if (act) {
                const double x = sx[tid];
                ring_s[(tick & (H - 1)) * T + tid] = x;
                p.ring[e * (long long)H * T + (long long)(tick & (H - 1)) * T + tid] = x;
            }
        }
        if (pf_ok && e + gridDim.x < p.E) {       // the next environment's sequence numbers, needed first
            const char *b0 = reinterpret_cast<const char *>(p.tab_seq + (e + gridDim.x) * envN);
            for (int o = tid * 128; o < (int)(envN * 4); o += TT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(b0 + o));
        }

        // ---- C + E: merges and the per-row epilogue -----------------------------------------------------------------
        const double *spill_prev = p.tab_x + ((tick & 1) ? 0 : p.spill_half) + tbase;   // written by the previous slot
        double *spill_cur = p.tab_x + ((tick & 1) ? p.spill_half : 0) + tbase;
        const int n_act = p.add_action ? (p.action_binary ? R : 1) : 0;
        const int o_vpd = n_act + (p.add_channel_obs ? R : 0);
        const int o_tail = o_vpd + B;
        const bool vec_rows = o_tail == S && (S & 3) == 0 && (n_act & 3) == 0 && (B & 3) == 0 && (!p.add_channel_obs || (R & 3) == 0);
        const double W = p.W, inv_binw = p.inv_binw, inv_binw_s = p.inv_binw * 1048576.0;
        const int age_thr = p.age_threshold;

        auto tables = [&](auto wide_c, auto flat_c) {
            constexpr bool WIDE = decltype(wide_c)::value;
            constexpr bool FL0 = decltype(flat_c)::value;            // every vehicle (and every phantom entry) on lane y = 0
            constexpr int NK = WIDE ? KPL : WPL;                     // key words per lane
            auto row_ptr = [&](int i) -> unsigned * { return WIDE ? Kw + (size_t)i * T + lane * KPL : K + i * T2 + lane * WPL; };
            auto merge_from = [&](unsigned (&my)[NK], int s) {       // Vehicle.received_update (vehicle.py:35-47) on a whole row
                unsigned o[NK];
                load_words<NK, WIDE>(row_ptr(s), o);
#pragma unroll
                for (int q = 0; q < NK; ++q) my[q] = WIDE ? max(my[q], o[q]) : __vmaxu2(my[q], o[q]);
            };
            volatile int *vflag = flag_s;
            // phase 1: snapshots in transmit order
            if (merge_mode) {
                for (int pidx = warp; pidx < N; pidx += NWARPS) {
                    const int t = order_s[pidx], at = sa[t];
                    unsigned my[NK];
                    load_words<NK, WIDE>(row_ptr(t), my);
                    bool changed = false;
                    for (int g = 0; g * 32 < at; ++g) {
                        unsigned m = rmask_s[t * RW + g];
                        if (at - g * 32 < 32) m &= (1u << (at - g * 32)) - 1u;      // passes before t's own
                        const int sb = src_s[t * R + min(g * 32 + lane, R - 1)];     // whom t hears on resource g*32 + lane
                        for (; m; m &= m - 1) {
                            const int s = __shfl_sync(FULL, sb, __ffs(m) - 1);
                            while (vflag[s] == 0) { }                               // s sorts before t: it will be published
                            asm volatile("" ::: "memory");
                            merge_from(my, s);
                            changed = true;
                        }
                    }
                    if (changed) store_words<NK, WIDE>(row_ptr(t), my);
                    if (WIDE) __threadfence_block();
                    __syncwarp();
                    if (lane == 0) { __threadfence_block(); vflag[t] = 1; }
                }
            }
            __syncthreads();

            // phase 2 + epilogue, one row at a time
            for (int k = 0; k < RPW; ++k) {
                const int i = warp + k * NWARPS;
                if (i >= N) break;
                const int ai = sa[i];
                unsigned my[NK];
                load_words<NK, WIDE>(row_ptr(i), my);
                unsigned s0[KPL], lu[KPL];
                load_words<KPL, false>(reinterpret_cast<const unsigned *>(seqg + (long long)i * T + lane * KPL), s0);
                load_words<KPL, false>(reinterpret_cast<const unsigned *>(lug + (long long)i * T + lane * KPL), lu);
                if (merge_mode) {
                    for (int g = ai >> 5; g < RW; ++g) {
                        unsigned m = rmask_s[i * RW + g];
                        if (g == (ai >> 5)) m &= ~((2u << (ai & 31)) - 1u);          // passes after i's own
                        const int sb = src_s[i * R + min(g * 32 + lane, R - 1)];
                        for (; m; m &= m - 1) merge_from(my, __shfl_sync(FULL, sb, __ffs(m) - 1));
                    }
                }
                if (vpd) { for (int b = lane; b <= B; b += 32) whist[b] = 0u; __syncwarp(); }
                const double xi = sxn[i], yi = sy[i];
                // Straight-line per entry: decode, age, position from the ring, bin.  Entries whose version is older than the
                // ring (sparse highways) or whose sample sits within 1e-6 of a bin edge are only flagged here and redone
                // exactly below -- the common path has no branch, samples that do not count go to bin B.
                unsigned slow = 0u;
#pragma unroll
                for (int q = 0; q < KPL; ++q) {
                    const int c = lane * KPL + q;
                    int sn;
                    if (WIDE) sn = (int)(my[q] >> SB);
                    else {
                        const unsigned hk = (q & 1) ? (my[q >> 1] >> 16) : (my[q >> 1] & 0xffffu), f = hk >> SB;
                        sn = f ? (int)f + kbase : 0;
                    }
                    // vehicle.py:41-47 / :56-70: a strictly newer version (or the own tick) resets the age, everything else ages
                    const int lun = (sn != (int)s0[q]) ? 0 : (int)lu[q] + 1;
                    s0[q] = (unsigned)sn; lu[q] = (unsigned)lun;
                    const bool old = sn > 0 && tick - sn >= H - 1;                   // leaves (or has left) the ring
                    const double xr = ring_s[(sn & (H - 1)) * T + c];
                    const double xn = sn > 0 ? xr : 0.0;
                    if (vpd) {
                        bool in = c != i && c < N && lun < age_thr;                  // network.py:547
                        double sv;
                        if (FL0) { sv = __dsub_rn(xn, xi); in = in && fabs(sv) < W; }       // network.py:487
                        else {
                            const double d = dist2d(xn, sn > 0 ? sy[min(c, N - 1)] : 0.0, xi, yi);
                            in = in && d < W;
                            sv = (__dsub_rn(xn, xi) > 0.0) ? d : -d;
                        }
                        // trunc(t) is NumPy's edge-corrected bin unless t is within 1e-6 of an edge
                        // (12.20 fixed point: integer part = bin, fraction within 2^-19 of an integer = near an edge)
                        const int ti = __double2int_rz(__dmul_rn(__dadd_rn(sv, W), inv_binw_s));
                        const bool near = ((unsigned)(ti + 2) & 0xfffffu) < 4u;
                        const bool again = old || (in && near);
                        const int kb = (in && !again) ? (ti >> 20) : B;
                        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(smem_addr(whist + kb)));
                        slow |= again ? 1u << q : 0u;
                    } else slow |= old ? 1u << q : 0u;
                }
                store_words<KPL, false>(reinterpret_cast<unsigned *>(seqg + (long long)i * T + lane * KPL), s0);
                store_words<KPL, false>(reinterpret_cast<unsigned *>(lug + (long long)i * T + lane * KPL), lu);
                if (__any_sync(FULL, slow != 0u)) {
#pragma unroll
                    for (int q = 0; q < KPL; ++q) {
                        if (!((slow >> q) & 1u)) continue;
                        const int c = lane * KPL + q, sn = (int)s0[q];
                        unsigned org;
                        if (WIDE) org = my[q] & srcmask;
                        else org = ((q & 1) ? (my[q >> 1] >> 16) : my[q >> 1]) & srcmask;
                        double xn = 0.0;
                        if (sn > 0) {       // position of that version: the ring while it is younger than H ticks, else the spill table
                            xn = (tick - sn < H) ? ring_s[(sn & (H - 1)) * T + c] : spill_prev[(long long)org * T + c];
                            if (tick - sn >= H - 1) spill_cur[(long long)i * T + c] = xn;
                        }
                        if (vpd) {
                            bool in = c != i && c < N && (int)lu[q] < age_thr;
                            double sv;
                            if (FL0) { sv = __dsub_rn(xn, xi); in = in && fabs(sv) < W; }
                            else {
                                const double d = dist2d(xn, sn > 0 ? sy[min(c, N - 1)] : 0.0, xi, yi);
                                in = in && d < W;
                                sv = (__dsub_rn(xn, xi) > 0.0) ? d : -d;
                            }
                            if (in) atomicAdd(&whist[vpd_bin(sv, W, inv_binw, B, s_edges)], 1u);
                        }
                    }
                }

                // state row (TestEnv.obtain_state, test_env.py:527-583)
                if (want_state) {
                    __syncwarp();
                    unsigned cnt0 = 0u;                               // this lane's bins (b = lane, lane + 32, ..): their sum
                    if (vpd) for (int b = lane; b < B; b += 32) cnt0 += whist[b];
                    int m_cnt = (int)cnt0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) m_cnt += __shfl_xor_sync(FULL, m_cnt, o);
                    const float den = (float)m_cnt, rcp = __frcp_rn(den);
                    const bool have = vpd && m_cnt > 0;
                    float *srow = p.state + (vbase + i) * S;
                    auto quotient = [&](unsigned cnt) {            // counts / len, exact (tests/test_host.py)
                        const float cf = (float)cnt, q0 = __fmul_rn(cf, rcp);
                        return have ? __fmaf_rn(__fmaf_rn(-q0, den, cf), rcp, q0) : 0.0f;
                    };
                    if (vec_rows) {           // every block is whole float4 groups: 16-byte stores straight from registers
                        for (int f4 = lane; f4 < (S >> 2); f4 += 32) {
                            const int s4 = f4 << 2;
                            float4 v;
                            if (s4 < n_act) {
                                const int d = ai - s4;
                                v = make_float4(d == 0 ? 1.0f : 0.0f, d == 1 ? 1.0f : 0.0f, d == 2 ? 1.0f : 0.0f, d == 3 ? 1.0f : 0.0f);
                            } else if (s4 < o_vpd) {
                                v = __ldcg(reinterpret_cast<const float4 *>(og + (long long)i * R + (s4 - n_act)));
                            } else {
                                const int b = s4 - o_vpd;
                                const unsigned c0 = have ? whist[b] : 0u, c1 = have ? whist[b + 1] : 0u, c2 = have ? whist[b + 2] : 0u,
                                               c3 = have ? whist[b + 3] : 0u;
                                v = make_float4(quotient(c0), quotient(c1), quotient(c2), quotient(c3));
                                if (p.vpd_counts)
                                    *reinterpret_cast<unsigned *>(p.vpd_counts + (vbase + i) * p.rec_stride + b) = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
                            }
                            *reinterpret_cast<float4 *>(srow + s4) = v;
                        }
                    } else {
                        if (p.add_action) {
                            if (p.action_binary) { for (int s = lane; s < R; s += 32) srow[s] = (ai == s) ? 1.0f : 0.0f; }
                            else if (lane == 0) srow[0] = (float)ai;
                        }
                        if (p.add_channel_obs) for (int s = lane; s < R; s += 32) srow[n_act + s] = __ldcg(og + (long long)i * R + s);
                        for (int b = lane; b < B; b += 32) {
                            const unsigned cnt = have ? whist[b] : 0u;
                            srow[o_vpd + b] = quotient(cnt);
                            if (p.vpd_counts) p.vpd_counts[(vbase + i) * p.rec_stride + b] = (unsigned char)cnt;
                        }
                        if (lane < S - o_tail) {
                            float val = 0.0f; int kk = lane;
                            if (p.add_reward)   { if (kk == 0) val = s_rew[i]; --kk; }
                            if (p.add_index)    { if (kk == 0) val = (float)(i + 1); --kk; }
                            if (p.add_position) { if (kk == 0) val = (float)__ddiv_rn(xi, p.L); if (kk == 1) val = (float)__ddiv_rn(yi, 2.0); kk -= 2; }
                            if (p.add_velocity) { if (kk == 0) val = (float)p.vel[vbase + i]; --kk; }
                            if (p.fingerprint)  { if (kk == 0) val = (float)p.episode; if (kk == 1) val = (float)p.epsilon; kk -= 2; }
                            srow[o_tail + lane] = val;
                        }
                    }
                    __syncwarp();
                }
            }
        };
        __syncthreads();                          // ring row, sxn, og, reception tables complete
        if (wide) { if (flat0) tables(std::true_type{}, std::true_type{}); else tables(std::true_type{}, std::false_type{}); }
        else { if (flat0) tables(std::false_type{}, std::true_type{}); else tables(std::false_type{}, std::false_type{}); }

        // ---- per-env metric accumulators (fixed-order block reduction for the reward sum) ----------------------------
        {
            double rs = act ? s_rewd[tid] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(FULL, rs, o);
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

int row_nw2(int N) { return std::max(1, (N + 63) / 64); }

int row_ring_depth(int N) { return row_nw2(N) == 1 ? 32 : (row_nw2(N) == 4 ? 8 : 16); }

template <int NW2>
cudaError_t prepare_t(size_t smem)
{
    if (smem <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(step_row_kernel<NW2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

template <int NW2>
cudaError_t launch_t(const Params &p, size_t smem, int SB, cudaStream_t stream)
{
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err == cudaSuccess) err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (err == cudaSuccess) err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_row_kernel<NW2>, row_threads(NW2), smem);
    if (err != cudaSuccess) return err;
    const long long resident = (long long)sms * std::max(per_sm, 1);
    const unsigned grid = (unsigned)std::min<long long>(p.E, resident);
    step_row_kernel<NW2><<<grid, row_threads(NW2), smem, stream>>>(p, SB);
    return cudaGetLastError();
}

}  // namespace

int step_row_stride(int N) { return 64 * row_nw2(N); }

int step_row_ring_depth(int N) { return row_ring_depth(N); }

size_t step_row_smem_bytes(const Params &p)
{
    const int nw2 = row_nw2(p.N);
    return RowSmem(64 * nw2, p.R, p.B, row_ring_depth(p.N), row_threads(nw2) / 32).bytes;
}

// The row kernel takes every configuration with neighbour tables and a fused state build between 33 and 256 vehicles
// whose reception tables fit next to the keys; everything else stays with the round-1 kernels.
bool step_row_supported(const Params &p)
{
    if (!p.piggy || p.N <= GROUP_MAX_N || p.N > 256) return false;
    if (p.add_positional_dist || p.pos_dist_type == 1) return false;      // un-fused State variants read the dense xpos table
    if (p.R > 256 || (long long)step_row_stride(p.N) * p.R > 32768) return false;
    return step_row_smem_bytes(p) <= ROW_SMEM_BUDGET;
}

size_t step_row_scratch_bytes(long long E, int N)
{
    const size_t T = (size_t)step_row_stride(N);
    return (size_t)E * T * T * sizeof(unsigned);
}

cudaError_t prepare_step_row(const Params &p)
{
    const size_t smem = step_row_smem_bytes(p);
    switch (row_nw2(p.N)) {
    case 1: return prepare_t<1>(smem);
    case 2: return prepare_t<2>(smem);
    case 3: return prepare_t<3>(smem);
    default: return prepare_t<4>(smem);
    }
}

cudaError_t launch_step_row(const Params &p, cudaStream_t stream)
{
    const size_t smem = step_row_smem_bytes(p);
    const int SB = key_src_bits(p.N);
    switch (row_nw2(p.N)) {
    case 1: return launch_t<1>(p, smem, SB, stream);
    case 2: return launch_t<2>(p, smem, SB, stream);
    case 3: return launch_t<3>(p, smem, SB, stream);
    default: return launch_t<4>(p, smem, SB, stream);
    }
}

}  // namespace diral
