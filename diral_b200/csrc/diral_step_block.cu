// diral_step_block.cu -- fused time-slot kernel for any N (one CTA per environment).
//
// Same slot semantics as diral_step_group.cu (which handles N <= 32 in registers); here thread u is
// vehicle u, and the packed table keys (seq << SB | origin-row) live in shared memory as
// K[u][j] with an odd row stride, or in the scratch buffer when N*(N+1)*4 bytes do not fit.
//
// The merge exploits that table COLUMNS are independent: for a fixed subject j the passes are
//     K[u][j] = max(K[u][j], K[nearest(u, r)][j])      for r = 0..R-1 in order, all receivers u
// and nearest(u, r) does not depend on j.  So phase 2 lets every vehicle find its nearest in-range
// transmitter per resource (Network.find_closest_tx, reference envs/network.py:378-398) into a
// small table ts[r][u]; phase 3 turns the CTA around -- thread j owns COLUMN j and walks the passes
// sequentially by itself: no barrier between resource passes, conflict-free shared-memory access
// (lanes differ in j), and exactly the merges that happen (Vehicle.received_update,
// vehicle.py:35-47).  Phase 5 turns it back (thread u = observer) to stream the columns once:
// gather xpos from the origin row, age, write back, and bin the positional distribution
// (network.py:473-513,538-558).
#include "diral_dev.cuh"
#include "diral_launch.h"

namespace diral {

namespace {

constexpr unsigned short TS_NONE = 0xFFFFu;

__host__ __device__ inline size_t align16z(size_t x) { return (x + 15) & ~(size_t)15; }

struct BlockSmem {
    size_t off_sx, off_sy, off_edges, off_sa, off_cnt, off_off, off_txl, off_inr, off_recv, off_red,
           off_ts, off_hist, off_keys, bytes;
    __host__ __device__ BlockSmem(int N, int R, int B, int T, bool vpd_state, bool keys_in_smem)
    {
        size_t o = 0;
        off_sx = o;    o += align16z(8 * (size_t)N);
        off_sy = o;    o += align16z(8 * (size_t)N);
        off_edges = o; o += align16z(8 * (size_t)(B + 1));
        off_sa = o;    o += align16z(4 * (size_t)N);
        off_cnt = o;   o += align16z(4 * (size_t)R);
        off_off = o;   o += align16z(4 * (size_t)(R + 1));
        off_txl = o;   o += align16z(4 * (size_t)N);
        off_inr = o;   o += align16z(4 * (size_t)N);
        off_recv = o;  o += align16z(4 * (size_t)N);
        off_red = o;   o += align16z(8 * 4 * 32);
        off_ts = o;    o += align16z(2 * (size_t)R * N);
        off_hist = o;  o += vpd_state ? align16z(4 * (size_t)B * T) : 0;
        off_keys = o;  o += keys_in_smem ? align16z(4 * (size_t)N * (N + 1)) : 0;
        bytes = o;
    }
};

constexpr size_t SMEM_BUDGET = 200 * 1024;

__device__ __forceinline__ int block_reward_weight(const Params &p, const double *sx, const double *sy,
                                                   const int *txl, int lo, int hi, double norm)
{
    PySum s; int pairs = 0;
    for (int i = lo; i < hi; ++i)
        for (int j = i + 1; j < hi; ++j) {
            s.add(dist2d(sx[txl[i]], sy[txl[i]], sx[txl[j]], sy[txl[j]]));
            ++pairs;
        }
    const double m = __ddiv_rn(s.result(), (double)pairs);
    return p.toy ? (m == norm) : (m > p.C);
}

__global__ void step_block_kernel(const Params p, const int SB, const int keys_in_smem)
{
    const int N = p.N, R = p.R, B = p.B, T = blockDim.x;
    const int u = threadIdx.x;
    const bool act = u < N;
    const long long e = blockIdx.x;
    const long long vbase = e * N, tbase = e * (long long)N * N;
    const bool want_state = p.build_state != 0;
    const bool vpd = want_state && p.vpd_enabled;
    const int ld = N + 1;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const BlockSmem lay(N, R, B, T, want_state && p.vpd_enabled, keys_in_smem != 0);
    double *sx = reinterpret_cast<double *>(smem_raw + lay.off_sx);
    double *sy = reinterpret_cast<double *>(smem_raw + lay.off_sy);
    double *s_edges = reinterpret_cast<double *>(smem_raw + lay.off_edges);
    int *sa = reinterpret_cast<int *>(smem_raw + lay.off_sa);
    int *cnt = reinterpret_cast<int *>(smem_raw + lay.off_cnt);
    int *off = reinterpret_cast<int *>(smem_raw + lay.off_off);
    int *txl = reinterpret_cast<int *>(smem_raw + lay.off_txl);
    int *s_inr = reinterpret_cast<int *>(smem_raw + lay.off_inr);
    int *s_recv = reinterpret_cast<int *>(smem_raw + lay.off_recv);
    double *s_red = reinterpret_cast<double *>(smem_raw + lay.off_red);
    unsigned short *ts = reinterpret_cast<unsigned short *>(smem_raw + lay.off_ts);
    unsigned *hist = reinterpret_cast<unsigned *>(smem_raw + lay.off_hist);
    unsigned *K = keys_in_smem ? reinterpret_cast<unsigned *>(smem_raw + lay.off_keys)
                               : p.scratch + (size_t)e * N * ld;
    const unsigned srcmask = (1u << SB) - 1u;

    // ---- phase 1: inputs, per-resource collision histogram, tick, publish keys ---------------------
    for (int i = u; i <= B; i += T) s_edges[i] = p.edges[i];
    for (int r = u; r < R; r += T) cnt[r] = 0;
    int a = -1; double x = 0.0, y = 0.0, v = 0.0; int bad = 0;
    if (act) {
        a = p.gen_actions ? philox_action(p.seed, u, p.env0 + e, p.timestep, R) : p.actions[vbase + u];
        if (a < 0 || a >= R) { bad = 1; a = min(max(a, 0), R - 1); }
        if (p.gen_actions && p.actions_out) p.actions_out[vbase + u] = a;
        x = p.pos_x[vbase + u]; y = p.pos_y[vbase + u]; v = p.vel[vbase + u];
        sx[u] = x; sy[u] = y; sa[u] = a;
        s_inr[u] = 0; s_recv[u] = 0;
    }
    __syncthreads();
    if (act) atomicAdd(&cnt[a], 1);                       // test_env.py:149-157
    if (p.piggy && act) {
        const int32_t *seqp = p.tab_seq + tbase;
        for (int j = 0; j < N; ++j) {
            int s = seqp[j * N + u];
            if (j == u) s += 1;                           // vehicle.py:58
            K[u * ld + j] = ((unsigned)s << SB) | (unsigned)u;
        }
    }
    __syncthreads();
    if (u == 0) { int o = 0; for (int r = 0; r < R; ++r) { off[r] = o; o += cnt[r]; } off[R] = o; }
    __syncthreads();
    if (act) {   // transmitter lists, ascending id inside every resource
        int rank = 0;
        for (int t = 0; t < u; ++t) rank += (sa[t] == a);
        txl[off[a] + rank] = u;
    }
    __syncthreads();

    // toy reward: first-min-x / first-max-x vehicle (network.py:225-246); every thread scans (N small)
    double norm = 0.0;
    if (p.toy && p.mode == MODE_STEP && act && cnt[a] > 1 && design_needs_weight(p.reward_design, cnt[a])) {
        double xmin = p.L + 1.0, xmax = -p.L - 1.0; int imin = 0, imax = 0;
        for (int t = 0; t < N; ++t) {
            if (sx[t] < xmin) { xmin = sx[t]; imin = t; }
            if (sx[t] > xmax) { xmax = sx[t]; imax = t; }
        }
        norm = dist2d(sx[imin], sy[imin], sx[imax], sy[imax]);
    }

    // ---- phase 2: nearest transmitter per resource, observations, rewards, last_arrival_time -------
    double rew = 0.0;
    int n_recv = 0, n_pairs = 0;
    int32_t *latp = p.track_lat ? p.lat + tbase : nullptr;
    const bool merge_mode = p.piggy && (p.mode != MODE_STEP || p.state_type == 1 || p.state_type == 2);
    if (act) {
        float *og = p.obs + (vbase + u) * R;
        for (int r = 0; r < R; ++r) {
            const int lo = off[r], hi = off[r + 1], tot = hi - lo;
            if (tot == 0) { og[r] = 0.0f; ts[r * N + u] = TS_NONE; continue; }
            const bool is_tx = (a == r);
            double best = p.sentinel; int tstar = -1;
            if (!is_tx) {
                for (int k = lo; k < hi; ++k) {
                    const int t = txl[k];
                    const double d = dist2d(sx[t], sy[t], x, y);
                    if (d < p.C) {
                        ++n_pairs;
                        if (d < best) { best = d; tstar = t; }
                        if (p.mode == MODE_CH && tot > 1) atomicAdd(&s_inr[t], 1);
                    } else if (latp) latp[t * N + u] = -1;                          // network.py:394
                }
                if (tstar >= 0) {
                    ++n_recv;
                    if (p.mode == MODE_CH) {
                        if (tot > 1) atomicAdd(&s_recv[tstar], 1);
                        if (latp) latp[tstar * N + u] = (int32_t)p.timestep;        // test_env.py:436
                    }
                }
            }
            ts[r * N + u] = (merge_mode && tstar >= 0) ? (unsigned short)tstar : TS_NONE;
            float o = 0.0f;
            if (!is_tx) {
                if (p.mode == MODE_STEP) o = p.state_type == 2 ? (float)best : (p.state_type == 1 ? 1.0f : 0.0f);
                else o = 1.0f;
            }
            og[r] = o;
            if (is_tx) {
                if (p.mode == MODE_STEP) {
                    if (tot == 1) rew = 1.0;
                    else {
                        int w = 0;
                        if (design_needs_weight(p.reward_design, tot)) w = block_reward_weight(p, sx, sy, txl, lo, hi, norm);
                        rew = collision_reward_step(p.reward_design, tot, w);
                    }
                } else if (p.mode == MODE_DESIGN) {
                    if (tot == 1) rew = 1.0;
                    else {   // TestEnv.calculate_reward_design (test_env.py:319-349)
                        int k = 1, last = u;
                        for (int q = lo; q < hi; ++q) {
                            const int t = txl[q];
                            if (t != u && dist2d(x, y, sx[t], sy[t]) < p.C2) { ++k; last = t; }
                        }
                        if (k == 1) rew = 1.0;
                        else if (k == 2) rew = (dist2d(x, y, sx[last], sy[last]) > p.C2) ? 0.0 : -2.0;
                        else rew = -(double)k;
                    }
                }
            }
        }
    }
    __syncthreads();
    if (act && p.mode == MODE_CH) rew = channel_reward(p.reward_design, cnt[a], s_recv[u], s_inr[u]);
    if (act) p.rews[vbase + u] = (float)rew;

    // ---- phase 3: thread j owns column j and replays the passes on it ------------------------------
    if (merge_mode && act) {
        const int j = u;
        for (int r = 0; r < R; ++r) {
            if (off[r + 1] == off[r]) continue;
            const unsigned short *tr = ts + r * N;
            for (int i = 0; i < N; ++i) {
                const unsigned t = tr[i];
                if (t != TS_NONE) {
                    const unsigned mine = K[i * ld + j], theirs = K[t * ld + j];
                    K[i * ld + j] = max(mine, theirs);
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 4: mobility -------------------------------------------------------------------------
    const double x_new = act ? mobility_step(p, x, v, u) : 0.0;
    if (act && p.mobility) p.pos_x[vbase + u] = x_new;

    // ---- phase 5: stream the columns (thread u = observer again) -----------------------------------
    int m_cnt = 0;
    if (vpd) { for (int k = 0; k < B; ++k) hist[k * T + u] = 0u; }
    if (p.piggy) {
        int32_t *seqp = p.tab_seq + tbase, *lup = p.tab_lu + tbase;
        double *xp = p.tab_x + tbase;
        for (int j = 0; j < N; ++j) {
            int sn = 0, lu = 0; double xn = 0.0;
            if (act) {
                int s0 = seqp[j * N + u];
                lu = lup[j * N + u];
                xn = xp[j * N + u];
                if (j == u) { s0 += 1; lu = 0; xn = x; } else lu += 1;     // vehicle.py:58-70
                const unsigned key = K[u * ld + j];
                sn = (int)(key >> SB);
                if (sn != s0) {                                          // vehicle.py:41-47
                    const int src = (int)(key & srcmask);
                    xn = (src == j) ? sx[j] : xp[j * N + src];
                    lu = 0;
                }
            }
            __syncthreads();          // every read of column j precedes every write of it
            if (act) {
                seqp[j * N + u] = sn; lup[j * N + u] = lu; xp[j * N + u] = xn;
                if (vpd && j != u && lu < p.age_threshold) {              // network.py:547
                    const double y1 = sn > 0 ? sy[j] : 0.0;
                    const double d = dist2d(xn, y1, x_new, y);
                    if (d < p.W) {                                        // network.py:487
                        const double s = (__dsub_rn(xn, x_new) > 0.0) ? d : -d;
                        hist[vpd_bin(s, p.W, p.inv_binw, B, s_edges) * T + u] += 1u;
                        ++m_cnt;
                    }
                }
            }
        }
    }

    // ---- phase 6: state rows (TestEnv.obtain_state, test_env.py:527-583) ---------------------------
    if (want_state && act) {
        float *row = p.state + (vbase + u) * p.S;
        const float *og = p.obs + (vbase + u) * R;
        int k = 0;
        if (p.add_action) {
            if (p.action_binary) { for (int r = 0; r < R; ++r) row[k++] = (a == r) ? 1.0f : 0.0f; }
            else row[k++] = (float)a;
        }
        if (p.add_channel_obs) { for (int r = 0; r < R; ++r) row[k++] = og[r]; }
        if (p.piggy) {
            const float den = (float)m_cnt;
            for (int b = 0; b < B; ++b)
                row[k++] = (vpd && m_cnt > 0) ? __fdiv_rn((float)hist[b * T + u], den) : 0.0f;
        }
        if (p.add_reward) row[k++] = (float)rew;
        if (p.add_index) row[k++] = (float)(u + 1);
        if (p.add_position) { row[k++] = (float)__ddiv_rn(x_new, p.L); row[k++] = (float)__ddiv_rn(y, 2.0); }
        if (p.add_velocity) row[k++] = (float)v;
        if (p.fingerprint) { row[k++] = (float)p.episode; row[k++] = (float)p.epsilon; }
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
        const int w = u >> 5, nw = T >> 5;
        if ((u & 31) == 0) { s_red[w * 4 + 0] = rs; s_red[w * 4 + 1] = nr; s_red[w * 4 + 2] = np; s_red[w * 4 + 3] = nb; }
        __syncthreads();
        if (u == 0) {
            double trs = 0.0, tnr = 0.0, tnp = 0.0, tnb = 0.0;
            for (int i = 0; i < nw; ++i) { trs += s_red[i * 4]; tnr += s_red[i * 4 + 1]; tnp += s_red[i * 4 + 2]; tnb += s_red[i * 4 + 3]; }
            p.acc_reward[e] += trs;
            long long *c = p.acc_count + e * ACC_COUNTS;
            c[0] += (long long)tnr; c[1] += (long long)tnp; c[2] += (long long)tnb; c[3] += 1;
        }
    }
}

int block_threads(int N) { return ((N + 31) / 32) * 32; }

}  // namespace

int key_src_bits(int N)
{
    int b = 1;
    while ((1 << b) < N) ++b;
    return b;
}

bool step_block_keys_fit_smem(const Params &p)
{
    const BlockSmem lay(p.N, p.R, p.B, block_threads(p.N), p.vpd_enabled != 0, true);
    return lay.bytes <= SMEM_BUDGET;
}

size_t step_block_smem_bytes(const Params &p, bool keys_in_smem)
{
    const BlockSmem lay(p.N, p.R, p.B, block_threads(p.N), p.build_state != 0 && p.vpd_enabled != 0, keys_in_smem);
    return lay.bytes;
}

size_t step_block_scratch_bytes(long long E, int N)
{
    return (size_t)E * N * (N + 1) * sizeof(unsigned);
}

cudaError_t prepare_step_block(const Params &p)
{
    Params q = p; q.build_state = 1;
    const size_t smem = step_block_smem_bytes(q, step_block_keys_fit_smem(q));
    if (smem <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(step_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

cudaError_t launch_step_block(const Params &p, cudaStream_t stream)
{
    const bool fit = step_block_keys_fit_smem(p);
    const size_t smem = step_block_smem_bytes(p, fit);
    step_block_kernel<<<(unsigned)p.E, block_threads(p.N), smem, stream>>>(p, key_src_bits(p.N), fit ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace diral
