// diral_host.cpp -- assembling TestEnv.obtain_state rows (reference envs/test_env.py:527-583) on the host from the
// compact per-agent record diral_step_host moves over PCIe.  See diral_host.h.
//
// The expander writes S float32 per agent (21.5 MB per slot at the headline configuration).  Every worker owns the
// same contiguous row range in every call, so when a worker's share fits its private L2 the rows are written with
// ordinary stores and stay cache-resident for the consumer (no DRAM round trip at all when the caller reuses its
// buffer); larger outputs leave as non-temporal 16-byte stores (no read-for-ownership of the destination lines).
#include "diral_host.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <immintrin.h>
#include <mutex>
#include <thread>
#include <vector>

namespace diral {

namespace {

// counts / len of the positional distribution (network.py:501): float32 rounding of the float64 quotient, exactly
// what the kernels store.  One table for every observer (len <= 255 samples at <= 256 vehicles).
const float *vpd_quotients()
{
    static std::vector<float> lut;
    static std::once_flag once;
    std::call_once(once, [] {
        lut.assign(256 * 256, 0.0f);
        for (int m = 1; m < 256; ++m)
            for (int c = 0; c < 256; ++c) lut[m * 256 + c] = (float)((double)c / (double)m);
    });
    return lut.data();
}

inline float rew_at(const HostJob &job, long long a)
{
    float r;
    std::memcpy(&r, reinterpret_cast<const char *>(job.rews) + a * job.rew_stride, sizeof r);
    return r;
}

void stream_out(float *dst, const float *src, long long n, bool nt)
{
    // dst is 16-byte aligned whenever the caller's buffer is and the row range starts at a multiple of 4 agents
    if (nt && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (n & 3) == 0) {
        for (long long i = 0; i < n; i += 4)
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), _mm_load_si128(reinterpret_cast<const __m128i *>(src + i)));
    } else {
        std::memcpy(dst, src, (size_t)n * sizeof(float));
    }
}

}  // namespace

namespace {

// Rows whose blocks are all whole 16-byte groups (one-hot over R % 4 == 0 resources, obs likewise, B % 4 == 0 bins,
// no scalar tail -- the shipped State block, S = R + B) are formed in registers and leave as non-temporal stores
// straight from them: per agent R/4 compares for the one-hot and B/4 divisions (IEEE float32 division of two small
// integers is the float32 rounding of the quotient, which is what the kernels store).
bool vector_rows_ok(const HostLayout &lay, const HostJob &job)
{
    if (lay.add_reward || lay.add_index || lay.add_position || lay.add_velocity || lay.fingerprint) return false;
    if (lay.add_action && (!lay.action_binary || (lay.R & 3))) return false;
    if (lay.add_channel_obs && (lay.R & 3)) return false;
    if (lay.piggy && (lay.B & 3)) return false;
    return (lay.S & 3) == 0 && (reinterpret_cast<uintptr_t>(job.out) & 15) == 0;
}

template <bool NT>
inline void put4(float *w, __m128i v)
{
    if (NT) _mm_stream_si128(reinterpret_cast<__m128i *>(w), v);
    else _mm_store_si128(reinterpret_cast<__m128i *>(w), v);
}

// SSE2 flavour (any x86-64): IEEE division per group of four bins
template <bool NT>
inline void expand_rows_vector_impl(const HostLayout &lay, const HostJob &job, long long a0, long long a1)
{
    const int R = lay.R, B = lay.B, S = lay.S;
    const __m128i lane_id = _mm_set_epi32(3, 2, 1, 0), four = _mm_set1_epi32(4), zero = _mm_setzero_si128();
    const __m128i one_bits = _mm_castps_si128(_mm_set1_ps(1.0f));
    for (long long a = a0; a < a1; ++a) {
        float *w = job.out + a * S;
        if (job.rews_out) job.rews_out[a] = rew_at(job, a);
        if (lay.add_action) {
            int act = job.actions[a];
            act = act < 0 ? 0 : (act >= R ? R - 1 : act);
            const __m128i av = _mm_set1_epi32(act);
            __m128i idx = lane_id;
            for (int r = 0; r < R; r += 4, w += 4) {
                put4<NT>(w, _mm_and_si128(_mm_cmpeq_epi32(idx, av), one_bits));
                idx = _mm_add_epi32(idx, four);
            }
        }
        if (lay.add_channel_obs) {
            const float *o = job.obs + a * R;
            for (int r = 0; r < R; r += 4, w += 4) put4<NT>(w, _mm_castps_si128(_mm_loadu_ps(o + r)));
        }
        if (lay.piggy) {
            const uint8_t *c = job.counts + a * job.count_stride;
            __m128i acc = zero;
            __m128i c32[64];                     // B <= 256 bins
            for (int b = 0; b < B; b += 4) {
                int word;
                std::memcpy(&word, c + b, 4);
                const __m128i v = _mm_unpacklo_epi16(_mm_unpacklo_epi8(_mm_cvtsi32_si128(word), zero), zero);
                c32[b >> 2] = v;
                acc = _mm_add_epi32(acc, v);
            }
            acc = _mm_add_epi32(acc, _mm_shuffle_epi32(acc, 0x4e));
            acc = _mm_add_epi32(acc, _mm_shuffle_epi32(acc, 0xb1));       // every lane = len(s)
            // len == 0: every count is 0 too; dividing by 1 leaves the all-zero vector (network.py:502-505)
            const __m128 den = _mm_cvtepi32_ps(_mm_max_epi16(acc, _mm_set1_epi32(1)));
            for (int b = 0; b < B; b += 4, w += 4) put4<NT>(w, _mm_castps_si128(_mm_div_ps(_mm_cvtepi32_ps(c32[b >> 2]), den)));
        }
    }
    if (NT) _mm_sfence();
}


// AVX2 + FMA flavour of the same row assembly (chosen at run time): counts / len without a division per bin -- one
// correctly rounded reciprocal per agent, then  q0 = c * rcp;  q = fma(fma(-q0, len, c), rcp, q0),  which equals the
// correctly rounded c / len for all 0 <= c <= len < 1024 (the lemma the kernels use, checked exhaustively in
// tests/test_host.py).  Eight floats per step (one 4-float step closes a block whose length is
// 4 mod 8), the sample count from one SAD over the count bytes, bytes widened with one VPMOVZXBD.
// a block may start 16 (not 32) bytes into a row: 32-byte non-temporal stores need 32-byte addresses
#define DIRAL_TARGET_FMA __attribute__((target("avx2,fma")))

template <bool NT>
DIRAL_TARGET_FMA inline void store8_avx(float *d, __m256 v, bool rows32)
{
    if (!NT) _mm256_storeu_ps(d, v);
    else if (rows32 && (reinterpret_cast<uintptr_t>(d) & 31) == 0) _mm256_stream_ps(d, v);
    else { _mm_stream_ps(d, _mm256_castps256_ps128(v)); _mm_stream_ps(d + 4, _mm256_extractf128_ps(v, 1)); }
}

template <bool NT>
DIRAL_TARGET_FMA void expand_rows_avx2(const HostLayout &lay, const HostJob &job, long long a0, long long a1)
{
    const int R = lay.R, B = lay.B, S = lay.S;
    const __m256i lane8 = _mm256_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7), eight = _mm256_set1_epi32(8);
    const __m256i one8 = _mm256_castps_si256(_mm256_set1_ps(1.0f));
    const __m128i one4 = _mm_castps_si128(_mm_set1_ps(1.0f));
    const bool rows32 = ((reinterpret_cast<uintptr_t>(job.out) | (uintptr_t)(4 * S)) & 31) == 0;   // every row 32-byte aligned
    for (long long a = a0; a < a1; ++a) {
        float *w = job.out + a * S;
        if (job.rews_out) job.rews_out[a] = rew_at(job, a);
#define store8(d, v) store8_avx<NT>((d), (v), rows32)
        if (lay.add_action) {
            int act = job.actions[a];
            act = act < 0 ? 0 : (act >= R ? R - 1 : act);
            const __m256i av = _mm256_set1_epi32(act);
            __m256i idx = lane8;
            int r = 0;
            for (; r + 8 <= R; r += 8, w += 8) {
                store8(w, _mm256_castsi256_ps(_mm256_and_si256(_mm256_cmpeq_epi32(idx, av), one8)));
                idx = _mm256_add_epi32(idx, eight);
            }
            if (r < R) { put4<NT>(w, _mm_and_si128(_mm_cmpeq_epi32(_mm256_castsi256_si128(idx), _mm256_castsi256_si128(av)), one4)); w += 4; }
        }
        if (lay.add_channel_obs) {
            const float *o = job.obs + a * R;
            int r = 0;
            for (; r + 8 <= R; r += 8, w += 8) store8(w, _mm256_loadu_ps(o + r));
            if (r < R) { put4<NT>(w, _mm_castps_si128(_mm_loadu_ps(o + r))); w += 4; }
        }
        if (lay.piggy) {
            const uint8_t *c = job.counts + a * job.count_stride;
            // len(s): sum of the B count bytes (SAD against zero, 16 / 8 / 4 bytes at a time)
            __m128i acc = _mm_setzero_si128();
            int b = 0;
            for (; b + 16 <= B; b += 16) acc = _mm_add_epi64(acc, _mm_sad_epu8(_mm_loadu_si128(reinterpret_cast<const __m128i *>(c + b)), _mm_setzero_si128()));
            for (; b + 8 <= B; b += 8) acc = _mm_add_epi64(acc, _mm_sad_epu8(_mm_loadl_epi64(reinterpret_cast<const __m128i *>(c + b)), _mm_setzero_si128()));
            for (; b < B; b += 4) { int word; std::memcpy(&word, c + b, 4); acc = _mm_add_epi64(acc, _mm_sad_epu8(_mm_cvtsi32_si128(word), _mm_setzero_si128())); }
            int m = _mm_cvtsi128_si32(acc) + _mm_extract_epi16(acc, 4);
            // len == 0: every count is 0 too; dividing by 1 leaves the all-zero vector (network.py:502-505)
            const float denf = (float)(m > 0 ? m : 1);
            const __m256 den = _mm256_set1_ps(denf), rcp = _mm256_set1_ps(1.0f / denf);
            for (b = 0; b + 8 <= B; b += 8, w += 8) {
                const __m256 cf = _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i *>(c + b))));
                const __m256 q0 = _mm256_mul_ps(cf, rcp);
                store8(w, _mm256_fmadd_ps(_mm256_fnmadd_ps(q0, den, cf), rcp, q0));
            }
            if (b < B) {
                int word; std::memcpy(&word, c + b, 4);
                const __m128 cf = _mm_cvtepi32_ps(_mm_cvtepu8_epi32(_mm_cvtsi32_si128(word)));
                const __m128 q0 = _mm_mul_ps(cf, _mm256_castps256_ps128(rcp));
                put4<NT>(w, _mm_castps_si128(_mm_fmadd_ps(_mm_fnmadd_ps(q0, _mm256_castps256_ps128(den), cf), _mm256_castps256_ps128(rcp), q0)));
                w += 4;
            }
        }
    }
    if (NT) _mm_sfence();
#undef store8
}

// Four agents per step (AVX-512 BW/VL/VBMI).  4 S floats are a whole number of 64-byte vectors when S % 4 == 0, so a
// group of four rows leaves as S / 4 aligned full-width stores.  What each output lane holds -- a one-hot lane, or the
// quotient of agent j's bin b -- depends only on the layout, so it is tabulated once per layout: per output vector a byte
// permute that drops every quotient lane's count byte (out of the group's 128-byte record block) into that lane's
// dword and zeroes the one-hot lanes, and the agent of every lane (to spread the four lengths / reciprocals).  The
// one-hot lanes leave as zeros; one scalar store per agent then sets the lane of its action.  Rewards embedded in the
// records come out with one more byte permute.
struct Rows512Plan {
    int Ra = -1, B = 0, S = 0;              // one-hot width (0: no action block), bins, row length
    long long stride = 0;                   // bytes between the count records of consecutive agents
    int nvec = 0, nsum = 0;
    alignas(64) uint8_t byte_idx[32][64];   // per output vector: source byte (in the 128-byte record block) of every lane, in
                                            // the low byte of the lane's dword
    unsigned long long bmask[32];           // ... and which of the 64 bytes are taken at all (the rest become 0)
    alignas(64) int32_t agent[32][16];      // per output vector: agent (0..3) of every lane
    alignas(64) int8_t weight[128];         // 1 on count bytes, 0 elsewhere (rewards, padding)
    alignas(64) int32_t sum_idx[8][16];     // t-th group of four count bytes of each agent (dword index in the block)
    unsigned long long hi_mask = 0;         // bytes of the record block beyond the first 64
    long long rew_off = -1;                 // rewards embedded in the records at this byte offset (-1: a separate array)
    alignas(64) uint8_t rew_idx[64];        // ... and the 16 bytes of the four agents' rewards inside the record block
};

bool rows512_layout_ok(const HostLayout &lay, const HostJob &job)
{
    if (lay.add_channel_obs || (lay.add_action && !lay.action_binary)) return false;
    const long long stride = lay.piggy ? job.count_stride : 0;
    return lay.S <= 128 && (lay.S & 3) == 0 && 4 * stride <= 128 && (stride & 3) == 0 && (!lay.piggy || lay.B <= 32)
           && (reinterpret_cast<uintptr_t>(job.out) & 63) == 0;
}

const Rows512Plan &rows512_plan(const HostLayout &lay, const HostJob &job)
{
    static thread_local Rows512Plan plan;
    const int Ra = lay.add_action ? lay.R : 0, B = lay.piggy ? lay.B : 0;
    const long long stride = lay.piggy ? job.count_stride : 0;
    const long long roff = reinterpret_cast<const uint8_t *>(job.rews) - job.counts;
    const long long rew_off = (lay.piggy && job.rews_out && job.rew_stride == stride && roff >= 0 && roff + 4 <= stride) ? roff : -1;
    if (plan.Ra == Ra && plan.B == B && plan.S == lay.S && plan.stride == stride && plan.rew_off == rew_off) return plan;
    plan.Ra = Ra; plan.B = B; plan.S = lay.S; plan.stride = stride; plan.rew_off = rew_off;
    memset(plan.rew_idx, 0, sizeof plan.rew_idx);
    if (rew_off >= 0) for (int i = 0; i < 16; ++i) plan.rew_idx[i] = (uint8_t)((i >> 2) * stride + rew_off + (i & 3));
    plan.nvec = lay.S / 4;
    memset(plan.byte_idx, 0, sizeof plan.byte_idx);
    for (int v = 0; v < plan.nvec; ++v) {
        plan.bmask[v] = 0ull;
        for (int l = 0; l < 16; ++l) {
            const int f = 16 * v + l, j = f / lay.S, sI = f % lay.S;
            plan.agent[v][l] = j;
            if (sI >= Ra) {
                plan.byte_idx[v][4 * l] = (uint8_t)(j * stride + (sI - Ra));
                plan.bmask[v] |= 1ull << (4 * l);
            }
        }
    }
    for (int i = 0; i < 128; ++i) plan.weight[i] = (stride > 0 && i < 4 * stride && (i % stride) < B) ? 1 : 0;
    plan.nsum = (B + 3) / 4;
    for (int t = 0; t < plan.nsum; ++t)
        for (int l = 0; l < 16; ++l) plan.sum_idx[t][l] = (int32_t)((l & 3) * (stride / 4) + t);
    const long long hi = 4 * stride - 64;
    plan.hi_mask = hi <= 0 ? 0ull : (hi >= 64 ? ~0ull : ((1ull << hi) - 1ull));
    return plan;
}

#define DIRAL_TARGET_512V __attribute__((target("avx512f,avx512bw,avx512vl,avx512dq,avx512vbmi,fma")))

// rows [a0, a1): a0 a multiple of 4, (a1 - a0) a multiple of 4, job.out 64-byte aligned
DIRAL_TARGET_512V void expand_rows_avx512x4(const Rows512Plan &pl, const HostJob &job, long long a0, long long a1)
{
    const __m512 ones = _mm512_set1_ps(1.0f);
    const __m512i w_lo = _mm512_load_si512(pl.weight), w_hi = _mm512_load_si512(pl.weight + 64), ones16 = _mm512_set1_epi16(1);
    const __m128i rmax = _mm_set1_epi32(pl.Ra - 1), zero4 = _mm_setzero_si128();
    const __m512i rew_idx = _mm512_load_si512(pl.rew_idx);
    const bool have_lo = pl.stride > 0, have_hi = pl.hi_mask != 0;
    for (long long a = a0; a < a1; a += 4) {
        float *w = job.out + a * pl.S;
        if (job.rews_out && pl.rew_off < 0) for (int j = 0; j < 4; ++j) job.rews_out[a + j] = rew_at(job, a + j);
        __m512i blk_lo = _mm512_setzero_si512(), blk_hi = _mm512_setzero_si512();
        __m512 den16 = ones, rcp16 = ones;
        if (have_lo) {
            const uint8_t *c = job.counts + a * pl.stride;
            blk_lo = have_hi ? _mm512_loadu_si512(c) : _mm512_maskz_loadu_epi8(4 * pl.stride >= 64 ? ~0ull : ((1ull << (4 * pl.stride)) - 1ull), c);
            if (have_hi) blk_hi = _mm512_maskz_loadu_epi8(pl.hi_mask, c + 64);
            if (pl.rew_off >= 0)           // the four rewards sit in the same block: one byte permute, one 16-byte store
                _mm_storeu_si128(reinterpret_cast<__m128i *>(job.rews_out + a),
                                 _mm512_castsi512_si128(_mm512_permutex2var_epi8(blk_lo, rew_idx, blk_hi)));
            // len(s) of the four agents: bytes -> sums of 4 (weights drop rewards / padding) -> one dword per agent
            const __m512i d_lo = _mm512_madd_epi16(_mm512_maddubs_epi16(blk_lo, w_lo), ones16);
            const __m512i d_hi = _mm512_madd_epi16(_mm512_maddubs_epi16(blk_hi, w_hi), ones16);
            __m128i m4 = zero4;
            for (int t = 0; t < pl.nsum; ++t)
                m4 = _mm_add_epi32(m4, _mm512_castsi512_si128(_mm512_permutex2var_epi32(d_lo, _mm512_load_si512(pl.sum_idx[t]), d_hi)));
            // len == 0: every count is 0 too; dividing by 1 leaves the all-zero vector (network.py:502-505)
            const __m128 den4 = _mm_cvtepi32_ps(_mm_max_epi32(m4, _mm_set1_epi32(1)));
            den16 = _mm512_broadcast_f32x4(den4);
            rcp16 = _mm512_broadcast_f32x4(_mm_div_ps(_mm_set1_ps(1.0f), den4));
        }
        alignas(16) int32_t act_a[4] = {0, 0, 0, 0};
        if (pl.Ra > 0)
            _mm_store_si128(reinterpret_cast<__m128i *>(act_a),
                            _mm_min_epi32(_mm_max_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i *>(job.actions + a)), zero4), rmax));
        for (int v = 0; v < pl.nvec; ++v) {
            __m512 q = _mm512_setzero_ps();
            if (pl.bmask[v]) {
                // one byte permute drops every lane's count into the low byte of its dword (zero elsewhere)
                const __m512 cf = _mm512_cvtepi32_ps(_mm512_maskz_permutex2var_epi8(pl.bmask[v], blk_lo, _mm512_load_si512(pl.byte_idx[v]), blk_hi));
                const __m512i ag = _mm512_load_si512(pl.agent[v]);
                const __m512 den = _mm512_permutexvar_ps(ag, den16), rcp = _mm512_permutexvar_ps(ag, rcp16);
                const __m512 q0 = _mm512_mul_ps(cf, rcp);
                q = _mm512_fmadd_ps(_mm512_fnmadd_ps(q0, den, cf), rcp, q0);
            }
            _mm512_store_ps(w + 16 * v, q);
        }
        // the one-hot lanes were written as zeros: one scalar store per agent sets its action's lane
        if (pl.Ra > 0)
            for (int j = 0; j < 4; ++j) w[j * pl.S + act_a[j]] = 1.0f;
    }
}

bool cpu_has_avx512vbmi()
{
    static const bool yes = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl")
                            && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vbmi") && __builtin_cpu_supports("fma")
                            && !getenv("DIRAL_HOST_NO_AVX512") && !getenv("DIRAL_HOST_NO_FMA");
    return yes;
}

void rows_sse_nt(const HostLayout &l, const HostJob &j, long long a0, long long a1) { expand_rows_vector_impl<true>(l, j, a0, a1); }
void rows_sse_st(const HostLayout &l, const HostJob &j, long long a0, long long a1) { expand_rows_vector_impl<false>(l, j, a0, a1); }

bool cpu_has_fma()
{
    static const bool yes = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma") && !getenv("DIRAL_HOST_NO_FMA");
    return yes;
}

}  // namespace

void expand_rows(const HostLayout &lay, const HostJob &job, long long a0, long long a1)
{
    if (vector_rows_ok(lay, job)) {
        if (!lay.nt_stores && (a0 & 3) == 0 && a1 - a0 >= 4 && cpu_has_avx512vbmi() && rows512_layout_ok(lay, job)) {
            const long long mid = a0 + ((a1 - a0) & ~3ll);
            expand_rows_avx512x4(rows512_plan(lay, job), job, a0, mid);
            a0 = mid;
            if (a0 == a1) return;
        }
        if (cpu_has_fma()) { if (lay.nt_stores) expand_rows_avx2<true>(lay, job, a0, a1); else expand_rows_avx2<false>(lay, job, a0, a1); }
        else { if (lay.nt_stores) rows_sse_nt(lay, job, a0, a1); else rows_sse_st(lay, job, a0, a1); }
        return;
    }
    const int N = lay.N, R = lay.R, B = lay.B, S = lay.S;
    const float *lut = vpd_quotients();
    // staging block: a multiple of 4 rows (so every flush is a whole number of 16-byte pieces), about 16 KB
    long long rows_per_block = (16384 / (4 * (long long)S)) & ~3ll;
    if (rows_per_block < 4) rows_per_block = 4;
    alignas(64) static thread_local float stage_small[4096 + 16];
    std::vector<float> stage_big;
    float *stage = stage_small;
    if (rows_per_block * S > 4096) { stage_big.resize((size_t)(rows_per_block * S) + 16); stage = stage_big.data(); }
    stage = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(stage) + 63) & ~(uintptr_t)63);

    for (long long b0 = a0; b0 < a1; b0 += rows_per_block) {
        const long long b1 = b0 + rows_per_block < a1 ? b0 + rows_per_block : a1;
        float *w = stage;
        for (long long a = b0; a < b1; ++a) {
            if (job.rews_out) job.rews_out[a] = rew_at(job, a);
            int act = job.actions[a];
            act = act < 0 ? 0 : (act >= R ? R - 1 : act);                     // the kernels clamp (and count) bad actions
            if (lay.add_action) {                                             // test_env.py:539-545
                if (lay.action_binary) { std::memset(w, 0, sizeof(float) * (size_t)R); w[act] = 1.0f; w += R; }
                else *w++ = (float)act;
            }
            if (lay.add_channel_obs) { std::memcpy(w, job.obs + a * R, sizeof(float) * (size_t)R); w += R; }   // :547-548
            if (lay.piggy) {                                                  // :554-562, network.py:495-505
                const uint8_t *c = job.counts + a * job.count_stride;
                int m = 0;
                for (int b = 0; b < B; ++b) m += c[b];
                const float *q = lut + (m > 255 ? 255 : m) * 256;
                for (int b = 0; b < B; ++b) w[b] = q[c[b]];
                w += B;
            }
            if (lay.add_reward) *w++ = rew_at(job, a);                           // :568-570
            if (lay.add_index) *w++ = (float)(a % N + 1);                     // :571-572
            if (lay.add_position) {                                           // :573-574, network.py:403-407
                *w++ = (float)(job.pos_x[a] / lay.L);
                *w++ = (float)(job.pos_y[a] / 2.0);
            }
            if (lay.add_velocity) *w++ = (float)job.vel[a];                   // :575-576
            if (lay.fingerprint) { *w++ = (float)job.episode; *w++ = (float)job.epsilon; }   // :577-579
        }
        stream_out(job.out + b0 * S, stage, (b1 - b0) * S, lay.nt_stores != 0);
    }
    _mm_sfence();
}

struct HostPool::Impl {
    static constexpr int RING = 8;       // jobs that can be queued at once
    struct Slot {
        std::atomic<unsigned long long> id{0};   // the job in this slot
        HostLayout lay{};
        HostJob job{};
        long long bounds[HostPool::MAX_CHUNKS + 1] = {};
        int nchunks = 0;
        std::atomic<int> ready{0};           // chunks whose inputs are in host memory ...
        const volatile unsigned *flags = nullptr;   // ... or per-chunk words raised by the device (== epoch: chunk landed)
        unsigned epoch = 0;
        std::atomic<bool> aborted{false};
        std::atomic<int> done{0};            // workers that have finished the job
        std::chrono::steady_clock::time_point t_begin{};
        double t_us[3] = {0.0, 0.0, 0.0};    // worker 0: first chunk released, last chunk released, last row written
    };
    std::vector<std::thread> workers;
    std::mutex mu, post_mu;
    std::condition_variable cv;
    std::atomic<unsigned long long> posted{0};   // jobs queued so far (ids 1 .. posted)
    bool stop = false;                   // guarded by mu
    Slot ring[RING];
    int hot_us = 200;                    // how long a worker keeps polling for the next job before it sleeps

    Slot &slot(unsigned long long id) { return ring[id % RING]; }

    void run(int idx, int nthreads)
    {
        unsigned long long seen = 0;
        for (;;) {
            // A caller stepping in a loop comes back within tens of microseconds: stay hot for a while (no futex wake,
            // no scheduler latency on the next call), then go to sleep on the condition variable.
            bool have = false;
            const auto t0 = std::chrono::steady_clock::now();
            for (int spins = 0; ; ++spins) {
                if (posted.load(std::memory_order_acquire) > seen) { have = true; break; }
                _mm_pause();
                if ((spins & 255) == 255 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(hot_us)) break;
            }
            if (!have) {
                std::unique_lock<std::mutex> lock(mu);
                cv.wait(lock, [&] { return stop || posted.load(std::memory_order_acquire) > seen; });
                if (stop) return;
            }
            ++seen;
            Slot &s = slot(seen);
            const volatile unsigned *flags = s.flags;
            for (int c = 0; c < s.nchunks; ++c) {
                int spins = 0;
                bool gone = false;
                while (flags ? flags[c] != s.epoch : s.ready.load(std::memory_order_acquire) <= c) {
                    if (s.aborted.load(std::memory_order_relaxed)) { gone = true; break; }
                    _mm_pause();
                    if (++spins > 20000) { std::this_thread::yield(); spins = 0; }
                }
                if (gone) break;
                std::atomic_thread_fence(std::memory_order_acquire);
                if (idx == 0 && (c == 0 || c == s.nchunks - 1))
                    s.t_us[c == 0 ? 0 : 1] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - s.t_begin).count();
                if (idx == 0 && s.nchunks == 1) s.t_us[1] = s.t_us[0];
                // this worker's share of chunk c, cut at multiples of 4 agents (16-byte aligned row ranges)
                const long long lo = s.bounds[c], n = s.bounds[c + 1] - lo;
                long long s0 = (n * idx / nthreads) & ~3ll, s1 = (n * (idx + 1) / nthreads) & ~3ll;
                if (idx == nthreads - 1) s1 = n;
                if (s1 > s0) expand_rows(s.lay, s.job, lo + s0, lo + s1);
            }
            if (idx == 0) s.t_us[2] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - s.t_begin).count();
            s.done.fetch_add(1, std::memory_order_release);
        }
    }
};

HostPool::HostPool(int threads) : impl(new Impl)
{
    if (const char *v = getenv("DIRAL_HOST_HOT_US")) impl->hot_us = atoi(v);       // measurement knob
    const int n = threads < 1 ? 1 : threads;
    vpd_quotients();
    for (int i = 0; i < n; ++i) impl->workers.emplace_back([this, i, n] { impl->run(i, n); });
}

HostPool::~HostPool()
{
    {
        std::lock_guard<std::mutex> lock(impl->mu);
        impl->stop = true;
    }
    impl->cv.notify_all();
    for (auto &t : impl->workers) t.join();
    delete impl;
}

int HostPool::threads() const { return (int)impl->workers.size(); }

unsigned long long HostPool::begin(const HostLayout &lay, const HostJob &job, const long long *bounds, int nchunks,
                                   const volatile unsigned *flags, unsigned epoch)
{
    std::lock_guard<std::mutex> post(impl->post_mu);         // callers on different threads queue one at a time
    const unsigned long long id = impl->posted.load(std::memory_order_relaxed) + 1;
    Impl::Slot &s = impl->slot(id);
    if (id > Impl::RING) finish(id - Impl::RING);            // the slot's previous job (long done in any sane use)
    s.lay = lay; s.job = job;
    s.nchunks = nchunks < MAX_CHUNKS ? nchunks : MAX_CHUNKS;
    for (int c = 0; c <= s.nchunks; ++c) s.bounds[c] = bounds[c];
    s.flags = flags; s.epoch = epoch;
    s.ready.store(0, std::memory_order_relaxed);
    s.done.store(0, std::memory_order_relaxed);
    s.aborted.store(false, std::memory_order_relaxed);
    s.id.store(id, std::memory_order_relaxed);
    s.t_begin = std::chrono::steady_clock::now();
    {
        std::lock_guard<std::mutex> lock(impl->mu);          // (a worker between its spin phase and cv.wait sees the new value)
        impl->posted.store(id, std::memory_order_release);
    }
    impl->cv.notify_all();
    return id;
}

void HostPool::publish(unsigned long long id, int chunk)
{
    Impl::Slot &s = impl->slot(id);
    if (s.id.load(std::memory_order_relaxed) == id) s.ready.store(chunk + 1, std::memory_order_release);
}

bool HostPool::done(unsigned long long id) const
{
    const Impl::Slot &s = impl->slot(id);
    return s.id.load(std::memory_order_acquire) != id || s.done.load(std::memory_order_acquire) >= (int)impl->workers.size();
}

void HostPool::abort(unsigned long long id)
{
    Impl::Slot &s = impl->slot(id);
    if (s.id.load(std::memory_order_relaxed) == id) s.aborted.store(true, std::memory_order_release);
}

void HostPool::timeline(unsigned long long id, double out_us[3]) const
{
    const Impl::Slot &s = impl->slot(id);
    for (int i = 0; i < 3; ++i) out_us[i] = s.id.load(std::memory_order_acquire) == id ? s.t_us[i] : 0.0;
}

void HostPool::finish(unsigned long long id)
{
    int spins = 0;
    while (!done(id)) {
        _mm_pause();
        if (++spins > 20000) { std::this_thread::yield(); spins = 0; }
    }
}

}  // namespace diral
