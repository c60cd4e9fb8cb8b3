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

#define DIRAL_TARGET_FMA_FWD __attribute__((target("avx2,fma")))

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

DIRAL_TARGET_FMA_FWD inline __m128 fma_quotient(__m128 c, __m128 den, __m128 rcp, __m128 q0)
{
    return _mm_fmadd_ps(_mm_fnmadd_ps(q0, den, c), rcp, q0);
}

template <bool NT>
inline void put4(float *w, __m128i v)
{
    if (NT) _mm_stream_si128(reinterpret_cast<__m128i *>(w), v);
    else _mm_store_si128(reinterpret_cast<__m128i *>(w), v);
}

// counts / len without a division per bin: one correctly rounded reciprocal per agent, then per group of four bins
//   q0 = c * rcp;  q = fma(fma(-q0, len, c), rcp, q0)
// which equals the correctly rounded c / len for all 0 <= c <= len < 1024 (the lemma the kernels use, checked
// exhaustively in tests/test_host.py).  Needs fused multiply-add: compiled for AVX2+FMA, chosen at run time.
#define DIRAL_TARGET_FMA __attribute__((target("avx2,fma")))

template <bool NT, bool FMA>
#if defined(__GNUC__)
__attribute__((always_inline))
#endif
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
            if (FMA) {
                const __m128 rcp = _mm_div_ps(_mm_set1_ps(1.0f), den);
                for (int b = 0; b < B; b += 4, w += 4) {
                    const __m128 c = _mm_cvtepi32_ps(c32[b >> 2]);
                    const __m128 q0 = _mm_mul_ps(c, rcp);
                    put4<NT>(w, _mm_castps_si128(fma_quotient(c, den, rcp, q0)));
                }
            } else {
                for (int b = 0; b < B; b += 4, w += 4) put4<NT>(w, _mm_castps_si128(_mm_div_ps(_mm_cvtepi32_ps(c32[b >> 2]), den)));
            }
        }
    }
    if (NT) _mm_sfence();
}


// AVX2 + FMA flavour of the same row assembly: eight floats per step (one 4-float step closes a block whose length is
// 4 mod 8), the sample count from one SAD over the count bytes, bytes widened with one VPMOVZXBD.
// a block may start 16 (not 32) bytes into a row: 32-byte non-temporal stores need 32-byte addresses
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

void rows_sse_nt(const HostLayout &l, const HostJob &j, long long a0, long long a1) { expand_rows_vector_impl<true, false>(l, j, a0, a1); }
void rows_sse_st(const HostLayout &l, const HostJob &j, long long a0, long long a1) { expand_rows_vector_impl<false, false>(l, j, a0, a1); }

bool cpu_has_fma()
{
    static const bool yes = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma") && !getenv("DIRAL_HOST_NO_FMA");
    return yes;
}

}  // namespace

void expand_rows(const HostLayout &lay, const HostJob &job, long long a0, long long a1)
{
    if (vector_rows_ok(lay, job)) {
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
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<unsigned long long> job_seq{0};
    bool stop = false;                   // guarded by mu
    // the running job (written by begin() before job_seq moves)
    HostLayout lay{};
    HostJob job{};
    std::vector<long long> bounds;
    int nchunks = 0;
    std::atomic<int> ready{0};           // chunks whose inputs are in host memory
    std::atomic<int> done{0};            // workers that have finished the job
    int hot_us = 200;                    // how long a worker keeps polling for the next job before it sleeps

    void run(int idx, int nthreads)
    {
        unsigned long long seen = 0;
        for (;;) {
            // A caller stepping in a loop comes back within tens of microseconds: stay hot for a while (no futex wake,
            // no scheduler latency on the next call), then go to sleep on the condition variable.
            bool have = false;
            const auto t0 = std::chrono::steady_clock::now();
            for (int spins = 0; ; ++spins) {
                if (job_seq.load(std::memory_order_acquire) != seen) { have = true; break; }
                _mm_pause();
                if ((spins & 255) == 255 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(hot_us)) break;
            }
            if (!have) {
                std::unique_lock<std::mutex> lock(mu);
                cv.wait(lock, [&] { return stop || job_seq.load(std::memory_order_acquire) != seen; });
                if (stop) return;
            }
            seen = job_seq.load(std::memory_order_acquire);
            for (int c = 0; c < nchunks; ++c) {
                int spins = 0;
                while (ready.load(std::memory_order_acquire) <= c) {
                    _mm_pause();
                    if (++spins > 20000) { std::this_thread::yield(); spins = 0; }
                }
                // this worker's share of chunk c, cut at multiples of 4 agents (16-byte aligned row ranges)
                const long long lo = bounds[c], n = bounds[c + 1] - lo;
                long long s0 = (n * idx / nthreads) & ~3ll, s1 = (n * (idx + 1) / nthreads) & ~3ll;
                if (idx == nthreads - 1) s1 = n;
                if (s1 > s0) expand_rows(lay, job, lo + s0, lo + s1);
            }
            done.fetch_add(1, std::memory_order_release);
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

void HostPool::begin(const HostLayout &lay, const HostJob &job, const long long *bounds, int nchunks)
{
    impl->lay = lay; impl->job = job;
    impl->bounds.assign(bounds, bounds + nchunks + 1);
    impl->nchunks = nchunks;
    impl->ready.store(0, std::memory_order_relaxed);
    impl->done.store(0, std::memory_order_relaxed);
    {
        std::lock_guard<std::mutex> lock(impl->mu);          // (a worker between its spin phase and cv.wait sees the new value)
        impl->job_seq.fetch_add(1, std::memory_order_release);
    }
    impl->cv.notify_all();
}

void HostPool::publish(int chunk) { impl->ready.store(chunk + 1, std::memory_order_release); }

void HostPool::finish()
{
    const int n = (int)impl->workers.size();
    int spins = 0;
    while (impl->done.load(std::memory_order_acquire) < n) {
        _mm_pause();
        if (++spins > 20000) { std::this_thread::yield(); spins = 0; }
    }
}

}  // namespace diral
