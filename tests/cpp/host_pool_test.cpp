// CPU test of the row-assembly pool (diral_host.cpp): queued jobs released by "device" flags in a scrambled order must
// give the rows a single-threaded expand_rows gives; an aborted job (flags that never come) must not hang.
#include "diral_host.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

using namespace diral;

int main()
{
    const int N = 32, R = 20, B = 20, S = R + B;
    const long long E = 96, A = E * N, rec = 24;
    HostLayout lay{};
    lay.N = N; lay.R = R; lay.B = B; lay.S = S; lay.add_action = 1; lay.action_binary = 1; lay.piggy = 1; lay.L = 800; lay.nt_stores = 0;
    HostPool pool(3);
    const int JOBS = 5, CHUNKS = 6;
    std::vector<std::vector<int32_t>> act(JOBS, std::vector<int32_t>(A));
    std::vector<std::vector<uint8_t>> recs(JOBS, std::vector<uint8_t>(A * rec + 64));
    std::vector<float *> out(JOBS), ref(JOBS), rw(JOBS);
    srand(7);
    for (int j = 0; j < JOBS; ++j) {
        for (long long a = 0; a < A; ++a) {
            act[j][a] = rand() % R;
            int left = 31;
            for (int b = 0; b < B; ++b) { const int c = rand() % 3; recs[j][a * rec + b] = (uint8_t)(c <= left ? c : 0); left -= c <= left ? c : 0; }
            const float r = (float)(rand() % 7) - 3.0f;
            memcpy(&recs[j][a * rec + 20], &r, 4);
        }
        out[j] = (float *)aligned_alloc(64, A * S * 4); ref[j] = (float *)aligned_alloc(64, A * S * 4); rw[j] = (float *)aligned_alloc(64, A * 4);
        memset(out[j], 0xff, A * S * 4);
    }
    auto job_of = [&](int j, float *dst) {
        HostJob job{};
        job.actions = act[j].data(); job.counts = recs[j].data(); job.count_stride = rec;
        job.rews = reinterpret_cast<const float *>(recs[j].data() + 20); job.rew_stride = rec; job.rews_out = rw[j]; job.out = dst;
        return job;
    };
    for (int j = 0; j < JOBS; ++j) { HostJob job = job_of(j, ref[j]); expand_rows(lay, job, 0, A); }

    long long bounds[CHUNKS + 1];
    for (int c = 0; c <= CHUNKS; ++c) bounds[c] = (E * c / CHUNKS) * N;
    static volatile unsigned flags[JOBS][CHUNKS];
    unsigned long long ids[JOBS];
    for (int j = 0; j < JOBS; ++j) {                       // all jobs queued before any chunk is released
        for (int c = 0; c < CHUNKS; ++c) flags[j][c] = 0;
        ids[j] = pool.begin(lay, job_of(j, out[j]), bounds, CHUNKS, flags[j], 41u + j);
    }
    // release: later jobs' chunks first, chunks of a job out of order (the workers must still go job by job, chunk by chunk)
    for (int c = CHUNKS - 1; c >= 0; --c)
        for (int j = JOBS - 1; j >= 0; --j) {
            std::this_thread::sleep_for(std::chrono::microseconds(200));
            flags[j][c] = 41u + j;
        }
    int bad = 0;
    for (int j = 0; j < JOBS; ++j) {
        pool.finish(ids[j]);
        if (!pool.done(ids[j])) { printf("job %d not done after finish\n", j); ++bad; }
        if (memcmp(out[j], ref[j], A * S * 4)) { printf("job %d rows differ\n", j); ++bad; }
        for (long long a = 0; a < A; ++a) { float r; memcpy(&r, &recs[j][a * rec + 20], 4); if (rw[j][a] != r) { printf("job %d reward %lld differs\n", j, a); ++bad; break; } }
    }
    // publish()-released job (no flags) after the flag-released ones
    memset(out[0], 0xff, A * S * 4);
    const unsigned long long idp = pool.begin(lay, job_of(0, out[0]), bounds, CHUNKS);
    for (int c = 0; c < CHUNKS; ++c) pool.publish(idp, c);
    pool.finish(idp);
    if (memcmp(out[0], ref[0], A * S * 4)) { printf("published job rows differ\n"); ++bad; }
    // abort: flags that never come
    static volatile unsigned never[CHUNKS] = {0, 0, 0, 0, 0, 0};
    const unsigned long long ida = pool.begin(lay, job_of(1, out[1]), bounds, CHUNKS, never, 9u);
    std::this_thread::sleep_for(std::chrono::milliseconds(2));
    if (pool.done(ida)) { printf("aborted job finished by itself\n"); ++bad; }
    pool.abort(ida);
    pool.finish(ida);
    // the pool still works afterwards
    memset(out[2], 0xff, A * S * 4);
    const unsigned long long idq = pool.begin(lay, job_of(2, out[2]), bounds, CHUNKS);
    pool.publish(idq, CHUNKS - 1);
    pool.finish(idq);
    if (memcmp(out[2], ref[2], A * S * 4)) { printf("job after abort differs\n"); ++bad; }
    // more jobs than ring slots, back to back
    for (int k = 0; k < 20; ++k) {
        const unsigned long long id = pool.begin(lay, job_of(k % JOBS, out[k % JOBS]), bounds, CHUNKS);
        pool.publish(id, CHUNKS - 1);
        if (k % 3 == 0) pool.finish(id);
    }
    for (int j = 0; j < JOBS; ++j) { /* drain */ }
    const unsigned long long idl = pool.begin(lay, job_of(3, out[3]), bounds, 1);
    pool.publish(idl, 0); pool.finish(idl);
    for (int j = 0; j < JOBS; ++j) if (memcmp(out[j], ref[j], A * S * 4)) { printf("ring reuse: job %d rows differ\n", j); ++bad; }
    printf(bad ? "FAILED\n" : "host pool ok\n");
    return bad ? 1 : 0;
}
