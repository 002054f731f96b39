/*
 * parallel.c -- persistent worker threads that run the oracle's (single-threaded, reference-faithful)
 * functions on independent slices of a large workload.
 *
 * TEST INFRASTRUCTURE (bench.py's cpu_baseline / --impl reference legs only).  The reference itself is
 * single-threaded per call (curve25519-dalek has no threading); a caller with many cores splits the work
 * exactly like this: independent sub-MSMs over contiguous slices whose results are added
 * (C/edwards.rs:795-800), independent verify_batch calls over batches of 256 signatures (the largest size the
 * reference benches, ed25519-dalek/benches/ed25519_benchmarks.rs:56-73), independent double-base MSMs.
 * Every slice runs the unmodified oracle function, so the per-core work is the reference algorithm's.
 *
 * A pool keeps `threads` pthreads alive between jobs; a job is a list of items that the workers pull
 * from a shared counter (slices of >= 2^13 pairs keep the w = 8 bucket reduction of pippenger.rs:81-87,
 * :146-151 amortised).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "oracle.h"

typedef void (*item_fn)(void *arg, size_t item, int worker);

typedef struct oracle_pool {
    int threads;
    pthread_t *tids;
    pthread_mutex_t mu;
    pthread_cond_t cv_start, cv_done;
    unsigned long generation;      /* bumped for every job */
    int running;                   /* workers still inside the current job */
    int quit;
    item_fn fn;
    void *arg;
    size_t nitems;
    atomic_size_t next;
} oracle_pool;

typedef struct { oracle_pool *pool; int id; } worker_arg;

static void *worker_main(void *p)
{
    worker_arg *wa = (worker_arg *)p;
    oracle_pool *pool = wa->pool;
    const int id = wa->id;
    free(wa);
    unsigned long seen = 0;
    for (;;) {
        pthread_mutex_lock(&pool->mu);
        while (pool->generation == seen && !pool->quit) pthread_cond_wait(&pool->cv_start, &pool->mu);
        if (pool->quit) { pthread_mutex_unlock(&pool->mu); return NULL; }
        seen = pool->generation;
        pthread_mutex_unlock(&pool->mu);
        for (;;) {
            size_t it = atomic_fetch_add(&pool->next, 1);
            if (it >= pool->nitems) break;
            pool->fn(pool->arg, it, id);
        }
        pthread_mutex_lock(&pool->mu);
        if (--pool->running == 0) pthread_cond_signal(&pool->cv_done);
        pthread_mutex_unlock(&pool->mu);
    }
}

oracle_pool *oracle_pool_create(int threads)
{
    if (threads < 1) threads = 1;
    oracle_pool *pool = (oracle_pool *)calloc(1, sizeof(oracle_pool));
    pool->threads = threads;
    pool->tids = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    pthread_mutex_init(&pool->mu, NULL);
    pthread_cond_init(&pool->cv_start, NULL);
    pthread_cond_init(&pool->cv_done, NULL);
    for (int i = 0; i < threads; i++) {
        worker_arg *wa = (worker_arg *)malloc(sizeof(worker_arg));
        wa->pool = pool; wa->id = i;
        pthread_create(&pool->tids[i], NULL, worker_main, wa);
    }
    return pool;
}

void oracle_pool_destroy(oracle_pool *pool)
{
    if (!pool) return;
    pthread_mutex_lock(&pool->mu);
    pool->quit = 1;
    pthread_cond_broadcast(&pool->cv_start);
    pthread_mutex_unlock(&pool->mu);
    for (int i = 0; i < pool->threads; i++) pthread_join(pool->tids[i], NULL);
    free(pool->tids);
    free(pool);
}

int oracle_pool_threads(const oracle_pool *pool) { return pool ? pool->threads : 0; }

static double now_s(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* run one job and wait for it; returns its wall time in seconds */
static double pool_run(oracle_pool *pool, item_fn fn, void *arg, size_t nitems)
{
    const double t0 = now_s();
    pthread_mutex_lock(&pool->mu);
    pool->fn = fn; pool->arg = arg; pool->nitems = nitems;
    atomic_store(&pool->next, 0);
    pool->running = pool->threads;
    pool->generation++;
    pthread_cond_broadcast(&pool->cv_start);
    while (pool->running) pthread_cond_wait(&pool->cv_done, &pool->mu);
    pthread_mutex_unlock(&pool->mu);
    return now_s() - t0;
}

/* ---- one MSM as independent Pippenger sub-MSMs over contiguous slices ------------------------------------- */
typedef struct { const uint8_t *scalars; const ge_p3 *points; size_t n, slice; ge_p3 *partial; } msm_job;

static void msm_item(void *arg, size_t item, int worker)
{
    (void)worker;
    msm_job *j = (msm_job *)arg;
    const size_t lo = item * j->slice, hi = lo + j->slice < j->n ? lo + j->slice : j->n;
    /* the reference's own dispatch (C/edwards.rs:1025-1029): Pippenger from 190 points */
    edwards_optional_multiscalar_mul(&j->partial[item], j->scalars + 32 * lo, j->points + lo, NULL, hi - lo);
}

/* out = sum scalars[i] * points[i] computed as ceil(n / slice) sub-MSMs on the pool's threads, partial sums added
 * on the calling thread.  Returns the wall time in seconds. */
double oracle_pool_msm(oracle_pool *pool, uint8_t out[32], const uint8_t *scalars, const ge_p3 *points, size_t n, size_t slice)
{
    if (slice == 0) slice = 8192;
    const size_t items = n ? (n + slice - 1) / slice : 0;
    msm_job j = {scalars, points, n, slice, (ge_p3 *)malloc(sizeof(ge_p3) * (items ? items : 1))};
    const double t0 = now_s();
    if (items) pool_run(pool, msm_item, &j, items);
    ge_p3 total; ge_identity(&total);
    for (size_t i = 0; i < items; i++) ge_p3_add(&total, &total, &j.partial[i]);
    ge_compress(out, &total);
    const double dt = now_s() - t0;
    free(j.partial);
    return dt;
}

/* ---- independent verify_batch calls over batches of `batch` signatures (equal message length) --------------- */
typedef struct { const uint8_t *msgs; size_t msg_len; const uint8_t *sigs, *keys; size_t n, batch; int *verdicts; } vb_job;

static void vb_item(void *arg, size_t item, int worker)
{
    (void)worker;
    vb_job *j = (vb_job *)arg;
    const size_t lo = item * j->batch, hi = lo + j->batch < j->n ? lo + j->batch : j->n, cnt = hi - lo;
    const uint8_t **ptrs = (const uint8_t **)malloc(sizeof(uint8_t *) * cnt);
    size_t *lens = (size_t *)malloc(sizeof(size_t) * cnt);
    for (size_t i = 0; i < cnt; i++) { ptrs[i] = j->msgs + (lo + i) * j->msg_len; lens[i] = j->msg_len; }
    j->verdicts[item] = ed25519_verify_batch(ptrs, lens, j->sigs + 64 * lo, j->keys + 32 * lo, cnt, NULL);
    free(ptrs); free(lens);
}

double oracle_pool_verify_batches(oracle_pool *pool, const uint8_t *msgs, size_t msg_len, const uint8_t *sigs,
                                  const uint8_t *keys, size_t n, size_t batch, int *verdicts)
{
    vb_job j = {msgs, msg_len, sigs, keys, n, batch, verdicts};
    return pool_run(pool, vb_item, &j, (n + batch - 1) / batch);
}

/* ---- independent single verifications (VerifyingKey::verify), slices of 64 signatures -------------------------- */
typedef struct { const uint8_t *msgs; size_t msg_len; const uint8_t *sigs, *keys; size_t n; int strict; uint8_t *results; } ve_job;

static void ve_item(void *arg, size_t item, int worker)
{
    (void)worker;
    ve_job *j = (ve_job *)arg;
    const size_t lo = item * 64, hi = lo + 64 < j->n ? lo + 64 : j->n;
    for (size_t i = lo; i < hi; i++)
        j->results[i] = (uint8_t)(j->strict ? ed25519_verify_strict(j->msgs + i * j->msg_len, j->msg_len, j->sigs + 64 * i, j->keys + 32 * i)
                                            : ed25519_verify(j->msgs + i * j->msg_len, j->msg_len, j->sigs + 64 * i, j->keys + 32 * i));
}

double oracle_pool_verify_each(oracle_pool *pool, const uint8_t *msgs, size_t msg_len, const uint8_t *sigs,
                               const uint8_t *keys, size_t n, int strict, uint8_t *results)
{
    ve_job j = {msgs, msg_len, sigs, keys, n, strict, results};
    return pool_run(pool, ve_item, &j, (n + 63) / 64);
}

/* ---- independent constant-time double-base MSMs a_i G + b_i H (config 5), slices of 64 pairs ------------------ */
typedef struct { uint8_t *out; const uint8_t *a, *b, *G, *H; size_t n; int ok; } db_job;

static void db_item(void *arg, size_t item, int worker)
{
    (void)worker;
    db_job *j = (db_job *)arg;
    const size_t lo = item * 64, hi = lo + 64 < j->n ? lo + 64 : j->n;
    if (ristretto_double_base_batch(j->out + 32 * lo, j->a + 32 * lo, j->b + 32 * lo, j->G, j->H, hi - lo)) j->ok = 0;
}

double oracle_pool_double_base(oracle_pool *pool, uint8_t *out, const uint8_t *a, const uint8_t *b, const uint8_t G[32],
                               const uint8_t H[32], size_t n, int *ok)
{
    db_job j = {out, a, b, G, H, n, 1};
    const double dt = pool_run(pool, db_item, &j, (n + 63) / 64);
    if (ok) *ok = j.ok;
    return dt;
}
