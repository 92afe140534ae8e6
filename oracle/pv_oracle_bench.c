/*
 * pv_oracle_bench.c -- multi-threaded driver of the f32 oracle, used ONLY as the timed CPU
 * baseline of bench.py (cpu_baseline / --impl reference, kind "port").  TEST INFRASTRUCTURE.
 * A pool of pthreads pulls streams from a shared counter; each stream runs the reference's two
 * host loops (src/main.cpp:228-297) through pvo_process_compat_f32.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "pv_oracle.h"

typedef struct job {
    int corrected;
    double beta;
    const float *x;
    long n_streams, n_in, n_frames;
    int N, Ha, Hs;
    const float *win;
    float *out;
    atomic_long next;
    atomic_int rc;
} job;

static void *worker(void *arg)
{
    job *j = (job *)arg;
    const int nb = j->N / 2 + 1;
    float *back = (float *)malloc(sizeof(float) * (size_t)j->N);
    uint32_t *P = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)nb);
    uint64_t *psi = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)nb);
    double *tail = (double *)malloc(sizeof(double) * (size_t)j->N);
    double *dout = j->corrected ? (double *)malloc(sizeof(double) * (size_t)(j->n_frames * j->Hs)) : NULL;
    for (;;) {
        long s = atomic_fetch_add(&j->next, 1);
        if (s >= j->n_streams) break;
        int r;
        if (!j->corrected) {
            memset(back, 0, sizeof(float) * (size_t)j->N);
            r = pvo_process_compat_f32(j->x + s * j->n_in, j->n_in, j->N, j->Ha, j->Hs, j->win, j->n_frames, 0,
                                       j->n_frames, 0, back, j->out + s * j->n_frames * (long)j->Hs);
        } else {
            pvo_corrected_state st = {0, P, psi, tail};
            memset(P, 0, sizeof(uint32_t) * (size_t)nb);
            memset(psi, 0, sizeof(uint64_t) * (size_t)nb);
            memset(tail, 0, sizeof(double) * (size_t)j->N);
            r = pvo_process_corrected(j->x + s * j->n_in, j->n_in, j->N, j->Ha, j->Hs, j->win, 1, &j->beta,
                                      j->n_frames, &st, 32, dout, j->n_frames * (long)j->Hs);
            float *o = j->out + s * j->n_frames * (long)j->Hs;
            for (long i = 0; i < j->n_frames * (long)j->Hs; i++) o[i] = (float)dout[i];
        }
        if (r) atomic_store(&j->rc, r);
    }
    free(back); free(P); free(psi); free(tail); free(dout);
    return NULL;
}

int pvo_bench_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* x: [n_streams][n_in], out: [n_streams][n_frames*Hs]; n_threads <= 0: all online cores.
 * corrected != 0: one voice with pitch ratio beta (f32 arithmetic), else the compat pipeline. */
int pvo_bench_f32(int corrected, double beta, const float *x, long n_streams, long n_in, int N, int Ha, int Hs,
                  const float *win, long n_frames, float *out, int n_threads)
{
    if (n_threads <= 0) n_threads = pvo_bench_threads();
    if (n_threads > n_streams) n_threads = (int)n_streams;
    job j = {corrected, beta, x, n_streams, n_in, n_frames, N, Ha, Hs, win, out, 0, 0};
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, worker, &j);
    for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
    free(th);
    return atomic_load(&j.rc);
}
