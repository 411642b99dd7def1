/*
 * oracle/pipeline.c -- CPU ORACLE batch driver (test infrastructure, NOT product code).
 *
 * Runs the whole reference path -- generate_features_for_clip-style frontend loop
 * (microwakeword/audio/audio_utils.py:50-64) followed by the predict_spectrogram loop
 * (microwakeword/inference.py:98-123) -- for many independent streams, one stream at a
 * time per host thread (the reference itself is single-stream, single-thread; threads over
 * independent streams is the most favourable CPU arrangement).  Used by tests for
 * batch parity and by bench.py as the timed CPU baseline ("port": the reference's own
 * native dependencies are not installable here).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "frontend.h"
#include "mixednet.h"

struct job {
    const void *blob; size_t blob_n;
    const int16_t *audio; size_t n_streams, n_samples;
    uint16_t *features; size_t max_rows;       /* optional [S][max_rows][40] */
    float *probs; size_t max_probs;            /* optional [S][max_probs] */
    size_t *rows_out, *probs_out;              /* optional [S] */
    int clip_loop;                             /* 1: strict-'<' clip loop; 0: plain streaming */
    size_t begin, end;
    int status;
};

static void *worker(void *arg) {
    struct job *j = (struct job *)arg;
    mwwo_frontend *fe = mwwo_frontend_create();
    mwwo_mixednet *nn = j->blob ? mwwo_mixednet_create(j->blob, j->blob_n) : NULL;
    const size_t cap = j->n_samples / 160 + 1;
    uint16_t *tmp = (uint16_t *)malloc(cap * MWWO_NUM_CHANNELS * sizeof(uint16_t));
    float *ptmp = (float *)malloc((cap + 1) * sizeof(float));
    if (!fe || !tmp || !ptmp || (j->blob && !nn)) { j->status = -1; goto done; }
    for (size_t s = j->begin; s < j->end; ++s) {
        const int16_t *a = j->audio + s * j->n_samples;
        size_t rows;
        if (j->clip_loop) {
            rows = mwwo_generate_features(a, j->n_samples, tmp, cap);
        } else {
            mwwo_frontend_reset(fe);
            rows = mwwo_frontend_stream(fe, a, j->n_samples, tmp, cap);
        }
        if (j->rows_out) j->rows_out[s] = rows;
        if (j->features) {
            size_t r = rows < j->max_rows ? rows : j->max_rows;
            memcpy(j->features + s * j->max_rows * MWWO_NUM_CHANNELS, tmp, r * MWWO_NUM_CHANNELS * sizeof(uint16_t));
        }
        if (nn) {
            mwwo_mixednet_reset(nn);
            size_t n = mwwo_mixednet_predict_u16(nn, tmp, rows, ptmp, cap);
            if (j->probs_out) j->probs_out[s] = n;
            if (j->probs) {
                size_t r = n < j->max_probs ? n : j->max_probs;
                memcpy(j->probs + s * j->max_probs, ptmp, r * sizeof(float));
            }
        }
    }
done:
    free(tmp); free(ptmp);
    if (fe) mwwo_frontend_free(fe);
    if (nn) mwwo_mixednet_free(nn);
    return NULL;
}

/* Every stream starts from reset state (fresh frontend + zero rings).  Returns 0 on success. */
int mwwo_pipeline_run(const void *blob, size_t blob_n, const int16_t *audio, size_t n_streams, size_t n_samples,
                      uint16_t *features, size_t max_rows, float *probs, size_t max_probs,
                      size_t *rows_out, size_t *probs_out, int clip_loop, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if ((size_t)n_threads > n_streams) n_threads = (int)(n_streams ? n_streams : 1);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    struct job *jobs = (struct job *)calloc(n_threads, sizeof(struct job));
    if (!th || !jobs) { free(th); free(jobs); return -1; }
    const size_t per = (n_streams + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        struct job *j = &jobs[t];
        j->blob = blob; j->blob_n = blob_n; j->audio = audio; j->n_streams = n_streams; j->n_samples = n_samples;
        j->features = features; j->max_rows = max_rows; j->probs = probs; j->max_probs = max_probs;
        j->rows_out = rows_out; j->probs_out = probs_out; j->clip_loop = clip_loop;
        j->begin = per * t; j->end = per * (t + 1) < n_streams ? per * (t + 1) : n_streams;
        if (j->begin > n_streams) j->begin = n_streams;
        pthread_create(&th[t], NULL, worker, j);
    }
    int status = 0;
    for (int t = 0; t < n_threads; ++t) { pthread_join(th[t], NULL); if (jobs[t].status) status = jobs[t].status; }
    free(th); free(jobs);
    return status;
}
