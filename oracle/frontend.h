/* oracle/frontend.h -- CPU ORACLE (test infrastructure only; see frontend.c header). */
#ifndef MWWO_FRONTEND_H_
#define MWWO_FRONTEND_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MWWO_NUM_CHANNELS 40

typedef struct mwwo_frontend mwwo_frontend;

mwwo_frontend *mwwo_frontend_create(void);   /* the okay_nabu configuration, audio_utils.py:71-78 */
/* any upstream FrontendConfig the static buffers can hold (window <= 512 samples, <= 40 channels); the noise
 * reduction / PCAN / log parameters stay the ones audio_utils.py and pymicro-features use.  Exists so the upstream
 * library's own unit-test configuration (1 kHz, 25 ms / 10 ms, 2 channels, 8..450 Hz) can pin this restatement. */
mwwo_frontend *mwwo_frontend_create_cfg(int sample_rate, int window_ms, int step_ms, int num_channels,
                                        float lower_hz, float upper_hz);
void mwwo_frontend_free(mwwo_frontend *s);
void mwwo_frontend_reset(mwwo_frontend *s);

/* One MicroFrontend.ProcessSamples call: consumes up to (480 - buffered) samples, returns 40 when a
 * feature row was produced (written to out[40]) else 0.  *n_read = samples consumed. */
int mwwo_frontend_process(mwwo_frontend *s, const int16_t *samples, size_t n, size_t *n_read, uint16_t *out);

/* generate_features_for_clip(use_c=True) chunk loop, audio_utils.py:50-64; returns the row count */
size_t mwwo_generate_features(const int16_t *audio, size_t n_samples, uint16_t *out, size_t max_rows);

/* stateful streaming over an arbitrary span (no strict-'<' quirk) */
size_t mwwo_frontend_stream(mwwo_frontend *s, const int16_t *audio, size_t n_samples, uint16_t *out, size_t max_rows);

void mwwo_frontend_tables(const mwwo_frontend *s, int16_t *window480, int16_t *bin_channel257,
                          int16_t *bin_weight257, int16_t *bin_unweight257, int16_t *chan_start42,
                          int16_t *gain_lut125, uint16_t *log_lut129, int16_t *twiddles512,
                          int16_t *super256, int32_t *scalars8);
void mwwo_frontend_taps(const mwwo_frontend *s, int32_t *shift, int16_t *fft_in512, int16_t *fft_out514,
                        uint32_t *energy257, uint64_t *work41, uint32_t *sqrt40, uint32_t *nr40,
                        uint32_t *pcan40, uint32_t *estimate40);
void mwwo_frontend_config(const mwwo_frontend *s, int32_t *cfg8);
void mwwo_frontend_window_tap(const mwwo_frontend *s, int16_t *win_out512, int32_t *max_abs);
void mwwo_frontend_get_state(const mwwo_frontend *s, int16_t *input480, int32_t *input_used, uint32_t *estimate40);

uint32_t mwwo_sqrt64(uint64_t x);
int32_t mwwo_wdf(const mwwo_frontend *s, uint32_t x);
uint32_t mwwo_pcan_shrink(uint32_t x);
uint32_t mwwo_log_scaled(const mwwo_frontend *s, uint32_t x);
void mwwo_fftr(const mwwo_frontend *s, const int16_t *in512, int16_t *out514);

#ifdef __cplusplus
}
#endif
#endif
