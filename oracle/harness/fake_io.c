/*
 * oracle/harness/fake_io.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Stand-ins for the three I/O entry points the reference's UNMODIFIED
 * src/audiosync.c links against: capture() (src/capture/linux_capture.c:359-382),
 * download() (src/download/linux_download.c:19-55) and pulseaudio_setup()
 * (linux_capture.c:143-352).  The real ones fork ffmpeg / talk to PulseAudio and
 * YouTube; these feed a pre-generated synthetic pair into the same
 * `struct ffmpeg_data` with the reader protocol of src/ffmpeg_pipe.c:68-149:
 * append BUFSIZE-double chunks to data->buf, bump data->len, signal
 * `interval_done` under `mutex` when an interval boundary is crossed, stop early
 * on ABORT_ST, zero-fill and signal at end of data.
 *
 * With these, audiosync_run() -- the reference's real interval loop -- runs with
 * no ffmpeg, PulseAudio or network, against either the reference's own
 * cross_correlation.c (CPU) or libaudiosync_cuda.so (GPU).
 */
#include <stdio.h>
#include <string.h>
#include <unistd.h>
#include <audiosync/audiosync.h>
#include <audiosync/capture/linux_capture.h>
#include <audiosync/download/linux_download.h>

#define BUFSIZE 4096                      /* src/ffmpeg_pipe.c chunk size (doubles) */

const double *harness_source = NULL;      /* 2 * 30 s of frames, set by main() */
const double *harness_sample = NULL;      /* 30 s of frames                      */
size_t harness_source_len = 0, harness_sample_len = 0;
unsigned harness_chunk_delay_us = 0;      /* optional pacing between chunks      */

static void *feed(struct ffmpeg_data *data, const double *from, size_t avail)
{
    size_t interval_count = 0;
    data->len = 0;
    while (1) {
        size_t n = BUFSIZE;
        if (data->len + n > avail) n = avail - data->len;
        if (data->len + n > data->total_len) n = data->total_len - data->len;
        memcpy(data->buf + data->len, from + data->len, n * sizeof(double));
        data->len += n;
        /* end of data, or the buffer would not take another chunk (ffmpeg_pipe.c:84-88) */
        if (n == 0 || data->len + BUFSIZE >= data->total_len) break;
        if (interval_count < data->n_intervals && data->len >= data->intervals[interval_count]) {
            pthread_mutex_lock(&mutex);
            pthread_cond_signal(&interval_done);
            pthread_mutex_unlock(&mutex);
            interval_count++;
        }
        if (audiosync_status() == ABORT_ST) return NULL;      /* ffmpeg_pipe.c:100-106 */
        if (harness_chunk_delay_us) usleep(harness_chunk_delay_us);
    }
    if (data->len < data->total_len) {                        /* ffmpeg_pipe.c:139-149 */
        for (size_t i = data->len; i < data->total_len; i++) data->buf[i] = 0.0;
        data->len = data->total_len;
        pthread_mutex_lock(&mutex);
        pthread_cond_signal(&interval_done);
        pthread_mutex_unlock(&mutex);
    }
    return NULL;
}

void *capture(void *arg)
{
    return feed((struct ffmpeg_data *) arg, harness_sample, harness_sample_len);
}

void *download(void *arg)
{
    return feed((struct ffmpeg_data *) arg, harness_source, harness_source_len);
}

int pulseaudio_setup(const char *stream_name)
{
    (void) stream_name;
    return 0;
}

int get_audio_url(const char *title, char **url)
{
    (void) title; (void) url;
    return -1;
}
