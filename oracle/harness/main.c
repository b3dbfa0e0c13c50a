/*
 * oracle/harness/main.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 *   audiosync_harness SOURCE.f64 SAMPLE.f64 [debug] [chunk_delay_us]
 *
 * Loads a pair of raw little-endian double files (2,880,000 and 1,440,000
 * frames, or shorter), hands them to fake_io.c and calls the reference's
 * audiosync_run() (src/audiosync.c:166-284, compiled unmodified).  Prints
 *   ret=<0|-1> lag=<ms on success, last frame lag otherwise>
 * which is exactly what the reference's apps/main.c would report.
 */
#include <stdio.h>
#include <stdlib.h>
#include <audiosync/audiosync.h>

extern const double *harness_source, *harness_sample;
extern size_t harness_source_len, harness_sample_len;
extern unsigned harness_chunk_delay_us;

static double *slurp(const char *path, size_t *n)
{
    FILE *f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END);
    long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    double *p = malloc((size_t) bytes);
    if (!p || fread(p, 1, (size_t) bytes, f) != (size_t) bytes) { perror("read"); exit(2); }
    fclose(f);
    *n = (size_t) bytes / sizeof(double);
    return p;
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s source.f64 sample.f64 [debug] [delay_us]\n", argv[0]); return 2; }
    harness_source = slurp(argv[1], &harness_source_len);
    harness_sample = slurp(argv[2], &harness_sample_len);
    if (argc > 3) audiosync_set_debug(atoi(argv[3]));
    if (argc > 4) harness_chunk_delay_us = (unsigned) atoi(argv[4]);
    long lag = 0;
    if (audiosync_setup("harness") != 0) return 3;
    int ret = audiosync_run("synthetic pair", &lag);
    printf("ret=%d lag=%ld\n", ret, lag);
    return 0;
}
