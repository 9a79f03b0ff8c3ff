/*
 * ref_csea.c -- thin wrapper that compiles the reference's OWN C decoder (c/sea.h, CBR only) where it
 * lies under /root/reference (never copied into this repo) into oracle/_ref/.  TEST INFRASTRUCTURE ONLY.
 *
 * Two uses: (1) libsea_cref.so validates the oracle restatement's CBR decode byte for byte;
 *           (2) csea_bench times the reference C decoder on the host cores (bench.py --impl reference).
 *
 * c/sea.h keeps its dequant table in static globals and frees it at the end of sea_decode without
 * clearing the cache key (c/sea.h:52-55, :225), so a second call in the same process would read freed
 * memory.  The wrapper resets those statics after every full decode; parallel timing uses fork().
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/*
 * LMS capture (tests only).  sea_read_chunk keeps its per-channel LMS in a malloc'ed block that it frees at the end of the
 * chunk (c/sea.h:147-185), so the state the REFERENCE decoder reached at each chunk end is never visible to a caller.  The
 * two macros below route c/sea.h's own malloc/free calls through this file; its source is still compiled where it lies,
 * unmodified.  sea_read_chunk frees exactly three blocks per chunk, in the order residuals, scale_factors, lms
 * (c/sea.h:181-183); the only other free is the dequant table (pointer == SEA_DQT).  The third one is copied out.
 */
static void *ref_hook_malloc(size_t n);
static void ref_hook_free(void *p);
#define malloc(n) ref_hook_malloc(n)
#define free(p) ref_hook_free(p)
#include "sea.h"
#undef malloc
#undef free

static int32_t *g_cap = NULL;      /* [chunk][channel][8]: history[4], weights[4] as c/sea.h holds them (int32) */
static uint32_t g_cap_chunks = 0;  /* capacity in chunks */
static uint32_t g_cap_channels = 0, g_cap_n = 0, g_free_no = 0;

static void *ref_hook_malloc(size_t n) { return malloc(n); }
static void ref_hook_free(void *p)
{
    if (p != (void *)SEA_DQT && g_cap) {
        if (++g_free_no % 3u == 0u && g_cap_n < g_cap_chunks) {
            memcpy(g_cap + (size_t)g_cap_n * g_cap_channels * 8u, p, (size_t)g_cap_channels * sizeof(SEA_LMS));
            g_cap_n++;
        }
    }
    free(p);
}

#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

int ref_csea_decode(uint8_t *encoded, uint32_t encoded_len, uint32_t *sample_rate, uint32_t *channels, int16_t *output,
                    uint32_t *total_frames)
{
    int rc = sea_decode(encoded, encoded_len, sample_rate, channels, output, total_frames);
    if (output != NULL) {
        SEA_DQT = NULL;
        SEA_DQT_COLUMNS = 0;
        SEA_DQT_SCALE_FACTOR_BITS = 0;
        SEA_DQT_RESIDUAL_BITS = 0;
    }
    return rc;
}

/* Full decode that also returns, for every chunk, the LMS state c/sea.h held when the chunk ended: lms_out[chunk][channel][8]
 * (history[4] then weights[4], int32).  *n_chunks receives how many were captured (<= max_chunks). */
int ref_csea_decode_capture_lms(uint8_t *encoded, uint32_t encoded_len, uint32_t *sample_rate, uint32_t *channels, int16_t *output,
                                uint32_t *total_frames, int32_t *lms_out, uint32_t max_chunks, uint32_t *n_chunks)
{
    if (encoded_len < 22 || !output || !lms_out) return 1;
    g_cap = lms_out;
    g_cap_chunks = max_chunks;
    g_cap_channels = encoded[5];
    g_cap_n = 0;
    g_free_no = 0;
    int rc = ref_csea_decode(encoded, encoded_len, sample_rate, channels, output, total_frames);
    g_cap = NULL;
    if (n_chunks) *n_chunks = g_cap_n;
    return rc;
}

#ifdef CSEA_BENCH_MAIN
/* usage: csea_bench <file.sea> <procs> <reps>  -> prints "<seconds> <samples_decoded>" */
int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    fseek(f, 0, SEEK_END);
    long len = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *enc = (uint8_t *)malloc((size_t)len);
    if (fread(enc, 1, (size_t)len, f) != (size_t)len) return 2;
    fclose(f);
    int procs = atoi(argv[2]), reps = atoi(argv[3]);
    uint32_t rate, ch, frames;
    if (ref_csea_decode(enc, (uint32_t)len, &rate, &ch, NULL, &frames)) return 3;
    size_t n = (size_t)frames * ch;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int p = 0; p < procs; p++) {
        pid_t pid = fork();
        if (pid == 0) {
            int16_t *out = (int16_t *)malloc(n * 2 + 4096 * ch);
            int bad = 0;
            for (int r = 0; r < reps; r++) bad |= ref_csea_decode(enc, (uint32_t)len, &rate, &ch, out, &frames);
            _exit(bad ? 1 : 0);
        }
    }
    int bad = 0, st;
    for (int p = 0; p < procs; p++) {
        wait(&st);
        if (!WIFEXITED(st) || WEXITSTATUS(st)) bad = 1;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    double s = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    printf("%.6f %llu\n", s, bad ? 0ULL : (unsigned long long)n * (unsigned long long)procs * (unsigned long long)reps);
    return bad;
}
#endif
