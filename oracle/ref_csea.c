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
#include "sea.h"

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
