/*
 * sea_oracle.c -- CPU restatement of the SEA codec hot path (chanderlud/sea-codec 0.5.3).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sea_codec_b200/ may include, link or call this
 * file; it is imported by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
 * --impl reference legs, and only as the checker or the timed CPU baseline.
 *
 * Parity status: the CBR *decode* half is pinned against the reference's own C decoder
 * (c/sea.h, compiled into oracle/_ref by oracle/Makefile) and every table against the
 * known-answer values of SURVEY.md Appendix D.  The encode half and everything VBR have no
 * golden vectors in the reference (tests/ hold none, cargo/rustc are absent), so for those the
 * parity is "unpinned": this file is a literal restatement, written after the Rust control flow
 * (sequential candidate loop with early exit, byte-at-a-time packer) so that it shares no
 * structure with the CUDA kernels it checks.
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fwrapv (see oracle/Makefile); f32 steps are
 * plain IEEE single operations in source order (SURVEY trap T16).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define LMS_LEN 4
#define ORACLE_OK 0
#define ORACLE_ERR_READ -1          /* SeaError::ReadError / io UnexpectedEof            */
#define ORACLE_ERR_INVALID_PARAMS -2
#define ORACLE_ERR_INVALID_FILE -3  /* SeaError::InvalidFile                             */
#define ORACLE_ERR_INVALID_FRAME -4 /* SeaError::InvalidFrame                            */
#define ORACLE_ERR_CLOSED -5        /* SeaError::EncoderClosed                           */
#define ORACLE_ERR_PANIC -100       /* the reference would panic (assert/unwrap/OOB)     */
#define ORACLE_ERR_CAPACITY -101    /* caller buffer too small (oracle-only condition)   */

/* ------------------------------------------------------------------ small helpers */

typedef struct {
    uint8_t *data;
    size_t len, cap;
} bytevec;

static void bv_push(bytevec *v, uint8_t b)
{
    if (v->len == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 256;
        v->data = (uint8_t *)realloc(v->data, v->cap);
    }
    v->data[v->len++] = b;
}
static void bv_extend(bytevec *v, const uint8_t *p, size_t n)
{
    for (size_t i = 0; i < n; i++) bv_push(v, p[i]);
}
static void bv_free(bytevec *v)
{
    free(v->data);
    v->data = NULL;
    v->len = v->cap = 0;
}

/* common.rs:5-8 */
static inline int16_t clamp_i16(int32_t v)
{
    if (v < -32768) return -32768;
    if (v > 32767) return 32767;
    return (int16_t)v;
}

/* ------------------------------------------------------------------ lms.rs */

typedef struct {
    int32_t history[LMS_LEN];
    int32_t weights[LMS_LEN];
} sea_lms;

/* lms.rs:19-32 */
static void lms_init(sea_lms *l)
{
    memset(l, 0, sizeof(*l));
    l->weights[LMS_LEN - 2] = -(1 << (16 - 3));
    l->weights[LMS_LEN - 1] = 1 << (17 - 3);
}

/* lms.rs:33-41; release builds wrap (trap T7) -> unsigned arithmetic */
static inline int32_t lms_predict(const sea_lms *l)
{
    uint32_t prediction = 0;
    for (int i = 0; i < LMS_LEN; i++) prediction += (uint32_t)l->weights[i] * (uint32_t)l->history[i];
    return ((int32_t)prediction) >> (16 - 3);
}

/* lms.rs:43-51 */
static inline void lms_update(sea_lms *l, int16_t sample, int32_t residual)
{
    int32_t delta = residual >> (3 + 1);
    for (int i = 0; i < LMS_LEN; i++) {
        uint32_t d = (uint32_t)(l->history[i] < 0 ? -delta : delta);
        l->weights[i] = (int32_t)((uint32_t)l->weights[i] + d);
    }
    l->history[0] = l->history[1];
    l->history[1] = l->history[2];
    l->history[2] = l->history[3];
    l->history[LMS_LEN - 1] = (int32_t)sample;
}

/* lms.rs:53-62 */
static inline uint64_t lms_weights_penalty(const sea_lms *l)
{
    uint64_t sum = 0; /* i64 in the reference; wrapping in release */
    for (int i = 0; i < LMS_LEN; i++) sum += (uint64_t)((int64_t)l->weights[i] * (int64_t)l->weights[i]);
    int64_t penalty = ((int64_t)sum >> 18) - 0x8ff;
    uint64_t p = penalty > 0 ? (uint64_t)penalty : 0;
    return p * p;
}

/* lms.rs:64-78 */
static void lms_serialize(const sea_lms *l, uint8_t out[16])
{
    for (int i = 0; i < LMS_LEN; i++) {
        uint32_t h = (uint32_t)l->history[i], w = (uint32_t)l->weights[i];
        out[i * 2] = (uint8_t)h;
        out[i * 2 + 1] = (uint8_t)(h >> 8);
        out[LMS_LEN * 2 + i * 2] = (uint8_t)w;
        out[LMS_LEN * 2 + i * 2 + 1] = (uint8_t)(w >> 8);
    }
}

/* lms.rs:80-94 */
static void lms_from_bytes(sea_lms *l, const uint8_t d[16])
{
    for (int i = 0; i < LMS_LEN; i++) {
        l->history[i] = (int16_t)(d[i * 2] | (d[i * 2 + 1] << 8));
        l->weights[i] = (int16_t)(d[LMS_LEN * 2 + i * 2] | (d[LMS_LEN * 2 + i * 2 + 1] << 8));
    }
}

/* ------------------------------------------------------------------ qt.rs */

#define QT_LEN (5 + 9 + 17 + 33 + 65 + 129 + 257 + 513)
typedef struct {
    size_t offsets[9];
    uint8_t quant_tab[QT_LEN];
} sea_quant_tab;

/* qt.rs:9-31 */
static void qt_fill(uint8_t *slice, size_t items)
{
    size_t midpoint = items / 2;
    int32_t x = (int32_t)(items / 2 - 1);
    slice[0] = (uint8_t)x;
    for (size_t i = 1; i < midpoint; i += 2) {
        slice[i] = (uint8_t)x;
        slice[i + 1] = (uint8_t)x;
        x -= 2;
    }
    x = 0;
    for (size_t i = midpoint; i < items - 1; i += 2) {
        slice[i] = (uint8_t)x;
        slice[i + 1] = (uint8_t)x;
        x += 2;
    }
    slice[items - 1] = (uint8_t)(x - 2);
    if (items == 9) {
        slice[2] = 1;
        slice[6] = 0;
    }
}

/* qt.rs:33-52 */
static void qt_init(sea_quant_tab *q)
{
    memset(q, 0, sizeof(*q));
    size_t current_offset = 0;
    for (int shift = 2; shift <= 9; shift++) {
        q->offsets[shift - 1] = current_offset;
        size_t items = ((size_t)1 << shift) + 1;
        qt_fill(&q->quant_tab[current_offset], items);
        current_offset += items;
    }
}

/* ------------------------------------------------------------------ dqt.rs */

static const float IDEAL_POW_FACTOR[8] = {12.0f, 11.65f, 11.20f, 10.58f, 9.64f, 8.75f, 7.66f, 6.63f}; /* dqt.rs:14 */

typedef struct {
    int scale_factor_bits;
    int sf_items;
    int32_t *recip[9];  /* [residual_bits][sf]          dqt.rs:57-69  */
    int32_t *dqt[9];    /* [residual_bits][sf*2^b+code] dqt.rs:99-126 */
} sea_dequant_tab;

/* dqt.rs:40-55 */
static void dqt_scale_factors(int residual_bits, int scale_factor_bits, int32_t *out)
{
    float power_factor = IDEAL_POW_FACTOR[residual_bits - 1] / (float)scale_factor_bits;
    int items = 1 << scale_factor_bits;
    for (int index = 1; index <= items; index++) {
        float value = powf((float)index, power_factor);
        out[index - 1] = (int32_t)value;
    }
}

/* dqt.rs:75-97 */
static int dqt_curve(int residual_bits, float *curve)
{
    if (residual_bits == 1) {
        curve[0] = 2.0f;
        return 1;
    }
    if (residual_bits == 2) {
        curve[0] = 1.115f;
        curve[1] = 4.0f;
        return 2;
    }
    float start = 0.75f;
    int steps = 1 << (residual_bits - 1);
    float end = (float)((1 << residual_bits) - 1);
    float step = (end - start) / (float)(steps - 1);
    float step_floor = floorf(step);
    for (int i = 0; i < steps; i++) curve[i] = 0.0f;
    for (int i = 1; i < steps; i++) curve[i] = 0.5f + (float)i * step_floor;
    curve[0] = start;
    curve[steps - 1] = end;
    return steps;
}

static void dqt_free(sea_dequant_tab *t)
{
    for (int b = 0; b < 9; b++) {
        free(t->recip[b]);
        free(t->dqt[b]);
        t->recip[b] = t->dqt[b] = NULL;
    }
}

/* dqt.rs:17-38, 57-69, 99-126 */
static void dqt_init(sea_dequant_tab *t, int scale_factor_bits)
{
    memset(t, 0, sizeof(*t));
    t->scale_factor_bits = scale_factor_bits;
    int items = 1 << scale_factor_bits;
    t->sf_items = items;
    int32_t *sf = (int32_t *)malloc(sizeof(int32_t) * items);
    for (int b = 1; b <= 8; b++) {
        dqt_scale_factors(b, scale_factor_bits, sf);
        t->recip[b] = (int32_t *)malloc(sizeof(int32_t) * items);
        for (int s = 0; s < items; s++) {
            float value = (float)(1 << 16) / (float)sf[s];
            t->recip[b][s] = (int32_t)value;
        }
        float curve[128];
        int n = dqt_curve(b, curve);
        int dqt_items = 1 << (b - 1);
        int cols = 1 << b;
        t->dqt[b] = (int32_t *)malloc(sizeof(int32_t) * items * cols);
        for (int s = 0; s < items; s++) {
            int k = 0;
            for (int q = 0; q < n && q < dqt_items; q++) {
                int32_t val = (int32_t)roundf((float)sf[s] * curve[q]);
                t->dqt[b][s * cols + k++] = val;
                t->dqt[b][s * cols + k++] = -val;
            }
        }
    }
    free(sf);
}

/* ------------------------------------------------------------------ bits.rs */

/* bits.rs:89-135 */
typedef struct {
    uint32_t accum, bits_stored;
    bytevec output;
} bit_packer;

static void packer_push(bit_packer *p, uint32_t input, uint8_t bits)
{
    uint32_t mask = (1u << bits) - 1;
    uint32_t value = input & mask;
    p->accum = (p->accum << bits) | value;
    p->bits_stored += bits;
    if (p->bits_stored >= 8) {
        uint32_t v = p->accum >> (p->bits_stored - 8);
        bv_push(&p->output, (uint8_t)v);
        p->bits_stored -= 8;
        p->accum &= (1u << p->bits_stored) - 1;
    }
}
static void packer_finish(bit_packer *p, bytevec *dst)
{
    if (p->bits_stored > 0) {
        uint8_t byte = (uint8_t)(p->accum << (8 - p->bits_stored));
        bv_push(&p->output, byte);
    }
    bv_extend(dst, p->output.data, p->output.len);
    bv_free(&p->output);
    p->accum = p->bits_stored = 0;
}

static const uint32_t UNPACK_MASKS[9] = {0, 1, 3, 7, 15, 31, 63, 127, 255};

/* bits.rs:34-50 */
static void unpack_const(uint8_t bits8, const uint8_t *input, size_t n, bytevec *out)
{
    uint32_t bits = bits8, mask = UNPACK_MASKS[bits], bits_stored = 0, carry = 0;
    for (size_t i = 0; i < n; i++) {
        uint32_t value = (carry << 8) | input[i];
        bits_stored += 8;
        while (bits_stored >= bits) {
            bv_push(out, (uint8_t)((value >> (bits_stored - bits)) & mask));
            bits_stored -= bits;
        }
        carry = value & ((1u << bits_stored) - 1);
    }
}

/* bits.rs:52-70 */
static void unpack_variable(const uint8_t *bitlengths, size_t n_lengths, const uint8_t *input, size_t n, bytevec *out)
{
    uint32_t bits_stored = 0, carry = 0;
    size_t idx = 0;
    for (size_t i = 0; i < n; i++) {
        uint32_t value = (carry << 8) | input[i];
        bits_stored += 8;
        while (idx < n_lengths && bits_stored >= bitlengths[idx]) {
            uint32_t bits = bitlengths[idx];
            bv_push(out, (uint8_t)((value >> (bits_stored - bits)) & UNPACK_MASKS[bits]));
            bits_stored -= bits;
            idx++;
        }
        carry = value & ((1u << bits_stored) - 1);
    }
}

/* bits.rs:72-78 (trap T17: a single bit length takes the const path) */
static void unpack_process(const uint8_t *bitlengths, size_t n_lengths, const uint8_t *input, size_t n, bytevec *out)
{
    if (n_lengths == 1) {
        unpack_const(bitlengths[0], input, n, out);
        return;
    }
    unpack_variable(bitlengths, n_lengths, input, n, out);
}

static void bv_resize(bytevec *v, size_t n)
{
    while (v->len < n) bv_push(v, 0);
    v->len = n;
}

/* ------------------------------------------------------------------ encoder settings / header */

typedef struct {
    uint8_t scale_factor_bits;
    uint8_t scale_factor_frames;
    float residual_bits;
    uint16_t frames_per_chunk;
    uint8_t vbr;
} oracle_settings; /* encoder.rs:16-23 */

typedef struct {
    uint8_t version, channels;
    uint16_t chunk_size, frames_per_chunk;
    uint32_t sample_rate, total_frames;
} sea_file_header; /* file.rs:21-30, metadata always empty (encoder.rs:65) */

/* file.rs:78-93 */
static void header_serialize(const sea_file_header *h, bytevec *out)
{
    const uint8_t magic[4] = {'s', 'e', 'a', 'c'};
    bv_extend(out, magic, 4);
    bv_push(out, h->version);
    bv_push(out, h->channels);
    bv_push(out, (uint8_t)h->chunk_size);
    bv_push(out, (uint8_t)(h->chunk_size >> 8));
    bv_push(out, (uint8_t)h->frames_per_chunk);
    bv_push(out, (uint8_t)(h->frames_per_chunk >> 8));
    for (int i = 0; i < 4; i++) bv_push(out, (uint8_t)(h->sample_rate >> (8 * i)));
    for (int i = 0; i < 4; i++) bv_push(out, (uint8_t)(h->total_frames >> (8 * i)));
    for (int i = 0; i < 4; i++) bv_push(out, 0); /* metadata_len = 0 */
}

/* ------------------------------------------------------------------ encoder_base.rs */

typedef struct {
    int channels, scale_factor_bits;
    int32_t *prev_scalefactor;
    sea_dequant_tab dequant_tab;
    sea_quant_tab quant_tab;
    sea_lms *lms;
    uint8_t *best_residual_bits, *current_residuals;
    size_t scratch_cap;
} encoder_base;

/* encoder_base.rs:22-26 */
static inline int32_t sea_div(int32_t v, int64_t recip)
{
    int64_t n = ((int64_t)v * recip + (1 << 15)) >> 16;
    int64_t sv = (v > 0) - (v < 0), sn = (n > 0) - (n < 0);
    return (int32_t)(n + (sv - sn));
}

/* encoder_base.rs:29-41 */
static void eb_init(encoder_base *e, int channels, int scale_factor_bits)
{
    memset(e, 0, sizeof(*e));
    e->channels = channels;
    e->scale_factor_bits = scale_factor_bits;
    e->prev_scalefactor = (int32_t *)calloc(channels, sizeof(int32_t));
    dqt_init(&e->dequant_tab, scale_factor_bits);
    qt_init(&e->quant_tab);
    e->lms = (sea_lms *)malloc(sizeof(sea_lms) * channels);
    for (int c = 0; c < channels; c++) lms_init(&e->lms[c]);
}
static void eb_free(encoder_base *e)
{
    free(e->prev_scalefactor);
    dqt_free(&e->dequant_tab);
    free(e->lms);
    free(e->best_residual_bits);
    free(e->current_residuals);
}

/* encoder_base.rs:44-92 -- one candidate trial over one block, with the early exit */
static uint64_t eb_calculate_residuals(const encoder_base *e, int channels, const int32_t *dequant_row,
                                       const int16_t *samples, size_t n_samples, int32_t scalefactor, sea_lms *lms,
                                       uint64_t best_rank, int residual_size, const int32_t *recips,
                                       uint8_t *current_residuals)
{
    uint64_t current_rank = 0;
    int32_t clamp_limit = 1 << residual_size;
    int32_t quant_tab_offset = clamp_limit + (int32_t)e->quant_tab.offsets[residual_size];
    size_t index = 0;
    for (size_t pos = 0; pos < n_samples; pos += (size_t)channels, index++) {
        int32_t sample = samples[pos];
        int32_t predicted = lms_predict(lms);
        int32_t residual = (int32_t)((uint32_t)sample - (uint32_t)predicted);
        int32_t scaled = sea_div(residual, (int64_t)recips[scalefactor]);
        int32_t clamped = scaled < -clamp_limit ? -clamp_limit : (scaled > clamp_limit ? clamp_limit : scaled);
        uint8_t quantized = e->quant_tab.quant_tab[quant_tab_offset + clamped];
        int32_t dequantized = dequant_row[quantized];
        int16_t reconstructed = clamp_i16((int32_t)((uint32_t)predicted + (uint32_t)dequantized));
        int64_t error = (int64_t)sample - (int64_t)reconstructed;
        uint64_t error_sq = (uint64_t)(error * error);
        current_rank += error_sq + lms_weights_penalty(lms);
        if (current_rank > best_rank) break;
        lms_update(lms, reconstructed, dequantized);
        current_residuals[index] = quantized;
    }
    return current_rank;
}

/* encoder_base.rs:95-144 -- rotated candidate order, strict '<' keeps the first minimum (trap T1) */
static uint64_t eb_best_scalefactor(encoder_base *e, int channels, const int32_t *dqt, int cols, const int32_t *recips,
                                    const int16_t *samples, size_t n_samples, int32_t prev_scalefactor,
                                    const sea_lms *ref_lms, int residual_size, size_t n_res, sea_lms *best_lms_out,
                                    int32_t *best_sf_out)
{
    uint64_t best_rank = UINT64_MAX;
    sea_lms best_lms;
    memset(&best_lms, 0, sizeof(best_lms));
    int32_t best_scalefactor = 0;
    sea_lms current_lms = *ref_lms;
    int32_t scalefactor_end = 1 << e->scale_factor_bits;
    for (int32_t sfi = 0; sfi < scalefactor_end; sfi++) {
        int32_t scalefactor = (sfi + prev_scalefactor) % scalefactor_end;
        current_lms = *ref_lms;
        const int32_t *row = &dqt[scalefactor * cols];
        uint64_t current_rank = eb_calculate_residuals(e, channels, row, samples, n_samples, scalefactor, &current_lms,
                                                       best_rank, residual_size, recips, e->current_residuals);
        if (current_rank < best_rank) {
            best_rank = current_rank;
            memcpy(e->best_residual_bits, e->current_residuals, n_res);
            best_lms = current_lms;
            best_scalefactor = scalefactor;
        }
    }
    *best_lms_out = best_lms;
    *best_sf_out = best_scalefactor;
    return best_rank;
}

/* encoder_base.rs:146-195 -- handles ONE scale-factor block of all channels */
static void eb_residuals_for_block(encoder_base *e, const int16_t *samples, size_t n_samples, const uint8_t *residual_size,
                                   uint8_t *scale_factors, uint8_t *residuals, uint64_t *ranks)
{
    size_t n_res = n_samples / (size_t)e->channels;
    if (n_res > e->scratch_cap) {
        e->best_residual_bits = (uint8_t *)realloc(e->best_residual_bits, n_res);
        e->current_residuals = (uint8_t *)realloc(e->current_residuals, n_res);
        e->scratch_cap = n_res;
    }
    memset(e->best_residual_bits, 0, n_res); /* resize(.., 0) keeps old contents; all slots are overwritten by the first
                                                candidate, which never exits early (best_rank = MAX) */
    for (int c = 0; c < e->channels; c++) {
        int b = residual_size[c];
        const int32_t *dqt = e->dequant_tab.dqt[b];
        const int32_t *recips = e->dequant_tab.recip[b];
        sea_lms best_lms;
        int32_t best_sf;
        uint64_t best_rank = eb_best_scalefactor(e, e->channels, dqt, 1 << b, recips, samples + c, n_samples - (size_t)c,
                                                 e->prev_scalefactor[c], &e->lms[c], b, n_res, &best_lms, &best_sf);
        e->prev_scalefactor[c] = best_sf;
        e->lms[c] = best_lms;
        scale_factors[c] = (uint8_t)best_sf;
        ranks[c] = best_rank;
        for (size_t i = 0; i < n_res; i++) residuals[i * (size_t)e->channels + (size_t)c] = e->best_residual_bits[i];
    }
}

/* ------------------------------------------------------------------ encoded samples */

typedef struct {
    uint8_t *scale_factors;
    size_t n_scale_factors;
    uint8_t *residuals;
    size_t n_residuals;
    uint8_t *residual_bits; /* empty (n=0) for CBR: trap T19 */
    size_t n_residual_bits;
} encoded_samples; /* common.rs:125-130 */

static void es_free(encoded_samples *s)
{
    free(s->scale_factors);
    free(s->residuals);
    free(s->residual_bits);
    memset(s, 0, sizeof(*s));
}

static size_t div_ceil(size_t a, size_t b) { return (a + b - 1) / b; }

/* ------------------------------------------------------------------ encoder_cbr.rs */

/* encoder_cbr.rs:36-66 */
static void cbr_encode(encoder_base *e, int residual_size, int scale_factor_frames, const int16_t *samples, size_t n,
                       encoded_samples *out)
{
    size_t channels = (size_t)e->channels;
    memset(out, 0, sizeof(*out));
    out->n_scale_factors = div_ceil(n / channels, (size_t)scale_factor_frames) * channels;
    out->scale_factors = (uint8_t *)calloc(out->n_scale_factors + 1, 1);
    out->n_residuals = n;
    out->residuals = (uint8_t *)calloc(n + 1, 1);
    uint64_t *ranks = (uint64_t *)calloc(channels, sizeof(uint64_t));
    size_t slice_size = (size_t)scale_factor_frames * channels;
    uint8_t *sizes = (uint8_t *)malloc(channels);
    memset(sizes, residual_size, channels);
    size_t slice_index = 0;
    for (size_t pos = 0; pos < n; pos += slice_size, slice_index++) {
        size_t len = n - pos < slice_size ? n - pos : slice_size;
        eb_residuals_for_block(e, samples + pos, len, sizes, out->scale_factors + slice_index * channels,
                               out->residuals + slice_index * slice_size, ranks);
    }
    free(ranks);
    free(sizes);
}

/* ------------------------------------------------------------------ encoder_vbr.rs */

static const float TARGET_RESIDUAL_DISTRIBUTION[6] = {0.00f, 0.00f, 0.95f, 0.05f, 0.00f, 0.00f}; /* encoder_vbr.rs:21 */

/* encoder_vbr.rs:40-63 */
static float vbr_normalized_bitrate(const oracle_settings *s)
{
    float vbr_bitrate = s->residual_bits;
    vbr_bitrate -= ((float)LMS_LEN * 16.0f * 2.0f) / (float)s->frames_per_chunk;
    vbr_bitrate -= (float)s->scale_factor_bits / (float)s->scale_factor_frames;
    vbr_bitrate -= 2.0f / (float)s->scale_factor_frames;
    float base_residuals = floorf(s->residual_bits);
    float new_bitrate = TARGET_RESIDUAL_DISTRIBUTION[1] * (base_residuals - 1.0f) +
                        TARGET_RESIDUAL_DISTRIBUTION[2] * base_residuals +
                        TARGET_RESIDUAL_DISTRIBUTION[3] * (base_residuals + 1.0f) +
                        TARGET_RESIDUAL_DISTRIBUTION[4] * (base_residuals + 2.0f);
    float diff = new_bitrate - base_residuals;
    vbr_bitrate -= diff;
    return vbr_bitrate;
}

/* Rust `f32 as usize`/`as u8`: saturating, NaN -> 0 */
static size_t f32_as_usize(float v)
{
    if (!(v > 0.0f)) return 0;
    if (v >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)v;
}
static uint8_t f32_as_u8(float v)
{
    if (!(v > 0.0f)) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}

/* encoder_vbr.rs:66-96 */
static void vbr_interpolate_distribution(size_t items, float target_rate, size_t res[4])
{
    float frac = target_rate - truncf(target_rate); /* f32::fract */
    float om_frac = 1.0f - frac;
    float percentages[4];
    for (int i = 0; i < 4; i++)
        percentages[i] = TARGET_RESIDUAL_DISTRIBUTION[i] * frac + TARGET_RESIDUAL_DISTRIBUTION[i + 1] * om_frac;
    res[0] = res[1] = res[2] = res[3] = 0;
    size_t sum = 0;
    while (sum < items) {
        size_t remaining = items - sum;
        for (int i = 0; i < 4; i++) {
            size_t value = f32_as_usize((float)remaining * percentages[i]);
            sum += value;
            res[i] += value;
        }
        if (items - sum == remaining) {
            sum += remaining;
            res[1] += remaining;
        }
    }
}

typedef struct {
    uint64_t err;
    uint16_t idx;
} err_idx;

static int cmp_err_idx(const void *a, const void *b)
{
    const err_idx *x = (const err_idx *)a, *y = (const err_idx *)b;
    if (x->err != y->err) return x->err < y->err ? -1 : 1;
    return (int)x->idx - (int)y->idx; /* trap T13: (error, index) order; ties counted below */
}

/* encoder_vbr.rs:98-137; returns 0 or ORACLE_ERR_PANIC; *ties += boundary-spanning ties (T13) */
static int vbr_choose_residual_len(float target, int scale_factor_frames, size_t input_len, const uint64_t *errors,
                                   size_t n_errors, uint8_t *residual_sizes, uint64_t *ties)
{
    size_t sortable_items = input_len / (size_t)scale_factor_frames;
    if (sortable_items > n_errors) return ORACLE_ERR_PANIC; /* errors[a] index out of bounds */
    err_idx *indices = (err_idx *)malloc(sizeof(err_idx) * (sortable_items + 1));
    for (size_t i = 0; i < sortable_items; i++) {
        indices[i].idx = (uint16_t)i;
        indices[i].err = errors[(uint16_t)i];
    }
    qsort(indices, sortable_items, sizeof(err_idx), cmp_err_idx);
    size_t counts[4];
    vbr_interpolate_distribution(sortable_items, target, counts);
    size_t minus_one = counts[0], plus_one = counts[2], plus_two = counts[3];
    uint8_t base = f32_as_u8(target);
    memset(residual_sizes, base, n_errors);
    for (size_t i = 0; i < minus_one && i < sortable_items; i++) residual_sizes[indices[i].idx] = (uint8_t)(base - 1);
    if (plus_two + plus_one > sortable_items) {
        free(indices);
        return ORACLE_ERR_PANIC;
    }
    size_t s1 = sortable_items - plus_two - plus_one;
    for (size_t i = 0; i < plus_one; i++) residual_sizes[indices[s1 + i].idx] = (uint8_t)(base + 1);
    size_t s2 = sortable_items - plus_two;
    for (size_t i = 0; i < plus_two; i++) residual_sizes[indices[s2 + i].idx] = (uint8_t)(base + 2);
    if (ties) {
        size_t bounds[3] = {minus_one, s1, s2};
        for (int k = 0; k < 3; k++) {
            size_t p = bounds[k];
            if (p > 0 && p < sortable_items && indices[p - 1].err == indices[p].err) {
                /* a tie only matters if the two sides get different sizes */
                if (residual_sizes[indices[p - 1].idx] != residual_sizes[indices[p].idx]) (*ties)++;
            }
        }
    }
    free(indices);
    return 0;
}

static int valid_residual_size(int v) { return v >= 1 && v <= 8; }

/* encoder_vbr.rs:139-171 (restores lms only: trap T2) */
static int vbr_analyze(encoder_base *e, float target, int scale_factor_frames, const int16_t *samples, size_t n,
                       uint8_t **sizes_out, size_t *n_sizes, uint64_t *ties)
{
    int analyze_size = (int)f32_as_u8(target) + 1;
    if (!valid_residual_size(analyze_size)) return ORACLE_ERR_PANIC; /* SeaResidualSize::from panics */
    size_t channels = (size_t)e->channels;
    size_t slice_size = (size_t)scale_factor_frames * channels;
    sea_lms *original_lms = (sea_lms *)malloc(sizeof(sea_lms) * channels);
    memcpy(original_lms, e->lms, sizeof(sea_lms) * channels);
    uint8_t *sizes = (uint8_t *)malloc(channels);
    memset(sizes, analyze_size, channels);
    uint8_t *scale_factors = (uint8_t *)calloc(slice_size + 1, 1);
    uint8_t *residuals = (uint8_t *)calloc(slice_size + 1, 1);
    size_t n_errors = div_ceil(n / channels, (size_t)scale_factor_frames) * channels;
    uint64_t *errors = (uint64_t *)calloc(n_errors + 1, sizeof(uint64_t));
    size_t slice_index = 0;
    for (size_t pos = 0; pos < n; pos += slice_size, slice_index++) {
        size_t len = n - pos < slice_size ? n - pos : slice_size;
        eb_residuals_for_block(e, samples + pos, len, sizes, scale_factors, residuals, errors + slice_index * channels);
    }
    memcpy(e->lms, original_lms, sizeof(sea_lms) * channels);
    uint8_t *residual_sizes = (uint8_t *)malloc(n_errors + 1);
    int rc = vbr_choose_residual_len(target, scale_factor_frames, n, errors, n_errors, residual_sizes, ties);
    free(original_lms);
    free(sizes);
    free(scale_factors);
    free(residuals);
    free(errors);
    if (rc) {
        free(residual_sizes);
        return rc;
    }
    *sizes_out = residual_sizes;
    *n_sizes = n_errors;
    return 0;
}

/* encoder_vbr.rs:175-214 */
static int vbr_encode(encoder_base *e, float target, int scale_factor_frames, const int16_t *samples, size_t n,
                      encoded_samples *out, uint64_t *ties)
{
    size_t channels = (size_t)e->channels;
    memset(out, 0, sizeof(*out));
    int rc = vbr_analyze(e, target, scale_factor_frames, samples, n, &out->residual_bits, &out->n_residual_bits, ties);
    if (rc) return rc;
    out->n_scale_factors = div_ceil(n / channels, (size_t)scale_factor_frames) * channels;
    out->scale_factors = (uint8_t *)calloc(out->n_scale_factors + 1, 1);
    out->n_residuals = n;
    out->residuals = (uint8_t *)calloc(n + 1, 1);
    size_t slice_size = (size_t)scale_factor_frames * channels;
    uint8_t *sizes = (uint8_t *)malloc(channels);
    uint64_t *ranks = (uint64_t *)calloc(channels, sizeof(uint64_t));
    size_t slice_index = 0;
    for (size_t pos = 0; pos < n; pos += slice_size, slice_index++) {
        size_t len = n - pos < slice_size ? n - pos : slice_size;
        for (size_t c = 0; c < channels; c++) {
            sizes[c] = out->residual_bits[slice_index * channels + c];
            if (!valid_residual_size(sizes[c])) { /* SeaResidualSize::from panics (trap T20) */
                free(sizes);
                free(ranks);
                es_free(out);
                return ORACLE_ERR_PANIC;
            }
        }
        eb_residuals_for_block(e, samples + pos, len, sizes, out->scale_factors + slice_index * channels,
                               out->residuals + slice_index * slice_size, ranks);
    }
    free(sizes);
    free(ranks);
    return 0;
}

/* ------------------------------------------------------------------ chunk.rs (serialise) */

/* chunk.rs:215-292 */
static int chunk_serialize(int channels, int frames_per_chunk, const oracle_settings *st, const sea_lms *lms,
                           const encoded_samples *es, bytevec *out)
{
    int is_vbr = es->n_residual_bits != 0; /* chunk.rs:46-51 */
    int residual_size = (int)f32_as_u8(floorf(st->residual_bits));
    if (!valid_residual_size(residual_size)) return ORACLE_ERR_PANIC; /* chunk.rs:60 */
    if (st->scale_factor_bits == 0 || st->scale_factor_frames == 0 ||
        frames_per_chunk % st->scale_factor_frames != 0)
        return ORACLE_ERR_PANIC; /* chunk.rs:216-218 asserts */
    bv_push(out, is_vbr ? 0x02 : 0x01);
    bv_push(out, (uint8_t)((st->scale_factor_bits << 4) | residual_size));
    bv_push(out, st->scale_factor_frames);
    bv_push(out, 0x5A);
    for (int c = 0; c < channels; c++) {
        uint8_t b[16];
        lms_serialize(&lms[c], b);
        bv_extend(out, b, 16);
    }
    bit_packer p;
    memset(&p, 0, sizeof(p));
    for (size_t i = 0; i < es->n_scale_factors; i++) packer_push(&p, es->scale_factors[i], st->scale_factor_bits);
    packer_finish(&p, out);
    if (is_vbr) {
        for (size_t i = 0; i < es->n_residual_bits; i++) {
            int32_t relative = (int32_t)es->residual_bits[i] - residual_size + 1;
            packer_push(&p, (uint32_t)relative, 2);
        }
        packer_finish(&p, out);
        size_t vbr_index = 0;
        int frames_written = 0;
        size_t frames = es->n_residuals / (size_t)channels; /* chunks_exact */
        for (size_t f = 0; f < frames; f++) {
            for (int c = 0; c < channels; c++)
                packer_push(&p, es->residuals[f * (size_t)channels + (size_t)c], es->residual_bits[vbr_index + (size_t)c]);
            frames_written++;
            if (frames_written == st->scale_factor_frames) {
                vbr_index += (size_t)channels;
                frames_written = 0;
            }
        }
    } else {
        for (size_t i = 0; i < es->n_residuals; i++) packer_push(&p, es->residuals[i], (uint8_t)residual_size);
    }
    packer_finish(&p, out);
    return 0;
}

/* ------------------------------------------------------------------ file.rs + encoder.rs: streaming encoder object */

typedef struct {
    sea_file_header header;
    oracle_settings settings;
    encoder_base base;
    float vbr_target;
    int state; /* 0 Start, 1 WritingFrames, 2 Finished  (encoder.rs:10-14) */
    uint32_t written_frames;
    uint64_t ties;
} oracle_encoder;

/* encoder.rs:50-86 + file.rs:111-129 */
oracle_encoder *oracle_encoder_new(uint8_t channels, uint32_t sample_rate, int has_total, uint32_t total_frames,
                                   const oracle_settings *settings, uint8_t *hdr_out, size_t *hdr_len)
{
    oracle_encoder *enc = (oracle_encoder *)calloc(1, sizeof(oracle_encoder));
    enc->header.version = 1;
    enc->header.channels = channels;
    enc->header.chunk_size = 0;
    enc->header.frames_per_chunk = settings->frames_per_chunk;
    enc->header.sample_rate = sample_rate;
    enc->header.total_frames = has_total ? total_frames : 0;
    enc->settings = *settings;
    eb_init(&enc->base, channels, settings->scale_factor_bits);
    enc->vbr_target = vbr_normalized_bitrate(settings);
    enc->state = 0;
    *hdr_len = 0;
    if (has_total && total_frames == 0) {
        bytevec h = {0};
        header_serialize(&enc->header, &h);
        memcpy(hdr_out, h.data, h.len);
        *hdr_len = h.len;
        bv_free(&h);
        enc->state = 1;
    }
    return enc;
}

void oracle_encoder_free(oracle_encoder *enc)
{
    if (!enc) return;
    eb_free(&enc->base);
    free(enc);
}

/* file.rs:142-178 */
static int file_make_chunk(oracle_encoder *enc, const int16_t *samples, size_t n, bytevec *out)
{
    int channels = enc->header.channels;
    sea_lms *initial_lms = (sea_lms *)malloc(sizeof(sea_lms) * (size_t)channels);
    memcpy(initial_lms, enc->base.lms, sizeof(sea_lms) * (size_t)channels);
    encoded_samples es;
    int rc = 0;
    if (enc->settings.vbr) {
        rc = vbr_encode(&enc->base, enc->vbr_target, enc->settings.scale_factor_frames, samples, n, &es, &enc->ties);
    } else {
        int residual_size = (int)f32_as_u8(floorf(enc->settings.residual_bits));
        if (!valid_residual_size(residual_size)) rc = ORACLE_ERR_PANIC; /* encoder_cbr.rs:22 (at construction) */
        else cbr_encode(&enc->base, residual_size, enc->settings.scale_factor_frames, samples, n, &es);
    }
    if (rc) {
        free(initial_lms);
        return rc;
    }
    size_t before = out->len;
    rc = chunk_serialize(channels, enc->header.frames_per_chunk, &enc->settings, initial_lms, &es, out);
    es_free(&es);
    free(initial_lms);
    if (rc) return rc;
    size_t produced = out->len - before;
    if (enc->header.chunk_size == 0) enc->header.chunk_size = (uint16_t)produced;
    size_t full_samples_len = (size_t)enc->header.frames_per_chunk * (size_t)channels;
    if (n == full_samples_len && enc->header.chunk_size != (uint16_t)produced) return ORACLE_ERR_PANIC; /* file.rs:173-175 */
    return 0;
}

/*
 * encoder.rs:106-149.  `avail` = samples the reader still holds.  Consumes up to one chunk from `samples`,
 * appends (header once +) chunk bytes to out.  Returns 1 = more, 0 = eof, <0 = error.  *consumed = samples read.
 */
int oracle_encoder_encode_frame(oracle_encoder *enc, const int16_t *samples, size_t avail, uint8_t *out, size_t out_cap,
                                size_t *out_len, size_t *consumed)
{
    *out_len = 0;
    *consumed = 0;
    if (enc->state == 2) return ORACLE_ERR_CLOSED;
    size_t channels = enc->header.channels;
    size_t frames;
    if (enc->header.total_frames > 0) {
        size_t left = (size_t)enc->header.total_frames - (size_t)enc->written_frames;
        frames = enc->header.frames_per_chunk < left ? enc->header.frames_per_chunk : left;
    } else {
        frames = enc->header.frames_per_chunk;
    }
    size_t full_size_samples = (size_t)enc->header.frames_per_chunk * channels;
    size_t samples_to_read = frames * channels;
    size_t got = avail < samples_to_read ? avail : samples_to_read; /* read_max_or_zero */
    if ((got * 2) % (2 * channels) != 0) return ORACLE_ERR_READ;    /* encoder.rs:95-99 */
    int eof = got == 0 || got < full_size_samples;
    bytevec buf = {0};
    if (got != 0) {
        bytevec chunk = {0};
        int rc = file_make_chunk(enc, samples, got, &chunk);
        if (rc) {
            bv_free(&chunk);
            return rc;
        }
        if (eof) {
            if (chunk.len > enc->header.chunk_size) { bv_free(&chunk); return ORACLE_ERR_PANIC; }
        } else if (chunk.len != enc->header.chunk_size) { bv_free(&chunk); return ORACLE_ERR_PANIC; }
        if (enc->state == 0) {
            header_serialize(&enc->header, &buf);
            enc->state = 1;
        }
        bv_extend(&buf, chunk.data, chunk.len);
        bv_free(&chunk);
        enc->written_frames += (uint32_t)frames;
        *consumed = got;
    }
    if (eof) enc->state = 2;
    if (buf.len > out_cap) {
        bv_free(&buf);
        return ORACLE_ERR_CAPACITY;
    }
    if (buf.len) memcpy(out, buf.data, buf.len);
    *out_len = buf.len;
    bv_free(&buf);
    return eof ? 0 : 1;
}

uint64_t oracle_encoder_ties(const oracle_encoder *enc) { return enc->ties; }

/* lib.rs:13-36.  Returns bytes written (>=0) or a negative error. *ties = VBR boundary ties (T13). */
int64_t oracle_sea_encode(const int16_t *samples, size_t n_samples, uint32_t sample_rate, uint32_t channels,
                          const oracle_settings *settings, uint8_t *out, size_t out_cap, uint64_t *ties)
{
    if (ties) *ties = 0;
    if (channels == 0) return ORACLE_ERR_PANIC; /* division by zero in lib.rs:25 */
    /* outside the reference's working domain (powf(inf) tables, division by zero): refuse instead of UB */
    if (settings->scale_factor_bits < 1 || settings->scale_factor_bits > 8 || settings->scale_factor_frames == 0 ||
        settings->frames_per_chunk == 0 || channels > 255)
        return ORACLE_ERR_PANIC;
    uint8_t hdr[32];
    size_t hdr_len = 0, total = 0;
    oracle_encoder *enc = oracle_encoder_new((uint8_t)channels, sample_rate, 1, (uint32_t)(n_samples / channels) /* as u32 */,
                                             settings, hdr, &hdr_len);
    if (hdr_len) {
        if (hdr_len > out_cap) { oracle_encoder_free(enc); return ORACLE_ERR_CAPACITY; }
        memcpy(out, hdr, hdr_len);
        total = hdr_len;
    }
    size_t pos = 0;
    for (;;) {
        size_t got = 0, used = 0;
        int rc = oracle_encoder_encode_frame(enc, samples + pos, n_samples - pos, out + total, out_cap - total, &got, &used);
        if (rc < 0) {
            oracle_encoder_free(enc);
            return rc;
        }
        total += got;
        pos += used;
        if (rc == 0) break;
    }
    if (ties) *ties = enc->ties;
    oracle_encoder_free(enc);
    return (int64_t)total;
}

/* ------------------------------------------------------------------ chunk.rs (parse) + codec/decoder.rs */

typedef struct {
    int channels, frames_per_chunk, chunk_type, scale_factor_bits, scale_factor_frames, residual_size;
    sea_lms *lms;
    bytevec scale_factors, vbr_residual_sizes, residuals;
} sea_chunk;

static void chunk_free(sea_chunk *c)
{
    free(c->lms);
    bv_free(&c->scale_factors);
    bv_free(&c->vbr_residual_sizes);
    bv_free(&c->residuals);
}

#define NEED(n)                                             \
    do {                                                    \
        if (encoded_index + (size_t)(n) > len) {            \
            chunk_free(c);                                  \
            return ORACLE_ERR_PANIC; /* slice OOB panic */  \
        }                                                   \
    } while (0)

/* chunk.rs:69-213; remaining_frames < 0 means None */
static int chunk_from_slice(const uint8_t *encoded, size_t len, const sea_file_header *h, int64_t remaining_frames,
                            sea_chunk *c)
{
    memset(c, 0, sizeof(*c));
    if (len > h->chunk_size) return ORACLE_ERR_PANIC;
    if (remaining_frames < 0 && len < h->chunk_size) return ORACLE_ERR_INVALID_FRAME;
    if (len < 4) return ORACLE_ERR_PANIC;
    if (encoded[0] != 0x01 && encoded[0] != 0x02) return ORACLE_ERR_INVALID_FRAME;
    c->chunk_type = encoded[0];
    c->scale_factor_bits = encoded[1] >> 4;
    c->residual_size = encoded[1] & 0x0f;
    if (!valid_residual_size(c->residual_size)) return ORACLE_ERR_PANIC;
    c->scale_factor_frames = encoded[2];
    c->channels = h->channels;
    c->frames_per_chunk = h->frames_per_chunk;
    size_t encoded_index = 4;
    size_t channels = h->channels;
    c->lms = (sea_lms *)malloc(sizeof(sea_lms) * channels);
    for (size_t ch = 0; ch < channels; ch++) {
        NEED(16);
        lms_from_bytes(&c->lms[ch], encoded + encoded_index);
        encoded_index += 16;
    }
    size_t frames_in_this_chunk = h->frames_per_chunk;
    if (remaining_frames >= 0 && (size_t)remaining_frames < frames_in_this_chunk) frames_in_this_chunk = (size_t)remaining_frames;
    if (c->scale_factor_frames == 0) { chunk_free(c); return ORACLE_ERR_PANIC; } /* div by zero */
    size_t scale_factor_items = div_ceil(frames_in_this_chunk, (size_t)c->scale_factor_frames) * channels;
    {
        size_t packed = div_ceil(scale_factor_items * (size_t)c->scale_factor_bits, 8);
        NEED(packed);
        if (c->scale_factor_bits == 0 || c->scale_factor_bits > 8) { chunk_free(c); return ORACLE_ERR_PANIC; }
        unpack_const((uint8_t)c->scale_factor_bits, encoded + encoded_index, packed, &c->scale_factors);
        encoded_index += packed;
        bv_resize(&c->scale_factors, scale_factor_items);
    }
    if (c->chunk_type == 0x02) {
        size_t packed = div_ceil(scale_factor_items * 2, 8);
        NEED(packed);
        unpack_const(2, encoded + encoded_index, packed, &c->vbr_residual_sizes);
        encoded_index += packed;
        bv_resize(&c->vbr_residual_sizes, scale_factor_items);
        for (size_t i = 0; i < c->vbr_residual_sizes.len; i++)
            c->vbr_residual_sizes.data[i] = (uint8_t)(c->vbr_residual_sizes.data[i] + c->residual_size - 1);
    }
    {
        size_t packed;
        bytevec bitlengths = {0};
        if (c->chunk_type == 0x02) {
            for (size_t blk = 0; blk + channels <= c->vbr_residual_sizes.len; blk += channels)
                for (int f = 0; f < c->scale_factor_frames; f++)
                    for (size_t ch = 0; ch < channels; ch++) bv_push(&bitlengths, c->vbr_residual_sizes.data[blk + ch]);
            if (c->vbr_residual_sizes.len < channels) { bv_free(&bitlengths); chunk_free(c); return ORACLE_ERR_PANIC; }
            uint32_t residual_bits = 0;
            size_t n = c->vbr_residual_sizes.len;
            for (size_t i = 0; i < n - channels; i++) residual_bits += c->vbr_residual_sizes.data[i];
            residual_bits *= (uint32_t)c->scale_factor_frames;
            uint32_t last_frame_samples = (uint32_t)frames_in_this_chunk % (uint32_t)c->scale_factor_frames;
            uint32_t multiplier = last_frame_samples == 0 ? (uint32_t)c->scale_factor_frames : last_frame_samples;
            for (size_t i = n - channels; i < n; i++) residual_bits += c->vbr_residual_sizes.data[i] * multiplier;
            packed = (residual_bits + 7) / 8;
        } else {
            packed = div_ceil(frames_in_this_chunk * (size_t)c->residual_size * channels, 8);
        }
        if (encoded_index + packed > len) { bv_free(&bitlengths); chunk_free(c); return ORACLE_ERR_PANIC; }
        if (c->chunk_type == 0x02) {
            for (size_t i = 0; i < bitlengths.len; i++)
                if (bitlengths.data[i] > 8) { bv_free(&bitlengths); chunk_free(c); return ORACLE_ERR_PANIC; } /* MASKS OOB */
            unpack_process(bitlengths.data, bitlengths.len, encoded + encoded_index, packed, &c->residuals);
        } else {
            uint8_t b = (uint8_t)c->residual_size;
            unpack_process(&b, 1, encoded + encoded_index, packed, &c->residuals);
        }
        bv_free(&bitlengths);
        bv_resize(&c->residuals, frames_in_this_chunk * channels);
    }
    return 0;
}

/* codec/decoder.rs:20-50 and 52-86 */
static int decode_chunk(const sea_dequant_tab *tab, const sea_chunk *c, int16_t *out)
{
    if (c->scale_factor_bits != tab->scale_factor_bits) return ORACLE_ERR_PANIC; /* assert_eq */
    size_t channels = (size_t)c->channels;
    sea_lms *lms = (sea_lms *)malloc(sizeof(sea_lms) * channels);
    memcpy(lms, c->lms, sizeof(sea_lms) * channels);
    size_t block = channels * (size_t)c->scale_factor_frames;
    size_t n = c->residuals.len, o = 0;
    int rc = 0;
    for (size_t pos = 0, sfi = 0; pos < n && !rc; pos += block, sfi++) {
        size_t len = n - pos < block ? n - pos : block;
        const uint8_t *scale_factors = c->scale_factors.data + sfi * channels;
        const uint8_t *vbr = c->chunk_type == 0x02 ? c->vbr_residual_sizes.data + sfi * channels : NULL;
        for (size_t i = 0; i < len; i++) {
            size_t ch = i % channels;
            int size = vbr ? vbr[ch] : c->residual_size;
            if (size < 1 || size > 8) { rc = ORACLE_ERR_PANIC; break; }
            int sf = scale_factors[ch];
            int q = c->residuals.data[pos + i];
            if (sf >= tab->sf_items || q >= (1 << size)) { rc = ORACLE_ERR_PANIC; break; }
            int32_t predicted = lms_predict(&lms[ch]);
            int32_t dequantized = tab->dqt[size][sf * (1 << size) + q];
            int16_t reconstructed = clamp_i16((int32_t)((uint32_t)predicted + (uint32_t)dequantized));
            out[o++] = reconstructed;
            lms_update(&lms[ch], reconstructed, dequantized);
        }
    }
    free(lms);
    return rc;
}

/* file.rs:40-72 (metadata bytes are never consumed: Vec::with_capacity has len 0, file.rs:53-54) */
int oracle_parse_header(const uint8_t *enc, size_t len, sea_file_header *h)
{
    if (len < 22) return ORACLE_ERR_READ;
    if (!(enc[0] == 's' && enc[1] == 'e' && enc[2] == 'a' && enc[3] == 'c')) return ORACLE_ERR_INVALID_FILE;
    h->version = enc[4];
    h->channels = enc[5];
    h->chunk_size = (uint16_t)(enc[6] | (enc[7] << 8));
    h->frames_per_chunk = (uint16_t)(enc[8] | (enc[9] << 8));
    h->sample_rate = (uint32_t)enc[10] | ((uint32_t)enc[11] << 8) | ((uint32_t)enc[12] << 16) | ((uint32_t)enc[13] << 24);
    h->total_frames = (uint32_t)enc[14] | ((uint32_t)enc[15] << 8) | ((uint32_t)enc[16] << 16) | ((uint32_t)enc[17] << 24);
    if (!(h->channels > 0 && h->chunk_size >= 16 && h->frames_per_chunk > 0 && h->sample_rate > 0))
        return ORACLE_ERR_INVALID_FILE;
    return 0;
}

/* lib.rs:44-63 + decoder.rs:33-59 + file.rs:180-209. Returns 0 or negative error; *n_out = samples written. */
int oracle_sea_decode(const uint8_t *enc, size_t len, int16_t *out, size_t out_cap, size_t *n_out, uint32_t *sample_rate,
                      uint32_t *channels)
{
    sea_file_header h;
    *n_out = 0;
    int rc = oracle_parse_header(enc, len, &h);
    if (rc) return rc;
    *sample_rate = h.sample_rate;
    *channels = h.channels;
    size_t pos = 22, frames_read = 0, o = 0;
    sea_dequant_tab tab;
    int have_tab = 0;
    for (;;) {
        if (h.total_frames != 0 && (size_t)h.total_frames <= frames_read) break;
        int64_t remaining = h.total_frames > 0 ? (int64_t)((size_t)h.total_frames - frames_read) : -1;
        size_t take = len - pos < h.chunk_size ? len - pos : h.chunk_size; /* read_max_or_zero */
        if (take == 0) break;
        sea_chunk c;
        rc = chunk_from_slice(enc + pos, take, &h, remaining, &c);
        if (rc) break;
        pos += take;
        if (!have_tab) {
            if (c.scale_factor_bits < 1 || c.scale_factor_bits > 8) { chunk_free(&c); rc = ORACLE_ERR_PANIC; break; }
            dqt_init(&tab, c.scale_factor_bits);
            have_tab = 1;
        }
        if (o + c.residuals.len > out_cap) { chunk_free(&c); rc = ORACLE_ERR_CAPACITY; break; }
        rc = decode_chunk(&tab, &c, out + o);
        if (rc) { chunk_free(&c); break; }
        o += c.residuals.len;
        frames_read += c.residuals.len / h.channels;
        chunk_free(&c);
    }
    if (have_tab) dqt_free(&tab);
    *n_out = o;
    return rc;
}

/* ------------------------------------------------------------------ table / parameter getters (known-answer tests) */

void oracle_scale_factors(int residual_bits, int scale_factor_bits, int32_t *out) { dqt_scale_factors(residual_bits, scale_factor_bits, out); }

void oracle_tables(int residual_bits, int scale_factor_bits, int32_t *recip_out, int32_t *dqt_out)
{
    sea_dequant_tab t;
    dqt_init(&t, scale_factor_bits);
    memcpy(recip_out, t.recip[residual_bits], sizeof(int32_t) * (size_t)t.sf_items);
    memcpy(dqt_out, t.dqt[residual_bits], sizeof(int32_t) * (size_t)t.sf_items * ((size_t)1 << residual_bits));
    dqt_free(&t);
}

/* quant table for residual_bits as the (2^(b+1)+1)-entry slice indexed by clamped + 2^b */
void oracle_quant_tab(int residual_bits, uint8_t *out)
{
    sea_quant_tab q;
    qt_init(&q);
    size_t items = ((size_t)1 << (residual_bits + 1)) + 1;
    memcpy(out, &q.quant_tab[q.offsets[residual_bits]], items);
}

void oracle_vbr_params(const oracle_settings *s, size_t items, float *target, int *base, size_t counts[4])
{
    *target = vbr_normalized_bitrate(s);
    *base = f32_as_u8(*target);
    vbr_interpolate_distribution(items, *target, counts);
}

int32_t oracle_sea_div(int32_t v, int32_t recip) { return sea_div(v, (int64_t)recip); }

/* ------------------------------------------------------------------ CPU baseline harness (bench.py cpu_baseline / --impl reference) */

typedef struct {
    int mode; /* 0 encode, 1 decode */
    const int16_t *pcm;
    size_t n_samples;
    uint32_t rate, channels;
    const oracle_settings *settings;
    const uint8_t *sea;
    size_t sea_len;
    int reps;
    int64_t result;
} bench_job;

static void *bench_worker(void *arg)
{
    bench_job *j = (bench_job *)arg;
    if (j->mode == 0) {
        size_t cap = j->n_samples * 2 + 4096;
        uint8_t *out = (uint8_t *)malloc(cap);
        for (int r = 0; r < j->reps; r++) j->result = oracle_sea_encode(j->pcm, j->n_samples, j->rate, j->channels, j->settings, out, cap, NULL);
        free(out);
    } else {
        int16_t *out = (int16_t *)malloc(j->n_samples * 2 + 64);
        size_t n;
        uint32_t r_, c_;
        for (int r = 0; r < j->reps; r++) {
            int rc = oracle_sea_decode(j->sea, j->sea_len, out, j->n_samples + 32, &n, &r_, &c_);
            j->result = rc ? rc : (int64_t)n;
        }
        free(out);
    }
    return NULL;
}

/*
 * Runs `reps` encodes (mode 0) or decodes (mode 1) of ONE stream on each of `threads` threads concurrently
 * (one stream per core) and returns wall seconds; *samples_done = threads*reps*n_samples.
 */
double oracle_bench(int mode, int threads, int reps, const int16_t *pcm, size_t n_samples, uint32_t rate, uint32_t channels,
                    const oracle_settings *settings, const uint8_t *sea, size_t sea_len, uint64_t *samples_done)
{
    pthread_t *tids = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    bench_job *jobs = (bench_job *)calloc((size_t)threads, sizeof(bench_job));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; t++) {
        jobs[t].mode = mode;
        jobs[t].pcm = pcm;
        jobs[t].n_samples = n_samples;
        jobs[t].rate = rate;
        jobs[t].channels = channels;
        jobs[t].settings = settings;
        jobs[t].sea = sea;
        jobs[t].sea_len = sea_len;
        jobs[t].reps = reps;
        pthread_create(&tids[t], NULL, bench_worker, &jobs[t]);
    }
    int64_t ok = 1;
    for (int t = 0; t < threads; t++) {
        pthread_join(tids[t], NULL);
        if (jobs[t].result < 0) ok = 0;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *samples_done = ok ? (uint64_t)threads * (uint64_t)reps * (uint64_t)n_samples : 0;
    free(tids);
    free(jobs);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
