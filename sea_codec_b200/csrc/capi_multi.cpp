// capi_multi.cpp -- in-process multi-GPU batch entry points (include/sea_b200.h: sea_b200_multi_*).
//
// BASELINE north star / SURVEY 8e: "work is partitioned across the 8 GPUs of one B200 box by independent streams ... no NCCL
// collectives because nothing is reduced across GPUs; only per-GPU output byte counts are gathered on the host".  That is all
// this file does: one context (one CUDA stream set) and one host thread per GPU, a contiguous range of the batch's streams
// per GPU balanced by work, the single-GPU host-buffer calls of capi.cu on every range, counts gathered when the threads
// join.  No peer access, no device-to-device traffic.  Nothing here touches a kernel; there is still no CPU codec path.
#include <stdint.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sea_b200.h"

struct sea_b200_multi {
    std::vector<int> devices;
    std::vector<sea_b200_ctx *> ctxs;
    std::string last_error;
};

namespace {

// Contiguous ranges [cut[d], cut[d+1]) of n items with weights w, as even as a prefix split allows.
std::vector<uint32_t> split_by_work(uint32_t n, const std::vector<uint64_t> &w, uint32_t parts)
{
    std::vector<uint32_t> cut(parts + 1, n);
    cut[0] = 0;
    uint64_t total = 0;
    for (uint64_t v : w) total += v;
    uint64_t acc = 0;
    uint32_t i = 0;
    for (uint32_t d = 1; d < parts; d++) {
        const uint64_t target = total / parts * d + (total % parts) * d / parts;
        while (i < n && acc + w[i] / 2 < target) acc += w[i++];
        cut[d] = i;
    }
    if (total == 0)  // no work at all: split by count so that empty streams still get their headers somewhere
        for (uint32_t d = 1; d < parts; d++) cut[d] = (uint32_t)((uint64_t)n * d / parts);
    return cut;
}

template <typename F>
int run_sharded(sea_b200_multi *m, const std::vector<uint32_t> &cut, F &&call)
{
    const uint32_t parts = (uint32_t)m->ctxs.size();
    std::vector<int> rc(parts, SEA_B200_OK);
    std::vector<std::thread> th;
    for (uint32_t d = 0; d < parts; d++)
        th.emplace_back([&, d] { rc[d] = cut[d + 1] > cut[d] ? call(d, cut[d], cut[d + 1] - cut[d]) : SEA_B200_OK; });
    for (auto &t : th) t.join();
    for (uint32_t d = 0; d < parts; d++)
        if (rc[d] != SEA_B200_OK) {
            m->last_error = "GPU " + std::to_string(m->devices[d]) + ": " + sea_b200_last_error(m->ctxs[d]);
            return rc[d];
        }
    return SEA_B200_OK;
}

}  // namespace

extern "C" {

int sea_b200_multi_create(const int *devices, uint32_t n_devices, sea_b200_multi **out)
{
    if (!out || n_devices == 0 || n_devices > 64) return SEA_B200_ERR_INVALID_PARAMETERS;
    *out = nullptr;
    sea_b200_multi *m = new sea_b200_multi();
    for (uint32_t d = 0; d < n_devices; d++) {
        sea_b200_ctx *c = nullptr;
        const int dev = devices ? devices[d] : (int)d;
        const int rc = sea_b200_ctx_create(dev, &c);
        if (rc != SEA_B200_OK) {
            sea_b200_multi_destroy(m);
            return rc;
        }
        m->devices.push_back(dev);
        m->ctxs.push_back(c);
    }
    *out = m;
    return SEA_B200_OK;
}

void sea_b200_multi_destroy(sea_b200_multi *m)
{
    if (!m) return;
    for (sea_b200_ctx *c : m->ctxs) sea_b200_ctx_destroy(c);
    delete m;
}

uint32_t sea_b200_multi_device_count(const sea_b200_multi *m) { return m ? (uint32_t)m->ctxs.size() : 0; }
sea_b200_ctx *sea_b200_multi_ctx(const sea_b200_multi *m, uint32_t index) { return m && index < m->ctxs.size() ? m->ctxs[index] : nullptr; }
const char *sea_b200_multi_last_error(const sea_b200_multi *m) { return m ? m->last_error.c_str() : ""; }

int sea_b200_multi_encode_batch(sea_b200_multi *m, uint32_t n_streams, const int16_t *pcm, const uint64_t *pcm_offsets, const uint32_t *n_frames,
                                uint32_t sample_rate, uint32_t channels, const sea_b200_settings *settings, uint8_t *out,
                                const uint64_t *out_offsets, uint64_t *out_lens, uint32_t *first_stream_of_device, uint64_t *bytes_per_device)
{
    if (!m || !pcm_offsets || !n_frames || !settings || !out || !out_offsets || !out_lens) return SEA_B200_ERR_INVALID_PARAMETERS;
    const uint32_t parts = (uint32_t)m->ctxs.size();
    std::vector<uint64_t> w(n_streams);
    for (uint32_t i = 0; i < n_streams; i++) w[i] = (uint64_t)n_frames[i] * channels + 1u;
    const std::vector<uint32_t> cut = split_by_work(n_streams, w, parts);
    const int rc = run_sharded(m, cut, [&](uint32_t d, uint32_t first, uint32_t count) {
        return sea_b200_encode_batch(m->ctxs[d], count, pcm, pcm_offsets + first, n_frames + first, sample_rate, channels, settings, out,
                                     out_offsets + first, out_lens + first);
    });
    for (uint32_t d = 0; d < parts; d++) {
        if (first_stream_of_device) first_stream_of_device[d] = cut[d];
        if (bytes_per_device) {
            bytes_per_device[d] = 0;
            if (rc == SEA_B200_OK)
                for (uint32_t i = cut[d]; i < cut[d + 1]; i++) bytes_per_device[d] += out_lens[i];
        }
    }
    return rc;
}

int sea_b200_multi_decode_batch(sea_b200_multi *m, uint32_t n_streams, const uint8_t *sea, const uint64_t *sea_offsets, const uint64_t *sea_lens,
                                int16_t *pcm, const uint64_t *pcm_offsets, const uint64_t *pcm_caps, uint64_t *n_samples,
                                uint32_t *first_stream_of_device, uint64_t *samples_per_device)
{
    if (!m || !sea || !sea_offsets || !sea_lens || !pcm_offsets || !n_samples) return SEA_B200_ERR_INVALID_PARAMETERS;
    const uint32_t parts = (uint32_t)m->ctxs.size();
    std::vector<uint64_t> w(n_streams);
    for (uint32_t i = 0; i < n_streams; i++) w[i] = sea_lens[i] + 1u;
    const std::vector<uint32_t> cut = split_by_work(n_streams, w, parts);
    const int rc = run_sharded(m, cut, [&](uint32_t d, uint32_t first, uint32_t count) {
        return sea_b200_decode_batch(m->ctxs[d], count, sea, sea_offsets + first, sea_lens + first, pcm, pcm_offsets + first,
                                     pcm_caps ? pcm_caps + first : nullptr, n_samples + first);
    });
    for (uint32_t d = 0; d < parts; d++) {
        if (first_stream_of_device) first_stream_of_device[d] = cut[d];
        if (samples_per_device) {
            samples_per_device[d] = 0;
            if (rc == SEA_B200_OK)
                for (uint32_t i = cut[d]; i < cut[d + 1]; i++) samples_per_device[d] += n_samples[i];
        }
    }
    return rc;
}

}  // extern "C"
