// sea_format.h -- host-side .sea container arithmetic: file header, chunk geometry, VBR plan, table generation.
// This is the scalar part of the path that stays on the host; the kernels read and
// write the wire layout directly.
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/sea_b200.h"
#include "sea_common.cuh"

namespace sea {

// ---- file header: file.rs:40-93 ------------------------------------------------------------------------
int parse_file_header(const uint8_t *p, uint64_t len, sea_b200_header *h);
void write_file_header(uint8_t *p, uint8_t channels, uint16_t chunk_size, uint16_t frames_per_chunk, uint32_t sample_rate,
                       uint32_t total_frames);

// ---- settings ------------------------------------------------------------------------------------------
// Resolved, validated view of EncoderSettings for one (channels) configuration.
struct EncodePlan {
    uint32_t channels, N, F, s;  // frames_per_chunk, scale_factor_frames, scale_factor_bits
    uint32_t hdr_bits;           // floor(residual_bits): chunk header nibble (chunk.rs:60)
    bool vbr;
    float vbr_target;            // encoder_vbr.rs:40-63
    uint32_t base;               // (u8)vbr_target
    uint32_t full_counts[4];     // interpolate_distribution(N*C/F) -> [base-1, base, base+1, base+2]
    bool full_chunk_valid;       // false when a full chunk would get a size outside 1..8 (the reference panics)
    uint32_t full_chunk_bytes;   // header.chunk_size of a full chunk
    uint32_t max_chunk_bytes;    // upper bound over partial chunks too (smem sizing)
};

// Returns SEA_B200_OK or INVALID_PARAMETERS / DOMAIN (the reference would panic).
int make_encode_plan(uint32_t channels, const sea_b200_settings *st, EncodePlan *plan);

float vbr_normalized_bitrate(const sea_b200_settings *st);                       // encoder_vbr.rs:40-63
void vbr_distribution(uint64_t items, float target, uint64_t counts[4]);         // encoder_vbr.rs:66-96

// chunk bytes of a CBR chunk with `frames` frames (chunk.rs:215-292; SURVEY App. D formula)
inline uint32_t cbr_chunk_bytes(uint32_t frames, uint32_t C, uint32_t s, uint32_t F, uint32_t b)
{
    uint32_t items = div_ceil_u32(frames, F) * C;
    return 4u + 16u * C + div_ceil_u32(items * s, 8u) + (uint32_t)(((uint64_t)frames * C * b + 7u) / 8u);
}
// bytes of a VBR chunk given the bucket counts (sizes base-1, base, base+1, base+2) when every block is full
uint32_t vbr_full_chunk_bytes(const EncodePlan &p);

// ---- quantiser tables: dqt.rs:40-126 ---------------------------------------------------------------------
// Flat table for scale_factor_bits = s in the device layout of sea_common.cuh (tab_words(s) int32 values).
std::vector<int32_t> build_tables(uint32_t s);

}  // namespace sea
