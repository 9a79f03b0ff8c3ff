/*
 * sea_compat.h -- source compatibility for existing callers of the reference's two C-level surfaces.
 *
 *   - src/wasm_api.rs:32-111 exports  setup, wasm_sea_encode, wasm_sea_decode, allocate, deallocate;
 *   - c/sea.h:189 defines             int sea_decode(encoded, encoded_len, &sample_rate, &channels, output, &total_frames).
 *
 * libsea_b200.so exports the same argument lists under a sea_b200_ prefix (include/sea_b200.h); including this header
 * instead of c/sea.h (or declaring the wasm imports through it) keeps the caller's source unchanged while every byte is
 * encoded/decoded by the sm_100a kernels.  Define SEA_COMPAT_NO_GENERIC_NAMES to skip setup/allocate/deallocate.
 */
#ifndef SEA_COMPAT_H
#define SEA_COMPAT_H
#include "sea_b200.h"

#define wasm_sea_encode sea_b200_wasm_sea_encode
#define wasm_sea_decode sea_b200_wasm_sea_decode
#define sea_decode sea_b200_csea_decode
#ifndef SEA_COMPAT_NO_GENERIC_NAMES
#define setup sea_b200_wasm_setup
#define allocate sea_b200_wasm_allocate
#define deallocate sea_b200_wasm_deallocate
#endif

#endif /* SEA_COMPAT_H */
