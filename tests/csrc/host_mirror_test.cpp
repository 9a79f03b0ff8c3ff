// Drives the C++ host mirror (include/sea_b200.hpp) the way tests/streaming.rs and examples/bench.rs drive the crate:
//   host_mirror_test <pcm.raw> <channels> <rate> <bits> <vbr 0|1> <out_oneshot.sea> <out_stream.sea> <out_dec.raw>
// one-shot sea_encode, chunk-at-a-time SeaEncoder, SeaDecoder over the result.  The pytest side compares the three
// outputs with the oracle.  Needs a GPU (no CPU fallback).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>

#include "../../include/sea_b200.hpp"

static std::vector<uint8_t> slurp(const char *path)
{
    std::ifstream f(path, std::ios::binary);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static void dump(const char *path, const void *p, size_t n)
{
    std::ofstream f(path, std::ios::binary);
    f.write(static_cast<const char *>(p), (std::streamsize)n);
}

int main(int argc, char **argv)
{
    if (argc < 9) return 2;
    try {
        std::vector<uint8_t> raw = slurp(argv[1]);
        const uint32_t channels = (uint32_t)atoi(argv[2]), rate = (uint32_t)atoi(argv[3]);
        sea::EncoderSettings st;
        st.residual_bits = (float)atof(argv[4]);
        st.vbr = atoi(argv[5]) != 0;
        sea::Context ctx(0);
        const int16_t *pcm = reinterpret_cast<const int16_t *>(raw.data());
        const size_t n = raw.size() / 2;
        std::vector<uint8_t> one = sea::sea_encode(ctx, pcm, n, rate, channels, st);
        dump(argv[6], one.data(), one.size());

        sea::SliceReader rd(raw.data(), raw.size());
        sea::VecWriter wr;
        sea::SeaEncoder<sea::SliceReader, sea::VecWriter> enc(ctx, (uint8_t)channels, rate, (uint32_t)(n / channels), st, rd, wr);
        while (enc.encode_frame()) {}
        enc.finalize();
        dump(argv[7], wr.data.data(), wr.data.size());

        sea::SliceReader rd2(wr.data.data(), wr.data.size());
        sea::VecWriter pcm_out;
        sea::SeaDecoder<sea::SliceReader, sea::VecWriter> dec(ctx, rd2, pcm_out);
        while (dec.decode_frame()) {}
        dec.finalize();
        dump(argv[8], pcm_out.data.data(), pcm_out.data.size());
        sea::SeaDecodeInfo info = sea::sea_decode(ctx, one.data(), one.size());
        if (info.samples.size() * 2 != pcm_out.data.size() || memcmp(info.samples.data(), pcm_out.data.data(), pcm_out.data.size()) != 0) {
            fprintf(stderr, "one-shot decode differs from streaming decode\n");
            return 1;
        }
        // the multi-chunk forms must give the same bytes / samples, and a range must equal the slice of the whole
        sea::SliceReader rd3(raw.data(), raw.size());
        sea::VecWriter wr3;
        sea::SeaEncoder<sea::SliceReader, sea::VecWriter> enc3(ctx, (uint8_t)channels, rate, (uint32_t)(n / channels), st, rd3, wr3);
        while (enc3.encode_frames(2)) {}
        if (wr3.data != wr.data) {
            fprintf(stderr, "encode_frames(2) differs from encode_frame()\n");
            return 1;
        }
        sea::SliceReader rd4(wr.data.data(), wr.data.size());
        sea::VecWriter pcm4;
        sea::SeaDecoder<sea::SliceReader, sea::VecWriter> dec4(ctx, rd4, pcm4);
        while (dec4.decode_frames(3)) {}
        if (pcm4.data != pcm_out.data) {
            fprintf(stderr, "decode_frames(3) differs from decode_frame()\n");
            return 1;
        }
        const uint64_t first = 5000, count = 6000;
        sea::SeaDecodeInfo part = sea::sea_decode_range(ctx, one.data(), one.size(), first, count);
        if (part.samples.size() != count * channels ||
            memcmp(part.samples.data(), info.samples.data() + first * channels, part.samples.size() * 2) != 0) {
            fprintf(stderr, "sea_decode_range differs from the slice of the full decode\n");
            return 1;
        }
        // the in-process multi-GPU calls (here: two contexts on GPU 0) on a batch of three copies: same bytes, same samples
        {
            sea::MultiContext multi({0, 0});
            const uint32_t nb = 3, frames = (uint32_t)(n / channels);
            std::vector<int16_t> batch(nb * n);
            for (uint32_t i = 0; i < nb; i++) memcpy(batch.data() + i * n, pcm, n * 2);
            uint64_t bound = 0;
            sea_b200_settings cst = st.c();
            sea_b200_encode_bound(frames, channels, &cst, &bound);
            std::vector<uint8_t> outb(nb * bound);
            std::vector<uint64_t> po(nb), oo(nb), lens(nb), ns(nb);
            std::vector<uint32_t> nf(nb, frames);
            for (uint32_t i = 0; i < nb; i++) { po[i] = (uint64_t)i * n; oo[i] = (uint64_t)i * bound; }
            sea::MultiContext::Shares sh = multi.encode_batch(nb, batch.data(), po.data(), nf.data(), rate, channels, cst, outb.data(), oo.data(), lens.data());
            uint64_t total = 0;
            for (uint64_t a : sh.amount) total += a;
            for (uint32_t i = 0; i < nb; i++)
                if (lens[i] != one.size() || memcmp(outb.data() + oo[i], one.data(), one.size()) != 0) {
                    fprintf(stderr, "multi encode_batch stream %u differs from the one-shot encode\n", i);
                    return 1;
                }
            if (total != nb * one.size() || sh.first_stream[0] != 0) {
                fprintf(stderr, "multi encode_batch per-device counts do not add up\n");
                return 1;
            }
            std::vector<int16_t> back(nb * n);
            multi.decode_batch(nb, outb.data(), oo.data(), lens.data(), back.data(), po.data(), nullptr, ns.data());
            for (uint32_t i = 0; i < nb; i++)
                if (ns[i] != info.samples.size() || memcmp(back.data() + po[i], info.samples.data(), ns[i] * 2) != 0) {
                    fprintf(stderr, "multi decode_batch stream %u differs from the one-shot decode\n", i);
                    return 1;
                }
        }
        printf("ok %zu %zu %zu\n", one.size(), wr.data.size(), pcm_out.data.size());
        return 0;
    } catch (const sea::SeaError &e) {
        fprintf(stderr, "SeaError %d: %s\n", e.code, e.what());
        return 3;
    }
}
