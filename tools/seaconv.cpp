// seaconv.cpp -- the reference's CLI (examples/seaconv.rs) as a compiled program over include/sea_b200.hpp:
//   seaconv INPUT OUTPUT [-c|--chunk-size N] [-b|--bitrate B] [-s|--scalefactor-bits S] [-d|--scalefactor-distance D] [-v|--vbr]
// Same arguments, defaults, validation ranges and messages (seaconv.rs:12-144); .wav handling follows tests/wav.rs (hound):
// 8/16/24/32-bit integer and 32-bit float LPCM in, 16-bit PCM out.  The codec work is done by libsea_b200 on the GPU.
//   g++ -std=c++17 -O2 -Iinclude tools/seaconv.cpp -Lsea_codec_b200 -l:libsea_b200.so -Wl,-rpath,$PWD/sea_codec_b200 -o seaconv
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sea_b200.hpp"

static void die(const char *msg)
{
    std::fprintf(stderr, "Error: %s\n", msg);
    std::exit(1);
}

struct Wave {
    std::vector<int16_t> samples;
    uint32_t channels = 0, sample_rate = 0;
};

static uint32_t le32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t le16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static int16_t sat16(float f)  // Rust `as i16`: saturating, NaN -> 0
{
    if (f != f) return 0;
    if (f <= -32768.0f) return -32768;
    if (f >= 32767.0f) return 32767;
    return (int16_t)f;
}

static bool read_file(const std::string &path, std::vector<uint8_t> *out)
{
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::rewind(f);
    out->resize(n > 0 ? (size_t)n : 0);
    const bool ok = out->empty() || std::fread(out->data(), 1, out->size(), f) == out->size();
    std::fclose(f);
    return ok;
}

static bool read_wav(const std::string &path, Wave *w)  // tests/wav.rs:11-50
{
    std::vector<uint8_t> d;
    if (!read_file(path, &d) || d.size() < 12 || std::memcmp(d.data(), "RIFF", 4) || std::memcmp(d.data() + 8, "WAVE", 4)) return false;
    const uint8_t *fmt = nullptr, *pcm = nullptr;
    size_t fmt_len = 0, pcm_len = 0;
    for (size_t pos = 12; pos + 8 <= d.size();) {
        const size_t size = le32(&d[pos + 4]), avail = d.size() - pos - 8;
        if (!std::memcmp(&d[pos], "fmt ", 4)) fmt = &d[pos + 8], fmt_len = size < avail ? size : avail;
        if (!std::memcmp(&d[pos], "data", 4)) {
            pcm = &d[pos + 8], pcm_len = size < avail ? size : avail;
            break;
        }
        pos += 8 + size + (size & 1);
    }
    if (!fmt || !pcm || fmt_len < 16) return false;
    uint32_t tag = le16(fmt);
    const uint32_t channels = le16(fmt + 2), rate = le32(fmt + 4), bits = le16(fmt + 14);
    if (tag == 0xFFFE && fmt_len >= 26) tag = le16(fmt + 24);
    if (channels == 0 || channels > 2 || rate == 0) return false;  // "More than 2 channels are not supported"
    const size_t width = bits / 8;
    size_t n = width ? pcm_len / width : 0;
    n -= n % channels;
    w->channels = channels;
    w->sample_rate = rate;
    w->samples.resize(n);
    for (size_t i = 0; i < n; i++) {
        const uint8_t *p = pcm + i * width;
        if (tag == 1 && bits == 8) w->samples[i] = (int16_t)(((int)p[0] - 128) * 256);
        else if (tag == 1 && bits == 16) w->samples[i] = (int16_t)le16(p);
        else if (tag == 1 && bits == 24) {
            int32_t v = p[0] | (p[1] << 8) | (p[2] << 16);
            if (v & 0x800000) v -= 1 << 24;
            w->samples[i] = sat16(std::round(((float)v / (float)(1 << 23)) * 32767.0f));
        } else if (tag == 1 && bits == 32) {
            w->samples[i] = sat16(std::round(((float)(int32_t)le32(p) / (float)INT32_MAX) * 32767.0f));
        } else if (tag == 3 && bits == 32) {
            float f;
            const uint32_t u = le32(p);
            std::memcpy(&f, &u, 4);
            w->samples[i] = sat16(std::round(f * 32767.0f));
        } else {
            return false;
        }
    }
    return true;
}

struct FileWriter {
    FILE *f;
    void write(const void *p, size_t n)
    {
        if (n && std::fwrite(p, 1, n, f) != n) die("Failed to write output file");
    }
};
struct FileReader {
    FILE *f;
    size_t read(void *p, size_t n) { return std::fread(p, 1, n, f); }
};

static bool ends_with(const std::string &s, const char *ext)
{
    const size_t n = std::strlen(ext);
    return s.size() > n && s.compare(s.size() - n, n, ext) == 0 && s.find_last_of('.') == s.size() - n;
}

int main(int argc, char **argv)
{
    std::string input, output, chunk = "5120", bitrate = "3", sfbits = "4", sfdist = "20";
    bool vbr = false;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto val = [&](std::string *dst) {
            if (i + 1 >= argc) die("missing option value");
            *dst = argv[++i];
        };
        if (a == "-c" || a == "--chunk-size") val(&chunk);
        else if (a == "-b" || a == "--bitrate") val(&bitrate);
        else if (a == "-s" || a == "--scalefactor-bits") val(&sfbits);
        else if (a == "-d" || a == "--scalefactor-distance") val(&sfdist);
        else if (a == "-v" || a == "--vbr") vbr = true;
        else if (input.empty()) input = a;
        else if (output.empty()) output = a;
        else die("unexpected argument");
    }
    if (input.empty() || output.empty()) die("usage: seaconv INPUT OUTPUT [-c N] [-b B] [-s S] [-d D] [-v]");

    auto parse_uint = [](const std::string &s, unsigned long max, unsigned long *out) {
        char *end = nullptr;
        if (s.empty() || s[0] == '-') return false;
        *out = std::strtoul(s.c_str(), &end, 10);
        return end && *end == 0 && *out <= max;
    };
    unsigned long fpc = 0, sfb = 0, sff = 0;  // seaconv.rs:12-91
    if (!parse_uint(chunk, 0xffff, &fpc)) die("Failed to parse chunk size");
    if (fpc < 200 || fpc > 32000) die("Chunk size must be between 200 and 32000");
    if (!parse_uint(sfbits, 255, &sfb)) die("Failed to parse scale factor bits");
    if (sfb < 3 || sfb > 5) die("Scale factor bits must be between 3 and 5");
    if (!parse_uint(sfdist, 255, &sff)) die("Failed to parse scale factor frames");
    if (sff < 1 || fpc % sff != 0) die("Scale factor frames must be a divisor of chunk size");
    char *end = nullptr;
    const float bits = std::strtof(bitrate.c_str(), &end);
    if (bitrate.empty() || !end || *end != 0) die("Failed to parse residual bits");
    if (!(bits >= 1.0f && bits <= 8.0f)) die("Bitrate must be between 1.0 and 8.0");
    if (vbr) {
        if (!(bits >= 1.5f && bits <= 8.0f)) die("With VBR, bitrate must be between 1.5 and 8.0");
    } else if (bits != std::floor(bits)) {
        die("Without VBR, bitrate must be an integer between 1 and 8");
    }
    sea::EncoderSettings st;
    st.frames_per_chunk = (uint16_t)fpc;
    st.scale_factor_bits = (uint8_t)sfb;
    st.scale_factor_frames = (uint8_t)sff;
    st.residual_bits = bits;
    st.vbr = vbr;

    try {
        if (ends_with(input, ".wav") && ends_with(output, ".sea")) {
            Wave w;
            if (!read_wav(input, &w)) die("Failed to decode .wav file");
            FILE *fo = std::fopen(output.c_str(), "wb");
            if (!fo) die("Failed to create output file");
            sea::Context ctx(std::getenv("SEA_B200_DEVICE") ? std::atoi(std::getenv("SEA_B200_DEVICE")) : 0);
            sea::SliceReader reader(w.samples.data(), w.samples.size() * 2);
            FileWriter writer{fo};
            try {
                sea::SeaEncoder<sea::SliceReader, FileWriter> enc(ctx, (uint8_t)w.channels, w.sample_rate,
                                                                  (uint32_t)(w.samples.size() / w.channels), st, reader, writer);
                try {
                    while (enc.encode_frame()) {}
                } catch (const sea::SeaError &) {
                    die("Failed to encode frame");
                }
                enc.finalize();
            } catch (const sea::SeaError &) {
                die("Failed to create encoder");
            }
            std::fclose(fo);
        } else if (ends_with(input, ".sea") && ends_with(output, ".wav")) {
            FILE *fi = std::fopen(input.c_str(), "rb");
            if (!fi) die("Failed to open input file");
            sea::Context ctx(std::getenv("SEA_B200_DEVICE") ? std::atoi(std::getenv("SEA_B200_DEVICE")) : 0);
            FileReader reader{fi};
            sea::VecWriter pcm;
            sea::SeaFileHeader info{};
            try {
                sea::SeaDecoder<FileReader, sea::VecWriter> dec(ctx, reader, pcm);
                while (dec.decode_frame()) {}
                dec.finalize();
                info = dec.get_header();
            } catch (const sea::SeaError &) {
                die("Failed to decode frame");
            }
            std::fclose(fi);
            FILE *fo = std::fopen(output.c_str(), "wb");
            if (!fo) die("Failed to encode wav file");
            const uint32_t data = (uint32_t)pcm.data.size(), rate = info.sample_rate, ch = info.channels;
            uint8_t h[44];  // tests/wav.rs:52-75: 16-bit integer PCM
            std::memcpy(h, "RIFF", 4);
            const uint32_t riff = 36 + data, fmt_len = 16, byte_rate = rate * ch * 2;
            const uint16_t tag = 1, chs = (uint16_t)ch, align = (uint16_t)(ch * 2), bps = 16;
            std::memcpy(h + 4, &riff, 4);
            std::memcpy(h + 8, "WAVEfmt ", 8);
            std::memcpy(h + 16, &fmt_len, 4);
            std::memcpy(h + 20, &tag, 2);
            std::memcpy(h + 22, &chs, 2);
            std::memcpy(h + 24, &rate, 4);
            std::memcpy(h + 28, &byte_rate, 4);
            std::memcpy(h + 32, &align, 2);
            std::memcpy(h + 34, &bps, 2);
            std::memcpy(h + 36, "data", 4);
            std::memcpy(h + 40, &data, 4);
            FileWriter writer{fo};
            writer.write(h, 44);
            writer.write(pcm.data.data(), pcm.data.size());
            std::fclose(fo);
        } else {
            die("Invalid file extensions. Supported conversions are .wav to .sea and .sea to .wav");
        }
    } catch (const sea::SeaError &e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    return 0;
}
