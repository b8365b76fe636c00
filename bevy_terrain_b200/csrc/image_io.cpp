// image_io.cpp — source-image decoders for the preprocessor's inputs (host only, no GPU):
//   * PNG (the `image` crate behind Bevy's ImageLoader in the reference): 16-bit gray -> R16, 8-bit RGB / RGBA -> Rgba8
//   * TIFF (src/formats/tiff.rs: the `tiff` crate, every integer result is reinterpreted as R16Unorm texels): 16-bit gray
//     (unsigned or signed samples, bytes kept as they are) -> R16, 8-bit RGB / RGBA -> Rgba8; strips or tiles; no compression,
//     Deflate (8 / 32946) or LZW (5); horizontal-differencing predictor; little- and big-endian files
// Written against the PNG (ISO/IEC 15948) and TIFF 6.0 specifications; zlib does the inflate.
#include <zlib.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bevy_terrain_b200.h"

namespace {

struct Failure { std::string message; };

[[noreturn]] void fail(const char* fmt, ...) {
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw Failure{buf};
}

std::vector<unsigned char> read_file(const char* path) {
    FILE* f = std::fopen(path, "rb");
    if (!f) fail("cannot open %s", path);
    std::vector<unsigned char> data;
    unsigned char chunk[1 << 16];
    size_t n;
    while ((n = std::fread(chunk, 1, sizeof(chunk), f)) > 0) data.insert(data.end(), chunk, chunk + n);
    std::fclose(f);
    return data;
}

std::vector<unsigned char> inflate_all(const unsigned char* src, size_t n, size_t expected) {
    std::vector<unsigned char> out(expected ? std::min<size_t>(expected, std::max<size_t>(n * 4, 1u << 16)) : 1);   // grows on demand up to `expected`
    z_stream z{};
    if (inflateInit(&z) != Z_OK) fail("zlib: inflateInit failed");
    z.next_in = const_cast<unsigned char*>(src);
    z.avail_in = (uInt)n;
    size_t produced = 0;
    int rc = Z_OK;
    while (rc == Z_OK) {
        if (produced == out.size()) out.resize(out.size() * 2);
        z.next_out = out.data() + produced;
        z.avail_out = (uInt)std::min<size_t>(out.size() - produced, 1u << 30);
        const size_t before = z.avail_out;
        rc = inflate(&z, Z_NO_FLUSH);
        produced += before - z.avail_out;
        if (rc == Z_BUF_ERROR && z.avail_in == 0) break;
    }
    inflateEnd(&z);
    if (rc != Z_STREAM_END) fail("zlib: corrupt or truncated deflate stream (%d)", rc);
    out.resize(produced);
    return out;
}

// Deflate expands at most ~1032 : 1 and TIFF-LZW less: a header that asks for more than that is corrupt (and would make the
// decoder allocate gigabytes for a few bytes of input)
size_t plausible_output(size_t compressed_bytes) { return compressed_bytes * 1100 + (1u << 16); }

constexpr uint32_t kMaxSide = 1u << 17;   // 131072: far above any terrain source, keeps size arithmetic and allocations sane

struct Decoded {
    uint32_t width = 0, height = 0, format = BT_FORMAT_R16;
    std::vector<unsigned char> pixels;   // R16: host-endian u16; Rgba8: r, g, b, a
};

// ---------------------------------------------------------------------------------------------- PNG
uint32_t be32(const unsigned char* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

Decoded decode_png(const std::vector<unsigned char>& file) {
    static const unsigned char kSig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() < 8 || std::memcmp(file.data(), kSig, 8) != 0) fail("not a PNG file");
    uint32_t w = 0, h = 0;
    int depth = 0, colour = 0;
    std::vector<unsigned char> idat;
    size_t at = 8;
    bool seen_end = false;
    while (!seen_end) {
        if (at + 12 > file.size()) fail("PNG: truncated chunk list");
        const uint32_t len = be32(&file[at]);
        const unsigned char* type = &file[at + 4];
        const unsigned char* data = &file[at + 8];
        if (at + 12 + (size_t)len > file.size()) fail("PNG: truncated chunk");
        if (be32(data + len) != (uint32_t)crc32(crc32(0L, type, 4), data, len)) fail("PNG: chunk CRC mismatch");
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) fail("PNG: bad IHDR");
            w = be32(data); h = be32(data + 4); depth = data[8]; colour = data[9];
            if (data[10] != 0 || data[11] != 0) fail("PNG: unknown compression / filter method");
            if (data[12] != 0) fail("PNG: interlaced images are not supported");
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            seen_end = true;
        }
        at += 12 + (size_t)len;
    }
    if (w == 0 || h == 0) fail("PNG: missing IHDR");
    if (w > kMaxSide || h > kMaxSide) fail("PNG: %u x %u exceeds the supported image size", w, h);
    int channels;
    Decoded out;
    if (colour == 0 && depth == 16) { channels = 1; out.format = BT_FORMAT_R16; }
    else if (colour == 2 && depth == 8) { channels = 3; out.format = BT_FORMAT_RGBA8; }
    else if (colour == 6 && depth == 8) { channels = 4; out.format = BT_FORMAT_RGBA8; }
    else fail("PNG: colour type %d with %d bits is not a terrain source format (16-bit gray, 8-bit RGB / RGBA)", colour, depth);
    const size_t bpp = (size_t)channels * depth / 8, row = (size_t)w * bpp;
    if ((row + 1) * h > plausible_output(idat.size())) fail("PNG: %u x %u pixels cannot come from %zu compressed bytes", w, h, idat.size());
    std::vector<unsigned char> raw = inflate_all(idat.data(), idat.size(), (row + 1) * h);
    if (raw.size() != (row + 1) * h) fail("PNG: image data has %zu bytes, expected %zu", raw.size(), (row + 1) * h);
    std::vector<unsigned char> img(row * h);
    for (uint32_t y = 0; y < h; ++y) {   // undo the per-row filter (PNG spec, section 9)
        const unsigned char* in = &raw[(row + 1) * y];
        unsigned char* cur = &img[row * y];
        const unsigned char* up = y ? cur - row : nullptr;
        const int filter = in[0];
        if (filter > 4) fail("PNG: unknown filter type %d", filter);
        for (size_t i = 0; i < row; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
            int pred = 0;
            switch (filter) { case 1: pred = a; break; case 2: pred = b; break; case 3: pred = (a + b) >> 1; break; case 4: pred = paeth(a, b, c); break; }
            cur[i] = (unsigned char)(in[1 + i] + pred);
        }
    }
    out.width = w; out.height = h;
    if (out.format == BT_FORMAT_R16) {
        out.pixels.resize((size_t)w * h * 2);
        uint16_t* dst = reinterpret_cast<uint16_t*>(out.pixels.data());
        for (size_t i = 0; i < (size_t)w * h; ++i) dst[i] = (uint16_t)(img[2 * i] << 8 | img[2 * i + 1]);   // network byte order
    } else {
        out.pixels.resize((size_t)w * h * 4);
        for (size_t i = 0; i < (size_t)w * h; ++i) {
            for (int k = 0; k < 3; ++k) out.pixels[4 * i + k] = img[channels * i + k];
            out.pixels[4 * i + 3] = channels == 4 ? img[4 * i + 3] : 255;
        }
    }
    return out;
}

// ---------------------------------------------------------------------------------------------- TIFF
struct Tiff {
    const std::vector<unsigned char>& f;
    bool big;
    uint16_t u16(size_t at) const { need(at, 2); return big ? (uint16_t)(f[at] << 8 | f[at + 1]) : (uint16_t)(f[at + 1] << 8 | f[at]); }
    uint32_t u32(size_t at) const {
        need(at, 4);
        return big ? ((uint32_t)f[at] << 24 | (uint32_t)f[at + 1] << 16 | (uint32_t)f[at + 2] << 8 | f[at + 3])
                   : ((uint32_t)f[at + 3] << 24 | (uint32_t)f[at + 2] << 16 | (uint32_t)f[at + 1] << 8 | f[at]);
    }
    void need(size_t at, size_t n) const { if (at + n > f.size()) fail("TIFF: truncated file"); }
};

// LZW as TIFF uses it (TIFF 6.0 section 13): MSB-first codes of 9..12 bits, ClearCode 256, EndOfInformation 257, the code
// width grows one code EARLY
std::vector<unsigned char> lzw_decode(const unsigned char* src, size_t n, size_t expected) {
    std::vector<unsigned char> out;
    out.reserve(expected);
    std::vector<uint32_t> prefix(4096), length(4096);
    std::vector<unsigned char> suffix(4096), first(4096);
    for (uint32_t i = 0; i < 256; ++i) { prefix[i] = 0xFFFF; suffix[i] = first[i] = (unsigned char)i; length[i] = 1; }
    uint32_t next = 258, width = 9, old = 0xFFFF;
    uint64_t acc = 0;
    int bits = 0;
    size_t at = 0;
    std::vector<unsigned char> scratch;
    for (;;) {
        while (bits < (int)width) {
            if (at >= n) return out;   // some writers omit the EndOfInformation code
            acc = acc << 8 | src[at++];
            bits += 8;
        }
        const uint32_t code = (uint32_t)(acc >> (bits - width)) & ((1u << width) - 1u);
        bits -= (int)width;
        if (code == 257) break;
        if (code == 256) { next = 258; width = 9; old = 0xFFFF; continue; }
        if (old == 0xFFFF) {
            if (code >= 256) fail("TIFF: corrupt LZW stream");
            out.push_back((unsigned char)code);
            old = code;
            continue;
        }
        uint32_t emit = code;
        unsigned char extra = 0;
        bool append_extra = false;
        if (code >= next) {            // the KwKwK case: the string of `old` plus its own first byte
            if (code != next) fail("TIFF: corrupt LZW stream");
            emit = old;
            extra = first[old];
            append_extra = true;
        }
        scratch.resize(length[emit]);
        for (uint32_t c = emit, i = length[emit]; i-- > 0; c = prefix[c]) scratch[i] = suffix[c];
        out.insert(out.end(), scratch.begin(), scratch.end());
        if (append_extra) out.push_back(extra);
        if (next < 4096) {
            prefix[next] = old;
            suffix[next] = append_extra ? extra : first[emit];
            first[next] = first[old];
            length[next] = length[old] + 1;
            ++next;
            if (next + 1 == (1u << width) && width < 12) ++width;   // early change
        }
        old = code;
    }
    return out;
}

Decoded decode_tiff(const std::vector<unsigned char>& file) {
    if (file.size() < 8) fail("not a TIFF file");
    const bool big = file[0] == 'M' && file[1] == 'M';
    if (!big && !(file[0] == 'I' && file[1] == 'I')) fail("not a TIFF file");
    Tiff t{file, big};
    if (t.u16(2) != 42) fail("TIFF: bad magic (BigTIFF is not supported)");
    const size_t ifd = t.u32(4);
    const uint32_t n_entries = t.u16(ifd);
    uint32_t width = 0, height = 0, bits = 1, compression = 1, samples = 1, rows_per_strip = 0xFFFFFFFFu, predictor = 1, planar = 1;
    uint32_t tile_w = 0, tile_h = 0, sample_format = 1;
    std::vector<uint32_t> offsets, counts;
    auto values = [&](size_t entry, std::vector<uint32_t>& out) {
        const uint32_t type = t.u16(entry + 2), count = t.u32(entry + 4);
        const size_t size = type == 3 ? 2 : (type == 4 ? 4 : (type == 1 ? 1 : 0));
        if (!size) fail("TIFF: unsupported field type %u", type);
        if ((size_t)count * size > file.size()) fail("TIFF: field with %u values in a %zu-byte file", count, file.size());
        const size_t base = (size_t)count * size <= 4 ? entry + 8 : t.u32(entry + 8);
        out.resize(count);
        for (uint32_t i = 0; i < count; ++i)
            out[i] = size == 2 ? t.u16(base + 2 * i) : (size == 4 ? t.u32(base + 4 * i) : (t.need(base + i, 1), file[base + i]));
    };
    for (uint32_t e = 0; e < n_entries; ++e) {
        const size_t entry = ifd + 2 + 12 * (size_t)e;
        std::vector<uint32_t> v;
        const uint32_t tag = t.u16(entry);
        switch (tag) {
            case 256: values(entry, v); width = v.at(0); break;
            case 257: values(entry, v); height = v.at(0); break;
            case 258: values(entry, v); bits = v.at(0); for (uint32_t b : v) if (b != bits) fail("TIFF: mixed BitsPerSample"); break;
            case 259: values(entry, v); compression = v.at(0); break;
            case 273: case 324: values(entry, offsets); break;
            case 277: values(entry, v); samples = v.at(0); break;
            case 278: values(entry, v); rows_per_strip = v.at(0); break;
            case 279: case 325: values(entry, counts); break;
            case 284: values(entry, v); planar = v.at(0); break;
            case 317: values(entry, v); predictor = v.at(0); break;
            case 322: values(entry, v); tile_w = v.at(0); break;
            case 323: values(entry, v); tile_h = v.at(0); break;
            case 339: values(entry, v); sample_format = v.at(0); break;
            default: break;
        }
    }
    if (!width || !height || offsets.empty() || offsets.size() != counts.size()) fail("TIFF: missing dimensions or strip / tile tables");
    if (width > kMaxSide || height > kMaxSide) fail("TIFF: %u x %u exceeds the supported image size", width, height);
    if (rows_per_strip == 0 || ((tile_w != 0) != (tile_h != 0)) || tile_w > kMaxSide || tile_h > kMaxSide) fail("TIFF: bad strip / tile geometry");
    if (planar != 1) fail("TIFF: planar sample layout is not supported");
    if (predictor != 1 && predictor != 2) fail("TIFF: predictor %u is not supported", predictor);
    Decoded out;
    uint32_t keep;   // samples copied per pixel
    if (bits == 16 && samples == 1 && (sample_format == 1 || sample_format == 2)) { out.format = BT_FORMAT_R16; keep = 1; }
    else if (bits == 8 && (samples == 3 || samples == 4)) { out.format = BT_FORMAT_RGBA8; keep = samples; }
    else fail("TIFF: %u x %u-bit samples (format %u) is not a terrain source format (16-bit gray, 8-bit RGB / RGBA)", samples, bits, sample_format);
    const size_t bytes_per_sample = bits / 8, bpp = bytes_per_sample * samples;
    const bool tiled = tile_w && tile_h;
    const uint32_t bw = tiled ? tile_w : width, bh = tiled ? tile_h : std::min(rows_per_strip, height);
    const uint32_t across = (width + bw - 1) / bw, down = (height + bh - 1) / bh;
    if ((size_t)across * down != offsets.size()) fail("TIFF: %zu strips / tiles, expected %u", offsets.size(), across * down);
    if ((size_t)width * height * bpp > plausible_output(file.size())) fail("TIFF: %u x %u pixels cannot come from a %zu-byte file", width, height, file.size());
    out.width = width; out.height = height;
    out.pixels.assign((size_t)width * height * (out.format == BT_FORMAT_R16 ? 2 : 4), 0);
    for (uint32_t by = 0; by < down; ++by)
        for (uint32_t bx = 0; bx < across; ++bx) {
            const size_t index = (size_t)by * across + bx;
            t.need(offsets[index], counts[index]);
            const unsigned char* src = &file[offsets[index]];
            const uint32_t rows = tiled ? bh : std::min(bh, height - by * bh);   // tiles are always full size, the last strip is short
            const size_t expected = (size_t)bw * rows * bpp;
            // bound the block BEFORE anything is allocated for it: a tiny crafted file must not make the decoder reserve
            // gigabytes (tile_w / tile_h / rows_per_strip are attacker-controlled up to kMaxSide each)
            if (compression == 1 ? counts[index] < expected : expected > plausible_output(counts[index]))
                fail("TIFF: strip / tile %zu: %zu stored bytes cannot hold %zu decoded bytes", index, (size_t)counts[index], expected);
            if (tiled && ((size_t)tile_w > 2 * (size_t)width + 16 || (size_t)tile_h > 2 * (size_t)height + 16)) fail("TIFF: tile %u x %u is larger than the image", tile_w, tile_h);
            std::vector<unsigned char> block;
            if (compression == 1) block.assign(src, src + counts[index]);
            else if (compression == 8 || compression == 32946) block = inflate_all(src, counts[index], expected);
            else if (compression == 5) block = lzw_decode(src, counts[index], expected);
            else fail("TIFF: compression %u is not supported (none, LZW, Deflate)", compression);
            if (block.size() < expected) fail("TIFF: strip / tile %zu has %zu bytes, expected %zu", index, block.size(), expected);
            for (uint32_t r = 0; r < rows; ++r) {
                const uint32_t y = by * bh + r;
                if (y >= height) break;
                unsigned char* line = &block[(size_t)r * bw * bpp];
                if (bits == 16) {   // to host order first: the predictor works on sample values
                    uint16_t* s = reinterpret_cast<uint16_t*>(line);
                    for (size_t i = 0; i < (size_t)bw * samples; ++i) s[i] = big ? (uint16_t)(line[2 * i] << 8 | line[2 * i + 1]) : (uint16_t)(line[2 * i + 1] << 8 | line[2 * i]);
                    if (predictor == 2) for (size_t i = samples; i < (size_t)bw * samples; ++i) s[i] = (uint16_t)(s[i] + s[i - samples]);
                } else if (predictor == 2) {
                    for (size_t i = samples; i < (size_t)bw * samples; ++i) line[i] = (unsigned char)(line[i] + line[i - samples]);
                }
                const uint32_t x0 = bx * bw, n = std::min(bw, width - x0);
                if (out.format == BT_FORMAT_R16) {
                    std::memcpy(&out.pixels[((size_t)y * width + x0) * 2], line, (size_t)n * 2);
                } else {
                    for (uint32_t x = 0; x < n; ++x) {
                        unsigned char* dst = &out.pixels[((size_t)y * width + x0 + x) * 4];
                        for (uint32_t k = 0; k < 3; ++k) dst[k] = line[(size_t)x * samples + k];
                        dst[3] = keep == 4 ? line[(size_t)x * samples + 3] : 255;
                    }
                }
            }
        }
    return out;
}

}  // namespace

extern "C" {

bt_status bt_image_load(const char* path, bt_image* out, char* error, size_t error_capacity) {
    if (error && error_capacity) error[0] = 0;
    if (!path || !out) return BT_ERR_INVALID_ARG;
    std::memset(out, 0, sizeof(*out));
    try {
        const std::vector<unsigned char> file = read_file(path);
        Decoded d = (file.size() >= 4 && file[0] == 0x89 && file[1] == 'P') ? decode_png(file) : decode_tiff(file);
        void* pixels = std::malloc(d.pixels.size());
        if (!pixels) return BT_ERR_OUT_OF_MEMORY;
        std::memcpy(pixels, d.pixels.data(), d.pixels.size());
        out->width = d.width; out->height = d.height; out->format = d.format; out->pixels = pixels;
        return BT_OK;
    } catch (const Failure& f) {
        if (error && error_capacity) std::snprintf(error, error_capacity, "%s: %s", path, f.message.c_str());
        return BT_ERR_IO;
    } catch (const std::exception& e) {
        if (error && error_capacity) std::snprintf(error, error_capacity, "%s: %s", path, e.what());
        return BT_ERR_IO;
    }
}

void bt_image_free(bt_image* image) {
    if (image && image->pixels) { std::free(image->pixels); image->pixels = nullptr; }
}

}  // extern "C"
