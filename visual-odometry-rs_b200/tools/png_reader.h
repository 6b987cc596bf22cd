// PNG input of the vors_track driver: what the reference gets from `image::open(path).to_luma()` (src/bin/vors_track.rs:143)
// and `helper::read_png_16bits` (src/misc/helper.rs:13-36), on top of zlib only.
//
// Decodes every PNG the reference's decoder accepts for these two calls: gray / gray+alpha / RGB / RGBA at 8 or 16 bits,
// gray at 1, 2, 4 bits (expanded to 8 bits by bit replication, i.e. v * 255 / (2^n - 1)), palette images at 1..8 bits
// (expanded to 8-bit RGB; tRNS is ignored, luma has no alpha), Adam7-interlaced or not.  Every chunk's CRC is checked.
// Host code only: tools/png_probe.cpp exposes it to the CPU tests (tests/test_tum_io.py).
#pragma once

#include <zlib.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

namespace vors_png {

struct Png {
    uint32_t width = 0, height = 0;
    int bit_depth = 0, channels = 0;  // of `data`: 8 or 16 bits; 1 (gray), 2 (gray+alpha), 3 (RGB), 4 (RGBA)
    std::vector<uint8_t> data;        // height x width x channels x (bit_depth / 8), row-major, big-endian samples
};

inline uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

inline bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    out.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    return true;
}

// PNG filter types 0..4 of one scanline, in place on `row` given the previous (unfiltered) scanline `up`.
inline bool unfilter_row(uint8_t filter, uint8_t* row, const uint8_t* up, size_t n, size_t bpp) {
    for (size_t i = 0; i < n; ++i) {
        const int a = i >= bpp ? row[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0;
        int pred = 0;
        switch (filter) {
            case 0: pred = 0; break;
            case 1: pred = a; break;
            case 2: pred = b; break;
            case 3: pred = (a + b) / 2; break;
            case 4: {
                const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                break;
            }
            default: return false;
        }
        row[i] = uint8_t(row[i] + pred);
    }
    return true;
}

inline bool decode_png(const std::string& path, Png& png, std::string& err) {
    std::vector<uint8_t> file;
    if (!read_file(path, file)) { err = "cannot open " + path; return false; }
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) { err = path + ": not a PNG file"; return false; }
    std::vector<uint8_t> idat, plte;
    int color_type = -1, interlace = 0, depth = 0;
    bool have_ihdr = false, have_iend = false;
    size_t pos = 8;
    while (pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        if (size_t(len) > file.size() || pos + 12 + size_t(len) > file.size()) break;
        const uint8_t* type = &file[pos + 4];
        const uint8_t* body = &file[pos + 8];
        if (uint32_t(crc32(crc32(0L, Z_NULL, 0), type, uInt(4 + len))) != be32(body + len)) {
            err = path + ": corrupt PNG (chunk CRC mismatch)";
            return false;
        }
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
            png.width = be32(body);
            png.height = be32(body + 4);
            depth = body[8];
            color_type = body[9];
            interlace = body[12];
            have_ihdr = true;
        } else if (!std::memcmp(type, "PLTE", 4)) {
            plte.assign(body, body + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            have_iend = true;
            break;
        }
        pos += 12 + size_t(len);
    }
    if (!have_ihdr || !have_iend || png.width == 0 || png.height == 0) { err = path + ": corrupt PNG (missing IHDR / IEND)"; return false; }
    int in_channels = 0;
    switch (color_type) {
        case 0: in_channels = 1; break;
        case 2: in_channels = 3; break;
        case 3: in_channels = 1; break;  // palette index
        case 4: in_channels = 2; break;
        case 6: in_channels = 4; break;
        default: err = path + ": unsupported PNG colour type"; return false;
    }
    const bool palette = color_type == 3;
    const bool depth_ok = (color_type == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
                          (palette && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                          ((color_type == 2 || color_type == 4 || color_type == 6) && (depth == 8 || depth == 16));
    if (!depth_ok || interlace > 1) { err = path + ": unsupported PNG (bit depth / interlace method)"; return false; }
    if (palette && (plte.empty() || plte.size() % 3 != 0)) { err = path + ": corrupt PNG (palette image without PLTE)"; return false; }

    // output samples: sub-byte gray -> 8-bit gray, palette -> 8-bit RGB, everything else as stored
    png.channels = palette ? 3 : in_channels;
    png.bit_depth = depth < 8 ? 8 : depth;
    const size_t out_px = size_t(png.channels) * size_t(png.bit_depth / 8);
    const size_t bits_px = size_t(in_channels) * size_t(depth);
    const size_t bpp = bits_px >= 8 ? bits_px / 8 : 1;  // the filters' "bytes per pixel", at least 1

    struct Pass { uint32_t x0, y0, dx, dy; };
    static const Pass adam7[7] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    static const Pass whole[1] = {{0, 0, 1, 1}};
    const Pass* passes = interlace ? adam7 : whole;
    const int n_passes = interlace ? 7 : 1;
    size_t raw_size = 0;
    for (int p = 0; p < n_passes; ++p) {
        const Pass& q = passes[p];
        if (png.width <= q.x0 || png.height <= q.y0) continue;
        const size_t pw = (png.width - q.x0 + q.dx - 1) / q.dx, ph = (png.height - q.y0 + q.dy - 1) / q.dy;
        raw_size += ph * (1 + (pw * bits_px + 7) / 8);
    }
    std::vector<uint8_t> raw(raw_size);
    uLongf raw_len = uLongf(raw.size());
    if (uncompress(raw.data(), &raw_len, idat.data(), uLong(idat.size())) != Z_OK || raw_len != raw.size()) {
        err = path + ": corrupt PNG data";
        return false;
    }
    png.data.assign(size_t(png.width) * png.height * out_px, 0);
    const unsigned maxv = depth < 8 ? (1u << depth) - 1u : 255u;
    size_t at = 0;
    for (int p = 0; p < n_passes; ++p) {
        const Pass& q = passes[p];
        if (png.width <= q.x0 || png.height <= q.y0) continue;
        const size_t pw = (png.width - q.x0 + q.dx - 1) / q.dx, ph = (png.height - q.y0 + q.dy - 1) / q.dy;
        const size_t row_bytes = (pw * bits_px + 7) / 8;
        std::vector<uint8_t> zero(row_bytes, 0);
        const uint8_t* up = zero.data();
        for (size_t j = 0; j < ph; ++j) {
            const uint8_t filter = raw[at];
            uint8_t* row = &raw[at + 1];
            if (!unfilter_row(filter, row, up, row_bytes, bpp)) { err = path + ": bad PNG filter"; return false; }
            up = row;
            at += 1 + row_bytes;
            const size_t y = q.y0 + j * q.dy;
            for (size_t i = 0; i < pw; ++i) {
                uint8_t* out = &png.data[(y * png.width + (q.x0 + i * q.dx)) * out_px];
                if (depth >= 8) {
                    if (palette) {
                        const size_t e = size_t(row[i]) * 3;
                        if (e + 3 > plte.size()) { err = path + ": corrupt PNG (palette index out of range)"; return false; }
                        out[0] = plte[e]; out[1] = plte[e + 1]; out[2] = plte[e + 2];
                    } else {
                        std::memcpy(out, row + i * bpp, bpp);
                    }
                } else {  // 1, 2, 4 bits per sample, one sample per pixel, most significant bits first
                    const size_t bit = i * size_t(depth);
                    const unsigned v = (row[bit >> 3] >> (8 - depth - int(bit & 7))) & maxv;
                    if (palette) {
                        const size_t e = size_t(v) * 3;
                        if (e + 3 > plte.size()) { err = path + ": corrupt PNG (palette index out of range)"; return false; }
                        out[0] = plte[e]; out[1] = plte[e + 1]; out[2] = plte[e + 2];
                    } else {
                        out[0] = uint8_t(v * 255u / maxv);
                    }
                }
            }
        }
    }
    return true;
}

// helper::read_png_16bits (src/misc/helper.rs:13-36): 16-bit gray PNG, samples are big-endian.
inline bool read_depth(const std::string& path, uint32_t& w, uint32_t& h, std::vector<uint16_t>& out, std::string& err) {
    Png png;
    if (!decode_png(path, png, err)) return false;
    if (png.bit_depth != 16 || png.channels != 1) { err = path + ": depth image must be a 16-bit gray PNG"; return false; }
    w = png.width; h = png.height;
    out.resize(size_t(w) * h);
    for (size_t i = 0; i < out.size(); ++i) out[i] = uint16_t((png.data[2 * i] << 8) | png.data[2 * i + 1]);
    return true;
}

// image::open(path).to_luma() (vors_track.rs:143).  image 0.19 converts RGB to luma in f32 with the BT.709 weights and a
// truncating cast (recalled behaviour of the un-vendored crate); gray images pass through; 16-bit samples keep the high byte.
inline bool read_gray(const std::string& path, uint32_t& w, uint32_t& h, std::vector<uint8_t>& out, std::string& err) {
    Png png;
    if (!decode_png(path, png, err)) return false;
    w = png.width; h = png.height;
    out.resize(size_t(w) * h);
    const int bytes = png.bit_depth / 8, ch = png.channels;
    for (size_t i = 0; i < out.size(); ++i) {
        const uint8_t* p = &png.data[i * size_t(ch) * size_t(bytes)];
        if (ch <= 2) {
            out[i] = p[0];
        } else {
            const float l = 0.2126f * float(p[0]) + 0.7152f * float(p[bytes]) + 0.0722f * float(p[2 * bytes]);
            out[i] = uint8_t(l);
        }
    }
    return true;
}

}  // namespace vors_png
