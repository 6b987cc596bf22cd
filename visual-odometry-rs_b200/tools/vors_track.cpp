// vors_track — the reference's command-line driver (src/bin/vors_track.rs:17-145) on top of libvors_b200.so.
//
//   Usage: vors_track [fr1|fr2|fr3|icl] associations_file        (same as the reference, vors_track.rs:24)
//
// Reads a TUM RGB-D associations file, loads each 16-bit depth PNG (big-endian, src/misc/helper.rs:13-36) and colour
// PNG (converted to luma like image::open().to_luma(), vors_track.rs:143), tracks every frame through the C ABI
// (Config::init / Tracker::track / Tracker::current_frame) and prints one TUM trajectory line per frame to stdout:
// `timestamp tx ty tz qx qy qz qw` (src/dataset/tum_rgbd.rs:76-86).  Diagnostics go to stderr like the reference's
// eprintln! calls.  PNG decoding is a small zlib-based reader (8/16-bit gray, RGB, RGBA, gray+alpha, non-interlaced).
#include <zlib.h>

#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/vors_b200.h"

namespace {

const char* kUsage = "Usage: ./vors_track [fr1|fr2|fr3|icl] associations_file";

struct Png {
    uint32_t width = 0, height = 0;
    int bit_depth = 0, channels = 0;
    std::vector<uint8_t> data;  // height x width x channels x (bit_depth / 8), row-major, big-endian samples
};

uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    out.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    return true;
}

bool decode_png(const std::string& path, Png& png, std::string& err) {
    std::vector<uint8_t> file;
    if (!read_file(path, file)) { err = "cannot open " + path; return false; }
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) { err = path + ": not a PNG file"; return false; }
    std::vector<uint8_t> idat;
    int color_type = -1, interlace = 0;
    size_t pos = 8;
    while (pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        const char* type = reinterpret_cast<const char*>(&file[pos + 4]);
        if (pos + 12 + len > file.size()) break;
        const uint8_t* body = &file[pos + 8];
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
            png.width = be32(body);
            png.height = be32(body + 4);
            png.bit_depth = body[8];
            color_type = body[9];
            interlace = body[12];
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + len;
    }
    switch (color_type) {
        case 0: png.channels = 1; break;
        case 2: png.channels = 3; break;
        case 4: png.channels = 2; break;
        case 6: png.channels = 4; break;
        default: err = path + ": unsupported PNG colour type (palette images are not supported)"; return false;
    }
    if (interlace != 0 || (png.bit_depth != 8 && png.bit_depth != 16)) { err = path + ": unsupported PNG (interlaced or bit depth != 8/16)"; return false; }
    const size_t bpp = size_t(png.channels) * size_t(png.bit_depth / 8);
    const size_t stride = size_t(png.width) * bpp;
    std::vector<uint8_t> raw((stride + 1) * png.height);
    uLongf raw_len = uLongf(raw.size());
    if (uncompress(raw.data(), &raw_len, idat.data(), uLong(idat.size())) != Z_OK || raw_len != raw.size()) {
        err = path + ": corrupt PNG data";
        return false;
    }
    png.data.assign(stride * png.height, 0);
    std::vector<uint8_t> zero(stride, 0);
    for (uint32_t y = 0; y < png.height; ++y) {
        const uint8_t filter = raw[(stride + 1) * y];
        const uint8_t* in = &raw[(stride + 1) * y + 1];
        uint8_t* out = &png.data[stride * y];
        const uint8_t* up = y ? &png.data[stride * (y - 1)] : zero.data();
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? out[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0;
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
                default: err = path + ": bad PNG filter"; return false;
            }
            out[i] = uint8_t(in[i] + pred);
        }
    }
    return true;
}

// helper::read_png_16bits (src/misc/helper.rs:13-36): 16-bit gray PNG, samples are big-endian.
bool read_depth(const std::string& path, uint32_t& w, uint32_t& h, std::vector<uint16_t>& out, std::string& err) {
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
bool read_gray(const std::string& path, uint32_t& w, uint32_t& h, std::vector<uint8_t>& out, std::string& err) {
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

struct Association {  // src/dataset/tum_rgbd.rs:64-74
    double depth_ts = 0, color_ts = 0;
    std::string depth_path, color_path;
};

// tum_rgbd::parse::associations (tum_rgbd.rs:97-145): every line is a comment (`#...`) or
// `depth_timestamp depth_file_path rgb_timestamp rgb_file_path`; anything else is a "Parsing error".
bool parse_associations(const std::string& file, std::vector<Association>& out, std::string& err) {
    std::ifstream f(file);
    if (!f) { err = "cannot open " + file; return false; }
    const size_t slash = file.find_last_of('/');
    const std::string parent = slash == std::string::npos ? std::string(".") : file.substr(0, slash);  // abs_path, vors_track.rs:126-138
    std::string line;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!line.empty() && line[0] == '#') continue;
        std::istringstream ss(line);
        Association a;
        std::string extra;
        if (!(ss >> a.depth_ts >> a.depth_path >> a.color_ts >> a.color_path)) { err = "Parsing error"; return false; }
        a.depth_path = parent + "/" + a.depth_path;
        a.color_path = parent + "/" + a.color_path;
        out.push_back(a);
    }
    return true;
}

template <typename T>
std::string shortest(T v) {  // Rust's `{}` for floats: shortest round-trip digits, never exponent notation
    char buf[512];
    const auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

}  // namespace

int main(int argc, char** argv) {
    // check_args / create_camera (vors_track.rs:75-110)
    if (argc != 3) {
        std::fprintf(stderr, "%s\n\"Wrong number of arguments\"\n", kUsage);
        return 0;  // the reference prints the error and exits normally (vors_track.rs:17-22)
    }
    vors_config cfg;
    vors_config_default(&cfg);  // nb_levels 6, threshold 7, depth scale 5000, idepth variance 1e-4 (vors_track.rs:34-40)
    const std::string cam = argv[1];
    if (cam == "fr1") { cfg.cx = 318.643040f; cfg.cy = 255.313989f; cfg.fx = 517.306408f; cfg.fy = 516.469215f; }        // tum_rgbd.rs:31-35
    else if (cam == "fr2") { cfg.cx = 325.141442f; cfg.cy = 249.701764f; cfg.fx = 520.908620f; cfg.fy = 521.007327f; }   // :39-43
    else if (cam == "fr3") { cfg.cx = 320.106653f; cfg.cy = 247.632132f; cfg.fx = 535.433105f; cfg.fy = 539.212524f; }   // :47-51
    else if (cam == "icl") { cfg.cx = 319.5f; cfg.cy = 239.5f; cfg.fx = 481.20f; cfg.fy = -480.00f; }                     // :23-27
    else {
        std::fprintf(stderr, "%s\n\"Unknown camera id: %s\"\n", kUsage, cam.c_str());
        return 0;
    }
    cfg.skew = 0.0f;
    std::vector<Association> assoc;
    std::string err;
    if (!parse_associations(argv[2], assoc, err)) {
        std::fprintf(stderr, "%s\n\"%s\"\n", kUsage, err.c_str());
        return 0;
    }
    if (assoc.empty()) {
        std::fprintf(stderr, "\"empty associations file\"\n");  // the reference would panic on associations[0]
        return 0;
    }

    uint32_t w = 0, h = 0, dw = 0, dh = 0;
    std::vector<uint16_t> depth;
    std::vector<uint8_t> gray;
    auto read_images = [&](const Association& a) {  // vors_track.rs:140-145
        return read_depth(a.depth_path, dw, dh, depth, err) && read_gray(a.color_path, w, h, gray, err) && (dw == w && dh == h ? true : (err = "depth / colour size mismatch", false));
    };
    if (!read_images(assoc[0])) { std::fprintf(stderr, "\"%s\"\n", err.c_str()); return 0; }
    vors_tracker* tracker = nullptr;
    // decoder output is row-major: the library transposes on the device (what DMatrix::from_row_slice does, vors_track.rs:142)
    if (vors_tracker_create(&cfg, assoc[0].depth_ts, depth.data(), assoc[0].color_ts, gray.data(), h, w, VORS_ROW_MAJOR, &tracker) != VORS_OK) {
        std::fprintf(stderr, "\"%s\"\n", vors_last_error());
        return 0;
    }
    double keyframe_ts = assoc[0].depth_ts;
    for (size_t i = 1; i < assoc.size(); ++i) {  // vors_track.rs:49-64
        if (!read_images(assoc[i])) { std::fprintf(stderr, "\"%s\"\n", err.c_str()); break; }
        vors_track_stats stats;
        const int rc = vors_tracker_track(tracker, assoc[i].depth_ts, depth.data(), assoc[i].color_ts, gray.data(), &stats);
        if (rc < 0) { std::fprintf(stderr, "\"%s\"\n", vors_last_error()); break; }
        if (rc == VORS_OPTIMIZATION_FAILED) std::fprintf(stderr, "Error at Cholesky decomposition of hessian\n");  // lm_optimizer.rs:133
        std::fprintf(stderr, "Optical_flow: %s\n", shortest(stats.optical_flow).c_str());                           // inverse_compositional.rs:222
        if (stats.keyframe_changed) {
            std::fprintf(stderr, "Changing keyframe after: %s seconds\n", shortest(assoc[i].depth_ts - keyframe_ts).c_str());  // :229
            keyframe_ts = assoc[i].depth_ts;
        }
        double ts = 0;
        vors_pose p;
        vors_tracker_current_frame(tracker, &ts, &p);
        std::printf("%s %s %s %s %s %s %s %s\n", shortest(ts).c_str(), shortest(p.t[0]).c_str(), shortest(p.t[1]).c_str(),
                    shortest(p.t[2]).c_str(), shortest(p.q[0]).c_str(), shortest(p.q[1]).c_str(), shortest(p.q[2]).c_str(),
                    shortest(p.q[3]).c_str());
    }
    vors_tracker_destroy(tracker);
    return 0;
}
