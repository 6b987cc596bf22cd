// vors_track — the reference's command-line driver (src/bin/vors_track.rs:17-145) on top of libvors_b200.so.
//
//   Usage: vors_track [fr1|fr2|fr3|icl] associations_file        (same as the reference, vors_track.rs:24)
//
// Reads a TUM RGB-D associations file, loads each 16-bit depth PNG (big-endian, src/misc/helper.rs:13-36) and colour
// PNG (converted to luma like image::open().to_luma(), vors_track.rs:143), tracks every frame through the C ABI
// (Config::init / Tracker::track / Tracker::current_frame) and prints one TUM trajectory line per frame to stdout:
// `timestamp tx ty tz qx qy qz qw` (src/dataset/tum_rgbd.rs:76-86).  Diagnostics go to stderr like the reference's
// eprintln! calls.  PNG decoding: tools/png_reader.h (zlib only; gray, RGB(A), palette, interlaced or not, CRC-checked).

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
#include "png_reader.h"

namespace {

const char* kUsage = "Usage: ./vors_track [fr1|fr2|fr3|icl] associations_file";

using vors_png::read_depth;
using vors_png::read_gray;

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
