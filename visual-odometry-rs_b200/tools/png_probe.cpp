// png_probe - the PNG reader of the vors_track driver (tools/png_reader.h) as a host-only command, for the CPU tests:
//   png_probe gray  in.png out.raw    image::open().to_luma()      -> "rows cols" on stdout, u8 row-major samples in out.raw
//   png_probe depth in.png out.raw    helper::read_png_16bits      -> "rows cols" on stdout, u16 little-endian samples in out.raw
// A decoding error goes to stderr with exit code 1.
#include <cstdio>
#include <string>

#include "png_reader.h"

int main(int argc, char** argv) {
    if (argc != 4) { std::fprintf(stderr, "usage: png_probe gray|depth in.png out.raw\n"); return 2; }
    const std::string mode = argv[1];
    uint32_t w = 0, h = 0;
    std::string err;
    std::vector<uint8_t> gray;
    std::vector<uint16_t> depth;
    const bool ok = mode == "depth" ? vors_png::read_depth(argv[2], w, h, depth, err) : vors_png::read_gray(argv[2], w, h, gray, err);
    if (!ok) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }
    std::FILE* f = std::fopen(argv[3], "wb");
    if (!f) { std::fprintf(stderr, "cannot write %s\n", argv[3]); return 2; }
    if (mode == "depth") std::fwrite(depth.data(), 2, depth.size(), f);
    else std::fwrite(gray.data(), 1, gray.size(), f);
    std::fclose(f);
    std::printf("%u %u\n", h, w);
    return 0;
}
