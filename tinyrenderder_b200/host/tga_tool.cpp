// tga_tool.cpp - read a TGA, optionally flip, write it back (RLE or raw).  Built against our
// tgaimage.cpp and (test only) against the reference's, to check the two writers byte for byte.
// usage: tga_tool in.tga out.tga [raw]
#include <tgaimage.h>
#include <cstring>
int main(int argc, char** argv) {
    if (argc < 3) return 2;
    TGAImage img;
    if (!img.read_tga_file(argv[1])) return 1;
    bool rle = !(argc > 3 && !strcmp(argv[3], "raw"));
    return img.write_tga_file(argv[2], true, rle) ? 0 : 1;
}
