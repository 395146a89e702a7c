// tgaimage.cpp - see tgaimage.h.  Behaviour follows the reference (file:line in comments).
#include "tgaimage.h"

#include <cstring>
#include <fstream>
#include <iostream>

namespace {
#pragma pack(push, 1)
struct Header {  // 18-byte TGA header, tgaimage.h:10-25
    std::uint8_t idlength = 0, colormaptype = 0, datatypecode = 0;
    std::uint16_t colormaporigin = 0, colormaplength = 0;
    std::uint8_t colormapdepth = 0;
    std::uint16_t x_origin = 0, y_origin = 0, width = 0, height = 0;
    std::uint8_t bitsperpixel = 0, imagedescriptor = 0;
};
#pragma pack(pop)
static_assert(sizeof(Header) == 18, "TGA header is 18 bytes");
}  // namespace

TGAImage::TGAImage(int width, int height, int bytespp, TGAColor clear)
    : w_(width), h_(height), bpp_((std::uint8_t)bytespp), data_((size_t)width * height * bytespp, 0) {
    for (size_t p = 0; p < (size_t)w_ * h_; ++p)           // tgaimage.cpp:8-17: every pixel = clear
        for (int i = 0; i < bpp_; ++i) data_[p * bpp_ + i] = clear.bgra[i];
}

TGAColor TGAImage::get(const int x, const int y) const {   // tgaimage.cpp:24-30
    if (data_.empty() || x < 0 || y < 0 || x >= w_ || y >= h_) return TGAColor();
    return TGAColor(&data_[((size_t)x + (size_t)y * w_) * bpp_], bpp_);
}
void TGAImage::set(const int x, const int y, const TGAColor& c) {  // tgaimage.cpp:32-39
    if (data_.empty() || x < 0 || y < 0 || x >= w_ || y >= h_) return;
    std::memcpy(&data_[((size_t)x + (size_t)y * w_) * bpp_], c.bgra, bpp_);
}

void TGAImage::flip_vertically() {
    if (data_.empty()) return;
    const size_t row = (size_t)w_ * bpp_;
    std::vector<std::uint8_t> tmp(row);
    for (int y = 0; y < h_ / 2; ++y) {
        std::uint8_t* a = &data_[y * row];
        std::uint8_t* b = &data_[(h_ - 1 - y) * row];
        std::memcpy(tmp.data(), a, row);
        std::memcpy(a, b, row);
        std::memcpy(b, tmp.data(), row);
    }
}
void TGAImage::flip_horizontally() {
    if (data_.empty()) return;
    for (int y = 0; y < h_; ++y)
        for (int x = 0; x < w_ / 2; ++x)
            for (int i = 0; i < bpp_; ++i)
                std::swap(data_[((size_t)x + (size_t)y * w_) * bpp_ + i],
                          data_[((size_t)(w_ - 1 - x) + (size_t)y * w_) * bpp_ + i]);
}

bool TGAImage::read_tga_file(const std::string filename) {  // tgaimage.cpp:76-122
    data_.clear();
    std::ifstream in(filename, std::ios::binary);
    if (!in.is_open()) { std::cerr << "can't open file " << filename << "\n"; return false; }
    Header hd;
    in.read((char*)&hd, sizeof(hd));
    if (!in.good()) { std::cerr << "can't read header\n"; return false; }
    w_ = hd.width;
    h_ = hd.height;
    bpp_ = hd.bitsperpixel >> 3;
    if (w_ <= 0 || h_ <= 0 || (bpp_ != 1 && bpp_ != 3 && bpp_ != 4)) { std::cerr << "invalid TGA format\n"; return false; }
    const size_t npx = (size_t)w_ * h_;
    data_.resize(npx * bpp_);
    in.seekg(hd.idlength, std::ios::cur);
    if (hd.datatypecode == 2 || hd.datatypecode == 3) {
        in.read((char*)data_.data(), data_.size());
    } else if (hd.datatypecode == 10 || hd.datatypecode == 11) {
        size_t px = 0;                                    // RLE packets, tgaimage.cpp:124-157
        std::uint8_t texel[4];
        while (px < npx) {
            int head = in.get();
            if (head < 0) return false;
            int count = (head & 127) + 1;
            if (px + count > npx) return false;
            if (head < 128) {
                in.read((char*)&data_[px * bpp_], (std::streamsize)count * bpp_);
            } else {
                in.read((char*)texel, bpp_);
                for (int i = 0; i < count; ++i) std::memcpy(&data_[(px + i) * bpp_], texel, bpp_);
            }
            px += count;
        }
    } else {
        std::cerr << "unknown TGA type\n";
        return false;
    }
    if (!(hd.imagedescriptor & 0x20)) flip_vertically();  // tgaimage.cpp:118-119
    if (hd.imagedescriptor & 0x10) flip_horizontally();
    return true;
}

bool TGAImage::write_tga_file(const std::string filename, const bool vflip, const bool rle) const {  // :161-191
    std::ofstream out(filename, std::ios::binary);
    if (!out.is_open()) { std::cerr << "can't open " << filename << "\n"; return false; }
    Header hd;
    hd.bitsperpixel = bpp_ * 8;
    hd.width = (std::uint16_t)w_;
    hd.height = (std::uint16_t)h_;
    hd.datatypecode = (bpp_ == 1 ? (rle ? 11 : 3) : (rle ? 10 : 2));
    hd.imagedescriptor = vflip ? 0x00 : 0x20;  // 0x00 = bottom-left origin: row 0 is the bottom of the picture
    out.write((const char*)&hd, sizeof(hd));
    if (!rle) {
        out.write((const char*)data_.data(), data_.size());
        return true;
    }
    // packetisation of tgaimage.cpp:193-242: a run packet for >= 2 equal pixels (max 128), else a raw
    // packet that stops right before the next pair of equal pixels (max 128)
    const size_t npx = (size_t)w_ * h_;
    auto same = [&](size_t a, size_t b) { return std::memcmp(&data_[a * bpp_], &data_[b * bpp_], bpp_) == 0; };
    size_t cur = 0;
    while (cur < npx) {
        size_t run = 1;
        while (cur + run < npx && run < 128 && same(cur + run, cur)) ++run;
        if (run > 1) {
            out.put((char)(run - 1 + 128));
            out.write((const char*)&data_[cur * bpp_], bpp_);
        } else {
            while (cur + run < npx && run < 128 && !same(cur + run, cur + run - 1)) ++run;
            out.put((char)(run - 1));
            out.write((const char*)&data_[cur * bpp_], (std::streamsize)run * bpp_);
        }
        cur += run;
    }
    return true;
}
