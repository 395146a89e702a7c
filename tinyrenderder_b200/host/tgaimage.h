// tgaimage.h - TGAColor / TGAImage with the API of the reference (tgaimage.h:29-104), written from
// scratch.  Kept on the host as the texture / framebuffer container; the device gets the raw byte
// array (buffer()).  Memory order: (x + y*w) * bpp, BGR(A).  write_tga_file produces byte-identical
// files to the reference's writer (same header, same RLE packetisation, tgaimage.cpp:161-242).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct TGAColor {
    std::uint8_t bgra[4];
    std::uint8_t bytespp;
    TGAColor() : bgra{0, 0, 0, 255}, bytespp(4) {}
    TGAColor(std::uint8_t R, std::uint8_t G, std::uint8_t B, std::uint8_t A = 255) : bgra{B, G, R, A}, bytespp(4) {}
    TGAColor(std::uint8_t v) : bgra{v, v, v, 255}, bytespp(1) {}
    TGAColor(const std::uint8_t* p, std::uint8_t bpp) : bgra{0, 0, 0, 0}, bytespp(bpp) {
        for (int i = 0; i < (int)bpp && i < 4; ++i) bgra[i] = p[i];
    }
    std::uint8_t& operator[](int i) { return bgra[i]; }
    const std::uint8_t& operator[](int i) const { return bgra[i]; }
    TGAColor operator*(float k) const {
        TGAColor r = *this;
        k = k < 0.f ? 0.f : (k > 1.f ? 1.f : k);
        for (int i = 0; i < 4; ++i) r.bgra[i] = (std::uint8_t)(bgra[i] * k);
        return r;
    }
};

class TGAImage {
public:
    enum Format { GRAYSCALE = 1, RGB = 3, RGBA = 4 };
    TGAImage() {}
    TGAImage(int width, int height, int bytespp, TGAColor clear = TGAColor());

    bool read_tga_file(const std::string filename);
    bool write_tga_file(const std::string filename, const bool vflip = true, const bool rle = true) const;
    void flip_horizontally();
    void flip_vertically();

    TGAColor get(const int x, const int y) const;
    void set(const int x, const int y, const TGAColor& c);
    int width() const { return w_; }
    int height() const { return h_; }
    int bytespp() const { return bpp_; }
    std::uint8_t* buffer() { return data_.empty() ? nullptr : data_.data(); }
    const std::uint8_t* buffer() const { return data_.empty() ? nullptr : data_.data(); }

private:
    int w_ = 0, h_ = 0;
    std::uint8_t bpp_ = 0;
    std::vector<std::uint8_t> data_;
};
