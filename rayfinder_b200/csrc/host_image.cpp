// PNG and baseline-JPEG decoding for the scene baker: what nlrs::Texture::fromMemory (common/texture.cpp:12-52) gets from
// stbi_load_from_memory(data, size, &w, &h, &channels, 4) — 8-bit RGBA, row 0 = top.
//
// stb_image (nothings/stb @ beebb24b, pinned by the reference's external/CMakeLists.txt) is not vendored under
// /root/reference, so its decoders are restated here from their published algorithm.  Every step that decides a texel value
// is integer arithmetic and is kept identical: for JPEG the dequantisation, the 12-bit fixed-point 8x8 IDCT
// (stbi__idct_block, which its SSE2 twin reproduces bit for bit by construction), the chroma upsampling filters
// (stbi__resample_row_*) and the 20-bit fixed-point YCbCr -> RGB conversion (stbi__YCbCr_to_RGB_row); for PNG the
// defiltering, the bit-depth scaling (1/2/4-bit grey x 255/85/17, 16-bit -> high byte) and palette expansion.  Inflate is
// zlib's (a valid stream has one decoding).  Not covered (rejected with an error, never approximated): progressive and
// arithmetic-coded JPEG, CMYK/YCCK JPEG, 12-bit JPEG.  Parity against a real stb_image build is UNPINNED (no copy of it
// here); tests/test_host.py pins PNG against Pillow exactly and JPEG against Pillow within the IDCT tolerance.
#include "rf_internal.h"

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <zlib.h>

namespace rfb200
{
namespace
{
struct DecodeError
{
    std::string message;
};
[[noreturn]] void fail(const char* what) { throw DecodeError{what}; }

// ------------------------------------------------------------------------------------------------- PNG
std::uint32_t be32(const std::uint8_t* p) { return (std::uint32_t(p[0]) << 24) | (std::uint32_t(p[1]) << 16) | (std::uint32_t(p[2]) << 8) | p[3]; }

int paeth(int a, int b, int c)
{
    // stbi__paeth
    const int thresh = c * 3 - (a + b);
    const int lo = a < b ? a : b;
    const int hi = a < b ? b : a;
    const int t0 = (hi <= thresh) ? lo : c;
    const int t1 = (thresh <= lo) ? hi : t0;
    return t1;
}

// Defilter one pass (an Adam7 sub-image or the whole image) of `w` x `h` pixels: `raw` holds h x (1 + rowBytes) bytes.
void defilter(const std::uint8_t* raw, std::size_t rawSize, std::uint32_t w, std::uint32_t h, int bitsPerPixel, std::vector<std::uint8_t>& out)
{
    const std::size_t rowBytes = (static_cast<std::size_t>(w) * bitsPerPixel + 7) / 8;
    const int         bpp = bitsPerPixel >= 8 ? bitsPerPixel / 8 : 1; // filter distance in bytes
    if (rawSize < (rowBytes + 1) * h) fail("not enough pixels");
    out.assign(rowBytes * h, 0);
    std::vector<std::uint8_t> zero(rowBytes, 0);
    for (std::uint32_t y = 0; y < h; ++y)
    {
        const std::uint8_t* src = raw + y * (rowBytes + 1);
        const int           filter = *src++;
        std::uint8_t*       cur = out.data() + y * rowBytes;
        const std::uint8_t* prior = y ? cur - rowBytes : zero.data();
        if (filter > 4) fail("invalid filter");
        for (std::size_t i = 0; i < rowBytes; ++i)
        {
            const int a = i >= static_cast<std::size_t>(bpp) ? cur[i - bpp] : 0;
            const int b = prior[i];
            const int c = i >= static_cast<std::size_t>(bpp) ? prior[i - bpp] : 0;
            int       v = src[i];
            switch (filter)
            {
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) >> 1; break;
            case 4: v += paeth(a, b, c); break;
            default: break;
            }
            cur[i] = static_cast<std::uint8_t>(v);
        }
    }
}

void decodePng(const std::uint8_t* data, std::size_t size, std::vector<std::uint8_t>& rgba, std::uint32_t& width, std::uint32_t& height)
{
    static const std::uint8_t SIGNATURE[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (size < 8 || std::memcmp(data, SIGNATURE, 8) != 0) fail("bad png sig");
    std::size_t               pos = 8;
    int                       depth = 0, colour = 0, interlace = 0;
    std::vector<std::uint8_t> idat;
    std::uint8_t              palette[256][3] = {};
    bool                      haveHeader = false, havePalette = false;
    for (;;)
    {
        if (pos + 8 > size) fail("truncated png");
        const std::uint32_t length = be32(data + pos);
        const std::uint32_t type = be32(data + pos + 4);
        const std::uint8_t* body = data + pos + 8;
        if (pos + 12 + static_cast<std::size_t>(length) > size) fail("truncated png");
        if (type == 0x49484452u) // IHDR
        {
            if (length != 13) fail("bad IHDR len");
            width = be32(body), height = be32(body + 4);
            depth = body[8], colour = body[9];
            interlace = body[12];
            if (width == 0 || height == 0) fail("0-pixel image");
            if (depth != 1 && depth != 2 && depth != 4 && depth != 8 && depth != 16) fail("1/2/4/8/16-bit only");
            if (colour > 6 || colour == 1 || colour == 5) fail("bad ctype");
            if (colour == 3 && depth == 16) fail("bad ctype");
            if (body[10] != 0 || body[11] != 0 || interlace > 1) fail("bad png header");
            if (static_cast<std::uint64_t>(width) * height > (1ull << 28)) fail("too large");
            haveHeader = true;
        }
        else if (type == 0x504C5445u) // PLTE
        {
            if (length > 256 * 3 || length % 3) fail("invalid PLTE");
            for (std::uint32_t i = 0; i < length / 3; ++i) std::memcpy(palette[i], body + 3 * i, 3);
            havePalette = true;
        }
        else if (type == 0x49444154u) // IDAT
        {
            idat.insert(idat.end(), body, body + length);
        }
        else if (type == 0x49454E44u) // IEND
        {
            break;
        }
        // tRNS and every ancillary chunk only affect alpha or metadata: fromMemory forces alpha to 255
        pos += 12 + length;
    }
    if (!haveHeader || idat.empty()) fail("no IDAT");
    if (colour == 3 && !havePalette) fail("no PLTE");
    const int channels = colour == 0 ? 1 : colour == 2 ? 3 : colour == 3 ? 1 : colour == 4 ? 2 : 4;
    const int bitsPerPixel = channels * depth;

    // inflate
    std::vector<std::uint8_t> raw;
    {
        const std::uint64_t rowBytes = (static_cast<std::uint64_t>(width) * bitsPerPixel + 7) / 8;
        raw.resize((rowBytes + 1) * height + (interlace ? 7 * height + 64 : 0) + 64);
        z_stream zs{};
        if (inflateInit(&zs) != Z_OK) fail("zlib init");
        zs.next_in = idat.data();
        zs.avail_in = static_cast<uInt>(idat.size());
        zs.next_out = raw.data();
        zs.avail_out = static_cast<uInt>(raw.size());
        const int rc = inflate(&zs, Z_FINISH);
        const std::size_t produced = raw.size() - zs.avail_out;
        inflateEnd(&zs);
        if (rc != Z_STREAM_END && rc != Z_BUF_ERROR && rc != Z_OK) fail("bad zlib stream");
        raw.resize(produced);
    }

    // defilter (+ de-interlace) into one packed image of `bitsPerPixel` per pixel
    const std::size_t         fullRowBytes = (static_cast<std::size_t>(width) * bitsPerPixel + 7) / 8;
    std::vector<std::uint8_t> image;
    const auto                sample = [&](const std::vector<std::uint8_t>& img, std::size_t rowBytes, std::uint32_t x, std::uint32_t y, int c) -> std::uint32_t {
        // channel c of pixel (x, y), as stored (depth bits)
        const std::uint8_t* row = img.data() + y * rowBytes;
        if (depth == 8) return row[x * channels + c];
        if (depth == 16) return (std::uint32_t(row[(x * channels + c) * 2]) << 8) | row[(x * channels + c) * 2 + 1];
        const std::size_t bit = static_cast<std::size_t>(x) * depth; // channels == 1 below 8 bits
        return (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
    };
    rgba.assign(static_cast<std::size_t>(width) * height * 4, 255);
    const std::uint32_t depthScale[9] = {0, 0xFF, 0x55, 0, 0x11, 0, 0, 0, 0x01}; // stbi__depth_scale_table
    const auto          emit = [&](const std::vector<std::uint8_t>& img, std::size_t rowBytes, std::uint32_t sx, std::uint32_t sy, std::uint32_t dx, std::uint32_t dy) {
        std::uint8_t* px = rgba.data() + (static_cast<std::size_t>(dy) * width + dx) * 4;
        const auto    to8 = [&](std::uint32_t v) -> std::uint8_t {
            if (depth == 16) return static_cast<std::uint8_t>(v >> 8); // stbi__convert_16_to_8
            if (depth < 8) return static_cast<std::uint8_t>(v * depthScale[depth]);
            return static_cast<std::uint8_t>(v);
        };
        if (colour == 3)
        {
            const std::uint32_t idx = sample(img, rowBytes, sx, sy, 0);
            px[0] = palette[idx][0], px[1] = palette[idx][1], px[2] = palette[idx][2];
        }
        else if (channels <= 2)
        {
            px[0] = px[1] = px[2] = to8(sample(img, rowBytes, sx, sy, 0));
        }
        else
        {
            for (int c = 0; c < 3; ++c) px[c] = to8(sample(img, rowBytes, sx, sy, c));
        }
    };
    if (!interlace)
    {
        defilter(raw.data(), raw.size(), width, height, bitsPerPixel, image);
        for (std::uint32_t y = 0; y < height; ++y)
            for (std::uint32_t x = 0; x < width; ++x) emit(image, fullRowBytes, x, y, x, y);
    }
    else
    {
        static const int xorig[7] = {0, 4, 0, 2, 0, 1, 0}, yorig[7] = {0, 0, 4, 0, 2, 0, 1};
        static const int xspc[7] = {8, 8, 4, 4, 2, 2, 1}, yspc[7] = {8, 8, 8, 4, 4, 2, 2};
        std::size_t      offset = 0;
        for (int p = 0; p < 7; ++p)
        {
            const std::uint32_t pw = (width - xorig[p] + xspc[p] - 1) / xspc[p];
            const std::uint32_t ph = (height - yorig[p] + yspc[p] - 1) / yspc[p];
            if (pw == 0 || ph == 0) continue;
            const std::size_t rowBytes = (static_cast<std::size_t>(pw) * bitsPerPixel + 7) / 8;
            if (offset > raw.size()) fail("not enough pixels");
            defilter(raw.data() + offset, raw.size() - offset, pw, ph, bitsPerPixel, image);
            for (std::uint32_t y = 0; y < ph; ++y)
                for (std::uint32_t x = 0; x < pw; ++x) emit(image, rowBytes, x, y, x * xspc[p] + xorig[p], y * yspc[p] + yorig[p]);
            offset += (rowBytes + 1) * ph;
        }
    }
}

// ------------------------------------------------------------------------------------------------ JPEG
constexpr std::uint8_t ZIGZAG[64 + 15] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
                                          6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
                                          39, 46, 53, 60, 61, 54, 47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

struct Huffman
{
    std::uint8_t  size[257] = {};
    std::uint16_t code[256] = {};
    std::uint8_t  values[256] = {};
    int           maxcode[18] = {}; // shifted left to 16 bits
    int           delta[17] = {};
    bool          present = false;

    void build(const int counts[16])
    {
        int k = 0;
        for (int i = 0; i < 16; ++i)
            for (int j = 0; j < counts[i]; ++j)
            {
                if (k >= 256) fail("bad size list");
                size[k++] = static_cast<std::uint8_t>(i + 1);
            }
        size[k] = 0;
        int code_ = 0;
        k = 0;
        for (int j = 1; j <= 16; ++j)
        {
            delta[j] = k - code_;
            if (size[k] == j)
            {
                while (size[k] == j) code[k++] = static_cast<std::uint16_t>(code_++);
                if (code_ - 1 >= (1 << j)) fail("bad code lengths");
            }
            maxcode[j] = code_ << (16 - j);
            code_ <<= 1;
        }
        maxcode[17] = 0x7FFFFFFF;
        present = true;
    }
};

struct Component
{
    int                       id = 0, h = 1, v = 1, tq = 0, hd = 0, ha = 0, dcPred = 0;
    int                       x = 0, y = 0, w2 = 0, h2 = 0;
    std::vector<std::uint8_t> data;
};

struct Jpeg
{
    const std::uint8_t* p = nullptr;
    const std::uint8_t* end = nullptr;
    std::uint32_t       codeBuffer = 0;
    int                 codeBits = 0;
    std::uint8_t        marker = 0xFF; // stbi's STBI__MARKER_none
    bool                noMore = false;
    Huffman             huffDc[4], huffAc[4];
    std::uint16_t       dequant[4][64] = {};
    Component           comp[4];
    int                 numComponents = 0, width = 0, height = 0, hMax = 1, vMax = 1, mcuW = 0, mcuH = 0, mcuX = 0, mcuY = 0;
    int                 restartInterval = 0, todo = 0;
    int                 scanN = 0, order[4] = {};
    bool                jfif = false;
    int                 app14Transform = -1;
    int                 rgb = 0;

    int  get8() { return p < end ? *p++ : 0; }
    int  get16() { const int hi = get8(); return (hi << 8) | get8(); }
    void skip(int n) { p = (end - p < n) ? end : p + n; }

    void growBuffer()
    {
        do
        {
            unsigned b = noMore ? 0u : static_cast<unsigned>(get8());
            if (b == 0xFF)
            {
                int c = get8();
                while (c == 0xFF) c = get8(); // consume fill bytes
                if (c != 0)
                {
                    marker = static_cast<std::uint8_t>(c);
                    noMore = true;
                    return;
                }
            }
            codeBuffer |= b << (24 - codeBits);
            codeBits += 8;
        } while (codeBits <= 24);
    }

    int huffDecode(const Huffman& h)
    {
        if (codeBits < 16) growBuffer();
        const unsigned temp = codeBuffer >> 16;
        int            k;
        for (k = 1; k <= 16; ++k)
            if (static_cast<int>(temp) < h.maxcode[k]) break;
        if (k == 17)
        {
            codeBits -= 16;
            return -1;
        }
        if (k > codeBits) return -1;
        const int c = static_cast<int>((codeBuffer >> (32 - k)) & ((1u << k) - 1u)) + h.delta[k];
        if (c < 0 || c >= 256) return -1;
        codeBits -= k;
        codeBuffer <<= k;
        return h.values[c];
    }

    int extendReceive(int n)
    {
        // stbi__extend_receive: n bits, sign-extended the JPEG way
        if (codeBits < n) growBuffer();
        if (codeBits < n) return 0;
        const int      sgn = static_cast<int>(codeBuffer >> 31);
        const unsigned k = (codeBuffer << n) | (codeBuffer >> (32 - n)); // rotate left
        static const std::uint32_t bmask[17] = {0, 1, 3, 7, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 16383, 32767, 65535};
        static const int           jbias[16] = {0, -1, -3, -7, -15, -31, -63, -127, -255, -511, -1023, -2047, -4095, -8191, -16383, -32767};
        codeBuffer = k & ~bmask[n];
        const unsigned value = k & bmask[n];
        codeBits -= n;
        return static_cast<int>(value) + (jbias[n] & (sgn - 1));
    }

    void decodeBlock(short data[64], Component& c)
    {
        // stbi__jpeg_decode_block
        if (codeBits < 16) growBuffer();
        const int t = huffDecode(huffDc[c.hd]);
        if (t < 0 || t > 15) fail("bad huffman code");
        std::memset(data, 0, 64 * sizeof(short));
        const int diff = t ? extendReceive(t) : 0;
        const int dc = c.dcPred + diff;
        c.dcPred = dc;
        data[0] = static_cast<short>(dc * dequant[c.tq][0]);
        int k = 1;
        do
        {
            if (codeBits < 16) growBuffer();
            const int rs = huffDecode(huffAc[c.ha]);
            if (rs < 0) fail("bad huffman code");
            const int s = rs & 15, r = rs >> 4;
            if (s == 0)
            {
                if (rs != 0xF0) break; // end of block
                k += 16;
            }
            else
            {
                k += r;
                const unsigned zig = ZIGZAG[k++];
                data[zig] = static_cast<short>(extendReceive(s) * dequant[c.tq][zig]);
            }
        } while (k < 64);
    }

    void reset()
    {
        codeBits = 0, codeBuffer = 0, noMore = false;
        for (Component& c : comp) c.dcPred = 0;
        marker = 0xFF;
        todo = restartInterval ? restartInterval : 0x7FFFFFFF;
    }
};

std::uint8_t clamp8(int x)
{
    if (static_cast<unsigned>(x) > 255u) return x < 0 ? 0 : 255;
    return static_cast<std::uint8_t>(x);
}

#define RF_F2F(x) (static_cast<int>(((x) * 4096 + 0.5)))
#define RF_FSH(x) ((x) * 4096)
// stbi's STBI__IDCT_1D: derived from jidctint
#define RF_IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)    \
    int t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3; \
    p2 = s2;                                          \
    p3 = s6;                                          \
    p1 = (p2 + p3) * RF_F2F(0.5411961f);              \
    t2 = p1 + p3 * RF_F2F(-1.847759065f);             \
    t3 = p1 + p2 * RF_F2F(0.765366865f);              \
    p2 = s0;                                          \
    p3 = s4;                                          \
    t0 = RF_FSH(p2 + p3);                             \
    t1 = RF_FSH(p2 - p3);                             \
    x0 = t0 + t3;                                     \
    x3 = t0 - t3;                                     \
    x1 = t1 + t2;                                     \
    x2 = t1 - t2;                                     \
    t0 = s7;                                          \
    t1 = s5;                                          \
    t2 = s3;                                          \
    t3 = s1;                                          \
    p3 = t0 + t2;                                     \
    p4 = t1 + t3;                                     \
    p1 = t0 + t3;                                     \
    p2 = t1 + t2;                                     \
    p5 = (p3 + p4) * RF_F2F(1.175875602f);            \
    t0 = t0 * RF_F2F(0.298631336f);                   \
    t1 = t1 * RF_F2F(2.053119869f);                   \
    t2 = t2 * RF_F2F(3.072711026f);                   \
    t3 = t3 * RF_F2F(1.501321110f);                   \
    p1 = p5 + p1 * RF_F2F(-0.899976223f);             \
    p2 = p5 + p2 * RF_F2F(-2.562915447f);             \
    p3 = p3 * RF_F2F(-1.961570560f);                  \
    p4 = p4 * RF_F2F(-0.390180644f);                  \
    t3 += p1 + p4;                                    \
    t2 += p2 + p3;                                    \
    t1 += p2 + p4;                                    \
    t0 += p1 + p3;

void idctBlock(std::uint8_t* out, int outStride, const short data[64])
{
    int          val[64], *v = val;
    const short* d = data;
    for (int i = 0; i < 8; ++i, ++d, ++v)
    {
        if (d[8] == 0 && d[16] == 0 && d[24] == 0 && d[32] == 0 && d[40] == 0 && d[48] == 0 && d[56] == 0)
        {
            const int dcterm = d[0] * 4;
            v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dcterm;
        }
        else
        {
            RF_IDCT_1D(d[0], d[8], d[16], d[24], d[32], d[40], d[48], d[56])
            x0 += 512, x1 += 512, x2 += 512, x3 += 512;
            v[0] = (x0 + t3) >> 10;
            v[56] = (x0 - t3) >> 10;
            v[8] = (x1 + t2) >> 10;
            v[48] = (x1 - t2) >> 10;
            v[16] = (x2 + t1) >> 10;
            v[40] = (x2 - t1) >> 10;
            v[24] = (x3 + t0) >> 10;
            v[32] = (x3 - t0) >> 10;
        }
    }
    v = val;
    std::uint8_t* o = out;
    for (int i = 0; i < 8; ++i, v += 8, o += outStride)
    {
        RF_IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
        x0 += 65536 + (128 << 17);
        x1 += 65536 + (128 << 17);
        x2 += 65536 + (128 << 17);
        x3 += 65536 + (128 << 17);
        o[0] = clamp8((x0 + t3) >> 17);
        o[7] = clamp8((x0 - t3) >> 17);
        o[1] = clamp8((x1 + t2) >> 17);
        o[6] = clamp8((x1 - t2) >> 17);
        o[2] = clamp8((x2 + t1) >> 17);
        o[5] = clamp8((x2 - t1) >> 17);
        o[3] = clamp8((x3 + t0) >> 17);
        o[4] = clamp8((x3 - t0) >> 17);
    }
}

// stbi__resample_row_*: `out` has room for w * hs bytes.
const std::uint8_t* resample1(std::uint8_t*, const std::uint8_t* inNear, const std::uint8_t*, int, int) { return inNear; }
const std::uint8_t* resampleV2(std::uint8_t* out, const std::uint8_t* inNear, const std::uint8_t* inFar, int w, int)
{
    for (int i = 0; i < w; ++i) out[i] = static_cast<std::uint8_t>((3 * inNear[i] + inFar[i] + 2) >> 2);
    return out;
}
const std::uint8_t* resampleH2(std::uint8_t* out, const std::uint8_t* in, const std::uint8_t*, int w, int)
{
    if (w == 1)
    {
        out[0] = out[1] = in[0];
        return out;
    }
    out[0] = in[0];
    out[1] = static_cast<std::uint8_t>((in[0] * 3 + in[1] + 2) >> 2);
    int i;
    for (i = 1; i < w - 1; ++i)
    {
        const int n = 3 * in[i] + 2;
        out[i * 2 + 0] = static_cast<std::uint8_t>((n + in[i - 1]) >> 2);
        out[i * 2 + 1] = static_cast<std::uint8_t>((n + in[i + 1]) >> 2);
    }
    out[i * 2 + 0] = static_cast<std::uint8_t>((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
    out[i * 2 + 1] = in[w - 1];
    return out;
}
const std::uint8_t* resampleHV2(std::uint8_t* out, const std::uint8_t* inNear, const std::uint8_t* inFar, int w, int)
{
    if (w == 1)
    {
        out[0] = out[1] = static_cast<std::uint8_t>((3 * inNear[0] + inFar[0] + 2) >> 2);
        return out;
    }
    int t1 = 3 * inNear[0] + inFar[0];
    out[0] = static_cast<std::uint8_t>((t1 + 2) >> 2);
    for (int i = 1; i < w; ++i)
    {
        const int t0 = t1;
        t1 = 3 * inNear[i] + inFar[i];
        out[i * 2 - 1] = static_cast<std::uint8_t>((3 * t0 + t1 + 8) >> 4);
        out[i * 2] = static_cast<std::uint8_t>((3 * t1 + t0 + 8) >> 4);
    }
    out[w * 2 - 1] = static_cast<std::uint8_t>((t1 + 2) >> 2);
    return out;
}
const std::uint8_t* resampleGeneric(std::uint8_t* out, const std::uint8_t* inNear, const std::uint8_t*, int w, int hs)
{
    for (int i = 0; i < w; ++i)
        for (int j = 0; j < hs; ++j) out[i * hs + j] = inNear[i];
    return out;
}

void decodeJpeg(const std::uint8_t* data, std::size_t size, std::vector<std::uint8_t>& rgba, std::uint32_t& width, std::uint32_t& height)
{
    Jpeg j{};
    j.p = data, j.end = data + size;
    if (j.get8() != 0xFF || j.get8() != 0xD8) fail("no SOI");
    const auto nextMarker = [&]() -> int {
        if (j.marker != 0xFF)
        {
            const int m = j.marker;
            j.marker = 0xFF;
            return m;
        }
        int x = j.get8();
        if (x != 0xFF) return 0xFF; // stbi: "none"
        while (x == 0xFF) x = j.get8();
        return x;
    };
    const auto processMarker = [&](int m) {
        switch (m)
        {
        case 0xDD: // DRI
            if (j.get16() != 4) fail("bad DRI len");
            j.restartInterval = j.get16();
            return;
        case 0xDB: // DQT
        {
            int L = j.get16() - 2;
            while (L > 0)
            {
                const int q = j.get8();
                const int p = q >> 4, t = q & 15;
                const bool sixteen = p != 0;
                if (p != 0 && p != 1) fail("bad DQT type");
                if (t > 3) fail("bad DQT table");
                for (int i = 0; i < 64; ++i) j.dequant[t][ZIGZAG[i]] = static_cast<std::uint16_t>(sixteen ? j.get16() : j.get8());
                L -= sixteen ? 129 : 65;
            }
            if (L != 0) fail("bad DQT len");
            return;
        }
        case 0xC4: // DHT
        {
            int L = j.get16() - 2;
            while (L > 0)
            {
                int       sizes[16], n = 0;
                const int q = j.get8();
                const int tc = q >> 4, th = q & 15;
                if (tc > 1 || th > 3) fail("bad DHT header");
                for (int i = 0; i < 16; ++i)
                {
                    sizes[i] = j.get8();
                    n += sizes[i];
                }
                if (n > 256) fail("bad DHT header");
                L -= 17;
                Huffman& h = tc == 0 ? j.huffDc[th] : j.huffAc[th];
                h = Huffman{};
                h.build(sizes);
                for (int i = 0; i < n; ++i) h.values[i] = static_cast<std::uint8_t>(j.get8());
                L -= n;
            }
            if (L != 0) fail("bad DHT len");
            return;
        }
        default: break;
        }
        if ((m >= 0xE0 && m <= 0xEF) || m == 0xFE)
        {
            int L = j.get16();
            if (L < 2) fail(m == 0xFE ? "bad COM len" : "bad APP len");
            L -= 2;
            if (m == 0xE0 && L >= 5)
            {
                static const unsigned char tag[5] = {'J', 'F', 'I', 'F', '\0'};
                bool                       ok = true;
                for (int i = 0; i < 5; ++i)
                    if (j.get8() != tag[i]) ok = false;
                L -= 5;
                if (ok) j.jfif = true;
            }
            else if (m == 0xEE && L >= 12)
            {
                static const unsigned char tag[6] = {'A', 'd', 'o', 'b', 'e', '\0'};
                bool                       ok = true;
                for (int i = 0; i < 6; ++i)
                    if (j.get8() != tag[i]) ok = false;
                L -= 6;
                if (ok)
                {
                    j.get8();  // version
                    j.get16(); // flags0
                    j.get16(); // flags1
                    j.app14Transform = j.get8();
                    L -= 6;
                }
            }
            j.skip(L);
            return;
        }
        fail("unknown marker");
    };

    // header: markers up to the frame header
    int m = nextMarker();
    while (!(m == 0xC0 || m == 0xC1 || m == 0xC2))
    {
        processMarker(m);
        m = nextMarker();
        while (m == 0xFF)
        {
            if (j.p >= j.end) fail("no SOF");
            m = nextMarker();
        }
    }
    if (m == 0xC2) fail("progressive JPEG is not supported by this decoder");
    {
        const int Lf = j.get16();
        if (Lf < 11) fail("bad SOF len");
        if (j.get8() != 8) fail("only 8-bit");
        j.height = j.get16();
        j.width = j.get16();
        if (j.height == 0) fail("no header height");
        if (j.width == 0) fail("0 width");
        const int c = j.get8();
        if (c != 3 && c != 1 && c != 4) fail("bad component count");
        if (c == 4) fail("CMYK / YCCK JPEG is not supported by this decoder");
        j.numComponents = c;
        if (Lf != 8 + 3 * c) fail("bad SOF len");
        static const unsigned char rgbIds[3] = {'R', 'G', 'B'};
        for (int i = 0; i < c; ++i)
        {
            Component& k = j.comp[i];
            k.id = j.get8();
            if (c == 3 && k.id == rgbIds[i]) ++j.rgb;
            const int q = j.get8();
            k.h = q >> 4, k.v = q & 15;
            if (!k.h || k.h > 4 || !k.v || k.v > 4) fail("bad H/V");
            k.tq = j.get8();
            if (k.tq > 3) fail("bad TQ");
            j.hMax = k.h > j.hMax ? k.h : j.hMax;
            j.vMax = k.v > j.vMax ? k.v : j.vMax;
        }
        for (int i = 0; i < c; ++i)
            if (j.hMax % j.comp[i].h != 0 || j.vMax % j.comp[i].v != 0) fail("bad H/V");
        j.mcuW = j.hMax * 8, j.mcuH = j.vMax * 8;
        j.mcuX = (j.width + j.mcuW - 1) / j.mcuW;
        j.mcuY = (j.height + j.mcuH - 1) / j.mcuH;
        for (int i = 0; i < c; ++i)
        {
            Component& k = j.comp[i];
            k.x = (j.width * k.h + j.hMax - 1) / j.hMax;
            k.y = (j.height * k.v + j.vMax - 1) / j.vMax;
            k.w2 = j.mcuX * k.h * 8;
            k.h2 = j.mcuY * k.v * 8;
            k.data.assign(static_cast<std::size_t>(k.w2) * k.h2, 0);
        }
    }

    // scans
    for (m = nextMarker();; m = nextMarker())
    {
        if (m == 0xDA)
        {
            const int Ls = j.get16();
            j.scanN = j.get8();
            if (j.scanN < 1 || j.scanN > 4 || j.scanN > j.numComponents) fail("bad SOS component count");
            if (Ls != 6 + 2 * j.scanN) fail("bad SOS len");
            for (int i = 0; i < j.scanN; ++i)
            {
                const int id = j.get8(), q = j.get8();
                int       which = 0;
                for (; which < j.numComponents; ++which)
                    if (j.comp[which].id == id) break;
                if (which == j.numComponents) return fail("bad SOS component");
                j.comp[which].hd = q >> 4;
                j.comp[which].ha = q & 15;
                if (j.comp[which].hd > 3 || j.comp[which].ha > 3) fail("bad huff table index");
                j.order[i] = which;
            }
            const int spectralStart = j.get8();
            j.get8(); // spectral end (baseline: ignored by stbi beyond the check below)
            const int aa = j.get8();
            if (spectralStart != 0 || (aa >> 4) != 0 || (aa & 15) != 0) fail("bad SOS");
            // entropy-coded segment
            j.reset();
            short block[64];
            const auto restartIfDue = [&]() -> bool {
                if (--j.todo <= 0)
                {
                    if (j.codeBits < 24) j.growBuffer();
                    if (!(j.marker >= 0xD0 && j.marker <= 0xD7)) return false; // "if it's NOT a restart, then just bail"
                    j.reset();
                }
                return true;
            };
            if (j.scanN == 1)
            {
                Component& c = j.comp[j.order[0]];
                const int  w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
                bool       go = true;
                for (int y = 0; y < h && go; ++y)
                    for (int x = 0; x < w && go; ++x)
                    {
                        j.decodeBlock(block, c);
                        idctBlock(c.data.data() + c.w2 * y * 8 + x * 8, c.w2, block);
                        go = restartIfDue();
                    }
            }
            else
            {
                bool go = true;
                for (int y = 0; y < j.mcuY && go; ++y)
                    for (int x = 0; x < j.mcuX && go; ++x)
                    {
                        for (int k = 0; k < j.scanN; ++k)
                        {
                            Component& c = j.comp[j.order[k]];
                            for (int v = 0; v < c.v; ++v)
                                for (int hh = 0; hh < c.h; ++hh)
                                {
                                    const int x2 = (x * c.h + hh) * 8, y2 = (y * c.v + v) * 8;
                                    j.decodeBlock(block, c);
                                    idctBlock(c.data.data() + c.w2 * y2 + x2, c.w2, block);
                                }
                        }
                        go = restartIfDue();
                    }
            }
            if (j.marker == 0xFF)
            {
                // stbi: scan ahead for the next marker
                while (j.p < j.end)
                {
                    const int x = j.get8();
                    if (x == 0xFF)
                    {
                        j.marker = static_cast<std::uint8_t>(j.get8());
                        break;
                    }
                }
            }
        }
        else if (m == 0xD9) // EOI
        {
            break;
        }
        else if (m == 0xDC) // DNL
        {
            const int Ld = j.get16();
            const int NL = j.get16();
            if (Ld != 4) fail("bad DNL len");
            if (NL != j.height) fail("bad DNL height");
        }
        else if (m == 0xFF)
        {
            if (j.p >= j.end) break; // ran off the end: stbi stops at the missing EOI too
        }
        else
        {
            processMarker(m);
        }
    }

    // resample + colour conversion (load_jpeg_image with req_comp = 4)
    width = static_cast<std::uint32_t>(j.width), height = static_cast<std::uint32_t>(j.height);
    const int  n = j.numComponents;
    const bool isRgb = n == 3 && (j.rgb == 3 || (j.app14Transform == 0 && !j.jfif));
    struct Resample
    {
        const std::uint8_t* (*fn)(std::uint8_t*, const std::uint8_t*, const std::uint8_t*, int, int);
        const std::uint8_t *line0, *line1;
        int                 hs, vs, wLores, ystep, ypos;
        std::vector<std::uint8_t> linebuf;
    } res[4];
    for (int k = 0; k < n; ++k)
    {
        Resample& r = res[k];
        r.linebuf.assign(static_cast<std::size_t>(j.width) + 3 + 8, 0);
        r.hs = j.hMax / j.comp[k].h;
        r.vs = j.vMax / j.comp[k].v;
        r.ystep = r.vs >> 1;
        r.wLores = (j.width + r.hs - 1) / r.hs;
        r.ypos = 0;
        r.line0 = r.line1 = j.comp[k].data.data();
        if (r.hs == 1 && r.vs == 1) r.fn = resample1;
        else if (r.hs == 1 && r.vs == 2) r.fn = resampleV2;
        else if (r.hs == 2 && r.vs == 1) r.fn = resampleH2;
        else if (r.hs == 2 && r.vs == 2) r.fn = resampleHV2;
        else r.fn = resampleGeneric;
        if (r.hs > 2 || r.vs > 2) r.linebuf.assign(static_cast<std::size_t>(r.wLores) * r.hs + 8, 0);
    }
    rgba.assign(static_cast<std::size_t>(width) * height * 4, 255);
    const int float2fixed_1_40200 = static_cast<int>(1.40200f * 4096.0f + 0.5f) << 8;
    const int float2fixed_0_71414 = static_cast<int>(0.71414f * 4096.0f + 0.5f) << 8;
    const int float2fixed_0_34414 = static_cast<int>(0.34414f * 4096.0f + 0.5f) << 8;
    const int float2fixed_1_77200 = static_cast<int>(1.77200f * 4096.0f + 0.5f) << 8;
    for (int y = 0; y < j.height; ++y)
    {
        const std::uint8_t* co[4] = {};
        for (int k = 0; k < n; ++k)
        {
            Resample&  r = res[k];
            const bool yBot = r.ystep >= (r.vs >> 1);
            co[k] = r.fn(r.linebuf.data(), yBot ? r.line1 : r.line0, yBot ? r.line0 : r.line1, r.wLores, r.hs);
            if (++r.ystep >= r.vs)
            {
                r.ystep = 0;
                r.line0 = r.line1;
                if (++r.ypos < j.comp[k].y) r.line1 += j.comp[k].w2;
            }
        }
        std::uint8_t* out = rgba.data() + static_cast<std::size_t>(y) * width * 4;
        if (n == 3 && !isRgb)
        {
            for (int i = 0; i < j.width; ++i)
            {
                // stbi__YCbCr_to_RGB_row
                const int yFixed = (co[0][i] << 20) + (1 << 19);
                const int cr = co[2][i] - 128, cb = co[1][i] - 128;
                int       r = yFixed + cr * float2fixed_1_40200;
                int       g = yFixed + (cr * -float2fixed_0_71414) + ((cb * -float2fixed_0_34414) & 0xffff0000);
                int       b = yFixed + cb * float2fixed_1_77200;
                r >>= 20, g >>= 20, b >>= 20;
                out[4 * i + 0] = clamp8(r), out[4 * i + 1] = clamp8(g), out[4 * i + 2] = clamp8(b), out[4 * i + 3] = 255;
            }
        }
        else if (n == 3)
        {
            for (int i = 0; i < j.width; ++i) out[4 * i + 0] = co[0][i], out[4 * i + 1] = co[1][i], out[4 * i + 2] = co[2][i], out[4 * i + 3] = 255;
        }
        else
        {
            for (int i = 0; i < j.width; ++i) out[4 * i + 0] = out[4 * i + 1] = out[4 * i + 2] = co[0][i], out[4 * i + 3] = 255;
        }
    }
}
} // namespace

// stbi_load_from_memory(..., 4) for the two container formats glTF allows.  Throws std::runtime_error-like DecodeError
// internally; the C-ABI wrappers turn it into a status.
bool decodeImageRgba8(const std::uint8_t* data, std::size_t size, std::vector<std::uint8_t>& rgba, std::uint32_t& width, std::uint32_t& height, std::string& error)
{
    try
    {
        if (size >= 2 && data[0] == 0xFF && data[1] == 0xD8)
            decodeJpeg(data, size, rgba, width, height);
        else if (size >= 8 && data[0] == 137 && data[1] == 80)
            decodePng(data, size, rgba, width, height);
        else
            fail("unknown image type");
        return true;
    }
    catch (const DecodeError& e)
    {
        error = e.message;
        return false;
    }
}
} // namespace rfb200

// Texture::fromMemory (common/texture.cpp:12-52): decode, then b | g << 8 | r << 16 | 255 << 24 per pixel.
extern "C" rf_status rf_texture_from_memory(const void* data, std::uint64_t size, std::uint32_t* outBgra, std::uint64_t capacityPixels, std::uint32_t* width, std::uint32_t* height)
{
    if (!data || !width || !height) return rfb200::setError(RF_ERROR_INVALID_ARGUMENT, "rf_texture_from_memory: null argument");
    try
    {
        std::vector<std::uint8_t> rgba;
        std::string               error;
        if (!rfb200::decodeImageRgba8(static_cast<const std::uint8_t*>(data), size, rgba, *width, *height, error))
            return rfb200::setError(RF_ERROR_FORMAT, "Failed to decode image: %s", error.c_str());
        const std::uint64_t n = static_cast<std::uint64_t>(*width) * *height;
        if (!outBgra) return RF_OK; // size query
        if (capacityPixels < n) return rfb200::setError(RF_ERROR_INVALID_ARGUMENT, "rf_texture_from_memory: need room for %llu pixels", (unsigned long long)n);
        for (std::uint64_t i = 0; i < n; ++i)
        {
            const std::uint32_t r = rgba[4 * i], g = rgba[4 * i + 1], b = rgba[4 * i + 2];
            outBgra[i] = b | (g << 8) | (r << 16) | (255u << 24);
        }
        return RF_OK;
    }
    catch (const std::exception& e)
    {
        return rfb200::setError(RF_ERROR_IO, "rf_texture_from_memory: %s", e.what());
    }
}
