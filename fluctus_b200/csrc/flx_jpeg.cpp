// flx_jpeg.cpp -- JPEG texture decoding for flx_image_load: host code, no dependency.
//
// The reference reads textures through DevIL's ilLoadImage (src/texture.cpp:16-40), which hands JPEG files to the IJG decoder with
// its defaults.  DevIL is not in this image, so the pin is the same decoder family as shipped with Pillow (libjpeg-turbo, IJG 6b
// behaviour): tests/test_scene_io_cpu.py demands BYTE-IDENTICAL pixels on the reference's eleven Country-Kitchen JPEGs and on
// generated files covering every path below.  Identical bytes need the identical arithmetic, so three pieces restate the IJG
// algorithms step for step (the published ones: jidctint.c "islow" inverse DCT after Loeffler, Ligtenberg and Moschytz with 13-bit
// constants; jdsample.c "fancy" triangle-filter chroma upsampling; jdcolor.c fixed-point YCbCr -> RGB); the bit-stream side
// (markers, Huffman, progressive refinement) only has to obey ITU-T T.81.
//
// Supported: baseline and extended sequential (SOF0 / SOF1) and progressive (SOF2) Huffman streams, 8-bit samples, 1 component
// (grey) or 3 (YCbCr, or RGB when an Adobe marker says so), any sampling factors with luma as the largest (fancy upsampling for
// 2x1, 1x2 and 2x2, replication otherwise), restart intervals, interleaved and per-component scans.  Refused with a message:
// arithmetic coding, lossless, hierarchical, 12-bit, CMYK / 4 components.
// Output: RGBA8, row 0 = BOTTOM row (the reference sets DevIL's origin to lower-left, src/main.cpp:69-71), alpha 255.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace
{
const int kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                         35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable
{
    bool defined = false;
    // canonical decoding (T.81 Annex F.2.2.3): per code length the smallest code, the largest code and the index of its first value
    int mincode[17], maxcode[18], valptr[17];
    unsigned char values[256];
    bool build(const unsigned char counts[16], const unsigned char *vals, int nvals)
    {
        int code = 0, k = 0;
        for (int len = 1; len <= 16; len++)
        {
            valptr[len] = k;
            mincode[len] = code;
            code += counts[len - 1];
            k += counts[len - 1];
            maxcode[len] = counts[len - 1] ? code - 1 : -1;
            if (code > (1 << len))
                return false; // over-subscribed
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        if (k != nvals || k > 256)
            return false;
        std::memcpy(values, vals, (size_t)k);
        defined = true;
        return true;
    }
};

struct Component
{
    int id = 0, h = 1, v = 1, tq = 0;
    int td = 0, ta = 0;           // Huffman table selectors of the current scan
    int blocksW = 0, blocksH = 0; // blocks covering the component's own size (non-interleaved scans)
    int allocW = 0, allocH = 0;   // blocks allocated: padded to whole MCUs
    int width = 0, height = 0;    // downsampled size in samples
    std::vector<int16_t> coef;    // allocW * allocH * 64, natural (not zigzag) order, not dequantised
    std::vector<unsigned char> plane; // allocW*8 x allocH*8 samples after the inverse DCT
    int pred = 0;                 // DC predictor
};

struct BitReader
{
    const unsigned char *p, *end;
    uint32_t acc = 0;
    int n = 0;
    bool hitMarker = false; // ran into a marker (or the end of the data): further bits read as zero, like the IJG decoder does
    void fill()
    {
        while (n <= 24)
        {
            int byte = 0;
            if (!hitMarker && p < end)
            {
                byte = *p;
                if (byte == 0xff)
                {
                    if (p + 1 < end && p[1] == 0x00)
                        p += 2;
                    else
                    {
                        hitMarker = true; // leave p ON the marker
                        byte = 0;
                    }
                }
                else
                    p++;
            }
            else
                hitMarker = true;
            acc |= (uint32_t)byte << (24 - n);
            n += 8;
        }
    }
    int bits(int k) // k in 0..16
    {
        if (k == 0)
            return 0;
        if (n < k)
            fill();
        const int v = (int)(acc >> (32 - k));
        acc <<= k;
        n -= k;
        return v;
    }
    int bit() { return bits(1); }
    void reset()
    {
        acc = 0;
        n = 0;
        hitMarker = false;
    }
};

int extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; } // T.81 F.2.2.1

struct Decoder
{
    std::string err;
    const unsigned char *data = nullptr;
    size_t size = 0;
    int width = 0, height = 0, ncomp = 0, hmax = 1, vmax = 1, mcusX = 0, mcusY = 0;
    bool progressive = false, sawJFIF = false, sawAdobe = false;
    int adobeTransform = 0, restartInterval = 0;
    uint16_t quant[4][64];
    bool quantDefined[4] = {false, false, false, false};
    HuffTable dc[4], ac[4];
    Component comp[3];
    BitReader br;
    int eobrun = 0;

    bool fail(const char *why)
    {
        if (err.empty())
            err = why;
        return false;
    }

    int decodeSymbol(const HuffTable &t)
    {
        int code = br.bit();
        for (int len = 1; len <= 16; len++)
        {
            if (t.maxcode[len] >= 0 && code <= t.maxcode[len] && code >= t.mincode[len])
                return t.values[t.valptr[len] + code - t.mincode[len]];
            code = (code << 1) | br.bit();
        }
        return -1; // not a code of this table (corrupt stream)
    }

    // ---- one 8x8 block, sequential mode (T.81 F.2.2)
    bool blockSequential(Component &c, int16_t *blk)
    {
        int t = decodeSymbol(dc[c.td]);
        if (t < 0 || t > 16)
            return fail("bad DC code");
        const int diff = t ? extend(br.bits(t), t) : 0;
        c.pred += diff;
        blk[0] = (int16_t)c.pred;
        for (int k = 1; k < 64;)
        {
            const int rs = decodeSymbol(ac[c.ta]);
            if (rs < 0)
                return fail("bad AC code");
            const int r = rs >> 4, s = rs & 15;
            if (s == 0)
            {
                if (r != 15)
                    break; // EOB
                k += 16;
                continue;
            }
            k += r;
            if (k > 63)
                return fail("AC coefficient index out of range");
            blk[kZigzag[k]] = (int16_t)extend(br.bits(s), s);
            k++;
        }
        return true;
    }

    // ---- progressive mode (T.81 G.1.2)
    bool blockDCFirst(Component &c, int16_t *blk, int al)
    {
        const int t = decodeSymbol(dc[c.td]);
        if (t < 0 || t > 16)
            return fail("bad DC code");
        const int diff = t ? extend(br.bits(t), t) : 0;
        c.pred += diff;
        blk[0] = (int16_t)(c.pred * (1 << al));
        return true;
    }
    void blockDCRefine(int16_t *blk, int al)
    {
        if (br.bit())
            blk[0] |= (int16_t)(1 << al);
    }
    bool blockACFirst(Component &c, int16_t *blk, int ss, int se, int al)
    {
        if (eobrun > 0)
        {
            eobrun--;
            return true;
        }
        for (int k = ss; k <= se;)
        {
            const int rs = decodeSymbol(ac[c.ta]);
            if (rs < 0)
                return fail("bad AC code");
            const int r = rs >> 4, s = rs & 15;
            if (s == 0)
            {
                if (r < 15)
                {
                    eobrun = (1 << r) - 1;
                    if (r)
                        eobrun += br.bits(r);
                    break;
                }
                k += 16;
                continue;
            }
            k += r;
            if (k > 63)
                return fail("AC coefficient index out of range");
            blk[kZigzag[k]] = (int16_t)(extend(br.bits(s), s) * (1 << al));
            k++;
        }
        return true;
    }
    bool blockACRefine(Component &c, int16_t *blk, int ss, int se, int al)
    {
        const int p1 = 1 << al, m1 = -(1 << al);
        int k = ss;
        if (eobrun == 0)
        {
            for (; k <= se;)
            {
                const int rs = decodeSymbol(ac[c.ta]);
                if (rs < 0)
                    return fail("bad AC code");
                int r = rs >> 4;
                const int s = rs & 15;
                int value = 0;
                if (s == 0)
                {
                    if (r < 15)
                    {
                        eobrun = 1 << r;
                        if (r)
                            eobrun += br.bits(r);
                        break; // the rest of the band is handled as part of the EOB run below
                    }
                    // r == 15: skip 16 zero-history coefficients
                }
                else
                {
                    if (s != 1)
                        return fail("bad refinement code");
                    value = br.bit() ? p1 : m1;
                }
                // advance over already-nonzero coefficients (each takes a correction bit) and r zero-history ones
                for (; k <= se; k++)
                {
                    int16_t &coef = blk[kZigzag[k]];
                    if (coef != 0)
                    {
                        if (br.bit() && (coef & p1) == 0)
                            coef = (int16_t)(coef >= 0 ? coef + p1 : coef + m1);
                    }
                    else
                    {
                        if (--r < 0)
                            break;
                    }
                }
                if (value && k <= se)
                    blk[kZigzag[k]] = (int16_t)value;
                k++;
            }
        }
        if (eobrun > 0)
        {
            // inside an end-of-band run: only correction bits for the coefficients that are already nonzero
            for (; k <= se; k++)
            {
                int16_t &coef = blk[kZigzag[k]];
                if (coef != 0 && br.bit() && (coef & p1) == 0)
                    coef = (int16_t)(coef >= 0 ? coef + p1 : coef + m1);
            }
            eobrun--;
        }
        return true;
    }

    // restart marker at the current position (after the bit reader has been byte-aligned)
    bool restart(int expected)
    {
        br.reset();
        const unsigned char *p = br.p;
        while (p < br.end && *p != 0xff)
            p++; // garbage before the marker: skip it
        while (p + 1 < br.end && p[1] == 0xff)
            p++; // fill bytes
        if (p + 1 >= br.end || p[1] != 0xd0 + (expected & 7))
            return fail("missing restart marker");
        br.p = p + 2;
        for (int i = 0; i < ncomp; i++)
            comp[i].pred = 0;
        eobrun = 0;
        return true;
    }

    bool decodeScan(const unsigned char *seg, int len, const unsigned char *entropy)
    {
        const int ns = seg[0];
        if (ns < 1 || ns > ncomp || len != 4 + 2 * ns)
            return fail("bad SOS header");
        Component *sc[3];
        for (int i = 0; i < ns; i++)
        {
            sc[i] = nullptr;
            for (int j = 0; j < ncomp; j++)
                if (comp[j].id == seg[1 + 2 * i])
                    sc[i] = &comp[j];
            if (!sc[i])
                return fail("scan names an unknown component");
            sc[i]->td = seg[2 + 2 * i] >> 4;
            sc[i]->ta = seg[2 + 2 * i] & 15;
            if (sc[i]->td > 3 || sc[i]->ta > 3)
                return fail("bad Huffman table selector");
        }
        const int ss = seg[1 + 2 * ns], se = seg[2 + 2 * ns], ah = seg[3 + 2 * ns] >> 4, al = seg[3 + 2 * ns] & 15;
        if (progressive)
        {
            if (ss > se || se > 63 || (ss == 0 && se != 0) || (ss > 0 && ns != 1) || al > 13 || (ah != 0 && ah != al + 1))
                return fail("bad progressive scan parameters");
        }
        else if (ss != 0 || se != 63 || ah != 0 || al != 0)
            return fail("bad sequential scan parameters");
        for (int i = 0; i < ns; i++)
        {
            const bool needDC = !progressive || ss == 0, needAC = !progressive || ss > 0;
            if ((needDC && !(progressive && ah) && !dc[sc[i]->td].defined) || (needAC && !ac[sc[i]->ta].defined))
                return fail("scan uses an undefined Huffman table");
            sc[i]->pred = 0;
        }
        br.p = entropy;
        br.end = data + size;
        br.reset();
        eobrun = 0;

        auto decodeBlock = [&](Component &c, int bx, int by) -> bool {
            int16_t *blk = &c.coef[((size_t)by * c.allocW + bx) * 64];
            if (!progressive)
                return blockSequential(c, blk);
            if (ss == 0)
            {
                if (ah == 0)
                    return blockDCFirst(c, blk, al);
                blockDCRefine(blk, al);
                return true;
            }
            return ah == 0 ? blockACFirst(c, blk, ss, se, al) : blockACRefine(c, blk, ss, se, al);
        };

        int untilRestart = restartInterval, nextRestart = 0;
        auto maybeRestart = [&](bool more) -> bool {
            if (restartInterval == 0 || --untilRestart > 0 || !more)
                return true;
            untilRestart = restartInterval;
            return restart(nextRestart++);
        };
        if (ns == 1)
        {
            Component &c = *sc[0];
            const long total = (long)c.blocksW * c.blocksH;
            long done = 0;
            for (int by = 0; by < c.blocksH; by++)
                for (int bx = 0; bx < c.blocksW; bx++)
                {
                    if (!decodeBlock(c, bx, by))
                        return false;
                    if (!maybeRestart(++done < total))
                        return false;
                }
        }
        else
        {
            const long total = (long)mcusX * mcusY;
            long done = 0;
            for (int my = 0; my < mcusY; my++)
                for (int mx = 0; mx < mcusX; mx++)
                {
                    for (int i = 0; i < ns; i++)
                        for (int v = 0; v < sc[i]->v; v++)
                            for (int h = 0; h < sc[i]->h; h++)
                                if (!decodeBlock(*sc[i], mx * sc[i]->h + h, my * sc[i]->v + v))
                                    return false;
                    if (!maybeRestart(++done < total))
                        return false;
                }
        }
        return true;
    }

    // ---- IJG "islow" inverse DCT (jidctint.c), 13-bit constants, two passes with a 2-bit intermediate scale
    typedef long long I64; // the IJG code computes in `long`; 64 bits also keep hostile coefficient data free of signed overflow
    static int descale(I64 x, int n) { return (int)((x + ((I64)1 << (n - 1))) >> n); }
    static unsigned char rangeLimit(int x)
    {
        // the IJG post-IDCT table: index (x & 1023) into { x+128 for -128..127, 255 up to 511, 0 from -512 } (wraps beyond)
        x &= 1023;
        if (x < 128)
            return (unsigned char)(x + 128);
        if (x < 512)
            return 255;
        if (x < 896)
            return 0;
        return (unsigned char)(x - 896);
    }
    static void idct(const int16_t *in, const uint16_t *q, unsigned char *out, int stride)
    {
        const I64 C_0_298 = 2446, C_0_390 = 3196, C_0_541 = 4433, C_0_765 = 6270, C_0_899 = 7373, C_1_175 = 9633, C_1_501 = 12299, C_1_847 = 15137, C_1_961 = 16069,
                  C_2_053 = 16819, C_2_562 = 20995, C_3_072 = 25172;
        const int CONST_BITS = 13, PASS1_BITS = 2;
        int ws[64];
        for (int col = 0; col < 8; col++)
        {
            const int16_t *i = in + col;
            const uint16_t *qq = q + col;
            if (i[8] == 0 && i[16] == 0 && i[24] == 0 && i[32] == 0 && i[40] == 0 && i[48] == 0 && i[56] == 0)
            {
                const int dcval = (i[0] * qq[0]) * (1 << PASS1_BITS);
                for (int r = 0; r < 8; r++)
                    ws[r * 8 + col] = dcval;
                continue;
            }
            I64 z2 = i[16] * qq[16], z3 = i[48] * qq[48];
            I64 z1 = (z2 + z3) * C_0_541;
            I64 tmp2 = z1 + z3 * (-C_1_847), tmp3 = z1 + z2 * C_0_765;
            z2 = i[0] * qq[0];
            z3 = i[32] * qq[32];
            I64 tmp0 = (z2 + z3) * (1 << CONST_BITS), tmp1 = (z2 - z3) * (1 << CONST_BITS);
            const I64 tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            tmp0 = i[56] * qq[56];
            tmp1 = i[40] * qq[40];
            tmp2 = i[24] * qq[24];
            tmp3 = i[8] * qq[8];
            z1 = tmp0 + tmp3;
            z2 = tmp1 + tmp2;
            z3 = tmp0 + tmp2;
            I64 z4 = tmp1 + tmp3;
            const I64 z5 = (z3 + z4) * C_1_175;
            tmp0 *= C_0_298;
            tmp1 *= C_2_053;
            tmp2 *= C_3_072;
            tmp3 *= C_1_501;
            z1 *= -C_0_899;
            z2 *= -C_2_562;
            z3 *= -C_1_961;
            z4 *= -C_0_390;
            z3 += z5;
            z4 += z5;
            tmp0 += z1 + z3;
            tmp1 += z2 + z4;
            tmp2 += z2 + z3;
            tmp3 += z1 + z4;
            ws[0 * 8 + col] = descale(tmp10 + tmp3, CONST_BITS - PASS1_BITS);
            ws[7 * 8 + col] = descale(tmp10 - tmp3, CONST_BITS - PASS1_BITS);
            ws[1 * 8 + col] = descale(tmp11 + tmp2, CONST_BITS - PASS1_BITS);
            ws[6 * 8 + col] = descale(tmp11 - tmp2, CONST_BITS - PASS1_BITS);
            ws[2 * 8 + col] = descale(tmp12 + tmp1, CONST_BITS - PASS1_BITS);
            ws[5 * 8 + col] = descale(tmp12 - tmp1, CONST_BITS - PASS1_BITS);
            ws[3 * 8 + col] = descale(tmp13 + tmp0, CONST_BITS - PASS1_BITS);
            ws[4 * 8 + col] = descale(tmp13 - tmp0, CONST_BITS - PASS1_BITS);
        }
        for (int row = 0; row < 8; row++)
        {
            const int *w = ws + row * 8;
            unsigned char *o = out + (size_t)row * stride;
            I64 z2 = w[2], z3 = w[6];
            I64 z1 = (z2 + z3) * C_0_541;
            I64 tmp2 = z1 + z3 * (-C_1_847), tmp3 = z1 + z2 * C_0_765;
            I64 tmp0 = ((I64)w[0] + w[4]) * (1 << CONST_BITS), tmp1 = ((I64)w[0] - w[4]) * (1 << CONST_BITS);
            const I64 tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            tmp0 = w[7];
            tmp1 = w[5];
            tmp2 = w[3];
            tmp3 = w[1];
            z1 = tmp0 + tmp3;
            z2 = tmp1 + tmp2;
            z3 = tmp0 + tmp2;
            I64 z4 = tmp1 + tmp3;
            const I64 z5 = (z3 + z4) * C_1_175;
            tmp0 *= C_0_298;
            tmp1 *= C_2_053;
            tmp2 *= C_3_072;
            tmp3 *= C_1_501;
            z1 *= -C_0_899;
            z2 *= -C_2_562;
            z3 *= -C_1_961;
            z4 *= -C_0_390;
            z3 += z5;
            z4 += z5;
            tmp0 += z1 + z3;
            tmp1 += z2 + z4;
            tmp2 += z2 + z3;
            tmp3 += z1 + z4;
            const int S = CONST_BITS + PASS1_BITS + 3;
            o[0] = rangeLimit(descale(tmp10 + tmp3, S));
            o[7] = rangeLimit(descale(tmp10 - tmp3, S));
            o[1] = rangeLimit(descale(tmp11 + tmp2, S));
            o[6] = rangeLimit(descale(tmp11 - tmp2, S));
            o[2] = rangeLimit(descale(tmp12 + tmp1, S));
            o[5] = rangeLimit(descale(tmp12 - tmp1, S));
            o[3] = rangeLimit(descale(tmp13 + tmp0, S));
            o[4] = rangeLimit(descale(tmp13 - tmp0, S));
        }
    }

    // ---- chroma upsampling to full resolution (jdsample.c).  `row(y)` clamps to the component's real rows, which is what the IJG
    // main controller's context rows amount to at the top and bottom of the image.
    void upsample(const Component &c, std::vector<unsigned char> &full)
    {
        const int stride = c.allocW * 8, W = c.width, H = c.height;
        const int hx = hmax / c.h, vx = vmax / c.v;
        const int outW = W * hx;
        full.assign((size_t)outW * H * vx, 0);
        auto row = [&](int y) { return &c.plane[(size_t)(y < 0 ? 0 : (y >= H ? H - 1 : y)) * stride]; };
        if (hx == 1 && vx == 1)
        {
            for (int y = 0; y < H; y++)
                std::memcpy(&full[(size_t)y * outW], row(y), (size_t)W);
        }
        else if (hx == 2 && vx == 1) // h2v1_fancy_upsample
        {
            for (int y = 0; y < H; y++)
            {
                const unsigned char *in = row(y);
                unsigned char *out = &full[(size_t)y * outW];
                if (W == 1)
                {
                    out[0] = out[1] = in[0];
                    continue;
                }
                out[0] = in[0];
                out[1] = (unsigned char)((in[0] * 3 + in[1] + 2) >> 2);
                for (int x = 1; x < W - 1; x++)
                {
                    const int v = in[x] * 3;
                    out[2 * x] = (unsigned char)((v + in[x - 1] + 1) >> 2);
                    out[2 * x + 1] = (unsigned char)((v + in[x + 1] + 2) >> 2);
                }
                out[2 * W - 2] = (unsigned char)((in[W - 1] * 3 + in[W - 2] + 1) >> 2);
                out[2 * W - 1] = in[W - 1];
            }
        }
        else if (hx == 1 && vx == 2) // h1v2_fancy_upsample (libjpeg-turbo)
        {
            for (int y = 0; y < H; y++)
                for (int v = 0; v < 2; v++)
                {
                    const unsigned char *in0 = row(y), *in1 = row(v == 0 ? y - 1 : y + 1);
                    unsigned char *out = &full[(size_t)(2 * y + v) * outW];
                    const int bias = v == 0 ? 1 : 2;
                    for (int x = 0; x < W; x++)
                        out[x] = (unsigned char)((in0[x] * 3 + in1[x] + bias) >> 2);
                }
        }
        else if (hx == 2 && vx == 2) // h2v2_fancy_upsample
        {
            for (int y = 0; y < H; y++)
                for (int v = 0; v < 2; v++)
                {
                    const unsigned char *in0 = row(y), *in1 = row(v == 0 ? y - 1 : y + 1);
                    unsigned char *out = &full[(size_t)(2 * y + v) * outW];
                    if (W == 1)
                    {
                        const int s = in0[0] * 3 + in1[0];
                        out[0] = (unsigned char)((s * 4 + 8) >> 4);
                        out[1] = (unsigned char)((s * 4 + 7) >> 4);
                        continue;
                    }
                    int thiscol = in0[0] * 3 + in1[0], nextcol = in0[1] * 3 + in1[1], lastcol;
                    out[0] = (unsigned char)((thiscol * 4 + 8) >> 4);
                    out[1] = (unsigned char)((thiscol * 3 + nextcol + 7) >> 4);
                    lastcol = thiscol;
                    thiscol = nextcol;
                    for (int x = 1; x < W - 1; x++)
                    {
                        nextcol = in0[x + 1] * 3 + in1[x + 1];
                        out[2 * x] = (unsigned char)((thiscol * 3 + lastcol + 8) >> 4);
                        out[2 * x + 1] = (unsigned char)((thiscol * 3 + nextcol + 7) >> 4);
                        lastcol = thiscol;
                        thiscol = nextcol;
                    }
                    out[2 * W - 2] = (unsigned char)((thiscol * 3 + lastcol + 8) >> 4);
                    out[2 * W - 1] = (unsigned char)((thiscol * 4 + 7) >> 4);
                }
        }
        else // int_upsample: plain replication
        {
            for (int y = 0; y < H * vx; y++)
            {
                const unsigned char *in = row(y / vx);
                unsigned char *out = &full[(size_t)y * outW];
                for (int x = 0; x < outW; x++)
                    out[x] = in[x / hx];
            }
        }
    }

    bool run(std::vector<unsigned char> &rgba)
    {
        if (size < 4 || data[0] != 0xff || data[1] != 0xd8)
            return fail("not a JPEG file");
        size_t at = 2;
        bool haveFrame = false, sawScan = false, done = false;
        std::memset(quant, 0, sizeof quant);
        while (!done)
        {
            // next marker
            while (at < size && data[at] != 0xff)
                at++;
            while (at < size && data[at] == 0xff)
                at++;
            if (at >= size)
                break; // no EOI: accept what has been decoded, like the IJG decoder (it warns)
            const int m = data[at++];
            if (m == 0xd9)
                break;
            if (m == 0x01 || (m >= 0xd0 && m <= 0xd7))
                continue;
            if (at + 2 > size)
                return fail("truncated marker segment");
            const int len = (data[at] << 8) | data[at + 1];
            if (len < 2 || at + len > size)
                return fail("truncated marker segment");
            const unsigned char *seg = data + at + 2;
            const int n = len - 2;
            at += len;
            switch (m)
            {
            case 0xc0: case 0xc1: case 0xc2:
            {
                if (haveFrame)
                    return fail("more than one frame header");
                if (n < 6 || seg[0] != 8)
                    return fail("only 8-bit samples are supported");
                progressive = m == 0xc2;
                height = (seg[1] << 8) | seg[2];
                width = (seg[3] << 8) | seg[4];
                ncomp = seg[5];
                if (width == 0 || height == 0)
                    return fail("empty image (or DNL-defined height, which is not supported)");
                if ((uint64_t)width * height > (1ull << 28))
                    return fail("unreasonable image size");
                if (ncomp != 1 && ncomp != 3)
                    return fail("only grey and three-component images are supported (CMYK / YCCK are not)");
                if (n != 6 + 3 * ncomp)
                    return fail("bad frame header");
                for (int i = 0; i < ncomp; i++)
                {
                    comp[i].id = seg[6 + 3 * i];
                    comp[i].h = seg[7 + 3 * i] >> 4;
                    comp[i].v = seg[7 + 3 * i] & 15;
                    comp[i].tq = seg[8 + 3 * i];
                    if (comp[i].h < 1 || comp[i].h > 4 || comp[i].v < 1 || comp[i].v > 4 || comp[i].tq > 3)
                        return fail("bad sampling factors");
                    hmax = comp[i].h > hmax ? comp[i].h : hmax;
                    vmax = comp[i].v > vmax ? comp[i].v : vmax;
                }
                if (ncomp == 1)
                    comp[0].h = comp[0].v = hmax = vmax = 1; // a single component is never interleaved: its factors are irrelevant
                mcusX = (width + 8 * hmax - 1) / (8 * hmax);
                mcusY = (height + 8 * vmax - 1) / (8 * vmax);
                for (int i = 0; i < ncomp; i++)
                {
                    Component &c = comp[i];
                    if (hmax % c.h || vmax % c.v)
                        return fail("fractional sampling ratios are not supported");
                    c.width = (width * c.h + hmax - 1) / hmax;
                    c.height = (height * c.v + vmax - 1) / vmax;
                    c.blocksW = (c.width + 7) / 8;
                    c.blocksH = (c.height + 7) / 8;
                    c.allocW = mcusX * c.h;
                    c.allocH = mcusY * c.v;
                    c.coef.assign((size_t)c.allocW * c.allocH * 64, 0);
                }
                haveFrame = true;
                break;
            }
            case 0xc3: case 0xc5: case 0xc6: case 0xc7: case 0xc9: case 0xca: case 0xcb: case 0xcd: case 0xce: case 0xcf:
                return fail("lossless, hierarchical and arithmetic-coded JPEG are not supported");
            case 0xc4: // DHT
            {
                int p = 0;
                while (p < n)
                {
                    if (p + 17 > n)
                        return fail("bad Huffman table");
                    const int tc = seg[p] >> 4, th = seg[p] & 15;
                    int total = 0;
                    for (int i = 0; i < 16; i++)
                        total += seg[p + 1 + i];
                    if (tc > 1 || th > 3 || p + 17 + total > n || total > 256)
                        return fail("bad Huffman table");
                    if (!(tc ? ac : dc)[th].build(seg + p + 1, seg + p + 17, total))
                        return fail("bad Huffman table");
                    p += 17 + total;
                }
                break;
            }
            case 0xdb: // DQT
            {
                int p = 0;
                while (p < n)
                {
                    const int pq = seg[p] >> 4, tq = seg[p] & 15;
                    if (pq > 1 || tq > 3 || p + 1 + 64 * (pq + 1) > n)
                        return fail("bad quantisation table");
                    for (int i = 0; i < 64; i++)
                        quant[tq][kZigzag[i]] = pq ? (uint16_t)((seg[p + 1 + 2 * i] << 8) | seg[p + 2 + 2 * i]) : seg[p + 1 + i];
                    quantDefined[tq] = true;
                    p += 1 + 64 * (pq + 1);
                }
                break;
            }
            case 0xdd:
                if (n != 2)
                    return fail("bad restart interval");
                restartInterval = (seg[0] << 8) | seg[1];
                break;
            case 0xe0:
                if (n >= 5 && !std::memcmp(seg, "JFIF", 5))
                    sawJFIF = true;
                break;
            case 0xee:
                if (n >= 12 && !std::memcmp(seg, "Adobe", 5))
                {
                    sawAdobe = true;
                    adobeTransform = seg[11];
                }
                break;
            case 0xda: // SOS
            {
                if (!haveFrame)
                    return fail("scan before the frame header");
                if (!decodeScan(seg, n, data + at))
                    return false;
                sawScan = true;
                at = (size_t)(br.p - data); // on the marker that ended the scan (or wherever the entropy decoder stopped)
                break;
            }
            default:
                break; // APPn, COM, ...: skipped
            }
        }
        if (!haveFrame || !sawScan)
            return fail("no image data");
        // dequantise + inverse DCT
        for (int i = 0; i < ncomp; i++)
        {
            Component &c = comp[i];
            if (!quantDefined[c.tq])
                return fail("component uses an undefined quantisation table");
            const int stride = c.allocW * 8;
            c.plane.assign((size_t)stride * c.allocH * 8, 0);
            for (int by = 0; by < c.allocH; by++)
                for (int bx = 0; bx < c.allocW; bx++)
                    idct(&c.coef[((size_t)by * c.allocW + bx) * 64], quant[c.tq], &c.plane[(size_t)by * 8 * stride + (size_t)bx * 8], stride);
            std::vector<int16_t>().swap(c.coef);
        }
        // colour space, the IJG rules (jdapimin.c default_decompress_parms): JFIF -> YCbCr; else Adobe transform 0 -> RGB, 1 -> YCbCr;
        // else component ids 'R','G','B' -> RGB, anything else YCbCr
        bool ycc = true;
        if (ncomp == 3 && !sawJFIF)
        {
            if (sawAdobe)
                ycc = adobeTransform != 0;
            else
                ycc = !(comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B');
        }
        rgba.assign((size_t)width * height * 4, 255);
        if (ncomp == 1)
        {
            const int stride = comp[0].allocW * 8;
            for (int y = 0; y < height; y++)
            {
                const unsigned char *src = &comp[0].plane[(size_t)y * stride];
                unsigned char *dst = &rgba[(size_t)(height - 1 - y) * width * 4];
                for (int x = 0; x < width; x++, dst += 4)
                    dst[0] = dst[1] = dst[2] = src[x];
            }
            return true;
        }
        std::vector<unsigned char> full[3];
        int fullW[3];
        for (int i = 0; i < 3; i++)
        {
            upsample(comp[i], full[i]);
            fullW[i] = comp[i].width * (hmax / comp[i].h);
        }
        auto clamp255 = [](int v) { return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v)); };
        for (int y = 0; y < height; y++)
        {
            const unsigned char *p0 = &full[0][(size_t)y * fullW[0]], *p1 = &full[1][(size_t)y * fullW[1]], *p2 = &full[2][(size_t)y * fullW[2]];
            unsigned char *dst = &rgba[(size_t)(height - 1 - y) * width * 4];
            for (int x = 0; x < width; x++, dst += 4)
            {
                if (!ycc)
                {
                    dst[0] = p0[x];
                    dst[1] = p1[x];
                    dst[2] = p2[x];
                    continue;
                }
                // jdcolor.c build_ycc_rgb_table / ycc_rgb_convert: 16-bit fixed point, the rounding constant folded into the Cb->G table
                const int Y = p0[x], cb = p1[x] - 128, cr = p2[x] - 128;
                const int r = Y + ((91881 * cr + 32768) >> 16);
                const int g = Y + ((-22554 * cb + 32768 + -46802 * cr) >> 16);
                const int b = Y + ((116130 * cb + 32768) >> 16);
                dst[0] = clamp255(r);
                dst[1] = clamp255(g);
                dst[2] = clamp255(b);
            }
        }
        return true;
    }
};
} // namespace

// used by flx_image_load (flx_scene_io.cpp)
bool flx_decode_jpeg(const std::string &path, uint32_t &w, uint32_t &h, std::vector<unsigned char> &rgba, std::string &error)
{
    FILE *fp = std::fopen(path.c_str(), "rb");
    if (!fp)
    {
        error = "cannot open " + path;
        return false;
    }
    std::fseek(fp, 0, SEEK_END);
    const long size = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    std::vector<unsigned char> file(size > 0 ? (size_t)size : 0);
    const bool ok = file.empty() || std::fread(file.data(), 1, file.size(), fp) == file.size();
    std::fclose(fp);
    if (!ok)
    {
        error = path + ": read error";
        return false;
    }
    Decoder d;
    d.data = file.data();
    d.size = file.size();
    if (!d.run(rgba))
    {
        error = path + ": " + (d.err.empty() ? "corrupt JPEG data" : d.err);
        return false;
    }
    w = (uint32_t)d.width;
    h = (uint32_t)d.height;
    return true;
}
