// Thumbnail JPEG on the GPU: replaces save_frame_as_jpeg (transvideo.c:164-198: one libavcodec MJPEG frame,
// YUVJ420P, written to /var/www/html/thumbnail<N>.jpg every 150th frame).  Baseline sequential JPEG, 4:2:0, the three planes
// taken as Y/Cb/Cr as they are (the reference hands its thumbnail planes to the encoder as full-range YUVJ420P the same way).
//
// What is kept from the reference's encoder: the picture layout (one quantisation table for all components, MCU = 2x2 luma
// + Cb + Cr) and the TABLE: libavcodec's MJPEG quantises with ff_mpeg1_default_intra_matrix scaled by the frame's qscale,
// q[i] = (m[i] * qscale) >> 3, DC entry 8 -- read out of the DQT segment the real encoder of this image writes
// (oracle/avcodec_so.py; it picks qscale 2-3 for natural content at the reference's 250 kb/s).  What is NOT reproduced: its
// SIMD forward DCT, its rate control and its Huffman table construction (un-vendored FFmpeg;
// no bit-exact oracle exists for them) -- this encoder is the public integer DCT of the IJG library (jfdctint: 13-bit
// constants, results scaled by 8; libavcodec's own C fallback), libavcodec's quantiser arithmetic (below) and, like
// libavcodec, the picture's own optimal Huffman tables (built on the host from a symbol histogram the GPU counts, by the
// T.81 K.2 / IJG procedure; libavcodec uses a package-merge variant), bit-exact to the plain C restatement in
// oracle/ref_jpeg.c and compared with the real libavcodec output at the coefficient level (every DC, 99.8 % of the AC levels
// identical), by decoding both, and by size (within 0.3 %: 5 133 against 5 128 bytes for a thumbnail) (tests).
//
// Work split (every stage one thread per 8x8 block unless noted), bitstream assembled in parallel:
//   jpeg_block_dct   : load (edge-replicated), level shift, forward DCT, quantise, zigzag -> coef[block][64]
//   jpeg_block_hist  : histogram of the picture's Huffman symbols (atomic adds) -> the host builds the OPTIMAL tables of this
//                      picture from it (IJG jpeg_gen_optimal_table), as libavcodec's MJPEG encoder does by default
//   jpeg_block_bits  : entropy-code LENGTH of the block (DC differences need the previous block of the component only)
//   (exclusive scan of the lengths -> bit position of every block; one CTA)
//   jpeg_block_emit  : codes written at the block's bit position with atomic ORs into the zeroed stream
//   jpeg_stuff_*     : 0xFF byte stuffing as count / scan / scatter, padding of the last byte with 1 bits, EOI
#pragma once
#include "ladder_core.cuh"

namespace ottgpu {

#if defined(OTTGPU_EMULATE)
inline void ott_atomic_or(uint32_t *p, uint32_t v) { *p |= v; }
inline void ott_atomic_add(uint32_t *p, uint32_t v) { *p += v; }
#else
OTT_DEV void ott_atomic_or(uint32_t *p, uint32_t v) { atomicOr(p, v); }
OTT_DEV void ott_atomic_add(uint32_t *p, uint32_t v) { atomicAdd(p, v); }
#endif

// natural index of the k-th zigzag coefficient
OTT_DEV int jpeg_zigzag(int k)
{
    const unsigned char z[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                 41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    return z[k];
}

// one 8-point pass of the IJG "islow" forward DCT (jfdctint.c): d[] in, results back into d[]; first = row pass
OTT_DEV void jpeg_fdct8(int (&d)[8], bool first)
{
    const int tmp0 = d[0] + d[7], tmp7 = d[0] - d[7], tmp1 = d[1] + d[6], tmp6 = d[1] - d[6];
    const int tmp2 = d[2] + d[5], tmp5 = d[2] - d[5], tmp3 = d[3] + d[4], tmp4 = d[3] - d[4];
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    const int sh = first ? 11 : 15;                              // CONST_BITS -/+ PASS1_BITS
    const int rnd = 1 << (sh - 1);
    if (first) {
        d[0] = (tmp10 + tmp11) << 2;
        d[4] = (tmp10 - tmp11) << 2;
    } else {
        d[0] = (tmp10 + tmp11 + 2) >> 2;
        d[4] = (tmp10 - tmp11 + 2) >> 2;
    }
    int z1 = (tmp12 + tmp13) * 4433;
    d[2] = (z1 + tmp13 * 6270 + rnd) >> sh;
    d[6] = (z1 - tmp12 * 15137 + rnd) >> sh;
    z1 = tmp4 + tmp7;
    int z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
    const int z5 = (z3 + z4) * 9633;
    const int t4 = tmp4 * 2446, t5 = tmp5 * 16819, t6 = tmp6 * 25172, t7 = tmp7 * 12299;
    z1 *= -7373;
    z2 *= -20995;
    z3 = z3 * -16069 + z5;
    z4 = z4 * -3196 + z5;
    d[7] = (t4 + z1 + z3 + rnd) >> sh;
    d[5] = (t5 + z2 + z4 + rnd) >> sh;
    d[3] = (t6 + z2 + z3 + rnd) >> sh;
    d[1] = (t7 + z1 + z4 + rnd) >> sh;
}

// which plane / where: block b of the scan
OTT_DEV void jpeg_block_place(const JpegJob &j, int b, int &comp, int &bx, int &by)
{
    const int mcu = b / 6, k = b - 6 * mcu, mx = mcu % j.mcus_x, my = mcu / j.mcus_x;
    if (k < 4) { comp = 0; bx = 2 * mx + (k & 1); by = 2 * my + (k >> 1); }
    else { comp = k - 3; bx = mx; by = my; }
}

OTT_DEV void jpeg_block_dct(const JpegJob &j, int b)
{
    if (b >= j.n_blocks) return;
    int comp, bx, by;
    jpeg_block_place(j, b, comp, bx, by);
    const int pw = comp ? (j.width + 1) >> 1 : j.width, ph = comp ? (j.height + 1) >> 1 : j.height;
    const uint8_t *p = j.plane[comp];
    const int pitch = j.pitch[comp];
    int blk[64];
    for (int r = 0; r < 8; ++r) {
        const int y = ott_min(8 * by + r, ph - 1);
        int d[8];
        OTT_UNROLL
        for (int c = 0; c < 8; ++c) d[c] = (int)ott_ldg(p + (size_t)y * pitch + ott_min(8 * bx + c, pw - 1)) - 128;
        jpeg_fdct8(d, true);
        OTT_UNROLL
        for (int c = 0; c < 8; ++c) blk[8 * r + c] = d[c];
    }
    for (int c = 0; c < 8; ++c) {
        int d[8];
        OTT_UNROLL
        for (int r = 0; r < 8; ++r) d[r] = blk[8 * r + c];
        jpeg_fdct8(d, false);
        OTT_UNROLL
        for (int r = 0; r < 8; ++r) blk[8 * r + c] = d[r];
    }
    // Quantiser = libavcodec's for MJPEG (mpegvideo_enc.c convert_matrix / dct_quantize, 16-bit form): AC levels
    // ((|x| + bias16) * qmat16) >> 16 with qmat16 = (2 << 16) / (16 q) and the 3/8 intra bias, DC (x + 32) >> 6 (its DC scale
    // is 8 whatever the qscale; floor).  Checked against the real encoder: every DC and 99.8 % of the AC levels are identical
    // (the rest is its SIMD forward DCT rounding differently from jfdctint), tests/test_parity_jpeg.py.
    int16_t *out = j.coef + (size_t)b * 64;
    for (int k = 0; k < 64; ++k) {
        const int n = jpeg_zigzag(k);
        const int v = blk[n];
        int lv;
        if (k == 0) {
            lv = ott_max(ott_min((v + 32) >> 6, 1023), -1023);
        } else {
            const int a = v < 0 ? -v : v;
            const int qd = ott_min((int)(((uint32_t)(a + (int)j.bias16[n]) * (uint32_t)j.qmat16[n]) >> 16), 1023);   // baseline: AC category <= 10
            lv = v < 0 ? -qd : qd;
        }
        out[k] = (int16_t)lv;
    }
}

OTT_DEV int jpeg_category(int a)       // bits needed for |value|
{
    int s = 0;
    while (a) { ++s; a >>= 1; }
    return s;
}

// DC predictor of block b: the previous block of the same component in scan order (0 for the first)
OTT_DEV int jpeg_dc_pred(const JpegJob &j, int b)
{
    const int mcu = b / 6, k = b - 6 * mcu;
    int prev;
    if (k >= 1 && k <= 3) prev = b - 1;
    else if (mcu == 0) return 0;
    else prev = k == 0 ? b - 3 : b - 6;                          // Y00 follows the previous MCU's Y11; chroma its own plane's block
    return j.coef[(size_t)prev * 64];
}

// Three walks over a block's symbols share one routine: the histogram of the (table, symbol) pairs (first pass: the picture's
// OPTIMAL Huffman tables are built from it on the host, like libavcodec's MJPEG encoder does by default), the code length of
// the block, and the emission of its codes.  kind: 0 DC luma, 1 DC chroma, 2 AC luma, 3 AC chroma.
struct JpegHist {
    uint32_t *hist;              // [4][256]
    OTT_DEV void sym(int kind, int symbol, const JpegHuff &) { ott_atomic_add(hist + 256 * kind + symbol, 1u); }
    OTT_DEV void raw(uint32_t, int) {}
};

struct JpegCount {
    uint32_t n = 0;
    OTT_DEV void sym(int kind, int symbol, const JpegHuff &H) { n += kind < 2 ? H.dc_len[kind][symbol] : H.ac_len[kind - 2][symbol]; }
    OTT_DEV void raw(uint32_t, int len) { n += (uint32_t)len; }
};

struct JpegEmit {
    uint32_t *stream;
    uint32_t pos, limit;
    OTT_DEV void raw(uint32_t code, int len)
    {
        if (len == 0 || pos + (uint32_t)len > limit) { pos += (uint32_t)len; return; }
        const uint32_t w = pos >> 5, o = pos & 31u;
        if (o + (uint32_t)len <= 32u) {
            ott_atomic_or(stream + w, code << (32u - o - (uint32_t)len));
        } else {
            const uint32_t spill = o + (uint32_t)len - 32u;
            ott_atomic_or(stream + w, code >> spill);
            ott_atomic_or(stream + w + 1, code << (32u - spill));
        }
        pos += (uint32_t)len;
    }
    OTT_DEV void sym(int kind, int symbol, const JpegHuff &H)
    {
        if (kind < 2) raw(H.dc_code[kind][symbol], H.dc_len[kind][symbol]);
        else raw(H.ac_code[kind - 2][symbol], H.ac_len[kind - 2][symbol]);
    }
};

template <class Sink>
OTT_DEV void jpeg_block_codes(const JpegJob &j, int b, Sink &sink)
{
    const JpegHuff &H = *j.huff;
    const int tbl = (b % 6) >= 4;
    const int16_t *c = j.coef + (size_t)b * 64;
    const int diff = (int)c[0] - jpeg_dc_pred(j, b);
    int s = jpeg_category(diff < 0 ? -diff : diff);
    sink.sym(tbl, s, H);
    if (s) sink.raw((uint32_t)(diff < 0 ? diff - 1 : diff) & ((1u << s) - 1u), s);
    int run = 0;
    for (int k = 1; k < 64; ++k) {
        const int v = c[k];
        if (v == 0) { ++run; continue; }
        while (run > 15) { sink.sym(2 + tbl, 0xF0, H); run -= 16; }
        s = jpeg_category(v < 0 ? -v : v);
        sink.sym(2 + tbl, (run << 4) | s, H);
        sink.raw((uint32_t)(v < 0 ? v - 1 : v) & ((1u << s) - 1u), s);
        run = 0;
    }
    if (run) sink.sym(2 + tbl, 0, H);
}

OTT_DEV void jpeg_block_hist(const JpegJob &j, int b)
{
    if (b >= j.n_blocks) return;
    JpegHist h{j.hist};
    jpeg_block_codes(j, b, h);
}

OTT_DEV void jpeg_block_bits(const JpegJob &j, int b)
{
    if (b >= j.n_blocks) return;
    JpegCount cnt;
    jpeg_block_codes(j, b, cnt);
    j.bits[b] = cnt.n;
}

OTT_DEV void jpeg_block_emit(const JpegJob &j, int b)
{
    if (b >= j.n_blocks) return;
    JpegEmit e{j.stream, j.bits[b], j.stream_words * 32u};
    jpeg_block_codes(j, b, e);
}

// byte k of the entropy-coded stream (the last byte is padded with 1 bits)
OTT_DEV uint32_t jpeg_stream_byte(const JpegJob &j, uint32_t k, uint32_t total_bits)
{
    uint32_t v = (j.stream[k >> 2] >> (24u - 8u * (k & 3u))) & 255u;
    if (8u * k + 8u > total_bits) v |= 0xFFu >> (total_bits & 7u);
    return v;
}

constexpr int kJpegChunk = 64;         // stream bytes per thread in the stuffing pass

OTT_DEV void jpeg_stuff_count(const JpegJob &j, int chunk)
{
    const uint32_t total_bits = j.bits[j.n_blocks], nbytes = (total_bits + 7u) >> 3;
    const uint32_t k0 = (uint32_t)chunk * kJpegChunk;
    if (k0 >= nbytes) return;
    uint32_t n = 0;
    for (uint32_t k = k0; k < ott_min((int)(k0 + kJpegChunk), (int)nbytes); ++k) n += jpeg_stream_byte(j, k, total_bits) == 0xFFu;
    j.ff_count[chunk] = n;
}

// ff_count[] holds the exclusive scan of the counts; writes the chunk's bytes (and the stuffed zeros) to their place
OTT_DEV void jpeg_stuff_write(const JpegJob &j, int chunk)
{
    const uint32_t total_bits = j.bits[j.n_blocks], nbytes = (total_bits + 7u) >> 3;
    const uint32_t k0 = (uint32_t)chunk * kJpegChunk;
    if (k0 >= nbytes) return;
    uint32_t at = (uint32_t)j.header_bytes + k0 + j.ff_count[chunk];
    for (uint32_t k = k0; k < ott_min((int)(k0 + kJpegChunk), (int)nbytes); ++k) {
        const uint32_t v = jpeg_stream_byte(j, k, total_bits);
        if (at < j.out_capacity) j.out[at] = (uint8_t)v;
        ++at;
        if (v == 0xFFu) {
            if (at < j.out_capacity) j.out[at] = 0;
            ++at;
        }
    }
}

// chunk 0 also closes the file: EOI after the stuffed data, total size and overflow flag for the host
OTT_DEV void jpeg_finish(const JpegJob &j, int chunks)
{
    const uint32_t total_bits = j.bits[j.n_blocks], nbytes = (total_bits + 7u) >> 3;
    const uint32_t at = (uint32_t)j.header_bytes + nbytes + j.ff_count[chunks];
    const bool fits = at + 2u <= j.out_capacity && total_bits <= j.stream_words * 32u;
    if (fits) { j.out[at] = 0xFF; j.out[at + 1] = 0xD9; }
    j.result[0] = at + 2u;
    j.result[1] = fits ? 0u : 1u;
}

}  // namespace ottgpu
