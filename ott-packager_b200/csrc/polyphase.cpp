// See polyphase.h.  Fixed-point bicubic (B=0, C=0.6) / double-precision lanczos-3 window evaluation, near-zero tap
// trimming, alignment padding, border folding and error-diffused normalisation, in that order.
#include "polyphase.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace ottgpu {
namespace {

using i64 = long long;

struct RawRows {                 // un-normalised 64-bit taps, `width` per output sample
    int width = 0;
    std::vector<i64> v;
    std::vector<int32_t> pos;
    i64 *row(int i) { return v.data() + (size_t)i * width; }
};

inline int floor_log2(unsigned x) { int r = 0; while (x >>= 1) ++r; return r; }

// weight of one tap at fixed-point distance d (30 fractional bits, already divided by the downscale ratio)
inline i64 bicubic_weight(i64 d, i64 fone)
{
    constexpr i64 B = 0, C = (i64)(0.6 * (1 << 24));
    if (d >= (1LL << 31)) return 0;
    const i64 d2 = (d * d) >> 30, d3 = (d2 * d) >> 30;
    i64 w;
    if (d < (1LL << 30))
        w = (12 * (1 << 24) - 9 * B - 6 * C) * d3 + (-18 * (1 << 24) + 12 * B + 6 * C) * d2 + (6 * (1 << 24) - 2 * B) * (1LL << 30);
    else
        w = (-B - 6 * C) * d3 + (6 * B + 30 * C) * d2 + (-12 * B - 48 * C) * d + (8 * B + 24 * C) * (1LL << 30);
    return w / ((1LL << 54) / fone);
}

inline i64 lanczos_weight(i64 d, i64 fone)
{
    const double a = 3.0, f = (double)d * (1.0 / (1 << 30));
    if (f > a) return 0;
    const double w = d ? std::sin(f * M_PI) * std::sin(f * M_PI / a) / (f * f * M_PI * M_PI / a) : 1.0;
    return (i64)(w * (double)fone);
}

RawRows evaluate_window(int src_n, int dst_n, Kernel k, int inc, i64 fone)
{
    RawRows r;
    r.pos.resize(dst_n);
    if (std::abs(inc - 0x10000) < 10) {                      // axis not scaled: identity taps
        r.width = 1;
        r.v.assign(dst_n, fone);
        for (int i = 0; i < dst_n; ++i) r.pos[i] = i;
        return r;
    }
    const int support = k == Kernel::Lanczos ? 6 : 4;
    int width = inc <= 0x10000 ? 1 + support : 1 + (int)(((i64)support * src_n + dst_n - 1) / dst_n);
    width = std::max(std::min(width, src_n - 2), 1);
    r.width = width;
    r.v.resize((size_t)dst_n * width);
    i64 centre = (i64)inc - 0x10000;                           // position of output 0 in source units * 2^17 ... /2
    for (int i = 0; i < dst_n; ++i, centre += 2LL * inc) {
        int first = (int)((centre - (i64)(width - 2) * 65536) / 131072);   // truncating division
        r.pos[i] = first;
        i64 *row = r.row(i);
        for (int j = 0; j < width; ++j) {
            i64 d = std::llabs((i64)(first + j) * 131072 - centre) << 13;
            if (inc > 0x10000) d = d * dst_n / src_n;
            row[j] = k == Kernel::Bicubic ? bicubic_weight(d, fone) : lanczos_weight(d, fone);
        }
    }
    return r;
}

// Drop leading near-zero taps (sliding the window right) and measure how many trailing taps matter.
int trim(RawRows &r, i64 fone)
{
    const double limit = 0.002 * (double)fone;
    const int n = (int)r.pos.size(), w = r.width;
    int needed = 0;
    for (int i = n - 1; i >= 0; --i) {
        i64 *row = r.row(i);
        i64 acc = 0;
        for (int step = 0; step < w; ++step) {
            acc += std::llabs(row[0]);
            if ((double)acc > limit) break;
            if (i < n - 1 && r.pos[i] >= r.pos[i + 1]) break;   // windows must stay monotone
            std::move(row + 1, row + w, row);
            row[w - 1] = 0;
            ++r.pos[i];
        }
        acc = 0;
        int live = w;
        for (int j = w - 1; j > 0; --j) {
            acc += std::llabs(row[j]);
            if ((double)acc > limit) break;
            --live;
        }
        needed = std::max(needed, live);
    }
    return needed;
}

void fold_borders(i64 *row, int32_t &pos, int taps, int src_n)
{
    if (pos < 0) {                                            // taps left of sample 0 pile onto sample 0
        for (int j = 1; j < taps; ++j) {
            row[std::max(j + pos, 0)] += row[j];
            row[j] = 0;
        }
        pos = 0;
    }
    if (pos + taps > src_n) {                                 // taps right of the last sample pile onto it
        const int shift = pos + std::min(taps - src_n, 0);
        i64 spill = 0;
        for (int j = taps - 1; j >= 0; --j)
            if (pos + j >= src_n) { spill += row[j]; row[j] = 0; }
        for (int j = taps - 1; j >= 0; --j) row[j] = j < shift ? 0 : row[j - shift];
        pos -= shift;
        row[src_n - 1 - pos] += spill;
    }
}

inline i64 div_round(i64 a, i64 b) { return (a >= 0 ? a + (b >> 1) : a - (b >> 1)) / b; }

}  // namespace

PolyphaseTable make_polyphase(int src_n, int dst_n, Kernel k, int align, int one, bool zero_tail)
{
    const int inc = (int)((((i64)src_n << 16) + (dst_n >> 1)) / dst_n);
    const int ratio = src_n / dst_n;
    const i64 fone = 1LL << (54 - std::min(ratio > 0 ? floor_log2((unsigned)ratio) : 0, 8));

    RawRows raw = evaluate_window(src_n, dst_n, k, inc, fone);
    const int needed = trim(raw, fone);
    if (needed == 1 && align == 2) align = 1;
    const int taps = (needed + align - 1) & ~(align - 1);

    PolyphaseTable t;
    t.taps = taps;
    t.n = dst_n;
    t.pos = raw.pos;
    t.coef.assign((size_t)dst_n * taps, 0);
    std::vector<i64> row(taps);
    for (int i = 0; i < dst_n; ++i) {
        const i64 *src = raw.row(i);
        for (int j = 0; j < taps; ++j) {
            row[j] = j < raw.width ? src[j] : 0;
            if (zero_tail && j >= needed) row[j] = 0;
        }
        fold_borders(row.data(), t.pos[i], taps, src_n);
        i64 total = 0;
        for (i64 v : row) total += v;
        total = (total + one / 2) / one;
        if (!total) total = 1;
        i64 carry = 0;
        for (int j = 0; j < taps; ++j) {                      // error diffusion keeps every row summing to `one`
            const i64 v = row[j] + carry;
            const i64 q = div_round(v, total);
            t.coef[(size_t)i * taps + j] = (int16_t)q;
            carry = v - q * total;
        }
    }
    return t;
}

}  // namespace ottgpu
