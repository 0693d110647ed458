// Decode-side converter: what the reference's `decode_converter` SwsContext plus the three row-copy loops produce
// (transvideo.c:2797-2840) - a decoder picture in yuv420p / yuv422p / yuv420p10le / yuv422p10le becomes the packed 8-bit
// 4:2:0 frame the prepare thread expects.  libswscale semantics (same-size conversion, flags = SWS_BICUBIC):
//   equal subsampling, 10 -> 8 bit : planarCopyWrapper, dst = clip_u8((s + d2x2[y&1][x&1]) >> 2)
//   4:2:2 sources                  : luma / horizontal chroma are 1-tap filters (identity; >8 bit: (s<<5 + bayer) >> 7,
//                                    the same function as the 2x2 dither), vertical chroma is the 2:1 bicubic polyphase
//                                    filter, target A rows with the per-tap (t*c)>>16 truncation, Bayer dither for >8 bit.
// One thread = 16 (copy / depth reduction) or 8 (vertical filter) adjacent output samples of one row of one plane.  Shared by the kernel and the CPU emulation used in tests.
#pragma once
#include "ladder_core.cuh"

namespace ottgpu {

constexpr int kConvertGroup = 16;      // samples per thread on the copy / depth-reduction planes
constexpr int kConvertVGroup = 8;      // samples per thread on vertically filtered planes (half as wide: same grid)

// N adjacent samples of one row; fast: the group is moved as 8/16-byte vectors
template <typename S, int N>
OTT_DEV void convert_load(const uint8_t *row, int x0, int n, bool fast, int (&v)[N])
{
    if (fast) {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(row + sizeof(S) * x0);
        constexpr int kWords = N * (int)sizeof(S) / 4;
        uint32_t q[kWords];
        if (kWords == 2) {
            const uint2 a = ott_ldg(reinterpret_cast<const uint2 *>(w));
            q[0] = a.x; q[1] = a.y;
        } else {
            OTT_UNROLL
            for (int i = 0; i < kWords / 4; ++i) {
                const uint4 a = ott_ldg(reinterpret_cast<const uint4 *>(w) + i);
                q[4 * i] = a.x; q[4 * i + 1] = a.y; q[4 * i + 2] = a.z; q[4 * i + 3] = a.w;
            }
        }
        OTT_UNROLL
        for (int k = 0; k < N; ++k) {
            if (sizeof(S) == 1) v[k] = ott_byte(q[k >> 2], k & 3);
            else v[k] = (k & 1) ? (int)(q[k >> 1] >> 16) : (int)(q[k >> 1] & 0xffffu);
        }
        return;
    }
    const S *s = reinterpret_cast<const S *>(row) + x0;
    OTT_UNROLL
    for (int k = 0; k < N; ++k) v[k] = k < n ? (int)ott_ldg(s + k) : 0;
}

template <int N>
OTT_DEV void convert_store(uint8_t *row, int x0, int n, bool fast, const int (&v)[N])
{
    if (fast) {
        uint32_t q[N / 4];
        OTT_UNROLL
        for (int i = 0; i < N / 4; ++i)
            q[i] = (uint32_t)v[4 * i] | ((uint32_t)v[4 * i + 1] << 8) | ((uint32_t)v[4 * i + 2] << 16) | ((uint32_t)v[4 * i + 3] << 24);
        if (N == 8) {
            uint2 a;
            a.x = q[0]; a.y = q[1];
            *reinterpret_cast<uint2 *>(row + x0) = a;
        } else {
            OTT_UNROLL
            for (int i = 0; i < N / 16; ++i) {
                uint4 a;
                a.x = q[4 * i]; a.y = q[4 * i + 1]; a.z = q[4 * i + 2]; a.w = q[4 * i + 3];
                reinterpret_cast<uint4 *>(row + x0)[i] = a;
            }
        }
        return;
    }
    OTT_UNROLL
    for (int k = 0; k < N; ++k)
        if (k < n) row[x0 + k] = (uint8_t)v[k];
}

// vertical polyphase over source rows vpos[y] .. vpos[y]+vtaps-1; S = source sample type
template <typename S>
OTT_DEV void convert_vfilt(const ConvertJob &j, const ConvertPlane &P, int x0, int n, int y, int (&out)[kConvertVGroup])
{
    constexpr int N = kConvertVGroup;
    const int shift = sizeof(S) == 1 ? 7 : 5;               // hScale8To15 / hScale16To15 with a 1-tap 1<<14 filter
    const int vtaps = j.vtaps;
    const int16_t *c = j.vc + (size_t)y * vtaps;
    const uint8_t *row = P.src + (size_t)ott_ldg(j.vpos + y) * P.src_pitch;
    const bool fast = P.fast != 0;
    int d[N];                                               // the Bayer row repeats every 8 columns, x0 is a multiple of 8
    OTT_UNROLL
    for (int k = 0; k < N; ++k) d[k] = sizeof(S) == 1 ? 64 : (int)j.bayer[(y & 7) * 8 + ((k + P.dither_offset) & 7)];
    int acc[N], v[N];
    if (vtaps == 1) {
        convert_load<S, N>(row, x0, n, fast, v);
        OTT_UNROLL
        for (int k = 0; k < N; ++k) out[k] = ott_clip(((v[k] << shift) + d[k]) >> 7, 255);
        return;
    }
    const bool approx = j.approx && y < j.exact_from;
    OTT_UNROLL
    for (int k = 0; k < N; ++k) acc[k] = approx ? (d[k] + 8 * (vtaps - 1)) >> 4 : d[k] << 12;
    int t = 0;
    for (; t + 2 <= vtaps; t += 2) {                        // two source rows in flight
        int v2[N];
        const int c0 = ott_ldg(c + t), c1 = ott_ldg(c + t + 1);
        convert_load<S, N>(row, x0, n, fast, v);
        convert_load<S, N>(row + P.src_pitch, x0, n, fast, v2);
        if (approx) {
            OTT_UNROLL
            for (int k = 0; k < N; ++k) acc[k] += (((v[k] << shift) * c0) >> 16) + (((v2[k] << shift) * c1) >> 16);   // pmulhw, paddw
        } else {
            OTT_UNROLL
            for (int k = 0; k < N; ++k) acc[k] += (v[k] << shift) * c0 + (v2[k] << shift) * c1;
        }
        row += 2 * P.src_pitch;
    }
    for (; t < vtaps; ++t) {
        const int ct = ott_ldg(c + t);
        convert_load<S, N>(row, x0, n, fast, v);
        if (approx) {
            OTT_UNROLL
            for (int k = 0; k < N; ++k) acc[k] += ((v[k] << shift) * ct) >> 16;
        } else {
            OTT_UNROLL
            for (int k = 0; k < N; ++k) acc[k] += (v[k] << shift) * ct;
        }
        row += P.src_pitch;
    }
    OTT_UNROLL
    for (int k = 0; k < N; ++k) out[k] = approx ? ott_clip((int)(int16_t)acc[k] >> 3, 255) : ott_clip(acc[k] >> 19, 255);
}

// Thread position (xg, r): r runs over the Y rows, then the U rows, then the V rows of the destination frame.
OTT_DEV void convert_position(const ConvertJob &j, int xg, int r)
{
    int p = 0;
    if (r >= j.plane[0].h) {
        r -= j.plane[0].h;
        p = 1;
        if (r >= j.plane[1].h) { r -= j.plane[1].h; p = 2; }
    }
    const ConvertPlane &P = j.plane[p];
    if (r >= P.h) return;
    const bool fast = P.fast != 0;
    if (P.mode == CONV_VFILT8 || P.mode == CONV_VFILT10) {
        const int x0 = xg * kConvertVGroup;
        if (x0 >= P.w) return;
        const int n = ott_min(kConvertVGroup, P.w - x0);
        int v[kConvertVGroup];
        if (P.mode == CONV_VFILT8) convert_vfilt<uint8_t>(j, P, x0, n, r, v);
        else convert_vfilt<uint16_t>(j, P, x0, n, r, v);
        convert_store<kConvertVGroup>(P.dst + (size_t)r * P.dst_pitch, x0, n, fast, v);
        return;
    }
    const int x0 = xg * kConvertGroup;
    if (x0 >= P.w) return;
    const int n = ott_min(kConvertGroup, P.w - x0);
    if (P.mode == CONV_COPY8 && fast) {                      // nothing to compute: move the vector as it is
        *reinterpret_cast<uint4 *>(P.dst + (size_t)r * P.dst_pitch + x0) =
            ott_ldg(reinterpret_cast<const uint4 *>(P.src + (size_t)r * P.src_pitch + x0));
        return;
    }
    int v[kConvertGroup];
    if (P.mode == CONV_COPY8) {
        convert_load<uint8_t, kConvertGroup>(P.src + (size_t)r * P.src_pitch, x0, n, fast, v);
    } else {
        convert_load<uint16_t, kConvertGroup>(P.src + (size_t)r * P.src_pitch, x0, n, fast, v);
        const int d0 = (r & 1) ? 3 : 1, d1 = (r & 1) ? 0 : 2;          // {{1,2},{3,0}}; x0 is even
        OTT_UNROLL
        for (int k = 0; k < kConvertGroup; ++k) v[k] = ott_min((v[k] + ((k & 1) ? d1 : d0)) >> 2, 255);
    }
    convert_store<kConvertGroup>(P.dst + (size_t)r * P.dst_pitch, x0, n, fast, v);
}

}  // namespace ottgpu
