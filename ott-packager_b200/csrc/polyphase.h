// Polyphase coefficient tables for the ladder scaler (host side).
//
// Replaces what sws_getContext() builds for the reference (/root/reference/source/transvideo.c:1563-1565, 2155-2157):
// one horizontal (14-bit) and one vertical (12-bit) table per plane per rung.  The numbers must equal libswscale's
// (legacy scaler, utils.c initFilter) exactly, because the kernels are bit-exact to sws_scale(); SURVEY.md Appendix A.1
// is the specification.  This is product code: it shares nothing with oracle/ (tests compare the two).
#pragma once
#include <cstdint>
#include <vector>

namespace ottgpu {

enum class Kernel { Bicubic = 0, Lanczos = 1 };

struct PolyphaseTable {
    int taps = 0;                 // taps per output sample (after trimming + alignment padding)
    int n = 0;                    // output samples
    std::vector<int32_t> pos;     // [n] first source sample of each window; pos[i] + taps <= src_n always holds
    std::vector<int16_t> coef;    // [n * taps]; every row sums to `one`
};

// src_n -> dst_n samples, centre-aligned sampling (swscale's default chroma/luma positions all resolve to 128).
//   align      : libswscale filterAlign (x86 default flags: 4 horizontal, 2 vertical; bit-exact/C: 1)
//   one        : 1<<14 horizontal, 1<<12 vertical
//   zero_tail  : bit-exact flag behaviour (taps inside the alignment padding forced to zero)
PolyphaseTable make_polyphase(int src_n, int dst_n, Kernel k, int align, int one, bool zero_tail);

}  // namespace ottgpu
