// lv_pixel.cuh -- per-pixel arithmetic of the hot path, written for sm_100a integer SIMD.
//
// Colour conversion: bit-exact with ColorSpaceConversions::YV12_to_RGB24 of the reference
// (Convert.h:26; code in cscc.lib, Convert.obj .text+0x350, tables built at .text+0x110..0x1c1):
//     t  = (76309*(Y-16))  >> 15                       tab_76309
//     cb = (132201*(U-128)) >> 15                      cbu_tab
//     cr = (104597*(V-128)) >> 15                      crv_tab
//     g  = (25675*(U-128) + 53279*(V-128)) >> 15       cgu_tab + cgv_tab, ONE shift of the sum
//     B = clp(t + cb), G = clp(t - g), R = clp(t + cr),  clp(x) = x<0 ? 0 : x>510 ? 255 : x>>1
// (all shifts arithmetic).  No tables on the GPU: every term is one IMAD whose multiplier is
// doubled so the wanted value is the upper halfword (x>>15 == (2x)>>16), two pixels are packed per
// register as s16x2 with PRMT, and add+clamp is one VIADDMNMX.S16x2.RELU (DPX) per channel per
// pixel pair:  clp(x) == min(max(x,0),511) >> 1.
//   -g is formed directly: -floor(s/2^15) == floor((-s + 32767)/2^15).
//
// Difference: IsNearlyZero (lib/JM/image.h:22) => changed <=> |cur - ref| > tol, on 4 pixels per
// 32-bit lane with VABSDIFF4 and a carry trick for the per-byte compare.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lv {

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    return __byte_perm(a, b, sel);
}

// Two instruction-mix switches, chosen by measurement (tools/ubench/pipes.cu shows PRMT/SHF/LOP3/VIADDMNMX on
// the ALU pipe and IMAD/IDP on the FMA pipe, 2 warp-instructions/clk/SM each; profiles/ has the A/B runs):
//   LV_EXTRACT_IDP : byte extraction with IDP.4A (FMA pipe) instead of PRMT (ALU pipe)
//   LV_CHAN_HI     : channel value moved to the high byte of each halfword with *128 (IMAD.SHL, FMA pipe)
//                    instead of >>1 into the low byte (SHF, ALU pipe)
#ifndef LV_EXTRACT_IDP
#define LV_EXTRACT_IDP 0
#endif
#ifndef LV_CHAN_HI
#define LV_CHAN_HI 0
#endif
#ifndef LV_DIFF_POPC
#define LV_DIFF_POPC 0
#endif

// byte k (0..3) of w, zero extended
template <int K>
__device__ __forceinline__ int ubyte(uint32_t w)
{
#if LV_EXTRACT_IDP
    return (int)__dp4a(w, 1u << (8 * K), 0u);
#else
    return (int)prmt(w, 0u, 0x4440u | (uint32_t)K);
#endif
}

// Chroma terms of one (U,V) sample, each duplicated into both halfwords (the two pixels of a
// horizontal pair share the sample; so do the two rows).
struct Chroma2 {
    uint32_t cb2, ng2, cr2;
};

__device__ __forceinline__ Chroma2 chroma_terms(int u, int v)
{
    // doubled multipliers: the result of interest is the upper halfword
    const int CB = u * (2 * 132201) - (2 * 132201) * 128;
    const int CR = v * (2 * 104597) - (2 * 104597) * 128;
    const int NG = u * (-2 * 25675) + (v * (-2 * 53279) + ((2 * 25675 + 2 * 53279) * 128 + 2 * 32767));
    Chroma2 c;
    c.cb2 = prmt((uint32_t)CB, 0u, 0x3232u);
    c.cr2 = prmt((uint32_t)CR, 0u, 0x3232u);
    c.ng2 = prmt((uint32_t)NG, 0u, 0x3232u);
    return c;
}

// luma terms of two horizontally adjacent pixels packed as s16x2 (low half = even pixel)
__device__ __forceinline__ uint32_t luma_pair(int ya, int yb)
{
    const int TA = ya * (2 * 76309) - (2 * 76309) * 16;
    const int TB = yb * (2 * 76309) - (2 * 76309) * 16;
    return prmt((uint32_t)TA, (uint32_t)TB, 0x7632u);
}

// add + clamp to [0,511] per halfword (one VIADDMNMX.S16x2.RELU), then drop the low bit:
//   LV_CHAN_HI = 0: >>1, the channel value is the LOW byte of each halfword (bytes 0 and 2; bit 15 may carry
//                   the upper halfword's dropped bit, it is never selected)
//   LV_CHAN_HI = 1: *128, the channel value is the HIGH byte of each halfword (bytes 1 and 3)
__device__ __forceinline__ uint32_t chan_pair(uint32_t t2, uint32_t c2)
{
#if LV_CHAN_HI
    return __viaddmin_s16x2_relu(t2, c2, 0x01ff01ffu) * 128u;
#else
    return __viaddmin_s16x2_relu(t2, c2, 0x01ff01ffu) >> 1;
#endif
}
// the two channel bytes of a chan_pair() result (even pixel, odd pixel)
__device__ __forceinline__ uint8_t chan_even(uint32_t c) { return (uint8_t)(c >> (LV_CHAN_HI ? 8 : 0)); }
__device__ __forceinline__ uint8_t chan_odd(uint32_t c) { return (uint8_t)(c >> (LV_CHAN_HI ? 24 : 16)); }

// 4 pixels (one 32-bit word of luma, two chroma samples) -> 12 bytes B,G,R,B,G,R,...
__device__ __forceinline__ void convert4(uint32_t yw, const Chroma2 &c0, const Chroma2 &c1,
                                         uint32_t &o0, uint32_t &o1, uint32_t &o2)
{
    const uint32_t ta = luma_pair(ubyte<0>(yw), ubyte<1>(yw));
    const uint32_t tb = luma_pair(ubyte<2>(yw), ubyte<3>(yw));
    const uint32_t Ba = chan_pair(ta, c0.cb2), Ga = chan_pair(ta, c0.ng2), Ra = chan_pair(ta, c0.cr2);
    const uint32_t Bb = chan_pair(tb, c1.cb2), Gb = chan_pair(tb, c1.ng2), Rb = chan_pair(tb, c1.cr2);
#if LV_CHAN_HI
    // pixel p of pair x sits in byte 1 (even) / byte 3 (odd) of Bx,Gx,Rx
    o0 = prmt(prmt(Ba, Ga, 0x3051u), Ra, 0x3510u);                       // b0 g0 r0 b1
    o1 = prmt(prmt(prmt(Ga, Ra, 0x0073u), Bb, 0x0510u), Gb, 0x5210u);    // g1 r1 b2 g2
    o2 = prmt(prmt(Rb, Bb, 0x3071u), Gb, 0x3710u);                       // r2 b3 g3 r3
#else
    // pixel p of pair x sits in byte 0 (even) / byte 2 (odd) of Bx,Gx,Rx
    o0 = prmt(prmt(Ba, Ga, 0x2040u), Ra, 0x3410u);                       // b0 g0 r0 b1
    o1 = prmt(prmt(prmt(Ga, Ra, 0x0062u), Bb, 0x0410u), Gb, 0x4210u);    // g1 r1 b2 g2
    o2 = prmt(prmt(Rb, Bb, 0x2060u), Gb, 0x3610u);                       // r2 b3 g3 r3
#endif
}

// per-byte (|a-b| > tol) for 4 pixels: returns a word whose bit 7 of byte k is the result for
// pixel k (other bits undefined).  ntol7 = (~tol4) & 0x7f7f7f7f, tol4 = tol * 0x01010101.
__device__ __forceinline__ uint32_t diff_gt4(uint32_t a, uint32_t b, uint32_t tol4, uint32_t ntol7)
{
    const uint32_t d = __vabsdiffu4(a, b);
    // d > tol per byte without cross-byte carries: add the low 7 bits, then resolve bit 7 with a
    // majority-style LOP3:  gt = (d & ~tol) | (~(d ^ tol) & carry)   evaluated on bit 7
    const uint32_t s = (d & 0x7f7f7f7fu) + ntol7;
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xb2;" : "=r"(r) : "r"(d), "r"(tol4), "r"(s));
    return r;
}

// 16 pixels -> 16-bit mask, bit i = pixel i changed.  The gather of the four compare bits of a word is one IDP.4A
// against the byte weights 1,2,4,8 (second word: 16..128), exactly the instruction the counting path spends per word
// (diff_count16), so the mask costs what the count costs: 128 * byte per word pair, two pairs joined by one IMAD + SHF.
__device__ __forceinline__ uint32_t diff_mask16(const uint4 &cur, const uint4 &ref, uint32_t tol4,
                                                uint32_t ntol7)
{
    uint32_t lo = __dp4a(diff_gt4(cur.x, ref.x, tol4, ntol7) & 0x80808080u, 0x08040201u, 0u);
    lo = __dp4a(diff_gt4(cur.y, ref.y, tol4, ntol7) & 0x80808080u, 0x80402010u, lo);
    uint32_t hi = __dp4a(diff_gt4(cur.z, ref.z, tol4, ntol7) & 0x80808080u, 0x08040201u, 0u);
    hi = __dp4a(diff_gt4(cur.w, ref.w, tol4, ntol7) & 0x80808080u, 0x80402010u, hi);
    return (hi * 256u + lo) >> 7;
}

// 16 pixels -> acc + 128 * (number of changed pixels): bit 7 of each byte summed by one IDP.4A per word
// (the dot product runs on the FMA pipe, off the ALU pipe that bounds this kernel)
__device__ __forceinline__ uint32_t diff_count16(const uint4 &cur, const uint4 &ref, uint32_t tol4, uint32_t ntol7,
                                                 uint32_t acc)
{
#if LV_DIFF_POPC
    // all-ALU variant: OR the four compare words into distinct bit positions, one POPC
    const uint32_t r0 = diff_gt4(cur.x, ref.x, tol4, ntol7) & 0x80808080u;
    const uint32_t r1 = diff_gt4(cur.y, ref.y, tol4, ntol7) & 0x80808080u;
    const uint32_t r2 = diff_gt4(cur.z, ref.z, tol4, ntol7) & 0x80808080u;
    const uint32_t r3 = diff_gt4(cur.w, ref.w, tol4, ntol7) & 0x80808080u;
    return acc + ((uint32_t)__popc(r0 | (r1 >> 1) | (r2 >> 2) | (r3 >> 3)) << 7);
#endif
    acc = __dp4a(diff_gt4(cur.x, ref.x, tol4, ntol7) & 0x80808080u, 0x01010101u, acc);
    acc = __dp4a(diff_gt4(cur.y, ref.y, tol4, ntol7) & 0x80808080u, 0x01010101u, acc);
    acc = __dp4a(diff_gt4(cur.z, ref.z, tol4, ntol7) & 0x80808080u, 0x01010101u, acc);
    acc = __dp4a(diff_gt4(cur.w, ref.w, tol4, ntol7) & 0x80808080u, 0x01010101u, acc);
    return acc;
}

// expand 4 mask bits (bits 4k..4k+3 of m) to 4 bytes of 0/1
__device__ __forceinline__ uint32_t mask_bytes4(uint32_t m, int k)
{
    const uint32_t n = (m >> (4 * k)) & 0xfu;
    return (n * 0x00204081u) & 0x01010101u;
}

}  // namespace lv
