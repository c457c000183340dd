// b2_upsample_core.cuh -- packed 2x16-bit arithmetic core of the fancy upsampler.
//
// Everything here works on 32-bit words holding TWO independent 16-bit lanes (two
// pixels, or the U and V sample of one pixel), so one IMAD / PRMT / VIADDMNMX does
// the work for two pixels.  It is bit-exact with the reference's scalar code:
//   interpolation  UPSAMPLE_FUNC, pkg/libwebp/dsp/upsampling.c.go:43-103
//   conversion     VP8YUVToR/G/B = MultHi + Clip8, pkg/color/yuv/conversion.go:28-48
// (tests/test_upsample_core_host.py runs this very header on the host over all
// 2^24 (y,u,v) triples and over the interpolation's input space.)
//
// Why this shape (profiles/r01c_pipebench_sm100a.txt): on sm_100a IMAD.HI issues at
// half rate and the first kernel spent ~19 ALU-pipe + ~14 FMA-pipe slots per pixel.
// Here no instruction is a multiply-high or a funnel shift: every ">> 8" is a byte
// permute (PRMT picks bytes 1 and 3 of a word = both lanes >> 8, masked), every
// other shift is folded into a multiplier, and Clip8 is one packed add-min-relu.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#else
#error "compile with nvcc (host pass is used by the exhaustive test)"
#endif

// PRMT: result byte i = byte (sel >> 4i & 7) of the 8-byte pool {b:a}.
B2_HD uint32_t b2_prmt(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef __CUDA_ARCH__
  return __byte_perm(a, b, sel);
#else
  const uint64_t pool = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= (uint32_t)((pool >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
#endif
}
// both lanes >> 8, masked to 8 bits: [w.b1, 0, w.b3, 0]
B2_HD uint32_t b2_hi8(uint32_t w) { return b2_prmt(w, 0u, 0x4341); }
// even / odd bytes of a word as two clean lanes
B2_HD uint32_t b2_even(uint32_t w) { return w & 0x00ff00ffu; }
B2_HD uint32_t b2_odd(uint32_t w) { return b2_hi8(w); }
#define B2_K2(x) ((uint32_t)((x) & 0xffff) * 0x10001u)

// ---- chroma interpolation --------------------------------------------------------
// For the four samples tl,t (upper chroma row) and l,c (lower chroma row) of two
// adjacent columns, the reference computes (upsampling.c.go:69-86)
//   avg = tl+t+l+c+8, d12 = (avg + 2(t+l)) >> 3, d03 = (avg + 2(tl+c)) >> 3,
//   top row: (d12+tl)>>1, (d03+t)>>1;  bottom row: (d03+l)>>1, (d12+c)>>1.
// floor((floor(a/8)+b)/2) == floor((a+8b)/16), so each output is bits 4..11 of
// x + 8*sample; carrying everything scaled by 16 puts it in byte 1 of the lane
// (max 16*2048 + 128*255 = 65408 < 2^16), where b2_hi8() picks it up.
struct B2Quad16 { uint32_t tf, ts, bf, bs; };  // value = byte 1 of each 16-bit lane
B2_HD B2Quad16 b2_interp16(uint32_t tl, uint32_t t, uint32_t l, uint32_t c) {
  const uint32_t X = tl + c, Z = t + l;
  const uint32_t avg16 = (X + Z) * 16u + 0x00800080u;
  const uint32_t x12 = Z * 32u + avg16, x03 = X * 32u + avg16;
  B2Quad16 q;
  q.tf = tl * 128u + x12;  // top row, pixel 2x-1
  q.ts = t * 128u + x03;   // top row, pixel 2x
  q.bf = l * 128u + x03;   // bottom row, pixel 2x-1
  q.bs = c * 128u + x12;   // bottom row, pixel 2x
  return q;
}

// ---- YUV -> RGB -------------------------------------------------------------------
// MultHi(v, k) = (v*k) >> 8 with k = 256*kh + kl is v*kh + ((v*kl) >> 8); for 8-bit
// v both products fit a 16-bit lane: 19077 = 74*256+133, 26149 = 102*256+37,
// 33050 = 129*256+26, 6419 = 25*256+19, 13320 = 52*256+8.
//
// VIADDMNMX computes relu(min(a + b, c)) per 16-bit lane with a wrapping add and only
// c may be an immediate, so the constants of the three channels are arranged to need
// as few register operands as possible:
//  * luma carries a free offset: the PRMT that extracts (y*133) >> 8 takes its zero
//    bytes from `k.yoff` instead (byte 0x23), i.e. yh2 = MultHi(y,19077) + 0x2300;
//  * G = a + b with a = yh2 - MultHi(u,6419) (>= 0 thanks to the offset) and
//    b = 65284 - 52 v - (v >> 5); 0x2300 + 65284 = 8708 + 65536, so the wrapped sum
//    is exactly the reference's G before clipping -- no constant operand at all;
//  * R and B subtract (14234 + 0x2300) and (17685 + 0x2300) through `k.kr`, `k.kb`.
struct B2YuvConsts { uint32_t yoff, kr, kb; };
B2_HD B2YuvConsts b2_yuv_consts() {
  B2YuvConsts k;
  k.yoff = 0x23u;
  k.kr = B2_K2(-14234 - 0x2300);
  k.kb = B2_K2(-17685 - 0x2300);
  return k;
}
// MultHi(y,19077) + 0x2300 in both lanes
B2_HD uint32_t b2_luma14(uint32_t y2, const B2YuvConsts& k) {
  return y2 * 74u + b2_prmt(y2 * 133u, k.yoff, 0x4341);
}

struct B2Rgb14 { uint32_t r, g, b; };  // lanes = pre-shift channel clamped to [0,16383]
B2_HD B2Rgb14 b2_yuv_pair(uint32_t yh2, uint32_t u2, uint32_t v2, const B2YuvConsts& k) {
  B2Rgb14 o;
  // R = Clip8(MultHi(y,19077) + MultHi(v,26149) - 14234); lanes of wr <= 54009
  const uint32_t wr = v2 * 102u + (yh2 + b2_hi8(v2 * 37u));
  o.r = __viaddmin_s16x2_relu(wr, k.kr, B2_K2(16383));
  // B = Clip8(MultHi(y,19077) + MultHi(u,33050) - 17685); the sum reaches 60882 and
  // sum - 26645 would leave int16, so the add saturates the top first (unsigned min)
  const uint32_t bu = u2 * 129u + b2_hi8(u2 * 26u);
  o.b = __viaddmin_s16x2_relu(__viaddmin_u16x2(bu, yh2, B2_K2(17685 + 0x2300 + 16383)), k.kb, B2_K2(16383));
  // G = Clip8(MultHi(y,19077) - MultHi(u,6419) - MultHi(v,13320) + 8708)
  const uint32_t ga = (yh2 - b2_hi8(u2 * 19u)) - u2 * 25u;              // lanes in [2560, 27962]
  const uint32_t gb = B2_K2(65284) - b2_hi8(v2 * 8u) - v2 * 52u;        // lanes in [52017, 65284]
  o.g = __viaddmin_s16x2_relu(ga, gb, B2_K2(16383));
  return o;
}

// Clip8's ">> 6": x4 puts the byte in byte 1 / byte 3 of the word.
B2_HD uint32_t b2_to_byte1(uint32_t c14) { return c14 * 4u; }
