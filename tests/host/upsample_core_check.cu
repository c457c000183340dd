// Host-side exhaustive check of go_webp_b200/csrc/b2_upsample_core.cuh (the packed
// arithmetic the CUDA fancy upsampler is built from) against the reference's scalar
// formulas.  Built and run by tests/test_upsample_core_host.py (no GPU needed).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include "../../go_webp_b200/csrc/b2_upsample_core.cuh"

// pkg/color/yuv/conversion.go:28-48
static inline int MultHi(int v, int c) { return (v * c) >> 8; }
static inline int Clip8(int v) { return ((v & ~16383) == 0) ? (v >> 6) : (v < 0) ? 0 : 255; }
static inline int RefR(int y, int v) { return Clip8(MultHi(y, 19077) + MultHi(v, 26149) - 14234); }
static inline int RefG(int y, int u, int v) { return Clip8(MultHi(y, 19077) - MultHi(u, 6419) - MultHi(v, 13320) + 8708); }
static inline int RefB(int y, int u) { return Clip8(MultHi(y, 19077) + MultHi(u, 33050) - 17685); }

int main() {
  long bad = 0;
  // 1. every (y,u,v); the other lane carries a different triple to prove lane independence
  for (int y = 0; y < 256; ++y)
    for (int u = 0; u < 256; ++u)
      for (int v = 0; v < 256; ++v) {
        const int y1 = 255 - y, u1 = (v * 7 + 3) & 255, v1 = (u * 13 + 5) & 255;
        const uint32_t y2 = (uint32_t)y | ((uint32_t)y1 << 16), u2 = (uint32_t)u | ((uint32_t)u1 << 16),
                       v2 = (uint32_t)v | ((uint32_t)v1 << 16);
        const B2YuvConsts kc = b2_yuv_consts();
        const B2Rgb14 o = b2_yuv_pair(b2_luma14(y2, kc), u2, v2, kc);
        const uint32_t r = b2_hi8(b2_to_byte1(o.r)), g = b2_hi8(b2_to_byte1(o.g)), b = b2_hi8(b2_to_byte1(o.b));
        const uint32_t er = (uint32_t)RefR(y, v) | ((uint32_t)RefR(y1, v1) << 16);
        const uint32_t eg = (uint32_t)RefG(y, u, v) | ((uint32_t)RefG(y1, u1, v1) << 16);
        const uint32_t eb = (uint32_t)RefB(y, u) | ((uint32_t)RefB(y1, u1) << 16);
        if (r != er || g != eg || b != eb) {
          if (bad < 5) printf("yuv mismatch y=%d u=%d v=%d: got %08x %08x %08x want %08x %08x %08x\n", y, u, v, r, g, b, er, eg, eb);
          ++bad;
        }
      }
  // 2. interpolation on a 64^4 lattice that holds both ends and the middle of the range
  int vals[64];
  for (int i = 0; i < 16; ++i) { vals[i] = i; vals[16 + i] = 120 + i; vals[32 + i] = 240 + i; vals[48 + i] = (i * 37 + 19) & 255; }
  for (int a = 0; a < 64; ++a)
    for (int b = 0; b < 64; ++b)
      for (int c = 0; c < 64; ++c)
        for (int d = 0; d < 64; ++d) {
          const int tl = vals[a], t = vals[b], l = vals[c], cc = vals[d];
          // second lane: the mirrored quad
          const uint32_t TL = tl | ((uint32_t)cc << 16), T = t | ((uint32_t)l << 16), L = l | ((uint32_t)t << 16), C = cc | ((uint32_t)tl << 16);
          const B2Quad16 q = b2_interp16(TL, T, L, C);
          auto ref = [](int tl_, int t_, int l_, int c_, int* o) {  // upsampling.c.go:69-86
            const int avg = tl_ + t_ + l_ + c_ + 8;
            const int d12 = (avg + 2 * (t_ + l_)) >> 3, d03 = (avg + 2 * (tl_ + c_)) >> 3;
            o[0] = (d12 + tl_) >> 1; o[1] = (d03 + t_) >> 1; o[2] = (d03 + l_) >> 1; o[3] = (d12 + c_) >> 1;
          };
          int e0[4], e1[4];
          ref(tl, t, l, cc, e0);
          ref(cc, l, t, tl, e1);
          const uint32_t got[4] = {b2_hi8(q.tf), b2_hi8(q.ts), b2_hi8(q.bf), b2_hi8(q.bs)};
          for (int k = 0; k < 4; ++k)
            if (got[k] != ((uint32_t)e0[k] | ((uint32_t)e1[k] << 16))) {
              if (bad < 10) printf("interp mismatch %d %d %d %d out %d: got %08x\n", tl, t, l, cc, k, got[k]);
              ++bad;
            }
        }
  printf("%s (%ld mismatches)\n", bad ? "FAIL" : "OK", bad);
  return bad ? 1 : 0;
}
