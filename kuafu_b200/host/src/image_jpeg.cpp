// Baseline JPEG (ITU-T T.81 sequential DCT, Huffman, 8-bit) -> RGBA8, for textures referenced by imported
// assets (.mtl / .dae / .gltf): the reference reads them through stb_image (vkCore.hpp:1761-1795,
// stbi_load(..., STBI_rgb_alpha)).  Grey and YCbCr images, any sampling factors up to 2 x 2 per component
// (4:4:4, 4:2:2, 4:4:0, 4:2:0), restart intervals, JFIF / Adobe markers skipped.  Progressive, arithmetic
// coded, 12-bit and CMYK files are refused (loud failure upstream: "texture will not be used").
// The file is untrusted: every length and index is checked before use.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "image_io.hpp"

namespace kuafu::io {
namespace {

constexpr uint32_t kMaxSide = 32768;

struct Huffman {
  // canonical code tables: for code length l (1..16), codes [first[l], first[l] + count[l]) map to
  // symbols[offset[l] ...]
  uint16_t count[17] = {0};
  int32_t first[17] = {0}, offset[17] = {0};
  uint8_t symbols[256] = {0};
  bool present = false;
};

struct Component {
  int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
  int pred = 0;
  int blocksW = 0, blocksH = 0;        // blocks per line / column, padded to whole MCUs
  std::vector<uint8_t> plane;          // blocksW * 8 x blocksH * 8 samples
};

struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint32_t bits = 0;
  int n = 0;
  bool marker = false;  // ran into a marker (or the end): further bits read as zero
  void fill() {
    while (n <= 24) {
      uint32_t b = 0;
      if (!marker && p < end) {
        b = *p;
        if (b == 0xff) {
          if (p + 1 < end && p[1] == 0x00) p += 2;  // stuffed byte
          else { marker = true; b = 0; }
        } else {
          p++;
        }
      } else {
        marker = true;
      }
      bits |= b << (24 - n);
      n += 8;
    }
  }
  int get(int k) {  // k in 0..16
    if (k == 0) return 0;
    if (n < k) fill();
    const int v = int(bits >> (32 - k));
    bits <<= k;
    n -= k;
    return v;
  }
  void reset() { bits = 0; n = 0; marker = false; }
};

int decodeSymbol(BitReader& br, const Huffman& h) {
  int code = 0;
  for (int l = 1; l <= 16; l++) {
    code = (code << 1) | br.get(1);
    if (h.count[l] && code >= h.first[l] && code < h.first[l] + int(h.count[l])) return h.symbols[h.offset[l] + code - h.first[l]];
  }
  return -1;
}

int extend(int v, int t) { return (t && v < (1 << (t - 1))) ? v - (1 << t) + 1 : v; }

const uint8_t kZigZag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// Separable float inverse DCT of one 8 x 8 block (direct form with a cosine table: 1024 multiply-adds, fine for
// textures that are decoded once), level shift and clamp.
void idctBlock(const int* coef, uint8_t* out, int stride) {
  static float c[8][8];
  static bool init = false;
  if (!init) {
    for (int x = 0; x < 8; x++)
      for (int u = 0; u < 8; u++) c[x][u] = float((u == 0 ? std::sqrt(0.125) : 0.5) * std::cos((2 * x + 1) * u * 3.14159265358979323846 / 16.0));
    init = true;
  }
  float tmp[64];
  for (int y = 0; y < 8; y++)      // rows: over u
    for (int x = 0; x < 8; x++) {
      float s = 0.0f;
      for (int u = 0; u < 8; u++) s += c[x][u] * float(coef[8 * y + u]);
      tmp[8 * y + x] = s;
    }
  for (int x = 0; x < 8; x++)      // columns: over v
    for (int y = 0; y < 8; y++) {
      float s = 0.0f;
      for (int v = 0; v < 8; v++) s += c[y][v] * tmp[8 * v + x];
      const int val = int(std::floor(s + 128.5f));
      out[y * stride + x] = uint8_t(std::min(255, std::max(0, val)));
    }
}

uint16_t be16(const uint8_t* p) { return uint16_t((p[0] << 8) | p[1]); }

// Sample of a (possibly subsampled) plane at full-resolution pixel (x, y): bilinear between the sample centres,
// edges clamped ("fancy upsampling" of libjpeg / stb in spirit; exact for 1 x 1 sampling).
float planeAt(const Component& c, int hmax, int vmax, int x, int y) {
  const int W = c.blocksW * 8, H = c.blocksH * 8;
  if (c.h == hmax && c.v == vmax) return float(c.plane[size_t(y) * W + x]);
  const float fx = (x + 0.5f) * float(c.h) / float(hmax) - 0.5f, fy = (y + 0.5f) * float(c.v) / float(vmax) - 0.5f;
  const int x0 = int(std::floor(fx)), y0 = int(std::floor(fy));
  const float ax = fx - float(x0), ay = fy - float(y0);
  auto at = [&](int xx, int yy) {
    xx = std::min(W - 1, std::max(0, xx));
    yy = std::min(H - 1, std::max(0, yy));
    return float(c.plane[size_t(yy) * W + xx]);
  };
  return (at(x0, y0) * (1 - ax) + at(x0 + 1, y0) * ax) * (1 - ay) + (at(x0, y0 + 1) * (1 - ax) + at(x0 + 1, y0 + 1) * ax) * ay;
}

}  // namespace

bool decodeJpeg(const uint8_t* d, size_t n, uint32_t& W, uint32_t& H, std::vector<uint8_t>& rgba) {
  if (n < 4 || d[0] != 0xff || d[1] != 0xd8) return false;
  uint16_t quant[4][64];
  bool haveQuant[4] = {false, false, false, false};
  Huffman dc[4], ac[4];
  std::vector<Component> comps;
  int restartInterval = 0, hmax = 1, vmax = 1;
  bool haveFrame = false, adobeTransform = false;
  int adobeTransformValue = -1;
  size_t pos = 2;
  while (pos + 4 <= n) {
    if (d[pos] != 0xff) return false;
    while (pos < n && d[pos] == 0xff) pos++;  // fill bytes
    if (pos >= n) return false;
    const uint8_t m = d[pos++];
    if (m == 0xd8 || (m >= 0xd0 && m <= 0xd7) || m == 0x01) continue;  // no payload
    if (m == 0xd9) return false;                                       // EOI before any scan
    if (pos + 2 > n) return false;
    const size_t len = be16(d + pos);
    if (len < 2 || pos + len > n) return false;
    const uint8_t* b = d + pos + 2;
    const size_t bl = len - 2;
    if (m == 0xdb) {  // DQT
      size_t q = 0;
      while (q < bl) {
        const int pq = b[q] >> 4, tq = b[q] & 15;
        q++;
        if (tq > 3 || pq > 1 || q + size_t(64) * (pq + 1) > bl) return false;
        for (int k = 0; k < 64; k++) {
          quant[tq][kZigZag[k]] = pq ? be16(b + q + 2 * k) : b[q + k];
        }
        q += size_t(64) * (pq + 1);
        haveQuant[tq] = true;
      }
    } else if (m == 0xc4) {  // DHT
      size_t q = 0;
      while (q < bl) {
        if (q + 17 > bl) return false;
        const int tc = b[q] >> 4, th = b[q] & 15;
        if (tc > 1 || th > 3) return false;
        Huffman& h = tc ? ac[th] : dc[th];
        int total = 0;
        for (int l = 1; l <= 16; l++) {
          h.count[l] = b[q + l];
          total += h.count[l];
        }
        q += 17;
        if (total > 256 || q + size_t(total) > bl) return false;
        std::memcpy(h.symbols, b + q, size_t(total));
        q += size_t(total);
        int code = 0, off = 0;
        for (int l = 1; l <= 16; l++) {
          h.first[l] = code;
          h.offset[l] = off;
          code += h.count[l];
          off += h.count[l];
          if (code > (1 << l)) return false;  // over-subscribed
          code <<= 1;
        }
        h.present = true;
      }
    } else if (m == 0xc0 || m == 0xc1) {  // SOF0 / SOF1: sequential, Huffman
      if (haveFrame || bl < 6) return false;
      if (b[0] != 8) return false;  // 12-bit samples
      H = be16(b + 1);
      W = be16(b + 3);
      const int nc = b[5];
      if (W == 0 || H == 0 || W > kMaxSide || H > kMaxSide || uint64_t(W) * H > (1ull << 28) || (nc != 1 && nc != 3) || bl < size_t(6 + 3 * nc)) return false;
      comps.resize(size_t(nc));
      for (int i = 0; i < nc; i++) {
        Component& c = comps[size_t(i)];
        c.id = b[6 + 3 * i];
        c.h = b[7 + 3 * i] >> 4;
        c.v = b[7 + 3 * i] & 15;
        c.tq = b[8 + 3 * i];
        if (c.h < 1 || c.h > 2 || c.v < 1 || c.v > 2 || c.tq > 3) return false;
        hmax = std::max(hmax, c.h);
        vmax = std::max(vmax, c.v);
      }
      haveFrame = true;
    } else if (m == 0xc2 || (m >= 0xc3 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc)) {
      return false;  // progressive, lossless, differential or arithmetic coding
    } else if (m == 0xdd) {  // DRI
      if (bl < 2) return false;
      restartInterval = be16(b);
    } else if (m == 0xee) {  // Adobe: colour transform flag
      if (bl >= 12 && !std::memcmp(b, "Adobe", 5)) {
        adobeTransform = true;
        adobeTransformValue = b[11];
      }
    } else if (m == 0xda) {  // SOS: the one scan of a baseline file
      if (!haveFrame || bl < 1) return false;
      const int ns = b[0];
      if (ns != int(comps.size()) || bl < size_t(1 + 2 * ns + 3)) return false;  // non-interleaved multi-scan files are not handled
      for (int i = 0; i < ns; i++) {
        const int cid = b[1 + 2 * i];
        Component* c = nullptr;
        for (auto& k : comps)
          if (k.id == cid) c = &k;
        if (!c) return false;
        c->td = b[2 + 2 * i] >> 4;
        c->ta = b[2 + 2 * i] & 15;
        if (c->td > 3 || c->ta > 3 || !dc[c->td].present || !ac[c->ta].present || !haveQuant[c->tq]) return false;
      }
      const int mcuW = 8 * hmax, mcuH = 8 * vmax;
      const int mcusX = (int(W) + mcuW - 1) / mcuW, mcusY = (int(H) + mcuH - 1) / mcuH;
      for (auto& c : comps) {
        c.blocksW = mcusX * c.h;
        c.blocksH = mcusY * c.v;
        c.plane.assign(size_t(c.blocksW) * 8 * size_t(c.blocksH) * 8, 0);
        c.pred = 0;
      }
      BitReader br{d + pos + len, d + n};
      int untilRestart = restartInterval;
      int expectRst = 0;
      for (int my = 0; my < mcusY; my++)
        for (int mx = 0; mx < mcusX; mx++) {
          if (restartInterval && untilRestart == 0) {
            // byte-align, expect RSTn
            const uint8_t* q = br.p;
            while (q + 1 < br.end && !(q[0] == 0xff && q[1] >= 0xd0 && q[1] <= 0xd7)) q++;
            if (q + 1 >= br.end) return false;
            if ((q[1] & 7) != expectRst) return false;
            expectRst = (expectRst + 1) & 7;
            br.p = q + 2;
            br.reset();
            for (auto& c : comps) c.pred = 0;
            untilRestart = restartInterval;
          }
          for (auto& c : comps)
            for (int by = 0; by < c.v; by++)
              for (int bx = 0; bx < c.h; bx++) {
                int coef[64] = {0};
                const int t = decodeSymbol(br, dc[c.td]);
                if (t < 0 || t > 11) return false;
                c.pred += extend(br.get(t), t);
                coef[0] = c.pred * int(quant[c.tq][0]);
                for (int k = 1; k < 64;) {
                  const int rs = decodeSymbol(br, ac[c.ta]);
                  if (rs < 0) return false;
                  const int r = rs >> 4, s = rs & 15;
                  if (s == 0) {
                    if (r != 15) break;  // end of block
                    k += 16;
                    continue;
                  }
                  k += r;
                  if (k > 63) return false;
                  coef[kZigZag[k]] = extend(br.get(s), s) * int(quant[c.tq][kZigZag[k]]);
                  k++;
                }
                const int px = (mx * c.h + bx) * 8, py = (my * c.v + by) * 8;
                idctBlock(coef, &c.plane[size_t(py) * c.blocksW * 8 + px], c.blocksW * 8);
              }
          if (restartInterval) untilRestart--;
        }
      // colour conversion (JFIF: YCbCr full range; Adobe transform 0 with three components means RGB)
      rgba.resize(size_t(W) * H * 4);
      const bool isRgb = comps.size() == 3 && adobeTransform && adobeTransformValue == 0;
      for (uint32_t y = 0; y < H; y++)
        for (uint32_t x = 0; x < W; x++) {
          uint8_t* o = &rgba[(size_t(y) * W + x) * 4];
          if (comps.size() == 1) {
            o[0] = o[1] = o[2] = comps[0].plane[size_t(y) * comps[0].blocksW * 8 + x];
          } else {
            const float Y = planeAt(comps[0], hmax, vmax, int(x), int(y));
            const float Cb = planeAt(comps[1], hmax, vmax, int(x), int(y));
            const float Cr = planeAt(comps[2], hmax, vmax, int(x), int(y));
            float r, g, bl2;
            if (isRgb) {
              r = Y; g = Cb; bl2 = Cr;
            } else {
              r = Y + 1.402f * (Cr - 128.0f);
              g = Y - 0.344136f * (Cb - 128.0f) - 0.714136f * (Cr - 128.0f);
              bl2 = Y + 1.772f * (Cb - 128.0f);
            }
            o[0] = uint8_t(std::min(255.0f, std::max(0.0f, std::floor(r + 0.5f))));
            o[1] = uint8_t(std::min(255.0f, std::max(0.0f, std::floor(g + 0.5f))));
            o[2] = uint8_t(std::min(255.0f, std::max(0.0f, std::floor(bl2 + 0.5f))));
          }
          o[3] = 255;
        }
      return true;
    }
    pos += len;
  }
  return false;
}
}  // namespace kuafu::io
