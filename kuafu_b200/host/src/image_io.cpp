#include "image_io.hpp"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "core/context/global.hpp"

namespace kuafu::io {
namespace {
bool readFile(const std::string& path, std::vector<uint8_t>& out) {
  std::ifstream in(path, std::ios::binary);
  if (!in.good()) return false;
  in.seekg(0, std::ios::end);
  const std::streamoff n = in.tellg();
  in.seekg(0, std::ios::beg);
  out.resize(size_t(n));
  in.read(reinterpret_cast<char*>(out.data()), n);
  return in.good() || in.eof();
}
uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
uint32_t le32(const uint8_t* p) { return (uint32_t(p[3]) << 24) | (uint32_t(p[2]) << 16) | (uint32_t(p[1]) << 8) | p[0]; }
int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace

// PNG (all colour types, bit depths 1 / 2 / 4 / 8 / 16, Adam7 interlace) -> RGBA8 the way the reference's
// stb loader hands it over (STBI_rgb_alpha, 8 bits per channel: 16-bit samples keep their high byte, low
// bit depths of grey images are scaled to 0..255, reference vkCore.hpp:1761-1795).  Chunk contents are
// untrusted: sizes are checked before anything is allocated or indexed.
namespace {
constexpr uint32_t kMaxImageSide = 32768;

// One pass (or the whole non-interlaced image): un-filters `h` scanlines of `w` pixels starting at
// raw[pos] and writes 8-bit samples (`ch` per pixel) to out(x, y).  Returns false on truncated data.
template <class Put>
bool unfilterPass(const std::vector<uint8_t>& raw, size_t& pos, uint32_t w, uint32_t h, int ch, int bitDepth, Put put) {
  if (w == 0 || h == 0) return true;
  const size_t bitsPerPixel = size_t(ch) * bitDepth;
  const size_t stride = (size_t(w) * bitsPerPixel + 7) / 8;
  const size_t bpp = std::max<size_t>(1, bitsPerPixel / 8);  // filter distance in bytes
  if (pos + (stride + 1) * size_t(h) > raw.size()) return false;
  std::vector<uint8_t> prev(stride, 0), cur(stride, 0);
  for (uint32_t y = 0; y < h; y++) {
    const uint8_t ft = raw[pos++];
    const uint8_t* src = &raw[pos];
    pos += stride;
    for (size_t x = 0; x < stride; x++) {
      const int a = x >= bpp ? cur[x - bpp] : 0;
      const int b = prev[x];
      const int c = x >= bpp ? prev[x - bpp] : 0;
      int v = src[x];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: return false;
      }
      cur[x] = uint8_t(v);
    }
    for (uint32_t x = 0; x < w; x++) {
      uint8_t sample[4] = {0, 0, 0, 0};
      for (int k = 0; k < ch; k++) {
        if (bitDepth == 8) sample[k] = cur[size_t(x) * ch + k];
        else if (bitDepth == 16) sample[k] = cur[(size_t(x) * ch + k) * 2];  // high byte
        else {
          const size_t bit = size_t(x) * bitDepth;  // ch == 1 below 8 bits
          sample[k] = uint8_t((cur[bit / 8] >> (8 - bitDepth - bit % 8)) & ((1u << bitDepth) - 1u));
        }
      }
      put(x, y, sample);
    }
    prev.swap(cur);
  }
  return true;
}
}  // namespace

bool decodePng(const uint8_t* d, size_t n, uint32_t& W, uint32_t& H, std::vector<uint8_t>& rgba) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (n < 8 || std::memcmp(d, sig, 8) != 0) return false;
  size_t pos = 8;
  int colorType = -1, bitDepth = 0, interlace = 0;
  bool haveHeader = false;
  std::vector<uint8_t> idat, palette, trns;
  while (pos + 12 <= n) {
    const uint32_t len = be32(d + pos);
    const uint8_t* type = d + pos + 4;
    const uint8_t* body = d + pos + 8;
    if (size_t(len) > n - pos - 12) return false;
    if (!std::memcmp(type, "IHDR", 4)) {
      if (len != 13 || haveHeader) return false;
      W = be32(body);
      H = be32(body + 4);
      bitDepth = body[8];
      colorType = body[9];
      if (body[10] != 0 || body[11] != 0) return false;  // compression / filter method
      interlace = body[12];
      haveHeader = true;
    } else if (!std::memcmp(type, "PLTE", 4)) {
      palette.assign(body, body + len);
    } else if (!std::memcmp(type, "tRNS", 4)) {
      trns.assign(body, body + len);
    } else if (!std::memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), body, body + len);
    } else if (!std::memcmp(type, "IEND", 4)) {
      break;
    }
    pos += 12 + size_t(len);
  }
  if (!haveHeader || W == 0 || H == 0 || W > kMaxImageSide || H > kMaxImageSide || uint64_t(W) * H > (1ull << 28) || interlace > 1)
    return false;
  int ch;
  switch (colorType) {
    case 0: ch = 1; break;
    case 2: ch = 3; break;
    case 3: ch = 1; break;
    case 4: ch = 2; break;
    case 6: ch = 4; break;
    default: return false;
  }
  const bool depthOk = (colorType == 0 && (bitDepth == 1 || bitDepth == 2 || bitDepth == 4 || bitDepth == 8 || bitDepth == 16)) ||
                       (colorType == 3 && (bitDepth == 1 || bitDepth == 2 || bitDepth == 4 || bitDepth == 8)) ||
                       ((colorType == 2 || colorType == 4 || colorType == 6) && (bitDepth == 8 || bitDepth == 16));
  if (!depthOk) return false;
  // size of the un-filtered stream: one pass, or the seven Adam7 passes
  static const int px0[7] = {0, 4, 0, 2, 0, 1, 0}, py0[7] = {0, 0, 4, 0, 2, 0, 1};
  static const int pdx[7] = {8, 8, 4, 4, 2, 2, 1}, pdy[7] = {8, 8, 8, 4, 4, 2, 2};
  const size_t bitsPerPixel = size_t(ch) * bitDepth;
  size_t rawSize = 0;
  if (!interlace) {
    rawSize = ((size_t(W) * bitsPerPixel + 7) / 8 + 1) * H;
  } else {
    for (int p = 0; p < 7; p++) {
      const uint32_t pw = (W + pdx[p] - 1 - px0[p]) / pdx[p], ph = (H + pdy[p] - 1 - py0[p]) / pdy[p];
      if (pw && ph) rawSize += ((size_t(pw) * bitsPerPixel + 7) / 8 + 1) * ph;
    }
  }
  std::vector<uint8_t> raw(rawSize);
  uLongf rawLen = uLongf(raw.size());
  if (idat.empty() || uncompress(raw.data(), &rawLen, idat.data(), uLong(idat.size())) != Z_OK || rawLen != raw.size())
    return false;
  rgba.assign(size_t(W) * H * 4, 255);
  bool bad = false;
  const int greyScale = bitDepth < 8 ? 255 / ((1 << bitDepth) - 1) : 1;
  auto store = [&](uint32_t x, uint32_t y, const uint8_t* s) {
    uint8_t* o = &rgba[(size_t(y) * W + x) * 4];
    if (colorType == 0) {
      const uint8_t g = uint8_t(s[0] * greyScale);
      o[0] = o[1] = o[2] = g;
    } else if (colorType == 2) {
      o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
    } else if (colorType == 3) {
      const size_t k = s[0];
      if (3 * k + 2 >= palette.size()) { bad = true; return; }
      o[0] = palette[3 * k]; o[1] = palette[3 * k + 1]; o[2] = palette[3 * k + 2];
      if (k < trns.size()) o[3] = trns[k];
    } else if (colorType == 4) {
      o[0] = o[1] = o[2] = s[0]; o[3] = s[1];
    } else {
      o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[3];
    }
  };
  size_t rp = 0;
  if (!interlace) {
    if (!unfilterPass(raw, rp, W, H, ch, bitDepth, store)) return false;
  } else {
    for (int p = 0; p < 7; p++) {
      const uint32_t pw = (W + pdx[p] - 1 - px0[p]) / pdx[p], ph = (H + pdy[p] - 1 - py0[p]) / pdy[p];
      if (!unfilterPass(raw, rp, pw, ph, ch, bitDepth, [&](uint32_t x, uint32_t y, const uint8_t* s) {
            store(px0[p] + x * pdx[p], py0[p] + y * pdy[p], s);
          }))
        return false;
    }
  }
  return !bad;
}

static bool decodePnm(const std::vector<uint8_t>& f, uint32_t& W, uint32_t& H, std::vector<uint8_t>& rgba) {
  if (f.size() < 3 || f[0] != 'P' || (f[1] != '6' && f[1] != '5')) return false;
  const int ch = f[1] == '6' ? 3 : 1;
  size_t pos = 2;
  long vals[3];
  for (int k = 0; k < 3; k++) {
    for (;;) {
      while (pos < f.size() && std::isspace(f[pos])) pos++;
      if (pos < f.size() && f[pos] == '#') { while (pos < f.size() && f[pos] != '\n') pos++; } else break;
    }
    long v = 0;
    bool any = false;
    while (pos < f.size() && std::isdigit(f[pos])) { v = v * 10 + (f[pos++] - '0'); any = true; }
    if (!any) return false;
    vals[k] = v;
  }
  pos++;  // single whitespace after maxval
  if (vals[0] <= 0 || vals[1] <= 0 || vals[0] > long(kMaxImageSide) || vals[1] > long(kMaxImageSide)) return false;
  W = uint32_t(vals[0]);
  H = uint32_t(vals[1]);
  if (vals[2] != 255 || pos + size_t(W) * H * ch > f.size()) return false;
  rgba.resize(size_t(W) * H * 4);
  for (size_t i = 0; i < size_t(W) * H; i++) {
    const uint8_t* p = &f[pos + i * ch];
    rgba[4 * i] = p[0];
    rgba[4 * i + 1] = ch == 3 ? p[1] : p[0];
    rgba[4 * i + 2] = ch == 3 ? p[2] : p[0];
    rgba[4 * i + 3] = 255;
  }
  return true;
}

bool loadTextureRGBA8(const std::string& path, uint32_t& W, uint32_t& H, std::vector<uint8_t>& rgba) {
  if (path.rfind("mem:", 0) == 0) {
    const global::MemoryTexture* t = global::findMemoryTexture(path);
    if (!t) return false;
    W = t->width;
    H = t->height;
    rgba = t->rgba;
    return true;
  }
  std::vector<uint8_t> f;
  if (!readFile(path, f) && !(path[0] != '/' && readFile(global::assetsPath + path, f))) return false;
  if (decodePng(f.data(), f.size(), W, H, rgba)) return true;
  if (decodeJpeg(f.data(), f.size(), W, H, rgba)) return true;
  return decodePnm(f, W, H, rgba);
}

bool loadKtxCubeRGBA8(const std::string& path, uint32_t& size, std::vector<uint8_t> faces[6]) {
  std::vector<uint8_t> f;
  if (!readFile(path, f) && !(path[0] != '/' && readFile(global::assetsPath + path, f))) return false;
  static const uint8_t id[12] = {0xAB, 'K', 'T', 'X', ' ', '1', '1', 0xBB, '\r', '\n', 0x1A, '\n'};
  if (f.size() < 64 || std::memcmp(f.data(), id, 12) != 0) return false;
  const uint8_t* h = f.data() + 12;
  if (le32(h) != 0x04030201u) return false;
  const uint32_t glType = le32(h + 4), glFormat = le32(h + 12), width = le32(h + 24), height = le32(h + 28);
  const uint32_t nFaces = le32(h + 40), kvBytes = le32(h + 48);
  if (glType != 0x1401 /*UNSIGNED_BYTE*/ || glFormat != 0x1908 /*RGBA*/ || nFaces != 6 || width != height || !width ||
      width > kMaxImageSide)
    return false;
  size_t pos = 64 + size_t(kvBytes);
  if (pos + 4 > f.size()) return false;
  const uint32_t faceBytes = le32(f.data() + pos);
  pos += 4;
  if (faceBytes != width * height * 4) return false;
  for (int k = 0; k < 6; k++) {
    if (pos + faceBytes > f.size()) return false;
    faces[k].assign(f.begin() + pos, f.begin() + pos + faceBytes);
    pos += (faceBytes + 3) & ~3u;
  }
  size = width;
  return true;
}

bool writePpm(const std::string& path, uint32_t W, uint32_t H, const uint8_t* rgb) {
  std::ofstream out(path, std::ios::binary);
  if (!out.good()) return false;
  out << "P6\n" << W << " " << H << "\n255\n";
  out.write(reinterpret_cast<const char*>(rgb), std::streamsize(size_t(W) * H * 3));
  return out.good();
}
}  // namespace kuafu::io
