#include "image_io.hpp"

#include <zlib.h>

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

bool decodePng(const uint8_t* d, size_t n, uint32_t& W, uint32_t& H, std::vector<uint8_t>& rgba) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (n < 8 || std::memcmp(d, sig, 8) != 0) return false;
  size_t pos = 8;
  int colorType = -1, bitDepth = 0, interlace = 0;
  std::vector<uint8_t> idat, palette, trns;
  while (pos + 12 <= n) {
    const uint32_t len = be32(d + pos);
    const uint8_t* type = d + pos + 4;
    const uint8_t* body = d + pos + 8;
    if (pos + 12 + len > n) return false;
    if (!std::memcmp(type, "IHDR", 4)) {
      W = be32(body);
      H = be32(body + 4);
      bitDepth = body[8];
      colorType = body[9];
      interlace = body[12];
    } else if (!std::memcmp(type, "PLTE", 4)) {
      palette.assign(body, body + len);
    } else if (!std::memcmp(type, "tRNS", 4)) {
      trns.assign(body, body + len);
    } else if (!std::memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), body, body + len);
    } else if (!std::memcmp(type, "IEND", 4)) {
      break;
    }
    pos += 12 + len;
  }
  if (bitDepth != 8 || interlace != 0 || W == 0 || H == 0) return false;
  int ch;
  switch (colorType) {
    case 0: ch = 1; break;
    case 2: ch = 3; break;
    case 3: ch = 1; break;
    case 4: ch = 2; break;
    case 6: ch = 4; break;
    default: return false;
  }
  const size_t stride = size_t(W) * ch;
  std::vector<uint8_t> raw((stride + 1) * H);
  uLongf rawLen = uLongf(raw.size());
  if (uncompress(raw.data(), &rawLen, idat.data(), uLong(idat.size())) != Z_OK || rawLen != raw.size()) return false;
  std::vector<uint8_t> img(stride * H);
  for (uint32_t y = 0; y < H; y++) {
    const uint8_t ft = raw[(stride + 1) * y];
    const uint8_t* src = &raw[(stride + 1) * y + 1];
    uint8_t* dst = &img[stride * y];
    const uint8_t* up = y ? &img[stride * (y - 1)] : nullptr;
    for (size_t x = 0; x < stride; x++) {
      const int a = x >= size_t(ch) ? dst[x - ch] : 0;
      const int b = up ? up[x] : 0;
      const int c = (up && x >= size_t(ch)) ? up[x - ch] : 0;
      int v = src[x];
      switch (ft) {
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: break;
      }
      dst[x] = uint8_t(v);
    }
  }
  rgba.resize(size_t(W) * H * 4);
  for (size_t i = 0; i < size_t(W) * H; i++) {
    uint8_t r, g, b, a = 255;
    const uint8_t* p = &img[i * ch];
    if (colorType == 0) { r = g = b = p[0]; }
    else if (colorType == 2) { r = p[0]; g = p[1]; b = p[2]; }
    else if (colorType == 3) {
      const size_t k = p[0];
      if (3 * k + 2 >= palette.size()) return false;
      r = palette[3 * k]; g = palette[3 * k + 1]; b = palette[3 * k + 2];
      if (k < trns.size()) a = trns[k];
    } else if (colorType == 4) { r = g = b = p[0]; a = p[1]; }
    else { r = p[0]; g = p[1]; b = p[2]; a = p[3]; }
    rgba[4 * i] = r; rgba[4 * i + 1] = g; rgba[4 * i + 2] = b; rgba[4 * i + 3] = a;
  }
  return true;
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
  if (glType != 0x1401 /*UNSIGNED_BYTE*/ || glFormat != 0x1908 /*RGBA*/ || nFaces != 6 || width != height || !width)
    return false;
  size_t pos = 64 + kvBytes;
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
