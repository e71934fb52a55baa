// Texture / cube-map file decoding for the host facade (stands in for the reference's stb_image +
// libktx use in vkCore.hpp:1761-1795,1853-1971).  Everything is expanded to RGBA8 like
// stbi_load(..., STBI_rgb_alpha).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace kuafu::io {
/// "mem:<name>" (global::registerMemoryTexture), PNG (all colour types and bit depths, Adam7), baseline JPEG,
/// .ppm/.pgm (P6/P5) -- recognised by content, not by extension.
bool loadTextureRGBA8(const std::string& path, uint32_t& width, uint32_t& height, std::vector<uint8_t>& rgba);
/// KTX1, uncompressed RGBA8 / SRGB8_ALPHA8, six faces, level 0.
bool loadKtxCubeRGBA8(const std::string& path, uint32_t& size, std::vector<uint8_t> faces[6]);
bool decodePng(const uint8_t* data, size_t n, uint32_t& width, uint32_t& height, std::vector<uint8_t>& rgba);
/// Baseline (sequential, Huffman, 8-bit) JPEG, grey or YCbCr; progressive / arithmetic / CMYK files return false.
bool decodeJpeg(const uint8_t* data, size_t n, uint32_t& width, uint32_t& height, std::vector<uint8_t>& rgba);
bool writePpm(const std::string& path, uint32_t width, uint32_t height, const uint8_t* rgb);
}  // namespace kuafu::io
