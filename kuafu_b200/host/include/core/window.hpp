// Inert stand-in for the reference's SDL window (include/core/window.hpp): a B200 node has no
// display, the renderer is offscreen-only.  Kept so that code written against kuafu.hpp links.
#pragma once
#include "stdafx.hpp"

namespace kuafu {
class KUAFU_API Window {
 public:
  Window(int width = 800, int height = 600, const char* title = "App", uint32_t flags = 0)
      : mWidth(width), mHeight(height), mTitle(title), mFlags(flags) {}
  virtual ~Window() = default;
  virtual bool init() { return true; }
  virtual bool update() { return true; }
  void resize(int width, int height) { mWidth = width; mHeight = height; }
  int getWidth() const { return mWidth; }
  int getHeight() const { return mHeight; }
  bool changed() { return false; }
  bool minimized() { return false; }

 protected:
  int mWidth, mHeight;
  std::string mTitle;
  uint32_t mFlags;
};
}  // namespace kuafu
