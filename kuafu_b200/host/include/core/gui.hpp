// Inert stand-in for the reference's (already stubbed-out) ImGui layer (include/core/gui.hpp).
#pragma once
namespace kuafu {
class Gui {
 public:
  virtual ~Gui() = default;
  virtual void configure() {}
  virtual void render() {}
};
}  // namespace kuafu
