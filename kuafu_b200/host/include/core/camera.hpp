// Camera of the kuafu API (reference include/core/camera.hpp:40-192, src/core/camera.cpp).
// Pure host math that fills CameraUBO; the frame itself lives on the device until
// downloadLatestFrame() copies it out.
#pragma once
#include "core/context/global.hpp"

namespace kuafu {
class Context;
class Scene;

/// What a camera keeps of its last rendered frame (stands in for the reference's per-camera
/// Frames/RenderTargets/Sync Vulkan objects).
struct FrameStore {
  Context* owner = nullptr;  ///< context whose device buffers hold the frame, or null
  uint32_t slot = 0;         ///< camera slot inside the context's last render
  uint64_t serial = 0;       ///< render serial the slot belongs to
  std::vector<uint8_t> stash;  ///< BGRA8 copy taken when another camera displaced this frame
  bool valid = false;
};

class KUAFU_API Camera {
 public:
  Camera(const Camera&) = delete;
  Camera& operator=(const Camera&) = delete;
  /// Unregisters from the context that still remembers this camera's last frame.
  ~Camera();

  void update();
  void resetView();

  auto getPosition() const -> const glm::vec3& { return mPosition; }
  auto getFront() const -> const glm::vec3& { return mDirFront; }

  void setPosition(const glm::vec3& position);
  void setFront(const glm::vec3& front);
  void setUp(const glm::vec3& up);

  [[nodiscard]] inline float getAperture() const { return mAperture; }
  [[nodiscard]] inline float getFocalLength() const { return mFocalLength; }
  /// Additive (commented out in the reference): thin-lens parameters read by the ray generator.
  void setAperture(float aperture) { mAperture = aperture; }
  void setFocalLength(float focalLength) { mFocalLength = focalLength; }

  void setSize(int width, int height);
  void setPose(glm::mat4 pose);
  void setFullPerspective(float width, float height, float fx, float fy, float cx, float cy, float skew);
  glm::mat4 getPose() const;

  [[nodiscard]] auto getWidth() const { return mWidth; }
  [[nodiscard]] auto getHeight() const { return mHeight; }

  auto getViewMatrix() const -> const glm::mat4& { return mViewMatrix; }
  auto getProjectionMatrix() const -> const glm::mat4& { return mProjMatrix; }
  auto getViewInverseMatrix() const -> glm::mat4 { return glm::inverse(mViewMatrix); }
  auto getProjectionInverseMatrix() const -> glm::mat4 { return glm::inverse(mProjMatrix); }

  void updateViewMatrix();
  void updateProjectionMatrix();

  void processMouse(float xOffset, float yOffset);
  void processKeyboard();

  [[nodiscard]] inline float getPrincipalPointX() const { return mCx; }
  [[nodiscard]] inline float getPrincipalPointY() const { return mCy; }
  [[nodiscard]] inline float getFocalX() const { return mFx; }
  [[nodiscard]] inline float getFocalY() const { return mFy; }
  [[nodiscard]] inline float getNear() const { return mNear; }
  [[nodiscard]] inline float getFar() const { return mFar; }
  [[nodiscard]] inline float getSkew() const { return mSkew; }

  /// width*height*4 bytes, BGRA, sRGB-encoded, alpha 255 (reference camera.cpp:188-207).
  std::vector<uint8_t> downloadLatestFrame();
  /// Additive: the same bytes straight into the caller's buffer (width*height*4), without the by-value
  /// vector in between -- for bindings that already own the destination (a numpy array, a ROS message).
  void downloadLatestFrameInto(uint8_t* dst, size_t nbytes);
  /// Additive outputs the reference lists as TODO (README.md:64-68), all from sample 0 / depth 0.
  std::vector<float> downloadDepth();
  std::vector<int32_t> downloadSegmentation();
  std::vector<int32_t> downloadHitIds();   ///< (instance, primitive) pairs
  std::vector<float> downloadRadiance();   ///< float4 running-mean image
  std::vector<float> downloadAlbedo();     ///< float4, denoiser hand-off
  std::vector<float> downloadNormal();     ///< float4, denoiser hand-off

  bool mFirst = true;
  FrameStore mFrames;

 private:
  friend class Scene;
  friend class Context;

  Camera(int width, int height, const glm::vec3& position = {0.0F, 0.0F, 3.0F});
  std::vector<uint8_t> downloadAuxBytes(int kind, size_t bytesPerPixel);

  int mWidth;
  int mHeight;
  glm::vec3 mPosition;
  float mFx, mFy, mCx, mCy;
  float mSkew = 0;

  glm::mat4 mViewMatrix = glm::mat4(1.0F);
  glm::mat4 mProjMatrix = glm::mat4(1.0F);

  glm::vec3 mDirUp = {0.0F, 0.0F, 1.0F};
  glm::vec3 mDirRight = {0.0F, -1.0F, 0.0F};
  glm::vec3 mDirFront = {1.0F, 0.0F, 0.0F};

  float mAperture = 0.0F;
  float mFocalLength = 5.0F;

  const float mFar = 100.F;
  const float mNear = 0.1F;
  glm::vec3 mResetPosition;
  glm::vec3 mPrevPosition;
};

/// 320-byte camera block read by the ray generator (== KfrtCamera).
struct CameraUBO {
  glm::mat4 view = glm::mat4(1.0F);
  glm::mat4 projection = glm::mat4(1.0F);
  glm::mat4 viewInverse = glm::mat4(1.0F);
  glm::mat4 projectionInverse = glm::mat4(1.0F);
  glm::vec4 position = glm::vec4(1.0F);  // position + aperture
  glm::vec4 front = glm::vec4(1.0F);     // front + focus distance
  glm::vec4 padding1 = glm::vec4(1.0F);
  glm::vec4 padding2 = glm::vec4(1.0F);
};
static_assert(sizeof(CameraUBO) == 320, "camera wire format");
}  // namespace kuafu
