// Camera host math (reference src/core/camera.cpp).  Projection from pinhole intrinsics with the
// reference's sign conventions (negative [1][1]: image y grows downwards), view = lookAt.
#include "core/camera.hpp"

#include <cstdlib>

#include "core/context/context.hpp"
#include "kf_rt.h"

namespace kuafu {
Camera::~Camera() {
  if (mFrames.owner) mFrames.owner->forgetCamera(this);
}

Camera::Camera(int width, int height, const glm::vec3& position)
    : mWidth(width), mHeight(height), mPosition(position), mResetPosition(position), mPrevPosition(position) {
  mCx = mWidth * 0.5f;
  mCy = mHeight * 0.5f;
  mFx = mFy = mWidth * 0.5f;
  updateProjectionMatrix();
  resetView();
}

void Camera::resetView() {
  mPosition = mResetPosition;
  mDirUp = {0.0F, 0.0F, 1.0F};
  mDirRight = {0.0F, -1.0F, 0.0F};
  mDirFront = {1.0F, 0.0F, 0.0F};
  updateViewMatrix();
  global::frameCount = -1;
}

glm::mat4 Camera::getPose() const {
  return glm::mat4(glm::vec4(mDirFront, 0.f), glm::vec4(-mDirRight, 0.f), glm::vec4(mDirUp, 0.f),
                   glm::vec4(mPosition, 1.f));
}

void Camera::update() {
  if (mPrevPosition != mPosition) {  // a moved camera restarts the accumulation
    global::frameCount = -1;
    mPrevPosition = mPosition;
  }
  processKeyboard();
  updateViewMatrix();
}

void Camera::setPosition(const glm::vec3& position) {
  mPosition = position;
  updateViewMatrix();
}
void Camera::setFront(const glm::vec3& front) {
  mDirFront = front;
  updateViewMatrix();
}
void Camera::setUp(const glm::vec3& up) {
  mDirUp = up;
  updateViewMatrix();
}

void Camera::setSize(int width, int height) {
  KF_WARN("Resize by dragging not recommended and will be deprecated! Hint: use offscreen mode for production.");
  mWidth = width;
  mHeight = height;
  mCx = mWidth / 2.f;
  mCy = mHeight / 2.f;
  updateProjectionMatrix();
}

void Camera::updateViewMatrix() { mViewMatrix = glm::lookAt(mPosition, mPosition + mDirFront, mDirUp); }

void Camera::setFullPerspective(float width, float height, float fx, float fy, float cx, float cy, float skew) {
  mWidth = static_cast<int>(width);
  mHeight = static_cast<int>(height);
  mFx = fx;
  mFy = fy;
  mCx = cx;
  mCy = cy;
  mSkew = skew;
  updateProjectionMatrix();
}

void Camera::updateProjectionMatrix() {
  glm::mat4 p(0.0f);
  p[0][0] = (2.f * mFx) / mWidth;
  p[1][0] = -2 * mSkew / mWidth;
  p[1][1] = -(2.f * mFy) / mHeight;
  p[2][0] = -2.f * mCx / mWidth + 1;
  p[2][1] = -2.f * mCy / mHeight + 1;
  p[2][2] = -mFar / (mFar - mNear);
  p[2][3] = -1.f;
  p[3][2] = -mFar * mNear / (mFar - mNear);
  p[3][3] = 0.f;
  mProjMatrix = p;
  mFrames.valid = false;  // the reference destroys the camera's frames here
  mFrames.stash.clear();
}

void Camera::setPose(glm::mat4 pose) {
  mPosition = {pose[3][0], pose[3][1], pose[3][2]};
  mDirFront = {-pose[2][0], -pose[2][1], -pose[2][2]};
  mDirUp = {pose[1][0], pose[1][1], pose[1][2]};
  global::frameCount = -1;
}

void Camera::processMouse(float xOffset, float /*yOffset*/) {
  const glm::mat4 urot = glm::rotate(glm::mat4(1.0f), xOffset * 0.01f, mDirUp);
  const glm::vec4 f = urot * glm::vec4(mDirFront, 0.f);
  mDirFront = {f.x, f.y, f.z};
  global::frameCount = -1;
}

void Camera::processKeyboard() {
  // viewer-only in the reference (WASD fly camera); there is no window here, but honour the key
  // flags for code that drives them programmatically
  const float step = 2.5F * (1.0f / 60.0f) * (global::keys::eLeftShift ? 4.0f : (global::keys::eLeftCtrl ? 0.2f : 1.0f));
  if (global::keys::eW) mPosition += mDirFront * step;
  if (global::keys::eS) mPosition -= mDirFront * step;
  if (global::keys::eA) mPosition -= mDirRight * step;
  if (global::keys::eD) mPosition += mDirRight * step;
}

std::vector<uint8_t> Camera::downloadAuxBytes(int kind, size_t bytesPerPixel) {
  KF_ASSERT(mFrames.valid && mFrames.owner, "Invalid call to Camera::downloadLatestFrame");
  Context* ctx = mFrames.owner;
  KF_ASSERT(mFrames.serial == ctx->mSerial, "This camera's auxiliary buffers were displaced by a later render");
  std::vector<uint8_t> out(size_t(mWidth) * mHeight * bytesPerPixel);
  ctx->check(kfrtDownloadAux(ctx->getDevice(), mFrames.slot, kind, out.data(), out.size()), "kfrtDownloadAux");
  return out;
}

std::vector<uint8_t> Camera::downloadLatestFrame() {
  KF_ASSERT(mFrames.valid, "Invalid call to Camera::downloadLatestFrame");
  if (mFrames.owner && mFrames.serial == mFrames.owner->mSerial) {
    // the frame is already on its way to pinned host memory (kfrtResolve started the copy): the vector
    // the reference's signature returns is built from it in one pass
    Context* ctx = mFrames.owner;
    const uint8_t* bytes = nullptr;
    size_t n = 0;
    if (kfrtMapBGRA8(ctx->getDevice(), mFrames.slot, &bytes, &n) == KFRT_OK && n == size_t(mWidth) * mHeight * 4)
      return std::vector<uint8_t>(bytes, bytes + n);
    return downloadAuxBytes(KFRT_AUX_BGRA8, 4);  // rendered but not resolved (deferred resolve): the device copy
  }
  KF_ASSERT(!mFrames.stash.empty(), "Invalid call to Camera::downloadLatestFrame");
  return mFrames.stash;
}

void Camera::downloadLatestFrameInto(uint8_t* dst, size_t nbytes) {
  KF_ASSERT(mFrames.valid && dst, "Invalid call to Camera::downloadLatestFrameInto");
  KF_ASSERT(nbytes == size_t(mWidth) * mHeight * 4, "downloadLatestFrameInto: destination size mismatch");
  if (mFrames.owner && mFrames.serial == mFrames.owner->mSerial) {
    Context* ctx = mFrames.owner;
    ctx->check(kfrtDownloadBGRA8(ctx->getDevice(), mFrames.slot, dst, nbytes), "kfrtDownloadBGRA8");
    return;
  }
  KF_ASSERT(mFrames.stash.size() == nbytes, "Invalid call to Camera::downloadLatestFrameInto");
  std::memcpy(dst, mFrames.stash.data(), nbytes);
}

template <typename T>
static std::vector<T> reinterpretBytes(const std::vector<uint8_t>& b) {
  std::vector<T> out(b.size() / sizeof(T));
  std::memcpy(out.data(), b.data(), out.size() * sizeof(T));
  return out;
}
std::vector<float> Camera::downloadDepth() { return reinterpretBytes<float>(downloadAuxBytes(KFRT_AUX_DEPTH, 4)); }
std::vector<int32_t> Camera::downloadSegmentation() { return reinterpretBytes<int32_t>(downloadAuxBytes(KFRT_AUX_SEGMENTATION, 4)); }
std::vector<int32_t> Camera::downloadHitIds() { return reinterpretBytes<int32_t>(downloadAuxBytes(KFRT_AUX_HIT_IDS, 8)); }
std::vector<float> Camera::downloadRadiance() { return reinterpretBytes<float>(downloadAuxBytes(KFRT_AUX_RGBA32F, 16)); }
std::vector<float> Camera::downloadAlbedo() { return reinterpretBytes<float>(downloadAuxBytes(KFRT_AUX_ALBEDO32F, 16)); }
std::vector<float> Camera::downloadNormal() { return reinterpretBytes<float>(downloadAuxBytes(KFRT_AUX_NORMAL32F, 16)); }
}  // namespace kuafu
