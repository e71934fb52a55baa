// Scene recipes for the five BASELINE.json configs, written against the public kuafu.hpp API the
// way the reference's example app builds its levels (reference example/Example.hpp:79-787).  They
// are the fixtures of the parity tests and of bench.py.
#pragma once
#include "kuafu.hpp"

namespace kuafu::scenes {

struct Recipe {
  std::string name;
  int width = 0, height = 0;       ///< 0: the config's own resolution
  int spp = 0, depth = -1;         ///< 0 / -1: the config's own values
  int scale = 0;                   ///< 0: full size; otherwise a reduced instance/camera count for tests
};

/// Builds the named scene into `renderer`'s current scene and returns the cameras to render
/// (1 for configs 1-3, 2 for the stereo pair of config 4, 64 for config 5).
///   "spheres"      config 1  eSpheres (Example.hpp:128-327), 277 444 triangles, 10 instances
///   "cornell"      config 2  synthetic Cornell box: area light, glass sphere, rough-specular sphere
///   "million"      config 3  204 x createSphere + floor = 999 602 triangles, 16 textured materials, env cube
///   "active"       config 4  eActive (Example.hpp:479-645) with an IR dot-pattern projector, stereo pair
///   "articulated"  config 5  2048 link instances (64 chains x 32 links), 64 cameras, animate() per frame
///   "million_obj"  config 3 with every mesh written to a Wavefront file and read back by loadScene()
///   "spheres_ref"  config 1 with the reference's resources/models/suzanne.dae   } only where the reference tree
///   "active_ref"   config 4 with suzanne.dae and resources/patterns/fakesense_j415.png } is mounted (KUAFU_REFERENCE)
///   "unique"       stress: config 3's layout with 204 un-instanced displaced blobs (1.0 M unique triangles)
///   "unique10m"    stress: 1 000 un-instanced blobs of 10 000 triangles (10.0 M unique triangles)
///   "file:<path>"  the meshes of an asset file (.obj / .dae / .stl / .gltf / .glb), framed by one camera
KUAFU_API std::vector<Camera*> load(Kuafu& renderer, const Recipe& recipe);

/// Per-frame actor motion of "articulated": deterministic joint angles -> GeometryInstance::setTransform.
KUAFU_API void animate(Kuafu& renderer, int frame);

/// Displaced UV sphere standing in for resources/models/suzanne.dae (not shippable): same triangle
/// count as the reference asset (251 904) at the default tessellation.  A non-zero seed gives the blob its
/// own shape (the stress recipes with unique geometry).
KUAFU_API std::shared_ptr<Geometry> createBlob(NiceMaterial mat, uint32_t slices = 512, uint32_t stacks = 247,
                                               uint32_t seed = 0);
}  // namespace kuafu::scenes
