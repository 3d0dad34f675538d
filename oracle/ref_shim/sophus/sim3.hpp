// oracle/ref_shim/sophus/sim3.hpp -- TEST INFRASTRUCTURE ONLY (oracle/_ref recipe).
// The two Sophus members lib/Pangolin_IOWrapper/Keyframe.h touches: Sim3f::scale() (computeVbo,
// Keyframe.h:83) and Sim3f::matrix() (drawPoints / drawCamera, never reached by the oracle).
#pragma once
namespace Sophus {
struct Matrix4f {
  float m[16];
  float *data() { return m; }
};
struct Sim3f {
  float s = 1.0f;
  float scale() const { return s; }
  Matrix4f matrix() const {
    Matrix4f r{};
    r.m[0] = r.m[5] = r.m[10] = s;
    r.m[15] = 1.0f;
    return r;
  }
};
}  // namespace Sophus
