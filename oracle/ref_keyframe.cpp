// oracle/ref_keyframe.cpp -- TEST INFRASTRUCTURE ONLY: C entry point over the REFERENCE'S OWN code.
//
// Compiles /root/reference/lib/Pangolin_IOWrapper/Keyframe.h (unmodified, included from where it lies;
// nothing of it is copied into this repository) against the stand-in headers of oracle/ref_shim/ and
// runs the reference's Keyframe::computeVbo (Keyframe.h:66-158) on caller data.  The vertex buffer the
// reference would upload with glBufferData is captured by the shim and returned.
// Built by oracle/Makefile into oracle/_ref/libref_keyframe.so (git-ignored, travels to the GPU box).
#include "Pangolin_IOWrapper/Keyframe.h"

static_assert(sizeof(InputPointDense) == 12, "InputPointDense layout");
static_assert(sizeof(Keyframe::MyVertex) == 16, "MyVertex layout");

extern "C" {

// in: width*height InputPointDense (12 B each).  out: capacity width*height MyVertex (16 B each).
// Returns Keyframe::points after computeVbo(), or -1 when the captured buffer and the count disagree.
int ref_keyframe_compute_vbo(const void *pointData, int width, int height, float fx, float fy, float cx, float cy, float camToWorldScale,
                             void *outVertices) {
  Keyframe kf;
  kf.width = width;
  kf.height = height;
  kf.fx = fx;
  kf.fy = fy;
  kf.cx = cx;
  kf.cy = cy;
  kf.camToWorld.s = camToWorldScale;
  const size_t bytes = (size_t)width * height * sizeof(InputPointDense);
  kf.pointData = new unsigned char[bytes];  // computeVbo delete[]s it (Keyframe.h:154)
  memcpy(kf.pointData, pointData, bytes);
  kf.computeVbo();
  const auto &cap = ref_shim::capture();
  if (cap.bufferData.size() != sizeof(Keyframe::MyVertex) * (size_t)kf.points) return -1;
  memcpy(outVertices, cap.bufferData.data(), cap.bufferData.size());
  return kf.points;
}

// Keyframe::updatePoints (Keyframe.h:54-64) followed by computeVbo: the path GUI::addKeyframe takes when a
// keyframe id is published again (lib/GUI.cpp:126-131).
int ref_keyframe_update_and_compute_vbo(const void *firstPointData, const void *secondPointData, int width, int height, float fx, float fy,
                                        float cx, float cy, float camToWorldScale, void *outVertices) {
  Keyframe kf, newer;
  const size_t bytes = (size_t)width * height * sizeof(InputPointDense);
  for (Keyframe *k : {&kf, &newer}) {
    k->width = width; k->height = height;
    k->fx = fx; k->fy = fy; k->cx = cx; k->cy = cy;
    k->camToWorld.s = camToWorldScale;
    k->pointData = new unsigned char[bytes];
  }
  memcpy(kf.pointData, firstPointData, bytes);
  memcpy(newer.pointData, secondPointData, bytes);
  kf.computeVbo();
  kf.updatePoints(&newer);
  kf.computeVbo();
  const auto &cap = ref_shim::capture();
  if (cap.bufferData.size() != sizeof(Keyframe::MyVertex) * (size_t)kf.points) return -1;
  memcpy(outVertices, cap.bufferData.data(), cap.bufferData.size());
  return kf.points;
}
}
