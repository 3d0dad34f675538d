// Compile-and-link check of the header-only adapter against liblsd_b200.so (built by `make host-check`).
// Runs a tiny track + map step when a GPU is present; prints "no device" and exits 0 otherwise.
#include <cstdio>
#include <vector>

#include "lsd_b200.hpp"

int main() {
  using namespace lsd_b200;
  const int w = 64, h = 48;
  try {
    Context ctx(w, h, 52.5f, 52.5f, 31.5f, 23.5f);
    std::vector<unsigned char> a(w * h), b(w * h);
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) {
        a[x + y * w] = (unsigned char)(128 + 60 * ((x / 4 + y / 4) & 1) + (x * 7 + y * 13) % 17);
        b[x + y * w] = a[x + y * w];
      }
    Frame kf(ctx, 0, 0.0, a.data()), fr(ctx, 1, 0.033, b.data());
    std::vector<float> depth(w * h, 2.0f);
    kf.setDepthFromGroundTruth(depth.data());
    DepthMap map(ctx);
    map.initializeFromGTDepth(&kf);
    TrackingReference ref(ctx);
    ref.importFrame(&kf);
    SE3Tracker tracker(ctx);
    const SE3 pose = tracker.trackFrame(&ref, &fr, SE3());
    std::vector<lsd_vertex> vbo(w * h);
    const int points = KeyframePublisher::computeVbo(ctx, kf, 0, 1.0f, vbo.data());
    SlamSystem slam(ctx);
    slam.gtDepthInit(0, a.data(), w, depth.data());
    slam.nextImage(1, b.data(), w);
    std::printf("vbo points = %d  pose line: %s", points, slam.poseLine().c_str());
    std::printf("tracked: t = %g %g %g  good = %g  diverged = %d\n", pose.d[4], pose.d[5], pose.d[6], tracker.lastGoodCount, (int)tracker.diverged);
  } catch (const Error &e) {
    std::printf("no device (%s)\n", e.what());
  }
  return 0;
}
