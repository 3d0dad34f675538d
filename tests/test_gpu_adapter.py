"""The C++ adapter on the device: a consumer written against the lsd-slam include tree of host/compat/ --
upstream-shaped Frame constructor, SE3Tracker::trackFrame, DepthMap, and the exact accessor sequence of the reference's
PangolinOutputIOWrapper::publishKeyframe / TextOutputIOWrapper::publishTrackedFrame -- compiled with g++ on the GPU box
and run against liblsd_b200.so."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lsd-slam-pangolin-gui_b200")

SRC = r'''
#include <cmath>
#include <cstdio>
#include <sstream>
#include <vector>
#include "DataStructures/Frame.h"
#include "GlobalMapping/KeyFrameGraph.h"
using namespace lsd_slam;
int main() {
  const int w = 320, h = 240;
  lsd_b200::Context ctx(w, h, 262.5f, 262.5f, 159.5f, 119.5f);
  std::vector<unsigned char> a(w * h), b(w * h);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      a[x + y * w] = (unsigned char)(128 + 50 * std::sin(x * 0.21) * std::cos(y * 0.17) + 30 * std::sin((x + 2 * y) * 0.05));
      b[x + y * w] = a[x + y * w];
    }
  Eigen::Matrix3f K;
  K(0, 0) = 262.5f; K(1, 1) = 262.5f; K(0, 2) = 159.5f; K(1, 2) = 119.5f; K(2, 2) = 1;
  Frame::SharedPtr kf(new Frame(0, w, h, K, 0.0, a.data())), fr(new Frame(1, w, h, K, 0.033, b.data()));
  std::vector<float> depth(w * h, 2.0f);
  kf->setDepthFromGroundTruth(depth.data());
  kf->pose->isRegisteredToGraph = true;  // the first keyframe: camToWorld = identity
  lsd_b200::DepthMap map(ctx);
  map.initializeFromGTDepth(kf.get());
  lsd_b200::TrackingReference ref(ctx);
  ref.importFrame(kf.get());
  lsd_b200::SE3Tracker tracker(ctx);
  const lsd_b200::SE3 pose = tracker.trackFrame(&ref, fr.get(), lsd_b200::SE3());
  if (tracker.diverged || !tracker.trackingWasGood) { std::printf("tracking failed\n"); return 1; }
  // ---- the accessor sequence of PangolinOutputIOWrapper::publishKeyframe (PangolinOutputIOWrapper.cpp:50-89)
  const int publishLvl = 1;
  boost::shared_lock<boost::shared_mutex> lock = kf->getActiveLock();
  const int pw = kf->width(publishLvl), ph = kf->height(publishLvl);
  const Sophus::Sim3f camToWorld = kf->getCamToWorld().cast<float>();
  const float fx = kf->fx(publishLvl), fy = kf->fy(publishLvl), cx = kf->cx(publishLvl), cy = kf->cy(publishLvl);
  if (!kf->hasIDepthBeenSet()) return 2;
  const float *idepth = kf->idepth(publishLvl), *idepthVar = kf->idepthVar(publishLvl), *color = kf->image(publishLvl);
  double s_id = 0, s_c = 0;
  int valid = 0;
  for (int idx = 0; idx < pw * ph; idx++) {
    if (idepthVar[idx] > 0) { s_id += idepth[idx]; valid++; }
    s_c += color[idx];
  }
  lock.unlock();
  if (pw != 160 || ph != 120 || fx != 131.25f || fy != 131.25f || cx != 79.5f || cy != 59.5f) return 3;
  if (valid < pw * ph / 2 || std::fabs(s_id / valid - 0.5) > 1e-3 || camToWorld.scale() != 1.0f) return 4;
  // ---- the stream operations of TextOutputIOWrapper::publishTrackedFrame (TextOutputIOWrapper.cpp:100-120)
  std::ostringstream os;
  os << fr->id();
  { const auto p = fr->getCamToWorld(); const auto t = p.translation(); os << "," << t.x() << "," << t.y() << "," << t.z(); }
  { const auto p = fr->pose->thisToParent_raw; const auto t = p.translation(); os << "," << t.x() << "," << t.y() << "," << t.z(); }
  // pose->thisToParent_raw is what the tracker returned; the keyframe is the world origin, so both triples agree
  const auto t = fr->pose->thisToParent_raw.translation();
  if (t.x() != pose.d[4] || t.y() != pose.d[5] || t.z() != pose.d[6]) return 5;
  if (fr->pose->trackingParent != kf->pose) return 6;
  if (std::fabs(t.x()) > 1e-3 || std::fabs(t.y()) > 1e-3 || std::fabs(t.z()) > 1e-3) return 7;  // identical images: no motion
  std::printf("pose line: %s\nok valid=%d mean idepth=%.6f mean colour=%.3f\n", os.str().c_str(), valid, s_id / valid, s_c / (pw * ph));
  return 0;
}
'''


def test_compat_consumer_runs_on_the_device(tmp_path):
    src = tmp_path / "consumer_gpu.cpp"
    src.write_text(SRC)
    exe = tmp_path / "consumer_gpu"
    r = subprocess.run(["g++", "-std=c++14", "-Wall", "-O1", str(src), "-o", str(exe), "-I", os.path.join(ROOT, "tests", "consumer_shim"),
                        "-I", os.path.join(PKG, "host", "compat"), "-L", PKG, "-llsd_b200", f"-Wl,-rpath,{PKG}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "ok valid=" in out.stdout and out.stdout.startswith("pose line: 1,")


def test_adapter_check_runs_on_the_device():
    subprocess.check_call(["make", "-C", PKG, "-s", "host-check"])
    out = subprocess.run([os.path.join(PKG, "build", "adapter_check")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "tracked:" in out.stdout, (out.stdout, out.stderr)
