"""Drop-in check of the upper face of the boundary: the reference's OWN output wrappers --
/root/reference/lib/Pangolin_IOWrapper/PangolinOutputIOWrapper.cpp and TextOutputIOWrapper.cpp, unmodified, compiled from where
they lie -- build against the lsd-slam include tree of lsd-slam-pangolin-gui_b200/host/compat/, i.e. against lsd_slam::Frame =
the device-resident frame of liblsd_b200.  Every accessor they call (id, timestamp, width/height/fx/fy/cx/cy(level),
image/idepth/idepthVar(level), hasIDepthBeenSet, getCamToWorld().cast<float>(), getActiveLock(), pose->thisToParent_raw,
KeyFrameGraph::keyframesAll / keyframesAllMutex) must exist with a compatible type.  Third-party headers (Eigen, Sophus, Boost,
OpenCV, Pangolin, glm, g3log, libvideoio) are stand-ins under tests/consumer_shim/ (test infrastructure only)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PKG = os.path.join(ROOT, "lsd-slam-pangolin-gui_b200")
INC = ["-I", os.path.join(ROOT, "tests", "consumer_shim"), "-I", os.path.join(PKG, "host", "compat"),
       "-I", os.path.join(ROOT, "oracle", "ref_shim"),  # GL/glew.h stand-in (shared with the oracle/_ref recipe)
       "-I", os.path.join(REF, "lib"), "-I", os.path.join(REF, "lib", "Pangolin_IOWrapper")]


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is not present on this box")
@pytest.mark.parametrize("src", ["TextOutputIOWrapper.cpp", "PangolinOutputIOWrapper.cpp"])
def test_reference_output_wrapper_compiles_against_the_adapter(src, tmp_path):
    path = os.path.join(REF, "lib", "Pangolin_IOWrapper", src)
    obj = tmp_path / (src + ".o")
    r = subprocess.run(["g++", "-std=c++14", "-Wall", "-c", path, "-o", str(obj)] + INC, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    assert obj.stat().st_size > 0
    syms = subprocess.run(["nm", "-C", str(obj)], capture_output=True, text=True).stdout
    cls = src[:-4]
    assert f"lsd_slam::{cls}::publishKeyframe" in syms and f"lsd_slam::{cls}::publishTrackedFrame" in syms


def test_adapter_consumer_runs_end_to_end_without_a_gpu(tmp_path):
    """Everything of the compat surface that does not touch the device, executed: the FramePoseStruct chain behind
    getCamToWorld(), the Sophus <-> C-ABI pose conversion, per-level intrinsics, and the pose.txt line the reference's
    TextOutputIOWrapper would write for those poses (same stream operations)."""
    src = tmp_path / "consumer.cpp"
    src.write_text(r'''
#include <cmath>
#include <cstdio>
#include <sstream>
#include "DataStructures/Frame.h"
#include "GlobalMapping/KeyFrameGraph.h"
#include "IOWrapper/OutputIOWrapper.h"
using namespace lsd_slam;
int main() {
  // pose chain: keyframe (registered, scale 2) <- frame tracked on it
  FramePoseStruct kf, fr;
  const double kfw[8] = {0, 0, 0, 1, 1.0, 2.0, 3.0, 2.0}, f2k[8] = {0, std::sin(0.05), 0, std::cos(0.05), 0.1, -0.2, 0.3, 1.0};
  kf.isRegisteredToGraph = true;
  kf.camToWorld = lsd_b200::toSophus(kfw);
  fr.trackingParent = &kf;
  fr.thisToParent_raw = lsd_b200::toSophus(f2k);
  const Sophus::Sim3d w = fr.getCamToWorld();
  double back[8];
  lsd_b200::fromSophus(w, back);
  // camToWorld = kf * f2k: t = t_kf + s_kf * R_kf * t_f = (1.2, 1.6, 3.6), scale 2
  if (std::fabs(back[4] - 1.2) > 1e-12 || std::fabs(back[5] - 1.6) > 1e-12 || std::fabs(back[6] - 3.6) > 1e-12) return 1;
  if (std::fabs(back[7] - 2.0) > 1e-12 || std::fabs(back[1] - std::sin(0.05)) > 1e-12) return 2;
  const Sophus::Sim3f wf = w.cast<float>();
  if (std::fabs(wf.scale() - 2.0f) > 1e-6f) return 3;
  float seven[7];
  std::memcpy(seven, wf.data(), sizeof(float) * 7);  // what publishKeyframeGraph copies into GraphFramePose::camToWorld
  if (std::fabs(seven[4] - 1.2f) > 1e-6f) return 4;
  // the pose.txt line, with the stream operations of TextOutputIOWrapper::publishTrackedFrame
  std::ostringstream os;
  os << 7;
  { const auto pose = fr.getCamToWorld(); const auto trans = pose.translation(); os << "," << trans.x() << "," << trans.y() << "," << trans.z(); }
  { const auto pose = fr.thisToParent_raw; const auto trans = pose.translation(); os << "," << trans.x() << "," << trans.y() << "," << trans.z(); }
  std::printf("%s\n", os.str().c_str());
  KeyFrameGraph g;
  g.keyframesAllMutex.lock_shared();
  g.keyframesAllMutex.unlock_shared();
  return 0;
}
''')
    exe = tmp_path / "consumer"
    r = subprocess.run(["g++", "-std=c++14", "-Wall", str(src), "-o", str(exe), "-I", os.path.join(ROOT, "tests", "consumer_shim"),
                        "-I", os.path.join(PKG, "host", "compat"), "-L", PKG, "-llsd_b200", f"-Wl,-rpath,{PKG}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert out.stdout.strip() == "7,1.2,1.6,3.6,0.1,-0.2,0.3"
