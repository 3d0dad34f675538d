"""CPU checks of the drop-in boundary: the library loads and exports every symbol include/lsd_b200.h
declares (no compute calls: there is no GPU here)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "lsd_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lsd_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported_and_bound(lsd):
    L = lsd.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/lsd_b200.h but not exported"
        assert n in lsd.SYMBOLS, f"{n} not bound in lsd_b200/binding.py"
    for n in lsd.SYMBOLS:
        assert n in names, f"{n} bound but not declared"


def test_no_cpu_fallback_without_device(lsd):
    """On a box without a GPU every compute entry point must fail loudly."""
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    with pytest.raises(lsd.LsdError):
        lsd.Context(640, 480, (525, 525, 319.5, 239.5))


def test_product_never_touches_oracle():
    pkg = os.path.join(ROOT, "lsd-slam-pangolin-gui_b200")
    for d, _, files in os.walk(pkg):
        if "build" in d.split(os.sep):
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp", ".py")) or f == "Makefile":
                src = open(os.path.join(d, f), errors="ignore").read()
                assert "pyoracle" not in src and "lsd_oracle" not in src and "oracle/" not in src.replace("the oracle", ""), f


def test_cpp_adapter_compiles_and_links():
    """host/lsd_b200.hpp (upstream class names over the C ABI) builds with plain g++ -std=c++11 and links the library."""
    import subprocess
    pkg = os.path.join(ROOT, "lsd-slam-pangolin-gui_b200")
    subprocess.check_call(["make", "-C", pkg, "-s", "host-check"])
    out = subprocess.run([os.path.join(pkg, "build", "adapter_check")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "no device" in out.stdout or "tracked:" in out.stdout


def test_library_holds_sm_100a_code_for_every_kernel():
    """The shipped library carries sm_100a SASS only (no other arch, no PTX-only fallback) and one entry per kernel named in
    DESIGN.md section 4 -- a kernel silently dropped from the build would otherwise only show up on the GPU box."""
    import shutil
    import subprocess
    import pytest
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    lib = os.path.join(ROOT, "lsd-slam-pangolin-gui_b200", "liblsd_b200.so")
    elfs = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout.split()
    archs = {m for e in elfs for m in re.findall(r"sm_\d+a?", e)}
    assert archs == {"sm_100a"}, archs
    syms = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    for k in ("k_ingest", "k_gradients", "k_maxgrad0", "k_idepth_pyramid", "k_make_pointcloud", "k_se3_track", "k_sim3_track",
              "k_depth_observe", "k_depth_fill_holes", "k_depth_regularize", "k_prop_scatter", "k_prop_replay", "k_depth_sums", "k_depth_set_depth", "k_vbo_extract", "k_publish_pack", "k_remap_u8",
              "k_permaref_overlap", "k_idepth_stats"):
        assert k in syms, f"kernel {k} missing from liblsd_b200.so"


def test_pose_line_format_equals_ostream_default(lsd, tmp_path):
    """TextOutputIOWrapper::publishTrackedFrame streams `id << "," << trans.x() << ...` with ostream's default formatting
    (/root/reference/lib/Pangolin_IOWrapper/TextOutputIOWrapper.cpp:100-120).  lsd_slam_pose_line must produce the same
    characters: checked against a real std::ostream compiled here, on values that exercise %g's corner cases."""
    import ctypes as C
    import subprocess
    vals = [0.0, -0.0, 1.0, -1.5, 0.1, 1e-5, 1.23456789e-5, 123456.7, 1234567.8, 0.000123456789, 3.14159265358979, -2.5e10, 1e100, 5e-324]
    src = tmp_path / "fmt.cpp"
    src.write_text('#include <iostream>\n#include <cstdlib>\nint main(int c, char** v){ std::cout << atoi(v[1]);'
                   ' for (int i = 2; i < c; i++) std::cout << "," << strtod(v[i], nullptr); std::cout << std::endl; }\n')
    exe = tmp_path / "fmt"
    subprocess.check_call(["g++", "-O1", str(src), "-o", str(exe)])
    L = lsd.load()
    for k in range(0, len(vals) - 5):
        six = vals[k:k + 6]
        st = lsd.SlamStatus()
        st.frameId = 40 + k
        for i in range(3):
            st.camToWorld[4 + i] = six[i]
            st.thisToParent_raw[4 + i] = six[3 + i]
        buf = C.create_string_buffer(256)
        assert L.lsd_slam_pose_line(C.byref(st), buf, 256) == 0
        want = subprocess.run([str(exe), str(40 + k)] + [repr(v) for v in six], capture_output=True, text=True).stdout
        assert buf.value.decode() == want, (buf.value, want)


def test_ref_frame_score_is_the_16_9_formula(lsd):
    """[UP] TrackableKeyFrameSearch::getRefFrameScore = distSq * KFDistWeight^2 + (1 - usage)^2 * KFUsageWeight^2 with
    KFDistWeight 4 / KFUsageWeight 3: the C driver and the Python mirror must both use 16 and 9."""
    import numpy as np
    from lsd_b200 import pipeline
    L = lsd.load()
    f = np.float32
    for d2, u in [(0.0, 1.0), (0.01, 0.9), (0.0234, 0.61), (0.5, 0.3), (1e-4, 0.999)]:
        want = float(f(d2) * f(4) * f(4) + (f(1) - f(u)) * (f(1) - f(u)) * f(3) * f(3))
        assert L.lsd_slam_ref_frame_score(d2, u) == want
        assert pipeline.ref_frame_score(d2, u) == want
        assert abs(want - (16.0 * d2 + 9.0 * (1.0 - u) ** 2)) < 1e-5
    # usage 0.6 with no motion must already trigger a keyframe after the initialisation phase (9 * 0.16 = 1.44 > 1)
    assert L.lsd_slam_ref_frame_score(0.0, 0.6) > 1.0
