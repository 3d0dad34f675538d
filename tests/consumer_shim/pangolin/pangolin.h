// stand-in for <pangolin/pangolin.h>: the three types lib/GUI.h declares members of
#pragma once
#include <string>
namespace pangolin {
struct GlTexture {};
struct OpenGlRenderState {};
template <typename T> struct Var {};
}  // namespace pangolin
