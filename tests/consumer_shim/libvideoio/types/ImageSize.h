// stand-in for libvideoio/types/ImageSize.h
#pragma once
namespace libvideoio {
struct ImageSize {
  int width = 0, height = 0;
  ImageSize() {}
  ImageSize(int w, int h) : width(w), height(h) {}
};
}  // namespace libvideoio
