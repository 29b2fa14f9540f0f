// Stand-in for the reference's Vulkan-backed Image (Image.h / Image.cpp need Vulkan, imgui and stb): Renderer only creates it,
// resizes it and hands it the finished framebuffer (Renderer.hpp:56-57,477). TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstdint>
enum class ImageFormat { None = 0, RGBA, RGBA32F };
class Image {
public:
	Image(uint32_t w, uint32_t h, ImageFormat, const void* = nullptr) : w_(w), h_(h) {}
	void Resize(uint32_t w, uint32_t h) { w_ = w; h_ = h; }
	void SetData(const void* p) { last_ = p; }
	uint32_t GetWidth() const { return w_; }
	uint32_t GetHeight() const { return h_; }
private:
	uint32_t w_, h_; const void* last_ = nullptr;
};
