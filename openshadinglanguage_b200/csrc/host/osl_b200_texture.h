// osl_b200_texture.h — host side of texture(): the process-wide image registry and the
// Radiance .hdr reader (product code, internal to libosl_b200.so).
//
// The reference resolves texture file names through OIIO's TextureSystem, which the
// renderer owns and hands to the ShadingSystem (include/OSL/oslexec.h:172); here the
// renderer either registers decoded images by name (b200_texture_add) or lets the
// library read Radiance RGBE files found on the group's `texturepath` option.
#pragma once
#include <string>
#include <vector>

namespace oslb200 {

struct TextureImage {
    int w = 0, h = 0, nch = 0;
    std::vector<float> rgba;  // float4 per texel, top scanline first; alpha 1 when absent
};

// Registers (or replaces) an in-memory image: `pixels` is [h][w][nch] float, nch 1..4.
void texture_add(const std::string& name, int w, int h, int nch, const float* pixels);

// Registered image, else a .hdr file at `name` (absolute, relative to the working
// directory, or under one of the ':'-separated `searchpath` entries) decoded on first
// use.  nullptr + `err` when it cannot be had.
const TextureImage* texture_get(const std::string& name, const std::string& searchpath, std::string& err);

// Device-side descriptor, must match osld::TexDesc (device/osl_b200_texture.cuh).
struct TexDescHost {
    const void* px;
    int w, h, nch, pad_;
};

// Uploads the images named by a loaded module's texture() calls to the CURRENT device and
// fills the module's `osl_tex_` table; `module` is the CUmodule.  Returns "" or the error;
// `allocations` receives the device buffers (the caller frees them with cudaFree).
// Defined in osl_b200_render.cu (next to the driver-API table).
std::string bind_module_textures(void* module, const std::vector<std::string>& names, const std::string& searchpath,
                                 std::vector<void*>& allocations);

}  // namespace oslb200
