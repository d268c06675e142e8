// osl_b200_texture.cpp — image registry + Radiance RGBE (.hdr) reader (product code).
//
// File format: "#?RADIANCE" text header up to an empty line, a resolution line
// "-Y h +X w", then per scanline either flat RGBE quads or the new run-length coding
// (2 2 hi lo, then the four channels separately as runs: count > 128 repeats the next
// byte count-128 times, otherwise `count` literal bytes follow).  Texel = mantissa *
// 2^(e-136), zero for e == 0 — the conversion OIIO's hdr reader applies (no +0.5).
#include "osl_b200_texture.h"

#include <cmath>
#include <cstdio>
#include <map>
#include <memory>
#include <mutex>

namespace oslb200 {
namespace {

std::mutex g_mutex;
std::map<std::string, std::unique_ptr<TextureImage>>&
registry()
{
    static std::map<std::string, std::unique_ptr<TextureImage>> r;
    return r;
}

bool
read_file(const std::string& path, std::vector<unsigned char>& data)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f)
        return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    data.resize(n > 0 ? (size_t)n : 0);
    bool ok = n >= 0 && fread(data.data(), 1, data.size(), f) == data.size();
    fclose(f);
    return ok;
}

bool
decode_hdr(const std::vector<unsigned char>& d, TextureImage& im, std::string& err)
{
    size_t pos = 0;
    auto line  = [&](std::string& s) {
        s.clear();
        while (pos < d.size() && d[pos] != '\n')
            s.push_back((char)d[pos++]);
        if (pos >= d.size())
            return false;
        ++pos;
        return true;
    };
    std::string s;
    if (!line(s) || s.compare(0, 2, "#?") != 0) {
        err = "not a Radiance .hdr file";
        return false;
    }
    while (true) {
        if (!line(s)) {
            err = "truncated .hdr header";
            return false;
        }
        if (s.empty())
            break;
    }
    int h = 0, w = 0;
    if (!line(s) || sscanf(s.c_str(), "-Y %d +X %d", &h, &w) != 2 || h <= 0 || w <= 0) {
        err = "unsupported .hdr orientation '" + s + "' (only -Y h +X w)";
        return false;
    }
    im.w = w, im.h = h, im.nch = 3;
    im.rgba.assign((size_t)w * h * 4, 1.0f);
    std::vector<unsigned char> row((size_t)w * 4);
    for (int y = 0; y < h; ++y) {
        if (pos + 4 <= d.size() && w >= 8 && w < 32768 && d[pos] == 2 && d[pos + 1] == 2
            && ((d[pos + 2] << 8) | d[pos + 3]) == w) {
            pos += 4;
            for (int c = 0; c < 4; ++c) {
                int x = 0;
                while (x < w) {
                    if (pos + 2 > d.size()) {
                        err = "truncated .hdr scanline";
                        return false;
                    }
                    int n = d[pos];
                    if (n > 128) {
                        n -= 128;
                        if (x + n > w) {
                            err = "bad .hdr run";
                            return false;
                        }
                        for (int k = 0; k < n; ++k)
                            row[(size_t)(x + k) * 4 + c] = d[pos + 1];
                        pos += 2;
                    } else {
                        if (n == 0 || x + n > w || pos + 1 + n > d.size()) {
                            err = "bad .hdr run";
                            return false;
                        }
                        for (int k = 0; k < n; ++k)
                            row[(size_t)(x + k) * 4 + c] = d[pos + 1 + k];
                        pos += 1 + n;
                    }
                    x += n;
                }
            }
        } else {
            if (pos + (size_t)w * 4 > d.size()) {
                err = "truncated .hdr pixels";
                return false;
            }
            for (size_t k = 0; k < (size_t)w * 4; ++k)
                row[k] = d[pos + k];
            pos += (size_t)w * 4;
        }
        float* o = &im.rgba[(size_t)y * w * 4];
        for (int x = 0; x < w; ++x) {
            int e   = row[(size_t)x * 4 + 3];
            float f = e ? ldexpf(1.0f, e - 136) : 0.0f;
            o[x * 4 + 0] = row[(size_t)x * 4 + 0] * f;
            o[x * 4 + 1] = row[(size_t)x * 4 + 1] * f;
            o[x * 4 + 2] = row[(size_t)x * 4 + 2] * f;
        }
    }
    return true;
}

}  // namespace

void
texture_add(const std::string& name, int w, int h, int nch, const float* pixels)
{
    std::unique_ptr<TextureImage> im(new TextureImage);
    im->w = w, im->h = h, im->nch = nch;
    im->rgba.assign((size_t)w * h * 4, 0.0f);
    for (size_t t = 0; t < (size_t)w * h; ++t) {
        for (int c = 0; c < 4; ++c)
            im->rgba[t * 4 + c] = c < nch ? pixels[t * nch + c] : (c == 3 ? 1.0f : 0.0f);
    }
    std::lock_guard<std::mutex> lock(g_mutex);
    registry()[name] = std::move(im);
}

const TextureImage*
texture_get(const std::string& name, const std::string& searchpath, std::string& err)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = registry().find(name);
    if (it != registry().end())
        return it->second.get();
    std::vector<std::string> dirs { "" };
    for (size_t b = 0; b <= searchpath.size();) {
        size_t e = searchpath.find(':', b);
        if (e == std::string::npos)
            e = searchpath.size();
        if (e > b)
            dirs.push_back(searchpath.substr(b, e - b) + "/");
        b = e + 1;
    }
    std::vector<unsigned char> data;
    for (const std::string& dir : dirs) {
        if (!dir.empty() && !name.empty() && name[0] == '/')
            continue;
        if (read_file(dir + name, data)) {
            std::unique_ptr<TextureImage> im(new TextureImage);
            if (!decode_hdr(data, *im, err)) {
                err = "texture '" + name + "': " + err;
                return nullptr;
            }
            const TextureImage* p = im.get();
            registry()[name]       = std::move(im);
            return p;
        }
    }
    err = "texture '" + name + "' not found (not registered with b200_texture_add, no such file on texturepath '"
          + searchpath + "')";
    return nullptr;
}

}  // namespace oslb200
